// CTA-pair (tcgen05.mma.cta_group::2) forms of the scaled 2 x FP16 GEMMs for the shipped block shape N = K = 256.
// Included by gemm_f16x2.cu inside namespace gasfm (same translation unit: GemmF16Args, the tile constants and the launchers
// live there); see profiles/r02_gemm_pair.md for the measurements that led here.
#pragma once
// ---------------------------------------------------------------------------------------------------------------
// cta_group::2 form of the same product for the shipped block shape N = K = 256 (the three grouped projections of a block,
// lin_r, the single projections): a CTA PAIR works on 256 rows with ONE tcgen05.mma.cta_group::2 stream (M = 256).
//
// Why: ncu on the cta_group::1 kernel above (profiles/r02_gemm_pair.md) shows the SM's shared-memory data pipe saturated --
// LSU wavefronts 68 % + tensor-core operand wavefronts 32 % of the cycles, 14.2 k wavefronts per 128 x 256 output tile against
// 6.1 k cycles of MMA -- not HBM (39 %) and not the tensor pipe (44 %).  Per tile the pipe carried: operand reads of the MMAs
// (12 KB per instruction: 4.6 k), the fp16 hi/lo planes of A written once PER GROUP (1.0 k), the epilogue's staging round trip
// (2.0 k), 16 broadcast loads of column scale / bias per 32-column chunk (1.0 k), global loads / stores (1.4 k), plus the
// cycles lost when those clients collide.  This kernel removes what can be removed:
//   * cta_group::2: each SM feeds its 128 rows of A and HALF of the weights' K block; the pair shares the halves, so an
//     instruction reads 8 KB per SM instead of 12, and the TMA writes half as much B into each SM;
//   * the halved B stages leave room for the whole fp16 A tile (4 K blocks x hi/lo = 128 KB) to stay resident: A is converted
//     ONCE per tile and reused by every group (slot kb is refilled with the next tile as soon as the last group has used it);
//   * column scale and bias are applied after the staging transposition, where a lane owns 8 columns: 4 loads per chunk.
// Roles per CTA (16 warps, setmaxnreg as above): warp 0 TMA of this CTA's B half; warp 1 MMA issue (leader CTA) or relay
// (peer CTA: forwards "my half is in place" to the leader, one remote arrive per K block); warps 4-11 A producers; warps 12-15
// epilogue of this CTA's 128 accumulator rows.
constexpr int kPKB = 4;                                   // K blocks of 64 (K = 256)
constexpr int kPN = 256;
constexpr int kPSlotBytes = 2 * kFATileBytes;             // A slot: hi | lo planes of one K block (32 KB)
constexpr int kPBPlaneBytes = (kPN / 2) * kFBlockK * 2;   // this CTA's 128 weight rows of one K block, one plane (16 KB)
constexpr int kPBStageBytes = 2 * kPBPlaneBytes;          // hi | lo
constexpr int kPBStages = 2;
constexpr size_t kPSmemBytes = (size_t)kPKB * kPSlotBytes + (size_t)kPBStages * kPBStageBytes + 4 * 4096 + 1024;

__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {     // observes arrivals made by the peer CTA
  const uint32_t addr = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!done);
}
// wait with a back-off between polls: every failed try_wait is a shared-memory wavefront, and the roles that wait long
// (A producers two thirds of the time, the TMA thread, the epilogue) polled ~290 M times per launch in the first version of
// this kernel -- as many wavefronts as the whole epilogue read-back
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity, uint32_t ns) {
  const uint32_t addr = smem_u32(bar);
  for (;;) {
    uint32_t done;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) break;
    __nanosleep(ns);
  }
}
// the same without publishing this thread's earlier writes (nothing but the arrival itself is communicated): no memory barrier
__device__ __forceinline__ void mbar_arrive_remote_relaxed(uint64_t* bar, uint32_t rank) {
  uint32_t raddr;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(smem_u32(bar)), "r"(rank));
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}
// arrive on the barrier at this shared-memory offset in CTA ``rank`` of the cluster.  Default semantics (release at CTA
// scope): what is handed over lives in this CTA's shared memory and is already complete when the arrive is issued (it was
// itself observed through an mbarrier).  A cluster-scope release compiles to MEMBAR.ALL.GPU, which waits for the SM's
// outstanding GLOBAL stores -- the epilogue streams them continuously -- and cost ~2,600 cycles per K block on the relay
// (device timeline, profiles/r02_gemm_pair.md).
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t rank) {
  uint32_t raddr;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(smem_u32(bar)), "r"(rank));
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {      // arrives on this barrier in BOTH CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}

// profiling (tools/gemm_pair_trace.py): SM-clock timestamps of cluster 0, [cta 2][role 4][virtual tile 32][slot 16]
constexpr int kPTraceTiles = 32;
#define GASFM_PTRACE(role, vt, slot)                                                                             \
  do {                                                                                                           \
    if (TRACE && p.trace && blockIdx.x < 2 && (vt) < kPTraceTiles)                                               \
      p.trace[(((int64_t)blockIdx.x * 4 + (role)) * kPTraceTiles + (vt)) * 16 + (slot)] = clock64();             \
  } while (0)

// Drain of one 128 x 256 accumulator by one epilogue warp (its 32 rows), shared by the pair kernels.
template <bool TRACE>
__device__ __forceinline__ void pair_epilogue_drain(const GemmF16Args& p, uint32_t taddr0, uint32_t sts_base, uint32_t lds_v, uint32_t lds_w,
                                                    uint32_t csg, uint32_t bsg, const float (&rs)[4], float* const (&rowp)[4],
                                                    int rows_left, int vt, bool tracer) {
    // 32-column chunks: TMEM -> registers (lane = row) -> swizzled staging tile -> registers (lane = 8 columns of 4 rows)
    // -> descale by row and column, bias -> 256-bit stores.  The next chunk's TMEM read is issued as soon as the registers
    // are free, and all read-back loads of a chunk are issued before the first result is used.
    uint32_t r[32];
#define GASFM_TMEM_LD32(ADDR)                                                                                                  \
    asm volatile(                                                                                                          \
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                          \
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];" \
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),      \
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),           \
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),          \
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])                        \
        : "r"(ADDR))
#define GASFM_LDS4(V, ADDR) \
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"((V).x), "=f"((V).y), "=f"((V).z), "=f"((V).w) : "r"(ADDR) : "memory")
    GASFM_TMEM_LD32(taddr0);
#pragma unroll
    for (int c0 = 0; c0 < kPN; c0 += 32) {
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      __syncwarp();                                  // the previous chunk's read-back is complete
      if (tracer) GASFM_PTRACE(2, vt, 2 + (c0 >> 5));
#pragma unroll
      for (int k = 0; k < 8; ++k)                    // lane = row: raw accumulators into the swizzled staging tile
        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(sts_base ^ (uint32_t)(k << 4)), "r"(r[4 * k]), "r"(r[4 * k + 1]),
                     "r"(r[4 * k + 2]), "r"(r[4 * k + 3]) : "memory");
      __syncwarp();
      if (c0 + 32 < kPN) GASFM_TMEM_LD32(taddr0 + (uint32_t)(c0 + 32));      // overlaps the read-back below
      float4 v[4], w[4], cs0, cs1, bs0, bs1;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        GASFM_LDS4(v[i], lds_v + (uint32_t)(i * 1024));
        GASFM_LDS4(w[i], lds_w + (uint32_t)(i * 1024));
      }
      GASFM_LDS4(cs0, csg + (uint32_t)(c0 * 4)); GASFM_LDS4(cs1, csg + (uint32_t)(c0 * 4 + 16));
      GASFM_LDS4(bs0, bsg + (uint32_t)(c0 * 4)); GASFM_LDS4(bs1, bsg + (uint32_t)(c0 * 4 + 16));
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (8 * i < rows_left) {
          const float f = rs[i];
          asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(rowp[i] + c0), "f"(fmaf(v[i].x, f * cs0.x, bs0.x)),
                       "f"(fmaf(v[i].y, f * cs0.y, bs0.y)), "f"(fmaf(v[i].z, f * cs0.z, bs0.z)), "f"(fmaf(v[i].w, f * cs0.w, bs0.w)),
                       "f"(fmaf(w[i].x, f * cs1.x, bs1.x)), "f"(fmaf(w[i].y, f * cs1.y, bs1.y)), "f"(fmaf(w[i].z, f * cs1.z, bs1.z)),
                       "f"(fmaf(w[i].w, f * cs1.w, bs1.w))
                       : "memory");
        }
      }
    }
#undef GASFM_TMEM_LD32
#undef GASFM_LDS4
}

// DIRECT: the epilogue stores straight from the 16x256b TMEM fragment (a quad of lanes owns one 32-byte sector of a row) instead
// of transposing through shared memory
// LN: the operand is relu(layer_norm(A) * gamma + beta), evaluated once per tile on the register-resident rows (the producers of
// this kernel are idle two thirds of the time: the normalisation is free, and the separate LayerNorm pass over [E, d] goes away)
template <bool TRACE, bool DIRECT, bool LN = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kFThreads, 1)
gemm_f16x2_pair_kernel(const __grid_constant__ CUtensorMap map_bhi, const __grid_constant__ CUtensorMap map_blo, GemmF16Args p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* a_ring = smem;                                           // [kb][hi 16 KB | lo 16 KB]
  uint8_t* b_ring = smem + (size_t)kPKB * kPSlotBytes;              // [stage][hi 16 KB | lo 16 KB], 128 weight rows each
  uint8_t* c_stage = b_ring + (size_t)kPBStages * kPBStageBytes;    // 4 epilogue warps x [32 rows x 128 B]
  __shared__ uint64_t a_full[kPKB], a_empty[kPKB], b_full[kPBStages], b_empty[kPBStages], peer_bar[kPBStages];
  __shared__ uint64_t tmem_full_bar[2], tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_slot;
  __shared__ __align__(16) float bias_s[kFMaxGroups * kPN], bscale_s[kFMaxGroups * kPN];
  __shared__ float row_descale[kFScaleSlots][kFBlockM];
  __shared__ __align__(16) float ln_gamma_s[LN ? kPN : 4], ln_beta_s[LN ? kPN : 4];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta_rank = cluster_ctarank();
  const int64_t num_pair_tiles = (p.M + 2 * kFBlockM - 1) / (2 * kFBlockM);
  const int64_t num_clusters = gridDim.x / 2, cluster_id = blockIdx.x / 2;
  const int64_t my_steps = cluster_id < num_pair_tiles ? (num_pair_tiles - cluster_id + num_clusters - 1) / num_clusters : 0;
  const int groups = p.groups;
  if constexpr (LN) {
    for (int j = threadIdx.x; j < kPN; j += kFThreads) { ln_gamma_s[j] = p.ln_gamma[j]; ln_beta_s[j] = p.ln_beta[j]; }
  }

  if (threadIdx.x == 0) {
    for (int k = 0; k < kPKB; ++k) { mbar_init(&a_full[k], 256); mbar_init(&a_empty[k], 1); }
    for (int s = 0; s < kPBStages; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); mbar_init(&peer_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full_bar[a], 1); mbar_init(&tmem_empty_bar[a], 8); }   // 4 epilogue warps x 2 CTAs
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int j = threadIdx.x; j < groups * kPN; j += kFThreads) {
    bias_s[j] = p.bias ? p.bias[j] : 0.f;
    bscale_s[j] = p.b_scale[j];
  }
  if (warp == 1) {        // one warp of EACH CTA of the pair
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_slot;

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;" ::: "memory");
    if (warp == 0 && lane == 0) {
      // ===================== TMA: this CTA's 128 rows of the group's weights, K block by K block =====================
      int stage = 0; uint32_t phase = 0;
      for (int64_t vt = 0; vt < my_steps * groups; ++vt) {
        const int b_row0 = (int)(vt % groups) * kPN + (int)cta_rank * (kPN / 2);
        for (int kb = 0; kb < kPKB; ++kb) {
          mbar_wait_backoff(&b_empty[stage], phase ^ 1, 32);
          GASFM_PTRACE(3, vt, kb);
          uint8_t* st = b_ring + (size_t)stage * kPBStageBytes;
          mbar_expect_tx(&b_full[stage], kPBStageBytes);
          tma_load_2d(st, &map_bhi, &b_full[stage], kb * kFBlockK, b_row0);
          tma_load_2d(st + kPBPlaneBytes, &map_blo, &b_full[stage], kb * kFBlockK, b_row0);
          if (++stage == kPBStages) { stage = 0; phase ^= 1; }
        }
      }
    } else if (warp == 1 && lane == 0) {
      int stage = 0; uint32_t phase = 0;
      if (cta_rank != 0) {
        // ===================== relay (peer CTA): "my B half and my A block are in place" -> the leader =====================
        for (int64_t it = 0; it < my_steps; ++it) {
          for (int g = 0; g < groups; ++g) {
            for (int kb = 0; kb < kPKB; ++kb) {
              mbar_wait_backoff(&b_full[stage], phase, 20);
              if (g == 0) mbar_wait_backoff(&a_full[kb], (uint32_t)(it & 1), 20);
              mbar_arrive_remote(&peer_bar[stage], 0);
              if (++stage == kPBStages) { stage = 0; phase ^= 1; }
            }
          }
        }
      } else {
        // ===================== MMA issuer (leader CTA), M = 256 over the pair =====================
        const uint32_t idesc = (1u << 4) | ((uint32_t)(kPN >> 3) << 17) | ((uint32_t)((2 * kFBlockM) >> 4) << 24);
        int acc = 0; uint32_t acc_phase = 0;
        const uint32_t a_base = smem_u32(a_ring), b_base = smem_u32(b_ring);
        for (int64_t it = 0; it < my_steps; ++it) {
          for (int g = 0; g < groups; ++g) {
            mbar_wait_cluster(&tmem_empty_bar[acc], acc_phase ^ 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            GASFM_PTRACE(1, it * groups + g, 0);
            const uint32_t d_tmem = tmem_base + (uint32_t)(acc * kPN);
            for (int kb = 0; kb < kPKB; ++kb) {
              mbar_wait(&b_full[stage], phase);
              GASFM_PTRACE(1, it * groups + g, 1 + kb);
              if (g == 0) mbar_wait(&a_full[kb], (uint32_t)(it & 1));
              GASFM_PTRACE(1, it * groups + g, 5 + kb);
              mbar_wait_cluster(&peer_bar[stage], phase);
              GASFM_PTRACE(1, it * groups + g, 9 + kb);
              asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
              const uint32_t a_hi = a_base + (uint32_t)kb * kPSlotBytes, a_lo = a_hi + kFATileBytes;
              const uint32_t b_hi = b_base + (uint32_t)stage * kPBStageBytes, b_lo = b_hi + kPBPlaneBytes;
#pragma unroll
              for (int k = 0; k < kFBlockK / kFUmmaK; ++k) {
                const uint32_t koff = k * kFUmmaK * 2;
                umma_f16_pair(d_tmem, make_desc(a_hi + koff), make_desc(b_hi + koff), idesc, (kb == 0 && k == 0) ? 0u : 1u);
                umma_f16_pair(d_tmem, make_desc(a_lo + koff), make_desc(b_hi + koff), idesc, 1u);
                umma_f16_pair(d_tmem, make_desc(a_hi + koff), make_desc(b_lo + koff), idesc, 1u);
              }
              umma_commit_pair(&b_empty[stage]);
              if (g == groups - 1) umma_commit_pair(&a_empty[kb]);      // the last group is done with this K block of A
              if (kb == kPKB - 1) umma_commit_pair(&tmem_full_bar[acc]);
              if (++stage == kPBStages) { stage = 0; phase ^= 1; }
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
          }
        }
      }
    }
  } else if (warp < 12) {
    // ===================== A producers: this CTA's 128 rows, converted once per tile =====================
    if constexpr (DIRECT) asm volatile("setmaxnreg.inc.sync.aligned.u32 176;" ::: "memory");
    else asm volatile("setmaxnreg.inc.sync.aligned.u32 184;" ::: "memory");
    const int t = threadIdx.x - 128;
    const int q = t & 15, rg = t >> 4;
    uint32_t soff[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int row = rg + 16 * i;
      soff[i] = (uint32_t)(row * 128 + ((((q >> 1) ^ (row & 7))) << 4) + ((q & 1) << 3));
    }
    const uint32_t ring_base = smem_u32(a_ring);
    const uint32_t toff = (uint32_t)(rg * p.lda + q * 4), tstep = (uint32_t)(16 * p.lda);
    float4 buf[kPKB][8];
    auto load_block = [&](int64_t it, int kb, float4 (&v)[8]) {
      const int64_t row0 = ((cluster_id + it * num_clusters) * 2 + cta_rank) * kFBlockM;
      if (it < my_steps && row0 + kFBlockM <= p.M) {
        const float* base = p.A + row0 * p.lda + kb * kFBlockK;
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = ld_stream4(base + (toff + (uint32_t)i * tstep));
        return;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int64_t row = row0 + rg + 16 * i;
        v[i] = (it < my_steps && row < p.M) ? ld_stream4(p.A + row * p.lda + kb * kFBlockK + q * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
#pragma unroll
    for (int kb = 0; kb < kPKB; ++kb) load_block(0, kb, buf[kb]);
    float seen_max = 0.f;
    for (int64_t it = 0; it < my_steps; ++it) {
      if constexpr (LN) {
        // LayerNorm + ReLU in place on the register tile: two-pass mean / variance over the 256 columns of each row (16 lanes
        // share a row), then y = max(0, (x - mean) rstd gamma + beta) -- the arithmetic of ln_relu_fwd_kernel
        const int64_t row0 = ((cluster_id + it * num_clusters) * 2 + cta_rank) * kFBlockM;
        const bool full = row0 + kFBlockM <= p.M;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int64_t row = row0 + rg + 16 * i;
          float sum = 0.f;
#pragma unroll
          for (int kb = 0; kb < kPKB; ++kb) sum += (buf[kb][i].x + buf[kb][i].y) + (buf[kb][i].z + buf[kb][i].w);
#pragma unroll
          for (int off = 1; off < 16; off <<= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
          const float mean = sum * (1.f / kPN);
          float sq = 0.f;
#pragma unroll
          for (int kb = 0; kb < kPKB; ++kb) {
            const float dx = buf[kb][i].x - mean, dy = buf[kb][i].y - mean, dz = buf[kb][i].z - mean, dw = buf[kb][i].w - mean;
            sq = fmaf(dx, dx, sq); sq = fmaf(dy, dy, sq); sq = fmaf(dz, dz, sq); sq = fmaf(dw, dw, sq);
          }
#pragma unroll
          for (int off = 1; off < 16; off <<= 1) sq += __shfl_xor_sync(0xffffffffu, sq, off);
          const float rstd = 1.f / sqrtf(sq * (1.f / kPN) + p.ln_eps);
          const bool live = full || row < p.M;              // rows past the end stay zero (they must not enter a_amax)
#pragma unroll
          for (int kb = 0; kb < kPKB; ++kb) {
            const int kcol = kb * kFBlockK + q * 4;
            const float4 g = *reinterpret_cast<const float4*>(&ln_gamma_s[kcol]);
            const float4 b = *reinterpret_cast<const float4*>(&ln_beta_s[kcol]);
            float4 y;
            y.x = live ? fmaxf(fmaf((buf[kb][i].x - mean) * rstd, g.x, b.x), 0.f) : 0.f;
            y.y = live ? fmaxf(fmaf((buf[kb][i].y - mean) * rstd, g.y, b.y), 0.f) : 0.f;
            y.z = live ? fmaxf(fmaf((buf[kb][i].z - mean) * rstd, g.z, b.z), 0.f) : 0.f;
            y.w = live ? fmaxf(fmaf((buf[kb][i].w - mean) * rstd, g.w, b.w), 0.f) : 0.f;
            buf[kb][i] = y;
            if (p.ln_y != nullptr && live) st_stream4(p.ln_y + row * p.ldy + kcol, y);
          }
          if (q == 0 && live) { p.ln_mean[row] = mean; p.ln_rstd[row] = rstd; }
        }
      }
      float scale[8];
      float* descale_slot = row_descale[it & (kFScaleSlots - 1)];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float m = 0.f;
#pragma unroll
        for (int kb = 0; kb < kPKB; ++kb)
          m = fmaxf(m, fmaxf(fmaxf(fabsf(buf[kb][i].x), fabsf(buf[kb][i].y)), fmaxf(fabsf(buf[kb][i].z), fabsf(buf[kb][i].w))));
#pragma unroll
        for (int off = 1; off < 16; off <<= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
        float descale;
        seen_max = fmaxf(seen_max, m);
        row_scale_from_amax(m, scale[i], descale);
        if (q == 0) descale_slot[rg + 16 * i] = descale;
      }
#pragma unroll
      for (int kb = 0; kb < kPKB; ++kb) {
        // every group of the previous tile has read this slot.  ONE lane polls the mbarrier, the other producer warps block on
        // a named barrier (256 polling threads made 212 M shared-memory wavefronts per launch)
        if (t == 0) mbar_wait_backoff(&a_empty[kb], (uint32_t)((it & 1) ^ 1), 64);
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (t == 0) GASFM_PTRACE(0, it * groups, kb);
        const uint32_t sb = ring_base + (uint32_t)kb * kPSlotBytes;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float s = scale[i];
          const float x0 = buf[kb][i].x * s, x1 = buf[kb][i].y * s, x2 = buf[kb][i].z * s, x3 = buf[kb][i].w * s;
          const __half2 h01 = __floats2half2_rn(x0, x1), h23 = __floats2half2_rn(x2, x3);
          const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
          const __half2 l01 = __floats2half2_rn(x0 - f01.x, x1 - f01.y), l23 = __floats2half2_rn(x2 - f23.x, x3 - f23.y);
          const uint32_t addr = sb + soff[i];
          asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(*reinterpret_cast<const uint32_t*>(&h01)),
                       "r"(*reinterpret_cast<const uint32_t*>(&h23)) : "memory");
          asm volatile("st.shared.v2.b32 [%0+%3], {%1, %2};" ::"r"(addr), "r"(*reinterpret_cast<const uint32_t*>(&l01)),
                       "r"(*reinterpret_cast<const uint32_t*>(&l23)), "n"(kFATileBytes) : "memory");
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive(&a_full[kb]);
        if (t == 0) GASFM_PTRACE(0, it * groups, 4 + kb);
        load_block(it + 1, kb, buf[kb]);                         // the freed registers take the next tile's K block
      }
    }
    if (p.a_amax != nullptr) warp_amax_to_global(seen_max, p.a_amax);
  } else {
    // ===================== epilogue: this CTA's 128 accumulator rows (warp -> TMEM lane quarter) =====================
    const int quarter = warp & 3;
    if constexpr (DIRECT) {
      asm volatile("setmaxnreg.dec.sync.aligned.u32 112;" ::: "memory");
      // 16x256b fragments: lane (rq = lane / 4, c = lane % 4) holds columns 8 n + 2 c, + 1 of rows rq and rq + 8 of a 16-lane
      // half; two loads (lanes 0-15, 16-31 of the quarter) give rows rq + 8 h, h < 4.  A store instruction then writes one full
      // 32-byte sector in each of 8 rows: no staging tile, no shared-memory round trip in the warp's dependency chain.
      const int rq = lane >> 2, c = lane & 3;
      const uint32_t cs_l = smem_u32(bscale_s) + (uint32_t)(c * 8), bs_l = smem_u32(bias_s) + (uint32_t)(c * 8);
      int acc = 0; uint32_t acc_phase = 0;
      for (int64_t it = 0; it < my_steps; ++it) {
        const int64_t row0 = ((cluster_id + it * num_clusters) * 2 + cta_rank) * kFBlockM + quarter * 32;
        const int rows_left = (int)((p.M - row0) < 32 ? (p.M - row0) : 32) - rq;        // row rq + 8 h exists iff 8 h < rows_left
        float rs[4];
        float* rowp[4];
        for (int g = 0; g < groups; ++g) {
          mbar_wait_backoff(&tmem_full_bar[acc], acc_phase, 32);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          if (warp == 12 && lane == 0) GASFM_PTRACE(2, it * groups + g, 0);
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            rs[h] = row_descale[it & (kFScaleSlots - 1)][quarter * 32 + rq + 8 * h];
            rowp[h] = p.C + (row0 + rq + 8 * h) * p.ldc + g * kPN + 2 * c;
          }
          const uint32_t csg = cs_l + (uint32_t)g * (kPN * 4), bsg = bs_l + (uint32_t)g * (kPN * 4);
          const uint32_t ta = tmem_base + (uint32_t)(acc * kPN) + ((uint32_t)(quarter * 32) << 16), tb = ta + (16u << 16);
          uint32_t f0[32], f1[32];                      // [0,16): lanes 0-15 of the quarter, [16,32): lanes 16-31; two sets in flight
#define GASFM_TMEM_LD16(R, O, ADDR)                                                                                            \
          asm volatile("tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"    \
                       : "=r"(R[O + 0]), "=r"(R[O + 1]), "=r"(R[O + 2]), "=r"(R[O + 3]), "=r"(R[O + 4]), "=r"(R[O + 5]),           \
                         "=r"(R[O + 6]), "=r"(R[O + 7]), "=r"(R[O + 8]), "=r"(R[O + 9]), "=r"(R[O + 10]), "=r"(R[O + 11]),         \
                         "=r"(R[O + 12]), "=r"(R[O + 13]), "=r"(R[O + 14]), "=r"(R[O + 15])                                        \
                       : "r"(ADDR))
#define GASFM_EPI_BLOCK(R, CB)                                                                                                 \
          _Pragma("unroll") for (int n = 0; n < 4; ++n) {                                                                      \
            float2 cs, bs;                                                                                                     \
            asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(cs.x), "=f"(cs.y) : "r"(csg + (uint32_t)(((CB) + 8 * n) * 4)) : "memory"); \
            asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(bs.x), "=f"(bs.y) : "r"(bsg + (uint32_t)(((CB) + 8 * n) * 4)) : "memory"); \
            _Pragma("unroll") for (int h = 0; h < 4; ++h) {                                                                    \
              if (8 * h < rows_left) {                                                                                         \
                const float a0 = __uint_as_float(R[16 * (h >> 1) + 4 * n + 2 * (h & 1)]);                                      \
                const float a1 = __uint_as_float(R[16 * (h >> 1) + 4 * n + 2 * (h & 1) + 1]);                                  \
                asm volatile("st.global.v2.f32 [%0], {%1,%2};" ::"l"(rowp[h] + (CB) + 8 * n), "f"(fmaf(a0, rs[h] * cs.x, bs.x)),  \
                             "f"(fmaf(a1, rs[h] * cs.y, bs.y)) : "memory");                                                  \
              }                                                                                                                \
            }                                                                                                                  \
          }
          GASFM_TMEM_LD16(f0, 0, ta); GASFM_TMEM_LD16(f0, 16, tb);
#pragma unroll
          for (int cb = 0; cb < kPN; cb += 64) {
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (warp == 12 && lane == 0) GASFM_PTRACE(2, it * groups + g, 2 + (cb >> 5));
            GASFM_TMEM_LD16(f1, 0, ta + (uint32_t)(cb + 32)); GASFM_TMEM_LD16(f1, 16, tb + (uint32_t)(cb + 32));
            GASFM_EPI_BLOCK(f0, cb)
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (warp == 12 && lane == 0) GASFM_PTRACE(2, it * groups + g, 3 + (cb >> 5));
            if (cb + 64 < kPN) { GASFM_TMEM_LD16(f0, 0, ta + (uint32_t)(cb + 64)); GASFM_TMEM_LD16(f0, 16, tb + (uint32_t)(cb + 64)); }
            GASFM_EPI_BLOCK(f1, cb + 32)
          }
#undef GASFM_TMEM_LD16
#undef GASFM_EPI_BLOCK
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive_remote_relaxed(&tmem_empty_bar[acc], 0);
          if (warp == 12 && lane == 0) GASFM_PTRACE(2, it * groups + g, 1);
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
      }
    } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 96;" ::: "memory");
    const uint32_t stg = smem_u32(c_stage + (warp - 12) * 4096);   // [32 rows x 128 B], 16-byte chunk ^= row % 8
    const int pc = lane & 3, rsub = lane >> 2;
    // everything that does not change from chunk to chunk is computed once (the first version spent ~270 instructions per
    // 32-column chunk, two thirds of them on addresses and predicates, and the epilogue warps' issue latency paced the kernel):
    // write side, lane = row: 16-byte piece k of the row goes to chunk k ^ (row % 8) -> one XOR per store
    const uint32_t sts_base = stg + (uint32_t)(lane * 128 + ((lane & 7) << 4));
    // read side, lane = (rows rsub + 8 i, 32-byte piece pc): row % 8 == rsub for all four rows -> two bases + immediates
    const uint32_t lds_v = stg + (uint32_t)(rsub * 128 + (((2 * pc) ^ rsub) << 4)), lds_w = lds_v ^ 16u;
    const uint32_t cs_base = smem_u32(bscale_s) + (uint32_t)(pc * 32), bs_base = smem_u32(bias_s) + (uint32_t)(pc * 32);
    int acc = 0; uint32_t acc_phase = 0;
    for (int64_t it = 0; it < my_steps; ++it) {
      const int64_t row0 = ((cluster_id + it * num_clusters) * 2 + cta_rank) * kFBlockM + quarter * 32;
      float rs[4];                                               // row descale of the 4 rows this lane stores
      float* rowp[4];
      const int rows_left = (int)((p.M - row0) < 32 ? (p.M - row0) : 32) - rsub;      // row 8 i + rsub exists iff 8 i < rows_left
      for (int g = 0; g < groups; ++g) {
        mbar_wait_backoff(&tmem_full_bar[acc], acc_phase, 32);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (warp == 12 && lane == 0) GASFM_PTRACE(2, it * groups + g, 0);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int64_t grow = row0 + 8 * i + rsub;
          rs[i] = row_descale[it & (kFScaleSlots - 1)][quarter * 32 + 8 * i + rsub];
          rowp[i] = p.C + grow * p.ldc + g * kPN + 8 * pc;
        }
        const uint32_t csg = cs_base + (uint32_t)g * (kPN * 4), bsg = bs_base + (uint32_t)g * (kPN * 4);
        const uint32_t taddr0 = tmem_base + (uint32_t)(acc * kPN) + ((uint32_t)(quarter * 32) << 16);
        pair_epilogue_drain<TRACE>(p, taddr0, sts_base, lds_v, lds_w, csg, bsg, rs, rowp, rows_left, (int)(it * groups + g),
                                   warp == 12 && lane == 0);
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive_remote_relaxed(&tmem_empty_bar[acc], 0);   // the leader's MMA thread owns the accumulator hand-back
        if (warp == 12 && lane == 0) GASFM_PTRACE(2, it * groups + g, 1);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
    }
  }
  __syncthreads();
  cluster_sync();                            // neither CTA leaves (or frees TMEM) while the pair's MMAs / remote arrives are in flight
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------------------
// cta_group::2 form of the concatenated input gradient dX[M, 256] = [dY_0 | dY_1 | ..] Wcat^T with row maxima from upstream
// (n_seg segments of 256 columns, K = 256 n_seg).  Same pair structure as gemm_f16x2_pair_kernel; the operand streams instead
// of staying resident: a stage holds one K block of this CTA's 128 rows (fp16 hi | lo, written by the producers) and this CTA's
// half of the weights' K block (TMA) -- 64 KB, so THREE stages fit where the cta_group::1 kernel has two of 96 KB.
constexpr int kCStageBytes = kPSlotBytes + kPBStageBytes;       // A hi | A lo | B hi half | B lo half
constexpr int kCStages = 3;
constexpr size_t kCSmemBytes = (size_t)kCStages * kCStageBytes + 4 * 4096 + 1024;

// PWG: producer warpgroups (2, default: 8 rows per thread as in the other GEMMs; 4: 16 producer warps, 4 rows per thread,
// GASFM_GEMM_CAT_PRODUCERS=16).  The timeline of the 8-warp form shows the operand producers setting the pace, but doubling them
// does not help (6.42 vs 6.25 ms at cfg3): what they wait for is the operand loads and the shared-memory data pipe, not issue slots.
template <bool TRACE, int PWG>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128 * (PWG + 2), 1)
gemm_f16x2_cat_pair_kernel(const __grid_constant__ CUtensorMap map_bhi, const __grid_constant__ CUtensorMap map_blo, GemmF16Args p) {
  constexpr int kThreads = 128 * (PWG + 2), kProducers = 128 * PWG, kRows = 16 / PWG, kRowStep = 8 * PWG;   // rows per thread / stride
  constexpr int kEpiWarp0 = 4 * (PWG + 1);                     // first epilogue warp (a multiple of 4: warp % 4 = TMEM lane quarter)
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* c_stage = smem + (size_t)kCStages * kCStageBytes;
  __shared__ uint64_t full_bar[kCStages], split_bar[kCStages], empty_bar[kCStages], peer_bar[kCStages];
  __shared__ uint64_t tmem_full_bar[2], tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_slot;
  __shared__ __align__(16) float bias_s[kPN], bscale_s[kPN];
  __shared__ float row_descale[kFScaleSlots][kFBlockM];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta_rank = cluster_ctarank();
  const int64_t num_pair_tiles = (p.M + 2 * kFBlockM - 1) / (2 * kFBlockM);
  const int64_t num_clusters = gridDim.x / 2, cluster_id = blockIdx.x / 2;
  const int64_t my_steps = cluster_id < num_pair_tiles ? (num_pair_tiles - cluster_id + num_clusters - 1) / num_clusters : 0;
  const int nkb = kPKB * p.n_seg;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kCStages; ++s) {
      mbar_init(&full_bar[s], 1); mbar_init(&split_bar[s], kProducers); mbar_init(&empty_bar[s], 1); mbar_init(&peer_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full_bar[a], 1); mbar_init(&tmem_empty_bar[a], 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int j = threadIdx.x; j < kPN; j += kThreads) {
    bias_s[j] = p.bias ? p.bias[j] : 0.f;
    bscale_s[j] = p.b_scale[j];
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_slot;

  if (warp < 4) {
    // PWG == 4: launched with 80 registers per thread (768 threads); the redistribution must stay within that allocation:
    // 128 x 24 + 512 x 88 + 128 x 96 = 60,416 <= 61,440 (asking for more makes the last setmaxnreg.inc wait forever)
    if constexpr (PWG == 4) asm volatile("setmaxnreg.dec.sync.aligned.u32 24;" ::: "memory");
    else asm volatile("setmaxnreg.dec.sync.aligned.u32 40;" ::: "memory");
    if (warp == 0 && lane == 0) {
      // ===================== TMA: this CTA's 128 weight rows, K block by K block =====================
      int stage = 0; uint32_t phase = 0;
      const int b_row0 = (int)cta_rank * (kPN / 2);
      for (int64_t it = 0; it < my_steps; ++it) {
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait_backoff(&empty_bar[stage], phase ^ 1, 32);
          if (kb < 16) GASFM_PTRACE(3, it, kb);
          uint8_t* st = smem + (size_t)stage * kCStageBytes + kPSlotBytes;
          mbar_expect_tx(&full_bar[stage], kPBStageBytes);
          tma_load_2d(st, &map_bhi, &full_bar[stage], kb * kFBlockK, b_row0);
          tma_load_2d(st + kPBPlaneBytes, &map_blo, &full_bar[stage], kb * kFBlockK, b_row0);
          if (++stage == kCStages) { stage = 0; phase ^= 1; }
        }
      }
    } else if (warp == 1 && lane == 0) {
      int stage = 0; uint32_t phase = 0;
      if (cta_rank != 0) {
        // ===================== relay (peer CTA) =====================
        for (int64_t it = 0; it < my_steps; ++it) {
          for (int kb = 0; kb < nkb; ++kb) {
            mbar_wait_backoff(&full_bar[stage], phase, 20);
            mbar_wait_backoff(&split_bar[stage], phase, 20);
            mbar_arrive_remote(&peer_bar[stage], 0);
            if (++stage == kCStages) { stage = 0; phase ^= 1; }
          }
        }
      } else {
        // ===================== MMA issuer (leader CTA), M = 256 over the pair =====================
        const uint32_t idesc = (1u << 4) | ((uint32_t)(kPN >> 3) << 17) | ((uint32_t)((2 * kFBlockM) >> 4) << 24);
        int acc = 0; uint32_t acc_phase = 0;
        const uint32_t s_base = smem_u32(smem);
        for (int64_t it = 0; it < my_steps; ++it) {
          mbar_wait_cluster(&tmem_empty_bar[acc], acc_phase ^ 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          GASFM_PTRACE(1, it, 0);
          const uint32_t d_tmem = tmem_base + (uint32_t)(acc * kPN);
          for (int kb = 0; kb < nkb; ++kb) {
            mbar_wait(&full_bar[stage], phase);
            mbar_wait(&split_bar[stage], phase);
            mbar_wait(&peer_bar[stage], phase);
            if (kb < 12) GASFM_PTRACE(1, it, 1 + kb);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a_hi = s_base + (uint32_t)stage * kCStageBytes, a_lo = a_hi + kFATileBytes;
            const uint32_t b_hi = a_hi + kPSlotBytes, b_lo = b_hi + kPBPlaneBytes;
#pragma unroll
            for (int k = 0; k < kFBlockK / kFUmmaK; ++k) {
              const uint32_t koff = k * kFUmmaK * 2;
              umma_f16_pair(d_tmem, make_desc(a_hi + koff), make_desc(b_hi + koff), idesc, (kb == 0 && k == 0) ? 0u : 1u);
              umma_f16_pair(d_tmem, make_desc(a_lo + koff), make_desc(b_hi + koff), idesc, 1u);
              umma_f16_pair(d_tmem, make_desc(a_hi + koff), make_desc(b_lo + koff), idesc, 1u);
            }
            umma_commit_pair(&empty_bar[stage]);
            if (kb == nkb - 1) umma_commit_pair(&tmem_full_bar[acc]);
            if (++stage == kCStages) { stage = 0; phase ^= 1; }
          }
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
      }
    }
  } else if (warp < kEpiWarp0) {
    // ===================== A producers: K blocks of this CTA's 128 rows stream through four register buffers =====================
    if constexpr (PWG == 4) asm volatile("setmaxnreg.inc.sync.aligned.u32 88;" ::: "memory");
    else asm volatile("setmaxnreg.inc.sync.aligned.u32 184;" ::: "memory");
    const int t = threadIdx.x - 128;
    const int q = t & 15, rg = t >> 4;
    int stage = 0; uint32_t phase = 0;
    uint32_t soff[kRows];
#pragma unroll
    for (int i = 0; i < kRows; ++i) {
      const int row = rg + kRowStep * i;
      soff[i] = (uint32_t)(row * 128 + ((((q >> 1) ^ (row & 7))) << 4) + ((q & 1) << 3));
    }
    const uint32_t smem_base = smem_u32(smem);
    int64_t tr_it = 0; int tr_kb = 0;                            // (profiling only)
    auto convert_block = [&](const float4 (&v)[kRows], const float (&scale)[kRows]) {
      if (t == 0) mbar_wait_backoff(&empty_bar[stage], phase ^ 1, 32);        // one polling lane, the rest block on the named barrier
      asm volatile("bar.sync 1, %0;" ::"n"(kProducers) : "memory");
      if (t == 0) { GASFM_PTRACE(0, tr_it, tr_kb); }
      const uint32_t sb = smem_base + (uint32_t)stage * (uint32_t)kCStageBytes;
#pragma unroll
      for (int i = 0; i < kRows; ++i) {
        const float s = scale[i];
        const float x0 = v[i].x * s, x1 = v[i].y * s, x2 = v[i].z * s, x3 = v[i].w * s;
        const __half2 h01 = __floats2half2_rn(x0, x1), h23 = __floats2half2_rn(x2, x3);
        const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
        const __half2 l01 = __floats2half2_rn(x0 - f01.x, x1 - f01.y), l23 = __floats2half2_rn(x2 - f23.x, x3 - f23.y);
        const uint32_t addr = sb + soff[i];
        asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(*reinterpret_cast<const uint32_t*>(&h01)),
                     "r"(*reinterpret_cast<const uint32_t*>(&h23)) : "memory");
        asm volatile("st.shared.v2.b32 [%0+%3], {%1, %2};" ::"r"(addr), "r"(*reinterpret_cast<const uint32_t*>(&l01)),
                     "r"(*reinterpret_cast<const uint32_t*>(&l23)), "n"(kFATileBytes) : "memory");
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive(&split_bar[stage]);
      if (TRACE) { if (++tr_kb == nkb) { tr_kb = 0; ++tr_it; } }
      if (++stage == kCStages) { stage = 0; phase ^= 1; }
    };
    auto load_kbg = [&](int64_t it, int kbg, float4 (&v)[kRows]) {
      if (kbg >= nkb) { kbg -= nkb; ++it; }                     // the stream runs across tiles
      const int seg = kbg / kPKB, kcol = (kbg % kPKB) * kFBlockK + q * 4;
      const float* base = seg == 0 ? p.A_seg[0] : (seg == 1 ? p.A_seg[1] : (seg == 2 ? p.A_seg[2] : p.A_seg[3]));
      const int64_t ld = seg == 0 ? p.lda_seg[0] : (seg == 1 ? p.lda_seg[1] : (seg == 2 ? p.lda_seg[2] : p.lda_seg[3]));
      const int64_t row0 = ((cluster_id + it * num_clusters) * 2 + cta_rank) * kFBlockM;
#pragma unroll
      for (int i = 0; i < kRows; ++i) {
        const int64_t row = row0 + rg + kRowStep * i;
        v[i] = (it < my_steps && row < p.M) ? ld_stream4(base + row * ld + kcol) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    float seg_seen[4] = {0.f, 0.f, 0.f, 0.f};
    float4 v0[kRows], v1[kRows], v2[kRows], v3[kRows];
    // row scales of tile ``it`` from the upstream row maxima.  The device timeline of the first version showed the MMA stream
    // idle for ~12 k cycles at the start of EVERY tile: these loads (cold, queued behind 96 KB of streaming operand loads) sat
    // between two tiles.  Now the next tile's maxima are prefetched into L2 at the top of a tile and turned into scales before
    // the tile's LAST conversion, while the MMA still has two staged K blocks to work on.
    auto tile_scales = [&](int64_t it, float (&scale)[kRows]) {
      const int64_t row0 = ((cluster_id + it * num_clusters) * 2 + cta_rank) * kFBlockM;
      float* descale_slot = row_descale[it & (kFScaleSlots - 1)];
#pragma unroll
      for (int i = 0; i < kRows; ++i) {
        const int64_t row = row0 + rg + kRowStep * i;
        float m = 0.f;
#pragma unroll
        for (int seg = 0; seg < 4; ++seg) {
          if (seg < p.n_seg && it < my_steps && row < p.M) {
            const float r = __ldg(p.rowmax_seg[seg] + row);
            m = fmaxf(m, r);
            seg_seen[seg] = fmaxf(seg_seen[seg], r);
          }
        }
        float descale;
        row_scale_from_amax(m, scale[i], descale);
        if (q == 0) descale_slot[rg + kRowStep * i] = descale;
      }
    };
    auto prefetch_rowmax = [&](int64_t it) {                    // 128 rows x 4 B = 4 lines per segment
      const int64_t row = ((cluster_id + it * num_clusters) * 2 + cta_rank) * kFBlockM + (t & 3) * 32;
      const int seg = t >> 2;
      if (t < 16 && seg < p.n_seg && it < my_steps && row < p.M) {
        const float* src = (seg == 0 ? p.rowmax_seg[0] : (seg == 1 ? p.rowmax_seg[1] : (seg == 2 ? p.rowmax_seg[2] : p.rowmax_seg[3]))) + row;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(src));
      }
    };
    load_kbg(0, 0, v0); load_kbg(0, 1, v1); load_kbg(0, 2, v2);
    float scale[kRows];
    tile_scales(0, scale);
    for (int64_t it = 0; it < my_steps; ++it) {
      prefetch_rowmax(it + 1);
      float scale_next[kRows];
      for (int kbg = 0; kbg < nkb; kbg += 4) {                  // nkb % 4 == 0
        load_kbg(it, kbg + 3, v3);
        convert_block(v0, scale);
        load_kbg(it, kbg + 4, v0);
        convert_block(v1, scale);
        load_kbg(it, kbg + 5, v1);
        convert_block(v2, scale);
        load_kbg(it, kbg + 6, v2);
        if (kbg + 4 >= nkb) tile_scales(it + 1, scale_next);
        convert_block(v3, scale);
      }
#pragma unroll
      for (int i = 0; i < kRows; ++i) scale[i] = scale_next[i];
    }
    if (p.seg_amax != nullptr) {
#pragma unroll
      for (int seg = 0; seg < 4; ++seg)
        if (seg < p.n_seg) warp_amax_to_global(seg_seen[seg], p.seg_amax + seg);
    }
  } else {
    // ===================== epilogue: this CTA's 128 accumulator rows =====================
    if constexpr (PWG == 4) asm volatile("setmaxnreg.inc.sync.aligned.u32 96;" ::: "memory");      // launched with 80
    else asm volatile("setmaxnreg.dec.sync.aligned.u32 96;" ::: "memory");
    const int quarter = warp & 3;
    const uint32_t stg = smem_u32(c_stage + (warp - kEpiWarp0) * 4096);
    const int pc = lane & 3, rsub = lane >> 2;
    const uint32_t sts_base = stg + (uint32_t)(lane * 128 + ((lane & 7) << 4));
    const uint32_t lds_v = stg + (uint32_t)(rsub * 128 + (((2 * pc) ^ rsub) << 4)), lds_w = lds_v ^ 16u;
    const uint32_t csg = smem_u32(bscale_s) + (uint32_t)(pc * 32), bsg = smem_u32(bias_s) + (uint32_t)(pc * 32);
    int acc = 0; uint32_t acc_phase = 0;
    for (int64_t it = 0; it < my_steps; ++it) {
      const int64_t row0 = ((cluster_id + it * num_clusters) * 2 + cta_rank) * kFBlockM + quarter * 32;
      const int rows_left = (int)((p.M - row0) < 32 ? (p.M - row0) : 32) - rsub;
      float rs[4];
      float* rowp[4];
      mbar_wait_backoff(&tmem_full_bar[acc], acc_phase, 32);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (warp == kEpiWarp0 && lane == 0) GASFM_PTRACE(2, it, 0);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        rs[i] = row_descale[it & (kFScaleSlots - 1)][quarter * 32 + 8 * i + rsub];
        rowp[i] = p.C + (row0 + 8 * i + rsub) * p.ldc + 8 * pc;
      }
      const uint32_t taddr0 = tmem_base + (uint32_t)(acc * kPN) + ((uint32_t)(quarter * 32) << 16);
      pair_epilogue_drain<TRACE>(p, taddr0, sts_base, lds_v, lds_w, csg, bsg, rs, rowp, rows_left, (int)it, warp == kEpiWarp0 && lane == 0);
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive_remote_relaxed(&tmem_empty_bar[acc], 0);
      if (warp == kEpiWarp0 && lane == 0) GASFM_PTRACE(2, it, 1);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  __syncthreads();
  cluster_sync();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

