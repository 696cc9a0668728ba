// C[M,N] = A[M,K] * B[N,K]^T (+ bias), fp32 in / fp32 out, on the tcgen05 tensor cores with a SCALED 2 x FP16 split
//     A*B ~= A_hi*B_hi + A_lo*B_hi + A_hi*B_lo,     x_hi = fp16(s x), x_lo = fp16(s x - x_hi)
// kind::f16 runs at twice the kind::tf32 rate, and fp16 carries the same 11 significant bits as tf32, so the
// three-product split keeps the ~2^-22 relative accuracy of the 3xTF32 kernel (gemm_tf32x3.cu) at half the tensor
// time -- which moves these projections from tensor-bound to HBM-bound (read A once, write C once).
//
// fp16 has a 5-bit exponent, so every operand row is scaled by a power of two (exact) before the split:
//   * A rows (observations):  s_m = 2^(14 - floor(log2 max_k |A[m,k]|)), computed by the producer warps from the
//     row they hold in registers;  |s_m A[m,:]| < 2^15, the low part is >= 2^-24 (fp16 subnormal quantum), i.e. the
//     split is exact to 2^-38 of the row maximum.
//   * B rows (weights):       t_n likewise, by gasfm_split_f16 (once per weight and step).
// The epilogue multiplies the accumulator by 1/(s_m t_n) (a power of two: exact) and adds the bias.
//
// Call sites: lin_l of the two GATv2 graphs, lin_r, lin_proj and the matching input gradients
// (reference: code/models/layers.py:329,426,941 through torch.nn.functional.linear).
//
// Structure: CTA pairs (B halves multicast), persistent over M tiles, 16 warps in 4 warpgroups (128 registers per
// thread at launch, redistributed with setmaxnreg):
//   WG0  warp 0 TMA producer of B_hi/B_lo [N x 64] fp16 K-blocks, warp 1 TMEM allocation + MMA issue  (40 registers)
//   WG1-2  A producers: the whole [128 x K] fp32 tile lives in registers (one LDG per element, no second pass
//          for the row maximum); scaled, split and written as fp16 in the 128B-swizzled K-major layout; each
//          K-block's registers are refilled with the NEXT tile as soon as they are consumed      (184 registers)
//   WG3  epilogue: tcgen05.ld, descale, bias, swizzled staging tile, 256-bit global stores (+= C)     (96 registers)
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include "common.cuh"
#include "umma.cuh"
#include "../../include/gasfm_b200.h"

namespace gasfm {

constexpr int kFBlockM = 128;
constexpr int kFBlockK = 64;                          // 64 fp16 = 128 bytes = one SWIZZLE_128B atom row
constexpr int kFUmmaK = 16;                           // fp16: 32 bytes per MMA along K
constexpr int kFStages = 2;
constexpr int kFThreads = 512;
constexpr int kFATileBytes = kFBlockM * kFBlockK * 2;   // 16 KB per plane
constexpr int kFScaleSlots = 8;                       // row-scale ring (tiles in flight between producers and epilogue)
constexpr int kFCluster = 2;
constexpr int kFMaxGroups = 3;                        // projections of the same A computed per M tile

struct GemmF16Args {
  const float* A; int64_t lda; const float* b_scale; const float* bias; float* C; int64_t ldc; int64_t M; int N; int K;
  int groups;                                 // B = [groups * N, K] stacked weights, C column offset g * N: one pass over A
  int tmem_cols; int accumulate; int debug;   // debug: phase-isolation bits for profiling (GASFM_GEMM_DEBUG)
  float* a_amax;                              // optional: max |A| over the whole matrix (atomicMax; zeroed by the launcher)
  // LN variant: the operand is relu(layer_norm(A) * gamma + beta), evaluated on the register-resident tile (A = x_raw is
  // read once, the normalised matrix never exists in memory); the row statistics are written out for backward
  const float* ln_gamma; const float* ln_beta; float ln_eps; float* ln_mean; float* ln_rstd;
  float* ln_y; int64_t ldy;                   // optional (pair kernel): the normalised operand itself, for the weight gradients
  // CAT variant: A = [A_0 | A_1 | ..] (n_seg matrices of seg_k = 64 KB columns each, separate buffers), K = n_seg * seg_k.
  // The tile no longer fits the registers, so the operand producer makes two passes over it: row maxima first (one scale per
  // row across all segments), then scale + split K block by K block; the second pass and -- through an L2 prefetch issued one
  // tile ahead -- most of the first are served by L2.  seg_amax[n_seg] receives max |A_i| (for the fp16 weight gradient).
  const float* A_seg[4]; int64_t lda_seg[4]; int n_seg; float* seg_amax;
  // ... unless the producers of the A_i left their row maxima behind (rowmax_seg[i][M], e.g. gasfm_gat_edge_bwd_rowmax,
  // gasfm_x0_bwd_rowmax): then ONE pass suffices, with four K blocks of loads in flight ahead of the conversion
  const float* rowmax_seg[4];
  long long* trace;                           // optional [3 roles][kTraceTiles][16] SM-clock timestamps of CTA 0 (profiling)
};

constexpr int kTraceTiles = 16;
#define GASFM_TRACE(role, it, slot)                                                                  \
  do {                                                                                               \
    if (p.trace && blockIdx.x == 0 && (it) < kTraceTiles) p.trace[((role) * kTraceTiles + (it)) * 16 + (slot)] = clock64(); \
  } while (0)

// KB: number of 64-wide K blocks (K <= 64 * KB; CAT: per segment); LN: LayerNorm + ReLU on the A tile; CAT: concatenated A
template <int KB, bool LN = false, bool CAT = false>
__global__ void __cluster_dims__(kFCluster, 1, 1) __launch_bounds__(kFThreads, 1)
gemm_f16x2_kernel(const __grid_constant__ CUtensorMap map_bhi, const __grid_constant__ CUtensorMap map_blo, GemmF16Args p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [stage][A_hi 16K | A_lo 16K | B_hi N*128 | B_lo N*128] (1024-aligned), then the epilogue staging tiles
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int b_tile_bytes = p.N * kFBlockK * 2;
  const int stage_bytes = 2 * kFATileBytes + 2 * b_tile_bytes;
  uint8_t* c_stage = smem + (size_t)kFStages * stage_bytes;      // 4 epilogue warps x [32 rows x 128 B]
  __shared__ uint64_t full_bar[kFStages], split_bar[kFStages], empty_bar[kFStages], tmem_full_bar[2], tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_slot;
  __shared__ __align__(16) float bias_s[kFMaxGroups * 256], bscale_s[kFMaxGroups * 256];   // [group][256]
  __shared__ float row_descale[kFScaleSlots][kFBlockM];
  __shared__ __align__(16) float ln_gamma_s[LN ? 256 : 4], ln_beta_s[LN ? 256 : 4];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t num_tiles = (p.M + kFBlockM - 1) / kFBlockM;
  const uint32_t cta_rank = cluster_ctarank();
  const int64_t num_clusters = gridDim.x / kFCluster, cluster_id = blockIdx.x / kFCluster;
  const int64_t num_pairs = (num_tiles + kFCluster - 1) / kFCluster;
  // BOTH CTAs of a pair run every step (a CTA whose tile is past the end feeds zeros and stores nothing) so that
  // the shared pipeline of multicast loads never deadlocks
  const int64_t my_steps = cluster_id < num_pairs ? (num_pairs - cluster_id + num_clusters - 1) / num_clusters : 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kFStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&split_bar[s], 256); mbar_init(&empty_bar[s], kFCluster); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full_bar[a], 1); mbar_init(&tmem_empty_bar[a], 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int j = threadIdx.x; j < p.groups * 256; j += kFThreads) {
    const int g = j >> 8, c = j & 255;
    bias_s[j] = (p.bias && c < p.N) ? p.bias[g * p.N + c] : 0.f;
    bscale_s[j] = c < p.N ? p.b_scale[g * p.N + c] : 1.f;
  }
  if constexpr (LN) {
    for (int j = threadIdx.x; j < 256; j += kFThreads) {
      ln_gamma_s[j] = j < p.K ? p.ln_gamma[j] : 0.f;
      ln_beta_s[j] = j < p.K ? p.ln_beta[j] : 0.f;
    }
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync();                            // the peer's barriers are initialised before any remote arrive / multicast
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_slot;
  const int acc_cols = p.tmem_cols / 2;     // column offset of the second accumulator
  const int nkb = CAT ? KB * p.n_seg : KB;  // K blocks per tile

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;" ::: "memory");
    if (warp == 0 && lane == 0) {
      // ===================== TMA producer =====================
      int stage = 0; uint32_t phase = 0;
      const int half_rows = p.N / kFCluster;                       // B rows this CTA fetches and multicasts
      const int half_bytes = half_rows * kFBlockK * 2;
      for (int64_t vt = 0; vt < my_steps * p.groups; ++vt) {       // virtual tile = (M tile, group)
        const int b_row0 = (int)(vt % p.groups) * p.N + (int)cta_rank * half_rows;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);                 // both CTAs have retired the MMAs of this stage
          uint8_t* st = smem + (size_t)stage * stage_bytes;
          if (p.debug & 4) { mbar_arrive(&full_bar[stage]); if (++stage == kFStages) { stage = 0; phase ^= 1; } continue; }
          mbar_expect_tx(&full_bar[stage], 2 * b_tile_bytes);      // halves from both CTAs land here
          tma_load_2d_mc(st + 2 * kFATileBytes + cta_rank * half_bytes, &map_bhi, &full_bar[stage], kb * kFBlockK, b_row0,
                         (uint16_t)((1u << kFCluster) - 1));
          tma_load_2d_mc(st + 2 * kFATileBytes + b_tile_bytes + cta_rank * half_bytes, &map_blo, &full_bar[stage], kb * kFBlockK,
                         b_row0, (uint16_t)((1u << kFCluster) - 1));
          if (++stage == kFStages) { stage = 0; phase ^= 1; }
        }
      }
    } else if (warp == 1 && lane == 0) {
      // ===================== MMA issuer =====================
      // instruction descriptor (cute::UMMA::InstrDescriptor): D = f32 (bit 4), A = B = f16 (format 0), K-major, N>>3, M>>4
      const uint32_t idesc = (1u << 4) | ((uint32_t)(p.N >> 3) << 17) | ((uint32_t)(kFBlockM >> 4) << 24);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int64_t it = 0; it < my_steps * p.groups; ++it) {       // virtual tiles
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        GASFM_TRACE(1, it, 0);
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * acc_cols);
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&full_bar[stage], phase);       // B_hi / B_lo landed (TMA)
          if (!CAT) GASFM_TRACE(1, it, 1 + kb);
          mbar_wait(&split_bar[stage], phase);      // A_hi / A_lo written by the producer warps
          if (!CAT) GASFM_TRACE(1, it, 5 + kb);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t a_hi = smem_u32(smem + (size_t)stage * stage_bytes);
          const uint32_t a_lo = a_hi + kFATileBytes;
          const uint32_t b_hi = a_hi + 2 * kFATileBytes;
          const uint32_t b_lo = b_hi + b_tile_bytes;
#pragma unroll
          for (int k = 0; k < kFBlockK / kFUmmaK; ++k) {
            if (p.debug & 2) break;
            const uint32_t koff = k * kFUmmaK * 2;   // bytes along K inside the 128-byte swizzle row
            const uint32_t first = (kb == 0 && k == 0) ? 0u : 1u;
            umma_f16(d_tmem, make_desc(a_hi + koff), make_desc(b_hi + koff), idesc, first);
            umma_f16(d_tmem, make_desc(a_lo + koff), make_desc(b_hi + koff), idesc, 1u);
            umma_f16(d_tmem, make_desc(a_hi + koff), make_desc(b_lo + koff), idesc, 1u);
          }
          umma_commit_mc(&empty_bar[stage], (uint16_t)((1u << kFCluster) - 1));   // frees the stage in BOTH CTAs' eyes
          if (kb == nkb - 1) umma_commit(&tmem_full_bar[acc]);
          if (!CAT) GASFM_TRACE(1, it, 9 + kb);
          if (++stage == kFStages) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp < 12) {
    // ===================== A producers (256 threads, 2 warpgroups) =====================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 184;" ::: "memory");
    // thread -> float4 q (columns 4q..4q+3 of a 64-wide K block) of rows rg + 16 i: a warp-wide load covers
    // two full 256-byte row segments.  buf holds this thread's share of the WHOLE tile.
    const int t = threadIdx.x - 128;
    const int q = t & 15, rg = t >> 4;
    int stage = 0; uint32_t phase = 0;
    // scale, split into fp16 hi + lo and store this thread's 8 row pieces of one 64-wide K block (128B-swizzled, K-major)
    // (everything that does not change from K block to K block is computed once: shared-memory offsets of this thread's 8
    //  row pieces, the 32-bit element offsets of its rows -- the generic loop spent 70 % of its instructions on that)
    uint32_t soff[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int row = rg + 16 * i;
      soff[i] = (uint32_t)(row * 128 + ((((q >> 1) ^ (row & 7))) << 4) + ((q & 1) << 3));
    }
    const uint32_t smem_base = smem_u32(smem);
    auto convert_block = [&](const float4 (&v)[8], const float (&scale)[8]) {
      const uint32_t sb = smem_base + (uint32_t)stage * (uint32_t)stage_bytes;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
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
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to tcgen05.mma
      mbar_arrive(&split_bar[stage]);
      if (++stage == kFStages) { stage = 0; phase ^= 1; }
    };
    if constexpr (CAT) {
      static_assert(KB % 2 == 0, "the concatenated producer alternates two register buffers per K block");
      // global K block kbg = seg * KB + kb of tile it: this thread's float4 of rows rg + 16 i
      auto load_kbg = [&](int64_t it, int kbg, float4 (&v)[8]) {
        const int seg = kbg / KB, kcol = (kbg % KB) * kFBlockK + q * 4;
        const float* base = seg == 0 ? p.A_seg[0] : (seg == 1 ? p.A_seg[1] : (seg == 2 ? p.A_seg[2] : p.A_seg[3]));
        const int64_t ld = seg == 0 ? p.lda_seg[0] : (seg == 1 ? p.lda_seg[1] : (seg == 2 ? p.lda_seg[2] : p.lda_seg[3]));
        const int64_t tile = (cluster_id + it * num_clusters) * kFCluster + cta_rank;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int64_t row = tile * kFBlockM + rg + 16 * i;
          v[i] = (it < my_steps && kbg < nkb && row < p.M) ? ld_stream4(base + row * ld + kcol) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      };
      auto absmax4 = [](const float4& a) { return fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(a.z), fabsf(a.w))); };
      float seg_seen[4] = {0.f, 0.f, 0.f, 0.f};
      float4 v0[8], v1[8], v2[8], v3[8];
      const int lines_per_row = KB * kFBlockK * 4 / 128;          // 128-byte lines of one segment row
      if (p.rowmax_seg[0] != nullptr) {
        // ---- single pass: row scales from the row maxima the upstream kernels emitted; K blocks stream through four
        //      register buffers, three blocks of loads in flight while one is converted (the stream runs across tiles) ----
        auto load_stream = [&](int64_t it, int kbg, float4 (&v)[8]) {         // kbg may run past the tile: next tile's blocks
          if (kbg >= nkb) { kbg -= nkb; ++it; }
          load_kbg(it, kbg, v);
        };
        load_stream(0, 0, v0); load_stream(0, 1, v1); load_stream(0, 2, v2);
        for (int64_t it = 0; it < my_steps; ++it) {
          const int64_t tile = (cluster_id + it * num_clusters) * kFCluster + cta_rank;
          float scale[8];
          float* descale_slot = row_descale[it & (kFScaleSlots - 1)];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int64_t row = tile * kFBlockM + rg + 16 * i;
            float m = 0.f;
#pragma unroll
            for (int seg = 0; seg < 4; ++seg) {
              if (seg < p.n_seg && row < p.M) {
                const float r = __ldg(p.rowmax_seg[seg] + row);
                m = fmaxf(m, r);
                seg_seen[seg] = fmaxf(seg_seen[seg], r);
              }
            }
            float descale;
            row_scale_from_amax(m, scale[i], descale);
            if (q == 0) descale_slot[rg + 16 * i] = descale;
          }
          for (int kbg = 0; kbg < nkb; kbg += 4) {               // nkb % 4 == 0 (launcher)
            load_stream(it, kbg + 3, v3);
            mbar_wait(&empty_bar[stage], phase ^ 1);
            convert_block(v0, scale);
            load_stream(it, kbg + 4, v0);
            mbar_wait(&empty_bar[stage], phase ^ 1);
            convert_block(v1, scale);
            load_stream(it, kbg + 5, v1);
            mbar_wait(&empty_bar[stage], phase ^ 1);
            convert_block(v2, scale);
            load_stream(it, kbg + 6, v2);
            mbar_wait(&empty_bar[stage], phase ^ 1);
            convert_block(v3, scale);
          }
        }
      } else
      for (int64_t it = 0; it < my_steps; ++it) {
        // L2 prefetch of the NEXT tile (its first pass then finds the lines in L2 instead of HBM)
        // (off by default: a whole prefetched tile doubles the L2 footprint and evicts the current one before its second pass --
        //  ncu: 3.4 GB read for 1.5 GB algorithmic; GASFM_GEMM_DEBUG bit 32 switches it on for A/B)
        if ((p.debug & 32) && it + 1 < my_steps) {
          const int64_t ntile = (cluster_id + (it + 1) * num_clusters) * kFCluster + cta_rank;
          const int total = kFBlockM * lines_per_row;
#pragma unroll
          for (int seg = 0; seg < 4; ++seg) {               // constant indices: the segment table stays in the parameter bank
            if (seg >= p.n_seg) break;
            for (int j = t; j < total; j += 256) {
              const int64_t row = ntile * kFBlockM + j / lines_per_row;
              if (row < p.M) {
                const float* a = p.A_seg[seg & 3] + row * p.lda_seg[seg & 3] + (j % lines_per_row) * 32;
                asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
              }
            }
          }
        }
        // pass 1: row maxima over every K block of the tile, four K blocks in flight
        float m[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) m[i] = 0.f;
#pragma unroll
        for (int seg = 0; seg < 4; ++seg) {
          if (seg >= p.n_seg) break;
          float sm = 0.f;
#pragma unroll
          for (int kb = 0; kb < KB; kb += 4) {
            load_kbg(it, seg * KB + kb, v0);
            load_kbg(it, seg * KB + kb + 1, v1);
            if (kb + 2 < KB) { load_kbg(it, seg * KB + kb + 2, v2); load_kbg(it, seg * KB + kb + 3, v3); }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float r = fmaxf(absmax4(v0[i]), absmax4(v1[i]));
              if (kb + 2 < KB) r = fmaxf(r, fmaxf(absmax4(v2[i]), absmax4(v3[i])));
              m[i] = fmaxf(m[i], r);
              sm = fmaxf(sm, r);
            }
          }
          seg_seen[seg] = fmaxf(seg_seen[seg], sm);
        }
        float scale[8];
        float* descale_slot = row_descale[it & (kFScaleSlots - 1)];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
#pragma unroll
          for (int off = 1; off < 16; off <<= 1) m[i] = fmaxf(m[i], __shfl_xor_sync(0xffffffffu, m[i], off));
          float descale;
          row_scale_from_amax(m[i], scale[i], descale);
          if (q == 0) descale_slot[rg + 16 * i] = descale;
        }
        // pass 2: K block by K block (from L2), the next block's loads in flight while this one is converted
        load_kbg(it, 0, v0);
#pragma unroll
        for (int seg = 0; seg < 4; ++seg) {
          if (seg >= p.n_seg) break;
#pragma unroll
          for (int kb = 0; kb < KB; kb += 2) {
            load_kbg(it, seg * KB + kb + 1, v1);
            mbar_wait(&empty_bar[stage], phase ^ 1);
            convert_block(v0, scale);
            load_kbg(it, seg * KB + kb + 2, v0);           // past the last block of the tile: zeros, never consumed
            mbar_wait(&empty_bar[stage], phase ^ 1);
            convert_block(v1, scale);
          }
        }
      }
      if (p.seg_amax != nullptr) {
#pragma unroll
        for (int seg = 0; seg < 4; ++seg)
          if (seg < p.n_seg) warp_amax_to_global(seg_seen[seg], p.seg_amax + seg);
      }
    } else {
    float4 buf[KB][8];
    const uint32_t toff = (uint32_t)(rg * p.lda + q * 4), tstep = (uint32_t)(16 * p.lda);   // elements (lda < 2^24: launcher)
    auto load_block = [&](int64_t it, int kb, float4 (&v)[8]) {
      const int64_t tile = (cluster_id + it * num_clusters) * kFCluster + cta_rank;
      const int kcol = kb * kFBlockK + q * 4;
      if (it < my_steps && (tile + 1) * kFBlockM <= p.M && (kb + 1) * kFBlockK <= p.K) {
        // whole block in range (all but the last tile / a ragged K): no per-row predicates, 32-bit offsets from a uniform base
        const float* base = p.A + tile * kFBlockM * p.lda + kb * kFBlockK;
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = ld_stream4(base + (toff + (uint32_t)i * tstep));
        return;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int64_t row = tile * kFBlockM + rg + 16 * i;
        v[i] = (it < my_steps && row < p.M && kcol < p.K) ? ld_stream4(p.A + row * p.lda + kcol) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
#pragma unroll
    for (int kb = 0; kb < KB; ++kb) load_block(0, kb, buf[kb]);
    float seen_max = 0.f;
    for (int64_t it = 0; it < my_steps; ++it) {
      if constexpr (LN) {
        // LayerNorm + ReLU in place on the register tile: two-pass mean / variance over the K columns of each row
        // (16 lanes share a row), then y = max(0, (x - mean) rstd gamma + beta) -- the arithmetic of ln_relu_fwd_kernel
        const int64_t tile = (cluster_id + it * num_clusters) * kFCluster + cta_rank;
        const float inv_k = 1.f / (float)p.K;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int64_t row = tile * kFBlockM + rg + 16 * i;
          float sum = 0.f;
#pragma unroll
          for (int kb = 0; kb < KB; ++kb) sum += (buf[kb][i].x + buf[kb][i].y) + (buf[kb][i].z + buf[kb][i].w);
#pragma unroll
          for (int off = 1; off < 16; off <<= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
          const float mean = sum * inv_k;
          float sq = 0.f;
#pragma unroll
          for (int kb = 0; kb < KB; ++kb) {
            if (kb * kFBlockK + q * 4 < p.K) {            // padded columns (K < 64 KB) hold zeros and do not count
              const float dx = buf[kb][i].x - mean, dy = buf[kb][i].y - mean, dz = buf[kb][i].z - mean, dw = buf[kb][i].w - mean;
              sq = fmaf(dx, dx, sq); sq = fmaf(dy, dy, sq); sq = fmaf(dz, dz, sq); sq = fmaf(dw, dw, sq);
            }
          }
#pragma unroll
          for (int off = 1; off < 16; off <<= 1) sq += __shfl_xor_sync(0xffffffffu, sq, off);
          const float rstd = 1.f / sqrtf(sq * inv_k + p.ln_eps);
          const bool live = row < p.M;                    // rows past the end stay zero (they must not enter a_amax)
#pragma unroll
          for (int kb = 0; kb < KB; ++kb) {
            const int kcol = kb * kFBlockK + q * 4;
            const float4 g = *reinterpret_cast<const float4*>(&ln_gamma_s[kcol & 255]);
            const float4 b = *reinterpret_cast<const float4*>(&ln_beta_s[kcol & 255]);
            const bool on = live && kcol < p.K;
            buf[kb][i].x = on ? fmaxf(fmaf((buf[kb][i].x - mean) * rstd, g.x, b.x), 0.f) : 0.f;
            buf[kb][i].y = on ? fmaxf(fmaf((buf[kb][i].y - mean) * rstd, g.y, b.y), 0.f) : 0.f;
            buf[kb][i].z = on ? fmaxf(fmaf((buf[kb][i].z - mean) * rstd, g.z, b.z), 0.f) : 0.f;
            buf[kb][i].w = on ? fmaxf(fmaf((buf[kb][i].w - mean) * rstd, g.w, b.w), 0.f) : 0.f;
          }
          if (q == 0 && live && it < my_steps) { p.ln_mean[row] = mean; p.ln_rstd[row] = rstd; }
        }
      }
      // row maxima -> power-of-two scales (16 lanes share a row)
      float scale[8];
      float* descale_slot = row_descale[it & (kFScaleSlots - 1)];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float m = 0.f;
#pragma unroll
        for (int kb = 0; kb < KB; ++kb)
          m = fmaxf(m, fmaxf(fmaxf(fabsf(buf[kb][i].x), fabsf(buf[kb][i].y)), fmaxf(fabsf(buf[kb][i].z), fabsf(buf[kb][i].w))));
#pragma unroll
        for (int off = 1; off < 16; off <<= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
        float descale;
        seen_max = fmaxf(seen_max, m);
        row_scale_from_amax(m, scale[i], descale);
        if (q == 0) descale_slot[rg + 16 * i] = descale;
      }
      if (t == 0) GASFM_TRACE(0, it, 0);
      for (int g = 0; g < p.groups; ++g) {            // the same A tile feeds every group: converted again, loaded once
      const bool last_group = g == p.groups - 1;
#pragma unroll
      for (int kb = 0; kb < KB; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (t == 0) GASFM_TRACE(0, it, 1 + kb);
        convert_block(buf[kb], scale);
        if (t == 0) GASFM_TRACE(0, it, 5 + kb);
        if (last_group) load_block(it + 1, kb, buf[kb]);   // refill the freed registers with the next tile's K block
      }
      }
    }
    if (p.a_amax != nullptr) warp_amax_to_global(seen_max, p.a_amax);
    }
  } else {
    // ===================== epilogue (warps 12..15 -> TMEM lane quarters 0..3) =====================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 96;" ::: "memory");
    const int quarter = warp & 3;
    uint8_t* stg = c_stage + (warp - 12) * 4096;                   // [32 rows x 128 B], 128B-swizzled
    int acc = 0; uint32_t acc_phase = 0;
    for (int64_t vt = 0; vt < my_steps * p.groups; ++vt) {
      const int64_t it = vt / p.groups;
      const int grp = (int)(vt % p.groups);
      const float* bias_g = bias_s + grp * 256;
      const float* bscale_g = bscale_s + grp * 256;
      float* c_grp = p.C + (int64_t)grp * p.N;
      const int64_t tile = (cluster_id + it * num_clusters) * kFCluster + cta_rank;
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (warp == 12 && lane == 0) GASFM_TRACE(2, vt, 0);
      const int row0 = (int)(tile * kFBlockM) + quarter * 32;       // first C row of this warp's 32-row slab
      const float rs = row_descale[it & (kFScaleSlots - 1)][quarter * 32 + lane];   // lane = row of the slab
      const uint32_t taddr0 = tmem_base + (uint32_t)(acc * acc_cols) + ((uint32_t)(quarter * 32) << 16);
      // 32-column chunks: TMEM -> registers (lane = row) -> descale + bias -> swizzled staging tile -> registers
      // (4 lanes = one 128-byte row segment) -> 256-bit global stores, 8 full lines per instruction.  All warp-local:
      // no async proxy, no fences.  (Measured alternatives: TMA bulk stores of 4 KB / 2 KB tiles, 128-bit stores,
      // stores straight from the lane = row registers, eight epilogue warps -- all slower or equal; the epilogue
      // is paced by the memory system's write throughput under the concurrent A reads, see DESIGN.md.)
      for (int c0 = 0; c0 < p.N; c0 += 32) {
        uint32_t r[32];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
              "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
              "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr0 + (uint32_t)c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        __syncwarp();                                  // the previous chunk's read-back is complete
#pragma unroll
        for (int j = 0; j < 32; j += 4) {              // lane = row; 16-byte chunk ^= row % 8: conflict-free
          const float4 cs = *reinterpret_cast<const float4*>(&bscale_g[(c0 + j) & 255]);   // same address in every lane: broadcast
          const float4 bs = *reinterpret_cast<const float4*>(&bias_g[(c0 + j) & 255]);
          float4 v;
          v.x = fmaf(__uint_as_float(r[j]), rs * cs.x, bs.x); v.y = fmaf(__uint_as_float(r[j + 1]), rs * cs.y, bs.y);
          v.z = fmaf(__uint_as_float(r[j + 2]), rs * cs.z, bs.z); v.w = fmaf(__uint_as_float(r[j + 3]), rs * cs.w, bs.w);
          *reinterpret_cast<float4*>(stg + lane * 128 + ((((j >> 2) ^ (lane & 7))) << 4)) = v;
        }
        __syncwarp();
        const int pc = lane & 3, col = c0 + 8 * pc;    // this lane's 8 columns of rows (lane / 4) + 8 i
        if (col >= p.N || (p.debug & 1)) continue;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int rr = 8 * i + (lane >> 2);
          const int64_t grow = (int64_t)row0 + rr;
          float4 v = *reinterpret_cast<const float4*>(stg + rr * 128 + (((2 * pc) ^ (rr & 7)) << 4));
          float4 w = *reinterpret_cast<const float4*>(stg + rr * 128 + (((2 * pc + 1) ^ (rr & 7)) << 4));
          if (grow < p.M) {
            float* dst = c_grp + grow * p.ldc + col;
            if (p.accumulate) {
              // C += : fire-and-forget vector reductions resolved in L2 (a read-add-write here would expose one
              // memory round trip per chunk; every element has a single writer per launch, so the result is deterministic)
              asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(dst), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
              asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(dst + 4), "f"(w.x), "f"(w.y), "f"(w.z), "f"(w.w) : "memory");
            } else {
              asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w),
                           "f"(w.x), "f"(w.y), "f"(w.z), "f"(w.w)
                           : "memory");
            }
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(&tmem_empty_bar[acc]);
      if (warp == 12 && lane == 0) GASFM_TRACE(2, vt, 1);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  __syncthreads();
  cluster_sync();                            // no CTA leaves while its peer may still multicast into its shared memory
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

#include "gemm_f16x2_pair.cuh"   // gemm_f16x2_pair_kernel, gemm_f16x2_cat_pair_kernel (cta_group::2)

// One warp per weight row n: t_n = 2^(14 - floor(log2 max_k |W[n,k]|)); hi = fp16(t W), lo = fp16(t W - hi); descale[n] = 1/t_n
__global__ void __launch_bounds__(256) split_f16_kernel(const float* __restrict__ w, int n_rows, int k, __half* __restrict__ hi,
                                                        __half* __restrict__ lo, float* __restrict__ descale) {
  const int row = (int)((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (row >= n_rows) return;
  const float* src = w + (int64_t)row * k;
  float m = 0.f;
  for (int j = lane; j < k; j += 32) m = fmaxf(m, fabsf(src[j]));
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
  float s, d;
  row_scale_from_amax(m, s, d);
  for (int j = lane; j < k; j += 32) {
    const float x = src[j] * s;
    const __half h = __float2half_rn(x);
    hi[(int64_t)row * k + j] = h;
    lo[(int64_t)row * k + j] = __float2half_rn(x - __half2float(h));
  }
  if (lane == 0) descale[row] = d;
}

}  // namespace gasfm

using namespace gasfm;

static long long* g_trace = nullptr;
// profiling hook (tools/gemm_trace.py): device buffer of 3 * 16 * 16 int64 that the next launches fill with the
// SM-clock timestamps of CTA 0's producer / MMA / epilogue milestones; NULL switches tracing off
extern "C" int gasfm_debug_set_gemm_trace(void* dev_buffer) { g_trace = (long long*)dev_buffer; return 0; }

extern "C" int gasfm_split_f16(const float* w, int n_rows, int k, void* hi, void* lo, float* descale, void* stream) {
  GASFM_REQUIRE(n_rows >= 0 && k > 0, "split_f16: bad shape");
  if (n_rows == 0) return 0;
  split_f16_kernel<<<ceil_div((int64_t)n_rows * 32, 256), 256, 0, (cudaStream_t)stream>>>(w, n_rows, k, (__half*)hi, (__half*)lo, descale);
  return check_launch("split_f16");
}

extern "C" int gasfm_linear_f16x2_supported(int64_t M, int N, int K, int64_t lda, int64_t ldc) {
  // K < 64 would leave half of every 64-wide K block (and of the producer lanes) empty: the 3xTF32 kernel with its
  // 32-wide blocks is the better fit there
  return (M > 0 && N >= 16 && N <= 256 && N % 16 == 0 && K >= 64 && K <= 256 && K % 8 == 0 && lda % 4 == 0 && lda < (1 << 24) && ldc % 4 == 0) ? 1 : 0;
}

static int linear_f16x2_impl(const float* A, int64_t lda, const void* B_hi, const void* B_lo, const float* b_descale,
                            const float* bias, float* C, int64_t ldc, int64_t M, int N, int K, int groups, int accumulate,
                            float* a_amax, const float* ln_gamma, const float* ln_beta, float ln_eps, float* ln_mean,
                            float* ln_rstd, float* ln_y, int64_t ldy, void* stream) {
  GASFM_REQUIRE(gasfm_linear_f16x2_supported(M, N, K, lda, ldc), "linear_f16x2: unsupported shape M=%lld N=%d K=%d lda=%lld ldc=%lld",
                (long long)M, N, K, (long long)lda, (long long)ldc);
  GASFM_REQUIRE(groups >= 1 && groups <= kFMaxGroups && ldc >= (int64_t)groups * N, "linear_f16x2: 1..%d groups, ldc >= groups * N", kFMaxGroups);
  GASFM_REQUIRE(b_descale != nullptr && ((uintptr_t)A | (uintptr_t)B_hi | (uintptr_t)B_lo | (uintptr_t)C) % 16 == 0,
                "linear_f16x2: pointers must be 16-byte aligned");
  const bool ln = ln_gamma != nullptr;
  GASFM_REQUIRE(!ln || (ln_beta && ln_mean && ln_rstd), "linear_f16x2_ln: gamma, beta, mean and rstd are all required");
  CUtensorMap mh, ml;
  if (make_map_f16(&mh, B_hi, (int64_t)groups * N, K, K, N / kFCluster, kFBlockK) ||
      make_map_f16(&ml, B_lo, (int64_t)groups * N, K, K, N / kFCluster, kFBlockK)) return 1;
  int tmem_cols = 32;
  while (tmem_cols < 2 * N) tmem_cols <<= 1;
  const size_t smem = (size_t)kFStages * (2 * kFATileBytes + 2 * (size_t)N * kFBlockK * 2) + 4 * 4096 + 1024;
  const int kb = (K + kFBlockK - 1) / kFBlockK;
  const int64_t tiles = (M + kFBlockM - 1) / kFBlockM;
  const int64_t pairs = (tiles + kFCluster - 1) / kFCluster;
  const int grid = (int)(pairs < kNumSMs / kFCluster ? pairs : kNumSMs / kFCluster) * kFCluster;
  static int debug = -1;
  if (debug < 0) { const char* env = getenv("GASFM_GEMM_DEBUG"); debug = env ? atoi(env) : 0; }
  static int use_pair = -1;                   // GASFM_GEMM_PAIR=0: the cta_group::1 kernel for every shape (A/B switch)
  if (use_pair < 0) { const char* env = getenv("GASFM_GEMM_PAIR"); use_pair = env ? atoi(env) : 1; }
  if (a_amax != nullptr) cudaMemsetAsync(a_amax, 0, sizeof(float), (cudaStream_t)stream);
  GemmF16Args args{};
  args.A = A; args.lda = lda; args.b_scale = b_descale; args.bias = bias; args.C = C; args.ldc = ldc; args.M = M; args.N = N; args.K = K;
  args.groups = groups; args.tmem_cols = tmem_cols; args.accumulate = accumulate; args.debug = debug; args.a_amax = a_amax;
  args.ln_gamma = ln_gamma; args.ln_beta = ln_beta; args.ln_eps = ln_eps; args.ln_mean = ln_mean; args.ln_rstd = ln_rstd; args.trace = g_trace;
  const bool pair_ok = use_pair && !accumulate && N == kPN && K == kPKB * kFBlockK && ldc % 8 == 0 && (uintptr_t)C % 32 == 0 && debug == 0;
  GASFM_REQUIRE(ln_y == nullptr || (ln && pair_ok && ldy >= K && ldy % 4 == 0 && (uintptr_t)ln_y % 16 == 0),
                "linear_f16x2_ln: the normalised operand can only be written by the N = K = 256 kernel (see gasfm_linear_f16x2_ln_y_supported)");
  args.ln_y = ln_y; args.ldy = ldy;
  if (pair_ok) {
    // the shipped block shape: CTA pairs on one cta_group::2 MMA stream, A converted once per tile for all groups
    const int64_t pair_tiles = (M + 2 * kFBlockM - 1) / (2 * kFBlockM);
    const int pgrid = (int)(pair_tiles < kNumSMs / 2 ? pair_tiles : kNumSMs / 2) * 2;
    // GASFM_GEMM_EPI=direct: epilogue stores straight from the 16x256b TMEM fragments (A/B switch).  Measured slower than the
    // staged default (5.73 vs 5.44 ms at cfg3): a store instruction that touches 8 lines costs the L1 data pipe 8 wavefronts
    // whether it carries 256 bytes or 1 KB, which eats what the missing staging round trip saves.
    static int direct = -1;
    if (direct < 0) { const char* env = getenv("GASFM_GEMM_EPI"); direct = (env && env[0] == 'd') ? 1 : 0; }
    auto kernel = ln ? gemm_f16x2_pair_kernel<false, false, true>
                     : direct ? (g_trace ? gemm_f16x2_pair_kernel<true, true> : gemm_f16x2_pair_kernel<false, true>)
                              : (g_trace ? gemm_f16x2_pair_kernel<true, false> : gemm_f16x2_pair_kernel<false, false>);
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPSmemBytes);
    if (e != cudaSuccess) {
      set_error("linear_f16x2: cannot reserve %zu bytes of shared memory (%s)", kPSmemBytes, cudaGetErrorString(e));
      return (int)e;
    }
    kernel<<<pgrid, kFThreads, kPSmemBytes, (cudaStream_t)stream>>>(mh, ml, args);
    return check_launch("linear_f16x2 (pair)");
  }
#define LAUNCH_F16(KB, LNV)                                                                                                \
  do {                                                                                                                     \
    /* per-device attribute: set on every call (static smem -- barriers, scales -- also counts against the 227 KB limit) */ \
    cudaError_t e = cudaFuncSetAttribute(gemm_f16x2_kernel<KB, LNV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    if (e != cudaSuccess) {                                                                                                \
      set_error("linear_f16x2: cannot reserve %zu bytes of shared memory (%s)", smem, cudaGetErrorString(e));              \
      return (int)e;                                                                                                       \
    }                                                                                                                      \
    gemm_f16x2_kernel<KB, LNV><<<grid, kFThreads, smem, (cudaStream_t)stream>>>(mh, ml, args);                        \
  } while (0)
  if (ln) {
    switch (kb) {
      case 1: LAUNCH_F16(1, true); break;
      case 2: LAUNCH_F16(2, true); break;
      case 3: LAUNCH_F16(3, true); break;
      default: LAUNCH_F16(4, true); break;
    }
  } else {
    switch (kb) {
      case 1: LAUNCH_F16(1, false); break;
      case 2: LAUNCH_F16(2, false); break;
      case 3: LAUNCH_F16(3, false); break;
      default: LAUNCH_F16(4, false); break;
    }
  }
#undef LAUNCH_F16
  return check_launch("linear_f16x2");
}

extern "C" int gasfm_linear_f16x2_cat_supported(int64_t M, int N, int n_seg, int seg_k, int64_t ldc) {
  return (M > 0 && N >= 16 && N <= 256 && N % 16 == 0 && n_seg >= 1 && n_seg <= 4 && (seg_k == 128 || seg_k == 256) && ldc % 4 == 0) ? 1 : 0;
}

static int linear_f16x2_cat_impl(const float* const* A, const int64_t* lda, const float* const* rowmax, int n_seg, int seg_k,
                                const void* B_hi, const void* B_lo, const float* b_descale, const float* bias, float* C,
                                int64_t ldc, int64_t M, int N, float* a_amax, void* stream);

extern "C" int gasfm_linear_f16x2_cat(const float* const* A, const int64_t* lda, int n_seg, int seg_k, const void* B_hi,
                                      const void* B_lo, const float* b_descale, const float* bias, float* C, int64_t ldc,
                                      int64_t M, int N, float* a_amax, void* stream) {
  return linear_f16x2_cat_impl(A, lda, nullptr, n_seg, seg_k, B_hi, B_lo, b_descale, bias, C, ldc, M, N, a_amax, stream);
}

extern "C" int gasfm_linear_f16x2_cat_rowmax(const float* const* A, const int64_t* lda, const float* const* rowmax, int n_seg,
                                             int seg_k, const void* B_hi, const void* B_lo, const float* b_descale,
                                             const float* bias, float* C, int64_t ldc, int64_t M, int N, float* a_amax,
                                             void* stream) {
  GASFM_REQUIRE(rowmax != nullptr && (n_seg * (seg_k / 64)) % 4 == 0, "linear_f16x2_cat_rowmax: needs row maxima and a multiple of 4 K blocks");
  return linear_f16x2_cat_impl(A, lda, rowmax, n_seg, seg_k, B_hi, B_lo, b_descale, bias, C, ldc, M, N, a_amax, stream);
}

static int linear_f16x2_cat_impl(const float* const* A, const int64_t* lda, const float* const* rowmax, int n_seg, int seg_k,
                                const void* B_hi, const void* B_lo, const float* b_descale, const float* bias, float* C,
                                int64_t ldc, int64_t M, int N, float* a_amax, void* stream) {
  GASFM_REQUIRE(A && lda && gasfm_linear_f16x2_cat_supported(M, N, n_seg, seg_k, ldc), "linear_f16x2_cat: unsupported shape M=%lld N=%d "
                "n_seg=%d seg_k=%d", (long long)M, N, n_seg, seg_k);
  GASFM_REQUIRE(b_descale != nullptr && ((uintptr_t)B_hi | (uintptr_t)B_lo | (uintptr_t)C) % 16 == 0, "linear_f16x2_cat: pointers must be 16-byte aligned");
  const int K = n_seg * seg_k;
  CUtensorMap mh, ml;
  if (make_map_f16(&mh, B_hi, N, K, K, N / kFCluster, kFBlockK) || make_map_f16(&ml, B_lo, N, K, K, N / kFCluster, kFBlockK)) return 1;
  int tmem_cols = 32;
  while (tmem_cols < 2 * N) tmem_cols <<= 1;
  const size_t smem = (size_t)kFStages * (2 * kFATileBytes + 2 * (size_t)N * kFBlockK * 2) + 4 * 4096 + 1024;
  const int64_t tiles = (M + kFBlockM - 1) / kFBlockM;
  const int64_t pairs = (tiles + kFCluster - 1) / kFCluster;
  const int grid = (int)(pairs < kNumSMs / kFCluster ? pairs : kNumSMs / kFCluster) * kFCluster;
  if (a_amax != nullptr) cudaMemsetAsync(a_amax, 0, n_seg * sizeof(float), (cudaStream_t)stream);
  GemmF16Args args{};
  args.b_scale = b_descale; args.bias = bias; args.C = C; args.ldc = ldc; args.M = M; args.N = N; args.K = K; args.groups = 1;
  args.tmem_cols = tmem_cols; args.n_seg = n_seg; args.seg_amax = a_amax;
  { const char* env = getenv("GASFM_GEMM_DEBUG"); args.debug = env ? atoi(env) : 0; }   // bit 32: prefetch the next tile into L2
  for (int i = 0; i < n_seg; ++i) {
    GASFM_REQUIRE(A[i] != nullptr && (uintptr_t)A[i] % 16 == 0 && lda[i] % 4 == 0, "linear_f16x2_cat: segment %d misaligned", i);
    args.A_seg[i] = A[i]; args.lda_seg[i] = lda[i];
    if (rowmax != nullptr) {
      GASFM_REQUIRE(rowmax[i] != nullptr, "linear_f16x2_cat_rowmax: row maxima of segment %d missing", i);
      args.rowmax_seg[i] = rowmax[i];
    }
  }
  {
    static int use_pair = -1;
    if (use_pair < 0) { const char* env = getenv("GASFM_GEMM_PAIR"); use_pair = env ? atoi(env) : 1; }
    if (use_pair && rowmax != nullptr && N == kPN && seg_k == kPKB * kFBlockK && ldc % 8 == 0 && (uintptr_t)C % 32 == 0 && args.debug == 0) {
      // the block shape with row maxima at hand: CTA pairs on one cta_group::2 MMA stream, three 64 KB stages
      bool lda_ok = true;
      for (int i = 0; i < n_seg; ++i) lda_ok = lda_ok && lda[i] >= seg_k;
      if (lda_ok) {
        const int64_t pair_tiles = (M + 2 * kFBlockM - 1) / (2 * kFBlockM);
        const int pgrid = (int)(pair_tiles < kNumSMs / 2 ? pair_tiles : kNumSMs / 2) * 2;
        args.trace = g_trace;
        static int pwg = -1;                  // GASFM_GEMM_CAT_PRODUCERS=16: sixteen producer warps (A/B switch; default eight)
        if (pwg < 0) { const char* env = getenv("GASFM_GEMM_CAT_PRODUCERS"); pwg = (env && atoi(env) == 16) ? 4 : 2; }
        auto kernel = pwg == 4 ? (g_trace ? gemm_f16x2_cat_pair_kernel<true, 4> : gemm_f16x2_cat_pair_kernel<false, 4>)
                               : (g_trace ? gemm_f16x2_cat_pair_kernel<true, 2> : gemm_f16x2_cat_pair_kernel<false, 2>);
        const int threads = 128 * (pwg + 2);
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCSmemBytes);
        if (e != cudaSuccess) {
          set_error("linear_f16x2_cat: cannot reserve %zu bytes of shared memory (%s)", kCSmemBytes, cudaGetErrorString(e));
          return (int)e;
        }
        kernel<<<pgrid, threads, kCSmemBytes, (cudaStream_t)stream>>>(mh, ml, args);
        return check_launch("linear_f16x2_cat (pair)");
      }
    }
  }
#define LAUNCH_CAT(KB)                                                                                                          \
  do {                                                                                                                          \
    cudaError_t e = cudaFuncSetAttribute(gemm_f16x2_kernel<KB, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    if (e != cudaSuccess) {                                                                                                     \
      set_error("linear_f16x2_cat: cannot reserve %zu bytes of shared memory (%s)", smem, cudaGetErrorString(e));              \
      return (int)e;                                                                                                            \
    }                                                                                                                           \
    gemm_f16x2_kernel<KB, false, true><<<grid, kFThreads, smem, (cudaStream_t)stream>>>(mh, ml, args);                     \
  } while (0)
  if (seg_k == 256) LAUNCH_CAT(4); else LAUNCH_CAT(2);
#undef LAUNCH_CAT
  return check_launch("linear_f16x2_cat");
}

extern "C" int gasfm_linear_f16x2(const float* A, int64_t lda, const void* B_hi, const void* B_lo, const float* b_descale,
                                  const float* bias, float* C, int64_t ldc, int64_t M, int N, int K, int groups, int accumulate,
                                  float* a_amax, void* stream) {
  return linear_f16x2_impl(A, lda, B_hi, B_lo, b_descale, bias, C, ldc, M, N, K, groups, accumulate, a_amax, nullptr, nullptr, 0.f,
                           nullptr, nullptr, nullptr, 0, stream);
}

extern "C" int gasfm_linear_f16x2_ln(const float* A, int64_t lda, const float* ln_gamma, const float* ln_beta, float ln_eps,
                                     float* ln_mean, float* ln_rstd, const void* B_hi, const void* B_lo, const float* b_descale,
                                     const float* bias, float* C, int64_t ldc, int64_t M, int N, int K, int groups,
                                     float* a_amax, void* stream) {
  GASFM_REQUIRE(ln_gamma != nullptr, "linear_f16x2_ln: gamma is required");
  return linear_f16x2_impl(A, lda, B_hi, B_lo, b_descale, bias, C, ldc, M, N, K, groups, 0, a_amax, ln_gamma, ln_beta, ln_eps,
                           ln_mean, ln_rstd, nullptr, 0, stream);
}

extern "C" int gasfm_linear_f16x2_ln_y_supported(int64_t M, int N, int K, int64_t lda, int64_t ldc) {
  const char* env = getenv("GASFM_GEMM_PAIR");
  return (gasfm_linear_f16x2_supported(M, N, K, lda, ldc) && N == kPN && K == kPKB * kFBlockK && ldc % 8 == 0 && !(env && atoi(env) == 0)) ? 1 : 0;
}

extern "C" int gasfm_linear_f16x2_ln_y(const float* A, int64_t lda, const float* ln_gamma, const float* ln_beta, float ln_eps,
                                       float* ln_mean, float* ln_rstd, float* y, int64_t ldy, const void* B_hi, const void* B_lo,
                                       const float* b_descale, const float* bias, float* C, int64_t ldc, int64_t M, int N, int K,
                                       int groups, float* a_amax, void* stream) {
  GASFM_REQUIRE(ln_gamma != nullptr && y != nullptr, "linear_f16x2_ln_y: gamma and y are required");
  return linear_f16x2_impl(A, lda, B_hi, B_lo, b_descale, bias, C, ldc, M, N, K, groups, 0, a_amax, ln_gamma, ln_beta, ln_eps,
                           ln_mean, ln_rstd, y, ldy, stream);
}
