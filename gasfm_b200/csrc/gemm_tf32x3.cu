// C[M,N] = A[M,K] * B[N,K]^T (+ bias), fp32 in / fp32 out, on the 5th-generation tensor cores
// (tcgen05.mma kind::tf32, accumulators in TMEM, operands staged by TMA) with the 3xTF32 split
//     A*B ~= A_hi*B_hi + A_lo*B_hi + A_hi*B_lo,     x_hi = tf32(x), x_lo = x - x_hi
// which keeps fp32-level accuracy (~1e-6 relative), as the reference's fp32 GEMMs have
// (TF32 is off by default in PyTorch), while leaving the SIMT pipes.
//
// This is the dense per-observation projection of the GASFM layers: lin_l of the two GATv2 graphs
// and lin_proj of the observation update ([E,d] x [d,d'] with E in the millions, d <= 256), and the
// matching input gradients dX = dY * W.  Replaces cuBLAS SGEMM behind torch.nn.functional.linear
// at reference call sites code/models/layers.py:329,426 (GATv2Conv.lin_l) and :941 (lin_proj).
//
// Kernel structure (one persistent CTA per SM, 14 warps):
//   warp 0      TMA producer: B_hi / B_lo tiles [N x 32] per K-block (SWIZZLE_128B; the weights stay in L2)
//   warp 1      TMEM allocation + single-thread tcgen05.mma issue, tcgen05.commit to the barriers
//   warps 2-9   A producers: coalesced 128-bit global loads of the [128 x 32] fp32 tile (4 K-blocks prefetched
//               in registers), split into tf32(A) and A - tf32(A), written to shared memory directly in
//               the 128B-swizzled K-major layout the MMA descriptors expect (no TMA -> LDS -> STS round trip:
//               ncu showed the shared-memory data pipe, not the tensor pipe, was the contended resource)
//   warps 10-13 epilogue: tcgen05.ld the 128 x N accumulator, add bias, store rows to global
// Pipelines: smem ring (B full, A split_done) -> mma -> empty; TMEM ring tmem_full <-> tmem_empty (2 accumulators).
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "col_reduce.cuh"
#include "umma.cuh"
#include "../../include/gasfm_b200.h"

namespace gasfm {

constexpr int kBlockM = 128;
constexpr int kBlockK = 32;            // 32 fp32 = 128 bytes = one SWIZZLE_128B atom row
constexpr int kUmmaK = 8;              // tf32: 32 bytes per MMA along K
constexpr int kStages = 2;
constexpr int kGemmThreads = 448;          // TMA warp, MMA warp, 8 A-producer warps, 4 epilogue warps
constexpr int kPrefetch = 4;               // K-blocks of A kept in flight per producer thread
constexpr int kATileBytes = kBlockM * kBlockK * 4;   // 16 KB

constexpr int kMaxASegments = 4;
// A may be the column-wise concatenation [A_0 | A_1 | ...] of up to kMaxASegments matrices of seg_k columns each
// (input gradient of several projections of the same input: dX = sum_i dY_i W_i = [dY_0 | dY_1 | ..] [W_0; W_1; ..])
struct GemmArgs {
  const float* A[kMaxASegments]; int64_t lda[kMaxASegments]; int seg_k;
  const float* bias; float* C; int64_t ldc; int64_t M; int N; int K; int tmem_cols; int accumulate; int debug;
  float* a_amax;   // optional [segments]: max |A_i| (atomicMax; zeroed by the launcher) -- scales for wgrad_f16x2
};

// CTA pairs (cluster of 2): the weight tiles B_hi / B_lo are identical for every M tile, and re-streaming them
// from L2 for each tile (80 KB per K-block per SM) ran into the L2 bandwidth cap.  Each CTA of a pair loads half
// of the B rows and multicasts them into both CTAs' shared memory, halving the L2 traffic for B.  The pair walks
// its M tiles in lockstep (tile = 2 * pair_step + cta_rank); the MMAs themselves stay cta_group::1.
constexpr int kCluster = 2;
template <int kDummy>
__global__ void __cluster_dims__(kCluster, 1, 1) __launch_bounds__(kGemmThreads, 1)
gemm_tf32x3_kernel(const __grid_constant__ CUtensorMap map_bhi, const __grid_constant__ CUtensorMap map_blo,
                   const __grid_constant__ CUtensorMap map_c, GemmArgs p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [stage][A_hi 16K | A_lo 16K | B_hi N*128 | B_lo N*128], all 1024-aligned (N*128 is a multiple of 1024 for N%8==0)
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int b_tile_bytes = p.N * kBlockK * 4;
  const int stage_bytes = 2 * kATileBytes + 2 * b_tile_bytes;
  uint8_t* c_stage = smem + (size_t)kStages * stage_bytes;      // 4 epilogue warps x [32 rows x 128 B], 1024-aligned
  __shared__ uint64_t full_bar[kStages], split_bar[kStages], empty_bar[kStages], tmem_full_bar[2], tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_slot;
  __shared__ float bias_s[256];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_k_blocks = (p.K + kBlockK - 1) / kBlockK;
  const int64_t num_tiles = (p.M + kBlockM - 1) / kBlockM;
  const uint32_t cta_rank = cluster_ctarank();
  const int64_t num_clusters = gridDim.x / kCluster, cluster_id = blockIdx.x / kCluster;
  const int64_t num_pairs = (num_tiles + kCluster - 1) / kCluster;
  // pair steps this cluster executes; BOTH CTAs run every step (a CTA whose tile is past the end feeds zeros
  // and stores nothing) so that the shared pipeline of multicast loads never deadlocks
  const int64_t my_steps = cluster_id < num_pairs ? (num_pairs - cluster_id + num_clusters - 1) / num_clusters : 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&split_bar[s], 256); mbar_init(&empty_bar[s], kCluster); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full_bar[a], 1); mbar_init(&tmem_empty_bar[a], 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int j = threadIdx.x; j < 256; j += kGemmThreads) bias_s[j] = (p.bias && j < p.N) ? p.bias[j] : 0.f;
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

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      const int half_rows = p.N / kCluster;                       // B rows this CTA fetches and multicasts
      const int half_bytes = half_rows * kBlockK * 4;
      for (int64_t it = 0; it < my_steps; ++it) {
        for (int kb = 0; kb < num_k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);                 // both CTAs have retired the MMAs of this stage
          uint8_t* st = smem + (size_t)stage * stage_bytes;
          if (p.debug & 4) { mbar_arrive(&full_bar[stage]); if (++stage == kStages) { stage = 0; phase ^= 1; } continue; }
          mbar_expect_tx(&full_bar[stage], 2 * b_tile_bytes);      // halves from both CTAs land here
          tma_load_2d_mc(st + 2 * kATileBytes + cta_rank * half_bytes, &map_bhi, &full_bar[stage], kb * kBlockK,
                         (int)cta_rank * half_rows, (uint16_t)((1u << kCluster) - 1));
          tma_load_2d_mc(st + 2 * kATileBytes + b_tile_bytes + cta_rank * half_bytes, &map_blo, &full_bar[stage], kb * kBlockK,
                         (int)cta_rank * half_rows, (uint16_t)((1u << kCluster) - 1));
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      // instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=tf32, both K-major, N>>3, M>>4
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.N >> 3) << 17) | ((uint32_t)(kBlockM >> 4) << 24);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int64_t it = 0; it < my_steps; ++it) {
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * acc_cols);
        for (int kb = 0; kb < num_k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);       // B_hi / B_lo landed (TMA)
          mbar_wait(&split_bar[stage], phase);      // A_hi / A_lo written by the producer warps
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t a_hi = smem_u32(smem + (size_t)stage * stage_bytes);
          const uint32_t a_lo = a_hi + kATileBytes;
          const uint32_t b_hi = a_hi + 2 * kATileBytes;
          const uint32_t b_lo = b_hi + b_tile_bytes;
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) {
            if (p.debug & 2) break;
            const uint32_t koff = k * kUmmaK * 4;   // bytes along K inside the 128-byte swizzle row
            const uint32_t first = (kb == 0 && k == 0) ? 0u : 1u;
            umma_tf32(d_tmem, make_desc(a_hi + koff), make_desc(b_hi + koff), idesc, first);
            umma_tf32(d_tmem, make_desc(a_lo + koff), make_desc(b_hi + koff), idesc, 1u);
            umma_tf32(d_tmem, make_desc(a_hi + koff), make_desc(b_lo + koff), idesc, 1u);
          }
          umma_commit_mc(&empty_bar[stage], (uint16_t)((1u << kCluster) - 1));   // frees the stage in BOTH CTAs' eyes
          if (kb == num_k_blocks - 1) umma_commit(&tmem_full_bar[acc]);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp < 10) {
    // ===================== A producers (256 threads) =====================
    // thread -> 16-byte chunk q of rows rg, rg+32, rg+64, rg+96: a warp-wide load covers 4 full 128-byte rows.
    // kPrefetch K-blocks are kept in flight in registers (a ring with compile-time slots) so that HBM latency
    // is covered by ~kPrefetch x 16 KB of outstanding loads per SM.
    const int t = threadIdx.x - 64;
    const int q = t & 7, rg = t >> 3;
    int stage = 0; uint32_t phase = 0;
    const int64_t total = my_steps * num_k_blocks;           // K-blocks this CTA produces
    float4 buf[kPrefetch][4];
    float seg_max[kMaxASegments] = {0.f, 0.f, 0.f, 0.f};
    auto load_block = [&](int64_t g, float4 (&v)[4]) {
      const int64_t tile = (cluster_id + (g / num_k_blocks) * num_clusters) * kCluster + cta_rank;
      const int kcol = (int)(g % num_k_blocks) * kBlockK + q * 4;
      const int seg = kcol < p.K ? kcol / p.seg_k : 0, kloc = kcol - seg * p.seg_k;   // seg_k % 4 == 0: a float4 never straddles
      const float* a_seg = p.A[seg];
      const int64_t lda_seg = p.lda[seg];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int64_t row = tile * kBlockM + rg + 32 * i;
        v[i] = (g < total && row < p.M && kcol < p.K && !(p.debug & 8)) ? ld_stream4(a_seg + row * lda_seg + kloc) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
#pragma unroll
    for (int u = 0; u < kPrefetch - 1; ++u) load_block(u, buf[u]);
    for (int64_t g0 = 0; g0 < total; g0 += kPrefetch) {
#pragma unroll
      for (int u = 0; u < kPrefetch; ++u) {
        const int64_t g = g0 + u;
        load_block(g + kPrefetch - 1, buf[(u + kPrefetch - 1) % kPrefetch]);
        if (g < total) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* a_hi = smem + (size_t)stage * stage_bytes;
          uint8_t* a_lo = a_hi + kATileBytes;
          if (p.a_amax != nullptr) {
            const int kc = (int)(g % num_k_blocks) * kBlockK + q * 4;
            const int sg = kc < p.K ? kc / p.seg_k : 0;
            float m = 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i)
              m = fmaxf(m, fmaxf(fmaxf(fabsf(buf[u][i].x), fabsf(buf[u][i].y)), fmaxf(fabsf(buf[u][i].z), fabsf(buf[u][i].w))));
#pragma unroll
            for (int j = 0; j < kMaxASegments; ++j) seg_max[j] = (j == sg) ? fmaxf(seg_max[j], m) : seg_max[j];
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int row = rg + 32 * i;
            const int off = row * 128 + ((q ^ (row & 7)) << 4);   // SWIZZLE_128B: chunk ^= row % 8
            float4 v = buf[u][i], h, l;
            h.x = tf32_hi(v.x); h.y = tf32_hi(v.y); h.z = tf32_hi(v.z); h.w = tf32_hi(v.w);
            l.x = v.x - h.x; l.y = v.y - h.y; l.z = v.z - h.z; l.w = v.w - h.w;
            *reinterpret_cast<float4*>(a_hi + off) = h;
            *reinterpret_cast<float4*>(a_lo + off) = l;
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to tcgen05.mma
          mbar_arrive(&split_bar[stage]);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
    if (p.a_amax != nullptr) {
#pragma unroll
      for (int j = 0; j < kMaxASegments; ++j)
        if (j * p.seg_k < p.K) warp_amax_to_global(seg_max[j], p.a_amax + j);
    }
  } else {
    // ===================== epilogue (warps 10..13 -> TMEM lane quarters 2,3,0,1) =====================
    const int quarter = warp & 3;
    int acc = 0; uint32_t acc_phase = 0;
    for (int64_t it = 0; it < my_steps; ++it) {
      const int64_t tile = (cluster_id + it * num_clusters) * kCluster + cta_rank;
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int row0 = (int)(tile * kBlockM) + quarter * 32;       // first C row of this warp's 32-row slab
      uint8_t* stg = c_stage + (warp - 10) * 4096;
      const uint32_t taddr0 = tmem_base + (uint32_t)(acc * acc_cols) + ((uint32_t)(quarter * 32) << 16);
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
        // the previous chunk's bulk store must have finished READING the staging tile
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
        // lane = row: write its 8 float4 into the 128B-swizzled tile (chunk ^= row % 8): conflict-free, and the
        // layout a SWIZZLE_128B tensor map expects.  One TMA bulk store then writes 32 full 128-byte lines
        // (rows >= M and columns >= N are clipped by the tensor map), instead of 32 x 16-byte pieces per instruction.
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float4 v;
          v.x = __uint_as_float(r[j]) + bias_s[(c0 + j) & 255];
          v.y = __uint_as_float(r[j + 1]) + bias_s[(c0 + j + 1) & 255];
          v.z = __uint_as_float(r[j + 2]) + bias_s[(c0 + j + 2) & 255];
          v.w = __uint_as_float(r[j + 3]) + bias_s[(c0 + j + 3) & 255];
          *reinterpret_cast<float4*>(stg + lane * 128 + ((((j >> 2) ^ (lane & 7))) << 4)) = v;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0 && !(p.debug & 1)) {
          if (p.accumulate)
            asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&map_c),
                         "r"(smem_u32(stg)), "r"(c0), "r"(row0)
                         : "memory");
          else
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&map_c),
                         "r"(smem_u32(stg)), "r"(c0), "r"(row0)
                         : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(&tmem_empty_bar[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // all bulk stores complete before exit
  }
  __syncthreads();
  cluster_sync();                            // no CTA leaves while its peer may still multicast into its shared memory
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

// =============================================================================================
// Weight gradient:  dW[Nout,Kout] = dY[E,Nout]^T * X[E,Kout]   (reduction over the E observation rows)
//
// Both operands are activations stored row-major with the reduction index E as the slow dimension,
// i.e. they are MN-major UMMA operands: a TMA box of [32 columns x 16 rows] lands in shared memory as
// 16 rows of 128 bytes (TMA SWIZZLE_128B_ATOM_32B), which is exactly the canonical MN-major tf32 layout
// (UMMA SWIZZLE_128B_BASE32B) ((4,8,m),(4,k)) : ((1,4,LBO),(32,SBO)) with LBO = box size, SBO = 512 B.  Each CTA owns a
// contiguous range of E, accumulates the full [Nout x Kout] result in TMEM (2 x 256 columns) and
// writes one partial; a small kernel sums the partials (deterministic split-K).
// Both operands need the hi/lo split, done in place by the 8 worker warps, which also run the epilogue.
// =============================================================================================
constexpr int kWgRows = 16;                       // E rows per pipeline stage
constexpr int kWgBoxBytes = 32 * kWgRows * 4;     // 2 KB
constexpr int kWgStages = 3;
constexpr int kWgThreads = 320;

__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)(kWgBoxBytes >> 4) << 16;        // LBO: next 32-column block
  d |= (uint64_t)(512 >> 4) << 32;                // SBO: next swizzle atom (4 rows of 128 B) along K
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;                         // SWIZZLE_128B_BASE32B: the only MN-major layout for tf32
  return d;
}

struct WgradArgs {
  float* ws; float* ws_db; int64_t E; int Nout; int Kout; int m_tiles; int a_boxes; int b_boxes; int tmem_cols; int64_t rows_per_cta; int pass_stages;
};

template <int kDummy>
__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_tf32x3_kernel(const __grid_constant__ CUtensorMap map_dy, const __grid_constant__ CUtensorMap map_x, WgradArgs p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  // stage: [A_hi a_boxes*2K | A_lo | B_hi b_boxes*2K | B_lo]
  const int a_bytes = p.a_boxes * kWgBoxBytes, b_bytes = p.b_boxes * kWgBoxBytes;
  const int stage_bytes = 2 * a_bytes + 2 * b_bytes;
  __shared__ uint64_t full_bar[kWgStages], split_bar[kWgStages], empty_bar[kWgStages], done_bar, drained_bar;
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row_begin = (int64_t)blockIdx.x * p.rows_per_cta;
  const int64_t row_end = min(row_begin + p.rows_per_cta, p.E);
  const int num_stages_total = row_end > row_begin ? (int)((row_end - row_begin + kWgRows - 1) / kWgRows) : 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kWgStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&split_bar[s], 256); mbar_init(&empty_bar[s], 1); }
    mbar_init(&done_bar, 1);
    mbar_init(&drained_bar, 256);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_slot;
  // The tensor core accumulates in fp32 without rounding to nearest, so the error of a TMEM accumulator
  // grows with the length of the chain.  The E range of a CTA is therefore processed in passes of
  // pass_stages stages; after each pass the accumulator is drained and added into the CTA's partial.
  const int num_passes = num_stages_total > 0 ? (num_stages_total + p.pass_stages - 1) / p.pass_stages : 1;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int it = 0; it < num_stages_total; ++it) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* st = smem + (size_t)stage * stage_bytes;
        const int e0 = (int)(row_begin + (int64_t)it * kWgRows);
        mbar_expect_tx(&full_bar[stage], a_bytes + b_bytes);
        for (int j = 0; j < p.a_boxes; ++j) tma_load_2d(st + j * kWgBoxBytes, &map_dy, &full_bar[stage], j * 32, e0);
        for (int j = 0; j < p.b_boxes; ++j) tma_load_2d(st + 2 * a_bytes + j * kWgBoxBytes, &map_x, &full_bar[stage], j * 32, e0);
        if (++stage == kWgStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // D=f32, A=B=tf32, both MN-major (bits 15, 16), N = Kout, M = 128
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) |
                             ((uint32_t)(p.Kout >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      int stage = 0; uint32_t phase = 0;
      for (int it = 0; it < num_stages_total; ++it) {
        const int in_pass = it % p.pass_stages;
        if (in_pass == 0 && it > 0) {
          // previous pass fully issued: signal it, then wait until the workers have drained TMEM
          umma_commit(&done_bar);
          mbar_wait(&drained_bar, (uint32_t)((it / p.pass_stages - 1) & 1));
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        mbar_wait(&split_bar[stage], phase);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a_hi = smem_u32(smem + (size_t)stage * stage_bytes);
        const uint32_t a_lo = a_hi + a_bytes, b_hi = a_hi + 2 * a_bytes, b_lo = b_hi + b_bytes;
#pragma unroll
        for (int k = 0; k < kWgRows / kUmmaK; ++k) {
          const uint32_t koff = k * 1024;                      // 8 rows of 128 B
          const uint32_t acc = (in_pass == 0 && k == 0) ? 0u : 1u;
          for (int mt = 0; mt < p.m_tiles; ++mt) {
            const uint32_t d = tmem_base + (uint32_t)(mt * p.Kout);
            const uint32_t moff = mt * 4 * kWgBoxBytes;        // 128 columns of dY = 4 boxes
            umma_tf32(d, make_desc_mn(a_hi + moff + koff), make_desc_mn(b_hi + koff), idesc, acc);
            umma_tf32(d, make_desc_mn(a_lo + moff + koff), make_desc_mn(b_hi + koff), idesc, 1u);
            umma_tf32(d, make_desc_mn(a_hi + moff + koff), make_desc_mn(b_lo + koff), idesc, 1u);
          }
        }
        umma_commit(&empty_bar[stage]);
        if (++stage == kWgStages) { stage = 0; phase ^= 1; }
      }
      umma_commit(&done_bar);
    }
  } else {
    // ===================== 8 worker warps: split both operands, then epilogue =====================
    const int t = threadIdx.x - 64;                            // 0..255
    int stage = 0; uint32_t phase = 0;
    const int a_vec = a_bytes / 16, tot_vec = (a_bytes + b_bytes) / 16;
    // epilogue role: warps 2..5 -> M tile 0, warps 6..9 -> M tile 1; TMEM lane quarter = warp % 4
    const int mt = (warp - 2) >> 2;
    const int quarter = warp & 3;
    const int n_out = mt * 128 + quarter * 32 + lane;
    float* dst = p.ws + ((int64_t)blockIdx.x * p.Nout + n_out) * p.Kout;
    const uint32_t taddr0 = tmem_base + (uint32_t)(mt * p.Kout) + ((uint32_t)(quarter * 32) << 16);
    // column sums of dY (= the bias gradient) ride along: a thread always meets the same (box, row, slot)
    // positions of the dY tile, so it keeps one float4 of running sums per position (<= 4 positions)
    float4 colsum[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) colsum[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int pass = 0; pass < num_passes; ++pass) {
      const int it_end = min(num_stages_total, (pass + 1) * p.pass_stages);
      for (int it = pass * p.pass_stages; it < it_end; ++it) {
        mbar_wait(&full_bar[stage], phase);
        uint8_t* st = smem + (size_t)stage * stage_bytes;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int idx = t + 256 * i;
          if (idx < tot_vec) {
            const bool is_a = idx < a_vec;
            float4* hi = reinterpret_cast<float4*>(is_a ? st : st + 2 * a_bytes) + (is_a ? idx : idx - a_vec);
            float4* lo = reinterpret_cast<float4*>(reinterpret_cast<uint8_t*>(hi) + (is_a ? a_bytes : b_bytes));
            float4 v = *hi, h, l;
            h.x = tf32_hi(v.x); h.y = tf32_hi(v.y); h.z = tf32_hi(v.z); h.w = tf32_hi(v.w);
            l.x = v.x - h.x; l.y = v.y - h.y; l.z = v.z - h.z; l.w = v.w - h.w;
            *hi = h;
            *lo = l;
            if (i < 4 && is_a) { colsum[i].x += v.x; colsum[i].y += v.y; colsum[i].z += v.z; colsum[i].w += v.w; }
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive(&split_bar[stage]);
        if (++stage == kWgStages) { stage = 0; phase ^= 1; }
      }
      // drain this pass's accumulator into the CTA's partial
      if (num_stages_total > 0) {
        mbar_wait(&done_bar, (uint32_t)(pass & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      }
      if (mt < p.m_tiles) {
        for (int c0 = 0; c0 < p.Kout; c0 += 16) {
          uint32_t r[16];
          if (num_stages_total > 0) {
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                  "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                : "r"(taddr0 + (uint32_t)c0));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) r[j] = 0u;
          }
          if (n_out < p.Nout) {
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              float4 v = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
              if (pass > 0) {
                const float4 o = *reinterpret_cast<const float4*>(dst + c0 + j);
                v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
              }
              *reinterpret_cast<float4*>(dst + c0 + j) = v;
            }
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(&drained_bar);
    }
    // bias gradient: scatter the per-position sums to S[row 0..15][logical column], then add the 16 rows.
    // TMA SWIZZLE_128B_ATOM_32B stores logical 32-byte chunk c of row r at physical chunk c ^ (r & 3).
    if (p.ws_db != nullptr) {
      asm volatile("bar.sync 1, 256;" ::: "memory");           // all 8 worker warps are past their last stage
      float* S = reinterpret_cast<float*>(smem);               // 16 x 256 floats, reuses stage 0 (MMAs have retired)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int idx = t + 256 * i;
        if (idx < a_vec) {
          const int j = idx >> 7, r = (idx & 127) >> 3, sl = idx & 7;
          const int col = j * 32 + (((sl >> 1) ^ (r & 3)) << 3) + ((sl & 1) << 2);
          *reinterpret_cast<float4*>(S + r * 256 + col) = colsum[i];
        }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (t < p.Nout) {
        float a = 0.f;
#pragma unroll
        for (int r = 0; r < kWgRows; ++r) a += S[r * 256 + t];
        p.ws_db[(int64_t)blockIdx.x * p.Nout + t] = a;
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

__global__ void split_tf32_kernel(const float* __restrict__ w, float* __restrict__ hi, float* __restrict__ lo, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float x = w[i], h = tf32_hi(x);
  hi[i] = h;
  lo[i] = x - h;
}

static int make_map(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  return make_map_box(map, base, rows, cols, ld, box_rows, kBlockK);
}

}  // namespace gasfm

using namespace gasfm;

extern "C" int gasfm_split_tf32(const float* w, float* hi, float* lo, int64_t n, void* stream) {
  if (n <= 0) return 0;
  split_tf32_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(w, hi, lo, n);
  return check_launch("split_tf32");
}

extern "C" int gasfm_linear_tf32x3_supported(int64_t M, int N, int K, int64_t lda, int64_t ldc) {
  return (M > 0 && N >= 16 && N <= 256 && N % 16 == 0 && K >= 4 && K % 4 == 0 && lda % 4 == 0 && ldc % 4 == 0) ? 1 : 0;
}

static int launch_linear_tf32x3(const float* const* A, const int64_t* lda, int n_seg, int seg_k, const float* B_hi, const float* B_lo,
                                const float* bias, float* C, int64_t ldc, int64_t M, int N, int K, int accumulate, float* a_amax,
                                void* stream) {
  CUtensorMap mh, ml, mc;
  if (make_map(&mh, B_hi, N, K, K, N / kCluster) || make_map(&ml, B_lo, N, K, K, N / kCluster) ||
      make_map_box(&mc, C, M, N, ldc, 32, 32)) return 1;
  int tmem_cols = 32;
  while (tmem_cols < 2 * N) tmem_cols <<= 1;
  const size_t smem = (size_t)kStages * (2 * kATileBytes + 2 * (size_t)N * kBlockK * 4) + 4 * 4096 + 1024;
  size_t& smem_allowed = smem_opt_in_slot(0);   // per device; static smem (barriers, bias) also counts against the 227 KB per-CTA limit
  if (smem > smem_allowed) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tf32x3_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("linear_tf32x3: cannot reserve %zu bytes of shared memory (%s)", smem, cudaGetErrorString(e));
      return (int)e;
    }
    smem_allowed = smem;
  }
  const int64_t tiles = (M + kBlockM - 1) / kBlockM;
  const int64_t pairs = (tiles + kCluster - 1) / kCluster;
  const int grid = (int)(pairs < kNumSMs / kCluster ? pairs : kNumSMs / kCluster) * kCluster;
  static int debug = -1;
  if (debug < 0) { const char* env = getenv("GASFM_GEMM_DEBUG"); debug = env ? atoi(env) : 0; }   // phase-isolation knob for profiling
  GemmArgs args{};
  for (int i = 0; i < n_seg; ++i) { args.A[i] = A[i]; args.lda[i] = lda[i]; }
  args.seg_k = seg_k; args.bias = bias; args.C = C; args.ldc = ldc; args.M = M; args.N = N; args.K = K;
  args.tmem_cols = tmem_cols; args.accumulate = accumulate; args.debug = debug; args.a_amax = a_amax;
  if (a_amax != nullptr) cudaMemsetAsync(a_amax, 0, (size_t)n_seg * sizeof(float), (cudaStream_t)stream);
  gemm_tf32x3_kernel<0><<<grid, kGemmThreads, smem, (cudaStream_t)stream>>>(mh, ml, mc, args);
  return check_launch("linear_tf32x3");
}

extern "C" int gasfm_linear_tf32x3(const float* A, int64_t lda, const float* B_hi, const float* B_lo, const float* bias,
                                   float* C, int64_t ldc, int64_t M, int N, int K, int accumulate, void* stream) {
  GASFM_REQUIRE(gasfm_linear_tf32x3_supported(M, N, K, lda, ldc), "linear_tf32x3: unsupported shape M=%lld N=%d K=%d lda=%lld ldc=%lld",
                (long long)M, N, K, (long long)lda, (long long)ldc);
  GASFM_REQUIRE(((uintptr_t)A | (uintptr_t)B_hi | (uintptr_t)B_lo | (uintptr_t)C) % 16 == 0, "linear_tf32x3: pointers must be 16-byte aligned");
  return launch_linear_tf32x3(&A, &lda, 1, K, B_hi, B_lo, bias, C, ldc, M, N, K, accumulate, nullptr, stream);
}

extern "C" int gasfm_linear_tf32x3_cat(const float* const* A, const int64_t* lda, int n_seg, int seg_k, const float* B_hi,
                                       const float* B_lo, const float* bias, float* C, int64_t ldc, int64_t M, int N,
                                       int accumulate, float* a_amax, void* stream) {
  GASFM_REQUIRE(A && lda && n_seg >= 1 && n_seg <= kMaxASegments && seg_k >= 4 && seg_k % 4 == 0,
                "linear_tf32x3_cat: 1..%d segments of a multiple of 4 columns", kMaxASegments);
  const int K = n_seg * seg_k;
  uintptr_t bits = (uintptr_t)B_hi | (uintptr_t)B_lo | (uintptr_t)C;
  for (int i = 0; i < n_seg; ++i) {
    GASFM_REQUIRE(gasfm_linear_tf32x3_supported(M, N, K, lda[i], ldc), "linear_tf32x3_cat: unsupported shape M=%lld N=%d K=%d lda=%lld",
                  (long long)M, N, K, (long long)lda[i]);
    bits |= (uintptr_t)A[i];
  }
  GASFM_REQUIRE(bits % 16 == 0, "linear_tf32x3_cat: pointers must be 16-byte aligned");
  return launch_linear_tf32x3(A, lda, n_seg, seg_k, B_hi, B_lo, bias, C, ldc, M, N, K, accumulate, a_amax, stream);
}

extern "C" int gasfm_wgrad_tf32x3_supported(int64_t E, int Nout, int Kout, int64_t lddy, int64_t ldx) {
  return (E > 0 && Nout >= 4 && Nout <= 256 && Nout % 4 == 0 && Kout >= 16 && Kout <= 256 && Kout % 16 == 0 && lddy % 4 == 0 && ldx % 4 == 0) ? 1 : 0;
}

extern "C" size_t gasfm_wgrad_tf32x3_ws_bytes(int Nout, int Kout) {
  return ((size_t)kNumSMs * Nout * Kout + (size_t)kNumSMs * 256) * sizeof(float);
}

extern "C" int gasfm_wgrad_tf32x3(const float* dY, int64_t lddy, const float* X, int64_t ldx, int64_t E, int Nout, int Kout,
                                  float* dW, float* dbias, void* ws, void* stream) {
  GASFM_REQUIRE(gasfm_wgrad_tf32x3_supported(E, Nout, Kout, lddy, ldx), "wgrad_tf32x3: unsupported shape E=%lld Nout=%d Kout=%d",
                (long long)E, Nout, Kout);
  GASFM_REQUIRE(ws != nullptr && ((uintptr_t)dY | (uintptr_t)X | (uintptr_t)dW | (uintptr_t)ws) % 16 == 0, "wgrad_tf32x3: bad pointers");
  CUtensorMap mdy, mx;
  // MN-major tf32 operands only exist in the 128B-swizzle-with-32B-atom layout (UMMA SWIZZLE_128B_BASE32B)
  if (make_map_box(&mdy, dY, E, Nout, lddy, kWgRows, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B) ||
      make_map_box(&mx, X, E, Kout, ldx, kWgRows, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return 1;
  const int m_tiles = (Nout + 127) / 128;
  const int a_boxes = m_tiles * 4, b_boxes = (Kout + 31) / 32;
  int tmem_cols = 32;
  while (tmem_cols < m_tiles * Kout) tmem_cols <<= 1;
  const size_t smem = (size_t)kWgStages * 2 * (a_boxes + b_boxes) * kWgBoxBytes + 1024;
  size_t& smem_allowed = smem_opt_in_slot(1);
  if (smem > smem_allowed) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_tf32x3_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("wgrad_tf32x3: cannot reserve %zu bytes of shared memory (%s)", smem, cudaGetErrorString(e)); return (int)e; }
    smem_allowed = smem;
  }
  const int64_t units = (E + kWgRows - 1) / kWgRows;
  const int grid = (int)(units < kNumSMs ? units : kNumSMs);
  const int64_t rows_per_cta = ((units + grid - 1) / grid) * kWgRows;
  static int pass_stages = 0;
  if (pass_stages == 0) {
    const char* env = getenv("GASFM_WGRAD_PASS_STAGES");   // tuning knob; default chosen from the A/B in profiles/
    pass_stages = env ? atoi(env) : 64;
    if (pass_stages < 1) pass_stages = 64;
  }
  float* ws_db = dbias ? (float*)ws + (size_t)kNumSMs * Nout * Kout : nullptr;
  WgradArgs a{(float*)ws, ws_db, E, Nout, Kout, m_tiles, a_boxes, b_boxes, tmem_cols, rows_per_cta, pass_stages};
  cudaStream_t st = (cudaStream_t)stream;
  wgrad_tf32x3_kernel<0><<<grid, kWgThreads, smem, st>>>(mdy, mx, a);
  int rc = check_launch("wgrad_tf32x3");
  if (rc) return rc;
  // dW = sum of the per-CTA partials, db likewise, in one launch
  const int64_t width = (int64_t)Nout * Kout;
  const ColReduceJob jw{(const float*)ws, width, width, dW, 0, 0}, jb{ws_db, Nout, Nout, dbias, 0, 0};
  launch_col_reduce(jw, dbias ? &jb : nullptr, grid, 1.f, st);
  return check_launch("wgrad_tf32x3(reduce)");
}
