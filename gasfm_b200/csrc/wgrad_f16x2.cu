// Weight gradient dW[Nout,Kout] = dY[E,Nout]^T X[E,Kout] (+ db = column sums of dY) on the fp16 tensor-core path with the
// scaled 2 x FP16 split of gemm_f16x2.cu:  dY^T X ~= dYh^T Xh + dYl^T Xh + dYh^T Xl, fp32 accumulation in TMEM.
//
// The reduction runs over the observation rows, so a per-row scale cannot be factored out of the sum; each OPERAND gets one
// power-of-two scale instead, from the largest magnitude of the whole matrix.  Those two maxima cost nothing: the projection
// GEMMs that read the same matrices just before (gemm_f16x2: x in the forward pass, gemm_tf32x3_cat: the dY_i in the backward
// pass) leave them in device memory (`a_amax`).  A split value v = s*x carries an absolute error of max(2^-22 |v|, 2^-25), i.e.
// full fp32-like relative accuracy over 18 binades below the matrix maximum and an absolute error of 2^-39 of the maximum below
// that -- invisible in a sum that the large entries dominate (tests: error / max|dW| against fp64).
//
// kind::f16 halves the tensor time of the 3xTF32 kernel (wgrad_tf32x3_kernel) and fp16 operands halve the shared-memory
// operand traffic that kernel is bound by; the operands are written by the producer warps directly (no TMA landing zone that
// has to be read back and split in place).
//
// Shared-memory operand layout: MN-major, SWIZZLE_128B, 16-bit elements = atoms of [8 rows (E, the MMA K index) x 128 bytes
// (64 consecutive columns, the MMA M or N index)], 16-byte chunk index XOR row; atom (j, kg) of a stage at
// ((j * 4 + kg) * 1024) bytes: SBO = 1024 (next 8 rows), LBO = 4096 (next 64 columns).
// Warps: 1 = TMEM allocation + MMA issue; 4..11 = producers (coalesced 128-bit loads, two 32-row stages in flight in registers),
// also drain the accumulator after every pass (fp32 accumulation in the tensor core truncates: bounded chain length).
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include "common.cuh"
#include "col_reduce.cuh"
#include "umma.cuh"
#include "../../include/gasfm_b200.h"

namespace gasfm {

constexpr int kHRows = 32;                  // E rows per pipeline stage = two K = 16 MMA steps (16-row stages: 0.43 vs 0.32 ms)
constexpr int kHStages = 3;
constexpr int kHRegBufs = 2;                // stages a producer thread keeps in flight in registers
constexpr int kHThreads = 384;               // WG0: warp 1 = MMA issue (40 regs); WG1-2: producers (232 regs, setmaxnreg)
constexpr int kHAtomBytes = 1024;           // 8 rows x 128 B
constexpr int kHKGroups = kHRows / 8;       // 4 atoms along E per stage
constexpr int kHLbo = kHKGroups * kHAtomBytes;

constexpr int kHMaxGroups = 3;               // weight gradients that share X, computed by one launch
struct WgradF16Args {
  const float* dY[kHMaxGroups]; int64_t lddy[kHMaxGroups]; const float* X; int64_t ldx;
  const float* amax_dy /* [n_groups] */; const float* amax_x;
  float* ws; float* ws_db; int64_t E; int Nout; int Kout; int m_tiles; int tmem_cols; int64_t rows_per_cta; int pass_stages;
  int prefetch;                               // L2 prefetch distance in stages beyond the register stages (0 = off)
  int n_groups;                              // CTA b works on group b % n_groups, row range b / n_groups: the CTAs that read the
};                                           // same rows of X are launched together, so X comes from HBM once and from L2 after

// MN-major SWIZZLE_128B descriptor (cute::UMMA canonical ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units)
__device__ __forceinline__ uint64_t make_desc_mn16(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)(kHLbo >> 4) << 16;          // LBO: next 64-column atom
  d |= (uint64_t)(kHAtomBytes >> 4) << 32;    // SBO: next 8 rows
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;                     // SWIZZLE_128B
  return d;
}

// FULL: Nout = Kout = 256 and 32-bit row offsets -- the shape of every GASFM block at d = 256.  The producer loop is then free
// of per-row predicates, 64-bit multiplies and shared-memory offset arithmetic (ncu on the generic loop: ~1,100 instructions
// per thread and 32-row stage for 16 float4, 70 % of them address / predicate overhead, schedulers 67 % busy issuing).
template <bool FULL>
__global__ void __launch_bounds__(kHThreads, 1) wgrad_f16x2_kernel(WgradF16Args p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  // stage: [A_hi | A_lo | B_hi | B_lo], A = dY (Nout/64 column atoms x 4 row atoms), B = X (Kout/64 x 4)
  const int a_bytes = (p.Nout / 64) * kHLbo, b_bytes = (p.Kout / 64) * kHLbo;
  const int stage_bytes = 2 * a_bytes + 2 * b_bytes;
  __shared__ uint64_t split_bar[kHStages], empty_bar[kHStages], done_bar, drained_bar;
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = (int)(blockIdx.x % p.n_groups), range = (int)(blockIdx.x / p.n_groups);
  const float* __restrict__ dYg = p.dY[grp];
  const int64_t lddyg = p.lddy[grp];
  const int64_t part = (int64_t)range * p.n_groups + grp;      // this CTA's row in the partial-result workspace
  const int64_t row_begin = (int64_t)range * p.rows_per_cta;
  const int64_t row_end = min(row_begin + p.rows_per_cta, p.E);
  const int num_stages_total = row_end > row_begin ? (int)((row_end - row_begin + kHRows - 1) / kHRows) : 0;
  const int num_passes = num_stages_total > 0 ? (num_stages_total + p.pass_stages - 1) / p.pass_stages : 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kHStages; ++s) { mbar_init(&split_bar[s], 256); mbar_init(&empty_bar[s], 1); }
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

  float s_dy, s_x, d_dy, d_x;
  row_scale_from_amax(__ldg(p.amax_dy + grp), s_dy, d_dy);
  row_scale_from_amax(__ldg(p.amax_x), s_x, d_x);

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;" ::: "memory");
    if (warp == 1 && lane == 0) {
      // D = f32, A = B = f16 (format 0), both MN-major (bits 15, 16), N = Kout, M = 128
      const uint32_t idesc = (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(p.Kout >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      int stage = 0; uint32_t phase = 0;
      for (int it = 0; it < num_stages_total; ++it) {
        const int in_pass = it % p.pass_stages;
        if (in_pass == 0 && it > 0) {
          umma_commit(&done_bar);                              // previous pass fully issued
          mbar_wait(&drained_bar, (uint32_t)((it / p.pass_stages - 1) & 1));
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        mbar_wait(&split_bar[stage], phase);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a_hi = smem_u32(smem + (size_t)stage * stage_bytes);
        const uint32_t a_lo = a_hi + a_bytes, b_hi = a_hi + 2 * a_bytes, b_lo = b_hi + b_bytes;
#pragma unroll
        for (int ks = 0; ks < kHRows / 16; ++ks) {
          const uint32_t koff = ks * 2 * kHAtomBytes;          // 16 rows = two row atoms
          const uint32_t acc = (in_pass == 0 && ks == 0) ? 0u : 1u;
          for (int mt = 0; mt < p.m_tiles; ++mt) {
            const uint32_t d = tmem_base + (uint32_t)(mt * p.Kout);
            const uint32_t moff = mt * 2 * kHLbo;              // 128 columns of dY = two column atoms
            umma_f16(d, make_desc_mn16(a_hi + moff + koff), make_desc_mn16(b_hi + koff), idesc, acc);
            umma_f16(d, make_desc_mn16(a_lo + moff + koff), make_desc_mn16(b_hi + koff), idesc, 1u);
            umma_f16(d, make_desc_mn16(a_hi + moff + koff), make_desc_mn16(b_lo + koff), idesc, 1u);
          }
        }
        umma_commit(&empty_bar[stage]);
        if (++stage == kHStages) { stage = 0; phase ^= 1; }
      }
      umma_commit(&done_bar);
    }
  } else {
    // ===================== 8 producer warps (also the drain) =====================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;" ::: "memory");
    const int t = threadIdx.x - 128;                           // 0..255
    const int col4 = t & 63, r0 = t >> 6;                      // float4 column of the row, rows r0 + 4 i of a stage
    const bool a_on = col4 * 4 < p.Nout, b_on = col4 * 4 < p.Kout;
    const int atom_off = (col4 >> 4) * kHLbo, chunk = (col4 & 15) >> 1, half_off = (col4 & 1) * 8;
    constexpr int RPT = kHRows / 4;                            // rows per thread and stage
    float4 bufA[kHRegBufs][RPT], bufB[kHRegBufs][RPT];
    // ---- lean path (FULL): everything that does not change from stage to stage is computed here, once ----
    const uint32_t toff_a = (uint32_t)(r0 * lddyg + 4 * col4), toff_b = (uint32_t)(r0 * p.ldx + 4 * col4);   // elements
    const uint32_t step_a = (uint32_t)(4 * lddyg), step_b = (uint32_t)(4 * p.ldx);                          // between this thread's rows
    uint32_t soff[RPT];                                        // byte offset of this thread's piece of row r0 + 4 i inside a plane
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
      const int row = r0 + 4 * i;
      soff[i] = (uint32_t)(atom_off + (row >> 3) * kHAtomBytes + (row & 7) * 128 + ((chunk ^ (row & 7)) << 4) + half_off);
    }
    const uint32_t smem_base = smem_u32(smem);
    auto load_stage = [&](int it, float4 (&a)[RPT], float4 (&b)[RPT]) {
      const int64_t e0 = row_begin + (int64_t)it * kHRows;     // first row of the stage (uniform)
      if (FULL && e0 + kHRows <= row_end) {                    // whole stage in range: no per-row predicates, 32-bit offsets
        const float* ba = dYg + e0 * lddyg;
        const float* bb = p.X + e0 * p.ldx;
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
          a[i] = ld_stream4(ba + (toff_a + (uint32_t)i * step_a));
          b[i] = ld_stream4(bb + (toff_b + (uint32_t)i * step_b));
        }
        // ncu (profiles/r02_wgrad_multi_cfg3_ncu_full.csv): 25 % of all samples are the first use of these loads -- two register
        // stages are one stage period of lead, less than the HBM latency under load.  Optional experiment (off by default,
        // GASFM_WGRAD_PREFETCH=<stages>): one lane per 128-byte line pulls the lines of a later stage into L2.  Measured SLOWER
        // (three dW of a block at cfg2: 0.672 ms without, 0.703 at 2 stages, 0.720 at 4): the prefetches are extra requests on
        // the same saturated path and do not shorten what the loads wait for.
        if (p.prefetch > 0 && (col4 & 7) == 0 && e0 + (int64_t)(p.prefetch + 1) * kHRows <= row_end) {
          const float* pa = ba + (int64_t)p.prefetch * kHRows * lddyg;
          const float* pb = bb + (int64_t)p.prefetch * kHRows * p.ldx;
#pragma unroll
          for (int i = 0; i < RPT; ++i) {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(pa + (toff_a + (uint32_t)i * step_a)));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(pb + (toff_b + (uint32_t)i * step_b)));
          }
        }
        return;
      }
#pragma unroll
      for (int i = 0; i < RPT; ++i) {
        const int64_t e = e0 + r0 + 4 * i;
        const bool ok = it < num_stages_total && e < row_end;
        a[i] = (ok && a_on) ? ld_stream4(dYg + e * lddyg + 4 * col4) : make_float4(0.f, 0.f, 0.f, 0.f);
        b[i] = (ok && b_on) ? ld_stream4(p.X + e * p.ldx + 4 * col4) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    // x -> fp16 hi + lo of s * x, packed: hv = hi pair words, lv = lo pair words
    auto split4 = [](const float4& v, float s, uint2& hv, uint2& lv) {
      const float x0 = v.x * s, x1 = v.y * s, x2 = v.z * s, x3 = v.w * s;
      const __half2 h01 = __floats2half2_rn(x0, x1), h23 = __floats2half2_rn(x2, x3);
      const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
      const __half2 l01 = __floats2half2_rn(x0 - f01.x, x1 - f01.y), l23 = __floats2half2_rn(x2 - f23.x, x3 - f23.y);
      hv.x = *reinterpret_cast<const uint32_t*>(&h01); hv.y = *reinterpret_cast<const uint32_t*>(&h23);
      lv.x = *reinterpret_cast<const uint32_t*>(&l01); lv.y = *reinterpret_cast<const uint32_t*>(&l23);
    };
    auto split_store = [&](uint8_t* hi_base, uint8_t* lo_base, const float4& v, float s, int row) {
      const int off = atom_off + (row >> 3) * kHAtomBytes + (row & 7) * 128 + ((chunk ^ (row & 7)) << 4) + half_off;
      uint2 hv, lv;
      split4(v, s, hv, lv);
      *reinterpret_cast<uint2*>(hi_base + off) = hv;
      *reinterpret_cast<uint2*>(lo_base + off) = lv;
    };
    float4 colsum = make_float4(0.f, 0.f, 0.f, 0.f);           // column sums of dY over this thread's rows (bias gradient)
    int stage = 0; uint32_t phase = 0;
    auto produce = [&](int it, float4 (&a)[RPT], float4 (&b)[RPT]) {
      mbar_wait(&empty_bar[stage], phase ^ 1);
      if constexpr (FULL) {
        // planes of a stage: A_hi | A_lo | B_hi | B_lo, 16 KB each (Nout = Kout = 256): constant displacements
        constexpr int kPlane = (256 / 64) * kHLbo;
        const uint32_t sb = smem_base + (uint32_t)stage * (4 * kPlane);
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
          const uint32_t addr = sb + soff[i];
          uint2 hv, lv;
          split4(a[i], s_dy, hv, lv);
          asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(hv.x), "r"(hv.y) : "memory");
          asm volatile("st.shared.v2.b32 [%0+%3], {%1, %2};" ::"r"(addr), "r"(lv.x), "r"(lv.y), "n"(kPlane) : "memory");
          colsum.x += a[i].x; colsum.y += a[i].y; colsum.z += a[i].z; colsum.w += a[i].w;
          split4(b[i], s_x, hv, lv);
          asm volatile("st.shared.v2.b32 [%0+%3], {%1, %2};" ::"r"(addr), "r"(hv.x), "r"(hv.y), "n"(2 * kPlane) : "memory");
          asm volatile("st.shared.v2.b32 [%0+%3], {%1, %2};" ::"r"(addr), "r"(lv.x), "r"(lv.y), "n"(3 * kPlane) : "memory");
        }
      } else {
        uint8_t* st = smem + (size_t)stage * stage_bytes;
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
          const int row = r0 + 4 * i;
          if (a_on) {
            split_store(st, st + a_bytes, a[i], s_dy, row);
            colsum.x += a[i].x; colsum.y += a[i].y; colsum.z += a[i].z; colsum.w += a[i].w;
          }
          if (b_on) split_store(st + 2 * a_bytes, st + 2 * a_bytes + b_bytes, b[i], s_x, row);
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to tcgen05.mma
      mbar_arrive(&split_bar[stage]);                          // (one elected arrival per warp measured slower: 0.36 vs 0.32 ms)
      if (++stage == kHStages) { stage = 0; phase ^= 1; }
      load_stage(it + kHRegBufs, a, b);                        // refill the freed registers kHRegBufs stages ahead
    };
    // drain role: warps 4..7 -> M tile 0, warps 8..11 -> M tile 1; TMEM lane quarter = warp % 4
    const int mt = (warp - 4) >> 2;
    const int quarter = warp & 3;
    const uint32_t taddr0 = tmem_base + (uint32_t)(mt * p.Kout) + ((uint32_t)(quarter * 32) << 16);
    const float descale = d_dy * d_x;

#pragma unroll
    for (int u = 0; u < kHRegBufs; ++u) load_stage(u, bufA[u], bufB[u]);
    for (int pass = 0; pass < num_passes; ++pass) {
      const int it_end = min(num_stages_total, (pass + 1) * p.pass_stages);
      // pass_stages is a multiple of kHRegBufs, so a pass always starts on buffer 0 (only the LAST pass can be ragged)
      for (int it = pass * p.pass_stages; it < it_end; it += kHRegBufs) {
#pragma unroll
        for (int u = 0; u < kHRegBufs; ++u)
          if (it + u < it_end) produce(it + u, bufA[u], bufB[u]);
      }
      if (num_stages_total > 0) {
        mbar_wait(&done_bar, (uint32_t)(pass & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      }
      if (mt < p.m_tiles) {
        // 32-column chunks through a swizzled staging tile (the operand stages are idle: every MMA of the pass has completed):
        // TMEM -> registers (lane = row of dW) -> smem -> registers (4 lanes = one 128-byte row segment) -> coalesced
        // read-add-write of the CTA's partial.  Writing lane = row pieces of 16 bytes straight from the TMEM registers cost
        // ~22 us per pass; the previous partial is requested before the TMEM load so that its latency is covered.
        uint8_t* stg = smem + (warp - 4) * 4096;
        const int pc = lane & 3;
        for (int c0 = 0; c0 < p.Kout; c0 += 32) {
          const int col = c0 + 8 * pc;
          float4 o0[4], o1[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int n = mt * 128 + quarter * 32 + 8 * i + (lane >> 2);
            const float* src = p.ws + (part * p.Nout + n) * p.Kout + col;
            const bool on = pass > 0 && n < p.Nout && col < p.Kout;
            o0[i] = on ? *reinterpret_cast<const float4*>(src) : make_float4(0.f, 0.f, 0.f, 0.f);
            o1[i] = on ? *reinterpret_cast<const float4*>(src + 4) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
          uint32_t r[32];
          if (num_stages_total > 0) {
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                  "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                  "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                  "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                : "r"(taddr0 + (uint32_t)c0));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) r[j] = 0u;
          }
          __syncwarp();                                // the previous chunk's read-back is complete
#pragma unroll
          for (int j = 0; j < 32; j += 4) {            // lane = row; 16-byte chunk ^= row % 8: conflict-free
            const float4 v = make_float4(__uint_as_float(r[j]) * descale, __uint_as_float(r[j + 1]) * descale,
                                         __uint_as_float(r[j + 2]) * descale, __uint_as_float(r[j + 3]) * descale);
            *reinterpret_cast<float4*>(stg + lane * 128 + ((((j >> 2) ^ (lane & 7))) << 4)) = v;
          }
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int rr = 8 * i + (lane >> 2);
            const int n = mt * 128 + quarter * 32 + rr;
            if (n < p.Nout && col < p.Kout) {
              float4 v = *reinterpret_cast<const float4*>(stg + rr * 128 + (((2 * pc) ^ (rr & 7)) << 4));
              float4 w = *reinterpret_cast<const float4*>(stg + rr * 128 + (((2 * pc + 1) ^ (rr & 7)) << 4));
              v.x += o0[i].x; v.y += o0[i].y; v.z += o0[i].z; v.w += o0[i].w;
              w.x += o1[i].x; w.y += o1[i].y; w.z += o1[i].z; w.w += o1[i].w;
              float* out = p.ws + (part * p.Nout + n) * p.Kout + col;
              asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(out), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w),
                           "f"(w.x), "f"(w.y), "f"(w.z), "f"(w.w)
                           : "memory");
            }
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(&drained_bar);
      // the staging tiles live in the operand stages: nobody may produce the next pass's first stage while another warp
      // is still draining through its tile
      asm volatile("bar.sync 1, 256;" ::: "memory");
    }
    // bias gradient: the four row groups r0 = 0..3 of a column meet in shared memory (operand stages are idle now)
    if (p.ws_db != nullptr) {
      asm volatile("bar.sync 1, 256;" ::: "memory");           // all producers are past their last MMA-visible write
      float* S = reinterpret_cast<float*>(smem);               // [4][256]
      if (a_on) *reinterpret_cast<float4*>(S + r0 * 256 + 4 * col4) = colsum;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (t < p.Nout) p.ws_db[part * p.Nout + t] = (S[t] + S[256 + t]) + (S[512 + t] + S[768 + t]);
    }
  }
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

}  // namespace gasfm

using namespace gasfm;

extern "C" int gasfm_wgrad_f16x2_supported(int64_t E, int Nout, int Kout, int64_t lddy, int64_t ldx) {
  return (E > 0 && (Nout == 128 || Nout == 256) && Kout >= 64 && Kout <= 256 && Kout % 64 == 0 && lddy % 4 == 0 && ldx % 4 == 0) ? 1 : 0;
}

extern "C" size_t gasfm_wgrad_f16x2_ws_bytes(int Nout, int Kout) {
  return ((size_t)kNumSMs * Nout * Kout + (size_t)kNumSMs * 256) * sizeof(float);
}

static int wgrad_f16x2_launch(const float* const* dY, const int64_t* lddy, int n_groups, const float* X, int64_t ldx,
                              const float* amax_dy, const float* amax_x, int64_t E, int Nout, int Kout, float* dW, float* dbias,
                              void* ws, void* stream) {
  GASFM_REQUIRE(n_groups >= 1 && n_groups <= kHMaxGroups, "wgrad_f16x2: 1..%d groups", kHMaxGroups);
  for (int g = 0; g < n_groups; ++g) {
    GASFM_REQUIRE(gasfm_wgrad_f16x2_supported(E, Nout, Kout, lddy[g], ldx), "wgrad_f16x2: unsupported shape E=%lld Nout=%d Kout=%d",
                  (long long)E, Nout, Kout);
    GASFM_REQUIRE(dY[g] != nullptr && (uintptr_t)dY[g] % 16 == 0, "wgrad_f16x2: bad dY pointer");
  }
  GASFM_REQUIRE(ws != nullptr && amax_dy != nullptr && amax_x != nullptr && ((uintptr_t)X | (uintptr_t)dW | (uintptr_t)ws) % 16 == 0,
                "wgrad_f16x2: bad pointers");
  const int m_tiles = Nout / 128;
  int tmem_cols = 32;
  while (tmem_cols < m_tiles * Kout) tmem_cols <<= 1;
  const size_t smem = (size_t)kHStages * 2 * ((size_t)(Nout / 64) + (Kout / 64)) * kHLbo + 1024;
  size_t& smem_allowed = smem_opt_in_slot(2);
  if (smem > smem_allowed) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_f16x2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(wgrad_f16x2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("wgrad_f16x2: cannot reserve %zu bytes of shared memory (%s)", smem, cudaGetErrorString(e));
      return (int)e;
    }
    smem_allowed = smem;
  }
  // split-K over the SMs: ``ranges`` row ranges x n_groups CTAs; every CTA gets a multiple of the stage size
  int64_t stages = (E + kHRows - 1) / kHRows;
  const int max_ranges = kNumSMs / n_groups;
  int ranges = (int)(stages < max_ranges ? stages : max_ranges);
  const int64_t rows_per_cta = ((stages + ranges - 1) / ranges) * kHRows;
  ranges = (int)((E + rows_per_cta - 1) / rows_per_cta);
  static int pass_stages = 0;
  if (pass_stages == 0) {
    const char* env = getenv("GASFM_WGRAD_F16_PASS_STAGES");     // 32-row stages per accumulation pass (multiple of kHRegBufs)
    pass_stages = env ? atoi(env) : 32;
    if (pass_stages < kHRegBufs) pass_stages = 32;
    pass_stages &= ~(kHRegBufs - 1);
  }
  float* ws_db = dbias ? (float*)ws + (size_t)kNumSMs * Nout * Kout : nullptr;
  WgradF16Args a{};
  for (int g = 0; g < n_groups; ++g) { a.dY[g] = dY[g]; a.lddy[g] = lddy[g]; }
  a.X = X; a.ldx = ldx; a.amax_dy = amax_dy; a.amax_x = amax_x; a.ws = (float*)ws; a.ws_db = ws_db; a.E = E; a.Nout = Nout; a.Kout = Kout;
  a.m_tiles = m_tiles; a.tmem_cols = tmem_cols; a.rows_per_cta = rows_per_cta; a.pass_stages = pass_stages; a.n_groups = n_groups;
  static int prefetch = -1;                   // GASFM_WGRAD_PREFETCH=<stages> (0 = off; A/B switch)
  if (prefetch < 0) { const char* env = getenv("GASFM_WGRAD_PREFETCH"); prefetch = env ? atoi(env) : 0; }
  a.prefetch = prefetch;
  cudaStream_t st = (cudaStream_t)stream;
  bool full = Nout == 256 && Kout == 256 && ldx < (1 << 24);
  for (int g = 0; g < n_groups; ++g) full = full && lddy[g] < (1 << 24);
  static int lean = -1;
  if (lean < 0) { const char* env = getenv("GASFM_WGRAD_LEAN"); lean = env ? atoi(env) : 1; }   // A/B switch
  if (full && lean) wgrad_f16x2_kernel<true><<<ranges * n_groups, kHThreads, smem, st>>>(a);
  else wgrad_f16x2_kernel<false><<<ranges * n_groups, kHThreads, smem, st>>>(a);
  int rc = check_launch("wgrad_f16x2");
  if (rc) return rc;
  // partials are [range][group][Nout x Kout]: ONE column reduction yields the stacked [n_groups x Nout, Kout] result
  const int64_t width = (int64_t)n_groups * Nout * Kout, bwidth = (int64_t)n_groups * Nout;
  const ColReduceJob jw{(const float*)ws, width, width, dW, 0, 0}, jb{ws_db, bwidth, bwidth, dbias, 0, 0};
  launch_col_reduce(jw, dbias ? &jb : nullptr, ranges, 1.f, st);
  return check_launch("wgrad_f16x2(reduce)");
}

extern "C" int gasfm_wgrad_f16x2(const float* dY, int64_t lddy, const float* X, int64_t ldx, const float* amax_dy, const float* amax_x,
                                 int64_t E, int Nout, int Kout, float* dW, float* dbias, void* ws, void* stream) {
  return wgrad_f16x2_launch(&dY, &lddy, 1, X, ldx, amax_dy, amax_x, E, Nout, Kout, dW, dbias, ws, stream);
}

extern "C" int gasfm_wgrad_f16x2_multi(const float* const* dY, const int64_t* lddy, int n_groups, const float* X, int64_t ldx,
                                       const float* amax_dy, const float* amax_x, int64_t E, int Nout, int Kout, float* dW,
                                       float* dbias, void* ws, void* stream) {
  GASFM_REQUIRE(dY != nullptr && lddy != nullptr, "wgrad_f16x2_multi: NULL argument");
  return wgrad_f16x2_launch(dY, lddy, n_groups, X, ldx, amax_dy, amax_x, E, Nout, Kout, dW, dbias, ws, stream);
}
