// Fused GATv2 edge attention over CSR / CSC segments of the observation graph (sm_100a).
//
// Replaces the message-passing core of torch_geometric.nn.GATv2Conv as the reference calls it
// (code/models/layers.py:329-335, 426-432, 550-556, 566-572): per edge e with target t(e)
//     z = XL[e] + XR[t];  s[e,h] = sum_c att[h,c] * leaky_relu(z)[h,c]
//     alpha = segment-softmax(s);  out[t] = sum_e alpha * XL[e] (+ bias)
// One pass over XL: scores, online max / denominator and the weighted sum are all kept in
// registers; no [E,H,C] intermediate ever reaches HBM.
//
// Data layout.  A row of XL is H*C fp32 = NVEC float4.  LPR = min(32, NVEC) lanes cooperate on
// one row, each owning NV = NVEC/LPR float4 (lane l, vector v covers channels 4*(l + LPR*v)..+3,
// so every warp-wide load instruction touches LPR*16 contiguous bytes).  A warp therefore works
// on RPW = 32/LPR rows at once.  Heads are contiguous channel ranges, so the per-head score is a
// butterfly over the LPH lanes that share the head.
//
// Schedules.
//   short  (chunk == 0): one lane group per segment -- tracks (CSC, ~10-20 edges, rows gathered
//          through perm) -- no inter-group communication at all.
//   chunked(chunk  > 0): one warp per chunk of <= chunk edges of one segment -- views (CSR,
//          thousands of contiguous edges) and the single-segment global graphs.  The RPW lane
//          groups stride over the chunk, merge their (max, sum, acc) triples with shuffles and
//          either finalise (segment == one chunk) or park the triple in the workspace for
//          gat_merge_kernel, which is the flash-attention style log-sum-exp combine.
#include "common.cuh"
#include "../../include/gasfm_b200.h"

namespace gasfm {

template <int H, int C>
struct Lay {
  static constexpr int HC = H * C;
  static_assert(HC % 4 == 0, "row must be a whole number of float4");
  static constexpr int NVEC = HC / 4;
  static constexpr int LPR = NVEC < 32 ? NVEC : 32;
  static_assert((LPR & (LPR - 1)) == 0 && NVEC % LPR == 0, "unsupported row width");
  static constexpr int NV = NVEC / LPR;
  static constexpr int RPW = 32 / LPR;
  static constexpr int VPHEAD = C >= 4 ? C / 4 : 1;            // float4 per head
  static constexpr int HPV = C >= 4 ? 1 : 4 / C;               // heads per float4 (C < 4)
  static constexpr int LPH = VPHEAD < LPR ? VPHEAD : LPR;      // lanes sharing a head
  static constexpr int VPH = VPHEAD > LPR ? VPHEAD / LPR : 1;  // vectors of a lane per head
  static constexpr int NS = C >= 4 ? NV / VPH : NV * HPV;      // score slots per lane
  static constexpr int U = NV >= 8 ? 1 : (NV >= 4 ? 2 : 4);    // rows in flight per group
  static_assert(C >= 4 ? (C % 4 == 0 && (VPHEAD & (VPHEAD - 1)) == 0) : (4 % C == 0), "head dim");

  __device__ static __forceinline__ int slot_of(int v, int k) { return C >= 4 ? v / VPH : v * HPV + k / C; }
  __device__ static __forceinline__ int head_of_slot(int lir, int s) {
    return C >= 4 ? (lir + LPR * (s * VPH)) / VPHEAD : (lir + LPR * (s / HPV)) * HPV + (s % HPV);
  }
  __device__ static __forceinline__ bool slot_writer(int lir) { return C >= 4 ? (lir % LPH) == 0 : true; }
};

// Per-head reduction of NS partial sums over the lanes of a head.
template <class L>
__device__ __forceinline__ void head_reduce(float (&a)[L::NS], unsigned mask) {
  if (L::LPH > 1) {
#pragma unroll
    for (int off = L::LPH / 2; off > 0; off >>= 1) {
#pragma unroll
      for (int s = 0; s < L::NS; ++s) a[s] += __shfl_xor_sync(mask, a[s], off);
    }
  }
}
template <class L>
__device__ __forceinline__ void head_reduce2(float (&a)[L::NS], float (&b)[L::NS], unsigned mask) {
  if (L::LPH > 1) {
#pragma unroll
    for (int off = L::LPH / 2; off > 0; off >>= 1) {
#pragma unroll
      for (int s = 0; s < L::NS; ++s) {
        a[s] += __shfl_xor_sync(mask, a[s], off);
        b[s] += __shfl_xor_sync(mask, b[s], off);
      }
    }
  }
}

// scores of one row: sc[s] = sum over the head's channels of att * leaky(x + xr)
template <class L>
__device__ __forceinline__ void row_scores(const float4 (&x)[L::NV], const float4 (&xr)[L::NV],
                                           const float4 (&att)[L::NV], float slope, unsigned mask,
                                           float (&sc)[L::NS]) {
#pragma unroll
  for (int s = 0; s < L::NS; ++s) sc[s] = 0.f;
#pragma unroll
  for (int v = 0; v < L::NV; ++v) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float z = comp(x[v], k) + comp(xr[v], k);
      sc[L::slot_of(v, k)] = fmaf(comp(att[v], k), leaky(z, slope), sc[L::slot_of(v, k)]);
    }
  }
  head_reduce<L>(sc, mask);
}

// running softmax state of one lane group
template <class L>
struct Acc {
  float m[L::NS], l[L::NS];
  float4 o[L::NV];
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int s = 0; s < L::NS; ++s) { m[s] = -INFINITY; l[s] = 0.f; }
#pragma unroll
    for (int v = 0; v < L::NV; ++v) o[v] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  // fold another (m,l,o) triple into this one
  __device__ __forceinline__ void merge(const float (&m2)[L::NS], const float (&l2)[L::NS],
                                        const float4 (&o2)[L::NV]) {
    float a[L::NS], b[L::NS];
#pragma unroll
    for (int s = 0; s < L::NS; ++s) {
      float mn = fmaxf(m[s], m2[s]);
      float ms = mn == -INFINITY ? 0.f : mn;
      a[s] = __expf(m[s] - ms);
      b[s] = __expf(m2[s] - ms);
      l[s] = l[s] * a[s] + l2[s] * b[s];
      m[s] = mn;
    }
#pragma unroll
    for (int v = 0; v < L::NV; ++v) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        int s = L::slot_of(v, k);
        comp(o[v], k) = comp(o[v], k) * a[s] + comp(o2[v], k) * b[s];
      }
    }
  }
  // butterfly merge across the RPW lane groups of the warp (all lanes must participate)
  __device__ __forceinline__ void merge_across_groups() {
    if (L::RPW > 1) {
#pragma unroll
      for (int off = L::LPR; off < 32; off <<= 1) {
        float m2[L::NS], l2[L::NS];
        float4 o2[L::NV];
#pragma unroll
        for (int s = 0; s < L::NS; ++s) {
          m2[s] = __shfl_xor_sync(0xffffffffu, m[s], off);
          l2[s] = __shfl_xor_sync(0xffffffffu, l[s], off);
        }
#pragma unroll
        for (int v = 0; v < L::NV; ++v) {
          o2[v].x = __shfl_xor_sync(0xffffffffu, o[v].x, off);
          o2[v].y = __shfl_xor_sync(0xffffffffu, o[v].y, off);
          o2[v].z = __shfl_xor_sync(0xffffffffu, o[v].z, off);
          o2[v].w = __shfl_xor_sync(0xffffffffu, o[v].w, off);
        }
        merge(m2, l2, o2);
      }
    }
  }
};

// Storage type of the projected sources: fp32, or bf16 (BASELINE.json configs[4] "fp32 vs bf16": half the bytes per
// edge; arithmetic stays fp32).  BF is a template parameter, so the fp32 kernels compile exactly as before.
template <bool BF>
__device__ __forceinline__ float4 load_xl4(const float* base, int64_t row, int64_t ld, int col) {
  if constexpr (BF) {
    const uint16_t* q = reinterpret_cast<const uint16_t*>(base) + row * ld + col;
    uint32_t a, b;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(a), "=r"(b) : "l"(q));
    return make_float4(__uint_as_float(a << 16), __uint_as_float(a & 0xffff0000u), __uint_as_float(b << 16),
                       __uint_as_float(b & 0xffff0000u));
  } else {
    return ld_stream4(base + row * ld + col);
  }
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
template <bool BF>
__device__ __forceinline__ void store_dxl4(float* base, int64_t row, int64_t ld, int col, const float4& g) {
  if constexpr (BF) {
    uint16_t* q = reinterpret_cast<uint16_t*>(base) + row * ld + col;
    asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1,%2};" ::"l"(q), "r"(pack_bf16x2(g.x, g.y)), "r"(pack_bf16x2(g.z, g.w)) : "memory");
  } else {
    st_stream4(base + row * ld + col, g);
  }
}

struct GatFwdArgs {
  const float* XL; int64_t ldxl;
  const float* XR; int64_t ldxr;
  const float* att; const float* bias;
  const int32_t* seg_ptr; const int32_t* perm; int n_seg;
  int chunk; const int32_t* chunk_ptr; const int32_t* chunk_seg; int max_chunks;
  float slope; int normalize;
  float* out; float* seg_max; float* seg_sum;
  float* ws_o; float* ws_m; float* ws_l;
};

// consume rows [begin, end) with stride `step` into acc
template <class L, bool BF = false>
__device__ __forceinline__ void consume_rows(const GatFwdArgs& p, int begin, int end, int step,
                                             int lir, unsigned mask, const float4 (&xr)[L::NV],
                                             const float4 (&att)[L::NV], Acc<L>& acc) {
  for (int i = begin; i < end; i += step * L::U) {
    float4 x[L::U][L::NV];
    float sc[L::U][L::NS];
#pragma unroll
    for (int u = 0; u < L::U; ++u) {
      int pos = i + u * step;
      int posc = pos < end ? pos : begin;  // clamp: loads stay in range, score masked below
      int e = p.perm ? __ldg(p.perm + posc) : posc;
#pragma unroll
      for (int v = 0; v < L::NV; ++v) x[u][v] = load_xl4<BF>(p.XL, e, p.ldxl, 4 * lir + 4 * L::LPR * v);
    }
#pragma unroll
    for (int u = 0; u < L::U; ++u) {
      row_scores<L>(x[u], xr, att, p.slope, mask, sc[u]);
      if (i + u * step >= end) {
#pragma unroll
        for (int s = 0; s < L::NS; ++s) sc[u][s] = -INFINITY;
      }
    }
    float corr[L::NS], pw[L::U][L::NS];
#pragma unroll
    for (int s = 0; s < L::NS; ++s) {
      float mn = acc.m[s];
#pragma unroll
      for (int u = 0; u < L::U; ++u) mn = fmaxf(mn, sc[u][s]);
      float ms = mn == -INFINITY ? 0.f : mn;
      corr[s] = __expf(acc.m[s] - ms);
      float lsum = acc.l[s] * corr[s];
#pragma unroll
      for (int u = 0; u < L::U; ++u) {
        pw[u][s] = __expf(sc[u][s] - ms);
        lsum += pw[u][s];
      }
      acc.l[s] = lsum;
      acc.m[s] = mn;
    }
#pragma unroll
    for (int v = 0; v < L::NV; ++v) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        int s = L::slot_of(v, k);
        float o = comp(acc.o[v], k) * corr[s];
#pragma unroll
        for (int u = 0; u < L::U; ++u) o = fmaf(pw[u][s], comp(x[u][v], k), o);
        comp(acc.o[v], k) = o;
      }
    }
  }
}

template <class L>
__device__ __forceinline__ void finalize_segment(const GatFwdArgs& p, int t, int lir, int H,
                                                 const Acc<L>& acc) {
  float inv[L::NS];
#pragma unroll
  for (int s = 0; s < L::NS; ++s) inv[s] = (p.normalize && acc.l[s] > 0.f) ? 1.f / acc.l[s] : (p.normalize ? 0.f : 1.f);
  float* orow = p.out + (int64_t)t * L::HC + 4 * lir;
#pragma unroll
  for (int v = 0; v < L::NV; ++v) {
    float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.normalize && p.bias) b = ld4(p.bias + 4 * (lir + L::LPR * v));
    float4 r;
    r.x = acc.o[v].x * inv[L::slot_of(v, 0)] + b.x;
    r.y = acc.o[v].y * inv[L::slot_of(v, 1)] + b.y;
    r.z = acc.o[v].z * inv[L::slot_of(v, 2)] + b.z;
    r.w = acc.o[v].w * inv[L::slot_of(v, 3)] + b.w;
    st4(orow + 4 * L::LPR * v, r);
  }
  if (L::slot_writer(lir)) {
#pragma unroll
    for (int s = 0; s < L::NS; ++s) {
      int h = L::head_of_slot(lir, s);
      p.seg_max[(int64_t)t * H + h] = acc.m[s];
      p.seg_sum[(int64_t)t * H + h] = acc.l[s];
    }
  }
}

// Resident CTAs per SM (register cap), measured on B200 at H*C = 256 (profiles/r01_edge_kernel_occupancy_ab.md):
// the gathered short-segment schedule wants more warps in flight (3 CTAs, <= 80 regs, a few spills),
// the contiguous chunked schedule is faster with 2 CTAs and no spills.
template <int H, int C, bool CHUNKED, bool BF = false>
__global__ void __launch_bounds__(256, (H * C <= 256) ? (CHUNKED ? 2 : 3) : 1) gat_fwd_kernel(GatFwdArgs p) {
  using L = Lay<H, C>;
  const int lane = threadIdx.x & 31;
  const int lir = lane % L::LPR;
  const int grp = lane / L::LPR;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned gmask = group_mask<L::LPR>(lane);

  float4 att[L::NV], xr[L::NV];
#pragma unroll
  for (int v = 0; v < L::NV; ++v) att[v] = ld4(p.att + 4 * (lir + L::LPR * v));
  Acc<L> acc;
  acc.init();

  if (!CHUNKED) {
    const int64_t t64 = warp * L::RPW + grp;
    if (t64 >= p.n_seg) return;
    const int t = (int)t64;
    const float* xrrow = p.XR + (int64_t)t * p.ldxr + 4 * lir;
#pragma unroll
    for (int v = 0; v < L::NV; ++v) xr[v] = ld4(xrrow + 4 * L::LPR * v);
    const int b = __ldg(p.seg_ptr + t), e = __ldg(p.seg_ptr + t + 1);
    consume_rows<L, BF>(p, b, e, 1, lir, gmask, xr, att, acc);
    finalize_segment<L>(p, t, lir, H, acc);
  } else {
    const int total = __ldg(p.chunk_ptr + p.n_seg);
    if (warp >= total) return;
    const int k = (int)warp;
    const int t = __ldg(p.chunk_seg + k);
    const float* xrrow = p.XR + (int64_t)t * p.ldxr + 4 * lir;
#pragma unroll
    for (int v = 0; v < L::NV; ++v) xr[v] = ld4(xrrow + 4 * L::LPR * v);
    const int c0 = __ldg(p.chunk_ptr + t), c1 = __ldg(p.chunk_ptr + t + 1);
    const int sb = __ldg(p.seg_ptr + t), se = __ldg(p.seg_ptr + t + 1);
    const int b = sb + (k - c0) * p.chunk;
    const int e = min(b + p.chunk, se);
    consume_rows<L, BF>(p, b + grp, e, L::RPW, lir, gmask, xr, att, acc);
    __syncwarp();
    acc.merge_across_groups();
    if (grp == 0) {
      if (c1 - c0 == 1) {
        finalize_segment<L>(p, t, lir, H, acc);
      } else {
        float* orow = p.ws_o + (int64_t)k * L::HC + 4 * lir;
#pragma unroll
        for (int v = 0; v < L::NV; ++v) st4(orow + 4 * L::LPR * v, acc.o[v]);
        if (L::slot_writer(lir)) {
#pragma unroll
          for (int s = 0; s < L::NS; ++s) {
            int h = L::head_of_slot(lir, s);
            p.ws_m[(int64_t)k * H + h] = acc.m[s];
            p.ws_l[(int64_t)k * H + h] = acc.l[s];
          }
        }
      }
    }
  }
}

// One CTA per segment: combine the chunk partials of segments that span several chunks, and
// write bias / empty statistics for segments without edges.
constexpr int kMergeThreads = 128;
template <int H, int C>
__global__ void __launch_bounds__(kMergeThreads) gat_merge_kernel(GatFwdArgs p) {
  using L = Lay<H, C>;
  constexpr int NW = kMergeThreads / 32;
  const int t = blockIdx.x;
  const int c0 = __ldg(p.chunk_ptr + t), c1 = __ldg(p.chunk_ptr + t + 1);
  if (c1 - c0 == 1) return;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int lir = lane % L::LPR, grp = lane / L::LPR;
  Acc<L> acc;
  acc.init();
  for (int k = c0 + wid * L::RPW + grp; k < c1; k += NW * L::RPW) {
    float m2[L::NS], l2[L::NS];
    float4 o2[L::NV];
#pragma unroll
    for (int s = 0; s < L::NS; ++s) {
      int h = L::head_of_slot(lir, s);
      m2[s] = p.ws_m[(int64_t)k * H + h];
      l2[s] = p.ws_l[(int64_t)k * H + h];
    }
    const float* orow = p.ws_o + (int64_t)k * L::HC + 4 * lir;
#pragma unroll
    for (int v = 0; v < L::NV; ++v) o2[v] = *reinterpret_cast<const float4*>(orow + 4 * L::LPR * v);
    acc.merge(m2, l2, o2);
  }
  __syncwarp();
  acc.merge_across_groups();
  __shared__ float sm_m[NW][H], sm_l[NW][H];
  __shared__ float4 sm_o[NW][L::NVEC];
  if (grp == 0) {
#pragma unroll
    for (int v = 0; v < L::NV; ++v) sm_o[wid][lir + L::LPR * v] = acc.o[v];
    if (L::slot_writer(lir)) {
#pragma unroll
      for (int s = 0; s < L::NS; ++s) {
        sm_m[wid][L::head_of_slot(lir, s)] = acc.m[s];
        sm_l[wid][L::head_of_slot(lir, s)] = acc.l[s];
      }
    }
  }
  __syncthreads();
  if (wid == 0 && grp == 0) {
#pragma unroll
    for (int w = 1; w < NW; ++w) {
      float m2[L::NS], l2[L::NS];
      float4 o2[L::NV];
#pragma unroll
      for (int s = 0; s < L::NS; ++s) {
        m2[s] = sm_m[w][L::head_of_slot(lir, s)];
        l2[s] = sm_l[w][L::head_of_slot(lir, s)];
      }
#pragma unroll
      for (int v = 0; v < L::NV; ++v) o2[v] = sm_o[w][lir + L::LPR * v];
      acc.merge(m2, l2, o2);
    }
    finalize_segment<L>(p, t, lir, H, acc);
  }
}

// ---------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------
struct GatBwdArgs {
  const float* XL; int64_t ldxl;
  const float* XR; int64_t ldxr;
  const float* att; const float* out_nobias; const float* seg_max; const float* seg_sum;
  const float* dOut;
  const int32_t* seg_ptr; const int32_t* perm; int n_seg;
  int chunk; const int32_t* chunk_ptr; const int32_t* chunk_seg; int max_chunks;
  float slope;
  float* dXL; int64_t lddxl; float* dXR; float* ws_datt; float* ws_dxr;
  float* dxl_rowmax;   // optional [E]: max |dXL[e, :]| per edge row (the row scale of the fp16 input-gradient GEMM that reads dXL next)
};

constexpr int kBwdThreads = 256;

// per-target quantities held in registers while a group walks the target's edges
template <class L>
struct BwdSeg {
  float4 xr[L::NV], dO[L::NV];
  float M[L::NS], invL[L::NS], D[L::NS];
  __device__ __forceinline__ void load(const GatBwdArgs& p, int t, int lir, int H, unsigned mask) {
    const float* xrrow = p.XR + (int64_t)t * p.ldxr + 4 * lir;
    const float* dorow = p.dOut + (int64_t)t * L::HC + 4 * lir;
    const float* orow = p.out_nobias + (int64_t)t * L::HC + 4 * lir;
#pragma unroll
    for (int s = 0; s < L::NS; ++s) D[s] = 0.f;
#pragma unroll
    for (int v = 0; v < L::NV; ++v) {
      xr[v] = ld4(xrrow + 4 * L::LPR * v);
      dO[v] = ld4(dorow + 4 * L::LPR * v);
      float4 o = ld4(orow + 4 * L::LPR * v);
#pragma unroll
      for (int k = 0; k < 4; ++k) D[L::slot_of(v, k)] = fmaf(comp(dO[v], k), comp(o, k), D[L::slot_of(v, k)]);
    }
    head_reduce<L>(D, mask);
#pragma unroll
    for (int s = 0; s < L::NS; ++s) {
      int h = L::head_of_slot(lir, s);
      M[s] = __ldg(p.seg_max + (int64_t)t * H + h);
      float l = __ldg(p.seg_sum + (int64_t)t * H + h);
      invL[s] = l > 0.f ? 1.f / l : 0.f;
    }
  }
};

template <class L, bool BF = false, bool RM = false>   // RM: also emit the row maxima of dXL (p.dxl_rowmax)
__device__ __forceinline__ void bwd_rows(const GatBwdArgs& p, int begin, int end, int step, int lir,
                                         unsigned mask, const BwdSeg<L>& sg, const float4 (&att)[L::NV],
                                         float4 (&dxr)[L::NV], float4 (&datt)[L::NV]) {
  constexpr int U = L::NV >= 4 ? 1 : 2;
  for (int i = begin; i < end; i += step * U) {
    float4 x[U][L::NV];
    int eid[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      int pos = i + u * step;
      int posc = pos < end ? pos : begin;
      eid[u] = p.perm ? __ldg(p.perm + posc) : posc;
#pragma unroll
      for (int v = 0; v < L::NV; ++v) x[u][v] = load_xl4<BF>(p.XL, eid[u], p.ldxl, 4 * lir + 4 * L::LPR * v);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const bool valid = i + u * step < end;
      float sc[L::NS], da[L::NS];
#pragma unroll
      for (int s = 0; s < L::NS; ++s) { sc[s] = 0.f; da[s] = 0.f; }
#pragma unroll
      for (int v = 0; v < L::NV; ++v) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          int s = L::slot_of(v, k);
          float z = comp(x[u][v], k) + comp(sg.xr[v], k);
          sc[s] = fmaf(comp(att[v], k), leaky(z, p.slope), sc[s]);
          da[s] = fmaf(comp(sg.dO[v], k), comp(x[u][v], k), da[s]);
        }
      }
      head_reduce2<L>(sc, da, mask);
      float alpha[L::NS], ds[L::NS];
#pragma unroll
      for (int s = 0; s < L::NS; ++s) {
        alpha[s] = valid ? __expf(sc[s] - sg.M[s]) * sg.invL[s] : 0.f;
        ds[s] = alpha[s] * (da[s] - sg.D[s]);
      }
      float rmax = 0.f;
#pragma unroll
      for (int v = 0; v < L::NV; ++v) {
        float4 g;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          int s = L::slot_of(v, k);
          float z = comp(x[u][v], k) + comp(sg.xr[v], k);
          float dz = ds[s] * comp(att[v], k) * (z > 0.f ? 1.f : p.slope);
          comp(g, k) = fmaf(alpha[s], comp(sg.dO[v], k), dz);
          comp(dxr[v], k) += dz;
          comp(datt[v], k) = fmaf(ds[s], leaky(z, p.slope), comp(datt[v], k));
        }
        if (valid) store_dxl4<BF>(p.dXL, eid[u], p.lddxl, 4 * lir + 4 * L::LPR * v, g);
        if constexpr (RM) rmax = fmaxf(rmax, fmaxf(fmaxf(fabsf(g.x), fabsf(g.y)), fmaxf(fabsf(g.z), fabsf(g.w))));
      }
      if constexpr (RM) {
#pragma unroll
        for (int off = L::LPR / 2; off > 0; off >>= 1) rmax = fmaxf(rmax, __shfl_xor_sync(mask, rmax, off));
        if (lir == 0 && valid) p.dxl_rowmax[eid[u]] = rmax;
      }
    }
  }
}

template <class L>
__device__ __forceinline__ void sum_across_groups(float4 (&a)[L::NV]) {
  if (L::RPW > 1) {
#pragma unroll
    for (int off = L::LPR; off < 32; off <<= 1) {
#pragma unroll
      for (int v = 0; v < L::NV; ++v) {
        a[v].x += __shfl_xor_sync(0xffffffffu, a[v].x, off);
        a[v].y += __shfl_xor_sync(0xffffffffu, a[v].y, off);
        a[v].z += __shfl_xor_sync(0xffffffffu, a[v].z, off);
        a[v].w += __shfl_xor_sync(0xffffffffu, a[v].w, off);
      }
    }
  }
}

template <int H, int C, bool CHUNKED, bool BF = false, bool RM = false>
__global__ void __launch_bounds__(kBwdThreads, (H * C <= 256) ? (CHUNKED ? 2 : 3) : 1) gat_bwd_kernel(GatBwdArgs p) {
  using L = Lay<H, C>;
  constexpr int NW = kBwdThreads / 32;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int lir = lane % L::LPR, grp = lane / L::LPR;
  const unsigned gmask = group_mask<L::LPR>(lane);
  const int64_t warp0 = (int64_t)blockIdx.x * NW + wid;
  const int64_t nwarps = (int64_t)gridDim.x * NW;

  float4 att[L::NV], datt[L::NV];
#pragma unroll
  for (int v = 0; v < L::NV; ++v) {
    att[v] = ld4(p.att + 4 * (lir + L::LPR * v));
    datt[v] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  BwdSeg<L> sg;

  if (!CHUNKED) {
    for (int64_t t64 = warp0 * L::RPW + grp; t64 < p.n_seg; t64 += nwarps * L::RPW) {
      const int t = (int)t64;
      sg.load(p, t, lir, H, gmask);
      float4 dxr[L::NV];
#pragma unroll
      for (int v = 0; v < L::NV; ++v) dxr[v] = make_float4(0.f, 0.f, 0.f, 0.f);
      bwd_rows<L, BF, RM>(p, __ldg(p.seg_ptr + t), __ldg(p.seg_ptr + t + 1), 1, lir, gmask, sg, att, dxr, datt);
      float* xrow = p.dXR + (int64_t)t * L::HC + 4 * lir;
#pragma unroll
      for (int v = 0; v < L::NV; ++v) st4(xrow + 4 * L::LPR * v, dxr[v]);
    }
  } else {
    const int total = __ldg(p.chunk_ptr + p.n_seg);
    for (int64_t k64 = warp0; k64 < total; k64 += nwarps) {
      const int k = (int)k64;
      const int t = __ldg(p.chunk_seg + k);
      sg.load(p, t, lir, H, gmask);
      const int c0 = __ldg(p.chunk_ptr + t), c1 = __ldg(p.chunk_ptr + t + 1);
      const int sb = __ldg(p.seg_ptr + t), se = __ldg(p.seg_ptr + t + 1);
      const int b = sb + (k - c0) * p.chunk;
      const int e = min(b + p.chunk, se);
      float4 dxr[L::NV];
#pragma unroll
      for (int v = 0; v < L::NV; ++v) dxr[v] = make_float4(0.f, 0.f, 0.f, 0.f);
      bwd_rows<L, BF, RM>(p, b + grp, e, L::RPW, lir, gmask, sg, att, dxr, datt);
      __syncwarp();
      sum_across_groups<L>(dxr);
      if (grp == 0) {
        float* xrow = (c1 - c0 == 1) ? p.dXR + (int64_t)t * L::HC + 4 * lir
                                     : p.ws_dxr + (int64_t)k * L::HC + 4 * lir;
#pragma unroll
        for (int v = 0; v < L::NV; ++v) st4(xrow + 4 * L::LPR * v, dxr[v]);
      }
    }
  }
  // datt: groups -> warp -> CTA -> one row of the workspace per CTA (reduced by col_sum_kernel)
  __syncwarp();
  sum_across_groups<L>(datt);
  __shared__ float4 sm[NW][L::NVEC];
  if (grp == 0) {
#pragma unroll
    for (int v = 0; v < L::NV; ++v) sm[wid][lir + L::LPR * v] = datt[v];
  }
  __syncthreads();
  for (int j = threadIdx.x; j < L::NVEC; j += kBwdThreads) {
    float4 a = sm[0][j];
#pragma unroll
    for (int w = 1; w < NW; ++w) {
      float4 b = sm[w][j];
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    st4(p.ws_datt + (int64_t)blockIdx.x * L::HC + 4 * j, a);
  }
}

// dXR of multi-chunk segments (sum of chunk partials) and zero rows for empty segments.
// grid = (segments, column tiles of 32); the 8 warps of a CTA split the segment's chunks.
__global__ void __launch_bounds__(256) gat_bwd_merge_kernel(GatBwdArgs p, int HC) {
  __shared__ float sm[8][33];
  const int t = blockIdx.x;
  const int c0 = __ldg(p.chunk_ptr + t), c1 = __ldg(p.chunk_ptr + t + 1);
  if (c1 - c0 == 1) return;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int j = blockIdx.y * 32 + lane;
  float a = 0.f;
  if (j < HC)
    for (int k = c0 + wid; k < c1; k += 8) a += p.ws_dxr[(int64_t)k * HC + j];
  sm[wid][lane] = a;
  __syncthreads();
  if (wid == 0 && j < HC) {
    float r = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) r += sm[w][lane];
    p.dXR[(int64_t)t * HC + j] = r;
  }
}

// out[j] = sum_r ws[r, j]; deterministic two-stage column sum of a small [rows, width] matrix
__global__ void col_sum_kernel(const float* __restrict__ ws, int rows, int width, float* __restrict__ out) {
  __shared__ float sm[8][33];
  const int j = blockIdx.x * 32 + threadIdx.x;
  float a = 0.f;
  if (j < width)
    for (int r = threadIdx.y; r < rows; r += 8) a += ws[(int64_t)r * width + j];
  sm[threadIdx.y][threadIdx.x] = a;
  __syncthreads();
  if (threadIdx.y == 0 && j < width) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += sm[w][threadIdx.x];
    out[j] = s;
  }
}

// ---------------------------------------------------------------------------------------------
// Generic fallback: any (H <= 8, C, H*C <= 1024); one warp per segment, lane-strided channels.
// Only used for head shapes without a vectorised instantiation (never for the shipped confs).
// ---------------------------------------------------------------------------------------------
constexpr int kGenMaxH = 8;
template <int NI>
__global__ void __launch_bounds__(128) gat_fwd_generic_kernel(GatFwdArgs p, int H, int C) {
  const int lane = threadIdx.x & 31;
  const int t = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (t >= p.n_seg) return;
  const int HC = H * C;
  float att[NI], xr[NI], o[NI];
  int hd[NI];
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    int j = lane + 32 * i;
    bool ok = j < HC;
    att[i] = ok ? p.att[j] : 0.f;
    xr[i] = ok ? p.XR[(int64_t)t * p.ldxr + j] : 0.f;
    hd[i] = ok ? j / C : -1;
    o[i] = 0.f;
  }
  float m[kGenMaxH], l[kGenMaxH];
#pragma unroll
  for (int h = 0; h < kGenMaxH; ++h) { m[h] = -INFINITY; l[h] = 0.f; }
  const int b = p.seg_ptr[t], e = p.seg_ptr[t + 1];
  for (int pos = b; pos < e; ++pos) {
    const int eid = p.perm ? p.perm[pos] : pos;
    float x[NI], lz[NI];
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      int j = lane + 32 * i;
      x[i] = j < HC ? p.XL[(int64_t)eid * p.ldxl + j] : 0.f;
      lz[i] = att[i] * leaky(x[i] + xr[i], p.slope);
    }
#pragma unroll
    for (int h = 0; h < kGenMaxH; ++h) {
      if (h < H) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < NI; ++i) s += hd[i] == h ? lz[i] : 0.f;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        float mn = fmaxf(m[h], s);
        float corr = __expf(m[h] - mn), pw = __expf(s - mn);
        l[h] = l[h] * corr + pw;
        m[h] = mn;
#pragma unroll
        for (int i = 0; i < NI; ++i)
          if (hd[i] == h) o[i] = o[i] * corr + pw * x[i];
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    int j = lane + 32 * i;
    if (j < HC) {
      float lh = 0.f;
#pragma unroll
      for (int h = 0; h < kGenMaxH; ++h) lh = hd[i] == h ? l[h] : lh;
      float r = o[i];
      if (p.normalize) r = (lh > 0.f ? r / lh : 0.f) + (p.bias ? p.bias[j] : 0.f);
      p.out[(int64_t)t * HC + j] = r;
    }
  }
  if (lane == 0) {
#pragma unroll
    for (int h = 0; h < kGenMaxH; ++h)
      if (h < H) { p.seg_max[(int64_t)t * H + h] = m[h]; p.seg_sum[(int64_t)t * H + h] = l[h]; }
  }
}

template <int NI>
__global__ void __launch_bounds__(128) gat_bwd_generic_kernel(GatBwdArgs p, int H, int C) {
  const int lane = threadIdx.x & 31;
  const int wid = threadIdx.x >> 5;
  const int HC = H * C;
  float att[NI], datt[NI];
  int hd[NI];
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    int j = lane + 32 * i;
    att[i] = j < HC ? p.att[j] : 0.f;
    hd[i] = j < HC ? j / C : -1;
    datt[i] = 0.f;
  }
  const int64_t nwarps = (int64_t)gridDim.x * 4;
  for (int64_t t64 = (int64_t)blockIdx.x * 4 + wid; t64 < p.n_seg; t64 += nwarps) {
    const int t = (int)t64;
    float xr[NI], dO[NI], dxr[NI], Dl[NI];
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      int j = lane + 32 * i;
      bool ok = j < HC;
      xr[i] = ok ? p.XR[(int64_t)t * p.ldxr + j] : 0.f;
      dO[i] = ok ? p.dOut[(int64_t)t * HC + j] : 0.f;
      Dl[i] = ok ? dO[i] * p.out_nobias[(int64_t)t * HC + j] : 0.f;
      dxr[i] = 0.f;
    }
    float M[kGenMaxH], invL[kGenMaxH], D[kGenMaxH];
#pragma unroll
    for (int h = 0; h < kGenMaxH; ++h) {
      M[h] = 0.f; invL[h] = 0.f; D[h] = 0.f;
      if (h < H) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < NI; ++i) s += hd[i] == h ? Dl[i] : 0.f;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        D[h] = s;
        M[h] = p.seg_max[(int64_t)t * H + h];
        float l = p.seg_sum[(int64_t)t * H + h];
        invL[h] = l > 0.f ? 1.f / l : 0.f;
      }
    }
    const int b = p.seg_ptr[t], e = p.seg_ptr[t + 1];
    for (int pos = b; pos < e; ++pos) {
      const int eid = p.perm ? p.perm[pos] : pos;
      float x[NI], z[NI], g[NI];
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        int j = lane + 32 * i;
        x[i] = j < HC ? p.XL[(int64_t)eid * p.ldxl + j] : 0.f;
        z[i] = x[i] + xr[i];
        g[i] = 0.f;
      }
#pragma unroll
      for (int h = 0; h < kGenMaxH; ++h) {
        if (h < H) {
          float s = 0.f, da = 0.f;
#pragma unroll
          for (int i = 0; i < NI; ++i) {
            s += hd[i] == h ? att[i] * leaky(z[i], p.slope) : 0.f;
            da += hd[i] == h ? dO[i] * x[i] : 0.f;
          }
#pragma unroll
          for (int off = 16; off > 0; off >>= 1) {
            s += __shfl_xor_sync(0xffffffffu, s, off);
            da += __shfl_xor_sync(0xffffffffu, da, off);
          }
          float alpha = __expf(s - M[h]) * invL[h];
          float ds = alpha * (da - D[h]);
#pragma unroll
          for (int i = 0; i < NI; ++i) {
            if (hd[i] == h) {
              float dz = ds * att[i] * (z[i] > 0.f ? 1.f : p.slope);
              g[i] = alpha * dO[i] + dz;
              dxr[i] += dz;
              datt[i] += ds * leaky(z[i], p.slope);
            }
          }
        }
      }
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        int j = lane + 32 * i;
        if (j < HC) p.dXL[(int64_t)eid * p.lddxl + j] = g[i];
      }
    }
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      int j = lane + 32 * i;
      if (j < HC) p.dXR[(int64_t)t * HC + j] = dxr[i];
    }
  }
  __shared__ float sm[4][32 * NI];
#pragma unroll
  for (int i = 0; i < NI; ++i) sm[wid][lane + 32 * i] = datt[i];
  __syncthreads();
  for (int j = threadIdx.x; j < HC; j += blockDim.x)
    p.ws_datt[(int64_t)blockIdx.x * HC + j] = sm[0][j] + sm[1][j] + sm[2][j] + sm[3][j];
}

// ---------------------------------------------------------------------------------------------
// dispatch
// ---------------------------------------------------------------------------------------------
static bool has_fast_path(int H, int C) {
  if (H != 4) return false;
  switch (C) { case 1: case 2: case 4: case 8: case 16: case 32: case 64: case 128: case 256: return true; }
  return false;
}

static int bwd_grid_blocks(int64_t work_warps, int warps_per_block) {
  int64_t need = (work_warps + warps_per_block - 1) / warps_per_block;
  int64_t cap = (int64_t)kNumSMs * 8;
  return (int)(need < 1 ? 1 : (need < cap ? need : cap));
}

template <int H, int C, bool BF = false>
static int launch_fwd(const GatFwdArgs& a, cudaStream_t st) {
  using L = Lay<H, C>;
  if (a.chunk == 0) {
    int64_t warps = ((int64_t)a.n_seg + L::RPW - 1) / L::RPW;
    int blocks = ceil_div(warps, 8);
    if (blocks > 0) gat_fwd_kernel<H, C, false, BF><<<blocks, 256, 0, st>>>(a);
  } else {
    int blocks = ceil_div(a.max_chunks, 8);
    if (blocks > 0) gat_fwd_kernel<H, C, true, BF><<<blocks, 256, 0, st>>>(a);
    if (a.n_seg > 0) gat_merge_kernel<H, C><<<a.n_seg, kMergeThreads, 0, st>>>(a);
  }
  return check_launch("gat_edge_fwd");
}

template <int H, int C, bool BF = false, bool RM = false>
static int launch_bwd(const GatBwdArgs& a, int* n_blocks_out, cudaStream_t st) {
  using L = Lay<H, C>;
  int blocks;
  if (a.chunk == 0) {
    blocks = bwd_grid_blocks(((int64_t)a.n_seg + L::RPW - 1) / L::RPW, kBwdThreads / 32);
    gat_bwd_kernel<H, C, false, BF, RM><<<blocks, kBwdThreads, 0, st>>>(a);
  } else {
    blocks = bwd_grid_blocks(a.max_chunks, kBwdThreads / 32);
    gat_bwd_kernel<H, C, true, BF, RM><<<blocks, kBwdThreads, 0, st>>>(a);
    if (a.n_seg > 0) gat_bwd_merge_kernel<<<dim3(a.n_seg, (H * C + 31) / 32), 256, 0, st>>>(a, H * C);
  }
  *n_blocks_out = blocks;
  return check_launch("gat_edge_bwd");
}

#define GASFM_DISPATCH_C(FN, ...)                     \
  switch (head_dim) {                                 \
    case 1: rc = FN<4, 1>(__VA_ARGS__); break;        \
    case 2: rc = FN<4, 2>(__VA_ARGS__); break;        \
    case 4: rc = FN<4, 4>(__VA_ARGS__); break;        \
    case 8: rc = FN<4, 8>(__VA_ARGS__); break;        \
    case 16: rc = FN<4, 16>(__VA_ARGS__); break;      \
    case 32: rc = FN<4, 32>(__VA_ARGS__); break;      \
    case 64: rc = FN<4, 64>(__VA_ARGS__); break;      \
    case 128: rc = FN<4, 128>(__VA_ARGS__); break;    \
    case 256: rc = FN<4, 256>(__VA_ARGS__); break;    \
    default: rc = 1; break;                           \
  }

static int generic_ni(int HC) {
  int ni = 1;
  while (ni * 32 < HC) ni <<= 1;
  return ni;
}

}  // namespace gasfm

using namespace gasfm;

extern "C" size_t gasfm_gat_ws_bytes(int max_chunks, int heads, int head_dim) {
  return (size_t)max_chunks * (size_t)(heads * head_dim + 2 * heads) * sizeof(float);
}

static int gat_edge_fwd_impl(bool bf16, const float* XL, int64_t ldxl, const float* XR, int64_t ldxr,
                                  const float* att, const float* bias, const int32_t* seg_ptr,
                                  const int32_t* perm, int n_seg, int chunk, const int32_t* chunk_ptr,
                                  const int32_t* chunk_seg, int max_chunks, int heads, int head_dim,
                                  float slope, int normalize, float* out, float* seg_max,
                                  float* seg_sum, void* ws, void* stream) {
  GASFM_REQUIRE(heads > 0 && head_dim > 0, "gat_edge_fwd: bad head shape %d x %d", heads, head_dim);
  GASFM_REQUIRE(n_seg >= 0, "gat_edge_fwd: negative segment count");
  if (n_seg == 0) return 0;
  const int HC = heads * head_dim;
  cudaStream_t st = (cudaStream_t)stream;
  GatFwdArgs a{XL, ldxl, XR, ldxr, att, bias, seg_ptr, perm, n_seg, chunk, chunk_ptr, chunk_seg,
               max_chunks, slope, normalize, out, seg_max, seg_sum, nullptr, nullptr, nullptr};
  if (has_fast_path(heads, head_dim)) {
    GASFM_REQUIRE(ldxl % 4 == 0 && ldxr % 4 == 0, "gat_edge_fwd: row strides must be multiples of 4 floats");
    GASFM_REQUIRE(((uintptr_t)XL | (uintptr_t)XR | (uintptr_t)att | (uintptr_t)out | (uintptr_t)bias) % 16 == 0,
                  "gat_edge_fwd: pointers must be 16-byte aligned");
    if (chunk > 0) {
      GASFM_REQUIRE(ws != nullptr && chunk_ptr && chunk_seg, "gat_edge_fwd: chunked plan needs workspace and chunk tables");
      a.ws_o = (float*)ws;
      a.ws_m = a.ws_o + (size_t)max_chunks * HC;
      a.ws_l = a.ws_m + (size_t)max_chunks * heads;
    }
    int rc;
    if (bf16) {
      GASFM_REQUIRE(heads == 4 && (head_dim == 32 || head_dim == 64), "gat_edge_fwd_bf16: head shapes 4 x 32 and 4 x 64 only");
      rc = head_dim == 32 ? launch_fwd<4, 32, true>(a, st) : launch_fwd<4, 64, true>(a, st);
      return rc;
    }
    GASFM_DISPATCH_C(launch_fwd, a, st);
    return rc;
  }
  GASFM_REQUIRE(!bf16, "gat_edge_fwd_bf16: head shapes 4 x 32 and 4 x 64 only");
  GASFM_REQUIRE(heads <= kGenMaxH && HC <= 1024, "gat_edge_fwd: unsupported head shape %d x %d", heads, head_dim);
  int blocks = ceil_div(n_seg, 4);
  switch (generic_ni(HC)) {
    case 1: gat_fwd_generic_kernel<1><<<blocks, 128, 0, st>>>(a, heads, head_dim); break;
    case 2: gat_fwd_generic_kernel<2><<<blocks, 128, 0, st>>>(a, heads, head_dim); break;
    case 4: gat_fwd_generic_kernel<4><<<blocks, 128, 0, st>>>(a, heads, head_dim); break;
    case 8: gat_fwd_generic_kernel<8><<<blocks, 128, 0, st>>>(a, heads, head_dim); break;
    case 16: gat_fwd_generic_kernel<16><<<blocks, 128, 0, st>>>(a, heads, head_dim); break;
    default: gat_fwd_generic_kernel<32><<<blocks, 128, 0, st>>>(a, heads, head_dim); break;
  }
  return check_launch("gat_edge_fwd(generic)");
}

extern "C" int gasfm_gat_edge_fwd(const float* XL, int64_t ldxl, const float* XR, int64_t ldxr,
                                  const float* att, const float* bias, const int32_t* seg_ptr,
                                  const int32_t* perm, int n_seg, int chunk, const int32_t* chunk_ptr,
                                  const int32_t* chunk_seg, int max_chunks, int heads, int head_dim,
                                  float slope, int normalize, float* out, float* seg_max,
                                  float* seg_sum, void* ws, void* stream) {
  return gat_edge_fwd_impl(false, XL, ldxl, XR, ldxr, att, bias, seg_ptr, perm, n_seg, chunk, chunk_ptr, chunk_seg, max_chunks,
                           heads, head_dim, slope, normalize, out, seg_max, seg_sum, ws, stream);
}

extern "C" int gasfm_gat_edge_fwd_bf16(const void* XL_bf16, int64_t ldxl, const float* XR, int64_t ldxr,
                                       const float* att, const float* bias, const int32_t* seg_ptr,
                                       const int32_t* perm, int n_seg, int chunk, const int32_t* chunk_ptr,
                                       const int32_t* chunk_seg, int max_chunks, int heads, int head_dim,
                                       float slope, int normalize, float* out, float* seg_max,
                                       float* seg_sum, void* ws, void* stream) {
  GASFM_REQUIRE((uintptr_t)XL_bf16 % 8 == 0, "gat_edge_fwd_bf16: XL must be 8-byte aligned");
  return gat_edge_fwd_impl(true, (const float*)XL_bf16, ldxl, XR, ldxr, att, bias, seg_ptr, perm, n_seg, chunk, chunk_ptr, chunk_seg,
                           max_chunks, heads, head_dim, slope, normalize, out, seg_max, seg_sum, ws, stream);
}

extern "C" size_t gasfm_gat_bwd_ws_bytes(int64_t n_obs, int n_seg, int max_chunks, int heads, int head_dim) {
  (void)n_obs;
  const size_t HC = (size_t)heads * head_dim;
  size_t blocks = (size_t)kNumSMs * 8;
  if (!has_fast_path(heads, head_dim)) blocks = (size_t)kNumSMs * 16;
  (void)n_seg;
  return (blocks * HC + (size_t)max_chunks * HC) * sizeof(float);
}

static int gat_edge_bwd_impl(bool bf16, float* dxl_rowmax, const float* XL, int64_t ldxl, const float* XR, int64_t ldxr,
                                  const float* att, const float* out_nobias, const float* seg_max,
                                  const float* seg_sum, const float* dOut, const int32_t* seg_ptr,
                                  const int32_t* perm, int n_seg, int chunk, const int32_t* chunk_ptr,
                                  const int32_t* chunk_seg, int max_chunks, int heads, int head_dim,
                                  float slope, float* dXL, int64_t lddxl, float* dXR, float* datt,
                                  void* ws, void* stream) {
  GASFM_REQUIRE(heads > 0 && head_dim > 0, "gat_edge_bwd: bad head shape");
  GASFM_REQUIRE(ws != nullptr, "gat_edge_bwd: workspace required");
  const int HC = heads * head_dim;
  cudaStream_t st = (cudaStream_t)stream;
  if (n_seg == 0) {
    cudaMemsetAsync(datt, 0, HC * sizeof(float), st);
    return check_launch("gat_edge_bwd(empty)");
  }
  GatBwdArgs a{XL, ldxl, XR, ldxr, att, out_nobias, seg_max, seg_sum, dOut, seg_ptr, perm, n_seg,
               chunk, chunk_ptr, chunk_seg, max_chunks, slope, dXL, lddxl, dXR, (float*)ws, nullptr, dxl_rowmax};
  int blocks = 0, rc;
  if (has_fast_path(heads, head_dim)) {
    GASFM_REQUIRE(ldxl % 4 == 0 && ldxr % 4 == 0 && lddxl % 4 == 0, "gat_edge_bwd: row strides must be multiples of 4 floats");
    GASFM_REQUIRE(((uintptr_t)XR | (uintptr_t)att | (uintptr_t)out_nobias | (uintptr_t)dOut | (uintptr_t)dXR) % 16 == 0 &&
                  ((uintptr_t)XL | (uintptr_t)dXL) % (bf16 ? 8 : 16) == 0, "gat_edge_bwd: pointers must be 16-byte aligned");
    a.ws_dxr = (float*)ws + (size_t)kNumSMs * 8 * HC;
    if (bf16) {
      GASFM_REQUIRE(heads == 4 && (head_dim == 32 || head_dim == 64), "gat_edge_bwd_bf16: head shapes 4 x 32 and 4 x 64 only");
      rc = head_dim == 32 ? launch_bwd<4, 32, true>(a, &blocks, st) : launch_bwd<4, 64, true>(a, &blocks, st);
    } else if (dxl_rowmax != nullptr) {
      switch (head_dim) {                          // row maxima: the widths the fp16 input-gradient GEMM takes (segments of 128 / 256)
        case 32: rc = launch_bwd<4, 32, false, true>(a, &blocks, st); break;
        case 64: rc = launch_bwd<4, 64, false, true>(a, &blocks, st); break;
        default: set_error("gat_edge_bwd_rowmax: head shapes 4 x 32 and 4 x 64 only"); rc = 1; break;
      }
    } else {
      GASFM_DISPATCH_C(launch_bwd, a, &blocks, st);
    }
    if (rc) return rc;
  } else {
    GASFM_REQUIRE(!bf16, "gat_edge_bwd_bf16: head shapes 4 x 32 and 4 x 64 only");
    GASFM_REQUIRE(heads <= kGenMaxH && HC <= 1024, "gat_edge_bwd: unsupported head shape %d x %d", heads, head_dim);
    int64_t need = ((int64_t)n_seg + 3) / 4;
    blocks = (int)(need < (int64_t)kNumSMs * 16 ? need : (int64_t)kNumSMs * 16);
    switch (generic_ni(HC)) {
      case 1: gat_bwd_generic_kernel<1><<<blocks, 128, 0, st>>>(a, heads, head_dim); break;
      case 2: gat_bwd_generic_kernel<2><<<blocks, 128, 0, st>>>(a, heads, head_dim); break;
      case 4: gat_bwd_generic_kernel<4><<<blocks, 128, 0, st>>>(a, heads, head_dim); break;
      case 8: gat_bwd_generic_kernel<8><<<blocks, 128, 0, st>>>(a, heads, head_dim); break;
      case 16: gat_bwd_generic_kernel<16><<<blocks, 128, 0, st>>>(a, heads, head_dim); break;
      default: gat_bwd_generic_kernel<32><<<blocks, 128, 0, st>>>(a, heads, head_dim); break;
    }
    rc = check_launch("gat_edge_bwd(generic)");
    if (rc) return rc;
  }
  col_sum_kernel<<<ceil_div(HC, 32), dim3(32, 8), 0, st>>>((const float*)ws, blocks, HC, datt);
  return check_launch("gat_edge_bwd(datt)");
}

extern "C" int gasfm_gat_edge_bwd(const float* XL, int64_t ldxl, const float* XR, int64_t ldxr,
                                  const float* att, const float* out_nobias, const float* seg_max,
                                  const float* seg_sum, const float* dOut, const int32_t* seg_ptr,
                                  const int32_t* perm, int n_seg, int chunk, const int32_t* chunk_ptr,
                                  const int32_t* chunk_seg, int max_chunks, int heads, int head_dim,
                                  float slope, float* dXL, int64_t lddxl, float* dXR, float* datt,
                                  void* ws, void* stream) {
  return gat_edge_bwd_impl(false, nullptr, XL, ldxl, XR, ldxr, att, out_nobias, seg_max, seg_sum, dOut, seg_ptr, perm, n_seg, chunk,
                           chunk_ptr, chunk_seg, max_chunks, heads, head_dim, slope, dXL, lddxl, dXR, datt, ws, stream);
}

extern "C" int gasfm_gat_edge_bwd_rowmax_supported(int heads, int head_dim) { return (heads == 4 && (head_dim == 32 || head_dim == 64)) ? 1 : 0; }

extern "C" int gasfm_gat_edge_bwd_rowmax(const float* XL, int64_t ldxl, const float* XR, int64_t ldxr,
                                         const float* att, const float* out_nobias, const float* seg_max,
                                         const float* seg_sum, const float* dOut, const int32_t* seg_ptr,
                                         const int32_t* perm, int n_seg, int chunk, const int32_t* chunk_ptr,
                                         const int32_t* chunk_seg, int max_chunks, int heads, int head_dim,
                                         float slope, float* dXL, int64_t lddxl, float* dXR, float* datt,
                                         float* dxl_rowmax, void* ws, void* stream) {
  GASFM_REQUIRE(dxl_rowmax != nullptr && gasfm_gat_edge_bwd_rowmax_supported(heads, head_dim), "gat_edge_bwd_rowmax: head shapes 4 x 32 and 4 x 64 only");
  return gat_edge_bwd_impl(false, dxl_rowmax, XL, ldxl, XR, ldxr, att, out_nobias, seg_max, seg_sum, dOut, seg_ptr, perm, n_seg, chunk,
                           chunk_ptr, chunk_seg, max_chunks, heads, head_dim, slope, dXL, lddxl, dXR, datt, ws, stream);
}

extern "C" int gasfm_gat_edge_bwd_bf16(const void* XL_bf16, int64_t ldxl, const float* XR, int64_t ldxr,
                                       const float* att, const float* out_nobias, const float* seg_max,
                                       const float* seg_sum, const float* dOut, const int32_t* seg_ptr,
                                       const int32_t* perm, int n_seg, int chunk, const int32_t* chunk_ptr,
                                       const int32_t* chunk_seg, int max_chunks, int heads, int head_dim,
                                       float slope, void* dXL_bf16, int64_t lddxl, float* dXR, float* datt,
                                       void* ws, void* stream) {
  return gat_edge_bwd_impl(true, nullptr, (const float*)XL_bf16, ldxl, XR, ldxr, att, out_nobias, seg_max, seg_sum, dOut, seg_ptr, perm,
                           n_seg, chunk, chunk_ptr, chunk_seg, max_chunks, heads, head_dim, slope, (float*)dXL_bf16, lddxl, dXR,
                           datt, ws, stream);
}
