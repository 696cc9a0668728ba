// Per-observation feature kernels: LayerNorm+ReLU, gather-add update, segment pooling (sm_100a).
// All of them are single-pass, HBM-bound streams over [E, width] fp32 matrices.
#include "common.cuh"
#include "col_reduce.cuh"
#include "../../include/gasfm_b200.h"

namespace gasfm {

// A row of `width` floats is handled by LPR lanes, each owning NV vectors of VEC floats
// (vector index lane + LPR*v), predicated on the true width.
template <int VEC> struct VecT;
template <> struct VecT<4> {
  using T = float4;
  __device__ static __forceinline__ T ld(const float* p) { return *reinterpret_cast<const float4*>(p); }
  __device__ static __forceinline__ T ld_stream(const float* p) { return ld_stream4(p); }
  __device__ static __forceinline__ void st(float* p, T v) { *reinterpret_cast<float4*>(p) = v; }
  __device__ static __forceinline__ T zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
};
template <> struct VecT<1> {
  using T = float;
  __device__ static __forceinline__ T ld(const float* p) { return *p; }
  __device__ static __forceinline__ T ld_stream(const float* p) { return __ldg(p); }
  __device__ static __forceinline__ void st(float* p, T v) { *p = v; }
  __device__ static __forceinline__ T zero() { return 0.f; }
};
__device__ __forceinline__ float get(const float& v, int) { return v; }
__device__ __forceinline__ float& get(float& v, int) { return v; }
__device__ __forceinline__ float get(const float4& v, int k) { return comp(v, k); }
__device__ __forceinline__ float& get(float4& v, int k) { return comp(v, k); }

template <int LPR>
__device__ __forceinline__ float group_sum(float a, unsigned mask) {
#pragma unroll
  for (int off = LPR / 2; off > 0; off >>= 1) a += __shfl_xor_sync(mask, a, off);
  return a;
}

// =============================================================================================
// segment sum / mean
// =============================================================================================
struct SegSumArgs {
  const float* X; int64_t ldx; int width;
  const int32_t* seg_ptr; const int32_t* perm; int n_seg;
  int chunk; const int32_t* chunk_ptr; const int32_t* chunk_seg; int max_chunks;
  float scale; int mean_mode;
  float* out; float* ws;
};

template <int VEC, int LPR, int NV>
__device__ __forceinline__ void seg_accumulate(const SegSumArgs& p, int begin, int end, int step, int lir,
                                               int nvec, typename VecT<VEC>::T (&acc)[NV]) {
  using V = VecT<VEC>;
  constexpr int U = NV >= 4 ? 1 : 4;
  for (int i = begin; i < end; i += step * U) {
    typename V::T x[U][NV];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      int pos = i + u * step;
      bool ok = pos < end;
      int posc = ok ? pos : begin;
      int e = p.perm ? __ldg(p.perm + posc) : posc;
      const float* row = p.X + (int64_t)e * p.ldx;
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        int vi = lir + LPR * v;
        x[u][v] = (ok && vi < nvec) ? V::ld_stream(row + VEC * vi) : V::zero();
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int v = 0; v < NV; ++v)
#pragma unroll
        for (int k = 0; k < VEC; ++k) get(acc[v], k) += get(x[u][v], k);
  }
}

template <int VEC, int LPR, int NV, bool CHUNKED>
__global__ void __launch_bounds__(256) seg_sum_kernel(SegSumArgs p) {
  using V = VecT<VEC>;
  constexpr int RPW = 32 / LPR;
  const int lane = threadIdx.x & 31;
  const int lir = lane % LPR, grp = lane / LPR;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nvec = p.width / VEC;
  typename V::T acc[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) acc[v] = V::zero();
  if (!CHUNKED) {
    const int64_t t64 = warp * RPW + grp;
    if (t64 >= p.n_seg) return;
    const int t = (int)t64;
    const int b = __ldg(p.seg_ptr + t), e = __ldg(p.seg_ptr + t + 1);
    seg_accumulate<VEC, LPR, NV>(p, b, e, 1, lir, nvec, acc);
    const float f = p.mean_mode ? (e > b ? p.scale / (float)(e - b) : 0.f) : p.scale;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      int vi = lir + LPR * v;
      if (vi < nvec) {
#pragma unroll
        for (int k = 0; k < VEC; ++k) get(acc[v], k) *= f;
        V::st(p.out + (int64_t)t * p.width + VEC * vi, acc[v]);
      }
    }
  } else {
    const int total = __ldg(p.chunk_ptr + p.n_seg);
    if (warp >= total) return;
    const int k = (int)warp;
    const int t = __ldg(p.chunk_seg + k);
    const int c0 = __ldg(p.chunk_ptr + t), c1 = __ldg(p.chunk_ptr + t + 1);
    const int sb = __ldg(p.seg_ptr + t), se = __ldg(p.seg_ptr + t + 1);
    const int b = sb + (k - c0) * p.chunk, e = min(b + p.chunk, se);
    seg_accumulate<VEC, LPR, NV>(p, b + grp, e, RPW, lir, nvec, acc);
    __syncwarp();
    if (RPW > 1) {
#pragma unroll
      for (int off = LPR; off < 32; off <<= 1)
#pragma unroll
        for (int v = 0; v < NV; ++v)
#pragma unroll
          for (int kk = 0; kk < VEC; ++kk) get(acc[v], kk) += __shfl_xor_sync(0xffffffffu, get(acc[v], kk), off);
    }
    if (grp == 0) {
      const bool single = (c1 - c0) == 1;
      const float f = single ? (p.mean_mode ? p.scale / (float)(se - sb) : p.scale) : 1.f;
      float* dst = single ? p.out + (int64_t)t * p.width : p.ws + (int64_t)k * p.width;
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        int vi = lir + LPR * v;
        if (vi < nvec) {
#pragma unroll
          for (int kk = 0; kk < VEC; ++kk) get(acc[v], kk) *= f;
          V::st(dst + VEC * vi, acc[v]);
        }
      }
    }
  }
}

// grid = (segments, column tiles of 32); the 8 warps of a CTA split the segment's chunks.
__global__ void __launch_bounds__(256) seg_sum_merge_kernel(SegSumArgs p) {
  __shared__ float sm[8][33];
  const int t = blockIdx.x;
  const int c0 = __ldg(p.chunk_ptr + t), c1 = __ldg(p.chunk_ptr + t + 1);
  if (c1 - c0 == 1) return;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int j = blockIdx.y * 32 + lane;
  const int len = __ldg(p.seg_ptr + t + 1) - __ldg(p.seg_ptr + t);
  const float f = p.mean_mode ? (len > 0 ? p.scale / (float)len : 0.f) : p.scale;
  float a = 0.f;
  if (j < p.width)
    for (int k = c0 + wid; k < c1; k += 8) a += p.ws[(int64_t)k * p.width + j];
  sm[wid][lane] = a;
  __syncthreads();
  if (wid == 0 && j < p.width) {
    float r = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) r += sm[w][lane];
    p.out[(int64_t)t * p.width + j] = r * f;
  }
}

template <int VEC, int LPR, int NV>
static void launch_seg_sum(const SegSumArgs& a, cudaStream_t st) {
  constexpr int RPW = 32 / LPR;
  if (a.chunk == 0) {
    int64_t warps = ((int64_t)a.n_seg + RPW - 1) / RPW;
    seg_sum_kernel<VEC, LPR, NV, false><<<ceil_div(warps, 8), 256, 0, st>>>(a);
  } else {
    seg_sum_kernel<VEC, LPR, NV, true><<<ceil_div(a.max_chunks, 8), 256, 0, st>>>(a);
    seg_sum_merge_kernel<<<dim3(a.n_seg, (a.width + 31) / 32), 256, 0, st>>>(a);
  }
}

// picks (LPR, NV) for a row of nvec vectors: LPR = min(32, pow2ceil(nvec)), NV = pow2ceil(nvec/LPR)
#define GASFM_ROW_DISPATCH(VEC, nvec, CALL)                        \
  do {                                                             \
    if ((nvec) <= 1) { CALL(VEC, 1, 1); }                          \
    else if ((nvec) <= 2) { CALL(VEC, 2, 1); }                     \
    else if ((nvec) <= 4) { CALL(VEC, 4, 1); }                     \
    else if ((nvec) <= 8) { CALL(VEC, 8, 1); }                     \
    else if ((nvec) <= 16) { CALL(VEC, 16, 1); }                   \
    else if ((nvec) <= 32) { CALL(VEC, 32, 1); }                   \
    else if ((nvec) <= 64) { CALL(VEC, 32, 2); }                   \
    else if ((nvec) <= 128) { CALL(VEC, 32, 4); }                  \
    else { CALL(VEC, 32, 8); }                                     \
  } while (0)

// =============================================================================================
// segment broadcast (backward of pooling)
// =============================================================================================
__global__ void seg_bcast_kernel(const float* __restrict__ dOut, int width, const int32_t* __restrict__ seg_of_edge,
                                 const int32_t* __restrict__ seg_ptr, int64_t total, float scale, int mean_mode,
                                 float* __restrict__ dX) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int64_t e = i / width;
  const int j = (int)(i % width);
  const int t = seg_of_edge[e];
  float f = scale;
  if (mean_mode) f /= (float)(seg_ptr[t + 1] - seg_ptr[t]);
  dX[i] = f * dOut[(int64_t)t * width + j];
}

// =============================================================================================
// LayerNorm + ReLU
// =============================================================================================
struct LnArgs {
  const float* x; int64_t n_rows; int width; const float* gamma; const float* beta; float eps;
  float* y; float* mean; float* rstd;
};

template <int VEC, int LPR, int NV>
__global__ void __launch_bounds__(256) ln_relu_fwd_kernel(LnArgs p) {
  using V = VecT<VEC>;
  constexpr int RPW = 32 / LPR;
  const int lane = threadIdx.x & 31;
  const int lir = lane % LPR, grp = lane / LPR;
  const unsigned mask = group_mask<LPR>(lane);
  const int nvec = p.width / VEC;
  const int64_t row = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * RPW + grp;
  if (row >= p.n_rows) return;
  typename V::T x[NV];
  const float* src = p.x + row * p.width;
  float s = 0.f;
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    int vi = lir + LPR * v;
    x[v] = vi < nvec ? V::ld_stream(src + VEC * vi) : V::zero();
#pragma unroll
    for (int k = 0; k < VEC; ++k) s += get(x[v], k);
  }
  float* dst = p.y + row * p.width;
  if (p.gamma == nullptr) {
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      int vi = lir + LPR * v;
      if (vi < nvec) {
#pragma unroll
        for (int k = 0; k < VEC; ++k) get(x[v], k) = fmaxf(get(x[v], k), 0.f);
        V::st(dst + VEC * vi, x[v]);
      }
    }
    return;
  }
  const float mean = group_sum<LPR>(s, mask) / (float)p.width;
  float q = 0.f;
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    int vi = lir + LPR * v;
    if (vi < nvec) {
#pragma unroll
      for (int k = 0; k < VEC; ++k) { float d = get(x[v], k) - mean; q = fmaf(d, d, q); }
    }
  }
  const float var = group_sum<LPR>(q, mask) / (float)p.width;
  const float rstd = 1.f / sqrtf(var + p.eps);
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    int vi = lir + LPR * v;
    if (vi < nvec) {
      typename V::T g = V::ld(p.gamma + VEC * vi), b = V::ld(p.beta + VEC * vi);
#pragma unroll
      for (int k = 0; k < VEC; ++k)
        get(x[v], k) = fmaxf(fmaf((get(x[v], k) - mean) * rstd, get(g, k), get(b, k)), 0.f);
      V::st(dst + VEC * vi, x[v]);
    }
  }
  if (lir == 0) { p.mean[row] = mean; p.rstd[row] = rstd; }
}

struct LnBwdArgs {
  const float* dy; const float* x; const float* mean; const float* rstd; const float* gamma; const float* beta;
  const float* add;   // optional [n_rows, width]: dx += add (gradient of a residual branch that also reads x)
  int64_t n_rows; int width; float* dx; float* ws;   // ws: [blocks, 2*width] partial dgamma | dbeta
};

constexpr int kLnBwdThreads = 256;
// rows in flight per lane group (all their loads are issued before the first shuffle); wide rows already
// carry enough loads per lane, and their accumulators need the registers
constexpr int ln_bwd_unroll(int nv) { return nv <= 2 ? 2 : 1; }

// The ReLU mask is recomputed from x (same expression as the forward kernel, so the same sign) instead of
// reading y back: dy, x, [add] in, dx out.
template <int VEC, int LPR, int NV>
__global__ void __launch_bounds__(kLnBwdThreads, NV <= 2 ? 2 : 1) ln_relu_bwd_kernel(LnBwdArgs p) {
  using V = VecT<VEC>;
  constexpr int RPW = 32 / LPR;
  constexpr int NW = kLnBwdThreads / 32;
  constexpr int U = ln_bwd_unroll(NV);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int lir = lane % LPR, grp = lane / LPR;
  const unsigned mask = group_mask<LPR>(lane);
  const int nvec = p.width / VEC;
  const bool affine = p.gamma != nullptr;
  typename V::T dg[NV], db[NV], gam[NV], bet[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    int vi = lir + LPR * v;
    dg[v] = V::zero(); db[v] = V::zero();
    gam[v] = (affine && vi < nvec) ? V::ld(p.gamma + VEC * vi) : V::zero();
    bet[v] = (affine && vi < nvec) ? V::ld(p.beta + VEC * vi) : V::zero();
  }
  const int64_t rows_per_pass = (int64_t)gridDim.x * NW * RPW;
  const float invw = 1.f / (float)p.width;
  for (int64_t row0 = ((int64_t)blockIdx.x * NW + wid) * RPW + grp; row0 < p.n_rows; row0 += U * rows_per_pass) {
    typename V::T g[U][NV], xh[U][NV], extra[U][NV];
    float mean[U], rstd[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t row = row0 + u * rows_per_pass;
      const bool live = row < p.n_rows;
      mean[u] = 0.f; rstd[u] = 1.f;
      if (affine && live) { mean[u] = __ldg(p.mean + row); rstd[u] = __ldg(p.rstd + row); }
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        int vi = lir + LPR * v;
        const bool on = live && vi < nvec;
        g[u][v] = on ? V::ld_stream(p.dy + row * p.width + VEC * vi) : V::zero();
        xh[u][v] = on ? V::ld_stream(p.x + row * p.width + VEC * vi) : V::zero();
        extra[u][v] = (on && p.add) ? V::ld_stream(p.add + row * p.width + VEC * vi) : V::zero();
      }
    }
    float s1[U], s2[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      s1[u] = 0.f; s2[u] = 0.f;
#pragma unroll
      for (int v = 0; v < NV; ++v) {
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
          const float xraw = get(xh[u][v], k);
          const float xk = (xraw - mean[u]) * rstd[u];
          const float pre = affine ? fmaf(xk, get(gam[v], k), get(bet[v], k)) : xraw;
          const float gk = pre > 0.f ? get(g[u][v], k) : 0.f;
          get(dg[v], k) = fmaf(gk, xk, get(dg[v], k));
          get(db[v], k) += gk;
          const float gg = gk * get(gam[v], k);
          get(g[u][v], k) = affine ? gg : gk;
          get(xh[u][v], k) = xk;
          s1[u] += gg;
          s2[u] = fmaf(gg, xk, s2[u]);
        }
      }
    }
    if (affine) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        s1[u] = group_sum<LPR>(s1[u], mask) * invw;
        s2[u] = group_sum<LPR>(s2[u], mask) * invw;
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t row = row0 + u * rows_per_pass;
      if (row >= p.n_rows) continue;
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        int vi = lir + LPR * v;
        if (vi < nvec) {
#pragma unroll
          for (int k = 0; k < VEC; ++k) {
            float r = get(g[u][v], k);
            if (affine) r = rstd[u] * (r - s1[u] - get(xh[u][v], k) * s2[u]);
            get(g[u][v], k) = r + get(extra[u][v], k);
          }
          V::st(p.dx + row * p.width + VEC * vi, g[u][v]);
        }
      }
    }
  }
  if (p.gamma == nullptr) return;
  // column reduction of dgamma / dbeta: groups -> warp -> CTA -> workspace row
  __syncwarp();
  if (RPW > 1) {
#pragma unroll
    for (int off = LPR; off < 32; off <<= 1)
#pragma unroll
      for (int v = 0; v < NV; ++v)
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
          get(dg[v], k) += __shfl_xor_sync(0xffffffffu, get(dg[v], k), off);
          get(db[v], k) += __shfl_xor_sync(0xffffffffu, get(db[v], k), off);
        }
  }
  extern __shared__ float sm[];  // [NW][2*width]
  const int W2 = 2 * p.width;
  if (grp == 0) {
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      int vi = lir + LPR * v;
      if (vi < nvec) {
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
          sm[wid * W2 + VEC * vi + k] = get(dg[v], k);
          sm[wid * W2 + p.width + VEC * vi + k] = get(db[v], k);
        }
      }
    }
  }
  __syncthreads();
  for (int j = threadIdx.x; j < W2; j += kLnBwdThreads) {
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < NW; ++w) a += sm[w * W2 + j];
    p.ws[(int64_t)blockIdx.x * W2 + j] = a;
  }
}

__global__ void col_sum2_kernel(const float* __restrict__ ws, int rows, int width, float* __restrict__ out_a,
                                float* __restrict__ out_b) {
  // ws rows are [a(width) | b(width)]
  __shared__ float sm[8][33];
  const int j = blockIdx.x * 32 + threadIdx.x;
  const int W2 = 2 * width;
  float a = 0.f;
  if (j < W2)
    for (int r = threadIdx.y; r < rows; r += 8) a += ws[(int64_t)r * W2 + j];
  sm[threadIdx.y][threadIdx.x] = a;
  __syncthreads();
  if (threadIdx.y == 0 && j < W2) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += sm[w][threadIdx.x];
    if (j < width) out_a[j] = s; else out_b[j - width] = s;
  }
}

// =============================================================================================
// edge update: out = pscale * P + scale * (x0 @ W0^T + S[col] + V[row] + g) + skip
// =============================================================================================
struct EdgeUpdArgs {
  const float* P; int64_t ldp; const float* x0; int d0; const float* W0;
  const float* S; const float* V; const float* g; const float* skip; int64_t ldskip;
  const int32_t* row_idx; const int32_t* col_idx; int64_t n_obs; int width; float pscale; float scale; float* out;
};

template <int VEC>
__global__ void __launch_bounds__(256) edge_update_kernel(EdgeUpdArgs p) {
  using V = VecT<VEC>;
  const int nvec = p.width / VEC;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n_obs * nvec) return;
  const int64_t e = i / nvec;
  const int c = (int)(i % nvec) * VEC;
  typename V::T a = V::zero();
  if (p.S) {
    typename V::T s = V::ld(p.S + (int64_t)__ldg(p.col_idx + e) * p.width + c);
#pragma unroll
    for (int k = 0; k < VEC; ++k) get(a, k) += get(s, k);
  }
  if (p.V) {
    typename V::T s = V::ld(p.V + (int64_t)__ldg(p.row_idx + e) * p.width + c);
#pragma unroll
    for (int k = 0; k < VEC; ++k) get(a, k) += get(s, k);
  }
  if (p.g) {
    typename V::T s = V::ld(p.g + c);
#pragma unroll
    for (int k = 0; k < VEC; ++k) get(a, k) += get(s, k);
  }
  for (int q = 0; q < p.d0; ++q) {
    const float xv = __ldg(p.x0 + e * p.d0 + q);
#pragma unroll
    for (int k = 0; k < VEC; ++k) get(a, k) = fmaf(xv, __ldg(p.W0 + (int64_t)(c + k) * p.d0 + q), get(a, k));
  }
  {
    typename V::T pv = p.P ? V::ld_stream(p.P + e * p.ldp + c) : V::zero();
#pragma unroll
    for (int k = 0; k < VEC; ++k) get(a, k) = fmaf(p.pscale, get(pv, k), p.scale * get(a, k));
  }
  if (p.skip) {
    typename V::T s = V::ld_stream(p.skip + e * p.ldskip + c);
#pragma unroll
    for (int k = 0; k < VEC; ++k) get(a, k) += get(s, k);
  }
  V::st(p.out + e * p.width + c, a);
}

// Fast path (width % 4 == 0): LPR lanes per observation row, every lane owns NV float4 of the row; W0 columns
// and the global term live in registers, row / col ids are read once per row, two rows are in flight per group.
template <int LPR, int NV>
__global__ void __launch_bounds__(256, 2) edge_update_rows_kernel(EdgeUpdArgs p) {
  constexpr int RPW = 32 / LPR;
  constexpr int U = 2;
  const int lane = threadIdx.x & 31;
  const int lir = lane % LPR, grp = lane / LPR;
  const int nvec = p.width / 4;
  float4 gl[NV], w0[NV][4];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int vi = lir + LPR * v;
    gl[v] = (p.g && vi < nvec) ? ld4(p.g + 4 * vi) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      w0[v][q] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (q < p.d0 && vi < nvec) {
        const float* wp = p.W0 + (int64_t)(4 * vi) * p.d0 + q;
        w0[v][q] = make_float4(wp[0], wp[p.d0], wp[2 * p.d0], wp[3 * p.d0]);
      }
    }
  }
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t stride = ((int64_t)gridDim.x * blockDim.x >> 5) * RPW;
  for (int64_t e0 = warp * RPW + grp; e0 < p.n_obs; e0 += stride * U) {
    float4 pv[U][NV], sk[U][NV], sv[U][NV], vv[U][NV];
    float x0v[U][4];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t e = e0 + u * stride;
      const bool ok = e < p.n_obs;
      const int64_t ec = ok ? e : e0;
      const int col = p.S ? __ldg(p.col_idx + ec) : 0, row = p.V ? __ldg(p.row_idx + ec) : 0;
#pragma unroll
      for (int q = 0; q < 4; ++q) x0v[u][q] = (q < p.d0) ? __ldg(p.x0 + ec * p.d0 + q) : 0.f;
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const int vi = lir + LPR * v;
        const bool on = vi < nvec;
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        pv[u][v] = (on && p.P) ? ld_stream4(p.P + ec * p.ldp + 4 * vi) : z;
        sk[u][v] = (on && p.skip) ? ld_stream4(p.skip + ec * p.ldskip + 4 * vi) : z;
        sv[u][v] = (on && p.S) ? ld4(p.S + (int64_t)col * p.width + 4 * vi) : z;
        vv[u][v] = (on && p.V) ? ld4(p.V + (int64_t)row * p.width + 4 * vi) : z;
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t e = e0 + u * stride;
      if (e >= p.n_obs) continue;
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const int vi = lir + LPR * v;
        if (vi >= nvec) continue;
        float4 a;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float t = comp(sv[u][v], k) + comp(vv[u][v], k) + comp(gl[v], k);
#pragma unroll
          for (int q = 0; q < 4; ++q) t = fmaf(x0v[u][q], comp(w0[v][q], k), t);
          comp(a, k) = fmaf(p.pscale, comp(pv[u][v], k), p.scale * t) + comp(sk[u][v], k);
        }
        st_stream4(p.out + e * p.width + 4 * vi, a);
      }
    }
  }
}

}  // namespace gasfm

using namespace gasfm;

extern "C" size_t gasfm_seg_sum_ws_bytes(int max_chunks, int width) {
  return (size_t)max_chunks * width * sizeof(float);
}

extern "C" int gasfm_seg_sum(const float* X, int64_t ldx, int width, const int32_t* seg_ptr, const int32_t* perm,
                             int n_seg, int chunk, const int32_t* chunk_ptr, const int32_t* chunk_seg,
                             int max_chunks, float scale, int mean_mode, float* out, void* ws, void* stream) {
  GASFM_REQUIRE(width > 0 && width <= 4096, "seg_sum: unsupported width %d", width);
  if (n_seg <= 0) return 0;
  GASFM_REQUIRE(chunk == 0 || (ws && chunk_ptr && chunk_seg), "seg_sum: chunked plan needs workspace and chunk tables");
  cudaStream_t st = (cudaStream_t)stream;
  SegSumArgs a{X, ldx, width, seg_ptr, perm, n_seg, chunk, chunk_ptr, chunk_seg, max_chunks, scale, mean_mode, out, (float*)ws};
  const bool vec4 = width % 4 == 0 && ldx % 4 == 0 && ((uintptr_t)X | (uintptr_t)out | (uintptr_t)ws) % 16 == 0 && width <= 1024;
#define CALL_SEG(VEC, LPR, NV) launch_seg_sum<VEC, LPR, NV>(a, st)
  if (vec4) {
    GASFM_ROW_DISPATCH(4, width / 4, CALL_SEG);
  } else {
    GASFM_REQUIRE(width <= 256, "seg_sum: width %d needs 16-byte aligned rows", width);
    GASFM_ROW_DISPATCH(1, width, CALL_SEG);
  }
#undef CALL_SEG
  return check_launch("seg_sum");
}

extern "C" int gasfm_seg_bcast(const float* dOut, int width, const int32_t* seg_of_edge, const int32_t* seg_ptr,
                               int64_t n_obs, float scale, int mean_mode, float* dX, void* stream) {
  if (n_obs <= 0) return 0;
  const int64_t total = n_obs * width;
  seg_bcast_kernel<<<ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(dOut, width, seg_of_edge, seg_ptr, total,
                                                                          scale, mean_mode, dX);
  return check_launch("seg_bcast");
}

extern "C" int gasfm_ln_relu_fwd(const float* x, int64_t n_rows, int width, const float* gamma, const float* beta,
                                 float eps, float* y, float* mean, float* rstd, void* stream) {
  GASFM_REQUIRE(width > 0 && width <= 1024, "ln_relu_fwd: unsupported width %d", width);
  GASFM_REQUIRE((gamma == nullptr) == (beta == nullptr), "ln_relu_fwd: gamma and beta must both be given or both be NULL");
  if (n_rows <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  LnArgs a{x, n_rows, width, gamma, beta, eps, y, mean, rstd};
  const bool vec4 = width % 4 == 0 && ((uintptr_t)x | (uintptr_t)y | (uintptr_t)gamma | (uintptr_t)beta) % 16 == 0;
#define CALL_LN(VEC, LPR, NV)                                                      \
  ln_relu_fwd_kernel<VEC, LPR, NV><<<ceil_div((n_rows + (32 / LPR) - 1) / (32 / LPR), 8), 256, 0, st>>>(a)
  if (vec4) {
    GASFM_ROW_DISPATCH(4, width / 4, CALL_LN);
  } else {
    GASFM_REQUIRE(width <= 256, "ln_relu_fwd: width %d needs to be a multiple of 4", width);
    GASFM_ROW_DISPATCH(1, width, CALL_LN);
  }
#undef CALL_LN
  return check_launch("ln_relu_fwd");
}

static int ln_bwd_blocks(int64_t n_rows) {
  int64_t need = (n_rows + 63) / 64;
  int64_t cap = (int64_t)kNumSMs * 8;
  return (int)(need < 1 ? 1 : (need < cap ? need : cap));
}

extern "C" size_t gasfm_ln_relu_bwd_ws_bytes(int64_t n_rows, int width) {
  return (size_t)ln_bwd_blocks(n_rows) * 2 * width * sizeof(float);
}

extern "C" int gasfm_ln_relu_bwd(const float* dy, const float* x, const float* mean, const float* rstd, const float* gamma,
                                 const float* beta, const float* add, int64_t n_rows, int width, float* dx, float* dgamma,
                                 float* dbeta, void* ws, void* stream) {
  GASFM_REQUIRE(width > 0 && width <= 1024, "ln_relu_bwd: unsupported width %d", width);
  cudaStream_t st = (cudaStream_t)stream;
  if (n_rows <= 0) {
    if (gamma) {
      cudaMemsetAsync(dgamma, 0, width * sizeof(float), st);
      cudaMemsetAsync(dbeta, 0, width * sizeof(float), st);
    }
    return check_launch("ln_relu_bwd(empty)");
  }
  GASFM_REQUIRE(gamma == nullptr || (ws != nullptr && beta != nullptr && mean != nullptr && rstd != nullptr),
                "ln_relu_bwd: the affine form needs beta, mean, rstd and a workspace");
  const int blocks = ln_bwd_blocks(n_rows);
  LnBwdArgs a{dy, x, mean, rstd, gamma, beta, add, n_rows, width, dx, (float*)ws};
  const size_t smem = (size_t)(kLnBwdThreads / 32) * 2 * width * sizeof(float);
  const bool vec4 = width % 4 == 0 && ((uintptr_t)x | (uintptr_t)dy | (uintptr_t)dx | (uintptr_t)gamma | (uintptr_t)beta | (uintptr_t)add) % 16 == 0;
#define CALL_LNB(VEC, LPR, NV)                                                                       \
  do {                                                                                               \
    if (smem > 48 * 1024)                                                                            \
      cudaFuncSetAttribute(ln_relu_bwd_kernel<VEC, LPR, NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    ln_relu_bwd_kernel<VEC, LPR, NV><<<blocks, kLnBwdThreads, smem, st>>>(a);                        \
  } while (0)
  if (vec4) {
    GASFM_ROW_DISPATCH(4, width / 4, CALL_LNB);
  } else {
    GASFM_REQUIRE(width <= 256, "ln_relu_bwd: width %d needs to be a multiple of 4", width);
    GASFM_ROW_DISPATCH(1, width, CALL_LNB);
  }
#undef CALL_LNB
  int rc = check_launch("ln_relu_bwd");
  if (rc || gamma == nullptr) return rc;
  col_sum2_kernel<<<ceil_div(2 * width, 32), dim3(32, 8), 0, st>>>((const float*)ws, blocks, width, dgamma, dbeta);
  return check_launch("ln_relu_bwd(reduce)");
}

static int col_sum_slices(int64_t rows) {
  int64_t s = rows / 256;
  return (int)(s < 1 ? 1 : (s > 2 * kNumSMs ? 2 * kNumSMs : s));
}

namespace gasfm {
// Stage 1 of the column sum for 16-byte aligned rows: a warp reads whole rows (row-contiguous, 4 rows in flight),
// every lane keeps NV float4 column accumulators; the CTA's 8 warps are combined through shared memory and one
// partial row per CTA goes to the workspace.  (Reading 32-column strips of every row instead left most of each
// DRAM page unused: 86 us for a [50k, 256] matrix, cold.)
template <int NV>
__global__ void __launch_bounds__(256) col_sum_rows_kernel(const float* __restrict__ x, int64_t ld, int64_t rows, int width,
                                                           int64_t rows_per_cta, float* __restrict__ ws) {
  extern __shared__ float part[];                      // [8 warps][width]
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int nvec = width / 4;
  float4 acc[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
  const int64_t r_begin = (int64_t)blockIdx.x * rows_per_cta;
  const int64_t r_end = r_begin + rows_per_cta < rows ? r_begin + rows_per_cta : rows;
  for (int64_t r0 = r_begin + wid; r0 < r_end; r0 += 8 * 4) {
    float4 v4[4][NV];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t r = r0 + 8 * u;
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const int vi = lane + 32 * v;
        v4[u][v] = (r < r_end && vi < nvec) ? ld_stream4(x + r * ld + 4 * vi) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        acc[v].x += v4[u][v].x; acc[v].y += v4[u][v].y; acc[v].z += v4[u][v].z; acc[v].w += v4[u][v].w;
      }
  }
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int vi = lane + 32 * v;
    if (vi < nvec) *reinterpret_cast<float4*>(part + wid * width + 4 * vi) = acc[v];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < width; c += 256) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += part[w * width + c];
    ws[(int64_t)blockIdx.x * width + c] = t;
  }
}
}  // namespace gasfm

extern "C" size_t gasfm_col_sum_ws_bytes(int64_t rows, int width) {
  const int s = col_sum_slices(rows);
  return s > 1 ? (size_t)s * width * sizeof(float) : 0;
}

extern "C" int gasfm_col_sum(const float* x, int64_t ld, int64_t rows, int width, float* out, void* ws, void* stream) {
  GASFM_REQUIRE(width > 0 && ld >= width && rows < (int64_t)INT32_MAX, "col_sum: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  if (rows <= 0) {
    cudaMemsetAsync(out, 0, (size_t)width * sizeof(float), st);
    return check_launch("col_sum(empty)");
  }
  const int slices = col_sum_slices(rows);
  if (slices == 1) {
    launch_col_reduce(ColReduceJob{x, ld, width, out, 0, 0}, nullptr, (int)rows, 1.f, st);
    return check_launch("col_sum");
  }
  GASFM_REQUIRE(ws != nullptr, "col_sum: workspace required");
  const int64_t per_slice = (rows + slices - 1) / slices;
  const bool vec4 = width % 4 == 0 && ld % 4 == 0 && width <= 1024 && (uintptr_t)x % 16 == 0;
  if (vec4) {
    const size_t smem = (size_t)8 * width * sizeof(float);
    const int nv = (width / 4 + 31) / 32;
    if (nv <= 1) col_sum_rows_kernel<1><<<slices, 256, smem, st>>>(x, ld, rows, width, per_slice, (float*)ws);
    else if (nv <= 2) col_sum_rows_kernel<2><<<slices, 256, smem, st>>>(x, ld, rows, width, per_slice, (float*)ws);
    else if (nv <= 4) col_sum_rows_kernel<4><<<slices, 256, smem, st>>>(x, ld, rows, width, per_slice, (float*)ws);
    else col_sum_rows_kernel<8><<<slices, 256, smem, st>>>(x, ld, rows, width, per_slice, (float*)ws);
  } else {
    const int blocks = ceil_div(width, 32);
    const ColReduceJob stage1{x, ld, width, (float*)ws, 0, 0};
    col_reduce_kernel<<<dim3(blocks, slices), dim3(32, kColReduceGroups), 0, st>>>(stage1, stage1, blocks, (int)rows, (int)per_slice, width, 1.f);
  }
  launch_col_reduce(ColReduceJob{(const float*)ws, width, width, out, 0, 0}, nullptr, slices, 1.f, st);
  return check_launch("col_sum");
}

extern "C" int gasfm_edge_update_fwd(const float* P, int64_t ldp, const float* x0, int d0, const float* W0,
                                     const float* S, const float* V, const float* g, const float* skip, int64_t ldskip,
                                     const int32_t* row_idx, const int32_t* col_idx, int64_t n_obs, int width,
                                     float pscale, float scale, float* out, void* stream) {
  GASFM_REQUIRE(width > 0, "edge_update_fwd: bad width");
  GASFM_REQUIRE(d0 >= 0 && d0 <= 4, "edge_update_fwd: init-feature width %d > 4 is not fused", d0);
  GASFM_REQUIRE((!S || col_idx) && (!V || row_idx), "edge_update_fwd: gather needs the index arrays");
  GASFM_REQUIRE(P || d0 > 0, "edge_update_fwd: neither a projected term nor init features given");
  if (n_obs <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  EdgeUpdArgs a{P, ldp, d0 > 0 ? x0 : nullptr, d0, W0, S, V, g, skip, ldskip, row_idx, col_idx, n_obs, width, pscale, scale, out};
  const bool vec4 = width % 4 == 0 && ldp % 4 == 0 && (!skip || ldskip % 4 == 0) &&
                    ((uintptr_t)P | (uintptr_t)S | (uintptr_t)V | (uintptr_t)g | (uintptr_t)skip | (uintptr_t)out) % 16 == 0;
  if (vec4 && width <= 1024) {
    const int nvec = width / 4;
    int64_t cap = (int64_t)kNumSMs * 16;
#define CALL_EU(VEC, LPR, NV)                                                                          \
    do {                                                                                               \
      int64_t need = (n_obs + (32 / LPR) * 8 * 2 - 1) / ((32 / LPR) * 8 * 2);                          \
      edge_update_rows_kernel<LPR, NV><<<(int)(need < cap ? (need < 1 ? 1 : need) : cap), 256, 0, st>>>(a); \
    } while (0)
    GASFM_ROW_DISPATCH(4, nvec, CALL_EU);
#undef CALL_EU
  } else if (vec4) {
    edge_update_kernel<4><<<ceil_div(n_obs * (width / 4), 256), 256, 0, st>>>(a);
  } else {
    edge_update_kernel<1><<<ceil_div(n_obs * width, 256), 256, 0, st>>>(a);
  }
  return check_launch("edge_update_fwd");
}
