// Small dense helpers on the SIMT pipes (sm_100a), for shapes where a 128-wide tensor-core tile would be
// mostly padding:
//   wgrad_small   dW[Nout,Kout] = dY[E,Nout]^T X[E,Kout] for Nout, Kout in {32, 64} (the shipped d = 32 widths)
//   x0_bwd        the rank-d0 (d0 <= 4) part of the observation update's backward:
//                 dx0[E,d0] = scale * dOut[E,W] W0[W,d0],  dW0[W,d0] = scale * dOut^T x0
// Both are single passes over E-sized inputs (HBM-bound) with fp32 round-to-nearest accumulation and a
// deterministic two-stage reduction (per-CTA partials -> column sum).
#include "common.cuh"
#include "col_reduce.cuh"
#include "../../include/gasfm_b200.h"

namespace gasfm {

constexpr int kWsBlocks = kNumSMs * 2;
constexpr int kWsWarps = 8;

// One warp owns a [Nout x Kout] accumulator: lane l holds rows l, l+32 (RPL = Nout/32) x all Kout columns.
// Rows are streamed through a per-warp shared tile; x values are read back as broadcast float4.
template <int NOUT, int KOUT>
__global__ void __launch_bounds__(kWsWarps * 32) wgrad_small_kernel(const float* __restrict__ dY, int64_t lddy,
                                                                    const float* __restrict__ X, int64_t ldx, int64_t E,
                                                                    float* __restrict__ ws, float* __restrict__ ws_db) {
  constexpr int RPL = NOUT / 32;
  constexpr int TR = 8;                                    // rows per tile
  // one static buffer: per-warp row tiles during the main loop, the CTA accumulator afterwards
  __shared__ __align__(16) float s_buf[kWsWarps * TR * (NOUT + KOUT)];
  static_assert(kWsWarps * TR * (NOUT + KOUT) >= NOUT * KOUT, "accumulator must fit in the tile buffer");
  float (*s_dy)[TR][NOUT] = reinterpret_cast<float (*)[TR][NOUT]>(s_buf);
  float (*s_x)[TR][KOUT] = reinterpret_cast<float (*)[TR][KOUT]>(s_buf + kWsWarps * TR * NOUT);
  float* s_acc = s_buf;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  float acc[RPL][KOUT], dsum[RPL];
#pragma unroll
  for (int r = 0; r < RPL; ++r) {
    dsum[r] = 0.f;
#pragma unroll
    for (int k = 0; k < KOUT; ++k) acc[r][k] = 0.f;
  }
  const int64_t n_tiles = (E + TR - 1) / TR;
  const int64_t warp = (int64_t)blockIdx.x * kWsWarps + wid, n_warps = (int64_t)gridDim.x * kWsWarps;
  for (int64_t t = warp; t < n_tiles; t += n_warps) {
    const int64_t e0 = t * TR;
    // cooperative, coalesced tile load (float4 per lane)
#pragma unroll
    for (int i = lane; i < TR * NOUT / 4; i += 32) {
      const int r = i / (NOUT / 4), c = (i % (NOUT / 4)) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (e0 + r < E) v = ld_stream4(dY + (e0 + r) * lddy + c);
      *reinterpret_cast<float4*>(&s_dy[wid][r][c]) = v;
    }
#pragma unroll
    for (int i = lane; i < TR * KOUT / 4; i += 32) {
      const int r = i / (KOUT / 4), c = (i % (KOUT / 4)) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (e0 + r < E) v = ld_stream4(X + (e0 + r) * ldx + c);
      *reinterpret_cast<float4*>(&s_x[wid][r][c]) = v;
    }
    __syncwarp();
#pragma unroll 4
    for (int r = 0; r < TR; ++r) {
      float d[RPL];
#pragma unroll
      for (int q = 0; q < RPL; ++q) { d[q] = s_dy[wid][r][lane + 32 * q]; dsum[q] += d[q]; }
#pragma unroll
      for (int k = 0; k < KOUT; k += 4) {
        const float4 xv = *reinterpret_cast<const float4*>(&s_x[wid][r][k]);   // same address in all lanes: broadcast
#pragma unroll
        for (int q = 0; q < RPL; ++q) {
          acc[q][k] = fmaf(d[q], xv.x, acc[q][k]);
          acc[q][k + 1] = fmaf(d[q], xv.y, acc[q][k + 1]);
          acc[q][k + 2] = fmaf(d[q], xv.z, acc[q][k + 2]);
          acc[q][k + 3] = fmaf(d[q], xv.w, acc[q][k + 3]);
        }
      }
    }
    __syncwarp();
  }
  // deterministic CTA reduction: warps add their accumulators one after the other
  __syncthreads();                                         // every warp is done with its tiles: reuse the buffer
  for (int j = threadIdx.x; j < NOUT * KOUT; j += blockDim.x) s_acc[j] = 0.f;
  __syncthreads();
  for (int w = 0; w < kWsWarps; ++w) {
    if (wid == w) {
#pragma unroll
      for (int q = 0; q < RPL; ++q)
#pragma unroll
        for (int k = 0; k < KOUT; ++k) s_acc[(lane + 32 * q) * KOUT + k] += acc[q][k];
    }
    __syncthreads();
  }
  for (int j = threadIdx.x; j < NOUT * KOUT; j += blockDim.x) ws[(int64_t)blockIdx.x * NOUT * KOUT + j] = s_acc[j];
  if (ws_db != nullptr) {                                   // bias gradient = column sums of dY, same deterministic scheme
    __syncthreads();
    for (int j = threadIdx.x; j < NOUT; j += blockDim.x) s_acc[j] = 0.f;
    __syncthreads();
    for (int w = 0; w < kWsWarps; ++w) {
      if (wid == w) {
#pragma unroll
        for (int q = 0; q < RPL; ++q) s_acc[lane + 32 * q] += dsum[q];
      }
      __syncthreads();
    }
    for (int j = threadIdx.x; j < NOUT; j += blockDim.x) ws_db[(int64_t)blockIdx.x * NOUT + j] = s_acc[j];
  }
}

// ---- x0 backward ------------------------------------------------------------------------------------
constexpr int kX0Threads = 256;
template <int LPR, int NV, int D0>
__global__ void __launch_bounds__(kX0Threads) x0_bwd_kernel(const float* __restrict__ dOut, int64_t E, int width,
                                                            const float* __restrict__ x0, const float* __restrict__ W0,
                                                            float scale, float* __restrict__ dx0, float* __restrict__ ws,
                                                            float* __restrict__ rowmax /* optional [E]: max |dOut[row, :]| */) {
  constexpr int RPW = 32 / LPR;
  constexpr int NW = kX0Threads / 32;
  constexpr int U = 2;                                        // rows in flight per lane group
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int lir = lane % LPR, grp = lane / LPR;
  const unsigned mask = group_mask<LPR>(lane);
  const int nvec = width / 4;
  // W0 rows of this lane's channels, and the dW0 accumulators: [NV][4 channels][D0]
  float w[NV][4][D0], dw[NV][4][D0];
#pragma unroll
  for (int v = 0; v < NV; ++v)
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int q = 0; q < D0; ++q) {
        const int c = 4 * (lir + LPR * v) + k;
        w[v][k][q] = (lir + LPR * v < nvec) ? W0[(int64_t)c * D0 + q] : 0.f;
        dw[v][k][q] = 0.f;
      }
  const int64_t stride = (int64_t)gridDim.x * NW * RPW;
  for (int64_t row0 = ((int64_t)blockIdx.x * NW + wid) * RPW + grp; row0 < E; row0 += stride * U) {
    float4 g[U][NV];
    float xr[U][D0];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t row = row0 + u * stride;
      const bool ok = row < E;
#pragma unroll
      for (int q = 0; q < D0; ++q) xr[u][q] = ok ? __ldg(x0 + row * D0 + q) : 0.f;
#pragma unroll
      for (int v = 0; v < NV; ++v)
        g[u][v] = (ok && lir + LPR * v < nvec) ? ld_stream4(dOut + row * width + 4 * (lir + LPR * v)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float s[D0];
#pragma unroll
      for (int q = 0; q < D0; ++q) s[q] = 0.f;
#pragma unroll
      for (int v = 0; v < NV; ++v)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float gk = comp(g[u][v], k);
#pragma unroll
          for (int q = 0; q < D0; ++q) {
            s[q] = fmaf(gk, w[v][k][q], s[q]);
            dw[v][k][q] = fmaf(gk, xr[u][q], dw[v][k][q]);
          }
        }
#pragma unroll
      for (int q = 0; q < D0; ++q) {
#pragma unroll
        for (int off = LPR / 2; off > 0; off >>= 1) s[q] += __shfl_xor_sync(mask, s[q], off);
      }
      const int64_t row = row0 + u * stride;
      if (rowmax != nullptr) {                                  // row maximum of |dOut|: the fp16 input-gradient GEMM's row scale
        float m = 0.f;
#pragma unroll
        for (int v = 0; v < NV; ++v)
          m = fmaxf(m, fmaxf(fmaxf(fabsf(g[u][v].x), fabsf(g[u][v].y)), fmaxf(fabsf(g[u][v].z), fabsf(g[u][v].w))));
#pragma unroll
        for (int off = LPR / 2; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(mask, m, off));
        if (lir == 0 && row < E) rowmax[row] = m;
      }
      if (lir == 0 && row < E) {
#pragma unroll
        for (int q = 0; q < D0; ++q) dx0[row * D0 + q] = scale * s[q];
      }
    }
  }
  // dW0: groups -> warp -> CTA -> workspace row [width * 4]
  __syncwarp();
  if (RPW > 1) {
#pragma unroll
    for (int off = LPR; off < 32; off <<= 1)
#pragma unroll
      for (int v = 0; v < NV; ++v)
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
          for (int q = 0; q < D0; ++q) dw[v][k][q] += __shfl_xor_sync(0xffffffffu, dw[v][k][q], off);
  }
  extern __shared__ float sm[];   // [NW][width*4]
  const int W4 = width * 4;
  for (int j = threadIdx.x; j < NW * W4; j += kX0Threads) sm[j] = 0.f;
  __syncthreads();
  if (grp == 0) {
#pragma unroll
    for (int v = 0; v < NV; ++v)
      if (lir + LPR * v < nvec) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
          for (int q = 0; q < D0; ++q) sm[wid * W4 + (4 * (lir + LPR * v) + k) * 4 + q] = dw[v][k][q];
      }
  }
  __syncthreads();
  for (int j = threadIdx.x; j < W4; j += kX0Threads) {
    float a = 0.f;
#pragma unroll
    for (int ww = 0; ww < NW; ++ww) a += sm[ww * W4 + j];
    ws[(int64_t)blockIdx.x * W4 + j] = a;
  }
}

// ---------------------------------------------------------------------------------------------
// The observation update's backward over dOut in ONE storage-order pass (round 2): the per-view segment sum
// dV[t] = scale * sum_{e in view t} dOut[e] (gradient of lin_view(..)[row], models/layers.py:941-945) together with the
// rank-d0 term's dx0 / dW0 and the row maxima of dOut -- x0_bwd_kernel and the chunked seg_sum kernel each made their
// own pass over the same [E, width] matrix.  One warp per chunk of the view plan (<= chunk contiguous rows of one view).
// ---------------------------------------------------------------------------------------------
struct UpdBwdArgs {
  const float* dOut; int64_t E; int width; const float* x0; const float* W0; float scale;
  const int32_t* seg_ptr; int n_seg; int chunk; const int32_t* chunk_ptr; const int32_t* chunk_seg;
  float* dV; float* ws_v; float* dx0; float* ws_w; float* rowmax;
};

template <int LPR, int NV, int D0>
__global__ void __launch_bounds__(kX0Threads) update_bwd_views_kernel(UpdBwdArgs p) {
  constexpr int RPW = 32 / LPR;
  constexpr int NW = kX0Threads / 32;
  constexpr int U = 2;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int lir = lane % LPR, grp = lane / LPR;
  const unsigned mask = group_mask<LPR>(lane);
  const int width = p.width, nvec = width / 4;
  float w[NV][4][D0], dw[NV][4][D0];
  float4 acc[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int q = 0; q < D0; ++q) {
        const int c = 4 * (lir + LPR * v) + k;
        w[v][k][q] = (lir + LPR * v < nvec) ? p.W0[(int64_t)c * D0 + q] : 0.f;
        dw[v][k][q] = 0.f;
      }
  }
  const int total = __ldg(p.chunk_ptr + p.n_seg);
  const int kc = blockIdx.x * NW + wid;                       // this warp's chunk
  if (kc < total) {
    const int t = __ldg(p.chunk_seg + kc);
    const int c0 = __ldg(p.chunk_ptr + t), c1 = __ldg(p.chunk_ptr + t + 1);
    const int sb = __ldg(p.seg_ptr + t), se = __ldg(p.seg_ptr + t + 1);
    const int b = sb + (kc - c0) * p.chunk, e = min(b + p.chunk, se);
    for (int row0 = b + grp; row0 < e; row0 += RPW * U) {
      float4 g[U][NV];
      float xr[U][D0];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int row = row0 + u * RPW;
        const bool ok = row < e;
#pragma unroll
        for (int q = 0; q < D0; ++q) xr[u][q] = ok ? __ldg(p.x0 + (int64_t)row * D0 + q) : 0.f;
#pragma unroll
        for (int v = 0; v < NV; ++v)
          g[u][v] = (ok && lir + LPR * v < nvec) ? ld_stream4(p.dOut + (int64_t)row * width + 4 * (lir + LPR * v)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        float s[D0];
#pragma unroll
        for (int q = 0; q < D0; ++q) s[q] = 0.f;
        float m = 0.f;
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          acc[v].x += g[u][v].x; acc[v].y += g[u][v].y; acc[v].z += g[u][v].z; acc[v].w += g[u][v].w;
          m = fmaxf(m, fmaxf(fmaxf(fabsf(g[u][v].x), fabsf(g[u][v].y)), fmaxf(fabsf(g[u][v].z), fabsf(g[u][v].w))));
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float gk = comp(g[u][v], k);
#pragma unroll
            for (int q = 0; q < D0; ++q) {
              s[q] = fmaf(gk, w[v][k][q], s[q]);
              dw[v][k][q] = fmaf(gk, xr[u][q], dw[v][k][q]);
            }
          }
        }
#pragma unroll
        for (int q = 0; q < D0; ++q) {
#pragma unroll
          for (int off = LPR / 2; off > 0; off >>= 1) s[q] += __shfl_xor_sync(mask, s[q], off);
        }
#pragma unroll
        for (int off = LPR / 2; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(mask, m, off));
        const int row = row0 + u * RPW;
        if (lir == 0 && row < e) {
#pragma unroll
          for (int q = 0; q < D0; ++q) p.dx0[(int64_t)row * D0 + q] = p.scale * s[q];
          if (p.rowmax != nullptr) p.rowmax[row] = m;
        }
      }
    }
    // view sum of the chunk: groups -> lane group 0 -> dV (single-chunk views) or the chunk partials
    __syncwarp();
    if (RPW > 1) {
#pragma unroll
      for (int off = LPR; off < 32; off <<= 1)
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          acc[v].x += __shfl_xor_sync(0xffffffffu, acc[v].x, off); acc[v].y += __shfl_xor_sync(0xffffffffu, acc[v].y, off);
          acc[v].z += __shfl_xor_sync(0xffffffffu, acc[v].z, off); acc[v].w += __shfl_xor_sync(0xffffffffu, acc[v].w, off);
        }
    }
    if (grp == 0) {
      const bool single = (c1 - c0) == 1;
      const float f = single ? p.scale : 1.f;
      float* dst = single ? p.dV + (int64_t)t * width : p.ws_v + (int64_t)kc * width;
#pragma unroll
      for (int v = 0; v < NV; ++v)
        if (lir + LPR * v < nvec)
          st4(dst + 4 * (lir + LPR * v), make_float4(acc[v].x * f, acc[v].y * f, acc[v].z * f, acc[v].w * f));
    }
  }
  // dW0: groups -> warp -> CTA -> workspace row [width * 4] (as x0_bwd_kernel)
  __syncwarp();
  if (RPW > 1) {
#pragma unroll
    for (int off = LPR; off < 32; off <<= 1)
#pragma unroll
      for (int v = 0; v < NV; ++v)
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
          for (int q = 0; q < D0; ++q) dw[v][k][q] += __shfl_xor_sync(0xffffffffu, dw[v][k][q], off);
  }
  extern __shared__ float sm[];   // [NW][width*4]
  const int W4 = width * 4;
  for (int j = threadIdx.x; j < NW * W4; j += kX0Threads) sm[j] = 0.f;
  __syncthreads();
  if (grp == 0) {
#pragma unroll
    for (int v = 0; v < NV; ++v)
      if (lir + LPR * v < nvec) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
          for (int q = 0; q < D0; ++q) sm[wid * W4 + (4 * (lir + LPR * v) + k) * 4 + q] = dw[v][k][q];
      }
  }
  __syncthreads();
  for (int j = threadIdx.x; j < W4; j += kX0Threads) {
    float a = 0.f;
#pragma unroll
    for (int ww = 0; ww < NW; ++ww) a += sm[ww * W4 + j];
    p.ws_w[(int64_t)blockIdx.x * W4 + j] = a;
  }
}

// dV[t] = scale * sum of the chunk partials of views that span several chunks; views without observations get 0
__global__ void __launch_bounds__(256) update_bwd_views_merge_kernel(UpdBwdArgs p) {
  __shared__ float sm[8][33];
  const int t = blockIdx.x;
  const int c0 = __ldg(p.chunk_ptr + t), c1 = __ldg(p.chunk_ptr + t + 1);
  if (c1 - c0 == 1) return;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int j = blockIdx.y * 32 + lane;
  float a = 0.f;
  if (j < p.width)
    for (int k = c0 + wid; k < c1; k += 8) a += p.ws_v[(int64_t)k * p.width + j];
  sm[wid][lane] = a;
  __syncthreads();
  if (wid == 0 && j < p.width) {
    float r = 0.f;
#pragma unroll
    for (int ww = 0; ww < 8; ++ww) r += sm[ww][lane];
    p.dV[(int64_t)t * p.width + j] = r * p.scale;
  }
}

static int x0_blocks(int64_t E) {
  int64_t need = (E + 63) / 64;
  int64_t cap = (int64_t)kNumSMs * 8;
  return (int)(need < 1 ? 1 : (need < cap ? need : cap));
}

}  // namespace gasfm

using namespace gasfm;

extern "C" int gasfm_wgrad_small_supported(int Nout, int Kout, int64_t lddy, int64_t ldx) {
  return ((Nout == 32 || Nout == 64) && (Kout == 32 || Kout == 64) && lddy % 4 == 0 && ldx % 4 == 0) ? 1 : 0;
}
extern "C" size_t gasfm_wgrad_small_ws_bytes(int Nout, int Kout) { return (size_t)kWsBlocks * (Nout * Kout + Nout) * sizeof(float); }

extern "C" int gasfm_wgrad_small(const float* dY, int64_t lddy, const float* X, int64_t ldx, int64_t E, int Nout, int Kout,
                                 float* dW, float* dbias, void* ws, void* stream) {
  GASFM_REQUIRE(gasfm_wgrad_small_supported(Nout, Kout, lddy, ldx), "wgrad_small: unsupported shape %d x %d", Nout, Kout);
  GASFM_REQUIRE(ws && ((uintptr_t)dY | (uintptr_t)X) % 16 == 0, "wgrad_small: bad pointers");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t tiles = (E + 7) / 8;
  int blocks = (int)((tiles + kWsWarps - 1) / kWsWarps);
  if (blocks > kWsBlocks) blocks = kWsBlocks;
  if (blocks < 1) blocks = 1;
  float* w = (float*)ws;
  float* wdb = dbias ? w + (size_t)kWsBlocks * Nout * Kout : nullptr;
  if (Nout == 32 && Kout == 32) wgrad_small_kernel<32, 32><<<blocks, kWsWarps * 32, 0, st>>>(dY, lddy, X, ldx, E, w, wdb);
  else if (Nout == 32 && Kout == 64) wgrad_small_kernel<32, 64><<<blocks, kWsWarps * 32, 0, st>>>(dY, lddy, X, ldx, E, w, wdb);
  else if (Nout == 64 && Kout == 32) wgrad_small_kernel<64, 32><<<blocks, kWsWarps * 32, 0, st>>>(dY, lddy, X, ldx, E, w, wdb);
  else wgrad_small_kernel<64, 64><<<blocks, kWsWarps * 32, 0, st>>>(dY, lddy, X, ldx, E, w, wdb);
  int rc = check_launch("wgrad_small");
  if (rc) return rc;
  const int64_t width = (int64_t)Nout * Kout;
  const ColReduceJob jw{w, width, width, dW, 0, 0}, jb{wdb, Nout, Nout, dbias, 0, 0};
  launch_col_reduce(jw, dbias ? &jb : nullptr, blocks, 1.f, st);
  return check_launch("wgrad_small(reduce)");
}

extern "C" size_t gasfm_x0_bwd_ws_bytes(int64_t E, int width) { return (size_t)x0_blocks(E) * width * 4 * sizeof(float); }

static int x0_bwd_impl(const float* dOut, int64_t E, int width, const float* x0, const float* W0, int d0, float scale,
                       float* dx0, float* dW0, void* ws, float* rowmax, void* stream);

extern "C" int gasfm_x0_bwd(const float* dOut, int64_t E, int width, const float* x0, const float* W0, int d0, float scale,
                            float* dx0, float* dW0, void* ws, void* stream) {
  return x0_bwd_impl(dOut, E, width, x0, W0, d0, scale, dx0, dW0, ws, nullptr, stream);
}

extern "C" int gasfm_x0_bwd_rowmax(const float* dOut, int64_t E, int width, const float* x0, const float* W0, int d0, float scale,
                                   float* dx0, float* dW0, void* ws, float* rowmax, void* stream) {
  return x0_bwd_impl(dOut, E, width, x0, W0, d0, scale, dx0, dW0, ws, rowmax, stream);
}

static int x0_bwd_impl(const float* dOut, int64_t E, int width, const float* x0, const float* W0, int d0, float scale,
                       float* dx0, float* dW0, void* ws, float* rowmax, void* stream) {
  GASFM_REQUIRE(width > 0 && width % 4 == 0 && width <= 1024, "x0_bwd: width %d must be a multiple of 4 and <= 1024", width);
  GASFM_REQUIRE(d0 >= 1 && d0 <= 4, "x0_bwd: d0 = %d not in 1..4", d0);
  GASFM_REQUIRE(ws && (uintptr_t)dOut % 16 == 0, "x0_bwd: bad pointers");
  cudaStream_t st = (cudaStream_t)stream;
  if (E <= 0) {
    cudaMemsetAsync(dW0, 0, (size_t)width * d0 * sizeof(float), st);
    return check_launch("x0_bwd(empty)");
  }
  const int blocks = x0_blocks(E);
  const size_t smem = (size_t)(kX0Threads / 32) * width * 4 * sizeof(float);
  const int nvec = width / 4;
#define CALL_X0D(LPR, NV, D0)                                                                                       \
  do {                                                                                                              \
    if (smem > 48 * 1024)                                                                                           \
      cudaFuncSetAttribute(x0_bwd_kernel<LPR, NV, D0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);     \
    x0_bwd_kernel<LPR, NV, D0><<<blocks, kX0Threads, smem, st>>>(dOut, E, width, x0, W0, scale, dx0, (float*)ws, rowmax);   \
  } while (0)
#define CALL_X0(VEC, LPR, NV)                          \
  do {                                                 \
    switch (d0) {                                      \
      case 1: CALL_X0D(LPR, NV, 1); break;             \
      case 2: CALL_X0D(LPR, NV, 2); break;             \
      case 3: CALL_X0D(LPR, NV, 3); break;             \
      default: CALL_X0D(LPR, NV, 4); break;            \
    }                                                  \
  } while (0)
  if (nvec <= 1) CALL_X0(4, 1, 1);
  else if (nvec <= 2) CALL_X0(4, 2, 1);
  else if (nvec <= 4) CALL_X0(4, 4, 1);
  else if (nvec <= 8) CALL_X0(4, 8, 1);
  else if (nvec <= 16) CALL_X0(4, 16, 1);
  else if (nvec <= 32) CALL_X0(4, 32, 1);
  else if (nvec <= 64) CALL_X0(4, 32, 2);
  else if (nvec <= 128) CALL_X0(4, 32, 4);
  else CALL_X0(4, 32, 8);
#undef CALL_X0D
#undef CALL_X0
  int rc = check_launch("x0_bwd");
  if (rc) return rc;
  // dW0[c, q] = scale * sum_blocks ws[b][c*4 + q], q < d0
  const ColReduceJob jw{(const float*)ws, (int64_t)width * 4, (int64_t)width * 4, dW0, 4, d0};
  launch_col_reduce(jw, nullptr, blocks, scale, st);
  return check_launch("x0_bwd(reduce)");
}

extern "C" size_t gasfm_update_bwd_views_ws_bytes(int max_chunks, int width) {
  const size_t blocks = ((size_t)max_chunks + kX0Threads / 32 - 1) / (kX0Threads / 32);
  return ((size_t)max_chunks * width + blocks * width * 4) * sizeof(float);
}

extern "C" int gasfm_update_bwd_views(const float* dOut, int64_t E, int width, const float* x0, const float* W0, int d0, float scale,
                                      const int32_t* seg_ptr, int n_seg, int chunk, const int32_t* chunk_ptr,
                                      const int32_t* chunk_seg, int max_chunks, float* dV, float* dx0, float* dW0,
                                      float* rowmax, void* ws, void* stream) {
  GASFM_REQUIRE(width > 0 && width % 4 == 0 && width <= 1024, "update_bwd_views: width %d must be a multiple of 4 and <= 1024", width);
  GASFM_REQUIRE(d0 >= 1 && d0 <= 4, "update_bwd_views: d0 = %d not in 1..4", d0);
  GASFM_REQUIRE(chunk > 0 && chunk_ptr && chunk_seg && seg_ptr && n_seg > 0 && max_chunks > 0, "update_bwd_views: needs a chunked view plan");
  GASFM_REQUIRE(ws && dV && dx0 && dW0 && ((uintptr_t)dOut | (uintptr_t)dV | (uintptr_t)ws) % 16 == 0, "update_bwd_views: bad pointers");
  cudaStream_t st = (cudaStream_t)stream;
  const int NW = kX0Threads / 32;
  const int blocks = (max_chunks + NW - 1) / NW;
  UpdBwdArgs a{dOut, E, width, x0, W0, scale, seg_ptr, n_seg, chunk, chunk_ptr, chunk_seg, dV, (float*)ws, dx0,
               (float*)ws + (size_t)max_chunks * width, rowmax};
  const int nvec = width / 4;
  const size_t smem = (size_t)NW * width * 4 * sizeof(float);
#define CALL_UBD(LPR, NV, D0)                                                                                          \
  do {                                                                                                                 \
    if (smem > 48 * 1024)                                                                                              \
      cudaFuncSetAttribute(update_bwd_views_kernel<LPR, NV, D0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    update_bwd_views_kernel<LPR, NV, D0><<<blocks, kX0Threads, smem, st>>>(a);                                   \
  } while (0)
#define CALL_UB(LPR, NV)                               \
  do {                                                 \
    switch (d0) {                                      \
      case 1: CALL_UBD(LPR, NV, 1); break;             \
      case 2: CALL_UBD(LPR, NV, 2); break;             \
      case 3: CALL_UBD(LPR, NV, 3); break;             \
      default: CALL_UBD(LPR, NV, 4); break;            \
    }                                                  \
  } while (0)
  if (nvec <= 1) CALL_UB(1, 1);
  else if (nvec <= 2) CALL_UB(2, 1);
  else if (nvec <= 4) CALL_UB(4, 1);
  else if (nvec <= 8) CALL_UB(8, 1);
  else if (nvec <= 16) CALL_UB(16, 1);
  else if (nvec <= 32) CALL_UB(32, 1);
  else if (nvec <= 64) CALL_UB(32, 2);
  else if (nvec <= 128) CALL_UB(32, 4);
  else CALL_UB(32, 8);
#undef CALL_UBD
#undef CALL_UB
  int rc = check_launch("update_bwd_views");
  if (rc) return rc;
  update_bwd_views_merge_kernel<<<dim3(n_seg, (width + 31) / 32), 256, 0, st>>>(a);
  rc = check_launch("update_bwd_views(merge)");
  if (rc) return rc;
  const ColReduceJob jw{(const float*)a.ws_w, (int64_t)width * 4, (int64_t)width * 4, dW0, 4, d0};
  launch_col_reduce(jw, nullptr, blocks, scale, st);
  return check_launch("update_bwd_views(reduce)");
}
