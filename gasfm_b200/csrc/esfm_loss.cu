// Sparse, on-device ESFM reprojection loss.
//
// The reference (code/loss_functions.py:69-123) forms the dense [m,3,n] tensor Ps @ pts3D, dense masks and
// the dense normalised measurement matrix, although only the E observed (view, point) pairs enter the loss
// (``[data.valid_pts].mean()``).  Here one thread handles one observation e = (i, j):
//     p = P_i X_j;  ok = p_z >= margin (hinge) | |p_z| >= margin;
//     term = ok ? || p_xy / p_z - u_e || : (margin - p_z) * hinge_weight;      loss = mean_e term
// and the backward applies the reference's gradient hook on d loss / d p per observation
// (F.normalize(grad, dim=1) / max(1, #ok) where ok; unchanged elsewhere -- loss_functions.py:101-110),
// then emits the per-observation contributions to dPs (3x4 outer product) and dpts3D (P_i^T g), which
// the segment-sum kernels reduce over views / tracks.  No dense [m,n] tensor is ever built.
#include "common.cuh"
#include "../../include/gasfm_b200.h"

namespace gasfm {

struct EsfmArgs {
  const float* Ps; const float* pts; int64_t n; const float* obs; const int32_t* row_idx; const int32_t* col_idx;
  int64_t E; float margin; int hinge; float hinge_weight;
};

__device__ __forceinline__ void project(const EsfmArgs& a, int64_t e, float (&P)[12], float (&X)[4], float (&p)[3]) {
  const int i = __ldg(a.row_idx + e), j = __ldg(a.col_idx + e);
#pragma unroll
  for (int k = 0; k < 12; ++k) P[k] = __ldg(a.Ps + (int64_t)i * 12 + k);
#pragma unroll
  for (int k = 0; k < 4; ++k) X[k] = __ldg(a.pts + (int64_t)k * a.n + j);
#pragma unroll
  for (int r = 0; r < 3; ++r) p[r] = P[4 * r] * X[0] + P[4 * r + 1] * X[1] + P[4 * r + 2] * X[2] + P[4 * r + 3] * X[3];
}

__global__ void __launch_bounds__(256) esfm_fwd_kernel(EsfmArgs a, float* __restrict__ partial) {
  __shared__ float sm[8][2];
  float sum = 0.f, cnt = 0.f;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < a.E; e += (int64_t)gridDim.x * blockDim.x) {
    float P[12], X[4], p[3];
    project(a, e, P, X, p);
    const bool ok = a.hinge ? (p[2] >= a.margin) : (fabsf(p[2]) >= a.margin);
    if (ok) {
      const float rx = p[0] / p[2] - __ldg(a.obs + 2 * e), ry = p[1] / p[2] - __ldg(a.obs + 2 * e + 1);
      sum += sqrtf(rx * rx + ry * ry);
      cnt += 1.f;
    } else {
      sum += (a.margin - p[2]) * a.hinge_weight;
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    sum += __shfl_xor_sync(0xffffffffu, sum, off);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, off);
  }
  if ((threadIdx.x & 31) == 0) { sm[threadIdx.x >> 5][0] = sum; sm[threadIdx.x >> 5][1] = cnt; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f, c = 0.f;
    for (int w = 0; w < 8; ++w) { s += sm[w][0]; c += sm[w][1]; }
    partial[2 * blockIdx.x] = s;
    partial[2 * blockIdx.x + 1] = c;
  }
}

// out[0] = mean loss, out[1] = number of observations with a valid depth
__global__ void esfm_finish_kernel(const float* __restrict__ partial, int blocks, int64_t E, float* __restrict__ out) {
  double s = 0.0, c = 0.0;
  for (int b = 0; b < blocks; ++b) { s += partial[2 * b]; c += partial[2 * b + 1]; }
  out[0] = (float)(s / (double)(E > 0 ? E : 1));
  out[1] = (float)c;
}

// G[e, 0..11] = g (x) X  (contribution to dPs[i]),  G[e, 12..15] = P_i^T g  (contribution to dpts3D[:, j])
__global__ void __launch_bounds__(256) esfm_bwd_kernel(EsfmArgs a, const float* __restrict__ upstream, const float* __restrict__ stats,
                                                       int grad_mode, float* __restrict__ G) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= a.E) return;
  const float c = __ldg(upstream);
  const float inv_e = 1.f / (float)a.E;
  const float n_ok = fmaxf(1.f, __ldg(stats + 1));
  float P[12], X[4], p[3], g[3];
  project(a, e, P, X, p);
  const bool ok = a.hinge ? (p[2] >= a.margin) : (fabsf(p[2]) >= a.margin);
  if (ok) {
    const float iz = 1.f / p[2];
    const float rx = p[0] * iz - __ldg(a.obs + 2 * e), ry = p[1] * iz - __ldg(a.obs + 2 * e + 1);
    const float err = sqrtf(rx * rx + ry * ry);
    const float s = err > 0.f ? c * inv_e / err : 0.f;          // norm backward is 0 at the origin
    g[0] = s * rx * iz;
    g[1] = s * ry * iz;
    g[2] = -s * (rx * p[0] + ry * p[1]) * iz * iz;
  } else {
    g[0] = 0.f; g[1] = 0.f; g[2] = -c * inv_e * a.hinge_weight;
  }
  if (grad_mode == 2 || (grad_mode == 1 && ok)) {
    const float nrm = fmaxf(sqrtf(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]), 1e-12f);   // F.normalize eps
    const float f = 1.f / (nrm * (grad_mode == 2 ? (float)a.E : n_ok));
    g[0] *= f; g[1] *= f; g[2] *= f;
  }
  float4* row = reinterpret_cast<float4*>(G + e * 16);
  row[0] = make_float4(g[0] * X[0], g[0] * X[1], g[0] * X[2], g[0] * X[3]);
  row[1] = make_float4(g[1] * X[0], g[1] * X[1], g[1] * X[2], g[1] * X[3]);
  row[2] = make_float4(g[2] * X[0], g[2] * X[1], g[2] * X[2], g[2] * X[3]);
  row[3] = make_float4(P[0] * g[0] + P[4] * g[1] + P[8] * g[2], P[1] * g[0] + P[5] * g[1] + P[9] * g[2],
                       P[2] * g[0] + P[6] * g[1] + P[10] * g[2], P[3] * g[0] + P[7] * g[1] + P[11] * g[2]);
}

// Mean reprojection error over the observed pairs (evaluation.compute_core_errors -> 'our_repro', evaluation.py:27-32;
// geo_utils.reprojection_error_with_points, geo_utils.py:371-391): err_e = || u_e - (P_i X_j)_xy / (P_i X_j)_z ||,
// nan-mean (a 0/0 projection is skipped, an infinite one is not -- as numpy's nanmean does).
__global__ void __launch_bounds__(256) reproj_err_kernel(EsfmArgs a, float* __restrict__ partial) {
  __shared__ float sm[8][2];
  float sum = 0.f, cnt = 0.f;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < a.E; e += (int64_t)gridDim.x * blockDim.x) {
    float P[12], X[4], p[3];
    project(a, e, P, X, p);
    const float rx = __ldg(a.obs + 2 * e) - p[0] / p[2], ry = __ldg(a.obs + 2 * e + 1) - p[1] / p[2];
    const float err = sqrtf(rx * rx + ry * ry);
    if (err == err) { sum += err; cnt += 1.f; }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    sum += __shfl_xor_sync(0xffffffffu, sum, off);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, off);
  }
  if ((threadIdx.x & 31) == 0) { sm[threadIdx.x >> 5][0] = sum; sm[threadIdx.x >> 5][1] = cnt; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f, c = 0.f;
    for (int w = 0; w < 8; ++w) { s += sm[w][0]; c += sm[w][1]; }
    partial[2 * blockIdx.x] = s;
    partial[2 * blockIdx.x + 1] = c;
  }
}

// out[0] = sum / count (nan if nothing was counted, like nanmean of an all-nan array), out[1] = count
__global__ void reproj_finish_kernel(const float* __restrict__ partial, int blocks, float* __restrict__ out) {
  double s = 0.0, c = 0.0;
  for (int b = 0; b < blocks; ++b) { s += partial[2 * b]; c += partial[2 * b + 1]; }
  out[0] = (float)(s / c);
  out[1] = (float)c;
}

static int esfm_blocks(int64_t E) {
  int64_t need = (E + 255) / 256;
  int64_t cap = (int64_t)kNumSMs * 8;
  return (int)(need < 1 ? 1 : (need < cap ? need : cap));
}

}  // namespace gasfm

using namespace gasfm;

extern "C" size_t gasfm_esfm_loss_ws_bytes(int64_t E) { return (size_t)esfm_blocks(E) * 2 * sizeof(float); }

extern "C" int gasfm_esfm_loss_fwd(const float* Ps, const float* pts3D, int64_t n, const float* obs, const int32_t* row_idx,
                                   const int32_t* col_idx, int64_t E, float margin, int hinge, float hinge_weight,
                                   float* out, void* ws, void* stream) {
  GASFM_REQUIRE(E > 0 && n > 0 && ws != nullptr, "esfm_loss_fwd: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  EsfmArgs a{Ps, pts3D, n, obs, row_idx, col_idx, E, margin, hinge, hinge ? hinge_weight : 0.f};
  const int blocks = esfm_blocks(E);
  esfm_fwd_kernel<<<blocks, 256, 0, st>>>(a, (float*)ws);
  esfm_finish_kernel<<<1, 1, 0, st>>>((const float*)ws, blocks, E, out);
  return check_launch("esfm_loss_fwd");
}

extern "C" int gasfm_esfm_loss_bwd(const float* Ps, const float* pts3D, int64_t n, const float* obs, const int32_t* row_idx,
                                   const int32_t* col_idx, int64_t E, float margin, int hinge, float hinge_weight,
                                   const float* upstream, const float* stats, int grad_mode, float* G, void* stream) {
  GASFM_REQUIRE(E > 0 && n > 0 && (uintptr_t)G % 16 == 0, "esfm_loss_bwd: bad arguments");
  GASFM_REQUIRE(grad_mode >= 0 && grad_mode <= 2, "esfm_loss_bwd: grad_mode %d", grad_mode);
  EsfmArgs a{Ps, pts3D, n, obs, row_idx, col_idx, E, margin, hinge, hinge ? hinge_weight : 0.f};
  esfm_bwd_kernel<<<ceil_div(E, 256), 256, 0, (cudaStream_t)stream>>>(a, upstream, stats, grad_mode, G);
  return check_launch("esfm_loss_bwd");
}

extern "C" int gasfm_reproj_error(const float* Ps, const float* pts3D, int64_t n, const float* obs, const int32_t* row_idx,
                                  const int32_t* col_idx, int64_t E, float* out, void* ws, void* stream) {
  GASFM_REQUIRE(E > 0 && n > 0 && ws != nullptr && out != nullptr, "reproj_error: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  EsfmArgs a{Ps, pts3D, n, obs, row_idx, col_idx, E, 0.f, 0, 0.f};
  const int blocks = esfm_blocks(E);
  reproj_err_kernel<<<blocks, 256, 0, st>>>(a, (float*)ws);
  reproj_finish_kernel<<<1, 1, 0, st>>>((const float*)ws, blocks, out);
  return check_launch("reproj_error");
}
