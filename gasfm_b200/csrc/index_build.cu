// Observation-index construction on the GPU (bit-exact integer work).
//
// Replaces, for the hot path, what the reference does on the host with numpy / torch.sparse:
//   get_M_valid_points + M2sparse            code/utils/dataset_utils.py:86-156
//   normalize_M                              code/utils/geo_utils.py:689-703
//   AxialAggregationGraphWrapper edge lists  code/utils/dataset_utils.py:511-537
//   the repeated .coalesce() sorts           code/utils/sparse_utils.py:436-449
// Everything is built once per scene; the attention / pooling kernels only read it.
#include "common.cuh"
#include "../../include/gasfm_b200.h"

namespace gasfm {

// ---- single-block exclusive scan with carry (n up to a few hundred thousand) ---------------
template <class F>
__global__ void __launch_bounds__(1024) scan_excl_kernel(F f, int n, int32_t* out) {
  __shared__ int warp_sums[32];
  __shared__ int carry_s;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < n; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = i < n ? f(i) : 0;
    int x = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      int y = __shfl_up_sync(0xffffffffu, x, off);
      if (lane >= off) x += y;
    }
    if (lane == 31) warp_sums[wid] = x;
    __syncthreads();
    if (wid == 0) {
      int w = warp_sums[lane];
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, w, off);
        if (lane >= off) w += y;
      }
      warp_sums[lane] = w;
    }
    __syncthreads();
    const int carry = carry_s;
    const int incl = x + (wid > 0 ? warp_sums[wid - 1] : 0) + carry;
    if (i < n) out[i] = incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) out[n] = carry_s;
}

// ---- M2sparse -------------------------------------------------------------------------------
// one thread per track (column): count views, drop tracks seen in < min_views views
__global__ void valid_columns_kernel(const float* __restrict__ M, int m, int n, int min_views,
                                     uint8_t* __restrict__ valid, int64_t* __restrict__ cam_per_pts) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  int cnt = 0;
  for (int i = 0; i < m; ++i) {
    float x = M[(int64_t)(2 * i) * n + j], y = M[(int64_t)(2 * i + 1) * n + j];
    cnt += (fabsf(x) + fabsf(y)) != 0.f;
  }
  const bool keep = cnt >= min_views;
  cam_per_pts[j] = keep ? cnt : 0;
  for (int i = 0; i < m; ++i) {
    float x = M[(int64_t)(2 * i) * n + j], y = M[(int64_t)(2 * i + 1) * n + j];
    valid[(int64_t)i * n + j] = (keep && (fabsf(x) + fabsf(y)) != 0.f) ? 1 : 0;
  }
}

// one CTA per view (row): pts_per_cam
__global__ void __launch_bounds__(256) row_count_kernel(const uint8_t* __restrict__ valid, int n,
                                                        int64_t* __restrict__ pts_per_cam) {
  __shared__ int sm[8];
  const int i = blockIdx.x;
  int c = 0;
  for (int j = threadIdx.x; j < n; j += 256) c += valid[(int64_t)i * n + j];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int s = 0;
    for (int w = 0; w < 8; ++w) s += sm[w];
    pts_per_cam[i] = s;
  }
}

__global__ void __launch_bounds__(256) total_kernel(const int64_t* __restrict__ v, int m, int64_t* out) {
  __shared__ long long sm[8];
  long long c = 0;
  for (int i = threadIdx.x; i < m; i += 256) c += v[i];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    long long s = 0;
    for (int w = 0; w < 8; ++w) s += sm[w];
    *out = s;
  }
}

// Row-major stream compaction of the flattened mask: tile counts -> scan -> scatter.
constexpr int kTile = 4096;  // elements per CTA (256 threads x 16)
__global__ void __launch_bounds__(256) tile_count_kernel(const uint8_t* __restrict__ valid, int64_t total,
                                                         int64_t* __restrict__ tile_cnt) {
  __shared__ int sm[8];
  const int64_t base = (int64_t)blockIdx.x * kTile;
  int c = 0;
  for (int k = threadIdx.x; k < kTile; k += 256) {
    int64_t idx = base + k;
    if (idx < total) c += valid[idx];
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int s = 0;
    for (int w = 0; w < 8; ++w) s += sm[w];
    tile_cnt[blockIdx.x] = s;
  }
}

__global__ void __launch_bounds__(1024) tile_scan_kernel(int64_t* tile_cnt, int n_tiles) {
  // in-place exclusive scan (int64) by one CTA
  __shared__ long long warp_sums[32];
  __shared__ long long carry_s;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < n_tiles; base += 1024) {
    const int i = base + threadIdx.x;
    const long long v = i < n_tiles ? tile_cnt[i] : 0;
    long long x = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      long long y = __shfl_up_sync(0xffffffffu, x, off);
      if (lane >= off) x += y;
    }
    if (lane == 31) warp_sums[wid] = x;
    __syncthreads();
    if (wid == 0) {
      long long w = warp_sums[lane];
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        long long y = __shfl_up_sync(0xffffffffu, w, off);
        if (lane >= off) w += y;
      }
      warp_sums[lane] = w;
    }
    __syncthreads();
    const long long carry = carry_s;
    const long long incl = x + (wid > 0 ? warp_sums[wid - 1] : 0) + carry;
    if (i < n_tiles) tile_cnt[i] = incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = incl;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) compact_kernel(const float* __restrict__ M, const float* __restrict__ Ns,
                                                      const uint8_t* __restrict__ valid, int m, int n,
                                                      int64_t n_obs, const int64_t* __restrict__ tile_off,
                                                      int64_t* __restrict__ indices, float* __restrict__ values) {
  // each thread owns 16 consecutive elements of the tile so that output order == input order
  __shared__ int warp_sums[8];
  const int64_t total = (int64_t)m * n;
  const int64_t base = (int64_t)blockIdx.x * kTile + (int64_t)threadIdx.x * 16;
  int flags = 0, c = 0;
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    int64_t idx = base + k;
    int f = idx < total ? valid[idx] : 0;
    flags |= f << k;
    c += f;
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int x = c;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    int y = __shfl_up_sync(0xffffffffu, x, off);
    if (lane >= off) x += y;
  }
  if (lane == 31) warp_sums[wid] = x;
  __syncthreads();
  int wbase = 0;
  for (int w = 0; w < wid; ++w) wbase += warp_sums[w];
  int64_t pos = tile_off[blockIdx.x] + wbase + (x - c);
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    if (flags & (1 << k)) {
      int64_t idx = base + k;
      int i = (int)(idx / n), j = (int)(idx % n);
      indices[pos] = i;
      indices[n_obs + pos] = j;
      float px = M[(int64_t)(2 * i) * n + j], py = M[(int64_t)(2 * i + 1) * n + j];
      float vx = px, vy = py;
      if (Ns) {
        const float* N = Ns + 9 * i;
        // (Ns @ [x; y; 1])[:2], accumulated left to right like a K=3 dot product
        vx = fmaf(N[1], py, N[0] * px) + N[2];
        vy = fmaf(N[4], py, N[3] * px) + N[5];
      }
      values[2 * pos] = vx;
      values[2 * pos + 1] = vy;
      ++pos;
    }
  }
}

// ---- CSR / CSC ------------------------------------------------------------------------------
__global__ void csr_cast_kernel(const int64_t* __restrict__ indices, int64_t E, int m, int n,
                                int32_t* __restrict__ row_idx, int32_t* __restrict__ col_idx,
                                int32_t* __restrict__ row_ptr, int32_t* __restrict__ col_cnt,
                                int32_t* __restrict__ status) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const int64_t r = indices[e], c = indices[E + e];
  if (r < 0 || r >= m || c < 0 || c >= n) { atomicOr(status, 1); return; }
  row_idx[e] = (int32_t)r;
  col_idx[e] = (int32_t)c;
  int64_t rp = -1;
  if (e > 0) {
    rp = indices[e - 1];
    const int64_t cp = indices[E + e - 1];
    if (rp > r || (rp == r && cp >= c)) atomicOr(status, 2);  // not row-major sorted / duplicate
    if (rp < 0 || rp >= m) return;
  }
  // storage order is CSR order: row_ptr[q] = first edge whose row is >= q
  for (int64_t q = rp + 1; q <= r; ++q) row_ptr[q] = (int32_t)e;
  if (e == E - 1)
    for (int64_t q = r + 1; q <= m; ++q) row_ptr[q] = (int32_t)E;
  atomicAdd(col_cnt + c, 1);
}

__global__ void fill_i32_kernel(int32_t* p, int64_t n, int32_t v) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

struct LoadI32 {
  const int32_t* p;
  __device__ int operator()(int i) const { return p[i]; }
};

__global__ void csc_claim_kernel(const int32_t* __restrict__ col_idx, int64_t E,
                                 const int32_t* __restrict__ col_ptr, int32_t* __restrict__ fill,
                                 int32_t* __restrict__ tmp) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const int c = col_idx[e];
  const int pos = col_ptr[c] + atomicAdd(fill + c, 1);
  tmp[pos] = (int32_t)e;
}

// one warp per track: rank-sort the (unordered) claimed edge ids -> stable CSC order
__global__ void __launch_bounds__(256) csc_sort_kernel(const int32_t* __restrict__ col_ptr, int n,
                                                       const int32_t* __restrict__ tmp, int32_t* __restrict__ perm) {
  const int c = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (c >= n) return;
  const int b = col_ptr[c], len = col_ptr[c + 1] - b;
  for (int i = lane; i < len; i += 32) {
    const int v = tmp[b + i];
    int rank = 0;
    for (int j = 0; j < len; ++j) rank += tmp[b + j] < v;
    perm[b + rank] = v;
  }
}

// ---- chunk tables ---------------------------------------------------------------------------
struct ChunkCount {
  const int32_t* seg_ptr; int chunk;
  __device__ int operator()(int t) const { return (seg_ptr[t + 1] - seg_ptr[t] + chunk - 1) / chunk; }
};

__global__ void chunk_seg_kernel(const int32_t* __restrict__ chunk_ptr, int n_seg, int max_chunks,
                                 int32_t* __restrict__ chunk_seg) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_seg) return;
  const int c1 = min(chunk_ptr[t + 1], max_chunks);
  for (int k = chunk_ptr[t]; k < c1; ++k) chunk_seg[k] = t;
}

}  // namespace gasfm

using namespace gasfm;

extern "C" int gasfm_m2sparse_count(const float* M, int m, int n, int min_views_per_point, uint8_t* valid,
                                    int64_t* cam_per_pts, int64_t* pts_per_cam, int64_t* n_obs, void* stream) {
  GASFM_REQUIRE(m > 0 && n > 0, "m2sparse_count: empty measurement matrix (%d x %d)", m, n);
  cudaStream_t st = (cudaStream_t)stream;
  valid_columns_kernel<<<ceil_div(n, 128), 128, 0, st>>>(M, m, n, min_views_per_point, valid, cam_per_pts);
  row_count_kernel<<<m, 256, 0, st>>>(valid, n, pts_per_cam);
  total_kernel<<<1, 256, 0, st>>>(pts_per_cam, m, n_obs);
  return check_launch("m2sparse_count");
}

extern "C" size_t gasfm_m2sparse_ws_bytes(int m, int n) {
  return ((size_t)m * n / kTile + 2) * sizeof(int64_t);
}

extern "C" int gasfm_m2sparse_fill(const float* M, const float* Ns, const uint8_t* valid, int m, int n,
                                   int64_t n_obs, int64_t* indices, float* values, void* scan_ws, void* stream) {
  GASFM_REQUIRE(m > 0 && n > 0 && n_obs >= 0, "m2sparse_fill: bad sizes");
  if (n_obs == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t total = (int64_t)m * n;
  const int tiles = ceil_div(total, kTile);
  int64_t* tile_off = (int64_t*)scan_ws;
  tile_count_kernel<<<tiles, 256, 0, st>>>(valid, total, tile_off);
  tile_scan_kernel<<<1, 1024, 0, st>>>(tile_off, tiles);
  compact_kernel<<<tiles, 256, 0, st>>>(M, Ns, valid, m, n, n_obs, tile_off, indices, values);
  return check_launch("m2sparse_fill");
}

extern "C" size_t gasfm_csr_build_ws_bytes(int64_t n_obs, int n) {
  return ((size_t)n + (size_t)n_obs + 8) * sizeof(int32_t);
}

extern "C" int gasfm_csr_build(const int64_t* indices, int64_t n_obs, int m, int n, int32_t* row_idx,
                               int32_t* col_idx, int32_t* row_ptr, int32_t* col_ptr, int32_t* csc_perm,
                               int32_t* status, void* ws, void* stream) {
  GASFM_REQUIRE(m > 0 && n > 0 && n_obs >= 0, "csr_build: bad sizes");
  GASFM_REQUIRE(n_obs < (int64_t)INT32_MAX, "csr_build: more than 2^31 observations");
  GASFM_REQUIRE(ws != nullptr, "csr_build: workspace of gasfm_csr_build_ws_bytes() bytes required");
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(status, 0, sizeof(int32_t), st);
  cudaMemsetAsync(col_ptr, 0, (size_t)(n + 1) * sizeof(int32_t), st);
  if (n_obs == 0) {
    cudaMemsetAsync(row_ptr, 0, (size_t)(m + 1) * sizeof(int32_t), st);
    return check_launch("csr_build(empty)");
  }
  const int blocks = ceil_div(n_obs, 256);
  int32_t* cnt = (int32_t*)ws;                 // [n] per-track counters
  int32_t* tmp = cnt + n;                      // [n_obs] claimed (unordered) CSC slots
  cudaMemsetAsync(cnt, 0, (size_t)n * sizeof(int32_t), st);
  csr_cast_kernel<<<blocks, 256, 0, st>>>(indices, n_obs, m, n, row_idx, col_idx, row_ptr, cnt, status);
  scan_excl_kernel<<<1, 1024, 0, st>>>(LoadI32{cnt}, n, col_ptr);
  cudaMemsetAsync(cnt, 0, (size_t)n * sizeof(int32_t), st);
  csc_claim_kernel<<<blocks, 256, 0, st>>>(col_idx, n_obs, col_ptr, cnt, tmp);
  csc_sort_kernel<<<ceil_div((int64_t)n * 32, 256), 256, 0, st>>>(col_ptr, n, tmp, csc_perm);
  return check_launch("csr_build");
}

extern "C" int gasfm_plan_chunks(const int32_t* seg_ptr, int n_seg, int chunk, int32_t* chunk_ptr,
                                 int32_t* chunk_seg, int max_chunks, void* stream) {
  GASFM_REQUIRE(n_seg > 0 && chunk > 0 && max_chunks > 0, "plan_chunks: bad sizes");
  cudaStream_t st = (cudaStream_t)stream;
  scan_excl_kernel<<<1, 1024, 0, st>>>(ChunkCount{seg_ptr, chunk}, n_seg, chunk_ptr);
  chunk_seg_kernel<<<ceil_div(n_seg, 128), 128, 0, st>>>(chunk_ptr, n_seg, max_chunks, chunk_seg);
  return check_launch("plan_chunks");
}
