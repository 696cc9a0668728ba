// Column sums of a row-major [rows, width] fp32 matrix: the second stage of every per-CTA-partials
// reduction (dW, dbias, dgamma/dbeta, datt, dW0) and the first stage of gasfm_col_sum.
#pragma once
#include <stdint.h>

namespace gasfm {

struct ColReduceJob {
  const float* src; int64_t ld; int64_t width; float* out;
  int pack_in, pack_out;   // pack_in > 0: columns come in groups of pack_in, of which the first pack_out are kept (dense)
};

constexpr int kColReduceGroups = 16;   // row groups per CTA (blockDim.y)

// grid.x = blocks(job a) + blocks(job b) with 32 columns per CTA; grid.y = row slices (each slice writes its own
// output row: out + slice * out_ld).  out[j] = scale * sum over the slice's rows of src[r, j].
static __global__ void __launch_bounds__(32 * kColReduceGroups)
col_reduce_kernel(ColReduceJob a, ColReduceJob b, int blocks_a, int rows, int rows_per_slice, int64_t out_ld, float scale) {
  __shared__ float part[kColReduceGroups][33];
  const bool second = (int)blockIdx.x >= blocks_a;
  const ColReduceJob& job = second ? b : a;
  const int64_t j = (int64_t)(blockIdx.x - (second ? blocks_a : 0)) * 32 + threadIdx.x;
  const int r_begin = blockIdx.y * rows_per_slice;
  const int r_end = min(rows, r_begin + rows_per_slice);
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  if (j < job.width) {
    const float* p = job.src + j;
    int r = r_begin + threadIdx.y;
    constexpr int G = kColReduceGroups;
    for (; r + 3 * G < r_end; r += 4 * G) {
      a0 += __ldg(p + (int64_t)r * job.ld);
      a1 += __ldg(p + (int64_t)(r + G) * job.ld);
      a2 += __ldg(p + (int64_t)(r + 2 * G) * job.ld);
      a3 += __ldg(p + (int64_t)(r + 3 * G) * job.ld);
    }
    for (; r < r_end; r += G) a0 += __ldg(p + (int64_t)r * job.ld);
  }
  part[threadIdx.y][threadIdx.x] = (a0 + a1) + (a2 + a3);
  __syncthreads();
  if (threadIdx.y != 0 || j >= job.width) return;
  float s = 0.f;
#pragma unroll
  for (int g = 0; g < kColReduceGroups; ++g) s += part[g][threadIdx.x];
  int64_t o = j;
  if (job.pack_in > 0) {
    const int q = (int)(j % job.pack_in);
    if (q >= job.pack_out) return;
    o = (j / job.pack_in) * job.pack_out + q;
  }
  job.out[(int64_t)blockIdx.y * out_ld + o] = s * scale;
}

// One launch for up to two independent jobs over the same number of partial rows.
static inline void launch_col_reduce(const ColReduceJob& a, const ColReduceJob* b, int rows, float scale, cudaStream_t st) {
  const int blocks_a = (int)((a.width + 31) / 32);
  const int blocks_b = b ? (int)((b->width + 31) / 32) : 0;
  col_reduce_kernel<<<dim3(blocks_a + blocks_b, 1), dim3(32, kColReduceGroups), 0, st>>>(a, b ? *b : a, blocks_a, rows, rows, 0, scale);
}

}  // namespace gasfm
