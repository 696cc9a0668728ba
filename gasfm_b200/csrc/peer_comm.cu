// Cross-GPU exchange of the track-sharded GASFM step over NVLink peer memory (no NCCL on the data path).
//
// A view's softmax runs over observations held by all ranks.  Each rank computes the un-normalised
// partial (max, sum, acc) of its local edges with the ordinary edge kernel; ONE kernel then
//   1. pushes the rank's partial rows into a region of every peer's exchange buffer (remote 128-bit stores over NVLink),
//   2. publishes them with a release flag per (source rank, CTA),
//   3. waits for the same CTA of every peer, and
//   4. combines the world partials locally (log-sum-exp merge, or a plain sum for gradient exchanges),
// in rank order, so that every rank computes bit-identical replicated results.  Rows are independent: a CTA only ever
// waits for the matching CTA of its peers, so the pushes of one CTA overlap the merges of the others, there is no
// grid-wide barrier, and nothing but these launches is needed -- the whole sharded step can be captured in a CUDA graph
// (NCCL collectives inside a capture deadlocked on this stack).
//
// Sequencing: the exchange number lives in DEVICE memory (state[0]) and is advanced by the last CTA of every launch, so
// a replayed graph needs no changing kernel argument.  Exchange s uses buffer slot s & 1; a rank can only be one exchange
// ahead of its slowest peer (it needs that peer's flags to finish), so two slots suffice.  Spins are bounded by a
// wall-clock limit (state[2] is set on expiry): a missing peer fails the step instead of hanging the GPU.
//
// The same merge arithmetic is exported for an already gathered buffer (gasfm_*_gathered): that is the torch.distributed
// (NCCL / gloo all_gather) arm used as the A/B baseline and by the single-GPU multi-process tests.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "common.cuh"
#include "../../include/gasfm_b200.h"

namespace gasfm {

constexpr int kPeerMaxWorld = 8;
constexpr int kPeerMaxCtas = 1024;          // flag slots per source rank
constexpr int kPeerThreads = 256;
constexpr int kPeerStateWords = 8;          // [0] exchange number, [1] finished CTAs, [2] error

struct PeerComm {
  float* buf[kPeerMaxWorld];                // exchange buffer of every rank (buf[rank] is local): [2 slots][world][region]
  uint32_t* flags[kPeerMaxWorld];           // flag array of every rank: [world sources][kPeerMaxCtas]
  uint32_t* state;                          // local
  int rank, world;
  int64_t region_floats;
  unsigned long long timeout_ns;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_volatile4(const float* p) {
  float4 r;
  asm volatile("ld.volatile.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ float ld_volatile1(const float* p) {
  float r;
  asm volatile("ld.volatile.global.f32 %0, [%1];" : "=f"(r) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// Publish this CTA's pushes to every peer, then wait for the same CTA of every peer.  Called by all threads.
__device__ __forceinline__ void publish_and_wait(const PeerComm& c, uint32_t seq) {
  __syncthreads();                                            // every thread's remote stores are issued
  const int r = threadIdx.x;
  if (r < c.world && r != c.rank) {
    __threadfence_system();
    st_release_sys(c.flags[r] + (int64_t)c.rank * kPeerMaxCtas + blockIdx.x, seq + 1);
    const uint32_t* mine = c.flags[c.rank] + (int64_t)r * kPeerMaxCtas + blockIdx.x;
    const unsigned long long t0 = global_ns();
    int spins = 0;
    const bool broken = *reinterpret_cast<volatile uint32_t*>(c.state + 2) != 0;   // an earlier exchange timed out: do not wait again
    while (!broken && (int32_t)(ld_acquire_sys(mine) - (seq + 1)) < 0) {
      if ((++spins & 1023) == 0 && global_ns() - t0 > c.timeout_ns) {
        atomicExch(c.state + 2, 1u);
        break;
      }
    }
  }
  __syncthreads();
}

__device__ __forceinline__ void finish_exchange(const PeerComm& c, uint32_t seq) {
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t done = atomicAdd(c.state + 1, 1u);
    if (done == gridDim.x - 1) {
      c.state[1] = 0;
      __threadfence();
      c.state[0] = seq + 1;
    }
  }
}

// ---- plain sum (gradient exchanges): out = scale * sum_r in_r, n floats (n % 4 == 0), each CTA owns a contiguous range ----
template <bool GATHERED>
__global__ void __launch_bounds__(kPeerThreads) peer_sum_kernel(PeerComm c, const float* __restrict__ in, float* __restrict__ out,
                                                                int64_t n4, float scale) {
  const uint32_t seq = GATHERED ? 0u : *reinterpret_cast<volatile uint32_t*>(c.state);
  const int64_t per = (n4 + gridDim.x - 1) / gridDim.x;
  const int64_t begin = blockIdx.x * per, end = min(n4, begin + per);
  const int64_t slot_base = GATHERED ? 0 : (int64_t)(seq & 1u) * c.world * c.region_floats;
  if (!GATHERED) {
    for (int64_t i = begin + threadIdx.x; i < end; i += kPeerThreads) {
      const float4 v = *reinterpret_cast<const float4*>(in + 4 * i);
#pragma unroll 1
      for (int r = 0; r < c.world; ++r)
        if (r != c.rank) st4(c.buf[r] + slot_base + (int64_t)c.rank * c.region_floats + 4 * i, v);
    }
    publish_and_wait(c, seq);
  }
  const float* local = GATHERED ? in : c.buf[c.rank] + slot_base;
  for (int64_t i = begin + threadIdx.x; i < end; i += kPeerThreads) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
    for (int r = 0; r < c.world; ++r) {
      const float4 v = (!GATHERED && r == c.rank) ? *reinterpret_cast<const float4*>(in + 4 * i)
                                                  : ld_volatile4(local + (int64_t)r * c.region_floats + 4 * i);
      a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
    st4(out + 4 * i, make_float4(a.x * scale, a.y * scale, a.z * scale, a.w * scale));
  }
  if (!GATHERED) finish_exchange(c, seq);
}

// ---- log-sum-exp merge of per-rank softmax partials -------------------------------------------------------------
// Region layout of one rank: acc [T, HC] | max [T, H] | sum [T, H].  One warp per target row; a CTA owns kPeerThreads / 32 rows.
struct LseArgs {
  const float* acc; const float* mx; const float* sm;   // this rank's partial (GATHERED: unused, everything is in the buffer)
  const float* bias;                                    // optional [HC], added to the normalised result
  float* out; float* M; float* L;                       // merged result [T, HC] and global statistics [T, H]
  int T, H, C;
};

template <bool GATHERED>
__global__ void __launch_bounds__(kPeerThreads) peer_lse_kernel(PeerComm c, LseArgs a, const float* __restrict__ gathered) {
  const uint32_t seq = GATHERED ? 0u : *reinterpret_cast<volatile uint32_t*>(c.state);
  const int HC = a.H * a.C;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int t = blockIdx.x * (kPeerThreads / 32) + wid;
  const int64_t slot_base = GATHERED ? 0 : (int64_t)(seq & 1u) * c.world * c.region_floats;
  const int64_t off_mx = (int64_t)a.T * HC, off_sm = off_mx + (int64_t)a.T * a.H;
  if (!GATHERED) {
    if (t < a.T) {
#pragma unroll 1
      for (int r = 0; r < c.world; ++r) {
        if (r == c.rank) continue;
        float* dst = c.buf[r] + slot_base + (int64_t)c.rank * c.region_floats;
        for (int j = lane; j < HC; j += 32) dst[(int64_t)t * HC + j] = a.acc[(int64_t)t * HC + j];
        if (lane < a.H) {
          dst[off_mx + (int64_t)t * a.H + lane] = a.mx[(int64_t)t * a.H + lane];
          dst[off_sm + (int64_t)t * a.H + lane] = a.sm[(int64_t)t * a.H + lane];
        }
      }
    }
    publish_and_wait(c, seq);
  }
  if (t < a.T) {
    const float* local = GATHERED ? gathered : c.buf[c.rank] + slot_base;
    for (int j = lane; j < HC; j += 32) {
      const int h = j / a.C;
      float mr[kPeerMaxWorld], M = -INFINITY;
#pragma unroll 1
      for (int r = 0; r < c.world; ++r) {
        mr[r] = (!GATHERED && r == c.rank) ? a.mx[(int64_t)t * a.H + h]
                                           : ld_volatile1(local + (int64_t)r * c.region_floats + off_mx + (int64_t)t * a.H + h);
        M = fmaxf(M, mr[r]);
      }
      float Lsum = 0.f, A = 0.f;
#pragma unroll 1
      for (int r = 0; r < c.world; ++r) {
        const float w = mr[r] == -INFINITY ? 0.f : __expf(mr[r] - M);      // ranks without edges of this target weigh 0
        const bool own = !GATHERED && r == c.rank;
        const float* reg = local + (int64_t)r * c.region_floats;
        const float s = own ? a.sm[(int64_t)t * a.H + h] : ld_volatile1(reg + off_sm + (int64_t)t * a.H + h);
        const float v = own ? a.acc[(int64_t)t * HC + j] : ld_volatile1(reg + (int64_t)t * HC + j);
        Lsum = fmaf(w, s, Lsum);
        A = fmaf(w, v, A);
      }
      float o = Lsum > 0.f ? A / Lsum : 0.f;                  // empty segment: 0 (+ bias), like the single-GPU kernel
      if (a.bias) o += a.bias[j];
      a.out[(int64_t)t * HC + j] = o;
      if (j % a.C == 0) {
        a.M[(int64_t)t * a.H + h] = M;
        a.L[(int64_t)t * a.H + h] = Lsum;
      }
    }
  }
  if (!GATHERED) finish_exchange(c, seq);
}

static int sum_grid(int64_t n4) {
  int64_t g = (n4 + 1023) / 1024;            // >= 16 KB per CTA
  if (g < 1) g = 1;
  if (g > 2 * kNumSMs) g = 2 * kNumSMs;
  return (int)g;
}

}  // namespace gasfm

using namespace gasfm;

#define CU_TRY(expr, what)                                                              \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess) {                                                            \
      set_error("%s: CUDA error %d (%s)", what, (int)_e, cudaGetErrorString(_e));       \
      return (int)_e;                                                                   \
    }                                                                                   \
  } while (0)

extern "C" int gasfm_peer_alloc(size_t bytes, void** ptr) {
  GASFM_REQUIRE(ptr != nullptr, "peer_alloc: NULL argument");
  CU_TRY(cudaMalloc(ptr, bytes ? bytes : 16), "peer_alloc");
  CU_TRY(cudaMemset(*ptr, 0, bytes ? bytes : 16), "peer_alloc(memset)");
  return 0;
}
extern "C" int gasfm_peer_free(void* ptr) {
  CU_TRY(cudaFree(ptr), "peer_free");
  return 0;
}
extern "C" int gasfm_peer_export(void* ptr, void* handle64) {
  GASFM_REQUIRE(ptr && handle64, "peer_export: NULL argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == GASFM_PEER_HANDLE_BYTES, "IPC handle size");
  cudaIpcMemHandle_t h;
  CU_TRY(cudaIpcGetMemHandle(&h, ptr), "peer_export");
  memcpy(handle64, &h, sizeof(h));
  return 0;
}
extern "C" int gasfm_peer_import(const void* handle64, void** ptr) {
  GASFM_REQUIRE(ptr && handle64, "peer_import: NULL argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  CU_TRY(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess), "peer_import");
  return 0;
}
extern "C" int gasfm_peer_close(void* ptr) {
  CU_TRY(cudaIpcCloseMemHandle(ptr), "peer_close");
  return 0;
}

extern "C" size_t gasfm_peer_buffer_bytes(int world, int64_t region_floats) {
  return (size_t)2 * world * region_floats * sizeof(float);
}
extern "C" size_t gasfm_peer_flags_bytes(int world) {
  return ((size_t)world * kPeerMaxCtas + kPeerStateWords) * sizeof(uint32_t);
}

extern "C" int gasfm_peer_comm_create(int rank, int world, void* const* bufs, void* const* flags, int64_t region_floats,
                                      double timeout_s, void** comm_out) {
  GASFM_REQUIRE(comm_out && bufs && flags, "peer_comm_create: NULL argument");
  GASFM_REQUIRE(world >= 1 && world <= kPeerMaxWorld && rank >= 0 && rank < world, "peer_comm_create: rank %d / world %d (max %d)",
                rank, world, kPeerMaxWorld);
  GASFM_REQUIRE(region_floats > 0 && region_floats % 4 == 0, "peer_comm_create: region must be a positive multiple of 4 floats");
  PeerComm* c = new PeerComm();
  memset(c, 0, sizeof(*c));
  for (int r = 0; r < world; ++r) {
    GASFM_REQUIRE(bufs[r] && flags[r], "peer_comm_create: NULL buffer for rank %d", r);
    c->buf[r] = (float*)bufs[r];
    c->flags[r] = (uint32_t*)flags[r];
  }
  c->state = (uint32_t*)flags[rank] + (size_t)world * kPeerMaxCtas;
  c->rank = rank; c->world = world; c->region_floats = region_floats;
  c->timeout_ns = (unsigned long long)((timeout_s > 0 ? timeout_s : 10.0) * 1e9);
  *comm_out = c;
  return 0;
}
extern "C" int gasfm_peer_comm_destroy(void* comm) {
  delete (PeerComm*)comm;
  return 0;
}
extern "C" int gasfm_peer_comm_error(void* comm, int* error_out) {
  GASFM_REQUIRE(comm && error_out, "peer_comm_error: NULL argument");
  uint32_t e = 0;
  CU_TRY(cudaMemcpy(&e, ((PeerComm*)comm)->state + 2, sizeof(e), cudaMemcpyDeviceToHost), "peer_comm_error");
  *error_out = (int)e;
  return 0;
}

extern "C" int gasfm_peer_allreduce_sum(void* comm, const float* in, float* out, int64_t n, float scale, void* stream) {
  GASFM_REQUIRE(comm != nullptr, "peer_allreduce_sum: NULL communicator");
  const PeerComm& c = *(PeerComm*)comm;
  GASFM_REQUIRE(n >= 0 && n % 4 == 0 && n <= c.region_floats, "peer_allreduce_sum: n=%lld must be a multiple of 4 and <= %lld",
                (long long)n, (long long)c.region_floats);
  GASFM_REQUIRE(n == 0 || (in && out && ((uintptr_t)in | (uintptr_t)out) % 16 == 0), "peer_allreduce_sum: bad pointers");
  peer_sum_kernel<false><<<sum_grid(n / 4), kPeerThreads, 0, (cudaStream_t)stream>>>(c, in, out, n / 4, scale);
  return check_launch("peer_allreduce_sum");
}

extern "C" int gasfm_peer_lse_merge(void* comm, const float* acc, const float* seg_max, const float* seg_sum, const float* bias,
                                    int n_seg, int heads, int head_dim, float* out, float* M, float* L, void* stream) {
  GASFM_REQUIRE(comm != nullptr, "peer_lse_merge: NULL communicator");
  const PeerComm& c = *(PeerComm*)comm;
  const int64_t need = (int64_t)n_seg * heads * (head_dim + 2);
  GASFM_REQUIRE(n_seg > 0 && heads > 0 && head_dim > 0 && need <= c.region_floats, "peer_lse_merge: %lld floats exceed the region (%lld)",
                (long long)need, (long long)c.region_floats);
  const int grid = (n_seg + kPeerThreads / 32 - 1) / (kPeerThreads / 32);
  GASFM_REQUIRE(grid <= kPeerMaxCtas, "peer_lse_merge: %d targets exceed %d per launch", n_seg, kPeerMaxCtas * (kPeerThreads / 32));
  LseArgs a{acc, seg_max, seg_sum, bias, out, M, L, n_seg, heads, head_dim};
  peer_lse_kernel<false><<<grid, kPeerThreads, 0, (cudaStream_t)stream>>>(c, a, nullptr);
  return check_launch("peer_lse_merge");
}

// The same arithmetic on a buffer that already holds every rank's region (torch.distributed all_gather):
// gathered = [world][region_floats], region = acc | max | sum (lse) or the n summands (sum).
extern "C" int gasfm_lse_merge_gathered(const float* gathered, int world, int64_t region_floats, const float* bias, int n_seg,
                                        int heads, int head_dim, float* out, float* M, float* L, void* stream) {
  GASFM_REQUIRE(gathered && out && M && L && world >= 1 && world <= kPeerMaxWorld, "lse_merge_gathered: bad arguments");
  GASFM_REQUIRE((int64_t)n_seg * heads * (head_dim + 2) <= region_floats, "lse_merge_gathered: region too small");
  PeerComm c;
  memset(&c, 0, sizeof(c));
  c.world = world; c.rank = -1; c.region_floats = region_floats;
  LseArgs a{nullptr, nullptr, nullptr, bias, out, M, L, n_seg, heads, head_dim};
  const int grid = (n_seg + kPeerThreads / 32 - 1) / (kPeerThreads / 32);
  if (n_seg > 0) peer_lse_kernel<true><<<grid, kPeerThreads, 0, (cudaStream_t)stream>>>(c, a, gathered);
  return check_launch("lse_merge_gathered");
}

extern "C" int gasfm_sum_gathered(const float* gathered, int world, int64_t region_floats, int64_t n, float scale, float* out,
                                  void* stream) {
  GASFM_REQUIRE(gathered && out && world >= 1 && world <= kPeerMaxWorld && n % 4 == 0 && n <= region_floats && region_floats % 4 == 0,
                "sum_gathered: bad arguments");
  PeerComm c;
  memset(&c, 0, sizeof(c));
  c.world = world; c.rank = -1; c.region_floats = region_floats;
  if (n > 0) peer_sum_kernel<true><<<sum_grid(n / 4), kPeerThreads, 0, (cudaStream_t)stream>>>(c, gathered, out, n / 4, scale);
  return check_launch("sum_gathered");
}
