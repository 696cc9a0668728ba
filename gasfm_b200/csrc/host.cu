// Error reporting and the host-buffer entry points of the C ABI.
#include <stdarg.h>
#include <string.h>
#include <vector>

#include "common.cuh"
#include "../../include/gasfm_b200.h"

namespace gasfm {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) return 0;
  set_error("%s: CUDA error %d (%s)", what, (int)e, cudaGetErrorString(e));
  return (int)e == 0 ? 1 : (int)e;
}

// RAII device buffer for the host-buffer entry points
struct DevBuf {
  void* p = nullptr;
  cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1); }
  ~DevBuf() { if (p) cudaFree(p); }
  template <class T> T* as() { return (T*)p; }
};

#define CU_OK(expr, what)                                                               \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess) {                                                            \
      set_error("%s: CUDA error %d (%s)", what, (int)_e, cudaGetErrorString(_e));       \
      return (int)_e;                                                                   \
    }                                                                                   \
  } while (0)

}  // namespace gasfm

using namespace gasfm;

extern "C" int gasfm_abi_version(void) { return GASFM_B200_ABI_VERSION; }
extern "C" const char* gasfm_last_error(void) { return g_err; }

extern "C" int gasfm_csr_build_host(const int64_t* indices_host, int64_t n_obs, int m, int n, int32_t* row_ptr_host,
                                    int32_t* col_ptr_host, int32_t* csc_perm_host) {
  GASFM_REQUIRE(indices_host && row_ptr_host && col_ptr_host && csc_perm_host, "csr_build_host: NULL argument");
  DevBuf idx, ri, ci, rp, cp, perm, status, ws;
  CU_OK(ws.alloc(gasfm_csr_build_ws_bytes(n_obs, n)), "csr_build_host");
  CU_OK(idx.alloc((size_t)2 * n_obs * sizeof(int64_t)), "csr_build_host");
  CU_OK(ri.alloc((size_t)n_obs * 4), "csr_build_host");
  CU_OK(ci.alloc((size_t)n_obs * 4), "csr_build_host");
  CU_OK(rp.alloc((size_t)(m + 1) * 4), "csr_build_host");
  CU_OK(cp.alloc((size_t)(n + 1) * 4), "csr_build_host");
  CU_OK(perm.alloc((size_t)n_obs * 4), "csr_build_host");
  CU_OK(status.alloc(4), "csr_build_host");
  CU_OK(cudaMemcpy(idx.p, indices_host, (size_t)2 * n_obs * sizeof(int64_t), cudaMemcpyHostToDevice), "csr_build_host");
  int rc = gasfm_csr_build(idx.as<int64_t>(), n_obs, m, n, ri.as<int32_t>(), ci.as<int32_t>(), rp.as<int32_t>(),
                           cp.as<int32_t>(), perm.as<int32_t>(), status.as<int32_t>(), ws.p, nullptr);
  if (rc) return rc;
  int32_t st = 0;
  CU_OK(cudaMemcpy(&st, status.p, 4, cudaMemcpyDeviceToHost), "csr_build_host");
  GASFM_REQUIRE(st == 0, "csr_build_host: indices are %s", (st & 1) ? "out of range" : "not sorted row-major / contain duplicates");
  CU_OK(cudaMemcpy(row_ptr_host, rp.p, (size_t)(m + 1) * 4, cudaMemcpyDeviceToHost), "csr_build_host");
  CU_OK(cudaMemcpy(col_ptr_host, cp.p, (size_t)(n + 1) * 4, cudaMemcpyDeviceToHost), "csr_build_host");
  CU_OK(cudaMemcpy(csc_perm_host, perm.p, (size_t)n_obs * 4, cudaMemcpyDeviceToHost), "csr_build_host");
  return 0;
}

extern "C" int gasfm_gat_edge_fwd_host(const float* XL_host, const float* XR_host, const float* att_host,
                                       const float* bias_host, const int64_t* target_host, int64_t n_obs, int n_seg,
                                       int heads, int head_dim, float slope, float* out_host) {
  GASFM_REQUIRE(XL_host && XR_host && att_host && target_host && out_host, "gat_edge_fwd_host: NULL argument");
  GASFM_REQUIRE(n_obs > 0 && n_seg > 0, "gat_edge_fwd_host: empty graph");
  const size_t HC = (size_t)heads * head_dim;
  // Segment the edges by target: a (edge id, target) pair list is row-major sorted by construction,
  // so the CSC half of csr_build yields the stable grouping by target.
  std::vector<int64_t> pairs((size_t)2 * n_obs);
  for (int64_t e = 0; e < n_obs; ++e) { pairs[e] = e; pairs[n_obs + e] = target_host[e]; }
  DevBuf idx, ri, ci, rp, cp, perm, status, xl, xr, att, bias, out, smax, ssum, ws, cptr, cseg, csr_ws;
  CU_OK(csr_ws.alloc(gasfm_csr_build_ws_bytes(n_obs, n_seg)), "gat_edge_fwd_host");
  CU_OK(idx.alloc(pairs.size() * sizeof(int64_t)), "gat_edge_fwd_host");
  CU_OK(ri.alloc(n_obs * 4), "gat_edge_fwd_host");
  CU_OK(ci.alloc(n_obs * 4), "gat_edge_fwd_host");
  CU_OK(rp.alloc((size_t)(n_obs + 1) * 4), "gat_edge_fwd_host");
  CU_OK(cp.alloc((size_t)(n_seg + 1) * 4), "gat_edge_fwd_host");
  CU_OK(perm.alloc(n_obs * 4), "gat_edge_fwd_host");
  CU_OK(status.alloc(4), "gat_edge_fwd_host");
  CU_OK(xl.alloc(n_obs * HC * 4), "gat_edge_fwd_host");
  CU_OK(xr.alloc(n_seg * HC * 4), "gat_edge_fwd_host");
  CU_OK(att.alloc(HC * 4), "gat_edge_fwd_host");
  CU_OK(bias.alloc(HC * 4), "gat_edge_fwd_host");
  CU_OK(out.alloc(n_seg * HC * 4), "gat_edge_fwd_host");
  CU_OK(smax.alloc((size_t)n_seg * heads * 4), "gat_edge_fwd_host");
  CU_OK(ssum.alloc((size_t)n_seg * heads * 4), "gat_edge_fwd_host");
  CU_OK(cudaMemcpy(idx.p, pairs.data(), pairs.size() * sizeof(int64_t), cudaMemcpyHostToDevice), "gat_edge_fwd_host");
  CU_OK(cudaMemcpy(xl.p, XL_host, n_obs * HC * 4, cudaMemcpyHostToDevice), "gat_edge_fwd_host");
  CU_OK(cudaMemcpy(xr.p, XR_host, n_seg * HC * 4, cudaMemcpyHostToDevice), "gat_edge_fwd_host");
  CU_OK(cudaMemcpy(att.p, att_host, HC * 4, cudaMemcpyHostToDevice), "gat_edge_fwd_host");
  if (bias_host) CU_OK(cudaMemcpy(bias.p, bias_host, HC * 4, cudaMemcpyHostToDevice), "gat_edge_fwd_host");
  GASFM_REQUIRE(n_obs < (int64_t)INT32_MAX, "gat_edge_fwd_host: too many edges");
  int rc = gasfm_csr_build(idx.as<int64_t>(), n_obs, (int)n_obs, n_seg, ri.as<int32_t>(), ci.as<int32_t>(),
                           rp.as<int32_t>(), cp.as<int32_t>(), perm.as<int32_t>(), status.as<int32_t>(), csr_ws.p, nullptr);
  if (rc) return rc;
  int32_t st = 0;
  CU_OK(cudaMemcpy(&st, status.p, 4, cudaMemcpyDeviceToHost), "gat_edge_fwd_host");
  GASFM_REQUIRE(st == 0, "gat_edge_fwd_host: target ids out of range");
  int chunk = 0, max_chunks = 0;
  if (n_obs / n_seg > 64) {
    chunk = 128;
    max_chunks = (int)(n_obs / chunk) + n_seg + 1;
    CU_OK(cptr.alloc((size_t)(n_seg + 1) * 4), "gat_edge_fwd_host");
    CU_OK(cseg.alloc((size_t)max_chunks * 4), "gat_edge_fwd_host");
    CU_OK(ws.alloc(gasfm_gat_ws_bytes(max_chunks, heads, head_dim)), "gat_edge_fwd_host");
    rc = gasfm_plan_chunks(cp.as<int32_t>(), n_seg, chunk, cptr.as<int32_t>(), cseg.as<int32_t>(), max_chunks, nullptr);
    if (rc) return rc;
  }
  rc = gasfm_gat_edge_fwd(xl.as<float>(), (int64_t)HC, xr.as<float>(), (int64_t)HC, att.as<float>(),
                          bias_host ? bias.as<float>() : nullptr, cp.as<int32_t>(), perm.as<int32_t>(), n_seg, chunk,
                          cptr.as<int32_t>(), cseg.as<int32_t>(), max_chunks, heads, head_dim, slope, 1,
                          out.as<float>(), smax.as<float>(), ssum.as<float>(), ws.p, nullptr);
  if (rc) return rc;
  CU_OK(cudaMemcpy(out_host, out.p, n_seg * HC * 4, cudaMemcpyDeviceToHost), "gat_edge_fwd_host");
  return 0;
}
