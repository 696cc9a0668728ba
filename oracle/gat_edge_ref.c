/*
 * Plain-C, double-precision restatement of the GATv2 edge arithmetic (oracle, TEST-ONLY; see
 * oracle/__init__.py).  Independent of torch, so it also guards the torch restatement in
 * oracle/gatv2conv.py.  Follows the published GATv2Conv algorithm as the reference calls it
 * (/root/reference/code/models/layers.py:329-335, 426-432, 550-556, 566-572):
 *
 *   s[e,h]  = sum_c att[h,c] * leaky_relu(XL[e,h,c] + XR[t(e),h,c], slope)
 *   a[e,h]  = exp(s - max_t) / (sum_t exp(s - max_t) + 1e-16)      (softmax over edges of t)
 *   out[t]  = sum_e a[e,h] * XL[e,h,c] + bias
 *
 * and the hand-derived backward of SURVEY.md section 8(a).  Inputs are float32 arrays (what the
 * CUDA kernels consume); all arithmetic is double.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

static double lrelu(double z, double slope) { return z > 0 ? z : z * slope; }

/* target[e] in [0,T) for every edge, arbitrary order.  out[T*H*C], smax[T*H], ssum[T*H] (double). */
int gat_edge_fwd_ref(const float* XL, const float* XR, int xr_broadcast, const float* att, const float* bias,
                     const int64_t* target, int64_t E, int64_t T, int H, int C, double slope,
                     double* out, double* smax, double* ssum) {
  const int64_t HC = (int64_t)H * C;
  double* s = (double*)malloc(sizeof(double) * (size_t)(E > 0 ? E : 1) * H);
  if (!s) return 1;
  for (int64_t i = 0; i < T * H; ++i) { smax[i] = -INFINITY; ssum[i] = 0; }
  for (int64_t i = 0; i < T * HC; ++i) out[i] = 0;
  for (int64_t e = 0; e < E; ++e) {
    const int64_t t = target[e];
    const float* xr = XR + (xr_broadcast ? 0 : t * HC);
    for (int h = 0; h < H; ++h) {
      double acc = 0;
      for (int c = 0; c < C; ++c)
        acc += (double)att[h * C + c] * lrelu((double)XL[e * HC + h * C + c] + (double)xr[h * C + c], slope);
      s[e * H + h] = acc;
      if (acc > smax[t * H + h]) smax[t * H + h] = acc;
    }
  }
  for (int64_t e = 0; e < E; ++e) {
    const int64_t t = target[e];
    for (int h = 0; h < H; ++h) ssum[t * H + h] += exp(s[e * H + h] - smax[t * H + h]);
  }
  for (int64_t e = 0; e < E; ++e) {
    const int64_t t = target[e];
    for (int h = 0; h < H; ++h) {
      const double a = exp(s[e * H + h] - smax[t * H + h]) / (ssum[t * H + h] + 1e-16);
      for (int c = 0; c < C; ++c) out[t * HC + h * C + c] += a * (double)XL[e * HC + h * C + c];
    }
  }
  if (bias)
    for (int64_t t = 0; t < T; ++t)
      for (int64_t j = 0; j < HC; ++j) out[t * HC + j] += (double)bias[j];
  free(s);
  return 0;
}

/* Backward given dOut[T*HC] (float).  dXL[E*HC], dXR[T*HC], datt[HC], dbias[HC] (double). */
int gat_edge_bwd_ref(const float* XL, const float* XR, int xr_broadcast, const float* att,
                     const int64_t* target, const float* dOut, int64_t E, int64_t T, int H, int C, double slope,
                     double* dXL, double* dXR, double* datt, double* dbias) {
  const int64_t HC = (int64_t)H * C;
  double* out = (double*)calloc((size_t)(T * HC > 0 ? T * HC : 1), sizeof(double));
  double* smax = (double*)malloc(sizeof(double) * (size_t)(T * H > 0 ? T * H : 1));
  double* ssum = (double*)malloc(sizeof(double) * (size_t)(T * H > 0 ? T * H : 1));
  double* D = (double*)calloc((size_t)(T * H > 0 ? T * H : 1), sizeof(double));
  if (!out || !smax || !ssum || !D) return 1;
  gat_edge_fwd_ref(XL, XR, xr_broadcast, att, 0, target, E, T, H, C, slope, out, smax, ssum);
  for (int64_t t = 0; t < T; ++t)
    for (int h = 0; h < H; ++h)
      for (int c = 0; c < C; ++c) D[t * H + h] += (double)dOut[t * HC + h * C + c] * out[t * HC + h * C + c];
  for (int64_t i = 0; i < T * HC; ++i) dXR[i] = 0;
  for (int64_t j = 0; j < HC; ++j) { datt[j] = 0; dbias[j] = 0; }
  for (int64_t t = 0; t < T; ++t)
    for (int64_t j = 0; j < HC; ++j) dbias[j] += (double)dOut[t * HC + j];
  for (int64_t e = 0; e < E; ++e) {
    const int64_t t = target[e];
    const float* xr = XR + (xr_broadcast ? 0 : t * HC);
    for (int h = 0; h < H; ++h) {
      double sc = 0, da = 0;
      for (int c = 0; c < C; ++c) {
        const double x = XL[e * HC + h * C + c];
        sc += (double)att[h * C + c] * lrelu(x + (double)xr[h * C + c], slope);
        da += (double)dOut[t * HC + h * C + c] * x;
      }
      const double a = exp(sc - smax[t * H + h]) / (ssum[t * H + h] + 1e-16);
      const double ds = a * (da - D[t * H + h]);
      for (int c = 0; c < C; ++c) {
        const double z = (double)XL[e * HC + h * C + c] + (double)xr[h * C + c];
        const double dz = ds * (double)att[h * C + c] * (z > 0 ? 1.0 : slope);
        dXL[e * HC + h * C + c] = a * (double)dOut[t * HC + h * C + c] + dz;
        dXR[t * HC + h * C + c] += dz;
        datt[h * C + c] += ds * lrelu(z, slope);
      }
    }
  }
  free(out); free(smax); free(ssum); free(D);
  return 0;
}
