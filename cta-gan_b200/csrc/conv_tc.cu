// tcgen05 / TMEM / TMA implicit-GEMM convolution engine (placeholder until the kernels land: nothing is eligible).
#include "common.cuh"

int ctagan_conv_gather_tc_eligible(const ctagan_conv_geom *g) { (void)g; return 0; }
int ctagan_conv_wgrad_tc_eligible(const ctagan_conv_geom *g) { (void)g; return 0; }
int ctagan_conv_gather_tc(const ctagan_conv_geom *g, const void *x, const void *wp, const float *bias, void *y, cudaStream_t st) {
  (void)g; (void)x; (void)wp; (void)bias; (void)y; (void)st;
  ctagan_set_error("conv_gather: geometry not supported by the tcgen05 engine");
  return CTAGAN_ERR_UNSUPPORTED;
}
int ctagan_conv_wgrad_tc(const ctagan_conv_geom *g, const void *gy, const void *gx, float *dw, float *db, cudaStream_t st) {
  (void)g; (void)gy; (void)gx; (void)dw; (void)db; (void)st;
  ctagan_set_error("conv_wgrad: geometry not supported by the tcgen05 engine");
  return CTAGAN_ERR_UNSUPPORTED;
}
