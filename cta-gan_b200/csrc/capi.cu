#include <stdlib.h>
// C-ABI glue: error state, device checks and engine dispatch for the convolution entry points.
#include <stdarg.h>
#include "common.cuh"

int ctagan_conv_gather_simt(const ctagan_conv_geom *g, const void *x, const void *wp, const float *bias, void *y, cudaStream_t st);
int ctagan_conv_wgrad_simt(const ctagan_conv_geom *g, const void *gy, const void *gx, float *dw, float *db, void *workspace,
                           size_t workspace_bytes, cudaStream_t st, int accumulate = 0);
size_t ctagan_conv_wgrad_simt_workspace(const ctagan_conv_geom *g);
// tcgen05 engine (conv_tc.cu): return CTAGAN_ERR_UNSUPPORTED when the geometry does not tile
int ctagan_conv_gather_tc(const ctagan_conv_geom *g, const void *x, const void *wp, const float *bias, void *y, unsigned int *stat_ticket,
                          void *stat_scratch, size_t stat_scratch_bytes, float *stat_out, cudaStream_t st,
                          const ctagan_conv_groups *gr = nullptr);
size_t ctagan_conv_gather_tc_stat_bytes(const ctagan_conv_geom *g);
int ctagan_conv_wgrad_tc(const ctagan_conv_geom *g, const void *gy, const void *gx, float *dw, float *db, void *workspace,
                         size_t workspace_bytes, cudaStream_t st, int n_groups = 1, int accumulate = 0);
size_t ctagan_conv_wgrad_tc_workspace(const ctagan_conv_geom *g, int n_groups = 1);
int ctagan_conv_gather_tc_eligible(const ctagan_conv_geom *g);
// degenerate (1-2 channel) convolutions (conv_small.cu)
int ctagan_conv_small_kind(const ctagan_conv_geom *g);
int ctagan_conv_gather_small(const ctagan_conv_geom *g, const void *x, const void *wp, const float *bias, void *y, cudaStream_t st);
int ctagan_conv_wgrad_thin_eligible(const ctagan_conv_geom *g);
int ctagan_conv_wgrad_thin(const ctagan_conv_geom *g, const void *gy, const void *gx, float *dw, float *db, void *workspace,
                           size_t workspace_bytes, cudaStream_t st, int accumulate = 0);
size_t ctagan_conv_wgrad_thin_workspace(const ctagan_conv_geom *g);
int ctagan_conv_wgrad_tc_eligible(const ctagan_conv_geom *g, int n_groups = 1);
int ctagan_conv_fewin_tc_eligible(const ctagan_conv_geom *g);
size_t ctagan_conv_fewin_tc_stat_bytes(const ctagan_conv_geom *g);
int ctagan_conv_fewin_tc(const ctagan_conv_geom *g, const void *x, const void *wp, const float *bias, void *y, unsigned int *stat_ticket,
                         void *stat_scratch, size_t stat_scratch_bytes, float *stat_out, cudaStream_t st);
int ctagan_conv_fewout_tc_eligible(const ctagan_conv_geom *g);
size_t ctagan_conv_fewout_tc_workspace(const ctagan_conv_geom *g);
int ctagan_conv_fewout_tc(const ctagan_conv_geom *g, const void *x, const void *wp, const float *bias, void *y, void *workspace,
                          size_t workspace_bytes, cudaStream_t st);
int ctagan_conv_wgrad_thin_tc_eligible(const ctagan_conv_geom *g);
size_t ctagan_conv_wgrad_thin_tc_workspace(const ctagan_conv_geom *g);
int ctagan_conv_wgrad_thin_tc(const ctagan_conv_geom *g, const void *gy, const void *gx, float *dw, float *db, void *workspace,
                              size_t workspace_bytes, cudaStream_t st, int accumulate);

static thread_local char g_err[512] = "";

void ctagan_set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int ctagan_num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  }
  return sms;
}

namespace {
template <bool ACC>
__global__ void __launch_bounds__(256) ordered_sum_kernel(const float *__restrict__ part, float *__restrict__ out, int parts, long long n,
                                                          long long row_stride) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int k = 0; k < parts; ++k) s += __ldcg(part + (long long)k * row_stride + i);
    if (ACC) s += out[i];
    out[i] = s;
  }
}
}  // namespace

// Many rows, few columns (the per-CTA partial rows of the thin-layer weight gradients: ~148 rows of a few thousand floats): one WARP per
// output element -- lane l adds rows l, l + 32, ... in order, then the lanes are combined by a fixed shuffle tree -- so the row loop is 5
// independent loads deep instead of 148 dependent-latency steps.  A fixed order again: the same inputs give the same bits.  The row may
// hold two results ([dw | db]) that go to two destinations.
template <bool ACC>
__global__ void __launch_bounds__(256) ordered_sum_warp_kernel(const float *__restrict__ part, float *__restrict__ out1, long long n1,
                                                               float *__restrict__ out2, long long n2, int parts, long long row_stride) {
  const int lane = threadIdx.x & 31;
  const long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long i = wid; i < n1 + n2; i += nw) {
    float s = 0.f;
    for (int k = lane; k < parts; k += 32) s += __ldcg(part + (long long)k * row_stride + i);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if (lane == 0) {
      float *dst = i < n1 ? out1 + i : out2 + (i - n1);
      if (ACC) s += *dst;
      *dst = s;
    }
  }
}

// rows [dw (n1) | db (n2)] `row_stride` floats apart -> out1[n1], out2[n2] (out2 may be NULL with n2 = 0)
int ctagan_ordered_sum_rows2(const float *part, float *out1, long long n1, float *out2, long long n2, int parts, long long row_stride,
                             cudaStream_t st, int accumulate) {
  const long long warps = n1 + n2;
  long long blocks = (warps + 7) / 8;
  if (blocks > 16LL * ctagan_num_sms()) blocks = 16LL * ctagan_num_sms();
  if (accumulate) ordered_sum_warp_kernel<true><<<(int)blocks, 256, 0, st>>>(part, out1, n1, out2, n2, parts, row_stride);
  else ordered_sum_warp_kernel<false><<<(int)blocks, 256, 0, st>>>(part, out1, n1, out2, n2, parts, row_stride);
  CTAGAN_LAUNCH_OK();
  return CTAGAN_OK;
}

int ctagan_ordered_sum_strided(const float *part, float *out, int parts, long long n, long long row_stride, cudaStream_t st, int accumulate) {
  long long blocks = (n + 255) / 256;
  if (blocks > 8LL * ctagan_num_sms()) blocks = 8LL * ctagan_num_sms();
  if (accumulate) ordered_sum_kernel<true><<<(int)blocks, 256, 0, st>>>(part, out, parts, n, row_stride);
  else ordered_sum_kernel<false><<<(int)blocks, 256, 0, st>>>(part, out, parts, n, row_stride);
  CTAGAN_LAUNCH_OK();
  return CTAGAN_OK;
}

int ctagan_ordered_sum(const float *part, float *out, int parts, long long n, cudaStream_t st, int accumulate) {
  return ctagan_ordered_sum_strided(part, out, parts, n, n, st, accumulate);
}

bool ctagan_pdl_enabled() {
  // measured on the Cyc step (batch 1): trigger at kernel start 6.87 ms, trigger at the epilogue 6.16-6.21 ms, off 6.20-6.22 ms --
  // the dependent-launch gap is not what bounds the chains, so the attribute stays opt-in
  const char *e = getenv("CTAGAN_PDL");
  return e && e[0] == '1';
}

extern "C" int ctagan_stream_create(int priority, void **stream_out) {
  CTAGAN_REQUIRE(stream_out, "stream_create: null pointer");
  int lo = 0, hi = 0;
  CTAGAN_CUDA_OK(cudaDeviceGetStreamPriorityRange(&lo, &hi));      // lo = least priority (numerically largest)
  int prio = priority < hi ? hi : (priority > lo ? lo : priority);
  cudaStream_t st;
  CTAGAN_CUDA_OK(cudaStreamCreateWithPriority(&st, cudaStreamNonBlocking, prio));
  *stream_out = (void *)st;
  return CTAGAN_OK;
}

extern "C" int ctagan_stream_destroy(void *stream) {
  CTAGAN_REQUIRE(stream, "stream_destroy: null pointer");
  CTAGAN_CUDA_OK(cudaStreamDestroy((cudaStream_t)stream));
  return CTAGAN_OK;
}

extern "C" int ctagan_version(void) { return 100; }
extern "C" const char *ctagan_last_error(void) { return g_err; }

extern "C" int ctagan_check_device(int dev) {
  int major = 0, minor = 0;
  CTAGAN_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  CTAGAN_CUDA_OK(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  if (major != 10) {
    ctagan_set_error("libctagan is built for sm_100a only; device %d is sm_%d%d (no fallback path)", dev, major, minor);
    return CTAGAN_ERR_UNSUPPORTED;
  }
  return CTAGAN_OK;
}

static int check_geom(const ctagan_conv_geom *g, const char *who) {
  CTAGAN_REQUIRE(g, "%s: null geometry", who);
  CTAGAN_REQUIRE(g->N > 0 && g->Hi > 0 && g->Wi > 0 && g->Ci > 0 && g->Ho > 0 && g->Wo > 0 && g->Co > 0, "%s: non-positive extent", who);
  CTAGAN_REQUIRE(g->KH > 0 && g->KW > 0 && g->stride > 0 && g->dil > 0, "%s: bad kernel/stride/dilation", who);
  CTAGAN_REQUIRE(g->dtype == CTAGAN_F32 || g->dtype == CTAGAN_BF16, "%s: bad dtype", who);
  CTAGAN_REQUIRE(g->act >= 0 && g->act <= 3, "%s: bad activation", who);
  CTAGAN_REQUIRE(g->gy_margin >= 0, "%s: bad gy_margin", who);
  return CTAGAN_OK;
}

extern "C" int ctagan_conv_gather(const ctagan_conv_geom *g, const void *x, const void *wp, const float *bias, void *y, int engine,
                                  void *stream) {
  int rc = check_geom(g, "conv_gather");
  if (rc) return rc;
  CTAGAN_REQUIRE(x && wp && y, "conv_gather: null pointer");
  CTAGAN_REQUIRE(engine >= 0 && engine <= 3, "conv_gather: bad engine");
  cudaStream_t st = (cudaStream_t)stream;
  if (engine == 2) return ctagan_conv_gather_tc(g, x, wp, bias, y, nullptr, nullptr, 0, nullptr, st);
  if ((engine == 0 || engine == 1) && ctagan_conv_fewin_tc_eligible(g)) return ctagan_conv_fewin_tc(g, x, wp, bias, y, nullptr, nullptr, 0, nullptr, st);
  if (engine != 3 && ctagan_conv_small_kind(g)) return ctagan_conv_gather_small(g, x, wp, bias, y, st);
  if (engine == 0 && ctagan_conv_gather_tc_eligible(g)) return ctagan_conv_gather_tc(g, x, wp, bias, y, nullptr, nullptr, 0, nullptr, st);
  return ctagan_conv_gather_simt(g, x, wp, bias, y, st);
}

extern "C" size_t ctagan_conv_gather_workspace_bytes(const ctagan_conv_geom *g, int engine) {
  if (!g || !(engine == 0 || engine == 1)) return 0;
  return ctagan_conv_fewout_tc_eligible(g) ? ctagan_conv_fewout_tc_workspace(g) : 0;
}

extern "C" int ctagan_conv_gather_ws(const ctagan_conv_geom *g, const void *x, const void *wp, const float *bias, void *y, void *workspace,
                                     size_t workspace_bytes, int engine, void *stream) {
  if (g && workspace && (engine == 0 || engine == 1) && ctagan_conv_fewout_tc_eligible(g)) {
    int rc = check_geom(g, "conv_gather_ws");
    if (rc) return rc;
    CTAGAN_REQUIRE(x && wp && y, "conv_gather_ws: null pointer");
    return ctagan_conv_fewout_tc(g, x, wp, bias, y, workspace, workspace_bytes, (cudaStream_t)stream);
  }
  return ctagan_conv_gather(g, x, wp, bias, y, engine, stream);
}

extern "C" size_t ctagan_conv_wgrad_workspace_bytes(const ctagan_conv_geom *g, int engine) {
  if (!g) return 0;
  if (engine == 2) return ctagan_conv_wgrad_tc_workspace(g);
  if ((engine == 0 || engine == 1) && ctagan_conv_wgrad_thin_tc_eligible(g)) return ctagan_conv_wgrad_thin_tc_workspace(g);
  if (engine != 3 && ctagan_conv_wgrad_thin_eligible(g)) return ctagan_conv_wgrad_thin_workspace(g);
  if (engine == 0 && ctagan_conv_wgrad_tc_eligible(g)) return ctagan_conv_wgrad_tc_workspace(g);
  return ctagan_conv_wgrad_simt_workspace(g);
}

extern "C" int ctagan_conv_wgrad(const ctagan_conv_geom *g, const void *gy, const void *gx, float *dw, float *db, void *workspace,
                                 size_t workspace_bytes, int engine, int accumulate, void *stream) {
  int rc = check_geom(g, "conv_wgrad");
  if (rc) return rc;
  CTAGAN_REQUIRE(gy && gx && dw, "conv_wgrad: null pointer");
  CTAGAN_REQUIRE(engine >= 0 && engine <= 3, "conv_wgrad: bad engine");
  cudaStream_t st = (cudaStream_t)stream;
  CTAGAN_REQUIRE((accumulate & ~(CTAGAN_WGRAD_ACCUMULATE | CTAGAN_WGRAD_PACKED)) == 0, "conv_wgrad: bad flags");
  if (engine == 2) return ctagan_conv_wgrad_tc(g, gy, gx, dw, db, workspace, workspace_bytes, st, 1, accumulate);
  // 1-2 channel layers: the patch-matrix tensor-core kernel where it applies (bf16, large maps), else the CUDA-core kernels
  if ((engine == 0 || engine == 1) && !(accumulate & CTAGAN_WGRAD_PACKED) && ctagan_conv_wgrad_thin_tc_eligible(g))
    return ctagan_conv_wgrad_thin_tc(g, gy, gx, dw, db, workspace, workspace_bytes, st, accumulate);
  if (engine != 3 && ctagan_conv_wgrad_thin_eligible(g)) {
    if (accumulate & CTAGAN_WGRAD_PACKED) {
      ctagan_set_error("conv_wgrad: the packed gradient layout is not available for 1-2 channel layers");
      return CTAGAN_ERR_UNSUPPORTED;
    }
    return ctagan_conv_wgrad_thin(g, gy, gx, dw, db, workspace, workspace_bytes, st, accumulate);
  }
  if (engine == 0 && ctagan_conv_wgrad_tc_eligible(g)) return ctagan_conv_wgrad_tc(g, gy, gx, dw, db, workspace, workspace_bytes, st, 1, accumulate);
  return ctagan_conv_wgrad_simt(g, gy, gx, dw, db, workspace, workspace_bytes, st, accumulate);
}

// Which engine ctagan_conv_gather would run for this geometry: 1 CUDA-core generic, 2 tcgen05, 4 CUDA-core specialised (1-2 channels)
extern "C" int ctagan_conv_gather_engine(const ctagan_conv_geom *g, int engine) {
  if (!g) return 0;
  if (engine == 2) return 2;
  if ((engine == 0 || engine == 1) && ctagan_conv_fewin_tc_eligible(g)) return 2;        // patch-matrix tensor-core kernel of the 1-2 input channel layers
  if (engine != 3 && ctagan_conv_small_kind(g)) return 4;
  if (engine == 0 && ctagan_conv_gather_tc_eligible(g)) return 2;
  return 1;
}

// ---- grouped launches: the batch is gr->groups consecutive image groups, group k convolved with the weights in slot gr->slot[k] of a
// packed buffer [slots][Co][taps][Ci]; tcgen05 engine only (callers fall back to one launch per group for other geometries) ----
extern "C" int ctagan_conv_gather_grouped_supported(const ctagan_conv_geom *g, const ctagan_conv_groups *gr) {
  if (!g || !gr || gr->groups < 1 || gr->groups > CTAGAN_MAX_GROUPS || g->N % gr->groups) return 0;
  return ctagan_conv_gather_tc_eligible(g) ? 1 : 0;
}

extern "C" int ctagan_conv_gather_grouped(const ctagan_conv_geom *g, const ctagan_conv_groups *gr, const void *x, const void *wp,
                                          const float *bias, void *y, uint32_t *stat_tickets, void *stat_scratch, size_t stat_scratch_bytes,
                                          float *stats_out, void *stream) {
  int rc = check_geom(g, "conv_gather_grouped");
  if (rc) return rc;
  CTAGAN_REQUIRE(gr && x && wp && y, "conv_gather_grouped: null pointer");
  if (!ctagan_conv_gather_grouped_supported(g, gr)) {
    ctagan_set_error("conv_gather_grouped: geometry / grouping not supported by the tcgen05 engine");
    return CTAGAN_ERR_UNSUPPORTED;
  }
  return ctagan_conv_gather_tc(g, x, wp, bias, y, stat_tickets, stat_scratch, stat_scratch_bytes, stats_out, (cudaStream_t)stream, gr);
}

extern "C" size_t ctagan_conv_wgrad_grouped_workspace_bytes(const ctagan_conv_geom *g, int groups) {
  if (!g || groups < 1 || groups > CTAGAN_MAX_GROUPS) return 0;
  return ctagan_conv_wgrad_tc_workspace(g, groups);
}

extern "C" int ctagan_conv_wgrad_grouped(const ctagan_conv_geom *g, int groups, const void *gy, const void *gx, float *dw, float *db,
                                         void *workspace, size_t workspace_bytes, void *stream) {
  int rc = check_geom(g, "conv_wgrad_grouped");
  if (rc) return rc;
  CTAGAN_REQUIRE(gy && gx && dw && groups >= 1 && groups <= CTAGAN_MAX_GROUPS, "conv_wgrad_grouped: bad arguments");
  if (!ctagan_conv_wgrad_tc_eligible(g, groups)) {
    ctagan_set_error("conv_wgrad_grouped: geometry / grouping not supported by the tcgen05 engine");
    return CTAGAN_ERR_UNSUPPORTED;
  }
  return ctagan_conv_wgrad_tc(g, gy, gx, dw, db, workspace, workspace_bytes, (cudaStream_t)stream, groups);
}

extern "C" size_t ctagan_conv_gather_stats_scratch_bytes(const ctagan_conv_geom *g, int engine) {
  if (!g || ctagan_conv_gather_engine(g, engine) != 2) return 0;
  if (engine != 2 && ctagan_conv_fewin_tc_eligible(g)) return ctagan_conv_fewin_tc_stat_bytes(g);
  return ctagan_conv_gather_tc_stat_bytes(g);
}

extern "C" int ctagan_conv_gather_stats(const ctagan_conv_geom *g, const void *x, const void *wp, const float *bias, void *y,
                                        uint32_t *stat_tickets, void *stat_scratch, size_t stat_scratch_bytes, float *stats_out, int engine,
                                        void *stream) {
  int rc = check_geom(g, "conv_gather_stats");
  if (rc) return rc;
  CTAGAN_REQUIRE(x && wp && y && stat_tickets && stat_scratch && stats_out, "conv_gather_stats: null pointer");
  if (ctagan_conv_gather_engine(g, engine) != 2) {
    ctagan_set_error("conv_gather_stats: fused statistics need the tcgen05 engine (use ctagan_conv_gather + ctagan_instnorm_stats)");
    return CTAGAN_ERR_UNSUPPORTED;
  }
  if (engine != 2 && ctagan_conv_fewin_tc_eligible(g))
    return ctagan_conv_fewin_tc(g, x, wp, bias, y, stat_tickets, stat_scratch, stat_scratch_bytes, stats_out, (cudaStream_t)stream);
  return ctagan_conv_gather_tc(g, x, wp, bias, y, stat_tickets, stat_scratch, stat_scratch_bytes, stats_out, (cudaStream_t)stream);
}
