// Registration warp (Transformer_2D == bilinear grid_sample, border padding, align_corners=True) and the fused
// single-pass loss reductions (L1, LSGAN-MSE vs constant, smoothness, masked L1).  All fp32 at the module boundary.
#include "common.cuh"

namespace {

// The reference normalises pixel coordinates to [-1,1] (trainer/transformer.py:24-25) and grid_sample un-normalises them
// again (ATen GridSampler.h:27-36); both are replayed in the same fp32 op order so that floor() picks the same cell.
struct SampleCoord {
  float c;      // clipped source coordinate
  float gmult;  // d(c)/d(flow): 0 when clipped (ATen clip_coordinates_set_grad), else (size-1)/2 * 2/(size-1) applied in ATen order
};

__device__ __forceinline__ float ref_coord(float pix_plus_flow, int size) {
  float t = pix_plus_flow / (float)(size - 1);
  t = 2.f * (t - 0.5f);
  t = ((t + 1.f) / 2.f) * (float)(size - 1);
  return t;
}

__device__ __forceinline__ float clip_coord(float c, int size, float &g) {
  if (c <= 0.f) { g = 0.f; return 0.f; }
  const float mx = (float)(size - 1);
  if (c >= mx) { g = 0.f; return mx; }
  g = 1.f;
  return c;
}

__global__ void __launch_bounds__(256) warp_fwd_kernel(const float *__restrict__ src, const float *__restrict__ flow,
                                                       float *__restrict__ out, int B, int C, int H, int W) {
  const long long HW = (long long)H * W, total = (long long)B * HW;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(idx / HW);
    const int p = (int)(idx - (long long)b * HW);
    const int i = p / W, j = p - i * W;
    const float fy = __ldg(flow + ((long long)b * 2 + 0) * HW + p), fx = __ldg(flow + ((long long)b * 2 + 1) * HW + p);
    float gdummy;
    const float iy = clip_coord(ref_coord((float)i + fy, H), H, gdummy);
    const float ix = clip_coord(ref_coord((float)j + fx, W), W, gdummy);
    const float fy0 = floorf(iy), fx0 = floorf(ix);
    const int y0 = (int)fy0, x0 = (int)fx0, y1 = y0 + 1, x1 = x0 + 1;
    const float wnw = (fx0 + 1.f - ix) * (fy0 + 1.f - iy), wne = (ix - fx0) * (fy0 + 1.f - iy);
    const float wsw = (fx0 + 1.f - ix) * (iy - fy0), wse = (ix - fx0) * (iy - fy0);
    const bool x1ok = x1 < W, y1ok = y1 < H;  // x0,y0 are always in range after clipping
    for (int c = 0; c < C; ++c) {
      const float *s = src + ((long long)b * C + c) * HW;
      float v = __ldg(s + (long long)y0 * W + x0) * wnw;
      if (x1ok) v += __ldg(s + (long long)y0 * W + x1) * wne;
      if (y1ok) v += __ldg(s + (long long)y1 * W + x0) * wsw;
      if (x1ok && y1ok) v += __ldg(s + (long long)y1 * W + x1) * wse;
      out[((long long)b * C + c) * HW + p] = v;
    }
  }
}

__global__ void __launch_bounds__(256) warp_bwd_kernel(const float *__restrict__ gout, const float *__restrict__ src,
                                                       const float *__restrict__ flow, float *__restrict__ gsrc,
                                                       float *__restrict__ gflow, int B, int C, int H, int W) {
  const long long HW = (long long)H * W, total = (long long)B * HW;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(idx / HW);
    const int p = (int)(idx - (long long)b * HW);
    const int i = p / W, j = p - i * W;
    const float fy = __ldg(flow + ((long long)b * 2 + 0) * HW + p), fx = __ldg(flow + ((long long)b * 2 + 1) * HW + p);
    float gmy, gmx;
    const float iy = clip_coord(ref_coord((float)i + fy, H), H, gmy);
    const float ix = clip_coord(ref_coord((float)j + fx, W), W, gmx);
    const float fy0 = floorf(iy), fx0 = floorf(ix);
    const int y0 = (int)fy0, x0 = (int)fx0, y1 = y0 + 1, x1 = x0 + 1;
    const float ax = fx0 + 1.f - ix, bx = ix - fx0, ay = fy0 + 1.f - iy, by = iy - fy0;
    const bool x1ok = x1 < W, y1ok = y1 < H;
    float gix = 0.f, giy = 0.f;
    for (int c = 0; c < C; ++c) {
      const long long plane = ((long long)b * C + c) * HW;
      const float g = __ldg(gout + plane + p);
      const float *s = src + plane;
      float *gs = gsrc ? gsrc + plane : nullptr;
      {
        const float v = __ldg(s + (long long)y0 * W + x0);
        gix -= v * ay * g; giy -= v * ax * g;
        if (gs) atomicAdd(gs + (long long)y0 * W + x0, ax * ay * g);
      }
      if (x1ok) {
        const float v = __ldg(s + (long long)y0 * W + x1);
        gix += v * ay * g; giy -= v * bx * g;
        if (gs) atomicAdd(gs + (long long)y0 * W + x1, bx * ay * g);
      }
      if (y1ok) {
        const float v = __ldg(s + (long long)y1 * W + x0);
        gix -= v * by * g; giy += v * ax * g;
        if (gs) atomicAdd(gs + (long long)y1 * W + x0, ax * by * g);
      }
      if (x1ok && y1ok) {
        const float v = __ldg(s + (long long)y1 * W + x1);
        gix += v * by * g; giy += v * bx * g;
        if (gs) atomicAdd(gs + (long long)y1 * W + x1, bx * by * g);
      }
    }
    if (gflow) {
      // ATen: grad_grid = gix * ((size-1)/2 * clip_grad); then autograd through 2*(x/(size-1) - 0.5): (g*2)/(size-1)
      const float gx = gix * (gmx * ((float)(W - 1) / 2.f));
      const float gy = giy * (gmy * ((float)(H - 1) / 2.f));
      gflow[((long long)b * 2 + 0) * HW + p] = (gy * 2.f) / (float)(H - 1);
      gflow[((long long)b * 2 + 1) * HW + p] = (gx * 2.f) / (float)(W - 1);
    }
  }
}

// ---- block reduction + "last block finalises" -----------------------------------------------------------------
// acc[0] = running sum, acc[1] = ticket counter (both zeroed by the host wrapper before launch).
__device__ __forceinline__ void block_finish(double local, double *acc, float *loss, double scale) {
  __shared__ double sm[32];
  local = warp_sum_d(local);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += sm[i];
    atomicAdd(acc, t);
    __threadfence();
    const double ticket = atomicAdd(acc + 1, 1.0);
    if (ticket == (double)(gridDim.x - 1)) {
      const double total = atomicAdd(acc, 0.0);
      *loss = (float)(total * scale);
    }
  }
}

__global__ void __launch_bounds__(256) l1_fwd_kernel(const float *__restrict__ a, const float *__restrict__ b, float *loss, double *acc,
                                                     long long n) {
  float s = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    s += fabsf(__ldg(a + i) - __ldg(b + i));
  block_finish((double)s, acc, loss, 1.0 / (double)n);
}

__global__ void l1_bwd_kernel(const float *__restrict__ a, const float *__restrict__ b, const float *__restrict__ gloss,
                              float *__restrict__ ga, long long n) {
  const float g = __ldg(gloss) / (float)n;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float d = a[i] - b[i];
    ga[i] = d > 0.f ? g : (d < 0.f ? -g : 0.f);
  }
}

__global__ void __launch_bounds__(256) mse_const_fwd_kernel(const float *__restrict__ p, float target, float *loss, double *acc,
                                                            long long n) {
  float s = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float d = p[i] - target;
    s = fmaf(d, d, s);
  }
  block_finish((double)s, acc, loss, 1.0 / (double)n);
}

__global__ void mse_const_bwd_kernel(const float *__restrict__ p, float target, const float *__restrict__ gloss, float *__restrict__ gp,
                                     long long n) {
  const float g = 2.f * __ldg(gloss) / (float)n;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    gp[i] = g * (p[i] - target);
}

// smoothness: sum_dy/(B*C*(H-1)*W) + sum_dx/(B*C*H*(W-1)); both terms folded into one accumulator with their divisors.
__global__ void __launch_bounds__(256) smooth_fwd_kernel(const float *__restrict__ f, float *loss, double *acc, int B, int C, int H, int W) {
  const long long total = (long long)B * C * H * W;
  const float inv_y = 1.f / ((float)B * C * (H - 1) * W), inv_x = 1.f / ((float)B * C * H * (W - 1));
  float sy = 0.f, sx = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int w = (int)(i % W);
    const int h = (int)((i / W) % H);
    const float v = f[i];
    if (h > 0) { const float d = v - f[i - W]; sy = fmaf(d, d, sy); }
    if (w > 0) { const float d = v - f[i - 1]; sx = fmaf(d, d, sx); }
  }
  block_finish((double)sy * (double)inv_y + (double)sx * (double)inv_x, acc, loss, 1.0);
}

__global__ void smooth_bwd_kernel(const float *__restrict__ f, const float *__restrict__ gloss, float *__restrict__ gf, int B, int C,
                                  int H, int W) {
  const long long total = (long long)B * C * H * W;
  const float g = __ldg(gloss);
  const float ky = 2.f * g / ((float)B * C * (H - 1) * W), kx = 2.f * g / ((float)B * C * H * (W - 1));
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int w = (int)(i % W);
    const int h = (int)((i / W) % H);
    const float v = f[i];
    float o = 0.f;
    if (h > 0) o += ky * (v - f[i - W]);
    if (h < H - 1) o -= ky * (f[i + W] - v);
    if (w > 0) o += kx * (v - f[i - 1]);
    if (w < W - 1) o -= kx * (f[i + 1] - v);
    gf[i] = o;
  }
}

__device__ __forceinline__ void masked_pair(float wv, float b1, float b2, float &sw, float &rb, float &gmask) {
  const float bb = b1 >= 0.3f ? 1.f : 0.f;
  rb = b2 * bb;
  if (rb == 0.f) rb = -1.f;
  sw = wv * bb;
  gmask = bb;
  if (sw == 0.f) { sw = -1.f; gmask = 0.f; }
}

__global__ void __launch_bounds__(256) masked_l1_fwd_kernel(const float *__restrict__ wv, const float *__restrict__ b1,
                                                            const float *__restrict__ b2, float *loss, double *acc, long long n) {
  float s = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float sw, rb, gm;
    masked_pair(wv[i], b1[i], b2[i], sw, rb, gm);
    s += fabsf(sw - rb);
  }
  block_finish((double)s, acc, loss, 1.0 / (double)n);
}

__global__ void masked_l1_bwd_kernel(const float *__restrict__ wv, const float *__restrict__ b1, const float *__restrict__ b2,
                                     const float *__restrict__ gloss, float *__restrict__ gw, long long n) {
  const float g = __ldg(gloss) / (float)n;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float sw, rb, gm;
    masked_pair(wv[i], b1[i], b2[i], sw, rb, gm);
    const float d = sw - rb;
    gw[i] = gm * (d > 0.f ? g : (d < 0.f ? -g : 0.f));
  }
}

inline int red_blocks(long long n) {
  long long b = (n + 256 * 8 - 1) / (256 * 8);
  const long long cap = (long long)ctagan_num_sms() * 4;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}
inline int ew_blocks2(long long n) {
  long long b = (n + 255) / 256;
  const long long cap = (long long)ctagan_num_sms() * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace

extern "C" int ctagan_warp_fwd(const float *src, const float *flow, float *out, int B, int C, int H, int W, void *stream) {
  CTAGAN_REQUIRE(src && flow && out && B > 0 && C > 0 && H > 1 && W > 1, "warp_fwd: bad arguments");
  warp_fwd_kernel<<<ew_blocks2((long long)B * H * W), 256, 0, (cudaStream_t)stream>>>(src, flow, out, B, C, H, W);
  CTAGAN_LAUNCH_OK();
  return CTAGAN_OK;
}

extern "C" int ctagan_warp_bwd(const float *gout, const float *src, const float *flow, float *gsrc, float *gflow, int B, int C, int H,
                               int W, void *stream) {
  CTAGAN_REQUIRE(gout && src && flow && (gsrc || gflow) && B > 0 && C > 0 && H > 1 && W > 1, "warp_bwd: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  if (gsrc) CTAGAN_CUDA_OK(cudaMemsetAsync(gsrc, 0, sizeof(float) * (size_t)B * C * H * W, st));
  warp_bwd_kernel<<<ew_blocks2((long long)B * H * W), 256, 0, st>>>(gout, src, flow, gsrc, gflow, B, C, H, W);
  CTAGAN_LAUNCH_OK();
  return CTAGAN_OK;
}

#define LOSS_PROLOGUE(name)                                                        \
  cudaStream_t st = (cudaStream_t)stream;                                          \
  CTAGAN_CUDA_OK(cudaMemsetAsync(acc, 0, sizeof(double) * 2, st));

extern "C" int ctagan_l1_fwd(const float *a, const float *b, float *loss, double *acc, int64_t n, void *stream) {
  CTAGAN_REQUIRE(a && b && loss && acc && n > 0, "l1_fwd: bad arguments");
  LOSS_PROLOGUE(l1);
  l1_fwd_kernel<<<red_blocks(n), 256, 0, st>>>(a, b, loss, acc, n);
  CTAGAN_LAUNCH_OK();
  return CTAGAN_OK;
}
extern "C" int ctagan_l1_bwd(const float *a, const float *b, const float *gloss, float *ga, int64_t n, void *stream) {
  CTAGAN_REQUIRE(a && b && gloss && ga && n > 0, "l1_bwd: bad arguments");
  l1_bwd_kernel<<<ew_blocks2(n), 256, 0, (cudaStream_t)stream>>>(a, b, gloss, ga, n);
  CTAGAN_LAUNCH_OK();
  return CTAGAN_OK;
}
extern "C" int ctagan_mse_const_fwd(const float *p, float target, float *loss, double *acc, int64_t n, void *stream) {
  CTAGAN_REQUIRE(p && loss && acc && n > 0, "mse_const_fwd: bad arguments");
  LOSS_PROLOGUE(mse);
  mse_const_fwd_kernel<<<red_blocks(n), 256, 0, st>>>(p, target, loss, acc, n);
  CTAGAN_LAUNCH_OK();
  return CTAGAN_OK;
}
extern "C" int ctagan_mse_const_bwd(const float *p, float target, const float *gloss, float *gp, int64_t n, void *stream) {
  CTAGAN_REQUIRE(p && gloss && gp && n > 0, "mse_const_bwd: bad arguments");
  mse_const_bwd_kernel<<<ew_blocks2(n), 256, 0, (cudaStream_t)stream>>>(p, target, gloss, gp, n);
  CTAGAN_LAUNCH_OK();
  return CTAGAN_OK;
}
extern "C" int ctagan_smooth_fwd(const float *flow, float *loss, double *acc, int B, int C, int H, int W, void *stream) {
  CTAGAN_REQUIRE(flow && loss && acc && B > 0 && C > 0 && H > 1 && W > 1, "smooth_fwd: bad arguments");
  LOSS_PROLOGUE(smooth);
  smooth_fwd_kernel<<<red_blocks((long long)B * C * H * W), 256, 0, st>>>(flow, loss, acc, B, C, H, W);
  CTAGAN_LAUNCH_OK();
  return CTAGAN_OK;
}
extern "C" int ctagan_smooth_bwd(const float *flow, const float *gloss, float *gflow, int B, int C, int H, int W, void *stream) {
  CTAGAN_REQUIRE(flow && gloss && gflow && B > 0 && C > 0 && H > 1 && W > 1, "smooth_bwd: bad arguments");
  smooth_bwd_kernel<<<ew_blocks2((long long)B * C * H * W), 256, 0, (cudaStream_t)stream>>>(flow, gloss, gflow, B, C, H, W);
  CTAGAN_LAUNCH_OK();
  return CTAGAN_OK;
}
extern "C" int ctagan_masked_l1_fwd(const float *warped, const float *b1, const float *b2, float *loss, double *acc, int64_t n,
                                    void *stream) {
  CTAGAN_REQUIRE(warped && b1 && b2 && loss && acc && n > 0, "masked_l1_fwd: bad arguments");
  LOSS_PROLOGUE(ml1);
  masked_l1_fwd_kernel<<<red_blocks(n), 256, 0, st>>>(warped, b1, b2, loss, acc, n);
  CTAGAN_LAUNCH_OK();
  return CTAGAN_OK;
}
extern "C" int ctagan_masked_l1_bwd(const float *warped, const float *b1, const float *b2, const float *gloss, float *gwarped, int64_t n,
                                    void *stream) {
  CTAGAN_REQUIRE(warped && b1 && b2 && gloss && gwarped && n > 0, "masked_l1_bwd: bad arguments");
  masked_l1_bwd_kernel<<<ew_blocks2(n), 256, 0, (cudaStream_t)stream>>>(warped, b1, b2, gloss, gwarped, n);
  CTAGAN_LAUNCH_OK();
  return CTAGAN_OK;
}
