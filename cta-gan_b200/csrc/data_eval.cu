// The two sides of the training hot path that the reference runs on the CPU, one slice at a time:
//   f1  input pipeline   trainer/datasets.py:36-82 (HU -> [-1,1], 50/400 window), trainer/utils.py:13-36 (nearest Resize),
//                        trainer/CycTrainer.py:91-95 (RandomAffine, nearest, fill -1)          -> batched kernels on int16 slices
//   f2  evaluation       trainer/CycTrainer.py:34-57 (to_windowdata), :286-330 (0.3-threshold masks), :362-398 (MAE / PSNR / UQI),
//                        skimage compare_ssim (7x7 uniform window), :337-343 (int16 DICOM pixels) -> fused per-slice reductions
// All reductions are deterministic (per-block partial sums added in block order).  Arithmetic that the reference does in numpy float64
// is done in double here, so the results match it to double round-off.
#include "common.cuh"

namespace {

// ---------------------------------------------------------------------------------------------------------------------
// f1
// ---------------------------------------------------------------------------------------------------------------------

// read_dicom (datasets.py:74-82) / the `image2` half of read_ori_w (:62-65): v = raw + add; v < 0 -> 0; (v / 4095 - 0.5) / 0.5
__global__ void hu_to_unit_kernel(const short *__restrict__ raw, float *__restrict__ out, long long n, int add) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    int v = (int)raw[i] + add;
    if (v < 0) v = 0;
    out[i] = (float)(((double)v / 4095.0 - 0.5) / 0.5);
  }
}

// the display window of read_ori_w (datasets.py:45-56) and to_windowdata (CycTrainer.py:34-57):
//   trunc((hu - win_min) * 255 / (win_max - win_min)) clipped to [0, 255], / 255, (x - 0.5) / 0.5
__device__ __forceinline__ double window_value(double hu, double center, double width) {
  const double win_min = (2.0 * center - width) / 2.0 + 0.5, win_max = (2.0 * center + width) / 2.0 + 0.5;
  const double f = 255.0 / (win_max - win_min);
  double t = trunc((hu - win_min) * f);
  if (t > 255.0) t = 255.0;
  if (t < 0.0) t = 0.0;
  return (t / 255.0 - 0.5) / 0.5;
}

__global__ void hu_window_kernel(const short *__restrict__ hu, float *__restrict__ out, long long n, double center, double width) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = (float)window_value((double)hu[i], center, width);
}

// F.interpolate(x, size) with the default mode 'nearest' (trainer/utils.py:28): src = min(floor(dst * scale), in - 1), scale = in / out in fp32
__global__ void resize_nearest_kernel(const float *__restrict__ src, float *__restrict__ dst, int B, int Hs, int Ws, int Hd, int Wd) {
  const float sh = (float)Hs / (float)Hd, sw = (float)Ws / (float)Wd;
  const long long total = (long long)B * Hd * Wd;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % Wd);
    const long long r = i / Wd;
    const int y = (int)(r % Hd), b = (int)(r / Hd);
    const int ys = min((int)floorf((float)y * sh), Hs - 1), xs = min((int)floorf((float)x * sw), Ws - 1);
    dst[i] = src[((long long)b * Hs + ys) * Ws + xs];
  }
}

// RandomAffine's resampling (PIL Image.transform(AFFINE, NEAREST), libImaging/Geometry.c affine_fixed): 16.16 fixed point,
//   xin = (FIX(c + a/2 + b/2) + x*FIX(a) + y*FIX(b)) >> 16,  FIX(v) = floor(v*65536 + 0.5);  pixels that map outside get `fill`.
// m[b][6] = the INVERSE affine matrix (a, b, c, d, e, f) of image b, as torchvision hands it to PIL.
__global__ void affine_nearest_kernel(const float *__restrict__ src, float *__restrict__ dst, int B, int H, int W, const double *__restrict__ m,
                                      float fill) {
  const long long total = (long long)B * H * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % W);
    const long long r = i / W;
    const int y = (int)(r % H), b = (int)(r / H);
    const double *a = m + 6 * b;
    const long long a0 = (long long)floor(a[0] * 65536.0 + 0.5), a1 = (long long)floor(a[1] * 65536.0 + 0.5);
    const long long a3 = (long long)floor(a[3] * 65536.0 + 0.5), a4 = (long long)floor(a[4] * 65536.0 + 0.5);
    const long long a2 = (long long)floor((a[2] + a[0] * 0.5 + a[1] * 0.5) * 65536.0 + 0.5);
    const long long a5 = (long long)floor((a[5] + a[3] * 0.5 + a[4] * 0.5) * 65536.0 + 0.5);
    // PIL keeps these in 32-bit ints; for images below 32768 pixels a side the sums cannot overflow, so 64-bit gives the same bits
    const int xin = (int)((int)(a2 + (long long)y * a1 + (long long)x * a0) >> 16);
    const int yin = (int)((int)(a5 + (long long)y * a4 + (long long)x * a3) >> 16);
    float v = fill;
    if (xin >= 0 && xin < W && yin >= 0 && yin < H) v = src[((long long)b * H + yin) * W + xin];
    dst[i] = v;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// f2
// ---------------------------------------------------------------------------------------------------------------------

// The four images the reference's test() compares per slice (CycTrainer.py:286-316), from fake / real in [-1, 1]:
//   bw = (W(real) >= 0.3 ? 1 : -1)                     "b":  the windowed target, thresholded IN PLACE through its alias `bb` (:289-293)
//   cw = (W(fake) * bb >= 0.3 ? 1 : -1)                "c":  likewise through `cc` (:294-299);  bb, cc in {0, 1} are the masks
//   rm = real * bb (== 0 -> -1),  fm = fake * cc (== 0 -> -1)                                   (:312-318)
// with W = to_windowdata (unit -> HU: (x+1)*0.5*4095, == 0 -> -2000, - 1024, then the display window).
struct EvalPix {
  double bw, cw, rm, fm;
};

// to_windowdata on the fp32 arrays test() feeds it (`.cpu().numpy()` of fp32 tensors): numpy keeps every step in fp32 (Python scalars
// are weak), so the same fp32 steps are taken here -- the 0.3 masks then agree with the reference's pixel for pixel
__device__ __forceinline__ float unit_to_window(float x, double wc, double ww) {
  float v = __fmul_rn(__fmul_rn(__fadd_rn(x, 1.f), 0.5f), 4095.f);
  if (v == 0.f) v = -2000.f;
  v = __fsub_rn(v, 1024.f);
  const double win_min = (2.0 * wc - ww) / 2.0 + 0.5, win_max = (2.0 * wc + ww) / 2.0 + 0.5;
  float t = truncf(__fmul_rn(__fsub_rn(v, (float)win_min), (float)(255.0 / (win_max - win_min))));
  if (t > 255.f) t = 255.f;
  if (t < 0.f) t = 0.f;
  return __fdiv_rn(__fsub_rn(__fdiv_rn(t, 255.f), 0.5f), 0.5f);
}

__device__ __forceinline__ EvalPix eval_pixel(float fake, float real, double wc, double ww) {
  EvalPix e;
  const double bb = unit_to_window(real, wc, ww) >= 0.3f ? 1.0 : 0.0;
  e.bw = bb == 1.0 ? 1.0 : -1.0;
  const double cc = __fmul_rn(unit_to_window(fake, wc, ww), (float)bb) >= 0.3f ? 1.0 : 0.0;
  e.cw = cc == 1.0 ? 1.0 : -1.0;
  e.rm = (double)real * bb;
  if (e.rm == 0.0) e.rm = -1.0;
  e.fm = (double)fake * cc;
  if (e.fm == 0.0) e.fm = -1.0;
  return e;
}

constexpr int EV_Q = 10;        // per pair: n_valid, sum|d| valid, sum d^2 valid, sum|d| all, sum d^2 all, sum f, sum r, sum f^2, sum r^2, sum f*r
constexpr int EV_NQ = 2 * EV_Q;

__device__ __forceinline__ void eval_accumulate(double f, double r, double *q) {
  const double d = f - r, h = (f + 1.0) / 2.0 - (r + 1.0) / 2.0;
  if (r != -1.0) { q[0] += 1.0; q[1] += fabs(d); q[2] += h * h; }
  q[3] += fabs(d); q[4] += h * h;
  q[5] += f; q[6] += r; q[7] += f * f; q[8] += r * r; q[9] += f * r;
}

// grid (chunks, B): partial sums of the 20 quantities of slice b over the block's pixel range -> part[b][chunk][20]
__global__ void __launch_bounds__(256) eval_partial_kernel(const float *__restrict__ fake, const float *__restrict__ real, double *__restrict__ part,
                                                           int HW, double wc, double ww) {
  const int b = blockIdx.y;
  const int per = (HW + gridDim.x - 1) / gridDim.x;
  const int p0 = blockIdx.x * per, p1 = min(HW, p0 + per);
  double q[EV_NQ];
#pragma unroll
  for (int k = 0; k < EV_NQ; ++k) q[k] = 0.0;
  for (int p = p0 + threadIdx.x; p < p1; p += blockDim.x) {
    const EvalPix e = eval_pixel(fake[(long long)b * HW + p], real[(long long)b * HW + p], wc, ww);
    eval_accumulate(e.cw, e.bw, q);
    eval_accumulate(e.fm, e.rm, q + EV_Q);
  }
  __shared__ double sm[8][EV_NQ];
#pragma unroll
  for (int k = 0; k < EV_NQ; ++k) q[k] = warp_sum_d(q[k]);
  if ((threadIdx.x & 31) == 0)
    for (int k = 0; k < EV_NQ; ++k) sm[threadIdx.x >> 5][k] = q[k];
  __syncthreads();
  if (threadIdx.x < EV_NQ) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += sm[w][threadIdx.x];
    part[((long long)b * gridDim.x + blockIdx.x) * EV_NQ + threadIdx.x] = t;
  }
}

// SSIM map (skimage compare_ssim defaults: 7x7 uniform window, sample covariance, K1 = 0.01, K2 = 0.03, data range 2 for float images in
// [-1, 1]), summed over the interior pixels (the mean is taken over the image cropped by 3 on every side).  grid (tiles, B);
// part[b][tile][2] = (sum of S over the tile's interior pixels for the (cw, bw) pair, for the (fm, rm) pair).
constexpr int SS_T = 32, SS_R = 3, SS_S = SS_T + 2 * SS_R;

__global__ void __launch_bounds__(256) ssim_partial_kernel(const float *__restrict__ fake, const float *__restrict__ real, double *__restrict__ part,
                                                           int H, int W, double wc, double ww) {
  __shared__ float t_f[2][SS_S][SS_S + 1], t_r[2][SS_S][SS_S + 1];
  const int b = blockIdx.y;
  const int tiles_w = (W + SS_T - 1) / SS_T;
  const int ty0 = (blockIdx.x / tiles_w) * SS_T, tx0 = (blockIdx.x % tiles_w) * SS_T;
  for (int idx = threadIdx.x; idx < SS_S * SS_S; idx += blockDim.x) {
    const int r = idx / SS_S, c = idx - r * SS_S;
    const int y = ty0 - SS_R + r, x = tx0 - SS_R + c;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (y >= 0 && y < H && x >= 0 && x < W) {
      const long long o = ((long long)b * H + y) * W + x;
      const EvalPix e = eval_pixel(fake[o], real[o], wc, ww);
      v[0] = (float)e.cw; v[1] = (float)e.bw; v[2] = (float)e.fm; v[3] = (float)e.rm;       // (+-1 and fp32 inputs: exact in float)
    }
    t_f[0][r][c] = v[0]; t_r[0][r][c] = v[1]; t_f[1][r][c] = v[2]; t_r[1][r][c] = v[3];
  }
  __syncthreads();
  const double C1 = (0.01 * 2.0) * (0.01 * 2.0), C2 = (0.03 * 2.0) * (0.03 * 2.0), NP = 49.0, cov_norm = NP / (NP - 1.0);
  double acc[2] = {0.0, 0.0};
  for (int idx = threadIdx.x; idx < SS_T * SS_T; idx += blockDim.x) {
    const int r = idx / SS_T, c = idx - r * SS_T;
    const int y = ty0 + r, x = tx0 + c;
    if (y < SS_R || y >= H - SS_R || x < SS_R || x >= W - SS_R) continue;
#pragma unroll
    for (int pr = 0; pr < 2; ++pr) {
      double sf = 0, sr = 0, sff = 0, srr = 0, sfr = 0;
      for (int dy = 0; dy < 7; ++dy)
#pragma unroll
        for (int dx = 0; dx < 7; ++dx) {
          const double f = t_f[pr][r + dy][c + dx], g = t_r[pr][r + dy][c + dx];
          sf += f; sr += g; sff += f * f; srr += g * g; sfr += f * g;
        }
      const double ux = sf / NP, uy = sr / NP;
      const double vx = cov_norm * (sff / NP - ux * ux), vy = cov_norm * (srr / NP - uy * uy), vxy = cov_norm * (sfr / NP - ux * uy);
      acc[pr] += ((2 * ux * uy + C1) * (2 * vxy + C2)) / ((ux * ux + uy * uy + C1) * (vx + vy + C2));
    }
  }
  __shared__ double sm[8][2];
  acc[0] = warp_sum_d(acc[0]); acc[1] = warp_sum_d(acc[1]);
  if ((threadIdx.x & 31) == 0) { sm[threadIdx.x >> 5][0] = acc[0]; sm[threadIdx.x >> 5][1] = acc[1]; }
  __syncthreads();
  if (threadIdx.x < 2) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += sm[w][threadIdx.x];
    part[((long long)b * gridDim.x + blockIdx.x) * 2 + threadIdx.x] = t;
  }
}

// one block per slice: add the partial sums in block order, then MAE / PSNR / UQI (CycTrainer.py:362-398) and the SSIM means;
// out[b][8] = (MAEw, PSNRw, SSIMw, UQIw, MAE, PSNR, SSIM, UQI)
__global__ void eval_finalize_kernel(const double *__restrict__ part, int chunks, const double *__restrict__ spart, int tiles, int H, int W,
                                     double *__restrict__ out) {
  const int b = blockIdx.x;
  __shared__ double q[EV_NQ], ss[2];
  if (threadIdx.x < EV_NQ) {
    double t = 0.0;
    for (int k = 0; k < chunks; ++k) t += part[((long long)b * chunks + k) * EV_NQ + threadIdx.x];
    q[threadIdx.x] = t;
  } else if (threadIdx.x < EV_NQ + 2) {
    const int j = threadIdx.x - EV_NQ;
    double t = 0.0;
    for (int k = 0; k < tiles; ++k) t += spart[((long long)b * tiles + k) * 2 + j];
    ss[j] = t;
  }
  __syncthreads();
  if (threadIdx.x < 2) {
    const double *p = q + threadIdx.x * EV_Q;
    const double N = (double)H * W;
    double mae, mse;
    if (p[0] == 0.0) { mae = p[3] / N + 1e-10; mse = p[4] / N + 1e-10; }
    else { mae = p[1] / p[0]; mse = p[2] / p[0]; }
    const double psnr = mse < 1.0e-10 ? 100.0 : 20.0 * log10(1.0 / (sqrt(mse) + 1e-10));
    const double mf = p[5] / N, mr = p[6] / N;
    const double varf = (p[7] - N * mf * mf) / (N - 1.0), varr = (p[8] - N * mr * mr) / (N - 1.0), cov = (p[9] - N * mf * mr) / (N - 1.0);
    const double uqi = 4.0 * mf * mr * cov / ((mf * mf + mr * mr) * (varf + varr) + 1e-10);
    const double interior = (double)(H - 2 * SS_R) * (double)(W - 2 * SS_R);
    double *o = out + (long long)b * 8 + threadIdx.x * 4;
    o[0] = mae / 2.0; o[1] = psnr; o[2] = ss[threadIdx.x] / interior; o[3] = uqi;
  }
}

// (fake + 1) * 0.5 * 4095 -> int16 (numpy astype: truncation toward zero), CycTrainer.py:337-341
__global__ void to_dicom_i16_kernel(const float *__restrict__ x, short *__restrict__ out, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = (short)(long long)(((double)x[i] + 1.0) * 0.5 * 4095.0);
}

inline int blocks_for(long long n) {
  long long b = (n + 255) / 256;
  const long long cap = 16LL * ctagan_num_sms();
  return (int)(b > cap ? cap : (b < 1 ? 1 : b));
}

}  // namespace

extern "C" int ctagan_hu_to_unit(const int16_t *raw, float *out, int64_t n, int add, void *stream) {
  CTAGAN_REQUIRE(raw && out && n > 0, "hu_to_unit: bad arguments");
  hu_to_unit_kernel<<<blocks_for(n), 256, 0, (cudaStream_t)stream>>>(raw, out, n, add);
  CTAGAN_LAUNCH_OK();
  return CTAGAN_OK;
}

extern "C" int ctagan_hu_window(const int16_t *hu, float *out, int64_t n, double center, double width, void *stream) {
  CTAGAN_REQUIRE(hu && out && n > 0 && width > 0, "hu_window: bad arguments");
  hu_window_kernel<<<blocks_for(n), 256, 0, (cudaStream_t)stream>>>(hu, out, n, center, width);
  CTAGAN_LAUNCH_OK();
  return CTAGAN_OK;
}

extern "C" int ctagan_resize_nearest(const float *src, float *dst, int B, int Hs, int Ws, int Hd, int Wd, void *stream) {
  CTAGAN_REQUIRE(src && dst && B > 0 && Hs > 0 && Ws > 0 && Hd > 0 && Wd > 0, "resize_nearest: bad arguments");
  resize_nearest_kernel<<<blocks_for((long long)B * Hd * Wd), 256, 0, (cudaStream_t)stream>>>(src, dst, B, Hs, Ws, Hd, Wd);
  CTAGAN_LAUNCH_OK();
  return CTAGAN_OK;
}

extern "C" int ctagan_affine_nearest(const float *src, float *dst, int B, int H, int W, const double *inv_matrix, float fill, void *stream) {
  CTAGAN_REQUIRE(src && dst && inv_matrix && B > 0 && H > 0 && W > 0 && H < 32768 && W < 32768, "affine_nearest: bad arguments");
  affine_nearest_kernel<<<blocks_for((long long)B * H * W), 256, 0, (cudaStream_t)stream>>>(src, dst, B, H, W, inv_matrix, fill);
  CTAGAN_LAUNCH_OK();
  return CTAGAN_OK;
}

static void eval_grid(int H, int W, int &chunks, int &tiles) {
  const int HW = H * W;
  chunks = (HW + 8191) / 8192;
  if (chunks > 64) chunks = 64;
  tiles = ((H + SS_T - 1) / SS_T) * ((W + SS_T - 1) / SS_T);
}

extern "C" size_t ctagan_eval_metrics_scratch_doubles(int B, int H, int W) {
  if (B <= 0 || H <= 0 || W <= 0) return 0;
  int chunks, tiles;
  eval_grid(H, W, chunks, tiles);
  return (size_t)B * ((size_t)chunks * EV_NQ + (size_t)tiles * 2);
}

extern "C" int ctagan_eval_metrics(const float *fake, const float *real, double *out, double *scratch, int B, int H, int W, double wc, double ww,
                                   void *stream) {
  CTAGAN_REQUIRE(fake && real && out && scratch && B > 0 && H > 2 * SS_R && W > 2 * SS_R && ww > 0, "eval_metrics: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  int chunks, tiles;
  eval_grid(H, W, chunks, tiles);
  double *part = scratch, *spart = scratch + (size_t)B * chunks * EV_NQ;
  eval_partial_kernel<<<dim3(chunks, B), 256, 0, st>>>(fake, real, part, H * W, wc, ww);
  CTAGAN_LAUNCH_OK();
  ssim_partial_kernel<<<dim3(tiles, B), 256, 0, st>>>(fake, real, spart, H, W, wc, ww);
  CTAGAN_LAUNCH_OK();
  eval_finalize_kernel<<<B, 32, 0, st>>>(part, chunks, spart, tiles, H, W, out);
  CTAGAN_LAUNCH_OK();
  return CTAGAN_OK;
}

extern "C" int ctagan_to_dicom_i16(const float *x, int16_t *out, int64_t n, void *stream) {
  CTAGAN_REQUIRE(x && out && n > 0, "to_dicom_i16: bad arguments");
  to_dicom_i16_kernel<<<blocks_for(n), 256, 0, (cudaStream_t)stream>>>(x, out, n);
  CTAGAN_LAUNCH_OK();
  return CTAGAN_OK;
}
