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

// Tiling shared by the forward and backward kernels: a block of 256 threads owns WT_H x WT_W output pixels (4 consecutive columns per
// thread: 16-byte loads of the flow / gout rows and 16-byte stores) and stages the source window
// [i0 - HALO, i0 + WT_H + HALO] x [j0 - HALO, j0 + WT_W + HALO] of the current channel plane in shared memory with coalesced row
// loads.  Registration flows are small (Reg is initialised to ~zero flow), so practically every bilinear tap is a shared-memory hit;
// a tap outside the window (a displacement of more than HALO pixels) falls back to a global load, so any flow is handled.
constexpr int WT_H = 16, WT_W = 64, WT_HALO = 8;
constexpr int WS_H = WT_H + 2 * WT_HALO + 1, WS_W = WT_W + 2 * WT_HALO + 1;      // window extent incl. the +1 neighbour

struct WarpTile {
  int b, i0, j0;         // image, tile origin
  int wi0, wj0;          // window origin in the image (may be negative: clipped on load)
};

__device__ __forceinline__ WarpTile warp_tile(int H, int W) {
  const int tiles_w = (W + WT_W - 1) / WT_W, tiles_h = (H + WT_H - 1) / WT_H;
  int t = blockIdx.x;
  WarpTile wt;
  wt.j0 = (t % tiles_w) * WT_W; t /= tiles_w;
  wt.i0 = (t % tiles_h) * WT_H;
  wt.b = t / tiles_h;
  wt.wi0 = wt.i0 - WT_HALO; wt.wj0 = wt.j0 - WT_HALO;
  return wt;
}

__device__ __forceinline__ void warp_load_window(const float *__restrict__ plane, float *__restrict__ win, const WarpTile &wt, int H, int W) {
  for (int idx = threadIdx.x; idx < WS_H * WS_W; idx += blockDim.x) {
    const int r = idx / WS_W, c = idx - r * WS_W;
    const int y = wt.wi0 + r, x = wt.wj0 + c;
    win[idx] = (y >= 0 && y < H && x >= 0 && x < W) ? __ldg(plane + (long long)y * W + x) : 0.f;
  }
}

__device__ __forceinline__ float warp_tap(const float *__restrict__ plane, const float *__restrict__ win, const WarpTile &wt, int y, int x,
                                          int W) {
  const int r = y - wt.wi0, c = x - wt.wj0;
  if ((unsigned)r < (unsigned)WS_H && (unsigned)c < (unsigned)WS_W) return win[r * WS_W + c];
  return __ldg(plane + (long long)y * W + x);
}

__global__ void __launch_bounds__(256) warp_fwd_kernel(const float *__restrict__ src, const float *__restrict__ flow,
                                                       float *__restrict__ out, int B, int C, int H, int W) {
  __shared__ float win[WS_H * WS_W];
  const WarpTile wt = warp_tile(H, W);
  const long long HW = (long long)H * W;
  const int i = wt.i0 + (int)(threadIdx.x >> 4), j = wt.j0 + (int)(threadIdx.x & 15) * 4;
  const bool row_ok = i < H;
  const bool vec = row_ok && ((W & 3) == 0) && j + 3 < W;
  float fy[4] = {0.f, 0.f, 0.f, 0.f}, fx[4] = {0.f, 0.f, 0.f, 0.f};
  const long long p = (long long)i * W + j;
  if (vec) {
    const float4 a = __ldg(reinterpret_cast<const float4 *>(flow + ((long long)wt.b * 2 + 0) * HW + p));
    const float4 c4 = __ldg(reinterpret_cast<const float4 *>(flow + ((long long)wt.b * 2 + 1) * HW + p));
    fy[0] = a.x; fy[1] = a.y; fy[2] = a.z; fy[3] = a.w;
    fx[0] = c4.x; fx[1] = c4.y; fx[2] = c4.z; fx[3] = c4.w;
  } else if (row_ok) {
    for (int e = 0; e < 4; ++e)
      if (j + e < W) {
        fy[e] = __ldg(flow + ((long long)wt.b * 2 + 0) * HW + p + e);
        fx[e] = __ldg(flow + ((long long)wt.b * 2 + 1) * HW + p + e);
      }
  }
  for (int c = 0; c < C; ++c) {
    const float *plane = src + ((long long)wt.b * C + c) * HW;
    __syncthreads();
    warp_load_window(plane, win, wt, H, W);
    __syncthreads();
    if (!row_ok) continue;
    float o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      o[e] = 0.f;
      if (j + e >= W) continue;
      float gdummy;
      const float iy = clip_coord(ref_coord((float)i + fy[e], H), H, gdummy);
      const float ix = clip_coord(ref_coord((float)(j + e) + fx[e], W), W, gdummy);
      const float fy0 = floorf(iy), fx0 = floorf(ix);
      const int y0 = (int)fy0, x0 = (int)fx0, y1 = y0 + 1, x1 = x0 + 1;
      const float wnw = (fx0 + 1.f - ix) * (fy0 + 1.f - iy), wne = (ix - fx0) * (fy0 + 1.f - iy);
      const float wsw = (fx0 + 1.f - ix) * (iy - fy0), wse = (ix - fx0) * (iy - fy0);
      const bool x1ok = x1 < W, y1ok = y1 < H;  // x0,y0 are always in range after clipping
      float v = warp_tap(plane, win, wt, y0, x0, W) * wnw;
      if (x1ok) v += warp_tap(plane, win, wt, y0, x1, W) * wne;
      if (y1ok) v += warp_tap(plane, win, wt, y1, x0, W) * wsw;
      if (x1ok && y1ok) v += warp_tap(plane, win, wt, y1, x1, W) * wse;
      o[e] = v;
    }
    float *dst = out + ((long long)wt.b * C + c) * HW + p;
    if (vec) *reinterpret_cast<float4 *>(dst) = make_float4(o[0], o[1], o[2], o[3]);
    else
      for (int e = 0; e < 4; ++e)
        if (j + e < W) dst[e] = o[e];
  }
}

// ---- backward ----------------------------------------------------------------------------------------------------------------
// gsrc is a scatter-add (several output pixels sample the same source pixel).  Floating-point atomics would make the result depend on
// the CTA schedule, so the contributions are accumulated in 64-bit FIXED POINT (integer addition is associative: any order gives the
// same bits): scale = 2^(40 - e) with 2^e > max|gout| (found by warp_absmax_kernel), i.e. every product w*g (|w| <= 1) is below 2^40
// and a source pixel can absorb 2^22 of them; the quantum is max|gout| * 2^-40, far below one fp32 ulp of any sum that matters.
// Contributions that land inside the block's window go to a shared-memory tile first (one global atomic per touched pixel and block).
__global__ void __launch_bounds__(256) warp_absmax_kernel(const float *__restrict__ g, unsigned int *__restrict__ out, long long n) {
  float m = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) m = fmaxf(m, fabsf(__ldg(g + i)));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f && !(m != m)) atomicMax(out, __float_as_uint(m));   // non-negative floats order like their bit patterns
}

__device__ __forceinline__ double warp_fixed_scale(unsigned int absmax_bits) {
  if (absmax_bits == 0u) return 1.0;
  const int e = (int)((absmax_bits >> 23) & 0xffu) - 127 + 1;       // max|gout| < 2^e   (denormals: e = -126)
  return ldexp(1.0, 40 - e);
}

__global__ void __launch_bounds__(256) warp_bwd_kernel(const float *__restrict__ gout, const float *__restrict__ src,
                                                       const float *__restrict__ flow, long long *__restrict__ gacc,
                                                       const unsigned int *__restrict__ absmax, float *__restrict__ gflow, int B, int C,
                                                       int H, int W) {
  __shared__ float win[WS_H * WS_W];
  __shared__ long long gwin[WS_H * WS_W];
  const WarpTile wt = warp_tile(H, W);
  const long long HW = (long long)H * W;
  const int i = wt.i0 + (int)(threadIdx.x >> 4), j = wt.j0 + (int)(threadIdx.x & 15) * 4;
  const bool row_ok = i < H;
  const bool vec = row_ok && ((W & 3) == 0) && j + 3 < W;
  const double scale = gacc ? warp_fixed_scale(*absmax) : 1.0;
  float fy[4] = {0.f, 0.f, 0.f, 0.f}, fx[4] = {0.f, 0.f, 0.f, 0.f};
  const long long p = (long long)i * W + j;
  if (vec) {
    const float4 a = __ldg(reinterpret_cast<const float4 *>(flow + ((long long)wt.b * 2 + 0) * HW + p));
    const float4 c4 = __ldg(reinterpret_cast<const float4 *>(flow + ((long long)wt.b * 2 + 1) * HW + p));
    fy[0] = a.x; fy[1] = a.y; fy[2] = a.z; fy[3] = a.w;
    fx[0] = c4.x; fx[1] = c4.y; fx[2] = c4.z; fx[3] = c4.w;
  } else if (row_ok) {
    for (int e = 0; e < 4; ++e)
      if (j + e < W) {
        fy[e] = __ldg(flow + ((long long)wt.b * 2 + 0) * HW + p + e);
        fx[e] = __ldg(flow + ((long long)wt.b * 2 + 1) * HW + p + e);
      }
  }
  float gix[4] = {0.f, 0.f, 0.f, 0.f}, giy[4] = {0.f, 0.f, 0.f, 0.f}, gmx[4] = {0.f, 0.f, 0.f, 0.f}, gmy[4] = {0.f, 0.f, 0.f, 0.f};
  for (int c = 0; c < C; ++c) {
    const long long plane_off = ((long long)wt.b * C + c) * HW;
    const float *plane = src + plane_off;
    __syncthreads();
    warp_load_window(plane, win, wt, H, W);
    if (gacc)
      for (int idx = threadIdx.x; idx < WS_H * WS_W; idx += blockDim.x) gwin[idx] = 0ll;
    __syncthreads();
    if (row_ok) {
      float g[4] = {0.f, 0.f, 0.f, 0.f};
      if (vec) {
        const float4 t = __ldg(reinterpret_cast<const float4 *>(gout + plane_off + p));
        g[0] = t.x; g[1] = t.y; g[2] = t.z; g[3] = t.w;
      } else {
        for (int e = 0; e < 4; ++e)
          if (j + e < W) g[e] = __ldg(gout + plane_off + p + e);
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (j + e >= W) continue;
        const float iy = clip_coord(ref_coord((float)i + fy[e], H), H, gmy[e]);
        const float ix = clip_coord(ref_coord((float)(j + e) + fx[e], W), W, gmx[e]);
        const float fy0 = floorf(iy), fx0 = floorf(ix);
        const int y0 = (int)fy0, x0 = (int)fx0, y1 = y0 + 1, x1 = x0 + 1;
        const float ax = fx0 + 1.f - ix, bx = ix - fx0, ay = fy0 + 1.f - iy, by = iy - fy0;
        const bool x1ok = x1 < W, y1ok = y1 < H;
        const float ge = g[e];
        auto scatter = [&](int y, int x, float v) {
          if (!gacc) return;
          const long long q = __double2ll_rn((double)v * scale);
          if (q == 0ll) return;
          const int r = y - wt.wi0, cc = x - wt.wj0;
          if ((unsigned)r < (unsigned)WS_H && (unsigned)cc < (unsigned)WS_W)
            atomicAdd(reinterpret_cast<unsigned long long *>(&gwin[r * WS_W + cc]), (unsigned long long)q);
          else
            atomicAdd(reinterpret_cast<unsigned long long *>(gacc + plane_off + (long long)y * W + x), (unsigned long long)q);
        };
        {
          const float v = warp_tap(plane, win, wt, y0, x0, W);
          gix[e] -= v * ay * ge; giy[e] -= v * ax * ge;
          scatter(y0, x0, ax * ay * ge);
        }
        if (x1ok) {
          const float v = warp_tap(plane, win, wt, y0, x1, W);
          gix[e] += v * ay * ge; giy[e] -= v * bx * ge;
          scatter(y0, x1, bx * ay * ge);
        }
        if (y1ok) {
          const float v = warp_tap(plane, win, wt, y1, x0, W);
          gix[e] -= v * by * ge; giy[e] += v * ax * ge;
          scatter(y1, x0, ax * by * ge);
        }
        if (x1ok && y1ok) {
          const float v = warp_tap(plane, win, wt, y1, x1, W);
          gix[e] += v * by * ge; giy[e] += v * bx * ge;
          scatter(y1, x1, bx * by * ge);
        }
      }
    }
    if (gacc) {
      __syncthreads();
      for (int idx = threadIdx.x; idx < WS_H * WS_W; idx += blockDim.x) {
        const long long q = gwin[idx];
        if (q == 0ll) continue;
        const int r = idx / WS_W, cc = idx - r * WS_W;
        atomicAdd(reinterpret_cast<unsigned long long *>(gacc + plane_off + (long long)(wt.wi0 + r) * W + wt.wj0 + cc), (unsigned long long)q);
      }
    }
  }
  if (gflow && row_ok) {
    // ATen: grad_grid = gix * ((size-1)/2 * clip_grad); then autograd through 2*(x/(size-1) - 0.5): (g*2)/(size-1)
    float oy[4], ox[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float gx = gix[e] * (gmx[e] * ((float)(W - 1) / 2.f));
      const float gy = giy[e] * (gmy[e] * ((float)(H - 1) / 2.f));
      oy[e] = (gy * 2.f) / (float)(H - 1);
      ox[e] = (gx * 2.f) / (float)(W - 1);
    }
    float *dy = gflow + ((long long)wt.b * 2 + 0) * HW + p, *dx = gflow + ((long long)wt.b * 2 + 1) * HW + p;
    if (vec) {
      *reinterpret_cast<float4 *>(dy) = make_float4(oy[0], oy[1], oy[2], oy[3]);
      *reinterpret_cast<float4 *>(dx) = make_float4(ox[0], ox[1], ox[2], ox[3]);
    } else {
      for (int e = 0; e < 4; ++e)
        if (j + e < W) { dy[e] = oy[e]; dx[e] = ox[e]; }
    }
  }
}

// fixed point -> fp32
__global__ void __launch_bounds__(256) warp_gsrc_finish_kernel(const long long *__restrict__ gacc, const unsigned int *__restrict__ absmax,
                                                               float *__restrict__ gsrc, long long n) {
  const double inv = 1.0 / warp_fixed_scale(*absmax);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    gsrc[i] = (float)((double)gacc[i] * inv);
}

// ---- block reduction + "last block finalises" -----------------------------------------------------------------
// acc[1] = ticket counter (zeroed by the host wrapper before launch), acc[2 + b] = partial sum of block b.  The last block to arrive
// adds the partial sums in block order: deterministic (no floating-point atomics on the sum), one launch, no host sync.
__device__ __forceinline__ void block_finish(double local, double *acc, float *loss, double scale) {
  __shared__ double sm[32];
  __shared__ bool last;
  local = warp_sum_d(local);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += sm[i];
    acc[2 + blockIdx.x] = t;
    __threadfence();
    const double ticket = atomicAdd(acc + 1, 1.0);
    last = ticket == (double)(gridDim.x - 1);
  }
  __syncthreads();
  if (last) {
    __threadfence();
    // 256 threads: strided partial sums in a fixed pattern, then a fixed-order tree
    double t = 0.0;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) t += __ldcg(acc + 2 + b);
    t = warp_sum_d(t);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
      double total = 0.0;
      for (int i = 0; i < (int)(blockDim.x >> 5); ++i) total += sm[i];
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

__global__ void __launch_bounds__(256) mse_const_fwd_kernel(const float *__restrict__ p, float target, const float *__restrict__ target_dev,
                                                            float *loss, double *acc, long long n) {
  if (target_dev) target = __ldg(target_dev);        // the constant lives on the device (a (1,1) target tensor): no host read
  float s = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float d = p[i] - target;
    s = fmaf(d, d, s);
  }
  block_finish((double)s, acc, loss, 1.0 / (double)n);
}

__global__ void mse_const_bwd_kernel(const float *__restrict__ p, float target, const float *__restrict__ target_dev,
                                     const float *__restrict__ gloss, float *__restrict__ gp, long long n) {
  if (target_dev) target = __ldg(target_dev);
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
  long long cap = (long long)ctagan_num_sms() * 4;
  if (cap > CTAGAN_LOSS_ACC_DOUBLES - 2) cap = CTAGAN_LOSS_ACC_DOUBLES - 2;
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

static int warp_tiles(int B, int H, int W) { return B * ((H + WT_H - 1) / WT_H) * ((W + WT_W - 1) / WT_W); }

extern "C" int ctagan_warp_fwd(const float *src, const float *flow, float *out, int B, int C, int H, int W, void *stream) {
  CTAGAN_REQUIRE(src && flow && out && B > 0 && C > 0 && H > 1 && W > 1, "warp_fwd: bad arguments");
  CTAGAN_REQUIRE(((reinterpret_cast<uintptr_t>(flow) | reinterpret_cast<uintptr_t>(out)) & 15) == 0, "warp_fwd: flow / out must be 16-byte aligned");
  warp_fwd_kernel<<<warp_tiles(B, H, W), 256, 0, (cudaStream_t)stream>>>(src, flow, out, B, C, H, W);
  CTAGAN_LAUNCH_OK();
  return CTAGAN_OK;
}

/* scratch of ctagan_warp_bwd when gsrc is wanted: the 64-bit fixed-point accumulator of the scatter-add + 16 bytes */
extern "C" size_t ctagan_warp_bwd_workspace_bytes(int B, int C, int H, int W) {
  if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return 0;
  return (size_t)B * C * H * W * sizeof(long long) + 16;
}

extern "C" int ctagan_warp_bwd(const float *gout, const float *src, const float *flow, float *gsrc, float *gflow, void *workspace,
                               size_t workspace_bytes, int B, int C, int H, int W, void *stream) {
  CTAGAN_REQUIRE(gout && src && flow && (gsrc || gflow) && B > 0 && C > 0 && H > 1 && W > 1, "warp_bwd: bad arguments");
  CTAGAN_REQUIRE(((reinterpret_cast<uintptr_t>(flow) | reinterpret_cast<uintptr_t>(gout) | reinterpret_cast<uintptr_t>(gflow)) & 15) == 0,
                 "warp_bwd: flow / gout / gflow must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const long long n = (long long)B * C * H * W;
  long long *gacc = nullptr;
  unsigned int *absmax = nullptr;
  if (gsrc) {
    const size_t need = ctagan_warp_bwd_workspace_bytes(B, C, H, W);
    CTAGAN_REQUIRE(workspace && workspace_bytes >= need && (reinterpret_cast<uintptr_t>(workspace) & 7) == 0,
                   "warp_bwd: 8-byte aligned workspace of %zu bytes required (got %zu)", need, workspace_bytes);
    gacc = (long long *)workspace;
    absmax = reinterpret_cast<unsigned int *>(gacc + n);
    CTAGAN_CUDA_OK(cudaMemsetAsync(workspace, 0, need, st));
    warp_absmax_kernel<<<red_blocks(n), 256, 0, st>>>(gout, absmax, n);
    CTAGAN_LAUNCH_OK();
  }
  warp_bwd_kernel<<<warp_tiles(B, H, W), 256, 0, st>>>(gout, src, flow, gacc, absmax, gflow, B, C, H, W);
  CTAGAN_LAUNCH_OK();
  if (gsrc) {
    warp_gsrc_finish_kernel<<<ew_blocks2(n), 256, 0, st>>>(gacc, absmax, gsrc, n);
    CTAGAN_LAUNCH_OK();
  }
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
extern "C" int ctagan_mse_const_fwd(const float *p, float target, const float *target_dev, float *loss, double *acc, int64_t n, void *stream) {
  CTAGAN_REQUIRE(p && loss && acc && n > 0, "mse_const_fwd: bad arguments");
  LOSS_PROLOGUE(mse);
  mse_const_fwd_kernel<<<red_blocks(n), 256, 0, st>>>(p, target, target_dev, loss, acc, n);
  CTAGAN_LAUNCH_OK();
  return CTAGAN_OK;
}
extern "C" int ctagan_mse_const_bwd(const float *p, float target, const float *target_dev, const float *gloss, float *gp, int64_t n,
                                    void *stream) {
  CTAGAN_REQUIRE(p && gloss && gp && n > 0, "mse_const_bwd: bad arguments");
  mse_const_bwd_kernel<<<ew_blocks2(n), 256, 0, (cudaStream_t)stream>>>(p, target, target_dev, gloss, gp, n);
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
