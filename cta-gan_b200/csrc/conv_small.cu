// CUDA-core kernels for the degenerate convolutions of the path: one or two channels on one side of the GEMM
// (generator 7x7 head Cin=1 and tail Cout=1, discriminator first layer Cin=1|2 and last layer Cout=1).  They are <1% of the FLOPs
// but touch the largest activations (64 x 256^2); on tensor-core tiles 63/64 of the work would be padding, so they are written
// as bandwidth-style kernels: weights in shared memory, activations streamed once with 16-byte accesses.
#include "common.cuh"

namespace {

__device__ __forceinline__ bool tap_in(const ctagan_conv_geom &g, int o, int k, int pad, int in_extent, int &i_out) {
  int num = o * g.stride + k - pad;
  if (num < 0) return false;
  if (g.dil > 1) {
    if (num % g.dil) return false;
    num /= g.dil;
  }
  i_out = num;
  return num < in_extent;
}

// ---------------------------------------------------------------------------------------------------------------------
// few -> many:  Ci <= 2, Co % 8 == 0.   thread = (pixel, 8 output channels); block = 32 pixels x 64 output channels.
// ---------------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) conv_fewin_kernel(ctagan_conv_geom g, const T *__restrict__ x, const T *__restrict__ wp,
                                                         const float *__restrict__ bias, T *__restrict__ y) {
  extern __shared__ float ws[];   // [ntaps*Ci][64]
  const int ntaps = g.KH * g.KW, KC = ntaps * g.Ci;
  const int co_blk = blockIdx.y * 64;
  for (int i = threadIdx.x; i < KC * 64; i += 256) {
    const int k = i / 64, c = i - k * 64;
    ws[i] = (co_blk + c < g.Co) ? to_f(wp[(long long)(co_blk + c) * KC + k]) : 0.f;
  }
  __syncthreads();
  const long long M = (long long)g.N * g.Ho * g.Wo;
  const int cg = threadIdx.x & 7;
  const long long m = (long long)blockIdx.x * 32 + (threadIdx.x >> 3);
  if (m >= M || co_blk + cg * 8 >= g.Co) return;
  const int ow = (int)(m % g.Wo);
  const long long q = m / g.Wo;
  const int oh = (int)(q % g.Ho), n = (int)(q / g.Ho);
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = bias ? bias[co_blk + cg * 8 + j] : 0.f;
  for (int kh = 0; kh < g.KH; ++kh) {
    int ih;
    if (!tap_in(g, oh, kh, g.pad_h, g.Hi, ih)) continue;
    for (int kw = 0; kw < g.KW; ++kw) {
      int iw;
      if (!tap_in(g, ow, kw, g.pad_w, g.Wi, iw)) continue;
      const T *xp = x + (((long long)n * g.Hi + ih) * g.Wi + iw) * g.Ci;
      const float *wrow = ws + (kh * g.KW + kw) * g.Ci * 64 + cg * 8;
      for (int ci = 0; ci < g.Ci; ++ci) {
        const float xv = to_f(xp[ci]);
        const float4 w0 = *reinterpret_cast<const float4 *>(wrow + ci * 64);
        const float4 w1 = *reinterpret_cast<const float4 *>(wrow + ci * 64 + 4);
        acc[0] = fmaf(xv, w0.x, acc[0]); acc[1] = fmaf(xv, w0.y, acc[1]); acc[2] = fmaf(xv, w0.z, acc[2]); acc[3] = fmaf(xv, w0.w, acc[3]);
        acc[4] = fmaf(xv, w1.x, acc[4]); acc[5] = fmaf(xv, w1.y, acc[5]); acc[6] = fmaf(xv, w1.z, acc[6]); acc[7] = fmaf(xv, w1.w, acc[7]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = apply_act(acc[j], g.act);
  store_vec<T, 8>(y + m * g.Co + co_blk + cg * 8, acc);
}

// ---------------------------------------------------------------------------------------------------------------------
// many -> few:  Co <= 2, Ci % 64 == 0.   8 threads per pixel, each owning 8 of every 64 input channels (one 16-byte load
// per tap per 64-channel group), shuffle-reduced at the end.
// ---------------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) conv_fewout_kernel(ctagan_conv_geom g, const T *__restrict__ x, const T *__restrict__ wp,
                                                          const float *__restrict__ bias, T *__restrict__ y) {
  extern __shared__ float ws[];   // [Co][ntaps][Ci]
  const int ntaps = g.KH * g.KW, KC = ntaps * g.Ci;
  for (int i = threadIdx.x; i < g.Co * KC; i += 256) ws[i] = to_f(wp[i]);
  __syncthreads();
  const long long M = (long long)g.N * g.Ho * g.Wo;
  const int cs = threadIdx.x & 7;
  long long m = (long long)blockIdx.x * 32 + (threadIdx.x >> 3);
  const bool live = m < M;
  if (!live) m = M - 1;
  const int ow = (int)(m % g.Wo);
  const long long q = m / g.Wo;
  const int oh = (int)(q % g.Ho), n = (int)(q / g.Ho);
  float acc[2] = {0.f, 0.f};
  for (int kh = 0; kh < g.KH; ++kh) {
    int ih;
    if (!tap_in(g, oh, kh, g.pad_h, g.Hi, ih)) continue;
    for (int kw = 0; kw < g.KW; ++kw) {
      int iw;
      if (!tap_in(g, ow, kw, g.pad_w, g.Wi, iw)) continue;
      const T *xp = x + (((long long)n * g.Hi + ih) * g.Wi + iw) * g.Ci + cs * 8;
      const float *wrow = ws + (kh * g.KW + kw) * g.Ci + cs * 8;
      for (int c0 = 0; c0 < g.Ci; c0 += 64) {
        if (c0 + cs * 8 >= g.Ci) break;
        float xv[8];
        load_vec<T, 8>(xp + c0, xv);
        for (int co = 0; co < g.Co; ++co) {
          const float4 w0 = *reinterpret_cast<const float4 *>(wrow + co * KC + c0);
          const float4 w1 = *reinterpret_cast<const float4 *>(wrow + co * KC + c0 + 4);
          float s = acc[co];
          s = fmaf(xv[0], w0.x, s); s = fmaf(xv[1], w0.y, s); s = fmaf(xv[2], w0.z, s); s = fmaf(xv[3], w0.w, s);
          s = fmaf(xv[4], w1.x, s); s = fmaf(xv[5], w1.y, s); s = fmaf(xv[6], w1.z, s); s = fmaf(xv[7], w1.w, s);
          acc[co] = s;
        }
      }
    }
  }
#pragma unroll
  for (int co = 0; co < 2; ++co) {
    acc[co] += __shfl_xor_sync(0xffffffffu, acc[co], 1);
    acc[co] += __shfl_xor_sync(0xffffffffu, acc[co], 2);
    acc[co] += __shfl_xor_sync(0xffffffffu, acc[co], 4);
  }
  if (live && cs == 0) {
    for (int co = 0; co < g.Co; ++co) y[m * g.Co + co] = from_f<T>(apply_act(acc[co] + (bias ? bias[co] : 0.f), g.act));
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Register-tiled versions for stride 1, dilation 1 (the 7x7 generator head/tail at full resolution, where these layers cost most).
// Each thread produces PX = 4 consecutive output pixels of one row, so that every weight fetched from shared memory feeds 4x the
// FMAs and the input row segment (PX + KW - 1 positions) is loaded once per kernel row and reused across the KW taps.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int PX = 4;

// many -> few (Co <= 2): 8 threads per pixel group, each owning 8 of every 64 input channels.
template <typename T, int KW>
__global__ void __launch_bounds__(256) conv_fewout_tiled_kernel(ctagan_conv_geom g, const T *__restrict__ x, const T *__restrict__ wp,
                                                                const float *__restrict__ bias, T *__restrict__ y) {
  extern __shared__ float ws[];   // [Co][ntaps][Ci]
  const int ntaps = g.KH * KW, KC = ntaps * g.Ci;
  for (int i = threadIdx.x; i < g.Co * KC; i += 256) ws[i] = to_f(wp[i]);
  __syncthreads();
  const int groups_w = (g.Wo + PX - 1) / PX;
  const long long G = (long long)g.N * g.Ho * groups_w;
  const int cs = threadIdx.x & 7;
  long long gi = (long long)blockIdx.x * 32 + (threadIdx.x >> 3);
  const bool live = gi < G;
  if (!live) gi = G - 1;
  const int gw = (int)(gi % groups_w);
  const long long q = gi / groups_w;
  const int oh = (int)(q % g.Ho), n = (int)(q / g.Ho);
  const int ow0 = gw * PX;
  float acc[2][PX];
#pragma unroll
  for (int co = 0; co < 2; ++co)
#pragma unroll
    for (int p = 0; p < PX; ++p) acc[co][p] = 0.f;
  for (int c0 = 0; c0 < g.Ci; c0 += 64) {
    if (c0 + cs * 8 >= g.Ci) break;
    for (int kh = 0; kh < g.KH; ++kh) {
      const int ih = oh + kh - g.pad_h;
      if (ih < 0 || ih >= g.Hi) continue;
      float xw[PX + KW - 1][8];
      const T *xrow = x + ((long long)n * g.Hi + ih) * g.Wi * g.Ci + c0 + cs * 8;
#pragma unroll
      for (int j = 0; j < PX + KW - 1; ++j) {
        const int iw = ow0 + j - g.pad_w;
        if (iw >= 0 && iw < g.Wi) load_vec<T, 8>(xrow + (long long)iw * g.Ci, xw[j]);
        else {
#pragma unroll
          for (int e = 0; e < 8; ++e) xw[j][e] = 0.f;
        }
      }
#pragma unroll
      for (int kw = 0; kw < KW; ++kw) {
        const float *wrow = ws + (kh * KW + kw) * g.Ci + c0 + cs * 8;
        for (int co = 0; co < g.Co; ++co) {
          const float4 w0 = *reinterpret_cast<const float4 *>(wrow + co * KC);
          const float4 w1 = *reinterpret_cast<const float4 *>(wrow + co * KC + 4);
#pragma unroll
          for (int p = 0; p < PX; ++p) {
            float s = acc[co][p];
            s = fmaf(xw[p + kw][0], w0.x, s); s = fmaf(xw[p + kw][1], w0.y, s); s = fmaf(xw[p + kw][2], w0.z, s); s = fmaf(xw[p + kw][3], w0.w, s);
            s = fmaf(xw[p + kw][4], w1.x, s); s = fmaf(xw[p + kw][5], w1.y, s); s = fmaf(xw[p + kw][6], w1.z, s); s = fmaf(xw[p + kw][7], w1.w, s);
            acc[co][p] = s;
          }
        }
      }
    }
  }
#pragma unroll
  for (int co = 0; co < 2; ++co)
#pragma unroll
    for (int p = 0; p < PX; ++p) {
      float v = acc[co][p];
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      acc[co][p] = v;
    }
  if (live && cs == 0) {
    const long long base = ((long long)n * g.Ho + oh) * g.Wo + ow0;
    for (int p = 0; p < PX && ow0 + p < g.Wo; ++p)
      for (int co = 0; co < g.Co; ++co)
        y[(base + p) * g.Co + co] = from_f<T>(apply_act(acc[co][p] + (bias ? bias[co] : 0.f), g.act));
  }
}

// few -> many (Ci <= 2): thread = (pixel group of 4, 8 output channels)
template <typename T, int KW>
__global__ void __launch_bounds__(256) conv_fewin_tiled_kernel(ctagan_conv_geom g, const T *__restrict__ x, const T *__restrict__ wp,
                                                               const float *__restrict__ bias, T *__restrict__ y) {
  extern __shared__ float ws[];   // [ntaps*Ci][64]
  const int ntaps = g.KH * KW, KC = ntaps * g.Ci;
  const int co_blk = blockIdx.y * 64;
  for (int i = threadIdx.x; i < KC * 64; i += 256) {
    const int k = i / 64, c = i - k * 64;
    ws[i] = (co_blk + c < g.Co) ? to_f(wp[(long long)(co_blk + c) * KC + k]) : 0.f;
  }
  __syncthreads();
  const int groups_w = (g.Wo + PX - 1) / PX;
  const long long G = (long long)g.N * g.Ho * groups_w;
  const int cg = threadIdx.x & 7;
  const long long gi = (long long)blockIdx.x * 32 + (threadIdx.x >> 3);
  if (gi >= G || co_blk + cg * 8 >= g.Co) return;
  const int gw = (int)(gi % groups_w);
  const long long q = gi / groups_w;
  const int oh = (int)(q % g.Ho), n = (int)(q / g.Ho);
  const int ow0 = gw * PX;
  float acc[PX][8];
#pragma unroll
  for (int p = 0; p < PX; ++p)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[p][j] = bias ? bias[co_blk + cg * 8 + j] : 0.f;
  for (int ci = 0; ci < g.Ci; ++ci)
    for (int kh = 0; kh < g.KH; ++kh) {
      const int ih = oh + kh - g.pad_h;
      if (ih < 0 || ih >= g.Hi) continue;
      float xw[PX + KW - 1];
      const T *xrow = x + ((long long)n * g.Hi + ih) * g.Wi * g.Ci + ci;
#pragma unroll
      for (int j = 0; j < PX + KW - 1; ++j) {
        const int iw = ow0 + j - g.pad_w;
        xw[j] = (iw >= 0 && iw < g.Wi) ? to_f(xrow[(long long)iw * g.Ci]) : 0.f;
      }
#pragma unroll
      for (int kw = 0; kw < KW; ++kw) {
        const float *wrow = ws + ((kh * KW + kw) * g.Ci + ci) * 64 + cg * 8;
        const float4 w0 = *reinterpret_cast<const float4 *>(wrow);
        const float4 w1 = *reinterpret_cast<const float4 *>(wrow + 4);
#pragma unroll
        for (int p = 0; p < PX; ++p) {
          const float xv = xw[p + kw];
          acc[p][0] = fmaf(xv, w0.x, acc[p][0]); acc[p][1] = fmaf(xv, w0.y, acc[p][1]); acc[p][2] = fmaf(xv, w0.z, acc[p][2]);
          acc[p][3] = fmaf(xv, w0.w, acc[p][3]); acc[p][4] = fmaf(xv, w1.x, acc[p][4]); acc[p][5] = fmaf(xv, w1.y, acc[p][5]);
          acc[p][6] = fmaf(xv, w1.z, acc[p][6]); acc[p][7] = fmaf(xv, w1.w, acc[p][7]);
        }
      }
    }
  const long long base = ((long long)n * g.Ho + oh) * g.Wo + ow0;
#pragma unroll
  for (int p = 0; p < PX; ++p) {
    if (ow0 + p >= g.Wo) break;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[p][j] = apply_act(acc[p][j], g.act);
    store_vec<T, 8>(y + (base + p) * g.Co + co_blk + cg * 8, acc[p]);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// thin weight gradients.  V = the wide tensor (C channels), S = the thin one (SC <= 2 channels).
//   gy_thin == 0 :  gy is wide (V at output positions), gx is thin   -> dw[c][sc][tap]   (head 7x7 Cin=1, discriminator layer 0)
//   gy_thin == 1 :  gx is wide (V at input positions),  gy is thin   -> dw[sc][c][tap]   (tail 7x7 Cout=1, discriminator last layer)
// thread = (channel c of a 64-channel block, kernel row kh); it walks along image rows keeping KW accumulators.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int THIN_MAXKW = 7;

// db[s] (this block's share) = sum of gy over the block's slice of pixels, warps combined in warp order (256-thread blocks)
template <typename T>
__device__ __forceinline__ void thin_gy_bias(const ctagan_conv_geom &g, const T *__restrict__ gy, float *__restrict__ db, int SC) {
  __shared__ float wsum[8];
  const long long total = (long long)g.N * g.Ho * g.Wo;
  const long long per = (total + gridDim.x - 1) / gridDim.x;
  const long long p0 = blockIdx.x * per, p1 = min(total, p0 + per);
  for (int s = 0; s < SC; ++s) {
    float t = 0.f;
    for (long long p = p0 + threadIdx.x; p < p1; p += blockDim.x) t += to_f(gy[p * SC + s]);
    t = warp_sum(t);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
      float a = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) a += wsum[w];
      db[s] = a;
    }
  }
}

// Register-window version: thread = (channel c, kernel row kh).  For U consecutive positions of the wide tensor it loads U wide
// values and the (U-1)*STRIDE+KW thin values they touch once, then issues U*KW FMAs (>= 60% of the instruction stream).
template <typename T, int KW, int STRIDE>
__global__ void __launch_bounds__(256) conv_wgrad_thin_kernel(ctagan_conv_geom g, const T *__restrict__ gy, const T *__restrict__ gx,
                                                              float *__restrict__ dw_part, float *__restrict__ db_part, int gy_thin,
                                                              int rows_per_block) {
  // deterministic split reduction: block x stores its partial sums in row x of dw_part[gridDim.x][Co*Ci*taps] / db_part[gridDim.x][Co]
  float *dw = dw_part + (long long)blockIdx.x * g.Co * g.Ci * g.KH * KW;
  float *db = db_part ? db_part + (long long)blockIdx.x * g.Co : nullptr;
  constexpr int U = 8;
  constexpr int WIN = (U - 1) * STRIDE + KW;
  const int c = blockIdx.y * 64 + (threadIdx.x & 63);
  const int kh = blockIdx.z * 4 + (threadIdx.x >> 6);
  const int C = gy_thin ? g.Ci : g.Co;            // wide channel count
  const int SC = gy_thin ? g.Co : g.Ci;           // thin channel count (1 or 2)
  const int ntaps = g.KH * KW;
  const int VH = gy_thin ? g.Hi : g.Ho, VW = gy_thin ? g.Wi : g.Wo;
  const int SW = gy_thin ? g.Wo : g.Wi, SH = gy_thin ? g.Ho : g.Hi;
  const long long row0 = (long long)blockIdx.x * rows_per_block;
  const long long row1 = min((long long)g.N * VH, row0 + rows_per_block);
  const bool active = (c < C) && (kh < g.KH);
  float bsum = 0.f;
  if (active) {
    for (int s = 0; s < SC; ++s) {
      float acc[KW];
#pragma unroll
      for (int k = 0; k < KW; ++k) acc[k] = 0.f;
      for (long long r = row0; r < row1; ++r) {
        const int n = (int)(r / VH), vh = (int)(r - (long long)n * VH);
        const int sh = gy_thin ? (vh + g.pad_h - kh) : (vh * STRIDE + kh - g.pad_h);
        const bool row_ok = sh >= 0 && sh < SH;
        const T *vrow = (gy_thin ? gx : gy) + ((long long)n * VH + vh) * VW * C + c;
        if (!row_ok) {
          if (!gy_thin && s == 0 && kh == 0 && blockIdx.z == 0 && db)
            for (int vw = 0; vw < VW; ++vw) bsum += to_f(vrow[(long long)vw * C]);
          continue;
        }
        const T *srow = (gy_thin ? gy : gx) + ((long long)n * SH + sh) * SW * SC + s;
        for (int vw0 = 0; vw0 < VW; vw0 += U) {
          float v[U], win[WIN];
#pragma unroll
          for (int u = 0; u < U; ++u) v[u] = (vw0 + u < VW) ? to_f(vrow[(long long)(vw0 + u) * C]) : 0.f;
          // window start in the thin row: gy wide: sw = vw*STRIDE + kw - pad ; gy thin (stride 1): sw = vw + pad - kw
          const int wbase = gy_thin ? (vw0 + g.pad_w - (KW - 1)) : (vw0 * STRIDE - g.pad_w);
#pragma unroll
          for (int j = 0; j < WIN; ++j) {
            const int sw = wbase + j;
            win[j] = (sw >= 0 && sw < SW) ? to_f(srow[(long long)sw * SC]) : 0.f;
          }
          if (!gy_thin && s == 0 && kh == 0 && blockIdx.z == 0) {
#pragma unroll
            for (int u = 0; u < U; ++u) bsum += v[u];
          }
#pragma unroll
          for (int u = 0; u < U; ++u)
#pragma unroll
            for (int kw = 0; kw < KW; ++kw) {
              const int j = gy_thin ? (u - kw + KW - 1) : (u * STRIDE + kw);
              acc[kw] = fmaf(v[u], win[j], acc[kw]);
            }
        }
      }
#pragma unroll
      for (int kw = 0; kw < KW; ++kw) {
        // dw is [A][B][KH][KW] with A = gy channels, B = gx channels
        const long long idx = gy_thin ? (((long long)s * C + c) * ntaps + kh * KW + kw) : (((long long)c * SC + s) * ntaps + kh * KW + kw);
        dw[idx] = acc[kw];
      }
    }
    if (db && !gy_thin && kh == 0 && blockIdx.z == 0) db[c] = bsum;
  }
  // bias gradient of the thin-gy case: db[sc] = sum of gy (first channel block / kernel-row group only)
  if (db && gy_thin && blockIdx.y == 0 && blockIdx.z == 0) thin_gy_bias(g, gy, db, SC);
}

// Vectorised thin weight gradient for stride 1: thread = (8-channel slice, kernel row); one 16-byte load of the wide tensor and
// KW (warp-broadcast) scalars of the thin one feed 8*KW FMAs.  block = 4 streams x (8 slices x 8 kernel rows); each stream walks a
// contiguous range of wide-tensor pixels; streams are combined in shared memory, blocks by fp32 atomics.
template <typename T, int KW>
__global__ void __launch_bounds__(256) conv_wgrad_thin_vec_kernel(ctagan_conv_geom g, const T *__restrict__ gy, const T *__restrict__ gx,
                                                                  float *__restrict__ dw_part, float *__restrict__ db_part, int gy_thin,
                                                                  long long pix_per_stream) {
  // deterministic split reduction: block x stores its partial sums in row x of dw_part[gridDim.x][Co*Ci*taps] / db_part[gridDim.x][Co]
  float *dw = dw_part + (long long)blockIdx.x * g.Co * g.Ci * g.KH * KW;
  float *db = db_part ? db_part + (long long)blockIdx.x * g.Co : nullptr;
  __shared__ float red[64 * KW * 8];
  __shared__ float redb[64];
  const int t = threadIdx.x;
  const int stream = t >> 6, cs = t & 7, kh = (t >> 3) & 7;
  const int C = gy_thin ? g.Ci : g.Co;            // wide channel count
  const int SC = gy_thin ? g.Co : g.Ci;           // thin channel count (1 or 2)
  const int c = blockIdx.y * 64 + cs * 8;
  const int ntaps = g.KH * KW;
  const int VH = gy_thin ? g.Hi : g.Ho, VW = gy_thin ? g.Wi : g.Wo;
  const int SW = gy_thin ? g.Wo : g.Wi, SH = gy_thin ? g.Ho : g.Hi;
  const long long total = (long long)g.N * VH * VW;
  const long long p0 = ((long long)blockIdx.x * 4 + stream) * pix_per_stream;
  const long long p1 = min(total, p0 + pix_per_stream);
  const bool active = (c < C) && (kh < g.KH);
  const T *V = gy_thin ? gx : gy;
  const T *S = gy_thin ? gy : gx;
  for (int s = 0; s < SC; ++s) {
    float acc[KW][8];
#pragma unroll
    for (int k = 0; k < KW; ++k)
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[k][e] = 0.f;
    float bsum[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (active && p0 < p1) {
      int vw = (int)(p0 % VW);
      long long r = p0 / VW;                      // n*VH + vh
      int vh = (int)(r % VH);
      int n = (int)(r / VH);
      for (long long p = p0; p < p1; ++p) {
        const int sh = gy_thin ? (vh + g.pad_h - kh) : (vh + kh - g.pad_h);
        float v[8];
        load_vec<T, 8>(V + p * C + c, v);
        if (!gy_thin && s == 0 && kh == 0 && db) {
#pragma unroll
          for (int e = 0; e < 8; ++e) bsum[e] += v[e];
        }
        if (sh >= 0 && sh < SH) {
          const T *srow = S + ((long long)n * SH + sh) * SW * SC + s;
#pragma unroll
          for (int kw = 0; kw < KW; ++kw) {
            const int sw = gy_thin ? (vw + g.pad_w - kw) : (vw + kw - g.pad_w);
            const float sv = (sw >= 0 && sw < SW) ? to_f(srow[(long long)sw * SC]) : 0.f;
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[kw][e] = fmaf(v[e], sv, acc[kw][e]);
          }
        }
        if (++vw == VW) { vw = 0; if (++vh == VH) { vh = 0; ++n; } }
      }
    }
    // combine the 4 streams of the block in stream order (plain read-modify-write, one stream at a time), then one store per output
    float *mine = red + (t & 63) * (KW * 8);
    for (int st_ = 0; st_ < 4; ++st_) {
      if (stream == st_ && active) {
#pragma unroll
        for (int k = 0; k < KW; ++k)
#pragma unroll
          for (int e = 0; e < 8; ++e) mine[k * 8 + e] = (st_ == 0 ? 0.f : mine[k * 8 + e]) + acc[k][e];
        if (kh == 0) {
#pragma unroll
          for (int e = 0; e < 8; ++e) redb[cs * 8 + e] = (st_ == 0 ? 0.f : redb[cs * 8 + e]) + bsum[e];
        }
      }
      __syncthreads();
    }
    if (stream == 0 && active) {
#pragma unroll
      for (int k = 0; k < KW; ++k)
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int cc = c + e;
          // dw is [A][B][KH][KW] with A = gy channels, B = gx channels
          const long long idx = gy_thin ? (((long long)s * C + cc) * ntaps + kh * KW + k) : (((long long)cc * SC + s) * ntaps + kh * KW + k);
          dw[idx] = mine[k * 8 + e];
        }
      if (db && !gy_thin && s == 0 && kh == 0) {
#pragma unroll
        for (int e = 0; e < 8; ++e) db[c + e] = redb[cs * 8 + e];
      }
    }
    __syncthreads();
  }
  // bias gradient of the thin-gy case: db[sc] = sum of gy (first channel block only)
  if (db && gy_thin && blockIdx.y == 0) thin_gy_bias(g, gy, db, SC);
}

}  // namespace

int ctagan_conv_small_kind(const ctagan_conv_geom *g) {
  if (g->Ci <= 2 && g->Co % 8 == 0 && g->KH * g->KW * g->Ci * 64 * 4 <= 48 * 1024) return 1;
  if (g->Co <= 2 && g->Ci % 8 == 0 && (long long)g->Co * g->KH * g->KW * g->Ci * 4 <= 96 * 1024) return 2;
  return 0;
}

int ctagan_conv_gather_small(const ctagan_conv_geom *g, const void *x, const void *wp, const float *bias, void *y, cudaStream_t st) {
  const int kind = ctagan_conv_small_kind(g);
  const long long M = (long long)g->N * g->Ho * g->Wo;
  const bool tiled = g->stride == 1 && g->dil == 1 && (g->KW == 7 || g->KW == 4 || g->KW == 3);
  const long long Gr = (long long)g->N * g->Ho * ((g->Wo + PX - 1) / PX);
  if (kind == 1 && tiled) {
    dim3 grid(cdiv(Gr, 32), cdiv(g->Co, 64));
    const size_t smem = (size_t)g->KH * g->KW * g->Ci * 64 * sizeof(float);
    CTAGAN_DISPATCH_DTYPE(g->dtype, T, {
      if (g->KW == 7) conv_fewin_tiled_kernel<T, 7><<<grid, 256, smem, st>>>(*g, (const T *)x, (const T *)wp, bias, (T *)y);
      else if (g->KW == 4) conv_fewin_tiled_kernel<T, 4><<<grid, 256, smem, st>>>(*g, (const T *)x, (const T *)wp, bias, (T *)y);
      else conv_fewin_tiled_kernel<T, 3><<<grid, 256, smem, st>>>(*g, (const T *)x, (const T *)wp, bias, (T *)y);
    });
  } else if (kind == 2 && tiled) {
    const size_t smem = (size_t)g->Co * g->KH * g->KW * g->Ci * sizeof(float);
    CTAGAN_DISPATCH_DTYPE(g->dtype, T, {
      static bool configured = false;
      if (!configured) {
        CTAGAN_CUDA_OK(cudaFuncSetAttribute(conv_fewout_tiled_kernel<T, 7>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        CTAGAN_CUDA_OK(cudaFuncSetAttribute(conv_fewout_tiled_kernel<T, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        CTAGAN_CUDA_OK(cudaFuncSetAttribute(conv_fewout_tiled_kernel<T, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        configured = true;
      }
      if (g->KW == 7) conv_fewout_tiled_kernel<T, 7><<<cdiv(Gr, 32), 256, smem, st>>>(*g, (const T *)x, (const T *)wp, bias, (T *)y);
      else if (g->KW == 4) conv_fewout_tiled_kernel<T, 4><<<cdiv(Gr, 32), 256, smem, st>>>(*g, (const T *)x, (const T *)wp, bias, (T *)y);
      else conv_fewout_tiled_kernel<T, 3><<<cdiv(Gr, 32), 256, smem, st>>>(*g, (const T *)x, (const T *)wp, bias, (T *)y);
    });
  } else if (kind == 1) {
    dim3 grid(cdiv(M, 32), cdiv(g->Co, 64));
    const size_t smem = (size_t)g->KH * g->KW * g->Ci * 64 * sizeof(float);
    CTAGAN_DISPATCH_DTYPE(g->dtype, T, { conv_fewin_kernel<T><<<grid, 256, smem, st>>>(*g, (const T *)x, (const T *)wp, bias, (T *)y); });
  } else if (kind == 2) {
    const size_t smem = (size_t)g->Co * g->KH * g->KW * g->Ci * sizeof(float);
    CTAGAN_DISPATCH_DTYPE(g->dtype, T, {
      static bool configured = false;
      if (!configured) {
        CTAGAN_CUDA_OK(cudaFuncSetAttribute(conv_fewout_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        configured = true;
      }
      conv_fewout_kernel<T><<<cdiv(M, 32), 256, smem, st>>>(*g, (const T *)x, (const T *)wp, bias, (T *)y);
    });
  } else {
    ctagan_set_error("conv_gather_small: geometry is not degenerate");
    return CTAGAN_ERR_UNSUPPORTED;
  }
  CTAGAN_LAUNCH_OK();
  return CTAGAN_OK;
}

int ctagan_conv_wgrad_thin_eligible(const ctagan_conv_geom *g) {
  if (g->dil != 1 || g->KH > 8) return 0;
  const bool shape_ok = (g->KW == 7 && g->stride == 1) || (g->KW == 4 && g->stride <= 2) || (g->KW == 3 && g->stride == 1) ||
                        (g->KW == 1 && g->stride == 1);
  if (!shape_ok) return 0;
  if (g->Ci <= 2 && g->Co >= 8) return 1;                       // gx thin
  if (g->Co <= 2 && g->Ci >= 8 && g->stride == 1) return 2;     // gy thin
  return 0;
}

namespace {
struct ThinPlan {
  bool vec;
  int gy_thin, C;
  dim3 grid;
  long long pps;      // vec: pixels per stream
  int rpb;            // row walker: rows per block
};

bool plan_thin(const ctagan_conv_geom *g, ThinPlan &pl) {
  const int kind = ctagan_conv_wgrad_thin_eligible(g);
  if (!kind) return false;
  pl.gy_thin = kind == 2;
  pl.C = pl.gy_thin ? g->Ci : g->Co;
  const int VH = pl.gy_thin ? g->Hi : g->Ho;
  const long long rows = (long long)g->N * VH;
  pl.vec = g->stride == 1 && pl.C % 8 == 0 && g->KH <= 8 && (g->KW == 7 || g->KW == 4 || g->KW == 3 || g->KW == 1);
  if (pl.vec) {
    const int VW = pl.gy_thin ? g->Wi : g->Wo;
    const long long total = rows * VW;
    const int chb = cdiv(pl.C, 64);
    long long blocks = (2LL * ctagan_num_sms() + chb - 1) / chb;
    if (blocks * 4 * 32 > total) blocks = (total + 127) / 128;       // at least 32 pixels per stream
    if (blocks < 1) blocks = 1;
    pl.pps = (total + blocks * 4 - 1) / (blocks * 4);
    pl.grid = dim3((unsigned)((total + pl.pps * 4 - 1) / (pl.pps * 4)), chb);
    return true;
  }
  const int ch_blocks = cdiv(pl.C, 64), kh_blocks = cdiv(g->KH, 4);
  long long want = (2LL * ctagan_num_sms()) / ((long long)ch_blocks * kh_blocks);
  if (want < 1) want = 1;
  if (want > rows) want = rows;
  pl.rpb = (int)((rows + want - 1) / want);
  pl.grid = dim3(cdiv(rows, pl.rpb), ch_blocks, kh_blocks);
  return true;
}
}  // namespace

// scratch: per-block partial sums [grid.x][Co*Ci*taps] + [grid.x][Co]
size_t ctagan_conv_wgrad_thin_workspace(const ctagan_conv_geom *g) {
  ThinPlan pl;
  if (!plan_thin(g, pl)) return 0;
  return (size_t)pl.grid.x * ((size_t)g->Co * g->Ci * g->KH * g->KW + g->Co) * sizeof(float);
}

int ctagan_conv_wgrad_thin(const ctagan_conv_geom *g, const void *gy, const void *gx, float *dw, float *db, void *workspace,
                           size_t workspace_bytes, cudaStream_t st, int accumulate) {
  ThinPlan pl;
  if (!plan_thin(g, pl)) {
    ctagan_set_error("conv_wgrad_thin: geometry is not degenerate");
    return CTAGAN_ERR_UNSUPPORTED;
  }
  const size_t need = ctagan_conv_wgrad_thin_workspace(g);
  CTAGAN_REQUIRE(workspace && workspace_bytes >= need, "conv_wgrad(thin): workspace of %zu bytes required (got %zu)", need, workspace_bytes);
  const long long dw_elems = (long long)g->Co * g->Ci * g->KH * g->KW;
  float *dw_part = (float *)workspace;
  float *db_part = db ? dw_part + (long long)pl.grid.x * dw_elems : nullptr;
  const int gy_thin = pl.gy_thin;
  if (pl.vec) {
#define THINV_LAUNCH(KW_) conv_wgrad_thin_vec_kernel<T, KW_><<<pl.grid, 256, 0, st>>>(*g, (const T *)gy, (const T *)gx, dw_part, db_part, gy_thin, pl.pps)
    CTAGAN_DISPATCH_DTYPE(g->dtype, T, {
      if (g->KW == 7) THINV_LAUNCH(7);
      else if (g->KW == 4) THINV_LAUNCH(4);
      else if (g->KW == 3) THINV_LAUNCH(3);
      else THINV_LAUNCH(1);
    });
#undef THINV_LAUNCH
  } else {
#define THIN_LAUNCH(KW_, S_) conv_wgrad_thin_kernel<T, KW_, S_><<<pl.grid, 256, 0, st>>>(*g, (const T *)gy, (const T *)gx, dw_part, db_part, gy_thin, pl.rpb)
    CTAGAN_DISPATCH_DTYPE(g->dtype, T, {
      if (g->KW == 7) THIN_LAUNCH(7, 1);
      else if (g->KW == 4 && g->stride == 1) THIN_LAUNCH(4, 1);
      else if (g->KW == 4) THIN_LAUNCH(4, 2);
      else if (g->KW == 3) THIN_LAUNCH(3, 1);
      else THIN_LAUNCH(1, 1);
    });
#undef THIN_LAUNCH
  }
  CTAGAN_LAUNCH_OK();
  int rc = ctagan_ordered_sum(dw_part, dw, (int)pl.grid.x, dw_elems, st, accumulate);
  if (rc) return rc;
  if (db) rc = ctagan_ordered_sum(db_part, db, (int)pl.grid.x, g->Co, st, accumulate);
  return rc;
}
