// Bandwidth-bound NHWC kernels around the convolutions: InstanceNorm statistics, fused normalise+activation+residual+
// reflection-pad and its backward, pooling, bilinear 2x upsample + concat, plane means, casts.
// All are coalesced along the channel (innermost) dimension with 16-byte vectors when C allows it.
#include <stdlib.h>
#include <cooperative_groups.h>
#include "common.cuh"

// All index arithmetic in this file is 32-bit (64-bit integer division costs ~100 instructions on the SM and dominated these
// latency-bound kernels); the host wrappers reject tensors with >= 2^31 elements.
typedef int idx_t;
#define CTAGAN_FITS32(expr) CTAGAN_REQUIRE((int64_t)(expr) < (int64_t)0x7fffffff, "tensor too large for 32-bit indexing")

namespace {

// (mean, rstd) of V consecutive channels = 2V consecutive floats of stats[N][C][2]: fetched with 16-byte loads
template <int V>
__device__ __forceinline__ void load_stats(const float *__restrict__ sp, float (&mean)[V], float (&rstd)[V]) {
  if constexpr (V >= 2) {
#pragma unroll
    for (int i = 0; i < V; i += 2) {
      const float4 t = *reinterpret_cast<const float4 *>(sp + 2 * i);
      mean[i] = t.x; rstd[i] = t.y; mean[i + 1] = t.z; rstd[i + 1] = t.w;
    }
  } else {
    mean[0] = sp[0]; rstd[0] = sp[1];
  }
}
template <int V>
__device__ __forceinline__ void load_acc_means(const double *__restrict__ ap, float inv_hw, float (&m1)[V], float (&m2)[V]) {
  if constexpr (V >= 2) {
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const double2 t = *reinterpret_cast<const double2 *>(ap + 2 * i);
      m1[i] = (float)(t.x * (double)inv_hw); m2[i] = (float)(t.y * (double)inv_hw);
    }
  } else {
    m1[0] = (float)(ap[0] * (double)inv_hw); m2[0] = (float)(ap[1] * (double)inv_hw);
  }
}

__device__ __forceinline__ int reflect_idx(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}

// ---------------------------------------------------------------------------------------------------------------
// InstanceNorm statistics.  Shifted sums (shift = first pixel of the plane) in fp32 per thread, fp64 across threads.
// grid (chunks, N); block 256 = PL pixel lanes x CV channel vectors.  Every block stores its partial sums in its own row of
// acc[N][chunks][C][2]; the finalize kernel adds the rows in chunk order (deterministic: no floating-point atomics).
// ---------------------------------------------------------------------------------------------------------------
template <typename T, int V>
__global__ void __launch_bounds__(256) instnorm_partial_kernel(const T *__restrict__ x, double *__restrict__ acc, int HW, int C,
                                                               int pix_per_block) {
  const int CV = C / V;
  const int n = blockIdx.y;
  const int lanes = 256 / CV > 0 ? 256 / CV : 1;  // pixel lanes (CV <= 256 guaranteed by host)
  const int cv = threadIdx.x % CV, pl = threadIdx.x / CV;
  const T *base = x + (idx_t)n * HW * C;
  float shift[V], s1[V], s2[V];
  load_vec<T, V>(base + cv * V, shift);
#pragma unroll
  for (int i = 0; i < V; ++i) s1[i] = s2[i] = 0.f;
  const int p0 = blockIdx.x * pix_per_block;
  const int p1 = min(HW, p0 + pix_per_block);
  if (pl < lanes) {
    constexpr int U = 4;                       // independent loads in flight per thread
    for (int p = p0 + pl; p < p1; p += lanes * U) {
      float v[U][V];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int pp = p + u * lanes;
        if (pp < p1) load_vec<T, V>(base + (idx_t)pp * C + cv * V, v[u]);
        else {
#pragma unroll
          for (int i = 0; i < V; ++i) v[u][i] = shift[i];
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int i = 0; i < V; ++i) {
          const float d = v[u][i] - shift[i];
          s1[i] += d;
          s2[i] = fmaf(d, d, s2[i]);
        }
    }
  }
  // combine pixel lanes through shared memory in fp64
  // combine the pixel lanes in ONE shared-memory pass: every thread then owns one (channel, sum|sumsq) pair group and issues its
  // atomics in parallel (the old per-channel loop serialised 16 block barriers and put all atomics on 32 threads)
  __shared__ float sm[2][V][256];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    sm[0][i][threadIdx.x] = (pl < lanes) ? s1[i] : 0.f;
    sm[1][i][threadIdx.x] = (pl < lanes) ? s2[i] : 0.f;
  }
  __syncthreads();
  for (int item = threadIdx.x; item < CV * V; item += 256) {
    const int ccv = item % CV, i = item / CV;
    double a = 0.0, b = 0.0;
    for (int l = 0; l < lanes; ++l) {
      a += (double)sm[0][i][l * CV + ccv];
      b += (double)sm[1][i][l * CV + ccv];
    }
    double *dst = acc + (((idx_t)n * gridDim.x + blockIdx.x) * C + ccv * V + i) * 2;
    dst[0] = a;
    dst[1] = b;
  }
}

template <typename T>
__global__ void instnorm_finalize_kernel(const T *__restrict__ x, const double *__restrict__ acc, float *__restrict__ stats, int N,
                                         int HW, int C, int chunks) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * C) return;
  const int n = i / C, c = i - n * C;
  const double shift = (double)to_f(x[(idx_t)n * HW * C + c]);
  double a1 = 0.0, a2 = 0.0;
  for (int k = 0; k < chunks; ++k) {
    const double2 t = *reinterpret_cast<const double2 *>(acc + (((idx_t)n * chunks + k) * C + c) * 2);
    a1 += t.x;
    a2 += t.y;
  }
  const double m1 = a1 / HW, m2 = a2 / HW;
  double var = m2 - m1 * m1;
  if (var < 0) var = 0;
  stats[2 * i] = (float)(shift + m1);
  stats[2 * i + 1] = (float)(1.0 / sqrt(var + 1e-5));
}

// ---------------------------------------------------------------------------------------------------------------
// out[n,hp,wp,c] = act((x-mean)*rstd) + res ; gather form, reflect indices.
// grid (pixel chunks, N); block = CV channel vectors x (256 / CV) pixel lanes: a thread keeps its channel vector, so (mean, rstd) are
// loaded once per thread instead of once per 16-byte output and the pixel walk needs one division per pixel.
// ---------------------------------------------------------------------------------------------------------------
template <typename T, int V>
__global__ void __launch_bounds__(256) norm_act_pad_kernel(const T *__restrict__ x, const float *__restrict__ stats,
                                                           const T *__restrict__ res, int res_pad, T *__restrict__ out, int N, int H,
                                                           int W, int C, int pad, int act, int pix_per_block) {
  pdl_wait();
  const int CV = C / V;
  const int n = blockIdx.y;
  const int lanes = 256 / CV > 0 ? 256 / CV : 1;
  const int cv = threadIdx.x % CV, pl = threadIdx.x / CV;
  if (pl >= lanes) return;
  const int Hp = H + 2 * pad, Wp = W + 2 * pad;
  float mean[V], rstd[V];
  if (stats) load_stats<V>(stats + ((idx_t)n * C + cv * V) * 2, mean, rstd);
  const int q0 = blockIdx.x * pix_per_block, q1 = min(Hp * Wp, q0 + pix_per_block);
  const T *xn = x + (idx_t)n * H * W * C + cv * V;
  T *on = out + (idx_t)n * Hp * Wp * C + cv * V;
  const int Hr = H + 2 * res_pad, Wr = W + 2 * res_pad;
  const T *rn = res ? res + (idx_t)n * Hr * Wr * C + cv * V : nullptr;
  constexpr int U = 4;                          // independent pixels in flight per thread
  for (int q = q0 + pl; q < q1; q += lanes * U) {
    float v[U][V], rv[U][V];
    bool ok[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int qq = q + u * lanes;
      ok[u] = qq < q1;
      const int qc = ok[u] ? qq : q;
      const int hp = qc / Wp, wp = qc - hp * Wp;
      const int h = reflect_idx(hp - pad, H), w = reflect_idx(wp - pad, W);
      load_vec<T, V>(xn + ((idx_t)h * W + w) * C, v[u]);
      if (rn) load_vec<T, V>(rn + ((idx_t)(h + res_pad) * Wr + w + res_pad) * C, rv[u]);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (!ok[u]) continue;
      if (stats) {
#pragma unroll
        for (int i = 0; i < V; ++i) v[u][i] = (v[u][i] - mean[i]) * rstd[i];
      }
#pragma unroll
      for (int i = 0; i < V; ++i) v[u][i] = apply_act(v[u][i], act);
      if (rn) {
#pragma unroll
        for (int i = 0; i < V; ++i) v[u][i] += rv[u][i];
      }
      store_vec<T, V>(on + (idx_t)(q + u * lanes) * C, v[u]);
    }
  }
}

// padded positions that reflect onto interior index h (<= 3 of them)
__device__ __forceinline__ int fold_sources(int h, int H, int pad, int (&src)[3]) {
  int cnt = 0;
  src[cnt++] = h + pad;
  if (pad > 0) {
    if (h >= 1 && h <= pad) src[cnt++] = pad - h;
    if (h <= H - 2 && h >= H - 1 - pad) src[cnt++] = 2 * (H - 1) + pad - h;
  }
  return cnt;
}

__device__ __forceinline__ bool fold_is_border(int h, int w, int H, int W, int pad) {
  return pad > 0 && ((h >= 1 && h <= pad) || (h <= H - 2 && h >= H - 1 - pad) || (w >= 1 && w <= pad) || (w <= W - 2 && w >= W - 1 - pad));
}

// adds the mirrored (reflection) sources of a border pixel; the main source (h+pad, w+pad) is loaded by the caller
template <typename T, int V>
__device__ __noinline__ void fold_extras(const T *__restrict__ gout, int n, int h, int w, int cv, int H, int W, int C, int pad,
                                         float (&g)[V]) {
  const int Hp = H + 2 * pad, Wp = W + 2 * pad;
  int hs[3], ws[3];
  const int nh = fold_sources(h, H, pad, hs), nw = fold_sources(w, W, pad, ws);
  for (int a = 0; a < nh; ++a)
    for (int b = 0; b < nw; ++b) {
      if (a == 0 && b == 0) continue;
      float t[V];
      load_vec<T, V>(gout + (((idx_t)n * Hp + hs[a]) * Wp + ws[b]) * C + cv * V, t);
#pragma unroll
      for (int i = 0; i < V; ++i) g[i] += t[i];
    }
}

template <typename T, int V>
__device__ __forceinline__ void folded_grad(const T *__restrict__ gout, int n, int h, int w, int cv, int H, int W, int C, int pad,
                                            float (&g)[V]) {
  const int Wp = W + 2 * pad;
  load_vec<T, V>(gout + (((idx_t)n * (H + 2 * pad) + h + pad) * Wp + w + pad) * C + cv * V, g);
  if (fold_is_border(h, w, H, W, pad)) fold_extras<T, V>(gout, n, h, w, cv, H, W, C, pad, g);
}

__device__ __forceinline__ float act_grad_from_sign(float pre, int act) {
  if (act == CTAGAN_ACT_RELU) return pre > 0.f ? 1.f : 0.f;
  if (act == CTAGAN_ACT_LRELU) return pre > 0.f ? 1.f : 0.2f;
  return 1.f;
}

// pass 1: per-(n,c) sums of g and g*xhat.  grid (chunks, N).  Deterministic: every block stores its partial sums in its own row of
// part[N][chunks][C][2]; the last block of an image to arrive (ticket = the double at acc[N*C*2 + n], zero on entry and again on
// exit) adds the rows in chunk order into acc[n][C][2].
template <typename T, int V>
__global__ void __launch_bounds__(256) norm_bwd_reduce_kernel(const T *__restrict__ gout, const T *__restrict__ x,
                                                              const float *__restrict__ stats, const T *__restrict__ addend,
                                                              double *__restrict__ acc, double *__restrict__ part, int N, int H, int W,
                                                              int C, int pad, int act, int pix_per_block) {
  pdl_wait();
  const int CV = C / V;
  const int n = blockIdx.y;
  const int lanes = 256 / CV > 0 ? 256 / CV : 1;
  const int cv = threadIdx.x % CV, pl = threadIdx.x / CV;
  const int HW = H * W;
  float mean[V], rstd[V], s1[V], s2[V];
  load_stats<V>(stats + ((idx_t)n * C + cv * V) * 2, mean, rstd);
#pragma unroll
  for (int i = 0; i < V; ++i) s1[i] = s2[i] = 0.f;
  const int p0 = blockIdx.x * pix_per_block, p1 = min(HW, p0 + pix_per_block);
  if (pl < lanes) {
    constexpr int U = 4;                       // independent loads in flight per thread
    const int Wp = W + 2 * pad;
    for (int p = p0 + pl; p < p1; p += lanes * U) {
      float g[U][V], xv[U][V];
      int hh[U], ww[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int pp = p + u * lanes;
        const bool ok = pp < p1;
        const int pc = ok ? pp : p;
        hh[u] = pc / W; ww[u] = pc - hh[u] * W;
        load_vec<T, V>(gout + (((idx_t)n * (H + 2 * pad) + hh[u] + pad) * Wp + ww[u] + pad) * C + cv * V, g[u]);
        load_vec<T, V>(x + ((idx_t)n * HW + pc) * C + cv * V, xv[u]);
        if (addend) {
          float av[V];
          load_vec<T, V>(addend + ((idx_t)n * HW + pc) * C + cv * V, av);
#pragma unroll
          for (int i = 0; i < V; ++i) g[u][i] += av[i];
        }
        if (!ok) {
#pragma unroll
          for (int i = 0; i < V; ++i) g[u][i] = 0.f;
          hh[u] = -1;
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (hh[u] >= 0 && fold_is_border(hh[u], ww[u], H, W, pad)) fold_extras<T, V>(gout, n, hh[u], ww[u], cv, H, W, C, pad, g[u]);
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int i = 0; i < V; ++i) {
          const float xh = (xv[u][i] - mean[i]) * rstd[i];
          const float gg = g[u][i] * act_grad_from_sign(xh, act);
          s1[i] += gg;
          s2[i] = fmaf(gg, xh, s2[i]);
        }
    }
  }
  // combine the pixel lanes in ONE shared-memory pass: every thread then owns one (channel, sum|sumsq) pair group and issues its
  // atomics in parallel (the old per-channel loop serialised 16 block barriers and put all atomics on 32 threads)
  __shared__ float sm[2][V][256];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    sm[0][i][threadIdx.x] = (pl < lanes) ? s1[i] : 0.f;
    sm[1][i][threadIdx.x] = (pl < lanes) ? s2[i] : 0.f;
  }
  __syncthreads();
  for (int item = threadIdx.x; item < CV * V; item += 256) {
    const int ccv = item % CV, i = item / CV;
    double a = 0.0, b = 0.0;
    for (int l = 0; l < lanes; ++l) {
      a += (double)sm[0][i][l * CV + ccv];
      b += (double)sm[1][i][l * CV + ccv];
    }
    double *dst = part + (((idx_t)n * gridDim.x + blockIdx.x) * C + ccv * V + i) * 2;
    dst[0] = a;
    dst[1] = b;
  }
  __shared__ unsigned long long ticket_s;
  __threadfence();
  __syncthreads();
  unsigned long long *ticket = reinterpret_cast<unsigned long long *>(acc + (idx_t)N * C * 2) + n;
  if (threadIdx.x == 0) ticket_s = atomicAdd(ticket, 1ull);
  __syncthreads();
  if (ticket_s == (unsigned long long)gridDim.x - 1ull) {
    __threadfence();
    const int chunks = (int)gridDim.x;
    for (int item = threadIdx.x; item < C * 2; item += 256) {
      const double *src = part + (idx_t)n * chunks * C * 2 + item;
      double t = 0.0;
      for (int k = 0; k < chunks; ++k) t += __ldcg(src + (idx_t)k * C * 2);
      acc[(idx_t)n * C * 2 + item] = t;
    }
    if (threadIdx.x == 0) *ticket = 0ull;
  }
}

// pass 2: g = fold(gout) + addend;  dx = rstd*(g*act' - mean - xhat*mean(. xhat)), written with a zero margin of `out_pad`
// pixels on every side (the layout the stride-1 input-gradient convolution consumes as a plain VALID convolution).
// grid (pixel chunks, N); block = CV channel vectors x (256 / CV) pixel lanes.  A thread keeps ITS channel vector for the whole kernel,
// so the per-(n, c) constants (mean, rstd and the two reduced means: 24 B per channel, doubles converted once) live in registers instead
// of being re-loaded for every 16-byte output, and the pixel walk needs one division per pixel instead of four per vector
// (the first version moved 224 B of loads per 16 B stored and ran at 12 % of the DRAM bandwidth on 128 x 128 maps).
template <typename T, int V>
__global__ void __launch_bounds__(256) norm_bwd_apply_kernel(const T *__restrict__ gout, const T *__restrict__ x,
                                                             const float *__restrict__ stats, const double *__restrict__ acc,
                                                             const T *__restrict__ addend, T *__restrict__ dx, T *__restrict__ g_out,
                                                             int N, int H, int W, int C, int pad, int act, int out_pad, int pix_per_block) {
  pdl_wait();
  const int CV = C / V;
  const int n = blockIdx.y;
  const int lanes = 256 / CV > 0 ? 256 / CV : 1;
  const int cv = threadIdx.x % CV, pl = threadIdx.x / CV;
  if (pl >= lanes) return;
  const int HW = H * W;
  const int Ho = H + 2 * out_pad, Wo = W + 2 * out_pad;
  const float inv_hw = 1.f / (float)HW;
  float mean[V], rstd[V], m1[V], m2[V];
  if (stats) {
    const idx_t sc0 = (idx_t)n * C + cv * V;
    load_stats<V>(stats + 2 * sc0, mean, rstd);
    load_acc_means<V>(acc + 2 * sc0, inv_hw, m1, m2);
  }
  const int q0 = blockIdx.x * pix_per_block, q1 = min(Ho * Wo, q0 + pix_per_block);
  T *dxn = dx + (idx_t)n * Ho * Wo * C;
  for (int q = q0 + pl; q < q1; q += lanes) {
    const int ho = q / Wo, wo = q - ho * Wo;
    const int h = ho - out_pad, w = wo - out_pad;
    float g[V], o[V];
    if (h < 0 || h >= H || w < 0 || w >= W) {
#pragma unroll
      for (int i = 0; i < V; ++i) o[i] = 0.f;
      store_vec<T, V>(dxn + ((idx_t)q * CV + cv) * V, o);
      continue;
    }
    const idx_t idx = (((idx_t)n * H + h) * W + w) * CV + cv;
    folded_grad<T, V>(gout, n, h, w, cv, H, W, C, pad, g);
    if (addend) {
      float av[V];
      load_vec<T, V>(addend + idx * V, av);
#pragma unroll
      for (int i = 0; i < V; ++i) g[i] += av[i];
    }
    if (g_out) store_vec<T, V>(g_out + idx * V, g);
    if (stats) {
      float xv[V];
      load_vec<T, V>(x + idx * V, xv);
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const float xh = (xv[i] - mean[i]) * rstd[i];
        const float gg = g[i] * act_grad_from_sign(xh, act);
        o[i] = rstd[i] * (gg - m1[i] - xh * m2[i]);
      }
    } else if (act != CTAGAN_ACT_NONE) {
      float xv[V];
      load_vec<T, V>(x + idx * V, xv);
#pragma unroll
      for (int i = 0; i < V; ++i) o[i] = g[i] * act_grad_from_sign(xv[i], act);
    } else {
#pragma unroll
      for (int i = 0; i < V; ++i) o[i] = g[i];
    }
    store_vec<T, V>(dxn + ((idx_t)q * CV + cv) * V, o);
  }
}


// ---------------------------------------------------------------------------------------------------------------------
// InstanceNorm backward in ONE kernel: thread-block clusters + distributed shared memory.
// The two-kernel form (reduce: few CTAs, fp64 atomics; apply: re-reads everything) costs ~20 us per layer at batch 1, more than the
// input-gradient convolution next to it.  Here a cluster of NB_CL CTAs owns NB_CH channels of one image: every thread loads its
// <= NB_PPT pixels ONCE (fold of the reflection-pad gradient, skip-connection addend, activation mask), keeps g and xhat in registers,
// the per-channel sums are combined warp -> CTA (shared memory) -> cluster (each CTA reads its peers' partial sums through DSMEM),
// and the same registers produce dx.  No accumulator, no memset, no atomics, one read of each operand.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int NB_CL = 8;        // CTAs per cluster (portable maximum)
// NB_CVG: channel vectors (of 8 bf16) per CTA; NB_PPT: pixels per thread held in registers.  <2, 4>: 16 channels per cluster, maps up to
// 64x64; <1, 8>: 8 channels per cluster, maps up to 128x128.
template <int NB_CVG, int NB_PPT>
__global__ void __launch_bounds__(256) norm_bwd_cluster_kernel(const bf16 *__restrict__ gout, const bf16 *__restrict__ x,
                                                               const float *__restrict__ stats, const bf16 *__restrict__ addend,
                                                               bf16 *__restrict__ dx, bf16 *__restrict__ g_out, int H, int W, int C, int pad,
                                                               int act, int out_pad) {
  constexpr int V = 8;
  constexpr int NB_LANES = 256 / NB_CVG;   // pixel lanes per CTA
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  pdl_wait();
  const int rank = (int)cluster.block_rank();
  const int group = blockIdx.x / NB_CL;                 // channel group of this cluster
  const int n = blockIdx.y;
  const int cvl = threadIdx.x % NB_CVG, pl = threadIdx.x / NB_CVG;
  const int cv = group * NB_CVG + cvl;
  const int CV = C / V;
  const int HW = H * W;
  const int per_cta = (HW + NB_CL - 1) / NB_CL;
  const int p0 = rank * per_cta, p1 = min(HW, p0 + per_cta);
  const int Ho = H + 2 * out_pad, Wo = W + 2 * out_pad;

  float mean[V], rstd[V];
  load_stats<V>(stats + 2 * ((idx_t)n * C + cv * V), mean, rstd);
  float gg[NB_PPT][V], xh[NB_PPT][V];
  float s1[V], s2[V];
#pragma unroll
  for (int i = 0; i < V; ++i) s1[i] = s2[i] = 0.f;
#pragma unroll
  for (int k = 0; k < NB_PPT; ++k) {
    const int pix = p0 + pl + k * NB_LANES;
    if (pix < p1) {
      const int h = pix / W, w = pix - h * W;
      float g[V], xv[V];
      folded_grad<bf16, V>(gout, n, h, w, cv, H, W, C, pad, g);
      const idx_t idx = ((idx_t)n * HW + pix) * CV + cv;
      if (addend) {
        float av[V];
        load_vec<bf16, V>(addend + idx * V, av);
#pragma unroll
        for (int i = 0; i < V; ++i) g[i] += av[i];
      }
      if (g_out) store_vec<bf16, V>(g_out + idx * V, g);      // the folded (+ skip) gradient itself, for the next skip connection
      load_vec<bf16, V>(x + idx * V, xv);
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const float t = (xv[i] - mean[i]) * rstd[i];
        const float q = g[i] * act_grad_from_sign(t, act);
        xh[k][i] = t;
        gg[k][i] = q;
        s1[i] += q;
        s2[i] = fmaf(q, t, s2[i]);
      }
    } else {
#pragma unroll
      for (int i = 0; i < V; ++i) xh[k][i] = gg[k][i] = 0.f;
    }
  }
  // warp: lanes with the same channel vector are NB_CVG apart
#pragma unroll
  for (int i = 0; i < V; ++i) {
#pragma unroll
    for (int o = NB_CVG; o < 32; o <<= 1) {
      s1[i] += __shfl_xor_sync(0xffffffffu, s1[i], o);
      s2[i] += __shfl_xor_sync(0xffffffffu, s2[i], o);
    }
  }
  __shared__ float warp_part[8][NB_CVG][2][V];
  __shared__ float cta_part[NB_CVG][2][V];          // read by the peers through DSMEM
  __shared__ float total[NB_CVG][2][V];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane < NB_CVG) {
#pragma unroll
    for (int i = 0; i < V; ++i) {
      warp_part[warp][lane][0][i] = s1[i];
      warp_part[warp][lane][1][i] = s2[i];
    }
  }
  __syncthreads();
  if (threadIdx.x < NB_CVG * 2 * V) {
    const int c = threadIdx.x / (2 * V), q = (threadIdx.x / V) % 2, i = threadIdx.x % V;
    float t = 0.f;
#pragma unroll
    for (int wv = 0; wv < 8; ++wv) t += warp_part[wv][c][q][i];
    cta_part[c][q][i] = t;
  }
  cluster.sync();                                     // every CTA's partial sums are visible cluster-wide
  if (threadIdx.x < NB_CVG * 2 * V) {
    const int c = threadIdx.x / (2 * V), q = (threadIdx.x / V) % 2, i = threadIdx.x % V;
    float t = 0.f;
#pragma unroll
    for (int r = 0; r < NB_CL; ++r) {
      const float *peer = cluster.map_shared_rank(&cta_part[0][0][0], r);
      t += peer[(c * 2 + q) * V + i];
    }
    total[c][q][i] = t;
  }
  __syncthreads();
  const float inv_hw = 1.f / (float)HW;
  float m1[V], m2[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    m1[i] = total[cvl][0][i] * inv_hw;
    m2[i] = total[cvl][1][i] * inv_hw;
  }
#pragma unroll
  for (int k = 0; k < NB_PPT; ++k) {
    const int pix = p0 + pl + k * NB_LANES;
    if (pix < p1) {
      const int h = pix / W, w = pix - h * W;
      float o[V];
#pragma unroll
      for (int i = 0; i < V; ++i) o[i] = rstd[i] * (gg[k][i] - m1[i] - xh[k][i] * m2[i]);
      store_vec<bf16, V>(dx + ((((idx_t)n * Ho + h + out_pad) * Wo + w + out_pad) * CV + cv) * V, o);
    }
  }
  if (out_pad > 0) {                                  // zero margin of the output (the dgrad convolution reads it as padding)
    const int margin = Ho * Wo - HW;
    float z[V];
#pragma unroll
    for (int i = 0; i < V; ++i) z[i] = 0.f;
    for (int m = rank * NB_LANES + pl; m < margin; m += NB_CL * NB_LANES) {
      // margin pixels in row-major order: full top rows, then left/right strips of the interior rows, then full bottom rows
      int ho, wo;
      const int top = out_pad * Wo, side = 2 * out_pad;
      if (m < top) { ho = m / Wo; wo = m - ho * Wo; }
      else if (m < top + H * side) { const int t = m - top; ho = out_pad + t / side; const int c2 = t % side; wo = c2 < out_pad ? c2 : W + c2; }
      else { const int t = m - top - H * side; ho = out_pad + H + t / Wo; wo = t % Wo; }
      store_vec<bf16, V>(dx + ((((idx_t)n * Ho + ho) * Wo + wo) * CV + cv) * V, z);
    }
  }
  cluster.sync();                                     // no CTA may exit while a peer still reads its shared memory
}

template <typename T, int V>
__global__ void act_bwd_kernel(const T *__restrict__ gy, const T *__restrict__ y, T *__restrict__ dx, idx_t nvec, int act) {
  for (idx_t idx = (idx_t)blockIdx.x * blockDim.x + threadIdx.x; idx < nvec; idx += (idx_t)gridDim.x * blockDim.x) {
    float g[V], yv[V], o[V];
    load_vec<T, V>(gy + idx * V, g);
    load_vec<T, V>(y + idx * V, yv);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      if (act == CTAGAN_ACT_TANH) o[i] = g[i] * (1.f - yv[i] * yv[i]);
      else o[i] = g[i] * act_grad_from_sign(yv[i], act);
    }
    store_vec<T, V>(dx + idx * V, o);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// MaxPool2d(2)
// ---------------------------------------------------------------------------------------------------------------
template <typename T, int V>
__global__ void maxpool2_fwd_kernel(const T *__restrict__ x, T *__restrict__ y, int N, int H, int W, int C) {
  const int CV = C / V, Ho = H / 2, Wo = W / 2;
  const idx_t total = (idx_t)N * Ho * Wo * CV;
  for (idx_t idx = (idx_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (idx_t)gridDim.x * blockDim.x) {
    const int cv = (int)(idx % CV);
    idx_t r = idx / CV;
    const int ow = (int)(r % Wo);
    r /= Wo;
    const int oh = (int)(r % Ho), n = (int)(r / Ho);
    float m[V];
#pragma unroll
    for (int i = 0; i < V; ++i) m[i] = -INFINITY;
    for (int dy = 0; dy < 2; ++dy)
      for (int dxx = 0; dxx < 2; ++dxx) {
        float v[V];
        load_vec<T, V>(x + (((idx_t)n * H + 2 * oh + dy) * W + 2 * ow + dxx) * C + cv * V, v);
#pragma unroll
        for (int i = 0; i < V; ++i) m[i] = (v[i] > m[i] || v[i] != v[i]) ? v[i] : m[i];
      }
    store_vec<T, V>(y + idx * V, m);
  }
}

template <typename T, int V>
__global__ void maxpool2_bwd_kernel(const T *__restrict__ gy, const T *__restrict__ x, const T *__restrict__ addend, T *__restrict__ gx, int N, int H,
                                    int W, int C) {
  const int CV = C / V, Ho = H / 2, Wo = W / 2;
  const idx_t total = (idx_t)N * Ho * Wo * CV;
  for (idx_t idx = (idx_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (idx_t)gridDim.x * blockDim.x) {
    const int cv = (int)(idx % CV);
    idx_t r = idx / CV;
    const int ow = (int)(r % Wo);
    r /= Wo;
    const int oh = (int)(r % Ho), n = (int)(r / Ho);
    float v[4][V], g[V];
    load_vec<T, V>(gy + idx * V, g);
    for (int k = 0; k < 4; ++k)
      load_vec<T, V>(x + (((idx_t)n * H + 2 * oh + (k >> 1)) * W + 2 * ow + (k & 1)) * C + cv * V, v[k]);
    float o[4][V];
#pragma unroll
    for (int i = 0; i < V; ++i) {
      int arg = 0;
      float m = v[0][i];
#pragma unroll
      for (int k = 1; k < 4; ++k)
        if (v[k][i] > m) { m = v[k][i]; arg = k; }
#pragma unroll
      for (int k = 0; k < 4; ++k) o[k][i] = (k == arg) ? g[i] : 0.f;
    }
    for (int k = 0; k < 4; ++k) {
      const idx_t off = (((idx_t)n * H + 2 * oh + (k >> 1)) * W + 2 * ow + (k & 1)) * C + cv * V;
      if (addend) {
        float av[V];
        load_vec<T, V>(addend + off, av);
#pragma unroll
        for (int i = 0; i < V; ++i) o[k][i] += av[i];
      }
      store_vec<T, V>(gx + off, o[k]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// bilinear 2x upsample (align_corners=False) + concat.  For exact 2x: src = (o + 0.5)/2 - 0.5 clamped at 0,
// i0 = floor(src), i1 = min(i0+1, H-1), lambda = src - i0   (ATen area_pixel_compute_source_index).
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void up2_coords(int o, int n_in, int &i0, int &i1, float &l1) {
  float s = (o + 0.5f) * 0.5f - 0.5f;
  if (s < 0.f) s = 0.f;
  i0 = (int)s;
  i1 = i0 + (i0 < n_in - 1 ? 1 : 0);
  l1 = s - (float)i0;
}

template <typename T, int V>
__global__ void upsample2x_cat_fwd_kernel(const T *__restrict__ x, const T *__restrict__ skip, T *__restrict__ out, int N, int H, int W,
                                          int C1, int C2) {
  const int C = C1 + C2, CV = C / V, CV1 = C1 / V;
  const int Ho = 2 * H, Wo = 2 * W;
  const idx_t total = (idx_t)N * Ho * Wo * CV;
  for (idx_t idx = (idx_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (idx_t)gridDim.x * blockDim.x) {
    const int cv = (int)(idx % CV);
    idx_t r = idx / CV;
    const int ow = (int)(r % Wo);
    r /= Wo;
    const int oh = (int)(r % Ho), n = (int)(r / Ho);
    float o[V];
    if (cv < CV1) {
      int h0, h1, w0, w1;
      float lh, lw;
      up2_coords(oh, H, h0, h1, lh);
      up2_coords(ow, W, w0, w1, lw);
      float a[V], b[V], c[V], d[V];
      const T *base = x + (idx_t)n * H * W * C1 + cv * V;
      load_vec<T, V>(base + ((idx_t)h0 * W + w0) * C1, a);
      load_vec<T, V>(base + ((idx_t)h0 * W + w1) * C1, b);
      load_vec<T, V>(base + ((idx_t)h1 * W + w0) * C1, c);
      load_vec<T, V>(base + ((idx_t)h1 * W + w1) * C1, d);
      const float hh0 = 1.f - lh, ww0 = 1.f - lw;
#pragma unroll
      for (int i = 0; i < V; ++i) o[i] = hh0 * (ww0 * a[i] + lw * b[i]) + lh * (ww0 * c[i] + lw * d[i]);
    } else {
      load_vec<T, V>(skip + (((idx_t)n * Ho + oh) * Wo + ow) * C2 + (cv - CV1) * V, o);
    }
    store_vec<T, V>(out + idx * V, o);
  }
}

// gather form of the transposed bilinear operator: input pixel (h,w) collects from the <=3x3 output pixels that read it.
template <typename T, int V>
__global__ void upsample2x_cat_bwd_kernel(const T *__restrict__ gout, T *__restrict__ gx, T *__restrict__ gskip, int N, int H, int W,
                                          int C1, int C2) {
  const int C = C1 + C2, CV1 = C1 / V, CV2 = C2 / V;
  const int Ho = 2 * H, Wo = 2 * W;
  const idx_t total1 = (idx_t)N * H * W * CV1;
  const idx_t total2 = gskip ? (idx_t)N * Ho * Wo * CV2 : 0;
  for (idx_t idx = (idx_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total1 + total2; idx += (idx_t)gridDim.x * blockDim.x) {
    if (idx < total1) {
      const int cv = (int)(idx % CV1);
      idx_t r = idx / CV1;
      const int w = (int)(r % W);
      r /= W;
      const int h = (int)(r % H), n = (int)(r / H);
      float acc[V];
#pragma unroll
      for (int i = 0; i < V; ++i) acc[i] = 0.f;
      for (int oh = max(0, 2 * h - 2); oh <= min(Ho - 1, 2 * h + 2); ++oh) {
        int h0, h1;
        float lh;
        up2_coords(oh, H, h0, h1, lh);
        float wh = 0.f;
        if (h0 == h) wh += 1.f - lh;
        if (h1 == h) wh += lh;
        if (wh == 0.f) continue;
        for (int ow = max(0, 2 * w - 2); ow <= min(Wo - 1, 2 * w + 2); ++ow) {
          int w0, w1;
          float lw;
          up2_coords(ow, W, w0, w1, lw);
          float ww = 0.f;
          if (w0 == w) ww += 1.f - lw;
          if (w1 == w) ww += lw;
          if (ww == 0.f) continue;
          float g[V];
          load_vec<T, V>(gout + (((idx_t)n * Ho + oh) * Wo + ow) * C + cv * V, g);
#pragma unroll
          for (int i = 0; i < V; ++i) acc[i] = fmaf(wh * ww, g[i], acc[i]);
        }
      }
      store_vec<T, V>(gx + idx * V, acc);
    } else {
      const idx_t j = idx - total1;
      const int cv = (int)(j % CV2);
      const idx_t pix = j / CV2;
      float g[V];
      load_vec<T, V>(gout + pix * C + C1 + cv * V, g);
      store_vec<T, V>(gskip + j * V, g);
    }
  }
}

template <typename T>
__global__ void copy_channels_kernel(const T *__restrict__ src, T *__restrict__ dst, idx_t pixels, int C, int ss, int so, int ds, int doff) {
  const idx_t total = pixels * C;
  for (idx_t idx = (idx_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (idx_t)gridDim.x * blockDim.x) {
    const idx_t p = idx / C;
    const int c = (int)(idx - p * C);
    dst[p * ds + doff + c] = src[p * ss + so + c];
  }
}

// plane mean: one warp per (n, c-group); small C (1) in practice.
template <typename T>
__global__ void plane_mean_fwd_kernel(const T *__restrict__ x, float *__restrict__ out, int HW, int C) {
  const int n = blockIdx.x, c = blockIdx.y;
  double s = 0.0;
  for (int p = threadIdx.x; p < HW; p += blockDim.x) s += (double)to_f(x[((idx_t)n * HW + p) * C + c]);
  __shared__ double sm[32];
  s = warp_sum_d(s);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < (blockDim.x >> 5); ++i) t += sm[i];
    out[n * C + c] = (float)(t / HW);
  }
}

template <typename T>
__global__ void plane_mean_bwd_kernel(const float *__restrict__ gout, T *__restrict__ gx, int N, int HW, int C) {
  const idx_t total = (idx_t)N * HW * C;
  const float inv = 1.f / (float)HW;
  for (idx_t idx = (idx_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (idx_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C);
    const int n = (int)(idx / ((idx_t)HW * C));
    gx[idx] = from_f<T>(gout[n * C + c] * inv);
  }
}

template <typename S, typename D>
__global__ void cast_kernel(const S *__restrict__ s, D *__restrict__ d, idx_t n) {
  for (idx_t idx = (idx_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (idx_t)gridDim.x * blockDim.x)
    d[idx] = from_f<D>(to_f(s[idx]));
}

// fp32 NCHW (module boundary) <-> T NHWC (internal)
template <typename T>
__global__ void nchw_to_nhwc_kernel(const float *__restrict__ src, T *__restrict__ dst, int N, int C, idx_t HW) {
  const idx_t total = (idx_t)N * HW * C;
  for (idx_t idx = (idx_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (idx_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C);
    const idx_t r = idx / C;
    const idx_t p = r % HW;
    const int n = (int)(r / HW);
    dst[idx] = from_f<T>(src[((idx_t)n * C + c) * HW + p]);
  }
}
template <typename T>
__global__ void nhwc_to_nchw_kernel(const T *__restrict__ src, float *__restrict__ dst, int N, int C, idx_t HW) {
  const idx_t total = (idx_t)N * HW * C;
  for (idx_t idx = (idx_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (idx_t)gridDim.x * blockDim.x) {
    const idx_t p = idx % HW;
    const idx_t r = idx / HW;
    const int c = (int)(r % C);
    const int n = (int)(r / C);
    dst[idx] = to_f(src[((idx_t)n * HW + p) * C + c]);
  }
}

// two 1-channel fp32 planes <-> one 2-channel NHWC tensor (torch.cat([a, b], 1) of trainer/reg.py:77 and its gradient split)
template <typename T>
__global__ void interleave2_kernel(const float *__restrict__ a, const float *__restrict__ b, T *__restrict__ dst, idx_t n) {
  for (idx_t i = (idx_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (idx_t)gridDim.x * blockDim.x) {
    dst[2 * i] = from_f<T>(a[i]);
    dst[2 * i + 1] = from_f<T>(b[i]);
  }
}
template <typename T>
__global__ void deinterleave2_kernel(const T *__restrict__ src, float *__restrict__ a, float *__restrict__ b, idx_t n) {
  for (idx_t i = (idx_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (idx_t)gridDim.x * blockDim.x) {
    if (a) a[i] = to_f(src[2 * i]);
    if (b) b[i] = to_f(src[2 * i + 1]);
  }
}

inline int ew_blocks(idx_t work) {
  idx_t b = (work + 255) / 256;
  const idx_t cap = (idx_t)ctagan_num_sms() * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

// choose the widest vector V in {max, 4, 2, 1} dividing all the given channel counts
template <typename T>
int pick_vec(int c1, int c2 = 0) {
  for (int v = max_vec<T>(); v > 1; v >>= 1)
    if (c1 % v == 0 && c2 % v == 0) return v;
  return 1;
}

}  // namespace

#define VEC_SWITCH(T, v, V, ...)                                   \
  do {                                                             \
    if (v == 8) { constexpr int V = sizeof(T) == 2 ? 8 : 4; __VA_ARGS__; }  \
    else if (v == 4) { constexpr int V = 4; __VA_ARGS__; }         \
    else if (v == 2) { constexpr int V = 2; __VA_ARGS__; }         \
    else { constexpr int V = 1; __VA_ARGS__; }                     \
  } while (0)

static int reduce_chunks(int N, int HW, int C, int v, int &pix_per_block) {
  const int CV = C / v;
  const int lanes = 256 / CV > 0 ? 256 / CV : 1;
  // aim for ~4 CTAs per SM overall, at least 8 pixels per lane
  idx_t want = ((idx_t)ctagan_num_sms() * 4 + N - 1) / N;
  static int ppl = 0;
  if (!ppl) { const char *e = getenv("CTAGAN_RED_PPL"); ppl = e ? atoi(e) : 16; if (ppl < 1) ppl = 16; }
  idx_t maxc = (HW + (idx_t)lanes * ppl - 1) / ((idx_t)lanes * ppl);
  if (want > maxc) want = maxc;
  if (want < 1) want = 1;
  pix_per_block = (int)((HW + want - 1) / want);
  return (int)((HW + pix_per_block - 1) / pix_per_block);
}

template <typename T>
static int stats_vec(int C) {
  int v = pick_vec<T>(C);
  while (C / v > 256 && v < max_vec<T>()) v *= 2;
  return v;
}

// doubles of scratch ctagan_instnorm_stats / the two-kernel ctagan_norm_act_pad_bwd need: one row of C*2 partial sums per block
static size_t reduce_scratch_doubles(int N, int HW, int C, int dtype) {
  int ppb;
  const int v = dtype == CTAGAN_BF16 ? stats_vec<bf16>(C) : stats_vec<float>(C);
  return (size_t)N * reduce_chunks(N, HW, C, v, ppb) * C * 2;
}

extern "C" size_t ctagan_instnorm_stats_scratch_doubles(int N, int HW, int C, int dtype) {
  if (N <= 0 || HW <= 0 || C <= 0) return 0;
  return reduce_scratch_doubles(N, HW, C, dtype);
}

extern "C" int ctagan_instnorm_stats(const void *x, float *stats, double *acc, int N, int HW, int C, int dtype, void *stream) {
  CTAGAN_REQUIRE(x && stats && acc && N > 0 && HW > 0 && C > 0, "instnorm_stats: bad arguments");
  CTAGAN_FITS32((int64_t)N * HW * C);
  cudaStream_t st = (cudaStream_t)stream;
  CTAGAN_DISPATCH_DTYPE(dtype, T, {
    int v = pick_vec<T>(C);
    while (C / v > 256) { CTAGAN_REQUIRE(v < max_vec<T>(), "instnorm_stats: C=%d too large", C); v *= 2; }
    int ppb;
    const int chunks = reduce_chunks(N, HW, C, v, ppb);
    dim3 grid(chunks, N);
    VEC_SWITCH(T, v, V, instnorm_partial_kernel<T, V><<<grid, 256, 0, st>>>((const T *)x, acc, HW, C, ppb));
    instnorm_finalize_kernel<T><<<cdiv((idx_t)N * C, 128), 128, 0, st>>>((const T *)x, acc, stats, N, HW, C, chunks);
  });
  CTAGAN_LAUNCH_OK();
  return CTAGAN_OK;
}

extern "C" int ctagan_norm_act_pad(const void *x, const float *stats, const void *res, int res_pad, void *out, int N, int H, int W, int C,
                                   int pad, int act, int dtype, void *stream) {
  CTAGAN_REQUIRE(x && out && N > 0 && H > 0 && W > 0 && C > 0 && pad >= 0 && pad < H && pad < W, "norm_act_pad: bad arguments");
  CTAGAN_FITS32((int64_t)N * (H + 2 * pad) * (W + 2 * pad) * C);
  cudaStream_t st = (cudaStream_t)stream;
  CTAGAN_DISPATCH_DTYPE(dtype, T, {
    const int v = pick_vec<T>(C);
    CTAGAN_REQUIRE(C / v <= 256, "norm_act_pad: at most 256 channel vectors per pixel");
    const int out_px = (H + 2 * pad) * (W + 2 * pad);
    const int lanes = 256 / (C / v) > 0 ? 256 / (C / v) : 1;
    int chunks = (16 * ctagan_num_sms() + N - 1) / N;             // ~16 blocks per SM (two waves of full occupancy), at least 4 pixels per lane
    const int max_chunks = (out_px + 4 * lanes - 1) / (4 * lanes);
    if (chunks > max_chunks) chunks = max_chunks;
    if (chunks < 1) chunks = 1;
    const int ppb = (out_px + chunks - 1) / chunks;
    chunks = (out_px + ppb - 1) / ppb;
    VEC_SWITCH(T, v, V, CTAGAN_CUDA_OK(launch_pdl(norm_act_pad_kernel<T, V>, dim3(chunks, N), dim3(256), 0, st, (const T *)x, stats, (const T *)res, res_pad, (T *)out, N, H, W, C, pad, act, ppb)));
  });
  CTAGAN_LAUNCH_OK();
  return CTAGAN_OK;
}

// 0: two-kernel path; 1: cluster kernel <2, 4> (maps <= 64x64); 2: cluster kernel <1, 8> (maps <= 128x128)
static int norm_bwd_cluster_kind(const float *stats, int H, int W, int C, int dtype) {
  const char *e = getenv("CTAGAN_NORM_BWD_CLUSTER");
  if ((e && e[0] == '0') || !stats || dtype != CTAGAN_BF16) return 0;
  const long long hw = (long long)H * W;
  if (C % 16 == 0 && hw <= (long long)NB_CL * 128 * 4) return 1;
  // the <1, 8> variant (203 registers) is correct but measured slower than the two-kernel form on the Reg step (17.15 vs 16.85 ms)
  // and neutral on the Cyc step: opt-in with CTAGAN_NORM_BWD_CLUSTER=2
  if (e && e[0] == '2' && C % 8 == 0 && hw <= (long long)NB_CL * 256 * 8) return 2;
  return 0;
}
static bool norm_bwd_uses_cluster(const float *stats, int H, int W, int C, int dtype) { return norm_bwd_cluster_kind(stats, H, W, C, dtype) != 0; }

extern "C" int ctagan_norm_act_pad_bwd_launches(int has_stats, int H, int W, int C, int dtype) {
  if (!has_stats) return 1;
  return norm_bwd_uses_cluster((const float *)1, H, W, C, dtype) ? 1 : 2;
}

extern "C" size_t ctagan_norm_act_pad_bwd_scratch_doubles(int has_stats, int N, int H, int W, int C, int dtype) {
  if (!has_stats || N <= 0 || H <= 0 || W <= 0 || C <= 0 || norm_bwd_uses_cluster((const float *)1, H, W, C, dtype)) return 0;
  return reduce_scratch_doubles(N, H * W, C, dtype);
}

extern "C" int ctagan_norm_act_pad_bwd(const void *gout, const void *x, const float *stats, const void *addend, void *dx, void *g_out,
                                       double *acc, int acc_is_zero, double *scratch, int N, int H, int W, int C, int pad, int act,
                                       int out_pad, int dtype, void *stream) {
  CTAGAN_REQUIRE(gout && dx && N > 0 && H > 0 && W > 0 && C > 0 && pad >= 0 && out_pad >= 0, "norm_act_pad_bwd: bad arguments");
  CTAGAN_FITS32((int64_t)N * (H + 2 * pad + 2 * out_pad) * (W + 2 * pad + 2 * out_pad) * C);
  CTAGAN_REQUIRE(!(stats || act != CTAGAN_ACT_NONE) || x, "norm_act_pad_bwd: x required when stats/act given");
  CTAGAN_REQUIRE(!stats || (acc && scratch) || norm_bwd_uses_cluster(stats, H, W, C, dtype), "norm_act_pad_bwd: acc and scratch required with stats");
  CTAGAN_REQUIRE(act != CTAGAN_ACT_TANH, "norm_act_pad_bwd: tanh unsupported here (use act_bwd)");
  cudaStream_t st = (cudaStream_t)stream;
  if (const int kind = norm_bwd_cluster_kind(stats, H, W, C, dtype)) {
    // one-kernel cluster path (bf16): everything a thread needs fits in registers
    if (kind == 1) {
      dim3 grid((unsigned)(C / 16 * NB_CL), (unsigned)N);
      CTAGAN_CUDA_OK(launch_cluster_pdl(norm_bwd_cluster_kernel<2, 4>, grid, dim3(256), 0, st, (unsigned)NB_CL, (const bf16 *)gout, (const bf16 *)x, stats,
                                        (const bf16 *)addend, (bf16 *)dx, (bf16 *)g_out, H, W, C, pad, act, out_pad));
    } else {
      dim3 grid((unsigned)(C / 8 * NB_CL), (unsigned)N);
      CTAGAN_CUDA_OK(launch_cluster_pdl(norm_bwd_cluster_kernel<1, 8>, grid, dim3(256), 0, st, (unsigned)NB_CL, (const bf16 *)gout, (const bf16 *)x, stats,
                                        (const bf16 *)addend, (bf16 *)dx, (bf16 *)g_out, H, W, C, pad, act, out_pad));
    }
    CTAGAN_LAUNCH_OK();
    return CTAGAN_OK;
  }
  CTAGAN_DISPATCH_DTYPE(dtype, T, {
    int v = pick_vec<T>(C);
    if (stats) {
      while (C / v > 256) { CTAGAN_REQUIRE(v < max_vec<T>(), "norm_act_pad_bwd: C=%d too large", C); v *= 2; }
      if (!acc_is_zero) CTAGAN_CUDA_OK(cudaMemsetAsync(acc + 2 * (size_t)N * C, 0, sizeof(double) * (size_t)N, st));     // the tickets
      int ppb;
      const int chunks = reduce_chunks(N, H * W, C, v, ppb);
      dim3 grid(chunks, N);
      VEC_SWITCH(T, v, V, CTAGAN_CUDA_OK(launch_pdl(norm_bwd_reduce_kernel<T, V>, grid, dim3(256), 0, st, (const T *)gout, (const T *)x, stats, (const T *)addend, acc, scratch, N, H, W, C, pad, act, ppb)));
    }
    {
      // pixel chunks x images: ~16 blocks per SM, at least 4 pixels per lane
      CTAGAN_REQUIRE(C / v <= 256, "norm_act_pad_bwd: at most 256 channel vectors per pixel");
      const int out_px = (H + 2 * out_pad) * (W + 2 * out_pad);
      const int lanes = 256 / (C / v) > 0 ? 256 / (C / v) : 1;
      int chunks = (16 * ctagan_num_sms() + N - 1) / N;
      const int max_chunks = (out_px + 4 * lanes - 1) / (4 * lanes);
      if (chunks > max_chunks) chunks = max_chunks;
      if (chunks < 1) chunks = 1;
      const int ppb = (out_px + chunks - 1) / chunks;
      chunks = (out_px + ppb - 1) / ppb;
      VEC_SWITCH(T, v, V, CTAGAN_CUDA_OK(launch_pdl(norm_bwd_apply_kernel<T, V>, dim3(chunks, N), dim3(256), 0, st, (const T *)gout, (const T *)x, stats, acc, (const T *)addend, (T *)dx, (T *)g_out, N, H, W, C, pad, act, out_pad, ppb)));
    }
  });
  CTAGAN_LAUNCH_OK();
  return CTAGAN_OK;
}

extern "C" int ctagan_act_bwd(const void *gy, const void *y, void *dx, int64_t n, int act, int dtype, void *stream) {
  CTAGAN_REQUIRE(gy && y && dx && n > 0, "act_bwd: bad arguments");
  CTAGAN_FITS32(n);
  cudaStream_t st = (cudaStream_t)stream;
  CTAGAN_DISPATCH_DTYPE(dtype, T, {
    int v = max_vec<T>();
    while (n % v) v >>= 1;
    VEC_SWITCH(T, v, V, act_bwd_kernel<T, V><<<ew_blocks(n / v), 256, 0, st>>>((const T *)gy, (const T *)y, (T *)dx, n / v, act));
  });
  CTAGAN_LAUNCH_OK();
  return CTAGAN_OK;
}

extern "C" int ctagan_maxpool2_fwd(const void *x, void *y, int N, int H, int W, int C, int dtype, void *stream) {
  CTAGAN_REQUIRE(x && y && N > 0 && H > 0 && W > 0 && C > 0 && H % 2 == 0 && W % 2 == 0, "maxpool2_fwd: bad arguments (even H, W required)");
  CTAGAN_FITS32((int64_t)N * H * W * C);
  cudaStream_t st = (cudaStream_t)stream;
  CTAGAN_DISPATCH_DTYPE(dtype, T, {
    const int v = pick_vec<T>(C);
    const idx_t total = (idx_t)N * (H / 2) * (W / 2) * (C / v);
    VEC_SWITCH(T, v, V, maxpool2_fwd_kernel<T, V><<<ew_blocks(total), 256, 0, st>>>((const T *)x, (T *)y, N, H, W, C));
  });
  CTAGAN_LAUNCH_OK();
  return CTAGAN_OK;
}

extern "C" int ctagan_maxpool2_bwd(const void *gy, const void *x, const void *addend, void *gx, int N, int H, int W, int C, int dtype, void *stream) {
  CTAGAN_REQUIRE(gy && x && gx && N > 0 && H > 0 && W > 0 && C > 0 && H % 2 == 0 && W % 2 == 0, "maxpool2_bwd: bad arguments");
  CTAGAN_FITS32((int64_t)N * H * W * C);
  cudaStream_t st = (cudaStream_t)stream;
  CTAGAN_DISPATCH_DTYPE(dtype, T, {
    const int v = pick_vec<T>(C);
    const idx_t total = (idx_t)N * (H / 2) * (W / 2) * (C / v);
    VEC_SWITCH(T, v, V, maxpool2_bwd_kernel<T, V><<<ew_blocks(total), 256, 0, st>>>((const T *)gy, (const T *)x, (const T *)addend, (T *)gx, N, H, W, C));
  });
  CTAGAN_LAUNCH_OK();
  return CTAGAN_OK;
}

extern "C" int ctagan_upsample2x_cat_fwd(const void *x, const void *skip, void *out, int N, int H, int W, int C1, int C2, int dtype,
                                         void *stream) {
  CTAGAN_REQUIRE(x && out && N > 0 && H > 0 && W > 0 && C1 > 0 && C2 >= 0 && (C2 == 0 || skip), "upsample2x_cat_fwd: bad arguments");
  CTAGAN_FITS32((int64_t)N * 4 * H * W * (C1 + C2));
  cudaStream_t st = (cudaStream_t)stream;
  CTAGAN_DISPATCH_DTYPE(dtype, T, {
    const int v = pick_vec<T>(C1, C2);
    const idx_t total = (idx_t)N * 4 * H * W * ((C1 + C2) / v);
    VEC_SWITCH(T, v, V, upsample2x_cat_fwd_kernel<T, V><<<ew_blocks(total), 256, 0, st>>>((const T *)x, (const T *)skip, (T *)out, N, H, W, C1, C2));
  });
  CTAGAN_LAUNCH_OK();
  return CTAGAN_OK;
}

extern "C" int ctagan_upsample2x_cat_bwd(const void *gout, void *gx, void *gskip, int N, int H, int W, int C1, int C2, int dtype,
                                         void *stream) {
  CTAGAN_REQUIRE(gout && gx && N > 0 && H > 0 && W > 0 && C1 > 0 && C2 >= 0, "upsample2x_cat_bwd: bad arguments");
  CTAGAN_FITS32((int64_t)N * 4 * H * W * (C1 + C2));
  cudaStream_t st = (cudaStream_t)stream;
  CTAGAN_DISPATCH_DTYPE(dtype, T, {
    const int v = pick_vec<T>(C1, C2);
    const idx_t total = (idx_t)N * H * W * (C1 / v) + (gskip ? (idx_t)N * 4 * H * W * (C2 / v) : 0);
    VEC_SWITCH(T, v, V, upsample2x_cat_bwd_kernel<T, V><<<ew_blocks(total), 256, 0, st>>>((const T *)gout, (T *)gx, (T *)gskip, N, H, W, C1, C2));
  });
  CTAGAN_LAUNCH_OK();
  return CTAGAN_OK;
}

extern "C" int ctagan_copy_channels(const void *src, void *dst, int64_t pixels, int C, int src_stride, int src_off, int dst_stride,
                                    int dst_off, int dtype, void *stream) {
  CTAGAN_REQUIRE(src && dst && pixels > 0 && C > 0 && src_off + C <= src_stride && dst_off + C <= dst_stride, "copy_channels: bad arguments");
  CTAGAN_FITS32((int64_t)pixels * (src_stride > dst_stride ? src_stride : dst_stride));
  cudaStream_t st = (cudaStream_t)stream;
  CTAGAN_DISPATCH_DTYPE(dtype, T, {
    copy_channels_kernel<T><<<ew_blocks(pixels * C), 256, 0, st>>>((const T *)src, (T *)dst, pixels, C, src_stride, src_off, dst_stride, dst_off);
  });
  CTAGAN_LAUNCH_OK();
  return CTAGAN_OK;
}

extern "C" int ctagan_plane_mean_fwd(const void *x, float *out, int N, int HW, int C, int dtype, void *stream) {
  CTAGAN_REQUIRE(x && out && N > 0 && HW > 0 && C > 0 && C <= 65535, "plane_mean_fwd: bad arguments");
  CTAGAN_FITS32((int64_t)N * HW * C);
  cudaStream_t st = (cudaStream_t)stream;
  CTAGAN_DISPATCH_DTYPE(dtype, T, { plane_mean_fwd_kernel<T><<<dim3(N, C), 256, 0, st>>>((const T *)x, out, HW, C); });
  CTAGAN_LAUNCH_OK();
  return CTAGAN_OK;
}

extern "C" int ctagan_plane_mean_bwd(const float *gout, void *gx, int N, int HW, int C, int dtype, void *stream) {
  CTAGAN_REQUIRE(gout && gx && N > 0 && HW > 0 && C > 0, "plane_mean_bwd: bad arguments");
  CTAGAN_FITS32((int64_t)N * HW * C);
  cudaStream_t st = (cudaStream_t)stream;
  CTAGAN_DISPATCH_DTYPE(dtype, T, { plane_mean_bwd_kernel<T><<<ew_blocks((idx_t)N * HW * C), 256, 0, st>>>(gout, (T *)gx, N, HW, C); });
  CTAGAN_LAUNCH_OK();
  return CTAGAN_OK;
}

extern "C" int ctagan_cast(const void *src, int sd, void *dst, int dd, int64_t n, void *stream) {
  CTAGAN_REQUIRE(src && dst && n > 0, "cast: bad arguments");
  CTAGAN_FITS32(n);
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = ew_blocks(n);
  if (sd == CTAGAN_F32 && dd == CTAGAN_BF16) cast_kernel<float, bf16><<<blocks, 256, 0, st>>>((const float *)src, (bf16 *)dst, n);
  else if (sd == CTAGAN_BF16 && dd == CTAGAN_F32) cast_kernel<bf16, float><<<blocks, 256, 0, st>>>((const bf16 *)src, (float *)dst, n);
  else if (sd == CTAGAN_F32 && dd == CTAGAN_F32) cast_kernel<float, float><<<blocks, 256, 0, st>>>((const float *)src, (float *)dst, n);
  else if (sd == CTAGAN_BF16 && dd == CTAGAN_BF16) cast_kernel<bf16, bf16><<<blocks, 256, 0, st>>>((const bf16 *)src, (bf16 *)dst, n);
  else { ctagan_set_error("cast: bad dtypes"); return CTAGAN_ERR_ARG; }
  CTAGAN_LAUNCH_OK();
  return CTAGAN_OK;
}

extern "C" int ctagan_nchw_to_nhwc(const float *src, void *dst, int N, int C, int64_t HW, int dtype, void *stream) {
  CTAGAN_REQUIRE(src && dst && N > 0 && C > 0 && HW > 0, "nchw_to_nhwc: bad arguments");
  CTAGAN_FITS32((int64_t)N * C * HW);
  cudaStream_t st = (cudaStream_t)stream;
  CTAGAN_DISPATCH_DTYPE(dtype, T, { nchw_to_nhwc_kernel<T><<<ew_blocks((idx_t)N * C * HW), 256, 0, st>>>(src, (T *)dst, N, C, HW); });
  CTAGAN_LAUNCH_OK();
  return CTAGAN_OK;
}

extern "C" int ctagan_nhwc_to_nchw(const void *src, float *dst, int N, int C, int64_t HW, int dtype, void *stream) {
  CTAGAN_REQUIRE(src && dst && N > 0 && C > 0 && HW > 0, "nhwc_to_nchw: bad arguments");
  CTAGAN_FITS32((int64_t)N * C * HW);
  cudaStream_t st = (cudaStream_t)stream;
  CTAGAN_DISPATCH_DTYPE(dtype, T, { nhwc_to_nchw_kernel<T><<<ew_blocks((idx_t)N * C * HW), 256, 0, st>>>((const T *)src, dst, N, C, HW); });
  CTAGAN_LAUNCH_OK();
  return CTAGAN_OK;
}

extern "C" int ctagan_interleave2(const float *a, const float *b, void *dst, int64_t n, int dtype, void *stream) {
  CTAGAN_REQUIRE(a && b && dst && n > 0, "interleave2: bad arguments");
  CTAGAN_FITS32(2 * n);
  cudaStream_t st = (cudaStream_t)stream;
  CTAGAN_DISPATCH_DTYPE(dtype, T, { interleave2_kernel<T><<<ew_blocks(n), 256, 0, st>>>(a, b, (T *)dst, n); });
  CTAGAN_LAUNCH_OK();
  return CTAGAN_OK;
}

extern "C" int ctagan_deinterleave2(const void *src, float *a, float *b, int64_t n, int dtype, void *stream) {
  CTAGAN_REQUIRE(src && (a || b) && n > 0, "deinterleave2: bad arguments");
  CTAGAN_FITS32(2 * n);
  cudaStream_t st = (cudaStream_t)stream;
  CTAGAN_DISPATCH_DTYPE(dtype, T, { deinterleave2_kernel<T><<<ew_blocks(n), 256, 0, st>>>((const T *)src, a, b, n); });
  CTAGAN_LAUNCH_OK();
  return CTAGAN_OK;
}
