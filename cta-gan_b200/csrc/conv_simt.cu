// CUDA-core (FFMA, fp32 accumulate) implicit-GEMM convolution kernels.
//
// These are the *validation-mode* engine (fp32 activations, <=1e-4 vs the fp32 oracle: tcgen05 has no IEEE-fp32 MMA) and the
// engine for shapes that do not tile onto tensor cores (Cin=1 7x7 head, Cout<=2 tails, <=16x16 maps of Reg).
// One "gather" geometry (ctagan_conv_geom) covers Conv2d fwd/dgrad and ConvTranspose2d fwd/dgrad; see include/ctagan.h.
#include "common.cuh"

namespace {

constexpr int BM = 64, BN = 64, BK = 16, PADM = 4;

struct PixelCoord {
  int n, oh, ow;
  bool valid;
};

__device__ __forceinline__ bool tap_coord(const ctagan_conv_geom &g, int o, int k, int pad, int in_extent, int &i_out) {
  int num = o * g.stride + k - pad;
  if (num < 0) return false;
  if (g.dil > 1) {
    if (num % g.dil) return false;
    num /= g.dil;
  }
  i_out = num;
  return num < in_extent;
}

// y[m, co] = act(bias + sum_k A[m,k] * W[co,k]);  A gathered from x on the fly.
// FLAT=false requires Ci % 16 == 0 (vector loads, K chunks never straddle taps); FLAT=true handles any Ci.
template <typename T, bool FLAT>
__global__ void __launch_bounds__(256) conv_gather_simt_kernel(ctagan_conv_geom g, const T *__restrict__ x,
                                                               const T *__restrict__ wp, const float *__restrict__ bias,
                                                               T *__restrict__ y) {
  __shared__ __align__(16) float As[BK][BM + PADM];
  __shared__ __align__(16) float Bs[BK][BN + PADM];
  const int t = threadIdx.x;
  const int ty = t >> 4, tx = t & 15;
  const long long M = (long long)g.N * g.Ho * g.Wo;
  const int ntaps = g.KH * g.KW;
  const int Ktot = ntaps * g.Ci;

  // loader coordinates: row r (pixel for A, cout for B), 4 consecutive k at kseg
  const int r = t >> 2, kseg = (t & 3) * 4;
  const long long m = (long long)blockIdx.x * BM + r;
  PixelCoord pc;
  pc.valid = m < M;
  {
    long long mm = pc.valid ? m : 0;
    pc.ow = (int)(mm % g.Wo);
    long long q = mm / g.Wo;
    pc.oh = (int)(q % g.Ho);
    pc.n = (int)(q / g.Ho);
  }
  const int co_l = blockIdx.y * BN + r;
  const bool co_ok = co_l < g.Co;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int nk = FLAT ? (Ktot + BK - 1) / BK : ntaps * (g.Ci / BK);
  const int cchunks = FLAT ? 1 : g.Ci / BK;
  for (int it = 0; it < nk; ++it) {
    float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
    if constexpr (!FLAT) {
      const int tap = it / cchunks, c0 = (it - tap * cchunks) * BK;
      const int kh = tap / g.KW, kw = tap - kh * g.KW;
      int ih, iw;
      if (pc.valid && tap_coord(g, pc.oh, kh, g.pad_h, g.Hi, ih) && tap_coord(g, pc.ow, kw, g.pad_w, g.Wi, iw)) {
        const T *p = x + (((long long)pc.n * g.Hi + ih) * g.Wi + iw) * g.Ci + c0 + kseg;
        load_vec<T, 4>(p, a);
      }
      if (co_ok) {
        const T *q = wp + ((long long)co_l * ntaps + tap) * g.Ci + c0 + kseg;
        load_vec<T, 4>(q, b);
      }
    } else {
      const int k0 = it * BK + kseg;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int kk = k0 + j;
        if (kk < Ktot) {
          const int tap = kk / g.Ci, ci = kk - tap * g.Ci;
          const int kh = tap / g.KW, kw = tap - kh * g.KW;
          int ih, iw;
          if (pc.valid && tap_coord(g, pc.oh, kh, g.pad_h, g.Hi, ih) && tap_coord(g, pc.ow, kw, g.pad_w, g.Wi, iw))
            a[j] = to_f(x[(((long long)pc.n * g.Hi + ih) * g.Wi + iw) * g.Ci + ci]);
          if (co_ok) b[j] = to_f(wp[(long long)co_l * Ktot + kk]);
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      As[kseg + j][r] = a[j];
      Bs[kseg + j][r] = b[j];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 av = *reinterpret_cast<const float4 *>(&As[k][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4 *>(&Bs[k][tx * 4]);
      const float aa[4] = {av.x, av.y, av.z, av.w}, bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
    }
  }

  const int co0 = blockIdx.y * BN + tx * 4;
  float bv[4] = {0.f, 0.f, 0.f, 0.f};
  if (bias) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (co0 + j < g.Co) bv[j] = bias[co0 + j];
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long mo = (long long)blockIdx.x * BM + ty * 4 + i;
    if (mo >= M) continue;
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = apply_act(acc[i][j] + bv[j], g.act);
    T *dst = y + mo * g.Co + co0;
    if ((g.Co & 3) == 0 && co0 + 3 < g.Co) {
      store_vec<T, 4>(dst, o);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (co0 + j < g.Co) dst[j] = from_f<T>(o[j]);
    }
  }
}

// dw[a, b, kh, kw] (+)= sum_p gy[p, a] * gx[gather(p, kh, kw), b];   tile 64 a x 64 n' (n' = tap*B + b), K = pixels.
template <typename T>
__global__ void __launch_bounds__(256) conv_wgrad_simt_kernel(ctagan_conv_geom g, const T *__restrict__ gy,
                                                              const T *__restrict__ gx, float *__restrict__ dw,
                                                              float *__restrict__ db, int pixels_per_split, int packed) {
  // split-K: split z stores its partial sums in row z of dw[gridDim.z][Co*Ci*taps] / db[gridDim.z][Co] (the host points dw / db at the
  // workspace and adds the rows in order with ctagan_ordered_sum when there is more than one split: no floating-point atomics)
  dw += (long long)blockIdx.z * g.Co * g.Ci * g.KH * g.KW;
  if (db != nullptr) db += (long long)blockIdx.z * g.Co;
  __shared__ __align__(16) float As[BK][BM + PADM];
  __shared__ __align__(16) float Bs[BK][BN + PADM];
  const int t = threadIdx.x;
  const int ty = t >> 4, tx = t & 15;
  const int A = g.Co, B = g.Ci;
  const int ntaps = g.KH * g.KW, NP = ntaps * B;
  const long long P = (long long)g.N * g.Ho * g.Wo;
  const long long p_begin = (long long)blockIdx.z * pixels_per_split;
  const long long p_end = min(P, p_begin + (long long)pixels_per_split);

  const int lp = t >> 4, seg = (t & 15) * 4;  // loader: pixel lp of the chunk, 4 consecutive columns at seg
  const int a_l = blockIdx.x * BM + seg;
  const int n_l = blockIdx.y * BN + seg;
  const bool a_vec = (A & 3) == 0 && a_l + 3 < A;
  const bool b_vec = (B & 3) == 0 && n_l + 3 < NP;
  int tapj[4], bj[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int np = n_l + j;
    tapj[j] = np < NP ? np / B : -1;
    bj[j] = np < NP ? np - tapj[j] * B : 0;
  }

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float colsum = 0.f;  // bias gradient partial for a = blockIdx.x*BM + t (t < 64)

  for (long long p0 = p_begin; p0 < p_end; p0 += BK) {
    const long long p = p0 + lp;
    float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
    if (p < p_end) {
      if (a_vec) {
        load_vec<T, 4>(gy + p * A + a_l, a);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (a_l + j < A) a[j] = to_f(gy[p * A + a_l + j]);
      }
      const int ow = (int)(p % g.Wo);
      const long long q = p / g.Wo;
      const int oh = (int)(q % g.Ho), n = (int)(q / g.Ho);
      if (b_vec) {
        const int kh = tapj[0] / g.KW, kw = tapj[0] - kh * g.KW;
        int ih, iw;
        if (tap_coord(g, oh, kh, g.pad_h, g.Hi, ih) && tap_coord(g, ow, kw, g.pad_w, g.Wi, iw))
          load_vec<T, 4>(gx + (((long long)n * g.Hi + ih) * g.Wi + iw) * B + bj[0], b);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (tapj[j] < 0) continue;
          const int kh = tapj[j] / g.KW, kw = tapj[j] - kh * g.KW;
          int ih, iw;
          if (tap_coord(g, oh, kh, g.pad_h, g.Hi, ih) && tap_coord(g, ow, kw, g.pad_w, g.Wi, iw))
            b[j] = to_f(gx[(((long long)n * g.Hi + ih) * g.Wi + iw) * B + bj[j]]);
        }
      }
    }
    __syncthreads();
    *reinterpret_cast<float4 *>(&As[lp][seg]) = make_float4(a[0], a[1], a[2], a[3]);
    *reinterpret_cast<float4 *>(&Bs[lp][seg]) = make_float4(b[0], b[1], b[2], b[3]);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 av = *reinterpret_cast<const float4 *>(&As[k][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4 *>(&Bs[k][tx * 4]);
      const float aa[4] = {av.x, av.y, av.z, av.w}, bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
    }
    if (db != nullptr && blockIdx.y == 0 && t < BM) {
#pragma unroll
      for (int k = 0; k < BK; ++k) colsum += As[k][t];
    }
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int a = blockIdx.x * BM + ty * 4 + i;
    if (a >= A) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int np = blockIdx.y * BN + tx * 4 + j;
      if (np >= NP) continue;
      const int tap = np / B, bb = np - tap * B;
      float *dst = packed ? dw + (long long)a * NP + np : dw + ((long long)a * B + bb) * ntaps + tap;      // [A][taps][B] or [A][B][taps]
      *dst = acc[i][j];
    }
  }
  if (db != nullptr && blockIdx.y == 0 && t < BM) {
    const int a = blockIdx.x * BM + t;
    if (a < A) db[a] = colsum;
  }
}

template <typename T>
__global__ void pack_weights_kernel(const float *__restrict__ w, T *__restrict__ wp, int O, int I, int KH, int KW, int mode) {
  const long long total = (long long)O * I * KH * KW;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    // idx enumerates the PACKED layout (coalesced writes)
    long long r = idx;
    if (mode == 0) {
      const int i = (int)(r % I); r /= I;
      const int kw = (int)(r % KW); r /= KW;
      const int kh = (int)(r % KH); r /= KH;
      const int o = (int)r;
      wp[idx] = from_f<T>(w[(((long long)o * I + i) * KH + kh) * KW + kw]);
    } else {
      const int o = (int)(r % O); r /= O;
      const int kw = (int)(r % KW); r /= KW;
      const int kh = (int)(r % KH); r /= KH;
      const int i = (int)r;
      wp[idx] = from_f<T>(w[(((long long)o * I + i) * KH + (KH - 1 - kh)) * KW + (KW - 1 - kw)]);
    }
  }
}

constexpr int PACK_MAX_ITEMS = 48;
constexpr int PACK_TILE = 32;          // 32 output channels x 32 input channels x all taps per block
struct PackTable {
  int n;
  int tile_start[PACK_MAX_ITEMS + 1];  // prefix sums of tile counts
  const float *w[PACK_MAX_ITEMS];
  void *wp[PACK_MAX_ITEMS];
  void *wp2[PACK_MAX_ITEMS];           // mode 2 (= both layouts from one read of the master weights): the mode-1 destination
  int O[PACK_MAX_ITEMS], I[PACK_MAX_ITEMS], KH[PACK_MAX_ITEMS], KW[PACK_MAX_ITEMS], mode[PACK_MAX_ITEMS];
};

// One launch re-packs every weight of a network (all layers x both layouts); the table rides in the kernel parameters.
// Each block moves a 32(O) x 32(I) x taps tile through shared memory: reads are contiguous runs of 32*taps fp32 of the OIHW master
// weights, writes are 32 consecutive elements of the packed layout (mode 0: [O][tap][I], mode 1: [I][flipped tap][O]).
template <typename T>
__global__ void __launch_bounds__(256) pack_weights_multi_kernel(const __grid_constant__ PackTable t) {
  extern __shared__ float tile[];        // [32 o][32 i * taps (+1 pad)]
  int lo = 0, hi = t.n - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (t.tile_start[mid] <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
  }
  const int it = lo;
  const int O = t.O[it], I = t.I[it], taps = t.KH[it] * t.KW[it];
  const int tiles_i = (I + PACK_TILE - 1) / PACK_TILE;
  const int local = blockIdx.x - t.tile_start[it];
  const int o0 = (local / tiles_i) * PACK_TILE, i0 = (local % tiles_i) * PACK_TILE;
  const int no = min(PACK_TILE, O - o0), ni = min(PACK_TILE, I - i0);
  const int row = ni * taps;             // contiguous fp32 per output channel in this tile
  const int pitch = PACK_TILE * taps + 1;
  const float *w = t.w[it];
  for (int idx = threadIdx.x; idx < no * row; idx += 256) {
    const int o = idx / row, r = idx - o * row;
    tile[o * pitch + r] = w[((long long)(o0 + o) * I + i0) * taps + r];
  }
  __syncthreads();
  const int mode = t.mode[it];
  if (mode == 0 || mode == 2) {
    // wp[o][tap][i]: consecutive threads -> consecutive i
    T *dst = reinterpret_cast<T *>(t.wp[it]);
    for (int idx = threadIdx.x; idx < no * taps * ni; idx += 256) {
      const int i = idx % ni;
      const int r = idx / ni;
      const int tap = r % taps, o = r / taps;
      dst[((long long)(o0 + o) * taps + tap) * I + i0 + i] = from_f<T>(tile[o * pitch + i * taps + tap]);
    }
  }
  if (mode == 1 || mode == 2) {
    // wp[i][taps-1-tap][o]: consecutive threads -> consecutive o
    T *dst = reinterpret_cast<T *>(mode == 2 ? t.wp2[it] : t.wp[it]);
    for (int idx = threadIdx.x; idx < ni * taps * no; idx += 256) {
      const int o = idx % no;
      const int r = idx / no;
      const int tap = r % taps, i = r / taps;
      dst[((long long)(i0 + i) * taps + (taps - 1 - tap)) * O + o0 + o] = from_f<T>(tile[o * pitch + i * taps + tap]);
    }
  }
}

}  // namespace

int ctagan_conv_gather_simt(const ctagan_conv_geom *g, const void *x, const void *wp, const float *bias, void *y, cudaStream_t st) {
  const long long M = (long long)g->N * g->Ho * g->Wo;
  dim3 grid(cdiv(M, BM), cdiv(g->Co, BN));
  const bool flat = (g->Ci % BK) != 0;
  CTAGAN_DISPATCH_DTYPE(g->dtype, T, {
    if (flat) conv_gather_simt_kernel<T, true><<<grid, 256, 0, st>>>(*g, (const T *)x, (const T *)wp, bias, (T *)y);
    else conv_gather_simt_kernel<T, false><<<grid, 256, 0, st>>>(*g, (const T *)x, (const T *)wp, bias, (T *)y);
  });
  CTAGAN_LAUNCH_OK();
  return CTAGAN_OK;
}

static int simt_wgrad_splits(const ctagan_conv_geom *g, int &pps) {
  const long long P = (long long)g->N * g->Ho * g->Wo;
  const int ntaps = g->KH * g->KW;
  const int gx_ = cdiv(g->Co, BM), gy_ = cdiv((long long)ntaps * g->Ci, BN);
  // split the pixel (K) dimension until the grid covers ~2 waves of the SMs
  const int target = 2 * ctagan_num_sms();
  int splits = (int)((target + (long long)gx_ * gy_ - 1) / ((long long)gx_ * gy_));
  const int max_splits = (int)((P + 255) / 256);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  pps = (int)((P + splits - 1) / splits);
  pps = ((pps + BK - 1) / BK) * BK;
  return (int)((P + pps - 1) / pps);
}

// scratch: split-K partial sums [splits][Co*Ci*taps] + [splits][Co] (nothing for a single split)
size_t ctagan_conv_wgrad_simt_workspace(const ctagan_conv_geom *g) {
  int pps;
  const int splits = simt_wgrad_splits(g, pps);
  return (size_t)splits * ((size_t)g->Co * g->Ci * g->KH * g->KW + g->Co) * sizeof(float);      // (a single split goes direct unless it accumulates)
}

int ctagan_conv_wgrad_simt(const ctagan_conv_geom *g, const void *gy, const void *gx, float *dw, float *db, void *workspace,
                           size_t workspace_bytes, cudaStream_t st, int accumulate) {
  const int ntaps = g->KH * g->KW;
  const int gx_ = cdiv(g->Co, BM), gy_ = cdiv((long long)ntaps * g->Ci, BN);
  int pps;
  const int splits = simt_wgrad_splits(g, pps);
  const long long dw_elems = (long long)g->Co * g->Ci * ntaps;
  float *dw_dst = dw, *db_dst = db;
  const int packed = (accumulate & CTAGAN_WGRAD_PACKED) ? 1 : 0;
  accumulate &= CTAGAN_WGRAD_ACCUMULATE;
  const bool staged = splits > 1 || accumulate;
  if (staged) {
    const size_t need = ctagan_conv_wgrad_simt_workspace(g);
    CTAGAN_REQUIRE(workspace && workspace_bytes >= need, "conv_wgrad(simt): workspace of %zu bytes required (got %zu)", need, workspace_bytes);
    dw_dst = (float *)workspace;
    db_dst = db ? dw_dst + (long long)splits * dw_elems : nullptr;
  }
  dim3 grid(gx_, gy_, splits);
  CTAGAN_DISPATCH_DTYPE(g->dtype, T, {
    conv_wgrad_simt_kernel<T><<<grid, 256, 0, st>>>(*g, (const T *)gy, (const T *)gx, dw_dst, db_dst, pps, packed);
  });
  CTAGAN_LAUNCH_OK();
  if (staged) {
    int rc = ctagan_ordered_sum(dw_dst, dw, splits, dw_elems, st, accumulate);
    if (rc) return rc;
    if (db) rc = ctagan_ordered_sum(db_dst, db, splits, g->Co, st, accumulate);
    return rc;
  }
  return CTAGAN_OK;
}

extern "C" int ctagan_pack_weights(const float *w, void *wp, int O, int I, int KH, int KW, int mode, int dtype, void *stream) {
  CTAGAN_REQUIRE(w && wp && O > 0 && I > 0 && KH > 0 && KW > 0 && (mode == 0 || mode == 1), "pack_weights: bad arguments");
  const long long total = (long long)O * I * KH * KW;
  const int blocks = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
  CTAGAN_DISPATCH_DTYPE(dtype, T, { pack_weights_kernel<T><<<blocks, 256, 0, (cudaStream_t)stream>>>(w, (T *)wp, O, I, KH, KW, mode); });
  CTAGAN_LAUNCH_OK();
  return CTAGAN_OK;
}

extern "C" int ctagan_pack_weights_multi(const ctagan_pack_item *items, int n_items, int dtype, void *stream) {
  CTAGAN_REQUIRE(items && n_items > 0, "pack_weights_multi: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  for (int k = 0; k < n_items; ++k) {
    const ctagan_pack_item &it = items[k];
    CTAGAN_REQUIRE(it.w && it.wp && it.O > 0 && it.I > 0 && it.KH > 0 && it.KW > 0 && (it.mode == 0 || it.mode == 1),
                   "pack_weights_multi: bad item %d", k);
  }
  // One launch per shared-memory class (taps <= 9, <= 16, larger): the tile buffer is sized by the largest filter of a launch, and a
  // single 7x7 layer in the list used to pin EVERY block to 200 KB (one block per SM).  A (mode 0, mode 1) pair of the same weight
  // becomes one entry that reads the master weights once and writes both layouts.
  const int class_taps[3] = {9, 16, 1 << 30};
  for (int cls = 0; cls < 3; ++cls) {
    PackTable t;
    t.n = 0;
    t.tile_start[0] = 0;
    int max_taps = 1;
    auto flush = [&]() -> int {
      if (t.n == 0) return CTAGAN_OK;
      const size_t smem = (size_t)PACK_TILE * (PACK_TILE * max_taps + 1) * sizeof(float);
      CTAGAN_REQUIRE(smem <= 200 * 1024, "pack_weights_multi: kernel window too large");
      CTAGAN_DISPATCH_DTYPE(dtype, T, {
        if (smem > 48 * 1024) CTAGAN_CUDA_OK(cudaFuncSetAttribute(pack_weights_multi_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        pack_weights_multi_kernel<T><<<t.tile_start[t.n], 256, smem, st>>>(t);
      });
      CTAGAN_LAUNCH_OK();
      t.n = 0;
      max_taps = 1;
      return CTAGAN_OK;
    };
    for (int k = 0; k < n_items; ++k) {
      const ctagan_pack_item &it = items[k];
      const int taps = it.KH * it.KW;
      if (taps > class_taps[cls] || (cls > 0 && taps <= class_taps[cls - 1])) continue;
      const bool pair = k + 1 < n_items && it.mode == 0 && items[k + 1].mode == 1 && items[k + 1].w == it.w && items[k + 1].O == it.O &&
                        items[k + 1].I == it.I && items[k + 1].KH == it.KH && items[k + 1].KW == it.KW;
      const int e = t.n++;
      t.w[e] = it.w; t.wp[e] = it.wp; t.wp2[e] = pair ? items[k + 1].wp : nullptr;
      t.O[e] = it.O; t.I[e] = it.I; t.KH[e] = it.KH; t.KW[e] = it.KW; t.mode[e] = pair ? 2 : it.mode;
      const int tiles = ((it.O + PACK_TILE - 1) / PACK_TILE) * ((it.I + PACK_TILE - 1) / PACK_TILE);
      t.tile_start[e + 1] = t.tile_start[e] + tiles;
      if (taps > max_taps) max_taps = taps;
      if (pair) ++k;
      if (t.n == PACK_MAX_ITEMS) { const int rc = flush(); if (rc) return rc; }
    }
    const int rc = flush();
    if (rc) return rc;
  }
  return CTAGAN_OK;
}
