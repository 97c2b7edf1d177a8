// f3: the optimiser step of a whole parameter group as ONE kernel.
//   torch.optim.Adam(lr, betas=(0.5, 0.999)) of trainer/CycTrainer.py:67-73, RegTrainer.py:97-101, HdTrainer.py:101-105, p2pTrainer.py:62-63
// per 32 x 32 x taps tile of a weight tensor: read the fp32 master weight, gradient and both moments once, apply the Adam update
// (the arithmetic of PyTorch's fused Adam, operation for operation: fp32, fma forms, powf bias corrections), write weight and moments
// back, and emit BOTH packed bf16 layouts the convolution engines read ([O][kh][kw][I] for fprop, flipped-transposed [I][kh][kw][O] for
// dgrad / ConvTranspose) from the tile while it is in shared memory.  28 B/param of optimiser traffic + 4 B/param of packed weights;
// the separate re-pack pass (8 B/param) and ~20 multi-tensor launches per group disappear.  The item table lives in device memory
// (built once: every pointer in it is persistent), so the launch is graph-capturable and one launch covers any number of tensors.
#include "common.cuh"

namespace {

constexpr int OPT_TILE = 32;          // 32 output channels x 32 input channels x all taps per block

__device__ __forceinline__ void adam_update(float &param, float grad, float &exp_avg, float &exp_avg_sq, float lr, float beta1, float beta2,
                                            float eps, float bias_correction1, float bias_correction2_sqrt) {
  // ATen/native/cuda/fused_adam_utils.cuh: adam_math<float, float, 4, ORIGINAL, false>, weight_decay = 0, no grad scaler
  exp_avg = fmaf(beta1, exp_avg, fmaf(-beta1, grad, grad));
  const float g2 = __fmul_rn(grad, grad);
  exp_avg_sq = fmaf(beta2, exp_avg_sq, fmaf(-beta2, g2, g2));
  const float step_size = __fdiv_rn(lr, bias_correction1);
  const float denom = __fadd_rn(__fdiv_rn(sqrtf(exp_avg_sq), bias_correction2_sqrt), eps);
  param = __fsub_rn(param, __fdiv_rn(__fmul_rn(step_size, exp_avg), denom));
}

template <typename T>
__global__ void __launch_bounds__(256, 5) adam_pack_kernel(const ctagan_adam_item *__restrict__ items, const int *__restrict__ tile_start, int n_items,
                                                        const float *__restrict__ lr_ptr, float *__restrict__ step_ptr, unsigned int *ticket,
                                                        float beta1, float beta2, float eps, int advance) {
  extern __shared__ float tile[];        // [no][ni * taps (+1 pad)]
  int lo = 0, hi = n_items - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (__ldg(tile_start + mid) <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
  }
  const ctagan_adam_item it = items[lo];
  const int O = it.O, I = it.I, taps = it.KH * it.KW;
  const int tiles_i = (I + OPT_TILE - 1) / OPT_TILE;
  const int local = blockIdx.x - __ldg(tile_start + lo);
  const int o0 = (local / tiles_i) * OPT_TILE, i0 = (local % tiles_i) * OPT_TILE;
  const int no = min(OPT_TILE, O - o0), ni = min(OPT_TILE, I - i0);
  const int row = ni * taps;             // contiguous fp32 per output channel in this tile
  const int pitch = row + 1;
  const float step = *step_ptr + 1.f;    // torch increments state['step'] before the update
  const float lr = *lr_ptr;
  const float bc1 = 1.f - powf(beta1, step);
  const float bc2_sqrt = sqrtf(1.f - powf(beta2, step));
  // The tile is no rows (output channels) of `row` = ni * taps contiguous master weights.  Threads walk it in row-major order, 256
  // consecutive elements per pass (coalesced, whole DRAM bursts), and keep (o, i, tap) of their element up to date incrementally: no
  // per-element integer division.  U passes are batched -- all 4 * U loads are issued before the first store -- because the compiler
  // must assume that the weight / moment stores alias the next pass's loads and would otherwise serialise the passes on DRAM latency.
  {
    constexpr int U = 2;
    const int di = 256 / taps, dt = 256 - di * taps;
    int o = (int)threadIdx.x / row, r0 = (int)threadIdx.x - o * row;
    int i = r0 / taps, tap = r0 - i * taps;
    const int row_stride = I * taps;                        // (a tensor has < 2^31 elements: 32-bit offsets)
    const int tile_base = (o0 * I + i0) * taps;             // + o * row_stride + (i * taps + tap)
    const int gpk_base = o0 * taps * I + i0;                // packed gradient [O][taps][I]: + (o * taps + tap) * I + i
    while (o < no) {
      float pv[U], mv[U], vv[U], gv[U];
      int gi[U], so[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int r = i * taps + tap;
        so[u] = o < no ? o * pitch + r : -1;
        gi[u] = tile_base + o * row_stride + r;
        if (o < no) {
          pv[u] = it.p[gi[u]]; mv[u] = it.m[gi[u]]; vv[u] = it.v[gi[u]];
          gv[u] = it.g[it.g_packed ? gpk_base + (o * taps + tap) * I + i : gi[u]];
        }
        i += di; tap += dt;                          // 256 elements further in row-major order
        if (tap >= taps) { tap -= taps; ++i; }
        while (i >= ni) { i -= ni; ++o; }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (so[u] < 0) continue;
        adam_update(pv[u], gv[u], mv[u], vv[u], lr, beta1, beta2, eps, bc1, bc2_sqrt);
        it.p[gi[u]] = pv[u]; it.m[gi[u]] = mv[u]; it.v[gi[u]] = vv[u];
        tile[so[u]] = pv[u];
      }
    }
  }
  if (it.wp0 != nullptr || it.wp1 != nullptr) {
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (it.wp0 != nullptr) {
      // wp[o][tap][i]: one warp per (o, tap) row, lanes along i (bank = (i * taps + tap) mod 32: conflict-free for odd taps)
      T *dst = reinterpret_cast<T *>(it.wp0);
      for (int q = warp; q < no * taps; q += 8) {
        const int o = q / taps, tap = q - o * taps;
        if (lane < ni) dst[((long long)(o0 + o) * taps + tap) * I + i0 + lane] = from_f<T>(tile[o * pitch + lane * taps + tap]);
      }
    }
    if (it.wp1 != nullptr) {
      // wp[i][taps-1-tap][o]: one warp per (i, tap) row, lanes along o (pitch is odd: conflict-free)
      T *dst = reinterpret_cast<T *>(it.wp1);
      for (int q = warp; q < ni * taps; q += 8) {
        const int i = q / taps, tap = q - i * taps;
        if (lane < no) dst[((long long)(i0 + i) * taps + (taps - 1 - tap)) * O + o0 + lane] = from_f<T>(tile[lane * pitch + i * taps + tap]);
      }
    }
  }
  // the step counter advances once every block has read it: the last block to finish (ticket) increments it and re-arms the ticket
  __shared__ bool last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    if (advance) *step_ptr = step;         // (an early launch over part of the group leaves the counter to the launch that completes the step)
    *ticket = 0u;
  }
}

}  // namespace

extern "C" size_t ctagan_adam_pack_smem_bytes(const ctagan_adam_item *items_host, int n_items) {
  size_t need = 0;
  for (int k = 0; k < n_items; ++k) {
    const ctagan_adam_item &it = items_host[k];
    const size_t no = it.O < OPT_TILE ? it.O : OPT_TILE, ni = it.I < OPT_TILE ? it.I : OPT_TILE;
    const size_t b = no * (ni * it.KH * it.KW + 1) * sizeof(float);
    if (b > need) need = b;
  }
  return need;
}

extern "C" int ctagan_adam_pack_tiles(const ctagan_adam_item *items_host, int n_items, int *tile_start_host) {
  CTAGAN_REQUIRE(items_host && tile_start_host && n_items > 0, "adam_pack_tiles: bad arguments");
  int t = 0;
  for (int k = 0; k < n_items; ++k) {
    const ctagan_adam_item &it = items_host[k];
    CTAGAN_REQUIRE(it.p && it.g && it.m && it.v && it.O > 0 && it.I > 0 && it.KH > 0 && it.KW > 0, "adam_pack_tiles: bad item %d", k);
    tile_start_host[k] = t;
    t += ((it.O + OPT_TILE - 1) / OPT_TILE) * ((it.I + OPT_TILE - 1) / OPT_TILE);
  }
  tile_start_host[n_items] = t;
  return CTAGAN_OK;
}

extern "C" int ctagan_adam_pack_multi(const ctagan_adam_item *items_dev, const int *tile_start_dev, int n_items, int total_tiles, size_t smem_bytes,
                                      const float *lr_dev, float *step_dev, uint32_t *ticket_dev, float beta1, float beta2, float eps,
                                      int packed_dtype, int advance_step, void *stream) {
  CTAGAN_REQUIRE(items_dev && tile_start_dev && n_items > 0 && total_tiles > 0 && lr_dev && step_dev && ticket_dev, "adam_pack_multi: bad arguments");
  CTAGAN_REQUIRE(smem_bytes <= 200 * 1024, "adam_pack_multi: filter window too large for one tile");
  cudaStream_t st = (cudaStream_t)stream;
  CTAGAN_DISPATCH_DTYPE(packed_dtype, T, {
    if (smem_bytes > 48 * 1024) CTAGAN_CUDA_OK(cudaFuncSetAttribute(adam_pack_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    adam_pack_kernel<T><<<total_tiles, 256, smem_bytes, st>>>(items_dev, tile_start_dev, n_items, lr_dev, step_dev, ticket_dev, beta1, beta2, eps, advance_step);
  });
  CTAGAN_LAUNCH_OK();
  return CTAGAN_OK;
}
