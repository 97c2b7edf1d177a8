// tcgen05 / TMEM / TMA implicit-GEMM convolution engine for sm_100a (bf16 operands, fp32 accumulation in tensor memory).
//
// Formulation ("pitch trick"): a stride-1 VALID convolution over a physically padded NHWC input [N][Hi][Wi][Ci] is
//     y_flat[q] = sum_{tap, ci} x_flat[q + kh*Wi + kw][ci] * wp[co][tap][ci],        q = (n*Hi + i)*Wi + j
// i.e. for every filter tap the A operand of an M=128 output-pixel tile is a CONTIGUOUS run of 128 rows of the flattened
// [pixels][Ci] input.  Each (tap, 64-channel chunk) is therefore one plain 2-D TMA box {64 ch, 128 px} landing in shared
// memory in exactly the 128B-swizzled K-major layout tcgen05.mma consumes -- no im2col buffer, no gather instructions.
// Outputs at virtual positions j >= Wo (the K-1 wrap-around columns per row) are computed and discarded: (Wi-K+1)/Wi useful.
// Reflection / zero padding is materialised by the producer kernel (ctagan_norm_act_pad), the input-gradient ("full"
// correlation) is the same kernel on a zero-padded dy with flipped+transposed weights.
//
// Roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread tcgen05.mma issuer,
// warps 2..5 = epilogue (tcgen05.ld -> bias/activation -> bf16 -> global).  STAGES-deep mbarrier ring between producer
// and MMA; tcgen05.commit releases stages and signals the epilogue.
#include <cuda.h>
#include <map>
#include <mutex>
#include <tuple>
#include "common.cuh"

namespace {

constexpr int TILE_M = 128;
constexpr int CHUNK_K = 64;     // bf16 elements per 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int MAX_TAPS = 49;

struct TcParams {
  int n_taps;
  int tap_pix_off[MAX_TAPS];   // row offset of the tap in the flattened input
  int tap_w_col[MAX_TAPS];     // column (element) offset of the tap in the packed weight row: tap * Ci
  int Ci, Co;
  int Hv, Wv;                  // virtual (input) grid per image
  int Hov, Wov;                // valid output extent in that grid
  int tiles_per_img;
  int out_H, out_W;            // output tensor spatial dims
  int act;
  const float *bias;
  bf16 *out;
};

// ---------------------------------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (reported as a launch failure) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(const CUtensorMap *map, uint64_t *bar, void *smem_dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "n"(COLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}

// D[tmem] (+)= A[smem] * B[smem]^T, bf16 x bf16 -> fp32, single-CTA, issued by one thread
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor, K-major operand, SWIZZLE_128B, 8-row groups 1024 B apart (cute::UMMA::SmemDescriptor):
//   [0,14) start>>4 | [16,30) LBO>>4 (=1, unused for swizzled K-major) | [32,46) SBO>>4 (=64) | [46,48) version=1 | [61,64) layout=2
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// Instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 (1) @4, a/b format BF16 (1) @7/@10, K-major both,
// N>>3 @17, M>>4 @24.
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

template <int BN>
struct TcConfig {
  static constexpr int A_BYTES = TILE_M * CHUNK_K * 2;   // 16 KB
  static constexpr int B_BYTES = BN * CHUNK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN >= 256) ? 4 : (BN >= 128 ? 5 : 6);
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
  static constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
};

template <int BN>
__global__ void __launch_bounds__(192, 1)
conv_tc_valid_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, const TcParams p) {
  using Cfg = TcConfig<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t *empty_bar = full_bar + Cfg::STAGES;
  uint64_t *tmem_full_bar = empty_bar + Cfg::STAGES;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int img = blockIdx.x / p.tiles_per_img;
  const int tile = blockIdx.x - img * p.tiles_per_img;
  const int q_local0 = tile * TILE_M;                         // first virtual position of this tile inside the image
  const long long q0 = (long long)img * p.Hv * p.Wv + q_local0;
  const int co0 = blockIdx.y * BN;
  const int k_chunks = p.Ci / CHUNK_K;
  const int n_iters = p.n_taps * k_chunks;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_x);
    tma_prefetch_desc(&map_w);
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int it = 0; it < n_iters; ++it) {
        const int s = it % Cfg::STAGES;
        const uint32_t ph = (uint32_t)(it / Cfg::STAGES) & 1u;
        mbar_wait(&empty_bar[s], ph ^ 1u);
        const int tap = it / k_chunks, kc = it - tap * k_chunks;
        uint8_t *a_dst = smem + s * Cfg::STAGE_BYTES;
        uint8_t *b_dst = a_dst + Cfg::A_BYTES;
        mbar_expect_tx(&full_bar[s], Cfg::STAGE_BYTES);
        tma_load_2d(&map_x, &full_bar[s], a_dst, kc * CHUNK_K, (int)(q0 + p.tap_pix_off[tap]));
        tma_load_2d(&map_w, &full_bar[s], b_dst, p.tap_w_col[tap] + kc * CHUNK_K, co0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(TILE_M, BN);
      for (int it = 0; it < n_iters; ++it) {
        const int s = it % Cfg::STAGES;
        const uint32_t ph = (uint32_t)(it / Cfg::STAGES) & 1u;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + s * Cfg::STAGE_BYTES);
        const uint32_t b_addr = a_addr + Cfg::A_BYTES;
        const uint64_t adesc = make_kmajor_sw128_desc(a_addr);
        const uint64_t bdesc = make_kmajor_sw128_desc(b_addr);
#pragma unroll
        for (int k = 0; k < CHUNK_K / UMMA_K; ++k) {
          // advance 16 bf16 = 32 bytes along K inside the 128-byte swizzle row: +2 in the (addr >> 4) field
          umma_bf16(tmem_base, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (it > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);   // frees the smem stage when these MMAs retire (implies fence::before_thread_sync)
      }
      umma_commit(tmem_full_bar);     // accumulator complete
    }
  } else {
    // ===== epilogue: warps 2..5; warp w may touch TMEM lanes 32*(w%4) .. +31 =====
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int ql = q_local0 + row;
    const int i = ql / p.Wv, j = ql - i * p.Wv;
    const bool valid = (i < p.Hov) && (j < p.Wov);
    bf16 *out_row = p.out + (((long long)img * p.out_H + i) * p.out_W + j) * p.Co + co0;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
#pragma unroll 1
    for (int c = 0; c < BN; c += 32) {
      uint32_t r[32];
      tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c, r);
      tmem_ld_wait();
      if (valid) {
        float v[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) v[e] = __uint_as_float(r[e]);
        if (p.bias) {
#pragma unroll
          for (int e = 0; e < 32; ++e) v[e] += __ldg(p.bias + co0 + c + e);
        }
        if (p.act != CTAGAN_ACT_NONE) {
#pragma unroll
          for (int e = 0; e < 32; ++e) v[e] = apply_act(v[e], p.act);
        }
        uint4 *dst = reinterpret_cast<uint4 *>(out_row + c);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 pk;
          __nv_bfloat162 h0 = __floats2bfloat162_rn(v[8 * g + 0], v[8 * g + 1]);
          __nv_bfloat162 h1 = __floats2bfloat162_rn(v[8 * g + 2], v[8 * g + 3]);
          __nv_bfloat162 h2 = __floats2bfloat162_rn(v[8 * g + 4], v[8 * g + 5]);
          __nv_bfloat162 h3 = __floats2bfloat162_rn(v[8 * g + 6], v[8 * g + 7]);
          pk.x = *reinterpret_cast<uint32_t *>(&h0);
          pk.y = *reinterpret_cast<uint32_t *>(&h1);
          pk.z = *reinterpret_cast<uint32_t *>(&h2);
          pk.w = *reinterpret_cast<uint32_t *>(&h3);
          dst[g] = pk;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// host side: tensor maps
// ---------------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void *ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  });
  return fn;
}

// 2-D bf16 row-major [rows][cols] tensor, box {64 cols, box_rows}, 128B swizzle, zero OOB fill
int make_map_2d(CUtensorMap *map, const void *base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    ctagan_set_error("cuTensorMapEncodeTiled unavailable from the driver");
    return CTAGAN_ERR_CUDA;
  }
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 2};
  cuuint32_t box[2] = {CHUNK_K, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    ctagan_set_error("cuTensorMapEncodeTiled failed (%d): rows=%llu cols=%llu box_rows=%u", (int)r, (unsigned long long)rows,
                     (unsigned long long)cols, box_rows);
    return CTAGAN_ERR_CUDA;
  }
  return CTAGAN_OK;
}

template <int BN>
int launch_tc(const CUtensorMap &mx, const CUtensorMap &mw, const TcParams &p, dim3 grid, cudaStream_t st) {
  using Cfg = TcConfig<BN>;
  static bool configured = false;
  if (!configured) {
    CTAGAN_CUDA_OK(cudaFuncSetAttribute(conv_tc_valid_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    configured = true;
  }
  conv_tc_valid_kernel<BN><<<grid, 192, Cfg::SMEM_BYTES, st>>>(mx, mw, p);
  CTAGAN_LAUNCH_OK();
  return CTAGAN_OK;
}

int pick_bn(const ctagan_conv_geom *g) {
  const long long m_tiles = (long long)g->N * (((long long)g->Ho * g->Wi + TILE_M - 1) / TILE_M);
  const int sms = ctagan_num_sms();
  // largest N tile that still gives every SM a CTA; N=64 tiles are shared-memory-bandwidth bound (A re-read per 32 cycles)
  for (int bn : {256, 128, 64}) {
    if (g->Co % bn) continue;
    if (m_tiles * (g->Co / bn) >= sms || bn == 64) return bn;
  }
  return 0;
}

}  // namespace

int ctagan_conv_gather_tc_eligible(const ctagan_conv_geom *g) {
  if (g->dtype != CTAGAN_BF16) return 0;
  if (g->stride != 1 || g->dil != 1 || g->pad_h != 0 || g->pad_w != 0) return 0;
  if (g->Ci % CHUNK_K) return 0;
  if (g->Co % 64) return 0;
  if (g->KH * g->KW > MAX_TAPS) return 0;
  if (g->Ho != g->Hi - g->KH + 1 || g->Wo != g->Wi - g->KW + 1) return 0;
  if ((long long)g->N * g->Ho * g->Wo < 1024) return 0;   // tiny maps: launch-latency bound either way, CUDA-core kernel
  return 1;
}

int ctagan_conv_gather_tc(const ctagan_conv_geom *g, const void *x, const void *wp, const float *bias, void *y, cudaStream_t st) {
  if (!ctagan_conv_gather_tc_eligible(g)) {
    ctagan_set_error("conv_gather: geometry not supported by the tcgen05 engine (needs bf16, stride 1, pad 0 on a padded input, Ci%%64==0, Co%%64==0)");
    return CTAGAN_ERR_UNSUPPORTED;
  }
  CTAGAN_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(wp) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(y) & 15) == 0,
                 "conv_gather(tc): pointers must be 16-byte aligned");
  TcParams p;
  memset(&p, 0, sizeof(p));
  p.n_taps = g->KH * g->KW;
  for (int kh = 0; kh < g->KH; ++kh)
    for (int kw = 0; kw < g->KW; ++kw) {
      p.tap_pix_off[kh * g->KW + kw] = kh * g->Wi + kw;
      p.tap_w_col[kh * g->KW + kw] = (kh * g->KW + kw) * g->Ci;
    }
  p.Ci = g->Ci; p.Co = g->Co;
  p.Hv = g->Hi; p.Wv = g->Wi; p.Hov = g->Ho; p.Wov = g->Wo;
  p.tiles_per_img = (int)(((long long)g->Ho * g->Wi + TILE_M - 1) / TILE_M);
  p.out_H = g->Ho; p.out_W = g->Wo;
  p.act = g->act; p.bias = bias; p.out = (bf16 *)y;
  CUtensorMap mx, mw;
  int rc = make_map_2d(&mx, x, (uint64_t)g->N * g->Hi * g->Wi, (uint64_t)g->Ci, TILE_M);
  if (rc) return rc;
  const int bn = pick_bn(g);
  rc = make_map_2d(&mw, wp, (uint64_t)g->Co, (uint64_t)p.n_taps * g->Ci, (uint32_t)bn);
  if (rc) return rc;
  dim3 grid((unsigned)(g->N * p.tiles_per_img), (unsigned)(g->Co / bn));
  switch (bn) {
    case 256: return launch_tc<256>(mx, mw, p, grid, st);
    case 128: return launch_tc<128>(mx, mw, p, grid, st);
    case 64: return launch_tc<64>(mx, mw, p, grid, st);
  }
  ctagan_set_error("conv_gather(tc): no tile configuration");
  return CTAGAN_ERR_UNSUPPORTED;
}

int ctagan_conv_wgrad_tc_eligible(const ctagan_conv_geom *g) { (void)g; return 0; }
int ctagan_conv_wgrad_tc(const ctagan_conv_geom *g, const void *gy, const void *gx, float *dw, float *db, cudaStream_t st) {
  (void)g; (void)gy; (void)gx; (void)dw; (void)db; (void)st;
  ctagan_set_error("conv_wgrad: geometry not supported by the tcgen05 engine");
  return CTAGAN_ERR_UNSUPPORTED;
}
