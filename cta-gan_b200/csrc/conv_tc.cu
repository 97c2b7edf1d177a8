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
#include <cooperative_groups.h>
#include <stdlib.h>
#include <map>
#include <mutex>
#include <tuple>
#include "common.cuh"

namespace cg = cooperative_groups;

// Developer probes (per-phase clock64 accounting of one CTA, printed by the host with CTAGAN_TC_PROF=1 / CTAGAN_WG_PROF=1) are compiled in
// only with -DCTAGAN_PROBES=1 (CTAGAN_PROBES=1 python cta-gan_b200/build.py --force): even a clock read in the MMA issue loop costs ~2 %.
#ifndef CTAGAN_PROBES
#define CTAGAN_PROBES 0
#endif
__device__ __forceinline__ long long probe_clock() { return CTAGAN_PROBES ? clock64() : 0LL; }

namespace {

constexpr int TILE_M = 128;
constexpr int CHUNK_K = 64;     // bf16 elements per 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int MAX_TAPS = 49;

struct TcParams {
  int mode;                    // 0: pitch trick (2-D map, stride-1 VALID conv on a padded input); 1: 4-D boxes (any stride / zero pad)
  int n_taps;
  short tap_dh[MAX_TAPS];      // input offset of the tap relative to (i*stride, j*stride)
  short tap_dw[MAX_TAPS];
  int tap_w_col[MAX_TAPS];     // tap index in the packed weight [O][tap][Ci] (3-D map: channels past Ci are zero-filled)
  int Ci, Co;
  int stride;                  // input step per output position (mode 1)
  int Hv, Wv;                  // mode 0: virtual (input) grid per image
  int bw_log2;                 // mode 1: tile = (128 >> bw_log2) rows x (1 << bw_log2) cols of output positions
  int tiles_w;                 // mode 1: tiles along the width
  int Hov, Wov;                // valid extent of the output-position grid (i, j)
  int tiles_per_img;
  int out_H, out_W;            // output tensor spatial dims; position (i,j) is stored at (i*sy+ay, j*sx+ax)
  int sy, sx, ay, ax;
  int act;
  int n_stages;                // pipeline depth actually used (<= TcConfig::STAGES)
  int dbg;                     // tuning/debug knobs (0 in production)
  long long *prof;             // optional per-phase clock64 stamps of CTA 0 (developer probe)
  long long dbg_t0;
  const float *bias;
  bf16 *out;
  float *stat_part;            // optional [N][stat_parts][Co][2]: per-tile (sum, sum of squares) of the fp32 conv output over the tile's rows
  unsigned int *stat_ticket;   // [N][stat_tpi], zero on entry (and again on exit): arrivals per (image, column tile)
  float *stat_out;             // [N][Co][2] (mean, rstd), written by the last CTA of an (image, column tile)
  int stat_hw;                 // pixels per (n,co) plane
  int stat_parts;              // partial-sum slots per image (over all launches that feed these statistics)
  int stat_part0;              // first slot of THIS launch (output-phase launches: phase index * tiles per image)
  int stat_tpi;                // tickets per image (>= column tiles)
  int imgs_per_group;          // grouped launch: image n uses the weights (and bias) at row offset w_row_off[n / imgs_per_group]
  int w_row_off[CTAGAN_MAX_GROUPS];
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
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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

// One lane of a fully converged warp (elect.sync): unlike `if (lane == 0)` the compiler knows that exactly one lane runs the
// guarded code, so the TMA / tcgen05 instructions (uniform-datapath operands) are emitted straight instead of inside a
// per-lane "waterfall" loop (ELECT + BRA.U.ANY around every UTCHMMA / UTMALDG).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n"
      ".reg .b32 rx;\n"
      ".reg .pred px;\n"
      "elect.sync rx|px, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, px;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

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

__device__ __forceinline__ void tma_load_3d(const CUtensorMap *map, uint64_t *bar, void *smem_dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

__device__ __forceinline__ void tma_load_4d(const CUtensorMap *map, uint64_t *bar, void *smem_dst, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// MN-major operand (the contiguous dimension is M or N, K runs across 128-byte rows), SWIZZLE_128B:
// canonical layout ((64 elems, m slabs),(8 rows, k groups)) : ((1, LBO),(128 B, SBO)); LBO = byte distance between 64-element slabs,
// SBO = 1024 B between 8-row groups.
__device__ __forceinline__ uint64_t make_mnmajor_sw128_desc(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc_bf16_mn(int M, int N) {
  return make_idesc_bf16(M, N) | (1u << 15) | (1u << 16);
}

struct WgParams {
  int ntaps, KW;
  int Co, Ci;
  int stride, pad;
  int margin;                 // gy zero border that is skipped
  int rb, cb;                 // chunk grid per image (row blocks x col blocks)
  int bkh, bkw;               // chunk = bkh x bkw output pixels (= 64)
  int total_chunks, chunks_per_split;   // chunks of ONE group (grouped launch: blockIdx.z = group, its images follow each other)
  int ci_tiles;
  int splits;                 // K splits of a tile == CTAs of its thread-block cluster (1, 2, 4 or 8)
  float *dw;                  // [groups][Co][Ci][KH][KW] fp32 (PyTorch OIHW) or, dw_packed, [groups][Co][KH][KW][Ci]: written (or added to)
  int dw_packed;
  float *ws;                  // split-K workspace [groups][tiles][splits][128][BNW] fp32 (splits > 1)
  long long *prof;            // developer probe (CTAGAN_WG_PROF=1): clock64 phase accounting of CTA (0,0,0), else nullptr
};

constexpr int WG_M = 128;       // Co tile
constexpr int WG_KPIX = 64;     // pixels per K chunk
constexpr int WG_SLAB = WG_KPIX * 128;   // bytes of one [64 px][64 ch] slab

template <int BNW>
struct WgConfig {
  static constexpr int A_BYTES = (WG_M / 64) * WG_SLAB;      // 16 KB
  static constexpr int B_BYTES = (BNW / 64) * WG_SLAB;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = BNW >= 256 ? 4 : 5;
  static constexpr int ACC_PITCH = BNW + 4;                  // fp32 words per row of the accumulator dump (16-byte rows, conflict-free float4 stores)
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
  static constexpr int TMEM_COLS = BNW < 32 ? 32 : BNW;
  static_assert(WG_M * ACC_PITCH * 4 <= STAGES * STAGE_BYTES, "the accumulator dump reuses the operand ring");
};

// dW[co][ci][tap] = sum_pix gy[pix][co] * gx[pix + tap][ci]   (both operands MN-major).
// One CTA = one (tap, Co tile, Ci tile) and one K split (a range of 64-pixel chunks); the K splits of a tile form ONE THREAD-BLOCK
// CLUSTER (grid.x = cluster size = splits).  Every CTA dumps its fp32 accumulator (TMEM) into its own shared memory -- the operand ring
// is free by then -- and after a cluster barrier CTA r adds rows [r*128/S, (r+1)*128/S) of all S dumps through distributed shared
// memory, IN RANK ORDER (deterministic), and stores the result straight into the OIHW gradient (optionally adding to what is
// there: the second use of a network in the same iteration).  No split-K workspace in global memory, no reduce launch.
template <int BNW, bool ACC>
__global__ void __launch_bounds__(192, 1)
conv_wgrad_tc_kernel(const __grid_constant__ CUtensorMap map_gy, const __grid_constant__ CUtensorMap map_gx, const WgParams p) {
  using Cfg = WgConfig<BNW>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t *empty_bar = full_bar + Cfg::STAGES;
  uint64_t *tmem_full_bar = empty_bar + Cfg::STAGES;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int t = blockIdx.y;
  const int ci_t = t % p.ci_tiles; t /= p.ci_tiles;
  const int co_tiles = (p.Co + WG_M - 1) / WG_M;
  const int co_t = t % co_tiles; t /= co_tiles;
  const int tap = t;
  const int kh = tap / p.KW, kw = tap - kh * p.KW;
  const int split = blockIdx.x, grp = blockIdx.z;          // split == rank of this CTA in its cluster
  const int chunk_lo = split * p.chunks_per_split;
  const int n_iters = max(0, min(p.total_chunks, chunk_lo + p.chunks_per_split) - chunk_lo);
  const int chunk0 = grp * p.total_chunks + chunk_lo;          // global chunk index: (image, row block, column block)
  const int co0 = co_t * WG_M, ci0 = ci_t * BNW;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_gy);
    tma_prefetch_desc(&map_gx);
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
  pdl_wait();

  __shared__ long long prof_s[10];                            // developer probe: phase clocks of this CTA
  if (warp == 0) {
    // TMA producer: converged warp, one elected lane issues (see elect_one)
    long long t_wait = 0, t_all = probe_clock();
    for (int it = 0; it < n_iters; ++it) {
      const int s = it % Cfg::STAGES;
      const uint32_t ph = (uint32_t)(it / Cfg::STAGES) & 1u;
      const long long tw = probe_clock();
      mbar_wait(&empty_bar[s], ph ^ 1u);
      t_wait += probe_clock() - tw;
      int ch = chunk0 + it;
      const int cbi = ch % p.cb; ch /= p.cb;
      const int rbi = ch % p.rb;
      const int n = ch / p.rb;
      const int oh0 = p.margin + rbi * p.bkh, ow0 = p.margin + cbi * p.bkw;     // gy coordinates of the chunk origin
      const int iw0 = ow0 * p.stride + kw - p.pad, ih0 = oh0 * p.stride + kh - p.pad;
      if (elect_one()) {
        uint8_t *a_dst = smem + s * Cfg::STAGE_BYTES;
        uint8_t *b_dst = a_dst + Cfg::A_BYTES;
        mbar_expect_tx(&full_bar[s], Cfg::STAGE_BYTES);
#pragma unroll
        for (int sl = 0; sl < WG_M / 64; ++sl) tma_load_4d(&map_gy, &full_bar[s], a_dst + sl * WG_SLAB, co0 + sl * 64, ow0, oh0, n);
#pragma unroll
        for (int sl = 0; sl < BNW / 64; ++sl) tma_load_4d(&map_gx, &full_bar[s], b_dst + sl * WG_SLAB, ci0 + sl * 64, iw0, ih0, n);
      }
      __syncwarp();
    }
    if (CTAGAN_PROBES && lane == 0) { prof_s[0] = probe_clock() - t_all; prof_s[1] = t_wait; prof_s[8] = n_iters; }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_bf16_mn(WG_M, BNW);
    long long t_wait = 0, t_all = probe_clock();
    for (int it = 0; it < n_iters; ++it) {
      const int s = it % Cfg::STAGES;
      const uint32_t ph = (uint32_t)(it / Cfg::STAGES) & 1u;
      const long long tw = probe_clock();
      mbar_wait(&full_bar[s], ph);
      t_wait += probe_clock() - tw;
      tc_fence_after();
      if (elect_one()) {
        const uint32_t a_addr = smem_u32(smem + s * Cfg::STAGE_BYTES);
        const uint32_t b_addr = a_addr + Cfg::A_BYTES;
#pragma unroll
        for (int k = 0; k < WG_KPIX / UMMA_K; ++k) {
          // 16 pixels (K) = 16 rows of 128 B = 2048 B further into every slab
          const uint64_t adesc = make_mnmajor_sw128_desc(a_addr + k * 2048, WG_SLAB);
          const uint64_t bdesc = make_mnmajor_sw128_desc(b_addr + k * 2048, WG_SLAB);
          umma_bf16(tmem_base, adesc, bdesc, idesc, (it > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);
        if (it == n_iters - 1) umma_commit(tmem_full_bar);
      }
      __syncwarp();
    }
    if (CTAGAN_PROBES && lane == 0) { prof_s[2] = probe_clock() - t_all; prof_s[3] = t_wait; }
  } else {
    // epilogue warps: wait for the accumulator (the commit behind the last MMA also means every operand read of the ring is done)
    const long long t_e0 = probe_clock();
    if (n_iters > 0) {
      mbar_wait(tmem_full_bar, 0);
      tc_fence_after();
    }
    if (CTAGAN_PROBES && threadIdx.x == 64) prof_s[4] = probe_clock() - t_e0;
  }
  // ---- split-K reduction.  The K splits of a tile are the CTAs of one cluster, so they are co-resident and a cluster barrier orders them
  // (no grid-wide synchronisation, no second launch).  CTA r OWNS rows [r*128/S, (r+1)*128/S) of the tile:
  //   1. the epilogue warps dump the accumulator (TMEM lane = co row) into this CTA's shared memory (the operand ring is free: the
  //      commit behind the last MMA means every operand read is done), row pitch BNW+4 words: conflict-free 16-byte stores;
  //   2. all warps copy the rows OTHER CTAs own to the split-K workspace in global memory (L2) with fully coalesced 16-byte stores;
  //   3. cluster barrier (release / acquire at cluster scope);
  //   4. every CTA adds the S partial sums of its own rows IN RANK ORDER (deterministic; its own from shared memory, the others from
  //      L2) and writes the gradient: coalesced when the destination is the packed [co][tap][ci] layout, a 4-byte scatter for OIHW.
  const int S = p.splits;
  const int rows_per = (WG_M + S - 1) / S;                    // (S need not divide 128: the last CTA then owns fewer rows)
  const int my_row0 = split * rows_per;
  const int my_rows = max(0, min(WG_M, my_row0 + rows_per) - my_row0);
  constexpr int VPR = BNW / 4;                                // float4 per row
  float *acc = reinterpret_cast<float *>(smem);               // [128][ACC_PITCH] fp32, over the ring
  const long long t_s0 = probe_clock();
  if (warp >= 2) {
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    float4 *dst = reinterpret_cast<float4 *>(acc + row * Cfg::ACC_PITCH);
#pragma unroll 1
    for (int c = 0; c < BNW; c += 32) {
      uint32_t r[32];
      if (n_iters > 0) {
        tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c, r);
        tmem_ld_wait();
      } else {                                    // a split without chunks (ragged division) contributes zeros
#pragma unroll
        for (int e = 0; e < 32; ++e) r[e] = 0u;
      }
#pragma unroll
      for (int g = 0; g < 8; ++g)
        dst[c / 4 + g] = make_float4(__uint_as_float(r[4 * g]), __uint_as_float(r[4 * g + 1]), __uint_as_float(r[4 * g + 2]), __uint_as_float(r[4 * g + 3]));
    }
  }
  tc_fence_before();
  __syncthreads();
  const long long t_s1 = probe_clock();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
  // workspace: [group][tile][split][128][BNW] fp32
  float *ws_tile = p.ws + ((long long)grp * gridDim.y + blockIdx.y) * S * (WG_M * BNW);
  if (S > 1) {
    // rows other CTAs own: (S-1) * rows_per rows of VPR float4 each, skipping this CTA's own block of rows; four independent
    // 16-byte copies per thread and iteration
    float4 *mine = reinterpret_cast<float4 *>(ws_tile + (long long)split * (WG_M * BNW));
    const int n_vec = (WG_M - my_rows) * VPR;
    for (int v0 = threadIdx.x; v0 < n_vec; v0 += 4 * 192) {
      float4 t[4];
      int dst_v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int v = v0 + u * 192;
        int row = v / VPR;
        const int c4 = v - row * VPR;
        if (row >= my_row0) row += my_rows;
        dst_v[u] = row * VPR + c4;
        if (v < n_vec) t[u] = *reinterpret_cast<const float4 *>(acc + row * Cfg::ACC_PITCH + c4 * 4);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (v0 + u * 192 < n_vec) mine[dst_v[u]] = t[u];
    }
    __threadfence();
  }
  const long long t_s2 = probe_clock();
  cg::cluster_group cluster = cg::this_cluster();
  if (S > 1) cluster.sync();
  const long long t_s3 = probe_clock();
  {
    float *dw = p.dw + (long long)grp * p.Co * p.Ci * p.ntaps;
    constexpr int UV = 2;                          // rows-of-4 per thread and iteration: UV * (S-1) independent L2 loads in flight
    for (int v0 = threadIdx.x; v0 < my_rows * VPR; v0 += UV * 192) {
      float4 u[UV][8];
      int rowv[UV], cv[UV];
      bool ok[UV];
#pragma unroll
      for (int q = 0; q < UV; ++q) {
        const int v = v0 + q * 192;
        const int lr = v / VPR;
        rowv[q] = my_row0 + lr;
        cv[q] = (v - lr * VPR) * 4;
        ok[q] = v < my_rows * VPR && co0 + rowv[q] < p.Co && ci0 + cv[q] < p.Ci;        // Ci is a multiple of 32 here
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (k < S && ok[q])
            u[q][k] = k == split ? *reinterpret_cast<const float4 *>(acc + rowv[q] * Cfg::ACC_PITCH + cv[q])
                                 : __ldcg(reinterpret_cast<const float4 *>(ws_tile + (long long)k * (WG_M * BNW) + rowv[q] * BNW + cv[q]));
      }
#pragma unroll
      for (int q = 0; q < UV; ++q) {
        if (!ok[q]) continue;
        float4 sum = u[q][0];
#pragma unroll
        for (int k = 1; k < 8; ++k)
          if (k < S) { sum.x += u[q][k].x; sum.y += u[q][k].y; sum.z += u[q][k].z; sum.w += u[q][k].w; }      // rank order: deterministic
        const int row = rowv[q], c = cv[q];
        if (p.dw_packed) {                          // [co][tap][ci]: ci contiguous
          float4 *d = reinterpret_cast<float4 *>(dw + ((long long)(co0 + row) * p.ntaps + tap) * p.Ci + ci0 + c);
          if (ACC) { const float4 o = *d; sum.x += o.x; sum.y += o.y; sum.z += o.z; sum.w += o.w; }
          *d = sum;
        } else {                                    // PyTorch OIHW: tap is the fastest index
          float *d = dw + ((long long)(co0 + row) * p.Ci + ci0 + c) * p.ntaps + tap;
          if (ACC) {                                // (a template parameter: the plain store must not carry a speculative load of the strided destination)
            sum.x += d[0]; sum.y += d[p.ntaps]; sum.z += d[2 * p.ntaps]; sum.w += d[3 * p.ntaps];
          }
          d[0] = sum.x; d[p.ntaps] = sum.y; d[2 * p.ntaps] = sum.z; d[3 * p.ntaps] = sum.w;
        }
      }
    }
  }
  if (CTAGAN_PROBES && p.prof && blockIdx.x + blockIdx.y + blockIdx.z == 0) {
    __syncthreads();
    if (threadIdx.x == 64) {
      prof_s[5] = t_s1 - t_s0; prof_s[6] = t_s2 - t_s1; prof_s[7] = t_s3 - t_s2; prof_s[9] = probe_clock() - t_s3;
      for (int k = 0; k < 10; ++k) p.prof[k] = prof_s[k];
    }
  }
}

// db[c] = sum over pixels of gy[p][c]   (bias gradient, only when requested).  grid.x = pixel chunks; block = 8 pixel lanes x 32
// channel pairs (C <= 64 per pass).  Every block stores its partial sums in its own row of part[blocks][C]; ctagan_ordered_sum adds
// the rows in block order (deterministic: no floating-point atomics).
__global__ void __launch_bounds__(256) colsum_kernel(const bf16 *__restrict__ gy, float *__restrict__ part, long long pixels, int C,
                                                     long long pix_per_block) {
  const int cp = threadIdx.x & 31, pl = threadIdx.x >> 5;      // channel pair, pixel lane
  const long long p0 = (long long)blockIdx.x * pix_per_block;
  const long long p1 = min(pixels, p0 + pix_per_block);
  __shared__ float sm[8][64];
  for (int c0 = 0; c0 < C; c0 += 64) {
    const int c = c0 + cp * 2;
    float s0 = 0.f, s1 = 0.f;
    if (c < C)
      for (long long p = p0 + pl; p < p1; p += 8) {
        const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162 *>(gy + p * C + c);
        s0 += __bfloat162float(v.x);
        s1 += __bfloat162float(v.y);
      }
    __syncthreads();
    sm[pl][cp * 2] = s0;
    sm[pl][cp * 2 + 1] = s1;
    __syncthreads();
    if (threadIdx.x < 64 && c0 + threadIdx.x < C) {
      float t = 0.f;
#pragma unroll
      for (int l = 0; l < 8; ++l) t += sm[l][threadIdx.x];
      part[(long long)blockIdx.x * C + c0 + threadIdx.x] = t;
    }
  }
}

template <int BN, int KCH>
struct TcConfig {
  static constexpr int A_BYTES = TILE_M * CHUNK_K * 2;   // 16 KB per chunk
  static constexpr int B_BYTES = BN * CHUNK_K * 2;
  static constexpr int STAGE_BYTES = KCH * (A_BYTES + B_BYTES);
  static constexpr int STAGES = (196 * 1024) / STAGE_BYTES > 6 ? 6 : (196 * 1024) / STAGE_BYTES;     // (+ ~25 KB of static epilogue buffers)
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
  static constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
};

// Epilogue of one 128 x BN accumulator tile (4 warps, warp w reads TMEM lanes 32*(w%4) .. +31): bias / activation / bf16 stores and,
// optionally, the InstanceNorm statistics of the layer.
//
// Statistics are DETERMINISTIC: every CTA stores its per-column (sum, sum of squares) over its 128 rows into its own slot of
// stat_part[N][parts][Co][2] (plain stores, fixed arithmetic order inside the tile), and the last CTA of an (image, column tile) to
// arrive -- a ticket per (image, column tile) -- adds the slots in slot order in fp64 and writes (mean, rstd).  No floating-point
// atomics anywhere: the same inputs give the same bits whatever the CTA schedule.
//
// Column sums inside the tile: a thread holds 32 columns of ONE row, the sums run over rows.  Each warp transposes its 32 x 32 block
// through a private, bank-conflict-free shared-memory tile (row pitch 33 words: 32 STS + 32 LDS per thread and chunk, against 62
// shuffles + 124 selects of a register butterfly), lane l then owns column l of the chunk; the per-warp results wait in shared memory
// and the four warps are combined ONCE per tile (one barrier per tile, not one per 32-column chunk).
constexpr int EPI_PITCH = 33;

template <int BN>
struct EpiSmem {
  float xpose[4][32 * EPI_PITCH];     // per-warp transpose tile
  float part[4][BN][2];               // per-warp column sums of the tile: [warp][column][sum | sumsq]
  unsigned int ticket;
};

template <int BN>
__device__ __forceinline__ void tc_epilogue(const TcParams &p, uint32_t tmem_acc, int warp, int lane, int img, int tile, int co0, int wrow0,
                                            int co_tile, EpiSmem<BN> &es, uint64_t *acc_drained = nullptr, int et_base = 64,
                                            bool commit_stats = true) {
  const int quarter = warp & 3;
  const int row = quarter * 32 + lane;
  const int BW = 1 << p.bw_log2;
  int i, j;
  if (p.mode == 0) {
    const int ql = tile * TILE_M + row;
    i = ql / p.Wv; j = ql - i * p.Wv;
  } else {
    i = (tile / p.tiles_w) * (TILE_M >> p.bw_log2) + (row >> p.bw_log2);
    j = (tile % p.tiles_w) * BW + (row & (BW - 1));
  }
  const int oi = i * p.sy + p.ay, oj = j * p.sx + p.ax;
  const bool valid = (i < p.Hov) && (j < p.Wov) && oi >= 0 && oi < p.out_H && oj >= 0 && oj < p.out_W;
  bf16 *out_row = p.out + (((long long)img * p.out_H + oi) * p.out_W + oj) * p.Co + co0;
  const bool stats = p.stat_part != nullptr;
  float *xp = es.xpose[quarter];
#pragma unroll 1
  for (int c = 0; c < BN; c += 32) {
    uint32_t r[32];
    tmem_ld32(tmem_acc + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c, r);
    tmem_ld_wait();
    if (stats) {
      __syncwarp();                                  // the previous chunk's column reads are done
#pragma unroll
      for (int e = 0; e < 32; ++e) xp[lane * EPI_PITCH + e] = valid ? __uint_as_float(r[e]) : 0.f;
      __syncwarp();
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int k = 0; k < 32; ++k) {                 // rows of the warp in order: fixed summation order
        const float t = xp[k * EPI_PITCH + lane];
        s1 += t;
        s2 = fmaf(t, t, s2);
      }
      es.part[quarter][c + lane][0] = s1;
      es.part[quarter][c + lane][1] = s2;
    }
    if (valid && co0 + c < p.Co) {            // Co is a multiple of 32 here; tiles may overhang it
      float v[32];
#pragma unroll
      for (int e = 0; e < 32; ++e) v[e] = __uint_as_float(r[e]);
      if (p.bias) {
#pragma unroll
        for (int e = 0; e < 32; ++e) v[e] += __ldg(p.bias + wrow0 + c + e);
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
  if (acc_drained != nullptr) {          // the accumulator is in registers / stored: hand it back to the MMA issuer (persistent callers)
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(acc_drained);
  }
  if (!stats || !commit_stats) return;      // (!commit_stats: the caller accumulates es.part over its tiles and publishes one slot per image)
  // combine the four warps (in warp order) and store this tile's slot; then "last CTA finalises": once every CTA of this (image,
  // column tile) has stored its slot, the last one adds the slots in slot order and turns them into (mean, rstd); it also re-arms
  // the ticket (the buffer is zero again when the kernel ends)
  const int et = (int)threadIdx.x - et_base;      // 0..127
  asm volatile("bar.sync 1, 128;" ::: "memory");
  float2 *slot = reinterpret_cast<float2 *>(p.stat_part) + ((long long)img * p.stat_parts + p.stat_part0 + tile) * p.Co + co0;
#pragma unroll
  for (int col = et; col < BN; col += 128) {
    if (co0 + col < p.Co) {
      const float a = (es.part[0][col][0] + es.part[1][col][0]) + (es.part[2][col][0] + es.part[3][col][0]);
      const float b = (es.part[0][col][1] + es.part[1][col][1]) + (es.part[2][col][1] + es.part[3][col][1]);
      slot[col] = make_float2(a, b);
    }
  }
  __threadfence();
  asm volatile("bar.sync 1, 128;" ::: "memory");
  unsigned int *ticket = p.stat_ticket + (long long)img * p.stat_tpi + co_tile;
  if (et == 0) es.ticket = atomicAdd(ticket, 1u);
  asm volatile("bar.sync 1, 128;" ::: "memory");
  if (es.ticket == (unsigned int)p.stat_parts - 1u) {
    __threadfence();
    const double inv = 1.0 / (double)p.stat_hw;
    const float2 *base = reinterpret_cast<const float2 *>(p.stat_part) + (long long)img * p.stat_parts * p.Co + co0;
    constexpr int NC = (BN + 127) / 128;         // columns per thread
    double s1[NC], s2[NC];
#pragma unroll
    for (int q = 0; q < NC; ++q) s1[q] = s2[q] = 0.0;
    // many independent loads in flight (the slots are L2 lines written by other SMs); the additions stay in slot order
    for (int k0 = 0; k0 < p.stat_parts; k0 += 16) {
      float2 v[NC][16];
#pragma unroll
      for (int q = 0; q < NC; ++q)
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          const int col = et + q * 128;
          v[q][u] = (k0 + u < p.stat_parts && col < BN && co0 + col < p.Co) ? __ldcg(base + (long long)(k0 + u) * p.Co + col) : make_float2(0.f, 0.f);
        }
#pragma unroll
      for (int q = 0; q < NC; ++q)
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          s1[q] += (double)v[q][u].x;
          s2[q] += (double)v[q][u].y;
        }
    }
#pragma unroll
    for (int q = 0; q < NC; ++q) {
      const int col = et + q * 128;
      if (col < BN && co0 + col < p.Co) {
        const double m = s1[q] * inv;
        double var = s2[q] * inv - m * m;
        if (var < 0) var = 0;
        *reinterpret_cast<float2 *>(p.stat_out + ((long long)img * p.Co + co0 + col) * 2) = make_float2((float)m, (float)(1.0 / sqrt(var + 1e-5)));
      }
    }
    if (et == 0) *ticket = 0u;
  }
}

// KS: 16-channel MMA K steps issued per 64-channel chunk.  4 = the whole chunk; 2 for layers with Ci <= 32, whose chunk is half TMA zero
// fill: these layers are bound by the ~70-cycle issue cost of an MMA (not by its size), so skipping the empty half halves their main
// loop.  A template parameter, not a run-time bound: the instruction stream between two tcgen05.mma issues is on the critical path
// (the MMA queue is shallow), and a run-time test in that loop measurably slows every layer down.
template <int BN, int KCH, int KS = CHUNK_K / UMMA_K>
__global__ void __launch_bounds__(192, 1)
conv_tc_valid_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, const TcParams p) {
  using Cfg = TcConfig<BN, KCH>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + p.n_stages * Cfg::STAGE_BYTES);   // (the ring may be shorter than Cfg::STAGES)
  uint64_t *empty_bar = full_bar + Cfg::STAGES;
  uint64_t *tmem_full_bar = empty_bar + Cfg::STAGES;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_full_bar + 1);
  __shared__ EpiSmem<BN> epi;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int img = blockIdx.x / p.tiles_per_img;
  const int tile = blockIdx.x - img * p.tiles_per_img;
  const int q_local0 = tile * TILE_M;                         // mode 0: first virtual position of this tile inside the image
  const long long q0 = (long long)img * p.Hv * p.Wv + q_local0;
  const int BW = 1 << p.bw_log2;
  const int tile_i0 = (tile / p.tiles_w) * (TILE_M >> p.bw_log2);   // mode 1: tile origin in output positions
  const int tile_j0 = (tile % p.tiles_w) * BW;
  const int co0 = blockIdx.y * BN;
  const int wrow0 = p.w_row_off[img / p.imgs_per_group] + co0;       // first row of this tile in the (grouped) packed-weight buffer
  const int groups = (p.Ci + CHUNK_K * KCH - 1) / (CHUNK_K * KCH);   // stages per tap (channel tail = TMA zero fill)
  const int NS = p.n_stages;

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
  pdl_wait();                          // producer grid complete + flushed: nothing above touches global memory
  __shared__ long long prof_s[12];     // developer probe (CTAGAN_TC_PROF=1): phase clocks of CTA (0,0)
  const long long t_start = probe_clock();

  if (warp == 0) {
    // ===== TMA producer: the whole warp walks the ring (coordinates stay warp-uniform), one elected lane issues =====
    int s = 0;
    uint32_t ph = 0;
    long long t_wait = 0;
    for (int tap = 0; tap < p.n_taps; ++tap) {
      const int dh = p.tap_dh[tap], dw = p.tap_dw[tap], wcol = p.tap_w_col[tap];
      const int row2d = (int)(q0 + dh * p.Wv + dw);
      const int c1 = tile_j0 * p.stride + dw, c2 = tile_i0 * p.stride + dh;
      for (int gk = 0; gk < groups; ++gk) {
        const long long tw = probe_clock();
        mbar_wait(&empty_bar[s], ph ^ 1u);
        t_wait += probe_clock() - tw;
        if (elect_one()) {
          uint8_t *a_dst = smem + s * Cfg::STAGE_BYTES;
          uint8_t *b_dst = a_dst + KCH * Cfg::A_BYTES;
          mbar_expect_tx(&full_bar[s], Cfg::STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < KCH; ++k) {
            const int ch = (gk * KCH + k) * CHUNK_K;
            if (p.mode == 0) tma_load_2d(&map_x, &full_bar[s], a_dst + k * Cfg::A_BYTES, ch, row2d);
            else tma_load_4d(&map_x, &full_bar[s], a_dst + k * Cfg::A_BYTES, ch, c1, c2, img);
            tma_load_3d(&map_w, &full_bar[s], b_dst + k * Cfg::B_BYTES, ch, wcol, wrow0);
          }
        }
        __syncwarp();
        if (++s == NS) { s = 0; ph ^= 1u; }
      }
    }
    if (CTAGAN_PROBES && lane == 0) { prof_s[0] = probe_clock() - t_start; prof_s[1] = t_wait; }
  } else if (warp == 1) {
    // ===== MMA issuer: converged warp, one elected lane issues the tcgen05.mma / commit =====
    constexpr uint32_t idesc = make_idesc_bf16(TILE_M, BN);
    const int n_iters = p.n_taps * groups;
    int s = 0;
    uint32_t ph = 0;
    long long t_wait = 0, t_first = 0;
    for (int it = 0; it < n_iters; ++it) {
      const long long tw = probe_clock();
      mbar_wait(&full_bar[s], ph);
      if (it == 0) t_first = probe_clock() - tw; else t_wait += probe_clock() - tw;
      tc_fence_after();
      if (elect_one()) {
        const uint32_t a_addr = smem_u32(smem + s * Cfg::STAGE_BYTES);
        const uint32_t b_addr = a_addr + KCH * Cfg::A_BYTES;
#pragma unroll
        for (int k = 0; k < KCH; ++k) {
          const uint64_t adesc = make_kmajor_sw128_desc(a_addr + k * Cfg::A_BYTES);
          const uint64_t bdesc = make_kmajor_sw128_desc(b_addr + k * Cfg::B_BYTES);
#pragma unroll
          for (int kk = 0; kk < KS; ++kk) {
            // advance 16 bf16 = 32 bytes along K inside the 128-byte swizzle row: +2 in the (addr >> 4) field
            umma_bf16(tmem_base, adesc + (uint64_t)(2 * kk), bdesc + (uint64_t)(2 * kk), idesc, (it > 0 || k > 0 || kk > 0) ? 1u : 0u);
          }
        }
        umma_commit(&empty_bar[s]);           // frees the smem stage when these MMAs retire (implies fence::before_thread_sync)
        if (it == n_iters - 1) umma_commit(tmem_full_bar);     // accumulator complete
      }
      __syncwarp();
      if (++s == NS) { s = 0; ph ^= 1u; }
    }
    if (CTAGAN_PROBES && lane == 0) { prof_s[2] = probe_clock() - t_start; prof_s[3] = t_wait; prof_s[4] = t_first; prof_s[5] = n_iters; }
  } else {
    // ===== epilogue: warps 2..5 =====
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    if (CTAGAN_PROBES && threadIdx.x == 64) prof_s[6] = probe_clock() - t_start;
    // accumulator complete: from here on only the epilogue is left, so let the consumer's CTAs be scheduled now (they run their
    // prologue and block in their own pdl_wait() until this grid has finished).  Triggering at kernel start instead made the
    // step SLOWER: early-resident consumers held shared memory that the other streams' kernels needed.
    pdl_trigger();
    tc_epilogue<BN>(p, tmem_base, warp, lane, img, tile, co0, wrow0, (int)blockIdx.y, epi);
    if (CTAGAN_PROBES && threadIdx.x == 64) prof_s[7] = probe_clock() - t_start;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
  if (CTAGAN_PROBES && p.prof && blockIdx.x + blockIdx.y == 0 && threadIdx.x == 64) {
    for (int k = 0; k < 8; ++k) p.prof[k] = prof_s[k];
    p.prof[8] = t_start - p.dbg_t0;
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

// packed weights [O][taps][Ci] as a 3-D map, box {64 ch, 1 tap, box_rows}: channel / row overhang is zero-filled by TMA
int make_map_w3d(CUtensorMap *map, const void *base, int O, int taps, int Ci, uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    ctagan_set_error("cuTensorMapEncodeTiled unavailable from the driver");
    return CTAGAN_ERR_CUDA;
  }
  cuuint64_t dims[3] = {(cuuint64_t)Ci, (cuuint64_t)taps, (cuuint64_t)O};
  cuuint64_t strides[2] = {(cuuint64_t)Ci * 2, (cuuint64_t)taps * Ci * 2};
  cuuint32_t box[3] = {CHUNK_K, 1, box_rows};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void *>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    ctagan_set_error("cuTensorMapEncodeTiled(w3d) failed (%d): O=%d taps=%d Ci=%d box_rows=%u", (int)r, O, taps, Ci, box_rows);
    return CTAGAN_ERR_CUDA;
  }
  return CTAGAN_OK;
}

template <int BN, int KCH, int KS = CHUNK_K / UMMA_K>
int launch_tc(const CUtensorMap &mx, const CUtensorMap &mw, const TcParams &p, dim3 grid, cudaStream_t st) {
  using Cfg = TcConfig<BN, KCH>;
  static bool configured = false;
  if (!configured) {
    CTAGAN_CUDA_OK(cudaFuncSetAttribute(conv_tc_valid_kernel<BN, KCH, KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    configured = true;
  }
  TcParams q = p;
  q.n_stages = Cfg::STAGES;
  // Short-K tiles (few taps x few channel chunks, e.g. the 32-channel 3x3 layers of Reg: 9 stages per tile) on a grid of many waves are
  // bound by the per-CTA prologue / epilogue, not by the ring depth: a ring that fits twice into an SM's shared memory lets two CTAs
  // share the SM, so one CTA's main loop runs under the other's epilogue.
  const int iters = p.n_taps * ((p.Ci + CHUNK_K * KCH - 1) / (CHUNK_K * KCH));
  static int short_k = -1;
  if (short_k < 0) { const char *e = getenv("CTAGAN_TC_SHORTK"); short_k = e ? atoi(e) : 20; }    // measured on the Reg step (b=8): off 18.05 ms, 12 -> 17.31 ms, 20 -> 17.21 ms
  if (iters <= short_k && (long long)grid.x * grid.y >= 4LL * ctagan_num_sms()) {
    const int fit = (int)((113 * 1024 - sizeof(EpiSmem<BN>) - 2048) / Cfg::STAGE_BYTES);
    if (fit >= 2 && fit < q.n_stages) q.n_stages = fit;
  }
  if (const char *env = getenv("CTAGAN_TC_STAGES")) {
    const int ns = atoi(env);
    if (ns >= 1 && ns <= Cfg::STAGES) q.n_stages = ns;
  }
  const size_t smem_bytes = (size_t)q.n_stages * Cfg::STAGE_BYTES + 1024 + 256;
  static long long *prof_buf = nullptr;
  static int prof_on = -1;
  if (prof_on < 0) { const char *e = getenv("CTAGAN_TC_PROF"); prof_on = e ? atoi(e) : 0; }
  if (prof_on) {
    if (!prof_buf) cudaMalloc(&prof_buf, 16 * sizeof(long long));
    q.prof = prof_buf;
  }
  CTAGAN_CUDA_OK(launch_pdl(conv_tc_valid_kernel<BN, KCH, KS>, grid, dim3(192), smem_bytes, st, mx, mw, q));
  CTAGAN_LAUNCH_OK();
  if (prof_on) {
    long long h[16];
    cudaStreamSynchronize(st);
    cudaMemcpy(h, prof_buf, sizeof(h), cudaMemcpyDeviceToHost);
    fprintf(stderr, "[conv prof] BN=%d KCH=%d stages=%d grid=(%u,%u) Ci=%d Co=%d taps=%d stats=%d iters=%lld | producer done %lld (wait_empty %lld) | mma done %lld "
            "(first full %lld, later waits %lld) | acc ready %lld | epilogue done %lld (clk from kernel start)\n", BN, KCH, q.n_stages, grid.x, grid.y, p.Ci, p.Co,
            p.n_taps, p.stat_part ? 1 : 0, h[5], h[0], h[1], h[2], h[4], h[3], h[6], h[7]);
  }
  return CTAGAN_OK;
}

int pick_bn(long long m_tiles, int Co) {
  const int sms = ctagan_num_sms();
  if (const char *env = getenv("CTAGAN_TC_BN")) {          // tuning override
    const int bn = atoi(env);
    if ((bn == 64 || bn == 128 || bn == 256) && Co % bn == 0) return bn;
  }
  // Operand traffic from L2 per FLOP falls with the tile size (each SM sustains ~60 B/clk of TMA traffic: a 128x64 tile is
  // bandwidth-bound at ~32% of the tensor peak, 128x128 at ~48%, 128x256 at ~64%), while small grids leave SMs idle.  Measured on B200
  // (3x3 256->256, 64x64 maps): N=256 wins once it fills the chip; N=128 is best from ~1/3 of the SMs up (and lets two independent
  // generator chains of a batch-1 step share the chip); N=64 only for the smallest grids.
  if (Co % 256 == 0 && m_tiles * (Co / 256) >= sms) return 256;
  if (Co % 128 == 0 && m_tiles * (Co / 128) >= sms / 3) return 128;
  return 64;                                 // also Co = 32 / 96 ...: the tile overhang is zero-filled and masked
  return 0;
}

int make_map_4d(CUtensorMap *map, const void *base, int N, int H, int W, int C, int bw, int bh, int estr_hw);

// launch one tcgen05 conv with the tap table already in p; x described by (mode 0) [rows][Ci] or (mode 1) [N][Hi][Wi][Ci]
int run_tc(TcParams &p, const void *x, int N, int Hi, int Wi, const void *wp, int w_taps, cudaStream_t st, int w_slots = 1) {
  CUtensorMap mx, mw;
  int rc;
  if (p.imgs_per_group <= 0) { p.imgs_per_group = N; for (int k = 0; k < CTAGAN_MAX_GROUPS; ++k) p.w_row_off[k] = 0; }
  if (p.mode == 0) {
    rc = make_map_2d(&mx, x, (uint64_t)N * Hi * Wi, (uint64_t)p.Ci, TILE_M);
  } else {
    const int BW = 1 << p.bw_log2, BH = TILE_M >> p.bw_log2;
    rc = make_map_4d(&mx, x, N, Hi, Wi, p.Ci, BW * p.stride, BH * p.stride, p.stride);
  }
  if (rc) return rc;
  const int bn = pick_bn((long long)N * p.tiles_per_img, p.Co);
  rc = make_map_w3d(&mw, wp, p.Co * w_slots, w_taps, p.Ci, (uint32_t)bn);
  if (rc) return rc;
  dim3 grid((unsigned)(N * p.tiles_per_img), (unsigned)((p.Co + bn - 1) / bn));
  CTAGAN_REQUIRE(p.stat_part == nullptr || (int)grid.y <= p.stat_tpi, "conv_gather(tc): ticket buffer too small");
  const int kch = (p.Ci % 128 == 0 && bn < 256) ? 2 : 1;      // BN=256 keeps 64-channel stages (4 of them fit; 2-chunk stages would leave 2)
  if (kch == 2) {
    switch (bn) {
      case 128: return launch_tc<128, 2>(mx, mw, p, grid, st);
      case 64: return launch_tc<64, 2>(mx, mw, p, grid, st);
    }
  } else {
    switch (bn) {
      case 256: return launch_tc<256, 1>(mx, mw, p, grid, st);
      case 128: return launch_tc<128, 1>(mx, mw, p, grid, st);
      case 64: return p.Ci <= 32 ? launch_tc<64, 1, 2>(mx, mw, p, grid, st) : launch_tc<64, 1>(mx, mw, p, grid, st);
    }
  }
  ctagan_set_error("conv_gather(tc): no tile configuration");
  return CTAGAN_ERR_UNSUPPORTED;
}

int ceil_log2(int v) {
  int l = 0;
  while ((1 << l) < v) ++l;
  return l;
}

}  // namespace

// Which tcgen05 formulation serves this geometry: 0 none, 1 pitch trick, 2 strided 4-D boxes, 3 output-phase decomposition
static int tc_gather_kind(const ctagan_conv_geom *g) {
  if (g->dtype != CTAGAN_BF16) return 0;
  if (g->Ci % 8 || g->Ci < 32 || g->Co % 32) return 0;       // 16-byte pixel rows; channel tails are TMA zero fill + masked stores
  if (g->KH * g->KW > MAX_TAPS) return 0;
  // tiny maps are launch-latency bound either way, but the generic CUDA-core kernel needs ~65 us for a 64-channel 3x3 layer on
  // 8 x 4 x 4 pixels (one thread per output, 576 dependent FMAs) where one masked tensor-core tile needs ~8 us
  static int min_px = 0, min_wo = 0;
  if (!min_px) { const char *e = getenv("CTAGAN_TC_MIN_PIXELS"); min_px = e ? atoi(e) : 32; if (min_px < 1) min_px = 1; }
  if (!min_wo) { const char *e = getenv("CTAGAN_TC_MIN_WO"); min_wo = e ? atoi(e) : 2; if (min_wo < 1) min_wo = 1; }
  if ((long long)g->N * g->Ho * g->Wo < min_px) return 0;
  if (g->dil == 1) {
    if (g->stride == 1 && g->pad_h == 0 && g->pad_w == 0 && g->Ho == g->Hi - g->KH + 1 && g->Wo == g->Wi - g->KW + 1) return 1;
    if (g->stride <= 2 && g->Wo >= min_wo) return 2;
    return 0;
  }
  if (g->dil == 2 && g->stride == 1 && g->Wo >= 32 && (g->Ho % 2 == 0) && (g->Wo % 2 == 0)) return 3;
  return 0;
}

int ctagan_conv_gather_tc_eligible(const ctagan_conv_geom *g) { return tc_gather_kind(g) != 0; }

// tiles per image of one launch, and the number of launches, for the statistics slots
static void tc_tile_counts(const ctagan_conv_geom *g, int kind, int &tiles_per_img, int &launches) {
  launches = 1;
  if (kind == 1) {
    tiles_per_img = (int)(((long long)g->Ho * g->Wi + TILE_M - 1) / TILE_M);
  } else if (kind == 2) {
    const int bwl = ceil_log2(g->Wo < 128 ? g->Wo : 128);
    const int BW = 1 << bwl, BH = TILE_M >> bwl;
    tiles_per_img = ((g->Wo + BW - 1) / BW) * ((g->Ho + BH - 1) / BH);
  } else {
    const int Hq = g->Ho / 2, Wq = g->Wo / 2;
    const int bwl = ceil_log2(Wq < 128 ? Wq : 128);
    const int BWq = 1 << bwl, BHq = TILE_M >> bwl;
    tiles_per_img = ((Wq + BWq - 1) / BWq) * ((Hq + BHq - 1) / BHq);
    launches = 0;
    for (int rh = 0; rh < 2; ++rh)
      for (int rw = 0; rw < 2; ++rw) {
        int t = 0;
        for (int kh = (rh + g->pad_h) & 1; kh < g->KH; kh += 2)
          for (int kw = (rw + g->pad_w) & 1; kw < g->KW; kw += 2) ++t;
        if (t) ++launches;
      }
  }
}

// scratch of the fused statistics: partial-sum slots [N][parts][Co][2] fp32 (0 when the geometry is not served by this engine)
size_t ctagan_conv_gather_tc_stat_bytes(const ctagan_conv_geom *g) {
  const int kind = tc_gather_kind(g);
  if (!kind) return 0;
  int tiles, launches;
  tc_tile_counts(g, kind, tiles, launches);
  return (size_t)g->N * tiles * launches * g->Co * 2 * sizeof(float);
}

int ctagan_conv_gather_tc(const ctagan_conv_geom *g, const void *x, const void *wp, const float *bias, void *y, unsigned int *stat_ticket,
                          void *stat_scratch, size_t stat_scratch_bytes, float *stat_out, cudaStream_t st, const ctagan_conv_groups *gr) {
  const int kind = tc_gather_kind(g);
  if (!kind) {
    ctagan_set_error("conv_gather: geometry not supported by the tcgen05 engine (bf16, Ci%%8==0, Ci>=32, Co%%32==0, stride<=2 / dil<=2)");
    return CTAGAN_ERR_UNSUPPORTED;
  }
  CTAGAN_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(wp) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(y) & 15) == 0,
                 "conv_gather(tc): pointers must be 16-byte aligned");
  TcParams p;
  memset(&p, 0, sizeof(p));
  p.Ci = g->Ci; p.Co = g->Co; p.stride = g->stride;
  p.out_H = g->Ho; p.out_W = g->Wo; p.sy = p.sx = 1; p.ay = p.ax = 0;
  p.act = g->act; p.bias = bias; p.out = (bf16 *)y;
  if (stat_out) {
    int tiles, launches;
    tc_tile_counts(g, kind, tiles, launches);
    CTAGAN_REQUIRE(stat_ticket && stat_scratch && stat_scratch_bytes >= ctagan_conv_gather_tc_stat_bytes(g),
                   "conv_gather_stats: ticket buffer and %zu bytes of scratch required", ctagan_conv_gather_tc_stat_bytes(g));
    CTAGAN_REQUIRE((reinterpret_cast<uintptr_t>(stat_scratch) & 7) == 0 && (reinterpret_cast<uintptr_t>(stat_out) & 7) == 0,
                   "conv_gather_stats: scratch / stats must be 8-byte aligned");
    p.stat_part = (float *)stat_scratch; p.stat_ticket = stat_ticket; p.stat_out = stat_out;
    p.stat_hw = g->Ho * g->Wo; p.stat_parts = tiles * launches; p.stat_part0 = 0;
    p.stat_tpi = (g->Co + 31) / 32;
  }
  int w_slots = 1;
  if (gr) {
    CTAGAN_REQUIRE(gr->groups >= 1 && gr->groups <= CTAGAN_MAX_GROUPS && g->N % gr->groups == 0, "conv_gather(grouped): N must split evenly into 1..%d groups", CTAGAN_MAX_GROUPS);
    p.imgs_per_group = g->N / gr->groups;
    for (int k = 0; k < gr->groups; ++k) {
      CTAGAN_REQUIRE(gr->slot[k] >= 0 && gr->slot[k] < CTAGAN_MAX_GROUPS, "conv_gather(grouped): bad weight slot");
      p.w_row_off[k] = gr->slot[k] * g->Co;
      if (gr->slot[k] + 1 > w_slots) w_slots = gr->slot[k] + 1;
    }
  }
  const int ntaps = g->KH * g->KW;
  const int w_taps = ntaps;
  if (kind == 1 || kind == 2) {
    p.n_taps = ntaps;
    for (int kh = 0; kh < g->KH; ++kh)
      for (int kw = 0; kw < g->KW; ++kw) {
        const int t = kh * g->KW + kw;
        p.tap_dh[t] = (short)(kh - g->pad_h);
        p.tap_dw[t] = (short)(kw - g->pad_w);
        p.tap_w_col[t] = t;
      }
    p.Hov = g->Ho; p.Wov = g->Wo;
    if (kind == 1) {
      p.mode = 0; p.Hv = g->Hi; p.Wv = g->Wi;
      p.tiles_per_img = (int)(((long long)g->Ho * g->Wi + TILE_M - 1) / TILE_M);
    } else {
      p.mode = 1;
      p.bw_log2 = ceil_log2(g->Wo < 128 ? g->Wo : 128);
      const int BW = 1 << p.bw_log2, BH = TILE_M >> p.bw_log2;
      p.tiles_w = (g->Wo + BW - 1) / BW;
      p.tiles_per_img = p.tiles_w * ((g->Ho + BH - 1) / BH);
    }
    return run_tc(p, x, g->N, g->Hi, g->Wi, wp, w_taps, st, w_slots);
  }
  // kind 3: input dilation 2 (input gradient of a stride-2 conv == ConvTranspose2d forward).  Output pixel h = 2i + r reads
  // x[i + e - u] with weight tap kh' = 2u + a (a = (r + pad) & 1, e = (r + pad - a) / 2): one stride-1 launch per output parity.
  int phase = 0;       // statistics: every output-parity launch owns its own slots, the ticket counts the CTAs of all of them
  for (int rh = 0; rh < 2; ++rh)
    for (int rw = 0; rw < 2; ++rw) {
      const int ah = (rh + g->pad_h) & 1, eh = (rh + g->pad_h - ah) / 2;
      const int aw = (rw + g->pad_w) & 1, ew = (rw + g->pad_w - aw) / 2;
      int t = 0;
      for (int kh = ah; kh < g->KH; kh += 2)
        for (int kw = aw; kw < g->KW; kw += 2) {
          // gather form: y[h] += x[(h + kh - pad)/2] * wp[kh]  ->  with h = 2i + rh:  x[i + (rh + kh - pad)/2]
          p.tap_dh[t] = (short)((rh + kh - g->pad_h) / 2);
          p.tap_dw[t] = (short)((rw + kw - g->pad_w) / 2);
          p.tap_w_col[t] = kh * g->KW + kw;
          ++t;
        }
      (void)eh; (void)ew;
      if (t == 0) continue;
      p.n_taps = t;
      p.mode = 1; p.stride = 1;
      p.Hov = g->Ho / 2; p.Wov = g->Wo / 2;
      p.sy = p.sx = 2; p.ay = rh; p.ax = rw;
      p.bw_log2 = ceil_log2(p.Wov < 128 ? p.Wov : 128);
      const int BW = 1 << p.bw_log2, BH = TILE_M >> p.bw_log2;
      p.tiles_w = (p.Wov + BW - 1) / BW;
      p.tiles_per_img = p.tiles_w * ((p.Hov + BH - 1) / BH);
      p.stat_part0 = phase * p.tiles_per_img;
      ++phase;
      int rc = run_tc(p, x, g->N, g->Hi, g->Wi, wp, w_taps, st, w_slots);
      if (rc) return rc;
    }
  return CTAGAN_OK;
}

namespace {

// 4-D bf16 NHWC tensor [N][H][W][C]; box {64 ch, bw, bh, 1}; traversal stride `estr` along W and H (strided convolutions)
int make_map_4d(CUtensorMap *map, const void *base, int N, int H, int W, int C, int bw, int bh, int estr_hw) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    ctagan_set_error("cuTensorMapEncodeTiled unavailable from the driver");
    return CTAGAN_ERR_CUDA;
  }
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)bw, (cuuint32_t)bh, 1};
  cuuint32_t estr[4] = {1, (cuuint32_t)estr_hw, (cuuint32_t)estr_hw, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void *>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    ctagan_set_error("cuTensorMapEncodeTiled(4d) failed (%d): N=%d H=%d W=%d C=%d box=%dx%d estr=%d", (int)r, N, H, W, C, bh, bw, estr_hw);
    return CTAGAN_ERR_CUDA;
  }
  return CTAGAN_OK;
}

struct WgPlan {
  int bnw, bkw, bkh, rb, cb, total_chunks, splits, cps, tiles;
};

bool plan_wgrad(const ctagan_conv_geom *g, WgPlan &pl, int n_groups = 1) {
  if (g->dtype != CTAGAN_BF16 || g->dil != 1) return false;
  if (g->stride > 2) return false;
  if (g->Co % 32 || g->Ci % 32) return false;              // tiles overhang with TMA zero fill, stores are masked
  if (g->KH * g->KW > MAX_TAPS) return false;
  const int m = g->gy_margin;
  const int Hvld = g->Ho - 2 * m, Wvld = g->Wo - 2 * m;
  if (Hvld <= 0 || Wvld <= 0) return false;
  static int min_px = 0;
  if (!min_px) { const char *e = getenv("CTAGAN_WG_MIN_PIXELS"); min_px = e ? atoi(e) : 512; if (min_px < 1) min_px = 1; }
  if ((long long)g->N * Hvld * Wvld < min_px) return false;   // tiny maps stay on the CUDA-core kernel
  pl.bnw = g->Ci > 128 ? 256 : (g->Ci > 64 ? 128 : 64);
  pl.bkw = Wvld >= 64 ? 64 : (Wvld > 16 ? 32 : 16);
  pl.bkh = 64 / pl.bkw;
  pl.rb = (Hvld + pl.bkh - 1) / pl.bkh;
  pl.cb = (Wvld + pl.bkw - 1) / pl.bkw;
  pl.total_chunks = (g->N / n_groups) * pl.rb * pl.cb;        // per group
  pl.tiles = g->KH * g->KW * ((g->Co + WG_M - 1) / WG_M) * ((g->Ci + pl.bnw - 1) / pl.bnw);
  // K splits = CTAs of the tile's cluster (1, 2, 4 or 8): enough of them to put a CTA on every SM, but at least `min_chunks` chunks
  // (64 pixels each) of main loop per CTA
  static int min_chunks = 0;
  if (!min_chunks) { const char *e = getenv("CTAGAN_WG_MIN_CHUNKS"); min_chunks = e ? atoi(e) : 32; if (min_chunks < 1) min_chunks = 1; }     // (measured: Cyc 4.67 -> 4.60 ms against 4)
  static int max_cluster = 0;
  if (!max_cluster) { const char *e = getenv("CTAGAN_WG_MAX_CLUSTER"); max_cluster = e ? atoi(e) : 8; if (max_cluster < 1 || max_cluster > 8) max_cluster = 8; }
  // a cluster of S CTAs (one per SM: the operand ring fills the shared memory) must fit into ONE GPC (~18 SMs), so at most
  // 8 * floor(18 / S) clusters are resident at a time (16 of 8, 24 of 6, 32 of 4, 48 of 3, 72 of 2) -- more tiles than that would run as
  // two waves.  Pick the S <= max_cluster that puts the most CTAs on the chip in ONE wave (18 tiles of a res-block layer: 6 x 18 = 108
  // CTAs, not 4 x 18 = 72), subject to >= min_chunks chunks of main loop per CTA.
  int splits = 1, best = pl.tiles * n_groups < ctagan_num_sms() ? pl.tiles * n_groups : ctagan_num_sms();
  for (int s2 = 2; s2 <= max_cluster; ++s2) {
    if (pl.total_chunks < s2 * min_chunks) break;
    const int clusters = pl.tiles * n_groups, cap = 8 * (18 / s2);
    if (clusters > cap || clusters * s2 > ctagan_num_sms()) continue;
    if (clusters * s2 > best) { best = clusters * s2; splits = s2; }
  }
  pl.splits = splits;
  pl.cps = (pl.total_chunks + splits - 1) / splits;
  return true;
}

template <int BNW, bool ACC>
int launch_wg(const CUtensorMap &my, const CUtensorMap &mx, const WgParams &p, dim3 grid, cudaStream_t st) {
  using Cfg = WgConfig<BNW>;
  static bool configured = false;
  if (!configured) {
    CTAGAN_CUDA_OK(cudaFuncSetAttribute(conv_wgrad_tc_kernel<BNW, ACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    configured = true;
  }
  CTAGAN_CUDA_OK(launch_cluster_pdl(conv_wgrad_tc_kernel<BNW, ACC>, grid, dim3(192), Cfg::SMEM_BYTES, st, (unsigned)p.splits, my, mx, p));
  CTAGAN_LAUNCH_OK();
  return CTAGAN_OK;
}

}  // namespace

int ctagan_conv_wgrad_tc_eligible(const ctagan_conv_geom *g, int n_groups) {
  WgPlan pl;
  return (n_groups >= 1 && g->N % n_groups == 0 && plan_wgrad(g, pl, n_groups)) ? 1 : 0;
}

// workspace: split-K partial sums [groups][tiles][splits][128][BNW] fp32 (splits > 1), then the bias-gradient partial sums
// [groups][WG_DB_BLOCKS][Co]
static size_t wg_split_bytes(const WgPlan &pl, int n_groups) {
  return pl.splits > 1 ? (size_t)n_groups * pl.tiles * pl.splits * WG_M * pl.bnw * sizeof(float) : 0;
}
static int wg_db_blocks() { return 2 * ctagan_num_sms(); }

size_t ctagan_conv_wgrad_tc_workspace(const ctagan_conv_geom *g, int n_groups) {
  WgPlan pl;
  if (n_groups < 1 || g->N % n_groups || !plan_wgrad(g, pl, n_groups)) return 0;
  return wg_split_bytes(pl, n_groups) + (size_t)n_groups * wg_db_blocks() * g->Co * sizeof(float);
}

// n_groups > 1: the batch is n_groups consecutive image groups and dw / db hold one gradient per group ([groups][Co][Ci][KH][KW])
int ctagan_conv_wgrad_tc(const ctagan_conv_geom *g, const void *gy, const void *gx, float *dw, float *db, void *workspace,
                         size_t workspace_bytes, cudaStream_t st, int n_groups, int accumulate) {
  WgPlan pl;
  if (n_groups < 1 || g->N % n_groups || !plan_wgrad(g, pl, n_groups)) {
    ctagan_set_error("conv_wgrad: geometry not supported by the tcgen05 engine");
    return CTAGAN_ERR_UNSUPPORTED;
  }
  const size_t need = ctagan_conv_wgrad_tc_workspace(g, n_groups);
  CTAGAN_REQUIRE(workspace && workspace_bytes >= need, "conv_wgrad(tc): workspace of %zu bytes required (got %zu)", need, workspace_bytes);
  CUtensorMap my, mx;
  int rc = make_map_4d(&my, gy, g->N, g->Ho, g->Wo, g->Co, pl.bkw, pl.bkh, 1);
  if (rc) return rc;
  rc = make_map_4d(&mx, gx, g->N, g->Hi, g->Wi, g->Ci, pl.bkw * g->stride, pl.bkh * g->stride, g->stride);
  if (rc) return rc;
  WgParams p;
  p.ntaps = g->KH * g->KW; p.KW = g->KW; p.Co = g->Co; p.Ci = g->Ci; p.stride = g->stride; p.pad = g->pad_h;
  p.margin = g->gy_margin; p.rb = pl.rb; p.cb = pl.cb; p.bkh = pl.bkh; p.bkw = pl.bkw;
  p.total_chunks = pl.total_chunks; p.chunks_per_split = pl.cps; p.ci_tiles = (g->Ci + pl.bnw - 1) / pl.bnw; p.dw = dw;
  p.splits = pl.splits;
  p.dw_packed = (accumulate & CTAGAN_WGRAD_PACKED) ? 1 : 0;
  accumulate &= CTAGAN_WGRAD_ACCUMULATE;
  p.ws = (float *)workspace;
  p.prof = nullptr;
  static long long *prof_buf = nullptr;
  static int prof_on = -1;
  if (prof_on < 0) { const char *e = getenv("CTAGAN_WG_PROF"); prof_on = e ? atoi(e) : 0; }
  if (prof_on) {
    if (!prof_buf) cudaMallocManaged(&prof_buf, 16 * sizeof(long long));
    p.prof = prof_buf;
  }
  dim3 grid((unsigned)pl.splits, (unsigned)pl.tiles, (unsigned)n_groups);        // x = K split = rank in the tile's cluster
  switch (pl.bnw) {
    case 256: rc = accumulate ? launch_wg<256, true>(my, mx, p, grid, st) : launch_wg<256, false>(my, mx, p, grid, st); break;
    case 128: rc = accumulate ? launch_wg<128, true>(my, mx, p, grid, st) : launch_wg<128, false>(my, mx, p, grid, st); break;
    default: rc = accumulate ? launch_wg<64, true>(my, mx, p, grid, st) : launch_wg<64, false>(my, mx, p, grid, st); break;
  }
  if (rc) return rc;
  if (prof_on) {
    cudaStreamSynchronize(st);
    fprintf(stderr, "[wgrad prof] Co=%d Ci=%d K=%d N=%d grid=(%d,%d,%d) bnw=%d iters=%lld | producer total %lld wait_empty %lld | mma total %lld wait_full %lld | epi wait_acc %lld | "
            "dump %lld to-L2 %lld sync %lld sum+store %lld (clk)\n", g->Co, g->Ci, g->KH, g->N, pl.splits, pl.tiles, n_groups, pl.bnw, prof_buf[8], prof_buf[0], prof_buf[1], prof_buf[2],
            prof_buf[3], prof_buf[4], prof_buf[5], prof_buf[6], prof_buf[7], prof_buf[9]);
  }
  if (db) {
    const long long pixels = (long long)(g->N / n_groups) * g->Ho * g->Wo;
    long long blocks = (pixels + 511) / 512;
    if (blocks > wg_db_blocks()) blocks = wg_db_blocks();
    const long long ppb = (pixels + blocks - 1) / blocks;
    const int nb = (int)((pixels + ppb - 1) / ppb);
    float *part = reinterpret_cast<float *>(reinterpret_cast<uint8_t *>(workspace) + wg_split_bytes(pl, n_groups));
    for (int k = 0; k < n_groups; ++k) {
      float *pk = part + (size_t)k * wg_db_blocks() * g->Co;
      colsum_kernel<<<nb, 256, 0, st>>>((const bf16 *)gy + (size_t)k * pixels * g->Co, pk, pixels, g->Co, ppb);
      CTAGAN_LAUNCH_OK();
      rc = ctagan_ordered_sum(pk, db + (size_t)k * g->Co, nb, g->Co, st, accumulate);
      if (rc) return rc;
    }
  }
  return CTAGAN_OK;
}

// =====================================================================================================================
// Weight gradient of the layers with 1-2 channels on one side (7x7 head Cin = 1, 7x7 tail Cout = 1, the discriminator's first and
// last layer, Reg's first conv) on the tensor cores.
//
// V = the wide tensor ([N][VH][VW][C], C <= 128), S = the thin one (SC <= 2 channels):
//   gx thin (gy_thin = 0):  dw[c][sc][kh][kw] = sum_v V[v][c] * S[n, vh*s + kh - pad, vw*s + kw - pad, sc]       (V = gy, v = output position)
//   gy thin (gy_thin = 1):  dw[sc][c][kh][kw] = sum_v V[v][c] * S[n, vh + pad - kh,  vw + pad - kw,  sc]        (V = gx, v = input position)
// Both are D[c][col] = sum_v V[v][c] * P[v][col] with the PATCH MATRIX P[v][col = sc*taps + tap] of the thin tensor: a GEMM with
// M = 128 (c, channels past C are TMA zero fill), N = 64 (sc*taps <= 64 columns, the rest zero), K = positions.  V tiles of 64
// positions arrive by TMA ([64 px][128 B] slabs = MN-major A operand, as in conv_wgrad_tc_kernel); the matching P tile -- [64 px]
// [64 cols] bf16 = the MN-major B operand -- is BUILT IN SHARED MEMORY by four builder warps from the few rows of the thin tensor the
// chunk touches (staged in shared memory first), in the 128-byte-swizzled layout the MMA reads.  Column SC*taps of P is set to 1, so
// the same GEMM also yields db[c] = sum_v gy[v][c] (gx thin).  Every CTA walks a contiguous range of chunks and stores its partial
// D in its own row of the workspace; ctagan_ordered_sum adds the rows in CTA order (deterministic).
// Against the CUDA-core kernels (conv_wgrad_thin_vec_kernel: FMA-issue bound): 7x7 head, batch 1, 256^2: ~100 us -> ~10 us.
// =====================================================================================================================
namespace {

struct ThinTcParams {
  int gy_thin;
  int C, SC, KH, KW, stride, pad;
  int VH, VW, SH, SW;
  int ncols;                  // SC * taps (+ 1 ones column when the bias gradient rides along)
  int ones_col;               // index of the ones column or -1
  int chunks_w;               // 64-position chunks per row of V
  int total_chunks, chunks_per_cta;
  int win_w;                  // window columns (positions of S) a chunk touches
  float *part;                // [ctas][part_elems]
  int part_elems;             // Co*Ci*taps (+ Co for db)
  int db_off;                 // offset of db inside a partial row, or -1
  const bf16 *S;
};

constexpr int TT_STAGES = 4;
constexpr int TT_V_BYTES = 2 * WG_SLAB;          // 128 channels x 64 positions
constexpr int TT_P_BYTES = WG_SLAB;              // 64 columns x 64 positions
constexpr int TT_STAGE_BYTES = TT_V_BYTES + TT_P_BYTES;
constexpr int TT_WIN_MAX = 7 * 136 * 2;          // KH <= 7 rows x (63 * 2 + 7 + pad) columns x SC <= 2, bf16 elements
constexpr int TT_SMEM_BYTES = TT_STAGES * TT_STAGE_BYTES + 1024 + 256;

__global__ void __launch_bounds__(192, 1)
conv_wgrad_thin_tc_kernel(const __grid_constant__ CUtensorMap map_v, const ThinTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + TT_STAGES * TT_STAGE_BYTES);
  uint64_t *empty_bar = full_bar + TT_STAGES;
  uint64_t *tmem_full_bar = empty_bar + TT_STAGES;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_full_bar + 1);
  __shared__ bf16 win[2][TT_WIN_MAX];             // double-buffered window of the thin tensor
  __shared__ short lut[64];                       // column -> offset inside the window (-1: zero column, -2: ones column)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int chunk_lo = blockIdx.x * p.chunks_per_cta;
  const int n_iters = max(0, min(p.total_chunks, chunk_lo + p.chunks_per_cta) - chunk_lo);
  const int taps = p.KH * p.KW;
  const int win_pitch = p.win_w * p.SC;           // elements per window row

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_v);
    for (int s = 0; s < TT_STAGES; ++s) {
      mbar_init(&full_bar[s], 1 + 4);             // the producer's expect_tx arrival + one arrival per builder warp
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (threadIdx.x < 64) {
    const int col = threadIdx.x;
    int v = -1;
    if (col < p.SC * taps) {
      const int sc = col / taps, tap = col - sc * taps;
      const int kh = tap / p.KW, kw = tap - kh * p.KW;
      // window row / column of the sample that position r = 0 of the chunk needs for this tap
      const int wr = p.gy_thin ? (p.KH - 1 - kh) : kh;
      const int wc = p.gy_thin ? (p.KW - 1 - kw) : kw;
      v = wr * win_pitch + wc * p.SC + sc;
    } else if (col == p.ones_col) {
      v = -2;
    }
    lut[col] = (short)v;
  }
  if (warp == 1) tmem_alloc<64>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if (warp == 0) {
    // ===== TMA producer: V tiles =====
    for (int it = 0; it < n_iters; ++it) {
      const int s = it % TT_STAGES;
      const uint32_t ph = (uint32_t)(it / TT_STAGES) & 1u;
      mbar_wait(&empty_bar[s], ph ^ 1u);
      int ch = chunk_lo + it;
      const int cw = ch % p.chunks_w; ch /= p.chunks_w;
      const int vh = ch % p.VH;
      const int n = ch / p.VH;
      if (elect_one()) {
        uint8_t *v_dst = smem + s * TT_STAGE_BYTES;
        mbar_expect_tx(&full_bar[s], TT_V_BYTES);
        tma_load_4d(&map_v, &full_bar[s], v_dst, 0, cw * 64, vh, n);
        tma_load_4d(&map_v, &full_bar[s], v_dst + WG_SLAB, 64, cw * 64, vh, n);       // channels >= C: zero fill
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    constexpr uint32_t idesc = make_idesc_bf16_mn(WG_M, 64);
    for (int it = 0; it < n_iters; ++it) {
      const int s = it % TT_STAGES;
      const uint32_t ph = (uint32_t)(it / TT_STAGES) & 1u;
      mbar_wait(&full_bar[s], ph);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t a_addr = smem_u32(smem + s * TT_STAGE_BYTES);
        const uint32_t b_addr = a_addr + TT_V_BYTES;
#pragma unroll
        for (int k = 0; k < WG_KPIX / UMMA_K; ++k) {
          const uint64_t adesc = make_mnmajor_sw128_desc(a_addr + k * 2048, WG_SLAB);
          const uint64_t bdesc = make_mnmajor_sw128_desc(b_addr + k * 2048, WG_SLAB);
          umma_bf16(tmem_base, adesc, bdesc, idesc, (it > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);
        if (it == n_iters - 1) umma_commit(tmem_full_bar);
      }
      __syncwarp();
    }
  } else {
    // ===== builders (warps 2..5, 128 threads): window of the thin tensor -> shared memory -> swizzled patch tile =====
    const int bt = (int)threadIdx.x - 64;          // 0..127
    const int c8 = bt & 7;                          // this thread's 16-byte piece (8 columns) of every row it builds
    int off[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) off[e] = lut[c8 * 8 + e];
    const int rstep = (p.gy_thin ? 1 : p.stride) * p.SC;       // window elements per position step
    const int win_elems = p.KH * win_pitch;
    constexpr int WE = (TT_WIN_MAX + 127) / 128;               // window elements per builder thread (at most)
    // this thread's window elements e = bt, bt + 128, ...: (row, column * SC + sc) once, without per-chunk divisions
    short e_wr[WE], e_rem[WE];
#pragma unroll
    for (int q = 0; q < WE; ++q) {
      const int e = bt + q * 128;
      e_wr[q] = (short)(e < win_elems ? e / win_pitch : -1);
      e_rem[q] = (short)(e < win_elems ? e - (e / win_pitch) * win_pitch : 0);
    }
    unsigned short pre[WE];                                    // the NEXT chunk's window values, loaded while the current tile is built
    auto prefetch = [&](int it) {
      int ch = chunk_lo + it;
      const int cw = ch % p.chunks_w; ch /= p.chunks_w;
      const int vh = ch % p.VH;
      const int n = ch / p.VH;
      // window origin in S: row of window row 0, column * SC of window column 0
      const int sh0 = p.gy_thin ? vh + p.pad - (p.KH - 1) : vh * p.stride - p.pad;
      const int sw0 = p.gy_thin ? cw * 64 + p.pad - (p.KW - 1) : cw * 64 * p.stride - p.pad;
      const unsigned short *S16 = reinterpret_cast<const unsigned short *>(p.S) + (long long)n * p.SH * p.SW * p.SC;
#pragma unroll
      for (int q = 0; q < WE; ++q) {
        const int sh = sh0 + e_wr[q];
        const int col = sw0 * p.SC + e_rem[q];                 // element index inside the row of S (column * SC + sc)
        pre[q] = (e_wr[q] >= 0 && sh >= 0 && sh < p.SH && col >= 0 && col < p.SW * p.SC) ? __ldg(S16 + (long long)sh * p.SW * p.SC + col) : (unsigned short)0;
      }
    };
    if (n_iters > 0) prefetch(0);
    for (int it = 0; it < n_iters; ++it) {
      const int s = it % TT_STAGES;
      const uint32_t ph = (uint32_t)(it / TT_STAGES) & 1u;
      bf16 *w = win[it & 1];
#pragma unroll
      for (int q = 0; q < WE; ++q)
        if (e_wr[q] >= 0) w[bt + q * 128] = __ushort_as_bfloat16(pre[q]);
      asm volatile("bar.sync 1, 128;" ::: "memory");           // window complete (the other window buffer is still being read by nobody:
                                                               // every builder passed the previous bar.sync after its last read of it)
      if (it + 1 < n_iters) prefetch(it + 1);                  // in flight while this tile is built
      mbar_wait(&empty_bar[s], ph ^ 1u);                       // the MMAs that read this stage's previous patch tile have retired
      uint8_t *p_dst = smem + s * TT_STAGE_BYTES + TT_V_BYTES;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int r = (bt >> 3) + 16 * k;                      // position inside the chunk
        const int base = r * rstep;
        uint32_t pk[4];
#pragma unroll
        for (int e = 0; e < 8; e += 2) {
          const unsigned short lo = off[e] >= 0 ? __bfloat16_as_ushort(w[off[e] + base]) : (off[e] == -2 ? (unsigned short)0x3F80 : (unsigned short)0);
          const unsigned short hi = off[e + 1] >= 0 ? __bfloat16_as_ushort(w[off[e + 1] + base]) : (off[e + 1] == -2 ? (unsigned short)0x3F80 : (unsigned short)0);
          pk[e / 2] = (uint32_t)lo | ((uint32_t)hi << 16);
        }
        *reinterpret_cast<uint4 *>(p_dst + r * 128 + ((c8 ^ (r & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
      fence_proxy_async();                                     // generic-proxy stores -> visible to the tensor core's async-proxy reads
      __syncwarp();
      if (lane == 0) mbar_arrive(&full_bar[s]);
    }
    // ===== epilogue: this CTA's partial D -> its row of the workspace, in the final dw (and db) layout =====
    const int quarter = warp & 3;
    const int c = quarter * 32 + lane;               // channel of V == TMEM lane
    float *row = p.part + (long long)blockIdx.x * p.part_elems;
    if (n_iters > 0) {
      mbar_wait(tmem_full_bar, 0);
      tc_fence_after();
    }
#pragma unroll 1
    for (int c0 = 0; c0 < 64; c0 += 32) {
      uint32_t r[32];
      if (n_iters > 0) {
        tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, r);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int e = 0; e < 32; ++e) r[e] = 0u;
      }
      if (c < p.C) {
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const int col = c0 + e;
          if (col < p.SC * taps) {
            const int sc = col / taps, tap = col - sc * taps;
            const long long idx = p.gy_thin ? ((long long)sc * p.C + c) * taps + tap : ((long long)c * p.SC + sc) * taps + tap;
            row[idx] = __uint_as_float(r[e]);
          } else if (col == p.ones_col && p.db_off >= 0) {
            row[p.db_off + c] = __uint_as_float(r[e]);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<64>(tmem_base);
  }
}

// sum of a thin (1-2 channel) tensor per channel: part[block][SC], fixed order inside the block (the bias gradient of a Cout <= 2 layer)
__global__ void __launch_bounds__(256) thin_sum_kernel(const bf16 *__restrict__ x, float *__restrict__ part, long long pixels, int SC, int part_stride) {
  __shared__ float wsum[8];
  const long long per = (pixels + gridDim.x - 1) / gridDim.x;
  const long long p0 = blockIdx.x * per, p1 = min(pixels, p0 + per);
  for (int s = 0; s < SC; ++s) {
    float t = 0.f;
    for (long long q = p0 + threadIdx.x; q < p1; q += blockDim.x) t += __bfloat162float(x[q * SC + s]);
    t = warp_sum(t);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
      float a = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) a += wsum[w];
      part[(long long)blockIdx.x * part_stride + s] = a;
    }
  }
}

struct ThinTcPlan {
  int gy_thin, C, SC, ctas, cpc, total_chunks, chunks_w, win_w;
};

bool plan_thin_tc(const ctagan_conv_geom *g, ThinTcPlan &pl) {
  if (g->dtype != CTAGAN_BF16 || g->dil != 1 || g->gy_margin != 0) return false;
  if (g->KH > 7 || g->KW > 7 || g->stride > 2) return false;
  static int enabled = -1;
  if (enabled < 0) { const char *e = getenv("CTAGAN_THIN_TC"); enabled = e ? atoi(e) : 1; }
  if (!enabled) return false;
  if (g->Ci <= 2 && g->Co >= 8) pl.gy_thin = 0;
  else if (g->Co <= 2 && g->Ci >= 8 && g->stride == 1) pl.gy_thin = 1;
  else return false;
  pl.C = pl.gy_thin ? g->Ci : g->Co;
  pl.SC = pl.gy_thin ? g->Co : g->Ci;
  if (pl.C % 8 || pl.C > 128) return false;
  if (pl.SC * g->KH * g->KW + 1 > 64) return false;
  const int VH = pl.gy_thin ? g->Hi : g->Ho, VW = pl.gy_thin ? g->Wi : g->Wo;
  if (VW < 64) return false;                                   // (a ragged last chunk of a row is TMA zero fill: it adds nothing)
  if ((long long)g->N * VH * VW < 4096) return false;          // tiny maps: the CUDA-core kernels
  pl.chunks_w = (VW + 63) / 64;
  pl.total_chunks = g->N * VH * pl.chunks_w;
  pl.win_w = pl.gy_thin ? 63 + g->KW : 63 * g->stride + g->KW;
  if (g->KH * pl.win_w * pl.SC > TT_WIN_MAX) return false;
  int ctas = ctagan_num_sms();
  if (ctas > pl.total_chunks / 2) ctas = pl.total_chunks / 2 > 0 ? pl.total_chunks / 2 : 1;     // at least two chunks per CTA
  pl.cpc = (pl.total_chunks + ctas - 1) / ctas;
  pl.ctas = (pl.total_chunks + pl.cpc - 1) / pl.cpc;
  return true;
}

}  // namespace

int ctagan_conv_wgrad_thin_tc_eligible(const ctagan_conv_geom *g) {
  ThinTcPlan pl;
  return plan_thin_tc(g, pl) ? 1 : 0;
}

// partial sums [ctas][Co*Ci*taps + Co]
size_t ctagan_conv_wgrad_thin_tc_workspace(const ctagan_conv_geom *g) {
  ThinTcPlan pl;
  if (!plan_thin_tc(g, pl)) return 0;
  return (size_t)pl.ctas * ((size_t)g->Co * g->Ci * g->KH * g->KW + g->Co) * sizeof(float);
}

int ctagan_conv_wgrad_thin_tc(const ctagan_conv_geom *g, const void *gy, const void *gx, float *dw, float *db, void *workspace,
                              size_t workspace_bytes, cudaStream_t st, int accumulate) {
  ThinTcPlan pl;
  if (!plan_thin_tc(g, pl)) {
    ctagan_set_error("conv_wgrad_thin_tc: geometry not supported");
    return CTAGAN_ERR_UNSUPPORTED;
  }
  const size_t need = ctagan_conv_wgrad_thin_tc_workspace(g);
  CTAGAN_REQUIRE(workspace && workspace_bytes >= need, "conv_wgrad(thin tc): workspace of %zu bytes required (got %zu)", need, workspace_bytes);
  const void *V = pl.gy_thin ? gx : gy;
  const void *S = pl.gy_thin ? gy : gx;
  CTAGAN_REQUIRE((reinterpret_cast<uintptr_t>(V) & 15) == 0, "conv_wgrad(thin tc): the wide tensor must be 16-byte aligned");
  const int VH = pl.gy_thin ? g->Hi : g->Ho, VW = pl.gy_thin ? g->Wi : g->Wo;
  CUtensorMap mv;
  int rc = make_map_4d(&mv, V, g->N, VH, VW, pl.C, 64, 1, 1);
  if (rc) return rc;
  const long long dw_elems = (long long)g->Co * g->Ci * g->KH * g->KW;
  ThinTcParams p;
  memset(&p, 0, sizeof(p));
  p.gy_thin = pl.gy_thin; p.C = pl.C; p.SC = pl.SC; p.KH = g->KH; p.KW = g->KW; p.stride = g->stride; p.pad = g->pad_h;
  p.VH = VH; p.VW = VW; p.SH = pl.gy_thin ? g->Ho : g->Hi; p.SW = pl.gy_thin ? g->Wo : g->Wi;
  p.ncols = pl.SC * g->KH * g->KW;
  const bool ones = db != nullptr && !pl.gy_thin;         // db[c] = sum of gy rides along as a column of ones
  p.ones_col = ones ? p.ncols : -1;
  p.chunks_w = pl.chunks_w; p.total_chunks = pl.total_chunks; p.chunks_per_cta = pl.cpc; p.win_w = pl.win_w;
  p.part = (float *)workspace;
  p.part_elems = (int)(dw_elems + g->Co);
  p.db_off = ones ? (int)dw_elems : -1;
  p.S = (const bf16 *)S;
  static bool configured = false;
  if (!configured) {
    CTAGAN_CUDA_OK(cudaFuncSetAttribute(conv_wgrad_thin_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TT_SMEM_BYTES));
    configured = true;
  }
  CTAGAN_CUDA_OK(launch_pdl(conv_wgrad_thin_tc_kernel, dim3((unsigned)pl.ctas), dim3(192), (size_t)TT_SMEM_BYTES, st, mv, p));
  CTAGAN_LAUNCH_OK();
  int parts = pl.ctas;
  if (db != nullptr && pl.gy_thin) {                      // bias gradient of the thin output: a plain sum of gy
    const long long pixels = (long long)g->N * g->Ho * g->Wo;
    int blocks = pl.ctas;
    thin_sum_kernel<<<blocks, 256, 0, st>>>((const bf16 *)gy, p.part + dw_elems, pixels, g->Co, p.part_elems);
    CTAGAN_LAUNCH_OK();
  }
  // rows are [dw | db]: one launch adds the CTA rows (one warp per element) into the two destinations
  return ctagan_ordered_sum_rows2(p.part, dw, dw_elems, db, db ? g->Co : 0, parts, p.part_elems, st, accumulate & CTAGAN_WGRAD_ACCUMULATE);
}

// =====================================================================================================================
// Forward convolution of the layers with 1-2 INPUT channels (7x7 head Cin = 1, the discriminator's first layer, Reg's first conv; also
// the input gradient of a Cout <= 2 layer, which is the same operation on flipped weights) on the tensor cores:
//   y[px][co] = act(bias[co] + sum_col P[px][col] * wp[co][col]),   col = (kh * KW + kw) * Ci + ci  (<= 64 columns)
// i.e. a GEMM with M = 128 output positions, N = 64 (Co <= 64), K = 64: the A operand is the PATCH MATRIX of the thin input, built in
// shared memory (128-byte-swizzled K-major rows) by four builder warps from a window of the input staged in shared memory; B = the packed
// weights, padded to 64 columns, resident in shared memory for the whole kernel.  Persistent CTAs walk the tiles; the accumulator is
// double-buffered in tensor memory so that the epilogue warps (bias / activation / bf16 stores / fused InstanceNorm statistics: the
// shared tc_epilogue) work on tile i while the builders and the MMAs are on tile i+1.  No TMA: nothing here is a dense tile in memory.
// Roles (288 threads): warp 0 = TMEM allocation + MMA issuer, warps 1-4 = builders, warps 5-8 = epilogue.
// =====================================================================================================================
namespace {

struct FewinParams {
  int SC, KH, KW, stride, pad;
  int Hi, Wi;
  int ncols;                   // KH * KW * SC
  int win_w, win_h;            // window of input positions a tile touches
  int n_tiles;                 // N * tiles_per_img
  const bf16 *x;
  const bf16 *wp;              // [Co][ncols]
};

constexpr int FI_STAGES = 3;
constexpr int FI_P_BYTES = TILE_M * 128;          // 128 positions x 64 columns
constexpr int FI_B_BYTES = 64 * 128;              // 64 output channels x 64 columns
constexpr int FI_WIN_MAX = 2304;                  // bf16 elements of the input window (7 x 134 x 2, 4 x 258 x 2, ...)
constexpr int FI_SMEM_BYTES = FI_STAGES * FI_P_BYTES + FI_B_BYTES + 1024 + 256;

__global__ void __launch_bounds__(288, 1)
conv_fewin_tc_kernel(const TcParams p, const FewinParams f) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t *b_smem = smem + FI_STAGES * FI_P_BYTES;
  uint64_t *p_full = reinterpret_cast<uint64_t *>(b_smem + FI_B_BYTES);
  uint64_t *p_empty = p_full + FI_STAGES;
  uint64_t *tmem_full_bar = p_empty + FI_STAGES;       // [2]
  uint64_t *tmem_empty_bar = tmem_full_bar + 2;        // [2]
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_empty_bar + 2);
  __shared__ bf16 win[2][FI_WIN_MAX];
  __shared__ short lut[64];
  __shared__ EpiSmem<64> epi;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int win_pitch = f.win_w * f.SC;
  const int BW = 1 << p.bw_log2;
  const int my_tiles = (int)blockIdx.x < f.n_tiles ? (f.n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < FI_STAGES; ++s) {
      mbar_init(&p_full[s], 4);                    // one arrival per builder warp
      mbar_init(&p_empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], 4);            // one arrival per epilogue warp
    }
    fence_barrier_init();
  }
  if (threadIdx.x < 64) {
    const int col = threadIdx.x;
    int v = -1;
    if (col < f.ncols) {
      const int tap = col / f.SC, sc = col - tap * f.SC;
      const int kh = tap / f.KW, kw = tap - kh * f.KW;
      v = kh * win_pitch + kw * f.SC + sc;
    }
    lut[col] = (short)v;
  }
  pdl_wait();
  // resident B operand: wp[co][col] -> K-major rows of 128 bytes, 16-byte pieces XOR-swizzled by (row & 7); rows >= Co and columns >= ncols zero
  for (int e = threadIdx.x; e < 64 * 8; e += blockDim.x) {
    const int co = e >> 3, c8 = e & 7;
    uint32_t pk[4];
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      unsigned short v2[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int col = c8 * 8 + h * 2 + u;
        v2[u] = (co < p.Co && col < f.ncols) ? __bfloat16_as_ushort(f.wp[(long long)co * f.ncols + col]) : (unsigned short)0;
      }
      pk[h] = (uint32_t)v2[0] | ((uint32_t)v2[1] << 16);
    }
    *reinterpret_cast<uint4 *>(b_smem + co * 128 + ((c8 ^ (co & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  }
  fence_proxy_async();
  if (warp == 0) tmem_alloc<128>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== MMA issuer =====
    constexpr uint32_t idesc = make_idesc_bf16(TILE_M, 64);
    const uint64_t bdesc = make_kmajor_sw128_desc(smem_u32(b_smem));
    for (int lt = 0; lt < my_tiles; ++lt) {
      const int s = lt % FI_STAGES, a = lt & 1;
      mbar_wait(&tmem_empty_bar[a], (((uint32_t)lt >> 1) & 1u) ^ 1u);
      mbar_wait(&p_full[s], (uint32_t)(lt / FI_STAGES) & 1u);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t adesc = make_kmajor_sw128_desc(smem_u32(smem + s * FI_P_BYTES));
#pragma unroll
        for (int kk = 0; kk < CHUNK_K / UMMA_K; ++kk)
          umma_bf16(tmem_base + (uint32_t)(a * 64), adesc + (uint64_t)(2 * kk), bdesc + (uint64_t)(2 * kk), idesc, kk > 0 ? 1u : 0u);
        umma_commit(&p_empty[s]);
        umma_commit(&tmem_full_bar[a]);
      }
      __syncwarp();
    }
  } else if (warp <= 4) {
    // ===== builders =====
    const int bt = (int)threadIdx.x - 32;            // 0..127
    const int c8 = bt & 7;
    int off[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) off[e] = lut[c8 * 8 + e];
    const int win_elems = f.win_h * win_pitch;
    constexpr int WE = FI_WIN_MAX / 128;             // 18
    short e_wr[WE], e_rem[WE];
#pragma unroll
    for (int q = 0; q < WE; ++q) {
      const int e = bt + q * 128;
      e_wr[q] = (short)(e < win_elems ? e / win_pitch : -1);
      e_rem[q] = (short)(e < win_elems ? e - (e / win_pitch) * win_pitch : 0);
    }
    unsigned short pre[WE];
    auto prefetch = [&](int lt) {
      const int t = (int)blockIdx.x + lt * (int)gridDim.x;
      const int img = t / p.tiles_per_img, tile = t - img * p.tiles_per_img;
      const int ti0 = (tile / p.tiles_w) * (TILE_M >> p.bw_log2), tj0 = (tile % p.tiles_w) * BW;
      const int sh0 = ti0 * f.stride - f.pad, sc0 = (tj0 * f.stride - f.pad) * f.SC;
      const unsigned short *X16 = reinterpret_cast<const unsigned short *>(f.x) + (long long)img * f.Hi * f.Wi * f.SC;
#pragma unroll
      for (int q = 0; q < WE; ++q) {
        const int sh = sh0 + e_wr[q];
        const int col = sc0 + e_rem[q];
        pre[q] = (e_wr[q] >= 0 && sh >= 0 && sh < f.Hi && col >= 0 && col < f.Wi * f.SC) ? __ldg(X16 + (long long)sh * f.Wi * f.SC + col) : (unsigned short)0;
      }
    };
    if (my_tiles > 0) prefetch(0);
    for (int lt = 0; lt < my_tiles; ++lt) {
      const int s = lt % FI_STAGES;
      bf16 *w = win[lt & 1];
#pragma unroll
      for (int q = 0; q < WE; ++q)
        if (e_wr[q] >= 0) w[bt + q * 128] = __ushort_as_bfloat16(pre[q]);
      asm volatile("bar.sync 2, 128;" ::: "memory");
      if (lt + 1 < my_tiles) prefetch(lt + 1);
      mbar_wait(&p_empty[s], ((uint32_t)(lt / FI_STAGES) & 1u) ^ 1u);
      uint8_t *p_dst = smem + s * FI_P_BYTES;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int r = (bt >> 3) + 16 * k;                          // position inside the tile
        const int base = ((r >> p.bw_log2) * win_pitch + (r & (BW - 1)) * f.SC) * f.stride;
        uint32_t pk[4];
#pragma unroll
        for (int e = 0; e < 8; e += 2) {
          const unsigned short lo = off[e] >= 0 ? __bfloat16_as_ushort(w[off[e] + base]) : (unsigned short)0;
          const unsigned short hi = off[e + 1] >= 0 ? __bfloat16_as_ushort(w[off[e + 1] + base]) : (unsigned short)0;
          pk[e / 2] = (uint32_t)lo | ((uint32_t)hi << 16);
        }
        *reinterpret_cast<uint4 *>(p_dst + r * 128 + ((c8 ^ (r & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[s]);
    }
  } else {
    // ===== epilogue (warps 5..8: TMEM lane quarters 1, 2, 3, 0) =====
    // InstanceNorm statistics: a layer of this kind has hundreds of tiles per image, so the per-tile slot + ticket of the dense kernel
    // would put one __threadfence per tile on this warpgroup and ~30 dependent rounds of slot reads on the last CTA.  Instead every
    // CTA keeps the column sums of ITS tiles of an image (tile order) and publishes ONE slot per image -- also for images it has no
    // tile of (zeros) -- so an image has exactly gridDim.x slots; the last CTA to arrive adds them in CTA order (deterministic).
    const int et = (int)threadIdx.x - 160;          // 0..127; thread et < 64 owns column et
    const bool stats = p.stat_part != nullptr;
    const int n_imgs = f.n_tiles / p.tiles_per_img;
    float acc_s = 0.f, acc_q = 0.f;
    int cur_img = 0;
    auto flush = [&](int img) {
      float2 *slots = reinterpret_cast<float2 *>(p.stat_part) + (long long)img * gridDim.x * p.Co;
      if (et < p.Co && et < 64) slots[(long long)blockIdx.x * p.Co + et] = make_float2(acc_s, acc_q);
      acc_s = acc_q = 0.f;
      __threadfence();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      unsigned int *ticket = p.stat_ticket + (long long)img * p.stat_tpi;
      if (et == 0) epi.ticket = atomicAdd(ticket, 1u);
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (epi.ticket == gridDim.x - 1u) {
        __threadfence();
        if (et < p.Co && et < 64) {
          double s1 = 0.0, s2 = 0.0;
          for (unsigned k0 = 0; k0 < gridDim.x; k0 += 16) {
            float2 v[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) v[u] = k0 + u < gridDim.x ? __ldcg(slots + (long long)(k0 + u) * p.Co + et) : make_float2(0.f, 0.f);
#pragma unroll
            for (int u = 0; u < 16; ++u) { s1 += (double)v[u].x; s2 += (double)v[u].y; }
          }
          const double inv = 1.0 / (double)p.stat_hw;
          const double m = s1 * inv;
          double var = s2 * inv - m * m;
          if (var < 0) var = 0;
          *reinterpret_cast<float2 *>(p.stat_out + ((long long)img * p.Co + et) * 2) = make_float2((float)m, (float)(1.0 / sqrt(var + 1e-5)));
        }
        if (et == 0) *ticket = 0u;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");       // epi.ticket is free again
    };
    for (int lt = 0; lt < my_tiles; ++lt) {
      const int a = lt & 1;
      const int t = (int)blockIdx.x + lt * (int)gridDim.x;
      const int img = t / p.tiles_per_img, tile = t - img * p.tiles_per_img;
      if (stats)
        for (; cur_img < img; ++cur_img) flush(cur_img);
      mbar_wait(&tmem_full_bar[a], ((uint32_t)lt >> 1) & 1u);
      tc_fence_after();
      tc_epilogue<64>(p, tmem_base + (uint32_t)(a * 64), warp, lane, img, tile, 0, 0, 0, epi, &tmem_empty_bar[a], 160, false);
      if (stats) {
        asm volatile("bar.sync 1, 128;" ::: "memory");     // the four warps' column sums of this tile are in epi.part
        if (et < 64) {
          acc_s += (epi.part[0][et][0] + epi.part[1][et][0]) + (epi.part[2][et][0] + epi.part[3][et][0]);
          acc_q += (epi.part[0][et][1] + epi.part[1][et][1]) + (epi.part[2][et][1] + epi.part[3][et][1]);
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");     // ... and may be overwritten by the next tile
      }
    }
    if (stats)
      for (; cur_img < n_imgs; ++cur_img) flush(cur_img);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<128>(tmem_base);
  }
}

bool plan_fewin_tc(const ctagan_conv_geom *g, int &bw_log2, int &win_w, int &win_h) {
  if (g->dtype != CTAGAN_BF16 || g->dil != 1 || g->stride > 2) return false;
  static int enabled = -1;
  if (enabled < 0) { const char *e = getenv("CTAGAN_THIN_TC"); enabled = e ? atoi(e) : 1; }
  if (!enabled) return false;
  if (g->Ci > 2 || g->Co > 64 || g->Co <= 32 || g->Co % 8) return false;
  if (g->KH * g->KW * g->Ci > 64 || g->KH > 7 || g->KW > 7) return false;
  // measured against the CUDA-core kernels (conv_fewin_tiled): a clear win for the 7x7 head (49 columns: 24 -> 14 us at batch 1, 145 ->
  // 61 us at batch 8) and for 4x4 on 2 channels (26 -> 19 us); slower for 3x3 / 4x4 on one channel and for 32 output channels
  if (g->KH * g->KW * g->Ci < 32) return false;
  if ((long long)g->N * g->Ho * g->Wo < 4096 || g->Wo < 16) return false;       // tiny maps: the CUDA-core kernels
  bw_log2 = ceil_log2(g->Wo < 128 ? g->Wo : 128);
  const int BW = 1 << bw_log2, BH = TILE_M >> bw_log2;
  win_w = (BW - 1) * g->stride + g->KW;
  win_h = (BH - 1) * g->stride + g->KH;
  return win_w * win_h * g->Ci <= FI_WIN_MAX;
}

}  // namespace

int ctagan_conv_fewin_tc_eligible(const ctagan_conv_geom *g) {
  int a, b, c;
  return plan_fewin_tc(g, a, b, c) ? 1 : 0;
}

size_t ctagan_conv_fewin_tc_stat_bytes(const ctagan_conv_geom *g) {
  int bwl, ww, wh;
  if (!plan_fewin_tc(g, bwl, ww, wh)) return 0;
  const int BW = 1 << bwl, BH = TILE_M >> bwl;
  const long long tiles = (long long)g->N * ((g->Wo + BW - 1) / BW) * ((g->Ho + BH - 1) / BH);
  const long long ctas = tiles < ctagan_num_sms() ? tiles : ctagan_num_sms();
  return (size_t)g->N * ctas * g->Co * 2 * sizeof(float);          // one slot per (image, CTA)
}

int ctagan_conv_fewin_tc(const ctagan_conv_geom *g, const void *x, const void *wp, const float *bias, void *y, unsigned int *stat_ticket,
                         void *stat_scratch, size_t stat_scratch_bytes, float *stat_out, cudaStream_t st) {
  int bwl, ww, wh;
  if (!plan_fewin_tc(g, bwl, ww, wh)) {
    ctagan_set_error("conv_fewin_tc: geometry not supported");
    return CTAGAN_ERR_UNSUPPORTED;
  }
  CTAGAN_REQUIRE((reinterpret_cast<uintptr_t>(y) & 15) == 0, "conv_fewin_tc: the output must be 16-byte aligned");
  TcParams p;
  memset(&p, 0, sizeof(p));
  p.mode = 1; p.Ci = g->Ci; p.Co = g->Co; p.stride = g->stride;
  p.out_H = g->Ho; p.out_W = g->Wo; p.sy = p.sx = 1; p.ay = p.ax = 0;
  p.Hov = g->Ho; p.Wov = g->Wo;
  p.act = g->act; p.bias = bias; p.out = (bf16 *)y;
  p.bw_log2 = bwl;
  const int BW = 1 << bwl, BH = TILE_M >> bwl;
  p.tiles_w = (g->Wo + BW - 1) / BW;
  p.tiles_per_img = p.tiles_w * ((g->Ho + BH - 1) / BH);
  p.imgs_per_group = g->N;
  if (stat_out) {
    CTAGAN_REQUIRE(stat_ticket && stat_scratch && stat_scratch_bytes >= ctagan_conv_fewin_tc_stat_bytes(g),
                   "conv_fewin_tc: ticket buffer and %zu bytes of scratch required", ctagan_conv_fewin_tc_stat_bytes(g));
    p.stat_part = (float *)stat_scratch; p.stat_ticket = stat_ticket; p.stat_out = stat_out;
    p.stat_hw = g->Ho * g->Wo; p.stat_parts = p.tiles_per_img; p.stat_part0 = 0;
    p.stat_tpi = (g->Co + 31) / 32;
  }
  FewinParams f;
  f.SC = g->Ci; f.KH = g->KH; f.KW = g->KW; f.stride = g->stride; f.pad = g->pad_h; f.Hi = g->Hi; f.Wi = g->Wi;
  f.ncols = g->KH * g->KW * g->Ci; f.win_w = ww; f.win_h = wh; f.n_tiles = g->N * p.tiles_per_img;
  f.x = (const bf16 *)x; f.wp = (const bf16 *)wp;
  static bool configured = false;
  if (!configured) {
    CTAGAN_CUDA_OK(cudaFuncSetAttribute(conv_fewin_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FI_SMEM_BYTES));
    configured = true;
  }
  const int ctas = f.n_tiles < ctagan_num_sms() ? f.n_tiles : ctagan_num_sms();
  CTAGAN_CUDA_OK(launch_pdl(conv_fewin_tc_kernel, dim3((unsigned)ctas), dim3(288), (size_t)FI_SMEM_BYTES, st, p, f));
  CTAGAN_LAUNCH_OK();
  return CTAGAN_OK;
}

// =====================================================================================================================
// Forward convolution of the layers with 1-2 OUTPUT channels and stride 1 (7x7 tail Cout = 1 + tanh, Reg's 2-channel flow head; also the
// input gradient of a Cin <= 2 layer = the same operation on flipped weights) on the tensor cores, in two steps:
//   1. Z^T[col][q] = sum_ci x[q][ci] * wp[col][ci],   col = co * taps + tap (<= 64 columns), q = every INPUT position:
//      ONE small GEMM per 128 positions (M = 128, N = 64, K = Ci <= 64) -- x is read exactly once, as plain 2-D TMA boxes;
//   2. y[n,i,j,co] = act(bias[co] + sum_tap Z^T[co * taps + tap][n, i + kh - pad, j + kw - pad]): a gather of taps fp32 values per
//      output, coalesced along j because Z is stored column-major (one plane per filter tap).
// Against the direct form (taps x 64 FMAs per output on CUDA cores, FMA-issue bound: 48 us at batch 1, 297 us at batch 8 for the 7x7
// tail) the arithmetic moves to the tensor pipe and what is left is bandwidth: x once, Z (taps x 4 bytes per position) written and read
// once (L2-resident at batch 1).
// =====================================================================================================================
namespace {

struct FewoutZParams {
  int Ci, ncols;               // ncols = Co * taps
  long long n_rows;            // N * Hi * Wi input positions
  int n_tiles;
  const bf16 *wp;              // [ncols][Ci]
  float *zt;                   // [ncols][n_rows]
};

constexpr int FO_STAGES = 4;
constexpr int FO_A_BYTES = TILE_M * 128;
constexpr int FO_SMEM_BYTES = FO_STAGES * FO_A_BYTES + FI_B_BYTES + 1024 + 256;

__global__ void __launch_bounds__(192, 1)
conv_fewout_z_kernel(const __grid_constant__ CUtensorMap map_x, const FewoutZParams f) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t *b_smem = smem + FO_STAGES * FO_A_BYTES;
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(b_smem + FI_B_BYTES);
  uint64_t *empty_bar = full_bar + FO_STAGES;
  uint64_t *tmem_full_bar = empty_bar + FO_STAGES;     // [2]
  uint64_t *tmem_empty_bar = tmem_full_bar + 2;        // [2]
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_empty_bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int my_tiles = (int)blockIdx.x < f.n_tiles ? (f.n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_x);
    for (int s = 0; s < FO_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], 4);
    }
    fence_barrier_init();
  }
  pdl_wait();
  // resident B operand: wp[col][ci] -> K-major rows of 128 bytes (swizzled); rows >= ncols and channels >= Ci zero
  for (int e = threadIdx.x; e < 64 * 8; e += blockDim.x) {
    const int col = e >> 3, c8 = e & 7;
    uint32_t pk[4];
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      unsigned short v2[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int ci = c8 * 8 + h * 2 + u;
        v2[u] = (col < f.ncols && ci < f.Ci) ? __bfloat16_as_ushort(f.wp[(long long)col * f.Ci + ci]) : (unsigned short)0;
      }
      pk[h] = (uint32_t)v2[0] | ((uint32_t)v2[1] << 16);
    }
    *reinterpret_cast<uint4 *>(b_smem + col * 128 + ((c8 ^ (col & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  }
  fence_proxy_async();
  if (warp == 1) tmem_alloc<128>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    for (int lt = 0; lt < my_tiles; ++lt) {
      const int s = lt % FO_STAGES;
      mbar_wait(&empty_bar[s], ((uint32_t)(lt / FO_STAGES) & 1u) ^ 1u);
      if (elect_one()) {
        const int t = (int)blockIdx.x + lt * (int)gridDim.x;
        mbar_expect_tx(&full_bar[s], FO_A_BYTES);
        tma_load_2d(&map_x, &full_bar[s], smem + s * FO_A_BYTES, 0, t * TILE_M);        // rows past the end: zero fill
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_bf16(TILE_M, 64);
    const uint64_t bdesc = make_kmajor_sw128_desc(smem_u32(b_smem));
    for (int lt = 0; lt < my_tiles; ++lt) {
      const int s = lt % FO_STAGES, a = lt & 1;
      mbar_wait(&tmem_empty_bar[a], (((uint32_t)lt >> 1) & 1u) ^ 1u);
      mbar_wait(&full_bar[s], (uint32_t)(lt / FO_STAGES) & 1u);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t adesc = make_kmajor_sw128_desc(smem_u32(smem + s * FO_A_BYTES));
#pragma unroll
        for (int kk = 0; kk < CHUNK_K / UMMA_K; ++kk)
          umma_bf16(tmem_base + (uint32_t)(a * 64), adesc + (uint64_t)(2 * kk), bdesc + (uint64_t)(2 * kk), idesc, kk > 0 ? 1u : 0u);
        umma_commit(&empty_bar[s]);
        umma_commit(&tmem_full_bar[a]);
      }
      __syncwarp();
    }
  } else {
    const int quarter = warp & 3;
    for (int lt = 0; lt < my_tiles; ++lt) {
      const int a = lt & 1;
      const int t = (int)blockIdx.x + lt * (int)gridDim.x;
      const long long q = (long long)t * TILE_M + quarter * 32 + lane;
      mbar_wait(&tmem_full_bar[a], ((uint32_t)lt >> 1) & 1u);
      tc_fence_after();
#pragma unroll 1
      for (int c0 = 0; c0 < 64; c0 += 32) {
        if (c0 >= f.ncols) break;
        uint32_t r[32];
        tmem_ld32(tmem_base + (uint32_t)(a * 64) + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, r);
        tmem_ld_wait();
        if (q < f.n_rows) {
#pragma unroll
          for (int e = 0; e < 32; ++e)
            if (c0 + e < f.ncols) f.zt[(long long)(c0 + e) * f.n_rows + q] = __uint_as_float(r[e]);      // a warp stores 32 consecutive positions
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[a]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<128>(tmem_base);
  }
}

// step 2: one thread per output position (all Co <= 2 channels)
__global__ void __launch_bounds__(256) conv_fewout_gather_kernel(ctagan_conv_geom g, const float *__restrict__ zt, const float *__restrict__ bias,
                                                                 bf16 *__restrict__ y, long long n_rows) {
  const long long total = (long long)g.N * g.Ho * g.Wo;
  const int taps = g.KH * g.KW;
  for (long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(o % g.Wo);
    const long long r = o / g.Wo;
    const int i = (int)(r % g.Ho), n = (int)(r / g.Ho);
    const float *zn = zt + (long long)n * g.Hi * g.Wi;
    for (int co = 0; co < g.Co; ++co) {
      float acc = bias ? bias[co] : 0.f;
      const float *zc = zn + (long long)co * taps * n_rows;
      for (int kh = 0; kh < g.KH; ++kh) {
        const int ih = i + kh - g.pad_h;
        if (ih < 0 || ih >= g.Hi) continue;
        const float *zr = zc + (long long)kh * g.KW * n_rows + (long long)ih * g.Wi + (j - g.pad_w);
#pragma unroll 7
        for (int kw = 0; kw < g.KW; ++kw) {
          const int iw = j + kw - g.pad_w;
          if (iw >= 0 && iw < g.Wi) acc += __ldg(zr + (long long)kw * n_rows + kw);
        }
      }
      y[o * g.Co + co] = __float2bfloat16_rn(apply_act(acc, g.act));
    }
  }
}

bool plan_fewout_tc(const ctagan_conv_geom *g) {
  if (g->dtype != CTAGAN_BF16 || g->dil != 1 || g->stride != 1) return false;
  static int enabled = -1;
  if (enabled < 0) { const char *e = getenv("CTAGAN_THIN_TC"); enabled = e ? atoi(e) : 1; }
  if (!enabled) return false;
  if (g->Co > 2 || g->Ci % 8 || g->Ci < 16 || g->Ci > 64) return false;
  if (g->Co * g->KH * g->KW > 64) return false;
  if ((long long)g->N * g->Ho * g->Wo < 32768) return false;     // small maps (the discriminator's patch head): the CUDA-core kernels are faster
  return true;
}

}  // namespace

int ctagan_conv_fewout_tc_eligible(const ctagan_conv_geom *g) { return plan_fewout_tc(g) ? 1 : 0; }

// scratch: Z^T [Co * taps][N * Hi * Wi] fp32
size_t ctagan_conv_fewout_tc_workspace(const ctagan_conv_geom *g) {
  if (!plan_fewout_tc(g)) return 0;
  return (size_t)g->Co * g->KH * g->KW * g->N * g->Hi * g->Wi * sizeof(float);
}

int ctagan_conv_fewout_tc(const ctagan_conv_geom *g, const void *x, const void *wp, const float *bias, void *y, void *workspace,
                          size_t workspace_bytes, cudaStream_t st) {
  if (!plan_fewout_tc(g)) {
    ctagan_set_error("conv_fewout_tc: geometry not supported");
    return CTAGAN_ERR_UNSUPPORTED;
  }
  const size_t need = ctagan_conv_fewout_tc_workspace(g);
  CTAGAN_REQUIRE(workspace && workspace_bytes >= need, "conv_fewout_tc: workspace of %zu bytes required (got %zu)", need, workspace_bytes);
  CTAGAN_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, "conv_fewout_tc: the input must be 16-byte aligned");
  FewoutZParams f;
  f.Ci = g->Ci; f.ncols = g->Co * g->KH * g->KW;
  f.n_rows = (long long)g->N * g->Hi * g->Wi;
  f.n_tiles = (int)((f.n_rows + TILE_M - 1) / TILE_M);
  f.wp = (const bf16 *)wp; f.zt = (float *)workspace;
  CUtensorMap mx;
  int rc = make_map_2d(&mx, x, (uint64_t)f.n_rows, (uint64_t)g->Ci, TILE_M);
  if (rc) return rc;
  static bool configured = false;
  if (!configured) {
    CTAGAN_CUDA_OK(cudaFuncSetAttribute(conv_fewout_z_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FO_SMEM_BYTES));
    configured = true;
  }
  const int ctas = f.n_tiles < ctagan_num_sms() ? f.n_tiles : ctagan_num_sms();
  CTAGAN_CUDA_OK(launch_pdl(conv_fewout_z_kernel, dim3((unsigned)ctas), dim3(192), (size_t)FO_SMEM_BYTES, st, mx, f));
  CTAGAN_LAUNCH_OK();
  const long long total = (long long)g->N * g->Ho * g->Wo;
  long long blocks = (total + 255) / 256;
  if (blocks > 16LL * ctagan_num_sms()) blocks = 16LL * ctagan_num_sms();
  conv_fewout_gather_kernel<<<(int)blocks, 256, 0, st>>>(*g, (const float *)workspace, bias, (bf16 *)y, f.n_rows);
  CTAGAN_LAUNCH_OK();
  return CTAGAN_OK;
}
