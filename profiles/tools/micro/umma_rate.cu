// Developer microbenchmark: issue rate of tcgen05.mma (SS mode, bf16 -> fp32, M = 128) for K-major and MN-major operands, with and
// without concurrent shared-memory write traffic from other warps.  Operands are whatever is in shared memory (timing only).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_rate umma_rate.cu && ./umma_rate
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ uint64_t desc_k(uint32_t a) {
  return (uint64_t)((a & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ uint64_t desc_mn(uint32_t a, uint32_t lbo) {
  return (uint64_t)((a & 0x3FFFF) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// mode 0: K-major (stage = [A 128 rows x 128 B][B BN rows x 128 B], 4 MMAs of K=16 per 64-wide chunk)
// mode 1: MN-major (stage = A 2 slabs of [64 k rows][128 B], B BN/64 slabs; 4 MMAs of K=16 per 64-pixel chunk)
template <int BN>
__global__ void __launch_bounds__(192, 1) rate_kernel(int mode, int n_chunks, int writers, long long *out) {
  extern __shared__ uint8_t raw[];
  uint8_t *smem = (uint8_t *)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int STAGE = (128 + BN) * 128;
  constexpr int NST = (180 * 1024) / STAGE;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); stop = 0; asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "n"(BN < 32 ? 32 : BN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = slot;
  if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24) | (mode ? ((1u << 15) | (1u << 16)) : 0u);
      const long long t0 = clock64();
      for (int c = 0; c < n_chunks; ++c) {
        const uint32_t a = smem_u32(smem + (c % NST) * STAGE), b = a + 128 * 128;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (mode == 0) umma(tm, desc_k(a) + 2 * k, desc_k(b) + 2 * k, idesc, 1);
          else umma(tm, desc_mn(a + k * 2048, 8192), desc_mn(b + k * 2048, 8192), idesc, 1);
        }
      }
      const long long t1 = clock64();
      commit(&bar);
      while (!mbar_try_wait(&bar, 0)) {}
      const long long t2 = clock64();
      stop = 1;
      if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    __syncwarp();
  } else if (warp >= 2 && warp - 2 < writers) {
    // synthetic "TMA" pressure: 16-byte shared-memory stores at full rate into the upper part of the buffer
    uint4 *dst = (uint4 *)(smem + NST * STAGE) ;
    uint4 v = make_uint4(1, 2, 3, 4);
    int i = 0;
    while (!stop) {
#pragma unroll
      for (int u = 0; u < 16; ++u) dst[((warp - 2) * 32 + lane + (i + u) * 128) & 511] = v;
      i += 16;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "n"(BN < 32 ? 32 : BN) : "memory");
  }
}

template <int BN>
void run(int mode, int writers, int grid, long long *d_out) {
  const int n_chunks = 512;
  cudaFuncSetAttribute(rate_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  rate_kernel<BN><<<grid, 192, 200 * 1024>>>(mode, n_chunks, writers, d_out);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2];
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  printf("BN=%3d %s writers=%d grid=%3d: issue %6.1f clk/MMA, complete %6.1f clk/MMA (floor %d)  %s\n", BN, mode ? "MN-major" : "K-major ", writers, grid,
         (double)h[0] / (n_chunks * 4), (double)h[1] / (n_chunks * 4), 128 * BN / 256, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  long long *d_out;
  cudaMalloc(&d_out, 16);
  for (int grid : {1, 148})
    for (int mode : {0, 1})
      for (int writers : {0, 4}) {
        run<64>(mode, writers, grid, d_out);
        run<128>(mode, writers, grid, d_out);
        run<256>(mode, writers, grid, d_out);
      }
  return 0;
}
