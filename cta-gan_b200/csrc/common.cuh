// Shared device/host helpers for libctagan (sm_100a only).
#pragma once
#include <utility>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "../../include/ctagan.h"

typedef __nv_bfloat16 bf16;

void ctagan_set_error(const char *fmt, ...);

#define CTAGAN_REQUIRE(cond, ...)                      \
  do {                                                 \
    if (!(cond)) {                                     \
      ctagan_set_error(__VA_ARGS__);                   \
      return CTAGAN_ERR_ARG;                           \
    }                                                  \
  } while (0)

#define CTAGAN_CUDA_OK(expr)                                                              \
  do {                                                                                    \
    cudaError_t e_ = (expr);                                                              \
    if (e_ != cudaSuccess) {                                                              \
      ctagan_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
      return CTAGAN_ERR_CUDA;                                                             \
    }                                                                                     \
  } while (0)

#define CTAGAN_LAUNCH_OK()                                                                 \
  do {                                                                                     \
    cudaError_t e_ = cudaGetLastError();                                                   \
    if (e_ != cudaSuccess) {                                                               \
      ctagan_set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e_), __FILE__, __LINE__); \
      return CTAGAN_ERR_CUDA;                                                              \
    }                                                                                      \
  } while (0)

// dispatch on the runtime activation dtype
#define CTAGAN_DISPATCH_DTYPE(dtype, T, ...)                               \
  do {                                                                     \
    if ((dtype) == CTAGAN_F32) { typedef float T; __VA_ARGS__; }           \
    else if ((dtype) == CTAGAN_BF16) { typedef bf16 T; __VA_ARGS__; }      \
    else { ctagan_set_error("bad dtype %d", (int)(dtype)); return CTAGAN_ERR_ARG; } \
  } while (0)

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

// Vector of V elements of T moved with one 16-byte (or smaller) access.
template <typename T, int V> struct Vec {
  T v[V];
};

template <typename T, int V>
__device__ __forceinline__ void load_vec(const T *p, float (&out)[V]) {
  if constexpr (V == 1) {
    out[0] = to_f(p[0]);
  } else if constexpr (sizeof(T) * V == 16) {
    uint4 raw = *reinterpret_cast<const uint4 *>(p);
    const T *e = reinterpret_cast<const T *>(&raw);
#pragma unroll
    for (int i = 0; i < V; ++i) out[i] = to_f(e[i]);
  } else if constexpr (sizeof(T) * V == 8) {
    uint2 raw = *reinterpret_cast<const uint2 *>(p);
    const T *e = reinterpret_cast<const T *>(&raw);
#pragma unroll
    for (int i = 0; i < V; ++i) out[i] = to_f(e[i]);
  } else {
#pragma unroll
    for (int i = 0; i < V; ++i) out[i] = to_f(p[i]);
  }
}

template <typename T, int V>
__device__ __forceinline__ void store_vec(T *p, const float (&in)[V]) {
  if constexpr (V == 1) {
    p[0] = from_f<T>(in[0]);
  } else if constexpr (sizeof(T) * V == 16) {
    uint4 raw;
    T *e = reinterpret_cast<T *>(&raw);
#pragma unroll
    for (int i = 0; i < V; ++i) e[i] = from_f<T>(in[i]);
    *reinterpret_cast<uint4 *>(p) = raw;
  } else if constexpr (sizeof(T) * V == 8) {
    uint2 raw;
    T *e = reinterpret_cast<T *>(&raw);
#pragma unroll
    for (int i = 0; i < V; ++i) e[i] = from_f<T>(in[i]);
    *reinterpret_cast<uint2 *>(p) = raw;
  } else {
#pragma unroll
    for (int i = 0; i < V; ++i) p[i] = from_f<T>(in[i]);
  }
}

// widest vector (in elements) that keeps 16-byte accesses legal for C channels of T
template <typename T> __host__ __device__ constexpr int max_vec() { return 16 / (int)sizeof(T); }

__device__ __forceinline__ float apply_act(float v, int act) {
  switch (act) {
    case CTAGAN_ACT_RELU: return v > 0.f ? v : 0.f;
    case CTAGAN_ACT_LRELU: return v > 0.f ? v : 0.2f * v;
    case CTAGAN_ACT_TANH: return tanhf(v);
    default: return v;
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------------------------------------------------------------
// Programmatic dependent launch (griddepcontrol).  At batch 1 a step is ~1000 dependent kernels of a few microseconds each, so the
// launch gap between a kernel and its consumer is a large share of the chain.  Hot kernels therefore (a) signal
// launch_dependents at their very top, so the consumer's CTAs are scheduled and run their prologue (barrier init, TMEM allocation,
// index arithmetic) while this grid is still working, and (b) execute griddepcontrol.wait before their first global-memory access,
// which blocks until the producer grid has completed and flushed.  Both instructions are no-ops for a normally launched kernel.
// RULE: a kernel launched through launch_pdl() must call pdl_wait() before touching global memory.
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

bool ctagan_pdl_enabled();     // env CTAGAN_PDL=1 (opt-in: measured neutral on the Cyc step, see capi.cu)

template <typename... KArgs, typename... Args>
inline cudaError_t launch_cluster_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, unsigned cluster_x,
                                      Args &&...args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = ctagan_pdl_enabled() ? 1 : 0;
  cfg.numAttrs = 1;
  if (cluster_x > 1) {                      // thread-block cluster along x (CTA pairs of the cta_group::2 kernels)
    attr[1].id = cudaLaunchAttributeClusterDimension;
    attr[1].val.clusterDim.x = cluster_x;
    attr[1].val.clusterDim.y = 1;
    attr[1].val.clusterDim.z = 1;
    cfg.numAttrs = 2;
  }
  cfg.attrs = attr;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args &&...args) {
  return launch_cluster_pdl(kernel, grid, block, smem, st, 1u, std::forward<Args>(args)...);
}

// number of SMs (B200: 148); cached
int ctagan_num_sms();

// out[i] = (accumulate ? out[i] : 0) + part[0][i] + part[1][i] + ... + part[parts-1][i]  (i < n), added in row order: the deterministic
// second half of every split reduction of the library (each CTA stores its partial result in its own row; no floating-point atomics)
int ctagan_ordered_sum(const float *part, float *out, int parts, long long n, cudaStream_t st, int accumulate = 0);
// many rows of [out1 (n1) | out2 (n2)] -> out1, out2: one warp per element (fixed order: lane-strided rows, then a shuffle tree)
int ctagan_ordered_sum_rows2(const float *part, float *out1, long long n1, float *out2, long long n2, int parts, long long row_stride,
                             cudaStream_t st, int accumulate = 0);
// the same with rows `row_stride` floats apart (partial rows that hold more than one result, e.g. [dw | db])
int ctagan_ordered_sum_strided(const float *part, float *out, int parts, long long n, long long row_stride, cudaStream_t st, int accumulate = 0);
