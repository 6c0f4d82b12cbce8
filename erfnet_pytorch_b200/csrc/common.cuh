// Shared helpers for the sm_100a kernels behind include/erfnet_b200.h.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../include/erfnet_b200.h"

namespace erfk {

void set_error(const char* fmt, ...);

#define ERF_CHECK_ARG(cond, ...)          \
  do {                                    \
    if (!(cond)) {                        \
      erfk::set_error(__VA_ARGS__);        \
      return ERF_ERR_INVALID_ARGUMENT;    \
    }                                     \
  } while (0)

// Launch errors only (never synchronises): the library enqueues on the caller's stream.
#define ERF_CHECK_LAUNCH(name)                                                    \
  do {                                                                            \
    cudaError_t e__ = cudaGetLastError();                                         \
    if (e__ != cudaSuccess) {                                                     \
      erfk::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));     \
      return ERF_ERR_CUDA;                                                        \
    }                                                                             \
  } while (0)

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// Load / store N consecutive elements (N*sizeof(T) must be 8 or 16 bytes and the address so aligned).
template <typename T, int N> struct Vec;
template <> struct Vec<float, 4> {
  float v[4];
  __device__ __forceinline__ void load(const float* p) { float4 t = *reinterpret_cast<const float4*>(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
  __device__ __forceinline__ void store(float* p) const { *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); }
};
template <> struct Vec<__nv_bfloat16, 4> {
  float v[4];
  __device__ __forceinline__ void load(const __nv_bfloat16* p) {
    uint2 t = *reinterpret_cast<const uint2*>(p);
    __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&t.x), b = *reinterpret_cast<__nv_bfloat162*>(&t.y);
    v[0] = __low2float(a); v[1] = __high2float(a); v[2] = __low2float(b); v[3] = __high2float(b);
  }
  __device__ __forceinline__ void store(__nv_bfloat16* p) const {
    __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
    uint2 t; t.x = *reinterpret_cast<uint32_t*>(&a); t.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = t;
  }
};
template <> struct Vec<float, 8> {
  float v[8];
  __device__ __forceinline__ void load(const float* p) {
    float4 a = reinterpret_cast<const float4*>(p)[0], b = reinterpret_cast<const float4*>(p)[1];
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  __device__ __forceinline__ void store(float* p) const {
    reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
    reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
  }
};
template <> struct Vec<__nv_bfloat16, 8> {
  float v[8];
  __device__ __forceinline__ void load(const __nv_bfloat16* p) {
    uint4 t = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[2 * i] = __low2float(h[i]); v[2 * i + 1] = __high2float(h[i]); }
  }
  __device__ __forceinline__ void store(__nv_bfloat16* p) const {
    uint4 t;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = t;
  }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- programmatic dependent launch (PDL) ------------------------------------------------------------------------------------
// A training step is ~600 kernels of 5-40 us: the 2-3 us between the end of one grid and the start of the next (grid launch latency,
// plus barrier / TMEM / descriptor set-up for the tcgen05 kernels) is ~20 % of the step.  Every hot-path kernel is launched with the
// programmatic-stream-serialization attribute and begins with pdl_trigger() (dependents may be scheduled) ... pdl_wait() (the
// producer grid has completed and its memory is visible): whatever precedes pdl_wait() overlaps the previous kernel's tail.
// RULE: a kernel launched through launch_pdl() must execute pdl_wait() in every thread before it reads or writes global memory that an
// earlier kernel of the stream may have written or read.  ERF_PDL=0 in the environment falls back to ordinary launches.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_enter() { pdl_trigger(); pdl_wait(); }

static inline bool pdl_enabled() {
  static const bool on = !(getenv("ERF_PDL") && atoi(getenv("ERF_PDL")) == 0);
  return on;
}

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---- grid-wide barrier for kernels whose WHOLE grid is co-resident (the launcher checks grid <= SMs x occupancy) ------------------------
// `counter` is a zeroed 32-bit word in global memory used by exactly one barrier of one launch (the trainer hands out slots of a scratch
// arena it zeroes with one memset per step).  Pattern of cooperative groups' grid.sync(): block barrier, one thread publishes the block's
// global writes (fence) and arrives, spins with acquire loads, block barrier.  Data written by other blocks before the barrier must be
// read with L1-bypassing loads afterwards (__ldcg) -- the lines may sit stale in this SM's L1.
// Deadlock freedom with programmatic dependent launch: a dependent grid is only scheduled once every block of this grid has executed
// pdl_trigger(), i.e. is resident, and the preceding grid never waits for this one.
__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned expected) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    unsigned v;
    for (;;) {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
      if (v >= expected) break;
      __nanosleep(40);
    }
  }
  __syncthreads();
}

// Kernels meant to run several CTAs per SM ask for the largest shared-memory carveout: the driver's default picks the smallest configuration
// that fits ONE block whenever another resource already limits the occupancy estimate (ncu: 102 KB configured for a 94 KB block).
template <typename K>
static inline cudaError_t prefer_max_smem(K kernel) {
  return cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
}

// Launch cap (erf_set_launch_cap): while weight gradients run on a second stream next to the data gradients, both kinds of kernel take only
// part of the SMs so that they really run side by side (persistent kernels sized one CTA per SM would just queue behind each other).
int launch_cap_main();   // CTAs per SM-slot for kernels on the main chain (0 = no cap)
int launch_cap_side();   // ... for the weight-gradient kernels

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

}  // namespace erfk
