// Shared device/host helpers for the hnet_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/hnet_b200.h"

namespace hnb {

// ---------------------------------------------------------------------------------------------
// error reporting + launch accounting (C-ABI: int status + thread-local message)
// ---------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define HNB_CHECK_ARG(cond, ...)                 \
  do {                                           \
    if (!(cond)) {                               \
      hnb::set_error(__VA_ARGS__);               \
      return HNB_ERR_INVALID_ARG;                \
    }                                            \
  } while (0)

#define HNB_LAUNCH_CHECK(name)                                                      \
  do {                                                                              \
    cudaError_t e__ = cudaGetLastError();                                           \
    if (e__ != cudaSuccess) {                                                       \
      hnb::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));       \
      return HNB_ERR_CUDA;                                                          \
    }                                                                               \
    hnb::count_launch();                                                            \
  } while (0)

#define HNB_CUDA_CALL(expr)                                                         \
  do {                                                                              \
    cudaError_t e__ = (expr);                                                       \
    if (e__ != cudaSuccess) {                                                       \
      hnb::set_error("%s failed: %s", #expr, cudaGetErrorString(e__));              \
      return HNB_ERR_CUDA;                                                          \
    }                                                                               \
  } while (0)

// dtype dispatch: T is the storage type of activations (all math is fp32)
#define HNB_DISPATCH_DTYPE(dt, T, ...)                                  \
  do {                                                                  \
    if ((dt) == HNB_F32) {                                              \
      using T = float;                                                  \
      __VA_ARGS__;                                                      \
    } else if ((dt) == HNB_BF16) {                                      \
      using T = __nv_bfloat16;                                          \
      __VA_ARGS__;                                                      \
    } else {                                                            \
      hnb::set_error("unsupported dtype %d", (int)(dt));                \
      return HNB_ERR_INVALID_ARG;                                       \
    }                                                                   \
  } while (0)

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// Division by a run-time constant as one multiply-high: exact while n * d < 2^32 (item counts and head / chunk /
// batch counts here are far below that).  The kernels decode an item index per loop iteration in every thread; the
// hardware has no integer divide and the emulation costs ~40 instructions per division.
struct FastDiv {
  uint32_t d, m;
  FastDiv() : d(1), m(0) {}
  explicit FastDiv(int div) : d((uint32_t)div), m(div > 1 ? (uint32_t)((0x100000000ULL + (uint32_t)div - 1) / (uint32_t)div) : 0u) {}
  __device__ __forceinline__ int div(int n) const { return d == 1 ? n : (int)__umulhi((uint32_t)n, m); }
  __device__ __forceinline__ void divmod(int n, int& q, int& r) const { q = div(n); r = n - q * (int)d; }
};

// ---------------------------------------------------------------------------------------------
// scalar conversions
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// ---------------------------------------------------------------------------------------------
// 16-byte vector access: Vec<T> carries 16 bytes worth of T (4 floats / 8 bf16)
// ---------------------------------------------------------------------------------------------
template <typename T> struct Vec;
template <> struct Vec<float> {
  static constexpr int N = 4;
  float v[4];
  __device__ __forceinline__ void load(const float* p) {
    float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  __device__ __forceinline__ void store(float* p) const {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <> struct Vec<__nv_bfloat16> {
  static constexpr int N = 8;
  float v[8];
  __device__ __forceinline__ void load(const __nv_bfloat16* p) {
    uint4 t = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 f = __bfloat1622float2(h[i]);
      v[2 * i] = f.x; v[2 * i + 1] = f.y;
    }
  }
  __device__ __forceinline__ void store(__nv_bfloat16* p) const {
    uint4 t;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = t;
  }
};

// generic small-vector load/store of VN elements (VN*sizeof(T) in {4,8,16} bytes), fp32 in registers
template <typename T, int VN> __device__ __forceinline__ void ldv(const T* p, float* r) {
#pragma unroll
  for (int i = 0; i < VN; ++i) r[i] = to_f(p[i]);
}
template <> __device__ __forceinline__ void ldv<float, 4>(const float* p, float* r) {
  float4 t = *reinterpret_cast<const float4*>(p); r[0] = t.x; r[1] = t.y; r[2] = t.z; r[3] = t.w;
}
template <> __device__ __forceinline__ void ldv<float, 8>(const float* p, float* r) {
  float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  r[0] = a.x; r[1] = a.y; r[2] = a.z; r[3] = a.w; r[4] = b.x; r[5] = b.y; r[6] = b.z; r[7] = b.w;
}
template <> __device__ __forceinline__ void ldv<float, 2>(const float* p, float* r) {
  float2 t = *reinterpret_cast<const float2*>(p); r[0] = t.x; r[1] = t.y;
}
template <> __device__ __forceinline__ void ldv<__nv_bfloat16, 2>(const __nv_bfloat16* p, float* r) {
  float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p)); r[0] = f.x; r[1] = f.y;
}
template <> __device__ __forceinline__ void ldv<__nv_bfloat16, 4>(const __nv_bfloat16* p, float* r) {
  uint2 t = *reinterpret_cast<const uint2*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
  float2 a = __bfloat1622float2(h[0]), b = __bfloat1622float2(h[1]);
  r[0] = a.x; r[1] = a.y; r[2] = b.x; r[3] = b.y;
}
template <> __device__ __forceinline__ void ldv<__nv_bfloat16, 8>(const __nv_bfloat16* p, float* r) {
  Vec<__nv_bfloat16> t; t.load(p);
#pragma unroll
  for (int i = 0; i < 8; ++i) r[i] = t.v[i];
}
template <typename T, int VN> __device__ __forceinline__ void stv(T* p, const float* r) {
#pragma unroll
  for (int i = 0; i < VN; ++i) p[i] = from_f<T>(r[i]);
}
template <> __device__ __forceinline__ void stv<float, 4>(float* p, const float* r) {
  *reinterpret_cast<float4*>(p) = make_float4(r[0], r[1], r[2], r[3]);
}
template <> __device__ __forceinline__ void stv<float, 2>(float* p, const float* r) {
  *reinterpret_cast<float2*>(p) = make_float2(r[0], r[1]);
}
template <> __device__ __forceinline__ void stv<__nv_bfloat16, 2>(__nv_bfloat16* p, const float* r) {
  *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(r[0], r[1]);
}
template <> __device__ __forceinline__ void stv<__nv_bfloat16, 4>(__nv_bfloat16* p, const float* r) {
  uint2 t; __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
  h[0] = __floats2bfloat162_rn(r[0], r[1]); h[1] = __floats2bfloat162_rn(r[2], r[3]);
  *reinterpret_cast<uint2*>(p) = t;
}
template <> __device__ __forceinline__ void stv<__nv_bfloat16, 8>(__nv_bfloat16* p, const float* r) {
  Vec<__nv_bfloat16> t;
#pragma unroll
  for (int i = 0; i < 8; ++i) t.v[i] = r[i];
  t.store(p);
}

// ---------------------------------------------------------------------------------------------
// warp / block reductions
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum; every thread gets the result. `red` is >= 32 floats of shared memory.
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float t = (lane < nw) ? red[lane] : 0.f;
  t = warp_sum(t);
  return t;
}

// fast reciprocal (MUFU.RCP, ~1 ulp): the IEEE division costs ~8 extra instructions per activation
__device__ __forceinline__ float silu_f(float x) { return __fdividef(x, 1.f + __expf(-x)); }
__device__ __forceinline__ float sigmoid_f(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float softplus_f(float x) { return x > 20.f ? x : log1pf(__expf(x)); }

// scan position s of direction `dir` -> natural time index.  dir 0: identity.  dir 1: the valid span
// [0,len) is traversed back to front and the right padding stays in place, which is exactly what
// reverse_sequences() does (reference src/dcasr/models/mamba_block.py:19-28).
__device__ __forceinline__ int scan_to_nat(int dir, int s, int len) {
  return (dir == 1 && s < len) ? (len - 1 - s) : s;
}

// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch.  An encoder step is ~375 launches of 5-50 us kernels on one stream: the launch
// latency and the drain / ramp between two of them are a measurable share of the step.  Kernels launched through
// launch_pdl() may become resident while their predecessor's last CTAs are still running; everything they do before
// pdl_wait() (barrier init, TMEM allocation, tensor-map prefetch, shared-memory fills) overlaps that tail.
// RULE: a kernel launched through launch_pdl() executes pdl_wait() in EVERY thread before its first global-memory
// access (read or write: the predecessor may still be reading what this kernel overwrites).  pdl_trigger() only allows
// the successor to be scheduled early; the successor's own pdl_wait() still waits for this grid to complete and flush.
// Without the launch attribute both instructions are no-ops.  HNB_PDL=0 switches the attribute off.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_enter() { pdl_trigger(); pdl_wait(); }
bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute at[1];
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = at; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---------------------------------------------------------------------------------------------
// host: per-launch driver queries cached (the host needs ~13 ms to enqueue one encoder step of ~650 launches, so a
// few microseconds per launch in cudaFuncSetAttribute / occupancy queries are a measurable share of it)
// ---------------------------------------------------------------------------------------------
inline cudaError_t hnb_set_max_smem(const void* fn, int bytes) {
  struct E { const void* fn; int dev; int bytes; };
  static E tab[128];
  static int n = 0;
  int dev = 0;
  cudaGetDevice(&dev);
  for (int i = 0; i < n; ++i)
    if (tab[i].fn == fn && tab[i].dev == dev && tab[i].bytes >= bytes) return cudaSuccess;
  const cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess && n < 128) tab[n++] = E{fn, dev, bytes};
  return e;
}

}  // namespace hnb
