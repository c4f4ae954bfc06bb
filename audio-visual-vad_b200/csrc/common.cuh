// Shared host/device helpers for libavvad (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>
#include <mutex>
#include <string>

#include "../../include/avvad.h"

namespace avvad {

// ---- error plumbing ---------------------------------------------------------------------------
void set_error(const std::string& msg);
extern std::atomic<uint64_t> g_launches;

#define AVVAD_CHECK_ARG(cond, msg)                                         \
  do {                                                                     \
    if (!(cond)) {                                                         \
      ::avvad::set_error(std::string("bad argument: ") + (msg));           \
      return AVVAD_ERR_ARG;                                                \
    }                                                                      \
  } while (0)

#define AVVAD_CUDA(call)                                                                      \
  do {                                                                                        \
    cudaError_t e__ = (call);                                                                 \
    if (e__ != cudaSuccess) {                                                                 \
      ::avvad::set_error(std::string(#call) + ": " + cudaGetErrorString(e__));                \
      return AVVAD_ERR_CUDA;                                                                  \
    }                                                                                         \
  } while (0)

// Call after every kernel launch: counts it and surfaces launch-configuration errors.  With AVVAD_SYNC_DEBUG=1 in the
// environment every launch is followed by a device synchronisation, so an asynchronous fault (illegal address, ...) is
// reported with the file:line of the kernel that caused it (debugging aid; never set in production or benchmarks).
bool sync_debug();
#define AVVAD_LAUNCHED()                                                                      \
  do {                                                                                        \
    ::avvad::g_launches.fetch_add(1, std::memory_order_relaxed);                              \
    cudaError_t e__ = cudaGetLastError();                                                     \
    if (e__ == cudaSuccess && ::avvad::sync_debug()) e__ = cudaDeviceSynchronize();           \
    if (e__ != cudaSuccess) {                                                                 \
      ::avvad::set_error(std::string("kernel launch failed: ") + cudaGetErrorString(e__) +    \
                         " at " + __FILE__ + ":" + std::to_string(__LINE__));                 \
      return AVVAD_ERR_CUDA;                                                                  \
    }                                                                                         \
  } while (0)

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t a, size_t b) { return (a + b - 1) / b * b; }

// ---- small device helpers ---------------------------------------------------------------------
__device__ __forceinline__ float sigmoidf_fast(float x) { return 1.0f / (1.0f + __expf(-x)); }
__device__ __forceinline__ float tanhf_fast(float x) {
  // 2*sigmoid(2x)-1, clamped so __expf never overflows to inf/inf
  float e = __expf(-2.0f * fminf(fmaxf(x, -15.0f), 15.0f));
  return (1.0f - e) / (1.0f + e);
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
// 256-bit global accesses (sm_100: LDG/STG.E.ENL2.256): one full 32-byte sector per thread and instruction; the
// address must be 32-byte aligned
struct alignas(32) u32x8 {
  uint32_t v[8];
};
__device__ __forceinline__ void st_global_256(void* p, const u32x8& x) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(x.v[0]), "r"(x.v[1]), "r"(x.v[2]),
               "r"(x.v[3]), "r"(x.v[4]), "r"(x.v[5]), "r"(x.v[6]), "r"(x.v[7])
               : "memory");
}
__device__ __forceinline__ u32x8 ld_global_256(const void* p) {
  u32x8 x;
  asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(x.v[0]), "=r"(x.v[1]), "=r"(x.v[2]), "=r"(x.v[3]), "=r"(x.v[4]), "=r"(x.v[5]), "=r"(x.v[6]),
                 "=r"(x.v[7])
               : "l"(p));
  return x;
}

// {bf16(max(lo,0)), bf16(max(hi,0))} in one instruction (F2FP.RELU)
__device__ __forceinline__ uint32_t pack_relu_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}

// One-time initialisation PER DEVICE.  Kernel attributes (cudaFuncSetAttribute) and device-resident lookup tables belong
// to one device's context, and nn.DataParallel -- what the reference's training scripts use -- drives several devices
// from the threads of one process: a process-wide std::once_flag would leave every device but the first unconfigured.
constexpr int kMaxDevices = 64;
struct PerDeviceOnce {
  std::mutex mu;
  uint64_t done = 0;
  cudaError_t err[kMaxDevices] = {};
  template <class F>
  cudaError_t run(F&& f) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return cudaErrorInvalidDevice;
    std::lock_guard<std::mutex> lk(mu);
    if (!((done >> dev) & 1ull)) {
      err[dev] = f();
      done |= 1ull << dev;
    }
    return err[dev];
  }
};

}  // namespace avvad
