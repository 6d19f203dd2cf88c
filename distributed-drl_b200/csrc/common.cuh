// common.cuh — shared host/device helpers for libddrl_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <string>

#include "ddrl_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libddrl_b200 is written for sm_100a (B200) only"
#endif

namespace ddrl {

// ---------------------------------------------------------------------------------------------
// host-side error plumbing
// ---------------------------------------------------------------------------------------------
std::string& last_error_ref();
int fail(int code, const char* fmt, ...);
extern std::atomic<int64_t> g_launches;

#define DDRL_CUDA(expr)                                                                        \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess)                                                                     \
      return ::ddrl::fail(DDRL_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),  \
                          __FILE__, __LINE__);                                                 \
  } while (0)

#define DDRL_LAUNCH_CHECK()                                                                    \
  do {                                                                                         \
    ::ddrl::g_launches.fetch_add(1, std::memory_order_relaxed);                                \
    cudaError_t _e = cudaGetLastError();                                                       \
    if (_e != cudaSuccess)                                                                     \
      return ::ddrl::fail(DDRL_ECUDA, "kernel launch failed: %s (%s:%d)",                      \
                          cudaGetErrorString(_e), __FILE__, __LINE__);                         \
  } while (0)

struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    // every entry point opens with a guard: drop a stale per-thread "last error" left by other runtime users of this host
    // thread (PyTorch probes that fail benignly), so that DDRL_LAUNCH_CHECK only ever reports our own launches
    (void)cudaGetLastError();
    if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
    if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

int sm_count(int device);
// device-side address of a host pointer when it is pinned, mapped host memory the current device can access directly
// (cudaHostAlloc / cudaHostRegister under unified addressing), else nullptr
void* host_device_pointer(const void* host_ptr);

#ifdef __CUDACC__
extern bool g_use_pdl;   // DDRL_PDL=1 turns programmatic dependent launch on (off by default: slower inside graphs)
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = g_use_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__

// Programmatic dependent launch: a kernel launched with launch_pdl() may start while its stream predecessor is
// still running.  pdl_wait() blocks until the predecessor grid has completed and its writes are visible (a no-op
// for a normal launch); EVERY kernel of a chain calls it before touching global memory, so completion is
// transitive.  pdl_trigger() lets the successor's CTAs be scheduled as soon as this grid's CTAs are all resident.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// 128-bit streaming load that does not allocate in L1 (rows of a random gather are not re-used)
__device__ __forceinline__ float4 ld_nc_f4(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_f4(float4* p, const float4& v) {
  asm volatile("st.global.v4.f32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// Philox4x32-10 (Random123; oracle/replay_oracle.py: philox4x32_10 restates the same rounds)
struct Philox4 { uint32_t x, y, z, w; };
__device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                 uint32_t k0, uint32_t k1) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += W0; k1 += W1;
  }
  return Philox4{c0, c1, c2, c3};
}

// index of sample ordinal i: uniform on [0,size) by multiply-shift of a 64-bit draw
__device__ __forceinline__ int64_t philox_index(uint64_t ordinal, uint64_t seed, uint64_t counter,
                                                uint32_t rng_stream, uint64_t size) {
  const Philox4 p = philox4x32_10((uint32_t)ordinal, (uint32_t)counter, (uint32_t)(counter >> 32),
                                  rng_stream, (uint32_t)seed, (uint32_t)(seed >> 32));
  const uint64_t u = ((uint64_t)p.y << 32) | (uint64_t)p.x;
  return (int64_t)__umul64hi(u, size);
}

#endif  // __CUDACC__

}  // namespace ddrl
