// common.cu — error plumbing, launch counter, device queries.
#include "common.cuh"

#include <cstdlib>
#include <mutex>

namespace ddrl {

std::string& last_error_ref() {
  static thread_local std::string s;
  return s;
}

int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  last_error_ref() = buf;
  return code;
}

std::atomic<int64_t> g_launches{0};
bool g_use_pdl = [] { const char* e = getenv("DDRL_PDL"); return e && e[0] == '1'; }();   // measured slower inside CUDA graphs on B200 (146 vs 130 us at C2): opt-in

int sm_count(int device) {
  static std::mutex mu;
  static int cache[64];
  static bool have[64];
  std::lock_guard<std::mutex> lk(mu);
  if (device >= 0 && device < 64 && have[device]) return cache[device];
  int n = 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || n <= 0)
    n = 148;
  if (device >= 0 && device < 64) { cache[device] = n; have[device] = true; }
  return n;
}

void* host_device_pointer(const void* host_ptr) {
  cudaPointerAttributes at{};
  if (cudaPointerGetAttributes(&at, host_ptr) != cudaSuccess) { (void)cudaGetLastError(); return nullptr; }
  if (at.type != cudaMemoryTypeHost || !at.devicePointer) return nullptr;
  return at.devicePointer;
}

}  // namespace ddrl

extern "C" {
int ddrl_abi_version(void) { return DDRL_ABI_VERSION; }
const char* ddrl_last_error(void) { return ddrl::last_error_ref().c_str(); }
int64_t ddrl_launch_count(void) { return ddrl::g_launches.load(std::memory_order_relaxed); }
}
