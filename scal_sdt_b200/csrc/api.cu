// Library-wide state of libsdt_b200.so: thread-local error message, device queries.
#include "sdt_common.cuh"

#include <atomic>
#include <stdlib.h>
#include <mutex>

namespace sdt {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
long long launch_count() { return g_launches.load(std::memory_order_relaxed); }

uint64_t debug_get(int key);   // lora_wgrad.cu

bool pdl_enabled() {
  static int env = -1;
  if (env < 0) {
    const char* e = getenv("SDT_PDL");
    env = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  const uint64_t dbg = debug_get(24);      // 1: off, 2: on (overrides the environment; A/B inside one process)
  if (dbg == 1) return false;
  if (dbg == 2) return true;
  return env != 0;
}

int num_sms() {
  static int cached[64];
  static std::mutex mu;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  std::lock_guard<std::mutex> lk(mu);
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

}  // namespace sdt

extern "C" int sdt_version(void) { return 100; }  // round 1

extern "C" long long sdt_launch_count(void) { return sdt::launch_count(); }

extern "C" const char* sdt_last_error(void) { return sdt::g_err; }

extern "C" int sdt_device_check(void) {
  int dev = 0, major = 0, minor = 0;
  SDT_CUDA_OK(cudaGetDevice(&dev));
  SDT_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  SDT_CUDA_OK(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  SDT_REQUIRE(major == 10, SDT_ERR_UNSUPPORTED,
              "libsdt_b200 is built for sm_100a only; device %d is sm_%d%d (there is no fallback path)", dev, major, minor);
  return SDT_OK;
}
