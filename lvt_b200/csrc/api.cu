// lvt_b200 :: library-level C-ABI entry points (error string, device check, launch counter).
#include <atomic>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../include/lvt_b200.h"
#include "common.cuh"

static thread_local char g_last_error[1024] = "";
static std::atomic<long long> g_launches{0};

void lvt_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
}

bool lvt_pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    // measured on B200 (DSFVT step, CUDA-graph replay): trigger at kernel start 12.08 ms vs 11.85 ms without PDL;
    // trigger at the last tile of the persistent GEMMs 10.59 ms vs 10.59 ms without -> no gain, off by default
    const char* e = getenv("LVT_PDL");
    on = (e && e[0] == '1') ? 1 : 0;
  }
  return on == 1;
}

void lvt_count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// SM budget of the persistent tensor-core kernels (0 = every SM).  While a gradient bucket is being all-reduced the
// collective's CTAs hold a few SMs for its whole duration; a persistent grid sized for ALL SMs would then run its
// last CTAs as a second wave (twice the kernel time), so the overlapped part of the backward is launched -- and
// captured into its CUDA graphs -- with that many SMs fewer.
static std::atomic<int> g_sm_limit{0};
int lvt_sm_limit() { return g_sm_limit.load(std::memory_order_relaxed); }
extern "C" void lvt_set_sm_limit(int n) { g_sm_limit.store(n > 0 ? n : 0, std::memory_order_relaxed); }

extern "C" int lvt_abi_version(void) { return LVT_B200_ABI_VERSION; }

extern "C" const char* lvt_last_error(void) { return g_last_error; }

extern "C" long long lvt_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" void lvt_launch_count_reset(void) { g_launches.store(0, std::memory_order_relaxed); }

extern "C" int lvt_device_check(void) {
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    lvt_set_error("no CUDA device: %s", cudaGetErrorString(e));
    cudaGetLastError();
    return LVT_ERR_NO_DEVICE;
  }
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (major != 10) {
    lvt_set_error("device %d is sm_%d%d; lvt_b200 is built for sm_100a only and has no fallback",
                  dev, major, minor);
    return LVT_ERR_NO_DEVICE;
  }
  return LVT_OK;
}
