// byol_b200 — C-ABI plumbing shared by every entry point: thread-local error string, launch checks, version.
#include <stdarg.h>
#include <stdio.h>

#include "common.cuh"

namespace byol {

static thread_local char g_last_error[512] = "";

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
}

// Launch-configuration errors surface here; asynchronous faults surface at the caller's next sync.
int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_last_error("%s: launch failed: %s", what, cudaGetErrorString(e));
    return -100;
  }
  return 0;
}

int device_slot() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) dev = 0;
  return dev < kMaxDevices ? dev : kMaxDevices - 1;
}

int device_sm_count() {
  static int n[kMaxDevices] = {};
  const int slot = device_slot();
  if (n[slot] == 0) {
    int dev = 0, v = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    n[slot] = v;
  }
  return n[slot];
}

}  // namespace byol

extern "C" const char* byol_last_error(void) { return byol::g_last_error; }

extern "C" int byol_abi_version(void) { return 2; }

// number of SMs of the current device (used by the host side to size persistent grids); < 0 on error
extern "C" int byol_device_sm_count(void) {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
  return n;
}
