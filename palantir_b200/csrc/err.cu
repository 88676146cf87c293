// Error-string plumbing shared by every C-ABI entry point.
#include "common.cuh"
#include <stdio.h>
#include <string.h>

namespace pb200 {
static thread_local char g_err[512] = "";
void set_error(const char* where, cudaError_t e) {
  snprintf(g_err, sizeof(g_err), "%s: %s (%d)", where, cudaGetErrorString(e), (int)e);
}
void set_error_msg(const char* msg) {
  strncpy(g_err, msg, sizeof(g_err) - 1);
  g_err[sizeof(g_err) - 1] = 0;
}
static unsigned long long g_launches = 0;
void count_launch() { __atomic_add_fetch(&g_launches, 1ull, __ATOMIC_RELAXED); }
}  // namespace pb200

// Number of kernel launches issued by this library since it was loaded.
extern "C" long long pb200_launch_count(void) { return (long long)pb200::g_launches; }
extern "C" const char* pb200_last_error(void) { return pb200::g_err; }
extern "C" int pb200_abi_version(void) { return 1; }
// Bind the calling thread to a device (the library links its own static CUDA
// runtime; the host side calls this with torch's current device index).
extern "C" int pb200_set_device(int dev) {
  cudaError_t e = cudaSetDevice(dev);
  if (e != cudaSuccess) { pb200::set_error("cudaSetDevice", e); return (int)e; }
  return 0;
}
extern "C" int pb200_device_sync(void) {
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { pb200::set_error("cudaDeviceSynchronize", e); return (int)e; }
  return 0;
}
