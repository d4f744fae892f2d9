// Error reporting, version and launch accounting for the mspi_b200 C ABI.
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace mspi {

static thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int num_sms() {   // of the CURRENT device (cached per device: one process may drive several GPUs)
  static int sms[64];
  static bool known[64];
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (dev >= 0 && dev < 64 && known[dev]) return sms[dev];
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
  if (dev >= 0 && dev < 64) { sms[dev] = n; known[dev] = true; }
  return n;
}

static std::atomic<int> g_pdl{-1};
bool pdl_enabled() {
  int v = g_pdl.load(std::memory_order_relaxed);
  if (v < 0) {
    const char* e = getenv("MSPI_PDL");
    v = (!e || atoi(e) != 0) ? 1 : 0;
    g_pdl.store(v, std::memory_order_relaxed);
  }
  return v != 0;
}

bool pdl_small_enabled() {
  static const bool on = [] { const char* e = getenv("MSPI_PDL_SMALL"); return !e || atoi(e) != 0; }();
  return on;
}

}  // namespace mspi

extern "C" int mspi_set_pdl(int on) {
  const int prev = mspi::pdl_enabled() ? 1 : 0;
  mspi::g_pdl.store(on ? 1 : 0);
  return prev;
}
extern "C" const char* mspi_last_error(void) { return mspi::g_err; }
extern "C" int mspi_version(void) { return 1; }
extern "C" const char* mspi_arch(void) { return "sm_100a"; }
extern "C" int64_t mspi_launch_count(void) { return mspi::g_launches.load(); }
