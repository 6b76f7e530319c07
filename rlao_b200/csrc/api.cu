// Error plumbing, ABI version and launch accounting for libaoenv_b200.
#include <stdarg.h>
#include <stdlib.h>

#include <mutex>

#include "common.cuh"
#include "tma.cuh"

namespace aoenv {

thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};
int g_pdl = [] {
  const char* v = getenv("AOENV_PDL");
  return (v != nullptr && v[0] == '0') ? 0 : 1;
}();

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

namespace tma {
EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}
}  // namespace tma

}  // namespace aoenv

extern "C" {

int aoenv_abi_version(void) { return AOENV_ABI_VERSION; }
const char* aoenv_last_error(void) { return aoenv::g_err; }
uint64_t aoenv_launch_count(void) { return aoenv::g_launches.load(); }
int aoenv_set_pdl(int enabled) {
  const int old = aoenv::g_pdl;
  aoenv::g_pdl = enabled ? 1 : 0;
  return old;
}

}  // extern "C"
