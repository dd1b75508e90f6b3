// TEST INFRASTRUCTURE: error reporting for the host-emulated kernels (the product's lives in csrc/sacb_api.cu).
#include "cuda_emul.h"
#include "sacb.h"

namespace sacb {
std::atomic<long long> g_launches{0};
static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace sacb

extern "C" const char* sacb_last_error(void) { return sacb::g_err; }
extern "C" int64_t sacb_launch_count(void) { return (int64_t)sacb::g_launches.load(); }
extern "C" int sacb_abi_version(void) { return SACB_ABI_VERSION; }
extern "C" int sacb_emul_marker(void) { return 1; }   // the product library does not export this symbol

// name of the last tcgen05 kernel instantiation launched through the emulated cudaLaunchKernelEx (tests assert on the variant)
namespace cuda_emul {
static std::mutex g_last_m;
static char g_last_kernel[256] = "";
void set_last_kernel(const char* name) { std::lock_guard<std::mutex> lk(g_last_m); snprintf(g_last_kernel, sizeof(g_last_kernel), "%s", name); }
}
extern "C" const char* sacb_emul_last_kernel(void) { return cuda_emul::g_last_kernel; }
