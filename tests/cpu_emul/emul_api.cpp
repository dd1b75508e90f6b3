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

// ---------------------------------------------------------------- multi-rank emulation of the peer-memory exchange (sacb_p2p.cu)
#include <thread>
#include <vector>
namespace cuda_emul {
#ifndef SACB_EMUL_TC      // (cuda_emul_tc.h has the same definition when this file is compiled together with a tensor-core unit)
struct Multicast { const char* base; size_t bytes; int world; char* replica[8]; };
#endif
static std::vector<Multicast> g_multicast;
Multicast* find_multicast(const void* p) {
  for (auto& m : g_multicast) if ((const char*)p >= m.base && (const char*)p < m.base + m.bytes) return &m;
  return nullptr;
}
}
// `key` is any address range of `bytes` bytes that stands for the multicast mapping of the `world` replicas
extern "C" int sacb_emul_register_multicast(const void* key, size_t bytes, int world, void* const* replicas) {
  if (world < 1 || world > 8) return -1;
  cuda_emul::Multicast m{(const char*)key, bytes, world, {}};
  for (int r = 0; r < world; ++r) m.replica[r] = (char*)replicas[r];
  cuda_emul::g_multicast.push_back(m);
  return 0;
}
extern "C" int sacb_emul_clear_multicast(void) { cuda_emul::g_multicast.clear(); return 0; }
// one OS thread per rank calls `fn(descs[r], stream)` (= sacb_allreduce_sgd of the library that also contains this file)
extern "C" int sacb_emul_run_ranks(int (*fn)(const void*, void*), const void* const* descs, int world) {
  std::vector<std::thread> th;
  std::vector<int> rc(world, 0);
  for (int r = 0; r < world; ++r) th.emplace_back([&, r] { cuda_emul::t_inline_launch = true; rc[r] = fn(descs[r], nullptr); });
  for (auto& t : th) t.join();
  for (int r = 0; r < world; ++r) if (rc[r]) return rc[r];
  return 0;
}
