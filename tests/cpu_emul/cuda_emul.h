// TEST INFRASTRUCTURE -- never part of the product, never loaded by da_sac_b200/ (lib.py refuses it).
//
// A small host emulation of the CUDA execution model, enough to compile the STREAMING kernels of da_sac_b200/csrc (everything
// except the tcgen05 / TMA GEMM kernels and the peer-memory exchange) with g++ and run them in the GPU-less build container:
//   * a CUDA thread is a fiber (ucontext) with its own stack; the fibers of a block share one OS thread; blocks are spread
//     over a pool of OS threads;
//   * threadIdx / blockIdx / blockDim / gridDim are thread_local and re-loaded whenever a fiber is resumed;
//   * __syncthreads() is a real barrier over the live fibers of the block; __shfl_xor_sync exchanges through per-thread slots
//     between two warp barriers (warps = 32 consecutive linear thread ids, as on the device);
//   * __shared__ is `static thread_local` (an OS thread holds one block at a time);
//   * atomics are std::atomic_ref (blocks on different OS threads do race, as on the device);
//   * kernel<<<grid, block, smem, stream>>>(args) is rewritten to cuda_emul::run_grid(...) by translate.py; the grid runs to
//     completion before the call returns ("stream order" is program order);
//   * kernels that contain no barrier / warp primitive (translate.py decides from the source) run as plain function calls,
//     one CUDA thread after the other, without fibers; a primitive reached in that mode aborts.
// What this checks: indexing, bounds (run it under -fsanitize=address), aliasing, arithmetic, barrier placement (a barrier that
// not every live thread reaches is reported as a deadlock), and the launch geometry computed by the C-ABI entry points.
// What it cannot check: tcgen05 / TMA / mbarrier code, memory-ordering bugs, performance.
#pragma once
#define SACB_HOST_EMUL 1
#include <math.h>
#include <stdarg.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <string>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static thread_local
#define SACB_DEVINL inline

struct uint3 { unsigned x, y, z; };
struct int3 { int x, y, z; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct alignas(8) uint2 { unsigned x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(8) float2 { float x, y; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline float2 make_float2(float x, float y) { return float2{x, y}; }

// CUDA's global min / max overloads
#include <type_traits>
template <class A, class B> static inline typename std::common_type<A, B>::type min(A a, B b) {
  typedef typename std::common_type<A, B>::type T; return (T)b < (T)a ? (T)b : (T)a;
}
template <class A, class B> static inline typename std::common_type<A, B>::type max(A a, B b) {
  typedef typename std::common_type<A, B>::type T; return (T)a < (T)b ? (T)b : (T)a;
}

typedef void* cudaStream_t;
typedef int cudaError_t;
constexpr cudaError_t cudaSuccess = 0;
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
static inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }

namespace cuda_emul {

constexpr size_t STACK_BYTES = 96 * 1024;

// Context switch between the scheduler and a fiber.  glibc's swapcontext makes a system call (signal mask) per switch, which
// dominated the barrier-heavy kernels; on x86-64 a hand-written switch of the callee-saved registers is used instead.
#if defined(__x86_64__) && !defined(SACB_EMUL_UCONTEXT)
extern "C" void sacb_emul_switch(void** save_sp, void* new_sp);
asm(R"(
.text
.weak sacb_emul_switch
.type sacb_emul_switch,@function
sacb_emul_switch:
  pushq %rbp
  pushq %rbx
  pushq %r12
  pushq %r13
  pushq %r14
  pushq %r15
  movq %rsp, (%rdi)
  movq %rsi, %rsp
  popq %r15
  popq %r14
  popq %r13
  popq %r12
  popq %rbx
  popq %rbp
  ret
.size sacb_emul_switch,.-sacb_emul_switch
)");
struct Ctx {
  void* sp = nullptr;
  void init(char* stack, size_t bytes, void (*entry)()) {
    uintptr_t top = ((uintptr_t)stack + bytes) & ~(uintptr_t)15;
    void** p = (void**)top;
    *--p = nullptr;                 // return address of `entry` (it never returns)
    *--p = (void*)entry;            // popped by the `ret` of the first switch
    for (int i = 0; i < 6; ++i) *--p = nullptr;
    sp = (void*)p;
  }
};
inline void ctx_switch(Ctx& from, Ctx& to) { sacb_emul_switch(&from.sp, to.sp); }
#else
struct Ctx {
  ucontext_t uc;
  void init(char* stack, size_t bytes, void (*entry)()) {
    getcontext(&uc);
    uc.uc_stack.ss_sp = stack; uc.uc_stack.ss_size = bytes; uc.uc_link = nullptr;
    makecontext(&uc, entry, 0);
  }
};
inline void ctx_switch(Ctx& from, Ctx& to) { swapcontext(&from.uc, &to.uc); }
#endif

struct Fiber {
  Ctx ctx;
  uint3 tid;
  bool done;
};

struct BlockState {              // one per OS worker thread: the block it is running
  Ctx sched;
  std::vector<Fiber> fibers;
  char* stacks = nullptr;
  size_t stacks_n = 0;
  int cur = -1;
  int live = 0, bar_count = 0, bar_gen = 0;
  std::vector<int> warp_live, warp_count, warp_gen;
  std::vector<uint64_t> slot;
  bool progress = false;
  long yields = 0;
  bool direct = false;           // plain function calls, no fibers
  const std::function<void()>* body = nullptr;
  std::vector<char> dyn_smem;
  ~BlockState() { free(stacks); }
};

inline thread_local uint3 t_threadIdx, t_blockIdx;
inline thread_local dim3 t_blockDim, t_gridDim;
inline thread_local BlockState* t_bs = nullptr;

[[noreturn]] inline void die(const char* what) {
  fprintf(stderr, "cuda_emul: %s (block %u,%u,%u thread %u,%u,%u)\n", what, t_blockIdx.x, t_blockIdx.y, t_blockIdx.z,
          t_threadIdx.x, t_threadIdx.y, t_threadIdx.z);
  abort();
}

inline void yield() {
  BlockState* bs = t_bs;
  if (bs->direct) die("synchronisation primitive reached in direct mode (the first block of this launch never synchronised)");
  bs->yields++;
  ctx_switch(bs->fibers[bs->cur].ctx, bs->sched);
}

inline void release_block_barrier(BlockState* bs) { bs->bar_count = 0; bs->bar_gen++; bs->progress = true; }
inline void release_warp_barrier(BlockState* bs, int w) { bs->warp_count[w] = 0; bs->warp_gen[w]++; bs->progress = true; }

inline void block_barrier() {
  BlockState* bs = t_bs;
  if (bs->direct) die("__syncthreads in direct mode");
  const int gen = bs->bar_gen;
  bs->progress = true;
  if (++bs->bar_count == bs->live) { release_block_barrier(bs); return; }
  while (bs->bar_gen == gen) yield();
}
inline void warp_barrier() {
  BlockState* bs = t_bs;
  if (bs->direct) die("warp primitive in direct mode");
  const int w = bs->cur / 32, gen = bs->warp_gen[w];
  bs->progress = true;
  if (++bs->warp_count[w] == bs->warp_live[w]) { release_warp_barrier(bs, w); return; }
  while (bs->warp_gen[w] == gen) yield();
}
template <class T>
inline T shfl_xor(T v, int lane_mask) {
  static_assert(sizeof(T) <= 8, "shuffle of more than 8 bytes");
  BlockState* bs = t_bs;
  const int me = bs->cur, base = me & ~31, partner = base | ((me & 31) ^ lane_mask);
  uint64_t bits = 0;
  memcpy(&bits, &v, sizeof(T));
  bs->slot[me] = bits;
  warp_barrier();
  T r = v;
  if (partner < (int)bs->fibers.size() && !bs->fibers[partner].done) memcpy(&r, &bs->slot[partner], sizeof(T));
  warp_barrier();
  return r;
}

inline void fiber_entry() {
  BlockState* bs = t_bs;
  (*bs->body)();
  Fiber& f = bs->fibers[bs->cur];
  f.done = true;
  bs->progress = true;
  const int w = bs->cur / 32;
  bs->live--;
  bs->warp_live[w]--;
  if (bs->live > 0 && bs->bar_count == bs->live) release_block_barrier(bs);          // exited threads do not hold a barrier
  if (bs->warp_live[w] > 0 && bs->warp_count[w] == bs->warp_live[w]) release_warp_barrier(bs, w);
  ctx_switch(f.ctx, bs->sched);
  die("finished fiber resumed");
}

// one block, fiber mode; returns the number of yields
inline long run_block_fibers(BlockState* bs, dim3 block) {
  const int n = (int)(block.x * block.y * block.z), nw = (n + 31) / 32;
  if (bs->stacks_n < (size_t)n) {
    free(bs->stacks);
    bs->stacks = (char*)malloc((size_t)n * STACK_BYTES);
    bs->stacks_n = n;
  }
  bs->fibers.resize(n);
  bs->slot.assign(n, 0);
  bs->warp_live.assign(nw, 0); bs->warp_count.assign(nw, 0); bs->warp_gen.assign(nw, 0);
  bs->live = n; bs->bar_count = 0; bs->bar_gen = 0; bs->yields = 0; bs->direct = false;
  for (int t = 0; t < n; ++t) {
    Fiber& f = bs->fibers[t];
    f.tid = uint3{t % block.x, (t / block.x) % block.y, t / (block.x * block.y)};
    f.done = false;
    bs->warp_live[t / 32]++;
    f.ctx.init(bs->stacks + (size_t)t * STACK_BYTES, STACK_BYTES, fiber_entry);
  }
  while (bs->live > 0) {
    bs->progress = false;
    for (int t = 0; t < n; ++t) {
      Fiber& f = bs->fibers[t];
      if (f.done) continue;
      bs->cur = t;
      t_threadIdx = f.tid;
      ctx_switch(bs->sched, f.ctx);
    }
    if (!bs->progress) die("deadlock: a barrier is not reached by every live thread of the block / warp");
  }
  return bs->yields;
}

inline void run_block_direct(BlockState* bs, dim3 block) {
  bs->direct = true;
  for (unsigned z = 0; z < block.z; ++z)
    for (unsigned y = 0; y < block.y; ++y)
      for (unsigned x = 0; x < block.x; ++x) {
        t_threadIdx = uint3{x, y, z};
        (*bs->body)();
      }
}

// ---- pool of OS threads; job(worker) runs on every worker
class Pool {
 public:
  static Pool& get() { static Pool* p = new Pool; return *p; }      // leaked on purpose: its threads are detached
  int size() const { return (int)threads_.size(); }
  void run(const std::function<void(int)>& job) {
    std::lock_guard<std::mutex> one_job(run_m_);
    std::unique_lock<std::mutex> lk(m_);
    job_ = &job; pending_ = size(); epoch_++;
    cv_.notify_all();
    done_.wait(lk, [&] { return pending_ == 0; });
    job_ = nullptr;
  }
 private:
  Pool() {
    int n = (int)std::thread::hardware_concurrency();
    if (const char* e = getenv("SACB_EMUL_THREADS")) n = atoi(e);
    if (n < 1) n = 1;
    if (n > 16) n = 16;
    for (int i = 0; i < n; ++i) threads_.emplace_back([this, i] { loop(i); });
    for (auto& t : threads_) t.detach();
  }
  void loop(int i) {
    long seen = 0;
    for (;;) {
      const std::function<void(int)>* job;
      {
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [&] { return epoch_ != seen; });
        seen = epoch_; job = job_;
      }
      (*job)(i);
      {
        std::unique_lock<std::mutex> lk(m_);
        if (--pending_ == 0) done_.notify_all();
      }
    }
  }
  std::mutex m_, run_m_;
  std::condition_variable cv_, done_;
  std::vector<std::thread> threads_;
  const std::function<void(int)>* job_ = nullptr;
  long epoch_ = 0;
  int pending_ = 0;
};

inline BlockState& my_block_state() { static thread_local BlockState bs; return bs; }
inline std::atomic<long long> g_emulated_launches{0};
inline std::mutex g_launch_mutex;      // one grid at a time

// SACB_EMUL_PROFILE=1: host seconds per kernel name, printed at exit (where the emulation spends its time, nothing more)
struct Profile {
  std::mutex m;
  std::vector<std::pair<std::string, std::pair<double, long>>> rows;
  bool on = getenv("SACB_EMUL_PROFILE") && getenv("SACB_EMUL_PROFILE")[0] == '1';
  void add(const char* name, double s) {
    std::lock_guard<std::mutex> lk(m);
    for (auto& r : rows) if (r.first == name) { r.second.first += s; r.second.second++; return; }
    rows.push_back({name, {s, 1}});
  }
  ~Profile() {
    if (!on) return;
    std::sort(rows.begin(), rows.end(), [](auto& a, auto& b) { return a.second.first > b.second.first; });
    for (auto& r : rows) fprintf(stderr, "cuda_emul %9.3f s %7ld launches  %s\n", r.second.first, r.second.second, r.first.c_str());
  }
};
inline Profile& profile() { static Profile p; return p; }
struct ProfileScope {
  const char* name; std::chrono::steady_clock::time_point t0;
  explicit ProfileScope(const char* n) : name(n), t0(std::chrono::steady_clock::now()) {}
  ~ProfileScope() { if (profile().on) profile().add(name, std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count()); }
};

template <class G, class B, class F>
void run_grid(const char* name, G grid_, B block_, size_t smem, bool syncing, F body_fn) {
  ProfileScope prof(name);
  const dim3 grid(grid_), block(block_);
  const std::function<void()> body(body_fn);
  if (block.x * block.y * block.z == 0 || block.x * block.y * block.z > 1024) {
    fprintf(stderr, "cuda_emul: invalid block size %u x %u x %u\n", block.x, block.y, block.z);
    abort();
  }
  if (smem > 227 * 1024) { fprintf(stderr, "cuda_emul: %zu bytes of dynamic shared memory\n", smem); abort(); }
  std::lock_guard<std::mutex> lk(g_launch_mutex);
  const long long nblocks = (long long)grid.x * grid.y * grid.z;
  if (nblocks == 0) { fprintf(stderr, "cuda_emul: empty grid\n"); abort(); }
  const char* force = getenv("SACB_EMUL_FIBERS");
  const bool fibers = syncing || (force && force[0] == '1');
  auto run_block = [&](BlockState* bs, long long b) {
    t_blockIdx = uint3{(unsigned)(b % grid.x), (unsigned)((b / grid.x) % grid.y), (unsigned)(b / ((long long)grid.x * grid.y))};
    if (fibers) run_block_fibers(bs, block);
    else run_block_direct(bs, block);
  };
  auto setup = [&](BlockState* bs) {
    t_bs = bs; t_blockDim = block; t_gridDim = grid; bs->body = &body;
    if (bs->dyn_smem.size() < smem) bs->dyn_smem.resize(smem);
  };
  if (nblocks < 4) {                                   // not worth waking the pool
    BlockState& bs = my_block_state();
    setup(&bs);
    for (long long b = 0; b < nblocks; ++b) run_block(&bs, b);
  } else {
    std::atomic<long long> next{0};
    const long long chunk = fibers ? 1 : 8;
    Pool::get().run([&](int) {
      BlockState& bs = my_block_state();
      setup(&bs);
      for (;;) {
        const long long b0 = next.fetch_add(chunk);
        if (b0 >= nblocks) break;
        for (long long b = b0; b < b0 + chunk && b < nblocks; ++b) run_block(&bs, b);
      }
      t_bs = nullptr;
    });
  }
  t_bs = nullptr;
  g_emulated_launches++;
}
}  // namespace cuda_emul

#define threadIdx (cuda_emul::t_threadIdx)
#define blockIdx (cuda_emul::t_blockIdx)
#define blockDim (cuda_emul::t_blockDim)
#define gridDim (cuda_emul::t_gridDim)
static inline void __syncthreads() { cuda_emul::block_barrier(); }
template <class T> static inline T __shfl_xor_sync(unsigned mask, T v, int lane_mask) {
  if (mask != 0xffffffffu) cuda_emul::die("only full-warp shuffles are emulated");
  return cuda_emul::shfl_xor(v, lane_mask);
}

template <class T> static inline T __ldg(const T* p) { return *p; }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline float __int_as_float(int u) { float f; memcpy(&f, &u, 4); return f; }
static inline int __float_as_int(float f) { int u; memcpy(&u, &f, 4); return u; }

template <class T> static inline T atomicAdd(T* p, T v) { return std::atomic_ref<T>(*p).fetch_add(v, std::memory_order_relaxed); }
static inline int atomicMax(int* p, int v) {
  std::atomic_ref<int> a(*p);
  int old = a.load(std::memory_order_relaxed);
  while (old < v && !a.compare_exchange_weak(old, v, std::memory_order_relaxed)) {}
  return old;
}

// ---- bf16 (round to nearest even, as __float2bfloat16_rn)
struct __nv_bfloat16 { uint16_t bits; };
struct __nv_bfloat162 { __nv_bfloat16 x, y; };
static inline __nv_bfloat16 __float2bfloat16_rn(float v) {
  unsigned u = __float_as_uint(v);
  if ((u & 0x7FFFFFFFu) > 0x7F800000u) return __nv_bfloat16{(uint16_t)0x7FFF};
  u += 0x7FFFu + ((u >> 16) & 1u);
  return __nv_bfloat16{(uint16_t)(u >> 16)};
}
static inline float __bfloat162float(__nv_bfloat16 h) { return __uint_as_float(((unsigned)h.bits) << 16); }
static inline __nv_bfloat162 __floats2bfloat162_rn(float a, float b) { return __nv_bfloat162{__float2bfloat16_rn(a), __float2bfloat16_rn(b)}; }
static inline float2 __bfloat1622float2(__nv_bfloat162 h) { return float2{__bfloat162float(h.x), __bfloat162float(h.y)}; }

namespace sacb {
void set_error(const char* fmt, ...);
// the helpers of sacb_common.cuh the streaming kernels use
static inline void split_bf16(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(v);
  lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}
static inline uint16_t float_to_bf16_bits(float v) { return __float2bfloat16_rn(v).bits; }
static inline float bf16_bits_to_float(uint16_t b) { return __uint_as_float(((unsigned)b) << 16); }
}  // namespace sacb

#define SACB_CHECK_CUDA(expr)                                                      \
  do {                                                                             \
    cudaError_t _e = (expr);                                                       \
    if (_e != cudaSuccess) { sacb::set_error("%s:%d %s", __FILE__, __LINE__, #expr); return -2; } \
  } while (0)
#define SACB_REQUIRE(cond, ...)                                                    \
  do {                                                                             \
    if (!(cond)) { sacb::set_error(__VA_ARGS__); return -1; }                      \
  } while (0)
