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

// the CUDA built-ins: real (thread-local) objects, re-loaded by the scheduler whenever a fiber is resumed
inline thread_local uint3 threadIdx, blockIdx;
inline thread_local dim3 blockDim, gridDim;

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
  uint3 tid, bid;
  int cta;                       // rank of the fiber's CTA in its cluster
  bool done;
  const void* waiting_on = nullptr;   // mbarrier the fiber spins on (deadlock report)
  // what the fiber is blocked on, so that the scheduler can skip it without two context switches:
  // a barrier generation counter (runnable once *gen != gen_seen) or a 64-bit word (runnable once (*word & mask) != value)
  const int* gen = nullptr; int gen_seen = 0;
  const uint64_t* word = nullptr; uint64_t mask = 0, value = 0;
  bool blocked() const { return (gen && *gen == gen_seen) || (word && ((*word & mask) == value)); }
};

// One per OS worker thread: the cluster (1..8 CTAs; 1 without a cluster launch) it is running.  Fibers are numbered
// cta * threads_per_cta + linear thread id, warps = 32 consecutive fibers (threads_per_cta is a multiple of 32 for clusters).
struct BlockState {
  Ctx sched;
  std::vector<Fiber> fibers;
  char* stacks = nullptr;
  size_t stacks_n = 0;
  int cur = -1;
  int ncta = 1, nthreads = 0;
  unsigned cluster_id = 0, ncluster = 1;
  std::vector<int> live, bar_count, bar_gen;             // per CTA
  int cl_live = 0, cl_count = 0, cl_gen = 0;             // cluster barrier
  std::vector<int> warp_live, warp_count, warp_gen;
  std::vector<uint64_t> slot;
  bool progress = false;
  long yields = 0;
  bool direct = false;           // plain function calls, no fibers
  const std::function<void()>* body = nullptr;
  // dynamic shared memory (1024-byte aligned) and tensor memory ([128 lanes][512 columns] fp32) of every CTA of the cluster
  std::vector<char*> smem_base; std::vector<std::vector<char>> smem_store; size_t smem_bytes = 0;
  std::vector<std::vector<float>> tmem; std::vector<int> tmem_next;
  // SACB_EMUL_ASYNC=1 (cuda_emul_tc.h): TMA copies land a random number of scheduler passes after they were issued and MMAs
  // execute only when a tcgen05.commit forces them -- the latest the hardware may legally be
  struct Deferred { int due; std::function<void()> op; const void* bar; };
  std::vector<Deferred> late_tma;                        // timed: a thread is already waiting on the barrier
  std::vector<Deferred> lazy_tma;                        // nobody observes the barrier yet: lands when somebody first looks at it
  std::vector<std::function<void()>> late_mma;
  int pass = 0;
  ~BlockState() { free(stacks); }
};

inline thread_local BlockState* t_bs = nullptr;

[[noreturn]] inline void die(const char* what) {
  fprintf(stderr, "cuda_emul: %s (block %u,%u,%u thread %u,%u,%u)\n", what, ::blockIdx.x, ::blockIdx.y, ::blockIdx.z,
          ::threadIdx.x, ::threadIdx.y, ::threadIdx.z);
  abort();
}

inline void yield() {
  BlockState* bs = t_bs;
  if (bs->direct) die("synchronisation primitive reached in direct mode (translate.py classified the kernel as barrier-free)");
  bs->yields++;
  ctx_switch(bs->fibers[bs->cur].ctx, bs->sched);
}
inline int my_cta() { return t_bs->fibers[t_bs->cur].cta; }

inline void release_block_barrier(BlockState* bs, int c) { bs->bar_count[c] = 0; bs->bar_gen[c]++; bs->progress = true; }
inline void release_warp_barrier(BlockState* bs, int w) { bs->warp_count[w] = 0; bs->warp_gen[w]++; bs->progress = true; }
inline void release_cluster_barrier(BlockState* bs) { bs->cl_count = 0; bs->cl_gen++; bs->progress = true; }

inline void block_barrier() {
  BlockState* bs = t_bs;
  if (bs->direct) die("__syncthreads in direct mode");
  const int c = my_cta(), gen = bs->bar_gen[c];
  bs->progress = true;
  if (++bs->bar_count[c] == bs->live[c]) { release_block_barrier(bs, c); return; }
  Fiber& f = bs->fibers[bs->cur];
  f.gen = &bs->bar_gen[c]; f.gen_seen = gen;
  while (bs->bar_gen[c] == gen) yield();
  f.gen = nullptr;
}
inline void cluster_barrier() {
  BlockState* bs = t_bs;
  const int gen = bs->cl_gen;
  bs->progress = true;
  if (++bs->cl_count == bs->cl_live) { release_cluster_barrier(bs); return; }
  Fiber& f = bs->fibers[bs->cur];
  f.gen = &bs->cl_gen; f.gen_seen = gen;
  while (bs->cl_gen == gen) yield();
  f.gen = nullptr;
}
inline void warp_barrier() {
  BlockState* bs = t_bs;
  if (bs->direct) die("warp primitive in direct mode");
  const int w = bs->cur / 32, gen = bs->warp_gen[w];
  bs->progress = true;
  if (++bs->warp_count[w] == bs->warp_live[w]) { release_warp_barrier(bs, w); return; }
  Fiber& f = bs->fibers[bs->cur];
  f.gen = &bs->warp_gen[w]; f.gen_seen = gen;
  while (bs->warp_gen[w] == gen) yield();
  f.gen = nullptr;
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
  if (partner < (int)bs->fibers.size() && bs->fibers[partner].cta == bs->fibers[me].cta && !bs->fibers[partner].done)
    memcpy(&r, &bs->slot[partner], sizeof(T));
  warp_barrier();
  return r;
}

inline void fiber_entry() {
  BlockState* bs = t_bs;
  (*bs->body)();
  Fiber& f = bs->fibers[bs->cur];
  f.done = true;
  bs->progress = true;
  const int w = bs->cur / 32, c = f.cta;
  bs->live[c]--; bs->cl_live--;
  bs->warp_live[w]--;
  if (bs->live[c] > 0 && bs->bar_count[c] == bs->live[c]) release_block_barrier(bs, c);   // exited threads do not hold a barrier
  if (bs->warp_live[w] > 0 && bs->warp_count[w] == bs->warp_live[w]) release_warp_barrier(bs, w);
  if (bs->cl_live > 0 && bs->cl_count == bs->cl_live) release_cluster_barrier(bs);
  ctx_switch(f.ctx, bs->sched);
  die("finished fiber resumed");
}

inline void report_deadlock(BlockState* bs) {
  fprintf(stderr, "cuda_emul: DEADLOCK in cluster %u -- no live thread can make progress.  Waiting threads:\n", bs->cluster_id);
  int shown = 0;
  for (size_t i = 0; i < bs->fibers.size() && shown < 40; ++i) {
    const Fiber& f = bs->fibers[i];
    if (f.done) continue;
    if (f.waiting_on || (i % 32) == 0) {
      fprintf(stderr, "  cta %d warp %d lane %d: %s %p\n", f.cta, (int)((i % bs->nthreads) / 32), (int)(i % 32),
              f.waiting_on ? "spins on mbarrier" : "at a block / warp / cluster barrier", f.waiting_on);
      ++shown;
    }
  }
  abort();
}

// one cluster of `ncta` CTAs (first block index `b0`, x-major), fiber mode; returns the number of yields
inline long run_cluster_fibers(BlockState* bs, dim3 grid, dim3 block, long long b0, int ncta) {
  const int nt = (int)(block.x * block.y * block.z), n = nt * ncta;
  if (ncta > 1 && nt % 32) die("cluster launches need a multiple of 32 threads per CTA");
  const int nw = (n + 31) / 32;
  if (bs->stacks_n < (size_t)n) {
    free(bs->stacks);
    bs->stacks = (char*)malloc((size_t)n * STACK_BYTES);
    bs->stacks_n = n;
  }
  bs->ncta = ncta; bs->nthreads = nt;
  bs->fibers.resize(n);
  bs->slot.assign(n, 0);
  bs->warp_live.assign(nw, 0); bs->warp_count.assign(nw, 0); bs->warp_gen.assign(nw, 0);
  bs->live.assign(ncta, nt); bs->bar_count.assign(ncta, 0); bs->bar_gen.assign(ncta, 0);
  bs->cl_live = n; bs->cl_count = 0; bs->cl_gen = 0;
  bs->yields = 0; bs->direct = false;
  bs->tmem_next.assign(ncta, 0);
  for (int i = 0; i < n; ++i) {
    Fiber& f = bs->fibers[i];
    const int t = i % nt;
    const long long b = b0 + i / nt;
    f.cta = i / nt;
    f.tid = uint3{t % block.x, (t / block.x) % block.y, t / (block.x * block.y)};
    f.bid = uint3{(unsigned)(b % grid.x), (unsigned)((b / grid.x) % grid.y), (unsigned)(b / ((long long)grid.x * grid.y))};
    f.done = false; f.waiting_on = nullptr; f.gen = nullptr; f.word = nullptr;
    bs->warp_live[i / 32]++;
    f.ctx.init(bs->stacks + (size_t)i * STACK_BYTES, STACK_BYTES, fiber_entry);
  }
  // SACB_EMUL_SCHED_SEED=s: every pass visits the WARPS in a pseudo-random order and lets a random half of them sit the pass
  // out, so that producers / consumers / epilogue warps overtake each other differently from run to run (other legal
  // interleavings of the same mbarrier protocol).  Warps, not threads: the lanes of a warp stay in step as on the device.
  static const char* seed_env = getenv("SACB_EMUL_SCHED_SEED");
  uint64_t rng = seed_env ? (uint64_t)atoll(seed_env) * 0x9E3779B97F4A7C15ull + (uint64_t)b0 + 1 : 0;
  auto next_rand = [&]() { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; return rng; };
  std::vector<int> order(nw);
  for (int w = 0; w < nw; ++w) order[w] = w;
  int idle_passes = 0;
  bs->late_tma.clear(); bs->lazy_tma.clear(); bs->late_mma.clear(); bs->pass = 0;
  while (bs->cl_live > 0) {
    bs->progress = false;
    bs->pass++;
    if (!bs->late_tma.empty()) {                           // asynchronous copies whose time has come (issue order preserved per due pass)
      std::vector<BlockState::Deferred> todo;
      todo.swap(bs->late_tma);
      for (auto& d : todo) { if (d.due <= bs->pass) { d.op(); bs->progress = true; } else bs->late_tma.push_back(std::move(d)); }
      if (!bs->late_tma.empty()) bs->progress = true;      // something is still in flight: not a deadlock
    }
    if (rng) for (int w = nw - 1; w > 0; --w) std::swap(order[w], order[next_rand() % (uint64_t)(w + 1)]);
    for (int k = 0; k < nw; ++k) {
      if (rng && idle_passes == 0 && (next_rand() & 1)) continue;
      for (int i = order[k] * 32; i < order[k] * 32 + 32 && i < n; ++i) {
        Fiber& f = bs->fibers[i];
        if (f.done || f.blocked()) continue;
        bs->cur = i;
        ::threadIdx = f.tid; ::blockIdx = f.bid;
        ctx_switch(bs->sched, f.ctx);
      }
    }
    if (bs->progress) idle_passes = 0;
    else if (!rng || ++idle_passes > 1) report_deadlock(bs);      // randomised: one full pass with nobody skipped must also be idle
  }
  return bs->yields;
}

inline void run_block_direct(BlockState* bs, dim3 block) {
  bs->direct = true;
  for (unsigned z = 0; z < block.z; ++z)
    for (unsigned y = 0; y < block.y; ++y)
      for (unsigned x = 0; x < block.x; ++x) {
        ::threadIdx = uint3{x, y, z};
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

// set by the multi-rank helper of emul_api.cpp: this OS thread plays one rank; its launches run on this thread alone (no pool,
// no launch mutex), so that the ranks' kernels really run concurrently and synchronise through their flag words
inline thread_local bool t_inline_launch = false;
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

// grid of clusters of `cl` CTAs (cl = 1: ordinary launch); clusters are spread over the pool, a cluster's CTAs share an OS thread
template <class F>
void run_grid_impl(const char* name, dim3 grid, dim3 block, size_t smem, bool syncing, int cl, F body_fn) {
  ProfileScope prof(name);
  const std::function<void()> body(body_fn);
  if (block.x * block.y * block.z == 0 || block.x * block.y * block.z > 1024) {
    fprintf(stderr, "cuda_emul: invalid block size %u x %u x %u\n", block.x, block.y, block.z);
    abort();
  }
  if (smem > 227 * 1024) { fprintf(stderr, "cuda_emul: %zu bytes of dynamic shared memory\n", smem); abort(); }
  std::unique_lock<std::mutex> lk(g_launch_mutex, std::defer_lock);
  if (!t_inline_launch) lk.lock();
  const long long nblocks = (long long)grid.x * grid.y * grid.z;
  if (nblocks == 0) { fprintf(stderr, "cuda_emul: empty grid\n"); abort(); }
  if (cl < 1 || nblocks % cl) { fprintf(stderr, "cuda_emul: grid of %lld blocks is not a multiple of the cluster size %d\n", nblocks, cl); abort(); }
  const long long nclusters = nblocks / cl;
  const char* force = getenv("SACB_EMUL_FIBERS");
  const bool fibers = syncing || cl > 1 || (force && force[0] == '1');
  auto setup = [&](BlockState* bs) {
    t_bs = bs; ::blockDim = block; ::gridDim = grid; bs->body = &body;
    bs->ncluster = (unsigned)nclusters;
    if ((int)bs->smem_store.size() < cl) { bs->smem_store.resize(cl); bs->tmem.resize(cl); }
    bs->smem_base.assign(cl, nullptr); bs->smem_bytes = smem;
    for (int c = 0; c < cl; ++c) {
      if (bs->smem_store[c].size() < smem + 1024) bs->smem_store[c].resize(smem + 1024);
      bs->smem_base[c] = (char*)(((uintptr_t)bs->smem_store[c].data() + 1023) & ~(uintptr_t)1023);
    }
  };
  auto run_cluster = [&](BlockState* bs, long long c) {
    bs->cluster_id = (unsigned)c;
    if (fibers) { run_cluster_fibers(bs, grid, block, c * cl, cl); return; }
    ::blockIdx = uint3{(unsigned)(c % grid.x), (unsigned)((c / grid.x) % grid.y), (unsigned)(c / ((long long)grid.x * grid.y))};
    bs->ncta = 1;
    run_block_direct(bs, block);
  };
  if (nclusters < 4 || t_inline_launch) {               // not worth waking the pool / this thread is one emulated rank
    BlockState& bs = my_block_state();
    setup(&bs);
    for (long long c = 0; c < nclusters; ++c) run_cluster(&bs, c);
  } else {
    std::atomic<long long> next{0};
    const long long chunk = fibers ? 1 : 8;
    Pool::get().run([&](int) {
      BlockState& bs = my_block_state();
      setup(&bs);
      for (;;) {
        const long long c0 = next.fetch_add(chunk);
        if (c0 >= nclusters) break;
        for (long long c = c0; c < c0 + chunk && c < nclusters; ++c) run_cluster(&bs, c);
      }
      t_bs = nullptr;
    });
  }
  t_bs = nullptr;
  g_emulated_launches++;
}
template <class G, class B, class F>
void run_grid(const char* name, G grid_, B block_, size_t smem, bool syncing, F body_fn) {
  run_grid_impl(name, dim3(grid_), dim3(block_), smem, syncing, 1, body_fn);
}
}  // namespace cuda_emul

static inline void __syncthreads() { cuda_emul::block_barrier(); }
template <class T> static inline T __shfl_xor_sync(unsigned mask, T v, int lane_mask) {
  if (mask != 0xffffffffu) cuda_emul::die("only full-warp shuffles are emulated");
  return cuda_emul::shfl_xor(v, lane_mask);
}

template <class T> static inline T __ldg(const T* p) { return *p; }
static inline long long clock64() { return (long long)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
static inline void __nanosleep(unsigned) { if (cuda_emul::t_bs && !cuda_emul::t_bs->direct) { cuda_emul::t_bs->progress = true; cuda_emul::yield(); } }
static inline void __threadfence_system() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline float sacb_bf16_to_float(uint16_t b) { unsigned u = ((unsigned)b) << 16; float f; memcpy(&f, &u, 4); return f; }
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

#ifdef SACB_EMUL_TC
// tensor-core translation units include the (translated) sacb_common.cuh itself; the PTX wrappers come from cuda_emul_tc.h
#include "cuda_emul_tc.h"
#else
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
#endif  // SACB_EMUL_TC
