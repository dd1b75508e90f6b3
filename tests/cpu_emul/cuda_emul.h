// TEST INFRASTRUCTURE -- never part of the product, never loaded by da_sac_b200/.
//
// Minimal host emulation of the CUDA execution model, just enough to compile the streaming kernels of
// da_sac_b200/csrc/*_kernels.cuh UNCHANGED with g++ and run them in the GPU-less build container:
//   * one std::thread per CUDA thread of a block, the blocks of a grid one after another;
//   * threadIdx / blockIdx / blockDim / gridDim are thread_local;
//   * __syncthreads() is a std::barrier over the block; __shared__ is `static` (one block is resident at a time);
//   * kernel<<<grid, block, smem, stream>>>(args) is spelled SACB_LAUNCH(kernel, grid, block, smem, stream, args) in the
//     .cu files that support emulation; here it runs the grid synchronously.
// What this checks: indexing, bounds, aliasing, arithmetic order, the launch geometry computed by the C-ABI entry points.
// What it cannot check: tcgen05 / TMA / mbarrier code, memory-model races, performance.
#pragma once
#define SACB_HOST_EMUL 1
#include <stdint.h>
#include <stddef.h>
#include <math.h>
#include <string.h>
#include <stdarg.h>
#include <stdio.h>
#include <atomic>
#include <barrier>
#include <thread>
#include <vector>
#include <functional>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
#define SACB_DEVINL inline

struct uint3 { unsigned x, y, z; };
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

typedef void* cudaStream_t;
typedef int cudaError_t;
constexpr cudaError_t cudaSuccess = 0;
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }

namespace cuda_emul {
inline thread_local uint3 t_threadIdx, t_blockIdx;
inline thread_local dim3 t_blockDim, t_gridDim;
inline thread_local std::barrier<>* t_barrier = nullptr;
inline std::atomic<long long> g_emulated_launches{0};

// Runs `body` for every thread of every block.  blockDim threads live for the whole launch and walk the blocks together
// (a barrier separates two blocks: `static` shared memory is reused).
template <class F>
void run_grid(dim3 grid, dim3 block, F body) {
  const unsigned nthreads = block.x * block.y * block.z;
  std::barrier<> bar((ptrdiff_t)nthreads);
  auto worker = [&](unsigned tid) {
    t_blockDim = block; t_gridDim = grid; t_barrier = &bar;
    t_threadIdx = uint3{tid % block.x, (tid / block.x) % block.y, tid / (block.x * block.y)};
    for (unsigned bz = 0; bz < grid.z; ++bz)
      for (unsigned by = 0; by < grid.y; ++by)
        for (unsigned bx = 0; bx < grid.x; ++bx) {
          t_blockIdx = uint3{bx, by, bz};
          body();
          bar.arrive_and_wait();
        }
  };
  std::vector<std::thread> pool;
  pool.reserve(nthreads);
  for (unsigned t = 0; t < nthreads; ++t) pool.emplace_back(worker, t);
  for (auto& th : pool) th.join();
  g_emulated_launches++;
}
}  // namespace cuda_emul

#define threadIdx (cuda_emul::t_threadIdx)
#define blockIdx (cuda_emul::t_blockIdx)
#define blockDim (cuda_emul::t_blockDim)
#define gridDim (cuda_emul::t_gridDim)
static inline void __syncthreads() { cuda_emul::t_barrier->arrive_and_wait(); }

#define SACB_LAUNCH(kernel, grid, block, smem, stream, ...) \
  cuda_emul::run_grid(dim3(grid), dim3(block), [&]() { kernel(__VA_ARGS__); })

template <class T> static inline T __ldg(const T* p) { return *p; }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }

namespace sacb {
void set_error(const char* fmt, ...);
// round-to-nearest-even fp32 -> bf16, NaN kept quiet: what __float2bfloat16_rn does
static inline uint16_t float_to_bf16_bits(float v) {
  unsigned u = __float_as_uint(v);
  if ((u & 0x7FFFFFFFu) > 0x7F800000u) return (uint16_t)0x7FFF;
  u += 0x7FFFu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
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
