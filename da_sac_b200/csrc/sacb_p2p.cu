// Gradient all-reduce fused with the SGD update over NVLink peer memory (one kernel, no NCCL):
//
//   reduce-scatter : rank r sums ITS 1/W slice of the flat gradient buffer over all W ranks with peer loads
//   SGD            : weight decay + momentum + update on that slice only (the momentum buffer is sharded, ZeRO-1 style)
//   all-gather     : the updated weights of the slice are stored straight into every rank's flat parameter buffer
//
// Replaces, for the data-parallel target step, DistributedDataParallel's bucketed gradient all-reduce
// (/root/reference/train.py:104,232) followed by torch.optim.SGD.step() (train.py:233, base_trainer.py:61-66).
// Every element is reduced by exactly one rank in the fixed order 0..W-1, so all replicas receive bit-identical weights
// and the result does not depend on timing.  Cross-GPU ordering uses epoch flags in peer-mapped memory
// (st.release.sys / ld.acquire.sys); the epoch counter lives on the device, so the kernel can sit inside a CUDA graph.
// Every wait is bounded and traps instead of hanging the GPU.
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "sacb_common.cuh"
#include "../../include/sacb.h"

namespace sacb {
extern std::atomic<long long> g_launches;

constexpr int P2P_MAXW = SACB_P2P_MAX_WORLD;
constexpr int FLAG_ARRIVE = 0;                 // [MAXW]  peer p -> "my gradients of epoch e are complete"
constexpr int FLAG_DONE = P2P_MAXW;            // [MAXW]  peer p -> "my slice of epoch e is stored in your parameters"
constexpr int FLAG_EPOCH = 2 * P2P_MAXW;       // completed epochs of THIS rank
constexpr int FLAG_BLOCKS = 2 * P2P_MAXW + 1;  // blocks of the running kernel that finished their share
// A rank that waits for a peer gives up (trap, never a silent hang) after SACB_P2P_TIMEOUT_S seconds of SM clocks: default
// 600 s, the order of NCCL's watchdog, so that ordinary rank skew -- a checkpoint write on rank 0, a dataloader stall at an epoch
// boundary, a debugger -- does not kill the context.
constexpr double SPIN_CLOCK_HZ = 2.0e9;

struct P2PArgs {
  float* grads[P2P_MAXW];
  float* params[P2P_MAXW];
  uint32_t* flags[P2P_MAXW];
  float* mc_grads; float* mc_params;     // NVLS instantiation: multicast mappings of the same gradient / parameter buffers
  float* mom;
  const int64_t* seg_ranges; const float* seg_lr; const float* seg_wd;
  int nseg, world, rank, first;
  long long vec_lo, vec_hi;      // this rank's slice in float4 units
  float mu, inv_world;
  long long spin_limit;          // SM clocks a cross-rank wait may take before the kernel traps
};

SACB_DEVINL void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
SACB_DEVINL uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
SACB_DEVINL float4 ld_peer_f4(const float* p) {       // relaxed system-scope load: never served from a stale L1 line
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
SACB_DEVINL void st_peer_f4(float* p, const float4& v) {
  asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// NVLS (NVLink SHARP): one multimem.ld_reduce returns the sum of the W replicas of an address -- the NVSwitch does the adds --
// and one multimem.st writes all W replicas, so a rank moves 1/W of the buffer once in each direction instead of W times.
SACB_DEVINL float4 multimem_ld_reduce_add_f4(const float* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc) : "memory");
  return v;
}
SACB_DEVINL void multimem_st_f4(float* mc, const float4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
SACB_DEVINL void spin_until(const uint32_t* flag, uint32_t epoch, long long spin_limit) {
  const long long t0 = clock64();
  while ((int)(ld_acquire_sys(flag) - epoch) < 0) {
    __nanosleep(200);
    if (clock64() - t0 > spin_limit) { printf("sacb allreduce_sgd: peer flag timeout (epoch %u)\n", epoch); __trap(); }
  }
}

template <bool NVLS>
__global__ void __launch_bounds__(512)
allreduce_sgd_kernel(const P2PArgs a) {
  extern __shared__ int64_t s_seg[];                  // [2*nseg] segment table (begin, end), sorted by begin
  __shared__ uint32_t s_epoch;
  uint32_t* my_flags = a.flags[a.rank];
  for (int i = threadIdx.x; i < 2 * a.nseg; i += blockDim.x) s_seg[i] = a.seg_ranges[i];
  if (threadIdx.x == 0) s_epoch = ld_acquire_sys(my_flags + FLAG_EPOCH) + 1;
  __syncthreads();
  const uint32_t epoch = s_epoch;
  // ---- arrive: this rank's gradients are complete (the kernel is stream-ordered after the backward pass)
  if (blockIdx.x == 0 && threadIdx.x < a.world) {
    __threadfence_system();
    st_release_sys(a.flags[threadIdx.x] + FLAG_ARRIVE + a.rank, epoch);
  }
  if (threadIdx.x < a.world) spin_until(my_flags + FLAG_ARRIVE + threadIdx.x, epoch, a.spin_limit);
  __syncthreads();

  // ---- reduce-scatter + SGD + all-gather on this rank's slice
  const long long stride = (long long)gridDim.x * blockDim.x;
  float* my_params = a.params[a.rank];
  for (long long v = a.vec_lo + (long long)blockIdx.x * blockDim.x + threadIdx.x; v < a.vec_hi; v += stride) {
    const int64_t i0 = v * 4;
    // segment that contains element i0 (tensors are 16-byte aligned: a float4 never straddles two of them)
    int lo = 0, hi = a.nseg;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (s_seg[2 * mid] <= i0) lo = mid + 1; else hi = mid; }
    const int seg = lo - 1;
    if (seg < 0) continue;
    const int64_t end = s_seg[2 * seg + 1];
    if (i0 >= end) continue;                         // BN running statistics / padding: not an optimiser tensor
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    if constexpr (NVLS) {
      g = multimem_ld_reduce_add_f4(a.mc_grads + i0);            // sum over the ranks, reduced inside the switch
    } else {
#pragma unroll
      for (int p = 0; p < P2P_MAXW; ++p) {
        if (p < a.world) {
          const float4 t = ld_peer_f4(a.grads[p] + i0);
          g.x += t.x; g.y += t.y; g.z += t.z; g.w += t.w;
        }
      }
    }
    const float l = a.seg_lr[seg], w = a.seg_wd[seg];
    const float4 pv = *reinterpret_cast<const float4*>(my_params + i0);
    float4 m = a.first ? make_float4(0.f, 0.f, 0.f, 0.f) : *reinterpret_cast<const float4*>(a.mom + i0);
    float gg[4] = {g.x * a.inv_world, g.y * a.inv_world, g.z * a.inv_world, g.w * a.inv_world};   // DDP: mean over ranks
    const float pp[4] = {pv.x, pv.y, pv.z, pv.w};
    float mm[4] = {m.x, m.y, m.z, m.w};
    float out[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float d = gg[k];
      if (w != 0.f) d = fmaf(w, pp[k], d);                       // grad.add(param, alpha=weight_decay)
      const float buf = a.first ? d : fmaf(a.mu, mm[k], d);      // buf.mul_(momentum).add_(grad)
      mm[k] = buf;
      out[k] = (i0 + k < end) ? pp[k] - l * buf : pp[k];         // param.add_(buf, alpha=-lr)
    }
    *reinterpret_cast<float4*>(a.mom + i0) = make_float4(mm[0], mm[1], mm[2], mm[3]);
    const float4 nv = make_float4(out[0], out[1], out[2], out[3]);
    if constexpr (NVLS) {
      multimem_st_f4(a.mc_params + i0, nv);                      // lands in every rank's parameter buffer
    } else {
#pragma unroll
      for (int p = 0; p < P2P_MAXW; ++p)
        if (p < a.world) st_peer_f4(a.params[p] + i0, nv);
    }
  }

  // ---- done: the last block of this rank tells every peer, then waits until every peer's slice has landed here
  __threadfence_system();
  __syncthreads();
  __shared__ int s_last;
  if (threadIdx.x == 0) s_last = (atomicAdd(my_flags + FLAG_BLOCKS, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!s_last) return;
  if (threadIdx.x < a.world) {
    __threadfence_system();
    st_release_sys(a.flags[threadIdx.x] + FLAG_DONE + a.rank, epoch);
    spin_until(my_flags + FLAG_DONE + threadIdx.x, epoch, a.spin_limit);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    my_flags[FLAG_BLOCKS] = 0;
    st_release_sys(my_flags + FLAG_EPOCH, epoch);
  }
}
}  // namespace sacb

using namespace sacb;
#define ST ((cudaStream_t)stream)

extern "C" int sacb_p2p_flag_words(void) { return 2 * P2P_MAXW + 2; }

extern "C" int sacb_symm_alloc(size_t bytes, void** dptr) {
  SACB_REQUIRE(dptr && bytes > 0, "sacb_symm_alloc: bad arguments");
  SACB_CHECK_CUDA(cudaMalloc(dptr, bytes));            // plain cudaMalloc: exportable with cudaIpcGetMemHandle
  SACB_CHECK_CUDA(cudaMemset(*dptr, 0, bytes));
  return 0;
}
extern "C" int sacb_symm_free(void* dptr) {
  SACB_CHECK_CUDA(cudaFree(dptr));
  return 0;
}
extern "C" int sacb_ipc_export(const void* dptr, void* handle64) {
  static_assert(sizeof(cudaIpcMemHandle_t) == SACB_IPC_HANDLE_BYTES, "IPC handle size");
  SACB_CHECK_CUDA(cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(handle64), const_cast<void*>(dptr)));
  return 0;
}
extern "C" int sacb_ipc_import(const void* handle64, void** peer_ptr) {
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  SACB_CHECK_CUDA(cudaIpcOpenMemHandle(peer_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}
extern "C" int sacb_ipc_close(void* peer_ptr) {
  SACB_CHECK_CUDA(cudaIpcCloseMemHandle(peer_ptr));
  return 0;
}

extern "C" int sacb_allreduce_sgd(const SacbAllreduceSgd* d, void* stream) {
  SACB_REQUIRE(d && d->size == sizeof(SacbAllreduceSgd), "sacb_allreduce_sgd: bad descriptor size");
  SACB_REQUIRE(d->world >= 1 && d->world <= P2P_MAXW && d->rank >= 0 && d->rank < d->world,
               "sacb_allreduce_sgd: world %d / rank %d out of range (max world %d)", d->world, d->rank, P2P_MAXW);
  SACB_REQUIRE(d->n % 4 == 0 && d->nseg > 0 && d->nseg <= 4096, "sacb_allreduce_sgd: n %% 4 != 0 or bad nseg");
  P2PArgs a;
  for (int p = 0; p < P2P_MAXW; ++p) {
    a.grads[p] = p < d->world ? d->grads[p] : nullptr;
    a.params[p] = p < d->world ? d->params[p] : nullptr;
    a.flags[p] = p < d->world ? d->flags[p] : nullptr;
    SACB_REQUIRE(p >= d->world || (a.grads[p] && a.params[p] && a.flags[p]), "sacb_allreduce_sgd: NULL peer pointer");
  }
  a.mc_grads = d->mc_grads; a.mc_params = d->mc_params;
  SACB_REQUIRE((d->mc_grads == nullptr) == (d->mc_params == nullptr), "sacb_allreduce_sgd: mc_grads and mc_params go together");
  a.mom = d->mom; a.seg_ranges = d->seg_ranges; a.seg_lr = d->seg_lr; a.seg_wd = d->seg_wd;
  a.nseg = d->nseg; a.world = d->world; a.rank = d->rank; a.first = d->first_step;
  const long long nvec = d->n / 4, per = (nvec + d->world - 1) / d->world;
  a.vec_lo = per * d->rank < nvec ? per * d->rank : nvec;
  a.vec_hi = a.vec_lo + per < nvec ? a.vec_lo + per : nvec;
  a.mu = d->momentum; a.inv_world = 1.f / (float)d->world;
  static double timeout_s = -1.0;
  if (timeout_s < 0) {
    const char* e = getenv("SACB_P2P_TIMEOUT_S");
    timeout_s = (e && atof(e) > 0) ? atof(e) : 600.0;
  }
  a.spin_limit = (long long)(timeout_s * SPIN_CLOCK_HZ);
  int dev = 0, sms = 148;
  SACB_CHECK_CUDA(cudaGetDevice(&dev));
  SACB_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const size_t smem = sizeof(int64_t) * 2 * d->nseg;
  static bool attr_set = false;
  if (!attr_set && smem > 48 * 1024) {
    SACB_CHECK_CUDA(cudaFuncSetAttribute(allreduce_sgd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SACB_CHECK_CUDA(cudaFuncSetAttribute(allreduce_sgd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  if (a.mc_grads) allreduce_sgd_kernel<true><<<2 * sms, 512, smem, ST>>>(a);
  else allreduce_sgd_kernel<false><<<2 * sms, 512, smem, ST>>>(a);
  g_launches++;
  SACB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
