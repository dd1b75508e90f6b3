// Training-mode batch normalisation for the ABN baseline (cfg.MODEL.BASELINE = True).
//
// In the SAC path every BN layer is frozen and folded into the GEMM epilogue.  The ABN baseline
// (/root/reference/models/__init__.py:29 -> freeze_bn = False; train.py:113-138,281-289) instead trains the backbone on the
// source domain with nn.SyncBatchNorm in training mode and lets the target domain move the running statistics.  A conv unit
// then is   z = conv(x)  ->  batch moments of z  ->  y = relu(gamma * (z - mean) * invstd + beta (+ residual))   and backward
//           g = dL/dy_pre  ->  sum g, sum g*xhat  ->  dz = gamma * invstd * (g - mean(g) - xhat * mean(g * xhat)).
// The convolutions stay on the tcgen05 GEMM kernels (raw epilogue); what is here are the streaming kernels around them.
// All activations are bf16 split planes [M][C] (include/sacb.h), C % 8 == 0; 16-byte accesses throughout.
//
// Moments are deterministic (no atomics): every block writes its partial sums, a second kernel adds them in block order, in
// double precision -- E[z^2] - E[z]^2 cancels badly in fp32 when |mean| >> std.
// Replaces: torch.nn.SyncBatchNorm.forward / backward in training mode (deeplabv2.py:15,28-31,60-71,124,149,183) and its
// running-statistics update.
// Launches are spelled SACB_LAUNCH(kernel, grid, block, smem, stream, args...) (sacb_common.cuh: kernel<<<...>>>(args)) so that
// tests/cpu_emul can compile this file, entry points included, for the host (SACB_HOST_EMUL: test infrastructure only).
#include <atomic>
#ifndef SACB_HOST_EMUL
#include "sacb_common.cuh"
#endif
#include "sacb_bn_kernels.cuh"
#include "../../include/sacb.h"

namespace sacb {

extern std::atomic<long long> g_launches;

static inline int bn_grid(size_t n, int block) {
  size_t g = (n + block - 1) / block;
  const size_t cap = 148 * 16;
  return (int)(g < cap ? (g ? g : 1) : cap);
}

}  // namespace sacb

using namespace sacb;
#define ST ((cudaStream_t)stream)
#define LAUNCHED() do { g_launches++; SACB_CHECK_CUDA(cudaGetLastError()); } while (0)

extern "C" size_t sacb_bn_moments_partial_elems(int64_t M, int C) {
  return (size_t)((M + BN_ROWS - 1) / BN_ROWS) * 2 * (size_t)C;
}

extern "C" int sacb_bn_moments(const void* a_hi, const void* a_lo, const void* z_hi, const void* z_lo, const float* mean,
                               const float* invstd, int mode, int64_t M, int C, double* partials, double* sums,
                               void* stream) {
  SACB_REQUIRE(C % 8 == 0 && M > 0, "sacb_bn_moments: C %% 8 == 0 and M > 0 required (M=%lld, C=%d)", (long long)M, C);
  SACB_REQUIRE(mode == 0 || mode == 1, "sacb_bn_moments: mode must be 0 (forward) or 1 (backward)");
  SACB_REQUIRE(a_hi && a_lo && partials && sums, "sacb_bn_moments: null argument");
  SACB_REQUIRE(mode == 0 || (z_hi && z_lo && mean && invstd), "sacb_bn_moments: mode 1 needs z planes, mean and invstd");
  const int cvs = C / 8;
  const int tx = cvs >= 32 ? 32 : (cvs >= 16 ? 16 : 8);
  dim3 block(tx, 256 / tx);
  const int nblk = (int)((M + BN_ROWS - 1) / BN_ROWS);
  dim3 grid((cvs + tx - 1) / tx, nblk);
  if (mode == 0)
    SACB_LAUNCH(bn_moments_kernel<0>, grid, block, 0, ST, (const uint16_t*)a_hi, (const uint16_t*)a_lo, (const uint16_t*)a_hi,
                                                 (const uint16_t*)a_lo, nullptr, nullptr, partials, (long long)M, C);
  else
    SACB_LAUNCH(bn_moments_kernel<1>, grid, block, 0, ST, (const uint16_t*)a_hi, (const uint16_t*)a_lo, (const uint16_t*)z_hi,
                                                 (const uint16_t*)z_lo, mean, invstd, partials, (long long)M, C);
  LAUNCHED();
  SACB_LAUNCH(bn_moments_reduce_kernel, (2 * C + 127) / 128, 128, 0, ST, partials, nblk, C, sums);
  LAUNCHED();
  return 0;
}

extern "C" int sacb_bn_train_finalize(const double* sums, double count, const float* gamma, float eps, float momentum,
                                      float* running_mean, float* running_var, float* mean, float* invstd, float* scale,
                                      int C, void* stream) {
  SACB_REQUIRE(sums && gamma && mean && invstd && scale && count >= 1.0, "sacb_bn_train_finalize: bad arguments");
  SACB_REQUIRE((running_mean == nullptr) == (running_var == nullptr), "sacb_bn_train_finalize: running_mean and running_var go together");
  SACB_LAUNCH(bn_train_finalize_kernel, (C + 127) / 128, 128, 0, ST, sums, count, gamma, eps, momentum, running_mean, running_var, mean,
                                                           invstd, scale, C);
  LAUNCHED();
  return 0;
}

extern "C" int sacb_bn_apply(const void* z_hi, const void* z_lo, const float* mean, const float* scale, const float* beta,
                             const void* res_hi, const void* res_lo, int relu, void* y_hi, void* y_lo, int64_t M, int C,
                             void* stream) {
  SACB_REQUIRE(C % 8 == 0 && M > 0, "sacb_bn_apply: C %% 8 == 0 and M > 0 required");
  SACB_REQUIRE(z_hi && z_lo && mean && scale && beta && y_hi && y_lo, "sacb_bn_apply: null argument");
  SACB_REQUIRE((res_hi == nullptr) == (res_lo == nullptr), "sacb_bn_apply: res_hi and res_lo go together");
  const size_t nvec = (size_t)M * (C / 8);
  SACB_LAUNCH(bn_apply_kernel, bn_grid(nvec, 256), 256, 0, ST, (const uint16_t*)z_hi, (const uint16_t*)z_lo, mean, scale, beta,
                                                     (const uint16_t*)res_hi, (const uint16_t*)res_lo, relu, (uint16_t*)y_hi,
                                                     (uint16_t*)y_lo, nvec, C);
  LAUNCHED();
  return 0;
}

extern "C" int sacb_bn_bwd_finalize(const double* sums_local, const double* sums_global, double count_global,
                                    const float* gamma, const float* invstd, float* dgamma, float* dbeta, float* coef, int C,
                                    void* stream) {
  SACB_REQUIRE(sums_local && sums_global && gamma && invstd && coef && count_global >= 1.0, "sacb_bn_bwd_finalize: bad arguments");
  SACB_LAUNCH(bn_bwd_finalize_kernel, (C + 127) / 128, 128, 0, ST, sums_local, sums_global, count_global, gamma, invstd, dgamma, dbeta,
                                                         coef, C);
  LAUNCHED();
  return 0;
}

extern "C" int sacb_bn_bwd_apply(const void* g_hi, const void* g_lo, const void* z_hi, const void* z_lo, const float* mean,
                                 const float* invstd, const float* coef, void* dz_hi, void* dz_lo, int64_t M, int C,
                                 void* stream) {
  SACB_REQUIRE(C % 8 == 0 && M > 0, "sacb_bn_bwd_apply: C %% 8 == 0 and M > 0 required");
  SACB_REQUIRE(g_hi && g_lo && z_hi && z_lo && mean && invstd && coef && dz_hi && dz_lo, "sacb_bn_bwd_apply: null argument");
  const size_t nvec = (size_t)M * (C / 8);
  SACB_LAUNCH(bn_bwd_apply_kernel, bn_grid(nvec, 256), 256, 0, ST, (const uint16_t*)g_hi, (const uint16_t*)g_lo, (const uint16_t*)z_hi,
                                                         (const uint16_t*)z_lo, mean, invstd, coef, (uint16_t*)dz_hi,
                                                         (uint16_t*)dz_lo, nvec, C);
  LAUNCHED();
  return 0;
}
