#line 1 "../../da_sac_b200/csrc/sacb_bn.cu"
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
#include <atomic>
// (sacb_common.cuh: see cuda_emul.h)
#include "../../include/sacb.h"

namespace sacb {

extern std::atomic<long long> g_launches;

constexpr int BN_ROWS = 512;          // rows swept by one block of the moment kernels

SACB_DEVINL void bn_unpack8(const uint4& u, float* f) {
  f[0] = bf16_bits_to_float(u.x & 0xFFFF); f[1] = bf16_bits_to_float(u.x >> 16);
  f[2] = bf16_bits_to_float(u.y & 0xFFFF); f[3] = bf16_bits_to_float(u.y >> 16);
  f[4] = bf16_bits_to_float(u.z & 0xFFFF); f[5] = bf16_bits_to_float(u.z >> 16);
  f[6] = bf16_bits_to_float(u.w & 0xFFFF); f[7] = bf16_bits_to_float(u.w >> 16);
}
SACB_DEVINL void bn_load8(const uint4* hi, const uint4* lo, size_t i, float* v) {
  float fh[8], fl[8];
  bn_unpack8(__ldg(hi + i), fh); bn_unpack8(__ldg(lo + i), fl);
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = fh[j] + fl[j];
}
SACB_DEVINL uint32_t bn_pack2(uint16_t a, uint16_t b) { return (uint32_t)a | ((uint32_t)b << 16); }
SACB_DEVINL void bn_store8(uint4* hi, uint4* lo, size_t i, const float* v) {
  uint16_t h[8], l[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    h[j] = float_to_bf16_bits(v[j]);
    l[j] = float_to_bf16_bits(v[j] - bf16_bits_to_float(h[j]));
  }
  hi[i] = make_uint4(bn_pack2(h[0], h[1]), bn_pack2(h[2], h[3]), bn_pack2(h[4], h[5]), bn_pack2(h[6], h[7]));
  lo[i] = make_uint4(bn_pack2(l[0], l[1]), bn_pack2(l[2], l[3]), bn_pack2(l[4], l[5]), bn_pack2(l[6], l[7]));
}

// MODE 0 (forward):  s1 = sum_m a,  s2 = sum_m a^2            (a = z, the conv output)
// MODE 1 (backward): s1 = sum_m a,  s2 = sum_m a * xhat(z)    (a = g, the gradient at the BN output; xhat = (z - mean) * invstd)
// block = (TX 8-channel vectors) x (TY row lanes), like colsum_kernel; partials[blockIdx.y][2][C] in double.
template <int MODE>
__global__ void __launch_bounds__(256)
bn_moments_kernel(const uint16_t* __restrict__ a_hi, const uint16_t* __restrict__ a_lo, const uint16_t* __restrict__ z_hi,
                  const uint16_t* __restrict__ z_lo, const float* __restrict__ mean, const float* __restrict__ invstd,
                  double* __restrict__ partials, long long M, int C) {
  const int TX = blockDim.x, TY = blockDim.y;
  const int cv = blockIdx.x * TX + threadIdx.x;
  const long long r0 = (long long)blockIdx.y * BN_ROWS;
  const long long r1 = r0 + BN_ROWS < M ? r0 + BN_ROWS : M;
  double d1[8], d2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { d1[j] = 0.0; d2[j] = 0.0; }
  if (cv < C / 8) {
    const uint4* ah = reinterpret_cast<const uint4*>(a_hi) + cv;
    const uint4* al = reinterpret_cast<const uint4*>(a_lo) + cv;
    const uint4* zh = reinterpret_cast<const uint4*>(z_hi) + cv;
    const uint4* zl = reinterpret_cast<const uint4*>(z_lo) + cv;
    const size_t stride = (size_t)C / 8;
    float mu[8], is[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { mu[j] = MODE ? mean[cv * 8 + j] : 0.f; is[j] = MODE ? invstd[cv * 8 + j] : 0.f; }
    // MODE 0: every element enters the double accumulators directly (the product of two fp32 values is exact in double), so
    // that var = E[z^2] - E[z]^2 keeps its digits when |mean| >> std; an fp32 partial sum of squares would already have lost
    // them (found by tests/test_emul_bn_cpu.py::test_single_row: var came out as 1e-7 z^2 instead of 0).  2 DP operations
    // and one conversion per 4 bytes read stay below the HBM time on a B200.
    // MODE 1: no cancellation downstream; groups of up to 4 rows are summed in fp32, the groups in double.
    for (long long r = r0 + threadIdx.y; r < r1; r += 4 * TY) {
      float f1[8], f2[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) { f1[j] = 0.f; f2[j] = 0.f; }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long rr = r + (long long)u * TY;
        if (rr < r1) {
          float a[8];
          bn_load8(ah, al, (size_t)rr * stride, a);
          if (MODE == 0) {
#pragma unroll
            for (int j = 0; j < 8; ++j) { const double da = (double)a[j]; d1[j] += da; d2[j] = fma(da, da, d2[j]); }
          } else {
            float z[8];
            bn_load8(zh, zl, (size_t)rr * stride, z);
#pragma unroll
            for (int j = 0; j < 8; ++j) { f1[j] += a[j]; f2[j] = fmaf(a[j], (z[j] - mu[j]) * is[j], f2[j]); }
          }
        }
      }
      if (MODE == 1) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { d1[j] += (double)f1[j]; d2[j] += (double)f2[j]; }
      }
    }
  }
  __shared__ double red[2][256][8];
  const int t = threadIdx.y * TX + threadIdx.x;
#pragma unroll
  for (int j = 0; j < 8; ++j) { red[0][t][j] = d1[j]; red[1][t][j] = d2[j]; }
  __syncthreads();
  if (threadIdx.y == 0 && cv < C / 8) {
    double* dst = partials + (size_t)blockIdx.y * 2 * C;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      double s1 = 0.0, s2 = 0.0;
      for (int y = 0; y < TY; ++y) { s1 += red[0][y * TX + threadIdx.x][j]; s2 += red[1][y * TX + threadIdx.x][j]; }
      dst[cv * 8 + j] = s1;
      dst[C + cv * 8 + j] = s2;
    }
  }
}

// sums[2][C] = sum over the row blocks, in block order
__global__ void bn_moments_reduce_kernel(const double* __restrict__ partials, int nblk, int C, double* __restrict__ sums) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * C) return;
  double s = 0.0;
  for (int b = 0; b < nblk; ++b) s += partials[(size_t)b * 2 * C + i];
  sums[i] = s;
}

// batch statistics -> (mean, invstd, scale = gamma * invstd) and the running-statistics update of nn.BatchNorm in training
// mode: running = (1 - momentum) * running + momentum * batch, with the UNBIASED batch variance (count / (count - 1)).
__global__ void bn_train_finalize_kernel(const double* __restrict__ sums, double count, const float* __restrict__ gamma,
                                         float eps, float momentum, float* __restrict__ running_mean,
                                         float* __restrict__ running_var, float* __restrict__ mean, float* __restrict__ invstd,
                                         float* __restrict__ scale, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double mu = sums[c] / count;
  double var = sums[C + c] / count - mu * mu;
  if (var < 0.0) var = 0.0;
  const float is = (float)(1.0 / sqrt(var + (double)eps));
  mean[c] = (float)mu;
  invstd[c] = is;
  scale[c] = gamma[c] * is;
  if (running_mean) {
    const double unbiased = count > 1.0 ? var * (count / (count - 1.0)) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mu;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
  }
}

// y = [relu]( (z - mean) * scale + beta [+ residual] )
__global__ void __launch_bounds__(256)
bn_apply_kernel(const uint16_t* __restrict__ z_hi, const uint16_t* __restrict__ z_lo, const float* __restrict__ mean,
                const float* __restrict__ scale, const float* __restrict__ beta, const uint16_t* __restrict__ res_hi,
                const uint16_t* __restrict__ res_lo, int relu, uint16_t* __restrict__ y_hi, uint16_t* __restrict__ y_lo,
                size_t nvec, int C) {
  const int cvs = C / 8;
  const uint4* zh = reinterpret_cast<const uint4*>(z_hi); const uint4* zl = reinterpret_cast<const uint4*>(z_lo);
  const uint4* rh = reinterpret_cast<const uint4*>(res_hi); const uint4* rl = reinterpret_cast<const uint4*>(res_lo);
  uint4* yh = reinterpret_cast<uint4*>(y_hi); uint4* yl = reinterpret_cast<uint4*>(y_lo);
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < nvec; t += (size_t)gridDim.x * blockDim.x) {
    const int c0 = (int)(t % cvs) * 8;
    float v[8];
    bn_load8(zh, zl, t, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = fmaf(v[j] - mean[c0 + j], scale[c0 + j], beta[c0 + j]);
    if (res_hi) {
      float r[8];
      bn_load8(rh, rl, t, r);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] += r[j];
    }
    if (relu) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
    }
    bn_store8(yh, yl, t, v);
  }
}

// d gamma = sum g * xhat, d beta = sum g over THIS rank's batch (DDP averages parameter gradients over the ranks afterwards);
// coef[0][c] = gamma * invstd, coef[1][c] = mean(g), coef[2][c] = mean(g * xhat) over the GLOBAL batch (SyncBatchNorm)
__global__ void bn_bwd_finalize_kernel(const double* __restrict__ sums_local, const double* __restrict__ sums_global,
                                       double count_global, const float* __restrict__ gamma, const float* __restrict__ invstd,
                                       float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ coef, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  if (dbeta) dbeta[c] = (float)sums_local[c];
  if (dgamma) dgamma[c] = (float)sums_local[C + c];
  coef[c] = gamma[c] * invstd[c];
  coef[C + c] = (float)(sums_global[c] / count_global);
  coef[2 * C + c] = (float)(sums_global[C + c] / count_global);
}

// dz = gamma * invstd * (g - mean(g) - xhat * mean(g * xhat));  may run in place (dz == g)
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(const uint16_t* g_hi, const uint16_t* g_lo, const uint16_t* __restrict__ z_hi,
                    const uint16_t* __restrict__ z_lo, const float* __restrict__ mean, const float* __restrict__ invstd,
                    const float* __restrict__ coef, uint16_t* dz_hi, uint16_t* dz_lo, size_t nvec, int C) {
  const int cvs = C / 8;
  const uint4* gh = reinterpret_cast<const uint4*>(g_hi); const uint4* gl = reinterpret_cast<const uint4*>(g_lo);
  const uint4* zh = reinterpret_cast<const uint4*>(z_hi); const uint4* zl = reinterpret_cast<const uint4*>(z_lo);
  uint4* oh = reinterpret_cast<uint4*>(dz_hi); uint4* ol = reinterpret_cast<uint4*>(dz_lo);
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < nvec; t += (size_t)gridDim.x * blockDim.x) {
    const int c0 = (int)(t % cvs) * 8;
    float g[8], z[8];
    {   // plain loads: g may alias the output
      float fh[8], fl[8];
      bn_unpack8(gh[t], fh); bn_unpack8(gl[t], fl);
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] = fh[j] + fl[j];
    }
    bn_load8(zh, zl, t, z);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = c0 + j;
      const float xhat = (z[j] - mean[c]) * invstd[c];
      g[j] = coef[c] * (g[j] - coef[C + c] - xhat * coef[2 * C + c]);
    }
    bn_store8(oh, ol, t, g);
  }
}

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
    cuda_emul::run_grid("bn_moments_kernel", grid, block, 0, true, [&]() { bn_moments_kernel<0>((const uint16_t*)a_hi, (const uint16_t*)a_lo, (const uint16_t*)a_hi,
                                                 (const uint16_t*)a_lo, nullptr, nullptr, partials, (long long)M, C); });
  else
    cuda_emul::run_grid("bn_moments_kernel", grid, block, 0, true, [&]() { bn_moments_kernel<1>((const uint16_t*)a_hi, (const uint16_t*)a_lo, (const uint16_t*)z_hi,
                                                 (const uint16_t*)z_lo, mean, invstd, partials, (long long)M, C); });
  LAUNCHED();
  cuda_emul::run_grid("bn_moments_reduce_kernel", (2 * C + 127) / 128, 128, 0, false, [&]() { bn_moments_reduce_kernel(partials, nblk, C, sums); });
  LAUNCHED();
  return 0;
}

extern "C" int sacb_bn_train_finalize(const double* sums, double count, const float* gamma, float eps, float momentum,
                                      float* running_mean, float* running_var, float* mean, float* invstd, float* scale,
                                      int C, void* stream) {
  SACB_REQUIRE(sums && gamma && mean && invstd && scale && count >= 1.0, "sacb_bn_train_finalize: bad arguments");
  SACB_REQUIRE((running_mean == nullptr) == (running_var == nullptr), "sacb_bn_train_finalize: running_mean and running_var go together");
  cuda_emul::run_grid("bn_train_finalize_kernel", (C + 127) / 128, 128, 0, false, [&]() { bn_train_finalize_kernel(sums, count, gamma, eps, momentum, running_mean, running_var, mean,
                                                           invstd, scale, C); });
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
  cuda_emul::run_grid("bn_apply_kernel", bn_grid(nvec, 256), 256, 0, false, [&]() { bn_apply_kernel((const uint16_t*)z_hi, (const uint16_t*)z_lo, mean, scale, beta,
                                                     (const uint16_t*)res_hi, (const uint16_t*)res_lo, relu, (uint16_t*)y_hi,
                                                     (uint16_t*)y_lo, nvec, C); });
  LAUNCHED();
  return 0;
}

extern "C" int sacb_bn_bwd_finalize(const double* sums_local, const double* sums_global, double count_global,
                                    const float* gamma, const float* invstd, float* dgamma, float* dbeta, float* coef, int C,
                                    void* stream) {
  SACB_REQUIRE(sums_local && sums_global && gamma && invstd && coef && count_global >= 1.0, "sacb_bn_bwd_finalize: bad arguments");
  cuda_emul::run_grid("bn_bwd_finalize_kernel", (C + 127) / 128, 128, 0, false, [&]() { bn_bwd_finalize_kernel(sums_local, sums_global, count_global, gamma, invstd, dgamma, dbeta,
                                                         coef, C); });
  LAUNCHED();
  return 0;
}

extern "C" int sacb_bn_bwd_apply(const void* g_hi, const void* g_lo, const void* z_hi, const void* z_lo, const float* mean,
                                 const float* invstd, const float* coef, void* dz_hi, void* dz_lo, int64_t M, int C,
                                 void* stream) {
  SACB_REQUIRE(C % 8 == 0 && M > 0, "sacb_bn_bwd_apply: C %% 8 == 0 and M > 0 required");
  SACB_REQUIRE(g_hi && g_lo && z_hi && z_lo && mean && invstd && coef && dz_hi && dz_lo, "sacb_bn_bwd_apply: null argument");
  const size_t nvec = (size_t)M * (C / 8);
  cuda_emul::run_grid("bn_bwd_apply_kernel", bn_grid(nvec, 256), 256, 0, false, [&]() { bn_bwd_apply_kernel((const uint16_t*)g_hi, (const uint16_t*)g_lo, (const uint16_t*)z_hi,
                                                         (const uint16_t*)z_lo, mean, invstd, coef, (uint16_t*)dz_hi,
                                                         (uint16_t*)dz_lo, nvec, C); });
  LAUNCHED();
  return 0;
}
