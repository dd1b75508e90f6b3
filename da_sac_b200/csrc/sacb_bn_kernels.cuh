// Device code of the training-mode BN kernels (see sacb_bn.cu for the entry points and the reference lines they replace).
// Kept in a header so that tests/cpu_emul can compile the SAME kernel bodies for the host (one std::thread per CUDA thread,
// std::barrier for __syncthreads) and check their indexing / arithmetic in the GPU-less build container.
#pragma once
#ifndef SACB_HOST_EMUL
#include "sacb_common.cuh"
#endif

namespace sacb {

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

}  // namespace sacb
