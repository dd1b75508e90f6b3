// Bandwidth-bound glue kernels around the conv GEMMs: stem conv (3 input channels, CUDA cores), max-pool,
// ReLU-mask / split / scatter, column sums, weight re-layout, BN folding and BN-gradient finalisation.
// All activations are bf16 split planes in NHWC (see include/sacb.h).
#include <atomic>
#include "sacb_common.cuh"
#include "../../include/sacb.h"

namespace sacb {
extern std::atomic<long long> g_launches;

SACB_DEVINL float ld_split(const uint16_t* hi, const uint16_t* lo, size_t i) {
  return bf16_bits_to_float(hi[i]) + bf16_bits_to_float(lo[i]);
}
SACB_DEVINL void st_split(uint16_t* hi, uint16_t* lo, size_t i, float v) {
  uint16_t h = float_to_bf16_bits(v);
  hi[i] = h;
  lo[i] = float_to_bf16_bits(v - bf16_bits_to_float(h));
}
SACB_DEVINL void unpack8f(const uint4& u, float* f) {
  f[0] = bf16_bits_to_float(u.x & 0xFFFF); f[1] = bf16_bits_to_float(u.x >> 16);
  f[2] = bf16_bits_to_float(u.y & 0xFFFF); f[3] = bf16_bits_to_float(u.y >> 16);
  f[4] = bf16_bits_to_float(u.z & 0xFFFF); f[5] = bf16_bits_to_float(u.z >> 16);
  f[6] = bf16_bits_to_float(u.w & 0xFFFF); f[7] = bf16_bits_to_float(u.w >> 16);
}
SACB_DEVINL void split8(const float* v, uint4& h, uint4& l) {
  uint32_t ph[4], pl[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint16_t h0 = float_to_bf16_bits(v[2 * j]), h1 = float_to_bf16_bits(v[2 * j + 1]);
    uint16_t l0 = float_to_bf16_bits(v[2 * j] - bf16_bits_to_float(h0));
    uint16_t l1 = float_to_bf16_bits(v[2 * j + 1] - bf16_bits_to_float(h1));
    ph[j] = (uint32_t)h0 | ((uint32_t)h1 << 16);
    pl[j] = (uint32_t)l0 | ((uint32_t)l1 << 16);
  }
  h = make_uint4(ph[0], ph[1], ph[2], ph[3]);
  l = make_uint4(pl[0], pl[1], pl[2], pl[3]);
}

// ------------------------------------------------------------------------------------------------
// Stem: conv 7x7 stride 2 pad 3, 3 -> 64, + BN affine + ReLU   (deeplabv2.py:160-163)
// block = 8x16 output pixels, 256 threads; thread = 2 horizontally adjacent pixels x 16 channels.
// ------------------------------------------------------------------------------------------------
constexpr int ST_TW = 16, ST_TH = 8;
constexpr int ST_PW = ST_TW * 2 + 5, ST_PH = ST_TH * 2 + 5;   // input patch 37x37

__global__ void __launch_bounds__(256)
stem_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ scale,
                const float* __restrict__ shift, uint16_t* __restrict__ out_hi, uint16_t* __restrict__ out_lo,
                int N, int H, int W, int P, int Q) {
  __shared__ float sw[147 * 64];                 // [tap][k]
  __shared__ float sx[3][ST_PH][ST_PW + 1];
  const int tid = threadIdx.x;
  const int n = blockIdx.z, p0 = blockIdx.y * ST_TH, q0 = blockIdx.x * ST_TW;
  for (int i = tid; i < 147 * 64; i += 256) {
    const int k = i & 63, tap = i >> 6;            // tap = (c*7 + r)*7 + s, w is [k][c][r][s]
    sw[i] = w[k * 147 + tap];
  }
  const int h0 = p0 * 2 - 3, w0 = q0 * 2 - 3;
  for (int i = tid; i < 3 * ST_PH * ST_PW; i += 256) {
    const int c = i / (ST_PH * ST_PW), rem = i - c * ST_PH * ST_PW;
    const int yy = rem / ST_PW, xx = rem - yy * ST_PW;
    const int hh = h0 + yy, ww = w0 + xx;
    sx[c][yy][xx] = (hh >= 0 && hh < H && ww >= 0 && ww < W) ? x[((size_t)(n * 3 + c) * H + hh) * W + ww] : 0.f;
  }
  __syncthreads();
  const int cg = tid & 3;            // 16-channel group
  constexpr int NPX = 2;
  const int pg = tid >> 2;           // 64 groups of 2 pixels: row = pg / 8, col group = pg % 8
  const int py = pg >> 3, px0 = (pg & 7) * NPX;
  float acc[NPX][16];
#pragma unroll
  for (int i = 0; i < NPX; ++i)
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[i][j] = 0.f;
  for (int c = 0; c < 3; ++c)
    for (int r = 0; r < 7; ++r) {
      const float* xr = &sx[c][py * 2 + r][px0 * 2];
#pragma unroll
      for (int s = 0; s < 7; ++s) {
        const float4* wp = reinterpret_cast<const float4*>(&sw[((c * 7 + r) * 7 + s) * 64 + cg * 16]);
        float wv[16];
#pragma unroll
        for (int j = 0; j < 4; ++j) { float4 t = wp[j]; wv[4 * j] = t.x; wv[4 * j + 1] = t.y; wv[4 * j + 2] = t.z; wv[4 * j + 3] = t.w; }
#pragma unroll
        for (int i = 0; i < NPX; ++i) {
          const float xv = xr[i * 2 + s];
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[i][j] = fmaf(xv, wv[j], acc[i][j]);
        }
      }
    }
  const int p = p0 + py;
  if (p >= P) return;
#pragma unroll
  for (int i = 0; i < NPX; ++i) {
    const int q = q0 + px0 + i;
    if (q >= Q) continue;
    float v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int k = cg * 16 + j;
      v[j] = fmaxf(fmaf(acc[i][j], scale[k], shift[k]), 0.f);
    }
    const size_t o = (((size_t)n * P + p) * Q + q) * 64 + cg * 16;
    uint4 h, l;
    split8(v, h, l);
    *reinterpret_cast<uint4*>(out_hi + o) = h; *reinterpret_cast<uint4*>(out_lo + o) = l;
    split8(v + 8, h, l);
    *reinterpret_cast<uint4*>(out_hi + o + 8) = h; *reinterpret_cast<uint4*>(out_lo + o + 8) = l;
  }
}

// Stem filter gradient: dw[k][c][r][s] += sum_{n,p,q} g[n,p,q,k] * x[n,c,2p-3+r,2q-3+s]
// persistent blocks loop over 8x16 output tiles; thread = (k, group of taps); one atomic per output per block.
constexpr int SW_TW = 16, SW_TH = 8;
constexpr int SW_PW = SW_TW * 2 + 5, SW_PH = SW_TH * 2 + 5;   // 37 x 21
constexpr int SW_TAPS_PER_GROUP = 37;                          // 147 taps over 4 groups (last group 36)

__global__ void __launch_bounds__(256)
stem_wgrad_kernel(const float* __restrict__ x, const uint16_t* __restrict__ g_hi, const uint16_t* __restrict__ g_lo,
                  float* __restrict__ dw, int N, int H, int W, int P, int Q) {
  __shared__ float sg[SW_TH * SW_TW][64];
  __shared__ float sx[3 * SW_PH * SW_PW];
  const int tid = threadIdx.x;
  const int k = tid & 63, grp = tid >> 6;
  const int tap0 = grp * SW_TAPS_PER_GROUP;
  const int ntaps = min(SW_TAPS_PER_GROUP, 147 - tap0);
  float acc[SW_TAPS_PER_GROUP];
#pragma unroll
  for (int i = 0; i < SW_TAPS_PER_GROUP; ++i) acc[i] = 0.f;
  const int tiles_x = (Q + SW_TW - 1) / SW_TW, tiles_y = (P + SW_TH - 1) / SW_TH;
  const int total = N * tiles_x * tiles_y;
  for (int t = blockIdx.x; t < total; t += gridDim.x) {
    const int n = t / (tiles_x * tiles_y);
    const int rem = t - n * tiles_x * tiles_y;
    const int p0 = (rem / tiles_x) * SW_TH, q0 = (rem % tiles_x) * SW_TW;
    __syncthreads();
    for (int i = tid; i < SW_TH * SW_TW * 64; i += 256) {
      const int kk = i & 63, px = i >> 6;
      const int p = p0 + px / SW_TW, q = q0 + px % SW_TW;
      float v = 0.f;
      if (p < P && q < Q) v = ld_split(g_hi, g_lo, (((size_t)n * P + p) * Q + q) * 64 + kk);
      sg[px][kk] = v;
    }
    const int h0 = p0 * 2 - 3, w0 = q0 * 2 - 3;
    for (int i = tid; i < 3 * SW_PH * SW_PW; i += 256) {
      const int c = i / (SW_PH * SW_PW), r2 = i - c * SW_PH * SW_PW;
      const int yy = r2 / SW_PW, xx = r2 - yy * SW_PW;
      const int hh = h0 + yy, ww = w0 + xx;
      sx[i] = (hh >= 0 && hh < H && ww >= 0 && ww < W) ? x[((size_t)(n * 3 + c) * H + hh) * W + ww] : 0.f;
    }
    __syncthreads();
    for (int px = 0; px < SW_TH * SW_TW; ++px) {
      const float gv = sg[px][k];
      const int py = px / SW_TW, pxx = px % SW_TW;
      const int base = (py * 2) * SW_PW + pxx * 2;
#pragma unroll
      for (int i = 0; i < SW_TAPS_PER_GROUP; ++i) {
        if (i < ntaps) {
          const int tap = tap0 + i;
          const int c = tap / 49, rs = tap - c * 49;
          const int r = rs / 7, s = rs - r * 7;
          acc[i] = fmaf(gv, sx[c * SW_PH * SW_PW + base + r * SW_PW + s], acc[i]);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < SW_TAPS_PER_GROUP; ++i)
    if (i < ntaps) atomicAdd(&dw[k * 147 + tap0 + i], acc[i]);
}

// ------------------------------------------------------------------------------------------------
// Stem on the tensor cores: explicit im2col of the 3-channel image (147 taps, zero-padded to 192 columns) so the
// 7x7 s2 conv becomes a 1x1 conv_gemm with C = 192 (and its filter gradient a plain conv_wgrad).
// ------------------------------------------------------------------------------------------------
// Generic for any 3-input-channel first conv: R x R taps, `stride`, `pad`; TAPS = 3*R*R columns padded to KP (multiple of 64).
// RT > 0: filter size known at compile time (7: ResNet stem, 3: VGG features.0) so the tap index arithmetic is constant-folded
template <int RT>
__global__ void __launch_bounds__(256)
stem_im2col_kernel(const float* __restrict__ x, uint16_t* __restrict__ a_hi, uint16_t* __restrict__ a_lo, int N, int H,
                   int W, int P, int Q, int R_rt, int stride, int pad, int KP) {
  const int R = RT > 0 ? RT : R_rt;
  const int RR = R * R, TAPS = 3 * RR;
  const size_t total = (size_t)N * P * Q * (KP / 8);
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int jv = (int)(t % (KP / 8));
    size_t pix = t / (KP / 8);
    const int q = (int)(pix % Q); pix /= Q;
    const int p = (int)(pix % P);
    const int n = (int)(pix / P);
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int j = jv * 8 + e;
      float val = 0.f;
      if (j < TAPS) {
        const int c = j / RR, rs = j - c * RR;
        const int r = rs / R, ss = rs - r * R;
        const int hh = stride * p - pad + r, ww = stride * q - pad + ss;
        if (hh >= 0 && hh < H && ww >= 0 && ww < W) val = __ldg(x + ((size_t)(n * 3 + c) * H + hh) * W + ww);
      }
      v[e] = val;
    }
    uint4 h, l;
    split8(v, h, l);
    reinterpret_cast<uint4*>(a_hi)[t] = h;
    reinterpret_cast<uint4*>(a_lo)[t] = l;
  }
}
// Shared-memory staged variant for compile-time (R, STRIDE): one block = 64 consecutive output pixels of one output row.
// The R input rows x (64*STRIDE + R - STRIDE) columns x 3 channels they read are staged once (the gather form above issues
// 8 scattered global loads per 16 output bytes and is bound by the load/store unit: 1.05 ms per launch for the ResNet
// stem, 5x off the 1.2 GB write roofline); the im2col rows are then written as coalesced 16-byte words.
template <int R, int STRIDE>
__global__ void __launch_bounds__(256)
stem_im2col_smem_kernel(const float* __restrict__ x, uint16_t* __restrict__ a_hi, uint16_t* __restrict__ a_lo, int H, int W,
                        int P, int Q, int pad, int KP) {
  constexpr int PIX = 64, RR = R * R, TAPS = 3 * RR;
  constexpr int IN_W = PIX * STRIDE + R - STRIDE;
  __shared__ float tile[3][R][IN_W + 1];
  const int q0 = blockIdx.x * PIX, p = blockIdx.y, n = blockIdx.z;
  const int h0 = STRIDE * p - pad, w0 = STRIDE * q0 - pad;
  for (int t = threadIdx.x; t < 3 * R * IN_W; t += 256) {
    const int c = t / (R * IN_W), rem = t - c * (R * IN_W);
    const int r = rem / IN_W, i = rem - r * IN_W;
    const int hh = h0 + r, ww = w0 + i;
    tile[c][r][i] = (hh >= 0 && hh < H && ww >= 0 && ww < W) ? __ldg(x + ((size_t)(n * 3 + c) * H + hh) * W + ww) : 0.f;
  }
  __syncthreads();
  const int vec_per_pix = KP / 8;
  const int npix = min(PIX, Q - q0);
  const size_t row0 = ((size_t)n * P + p) * Q + q0;                 // first im2col row of this block
  for (int it = threadIdx.x; it < npix * vec_per_pix; it += 256) {
    const int pl = it / vec_per_pix, jv = it - pl * vec_per_pix;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int j = jv * 8 + e;
      float val = 0.f;
      if (j < TAPS) {
        const int c = j / RR, rs = j - c * RR;
        const int r = rs / R, ss = rs - r * R;
        val = tile[c][r][pl * STRIDE + ss];
      }
      v[e] = val;
    }
    uint4 hq, lq;
    split8(v, hq, lq);
    const size_t o = (row0 + pl) * vec_per_pix + jv;
    reinterpret_cast<uint4*>(a_hi)[o] = hq;
    reinterpret_cast<uint4*>(a_lo)[o] = lq;
  }
}
// wf[k][j] (j < KP) from OIHW [K][TAPS]
__global__ void stem_pack_weight_kernel(const float* __restrict__ w, uint16_t* __restrict__ hi, uint16_t* __restrict__ lo,
                                        int K, int TAPS, int KP) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= K * KP) return;
  const int k = i / KP, j = i - k * KP;
  st_split(hi, lo, i, j < TAPS ? w[k * TAPS + j] : 0.f);
}
// dwraw[k][j<TAPS] = sum_split parts[split][k][j]  (parts rows are KP wide)
__global__ void stem_unpack_wgrad_kernel(const float* __restrict__ parts, int splits, float* __restrict__ dwraw, int K,
                                         int TAPS, int KP) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= K * TAPS) return;
  const int k = i / TAPS, j = i - k * TAPS;
  float s = 0.f;
  for (int sp = 0; sp < splits; ++sp) s += parts[((size_t)sp * K + k) * KP + j];
  dwraw[i] = s;
}

// MaxPool2d(2, 2) on split planes (torchvision vgg16 features.6/13/23); idx = argmax tap 0..3 for backward
__global__ void maxpool2_fwd_kernel(const uint16_t* __restrict__ in_hi, const uint16_t* __restrict__ in_lo,
                                    uint16_t* __restrict__ out_hi, uint16_t* __restrict__ out_lo,
                                    uint8_t* __restrict__ idx, int N, int H, int W, int C, int P, int Q) {
  const size_t total = (size_t)N * P * Q * (C / 8);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int cv = (int)(i % (C / 8));
    size_t t = i / (C / 8);
    const int q = (int)(t % Q); t /= Q;
    const int p = (int)(t % P);
    const int n = (int)(t / P);
    float best[8]; uint32_t bhh[8], bll[8]; int bi[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { best[j] = -INFINITY; bi[j] = 0; bhh[j] = 0; bll[j] = 0; }
#pragma unroll
    for (int tap = 0; tap < 4; ++tap) {
      const int hh = 2 * p + (tap >> 1), ww = 2 * q + (tap & 1);
      if (hh >= H || ww >= W) continue;
      const size_t o = (((size_t)n * H + hh) * W + ww) * C + cv * 8;
      const uint4 h = *reinterpret_cast<const uint4*>(in_hi + o);
      const uint4 l = *reinterpret_cast<const uint4*>(in_lo + o);
      float fh[8], fl[8];
      unpack8f(h, fh); unpack8f(l, fl);
      const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float v = fh[j] + fl[j];
        if (v > best[j]) {
          best[j] = v; bi[j] = tap;
          bhh[j] = (hw[j >> 1] >> ((j & 1) * 16)) & 0xFFFF;
          bll[j] = (lw[j >> 1] >> ((j & 1) * 16)) & 0xFFFF;
        }
      }
    }
    const size_t oo = (((size_t)n * P + p) * Q + q) * C + cv * 8;
    *reinterpret_cast<uint4*>(out_hi + oo) = make_uint4(bhh[0] | (bhh[1] << 16), bhh[2] | (bhh[3] << 16), bhh[4] | (bhh[5] << 16), bhh[6] | (bhh[7] << 16));
    *reinterpret_cast<uint4*>(out_lo + oo) = make_uint4(bll[0] | (bll[1] << 16), bll[2] | (bll[3] << 16), bll[4] | (bll[5] << 16), bll[6] | (bll[7] << 16));
    *reinterpret_cast<uint2*>(idx + oo) = make_uint2(bi[0] | (bi[1] << 8) | (bi[2] << 16) | (bi[3] << 24),
                                                     bi[4] | (bi[5] << 8) | (bi[6] << 16) | (bi[7] << 24));
  }
}
// g_in[n,h,w,c] = (idx[n,h/2,w/2,c] == tap(h,w)) ? g_out[n,h/2,w/2,c] : 0, masked by in_hi > 0 -> split planes
__global__ void maxpool2_bwd_kernel(const float* __restrict__ g_out, const uint8_t* __restrict__ idx,
                                    const uint16_t* __restrict__ in_hi, uint16_t* __restrict__ gin_hi,
                                    uint16_t* __restrict__ gin_lo, int N, int H, int W, int C, int P, int Q) {
  const size_t total = (size_t)N * H * W * (C / 8);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int cv = (int)(i % (C / 8));
    size_t t = i / (C / 8);
    const int w = (int)(t % W); t /= W;
    const int h = (int)(t % H);
    const int n = (int)(t / H);
    float acc[8], m[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    unpack8f(reinterpret_cast<const uint4*>(in_hi)[i], m);
    const int p = h >> 1, q = w >> 1;
    if (p < P && q < Q) {
      const size_t o = (((size_t)n * P + p) * Q + q) * C + cv * 8;
      const uint2 iv = *reinterpret_cast<const uint2*>(idx + o);
      const float4 g0 = *reinterpret_cast<const float4*>(g_out + o), g1 = *reinterpret_cast<const float4*>(g_out + o + 4);
      const float gv[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const int tap = ((h & 1) << 1) | (w & 1);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int id = ((j < 4 ? iv.x : iv.y) >> ((j & 3) * 8)) & 0xFF;
        if (id == tap && m[j] > 0.f) acc[j] = gv[j];
      }
    }
    uint4 hh, ll;
    split8(acc, hh, ll);
    reinterpret_cast<uint4*>(gin_hi)[i] = hh;
    reinterpret_cast<uint4*>(gin_lo)[i] = ll;
  }
}

// ------------------------------------------------------------------------------------------------
// MaxPool2d(3, stride 2, pad 1, ceil_mode=True)   (deeplabv2.py:126)
// ------------------------------------------------------------------------------------------------
__global__ void maxpool_fwd_kernel(const uint16_t* __restrict__ in_hi, const uint16_t* __restrict__ in_lo,
                                   uint16_t* __restrict__ out_hi, uint16_t* __restrict__ out_lo,
                                   uint8_t* __restrict__ idx, int N, int H, int W, int C, int P, int Q) {
  const size_t total = (size_t)N * P * Q * (C / 8);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int cv = (int)(i % (C / 8));
    size_t t = i / (C / 8);
    const int q = (int)(t % Q); t /= Q;
    const int p = (int)(t % P);
    const int n = (int)(t / P);
    float best[8]; uint4 bh, bl; int bi[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { best[j] = -INFINITY; bi[j] = 0; }
    uint32_t bhh[8], bll[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { bhh[j] = 0; bll[j] = 0; }
    for (int r = 0; r < 3; ++r) {
      const int hh = p * 2 - 1 + r;
      if (hh < 0 || hh >= H) continue;
      for (int s = 0; s < 3; ++s) {
        const int ww = q * 2 - 1 + s;
        if (ww < 0 || ww >= W) continue;
        const size_t o = (((size_t)n * H + hh) * W + ww) * C + cv * 8;
        const uint4 h = *reinterpret_cast<const uint4*>(in_hi + o);
        const uint4 l = *reinterpret_cast<const uint4*>(in_lo + o);
        float fh[8], fl[8];
        unpack8f(h, fh); unpack8f(l, fl);
        const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float v = fh[j] + fl[j];
          if (v > best[j]) {                      // first maximum in scan order wins (ATen max_pool2d)
            best[j] = v; bi[j] = r * 3 + s;
            bhh[j] = (hw[j >> 1] >> ((j & 1) * 16)) & 0xFFFF;
            bll[j] = (lw[j >> 1] >> ((j & 1) * 16)) & 0xFFFF;
          }
        }
      }
    }
    bh = make_uint4(bhh[0] | (bhh[1] << 16), bhh[2] | (bhh[3] << 16), bhh[4] | (bhh[5] << 16), bhh[6] | (bhh[7] << 16));
    bl = make_uint4(bll[0] | (bll[1] << 16), bll[2] | (bll[3] << 16), bll[4] | (bll[5] << 16), bll[6] | (bll[7] << 16));
    const size_t oo = (((size_t)n * P + p) * Q + q) * C + cv * 8;
    *reinterpret_cast<uint4*>(out_hi + oo) = bh;
    *reinterpret_cast<uint4*>(out_lo + oo) = bl;
    uint32_t i0 = bi[0] | (bi[1] << 8) | (bi[2] << 16) | (bi[3] << 24);
    uint32_t i1 = bi[4] | (bi[5] << 8) | (bi[6] << 16) | (bi[7] << 24);
    *reinterpret_cast<uint2*>(idx + oo) = make_uint2(i0, i1);
  }
}

__global__ void maxpool_bwd_kernel(const float* __restrict__ g_out, const uint8_t* __restrict__ idx,
                                   const uint16_t* __restrict__ in_hi, uint16_t* __restrict__ gin_hi,
                                   uint16_t* __restrict__ gin_lo, int N, int H, int W, int C, int P, int Q) {
  // thread = 8 channels of one input pixel; gathers from the (at most 4) pooling windows that cover it
  const size_t total = (size_t)N * H * W * (C / 8);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int cv = (int)(i % (C / 8));
    size_t t = i / (C / 8);
    const int w = (int)(t % W); t /= W;
    const int h = (int)(t % H);
    const int n = (int)(t / H);
    float acc[8], m[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    unpack8f(reinterpret_cast<const uint4*>(in_hi)[i], m);
    const int pl = h / 2, ph = (h + 1) / 2;         // windows p with 2p-1 <= h <= 2p+1
    const int ql = w / 2, qh = (w + 1) / 2;
    for (int p = pl; p <= ph; ++p) {
      if (p >= P) continue;
      const int r = h - (2 * p - 1);
      if (r < 0 || r > 2) continue;
      for (int q = ql; q <= qh; ++q) {
        if (q >= Q) continue;
        const int sft = w - (2 * q - 1);
        if (sft < 0 || sft > 2) continue;
        const size_t o = (((size_t)n * P + p) * Q + q) * C + cv * 8;
        const uint2 iv = *reinterpret_cast<const uint2*>(idx + o);
        const float4 g0 = *reinterpret_cast<const float4*>(g_out + o), g1 = *reinterpret_cast<const float4*>(g_out + o + 4);
        const float gv[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        const int tap = r * 3 + sft;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int id = ((j < 4 ? iv.x : iv.y) >> ((j & 3) * 8)) & 0xFF;
          if (id == tap) acc[j] += gv[j];
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = m[j] > 0.f ? acc[j] : 0.f;      // ReLU mask of the stem output
    uint4 hh, ll;
    split8(acc, hh, ll);
    reinterpret_cast<uint4*>(gin_hi)[i] = hh;
    reinterpret_cast<uint4*>(gin_lo)[i] = ll;
  }
}

// ------------------------------------------------------------------------------------------------
// elementwise: (a [+ b]) masked by forward activation > 0, written as split planes
// ------------------------------------------------------------------------------------------------
__global__ void add_mask_split_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                      const uint16_t* __restrict__ mask_hi, uint16_t* __restrict__ out_hi,
                                      uint16_t* __restrict__ out_lo, size_t n8) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
    float v[8];
    const float4 a0 = reinterpret_cast<const float4*>(a)[2 * i], a1 = reinterpret_cast<const float4*>(a)[2 * i + 1];
    v[0] = a0.x; v[1] = a0.y; v[2] = a0.z; v[3] = a0.w; v[4] = a1.x; v[5] = a1.y; v[6] = a1.z; v[7] = a1.w;
    if (b) {
      const float4 b0 = reinterpret_cast<const float4*>(b)[2 * i], b1 = reinterpret_cast<const float4*>(b)[2 * i + 1];
      v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w; v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
    }
    if (mask_hi) {
      float m[8];
      unpack8f(reinterpret_cast<const uint4*>(mask_hi)[i], m);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = m[j] > 0.f ? v[j] : 0.f;
    }
    uint4 h, l;
    split8(v, h, l);
    reinterpret_cast<uint4*>(out_hi)[i] = h;
    reinterpret_cast<uint4*>(out_lo)[i] = l;
  }
}

__global__ void scatter2_mask_split_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                           const uint16_t* __restrict__ mask_hi, uint16_t* __restrict__ out_hi,
                                           uint16_t* __restrict__ out_lo, int N, int H, int W, int C, int P, int Q) {
  const size_t total = (size_t)N * H * W * (C / 8);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int cv = (int)(i % (C / 8));
    size_t t = i / (C / 8);
    const int w = (int)(t % W); t /= W;
    const int h = (int)(t % H);
    const int n = (int)(t / H);
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 0.f;
    if (!(h & 1) && !(w & 1) && (h >> 1) < P && (w >> 1) < Q) {
      const size_t o = ((((size_t)n * P + (h >> 1)) * Q + (w >> 1)) * C + cv * 8) / 4;
      const float4 a0 = reinterpret_cast<const float4*>(a)[o], a1 = reinterpret_cast<const float4*>(a)[o + 1];
      v[0] = a0.x; v[1] = a0.y; v[2] = a0.z; v[3] = a0.w; v[4] = a1.x; v[5] = a1.y; v[6] = a1.z; v[7] = a1.w;
      if (b) {
        const float4 b0 = reinterpret_cast<const float4*>(b)[o], b1 = reinterpret_cast<const float4*>(b)[o + 1];
        v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w; v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
      }
      if (mask_hi) {
        float m[8];
        unpack8f(reinterpret_cast<const uint4*>(mask_hi)[i], m);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = m[j] > 0.f ? v[j] : 0.f;
      }
    }
    uint4 hh, ll;
    split8(v, hh, ll);
    reinterpret_cast<uint4*>(out_hi)[i] = hh;
    reinterpret_cast<uint4*>(out_lo)[i] = ll;
  }
}

// column sums of a split-plane matrix [M, C]: block = (TX channel-vectors of 8) x (TY row lanes); each block
// sweeps COLSUM_ROWS rows with TY rows in flight, reduces over TY in shared memory, then one atomic per channel.
constexpr int COLSUM_ROWS = 512;
__global__ void __launch_bounds__(256)
colsum_kernel(const uint16_t* __restrict__ hi, const uint16_t* __restrict__ lo, float* __restrict__ colsum,
              long long M, int C) {
  const int TX = blockDim.x, TY = blockDim.y;
  const int cv = blockIdx.x * TX + threadIdx.x;
  const long long r0 = (long long)blockIdx.y * COLSUM_ROWS;
  const long long r1 = r0 + COLSUM_ROWS < M ? r0 + COLSUM_ROWS : M;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  if (cv < C / 8) {
    const uint4* ph = reinterpret_cast<const uint4*>(hi) + cv;
    const uint4* pl = reinterpret_cast<const uint4*>(lo) + cv;
    const size_t stride = (size_t)C / 8;
    long long r = r0 + threadIdx.y;
    for (; r + 3 * TY < r1; r += 4 * TY) {
      uint4 h[4], l[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) { h[u] = __ldg(ph + (size_t)(r + u * TY) * stride); l[u] = __ldg(pl + (size_t)(r + u * TY) * stride); }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float fh[8], fl[8];
        unpack8f(h[u], fh); unpack8f(l[u], fl);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += fh[j] + fl[j];
      }
    }
    for (; r < r1; r += TY) {
      float fh[8], fl[8];
      unpack8f(__ldg(ph + (size_t)r * stride), fh); unpack8f(__ldg(pl + (size_t)r * stride), fl);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += fh[j] + fl[j];
    }
  }
  __shared__ float red[256][9];
  const int t = threadIdx.y * TX + threadIdx.x;
#pragma unroll
  for (int j = 0; j < 8; ++j) red[t][j] = acc[j];
  __syncthreads();
  if (threadIdx.y == 0 && cv < C / 8) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float s = 0.f;
      for (int y = 0; y < TY; ++y) s += red[y * TX + threadIdx.x][j];
      atomicAdd(&colsum[cv * 8 + j], s);
    }
  }
}

// weight re-layout (per optimiser step):  OIHW fp32 -> fprop planes [RS][Kf][C] and dgrad planes [RS][C][Kt]
__global__ void prep_weight_kernel(const float* __restrict__ w, const float* __restrict__ scale, int K, int C, int R,
                                   int S, int Kf, int Kt, uint16_t* __restrict__ wf_hi, uint16_t* __restrict__ wf_lo,
                                   uint16_t* __restrict__ wt_hi, uint16_t* __restrict__ wt_lo) {
  const int RS = R * S;
  const size_t nf = wf_hi ? (size_t)RS * Kf * C : 0;
  const size_t nt = wt_hi ? (size_t)RS * C * Kt : 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nf + nt; i += (size_t)gridDim.x * blockDim.x) {
    if (i < nf) {
      const int c = (int)(i % C);
      size_t t = i / C;
      const int k = (int)(t % Kf);
      const int rs = (int)(t / Kf);
      float v = 0.f;
      if (k < K) v = w[((size_t)k * C + c) * RS + rs];
      st_split(wf_hi, wf_lo, i, v);
    } else {
      const size_t j = i - nf;
      const int k = (int)(j % Kt);
      size_t t = j / Kt;
      const int c = (int)(t % C);
      const int rs = (int)(t / C);
      float v = 0.f;
      if (k < K) {
        v = w[((size_t)k * C + c) * RS + (RS - 1 - rs)];     // 180-degree flipped tap
        if (scale) v *= scale[k];
      }
      st_split(wt_hi, wt_lo, j, v);
    }
  }
}

__global__ void bn_fold_kernel(const float* __restrict__ gamma, const float* __restrict__ beta,
                               const float* __restrict__ mean, const float* __restrict__ var, float eps,
                               const float* __restrict__ conv_bias, float* __restrict__ scale,
                               float* __restrict__ shift, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  if (!gamma) {                                   // conv + bias without BN: y = conv + bias
    scale[c] = 1.f;
    shift[c] = conv_bias ? conv_bias[c] : 0.f;
    return;
  }
  // same operation order as ATen's batch_norm inference path: invstd = 1/sqrt(var+eps)
  const float invstd = 1.0f / sqrtf(var[c] + eps);
  const float sc = gamma[c] * invstd;
  scale[c] = sc;
  // y = gamma * (conv + bias - mean) * invstd + beta
  shift[c] = beta[c] + ((conv_bias ? conv_bias[c] : 0.f) - mean[c]) * sc;
}

// block per output channel k
__global__ void __launch_bounds__(256)
wgrad_finalize_kernel(const float* __restrict__ dwraw, const float* __restrict__ w, const float* __restrict__ scale,
                      const float* __restrict__ mean, const float* __restrict__ var, float eps,
                      const float* __restrict__ dbeta, float* __restrict__ dw, float* __restrict__ dgamma,
                      const float* __restrict__ conv_bias, float* __restrict__ dbias, int K, int C, int RS, int splits) {
  const int k = blockIdx.x;
  const float sc = scale ? scale[k] : 1.f;
  float dot = 0.f;
  const size_t base = (size_t)k * RS * C;
  for (int i = threadIdx.x; i < RS * C; i += blockDim.x) {
    const int rs = i / C, c = i - rs * C;
    float g = 0.f;                                   // [split][k][rs][c], summed in split order (deterministic)
    for (int sp = 0; sp < splits; ++sp) g += dwraw[(size_t)sp * K * RS * C + base + i];
    const size_t o = base + (size_t)c * RS + rs;     // [k][c][rs]
    dot = fmaf(w[o], g, dot);
    dw[o] = sc * g;
  }
  if (dgamma) {
    __shared__ float red[8];
    for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffff, dot, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = dot;
    __syncthreads();
    if (threadIdx.x == 0) {
      float s = 0.f;
      for (int i = 0; i < 8; ++i) s += red[i];
      // z = conv + bias: sum g*z = <W, dW_raw> + bias * dbeta
      const float b = conv_bias ? conv_bias[k] : 0.f;
      dgamma[k] = (s + (b - mean[k]) * dbeta[k]) * (1.0f / sqrtf(var[k] + eps));
    }
  }
  if (dbias && threadIdx.x == 0) dbias[k] = sc * dbeta[k];        // d(conv bias) = sum g * dy/dz = scale * d(beta)
}

// ---------------------------------------------------------------- batched (multi-layer) variants: ONE launch per network
// block -> item through a prefix table of blocks (binary search, <= 8 steps)
SACB_DEVINL int find_item(const int32_t* __restrict__ block_begin, int n_items, int blk) {
  int lo = 0, hi = n_items;
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (block_begin[mid] <= blk) lo = mid + 1; else hi = mid; }
  return lo - 1;
}

constexpr int PREP_PER_BLOCK = 2048;       // element-wise path (more than 9 taps: FCN-8s' 7x7 head conv)
constexpr int PREP_TILE = 32;              // tiled path: 32 output channels x 32 input channels x all taps per block
constexpr int PREP_MAX_RS = 9;
__global__ void __launch_bounds__(256)
prepare_batched_kernel(const SacbPrepItem* __restrict__ items, const int32_t* __restrict__ block_begin, int n_items, float eps) {
  const int it = find_item(block_begin, n_items, blockIdx.x);
  const SacbPrepItem d = items[it];
  const int cb = blockIdx.x - block_begin[it];
  if (cb == 0 && d.scale) {                        // BN fold (same expressions as bn_fold_kernel)
    for (int c = threadIdx.x; c < d.K; c += 256) {
      const float bias = d.conv_bias ? d.conv_bias[c] : 0.f;
      if (!d.gamma) { d.scale[c] = 1.f; d.shift[c] = bias; continue; }
      const float invstd = 1.0f / sqrtf(d.var[c] + eps);
      const float sc = d.gamma[c] * invstd;
      d.scale[c] = sc;
      d.shift[c] = d.beta[c] + (bias - d.mean[c]) * sc;
    }
  }
  if (!d.wf_hi) return;
  const int RS = d.R * d.S, K = d.K, C = d.C, Kf = d.Kf, Kt = d.Kt;
  const size_t nf = (size_t)RS * Kf * C;
  const size_t nt = d.wt_hi ? (size_t)RS * C * Kt : 0;
  uint16_t* wf_hi = (uint16_t*)d.wf_hi; uint16_t* wf_lo = (uint16_t*)d.wf_lo;
  uint16_t* wt_hi = (uint16_t*)d.wt_hi; uint16_t* wt_lo = (uint16_t*)d.wt_lo;
  if (RS <= PREP_MAX_RS) {
    // Tiled path (every layer up to 3x3): a block owns 32 output channels x 32 input channels x all taps.  The OIHW source of
    // one output channel is ONE contiguous run of 32*RS floats (coalesced), both plane layouts are written as 64-byte runs
    // (32 consecutive c of a [rs][k] row, 32 consecutive k of a [rs][c] row); the permutation happens in shared memory.  The
    // element-wise form below read the source with a stride of RS (fprop planes) or C*RS (dgrad planes) floats between lanes:
    // 0.47 ms per step for 0.5 GB of traffic.  Values are bit-identical: same expressions, no arithmetic between elements.
    __shared__ float s_w[PREP_TILE * (PREP_TILE * PREP_MAX_RS + 1)];
    __shared__ float s_g[PREP_TILE];
    const int c_tiles = (C + PREP_TILE - 1) / PREP_TILE;
    const int kt = cb / c_tiles, ct = cb - kt * c_tiles;
    const int k0 = kt * PREP_TILE, c0 = ct * PREP_TILE;
    const int cr = min(PREP_TILE, C - c0);
    const int run = cr * RS, ld = PREP_TILE * RS + 1;         // odd row stride: the k-fastest read below is conflict-free
    if (threadIdx.x < PREP_TILE) {
      const int k = k0 + threadIdx.x;
      s_g[threadIdx.x] = (d.gamma && k < K) ? d.gamma[k] * (1.0f / sqrtf(d.var[k] + eps)) : 1.f;     // same expression as the scale vector
    }
    for (int idx = threadIdx.x; idx < PREP_TILE * run; idx += 256) {
      const int kk = idx / run, off = idx - kk * run;
      const int k = k0 + kk;
      s_w[kk * ld + off] = k < K ? d.w[((size_t)k * C + c0) * RS + off] : 0.f;
    }
    __syncthreads();
    const bool fold_f = d.fold_wf && d.gamma;
    for (int idx = threadIdx.x; idx < RS * PREP_TILE * PREP_TILE; idx += 256) {          // fprop planes [rs][Kf][C]
      const int cc = idx & (PREP_TILE - 1), kk = (idx / PREP_TILE) & (PREP_TILE - 1), rs = idx / (PREP_TILE * PREP_TILE);
      const int k = k0 + kk;
      if (cc < cr && k < Kf) {
        float v = s_w[kk * ld + cc * RS + rs];
        if (fold_f) v *= s_g[kk];
        st_split(wf_hi, wf_lo, ((size_t)rs * Kf + k) * C + c0 + cc, v);
      }
    }
    if (wt_hi) {
      const bool fold_t = d.gamma != nullptr;
      for (int idx = threadIdx.x; idx < RS * PREP_TILE * PREP_TILE; idx += 256) {        // dgrad planes [rs][C][Kt], taps flipped
        const int kk = idx & (PREP_TILE - 1), cc = (idx / PREP_TILE) & (PREP_TILE - 1), rs = idx / (PREP_TILE * PREP_TILE);
        const int k = k0 + kk;
        if (cc < cr && k < Kt) {
          float v = s_w[kk * ld + cc * RS + (RS - 1 - rs)];
          if (fold_t) v *= s_g[kk];
          st_split(wt_hi, wt_lo, ((size_t)rs * C + c0 + cc) * Kt + k, v);
        }
      }
    }
    return;
  }
  const size_t i0 = (size_t)cb * PREP_PER_BLOCK;
  for (int e = threadIdx.x; e < PREP_PER_BLOCK; e += 256) {
    const size_t i = i0 + e;
    if (i >= nf + nt) break;
    if (i < nf) {
      const int c = (int)(i % C);
      size_t t = i / C;
      const int k = (int)(t % Kf);
      const int rs = (int)(t / Kf);
      float v = 0.f;
      if (k < K) {
        v = d.w[((size_t)k * C + c) * RS + rs];
        if (d.fold_wf && d.gamma) v *= d.gamma[k] * (1.0f / sqrtf(d.var[k] + eps));     // same expression as the scale vector
      }
      st_split(wf_hi, wf_lo, i, v);
    } else {
      const size_t j = i - nf;
      const int k = (int)(j % Kt);
      size_t t = j / Kt;
      const int c = (int)(t % C);
      const int rs = (int)(t / C);
      float v = 0.f;
      if (k < K) {
        v = d.w[((size_t)k * C + c) * RS + (RS - 1 - rs)];     // 180-degree flipped tap
        if (d.gamma) v *= d.gamma[k] * (1.0f / sqrtf(d.var[k] + eps));     // folded BN scale, recomputed (bit-identical)
      }
      st_split(wt_hi, wt_lo, j, v);
    }
  }
}

constexpr int FIN_SMEM_FLOATS = 9 * (512 + 1);
// block per (item, output channel)
__global__ void __launch_bounds__(256)
wgrad_finalize_batched_kernel(const SacbFinalizeItem* __restrict__ items, const int32_t* __restrict__ block_begin, int n_items,
                              float eps) {
  const int it = find_item(block_begin, n_items, blockIdx.x);
  const SacbFinalizeItem d = items[it];
  const int k = blockIdx.x - block_begin[it];
  const int K = d.K, C = d.C, RS = d.RS;
  const float sc = d.scale ? d.scale[k] : 1.f;
  float dot = 0.f;
  const size_t base = (size_t)k * RS * C;
  const size_t plane = (size_t)K * RS * C;           // one split's partial sums
  const bool vec = (C & 3) == 0 && (reinterpret_cast<uintptr_t>(d.dwraw) & 15) == 0;
  __shared__ float s_gsum[FIN_SMEM_FLOATS];
  if (vec && RS > 1 && RS * (C + 1) <= FIN_SMEM_FLOATS) {
    // 3x3 layers up to 512 input channels: the partials arrive as [rs][c], the gradient (and the weights it is multiplied
    // with) are [c][rs].  Sum the splits with 16-byte loads into shared memory, then walk the OUTPUT order: coalesced reads of w
    // and writes of dw instead of 4-byte accesses RS floats apart.  dw is bit-identical; the order of the d(gamma) dot product
    // differs from the direct form (same terms).
    for (int i = threadIdx.x * 4; i < RS * C; i += 1024) {
      const int rs = i / C, c = i - rs * C;
      const float* src = d.dwraw + base + i;
      float4 g = make_float4(0.f, 0.f, 0.f, 0.f);    // [split][k][rs][c], summed in split order (deterministic)
#pragma unroll 4
      for (int sp = 0; sp < d.splits; ++sp) {
        const float4 v = *reinterpret_cast<const float4*>(src + (size_t)sp * plane);
        g.x += v.x; g.y += v.y; g.z += v.z; g.w += v.w;
      }
      float* p = s_gsum + rs * (C + 1) + c;
      p[0] = g.x; p[1] = g.y; p[2] = g.z; p[3] = g.w;
    }
    __syncthreads();
    for (int o = threadIdx.x; o < RS * C; o += 256) {
      const int c = o / RS, rs = o - c * RS;
      const float g = s_gsum[rs * (C + 1) + c];
      dot = fmaf(d.w[base + o], g, dot);
      d.dw[base + o] = sc * g;
    }
  } else if (vec) {
    // 16-byte reads of the split-K partials (4 consecutive input channels of one tap): this kernel streams ~3.6 GB of
    // partials per step and was latency-bound with 4-byte loads (31 % of the HBM roofline, profiles/stream_kernels_r1p.txt)
    for (int i = threadIdx.x * 4; i < RS * C; i += 1024) {
      const int rs = i / C, c = i - rs * C;
      const float* src = d.dwraw + base + i;
      float4 g = make_float4(0.f, 0.f, 0.f, 0.f);    // [split][k][rs][c], summed in split order (deterministic)
#pragma unroll 4
      for (int sp = 0; sp < d.splits; ++sp) {
        const float4 v = *reinterpret_cast<const float4*>(src + (size_t)sp * plane);
        g.x += v.x; g.y += v.y; g.z += v.z; g.w += v.w;
      }
      const size_t o = base + (size_t)c * RS + rs;   // [k][c][rs]
      if (RS == 1 && ((reinterpret_cast<uintptr_t>(d.w + o) | reinterpret_cast<uintptr_t>(d.dw + o)) & 15) == 0) {
        const float4 wv = *reinterpret_cast<const float4*>(d.w + o);              // 1x1 layers: [k][c] is the partials' own order
        dot = fmaf(wv.x, g.x, dot); dot = fmaf(wv.y, g.y, dot); dot = fmaf(wv.z, g.z, dot); dot = fmaf(wv.w, g.w, dot);
        *reinterpret_cast<float4*>(d.dw + o) = make_float4(sc * g.x, sc * g.y, sc * g.z, sc * g.w);
        continue;
      }
      dot = fmaf(d.w[o], g.x, dot);              d.dw[o] = sc * g.x;
      dot = fmaf(d.w[o + RS], g.y, dot);         d.dw[o + RS] = sc * g.y;
      dot = fmaf(d.w[o + 2 * (size_t)RS], g.z, dot); d.dw[o + 2 * (size_t)RS] = sc * g.z;
      dot = fmaf(d.w[o + 3 * (size_t)RS], g.w, dot); d.dw[o + 3 * (size_t)RS] = sc * g.w;
    }
  } else {
    for (int i = threadIdx.x; i < RS * C; i += 256) {
      const int rs = i / C, c = i - rs * C;
      float g = 0.f;
      for (int sp = 0; sp < d.splits; ++sp) g += d.dwraw[sp * plane + base + i];
      const size_t o = base + (size_t)c * RS + rs;
      dot = fmaf(d.w[o], g, dot);
      d.dw[o] = sc * g;
    }
  }
  if (d.dgamma) {
    __shared__ float red[8];
    for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffff, dot, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = dot;
    __syncthreads();
    if (threadIdx.x == 0) {
      float s = 0.f;
      for (int i = 0; i < 8; ++i) s += red[i];
      const float b = d.conv_bias ? d.conv_bias[k] : 0.f;
      d.dgamma[k] = (s + (b - d.mean[k]) * d.dbeta[k]) * (1.0f / sqrtf(d.var[k] + eps));
    }
  }
  if (threadIdx.x == 0) {
    if (d.dbias) d.dbias[k] = sc * d.dbeta[k];
    if (d.dbeta_out) d.dbeta_out[k] = d.dbeta[k];    // d(beta) lands in the flat gradient buffer
  }
}

static inline int grid_for(size_t n, int block) {
  size_t g = (n + block - 1) / block;
  const size_t cap = 148 * 16;
  return (int)(g < cap ? (g ? g : 1) : cap);
}

}  // namespace sacb

using namespace sacb;
#define ST ((cudaStream_t)stream)
#define LAUNCHED() do { g_launches++; SACB_CHECK_CUDA(cudaGetLastError()); } while (0)

extern "C" int sacb_stem_fwd(const float* x, const float* w, const float* scale, const float* shift, void* out_hi,
                             void* out_lo, int N, int H, int W, int P, int Q, void* stream) {
  SACB_REQUIRE(P == (H + 6 - 7) / 2 + 1 && Q == (W + 6 - 7) / 2 + 1, "sacb_stem_fwd: bad output size");
  dim3 grid((Q + ST_TW - 1) / ST_TW, (P + ST_TH - 1) / ST_TH, N);
  stem_fwd_kernel<<<grid, 256, 0, ST>>>(x, w, scale, shift, (uint16_t*)out_hi, (uint16_t*)out_lo, N, H, W, P, Q);
  LAUNCHED();
  return 0;
}

extern "C" int sacb_stem_wgrad(const float* x, const void* g_hi, const void* g_lo, float* dw, int N, int H, int W, int P,
                               int Q, void* stream) {
  stem_wgrad_kernel<<<148 * 2, 256, 0, ST>>>(x, (const uint16_t*)g_hi, (const uint16_t*)g_lo, dw, N, H, W, P, Q);
  LAUNCHED();
  return 0;
}

extern "C" int sacb_stem_im2col(const float* x, void* a_hi, void* a_lo, int N, int H, int W, int P, int Q, int R,
                                int stride, int pad, int KP, void* stream) {
  SACB_REQUIRE(P == (H + 2 * pad - R) / stride + 1 && Q == (W + 2 * pad - R) / stride + 1, "sacb_stem_im2col: bad output size");
  SACB_REQUIRE(KP % 64 == 0 && KP >= 3 * R * R, "sacb_stem_im2col: KP must be a multiple of 64 covering 3*R*R taps");
  const size_t total = (size_t)N * P * Q * (KP / 8);
  if (R == 7 && stride == 2) {
    stem_im2col_smem_kernel<7, 2><<<dim3((Q + 63) / 64, P, N), 256, 0, ST>>>(x, (uint16_t*)a_hi, (uint16_t*)a_lo, H, W, P, Q, pad, KP);
  } else if (R == 3 && stride == 1) {
    stem_im2col_smem_kernel<3, 1><<<dim3((Q + 63) / 64, P, N), 256, 0, ST>>>(x, (uint16_t*)a_hi, (uint16_t*)a_lo, H, W, P, Q, pad, KP);
  } else if (R == 7) stem_im2col_kernel<7><<<grid_for(total, 256), 256, 0, ST>>>(x, (uint16_t*)a_hi, (uint16_t*)a_lo, N, H, W, P, Q, R, stride, pad, KP);
  else if (R == 3) stem_im2col_kernel<3><<<grid_for(total, 256), 256, 0, ST>>>(x, (uint16_t*)a_hi, (uint16_t*)a_lo, N, H, W, P, Q, R, stride, pad, KP);
  else stem_im2col_kernel<0><<<grid_for(total, 256), 256, 0, ST>>>(x, (uint16_t*)a_hi, (uint16_t*)a_lo, N, H, W, P, Q, R, stride, pad, KP);
  LAUNCHED();
  return 0;
}
extern "C" int sacb_stem_pack_weight(const float* w, void* hi, void* lo, int K, int taps, int KP, void* stream) {
  stem_pack_weight_kernel<<<(K * KP + 255) / 256, 256, 0, ST>>>(w, (uint16_t*)hi, (uint16_t*)lo, K, taps, KP);
  LAUNCHED();
  return 0;
}
extern "C" int sacb_stem_unpack_wgrad(const float* parts, int splits, float* dwraw, int K, int taps, int KP, void* stream) {
  stem_unpack_wgrad_kernel<<<(K * taps + 255) / 256, 256, 0, ST>>>(parts, splits, dwraw, K, taps, KP);
  LAUNCHED();
  return 0;
}
extern "C" int sacb_maxpool2_fwd(const void* in_hi, const void* in_lo, void* out_hi, void* out_lo, uint8_t* idx, int N,
                                 int H, int W, int C, int P, int Q, void* stream) {
  SACB_REQUIRE(C % 8 == 0 && P == H / 2 && Q == W / 2, "sacb_maxpool2_fwd: C %% 8, P = H/2, Q = W/2");
  const size_t total = (size_t)N * P * Q * (C / 8);
  maxpool2_fwd_kernel<<<grid_for(total, 256), 256, 0, ST>>>((const uint16_t*)in_hi, (const uint16_t*)in_lo,
                                                           (uint16_t*)out_hi, (uint16_t*)out_lo, idx, N, H, W, C, P, Q);
  LAUNCHED();
  return 0;
}
extern "C" int sacb_maxpool2_bwd(const float* g_out, const uint8_t* idx, const void* in_hi, void* gin_hi, void* gin_lo,
                                 int N, int H, int W, int C, int P, int Q, void* stream) {
  SACB_REQUIRE(C % 8 == 0, "sacb_maxpool2_bwd: C %% 8");
  const size_t total = (size_t)N * H * W * (C / 8);
  maxpool2_bwd_kernel<<<grid_for(total, 256), 256, 0, ST>>>(g_out, idx, (const uint16_t*)in_hi, (uint16_t*)gin_hi,
                                                           (uint16_t*)gin_lo, N, H, W, C, P, Q);
  LAUNCHED();
  return 0;
}

extern "C" int sacb_maxpool_fwd(const void* in_hi, const void* in_lo, void* out_hi, void* out_lo, uint8_t* idx, int N,
                                int H, int W, int C, int P, int Q, void* stream) {
  SACB_REQUIRE(C % 8 == 0, "sacb_maxpool_fwd: C %% 8");
  int Pe = (H + 2 - 3 + 1) / 2 + 1; if ((Pe - 1) * 2 >= H + 1) --Pe;
  int Qe = (W + 2 - 3 + 1) / 2 + 1; if ((Qe - 1) * 2 >= W + 1) --Qe;
  SACB_REQUIRE(P == Pe && Q == Qe, "sacb_maxpool_fwd: bad ceil-mode output size (%d,%d) vs (%d,%d)", P, Q, Pe, Qe);
  const size_t total = (size_t)N * P * Q * (C / 8);
  maxpool_fwd_kernel<<<grid_for(total, 256), 256, 0, ST>>>((const uint16_t*)in_hi, (const uint16_t*)in_lo,
                                                          (uint16_t*)out_hi, (uint16_t*)out_lo, idx, N, H, W, C, P, Q);
  LAUNCHED();
  return 0;
}

extern "C" int sacb_maxpool_bwd(const float* g_out, const uint8_t* idx, const void* in_hi, void* gin_hi, void* gin_lo,
                                int N, int H, int W, int C, int P, int Q, void* stream) {
  SACB_REQUIRE(C % 8 == 0, "sacb_maxpool_bwd: C %% 8");
  const size_t total = (size_t)N * H * W * (C / 8);
  maxpool_bwd_kernel<<<grid_for(total, 256), 256, 0, ST>>>(g_out, idx, (const uint16_t*)in_hi, (uint16_t*)gin_hi,
                                                          (uint16_t*)gin_lo, N, H, W, C, P, Q);
  LAUNCHED();
  return 0;
}

extern "C" int sacb_add_mask_split(const float* a, const float* b, const void* mask_hi, void* out_hi, void* out_lo,
                                   int64_t n, void* stream) {
  SACB_REQUIRE(n % 8 == 0, "sacb_add_mask_split: n %% 8");
  add_mask_split_kernel<<<grid_for((size_t)n / 8, 256), 256, 0, ST>>>(a, b, (const uint16_t*)mask_hi, (uint16_t*)out_hi,
                                                                     (uint16_t*)out_lo, (size_t)n / 8);
  LAUNCHED();
  return 0;
}

extern "C" int sacb_scatter2_mask_split(const float* a, const float* b, const void* mask_hi, void* out_hi, void* out_lo,
                                        int N, int H, int W, int C, int P, int Q, void* stream) {
  SACB_REQUIRE(C % 8 == 0, "sacb_scatter2_mask_split: C %% 8");
  const size_t total = (size_t)N * H * W * (C / 8);
  scatter2_mask_split_kernel<<<grid_for(total, 256), 256, 0, ST>>>(a, b, (const uint16_t*)mask_hi, (uint16_t*)out_hi,
                                                                  (uint16_t*)out_lo, N, H, W, C, P, Q);
  LAUNCHED();
  return 0;
}

extern "C" int sacb_colsum(const void* hi, const void* lo, float* colsum, int64_t M, int C, void* stream) {
  SACB_REQUIRE(C % 8 == 0, "sacb_colsum: C %% 8");
  const int cvs = C / 8;
  const int tx = cvs >= 32 ? 32 : (cvs >= 16 ? 16 : 8);
  dim3 block(tx, 256 / tx);
  dim3 grid((cvs + tx - 1) / tx, (unsigned)((M + COLSUM_ROWS - 1) / COLSUM_ROWS));
  colsum_kernel<<<grid, block, 0, ST>>>((const uint16_t*)hi, (const uint16_t*)lo, colsum, (long long)M, C);
  LAUNCHED();
  return 0;
}

extern "C" int sacb_prep_weight(const float* w, const float* scale, int K, int C, int R, int S, int Kf, int Kt,
                                void* wf_hi, void* wf_lo, void* wt_hi, void* wt_lo, void* stream) {
  const size_t n = (wf_hi ? (size_t)R * S * Kf * C : 0) + (wt_hi ? (size_t)R * S * C * Kt : 0);
  if (n == 0) return 0;
  prep_weight_kernel<<<grid_for(n, 256), 256, 0, ST>>>(w, scale, K, C, R, S, Kf, Kt, (uint16_t*)wf_hi, (uint16_t*)wf_lo,
                                                      (uint16_t*)wt_hi, (uint16_t*)wt_lo);
  LAUNCHED();
  return 0;
}

extern "C" int sacb_bn_fold(const float* gamma, const float* beta, const float* mean, const float* var, float eps,
                            const float* conv_bias, float* scale, float* shift, int C, void* stream) {
  bn_fold_kernel<<<(C + 127) / 128, 128, 0, ST>>>(gamma, beta, mean, var, eps, conv_bias, scale, shift, C);
  LAUNCHED();
  return 0;
}

extern "C" int sacb_wgrad_finalize(const float* dwraw, const float* w, const float* scale, const float* mean,
                                   const float* var, float eps, const float* dbeta, float* dw, float* dgamma,
                                   const float* conv_bias, float* dbias, int K, int C, int R, int S, int splits,
                                   void* stream) {
  wgrad_finalize_kernel<<<K, 256, 0, ST>>>(dwraw, w, scale, mean, var, eps, dbeta, dw, dgamma, conv_bias, dbias, K, C,
                                          R * S, splits);
  LAUNCHED();
  return 0;
}

extern "C" int sacb_prep_item_blocks(int K, int C, int R, int S, int Kf, int Kt, int with_wf, int with_wt) {
  if (with_wf && R * S <= PREP_MAX_RS) {
    const int kmax = (with_wt && Kt > Kf) ? Kt : Kf;
    return ((kmax + PREP_TILE - 1) / PREP_TILE) * ((C + PREP_TILE - 1) / PREP_TILE);
  }
  const size_t n = (with_wf ? (size_t)R * S * Kf * C : 0) + (with_wf && with_wt ? (size_t)R * S * C * Kt : 0);
  const size_t b = (n + PREP_PER_BLOCK - 1) / PREP_PER_BLOCK;
  return (int)(b ? b : 1);
}

extern "C" int sacb_prepare_batched(const SacbPrepItem* items_dev, const int32_t* block_begin_dev, int n_items, int total_blocks,
                                    float eps, void* stream) {
  SACB_REQUIRE(items_dev && block_begin_dev && n_items > 0 && total_blocks > 0, "sacb_prepare_batched: bad arguments");
  prepare_batched_kernel<<<total_blocks, 256, 0, ST>>>(items_dev, block_begin_dev, n_items, eps);
  LAUNCHED();
  return 0;
}

extern "C" int sacb_wgrad_finalize_batched(const SacbFinalizeItem* items_dev, const int32_t* block_begin_dev, int n_items,
                                           int total_blocks, float eps, void* stream) {
  SACB_REQUIRE(items_dev && block_begin_dev && n_items > 0 && total_blocks > 0, "sacb_wgrad_finalize_batched: bad arguments");
  wgrad_finalize_batched_kernel<<<total_blocks, 256, 0, ST>>>(items_dev, block_begin_dev, n_items, eps);
  LAUNCHED();
  return 0;
}
