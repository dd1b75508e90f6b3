#line 1 "../../da_sac_b200/csrc/sacb_aug.cu"
// Target-view augmentation on the device: K views per base crop (guided h-flip, zoom-crop + bilinear resize with the
// matching pad mask), then on the "noisy" copy Gaussian blur, colour jitter in a random op order, greyscale; normalise;
// padded pixels -> image 0 / label -1.
//
// Replaces the CPU/PIL pipeline of /root/reference/datasets/dataloader_target.py:264-306 built from
// /root/reference/datasets/tf_target.py: GuidedRandHFlip (:140-156), MaskRandScaleCrop (:158-239), RandGaussianBlur
// (:331-349), MaskRandJitter (:366-390), MaskRandGreyscale (:351-364), ToTensorMask / Normalize / ApplyMask (:32-98).
// The random parameters are drawn on the host in the reference's order (da_sac_b200/augment.py), so the affine
// operators are identical; pixel arithmetic restates Pillow's 8-bit rules (resample coefficients, ImageEnhance blends with
// truncation, 'L' / HSV conversions) in fp32 -- parity with PIL is to ~1 grey level, not bit-exact.
// All kernels are HBM-bound streaming passes; compiled with -fmad=false so the CPU restatement (oracle/aug_oracle.py)
// reproduces every rounding decision.
#include <atomic>
// (sacb_common.cuh: see cuda_emul.h)
#include "../../include/sacb.h"

namespace sacb {
extern std::atomic<long long> g_launches;

constexpr int NP = SACB_AUG_NPARAM;
constexpr int MAXT = 6;            // resample taps per axis (crop/out scale <= 2)
constexpr int BLUR_R = 6;          // Gaussian taps: radius ceil(3 sigma) <= 6 (sigma <= 2, tf_target.py:337)

// view parameter slots
enum { P_FLIP = 0, P_TOP, P_LEFT, P_CH, P_CW, P_SIGMA, P_JITTER, P_OP0, P_OP1, P_OP2, P_OP3, P_BRIGHT, P_CONTRAST, P_SAT,
       P_HUE, P_GREY };

SACB_DEVINL float tri(float x) { x = fabsf(x); return x < 1.f ? 1.f - x : 0.f; }

// Pillow precompute_coeffs() for the BILINEAR filter (support 1, scaled by max(scale,1) when shrinking)
SACB_DEVINL void resample_coeffs(int o, float scale, int in_size, int& kmin, int& kn, float* w) {
  const float fs = fmaxf(scale, 1.f);
  const float support = fs;
  const float center = ((float)o + 0.5f) * scale;
  const float ss = 1.f / fs;
  int xmin = (int)(center - support + 0.5f);
  if (xmin < 0) xmin = 0;
  int xmax = (int)(center + support + 0.5f);
  if (xmax > in_size) xmax = in_size;
  int n = xmax - xmin;
  if (n > MAXT) n = MAXT;
  float ww = 0.f;
  for (int x = 0; x < MAXT; ++x) {
    const float v = x < n ? tri(((float)(x + xmin) - center + 0.5f) * ss) : 0.f;
    w[x] = v;
    ww += v;
  }
  if (ww != 0.f)
    for (int x = 0; x < MAXT; ++x) w[x] /= ww;
  kmin = xmin; kn = n;
}

// ---------------------------------------------------------------- K1: geometry (flip -> zoom-crop -> resize)
// grid (ceil(HW/256), BT)
__global__ void __launch_bounds__(256)
aug_geom_kernel(const uint8_t* __restrict__ base, const uint8_t* __restrict__ base_mask, const uint8_t* __restrict__ base_label,
                const float* __restrict__ vp, uint8_t* __restrict__ raw, uint8_t* __restrict__ mask,
                float* __restrict__ frames2, int64_t* __restrict__ gt, int T, int H, int W, float m0, float m1, float m2,
                float s0, float s1, float s2) {
  const int b = blockIdx.y, g = b / T;
  const int pix = blockIdx.x * 256 + threadIdx.x;
  const int HW = H * W;
  if (pix >= HW) return;
  const int y = pix / W, x = pix - y * W;
  const float* p = vp + b * NP;
  const bool flip = p[P_FLIP] < 0.f;
  const int top = (int)p[P_TOP], left = (int)p[P_LEFT], ch = (int)p[P_CH], cw = (int)p[P_CW];
  const float sy = (float)ch / (float)H, sx = (float)cw / (float)W;
  float wy[MAXT], wx[MAXT];
  int ky0, kyn, kx0, kxn;
  resample_coeffs(y, sy, ch, ky0, kyn, wy);
  resample_coeffs(x, sx, cw, kx0, kxn, wx);
  const uint8_t* img = base + (size_t)g * HW * 3;
  float acc[3] = {0.f, 0.f, 0.f};
  for (int a = 0; a < kyn; ++a) {
    const int yy = top + ky0 + a;
    if (yy < 0 || yy >= H) continue;                  // zoom-out: the crop window extends into the (black) padding
    float row[3] = {0.f, 0.f, 0.f};
    for (int c = 0; c < kxn; ++c) {
      const int xx = left + kx0 + c;
      if (xx < 0 || xx >= W) continue;
      const int xb = flip ? W - 1 - xx : xx;          // GuidedRandHFlip precedes the crop (dataloader_target.py:108-113)
      const uint8_t* q = img + ((size_t)yy * W + xb) * 3;
      row[0] += wx[c] * (float)q[0]; row[1] += wx[c] * (float)q[1]; row[2] += wx[c] * (float)q[2];
    }
    acc[0] += wy[a] * row[0]; acc[1] += wy[a] * row[1]; acc[2] += wy[a] * row[2];
  }
  // nearest-neighbour source of the mask / label (Pillow: floor((o + 0.5) * scale))
  int ys = (int)(((float)y + 0.5f) * sy), xs = (int)(((float)x + 0.5f) * sx);
  ys = min(ys, ch - 1); xs = min(xs, cw - 1);
  const int yy = top + ys, xx = left + xs;
  uint8_t mk = 1, lab = 255;
  if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
    const int xb = flip ? W - 1 - xx : xx;
    const size_t o = (size_t)g * HW + (size_t)yy * W + xb;
    mk = base_mask ? (base_mask[o] > 0 ? 1 : 0) : 0;
    lab = base_label ? base_label[o] : 255;
  }
  const size_t ob = (size_t)b * HW + pix;
  const float mean[3] = {m0, m1, m2}, stdv[3] = {s0, s1, s2};
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    int v = (int)(acc[c] + 0.5f);
    v = v < 0 ? 0 : (v > 255 ? 255 : v);
    raw[ob * 3 + c] = (uint8_t)v;
    const float t = ((float)v / 255.f - mean[c]) / stdv[c];               // to_tensor, Normalize (tf_target.py:32-80)
    frames2[((size_t)b * 3 + c) * HW + pix] = mk ? 0.f : t;                // ApplyMask (tf_target.py:82-98)
  }
  mask[ob] = mk;
  gt[ob] = mk ? -1 : (int64_t)lab;
}

// ---------------------------------------------------------------- photometric ops on integer grey levels held in floats
SACB_DEVINL float clip_trunc(float t) {            // ImagingBlend: clip to [0,255], then (UINT8) cast
  if (t <= 0.f) return 0.f;
  if (t >= 255.f) return 255.f;
  return floorf(t);
}
SACB_DEVINL float grey_L(const float* v) {         // Pillow rgb -> L: (r*19595 + g*38470 + b*7471 + 0x8000) >> 16
  const unsigned r = (unsigned)v[0], g = (unsigned)v[1], b = (unsigned)v[2];
  return (float)((r * 19595u + g * 38470u + b * 7471u + 0x8000u) >> 16);
}
SACB_DEVINL void op_hue(float* v, float hue) {     // torchvision adjust_hue on PIL: uint8 HSV round trip with a wrapped H shift
  const int r = (int)v[0], g = (int)v[1], b = (int)v[2];
  const int maxc = max(r, max(g, b)), minc = min(r, min(g, b));
  int uh = 0, us = 0;
  const int uv = maxc;
  if (minc != maxc) {
    const float cr = (float)(maxc - minc);
    const float s = cr / (float)maxc;
    const float rc = (float)(maxc - r) / cr, gc = (float)(maxc - g) / cr, bc = (float)(maxc - b) / cr;
    float h;
    if (r == maxc) h = bc - gc;
    else if (g == maxc) h = 2.f + rc - bc;
    else h = 4.f + gc - rc;
    h = fmodf(h / 6.f + 1.f, 1.f);
    uh = min(255, max(0, (int)(h * 255.f)));
    us = min(255, max(0, (int)(s * 255.f)));
  }
  const int shift = ((int)(hue * 255.f)) & 255;      // np.int32(hue_factor * 255).astype(np.uint8)
  uh = (uh + shift) & 255;
  if (us == 0) { v[0] = v[1] = v[2] = (float)uv; return; }
  const float hf = (float)uh * 6.f / 255.f;
  const int i = (int)floorf(hf);
  const float f = hf - (float)i;
  const float fs = (float)us / 255.f;
  const float fv = (float)uv;
  const float pp = fminf(255.f, fmaxf(0.f, rintf(fv * (1.f - fs))));
  const float qq = fminf(255.f, fmaxf(0.f, rintf(fv * (1.f - fs * f))));
  const float tt = fminf(255.f, fmaxf(0.f, rintf(fv * (1.f - fs * (1.f - f)))));
  switch (i % 6) {
    case 0: v[0] = fv; v[1] = tt; v[2] = pp; break;
    case 1: v[0] = qq; v[1] = fv; v[2] = pp; break;
    case 2: v[0] = pp; v[1] = fv; v[2] = tt; break;
    case 3: v[0] = pp; v[1] = qq; v[2] = fv; break;
    case 4: v[0] = tt; v[1] = pp; v[2] = fv; break;
    default: v[0] = fv; v[1] = pp; v[2] = qq; break;
  }
}
// apply jitter ops order[from..to) ; `mean_level` is only read by the contrast op
SACB_DEVINL void apply_ops(float* v, const float* p, int from, int to, float mean_level) {
  for (int k = from; k < to; ++k) {
    const int op = (int)p[P_OP0 + k];
    if (op == 0) {                                   // ImageEnhance.Brightness: blend(black, img, f)
      const float f = p[P_BRIGHT];
      for (int c = 0; c < 3; ++c) v[c] = clip_trunc(0.f + f * (v[c] - 0.f));
    } else if (op == 1) {                            // ImageEnhance.Contrast: blend(mean grey, img, f)
      const float f = p[P_CONTRAST];
      for (int c = 0; c < 3; ++c) v[c] = clip_trunc(mean_level + f * (v[c] - mean_level));
    } else if (op == 2) {                            // ImageEnhance.Color: blend(L(img), img, f)
      const float f = p[P_SAT], L = grey_L(v);
      for (int c = 0; c < 3; ++c) v[c] = clip_trunc(L + f * (v[c] - L));
    } else {
      op_hue(v, p[P_HUE]);
    }
  }
}
SACB_DEVINL int contrast_pos(const float* p) {      // index of the contrast op in this view's order (4 = jitter off)
  if (p[P_JITTER] == 0.f) return 4;
  for (int k = 0; k < 4; ++k) if ((int)p[P_OP0 + k] == 1) return k;
  return 4;
}

// ---------------------------------------------------------------- K2a: horizontal Gaussian pass (u8 -> fp32)
__global__ void __launch_bounds__(256)
aug_blur_h_kernel(const uint8_t* __restrict__ raw, const float* __restrict__ vp, float* __restrict__ tmp, int H, int W) {
  const int b = blockIdx.y;
  const int pix = blockIdx.x * 256 + threadIdx.x;
  const int HW = H * W;
  if (pix >= HW) return;
  const int y = pix / W, x = pix - y * W;
  const float sigma = vp[b * NP + P_SIGMA];
  const uint8_t* img = raw + (size_t)b * HW * 3;
  float* dst = tmp + ((size_t)b * HW + pix) * 3;
  if (!(sigma > 0.f)) {
    const uint8_t* q = img + (size_t)pix * 3;
    dst[0] = (float)q[0]; dst[1] = (float)q[1]; dst[2] = (float)q[2];
    return;
  }
  const int R = min(BLUR_R, (int)ceilf(3.f * sigma));
  const float inv = 1.f / (2.f * sigma * sigma);
  float acc[3] = {0.f, 0.f, 0.f}, ws = 0.f;
  for (int k = -R; k <= R; ++k) {
    const float w = expf(-(float)(k * k) * inv);
    const int xx = min(W - 1, max(0, x + k));            // edge pixels are extended
    const uint8_t* q = img + ((size_t)y * W + xx) * 3;
    acc[0] += w * (float)q[0]; acc[1] += w * (float)q[1]; acc[2] += w * (float)q[2];
    ws += w;
  }
  dst[0] = acc[0] / ws; dst[1] = acc[1] / ws; dst[2] = acc[2] / ws;
}

// ---------------------------------------------------------------- K2b: vertical pass -> 8-bit levels -> ops before contrast -> grey sum
__global__ void __launch_bounds__(256)
aug_blur_v_kernel(const float* __restrict__ tmp, const float* __restrict__ vp, float* __restrict__ lev,
                  unsigned long long* __restrict__ grey_sum, int H, int W) {
  const int b = blockIdx.y;
  const int pix = blockIdx.x * 256 + threadIdx.x;
  const int HW = H * W;
  const float* p = vp + b * NP;
  unsigned L = 0;
  if (pix < HW) {
    const int y = pix / W, x = pix - y * W;
    const float sigma = p[P_SIGMA];
    const float* img = tmp + (size_t)b * HW * 3;
    float v[3];
    if (!(sigma > 0.f)) {
      const float* q = img + (size_t)pix * 3;
      v[0] = q[0]; v[1] = q[1]; v[2] = q[2];
    } else {
      const int R = min(BLUR_R, (int)ceilf(3.f * sigma));
      const float inv = 1.f / (2.f * sigma * sigma);
      float acc[3] = {0.f, 0.f, 0.f}, ws = 0.f;
      for (int k = -R; k <= R; ++k) {
        const float w = expf(-(float)(k * k) * inv);
        const int yy = min(H - 1, max(0, y + k));
        const float* q = img + ((size_t)yy * W + x) * 3;
        acc[0] += w * q[0]; acc[1] += w * q[1]; acc[2] += w * q[2];
        ws += w;
      }
      for (int c = 0; c < 3; ++c) v[c] = fminf(255.f, fmaxf(0.f, floorf(acc[c] / ws + 0.5f)));   // back to an 8-bit image
    }
    const int cp = contrast_pos(p);
    if (p[P_JITTER] != 0.f) apply_ops(v, p, 0, cp < 4 ? cp : 4, 0.f);
    float* dst = lev + ((size_t)b * HW + pix) * 3;
    dst[0] = v[0]; dst[1] = v[1]; dst[2] = v[2];
    L = (unsigned)grey_L(v);
  }
  // exact integer sum of the grey levels (ImageStat mean of convert("L")): order-independent, hence deterministic
  for (int o = 16; o > 0; o >>= 1) L += __shfl_xor_sync(0xffffffff, L, o);
  __shared__ unsigned red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = L;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned s = 0;
    for (int k = 0; k < 8; ++k) s += red[k];
    atomicAdd(grey_sum + b, (unsigned long long)s);
  }
}

// ---------------------------------------------------------------- K3: contrast + remaining ops, greyscale, normalise, mask
__global__ void __launch_bounds__(256)
aug_finish_kernel(const float* __restrict__ lev, const float* __restrict__ vp, const unsigned long long* __restrict__ grey_sum,
                  const uint8_t* __restrict__ mask, float* __restrict__ frames1, int H, int W, float m0, float m1, float m2,
                  float s0, float s1, float s2) {
  const int b = blockIdx.y;
  const int pix = blockIdx.x * 256 + threadIdx.x;
  const int HW = H * W;
  if (pix >= HW) return;
  const float* p = vp + b * NP;
  const float* q = lev + ((size_t)b * HW + pix) * 3;
  float v[3] = {q[0], q[1], q[2]};
  const int cp = contrast_pos(p);
  if (cp < 4) {
    // ImageEnhance.Contrast: mean = int(ImageStat.Stat(image.convert("L")).mean[0] + 0.5)
    const double mean = (double)grey_sum[b] / (double)HW;
    apply_ops(v, p, cp, 4, (float)(int)(mean + 0.5));
  }
  if (p[P_GREY] != 0.f) { const float L = grey_L(v); v[0] = v[1] = v[2] = L; }      // to_grayscale(num_output_channels=3)
  const bool mk = mask[(size_t)b * HW + pix] != 0;
  const float mean3[3] = {m0, m1, m2}, std3[3] = {s0, s1, s2};
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float t = (v[c] / 255.f - mean3[c]) / std3[c];
    frames1[((size_t)b * 3 + c) * HW + pix] = mk ? 0.f : t;
  }
}
}  // namespace sacb

using namespace sacb;
#define ST ((cudaStream_t)stream)
#define LAUNCHED() do { g_launches++; SACB_CHECK_CUDA(cudaGetLastError()); } while (0)

extern "C" int sacb_target_augment(const SacbAug* d, void* stream) {
  SACB_REQUIRE(d && d->size == sizeof(SacbAug), "sacb_target_augment: bad descriptor size");
  SACB_REQUIRE(d->G > 0 && d->T > 0 && d->H > 0 && d->W > 0, "sacb_target_augment: bad shape");
  SACB_REQUIRE(d->base && d->view_params && d->raw && d->mask && d->tmp && d->levels && d->grey_sum && d->frames1 && d->frames2 &&
               d->gt, "sacb_target_augment: NULL buffer");
  const int BT = d->G * d->T, HW = d->H * d->W;
  dim3 grid((HW + 255) / 256, BT);
  cuda_emul::run_grid("aug_geom_kernel", grid, 256, 0, false, [&]() { aug_geom_kernel(d->base, d->base_mask, d->base_label, d->view_params, d->raw, d->mask, d->frames2, d->gt,
                                       d->T, d->H, d->W, d->mean[0], d->mean[1], d->mean[2], d->std[0], d->std[1], d->std[2]); });
  LAUNCHED();
  SACB_CHECK_CUDA(cudaMemsetAsync(d->grey_sum, 0, sizeof(unsigned long long) * BT, ST));
  cuda_emul::run_grid("aug_blur_h_kernel", grid, 256, 0, false, [&]() { aug_blur_h_kernel(d->raw, d->view_params, d->tmp, d->H, d->W); });
  LAUNCHED();
  cuda_emul::run_grid("aug_blur_v_kernel", grid, 256, 0, true, [&]() { aug_blur_v_kernel(d->tmp, d->view_params, d->levels, reinterpret_cast<unsigned long long*>(d->grey_sum),
                                         d->H, d->W); });
  LAUNCHED();
  cuda_emul::run_grid("aug_finish_kernel", grid, 256, 0, false, [&]() { aug_finish_kernel(d->levels, d->view_params, reinterpret_cast<const unsigned long long*>(d->grey_sum),
                                         d->mask, d->frames1, d->H, d->W, d->mean[0], d->mean[1], d->mean[2], d->std[0],
                                         d->std[1], d->std[2]); });
  LAUNCHED();
  return 0;
}
