// SAC-specific tail: teacher probabilities -> multi-view average in the reference frame -> pseudo labels,
// student loss (fused upsample + log-softmax + weighted NLL) and its backward, plus the multi-tensor
// EMA / norm and SGD kernels.  All HBM-bound; nothing of size [BT,19,H,W] is materialised except the
// teacher probability scratch.
//
// Replaces /root/reference/models/sac.py:104-117 (_update_running_conf), :134-149 (_focal_ce_conf),
// :151-187 (_pseudo_labels_probs), :238-269 (_avg_pool), :271-313 (_refine), :70-102 (_momentum_update)
// and F.interpolate + CrossEntropyLoss in models/deeplabv2.py:217-224.
#include <atomic>
#include <cstdlib>
#include "sacb_common.cuh"
#include "../../include/sacb.h"

namespace sacb {
extern std::atomic<long long> g_launches;


struct UpCoef { int i0, i1; float l0, l1; };
// aten upsample_bilinear2d, align_corners=True: src = dst * (in-1)/(out-1)
SACB_DEVINL UpCoef up_coef(int dst, int in_size, int out_size) {
  const float scale = out_size > 1 ? (float)(in_size - 1) / (float)(out_size - 1) : 0.f;
  const float src = scale * (float)dst;
  UpCoef c;
  c.i0 = min((int)src, in_size - 1);
  c.i1 = c.i0 + (c.i0 < in_size - 1 ? 1 : 0);
  c.l1 = src - (float)c.i0;
  c.l0 = 1.f - c.l1;
  return c;
}
// upsampled logits of all classes at (i, j); L is [C,h,w] of one image
template <int C_>
SACB_DEVINL void up_logits(const float* __restrict__ L, int C, int h, int w, const UpCoef& cy, const UpCoef& cx,
                           float* out) {
  const int o00 = cy.i0 * w + cx.i0, o01 = cy.i0 * w + cx.i1, o10 = cy.i1 * w + cx.i0, o11 = cy.i1 * w + cx.i1;
  const int hw = h * w;
#pragma unroll
  for (int c = 0; c < C_; ++c) {
    if (c < C) {
      const float* p = L + c * hw;
      out[c] = cy.l0 * (cx.l0 * __ldg(p + o00) + cx.l1 * __ldg(p + o01)) +
               cy.l1 * (cx.l0 * __ldg(p + o10) + cx.l1 * __ldg(p + o11));
    }
  }
}

// The same, from low-resolution logits staged in shared memory.  The three kernels that up-sample the 19-class logits and take a
// softmax per full-resolution pixel (tail_probs, loss_fwd, loss_grad_rows) spent most of their issue slots on 64-bit address
// arithmetic for 76 scalar global loads per pixel (profiles/r2ab_*: LEA + IADD3 + IMAD + ISETP + SHF = 46 % of the instructions).
// A block touches two to four low-resolution rows; staged pixel-major ([row][column][20]: 16-byte records) they are read with
// 20 LDS.128 at immediate class offsets from four base addresses.  Same expression per class as up_logits: bit-identical.
// Measured (profiles/r2ac_*): tail_probs 449 -> 389 us, loss_fwd 372 -> 309 us at 24 x 512^2.  The same idea for the loss backward
// (256-pixel segments, staged logits, a table of the column coefficients) was bit-identical too but SLOWER than
// loss_grad_rows_kernel in three builds (1.23 / 1.21 / 0.93 ms against 0.61 ms; profiles/r2ac..r2ae_*) and was removed.
constexpr int UPS_FLOATS = 5200;            // 2 rows x 130 columns x 20 floats, or 4 rows x 65 columns
template <int C_>
SACB_DEVINL bool stage_up_rows(float* s_up, int cap_floats, const float* __restrict__ L, int C, int h, int w, int H,
                               int i_first, int i_last, int& r_lo) {
  constexpr int CS = (C_ + 3) / 4 * 4;
  const UpCoef a = up_coef(i_first, h, H), z = up_coef(i_last, h, H);
  r_lo = a.i0;
  const int n = (z.i1 - a.i0 + 1) * w;       // staged low-resolution pixels: rows r_lo .. z.i1, contiguous in the source
  if (C != C_ || n * CS > cap_floats) return false;             // block-uniform
  for (int e = threadIdx.x; e < n * C_; e += blockDim.x) {
    const int c = e / n, rx = e - c * n;
    s_up[rx * CS + c] = __ldg(L + (size_t)c * h * w + (size_t)r_lo * w + rx);
  }
  __syncthreads();
  return true;
}
template <int C_>
SACB_DEVINL void up_logits_staged(const float* s_up, int w, int r_lo, const UpCoef& cy, const UpCoef& cx, float* out) {
  constexpr int CS = (C_ + 3) / 4 * 4;
  const int r0 = (cy.i0 - r_lo) * w, r1 = (cy.i1 - r_lo) * w;
  const float4* p00 = reinterpret_cast<const float4*>(s_up + (r0 + cx.i0) * CS);
  const float4* p01 = reinterpret_cast<const float4*>(s_up + (r0 + cx.i1) * CS);
  const float4* p10 = reinterpret_cast<const float4*>(s_up + (r1 + cx.i0) * CS);
  const float4* p11 = reinterpret_cast<const float4*>(s_up + (r1 + cx.i1) * CS);
#pragma unroll
  for (int q = 0; q < CS / 4; ++q) {
    const float4 a = p00[q], b = p01[q], c = p10[q], d = p11[q];
    const float va[4] = {a.x, a.y, a.z, a.w}, vb[4] = {b.x, b.y, b.z, b.w}, vc[4] = {c.x, c.y, c.z, c.w}, vd[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (4 * q + k < C_)
        out[4 * q + k] = cy.l0 * (cx.l0 * va[k] + cx.l1 * vb[k]) + cy.l1 * (cx.l0 * vc[k] + cx.l1 * vd[k]);
    }
  }
}

// affine_grid(align_corners=False) base coordinate: linspace(-1,1,n)[j] * (n-1)/n
SACB_DEVINL float base_coord(int j, int n) {
  if (n <= 1) return 0.f;
  const float step = 2.f / (float)(n - 1);
  const float v = (j < n / 2) ? (-1.f + step * (float)j) : (1.f - step * (float)(n - 1 - j));
  return v * (float)(n - 1) / (float)n;
}

struct Taps { int x0, y0; float nw, ne, sw, se; bool in_x0, in_x1, in_y0, in_y1; };
// grid_sample(bilinear, zeros, align_corners=False) sampling parameters for output pixel (i,j) under theta
SACB_DEVINL Taps make_taps(const float* __restrict__ th, int i, int j, int H, int W) {
  const float x = base_coord(j, W), y = base_coord(i, H);
  float u = x * th[0]; u = fmaf(y, th[1], u); u += th[2];
  float v = x * th[3]; v = fmaf(y, th[4], v); v += th[5];
  const float ix = (u + 1.f) * (0.5f * (float)W) - 0.5f;
  const float iy = (v + 1.f) * (0.5f * (float)H) - 0.5f;
  const float xw = floorf(ix), yn = floorf(iy);
  const float we = ix - xw, ww = 1.f - we;     // east / west weights
  const float ws = iy - yn, wn = 1.f - ws;     // south / north weights
  Taps t;
  // clamp before the int conversion so that wild coordinates stay representable
  const float xc = fminf(fmaxf(xw, -2.f), (float)W + 1.f), yc = fminf(fmaxf(yn, -2.f), (float)H + 1.f);
  t.x0 = (int)xc; t.y0 = (int)yc;
  t.nw = wn * ww; t.ne = wn * we; t.sw = ws * ww; t.se = ws * we;
  t.in_x0 = t.x0 >= 0 && t.x0 < W; t.in_x1 = t.x0 + 1 >= 0 && t.x0 + 1 < W;
  t.in_y0 = t.y0 >= 0 && t.y0 < H; t.in_y1 = t.y0 + 1 >= 0 && t.y0 + 1 < H;
  return t;
}
SACB_DEVINL float taps_of_ones(const Taps& t) {
  float v = 0.f;
  if (t.in_y0 && t.in_x0) v += t.nw;
  if (t.in_y0 && t.in_x1) v += t.ne;
  if (t.in_y1 && t.in_x0) v += t.sw;
  if (t.in_y1 && t.in_x1) v += t.se;
  return v;
}

// ---------------------------------------------------------------- T1: teacher probabilities + class sums
// grid (ceil(HW/256), BT); probs scratch is pixel-major [BT][H*W][CP] with CP = C rounded up to 4
template <int C_>
__global__ void __launch_bounds__(256)
tail_probs_kernel(const float* __restrict__ logits, const int64_t* __restrict__ y, float* __restrict__ probs,
                  float* __restrict__ part_sums, int C, int CP, int h, int w, int H, int W, int stage) {
  const int b = blockIdx.y;
  const int pix = blockIdx.x * 256 + threadIdx.x;
  const int HW = H * W;
  float p[C_];
#pragma unroll
  for (int c = 0; c < C_; ++c) p[c] = 0.f;
  const bool valid = pix < HW;
  __shared__ float4 s_up4[UPS_FLOATS / 4];            // float4: 16-byte aligned records
  float* s_up = reinterpret_cast<float*>(s_up4);
  int r_lo;
  const float* Lb = logits + (size_t)b * C * h * w;
  const bool staged = stage_up_rows<C_>(s_up, stage ? UPS_FLOATS : 0, Lb, C, h, w, H, (blockIdx.x * 256) / W, (min(HW, blockIdx.x * 256 + 256) - 1) / W, r_lo);
  if (valid) {
    const int i = pix / W, j = pix - i * W;
    const UpCoef cy = up_coef(i, h, H), cx = up_coef(j, w, W);
    if (staged) up_logits_staged<C_>(s_up, w, r_lo, cy, cx, p);
    else up_logits<C_>(Lb, C, h, w, cy, cx, p);
    float mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < C_; ++c) if (c < C) mx = fmaxf(mx, p[c]);
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < C_; ++c) if (c < C) { p[c] = expf(p[c] - mx); sum += p[c]; }
#pragma unroll
    for (int c = 0; c < C_; ++c) if (c < C) p[c] = p[c] / sum;
    // sac.py:282 -- zero the probabilities inside augmentation padding (after the class sums, sac.py:278)
    const float keep = (y[(size_t)b * HW + pix] == -1) ? 0.f : 1.f;
    float* dst = probs + ((size_t)b * HW + pix) * CP;
    if (C == C_ && CP == (C_ + 3) / 4 * 4 && (reinterpret_cast<uintptr_t>(probs) & 15) == 0) {
      // one 80-byte record per pixel: five 16-byte stores instead of 19 scalar ones at an 80-byte lane stride (the pad
      // channel is written as 0; nobody reads it)
      constexpr int V = (C_ + 3) / 4;
      float4* dst4 = reinterpret_cast<float4*>(dst);
#pragma unroll
      for (int v = 0; v < V; ++v) {
        float4 o;
        o.x = p[4 * v] * keep;
        o.y = 4 * v + 1 < C_ ? p[(4 * v + 1) % C_] * keep : 0.f;
        o.z = 4 * v + 2 < C_ ? p[(4 * v + 2) % C_] * keep : 0.f;
        o.w = 4 * v + 3 < C_ ? p[(4 * v + 3) % C_] * keep : 0.f;
        dst4[v] = o;
      }
    } else {
#pragma unroll
      for (int c = 0; c < C_; ++c) if (c < C) dst[c] = p[c] * keep;
    }
  }
  // deterministic block reduction of the (unmasked) class sums
  __shared__ float red[8][C_];
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
#pragma unroll
  for (int c = 0; c < C_; ++c) {
    float v = p[c];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffff, v, o);
    if (lane == 0) red[wp][c] = v;
  }
  __syncthreads();
  if (threadIdx.x < C) {
    float s = 0.f;
    for (int k = 0; k < 8; ++k) s += red[k][threadIdx.x];
    part_sums[((size_t)b * gridDim.x + blockIdx.x) * C + threadIdx.x] = s;
  }
}

// running_conf update (sac.py:104-117): one block, thread = (class, slice); fixed-order double reduction
__global__ void __launch_bounds__(1024)
tail_running_conf_kernel(const float* __restrict__ part_sums, int nparts, int C, double inv_count, float beta,
                         float momentum, float* __restrict__ running_conf) {
  __shared__ double red[32][32];
  const int c = threadIdx.x, sl = threadIdx.y;       // blockDim = (32, 32)
  double s = 0.0;
  if (c < C)
    for (int k = sl; k < nparts; k += 32) s += (double)part_sums[(size_t)k * C + c];
  red[sl][c] = s;
  __syncthreads();
  if (sl == 0 && c < C) {
    double tot = 0.0;
    for (int k = 0; k < 32; ++k) tot += red[k][c];
    const float avg = (float)(tot * inv_count);
    float rc = running_conf[c];
    if (avg > 1e-8f && rc == beta) rc = avg;
    rc = rc * momentum;
    rc = rc + (1.f - momentum) * avg;
    running_conf[c] = rc;
  }
}

// ---------------------------------------------------------------- T2: warp to the reference frame + K-view average
// grid (ceil(HW/256), BT/T); pooled is [G][H*W][CP2] with channels 0..C-1 = averaged probs, channel C = mask
template <int C_>
__global__ void __launch_bounds__(256)
tail_pool_kernel(const float* __restrict__ probs, const float* __restrict__ affine, const float* __restrict__ affine_inv,
                 float* __restrict__ pooled, int T, int C, int CP, int CP2, int H, int W, int partial, int minent) {
  const int g = blockIdx.y;
  const int pix = blockIdx.x * 256 + threadIdx.x;
  const int HW = H * W;
  if (pix >= HW) return;
  const int i = pix / W, j = pix - i * W;
  float S[C_];
#pragma unroll
  for (int c = 0; c < C_; ++c) S[c] = 0.f;
  float best_ent = INFINITY, Zall = 0.f;     // CONF_POOL = minentropy_pool (sac.py:218-236)
  float Bst[C_];
#pragma unroll
  for (int c = 0; c < C_; ++c) Bst[c] = 0.f;
  for (int t = 0; t < T; ++t) {
    const int b = g * T + t;
    const Taps ta = make_taps(affine + b * 6, i, j, H, W);          // view -> reference (sac.py:289-290)
    const Taps tv = make_taps(affine_inv + b * 6, i, j, H, W);      // valid map = warp(ones, theta^-1) (sac.py:299-301)
    const float V = taps_of_ones(tv);
    const float* P = probs + (size_t)b * HW * CP;
    float A[C_];
#pragma unroll
    for (int c = 0; c < C_; ++c) A[c] = 0.f;
    const float wts[4] = {ta.nw, ta.ne, ta.sw, ta.se};
    const bool inb[4] = {ta.in_y0 && ta.in_x0, ta.in_y0 && ta.in_x1, ta.in_y1 && ta.in_x0, ta.in_y1 && ta.in_x1};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (!inb[k]) continue;
      const float* src = P + ((size_t)(ta.y0 + (k >> 1)) * W + (ta.x0 + (k & 1))) * CP;
#pragma unroll
      for (int c4 = 0; c4 < (C_ + 3) / 4; ++c4) {
        if (c4 * 4 < CP) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(src) + c4);
          if (c4 * 4 + 0 < C_) A[c4 * 4 + 0] += v.x * wts[k];
          if (c4 * 4 + 1 < C_) A[c4 * 4 + 1] += v.y * wts[k];
          if (c4 * 4 + 2 < C_) A[c4 * 4 + 2] += v.z * wts[k];
          if (c4 * 4 + 3 < C_) A[c4 * 4 + 3] += v.w * wts[k];
        }
      }
    }
    if (minent) {
      // entropy of this view's (aligned * valid) distribution (sac.py:189-196), views with no mass get 1/eps
      const float eps = 1e-5f;
      float ent = 0.f, z = 0.f;
#pragma unroll
      for (int c = 0; c < C_; ++c) if (c < C) {
        const float pv = A[c] * V;
        A[c] = pv;
        z += pv;
        ent -= pv * logf((pv + eps) / (1.f + eps));
      }
      if (z < 0.1f) ent = 1.f / eps;
      Zall += z;
      if (ent < best_ent) {                      // strict: the first view wins ties (torch.argmin)
        best_ent = ent;
#pragma unroll
        for (int c = 0; c < C_; ++c) Bst[c] = A[c];
      }
      continue;
    }
#pragma unroll
    for (int c = 0; c < C_; ++c) if (c < C) S[c] += A[c] * V;        // sac.py:305 aligned * valid, summed over T
  }
  float* dst = pooled + ((size_t)g * HW + pix) * CP2;
  if (minent) {
    // every view of the group receives the distribution of its min-entropy view; mask = total mass over views > 0.1
#pragma unroll
    for (int c = 0; c < C_; ++c) if (c < C) dst[c] = Bst[c];
    dst[C] = Zall > 0.1f ? 1.f : 0.f;
    return;
  }
  if (partial) {
    // fractional group (sac.py:198-216,243-245): this rank holds only T of the group's views; the un-normalised sums
    // are exchanged (sum over the ranks that share the group) and tail_pool_finalize_kernel normalises them
#pragma unroll
    for (int c = 0; c < C_; ++c) if (c < C) dst[c] = S[c];
    dst[C] = 0.f;
    return;
  }
  float Z = 0.f;
#pragma unroll
  for (int c = 0; c < C_; ++c) if (c < C) Z += S[c];
  const float mask = Z > 0.1f ? 1.f : 0.f;                           // sac.py:258
  const float denom = fmaxf(Z, 1e-3f);                               // sac.py:261
#pragma unroll
  for (int c = 0; c < C_; ++c) if (c < C) dst[c] = S[c] / denom;
  dst[C] = mask;
}

// second half of T2 for fractional groups: pooled holds the summed S over ALL views of the group (after the exchange)
template <int C_>
__global__ void __launch_bounds__(256)
tail_pool_finalize_kernel(float* __restrict__ pooled, int C, int CP2, size_t npix) {
  const size_t pix = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (pix >= npix) return;
  float* dst = pooled + pix * CP2;
  float S[C_];
  float Z = 0.f;
#pragma unroll
  for (int c = 0; c < C_; ++c) if (c < C) { S[c] = dst[c]; Z += S[c]; }
  const float mask = Z > 0.1f ? 1.f : 0.f;                           // sac.py:258
  const float denom = fmaxf(Z, 1e-3f);                               // sac.py:261
#pragma unroll
  for (int c = 0; c < C_; ++c) if (c < C) dst[c] = S[c] / denom;
  dst[C] = mask;
}

// ---------------------------------------------------------------- T3: warp back, max / argmax, per-(view, class) peaks
template <int C_>
__global__ void __launch_bounds__(256)
tail_refine_kernel(const float* __restrict__ pooled, const float* __restrict__ affine_inv, float* __restrict__ conf,
                   uint8_t* __restrict__ idx, int* __restrict__ peaks, float* __restrict__ refined, int T, int C,
                   int CP2, int H, int W, const float* __restrict__ probs_direct, int CP) {
  __shared__ int speak[C_];
  const int b = blockIdx.y;
  const int g = b / T;
  const int pix = blockIdx.x * 256 + threadIdx.x;
  const int HW = H * W;
  if (threadIdx.x < C_) speak[threadIdx.x] = 0;
  __syncthreads();
  if (pix < HW) {
    const int i = pix / W, j = pix - i * W;
    float R[C_];
#pragma unroll
    for (int c = 0; c < C_; ++c) R[c] = 0.f;
    float Mv = 0.f;
    if (probs_direct) {                                             // CONF_POOL_ON = False: _refine(pool=False), sac.py:284-285
      const float* src = probs_direct + ((size_t)b * HW + pix) * CP;
#pragma unroll
      for (int c = 0; c < C_; ++c) if (c < C) R[c] = src[c];
      Mv = 1.f;
    } else {
    const Taps t = make_taps(affine_inv + b * 6, i, j, H, W);       // reference -> view (sac.py:309-310)
    const float wts[4] = {t.nw, t.ne, t.sw, t.se};
    const bool inb[4] = {t.in_y0 && t.in_x0, t.in_y0 && t.in_x1, t.in_y1 && t.in_x0, t.in_y1 && t.in_x1};
    const float* base = pooled + (size_t)g * HW * CP2;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (!inb[k]) continue;
      const float* src = base + ((size_t)(t.y0 + (k >> 1)) * W + (t.x0 + (k & 1))) * CP2;
#pragma unroll
      for (int c4 = 0; c4 < (C_ + 4) / 4; ++c4) {
        if (c4 * 4 < CP2) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(src) + c4);
          const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int c = c4 * 4 + e;
            if (c < C_) { if (c < C) R[c] += vv[e] * wts[k]; }
            if (c == C) Mv += vv[e] * wts[k];
          }
        }
      }
    }
    }
    float best = -INFINITY; int bi = 0;
#pragma unroll
    for (int c = 0; c < C_; ++c) {
      if (c < C) {
        if (!probs_direct) R[c] *= Mv;                                 // sac.py:311
        if (R[c] > best) { best = R[c]; bi = c; }                      // first index on ties (torch.max)
      }
    }
    conf[(size_t)b * HW + pix] = best;
    idx[(size_t)b * HW + pix] = (uint8_t)bi;
    atomicMax(&speak[bi], __float_as_int(best));                       // best >= 0
    if (refined) {
#pragma unroll
      for (int c = 0; c < C_; ++c) if (c < C) refined[((size_t)b * C + c) * HW + pix] = R[c];
    }
  }
  __syncthreads();
  if (threadIdx.x < C && speak[threadIdx.x] != 0) atomicMax(&peaks[b * C + threadIdx.x], speak[threadIdx.x]);
}

// thresholds (sac.py:168-177): thr = clamp_min(peak * upper * (1 - exp(-rc/beta)), lower)
__global__ void tail_threshold_kernel(const int* __restrict__ peaks, const float* __restrict__ running_conf,
                                      float* __restrict__ thr, int BT, int C, float upper, float lower, float beta,
                                      int discount) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= BT * C) return;
  const int c = i % C;
  float t = __int_as_float(peaks[i]) * upper;
  if (discount) t *= 1.f - expf(-running_conf[c] / beta);
  thr[i] = fmaxf(t, lower);
}

// ---------------------------------------------------------------- T4: labels + batch-mean confidence
__global__ void __launch_bounds__(256)
tail_labels_kernel(const float* __restrict__ conf, const uint8_t* __restrict__ idx, const float* __restrict__ thr,
                   const int64_t* __restrict__ y, uint8_t* __restrict__ labels, float* __restrict__ conf_mean, int BT,
                   int C, int HW) {
  const int pix = blockIdx.x * 256 + threadIdx.x;
  if (pix >= HW) return;
  float s = 0.f;
  for (int b = 0; b < BT; ++b) {
    const size_t o = (size_t)b * HW + pix;
    const float cf = conf[o];
    const int id = idx[o];
    uint8_t lab = cf > thr[b * C + id] ? (uint8_t)id : (uint8_t)255;     // sac.py:175-181
    if (y[o] == -1) lab = 255;                                            // sac.py:185
    labels[o] = lab;
    s += cf;
  }
  conf_mean[pix] = s / (float)BT;
}

// ---------------------------------------------------------------- student loss forward
template <int C_>
__global__ void __launch_bounds__(256)
loss_fwd_kernel(const float* __restrict__ logits, const int64_t* __restrict__ y, const uint8_t* __restrict__ labels,
                const float* __restrict__ conf_mean, const float* __restrict__ running_conf, float focal_p,
                double* __restrict__ scratch, int C, int h, int w, int H, int W, int stage) {
  const int b = blockIdx.y;
  const int pix = blockIdx.x * 256 + threadIdx.x;
  const int HW = H * W;
  float l_ce = 0.f, l_self = 0.f;
  __shared__ float4 s_up4[UPS_FLOATS / 4];            // float4: 16-byte aligned records
  float* s_up = reinterpret_cast<float*>(s_up4);
  int r_lo;
  const float* Lb = logits + (size_t)b * C * h * w;
  const bool staged = stage_up_rows<C_>(s_up, stage ? UPS_FLOATS : 0, Lb, C, h, w, H, (blockIdx.x * 256) / W, (min(HW, blockIdx.x * 256 + 256) - 1) / W, r_lo);
  if (pix < HW) {
    const int i = pix / W, j = pix - i * W;
    float v[C_];
    if (staged) up_logits_staged<C_>(s_up, w, r_lo, up_coef(i, h, H), up_coef(j, w, W), v);
    else up_logits<C_>(Lb, C, h, w, up_coef(i, h, H), up_coef(j, w, W), v);
    float mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < C_; ++c) if (c < C) mx = fmaxf(mx, v[c]);
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < C_; ++c) if (c < C) sum += expf(v[c] - mx);
    const float lse = mx + logf(sum);
    const long long yy = y[(size_t)b * HW + pix];
    if (yy >= 0 && yy < C) {                      // deeplabv2.py:223-224 (ignore 255; -1 is mapped to 255 by sac.py:338)
      float t = 0.f;
#pragma unroll
      for (int c = 0; c < C_; ++c) if (c == (int)yy) t = v[c];
      l_ce = lse - t;
    }
    const int lab = labels ? labels[(size_t)b * HW + pix] : 255;
    if (lab < C) {
      float t = 0.f;
#pragma unroll
      for (int c = 0; c < C_; ++c) if (c == lab) t = v[c];
      const float base = 1.f - fmaxf(running_conf[lab], 0.f);          // sac.py:135
      const float fw = (focal_p == 3.f) ? base * base * base : powf(base, focal_p);
      l_self = fw * (lse - t) * conf_mean[pix];                          // sac.py:136,148
    }
  }
  __shared__ double red[2][8];
  double a = l_ce, s = l_self;
  for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffff, a, o); s += __shfl_xor_sync(0xffffffff, s, o); }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = a; red[1][threadIdx.x >> 5] = s; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ta = 0, ts = 0;
    for (int k = 0; k < 8; ++k) { ta += red[0][k]; ts += red[1][k]; }
    atomicAdd(&scratch[0], ta);
    atomicAdd(&scratch[1], ts);
  }
}
__global__ void loss_finalize_kernel(const double* __restrict__ scratch, float* __restrict__ losses, double inv_count) {
  if (threadIdx.x < 2) losses[threadIdx.x] = (float)(scratch[threadIdx.x] * inv_count);
}

// ---------------------------------------------------------------- student loss backward (gather form, deterministic)
// one warp per low-resolution logit pixel (b, yy, xx); lanes sweep the up-sampled pixels that touch it.
template <int C_>
__global__ void __launch_bounds__(256)
loss_bwd_kernel(const float* __restrict__ logits, const uint8_t* __restrict__ labels, const int64_t* __restrict__ y,
                const float* __restrict__ conf_mean, const float* __restrict__ running_conf, float focal_p, float coef,
                float* __restrict__ dlogits, int BT, int C, int h, int w, int H, int W) {
  // labels != NULL: gradient of self_ce (pseudo labels, focal weight, batch-mean confidence; sac.py:134-149)
  // labels == NULL: gradient of the plain loss_ce against y (deeplabv2.py:223-224; used by the source-domain pass)
  const int wid = (blockIdx.x * 256 + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (wid >= BT * h * w) return;
  const int b = wid / (h * w);
  const int rem = wid - b * h * w;
  const int yy = rem / w, xx = rem - yy * w;
  const float sy = H > 1 ? (float)(h - 1) / (float)(H - 1) : 0.f, sx = W > 1 ? (float)(w - 1) / (float)(W - 1) : 0.f;
  // conservative ranges of up-sampled rows / cols whose taps may include yy / xx
  const int i_lo = sy > 0.f ? max(0, (int)floorf((float)(yy - 1) / sy) - 1) : 0;
  const int i_hi = sy > 0.f ? min(H - 1, (int)ceilf((float)(yy + 1) / sy) + 1) : H - 1;
  const int j_lo = sx > 0.f ? max(0, (int)floorf((float)(xx - 1) / sx) - 1) : 0;
  const int j_hi = sx > 0.f ? min(W - 1, (int)ceilf((float)(xx + 1) / sx) + 1) : W - 1;
  const int nj = j_hi - j_lo + 1, ni = i_hi - i_lo + 1;
  float acc[C_];
#pragma unroll
  for (int c = 0; c < C_; ++c) acc[c] = 0.f;
  const float* L = logits + (size_t)b * C * h * w;
  const int HW = H * W;
  for (int t = lane; t < ni * nj; t += 32) {
    const int i = i_lo + t / nj, j = j_lo + t % nj;
    const UpCoef cy = up_coef(i, h, H), cx = up_coef(j, w, W);
    const float wy = (cy.i0 == yy ? cy.l0 : 0.f) + (cy.i1 == yy ? cy.l1 : 0.f);
    const float wx = (cx.i0 == xx ? cx.l0 : 0.f) + (cx.i1 == xx ? cx.l1 : 0.f);
    const float wgt = wy * wx;
    if (wgt == 0.f) continue;
    int lab;
    if (labels) lab = labels[(size_t)b * HW + i * W + j];
    else { const long long t = y[(size_t)b * HW + i * W + j]; lab = (t >= 0 && t < C) ? (int)t : 255; }
    if (lab >= C) continue;
    float v[C_];
    up_logits<C_>(L, C, h, w, cy, cx, v);
    float mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < C_; ++c) if (c < C) mx = fmaxf(mx, v[c]);
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < C_; ++c) if (c < C) { v[c] = expf(v[c] - mx); sum += v[c]; }
    float k = coef * wgt;
    if (labels) {
      const float base = 1.f - fmaxf(running_conf[lab], 0.f);
      const float fw = (focal_p == 3.f) ? base * base * base : powf(base, focal_p);
      k *= fw * conf_mean[i * W + j];
    }
    const float inv = 1.f / sum;
#pragma unroll
    for (int c = 0; c < C_; ++c) if (c < C) acc[c] += k * (v[c] * inv - (c == lab ? 1.f : 0.f));
  }
#pragma unroll
  for (int c = 0; c < C_; ++c) {
    float v = acc[c];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffff, v, o);
    if (lane == 0 && c < C) dlogits[((size_t)b * C + c) * h * w + rem] = v;
  }
}

// ---------------------------------------------------------------- student loss backward, two-stage form
// Stage A (one thread per up-sampled pixel): dL/d(logits_up)[b,c,i,j] = k * (softmax_c - [c == label]) written once as
// fp32 NCHW; stage B: the adjoint of the align_corners=True bilinear upsample, separable: rows first (H x W -> H x w),
// then columns (H x w -> h x w).  The gather form above recomputes the softmax of every up-sampled pixel for each of the
// (up to) four low-resolution pixels it touches, over a conservative window: 1.76 ms vs ~0.7 ms here at 24 x 19 x 512^2.
template <int C_>
__global__ void __launch_bounds__(256)
loss_grad_px_kernel(const float* __restrict__ logits, const uint8_t* __restrict__ labels, const int64_t* __restrict__ y,
                    const float* __restrict__ conf_mean, const float* __restrict__ running_conf, float focal_p, float coef,
                    float* __restrict__ g_px, int C, int h, int w, int H, int W) {
  const int b = blockIdx.y;
  const int pix = blockIdx.x * 256 + threadIdx.x;
  const int HW = H * W;
  if (pix >= HW) return;
  const int i = pix / W, j = pix - i * W;
  int lab;
  if (labels) lab = labels[(size_t)b * HW + pix];
  else { const long long t = y[(size_t)b * HW + pix]; lab = (t >= 0 && t < C) ? (int)t : 255; }
  float* dst = g_px + (size_t)b * C * HW + pix;
  if (lab >= C) {
#pragma unroll
    for (int c = 0; c < C_; ++c) if (c < C) dst[(size_t)c * HW] = 0.f;
    return;
  }
  float v[C_];
  up_logits<C_>(logits + (size_t)b * C * h * w, C, h, w, up_coef(i, h, H), up_coef(j, w, W), v);
  float mx = -INFINITY;
#pragma unroll
  for (int c = 0; c < C_; ++c) if (c < C) mx = fmaxf(mx, v[c]);
  float sum = 0.f;
#pragma unroll
  for (int c = 0; c < C_; ++c) if (c < C) { v[c] = expf(v[c] - mx); sum += v[c]; }
  float k = coef;
  if (labels) {
    const float base = 1.f - fmaxf(running_conf[lab], 0.f);
    const float fw = (focal_p == 3.f) ? base * base * base : powf(base, focal_p);
    k *= fw * conf_mean[pix];
  }
  const float inv = 1.f / sum;
#pragma unroll
  for (int c = 0; c < C_; ++c) if (c < C) dst[(size_t)c * HW] = k * (v[c] * inv - (c == lab ? 1.f : 0.f));
}
// rows: t[bc, i, xx] = sum_j wx(j, xx) g[bc, i, j]
__global__ void __launch_bounds__(256)
upsample_adj_rows_kernel(const float* __restrict__ g, float* __restrict__ t, size_t rows /* BC*H */, int w, int W) {
  const size_t total = rows * w;
  const float sx = W > 1 ? (float)(w - 1) / (float)(W - 1) : 0.f;
  for (size_t o = blockIdx.x * (size_t)blockDim.x + threadIdx.x; o < total; o += (size_t)gridDim.x * blockDim.x) {
    const int xx = (int)(o % w);
    const size_t r = o / w;
    const int j_lo = sx > 0.f ? max(0, (int)floorf((float)(xx - 1) / sx) - 1) : 0;
    const int j_hi = sx > 0.f ? min(W - 1, (int)ceilf((float)(xx + 1) / sx) + 1) : W - 1;
    const float* src = g + r * W;
    float acc = 0.f;
    for (int j = j_lo; j <= j_hi; ++j) {
      const UpCoef cx = up_coef(j, w, W);
      const float wx = (cx.i0 == xx ? cx.l0 : 0.f) + (cx.i1 == xx ? cx.l1 : 0.f);
      if (wx != 0.f) acc += wx * src[j];
    }
    t[o] = acc;
  }
}
// Stages A + B(rows) fused, one block per (image, up-sampled row): the per-pixel gradients of the row stay in shared memory
// and only their row adjoint t[b, c, i, 0..w) leaves the SM -- the full-resolution gradient (BT x C x H x W fp32 = 478 MB at
// 24 x 19 x 512^2) is never written to or read back from HBM.  Per-pixel arithmetic is loss_grad_px_kernel's, the adjoint sums
// the same terms in the same order as upsample_adj_rows_kernel: bit-identical results.  The block keeps classes
// [c_begin, c_end) of its row (static shared memory, 48 640 bytes): all 19 for rows up to 640 pixels, two passes of 10 + 9
// classes (softmax recomputed) up to 1216.
constexpr int GRAD_ROW_FLOATS = 19 * 640;
template <int C_>
__global__ void __launch_bounds__(256)
loss_grad_rows_kernel(const float* __restrict__ logits, const uint8_t* __restrict__ labels, const int64_t* __restrict__ y,
                      const float* __restrict__ conf_mean, const float* __restrict__ running_conf, float focal_p, float coef,
                      float* __restrict__ t, int C, int h, int w, int H, int W, int c_begin, int c_end) {
  __shared__ float s_g[GRAD_ROW_FLOATS];               // [c_end - c_begin][W]
  const int b = blockIdx.y, i = blockIdx.x;
  const int HW = H * W;
  const UpCoef cy = up_coef(i, h, H);
  for (int j = threadIdx.x; j < W; j += 256) {
    const int pix = i * W + j;
    int lab;
    if (labels) lab = labels[(size_t)b * HW + pix];
    else { const long long tt = y[(size_t)b * HW + pix]; lab = (tt >= 0 && tt < C) ? (int)tt : 255; }
    if (lab >= C) {
#pragma unroll
      for (int c = 0; c < C_; ++c) if (c >= c_begin && c < c_end) s_g[(c - c_begin) * W + j] = 0.f;
      continue;
    }
    float v[C_];
    up_logits<C_>(logits + (size_t)b * C * h * w, C, h, w, cy, up_coef(j, w, W), v);
    float mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < C_; ++c) if (c < C) mx = fmaxf(mx, v[c]);
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < C_; ++c) if (c < C) { v[c] = expf(v[c] - mx); sum += v[c]; }
    float k = coef;
    if (labels) {
      const float base = 1.f - fmaxf(running_conf[lab], 0.f);
      const float fw = (focal_p == 3.f) ? base * base * base : powf(base, focal_p);
      k *= fw * conf_mean[pix];
    }
    const float inv = 1.f / sum;
#pragma unroll
    for (int c = 0; c < C_; ++c) if (c >= c_begin && c < c_end) s_g[(c - c_begin) * W + j] = k * (v[c] * inv - (c == lab ? 1.f : 0.f));
  }
  __syncthreads();
  const float sx = W > 1 ? (float)(w - 1) / (float)(W - 1) : 0.f;
  for (int o = threadIdx.x; o < (c_end - c_begin) * w; o += 256) {
    const int cc = o / w, xx = o - cc * w;
    const int j_lo = sx > 0.f ? max(0, (int)floorf((float)(xx - 1) / sx) - 1) : 0;
    const int j_hi = sx > 0.f ? min(W - 1, (int)ceilf((float)(xx + 1) / sx) + 1) : W - 1;
    const float* src = s_g + cc * W;
    float acc = 0.f;
    for (int j = j_lo; j <= j_hi; ++j) {
      const UpCoef cx = up_coef(j, w, W);
      const float wx = (cx.i0 == xx ? cx.l0 : 0.f) + (cx.i1 == xx ? cx.l1 : 0.f);
      if (wx != 0.f) acc += wx * src[j];
    }
    t[(((size_t)b * C + c_begin + cc) * H + i) * w + xx] = acc;
  }
}
// columns: out[bc, yy, xx] = sum_i wy(i, yy) t[bc, i, xx]
__global__ void __launch_bounds__(256)
upsample_adj_cols_kernel(const float* __restrict__ t, float* __restrict__ out, int BC, int h, int w, int H) {
  const size_t total = (size_t)BC * h * w;
  const float sy = H > 1 ? (float)(h - 1) / (float)(H - 1) : 0.f;
  for (size_t o = blockIdx.x * (size_t)blockDim.x + threadIdx.x; o < total; o += (size_t)gridDim.x * blockDim.x) {
    const int xx = (int)(o % w);
    const int yy = (int)((o / w) % h);
    const size_t bc = o / ((size_t)w * h);
    const int i_lo = sy > 0.f ? max(0, (int)floorf((float)(yy - 1) / sy) - 1) : 0;
    const int i_hi = sy > 0.f ? min(H - 1, (int)ceilf((float)(yy + 1) / sy) + 1) : H - 1;
    const float* src = t + bc * H * w + xx;
    float acc = 0.f;
    for (int i = i_lo; i <= i_hi; ++i) {
      const UpCoef cy = up_coef(i, h, H);
      const float wy = (cy.i0 == yy ? cy.l0 : 0.f) + (cy.i1 == yy ? cy.l1 : 0.f);
      if (wy != 0.f) acc += wy * src[(size_t)i * w];
    }
    out[o] = acc;
  }
}

__global__ void upsample_kernel(const float* __restrict__ in, float* __restrict__ out, int BC, int h, int w, int H, int W) {
  const size_t total = (size_t)BC * H * W;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int j = (int)(t % W);
    const int i = (int)((t / W) % H);
    const size_t bc = t / ((size_t)W * H);
    const UpCoef cy = up_coef(i, h, H), cx = up_coef(j, w, W);
    const float* p = in + bc * h * w;
    out[t] = cy.l0 * (cx.l0 * p[cy.i0 * w + cx.i0] + cx.l1 * p[cy.i0 * w + cx.i1]) +
             cy.l1 * (cx.l0 * p[cy.i1 * w + cx.i0] + cx.l1 * p[cy.i1 * w + cx.i1]);
  }
}

// out = up(in) (+ addend): bilinear, align_corners=True, [BC,h,w] -> [BC,H,W]  (fcn.py:107-109 up_x2 and score fusion :111-134)
__global__ void upsample_add_kernel(const float* __restrict__ in, const float* __restrict__ addend, float* __restrict__ out,
                                    int BC, int h, int w, int H, int W) {
  const size_t total = (size_t)BC * H * W;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int j = (int)(t % W);
    const int i = (int)((t / W) % H);
    const size_t bc = t / ((size_t)W * H);
    const UpCoef cy = up_coef(i, h, H), cx = up_coef(j, w, W);
    const float* p = in + bc * h * w;
    float v = cy.l0 * (cx.l0 * p[cy.i0 * w + cx.i0] + cx.l1 * p[cy.i0 * w + cx.i1]) +
              cy.l1 * (cx.l0 * p[cy.i1 * w + cx.i0] + cx.l1 * p[cy.i1 * w + cx.i1]);
    if (addend) v += addend[t];
    out[t] = v;
  }
}
// adjoint of the bilinear align_corners=True upsample: g_in[bc,y,x] = sum_{(i,j)} wy(i,y) wx(j,x) g_out[bc,i,j]  (gather form)
__global__ void upsample_bwd_kernel(const float* __restrict__ g_out, float* __restrict__ g_in, int BC, int h, int w, int H,
                                    int W) {
  const size_t total = (size_t)BC * h * w;
  const float sy = H > 1 ? (float)(h - 1) / (float)(H - 1) : 0.f, sx = W > 1 ? (float)(w - 1) / (float)(W - 1) : 0.f;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int xx = (int)(t % w);
    const int yy = (int)((t / w) % h);
    const size_t bc = t / ((size_t)w * h);
    const int i_lo = sy > 0.f ? max(0, (int)floorf((float)(yy - 1) / sy) - 1) : 0;
    const int i_hi = sy > 0.f ? min(H - 1, (int)ceilf((float)(yy + 1) / sy) + 1) : H - 1;
    const int j_lo = sx > 0.f ? max(0, (int)floorf((float)(xx - 1) / sx) - 1) : 0;
    const int j_hi = sx > 0.f ? min(W - 1, (int)ceilf((float)(xx + 1) / sx) + 1) : W - 1;
    const float* g = g_out + bc * H * W;
    float acc = 0.f;
    for (int i = i_lo; i <= i_hi; ++i) {
      const UpCoef cy = up_coef(i, h, H);
      const float wy = (cy.i0 == yy ? cy.l0 : 0.f) + (cy.i1 == yy ? cy.l1 : 0.f);
      if (wy == 0.f) continue;
      for (int j = j_lo; j <= j_hi; ++j) {
        const UpCoef cx = up_coef(j, w, W);
        const float wx = (cx.i0 == xx ? cx.l0 : 0.f) + (cx.i1 == xx ? cx.l1 : 0.f);
        if (wx != 0.f) acc += wy * wx * g[i * W + j];
      }
    }
    g_in[t] = acc;
  }
}
// fp32 NCHW [N,C,P,Q] -> split planes NHWC [N,P,Q,Cp] (channels >= C zero): feeds the GEMM gradients of the 19-class score convs
__global__ void nchw_to_planes_kernel(const float* __restrict__ g, uint16_t* __restrict__ hi, uint16_t* __restrict__ lo, int N,
                                      int C, int PQ, int Cp) {
  const size_t total = (size_t)N * PQ * Cp;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(t % Cp);
    const size_t pix = t / Cp;
    const int n = (int)(pix / PQ);
    const int r = (int)(pix - (size_t)n * PQ);
    const float v = c < C ? g[((size_t)n * C + c) * PQ + r] : 0.f;
    const uint16_t h = float_to_bf16_bits(v);
    hi[t] = h;
    lo[t] = float_to_bf16_bits(v - bf16_bits_to_float(h));
  }
}
// Dropout2d (fcn.py:52,56) on split planes: x[n,p,q,c] *= m[n,c]   (m = 0 or 1/(1-p), drawn by the caller)
__global__ void channel_scale_kernel(uint16_t* __restrict__ hi, uint16_t* __restrict__ lo, const float* __restrict__ m, int N,
                                     int PQ, int C) {
  const size_t total = (size_t)N * PQ * C;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(t % C);
    const int n = (int)(t / ((size_t)PQ * C));
    const float v = (bf16_bits_to_float(hi[t]) + bf16_bits_to_float(lo[t])) * m[(size_t)n * C + c];
    const uint16_t h = float_to_bf16_bits(v);
    hi[t] = h;
    lo[t] = float_to_bf16_bits(v - bf16_bits_to_float(h));
  }
}

// ---------------------------------------------------------------- multi-tensor EMA / norm and SGD
constexpr int SEG_CHUNKS = 16;
__global__ void __launch_bounds__(256)
ema_norm_kernel(float* __restrict__ slow, const float* __restrict__ fast, const int64_t* __restrict__ offs, float m,
                int update, float* __restrict__ seg_sq) {
  const int seg = blockIdx.x;
  const int64_t b0 = offs[2 * seg], b1 = offs[2 * seg + 1];
  const int64_t n = b1 - b0;
  const int64_t per = (n + SEG_CHUNKS - 1) / SEG_CHUNKS;
  const int64_t s0 = b0 + per * blockIdx.y, s1 = min(b1, s0 + per);
  float acc = 0.f;
  for (int64_t i = s0 + threadIdx.x; i < s1; i += 256) {
    const float sv = slow[i], fv = fast[i];
    const float d = sv - fv;
    acc = fmaf(d, d, acc);
    if (update) slow[i] = sv * m + fv * (1.f - m);          // sac.py:95-97
  }
  __shared__ float red[8];
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffff, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int k = 0; k < 8; ++k) s += red[k];
    if (s != 0.f) atomicAdd(&seg_sq[seg], s);
  }
}
__global__ void ema_norm_finalize_kernel(const float* __restrict__ seg_sq, int nseg, float* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < nseg; ++i) s += sqrtf(seg_sq[i]);      // diff_sum += torch.norm(...) in key order (sac.py:93,100)
    out[0] = s;
  }
}
// One (segment, chunk) per block.  A segment is one parameter tensor (lr / weight decay differ per tensor); the largest
// (3x3 512->512: 2.4 M elements) used to be walked by 16 blocks with scalar accesses -- 576 dependent round trips per thread,
// 415 us for the 0.8 GB the step touches (24 % of the HBM roofline, ncu profiles/stream_kernels_r1p.txt).  Now 64 chunks per
// segment, 16-byte accesses whenever the chunk is aligned: the arithmetic per element is unchanged (bit-identical update).
constexpr int SGD_CHUNKS = 64;
SACB_DEVINL void sgd_update(float gv, float pv, float mv, float w, float mu, float l, int first, float& mo, float& po) {
  float d = gv;
  if (w != 0.f) d = fmaf(w, pv, d);                          // grad.add(param, alpha=weight_decay)
  const float buf = first ? d : fmaf(mu, mv, d);             // buf.mul_(momentum).add_(grad)
  mo = buf;
  po = pv - l * buf;                                         // param.add_(buf, alpha=-lr)
}
__global__ void __launch_bounds__(256)
sgd_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ mom, const int64_t* __restrict__ offs,
           const float* __restrict__ lr, const float* __restrict__ wd, float mu, int first) {
  const int seg = blockIdx.x;
  const int64_t b0 = offs[2 * seg], b1 = offs[2 * seg + 1];
  int64_t per = (b1 - b0 + SGD_CHUNKS - 1) / SGD_CHUNKS;
  per = (per + 3) & ~(int64_t)3;                             // chunk starts stay 16-byte aligned when the segment start is
  const int64_t s0 = b0 + per * blockIdx.y, s1 = min(b1, s0 + per);
  if (s0 >= s1) return;
  const float l = lr[seg], w = wd[seg];
  int64_t i = s0;
  const bool aligned = (s0 & 3) == 0 &&
      ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(mom)) & 15) == 0;
  if (aligned) {
    const int64_t n4 = (s1 - s0) >> 2;
    float4* p4 = reinterpret_cast<float4*>(p + s0);
    const float4* g4 = reinterpret_cast<const float4*>(g + s0);
    float4* m4 = reinterpret_cast<float4*>(mom + s0);
#pragma unroll 2
    for (int64_t k = threadIdx.x; k < n4; k += 256) {
      const float4 gv = g4[k], pv = p4[k];
      const float4 mv = first ? make_float4(0.f, 0.f, 0.f, 0.f) : m4[k];
      float4 mo, po;
      sgd_update(gv.x, pv.x, mv.x, w, mu, l, first, mo.x, po.x);
      sgd_update(gv.y, pv.y, mv.y, w, mu, l, first, mo.y, po.y);
      sgd_update(gv.z, pv.z, mv.z, w, mu, l, first, mo.z, po.z);
      sgd_update(gv.w, pv.w, mv.w, w, mu, l, first, mo.w, po.w);
      m4[k] = mo;
      p4[k] = po;
    }
    i = s0 + (n4 << 2);
  }
  for (i += threadIdx.x; i < s1; i += 256) {
    float mo, po;
    sgd_update(g[i], p[i], first ? 0.f : mom[i], w, mu, l, first, mo, po);
    mom[i] = mo;
    p[i] = po;
  }
}

static inline int grid1(size_t n, int block) {
  size_t g = (n + block - 1) / block;
  const size_t cap = 148 * 16;
  return (int)(g < cap ? (g ? g : 1) : cap);
}
}  // namespace sacb

using namespace sacb;
// A/B switch (read once): SACB_UP_STAGED=0 -> tail_probs / loss_fwd read the low-resolution logits from global memory again.
// Both forms are bit-identical (tests/test_tail_loss_variants_gpu.py); the switch exists for that test and for timing runs.
static int up_staged() { static const int v = [] { const char* e = getenv("SACB_UP_STAGED"); return (e && e[0] == '0') ? 0 : 1; }(); return v; }

#define ST ((cudaStream_t)stream)
#define LAUNCHED() do { g_launches++; SACB_CHECK_CUDA(cudaGetLastError()); } while (0)

extern "C" size_t sacb_tail_part_sums_elems(int BT, int C, int H, int W) {
  return (size_t)BT * ((H * W + 255) / 256) * C;
}
/* workspace sizes (elements of float): probs [BT*H*W*CP], pooled [BT/T*H*W*CP2] */
extern "C" size_t sacb_tail_probs_elems(int BT, int C, int H, int W) { return (size_t)BT * H * W * ((C + 3) / 4 * 4); }
extern "C" size_t sacb_tail_pooled_elems(int G, int C, int H, int W) { return (size_t)G * H * W * ((C + 1 + 3) / 4 * 4); }

extern "C" int sacb_teacher_tail(const SacbTail* d, void* stream) {
  SACB_REQUIRE(d && d->size == sizeof(SacbTail), "sacb_teacher_tail: bad descriptor size");
  SACB_REQUIRE(d->C == 19, "sacb_teacher_tail: built for 19 classes (got %d)", d->C);
  SACB_REQUIRE(d->BT % d->T == 0, "sacb_teacher_tail: BT %% T != 0");
  constexpr int C_ = 19;
  const int C = d->C, HW = d->H * d->W, CP = (C + 3) / 4 * 4, CP2 = (C + 1 + 3) / 4 * 4;
  const int nb = (HW + 255) / 256;
  dim3 gridB(nb, d->BT), gridG(nb, d->BT / d->T);
  SACB_REQUIRE(d->phase >= 0 && d->phase <= 3, "sacb_teacher_tail: phase must be 0, 1, 2 or 3");
  if (d->phase == 3) {
    // diagnostics only: `pooled` (and `probs` when pooling is off) still hold this forward's normalised reference-frame
    // probabilities; warp them back once more to materialise `refined` (sac.py:309-311).  conf / idx / peaks are rewritten with
    // the values they already hold (atomicMax onto itself); running_conf, thresholds, labels and conf_mean are not touched and no
    // exchange between ranks is needed.
    SACB_REQUIRE(d->refined != nullptr, "sacb_teacher_tail: phase 3 materialises `refined`");
    tail_refine_kernel<C_><<<gridB, 256, 0, ST>>>(d->pooled, d->affine_inv, d->conf, d->idx, reinterpret_cast<int*>(d->peaks),
                                                 d->refined, d->T, C, CP2, d->H, d->W, d->pool_mode == 2 ? d->probs : nullptr, CP);
    LAUNCHED();
    return 0;
  }
  SACB_REQUIRE(d->pool_mode >= 0 && d->pool_mode <= 2, "sacb_teacher_tail: pool_mode must be 0 (avg), 1 (min-entropy) or 2 (off)");
  SACB_REQUIRE(d->pool_mode == 0 || d->phase == 0 || d->phase == 3, "sacb_teacher_tail: fractional groups need the average pool");
  if (d->phase != 2) {
    tail_probs_kernel<C_><<<gridB, 256, 0, ST>>>(d->teacher_logits, d->y, d->probs, d->part_sums, C, CP, d->h, d->w, d->H, d->W, up_staged());
    LAUNCHED();
    if (d->training) {
      tail_running_conf_kernel<<<1, dim3(32, 32), 0, ST>>>(d->part_sums, nb * d->BT, C, 1.0 / ((double)d->BT * HW), d->beta,
                                                d->stat_momentum, d->running_conf);
      LAUNCHED();
    }
    if (d->pool_mode != 2) {
      tail_pool_kernel<C_><<<gridG, 256, 0, ST>>>(d->probs, d->affine, d->affine_inv, d->pooled, d->T, C, CP, CP2, d->H, d->W,
                                                 d->phase == 1, d->pool_mode == 1);
      LAUNCHED();
    }
    if (d->phase == 1) return 0;          // caller sums `pooled` over the ranks that share the group, then phase 2
  } else {
    const size_t npix = (size_t)(d->BT / d->T) * HW;
    tail_pool_finalize_kernel<C_><<<(unsigned)((npix + 255) / 256), 256, 0, ST>>>(d->pooled, C, CP2, npix);
    LAUNCHED();
  }
  SACB_CHECK_CUDA(cudaMemsetAsync(d->peaks, 0, sizeof(float) * d->BT * C, ST));
  tail_refine_kernel<C_><<<gridB, 256, 0, ST>>>(d->pooled, d->affine_inv, d->conf, d->idx, reinterpret_cast<int*>(d->peaks),
                                               d->refined, d->T, C, CP2, d->H, d->W, d->pool_mode == 2 ? d->probs : nullptr, CP);
  LAUNCHED();
  tail_threshold_kernel<<<(d->BT * C + 127) / 128, 128, 0, ST>>>(reinterpret_cast<const int*>(d->peaks), d->running_conf,
                                                               d->thresholds, d->BT, C, d->conf_upper, d->conf_lower,
                                                               d->beta, d->discount);
  LAUNCHED();
  tail_labels_kernel<<<nb, 256, 0, ST>>>(d->conf, d->idx, d->thresholds, d->y, d->labels, d->conf_mean, d->BT, C, HW);
  LAUNCHED();
  return 0;
}

extern "C" int sacb_student_loss_fwd(const SacbLoss* d, void* stream) {
  SACB_REQUIRE(d && d->size == sizeof(SacbLoss), "sacb_student_loss_fwd: bad descriptor size");
  SACB_REQUIRE(d->C == 19, "sacb_student_loss_fwd: built for 19 classes");
  const int HW = d->H * d->W;
  SACB_CHECK_CUDA(cudaMemsetAsync(d->scratch, 0, 2 * sizeof(double), ST));
  dim3 grid((HW + 255) / 256, d->BT);
  loss_fwd_kernel<19><<<grid, 256, 0, ST>>>(d->logits, d->y, d->labels, d->conf_mean, d->running_conf, d->focal_p,
                                           d->scratch, d->C, d->h, d->w, d->H, d->W, up_staged());
  LAUNCHED();
  loss_finalize_kernel<<<1, 32, 0, ST>>>(d->scratch, d->losses, 1.0 / ((double)d->BT * HW));
  LAUNCHED();
  return 0;
}

extern "C" int sacb_student_loss_bwd(const SacbLoss* d, void* stream) {
  SACB_REQUIRE(d && d->size == sizeof(SacbLoss), "sacb_student_loss_bwd: bad descriptor size");
  SACB_REQUIRE(d->C == 19 && d->dlogits, "sacb_student_loss_bwd: built for 19 classes, needs dlogits");
  const int HW = d->H * d->W;
  const float coef = d->grad_scale / (float)((double)d->BT * HW);
  if (d->grad_rows && d->C == 19 && 10 * d->W <= GRAD_ROW_FLOATS) {
    // fused form: per-pixel gradient + row adjoint in one pass over the row (no full-resolution gradient in HBM), then the
    // column adjoint.  One pass with all classes when they fit the block's shared memory, else two passes of 10 + 9 classes.
    const int split = d->C * d->W <= GRAD_ROW_FLOATS ? d->C : 10;
    for (int c0 = 0; c0 < d->C; c0 += split) {
      loss_grad_rows_kernel<19><<<dim3(d->H, d->BT), 256, 0, ST>>>(d->logits, d->labels, d->y, d->conf_mean, d->running_conf, d->focal_p,
                                                                   coef, d->grad_rows, d->C, d->h, d->w, d->H, d->W, c0,
                                                                   c0 + split < d->C ? c0 + split : d->C);
      LAUNCHED();
    }
    upsample_adj_cols_kernel<<<grid1((size_t)d->BT * d->C * d->h * d->w, 256), 256, 0, ST>>>(d->grad_rows, d->dlogits, d->BT * d->C,
                                                                                             d->h, d->w, d->H);
    LAUNCHED();
    return 0;
  }
  if (d->grad_px && d->grad_rows) {                 // two-stage form (workspace provided)
    dim3 grid((HW + 255) / 256, d->BT);
    loss_grad_px_kernel<19><<<grid, 256, 0, ST>>>(d->logits, d->labels, d->y, d->conf_mean, d->running_conf, d->focal_p, coef,
                                                 d->grad_px, d->C, d->h, d->w, d->H, d->W);
    LAUNCHED();
    const size_t rows = (size_t)d->BT * d->C * d->H;
    upsample_adj_rows_kernel<<<grid1(rows * d->w, 256), 256, 0, ST>>>(d->grad_px, d->grad_rows, rows, d->w, d->W);
    LAUNCHED();
    upsample_adj_cols_kernel<<<grid1((size_t)d->BT * d->C * d->h * d->w, 256), 256, 0, ST>>>(d->grad_rows, d->dlogits, d->BT * d->C,
                                                                                             d->h, d->w, d->H);
    LAUNCHED();
    return 0;
  }
  const size_t warps = (size_t)d->BT * d->h * d->w;
  loss_bwd_kernel<19><<<(unsigned)((warps * 32 + 255) / 256), 256, 0, ST>>>(d->logits, d->labels, d->y, d->conf_mean,
                                                                          d->running_conf, d->focal_p, coef, d->dlogits,
                                                                          d->BT, d->C, d->h, d->w, d->H, d->W);
  LAUNCHED();
  return 0;
}

extern "C" int sacb_upsample(const float* in, float* out, int B, int C, int h, int w, int H, int W, void* stream) {
  const size_t total = (size_t)B * C * H * W;
  upsample_kernel<<<grid1(total, 256), 256, 0, ST>>>(in, out, B * C, h, w, H, W);
  LAUNCHED();
  return 0;
}

extern "C" int sacb_upsample_add(const float* in, const float* addend, float* out, int B, int C, int h, int w, int H, int W,
                                 void* stream) {
  const size_t total = (size_t)B * C * H * W;
  upsample_add_kernel<<<grid1(total, 256), 256, 0, ST>>>(in, addend, out, B * C, h, w, H, W);
  LAUNCHED();
  return 0;
}
extern "C" int sacb_upsample_bwd(const float* g_out, float* g_in, int B, int C, int h, int w, int H, int W, void* stream) {
  const size_t total = (size_t)B * C * h * w;
  upsample_bwd_kernel<<<grid1(total, 256), 256, 0, ST>>>(g_out, g_in, B * C, h, w, H, W);
  LAUNCHED();
  return 0;
}
extern "C" int sacb_nchw_to_planes(const float* g, void* hi, void* lo, int N, int C, int P, int Q, int Cp, void* stream) {
  const size_t total = (size_t)N * P * Q * Cp;
  nchw_to_planes_kernel<<<grid1(total, 256), 256, 0, ST>>>(g, (uint16_t*)hi, (uint16_t*)lo, N, C, P * Q, Cp);
  LAUNCHED();
  return 0;
}
extern "C" int sacb_channel_scale(void* hi, void* lo, const float* m, int N, int P, int Q, int C, void* stream) {
  const size_t total = (size_t)N * P * Q * C;
  channel_scale_kernel<<<grid1(total, 256), 256, 0, ST>>>((uint16_t*)hi, (uint16_t*)lo, m, N, P * Q, C);
  LAUNCHED();
  return 0;
}

extern "C" int sacb_ema_norm(float* slow, const float* fast, const int64_t* seg_offsets, int nseg, float momentum,
                             int update, float* seg_sq, float* out, void* stream) {
  SACB_CHECK_CUDA(cudaMemsetAsync(seg_sq, 0, sizeof(float) * nseg, ST));
  ema_norm_kernel<<<dim3(nseg, SEG_CHUNKS), 256, 0, ST>>>(slow, fast, seg_offsets, momentum, update, seg_sq);
  LAUNCHED();
  ema_norm_finalize_kernel<<<1, 32, 0, ST>>>(seg_sq, nseg, out);
  LAUNCHED();
  return 0;
}

extern "C" int sacb_sgd(float* p, const float* g, float* mom, const int64_t* seg_offsets, const float* seg_lr,
                        const float* seg_wd, int nseg, float momentum, int first_step, void* stream) {
  sgd_kernel<<<dim3(nseg, SGD_CHUNKS), 256, 0, ST>>>(p, g, mom, seg_offsets, seg_lr, seg_wd, momentum, first_step);
  LAUNCHED();
  return 0;
}
