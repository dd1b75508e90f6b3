#line 1 "../../da_sac_b200/csrc/sacb_aspp.cu"
// ASPP head (Classifier_Module, /root/reference/models/deeplabv2.py:101-116) as ONE tap-unrolled 1x1 GEMM.
//
// The head is the sum of four 3x3 dilated convs 2048 -> 19.  As an implicit GEMM it is hopelessly skinny
// (N = 19) and re-reads the 830 MB activation 36 times.  Because the output has only 19 channels we instead
// compute, for every pixel, the contribution of that pixel to all 36 taps at once,
//     Z[pix, (i*9 + r*3 + s)*19 + k] = sum_c X[pix, c] * W_i[k, c, r, s]        (one dense GEMM, N = 684 -> 768)
// and then shift-and-add the 19-channel slices:  out[p] = sum_taps Z[p + off_tap, tap].  X is read once.
// The backward pass uses the transposed trick: Gcol[pix, tap*19 + k] = g[pix - off_tap, k] turns both the data
// gradient (Gcol x W_all) and the filter gradient (Gcol^T x X) into plain GEMMs as well.
#include <atomic>
// (sacb_common.cuh: see cuda_emul.h)
#include "../../include/sacb.h"

namespace sacb {
extern std::atomic<long long> g_launches;

constexpr int ASPP_K = 19, ASPP_TAPS = 36, ASPP_J = ASPP_K * ASPP_TAPS;   // 684

struct Ptr4 { const float* p[4]; };
struct MPtr4 { float* p[4]; };
struct Dil4 { int d[4]; };

// wf[j][c] (fprop planes, [J_pad][C]) and wt[c][j] (dgrad planes, [C][J_pad]); j = (i*9 + rs)*19 + k
__global__ void aspp_pack_kernel(Ptr4 w, int C, int Jp, uint16_t* __restrict__ wf_hi, uint16_t* __restrict__ wf_lo,
                                 uint16_t* __restrict__ wt_hi, uint16_t* __restrict__ wt_lo) {
  const size_t total = (size_t)Jp * C;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(t % C);
    const int j = (int)(t / C);
    float v = 0.f;
    if (j < ASPP_J) {
      const int k = j % ASPP_K, tap = j / ASPP_K;
      const int i = tap / 9, rs = tap - i * 9;
      v = w.p[i][((size_t)k * C + c) * 9 + rs];
    }
    const uint16_t h = float_to_bf16_bits(v);
    const uint16_t l = float_to_bf16_bits(v - bf16_bits_to_float(h));
    wf_hi[t] = h; wf_lo[t] = l;
    if (wt_hi) { wt_hi[(size_t)c * Jp + j] = h; wt_lo[(size_t)c * Jp + j] = l; }
  }
}

// one warp per output pixel, lane = class; every Z element is read exactly once
__global__ void __launch_bounds__(256)
aspp_gather_kernel(const float* __restrict__ Z, Ptr4 bias, Dil4 dil, float* __restrict__ out, int N, int P, int Q, int Jp) {
  const int wid = (blockIdx.x * 256 + threadIdx.x) >> 5;
  const int k = threadIdx.x & 31;
  const int M = N * P * Q;
  if (wid >= M || k >= ASPP_K) return;
  const int n = wid / (P * Q);
  const int rem = wid - n * P * Q;
  const int p = rem / Q, q = rem - p * Q;
  float acc = bias.p[0][k];
  acc += bias.p[1][k]; acc += bias.p[2][k]; acc += bias.p[3][k];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int d = dil.d[i];
#pragma unroll
    for (int rs = 0; rs < 9; ++rs) {
      const int pp = p + (rs / 3 - 1) * d, qq = q + (rs % 3 - 1) * d;
      if (pp >= 0 && pp < P && qq >= 0 && qq < Q)
        acc += __ldg(Z + ((size_t)(n * P + pp) * Q + qq) * Jp + (i * 9 + rs) * ASPP_K + k);
    }
  }
  out[((size_t)n * ASPP_K + k) * P * Q + rem] = acc;
}

// Gcol[pix, (i*9+rs)*19 + k] = g[n, k, p - (r-1)d_i, q - (s-1)d_i]  (0 outside / padded columns); split planes
__global__ void __launch_bounds__(256)
aspp_gcol_kernel(const float* __restrict__ g, Dil4 dil, uint16_t* __restrict__ hi, uint16_t* __restrict__ lo, int N, int P,
                 int Q, int Jp) {
  const size_t total = (size_t)N * P * Q * (Jp / 8);
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int jv = (int)(t % (Jp / 8));
    const int pix = (int)(t / (Jp / 8));
    const int n = pix / (P * Q);
    const int rem = pix - n * P * Q;
    const int p = rem / Q, q = rem - p * Q;
    uint32_t ph[4], pl[4];
#pragma unroll
    for (int e2 = 0; e2 < 4; ++e2) {
      float v[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int j = jv * 8 + e2 * 2 + e;
        float x = 0.f;
        if (j < ASPP_J) {
          const int k = j % ASPP_K, tap = j / ASPP_K;
          const int i = tap / 9, rs = tap - i * 9;
          const int pp = p - (rs / 3 - 1) * dil.d[i], qq = q - (rs % 3 - 1) * dil.d[i];
          if (pp >= 0 && pp < P && qq >= 0 && qq < Q) x = __ldg(g + ((size_t)(n * ASPP_K + k) * P + pp) * Q + qq);
        }
        v[e] = x;
      }
      const __nv_bfloat162 h = __floats2bfloat162_rn(v[0], v[1]);
      const float2 hf = __bfloat1622float2(h);
      const __nv_bfloat162 l = __floats2bfloat162_rn(v[0] - hf.x, v[1] - hf.y);
      ph[e2] = *reinterpret_cast<const uint32_t*>(&h);
      pl[e2] = *reinterpret_cast<const uint32_t*>(&l);
    }
    reinterpret_cast<uint4*>(hi)[t] = make_uint4(ph[0], ph[1], ph[2], ph[3]);
    reinterpret_cast<uint4*>(lo)[t] = make_uint4(pl[0], pl[1], pl[2], pl[3]);
  }
}

// dw_i[k][c][rs] = sum_split parts[split][(i*9+rs)*19 + k][c]
__global__ void aspp_unpack_wgrad_kernel(const float* __restrict__ parts, int splits, int Jp, int C, MPtr4 dw) {
  const size_t total = (size_t)ASPP_J * C;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(t % C);
    const int j = (int)(t / C);
    float s = 0.f;
    for (int sp = 0; sp < splits; ++sp) s += parts[((size_t)sp * Jp + j) * C + c];
    const int k = j % ASPP_K, tap = j / ASPP_K;
    const int i = tap / 9, rs = tap - i * 9;
    dw.p[i][((size_t)k * C + c) * 9 + rs] = s;
  }
}

static inline int grid1(size_t n, int block) {
  size_t g = (n + block - 1) / block;
  const size_t cap = 148 * 16;
  return (int)(g < cap ? (g ? g : 1) : cap);
}
}  // namespace sacb

using namespace sacb;
#define ST ((cudaStream_t)stream)
#define LAUNCHED() do { g_launches++; SACB_CHECK_CUDA(cudaGetLastError()); } while (0)

extern "C" int sacb_aspp_jpad(void) { return 768; }

extern "C" int sacb_aspp_pack_weights(const float* w0, const float* w1, const float* w2, const float* w3, int C,
                                      void* wf_hi, void* wf_lo, void* wt_hi, void* wt_lo, void* stream) {
  Ptr4 w = {{w0, w1, w2, w3}};
  const int Jp = 768;
  cuda_emul::run_grid("aspp_pack_kernel", grid1((size_t)Jp * C, 256), 256, 0, false, [&]() { aspp_pack_kernel(w, C, Jp, (uint16_t*)wf_hi, (uint16_t*)wf_lo,
                                                              (uint16_t*)wt_hi, (uint16_t*)wt_lo); });
  LAUNCHED();
  return 0;
}

extern "C" int sacb_aspp_gather(const float* Z, const float* b0, const float* b1, const float* b2, const float* b3,
                                const int32_t* dil4, float* out_nchw, int N, int P, int Q, void* stream) {
  Ptr4 b = {{b0, b1, b2, b3}};
  Dil4 d = {{dil4[0], dil4[1], dil4[2], dil4[3]}};
  const size_t warps = (size_t)N * P * Q;
  cuda_emul::run_grid("aspp_gather_kernel", (unsigned)((warps * 32 + 255) / 256), 256, 0, false, [&]() { aspp_gather_kernel(Z, b, d, out_nchw, N, P, Q, 768); });
  LAUNCHED();
  return 0;
}

extern "C" int sacb_aspp_gcol(const float* g_nchw, const int32_t* dil4, void* hi, void* lo, int N, int P, int Q,
                              void* stream) {
  Dil4 d = {{dil4[0], dil4[1], dil4[2], dil4[3]}};
  const size_t total = (size_t)N * P * Q * (768 / 8);
  cuda_emul::run_grid("aspp_gcol_kernel", grid1(total, 256), 256, 0, false, [&]() { aspp_gcol_kernel(g_nchw, d, (uint16_t*)hi, (uint16_t*)lo, N, P, Q, 768); });
  LAUNCHED();
  return 0;
}

extern "C" int sacb_aspp_unpack_wgrad(const float* parts, int splits, int C, float* dw0, float* dw1, float* dw2,
                                      float* dw3, void* stream) {
  MPtr4 dw = {{dw0, dw1, dw2, dw3}};
  cuda_emul::run_grid("aspp_unpack_wgrad_kernel", grid1((size_t)ASPP_J * C, 256), 256, 0, false, [&]() { aspp_unpack_wgrad_kernel(parts, splits, 768, C, dw); });
  LAUNCHED();
  return 0;
}
