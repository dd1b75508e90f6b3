// TEST INFRASTRUCTURE -- the checker's model of the two tensor-core entry points, NOT an emulation of the kernels.
//
// sacb_conv_gemm / sacb_conv_wgrad are tcgen05 + TMA kernels (csrc/sacb_gemm.cu) and cannot run on a host.  So that the
// host-side layer schedules (engine.py, engine_abn.py) can still be EXECUTED end to end in the GPU-less container -- with the
// streaming kernels running from their real source under cuda_emul.h -- this file evaluates the formulas stated in
// include/sacb.h for the two entry points with plain loops (fp32 products, x = hi + lo).  It proves nothing about the GPU
// kernels (tests/test_conv_gpu.py does that on a B200); it lets the schedules around them be checked against the goldens.
#include "cuda_emul.h"
#include "sacb.h"

namespace sacb { extern std::atomic<long long> g_launches; }
using sacb::bf16_bits_to_float;

namespace {

void to_float(const void* hi, const void* lo, size_t n, bool hi_only, std::vector<float>& out) {
  const uint16_t* h = (const uint16_t*)hi; const uint16_t* l = (const uint16_t*)lo;
  out.resize(n);
  for (size_t i = 0; i < n; ++i) out[i] = bf16_bits_to_float(h[i]) + (hi_only ? 0.f : bf16_bits_to_float(l[i]));
}

template <class F>
void parallel_for(long long n, long long chunk, F f) {       // f(begin, end)
  std::atomic<long long> next{0};
  cuda_emul::Pool::get().run([&](int) {
    for (;;) {
      const long long b = next.fetch_add(chunk);
      if (b >= n) break;
      f(b, b + chunk < n ? b + chunk : n);
    }
  });
}

inline float dot(const float* a, const float* b, int n) {
  float s = 0.f;
#pragma omp simd reduction(+ : s)
  for (int i = 0; i < n; ++i) s += a[i] * b[i];
  return s;
}

int geometry(int H, int W, int pad, int R, int S, int dil, int stride, int P, int Q, const char* who) {
  const int p = (H + 2 * pad - (R - 1) * dil - 1) / stride + 1, q = (W + 2 * pad - (S - 1) * dil - 1) / stride + 1;
  SACB_REQUIRE(p == P && q == Q, "%s: P,Q (%d,%d) inconsistent with geometry (%d,%d)", who, P, Q, p, q);
  return 0;
}

}  // namespace

extern "C" int sacb_conv_gemm(const SacbConvGemm* d, void*) {
  cuda_emul::ProfileScope prof("(model) sacb_conv_gemm");
  SACB_REQUIRE(d && d->size == sizeof(SacbConvGemm), "sacb_conv_gemm: bad descriptor size");
  SACB_REQUIRE(d->C % 64 == 0, "sacb_conv_gemm: C=%d must be a multiple of 64", d->C);
  SACB_REQUIRE(d->K % 32 == 0, "sacb_conv_gemm: K=%d must be a multiple of 32", d->K);
  SACB_REQUIRE(d->R == d->S, "sacb_conv_gemm: square filters only");
  if (int e = geometry(d->H, d->W, d->pad, d->R, d->S, d->dil, d->stride, d->P, d->Q, "sacb_conv_gemm")) return e;
  SACB_REQUIRE((d->scale == nullptr) == (d->shift == nullptr), "sacb_conv_gemm: scale and shift go together");
  SACB_REQUIRE((d->out_hi == nullptr) == (d->out_lo == nullptr), "sacb_conv_gemm: out_hi and out_lo go together");
  SACB_REQUIRE(d->precision == SACB_PRECISION_BF16X3 || d->precision == SACB_PRECISION_BF16, "sacb_conv_gemm: unknown precision %d", d->precision);
  SACB_REQUIRE(d->k_valid <= d->K, "sacb_conv_gemm: k_valid > K");
  const bool fast = d->precision == SACB_PRECISION_BF16;
  const int N = d->N, H = d->H, W = d->W, C = d->C, K = d->K, R = d->R, S = d->S, P = d->P, Q = d->Q;
  const long long M = (long long)N * P * Q;
  std::vector<float> xf, wf;
  to_float(d->x_hi, d->x_lo, (size_t)N * H * W * C, fast, xf);
  to_float(d->wt_hi, d->wt_lo, (size_t)R * S * K * C, fast, wf);
  const uint16_t* add_hi = (const uint16_t*)d->add_hi; const uint16_t* add_lo = (const uint16_t*)d->add_lo;
  const uint16_t* mask_hi = (const uint16_t*)d->mask_hi;
  uint16_t* out_hi = (uint16_t*)d->out_hi; uint16_t* out_lo = (uint16_t*)d->out_lo;
  std::mutex colsum_mutex;
  constexpr int MT = 8;
  parallel_for((M + MT - 1) / MT, 4, [&](long long t0, long long t1) {
    std::vector<float> acc((size_t)MT * K);
    std::vector<double> cs(d->colsum ? K : 0, 0.0);
    for (long long t = t0; t < t1; ++t) {
      const long long m0 = t * MT;
      const int rows = (int)(M - m0 < MT ? M - m0 : MT);
      std::fill(acc.begin(), acc.end(), 0.f);
      for (int r = 0; r < R; ++r)
        for (int s = 0; s < S; ++s) {
          const float* xrow[MT];
          bool any = false;
          for (int i = 0; i < rows; ++i) {
            const long long m = m0 + i;
            const int n = (int)(m / ((long long)P * Q)), p = (int)((m / Q) % P), q = (int)(m % Q);
            const int h = p * d->stride - d->pad + r * d->dil, w = q * d->stride - d->pad + s * d->dil;
            xrow[i] = (h >= 0 && h < H && w >= 0 && w < W) ? &xf[(((size_t)n * H + h) * W + w) * C] : nullptr;
            any |= xrow[i] != nullptr;
          }
          if (!any) continue;
          const float* wtap = &wf[(size_t)(r * S + s) * K * C];
          for (int k = 0; k < K; ++k) {
            const float* wrow = wtap + (size_t)k * C;
            for (int i = 0; i < rows; ++i)
              if (xrow[i]) acc[(size_t)i * K + k] += dot(xrow[i], wrow, C);
          }
        }
      for (int i = 0; i < rows; ++i) {
        const long long m = m0 + i;
        const size_t row = (size_t)m * K;
        for (int k = 0; k < K; ++k) {
          float v = acc[(size_t)i * K + k];
          if (d->scale) v = fmaf(v, d->scale[k], d->shift[k]);
          if (d->add_f32) v += d->add_f32[row + k];
          if (add_hi) v += bf16_bits_to_float(add_hi[row + k]) + bf16_bits_to_float(add_lo[row + k]);
          if (d->relu) v = fmaxf(v, 0.f);
          if (mask_hi) v = bf16_bits_to_float(mask_hi[row + k]) > 0.f ? v : 0.f;
          if (out_hi) {
            const uint16_t h = sacb::float_to_bf16_bits(v);
            out_hi[row + k] = h;
            out_lo[row + k] = sacb::float_to_bf16_bits(v - bf16_bits_to_float(h));
          }
          if (d->out_f32) d->out_f32[row + k] = v;
          if (d->out_nchw && k < d->k_valid) {
            const long long pq = (long long)P * Q, n = m / pq, rem = m - n * pq;
            d->out_nchw[((size_t)n * d->k_valid + k) * pq + rem] = v;
          }
          if (d->colsum) cs[k] += v;
        }
      }
    }
    if (d->colsum) {
      std::lock_guard<std::mutex> lk(colsum_mutex);
      for (int k = 0; k < K; ++k) d->colsum[k] += (float)cs[k];
    }
  });
  sacb::g_launches++;
  return 0;
}

static int plan_splits(const SacbConvWgrad* d, int& splits, long long& rows_per_split) {
  SACB_REQUIRE(d && d->size == sizeof(SacbConvWgrad), "sacb_conv_wgrad: bad descriptor size");
  SACB_REQUIRE(d->C % 64 == 0 && d->K % 64 == 0, "sacb_conv_wgrad: C=%d, K=%d must be multiples of 64", d->C, d->K);
  SACB_REQUIRE(d->R == d->S, "sacb_conv_wgrad: square filters only");
  if (int e = geometry(d->H, d->W, d->pad, d->R, d->S, d->dil, d->stride, d->P, d->Q, "sacb_conv_wgrad")) return e;
  SACB_REQUIRE(d->k_valid <= d->K, "sacb_conv_wgrad: k_valid > K");
  const long long M = (long long)d->N * d->P * d->Q;
  const int blocks = (int)((M + 63) / 64);
  splits = d->splits > 0 ? d->splits : 3;                 // auto: a fixed small number, so that the finalize kernels do sum planes
  if (splits > blocks) splits = blocks;
  const int per = (blocks + splits - 1) / splits;
  splits = (blocks + per - 1) / per;
  rows_per_split = 64LL * per;
  return 0;
}

extern "C" int sacb_conv_wgrad_splits(const SacbConvWgrad* d) {
  int splits; long long rps;
  if (int e = plan_splits(d, splits, rps)) return e;
  return splits;
}

extern "C" int sacb_conv_wgrad(const SacbConvWgrad* d, void*) {
  cuda_emul::ProfileScope prof("(model) sacb_conv_wgrad");
  int splits; long long rps;
  if (int e = plan_splits(d, splits, rps)) return e;
  SACB_REQUIRE(d->precision == SACB_PRECISION_BF16X3 || d->precision == SACB_PRECISION_BF16, "sacb_conv_wgrad: unknown precision %d", d->precision);
  const bool fast = d->precision == SACB_PRECISION_BF16;
  const int N = d->N, H = d->H, W = d->W, C = d->C, K = d->K, R = d->R, S = d->S, P = d->P, Q = d->Q, kv = d->k_valid;
  const long long M = (long long)N * P * Q;
  std::vector<float> xf, gf;
  to_float(d->x_hi, d->x_lo, (size_t)N * H * W * C, fast, xf);
  to_float(d->g_hi, d->g_lo, (size_t)M * K, fast, gf);
  const size_t plane = (size_t)kv * R * S * C;
  // one work item = (split, output row k): private accumulators, no races
  parallel_for((long long)splits * kv, 4, [&](long long i0, long long i1) {
    std::vector<float> acc((size_t)R * S * C);        // fp32 accumulation over one split's pixels, as the tensor core's TMEM
    for (long long it = i0; it < i1; ++it) {
      const int split = (int)(it / kv), k = (int)(it % kv);
      std::fill(acc.begin(), acc.end(), 0.f);
      const long long m0 = split * rps, m1 = m0 + rps < M ? m0 + rps : M;
      for (long long m = m0; m < m1; ++m) {
        const float g = gf[(size_t)m * K + k];
        if (g == 0.f) continue;
        const int n = (int)(m / ((long long)P * Q)), p = (int)((m / Q) % P), q = (int)(m % Q);
        for (int r = 0; r < R; ++r) {
          const int h = p * d->stride - d->pad + r * d->dil;
          if (h < 0 || h >= H) continue;
          for (int s = 0; s < S; ++s) {
            const int w = q * d->stride - d->pad + s * d->dil;
            if (w < 0 || w >= W) continue;
            const float* x = &xf[(((size_t)n * H + h) * W + w) * C];
            float* a = &acc[(size_t)(r * S + s) * C];
#pragma omp simd
            for (int c = 0; c < C; ++c) a[c] += g * x[c];
          }
        }
      }
      float* out = d->dw + (size_t)split * plane + (size_t)k * R * S * C;
      memcpy(out, acc.data(), sizeof(float) * R * S * C);
    }
  });
  sacb::g_launches++;
  return 0;
}
