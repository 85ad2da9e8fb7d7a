/* CPU ORACLE (test infrastructure, not product code): plain-C restatement of the integer/byte
 * arithmetic on the hot path.  Compiled by oracle/build_c.py into oracle/_build/liboracle_c.so and
 * called from tests/ and bench.py's cpu_baseline leg only.
 *
 * Follows (relative to /root/reference/ViDiT-Q):
 *   quant_rows_f32      quant_utils/qdiff/base/base_quantizer.py:110-157 (dynamic) / :58-99 (static)
 *   gemm_i8_i32         the int32 accumulators of quant_layer.py:70 in its algebraic integer form
 *                       (SURVEY appendix A); kernels/bench/bench_gemm.py:27-29 states the same algebra
 *   div_hoisted_check   the row-shared-reciprocal quotient used by the CUDA quantizer
 *                       (wan2.1-quantization_b200/csrc/common.cuh: div_rn_hoisted) against IEEE division
 *   attn_rowstep_i8     examples/Wan2.1/models/quant_opensora.py:430-478 in its integer form with the row-step
 *                       attention-map grid of base_quantizer.py:197-199 (the fused attention kernel's semantics)
 * Build with -ffp-contract=off: every '/' must stay an IEEE fp32 division, fmaf the only fused op.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

static float n_levels_of(int n_bits, int sym) { return sym ? (float)((1 << (n_bits - 1)) - 1) : (float)(1 << n_bits); }

/* codes (as float, like the reference), delta[rows], zero_point[rows] */
void quant_rows_f32(const float* x, long rows, long cols, int n_bits, int sym, int dynamic, float* codes, float* delta,
                    float* zero_point) {
  const float nl = n_levels_of(n_bits, sym);
  for (long r = 0; r < rows; ++r) {
    const float* xr = x + r * cols;
    float d, zp;
    if (sym) {
      float amax = 0.f;
      for (long c = 0; c < cols; ++c) { float a = fabsf(xr[c]); if (a > amax) amax = a; }
      d = amax / nl;
      if (dynamic && d < 1.0e-6f) d = 1.0e-6f;
      zp = 0.f;
    } else {
      float mx = 0.f, mn = 0.f;
      for (long c = 0; c < cols; ++c) { if (xr[c] > mx) mx = xr[c]; if (xr[c] < mn) mn = xr[c]; }
      d = (mx - mn) / (nl - 1.f);
      zp = rintf(mn / d) + nl / 2.f;
    }
    delta[r] = d; zero_point[r] = zp;
    for (long c = 0; c < cols; ++c) {
      float q = rintf(xr[c] / d) - zp;
      if (q < -nl - 1.f) q = -nl - 1.f;
      if (q > nl) q = nl;
      codes[r * cols + c] = q;
    }
  }
}

/* acc[m,n] = sum_k a[m,k]*w[n,k]  (int8 x int8 -> int32, exact) */
void gemm_i8_i32(const int8_t* a, const int8_t* w, long M, long N, long K, int32_t* acc) {
#pragma omp parallel for schedule(static)
  for (long m = 0; m < M; ++m)
    for (long n = 0; n < N; ++n) {
      int32_t s = 0;
      const int8_t* am = a + m * K; const int8_t* wn = w + n * K;
      for (long k = 0; k < K; ++k) s += (int32_t)am[k] * (int32_t)wn[k];
      acc[m * N + n] = s;
    }
}

static inline uint64_t xorshift(uint64_t* s) { *s ^= *s << 13; *s ^= *s >> 7; *s ^= *s << 17; return *s; }

/* Returns the number of (x, d) pairs, out of `trials` x 8, for which two Markstein corrections with
 * r = RN(1/d) do NOT reproduce the IEEE quotient x/d bit-for-bit.  Mix of uniform and near-tie x. */
long div_hoisted_check(long trials, uint64_t seed) {
  long bad = 0;
  uint64_t s = seed | 1;
  for (long it = 0; it < trials; ++it) {
    uint32_t m = (uint32_t)xorshift(&s) & 0x7fffffu;
    int e = (int)(xorshift(&s) % 30) - 20;
    uint32_t bits = ((uint32_t)(127 + e) << 23) | m;
    float d; memcpy(&d, &bits, 4);
    float r = 1.0f / d;                      /* IEEE division: correctly rounded reciprocal (== __frcp_rn) */
    for (int j = 0; j < 8; ++j) {
      float x;
      if (xorshift(&s) % 3 == 0) {
        double u = (double)(xorshift(&s) >> 11) / (double)(1ULL << 53);
        x = (float)((u * 2 - 1) * 127.0 * d);
      } else {
        int k = (int)(xorshift(&s) % 255) - 127;
        x = (float)((k + 0.5) * (double)d);
        uint32_t xb; memcpy(&xb, &x, 4); xb += (uint32_t)((int)(xorshift(&s) % 5) - 2); memcpy(&x, &xb, 4);
      }
      float q = x * r;
      float rem = fmaf(-q, d, x); q = fmaf(rem, r, q);
      rem = fmaf(-q, d, x); q = fmaf(rem, r, q);
      if (q != x / d) ++bad;
      if ((q + 12582912.0f) - 12582912.0f != rintf(x / d)) ++bad;   /* magic-number RNE == rintf */
    }
  }
  return bad;
}

/* Fused int8 attention in its integer form (what wan2.1-quantization_b200/csrc/attn.cu computes), one head:
 *   S = qq.kq^T exact in int32 ; x = S*dq[i]*dk[j]*scale ; P~ = exp(x - rowmax x) ; codes = rint(255*P~) in [0,255]
 *   (= DynamicQuantizer.forward_with_quant_params' unsigned grid, base_quantizer.py:197-199, delta = row maximum of the
 *   softmax: P = P~/l, rowmax P = 1/l) ; acc = codes . vq (exact integer) ; out = acc*dv[c] / (255*l), l = sum_j P~.
 * Follows quant_opensora.py:430-478 for the q/k/v operands (codes and scales are inputs here) and :456-462 for the
 * softmax; evaluated in double so that the codes are the correctly rounded ones the GPU test compares against.
 * qq [Lq,hd], kq [Lk,hd], vt [hd,Lk] int8 ; dq [Lq], dk [Lk], dv [hd] ; codes [Lq,Lk] ; acc [Lq,hd] ; out [Lq,hd]. */
void attn_rowstep_i8(const int8_t* qq, const int8_t* kq, const int8_t* vt, const float* dq, const float* dk, const float* dv,
                     long Lq, long Lk, long hd, double scale, uint8_t* codes, int64_t* acc, float* out, double* l_out) {
#pragma omp parallel for schedule(static)
  for (long i = 0; i < Lq; ++i) {
    double m = -INFINITY;
    double* x = (double*)__builtin_alloca(sizeof(double) * (size_t)Lk);
    for (long j = 0; j < Lk; ++j) {
      int32_t s = 0;
      for (long k = 0; k < hd; ++k) s += (int32_t)qq[i * hd + k] * (int32_t)kq[j * hd + k];
      x[j] = (double)s * (double)dq[i] * (double)dk[j] * scale;
      if (x[j] > m) m = x[j];
    }
    double l = 0.0;
    for (long j = 0; j < Lk; ++j) {
      const double p = exp(x[j] - m);
      l += p;
      codes[i * Lk + j] = (uint8_t)rint(255.0 * p);
    }
    for (long c = 0; c < hd; ++c) {
      int64_t a = 0;
      for (long j = 0; j < Lk; ++j) a += (int64_t)codes[i * Lk + j] * (int64_t)vt[c * Lk + j];
      acc[i * hd + c] = a;
      out[i * hd + c] = (float)((double)a * (double)dv[c] / (255.0 * l));
    }
    if (l_out) l_out[i] = l;
  }
}
