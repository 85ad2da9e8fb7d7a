// Token-local fused operators around the quantized linears.
//
//  b200q_ln_mod_quant : WanLayerNorm (fp32 statistics, wan/modules/model.py:92-102) -> optional affine
//                       (norm3) -> adaLN modulate `ln*(1+scale)+shift` (model.py:327,359) -> per-token
//                       symmetric quantizer (base_quantizer.py:110-157).  One HBM read of the fp32
//                       residual stream, one int8 write; any hidden size up to 32K channels (the reference
//                       kernel LayernormT2iQuantFuse, kernels/csrc/fused/fused.cu:234-380, needs
//                       hidden/4 threads and cannot launch for hidden > 4096).
//  b200q_gate_residual: out = residual + y*gate (model.py:337,362; reference fused.cu:382-483).
#include "common.cuh"
#include <type_traits>
#include <stdlib.h>

namespace b200q {

struct LnArgs {
  const void* x;
  int64_t rows, cols, ldx;
  const float* ln_w;
  const float* ln_b;
  float eps;
  const float* shift;
  const float* scale;
  float n_levels;
  int8_t* q;
  int64_t ldq;
  float* delta;
  int32_t* rowsum;
  void* y_out;
  int64_t ldy;
};

__device__ __forceinline__ uint32_t pack4i(int c0, int c1, int c2, int c3) {
  uint32_t t0 = __byte_perm((uint32_t)c0, (uint32_t)c1, 0x0040);
  uint32_t t1 = __byte_perm((uint32_t)c2, (uint32_t)c3, 0x0040);
  return __byte_perm(t0, t1, 0x5410);
}

// cross-warp (within one row's warps) reductions: one smem hop + one barrier; every call site owns its smem slot
__device__ __forceinline__ float row_reduce_sum(float v, float* s_slot, int warp, int lane, int w0, int wpr) {
  v = warp_sum(v);
  if (wpr > 1) {
    if (lane == 0) s_slot[warp] = v;
    __syncthreads();
    v = 0.f;
    for (int i = 0; i < wpr; ++i) v += s_slot[w0 + i];
  }
  return v;
}
__device__ __forceinline__ float row_reduce_max(float v, float* s_slot, int warp, int lane, int w0, int wpr) {
  v = warp_max(v);
  if (wpr > 1) {
    if (lane == 0) s_slot[warp] = v;
    __syncthreads();
    for (int i = 0; i < wpr; ++i) v = fmaxf(v, s_slot[w0 + i]);
  }
  return v;
}

// MODE bit 0: LayerNorm affine (ln_w, ln_b) ; bit 1: adaLN modulate (scale, shift).  Compile-time so the inner loops
// carry no per-element selects.  All element math is packed fp32x2 (FADD2 / FMUL2 / FFMA2).
// FULL: cols == V * warps_per_row * 32 vectors exactly, so every lane holds live data whenever its row exists and the
// per-vector predicates, selects and index clamps disappear (the Wan dims: 1536 = 12 x 32 float4, 5120 = 10 x 128).
template <typename T, typename YT, int V, int THREADS, int MODE, bool FULL>
__global__ void __launch_bounds__(THREADS) ln_mod_quant_kernel(const LnArgs a, const int warps_per_row) {
  using VT = Vec16<T>;
  constexpr int N = VT::N;
  constexpr int P = N / 2;
  constexpr bool AFFINE = (MODE & 1) != 0, MOD = (MODE & 2) != 0;
  __shared__ float s_buf[3][32];
  __shared__ int s_sum[32];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rows_per_cta = (THREADS / 32) / warps_per_row;
  const int row_in_cta = warp / warps_per_row;
  const int w0 = row_in_cta * warps_per_row;
  const int wr = warp - w0;
  const int row = (int)blockIdx.x * rows_per_cta + row_in_cta;
  const bool row_ok = row < (int)a.rows;
  const int tpr = warps_per_row * 32;
  const int t = wr * 32 + lane;
  const int kv = (int)(a.cols / N);
  const float inv_c = 1.f / (float)a.cols;

  const T* xrow = reinterpret_cast<const T*>(a.x) + (int64_t)(row_ok ? row : 0) * a.ldx + (int64_t)t * N;
  uint64_t x[V][P];
  bool live[V];
  uint64_t acc2 = pack_f32x2(0.f, 0.f);
#pragma unroll
  for (int v = 0; v < V; ++v) {
    live[v] = FULL ? row_ok : (row_ok && (v * tpr + t) < kv);
    const uint4 raw = live[v] ? ldg_stream16(xrow + (int64_t)v * tpr * N) : make_uint4(0, 0, 0, 0);
    VT::unpack_pairs(raw, x[v]);
#pragma unroll
    for (int i = 0; i < P; ++i) acc2 = add_f32x2(acc2, x[v][i]);
  }
  float s_lo, s_hi;
  unpack_f32x2(acc2, s_lo, s_hi);
  const float mean = row_reduce_sum(s_lo + s_hi, s_buf[0], warp, lane, w0, warps_per_row) * inv_c;
  const uint64_t nmean2 = pack_f32x2(-mean, -mean);
  uint64_t ss2 = pack_f32x2(0.f, 0.f);
#pragma unroll
  for (int v = 0; v < V; ++v) {
#pragma unroll
    for (int i = 0; i < P; ++i) {
      x[v][i] = (FULL || live[v]) ? add_f32x2(x[v][i], nmean2) : pack_f32x2(0.f, 0.f);   // d = x - mean (0 for padding lanes)
      ss2 = fma_f32x2(x[v][i], x[v][i], ss2);
    }
  }
  unpack_f32x2(ss2, s_lo, s_hi);
  const float var = row_reduce_sum(s_lo + s_hi, s_buf[1], warp, lane, w0, warps_per_row) * inv_c;
  const float rstd = __frsqrt_rn(var + a.eps);
  const uint64_t rstd2 = pack_f32x2(rstd, rstd), one2 = pack_f32x2(1.f, 1.f);

  float amax = 0.f;
#pragma unroll
  for (int v = 0; v < V; ++v) {
    const int c0 = (FULL || live[v]) ? (v * tpr + t) * N : 0;
#pragma unroll
    for (int h = 0; h < N / 4; ++h) {
      uint64_t y0 = mul_f32x2(x[v][2 * h], rstd2), y1 = mul_f32x2(x[v][2 * h + 1], rstd2);
      if (AFFINE) {
        const float4 w4 = __ldg(reinterpret_cast<const float4*>(a.ln_w + c0 + 4 * h));
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.ln_b + c0 + 4 * h));
        y0 = add_f32x2(mul_f32x2(y0, pack_f32x2(w4.x, w4.y)), pack_f32x2(b4.x, b4.y));
        y1 = add_f32x2(mul_f32x2(y1, pack_f32x2(w4.z, w4.w)), pack_f32x2(b4.z, b4.w));
      }
      if (MOD) {      // ln*(1+e1) + e0, two roundings like the reference's separate ops (model.py:327)
        const float4 s4 = __ldg(reinterpret_cast<const float4*>(a.scale + c0 + 4 * h));
        const float4 h4 = __ldg(reinterpret_cast<const float4*>(a.shift + c0 + 4 * h));
        y0 = add_f32x2(mul_f32x2(y0, add_f32x2(one2, pack_f32x2(s4.x, s4.y))), pack_f32x2(h4.x, h4.y));
        y1 = add_f32x2(mul_f32x2(y1, add_f32x2(one2, pack_f32x2(s4.z, s4.w))), pack_f32x2(h4.z, h4.w));
      }
      if (!FULL && !live[v]) { y0 = pack_f32x2(0.f, 0.f); y1 = y0; }
      x[v][2 * h] = y0; x[v][2 * h + 1] = y1;
      float f0, f1, f2, f3;
      unpack_f32x2(y0, f0, f1); unpack_f32x2(y1, f2, f3);
      amax = fmaxf(fmaxf(fabsf(f0), fabsf(f1)), amax);
      amax = fmaxf(fmaxf(fabsf(f2), fabsf(f3)), amax);
    }
    if (a.y_out != nullptr && live[v]) {
      YT* yrow = reinterpret_cast<YT*>(a.y_out) + (int64_t)row * a.ldy + c0;
      float f[N];
#pragma unroll
      for (int i = 0; i < P; ++i) unpack_f32x2(x[v][i], f[2 * i], f[2 * i + 1]);
      if constexpr (sizeof(YT) == 4) {
#pragma unroll
        for (int h = 0; h < N / 4; ++h)
          *reinterpret_cast<float4*>(yrow + 4 * h) = make_float4(f[4 * h], f[4 * h + 1], f[4 * h + 2], f[4 * h + 3]);
      } else {
        __nv_bfloat162 o[P];
#pragma unroll
        for (int i = 0; i < P; ++i) o[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
        if (N == 8) *reinterpret_cast<uint4*>(yrow) = *reinterpret_cast<uint4*>(o);
        else *reinterpret_cast<uint2*>(yrow) = *reinterpret_cast<uint2*>(o);
      }
    }
  }
  if (a.q == nullptr) return;       // uniform across the CTA

  amax = row_reduce_max(amax, s_buf[2], warp, lane, w0, warps_per_row);
  float delta = __fdiv_rn(amax, a.n_levels);
  if (delta < 1.0e-6f) delta = 1.0e-6f;                                   // base_quantizer.py:122-128
  const float r = __frcp_rn(delta);
  const uint64_t r2 = pack_f32x2(r, r), nd2 = pack_f32x2(-delta, -delta), magic2 = pack_f32x2(12582912.0f, 12582912.0f);
  int8_t* qrow = a.q + (int64_t)(row_ok ? row : 0) * a.ldq + (int64_t)t * N;
  int sum = 0;
#pragma unroll
  for (int v = 0; v < V; ++v) {
    uint32_t packed[N / 4];
#pragma unroll
    for (int g = 0; g < N / 4; ++g) {
      uint32_t c[4];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const uint64_t qb = div_rn_hoisted_rne2(x[v][2 * g + i], nd2, r2, magic2);
        unpack_u32x2(qb, c[2 * i], c[2 * i + 1]);
      }
      packed[g] = pack4i((int)c[0], (int)c[1], (int)c[2], (int)c[3]);
      sum = __dp4a((int)packed[g], 0x01010101, sum);
    }
    if (live[v]) {
      if (N == 4) stg_stream4(qrow + (int64_t)v * tpr * N, packed[0]);
      else stg_stream8(qrow + (int64_t)v * tpr * N, make_uint2(packed[0], packed[N / 4 - 1]));
    }
  }
  if (a.rowsum != nullptr) {
    sum = warp_sum(sum);
    if (warps_per_row > 1) {
      if (lane == 0) s_sum[warp] = sum;
      __syncthreads();
      if (t == 0) {
        sum = 0;
        for (int i = 0; i < warps_per_row; ++i) sum += s_sum[w0 + i];
      }
    }
    if (row_ok && t == 0) a.rowsum[row] = sum;
  }
  if (row_ok && t == 0) a.delta[row] = delta;
}

template <typename T, typename YT, int MODE>
static int launch_ln_mode(const LnArgs& a, cudaStream_t st) {
  constexpr int N = Vec16<T>::N;
  const int kv = (int)(a.cols / N);
  static int vmax_env = -1;          // tuning knob (B200Q_LN_VMAX): vectors per thread before a row is split over more warps
  if (vmax_env < 0) { const char* e = getenv("B200Q_LN_VMAX"); vmax_env = e ? atoi(e) : 0; }
  const int vmax = vmax_env > 0 ? (N == 4 ? vmax_env : (vmax_env + 1) / 2) : (N == 4 ? 12 : 6);
  const RowLayout lay = pick_row_layout(kv, vmax);                    // fp32 values live in registers: <= 48 per thread
  const int W = lay.W, V = lay.V;
  B200Q_REQUIRE(V <= 12 && (lay.threads == 256 || V <= 8), B200Q_ERR_UNSUPPORTED, "ln_mod_quant: cols=%lld too large",
                (long long)a.cols);
  const int rows_per_cta = (lay.threads / 32) / W;
  const unsigned grid = (unsigned)((a.rows + rows_per_cta - 1) / rows_per_cta);
  const bool full = (kv == V * W * 32);
#define B200Q_LN(VV, TH)                                                                              \
  do {                                                                                                \
    if (full && V == (VV)) ln_mod_quant_kernel<T, YT, VV, TH, MODE, true><<<grid, TH, 0, st>>>(a, W);  \
    else ln_mod_quant_kernel<T, YT, VV, TH, MODE, false><<<grid, TH, 0, st>>>(a, W);                   \
  } while (0)
  if (lay.threads == 1024) B200Q_LN(8, 1024);
  else if (V <= 2) B200Q_LN(2, 256);
  else if (V <= 4) B200Q_LN(4, 256);
  else if (V <= 8) B200Q_LN(8, 256);
  else if (V <= 10) B200Q_LN(10, 256);
  else B200Q_LN(12, 256);
#undef B200Q_LN
  B200Q_CHECK_LAUNCH();
  return B200Q_OK;
}

template <typename T, typename YT>
static int launch_ln(const LnArgs& a, cudaStream_t st) {
  const int mode = ((a.ln_w || a.ln_b) ? 1 : 0) | ((a.scale || a.shift) ? 2 : 0);
  switch (mode) {
    case 0: return launch_ln_mode<T, YT, 0>(a, st);
    case 1: return launch_ln_mode<T, YT, 1>(a, st);
    case 2: return launch_ln_mode<T, YT, 2>(a, st);
    default: return launch_ln_mode<T, YT, 3>(a, st);
  }
}

// ---- RMSNorm over the full model dim + 3-axis RoPE --------------------------------------------
// WanRMSNorm (wan/modules/model.py:73-89: fp32 statistics, result cast back to x's dtype, then * weight) followed by
// rope_apply (model.py:43-70; the reference multiplies complex128 numbers per sample in a Python loop): adjacent
// channel pairs (2i, 2i+1) of every head are rotated by the angle of the token's (frame, row, col) position, which the
// host passes as fp32 cos/sin tables [rows, head_dim/2] computed in float64.  One read of q (or k), one bf16 write.
struct RopeArgs {
  const void* x;
  int64_t rows, cols, ldx;
  const float* weight;
  float eps;
  const float* cos_t;
  const float* sin_t;
  int head_dim;
  void* out;
  int64_t ldo;
  // optional fused per-(token, head) symmetric quantizer of the rotated fp32 values (attention Q/K operands,
  // quant_opensora.py:430-435 = DynamicQuantizer on [tokens*heads, head_dim] rows): codes [rows, cols], scales [rows, heads]
  int8_t* q_out;
  int64_t ldq;
  float* dq_out;
  float n_levels;
  // optional: per 128-column head, the maximum over the rows of the squared norm of the bf16 OUTPUT (what the attention
  // kernel will read): head_sq_max[cols / 128], merged with atomicMax (the caller zeroes it).  b200q_attn_bf16_prenorm
  // classifies bounded heads from these instead of re-reading q and k in a pre-pass.
  float* head_sq_max;
};

template <typename T, int V, int THREADS>
__global__ void __launch_bounds__(THREADS) rmsnorm_rope_kernel(const RopeArgs a, const int warps_per_row) {
  using VT = Vec16<T>;
  constexpr int N = VT::N;            // 8 (bf16 / fp16)
  __shared__ float s_buf[32];
  __shared__ int s_hmax[64];          // per-head maxima of this CTA's rows (bit patterns of non-negative floats)
  if (a.head_sq_max != nullptr) {
    if (threadIdx.x < 64) s_hmax[threadIdx.x] = 0;
    __syncthreads();
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rows_per_cta = (THREADS / 32) / warps_per_row;
  const int row_in_cta = warp / warps_per_row;
  const int w0 = row_in_cta * warps_per_row;
  const int wr = warp - w0;
  const int64_t row = (int64_t)blockIdx.x * rows_per_cta + row_in_cta;
  const bool row_ok = row < a.rows;
  const int tpr = warps_per_row * 32;
  const int t = wr * 32 + lane;
  const int kv = (int)(a.cols / N);
  const T* xrow = reinterpret_cast<const T*>(a.x) + (row_ok ? row : 0) * a.ldx;
  float f[V][N];
  float ss = 0.f;
#pragma unroll
  for (int v = 0; v < V; ++v) {
    const int j = v * tpr + t;
    const uint4 raw = (row_ok && j < kv) ? ldg_stream16(xrow + (int64_t)j * N) : make_uint4(0, 0, 0, 0);
    VT::unpack(raw, f[v]);
#pragma unroll
    for (int i = 0; i < N; ++i) ss += f[v][i] * f[v][i];
  }
  const float ms = row_reduce_sum(ss, s_buf, warp, lane, w0, warps_per_row) / (float)a.cols;
  const float rstd = __frsqrt_rn(ms + a.eps);
  const int half = a.head_dim >> 1;
  const bool hd_pow2 = (a.head_dim & (a.head_dim - 1)) == 0;
  float hss[V];                                                  // head_sq_max only (dead code otherwise)
#pragma unroll
  for (int v = 0; v < V; ++v) hss[v] = 0.f;
#pragma unroll
  for (int v = 0; v < V; ++v) {
    const int j = v * tpr + t;
    if (!(row_ok && j < kv)) continue;
    const int c0 = j * N;
    float y[N];
#pragma unroll
    for (int h = 0; h < N / 4; ++h) {
      const float4 w4 = __ldg(reinterpret_cast<const float4*>(a.weight + c0 + 4 * h));
      const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
      for (int i = 0; i < 4; i += 2) {                       // .type_as(x) * weight: round to x's 16-bit type, back to fp32
        const float u0 = f[v][4 * h + i] * rstd, u1 = f[v][4 * h + i + 1] * rstd;
        float r0, r1;
        if constexpr (std::is_same<T, __nv_bfloat16>::value) {
          const __nv_bfloat162 pk = __floats2bfloat162_rn(u0, u1);          // one F2FP, then two shifts
          const uint32_t w = *reinterpret_cast<const uint32_t*>(&pk);
          r0 = __uint_as_float(w << 16); r1 = __uint_as_float(w & 0xffff0000u);
        } else {
          const float2 t2 = __half22float2(__floats2half2_rn(u0, u1));
          r0 = t2.x; r1 = t2.y;
        }
        y[4 * h + i] = r0 * wv[i]; y[4 * h + i + 1] = r1 * wv[i + 1];
      }
    }
    if (a.cos_t != nullptr) {
      // first pair index inside the head (multiple of 4); head_dim is a power of two for every Wan model - a runtime
      // integer modulo costs ~25 instructions per vector here
      const int p0 = (hd_pow2 ? (c0 & (a.head_dim - 1)) : (c0 % a.head_dim)) >> 1;
      const float4 c4 = __ldg(reinterpret_cast<const float4*>(a.cos_t + row * half + p0));
      const float4 s4 = __ldg(reinterpret_cast<const float4*>(a.sin_t + row * half + p0));
      const float cv[4] = {c4.x, c4.y, c4.z, c4.w}, sv[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float re = y[2 * i], im = y[2 * i + 1];
        y[2 * i] = re * cv[i] - im * sv[i];
        y[2 * i + 1] = re * sv[i] + im * cv[i];
      }
    }
    if (a.out != nullptr) {
      __nv_bfloat162 o[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) o[i] = __floats2bfloat162_rn(y[2 * i], y[2 * i + 1]);
      stg_stream16(reinterpret_cast<__nv_bfloat16*>(a.out) + row * a.ldo + c0, *reinterpret_cast<uint4*>(o));
      if (a.head_sq_max != nullptr) {                            // this thread's share of |row, head|^2; reduced after the loop
        float ss2 = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 f2 = __bfloat1622float2(o[i]);
          ss2 = fmaf(f2.x, f2.x, fmaf(f2.y, f2.y, ss2));
        }
        hss[v] = ss2;
      }
    }
    if (a.q_out != nullptr) {
      // a head (head_dim == 128 == 16 vectors) is held by 16 consecutive lanes: |y| max by 4 shuffles.  Every lane of
      // the group is live (cols is a multiple of head_dim), so the partial-warp shuffles below are well defined.
      float am = 0.f;
#pragma unroll
      for (int i = 0; i < N; ++i) am = fmaxf(am, fabsf(y[i]));
      const unsigned grp = 0xffffu << (lane & 16);
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) am = fmaxf(am, __shfl_xor_sync(grp, am, o));
      float d = am / a.n_levels;                              // base_quantizer.py:119
      if (d < 1e-6f) d = 1e-6f;                               // :122-128
      const float rcp = 1.0f / d;
      uint32_t w[2] = {0u, 0u};
#pragma unroll
      for (int i = 0; i < N; ++i) {
        int qv = rne_to_int(div_rn_hoisted(y[i], d, rcp));
        qv = max(-(int)a.n_levels - 1, min((int)a.n_levels, qv));
        w[i >> 2] |= ((uint32_t)qv & 0xffu) << (8 * (i & 3));
      }
      stg_stream8(a.q_out + row * a.ldq + c0, make_uint2(w[0], w[1]));
      if ((lane & 15) == 0) a.dq_out[row * (a.cols >> 7) + (c0 >> 7)] = d;      // head_dim == 128 on this path
    }
  }
  if (a.head_sq_max != nullptr) {
    // a 128-column head = 16 consecutive lanes of one vector slot (cols is a multiple of 128, so every lane of a live group
    // is live): the V slots' shuffle trees run side by side instead of one dependent chain per slot inside the loop
    if (row_ok) {
#pragma unroll
      for (int sh = 8; sh > 0; sh >>= 1) {
#pragma unroll
        for (int v = 0; v < V; ++v) hss[v] += __shfl_xor_sync(0xffffffffu, hss[v], sh);
      }
#pragma unroll
      for (int v = 0; v < V; ++v) {
        const int j = v * tpr + t;
        if (j < kv && (lane & 15) == 0) {
          float m = hss[v];
          if (m != m) m = INFINITY;                              // NaN row: the head is unbounded
          atomicMax(&s_hmax[(j * N) >> 7], __float_as_int(m));
        }
      }
    }
    __syncthreads();
    // thousands of CTAs, a dozen addresses: an atomic per CTA and head would serialise in L2 (measured: +12 us per launch).
    // The maximum only grows, so a CTA that does not exceed the value it reads (L2, uncached in L1) has nothing to add.
    if (threadIdx.x < (int)(a.cols >> 7)) {
      int* dst = reinterpret_cast<int*>(a.head_sq_max) + threadIdx.x;
      const int mine = s_hmax[threadIdx.x];
      if (mine > __ldcg(dst)) atomicMax(dst, mine);
    }
  }
}

template <typename T>
static int launch_rope(const RopeArgs& a, cudaStream_t st) {
  constexpr int N = Vec16<T>::N;
  const int kv = (int)(a.cols / N);
  const RowLayout lay = pick_row_layout(kv, 6);                     // 8 fp32 values per vector: 48 registers per thread
  const int W = lay.W, V = lay.V;
  B200Q_REQUIRE(V <= 8, B200Q_ERR_UNSUPPORTED, "rmsnorm_rope: cols=%lld too large", (long long)a.cols);
  const int rows_per_cta = (lay.threads / 32) / W;
  const unsigned grid = (unsigned)((a.rows + rows_per_cta - 1) / rows_per_cta);
  if (lay.threads == 1024) {
    if (V <= 4) rmsnorm_rope_kernel<T, 4, 1024><<<grid, 1024, 0, st>>>(a, W);
    else rmsnorm_rope_kernel<T, 8, 1024><<<grid, 1024, 0, st>>>(a, W);
  } else {
#define B200Q_RR(VV) rmsnorm_rope_kernel<T, VV, 256><<<grid, 256, 0, st>>>(a, W)
    B200Q_DISPATCH_V(V, B200Q_RR);
#undef B200Q_RR
  }
  B200Q_CHECK_LAUNCH();
  return B200Q_OK;
}

// ---- gate residual ---------------------------------------------------------------------------------------
template <typename YT>
__global__ void __launch_bounds__(256) gate_residual_kernel(const YT* __restrict__ y, int64_t ldy,
                                                             const float* __restrict__ gate,
                                                             const float* residual, int64_t ldr, float* out,
                                                             int64_t ldo, int64_t rows, int64_t cols4) {
  const int64_t total = rows * cols4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cols4, c = (i - r * cols4) * 4;
    float yv[4];
    if constexpr (sizeof(YT) == 4) {
      const float4 t = *reinterpret_cast<const float4*>(y + r * ldy + c);
      yv[0] = t.x; yv[1] = t.y; yv[2] = t.z; yv[3] = t.w;
    } else {
      const uint2 t = *reinterpret_cast<const uint2*>(y + r * ldy + c);
      const YT* h = reinterpret_cast<const YT*>(&t);
#pragma unroll
      for (int k = 0; k < 4; ++k) yv[k] = to_f32(h[k]);
    }
    const float4 res = *reinterpret_cast<const float4*>(residual + r * ldr + c);
    float4 g = make_float4(1.f, 1.f, 1.f, 1.f);
    if (gate) g = *reinterpret_cast<const float4*>(gate + c);
    float4 o;
    // x + y*e : two roundings like the reference's separate mul and add (model.py:337)
    o.x = __fadd_rn(res.x, __fmul_rn(yv[0], g.x));
    o.y = __fadd_rn(res.y, __fmul_rn(yv[1], g.y));
    o.z = __fadd_rn(res.z, __fmul_rn(yv[2], g.z));
    o.w = __fadd_rn(res.w, __fmul_rn(yv[3], g.w));
    *reinterpret_cast<float4*>(out + r * ldo + c) = o;
  }
}

template <typename YT>
__global__ void __launch_bounds__(256) gate_residual_scalar_kernel(const YT* __restrict__ y, int64_t ldy,
                                                                    const float* __restrict__ gate,
                                                                    const float* residual, int64_t ldr, float* out,
                                                                    int64_t ldo, int64_t rows, int64_t cols) {
  const int64_t total = rows * cols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cols, c = i - r * cols;
    const float g = gate ? gate[c] : 1.f;
    out[r * ldo + c] = __fadd_rn(residual[r * ldr + c], __fmul_rn(to_f32(y[r * ldy + c]), g));
  }
}

template <typename YT>
static int launch_gate(const void* y, int64_t ldy, const float* gate, const float* residual, int64_t ldr, float* out,
                       int64_t ldo, int64_t rows, int64_t cols, cudaStream_t st) {
  const YT* yp = reinterpret_cast<const YT*>(y);
  const bool vec = cols % 4 == 0 && ldy % 4 == 0 && ldr % 4 == 0 && ldo % 4 == 0 && aligned(y, 4 * sizeof(YT)) &&
                   aligned(residual, 16) && aligned(out, 16) && (!gate || aligned(gate, 16));
  const int64_t work = vec ? rows * (cols / 4) : rows * cols;
  int64_t blocks = (work + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 32;
  if (blocks > cap) blocks = cap;
  if (vec) gate_residual_kernel<YT><<<(unsigned)blocks, 256, 0, st>>>(yp, ldy, gate, residual, ldr, out, ldo, rows, cols / 4);
  else gate_residual_scalar_kernel<YT><<<(unsigned)blocks, 256, 0, st>>>(yp, ldy, gate, residual, ldr, out, ldo, rows, cols);
  B200Q_CHECK_LAUNCH();
  return B200Q_OK;
}

}  // namespace b200q

using namespace b200q;

extern "C" int b200q_ln_mod_quant(const void* x, int x_dtype, int64_t rows, int64_t cols, int64_t ldx,
                                  const float* ln_w, const float* ln_b, float eps, const float* shift,
                                  const float* scale, int n_bits, int8_t* q, int64_t ldq, float* delta,
                                  int32_t* rowsum, void* y_out, int y_dtype, int64_t ldy, b200q_stream_t stream) {
  clear_error();
  B200Q_REQUIRE(rows >= 0 && cols >= 0, B200Q_ERR_BAD_ARG, "ln_mod_quant: negative shape");
  if (rows == 0 || cols == 0) return B200Q_OK;
  B200Q_REQUIRE(x != nullptr && (q != nullptr || y_out != nullptr), B200Q_ERR_BAD_ARG, "ln_mod_quant: null pointer");
  B200Q_REQUIRE(q == nullptr || delta != nullptr, B200Q_ERR_BAD_ARG, "ln_mod_quant: delta required with q");
  B200Q_REQUIRE(n_bits >= 2 && n_bits <= 8, B200Q_ERR_BAD_ARG, "ln_mod_quant: n_bits=%d out of [2,8]", n_bits);
  B200Q_REQUIRE(ldx >= cols && (q == nullptr || ldq >= cols) && (y_out == nullptr || ldy >= cols), B200Q_ERR_BAD_ARG,
                "ln_mod_quant: leading dimension < cols");
  const int vecn = x_dtype == B200Q_F32 ? 4 : 8;
  B200Q_REQUIRE(cols % vecn == 0 && ldx % vecn == 0 && aligned(x, 16), B200Q_ERR_UNSUPPORTED,
                "ln_mod_quant: cols and ldx must be multiples of %d and x 16-byte aligned", vecn);
  B200Q_REQUIRE(q == nullptr || (ldq % vecn == 0 && aligned(q, vecn)), B200Q_ERR_UNSUPPORTED, "ln_mod_quant: q misaligned");
  B200Q_REQUIRE(y_out == nullptr || (ldy % vecn == 0 && aligned(y_out, 16)), B200Q_ERR_UNSUPPORTED, "ln_mod_quant: y_out misaligned");
  B200Q_REQUIRE((!ln_w || aligned(ln_w, 16)) && (!ln_b || aligned(ln_b, 16)) && (!shift || aligned(shift, 16)) &&
                    (!scale || aligned(scale, 16)),
                B200Q_ERR_BAD_ARG, "ln_mod_quant: per-channel vectors must be 16-byte aligned");
  B200Q_REQUIRE((ln_w == nullptr) == (ln_b == nullptr) && (scale == nullptr) == (shift == nullptr), B200Q_ERR_BAD_ARG,
                "ln_mod_quant: ln_w/ln_b and scale/shift are given in pairs");
  B200Q_REQUIRE(rows <= 0x7fffffff, B200Q_ERR_UNSUPPORTED, "ln_mod_quant: rows > 2^31-1");
  LnArgs a{};
  a.x = x; a.rows = rows; a.cols = cols; a.ldx = ldx; a.ln_w = ln_w; a.ln_b = ln_b; a.eps = eps;
  a.shift = shift; a.scale = scale; a.n_levels = (float)((1 << (n_bits - 1)) - 1);
  a.q = q; a.ldq = ldq; a.delta = delta; a.rowsum = rowsum; a.y_out = y_out; a.ldy = ldy;
  cudaStream_t st = (cudaStream_t)stream;
#define B200Q_LN_Y(T)                                                                  \
  switch (y_out ? y_dtype : B200Q_F32) {                                               \
    case B200Q_F32: return launch_ln<T, float>(a, st);                                 \
    case B200Q_BF16: return launch_ln<T, __nv_bfloat16>(a, st);                        \
    default: set_error("ln_mod_quant: y_dtype must be f32 or bf16"); return B200Q_ERR_BAD_ARG; \
  }
  switch (x_dtype) {
    case B200Q_F32: B200Q_LN_Y(float)
    case B200Q_BF16: B200Q_LN_Y(__nv_bfloat16)
  }
#undef B200Q_LN_Y
  set_error("ln_mod_quant: x_dtype must be f32 or bf16 (got %d)", x_dtype);
  return B200Q_ERR_BAD_ARG;
}

extern "C" int b200q_gate_residual(const void* y, int y_dtype, int64_t ldy, const float* gate, const float* residual,
                                   int64_t ldr, float* out, int64_t ldo, int64_t rows, int64_t cols,
                                   b200q_stream_t stream) {
  clear_error();
  B200Q_REQUIRE(rows >= 0 && cols >= 0, B200Q_ERR_BAD_ARG, "gate_residual: negative shape");
  if (rows == 0 || cols == 0) return B200Q_OK;
  B200Q_REQUIRE(y && residual && out, B200Q_ERR_BAD_ARG, "gate_residual: null pointer");
  B200Q_REQUIRE(ldy >= cols && ldr >= cols && ldo >= cols, B200Q_ERR_BAD_ARG, "gate_residual: leading dimension < cols");
  cudaStream_t st = (cudaStream_t)stream;
  switch (y_dtype) {
    case B200Q_F32: return launch_gate<float>(y, ldy, gate, residual, ldr, out, ldo, rows, cols, st);
    case B200Q_BF16: return launch_gate<__nv_bfloat16>(y, ldy, gate, residual, ldr, out, ldo, rows, cols, st);
    case B200Q_F16: return launch_gate<__half>(y, ldy, gate, residual, ldr, out, ldo, rows, cols, st);
  }
  set_error("gate_residual: bad y_dtype %d", y_dtype);
  return B200Q_ERR_BAD_ARG;
}

extern "C" int b200q_rmsnorm_rope(const void* x, int x_dtype, int64_t rows, int64_t cols, int64_t ldx, const float* weight,
                                  float eps, const float* cos_t, const float* sin_t, int head_dim, void* out, int64_t ldo,
                                  b200q_stream_t stream) {
  B200Q_REQUIRE(out != nullptr || rows == 0 || cols == 0, B200Q_ERR_BAD_ARG, "rmsnorm_rope: null pointer");
  return b200q_rmsnorm_rope_quant(x, x_dtype, rows, cols, ldx, weight, eps, cos_t, sin_t, head_dim, out, ldo, nullptr, 0,
                                  nullptr, 8, stream);
}

static int rmsnorm_rope_impl(const void* x, int x_dtype, int64_t rows, int64_t cols, int64_t ldx, const float* weight, float eps,
                             const float* cos_t, const float* sin_t, int head_dim, void* out, int64_t ldo, int8_t* q_out,
                             int64_t ldq, float* dq_out, int n_bits, float* head_sq_max, b200q_stream_t stream);

extern "C" int b200q_rmsnorm_rope_quant(const void* x, int x_dtype, int64_t rows, int64_t cols, int64_t ldx,
                                        const float* weight, float eps, const float* cos_t, const float* sin_t,
                                        int head_dim, void* out, int64_t ldo, int8_t* q_out, int64_t ldq, float* dq_out,
                                        int n_bits, b200q_stream_t stream) {
  return rmsnorm_rope_impl(x, x_dtype, rows, cols, ldx, weight, eps, cos_t, sin_t, head_dim, out, ldo, q_out, ldq, dq_out, n_bits,
                           nullptr, stream);
}

extern "C" int b200q_rmsnorm_rope_stats(const void* x, int x_dtype, int64_t rows, int64_t cols, int64_t ldx, const float* weight,
                                        float eps, const float* cos_t, const float* sin_t, int head_dim, void* out, int64_t ldo,
                                        float* head_sq_max, b200q_stream_t stream) {
  B200Q_REQUIRE(out != nullptr || rows == 0 || cols == 0, B200Q_ERR_BAD_ARG, "rmsnorm_rope_stats: null pointer");
  B200Q_REQUIRE(head_sq_max != nullptr && cols % 128 == 0 && cols <= 64 * 128, B200Q_ERR_BAD_ARG,
                "rmsnorm_rope_stats: head_sq_max [cols / 128] required, cols a multiple of 128 (at most 64 heads)");
  return rmsnorm_rope_impl(x, x_dtype, rows, cols, ldx, weight, eps, cos_t, sin_t, head_dim, out, ldo, nullptr, 0, nullptr, 8,
                           head_sq_max, stream);
}

static int rmsnorm_rope_impl(const void* x, int x_dtype, int64_t rows, int64_t cols, int64_t ldx, const float* weight, float eps,
                             const float* cos_t, const float* sin_t, int head_dim, void* out, int64_t ldo, int8_t* q_out,
                             int64_t ldq, float* dq_out, int n_bits, float* head_sq_max, b200q_stream_t stream) {
  clear_error();
  B200Q_REQUIRE(rows >= 0 && cols >= 0, B200Q_ERR_BAD_ARG, "rmsnorm_rope: negative shape");
  if (rows == 0 || cols == 0) return B200Q_OK;
  B200Q_REQUIRE(x && weight && (out || q_out), B200Q_ERR_BAD_ARG, "rmsnorm_rope: null pointer");
  if (q_out) {
    B200Q_REQUIRE(dq_out != nullptr, B200Q_ERR_BAD_ARG, "rmsnorm_rope_quant: dq_out required with q_out");
    B200Q_REQUIRE(head_dim == 128 && cols % 128 == 0, B200Q_ERR_UNSUPPORTED,
                  "rmsnorm_rope_quant: the fused per-(token, head) quantizer needs head_dim == 128");
    B200Q_REQUIRE(ldq >= cols && ldq % 8 == 0 && aligned(q_out, 8), B200Q_ERR_BAD_ARG, "rmsnorm_rope_quant: bad q_out layout");
    B200Q_REQUIRE(n_bits >= 2 && n_bits <= 8, B200Q_ERR_BAD_ARG, "rmsnorm_rope_quant: n_bits out of [2,8]");
  }
  if (out == nullptr) ldo = cols;
  B200Q_REQUIRE(x_dtype == B200Q_BF16 || x_dtype == B200Q_F16, B200Q_ERR_BAD_ARG, "rmsnorm_rope: x must be bf16 or fp16");
  B200Q_REQUIRE(cols % 8 == 0 && ldx % 8 == 0 && ldo % 8 == 0 && ldx >= cols && ldo >= cols && aligned(x, 16) &&
                    (out == nullptr || aligned(out, 16)) && aligned(weight, 16),
                B200Q_ERR_UNSUPPORTED, "rmsnorm_rope: cols/ldx/ldo must be multiples of 8 and pointers 16-byte aligned");
  B200Q_REQUIRE((cos_t == nullptr) == (sin_t == nullptr), B200Q_ERR_BAD_ARG, "rmsnorm_rope: cos and sin go together");
  if (cos_t) {
    B200Q_REQUIRE(head_dim > 0 && head_dim % 8 == 0 && cols % head_dim == 0 && aligned(cos_t, 16) && aligned(sin_t, 16),
                  B200Q_ERR_BAD_ARG, "rmsnorm_rope: head_dim must be a multiple of 8 dividing cols; tables 16-byte aligned");
  }
  RopeArgs a{};
  a.x = x; a.rows = rows; a.cols = cols; a.ldx = ldx; a.weight = weight; a.eps = eps; a.cos_t = cos_t; a.sin_t = sin_t;
  a.head_dim = head_dim > 0 ? head_dim : (int)cols; a.out = out; a.ldo = ldo;
  a.q_out = q_out; a.ldq = ldq; a.dq_out = dq_out; a.n_levels = (float)((1 << (n_bits - 1)) - 1);
  a.head_sq_max = head_sq_max;
  if (x_dtype == B200Q_BF16) return launch_rope<__nv_bfloat16>(a, (cudaStream_t)stream);
  return launch_rope<__half>(a, (cudaStream_t)stream);
}
