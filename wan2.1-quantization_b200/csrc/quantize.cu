// (a) Per-row quantizer: one HBM read of x, one write of int8 codes (+ delta/zp/rowsum).
//
// Replaces the ~12 torch elementwise/reduction kernels + 2 host syncs of
// DynamicQuantizer.quantize (ViDiT-Q/quant_utils/qdiff/base/base_quantizer.py:110-157), the
// offline StaticQuantizer (base_quantizer.py:58-99) and the reference CUDA kernel
// QuantKernel / fused.quant_sum (ViDiT-Q/kernels/csrc/fused/fused.cu:30-131) — whose
// x*(127/amax)+fast-math arithmetic is NOT what the fake-quant path computes; here the
// quotient is the correctly rounded fp32 x/delta followed by round-half-even.
//
// Layout: a row is owned by W warps (W*32 threads); each thread keeps V 16-byte vectors of the
// row in registers (single HBM read), abs-max / min-max go through warp shuffles (+ one smem
// hop when W > 1), then the same registers are quantized and streamed out.  HBM-bound:
// algorithmic bytes = rows*cols*(sizeof(in)+1) + 12*rows.
#include "common.cuh"

namespace b200q {

enum QuantMode { kSymDyn = 0, kAsymDyn = 1, kStatic = 2 };

struct QuantArgs {
  const void* x;
  int64_t rows, cols, ldx;
  int8_t* q;
  int64_t ldq;
  float* delta;            // out (dyn) / in (static)
  float* zero_point;       // out (dyn) / in (static)
  int32_t* rowsum;         // optional
  float* stat_max;         // optional: sym absmax | asym max(rowmax,0)
  float* stat_min;         // optional: asym min(rowmin,0)
  float n_levels;          // sym: 2^(b-1)-1 ; asym: 2^b
  float clamp_lo, clamp_hi;
  int apply_eps_floor;     // dynamic quantizers only
};

__device__ __forceinline__ uint32_t pack4(int c0, int c1, int c2, int c3) {
  uint32_t t0 = __byte_perm((uint32_t)c0, (uint32_t)c1, 0x0040);
  uint32_t t1 = __byte_perm((uint32_t)c2, (uint32_t)c3, 0x0040);
  return __byte_perm(t0, t1, 0x5410);
}

// Per-row parameters from the reduced statistics (thread-uniform within a row).
template <int MODE>
__device__ __forceinline__ void row_params(float s0, float s1, const QuantArgs& a, float& delta, float& zp) {
  if (MODE == kSymDyn) {                       // s0 = absmax
    delta = __fdiv_rn(s0, a.n_levels);
    if (a.apply_eps_floor && delta < 1.0e-6f) delta = 1.0e-6f;   // base_quantizer.py:122-128
    if (!a.apply_eps_floor && !(delta > 0.f)) delta = 1.0e-8f;   // all-zero weight row: reference drops into ipdb
    zp = 0.f;
  } else {                                     // s0 = max(rowmax,0), s1 = min(rowmin,0)
    delta = __fdiv_rn(s0 - s1, a.n_levels - 1.f);
    if (!(delta > 1.0e-8f)) delta = 1.0e-8f;                     // base_quantizer.py:139-147 (post-ipdb branch)
    zp = rintf(__fdiv_rn(s1, delta)) + 0.5f * a.n_levels;
  }
}

template <typename T, int V, int MODE, int THREADS>
__global__ void __launch_bounds__(THREADS) quant_rows_kernel(const QuantArgs a, const int warps_per_row) {
  using VT = Vec16<T>;
  constexpr int N = VT::N;
  __shared__ float s_red[2][32];
  __shared__ int s_sum[32];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rows_per_cta = (THREADS / 32) / warps_per_row;
  const int row_in_cta = warp / warps_per_row;
  const int w0 = row_in_cta * warps_per_row;
  const int wr = warp - w0;
  const int row = (int)blockIdx.x * rows_per_cta + row_in_cta;          // rows < 2^31 (checked on the host)
  const bool row_ok = row < (int)a.rows;
  const int tpr = warps_per_row * 32;
  const int t = wr * 32 + lane;
  const int kv = (int)(a.cols / N);

  const T* xrow = reinterpret_cast<const T*>(a.x) + (int64_t)(row_ok ? row : 0) * a.ldx + (int64_t)t * N;
  uint4 raw[V];
  bool live[V];
#pragma unroll
  for (int v = 0; v < V; ++v) {
    live[v] = row_ok && (v * tpr + t) < kv;
    raw[v] = live[v] ? ldg_stream16(xrow + (int64_t)v * tpr * N) : make_uint4(0, 0, 0, 0);
  }

  float delta, zp;
  float s0 = 0.f, s1 = 0.f;     // sym: s0 = absmax ; asym: s0 = max(.,0), s1 = min(.,0)
  if (MODE == kStatic) {
    delta = row_ok ? a.delta[row] : 1.f;
    zp = row_ok ? a.zero_point[row] : 0.f;
  } else {
    typename VT::Stat st = VT::stat_init();
#pragma unroll
    for (int v = 0; v < V; ++v) VT::template stat_update<MODE == kSymDyn>(st, raw[v]);
    VT::stat_final(st, s0, s1);
    s0 = warp_max(s0);
    if (MODE == kAsymDyn) s1 = warp_min(s1);
    if (warps_per_row > 1) {                   // one smem hop; every warp folds the W partials itself (W <= 32 broadcasts)
      if (lane == 0) { s_red[0][warp] = s0; s_red[1][warp] = s1; }
      __syncthreads();
      for (int i = 0; i < warps_per_row; ++i) {
        s0 = fmaxf(s0, s_red[0][w0 + i]);
        if (MODE == kAsymDyn) s1 = fminf(s1, s_red[1][w0 + i]);
      }
    }
    row_params<MODE>(s0, s1, a, delta, zp);
  }

  const float r = __frcp_rn(delta);
  const uint64_t r2 = pack_f32x2(r, r), nd2 = pack_f32x2(-delta, -delta), magic2 = pack_f32x2(12582912.0f, 12582912.0f);
  const int zpi = __float2int_rn(zp) + 0x4B400000;      // also strips the magic-number exponent bits
  const int lo = (int)a.clamp_lo, hi = (int)a.clamp_hi;
  int8_t* qrow = a.q + (int64_t)(row_ok ? row : 0) * a.ldq + (int64_t)t * N;
  int sum = 0;
#pragma unroll
  for (int v = 0; v < V; ++v) {
    uint64_t xp[N / 2];
    VT::unpack_pairs(raw[v], xp);
    uint32_t packed[N / 4];
#pragma unroll
    for (int g = 0; g < N / 4; ++g) {
      uint32_t c[4];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        // exact RN(x/delta) on two lanes (FFMA2), then round-half-even via the magic-number add
        const uint64_t qb = div_rn_hoisted_rne2(xp[2 * g + i], nd2, r2, magic2);
        unpack_u32x2(qb, c[2 * i], c[2 * i + 1]);
      }
      if (MODE != kSymDyn) {                    // sym dynamic: |x/delta| <= n_levels, the low byte is the code
#pragma unroll
        for (int i = 0; i < 4; ++i) c[i] = (uint32_t)min(max((int)c[i] - zpi, lo), hi);
      }
      packed[g] = pack4((int)c[0], (int)c[1], (int)c[2], (int)c[3]);
      // dead vectors hold x = 0 -> code 0 for symmetric quantizers; with a zero point they must be masked
      if (MODE == kSymDyn || live[v]) sum = __dp4a((int)packed[g], 0x01010101, sum);
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
  if (MODE != kStatic && row_ok && t == 0) {
    a.delta[row] = delta;
    a.zero_point[row] = zp;
    if (a.stat_max) a.stat_max[row] = s0;
    if (a.stat_min) a.stat_min[row] = s1;
  }
}

// Generic path: any cols / alignment / row length. One CTA per row, two passes (second pass
// re-reads the row, normally from L2).
template <typename T, int MODE>
__global__ void __launch_bounds__(256) quant_rows_generic_kernel(const QuantArgs a) {
  __shared__ float s_red[2][8];
  __shared__ int s_sum[8];
  const int64_t row = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const T* xrow = reinterpret_cast<const T*>(a.x) + row * a.ldx;
  float delta, zp;
  float s0 = 0.f, s1 = 0.f;
  if (MODE == kStatic) {
    delta = a.delta[row]; zp = a.zero_point[row];
  } else {
    for (int64_t c = threadIdx.x; c < a.cols; c += blockDim.x) {
      const float f = to_f32(xrow[c]);
      if (MODE == kSymDyn) s0 = fmaxf(s0, fabsf(f));
      else { s0 = fmaxf(s0, f); s1 = fminf(s1, f); }
    }
    s0 = warp_max(s0); s1 = warp_min(s1);
    if (lane == 0) { s_red[0][warp] = s0; s_red[1][warp] = s1; }
    __syncthreads();
    s0 = warp_max(lane < 8 ? s_red[0][lane] : 0.f);
    s1 = warp_min(lane < 8 ? s_red[1][lane] : 0.f);
    row_params<MODE>(s0, s1, a, delta, zp);
  }
  const float r = __frcp_rn(delta);
  const int zpi = __float2int_rn(zp);
  const int lo = (int)a.clamp_lo, hi = (int)a.clamp_hi;
  int8_t* qrow = a.q + row * a.ldq;
  int sum = 0;
  for (int64_t c = threadIdx.x; c < a.cols; c += blockDim.x) {
    const float qf = div_rn_hoisted(to_f32(xrow[c]), delta, r);
    const int code = min(max(rne_to_int(qf) - zpi, lo), hi);
    qrow[c] = (int8_t)code;
    sum += code;
  }
  if (a.rowsum != nullptr) {
    sum = warp_sum(sum);
    if (lane == 0) s_sum[warp] = sum;
    __syncthreads();
    sum = warp_sum(lane < 8 ? s_sum[lane] : 0);
    if (threadIdx.x == 0) a.rowsum[row] = sum;
  }
  if (MODE != kStatic && threadIdx.x == 0) {
    a.delta[row] = delta; a.zero_point[row] = zp;
    if (a.stat_max) a.stat_max[row] = s0;
    if (a.stat_min) a.stat_min[row] = s1;
  }
}

template <typename T, int MODE>
static int launch_quant(const QuantArgs& a, cudaStream_t st) {
  constexpr int N = Vec16<T>::N;
  const bool fast = (a.cols % N == 0) && (a.ldx % N == 0) && aligned(a.x, 16) &&
                    (a.ldq % (N == 4 ? 4 : 8) == 0) && aligned(a.q, N == 4 ? 4 : 8) &&
                    (a.cols / N <= 8 * 32 * 32);   // <= 8 vectors per thread of a 1024-thread CTA
  if (!fast) {
    quant_rows_generic_kernel<T, MODE><<<(unsigned)a.rows, 256, 0, st>>>(a);
    B200Q_CHECK_LAUNCH();
    return B200Q_OK;
  }
  const int kv = (int)(a.cols / N);
  const RowLayout lay = pick_row_layout(kv, N == 4 ? 12 : 10);       // 48 / 40 data registers per thread
  const int W = lay.W, V = lay.V;
  const int rows_per_cta = (lay.threads / 32) / W;
  const unsigned grid = (unsigned)((a.rows + rows_per_cta - 1) / rows_per_cta);
  if (lay.threads == 1024) {      // very long rows: one 1024-thread CTA per row
    if (V <= 4) quant_rows_kernel<T, 4, MODE, 1024><<<grid, 1024, 0, st>>>(a, W);
    else if (V <= 8) quant_rows_kernel<T, 8, MODE, 1024><<<grid, 1024, 0, st>>>(a, W);
    else { set_error("quant_rows: internal layout error"); return B200Q_ERR_UNSUPPORTED; }
  } else {
#define B200Q_LAUNCH_V(VV) quant_rows_kernel<T, VV, MODE, 256><<<grid, 256, 0, st>>>(a, W)
    B200Q_DISPATCH_V(V, B200Q_LAUNCH_V);
#undef B200Q_LAUNCH_V
  }
  B200Q_CHECK_LAUNCH();
  return B200Q_OK;
}

template <int MODE>
static int dispatch_dtype(int dtype, const QuantArgs& a, cudaStream_t st) {
  switch (dtype) {
    case B200Q_F32: return launch_quant<float, MODE>(a, st);
    case B200Q_BF16: return launch_quant<__nv_bfloat16, MODE>(a, st);
    case B200Q_F16: return launch_quant<__half, MODE>(a, st);
  }
  set_error("quant_rows: unsupported x_dtype %d", dtype);
  return B200Q_ERR_BAD_ARG;
}

static int fill_levels(QuantArgs& a, int n_bits, int sym) {
  B200Q_REQUIRE(n_bits >= 2 && n_bits <= 8, B200Q_ERR_BAD_ARG, "n_bits=%d out of [2,8]", n_bits);
  a.n_levels = sym ? (float)((1 << (n_bits - 1)) - 1) : (float)(1 << n_bits);   // base_quantizer.py:32
  // reference clamp is [-n_levels-1, n_levels] (base_quantizer.py:67,156); int8 storage bounds it to [-128,127]
  a.clamp_lo = fmaxf(-a.n_levels - 1.f, -128.f);
  a.clamp_hi = fminf(a.n_levels, 127.f);
  return B200Q_OK;
}

// ---- dequant: out = (q + zp) * delta ----------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) dequant_rows_kernel(const int8_t* __restrict__ q, int64_t ldq, int64_t rows,
                                                            int64_t cols, const float* __restrict__ delta,
                                                            const float* __restrict__ zp, T* __restrict__ out,
                                                            int64_t ldo) {
  const int64_t total = rows * cols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cols, c = i - r * cols;
    const float z = zp ? zp[r] : 0.f;
    out[r * ldo + c] = from_f32<T>(((float)q[r * ldq + c] + z) * delta[r]);
  }
}

}  // namespace b200q

using namespace b200q;

extern "C" int b200q_quant_rows(const void* x, int x_dtype, int64_t rows, int64_t cols, int64_t ldx, int n_bits,
                                int sym, int dynamic, int8_t* q, int64_t ldq, float* delta, float* zero_point,
                                int32_t* rowsum, float* stat_max, float* stat_min, b200q_stream_t stream) {
  clear_error();
  B200Q_REQUIRE(rows >= 0 && cols >= 0, B200Q_ERR_BAD_ARG, "quant_rows: negative shape");
  if (rows == 0 || cols == 0) return B200Q_OK;
  B200Q_REQUIRE(x && q && delta && zero_point, B200Q_ERR_BAD_ARG, "quant_rows: null pointer");
  B200Q_REQUIRE(ldx >= cols && ldq >= cols, B200Q_ERR_BAD_ARG, "quant_rows: leading dimension < cols");
  B200Q_REQUIRE(rows <= 0x7fffffff, B200Q_ERR_UNSUPPORTED, "quant_rows: rows > 2^31-1");
  QuantArgs a{};
  a.x = x; a.rows = rows; a.cols = cols; a.ldx = ldx; a.q = q; a.ldq = ldq;
  a.delta = delta; a.zero_point = zero_point; a.rowsum = rowsum; a.apply_eps_floor = dynamic ? 1 : 0;
  a.stat_max = stat_max; a.stat_min = stat_min;
  int rc = fill_levels(a, n_bits, sym);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  return sym ? dispatch_dtype<kSymDyn>(x_dtype, a, st) : dispatch_dtype<kAsymDyn>(x_dtype, a, st);
}

extern "C" int b200q_quant_rows_static(const void* x, int x_dtype, int64_t rows, int64_t cols, int64_t ldx,
                                       int n_bits, int sym, const float* delta, const float* zero_point, int8_t* q,
                                       int64_t ldq, int32_t* rowsum, b200q_stream_t stream) {
  clear_error();
  B200Q_REQUIRE(rows >= 0 && cols >= 0, B200Q_ERR_BAD_ARG, "quant_rows_static: negative shape");
  if (rows == 0 || cols == 0) return B200Q_OK;
  B200Q_REQUIRE(x && q && delta && zero_point, B200Q_ERR_BAD_ARG, "quant_rows_static: null pointer");
  B200Q_REQUIRE(ldx >= cols && ldq >= cols, B200Q_ERR_BAD_ARG, "quant_rows_static: leading dimension < cols");
  B200Q_REQUIRE(rows <= 0x7fffffff, B200Q_ERR_UNSUPPORTED, "quant_rows_static: rows > 2^31-1");
  QuantArgs a{};
  a.x = x; a.rows = rows; a.cols = cols; a.ldx = ldx; a.q = q; a.ldq = ldq;
  a.delta = const_cast<float*>(delta); a.zero_point = const_cast<float*>(zero_point); a.rowsum = rowsum;
  int rc = fill_levels(a, n_bits, sym);
  if (rc) return rc;
  return dispatch_dtype<kStatic>(x_dtype, a, (cudaStream_t)stream);
}

extern "C" int b200q_dequant_rows(const int8_t* q, int64_t ldq, int64_t rows, int64_t cols, const float* delta,
                                  const float* zero_point, void* out, int out_dtype, int64_t ldo,
                                  b200q_stream_t stream) {
  clear_error();
  B200Q_REQUIRE(rows >= 0 && cols >= 0, B200Q_ERR_BAD_ARG, "dequant_rows: negative shape");
  if (rows == 0 || cols == 0) return B200Q_OK;
  B200Q_REQUIRE(q && delta && out, B200Q_ERR_BAD_ARG, "dequant_rows: null pointer");
  B200Q_REQUIRE(ldq >= cols && ldo >= cols, B200Q_ERR_BAD_ARG, "dequant_rows: leading dimension < cols");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t total = rows * cols;
  const unsigned grid = (unsigned)((total + 255) / 256 < (int64_t)sm_count() * 16 ? (total + 255) / 256
                                                                                   : (int64_t)sm_count() * 16);
  switch (out_dtype) {
    case B200Q_F32: dequant_rows_kernel<float><<<grid, 256, 0, st>>>(q, ldq, rows, cols, delta, zero_point, (float*)out, ldo); break;
    case B200Q_BF16: dequant_rows_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(q, ldq, rows, cols, delta, zero_point, (__nv_bfloat16*)out, ldo); break;
    case B200Q_F16: dequant_rows_kernel<__half><<<grid, 256, 0, st>>>(q, ldq, rows, cols, delta, zero_point, (__half*)out, ldo); break;
    default: set_error("dequant_rows: unsupported out_dtype %d", out_dtype); return B200Q_ERR_BAD_ARG;
  }
  B200Q_CHECK_LAUNCH();
  return B200Q_OK;
}
