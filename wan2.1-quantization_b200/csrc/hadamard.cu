// (f-2) ViDiT-Q / QuaRot / SmoothQuant activation pre-processing fused into the per-token quantizer:
//
//     y = (x * colscale) . (H_K (x) H_{2^m})            colscale[c] = channel_mask[c] * sign[c] / sqrt(n)
//     q = rne(y / delta),  delta = max|y| / n_levels     (DynamicQuantizer sym, base_quantizer.py:110-157)
//
// Replaces `x = x*channel_mask; x = torch.matmul(x.double(), rotation_matrix)` followed by the activation quantizer
// (ViDiT-Q/quant_utils/qdiff/viditq/viditq_quant_layer.py:58-66, quarot/quarot_quant_layer.py:55-62,
// smooth_quant/sq_quant_layer.py:55-58): a dense fp64 [L, n] x [n, n] product per linear in the reference.  The rotation
// R = diag(s) . H_n / sqrt(n) produced by random_hadamard_matrix (quarot_utils.py:186-192) is applied in its factored
// form H_n = H_K (x) H_{2^m} (matmul_hadU, :158-179): a fast Walsh-Hadamard transform over the 2^m-wide segments of the
// row (in registers + warp shuffles) and the order-K base block across the K segments.
//
// One CTA owns R rows in shared memory (fp32) and walks five phases: load*colscale -> FWHT (warp per segment) ->
// K-block (thread per 4-column group, K inputs in registers) -> row abs-max -> quantize + store.  Several CTAs are
// resident per SM, so the HBM phases of one overlap the arithmetic phases of another.  HBM-bound:
// algorithmic bytes = rows*cols*(sizeof(in)+1) + 8*rows.
#include "common.cuh"

namespace b200q {

struct HadArgs {
  const void* x;
  int64_t rows, cols, ldx;
  const float* colscale;   // [cols] or null
  const float* hadK;       // [K*K] row-major +-1, null iff K == 1
  int K, log2w;            // cols == K << log2w ; log2w == 0: no transform at all (colscale only)
  float n_levels;
  int8_t* q; int64_t ldq;
  float* delta; int32_t* rowsum;
  float* y_out; int64_t ldy;
  int R;                   // rows per CTA
};

constexpr int HAD_THREADS = 256;
__device__ __forceinline__ void named_bar_sync_64(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }
constexpr int HAD_KMAX = 32;

// natural-order FWHT of one 32*E-wide segment held as E consecutive values per lane
template <int E>
__device__ __forceinline__ void fwht_segment(float* seg, int lane) {
  float v[E];
  if constexpr (E >= 4) {
#pragma unroll
    for (int i = 0; i < E; i += 4) {
      const float4 t = *reinterpret_cast<const float4*>(seg + lane * E + i);
      v[i] = t.x; v[i + 1] = t.y; v[i + 2] = t.z; v[i + 3] = t.w;
    }
  } else if constexpr (E == 2) {
    const float2 t = *reinterpret_cast<const float2*>(seg + lane * 2);
    v[0] = t.x; v[1] = t.y;
  } else {
    v[0] = seg[lane];
  }
#pragma unroll
  for (int h = 1; h < E; h <<= 1) {
#pragma unroll
    for (int i = 0; i < E; ++i) {
      if ((i & h) == 0) {
        const float lo = v[i], hi = v[i + h];
        v[i] = lo + hi; v[i + h] = lo - hi;
      }
    }
  }
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float sgn = (lane & o) ? -1.f : 1.f;        // upper half of the butterfly: other - v, lower: v + other
#pragma unroll
    for (int i = 0; i < E; ++i) {
      const float other = __shfl_xor_sync(0xffffffffu, v[i], o);
      v[i] = fmaf(v[i], sgn, other);
    }
  }
  if constexpr (E >= 4) {
#pragma unroll
    for (int i = 0; i < E; i += 4)
      *reinterpret_cast<float4*>(seg + lane * E + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
  } else if constexpr (E == 2) {
    *reinterpret_cast<float2*>(seg + lane * 2) = make_float2(v[0], v[1]);
  } else {
    seg[lane] = v[0];
  }
}

// out[i][c..c+3] = sum_j h[i][j] * y[j][c..c+3], in place on one 4-column group (all K inputs are read first)
template <int KM>
__device__ __forceinline__ void kblock_group(float* base, int W, int K, const float2* s_h2) {
  uint64_t y01[KM], y23[KM];
#pragma unroll
  for (int j = 0; j < KM; ++j) {
    if (j < K) {
      const float4 t = *reinterpret_cast<const float4*>(base + (size_t)j * W);
      y01[j] = pack_f32x2(t.x, t.y); y23[j] = pack_f32x2(t.z, t.w);
    } else {
      y01[j] = 0; y23[j] = 0;
    }
  }
  for (int i = 0; i < K; ++i) {
    uint64_t a01 = 0, a23 = 0;                      // bit pattern of (0.f, 0.f)
    const float4* hrow = reinterpret_cast<const float4*>(s_h2 + i * K);   // K is even: 16-byte aligned rows
#pragma unroll
    for (int j = 0; j < KM; j += 2) {
      if (j < K) {
        const float4 hh = hrow[j >> 1];             // (h_j, h_j, h_j+1, h_j+1): one broadcast LDS.128
        const uint64_t h2a = pack_f32x2(hh.x, hh.y), h2b = pack_f32x2(hh.z, hh.w);
        a01 = fma_f32x2(h2a, y01[j], a01);
        a23 = fma_f32x2(h2a, y23[j], a23);
        a01 = fma_f32x2(h2b, y01[j + 1], a01);
        a23 = fma_f32x2(h2b, y23[j + 1], a23);
      }
    }
    float o0, o1, o2, o3;
    unpack_f32x2(a01, o0, o1); unpack_f32x2(a23, o2, o3);
    *reinterpret_cast<float4*>(base + (size_t)i * W) = make_float4(o0, o1, o2, o3);
  }
}

template <typename T, int KM>
__global__ void __launch_bounds__(HAD_THREADS, KM <= 12 ? 3 : (KM <= 20 ? 2 : 1)) had_quant_kernel(const HadArgs a) {
  using VT = Vec16<T>;
  constexpr int N = VT::N;
  extern __shared__ __align__(16) uint8_t had_smem[];
  const int n = (int)a.cols, R = a.R, K = a.K;
  float* buf = reinterpret_cast<float*>(had_smem);                       // [R][n]
  float2* s_h2 = reinterpret_cast<float2*>(buf + (size_t)R * n);         // [K*K] (h, h)
  int* s_amax = reinterpret_cast<int*>(s_h2 + K * K);                    // [R] fp32 bits (non-negative: int order == float order)
  int* s_sum = s_amax + R;                                               // [R]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int W = 1 << a.log2w;
  const int64_t row0 = (int64_t)blockIdx.x * R;

  if (K > 1) for (int i = tid; i < K * K; i += HAD_THREADS) { const float h = a.hadK[i]; s_h2[i] = make_float2(h, h); }
  if (tid < R) { s_amax[tid] = 0; s_sum[tid] = 0; }

  // ---- phase 1: load, * colscale -> smem (LD loads in flight per thread before the first use) ----
  const int kv = n / N;
  constexpr int LD = 4;
  for (int f0 = tid; f0 < R * kv; f0 += LD * HAD_THREADS) {
    uint4 raw[LD];
#pragma unroll
    for (int u = 0; u < LD; ++u) {
      const int f = f0 + u * HAD_THREADS;
      const int r = f / kv, v = f - r * kv;
      const int64_t row = row0 + r;
      raw[u] = (f < R * kv && row < a.rows) ? ldg_stream16(reinterpret_cast<const T*>(a.x) + row * a.ldx + (int64_t)v * N)
                                             : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < LD; ++u) {
      const int f = f0 + u * HAD_THREADS;
      if (f >= R * kv) break;
      const int r = f / kv, v = f - r * kv;
      float x[N];
      VT::unpack(raw[u], x);
      float* dst = buf + (size_t)r * n + v * N;
#pragma unroll
      for (int h = 0; h < N / 4; ++h) {
        float4 o = make_float4(x[4 * h], x[4 * h + 1], x[4 * h + 2], x[4 * h + 3]);
        if (a.colscale != nullptr) {
          const float4 c = __ldg(reinterpret_cast<const float4*>(a.colscale + v * N + 4 * h));
          o.x *= c.x; o.y *= c.y; o.z *= c.z; o.w *= c.w;
        }
        *reinterpret_cast<float4*>(dst + 4 * h) = o;
      }
    }
  }
  __syncthreads();

  // ---- phase 2: FWHT over every 2^m-wide segment (one warp per segment) ----
  if (a.log2w > 0) {
    const int nseg = R * K;
    for (int s = warp; s < nseg; s += HAD_THREADS / 32) {
      float* seg = buf + (size_t)s * W;               // segments of a row are contiguous, rows are contiguous
      switch (a.log2w) {
        case 5: fwht_segment<1>(seg, lane); break;
        case 6: fwht_segment<2>(seg, lane); break;
        case 7: fwht_segment<4>(seg, lane); break;
        default: fwht_segment<8>(seg, lane); break;   // 8
      }
    }
    __syncthreads();
  }

  // ---- phase 3: order-K base block across the segments ----
  if (K > 1) {
    const int G = W >> 2;
    for (int it = tid; it < R * G; it += HAD_THREADS) {
      const int r = it / G, c = it - r * G;
      float* base = buf + (size_t)r * n + c * 4;
      kblock_group<KM>(base, W, K, s_h2);
    }
    __syncthreads();
  }

  // ---- phases 4 + 5: per-row abs-max, then quantize + store.  A row is owned by WPR = 8 / R warps (R in {1,2,4,8}), so
  // the reductions are thread-local sums plus one warp reduction and one shared-memory atomic per warp and row. ----
  const int n4 = n >> 2;
  const int WPR = (HAD_THREADS / 32) / R;             // warps per row
  const int r = warp / WPR, part = warp - r * WPR;
  const int64_t row = row0 + r;
  const float4* rowp = reinterpret_cast<const float4*>(buf + (size_t)r * n);
  {
    float m = 0.f;
    for (int f = part * 32 + lane; f < n4; f += WPR * 32) {
      const float4 t = rowp[f];
      m = fmaxf(m, fmaxf(fmaxf(fabsf(t.x), fabsf(t.y)), fmaxf(fabsf(t.z), fabsf(t.w))));
    }
    m = warp_max(m);
    if (lane == 0) atomicMax(&s_amax[r], __float_as_int(m));
  }
  __syncthreads();
  {
    float delta = __fdiv_rn(__int_as_float(s_amax[r]), a.n_levels);
    if (delta < 1.0e-6f) delta = 1.0e-6f;                                // base_quantizer.py:122-128
    const float rc = __frcp_rn(delta);
    const uint64_t r2 = pack_f32x2(rc, rc), nd2 = pack_f32x2(-delta, -delta), magic2 = pack_f32x2(12582912.0f, 12582912.0f);
    int sum = 0;
    const bool row_ok = row < a.rows;
    for (int f = part * 32 + lane; f < n4; f += WPR * 32) {
      const float4 t = rowp[f];
      uint32_t c0, c1, c2, c3;
      unpack_u32x2(div_rn_hoisted_rne2(pack_f32x2(t.x, t.y), nd2, r2, magic2), c0, c1);
      unpack_u32x2(div_rn_hoisted_rne2(pack_f32x2(t.z, t.w), nd2, r2, magic2), c2, c3);
      const uint32_t packed = __byte_perm(__byte_perm(c0, c1, 0x0040), __byte_perm(c2, c3, 0x0040), 0x5410);
      sum = __dp4a((int)packed, 0x01010101, sum);
      if (row_ok) {
        stg_stream4(a.q + row * a.ldq + (int64_t)f * 4, packed);
        if (a.y_out != nullptr) *reinterpret_cast<float4*>(a.y_out + row * a.ldy + (int64_t)f * 4) = t;
      }
    }
    if (a.rowsum != nullptr) {
      sum = warp_sum(sum);
      if (lane == 0) atomicAdd(&s_sum[r], sum);
    }
    if (part == 0 && lane == 0 && row_ok) a.delta[row] = delta;
  }
  if (a.rowsum != nullptr) {
    __syncthreads();
    if (tid < R && row0 + tid < a.rows) a.rowsum[row0 + tid] = s_sum[tid];
  }
}

// ---- register-resident variant: one warp owns one row of n = K * 128 channels ---------------------------------------
// Lane l holds positions 4l .. 4l+3 of every one of the K segments (4K registers), so
//   * the order-K base block (same position, all segments) is lane-local: K*K multiply-adds per position from registers,
//     the +-1 matrix broadcast from shared memory; it runs first - H_K (x) I and I (x) H_128 commute;
//   * the 128-point Walsh-Hadamard transform of a segment is 2 in-lane stages + 5 shuffle stages;
//   * abs-max, exact division, packing and the code row sum follow without the row ever leaving the registers:
//     one HBM read, one HBM write, no shared-memory round trips (the CTA-tile kernel below makes five).
// ~1,100 warp instructions per row against ~3,000 (ncu) for the shared-memory kernel at n = 1536.
template <typename T, int K>
__global__ void __launch_bounds__(128, 3) had_quant_warp_kernel(const HadArgs a) {
  __shared__ float s_h[K * K];
  for (int i = threadIdx.x; i < K * K; i += 128) s_h[i] = a.hadK[i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 4 + warp;
  if (row >= a.rows) return;
  const T* xrow = reinterpret_cast<const T*>(a.x) + row * a.ldx + lane * 4;

  // ---- load (one 16- or 8-byte vector per segment), * colscale ----
  uint64_t v01[K], v23[K];                          // (position 4l, 4l+1) and (4l+2, 4l+3) of segment s, packed fp32x2
#pragma unroll
  for (int s = 0; s < K; ++s) {
    float f0, f1, f2, f3;
    if constexpr (sizeof(T) == 4) {
      const uint4 r = ldg_stream16(xrow + s * 128);
      f0 = __uint_as_float(r.x); f1 = __uint_as_float(r.y); f2 = __uint_as_float(r.z); f3 = __uint_as_float(r.w);
    } else {
      const uint2 r = *reinterpret_cast<const uint2*>(xrow + s * 128);
      if constexpr (std::is_same<T, __nv_bfloat16>::value) {
        f0 = __uint_as_float(r.x << 16); f1 = __uint_as_float(r.x & 0xffff0000u);
        f2 = __uint_as_float(r.y << 16); f3 = __uint_as_float(r.y & 0xffff0000u);
      } else {
        const float2 a2 = __half22float2(*reinterpret_cast<const __half2*>(&r.x)), b2 = __half22float2(*reinterpret_cast<const __half2*>(&r.y));
        f0 = a2.x; f1 = a2.y; f2 = b2.x; f3 = b2.y;
      }
    }
    v01[s] = pack_f32x2(f0, f1); v23[s] = pack_f32x2(f2, f3);
    if (a.colscale != nullptr) {
      const float4 c = __ldg(reinterpret_cast<const float4*>(a.colscale + s * 128 + lane * 4));
      v01[s] = mul_f32x2(v01[s], pack_f32x2(c.x, c.y));
      v23[s] = mul_f32x2(v23[s], pack_f32x2(c.z, c.w));
    }
  }

  // ---- order-K base block across the segments, one position pair at a time (keeps 6K live registers, not 8K) ----
  {
    uint64_t o[K];
#pragma unroll
    for (int i = 0; i < K; ++i) {
      uint64_t acc = 0;                              // (0.f, 0.f)
#pragma unroll
      for (int j = 0; j < K; ++j) { const float h = s_h[i * K + j]; acc = fma_f32x2(pack_f32x2(h, h), v01[j], acc); }
      o[i] = acc;
    }
#pragma unroll
    for (int i = 0; i < K; ++i) v01[i] = o[i];
#pragma unroll
    for (int i = 0; i < K; ++i) {
      uint64_t acc = 0;
#pragma unroll
      for (int j = 0; j < K; ++j) { const float h = s_h[i * K + j]; acc = fma_f32x2(pack_f32x2(h, h), v23[j], acc); }
      o[i] = acc;
    }
#pragma unroll
    for (int i = 0; i < K; ++i) v23[i] = o[i];
  }

  // ---- 128-point FWHT of every segment: stages over the 4 in-lane positions, then over the 32 lanes ----
  const uint64_t pm = pack_f32x2(1.f, -1.f);
#pragma unroll
  for (int s = 0; s < K; ++s) {
    float a0, a1, a2, a3;
    unpack_f32x2(v01[s], a0, a1); unpack_f32x2(v23[s], a2, a3);
    // stage h = 1: (a0, a1) -> (a0 + a1, a0 - a1); stage h = 2: pairs (0,2), (1,3)
    const uint64_t p = fma_f32x2(pack_f32x2(a1, a1), pm, pack_f32x2(a0, a0));      // (a0 + a1, a0 - a1)
    const uint64_t q = fma_f32x2(pack_f32x2(a3, a3), pm, pack_f32x2(a2, a2));      // (a2 + a3, a2 - a3)
    v01[s] = add_f32x2(p, q);
    v23[s] = fma_f32x2(q, pack_f32x2(-1.f, -1.f), p);
  }
#pragma unroll
  for (int ofs = 1; ofs < 32; ofs <<= 1) {
    const float sg = (lane & ofs) ? -1.f : 1.f;     // upper half of the butterfly: other - v, lower: v + other
    const uint64_t sg2 = pack_f32x2(sg, sg);
#pragma unroll
    for (int s = 0; s < K; ++s) {
      uint32_t lo, hi;
      unpack_u32x2(v01[s], lo, hi);
      const uint64_t o01 = pack_u32x2(__shfl_xor_sync(0xffffffffu, lo, ofs), __shfl_xor_sync(0xffffffffu, hi, ofs));
      v01[s] = fma_f32x2(v01[s], sg2, o01);
      unpack_u32x2(v23[s], lo, hi);
      const uint64_t o23 = pack_u32x2(__shfl_xor_sync(0xffffffffu, lo, ofs), __shfl_xor_sync(0xffffffffu, hi, ofs));
      v23[s] = fma_f32x2(v23[s], sg2, o23);
    }
  }

  // ---- per-token abs-max -> delta ----
  float m = 0.f;
#pragma unroll
  for (int s = 0; s < K; ++s) {
    float b0, b1, b2, b3;
    unpack_f32x2(v01[s], b0, b1); unpack_f32x2(v23[s], b2, b3);
    m = fmaxf(m, fmaxf(fmaxf(fabsf(b0), fabsf(b1)), fmaxf(fabsf(b2), fabsf(b3))));
  }
  m = warp_max(m);
  float delta = __fdiv_rn(m, a.n_levels);
  if (delta < 1.0e-6f) delta = 1.0e-6f;                                  // base_quantizer.py:122-128
  const float rc = __frcp_rn(delta);
  const uint64_t r2 = pack_f32x2(rc, rc), nd2 = pack_f32x2(-delta, -delta), magic2 = pack_f32x2(12582912.0f, 12582912.0f);

  // ---- quantize, pack 4 codes per lane and segment, store ----
  int sum = 0;
  int8_t* qrow = a.q + row * a.ldq + lane * 4;
#pragma unroll
  for (int s = 0; s < K; ++s) {
    uint32_t c0, c1, c2, c3;
    unpack_u32x2(div_rn_hoisted_rne2(v01[s], nd2, r2, magic2), c0, c1);
    unpack_u32x2(div_rn_hoisted_rne2(v23[s], nd2, r2, magic2), c2, c3);
    const uint32_t packed = __byte_perm(__byte_perm(c0, c1, 0x0040), __byte_perm(c2, c3, 0x0040), 0x5410);
    sum = __dp4a((int)packed, 0x01010101, sum);
    stg_stream4(qrow + s * 128, packed);
    if (a.y_out != nullptr) {
      float b0, b1, b2, b3;
      unpack_f32x2(v01[s], b0, b1); unpack_f32x2(v23[s], b2, b3);
      *reinterpret_cast<float4*>(a.y_out + row * a.ldy + s * 128 + lane * 4) = make_float4(b0, b1, b2, b3);
    }
  }
  if (a.rowsum != nullptr) {
    sum = warp_sum(sum);
    if (lane == 0) a.rowsum[row] = sum;
  }
  if (lane == 0) a.delta[row] = delta;
}

template <typename T, int K>
static int launch_had_warp(const HadArgs& a, cudaStream_t st) {
  had_quant_warp_kernel<T, K><<<(unsigned)((a.rows + 3) / 4), 128, 0, st>>>(a);
  B200Q_CHECK_LAUNCH();
  return B200Q_OK;
}

// ---- register-resident variant for n = K * 256 channels (Wan-14B hidden: 5120 = 20 * 256): TWO warps own one row ---------
// Thread (w, l) holds positions w*128 + 4l .. + 3 of every one of the K segments.  Base block and the first seven
// Walsh-Hadamard stages are those of the one-warp kernel; the eighth stage pairs the two warps through shared memory
// (one float4 per segment and thread, [segment][thread] so that a warp's accesses are conflict-free); the row abs-max and
// the code row sum meet through two shared words.  One HBM read, one HBM write.
template <typename T, int K, int MINB>
__global__ void __launch_bounds__(128, MINB) had_quant_warp2_kernel(const HadArgs a) {
  __shared__ __align__(16) float s_h[K * K];
  __shared__ float4 s_x[2][K][64];                  // [row of the CTA][segment][thread of the row]
  __shared__ float s_m[2][2];
  __shared__ int s_s[2][2];
  for (int i = threadIdx.x; i < K * K; i += 128) s_h[i] = a.hadK[i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = warp >> 1, w = warp & 1;            // row of the CTA, half of the row
  const int t = w * 32 + lane;
  const int64_t row = (int64_t)blockIdx.x * 2 + r;
  if (row >= a.rows) return;                         // both warps of a row leave together
  const int pos = w * 128 + lane * 4;
  const T* xrow = reinterpret_cast<const T*>(a.x) + row * a.ldx + pos;

  uint64_t v01[K], v23[K];
#pragma unroll
  for (int s = 0; s < K; ++s) {
    float f0, f1, f2, f3;
    if constexpr (sizeof(T) == 4) {
      const uint4 q = ldg_stream16(xrow + s * 256);
      f0 = __uint_as_float(q.x); f1 = __uint_as_float(q.y); f2 = __uint_as_float(q.z); f3 = __uint_as_float(q.w);
    } else {
      const uint2 q = *reinterpret_cast<const uint2*>(xrow + s * 256);
      if constexpr (std::is_same<T, __nv_bfloat16>::value) {
        f0 = __uint_as_float(q.x << 16); f1 = __uint_as_float(q.x & 0xffff0000u);
        f2 = __uint_as_float(q.y << 16); f3 = __uint_as_float(q.y & 0xffff0000u);
      } else {
        const float2 a2 = __half22float2(*reinterpret_cast<const __half2*>(&q.x)), b2 = __half22float2(*reinterpret_cast<const __half2*>(&q.y));
        f0 = a2.x; f1 = a2.y; f2 = b2.x; f3 = b2.y;
      }
    }
    v01[s] = pack_f32x2(f0, f1); v23[s] = pack_f32x2(f2, f3);
    if (a.colscale != nullptr) {
      const float4 c = __ldg(reinterpret_cast<const float4*>(a.colscale + s * 256 + pos));
      v01[s] = mul_f32x2(v01[s], pack_f32x2(c.x, c.y));
      v23[s] = mul_f32x2(v23[s], pack_f32x2(c.z, c.w));
    }
  }

  // ---- order-K base block across the segments (lane-local), one position pair at a time ----
  {
    uint64_t o[K];
#pragma unroll
    for (int i = 0; i < K; ++i) {
      uint64_t acc = 0;
#pragma unroll
      for (int j = 0; j < K; ++j) { const float h = s_h[i * K + j]; acc = fma_f32x2(pack_f32x2(h, h), v01[j], acc); }
      o[i] = acc;
    }
#pragma unroll
    for (int i = 0; i < K; ++i) v01[i] = o[i];
#pragma unroll
    for (int i = 0; i < K; ++i) {
      uint64_t acc = 0;
#pragma unroll
      for (int j = 0; j < K; ++j) { const float h = s_h[i * K + j]; acc = fma_f32x2(pack_f32x2(h, h), v23[j], acc); }
      o[i] = acc;
    }
#pragma unroll
    for (int i = 0; i < K; ++i) v23[i] = o[i];
  }

  // ---- 256-point FWHT of every segment: 2 in-lane stages, 5 shuffle stages, 1 stage across the two warps ----
  const uint64_t pm = pack_f32x2(1.f, -1.f);
#pragma unroll
  for (int s = 0; s < K; ++s) {
    float a0, a1, a2, a3;
    unpack_f32x2(v01[s], a0, a1); unpack_f32x2(v23[s], a2, a3);
    const uint64_t p = fma_f32x2(pack_f32x2(a1, a1), pm, pack_f32x2(a0, a0));      // (a0 + a1, a0 - a1)
    const uint64_t q = fma_f32x2(pack_f32x2(a3, a3), pm, pack_f32x2(a2, a2));      // (a2 + a3, a2 - a3)
    v01[s] = add_f32x2(p, q);
    v23[s] = fma_f32x2(q, pack_f32x2(-1.f, -1.f), p);
  }
#pragma unroll
  for (int ofs = 1; ofs < 32; ofs <<= 1) {
    const float sg = (lane & ofs) ? -1.f : 1.f;     // upper half of the butterfly: other - v, lower: v + other
    const uint64_t sg2 = pack_f32x2(sg, sg);
#pragma unroll
    for (int s = 0; s < K; ++s) {
      uint32_t lo, hi;
      unpack_u32x2(v01[s], lo, hi);
      const uint64_t o01 = pack_u32x2(__shfl_xor_sync(0xffffffffu, lo, ofs), __shfl_xor_sync(0xffffffffu, hi, ofs));
      v01[s] = fma_f32x2(v01[s], sg2, o01);
      unpack_u32x2(v23[s], lo, hi);
      const uint64_t o23 = pack_u32x2(__shfl_xor_sync(0xffffffffu, lo, ofs), __shfl_xor_sync(0xffffffffu, hi, ofs));
      v23[s] = fma_f32x2(v23[s], sg2, o23);
    }
  }
  {
#pragma unroll
    for (int s = 0; s < K; ++s) {
      float b0, b1, b2, b3;
      unpack_f32x2(v01[s], b0, b1); unpack_f32x2(v23[s], b2, b3);
      s_x[r][s][t] = make_float4(b0, b1, b2, b3);
    }
    named_bar_sync_64(1 + r);
    const uint64_t sg2 = w ? pack_f32x2(-1.f, -1.f) : pack_f32x2(1.f, 1.f);         // warp 1 (upper half): other - v
#pragma unroll
    for (int s = 0; s < K; ++s) {
      const float4 o = s_x[r][s][t ^ 32];
      v01[s] = fma_f32x2(v01[s], sg2, pack_f32x2(o.x, o.y));
      v23[s] = fma_f32x2(v23[s], sg2, pack_f32x2(o.z, o.w));
    }
  }

  // ---- per-token abs-max -> delta ----
  float m = 0.f;
#pragma unroll
  for (int s = 0; s < K; ++s) {
    float b0, b1, b2, b3;
    unpack_f32x2(v01[s], b0, b1); unpack_f32x2(v23[s], b2, b3);
    m = fmaxf(m, fmaxf(fmaxf(fabsf(b0), fabsf(b1)), fmaxf(fabsf(b2), fabsf(b3))));
  }
  m = warp_max(m);
  if (lane == 0) s_m[r][w] = m;
  named_bar_sync_64(1 + r);
  m = fmaxf(s_m[r][0], s_m[r][1]);
  float delta = __fdiv_rn(m, a.n_levels);
  if (delta < 1.0e-6f) delta = 1.0e-6f;                                  // base_quantizer.py:122-128
  const float rc = __frcp_rn(delta);
  const uint64_t r2 = pack_f32x2(rc, rc), nd2 = pack_f32x2(-delta, -delta), magic2 = pack_f32x2(12582912.0f, 12582912.0f);

  // ---- quantize, pack 4 codes per lane and segment, store ----
  int sum = 0;
  int8_t* qrow = a.q + row * a.ldq + pos;
#pragma unroll
  for (int s = 0; s < K; ++s) {
    uint32_t c0, c1, c2, c3;
    unpack_u32x2(div_rn_hoisted_rne2(v01[s], nd2, r2, magic2), c0, c1);
    unpack_u32x2(div_rn_hoisted_rne2(v23[s], nd2, r2, magic2), c2, c3);
    const uint32_t packed = __byte_perm(__byte_perm(c0, c1, 0x0040), __byte_perm(c2, c3, 0x0040), 0x5410);
    sum = __dp4a((int)packed, 0x01010101, sum);
    stg_stream4(qrow + s * 256, packed);
    if (a.y_out != nullptr) {
      float b0, b1, b2, b3;
      unpack_f32x2(v01[s], b0, b1); unpack_f32x2(v23[s], b2, b3);
      *reinterpret_cast<float4*>(a.y_out + row * a.ldy + s * 256 + pos) = make_float4(b0, b1, b2, b3);
    }
  }
  if (a.rowsum != nullptr) {
    sum = warp_sum(sum);
    if (lane == 0) s_s[r][w] = sum;
    named_bar_sync_64(1 + r);
    if (t == 0) a.rowsum[row] = s_s[r][0] + s_s[r][1];
  }
  if (t == 0) a.delta[row] = delta;
}

template <typename T, int K, int MINB>
static int launch_had_warp2(const HadArgs& a, cudaStream_t st) {
  had_quant_warp2_kernel<T, K, MINB><<<(unsigned)((a.rows + 1) / 2), 128, 0, st>>>(a);
  B200Q_CHECK_LAUNCH();
  return B200Q_OK;
}

template <typename T, int KM>
static int launch_had_km(const HadArgs& a0, cudaStream_t st) {
  HadArgs a = a0;
  const int n = (int)a.cols;
  const int W = 1 << a.log2w;
  // rows per CTA: enough 4-column groups for 256 threads in the K-block phase, within ~96 KB of shared memory
  (void)W;
  int R = 8;                                           // power of two <= 8: a row is owned by 8 / R warps
  while (R > 1 && (size_t)R * n * 4 > 98304) R >>= 1;
  a.R = R;
  const size_t smem = (size_t)R * n * 4 + (size_t)a.K * a.K * 8 + (size_t)R * 8 + 16;
  B200Q_REQUIRE(smem <= 232448, B200Q_ERR_UNSUPPORTED, "had_quant_rows: cols=%d does not fit in shared memory", n);
  static bool configured = false;
  if (!configured) {
    B200Q_CUDA_OK(cudaFuncSetAttribute(had_quant_kernel<T, KM>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    configured = true;
  }
  const unsigned grid = (unsigned)((a.rows + R - 1) / R);
  had_quant_kernel<T, KM><<<grid, HAD_THREADS, smem, st>>>(a);
  B200Q_CHECK_LAUNCH();
  return B200Q_OK;
}

static int g_had_warp = 1;          // 0: always the shared-memory tile kernel (tests compare the two)

template <typename T>
static int launch_had(const HadArgs& a, cudaStream_t st) {
  const bool vec_ok = aligned(a.x, 16) && (a.ldx * sizeof(T)) % 16 == 0 && a.ldq % 4 == 0;
  if (g_had_warp && a.log2w == 7 && vec_ok) {       // n = K * 128: the register-resident warp-per-row kernel
    if (a.K == 8) return launch_had_warp<T, 8>(a, st);
    if (a.K == 12) return launch_had_warp<T, 12>(a, st);   // 1536 = Wan-1.3B hidden
  }
  if (g_had_warp && a.log2w == 8 && a.K == 20 && vec_ok)   // 5120 = Wan-14B hidden: two warps per row
    return g_had_warp == 2 ? launch_had_warp2<T, 20, 3>(a, st) : launch_had_warp2<T, 20, 2>(a, st);
  if (a.K <= 12) return launch_had_km<T, 12>(a, st);
  if (a.K <= 20) return launch_had_km<T, 20>(a, st);
  return launch_had_km<T, HAD_KMAX>(a, st);
}

}  // namespace b200q

using namespace b200q;

extern "C" int b200q_had_set_mode(int warp_kernel) {
  g_had_warp = warp_kernel < 0 ? 0 : (warp_kernel > 2 ? 1 : warp_kernel);   // 2: probe variant (3 CTAs per SM, spills)
  return B200Q_OK;
}

extern "C" int b200q_had_quant_rows(const void* x, int x_dtype, int64_t rows, int64_t cols, int64_t ldx,
                                    const float* colscale, const float* hadK, int K, int log2_width, int n_bits,
                                    int8_t* q, int64_t ldq, float* delta, int32_t* rowsum, float* y_out, int64_t ldy,
                                    b200q_stream_t stream) {
  clear_error();
  B200Q_REQUIRE(rows >= 0 && cols >= 0, B200Q_ERR_BAD_ARG, "had_quant_rows: negative shape");
  if (rows == 0 || cols == 0) return B200Q_OK;
  B200Q_REQUIRE(x && q && delta, B200Q_ERR_BAD_ARG, "had_quant_rows: null pointer");
  B200Q_REQUIRE(n_bits >= 2 && n_bits <= 8, B200Q_ERR_BAD_ARG, "had_quant_rows: n_bits=%d out of [2,8]", n_bits);
  B200Q_REQUIRE(K >= 1 && K <= HAD_KMAX, B200Q_ERR_UNSUPPORTED, "had_quant_rows: base block order K=%d not in [1,%d]", K, HAD_KMAX);
  B200Q_REQUIRE((K == 1) == (hadK == nullptr), B200Q_ERR_BAD_ARG, "had_quant_rows: hadK is given iff K > 1");
  B200Q_REQUIRE(log2_width == 0 || (log2_width >= 5 && log2_width <= 8), B200Q_ERR_UNSUPPORTED,
                "had_quant_rows: segment width 2^%d not in {1 (no transform), 32..256}", log2_width);
  B200Q_REQUIRE(log2_width != 0 || K == 1, B200Q_ERR_BAD_ARG, "had_quant_rows: log2_width == 0 means no transform (K must be 1)");
  B200Q_REQUIRE(log2_width == 0 || cols == ((int64_t)K << log2_width), B200Q_ERR_BAD_ARG,
                "had_quant_rows: cols=%lld != K * 2^log2_width", (long long)cols);
  B200Q_REQUIRE(cols % 128 == 0 && cols <= 57344, B200Q_ERR_UNSUPPORTED, "had_quant_rows: cols must be a multiple of 128, <= 57344");
  const int vecn = x_dtype == B200Q_F32 ? 4 : 8;
  B200Q_REQUIRE(x_dtype >= B200Q_F32 && x_dtype <= B200Q_F16, B200Q_ERR_BAD_ARG, "had_quant_rows: bad x_dtype");
  B200Q_REQUIRE(ldx >= cols && ldx % vecn == 0 && aligned(x, 16), B200Q_ERR_UNSUPPORTED, "had_quant_rows: x misaligned");
  B200Q_REQUIRE(ldq >= cols && ldq % 4 == 0 && aligned(q, 4), B200Q_ERR_UNSUPPORTED, "had_quant_rows: q misaligned");
  B200Q_REQUIRE(!colscale || aligned(colscale, 16), B200Q_ERR_BAD_ARG, "had_quant_rows: colscale must be 16-byte aligned");
  B200Q_REQUIRE(!y_out || (ldy >= cols && ldy % 4 == 0 && aligned(y_out, 16)), B200Q_ERR_BAD_ARG, "had_quant_rows: y_out misaligned");
  HadArgs a{};
  a.x = x; a.rows = rows; a.cols = cols; a.ldx = ldx; a.colscale = colscale; a.hadK = hadK; a.K = K; a.log2w = log2_width;
  a.n_levels = (float)((1 << (n_bits - 1)) - 1);
  a.q = q; a.ldq = ldq; a.delta = delta; a.rowsum = rowsum; a.y_out = y_out; a.ldy = ldy;
  cudaStream_t st = (cudaStream_t)stream;
  switch (x_dtype) {
    case B200Q_F32: return launch_had<float>(a, st);
    case B200Q_BF16: return launch_had<__nv_bfloat16>(a, st);
    default: return launch_had<__half>(a, st);
  }
}
