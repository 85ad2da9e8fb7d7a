// (d) Calibration reduction: ONE read of x[rows, cols] -> running per-channel abs-max / min / max.
//
// Replaces the forward-hook body `module_in[0].reshape(-1,C).abs().max(dim=0)[0]`
// (ViDiT-Q/examples/Wan2.1/get_calib_data_wanx.py:262-263; two extra full passes per linear
// per call) and the stack/cat/max merge (:443-468, ptq_wanx.py:336), which is algebraically a
// running elementwise max.  HBM-bound: algorithmic bytes = rows*cols*sizeof(in) + 12*cols.
//
// Layout: thread (tx) owns one 16-byte vector column group; blockDim.y row lanes stride through
// the CTA's row slab with 4 independent loads in flight; the CTA combines its row lanes in shared
// memory and issues one atomic per (channel, statistic) — max is exact and order independent, so
// the result is bit-identical to the reference for any slab decomposition (and any rank split).
#include "common.cuh"

namespace b200q {

// float atomic max/min through integer atomics (valid for mixed signs, NaN never written)
__device__ __forceinline__ void atomic_max_f32(float* addr, float v) {
  if (v >= 0.f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}
__device__ __forceinline__ void atomic_min_f32(float* addr, float v) {
  if (v >= 0.f) atomicMin(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMax(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

constexpr int kCalibTX = 64;   // vector columns per CTA
constexpr int kCalibTY = 4;    // row lanes per CTA

template <typename T, bool MINMAX>
__global__ void __launch_bounds__(kCalibTX * kCalibTY) calib_kernel(const T* __restrict__ x, int64_t rows,
                                                                     int64_t cols, int64_t ldx, int64_t rows_per_slab,
                                                                     float* absmax_io, float* min_io, float* max_io) {
  using VT = Vec16<T>;
  constexpr int N = VT::N;
  __shared__ float s_amax[kCalibTY][kCalibTX * N + 1];
  __shared__ float s_min[MINMAX ? kCalibTY : 1][MINMAX ? kCalibTX * N + 1 : 1];
  __shared__ float s_max[MINMAX ? kCalibTY : 1][MINMAX ? kCalibTX * N + 1 : 1];

  const int tx = threadIdx.x, ty = threadIdx.y;
  const int64_t vcol = (int64_t)blockIdx.x * kCalibTX + tx;          // vector column index
  const int64_t c0 = vcol * N;
  const bool col_ok = c0 < cols;                                      // cols % N == 0 on this path
  const int64_t r_begin = (int64_t)blockIdx.y * rows_per_slab;
  const int64_t r_end = min(rows, r_begin + rows_per_slab);

  float amax[N], vmin[N], vmax[N];
#pragma unroll
  for (int i = 0; i < N; ++i) { amax[i] = 0.f; vmin[i] = INFINITY; vmax[i] = -INFINITY; }

  if (col_ok) {
    const T* p = x + c0;
    int64_t r = r_begin + ty;
    for (; r + 3 * kCalibTY < r_end; r += 4 * kCalibTY) {            // 4 loads in flight per thread
      uint4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = ldg_stream16(p + (r + u * kCalibTY) * ldx);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float f[N];
        VT::unpack(v[u], f);
#pragma unroll
        for (int i = 0; i < N; ++i) {
          amax[i] = fmaxf(amax[i], fabsf(f[i]));
          if (MINMAX) { vmin[i] = fminf(vmin[i], f[i]); vmax[i] = fmaxf(vmax[i], f[i]); }
        }
      }
    }
    for (; r < r_end; r += kCalibTY) {
      float f[N];
      VT::unpack(ldg_stream16(p + r * ldx), f);
#pragma unroll
      for (int i = 0; i < N; ++i) {
        amax[i] = fmaxf(amax[i], fabsf(f[i]));
        if (MINMAX) { vmin[i] = fminf(vmin[i], f[i]); vmax[i] = fmaxf(vmax[i], f[i]); }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < N; ++i) {
    s_amax[ty][tx * N + i] = amax[i];
    if (MINMAX) { s_min[ty][tx * N + i] = vmin[i]; s_max[ty][tx * N + i] = vmax[i]; }
  }
  __syncthreads();
  // one thread per channel of the CTA's column range
  for (int c = ty * kCalibTX + tx; c < kCalibTX * N; c += kCalibTX * kCalibTY) {
    const int64_t gc = (int64_t)blockIdx.x * kCalibTX * N + c;
    if (gc >= cols) continue;
    float a = s_amax[0][c];
#pragma unroll
    for (int y = 1; y < kCalibTY; ++y) a = fmaxf(a, s_amax[y][c]);
    if (absmax_io) atomicMax(reinterpret_cast<int*>(absmax_io + gc), __float_as_int(a));  // a >= 0
    if (MINMAX) {
      float lo = s_min[0][c], hi = s_max[0][c];
#pragma unroll
      for (int y = 1; y < kCalibTY; ++y) { lo = fminf(lo, s_min[y][c]); hi = fmaxf(hi, s_max[y][c]); }
      if (min_io && lo != INFINITY) atomic_min_f32(min_io + gc, lo);
      if (max_io && hi != -INFINITY) atomic_max_f32(max_io + gc, hi);
    }
  }
}

// generic path (cols % N != 0 or unaligned): one thread per channel, slab of rows per CTA row
template <typename T>
__global__ void __launch_bounds__(256) calib_generic_kernel(const T* __restrict__ x, int64_t rows, int64_t cols,
                                                             int64_t ldx, int64_t rows_per_slab, float* absmax_io,
                                                             float* min_io, float* max_io) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  const int64_t r_begin = (int64_t)blockIdx.y * rows_per_slab;
  const int64_t r_end = min(rows, r_begin + rows_per_slab);
  float a = 0.f, lo = INFINITY, hi = -INFINITY;
  for (int64_t r = r_begin; r < r_end; ++r) {
    const float f = to_f32(x[r * ldx + c]);
    a = fmaxf(a, fabsf(f)); lo = fminf(lo, f); hi = fmaxf(hi, f);
  }
  if (absmax_io) atomicMax(reinterpret_cast<int*>(absmax_io + c), __float_as_int(a));
  if (min_io && lo != INFINITY) atomic_min_f32(min_io + c, lo);
  if (max_io && hi != -INFINITY) atomic_max_f32(max_io + c, hi);
}

template <typename T>
static int launch_calib(const void* x, int64_t rows, int64_t cols, int64_t ldx, float* absmax_io, float* min_io,
                        float* max_io, cudaStream_t st) {
  constexpr int N = Vec16<T>::N;
  const T* xp = reinterpret_cast<const T*>(x);
  const bool fast = (cols % N == 0) && (ldx % N == 0) && aligned(x, 16);
  // slabs: enough CTAs for ~8 per SM, but at least 32 rows each so the atomics stay negligible
  const int64_t col_ctas = fast ? (cols / N + kCalibTX - 1) / kCalibTX : (cols + 255) / 256;
  int64_t slabs = ((int64_t)sm_count() * 8 + col_ctas - 1) / col_ctas;
  if (slabs > (rows + 31) / 32) slabs = (rows + 31) / 32;
  if (slabs < 1) slabs = 1;
  if (slabs > 65535) slabs = 65535;
  const int64_t rows_per_slab = (rows + slabs - 1) / slabs;
  slabs = (rows + rows_per_slab - 1) / rows_per_slab;
  dim3 grid((unsigned)col_ctas, (unsigned)slabs);
  if (fast) {
    dim3 block(kCalibTX, kCalibTY);
    if (min_io || max_io) calib_kernel<T, true><<<grid, block, 0, st>>>(xp, rows, cols, ldx, rows_per_slab, absmax_io, min_io, max_io);
    else calib_kernel<T, false><<<grid, block, 0, st>>>(xp, rows, cols, ldx, rows_per_slab, absmax_io, min_io, max_io);
  } else {
    calib_generic_kernel<T><<<grid, 256, 0, st>>>(xp, rows, cols, ldx, rows_per_slab, absmax_io, min_io, max_io);
  }
  B200Q_CHECK_LAUNCH();
  return B200Q_OK;
}

}  // namespace b200q

using namespace b200q;

extern "C" int b200q_calib_absmax_minmax(const void* x, int x_dtype, int64_t rows, int64_t cols, int64_t ldx,
                                         float* absmax_io, float* min_io, float* max_io, b200q_stream_t stream) {
  clear_error();
  B200Q_REQUIRE(rows >= 0 && cols >= 0, B200Q_ERR_BAD_ARG, "calib: negative shape");
  if (rows == 0 || cols == 0) return B200Q_OK;
  B200Q_REQUIRE(x != nullptr, B200Q_ERR_BAD_ARG, "calib: null x");
  B200Q_REQUIRE(absmax_io || min_io || max_io, B200Q_ERR_BAD_ARG, "calib: no output buffer given");
  B200Q_REQUIRE(ldx >= cols, B200Q_ERR_BAD_ARG, "calib: ldx < cols");
  cudaStream_t st = (cudaStream_t)stream;
  switch (x_dtype) {
    case B200Q_F32: return launch_calib<float>(x, rows, cols, ldx, absmax_io, min_io, max_io, st);
    case B200Q_BF16: return launch_calib<__nv_bfloat16>(x, rows, cols, ldx, absmax_io, min_io, max_io, st);
    case B200Q_F16: return launch_calib<__half>(x, rows, cols, ldx, absmax_io, min_io, max_io, st);
  }
  set_error("calib: unsupported x_dtype %d", x_dtype);
  return B200Q_ERR_BAD_ARG;
}
