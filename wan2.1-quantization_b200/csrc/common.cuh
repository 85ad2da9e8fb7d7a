// Shared device/host helpers for libb200q (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/b200q.h"

namespace b200q {

// ---- error plumbing (thread-local message, status codes of include/b200q.h) ------------
void set_error(const char* fmt, ...);
void clear_error();

#define B200Q_REQUIRE(cond, code, ...)        \
  do {                                        \
    if (!(cond)) {                            \
      ::b200q::set_error(__VA_ARGS__);        \
      return (code);                          \
    }                                         \
  } while (0)

#define B200Q_CUDA_OK(expr)                                                          \
  do {                                                                               \
    cudaError_t _e = (expr);                                                         \
    if (_e != cudaSuccess) {                                                         \
      ::b200q::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),     \
                         __FILE__, __LINE__);                                        \
      return B200Q_ERR_CUDA;                                                         \
    }                                                                                \
  } while (0)

// every launch is followed by this (the reference never checks cudaGetLastError)
#define B200Q_CHECK_LAUNCH() B200Q_CUDA_OK(cudaGetLastError())

int sm_count();   // cached multiProcessorCount of the current device (148 on B200)

static inline bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

// Row-in-registers kernels (quantizer, LN+quant, RMSNorm+RoPE): a row of `kv` 16-byte vectors is owned by W warps, each
// thread holding V vectors.  Per-warp fixed work (reductions, row parameters) is amortised over V*32 vectors, so pick
// the FEWEST warps whose V fits the register budget; W == 1 needs no block barrier at all.
struct RowLayout { int W, V, threads; };
static inline RowLayout pick_row_layout(int kv, int vmax) {
  RowLayout r{1, 0, 256};
  while (r.W < 8 && (kv + 32 * r.W - 1) / (32 * r.W) > vmax) r.W *= 2;
  if ((kv + 32 * r.W - 1) / (32 * r.W) > vmax) { r.W = 32; r.threads = 1024; }
  r.V = (kv + 32 * r.W - 1) / (32 * r.W);
  return r;
}
// dispatch a runtime V onto the instantiated set {1,2,3,4,6,8,10,12}
#define B200Q_DISPATCH_V(V, CALL)                 \
  do {                                            \
    if ((V) <= 1) { CALL(1); }                    \
    else if ((V) <= 2) { CALL(2); }               \
    else if ((V) <= 3) { CALL(3); }               \
    else if ((V) <= 4) { CALL(4); }               \
    else if ((V) <= 6) { CALL(6); }               \
    else if ((V) <= 8) { CALL(8); }               \
    else if ((V) <= 10) { CALL(10); }             \
    else { CALL(12); }                            \
  } while (0)

// ---- dtype helpers ------------------------------------------------------------------------
template <typename T> struct DType;
template <> struct DType<float>         { static constexpr int id = B200Q_F32;  static constexpr int vec = 4; };
template <> struct DType<__nv_bfloat16> { static constexpr int id = B200Q_BF16; static constexpr int vec = 8; };
template <> struct DType<__half>        { static constexpr int id = B200Q_F16;  static constexpr int vec = 8; };

__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }
__device__ __forceinline__ float to_f32(__half v) { return __half2float(v); }

template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }

// 16-byte streaming global load/store (read-once data: bypass L1 allocation)
__device__ __forceinline__ uint4 ldg_stream16(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_stream16(void* p, uint4 v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void stg_stream8(void* p, uint2 v) {
  asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1,%2};" :: "l"(p), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ void stg_stream4(void* p, uint32_t v) {
  asm volatile("st.global.L1::no_allocate.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}

// ---- packed fp32x2 arithmetic (FFMA2 / FMUL2 / FADD2: two fp32 lanes per issue slot on sm_100) -------
__device__ __forceinline__ uint64_t pack_f32x2(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ uint64_t pack_u32x2(uint32_t a, uint32_t b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(a), "r"(b));
  return r;
}
__device__ __forceinline__ void unpack_u32x2(uint64_t v, uint32_t& a, uint32_t& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=r"(a), "=r"(b) : "l"(v));
}
__device__ __forceinline__ void unpack_f32x2(uint64_t v, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ uint64_t fma_f32x2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t mul_f32x2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t add_f32x2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
// div_rn_hoisted on two lanes, then + 1.5*2^23: returns the two "integer in the mantissa" bit patterns.
//   nd2 = (-d,-d), r2 = (r,r), magic2 = (12582912, 12582912)
__device__ __forceinline__ uint64_t div_rn_hoisted_rne2(uint64_t x2, uint64_t nd2, uint64_t r2, uint64_t magic2) {
  uint64_t q = mul_f32x2(x2, r2);
  uint64_t e = fma_f32x2(q, nd2, x2);
  q = fma_f32x2(e, r2, q);
  e = fma_f32x2(q, nd2, x2);
  q = fma_f32x2(e, r2, q);
  return add_f32x2(q, magic2);
}

// Unpack one 16-byte vector of T into fp32 lanes.
template <typename T> struct Vec16;
template <> struct Vec16<float> {
  static constexpr int N = 4;
  __device__ static __forceinline__ void unpack_pairs(const uint4& v, uint64_t* p) {
    p[0] = pack_u32x2(v.x, v.y); p[1] = pack_u32x2(v.z, v.w);
  }
  // running statistics: amax (sym) or max/min (asym), kept in fp32
  struct Stat { float a, b; };
  __device__ static __forceinline__ Stat stat_init() { return {0.f, 0.f}; }
  template <bool SYM> __device__ static __forceinline__ void stat_update(Stat& s, const uint4& v) {
    const float f0 = __uint_as_float(v.x), f1 = __uint_as_float(v.y), f2 = __uint_as_float(v.z), f3 = __uint_as_float(v.w);
    if (SYM) {
      s.a = fmaxf(fmaxf(fabsf(f0), fabsf(f1)), s.a);
      s.a = fmaxf(fmaxf(fabsf(f2), fabsf(f3)), s.a);
    } else {
      s.a = fmaxf(fmaxf(f0, f1), s.a); s.a = fmaxf(fmaxf(f2, f3), s.a);
      s.b = fminf(fminf(f0, f1), s.b); s.b = fminf(fminf(f2, f3), s.b);
    }
  }
  __device__ static __forceinline__ void stat_final(const Stat& s, float& a, float& b) { a = s.a; b = s.b; }
  __device__ static __forceinline__ void unpack(const uint4& v, float* f) {
    f[0] = __uint_as_float(v.x); f[1] = __uint_as_float(v.y);
    f[2] = __uint_as_float(v.z); f[3] = __uint_as_float(v.w);
  }
};
template <> struct Vec16<__nv_bfloat16> {
  static constexpr int N = 8;
  __device__ static __forceinline__ void unpack_pairs(const uint4& v, uint64_t* p) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) p[i] = pack_u32x2(w[i] << 16, w[i] & 0xffff0000u);
  }
  // statistics stay packed in bf16x2 (max/min of bf16 values is exact), one 3-input VHMNMX per 4 elements
  struct Stat { __nv_bfloat162 a, b; };
  __device__ static __forceinline__ Stat stat_init() {
    Stat s; s.a = __float2bfloat162_rn(0.f); s.b = __float2bfloat162_rn(0.f); return s;
  }
  template <bool SYM> __device__ static __forceinline__ void stat_update(Stat& s, const uint4& v) {
    uint32_t w[4] = {v.x, v.y, v.z, v.w};
    if (SYM) {
#pragma unroll
      for (int i = 0; i < 4; ++i) w[i] &= 0x7fff7fffu;
    }
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(w);
    s.a = __hmax2(__hmax2(h[0], h[1]), s.a);
    s.a = __hmax2(__hmax2(h[2], h[3]), s.a);
    if (!SYM) {
      s.b = __hmin2(__hmin2(h[0], h[1]), s.b);
      s.b = __hmin2(__hmin2(h[2], h[3]), s.b);
    }
  }
  __device__ static __forceinline__ void stat_final(const Stat& s, float& a, float& b) {
    a = fmaxf(__low2float(s.a), __high2float(s.a));
    b = fminf(__low2float(s.b), __high2float(s.b));
  }
  __device__ static __forceinline__ void unpack(const uint4& v, float* f) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {           // bf16 -> fp32 is a 16-bit left shift
      f[2 * i]     = __uint_as_float(w[i] << 16);
      f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
};
template <> struct Vec16<__half> {
  static constexpr int N = 8;
  __device__ static __forceinline__ void unpack_pairs(const uint4& v, uint64_t* p) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
      p[i] = pack_f32x2(t.x, t.y);
    }
  }
  struct Stat { __half2 a, b; };
  __device__ static __forceinline__ Stat stat_init() {
    Stat s; s.a = __float2half2_rn(0.f); s.b = __float2half2_rn(0.f); return s;
  }
  template <bool SYM> __device__ static __forceinline__ void stat_update(Stat& s, const uint4& v) {
    uint32_t w[4] = {v.x, v.y, v.z, v.w};
    if (SYM) {
#pragma unroll
      for (int i = 0; i < 4; ++i) w[i] &= 0x7fff7fffu;
    }
    const __half2* h = reinterpret_cast<const __half2*>(w);
    s.a = __hmax2(__hmax2(h[0], h[1]), s.a);
    s.a = __hmax2(__hmax2(h[2], h[3]), s.a);
    if (!SYM) {
      s.b = __hmin2(__hmin2(h[0], h[1]), s.b);
      s.b = __hmin2(__hmin2(h[2], h[3]), s.b);
    }
  }
  __device__ static __forceinline__ void stat_final(const Stat& s, float& a, float& b) {
    a = fmaxf(__low2float(s.a), __high2float(s.a));
    b = fminf(__low2float(s.b), __high2float(s.b));
  }
  __device__ static __forceinline__ void unpack(const uint4& v, float* f) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 t = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
      f[2 * i] = t.x; f[2 * i + 1] = t.y;
    }
  }
};

// ---- exact quotient with a row-shared divisor ----------------------------------------------
// The reference computes torch.round(x / delta) (base_quantizer.py:155): a correctly rounded
// IEEE fp32 division followed by round-half-even.  With r = RN(1/d) hoisted per row, two
// Markstein correction steps (exact fma remainders) return RN(x/d) bit-for-bit; this is what
// div.rn.f32 does internally, minus the per-element reciprocal.  Verified against true
// division on 2.56e9 random and near-tie cases (oracle/c/quant_oracle.c:div_hoisted_check, run by tests/test_oracle_golden.py).
__device__ __forceinline__ float div_rn_hoisted(float x, float d, float r) {
  float q = x * r;
  float e = fmaf(-q, d, x);
  q = fmaf(e, r, q);
  e = fmaf(-q, d, x);
  return fmaf(e, r, q);
}

// round-half-even of |v| < 2^22 to an integer, returned in the low bits of the float:
// (v + 1.5*2^23) has the integer in its mantissa, two's complement in the low byte.
__device__ __forceinline__ int rne_to_int_bits(float v) { return __float_as_int(v + 12582912.0f); }
__device__ __forceinline__ int rne_to_int(float v) { return rne_to_int_bits(v) - 0x4B400000; }

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace b200q
