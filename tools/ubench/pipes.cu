// Issue-rate microbenchmark of the instructions the attention softmax is made of (sm_100a): which pipe does each one use
// and which of them share one?  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu && ./pipes
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(512) k(float* out, int iters, float seed) {
  float a0 = seed + threadIdx.x, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f, a4 = a0 * .5f, a5 = a0 * .25f, a6 = a0 * .125f, a7 = a0 * .0625f;
  uint32_t u0 = threadIdx.x, u1 = u0 + 1, u2 = u0 + 2, u3 = u0 + 3;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      if (MODE == 0 || MODE == 3 || MODE == 5) {            // MUFU.EX2 x8
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a0)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a1));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a2)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a3));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a4)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a5));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a6)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a7));
      }
      if (MODE == 1 || MODE == 3) {                         // F2FP.BF16 pack x4
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u0) : "f"(a0), "f"(a1)); asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u1) : "f"(a2), "f"(a3));
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u2) : "f"(a4), "f"(a5)); asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u3) : "f"(a6), "f"(a7));
        a0 += __uint_as_float(u0 & 1); a2 += __uint_as_float(u1 & 1); a4 += __uint_as_float(u2 & 1); a6 += __uint_as_float(u3 & 1);
      }
      if (MODE == 2) {                                      // fma.rn.f32x2 x4
        uint64_t p0, p1, p2, p3;
        asm volatile("mov.b64 %0, {%1, %2};" : "=l"(p0) : "f"(a0), "f"(a1)); asm volatile("mov.b64 %0, {%1, %2};" : "=l"(p1) : "f"(a2), "f"(a3));
        asm volatile("mov.b64 %0, {%1, %2};" : "=l"(p2) : "f"(a4), "f"(a5)); asm volatile("mov.b64 %0, {%1, %2};" : "=l"(p3) : "f"(a6), "f"(a7));
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          asm volatile("fma.rn.f32x2 %0, %0, %0, %0;" : "+l"(p0)); asm volatile("fma.rn.f32x2 %0, %0, %0, %0;" : "+l"(p1));
          asm volatile("fma.rn.f32x2 %0, %0, %0, %0;" : "+l"(p2)); asm volatile("fma.rn.f32x2 %0, %0, %0, %0;" : "+l"(p3));
        }
        asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(a0), "=f"(a1) : "l"(p0)); asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(a2), "=f"(a3) : "l"(p1));
        asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(a4), "=f"(a5) : "l"(p2)); asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(a6), "=f"(a7) : "l"(p3));
      }
      if (MODE == 4 || MODE == 5) {                         // PRMT pack x4 (+ IADD rounding x8 in mode 6)
        asm volatile("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(u0) : "r"(__float_as_uint(a0)), "r"(__float_as_uint(a1)));
        asm volatile("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(u1) : "r"(__float_as_uint(a2)), "r"(__float_as_uint(a3)));
        asm volatile("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(u2) : "r"(__float_as_uint(a4)), "r"(__float_as_uint(a5)));
        asm volatile("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(u3) : "r"(__float_as_uint(a6)), "r"(__float_as_uint(a7)));
        a0 += __uint_as_float(u0 & 1); a2 += __uint_as_float(u1 & 1); a4 += __uint_as_float(u2 & 1); a6 += __uint_as_float(u3 & 1);
      }
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + __uint_as_float(u0 ^ u1 ^ u2 ^ u3);
}

template <int MODE>
void run(const char* name, double ops_per_iter) {
  float* out; cudaMalloc(&out, 148 * 4 * 512 * 4);
  const int iters = 4000;
  cudaEvent_t s, e; cudaEventCreate(&s); cudaEventCreate(&e);
  k<MODE><<<148, 512>>>(out, 100, 0.001f);
  cudaEventRecord(s);
  k<MODE><<<148, 512>>>(out, iters, 0.001f);
  cudaEventRecord(e); cudaEventSynchronize(e);
  float ms; cudaEventElapsedTime(&ms, s, e);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  // thread-ops per SM per ns
  const double ops = ops_per_iter * 8 * iters * 512.0;      // per SM
  printf("%-44s %8.3f ms  %7.2f thread-ops/ns/SM  (= %.1f per clk at 1.9 GHz)\n", name, ms, ops / (ms * 1e6), ops / (ms * 1e6) / 1.9);
  cudaFree(out);
}

int main() {
  run<0>("ex2.approx.ftz.f32", 8);
  run<1>("cvt.rn.bf16x2.f32 (+dep ops)", 4);
  run<2>("fma.rn.f32x2", 16);
  run<3>("ex2 x8 + cvt.bf16x2 x4 (ops = ex2)", 8);
  run<4>("prmt pack (+dep ops)", 4);
  run<5>("ex2 x8 + prmt x4 (ops = ex2)", 8);
  return 0;
}
