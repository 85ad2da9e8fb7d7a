// W4 weight packing for the W4A8 GEMM (mixed-precision configs: weight.n_bits=[4,8],
// ViDiT-Q/quant_utils/qdiff/base/mixed_precision_quantizer.py:56-125).
//
// Packed format (defined here, consumed by the in-smem unpacker of gemm_w4a8): K is split in groups of
// 8 codes; byte i (i = 0..3) of a group's 32-bit word holds  (code[i] + 8)  in bits 0-3 and
// (code[4+i] + 8) in bits 4-7.  The +8 bias makes nibbles unsigned so the unpack is two AND/SHIFT ops per
// 4 codes; the GEMM epilogue folds it back through the zero-point term (zp_eff = zp_w - 8), the same
// algebra the reference's QServe kernel uses (w4a8_per_channel_gemm_cuda_qserve.cu:290-297, 585-586).
#include "common.cuh"

namespace b200q {

__global__ void __launch_bounds__(256) pack_w4_kernel(const int8_t* __restrict__ codes, int64_t ld, int64_t N, int64_t K,
                                                       uint8_t* __restrict__ packed, int64_t ldp) {
  const int64_t groups = (K + 7) / 8;
  const int64_t total = N * groups;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = i / groups, g = i - n * groups;
    const int8_t* src = codes + n * ld + g * 8;
    uint32_t w = 0;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int64_t k0 = g * 8 + b, k1 = g * 8 + 4 + b;
      // saturate to [-8, 7]: the asymmetric quantizer can emit +8 on an exact double rounding tie (rne(xmax/d) and
      // rne(xmin/d) both rounding outward) - a 17th level that 4 bits cannot hold; without the clamp
      // the nibble would wrap to -8.  Same policy as the int8 storage of 8-bit codes (+128 -> +127).
      const uint32_t lo = k0 < K ? (uint32_t)(min(max((int)src[b], -8), 7) + 8) : 8u;       // pad codes are 0 -> nibble 8
      const uint32_t hi = k1 < K ? (uint32_t)(min(max((int)src[4 + b], -8), 7) + 8) : 8u;
      w |= (lo | (hi << 4)) << (8 * b);
    }
    *reinterpret_cast<uint32_t*>(packed + n * ldp + g * 4) = w;
  }
}

// Expansion of a packed weight matrix to one byte per code (the unsigned nibble, code + 8): thread = one 16-byte packed
// chunk = 32 codes -> two 16-byte stores.  HBM-bound and tiny next to the GEMM it feeds: 8960 x 1536 is 6.9 MB in,
// 13.8 MB out.
__global__ void __launch_bounds__(256) unpack_w4_kernel(const uint8_t* __restrict__ packed, int64_t ldp, int64_t N, int64_t chunks,
                                                         int8_t* __restrict__ out, int64_t ldu) {
  const int64_t total = N * chunks;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = i / chunks, c = i - n * chunks;
    const uint4 w = *reinterpret_cast<const uint4*>(packed + n * ldp + c * 16);
    uint4 o0, o1;
    o0.x = w.x & 0x0F0F0F0Fu; o0.y = (w.x >> 4) & 0x0F0F0F0Fu; o0.z = w.y & 0x0F0F0F0Fu; o0.w = (w.y >> 4) & 0x0F0F0F0Fu;
    o1.x = w.z & 0x0F0F0F0Fu; o1.y = (w.z >> 4) & 0x0F0F0F0Fu; o1.z = w.w & 0x0F0F0F0Fu; o1.w = (w.w >> 4) & 0x0F0F0F0Fu;
    uint4* dst = reinterpret_cast<uint4*>(out + n * ldu + c * 32);
    dst[0] = o0; dst[1] = o1;
  }
}

}  // namespace b200q

using namespace b200q;

extern "C" int b200q_pack_w4(const int8_t* codes, int64_t ld, int64_t N, int64_t K, uint8_t* packed, int64_t ldp,
                             b200q_stream_t stream) {
  clear_error();
  B200Q_REQUIRE(N >= 0 && K >= 0, B200Q_ERR_BAD_ARG, "pack_w4: negative shape");
  if (N == 0 || K == 0) return B200Q_OK;
  B200Q_REQUIRE(codes && packed, B200Q_ERR_BAD_ARG, "pack_w4: null pointer");
  B200Q_REQUIRE(ld >= K && ldp >= ((K + 7) / 8) * 4, B200Q_ERR_BAD_ARG, "pack_w4: leading dimension too small");
  B200Q_REQUIRE(ldp % 4 == 0 && aligned(packed, 4), B200Q_ERR_BAD_ARG, "pack_w4: packed must be 4-byte aligned, ldp % 4 == 0");
  const int64_t total = N * ((K + 7) / 8);
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  pack_w4_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(codes, ld, N, K, packed, ldp);
  B200Q_CHECK_LAUNCH();
  return B200Q_OK;
}

namespace b200q {
int gemm_w4a8_impl(const int8_t* qa, int64_t lda, const uint8_t* qw4, int64_t ldw4, const float* delta_a,
                   const float* delta_w, const float* zp_w, const int32_t* rowsum_a, const void* bias, int bias_dtype,
                   void* out, int out_dtype, int64_t ldo, int64_t M, int64_t N, int64_t K, int epilogue,
                   const float* residual, int64_t ldr, const float* gate, cudaStream_t st);
int gemm_w4a8_expanded_impl(const int8_t* qa, int64_t lda, const int8_t* qw_u8, int64_t ldu, const float* delta_a,
                            const float* delta_w, const float* zp_w, const int32_t* rowsum_a, const void* bias, int bias_dtype,
                            void* out, int out_dtype, int64_t ldo, int64_t M, int64_t N, int64_t K, int epilogue,
                            const float* residual, int64_t ldr, const float* gate, cudaStream_t st);
}

extern "C" int b200q_gemm_w4a8(const int8_t* qa, int64_t lda, const uint8_t* qw4, int64_t ldw4, const float* delta_a,
                               const float* delta_w, const float* zp_w, const int32_t* rowsum_a, const void* bias,
                               int bias_dtype, void* out, int out_dtype, int64_t ldo, int64_t M, int64_t N, int64_t K,
                               int epilogue, const float* residual, int64_t ldr, const float* gate, void* expand_ws,
                               b200q_stream_t stream) {
  clear_error();
  // Many-token GEMMs (the DiT: M = 32,760 ... 75,600): every one of the M/128 row tiles would expand the same weight tile
  // again in shared memory.  Expanding the matrix ONCE per call into caller scratch costs microseconds (N*K bytes written)
  // and lets the product run on the W8A8 kernel at its full rate; small M keeps the in-kernel converter.
  if (expand_ws != nullptr && M >= 1024 && qw4 != nullptr && N > 0 && K > 0) {
    const int64_t chunks = (K + 31) / 32, ldu = chunks * 32;
    B200Q_REQUIRE(aligned(expand_ws, 16) && aligned(qw4, 16) && ldw4 % 16 == 0 && ldw4 >= chunks * 16, B200Q_ERR_BAD_ARG,
                  "gemm_w4a8: expand_ws / qw4 must be 16-byte aligned and ldw4 a multiple of 16 covering ceil(K/32)*16 bytes");
    const int64_t total = N * chunks;
    int64_t blocks = (total + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    unpack_w4_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(qw4, ldw4, N, chunks, (int8_t*)expand_ws, ldu);
    B200Q_CHECK_LAUNCH();
    return gemm_w4a8_expanded_impl(qa, lda, (const int8_t*)expand_ws, ldu, delta_a, delta_w, zp_w, rowsum_a, bias, bias_dtype, out,
                                   out_dtype, ldo, M, N, K, epilogue, residual, ldr, gate, (cudaStream_t)stream);
  }
  return gemm_w4a8_impl(qa, lda, qw4, ldw4, delta_a, delta_w, zp_w, rowsum_a, bias, bias_dtype, out, out_dtype, ldo, M,
                        N, K, epilogue, residual, ldr, gate, (cudaStream_t)stream);
}
