// (b) Quantized linear: int8 x int8 -> int32 on the 5th-gen tensor cores (tcgen05.mma.kind::i8,
// accumulators in TMEM), operands streamed by TMA, fused dequant / zero-point / bias /
// (GELU | gate-residual) epilogue, TMA store.
//
// Replaces F.linear on two dequantised operands (ViDiT-Q/quant_utils/qdiff/base/quant_layer.py:70)
// and the reference's Ampere mma.sync kernels (ViDiT-Q/kernels/csrc/qgemm/w8a8/w8a8_gemm_cuda.cu:14-622,
// epilogue :416-441;  w4a8/w4a8_per_channel_gemm_cuda_qserve.cu:304-597).
//
//   out[m,n] = epi( da[m]*dw[n] * ( sum_k qa[m,k]*qw[n,k] + zp_w[n]*rowsum_a[m] ) + bias[n] )
//
// Kernel shape (one persistent CTA per SM, 256 threads, warp-specialised):
//   warp 0      TMA producer: A tile [128 x 128B], B tile [256 x 128B] per K-block into a 4-deep
//               smem ring (SWIZZLE_128B), mbarrier expect_tx/complete_tx
//   warp 1      MMA issuer: one elected lane, 4 x tcgen05.mma (M128,N256,K32) per K-block,
//               tcgen05.commit -> frees the smem slot / publishes the accumulator
//   warp 2      TMEM allocator (512 columns = two 128x256 int32 accumulators, double-buffered so
//               the epilogue of tile i overlaps the main loop of tile i+1)
//   warps 4-7   epilogue: tcgen05.ld 32 lanes x 32 columns, int zero-point fix-up, fp32 scale/bias,
//               activation, pack, swizzled smem staging, TMA store (clips ragged M/N tails)
// Ragged M, N, K need no host padding: TMA zero-fills out-of-bounds loads and clips stores.
#include "common.cuh"
#include "ptx.cuh"
#include <type_traits>

namespace b200q {
using namespace ptx;

constexpr int BM = 128, BN = 256, BK = 128;        // BK in bytes == int8 elements: one 128B swizzle row
constexpr int UMMA_K = 32;                         // kind::i8: 32 bytes of K per instruction
constexpr int STAGES = 4;
constexpr int A_BYTES = BM * BK, B_BYTES = BN * BK, STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int EPI_WARPS = 4;
constexpr int STAGING_BYTES = 32 * 128;            // one warp: 32 rows x 128 B (swizzled)
constexpr int GEMM_THREADS = 256;
constexpr int TMEM_COLS = 512;

struct GemmSmem {
  // offsets into dynamic smem (base aligned to 1024 B)
  static constexpr int ring = 0;
  static constexpr int staging = STAGES * STAGE_BYTES;                       // EPI_WARPS x 2 x 4 KB
  static constexpr int colparams = staging + EPI_WARPS * 2 * STAGING_BYTES;   // dw[BN] f32, bias[BN] f32, zp[BN] i16
  static constexpr int barriers = colparams + BN * 4 + BN * 4 + BN * 2;
  static constexpr int total = barriers + 256;
};
static_assert(GemmSmem::total <= 232448, "dynamic smem budget (227 KB) exceeded");

struct GemmParams {
  int M, N, K;
  const float* delta_a;
  const float* delta_w;
  const float* zp_w;        // may be null
  const int32_t* rowsum_a;  // may be null
  const void* bias;         // may be null
  int bias_dtype;
  const float* gate;        // EPI_GATE_RESIDUAL
  int epilogue;
  int zp_offset;            // added to zp_w (W4A8: -8, the unsigned-nibble bias of b200q_pack_w4)
};

__device__ __forceinline__ float load_bias(const void* bias, int dtype, int n) {
  if (dtype == B200Q_F32) return reinterpret_cast<const float*>(bias)[n];
  if (dtype == B200Q_BF16) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(bias)[n]);
  return __half2float(reinterpret_cast<const __half*>(bias)[n]);
}

__device__ __forceinline__ float gelu_tanh(float x) {
  // torch GELU(approximate='tanh') (wan/modules/model.py:287); tanh on the MUFU (tanh.approx.f32)
  const float u = 0.7978845608028654f * fmaf(0.044715f * x * x, x, x);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
  return 0.5f * x * (1.f + t);
}

template <typename OutT> struct OutPack;    // 16-byte chunk = ELEMS outputs
template <> struct OutPack<__nv_bfloat16> { static constexpr int ELEMS = 8; };
template <> struct OutPack<__half>        { static constexpr int ELEMS = 8; };
template <> struct OutPack<float>         { static constexpr int ELEMS = 4; };
template <> struct OutPack<int32_t>       { static constexpr int ELEMS = 4; };

template <typename OutT>
__device__ __forceinline__ uint4 pack_chunk(const float* y) {
  uint4 r;
  if constexpr (sizeof(OutT) == 4) {
    r.x = __float_as_uint(y[0]); r.y = __float_as_uint(y[1]); r.z = __float_as_uint(y[2]); r.w = __float_as_uint(y[3]);
  } else if constexpr (std::is_same<OutT, __nv_bfloat16>::value) {
    __nv_bfloat162 a = __floats2bfloat162_rn(y[0], y[1]), b = __floats2bfloat162_rn(y[2], y[3]);
    __nv_bfloat162 c = __floats2bfloat162_rn(y[4], y[5]), d = __floats2bfloat162_rn(y[6], y[7]);
    r.x = *reinterpret_cast<uint32_t*>(&a); r.y = *reinterpret_cast<uint32_t*>(&b);
    r.z = *reinterpret_cast<uint32_t*>(&c); r.w = *reinterpret_cast<uint32_t*>(&d);
  } else {
    __half2 a = __floats2half2_rn(y[0], y[1]), b = __floats2half2_rn(y[2], y[3]);
    __half2 c = __floats2half2_rn(y[4], y[5]), d = __floats2half2_rn(y[6], y[7]);
    r.x = *reinterpret_cast<uint32_t*>(&a); r.y = *reinterpret_cast<uint32_t*>(&b);
    r.z = *reinterpret_cast<uint32_t*>(&c); r.w = *reinterpret_cast<uint32_t*>(&d);
  }
  return r;
}

// OutT = int32_t -> raw accumulators.  EPI: b200q_epilogue.
template <typename OutT, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_w8a8_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                 const __grid_constant__ CUtensorMap tm_out, const __grid_constant__ CUtensorMap tm_res,
                 const GemmParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];          // SWIZZLE_128B tiles need 1024-byte alignment
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) {
    printf("b200q: dynamic smem base not 1024-byte aligned\n");
    __trap();
  }
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + GemmSmem::barriers);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint64_t* res_bar = tmem_empty_bar + 2;                      // EPI_WARPS barriers: residual tile landed
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(res_bar + EPI_WARPS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_n = (p.N + BN - 1) / BN, tiles_m = (p.M + BM - 1) / BM;
  const int num_tiles = tiles_m * tiles_n;
  const int num_kb = (p.K + BK - 1) / BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full_bar[s], 1); mbar_init(&tmem_empty_bar[s], EPI_WARPS); }
    for (int s = 0; s < EPI_WARPS; ++s) mbar_init(&res_bar[s], 1);
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tm_a); prefetch_tmap(&tm_b); prefetch_tmap(&tm_out);
    if (EPI == B200Q_EPI_GATE_RESIDUAL) prefetch_tmap(&tm_res);
  }
  if (warp == 2) tmem_alloc<TMEM_COLS>(tmem_base_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + GemmSmem::ring + stage * STAGE_BYTES;
          mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
          tma_load_2d(sa, &tm_a, &full_bar[stage], kb * BK, m0);
          tma_load_2d(sa + A_BYTES, &tm_b, &full_bar[stage], kb * BK, n0);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_i8_idesc(BM, BN);
      int stage = 0; uint32_t phase = 0; int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int as = it & 1; const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(&tmem_empty_bar[as], aphase ^ 1);          // epilogue has drained this accumulator
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + as * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          const uint32_t sa = smem_u32(smem + GemmSmem::ring + stage * STAGE_BYTES);
          const uint64_t adesc = make_kmajor_sw128_desc(sa);
          const uint64_t bdesc = make_kmajor_sw128_desc(sa + A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // advance both descriptors by k*32 bytes inside the 128B swizzle row (address field is >>4)
            mma_i8_ss(tmem_d, adesc + (uint64_t)(k * (UMMA_K >> 4)), bdesc + (uint64_t)(k * (UMMA_K >> 4)), idesc,
                      (kb | k) != 0 ? 1u : 0u);
          }
          mma_commit(&empty_bar[stage]);                     // slot free once these MMAs have read it
          if (kb == num_kb - 1) mma_commit(&tmem_full_bar[as]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int ew = warp - 4;                                 // == warp % 4: TMEM lane quarter
    const int etid = threadIdx.x - 128;
    float* s_dw = reinterpret_cast<float*>(smem + GemmSmem::colparams);
    float* s_bias = s_dw + BN;
    int16_t* s_zp = reinterpret_cast<int16_t*>(s_bias + BN);
    uint8_t* my_staging = smem + GemmSmem::staging + ew * 2 * STAGING_BYTES;
    constexpr int ELEMS = OutPack<OutT>::ELEMS;              // outputs per 16 B
    constexpr int CPS = 8 * ELEMS;                           // columns per 128B staging row / TMA store box
    constexpr bool RAW = std::is_same<OutT, int32_t>::value;
    uint32_t res_phase = 0;
    int it = 0; int sbuf = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * BN;
      const int as = it & 1; const uint32_t aphase = (it >> 1) & 1;
      if (!RAW) {
        named_bar_sync(1, EPI_WARPS * 32);                   // everyone finished reading the previous tile's params
        for (int i = etid; i < BN; i += EPI_WARPS * 32) {
          const int n = n0 + i; const bool ok = n < p.N;
          float dw = ok ? p.delta_w[n] : 0.f;
          float bs = (ok && p.bias) ? load_bias(p.bias, p.bias_dtype, n) : 0.f;
          if (EPI == B200Q_EPI_GATE_RESIDUAL && ok && p.gate) {   // (y*dw + b)*g == y*(dw*g) + b*g: fold the gate in
            const float g = p.gate[n];
            dw *= g; bs *= g;
          }
          s_dw[i] = dw;
          s_bias[i] = bs;
          s_zp[i] = (int16_t)((ok && p.zp_w) ? __float2int_rn(p.zp_w[n]) + p.zp_offset : p.zp_offset);
        }
        named_bar_sync(1, EPI_WARPS * 32);
      }
      const int row = m0 + ew * 32 + lane;
      const float da = (!RAW && row < p.M) ? p.delta_a[row] : 0.f;
      const int rs = (!RAW && p.rowsum_a && row < p.M) ? p.rowsum_a[row] : 0;   // only read when a zero point is in play

      mbar_wait(&tmem_full_bar[as], aphase);
      tcgen05_fence_after();
      const uint32_t t_row = tmem_base + as * BN + ((uint32_t)(ew * 32) << 16);

      for (int u = 0; u < BN / CPS; ++u) {
        uint8_t* sbuf_ptr = my_staging + sbuf * STAGING_BYTES;
        // the TMA store that last read this staging buffer must have finished reading it
        if (lane == 0) tma_store_wait_read<1>();
        __syncwarp();
        const bool chunk_live = (n0 + u * CPS < p.N) && (m0 + ew * 32 < p.M);
        if (EPI == B200Q_EPI_GATE_RESIDUAL && chunk_live) {
          // pull the fp32 residual box [32 rows x CPS cols] into the staging buffer first
          if (lane == 0) {
            mbar_expect_tx(&res_bar[ew], STAGING_BYTES);
            tma_load_2d(sbuf_ptr, &tm_res, &res_bar[ew], n0 + u * CPS, m0 + ew * 32);
          }
        }
#pragma unroll
        for (int h = 0; h < CPS / 32; ++h) {
          uint32_t v[32];
          tmem_ld_32x32(t_row + u * CPS + h * 32, v);
          tmem_ld_wait();
          if (u == BN / CPS - 1 && h == CPS / 32 - 1) {      // accumulator fully read: hand TMEM back to the MMA warp
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty_bar[as]);
          }
          if (EPI == B200Q_EPI_GATE_RESIDUAL && chunk_live && h == 0) {
            mbar_wait(&res_bar[ew], res_phase);
            res_phase ^= 1;
          }
#pragma unroll
          for (int c = 0; c < 32 / ELEMS; ++c) {             // 16-byte chunks of this thread's row
            const int col = u * CPS + h * 32 + c * ELEMS;    // column inside the tile
            const int j = (h * 32) / ELEMS + c;              // chunk index inside the 128B staging row
            uint4* dst = reinterpret_cast<uint4*>(sbuf_ptr + lane * 128 + ((j ^ (lane & 7)) << 4));
            uint4 o;
            if (RAW) {
              o = make_uint4(v[c * 4], v[c * 4 + 1], v[c * 4 + 2], v[c * 4 + 3]);
            } else {
              float y[ELEMS];
#pragma unroll
              for (int e = 0; e < ELEMS; ++e) {
                const int acc = (int)v[c * ELEMS + e] + (int)s_zp[col + e] * rs;
                float t = fmaf((float)acc, da * s_dw[col + e], s_bias[col + e]);
                if (EPI == B200Q_EPI_GELU_TANH) t = gelu_tanh(t);
                y[e] = t;
              }
              if (EPI == B200Q_EPI_GATE_RESIDUAL) {
                const uint4 r4 = *dst;                        // residual (fp32 x4) sits where the output goes
                y[0] += __uint_as_float(r4.x);
                y[1] += __uint_as_float(r4.y);
                y[2] += __uint_as_float(r4.z);
                y[3] += __uint_as_float(r4.w);
              }
              o = pack_chunk<OutT>(y);
            }
            *dst = o;
          }
        }
        fence_proxy_async_smem();                            // generic-proxy smem writes -> visible to TMA
        __syncwarp();
        if (lane == 0) {
          if (chunk_live) tma_store_2d(&tm_out, sbuf_ptr, n0 + u * CPS, m0 + ew * 32);
          tma_store_commit();
        }
        sbuf ^= 1;
      }
    }
    if (lane == 0) tma_store_wait<0>();
  }

  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  if (warp == 2) tmem_dealloc<TMEM_COLS>(tmem_base);
}

// ---- host side ----------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}

// 2-D row-major tensor [rows, cols] with leading dimension ld (elements); box = [box_rows, box_cols]
int make_tmap_2d(CUtensorMap* out, const void* base, CUtensorMapDataType dt, int elem_bytes, int64_t rows, int64_t cols,
                 int64_t ld, int box_rows, int box_cols, CUtensorMapSwizzle swz) {
  EncodeTiledFn fn = get_encode_fn();
  B200Q_REQUIRE(fn != nullptr, B200Q_ERR_CUDA, "cuTensorMapEncodeTiled entry point not found (driver too old?)");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * elem_bytes};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  B200Q_REQUIRE(r == CUDA_SUCCESS, B200Q_ERR_CUDA,
                "cuTensorMapEncodeTiled failed (%d): base=%p rows=%lld cols=%lld ld=%lld elem=%d box=[%d,%d]", (int)r, base,
                (long long)rows, (long long)cols, (long long)ld, elem_bytes, box_rows, box_cols);
  return B200Q_OK;
}

template <typename OutT, int EPI>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, const CUtensorMap& tr,
                       const GemmParams& p, cudaStream_t st) {
  auto kern = gemm_w8a8_kernel<OutT, EPI>;
  static bool configured = false;    // cudaFuncSetAttribute once per instantiation, not per call
  const int smem_bytes = GemmSmem::total;
  if (!configured) {
    B200Q_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    configured = true;
  }
  const int tiles = ((p.M + BM - 1) / BM) * ((p.N + BN - 1) / BN);
  const int grid = tiles < sm_count() ? tiles : sm_count();
  kern<<<grid, GEMM_THREADS, smem_bytes, st>>>(ta, tb, to, tr, p);
  B200Q_CHECK_LAUNCH();
  return B200Q_OK;
}

template <typename OutT>
static int dispatch_epi(int epi, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to,
                        const CUtensorMap& tr, const GemmParams& p, cudaStream_t st) {
  switch (epi) {
    case B200Q_EPI_NONE: return launch_gemm<OutT, B200Q_EPI_NONE>(ta, tb, to, tr, p, st);
    case B200Q_EPI_GELU_TANH: return launch_gemm<OutT, B200Q_EPI_GELU_TANH>(ta, tb, to, tr, p, st);
  }
  set_error("gemm: unsupported epilogue %d for this out_dtype", epi);
  return B200Q_ERR_BAD_ARG;
}

int gemm_i8_common(const int8_t* qa, int64_t lda, const int8_t* qw, int64_t ldw, const float* delta_a,
                   const float* delta_w, const float* zp_w, const int32_t* rowsum_a, const void* bias, int bias_dtype,
                   void* out, int out_dtype, int64_t ldo, int64_t M, int64_t N, int64_t K, int epilogue,
                   const float* residual, int64_t ldr, const float* gate, cudaStream_t st) {
  B200Q_REQUIRE(M >= 0 && N >= 0 && K >= 0, B200Q_ERR_BAD_ARG, "gemm: negative shape");
  if (M == 0 || N == 0) return B200Q_OK;
  B200Q_REQUIRE(K > 0, B200Q_ERR_BAD_ARG, "gemm: K must be > 0");
  B200Q_REQUIRE(M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31), B200Q_ERR_UNSUPPORTED, "gemm: dimension >= 2^31");
  B200Q_REQUIRE(qa && qw && out, B200Q_ERR_BAD_ARG, "gemm: null operand pointer");
  B200Q_REQUIRE(lda >= K && ldw >= K && ldo >= N, B200Q_ERR_BAD_ARG, "gemm: leading dimension too small");
  B200Q_REQUIRE(lda % 16 == 0 && ldw % 16 == 0 && aligned(qa, 16) && aligned(qw, 16), B200Q_ERR_BAD_ARG,
                "gemm: qa/qw must be 16-byte aligned with lda, ldw multiples of 16 (TMA global-stride rule)");
  const bool raw = out_dtype == B200Q_I32;
  if (!raw) {
    B200Q_REQUIRE(delta_a && delta_w, B200Q_ERR_BAD_ARG, "gemm: delta_a / delta_w required for dequantised output");
    B200Q_REQUIRE(!zp_w || rowsum_a, B200Q_ERR_BAD_ARG, "gemm: rowsum_a required when zp_w is given");
    B200Q_REQUIRE(!bias || (bias_dtype >= B200Q_F32 && bias_dtype <= B200Q_F16), B200Q_ERR_BAD_ARG, "gemm: bad bias_dtype");
  }
  const int osz = (out_dtype == B200Q_BF16 || out_dtype == B200Q_F16) ? 2 : 4;
  B200Q_REQUIRE(out_dtype >= B200Q_F32 && out_dtype <= B200Q_I32, B200Q_ERR_BAD_ARG, "gemm: bad out_dtype %d", out_dtype);
  B200Q_REQUIRE(aligned(out, 16) && (ldo * osz) % 16 == 0, B200Q_ERR_BAD_ARG,
                "gemm: out must be 16-byte aligned with a 16-byte-multiple row pitch");

  CUtensorMap ta, tb, to, tr;
  int rc;
  if ((rc = make_tmap_2d(&ta, qa, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, M, K, lda, BM, BK, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  if ((rc = make_tmap_2d(&tb, qw, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, N, K, ldw, BN, BK, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  CUtensorMapDataType odt = out_dtype == B200Q_BF16  ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                            : out_dtype == B200Q_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16
                            : out_dtype == B200Q_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                                     : CU_TENSOR_MAP_DATA_TYPE_INT32;
  if ((rc = make_tmap_2d(&to, out, odt, osz, M, N, ldo, 32, 128 / osz, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  tr = to;

  GemmParams p{};
  p.M = (int)M; p.N = (int)N; p.K = (int)K;
  p.delta_a = delta_a; p.delta_w = delta_w; p.zp_w = zp_w; p.rowsum_a = rowsum_a;
  p.bias = bias; p.bias_dtype = bias_dtype; p.gate = gate; p.epilogue = epilogue;

  if (raw) return launch_gemm<int32_t, B200Q_EPI_NONE>(ta, tb, to, tr, p, st);
  if (epilogue == B200Q_EPI_GATE_RESIDUAL) {
    B200Q_REQUIRE(out_dtype == B200Q_F32, B200Q_ERR_BAD_ARG, "gemm: gate-residual epilogue writes the fp32 residual stream");
    B200Q_REQUIRE(residual != nullptr, B200Q_ERR_BAD_ARG, "gemm: residual required");
    B200Q_REQUIRE(ldr >= N && aligned(residual, 16) && (ldr * 4) % 16 == 0, B200Q_ERR_BAD_ARG, "gemm: bad residual layout");
    if ((rc = make_tmap_2d(&tr, residual, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, M, N, ldr, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    return launch_gemm<float, B200Q_EPI_GATE_RESIDUAL>(ta, tb, to, tr, p, st);
  }
  switch (out_dtype) {
    case B200Q_BF16: return dispatch_epi<__nv_bfloat16>(epilogue, ta, tb, to, tr, p, st);
    case B200Q_F16: return dispatch_epi<__half>(epilogue, ta, tb, to, tr, p, st);
    case B200Q_F32: return dispatch_epi<float>(epilogue, ta, tb, to, tr, p, st);
  }
  return B200Q_ERR_BAD_ARG;
}

}  // namespace b200q

using namespace b200q;

extern "C" int b200q_gemm_w8a8(const int8_t* qa, int64_t lda, const int8_t* qw, int64_t ldw, const float* delta_a,
                               const float* delta_w, const float* zp_w, const int32_t* rowsum_a, const void* bias,
                               int bias_dtype, void* out, int out_dtype, int64_t ldo, int64_t M, int64_t N, int64_t K,
                               int epilogue, const float* residual, int64_t ldr, const float* gate,
                               b200q_stream_t stream) {
  clear_error();
  return gemm_i8_common(qa, lda, qw, ldw, delta_a, delta_w, zp_w, rowsum_a, bias, bias_dtype, out, out_dtype, ldo, M, N,
                        K, epilogue, residual, ldr, gate, (cudaStream_t)stream);
}

namespace b200q {
int gemm_w4a8_impl(const int8_t*, int64_t, const uint8_t*, int64_t, const float*, const float*, const float*,
                   const int32_t*, const void*, int, void*, int, int64_t, int64_t, int64_t, int64_t, int, const float*,
                   int64_t, const float*, cudaStream_t) {
  set_error("gemm_w4a8: not built yet");
  return B200Q_ERR_UNSUPPORTED;
}
}  // namespace b200q
