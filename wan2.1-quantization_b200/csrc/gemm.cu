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
constexpr int A_BYTES = BM * BK, B_BYTES = BN * BK;
constexpr int CTRL_WARPS = 4;                      // TMA producer, MMA issuer, TMEM allocator, spare
constexpr int EPI_WARPS = 8;                       // 2 per TMEM lane quarter: each owns 32 rows x 128 columns of a tile
constexpr int CVT_WARPS = 4;                       // W4A8 only: int4 -> int8 unpackers
constexpr int UNPACK_BUFS = 2;
constexpr int STAGING_BYTES = 32 * 128;            // one warp: 32 rows x 128 B (swizzled)
constexpr int TMEM_COLS = 512;

// W8A8: ring stage = A tile (16 KB) + B tile (32 KB).
// W4A8: ring stage = A tile (16 KB) + PACKED B tile (256 rows x 64 B = 16 KB); the converter warps expand it into one
//       of two 32 KB SWIZZLE_128B int8 tiles that the MMA reads.
// The gate-residual epilogue (8 B/element of HBM traffic: memory-bound) trades one ring stage for a second staging
// buffer per epilogue warp so the residual box of chunk u+1 is in flight while chunk u is processed.
// PAIR: 0 = single CTA per tile; 1 = 2-CTA cluster, single-CTA MMAs, B tile multicast; 2 = 2-CTA cluster with
// tcgen05.mma.cta_group::2 (M = 256 across the pair): each CTA stages only HALF of the B tile, so the smem traffic per
// MMA (operand fill + operand read) drops by a third and two more ring stages fit.
template <bool W4, int NSTAGE, int PAIR>
struct GemmSmem {
  static constexpr bool mma2 = (PAIR == 2);
  // NSTAGE 4, or 3 for the short-K gate-residual GEMMs: their epilogue is bound by the LATENCY of the fp32 residual boxes
  // (one 4 KB box per warp in flight moves 4.7 MB over the chip per HBM round trip), so they trade ring stages for
  // staging buffers - two residual boxes in flight per warp with cta_group::2 (half-size B stages), one otherwise
  static constexpr bool deep = mma2 && !W4 && NSTAGE == 3;
  static constexpr int stages = mma2 ? (deep ? 4 : NSTAGE + 2) : NSTAGE;
  static constexpr int staging_bufs = (NSTAGE == 3) ? (deep ? 3 : 2) : 1;
  static constexpr int b_rows = mma2 ? BN / 2 : BN;            // B rows staged per CTA
  static constexpr int b_unpacked = b_rows * BK;
  static constexpr int b_stage = W4 ? b_unpacked / 2 : b_unpacked;
  static constexpr int stage = A_BYTES + b_stage;
  static constexpr int ring = 0;
  static constexpr int unpack = stages * stage;
  static constexpr int staging = unpack + (W4 ? UNPACK_BUFS * b_unpacked : 0);
  static constexpr int colparams = staging + EPI_WARPS * staging_bufs * STAGING_BYTES;   // dw[BN] f32, bias[BN] f32, zp[BN] i16
  static constexpr int barriers = colparams + BN * 4 + BN * 4 + BN * 2;
  static constexpr int total = barriers + 384;
  static constexpr int threads = (CTRL_WARPS + EPI_WARPS + (W4 ? CVT_WARPS : 0)) * 32;
  static_assert(total <= 232448, "dynamic smem budget (227 KB) exceeded");
};

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

// torch GELU(approximate='tanh') (wan/modules/model.py:287) on two values, packed fp32x2 math + tanh on the MUFU:
//   0.5 x (1 + tanh(c (x + 0.044715 x^3))) = h + h tanh(x (c + 0.044715 c x^2)),  h = x / 2, c = sqrt(2 / pi)
// five FMA-pipe issue slots per PAIR (FMUL2, FFMA2, FMUL2, FMUL2, FFMA2) and one MUFU.TANH per value: the epilogue of the
// ffn.0 GEMM was issue-bound with the scalar form (ncu: 49 % issue, 64 % tensor pipe).
__device__ __forceinline__ uint64_t gelu_tanh2(uint64_t x2) {
  const uint64_t s2 = mul_f32x2(x2, x2);
  const uint64_t w2 = fma_f32x2(s2, pack_f32x2(0.035677408136300125f, 0.035677408136300125f),
                                pack_f32x2(0.7978845608028654f, 0.7978845608028654f));
  float u0, u1, t0, t1;
  unpack_f32x2(mul_f32x2(w2, x2), u0, u1);
  asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(u0));
  asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(u1));
  const uint64_t h2 = mul_f32x2(x2, pack_f32x2(0.5f, 0.5f));
  return fma_f32x2(h2, pack_f32x2(t0, t1), h2);
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
// CL = CTAs per cluster (1 or 2).  CL == 2: the two CTAs of a cluster work on vertically adjacent tiles (same n0);
// each loads its own A tile and HALF of the shared B tile, multicast into both CTAs' smem, which cuts the L2->SM
// operand traffic per tile from 48 KB to 32 KB per K-block (the int8 MMA rate is L2-bandwidth-bound at 128x256 tiles).
template <typename OutT, int EPI, bool W4, int NSTAGE, int PAIR>
__global__ void __launch_bounds__(GemmSmem<W4, NSTAGE, PAIR>::threads, 1)
gemm_i8_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                 const __grid_constant__ CUtensorMap tm_out, const __grid_constant__ CUtensorMap tm_res,
                 const GemmParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];          // SWIZZLE_128B tiles need 1024-byte alignment
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) {
    printf("b200q: dynamic smem base not 1024-byte aligned\n");
    __trap();
  }
  using SM = GemmSmem<W4, NSTAGE, PAIR>;
  constexpr int STAGE_BYTES = SM::stage;
  constexpr int STAGES = SM::stages;
  constexpr int CL = PAIR ? 2 : 1;
  constexpr bool MMA2 = SM::mma2;
  constexpr int UNPACK_BYTES = SM::b_unpacked;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + SM::barriers);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint64_t* res_bar = tmem_empty_bar + 2;                      // [EPI_WARPS][3]: residual box landed in staging buffer b
  uint64_t* bready_bar = res_bar + EPI_WARPS * 3;              // W4: unpacked B tile ready (converter -> MMA)
  uint64_t* bfree_bar = bready_bar + UNPACK_BUFS;              // W4: unpacked B tile consumed (MMA -> converter)
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(bfree_bar + UNPACK_BUFS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_n = (p.N + BN - 1) / BN, tiles_m = (p.M + BM - 1) / BM;
  // a "tile" index below is a CLUSTER tile: CL vertically adjacent 128-row blocks sharing one 256-column block
  const int num_tiles = ((tiles_m + CL - 1) / CL) * tiles_n;
  const int num_kb = (p.K + BK - 1) / BK;
  const int cta_rank = (CL == 2) ? (int)cluster_ctarank() : 0;
  const int first_tile = (int)blockIdx.x / CL, tile_step = (int)gridDim.x / CL;
  auto tile_m0 = [&](int t) { return ((t / tiles_n) * CL + cta_rank) * BM; };
  auto tile_n0 = [&](int t) { return (t % tiles_n) * BN; };
  constexpr uint16_t kAllCtas = (uint16_t)((1u << CL) - 1);

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      // a ring slot is free when every MMA warp (and, for W4, every converter warp) of the CLUSTER is done with it:
      // with CL == 2 the peer multicasts half of B into this CTA's slot
      // MMA2: one MMA issuer (the leader) releases both CTAs' slots; each CTA's converters release their own packed tile
      mbar_init(&empty_bar[s], MMA2 ? (W4 ? 1 + CVT_WARPS : 1) : CL * (W4 ? 1 + CVT_WARPS : 1));
    }
    for (int s = 0; s < UNPACK_BUFS; ++s) {
      mbar_init(&bready_bar[s], MMA2 ? 2 * CVT_WARPS : CVT_WARPS);   // MMA2: the leader waits for both CTAs' converters
      mbar_init(&bfree_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], MMA2 ? 2 * EPI_WARPS : EPI_WARPS);   // MMA2: both CTAs' epilogues drain before the leader reuses TMEM
    }
    for (int s = 0; s < EPI_WARPS * 3; ++s) mbar_init(&res_bar[s], 1);
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tm_a); prefetch_tmap(&tm_b); prefetch_tmap(&tm_out);
    if (EPI == B200Q_EPI_GATE_RESIDUAL) prefetch_tmap(&tm_res);
  }
  if (warp == 2) { if (MMA2) tmem_alloc_2sm<TMEM_COLS>(tmem_base_slot); else tmem_alloc<TMEM_COLS>(tmem_base_slot); }
  tcgen05_fence_before();
  __syncthreads();
  if (CL == 2) cluster_sync();                                 // peer's barriers are initialised before anyone signals them
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = first_tile; tile < num_tiles; tile += tile_step) {
        const int m0 = tile_m0(tile), n0 = tile_n0(tile);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + SM::ring + stage * STAGE_BYTES;
          if (MMA2) {
            // both CTAs stage their A rows and their half of B locally.  The leader issues the pair's MMAs, so operand
            // bytes the tensor core reads directly are credited to the LEADER's barrier; a packed W4 tile is consumed by
            // the local converter warps first and is credited to the local barrier.
            const uint32_t lead_bar = mapa_u32(&full_bar[stage], 0);
            if (W4) {
              mbar_expect_tx(&full_bar[stage], cta_rank == 0 ? 2 * A_BYTES + SM::b_stage : SM::b_stage);
              tma_load_2d_2sm(sa, &tm_a, lead_bar, kb * BK, m0);
              tma_load_2d(sa + A_BYTES, &tm_b, &full_bar[stage], kb * (BK / 2), n0 + cta_rank * (BN / 2));
            } else {
              if (cta_rank == 0) mbar_expect_tx(&full_bar[stage], 2 * STAGE_BYTES);
              tma_load_2d_2sm(sa, &tm_a, lead_bar, kb * BK, m0);
              tma_load_2d_2sm(sa + A_BYTES, &tm_b, lead_bar, kb * BK, n0 + cta_rank * (BN / 2));
            }
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
            continue;
          }
          mbar_expect_tx(&full_bar[stage], STAGE_BYTES);       // own A + both halves of B (one arrives from the peer)
          tma_load_2d(sa, &tm_a, &full_bar[stage], kb * BK, m0);
          if (CL == 2) {
            tma_load_2d_mc(sa + A_BYTES + cta_rank * (SM::b_stage / 2), &tm_b, &full_bar[stage],
                           W4 ? kb * (BK / 2) : kb * BK, n0 + cta_rank * (BN / 2), kAllCtas);
          } else {
            tma_load_2d(sa + A_BYTES, &tm_b, &full_bar[stage], W4 ? kb * (BK / 2) : kb * BK, n0);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0 && (!MMA2 || cta_rank == 0)) {
      constexpr uint32_t idesc = make_i8_idesc(MMA2 ? 2 * BM : BM, BN);
      int stage = 0; uint32_t phase = 0; int it = 0;
      int ub = 0; uint32_t uphase = 0;
      for (int tile = first_tile; tile < num_tiles; tile += tile_step, ++it) {
        const int as = it & 1; const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(&tmem_empty_bar[as], aphase ^ 1);          // epilogue has drained this accumulator
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + as * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          if (W4) mbar_wait(&bready_bar[ub], uphase);        // int4 tile expanded to int8 by the converter warps
          tcgen05_fence_after();
          const uint32_t sa = smem_u32(smem + SM::ring + stage * STAGE_BYTES);
          const uint64_t adesc = make_kmajor_sw128_desc(sa);
          const uint64_t bdesc = make_kmajor_sw128_desc(W4 ? smem_u32(smem + SM::unpack + ub * UNPACK_BYTES) : sa + A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // advance both descriptors by k*32 bytes inside the 128B swizzle row (address field is >>4)
            if (MMA2) mma_i8_ss_2sm(tmem_d, adesc + (uint64_t)(k * (UMMA_K >> 4)), bdesc + (uint64_t)(k * (UMMA_K >> 4)), idesc,
                                    (kb | k) != 0 ? 1u : 0u);
            else mma_i8_ss(tmem_d, adesc + (uint64_t)(k * (UMMA_K >> 4)), bdesc + (uint64_t)(k * (UMMA_K >> 4)), idesc,
                           (kb | k) != 0 ? 1u : 0u);
          }
          if (MMA2) mma_commit_2sm_mc(&empty_bar[stage], kAllCtas);
          else if (CL == 2) mma_commit_mc(&empty_bar[stage], kAllCtas);   // releases the slot in BOTH CTAs
          else mma_commit(&empty_bar[stage]);                // slot free once these MMAs have read it
          if (W4) {
            if (MMA2) mma_commit_2sm_mc(&bfree_bar[ub], kAllCtas);
            else mma_commit(&bfree_bar[ub]);
            if (++ub == UNPACK_BUFS) { ub = 0; uphase ^= 1; }
          }
          if (kb == num_kb - 1) {
            if (MMA2) mma_commit_2sm_mc(&tmem_full_bar[as], kAllCtas);   // accumulator ready in both CTAs' TMEM
            else mma_commit(&tmem_full_bar[as]);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (W4 && warp >= CTRL_WARPS + EPI_WARPS) {
    // ===================== W4A8 converter: packed int4 (TMA, no swizzle) -> int8 SWIZZLE_128B tile =====================
    // packed word (csrc/w4.cu): byte i = (code[i]+8) | (code[4+i]+8) << 4 for a group of 8 codes, so the low nibbles
    // are 4 consecutive K codes and the high nibbles the next 4: two AND/SHIFT ops per output word, no sign extension
    // (the +8 bias is folded into the zero-point term of the epilogue).
    const int ct = threadIdx.x - (CTRL_WARPS + EPI_WARPS) * 32;
    int stage = 0; uint32_t phase = 0; int ub = 0; uint32_t uphase = 0;
    for (int tile = first_tile; tile < num_tiles; tile += tile_step) {
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        mbar_wait(&bfree_bar[ub], uphase ^ 1);
        const uint8_t* src = smem + SM::ring + stage * STAGE_BYTES + A_BYTES;     // [b_rows][64 B]
        uint8_t* dst = smem + SM::unpack + ub * UNPACK_BYTES;                      // [b_rows][128 B] swizzled
#pragma unroll
        for (int i = 0; i < (SM::b_rows * 4) / (CVT_WARPS * 32); ++i) {
          const int ci = i * (CVT_WARPS * 32) + ct;       // 16-byte packed chunk index: row = ci/4, j = ci%4
          const int r = ci >> 2, j = ci & 3;
          const uint4 w = *reinterpret_cast<const uint4*>(src + ci * 16);
          uint4 o0, o1;
          o0.x = w.x & 0x0F0F0F0Fu; o0.y = (w.x >> 4) & 0x0F0F0F0Fu; o0.z = w.y & 0x0F0F0F0Fu; o0.w = (w.y >> 4) & 0x0F0F0F0Fu;
          o1.x = w.z & 0x0F0F0F0Fu; o1.y = (w.z >> 4) & 0x0F0F0F0Fu; o1.z = w.w & 0x0F0F0F0Fu; o1.w = (w.w >> 4) & 0x0F0F0F0Fu;
          uint8_t* row = dst + r * 128;
          *reinterpret_cast<uint4*>(row + (((2 * j) ^ (r & 7)) << 4)) = o0;
          *reinterpret_cast<uint4*>(row + (((2 * j + 1) ^ (r & 7)) << 4)) = o1;
        }
        fence_proxy_async_smem();                          // generic-proxy writes -> visible to the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) {
          if (MMA2 && cta_rank != 0) mbar_arrive_remote(&bready_bar[ub], 0);   // the leader issues the pair's MMAs
          else mbar_arrive(&bready_bar[ub]);
          mbar_arrive(&empty_bar[stage]);                  // this warp is done reading the packed tile
          if (CL == 2 && !MMA2) mbar_arrive_remote(&empty_bar[stage], cta_rank ^ 1);   // ... which the peer half-fills
        }
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
        if (++ub == UNPACK_BUFS) { ub = 0; uphase ^= 1; }
      }
    }
  } else if (warp >= CTRL_WARPS && warp < CTRL_WARPS + EPI_WARPS) {
    // ===================== epilogue =====================
    const int ew = warp - CTRL_WARPS;
    const int quarter = ew & 3;                              // TMEM lane quarter this warp may read (== warp % 4)
    const int half = ew >> 2;                                // which 128-column half of the tile
    const int etid = threadIdx.x - CTRL_WARPS * 32;
    float* s_dw = reinterpret_cast<float*>(smem + SM::colparams);
    float* s_bias = s_dw + BN;
    int16_t* s_zp = reinterpret_cast<int16_t*>(s_bias + BN);
    constexpr int ELEMS = OutPack<OutT>::ELEMS;              // outputs per 16 B
    constexpr int CPS = 8 * ELEMS;                           // columns per 128B staging row / TMA store box
    constexpr int CHUNKS = (BN / 2) / CPS;                   // store boxes per warp per tile (even)
    constexpr int NBUF = SM::staging_bufs;
    constexpr bool RAW = std::is_same<OutT, int32_t>::value;
    constexpr bool GATE = (EPI == B200Q_EPI_GATE_RESIDUAL);
    uint8_t* my_staging = smem + SM::staging + ew * NBUF * STAGING_BYTES;
    const int col_base = half * (BN / 2);
    uint32_t res_phase = 0;                                  // bit b: parity of res_bar[ew][b]'s next phase
    int sb = 0;                                              // staging buffer of the current chunk (chunk counter mod NBUF)
    constexpr int DIST = NBUF - 1;                           // residual boxes in flight per warp

    // residual prefetch (GATE): box [32 rows x CPS cols] of tile `t`, chunk `u` -> staging buffer `b`
    auto box_live = [&](int t, int u) {
      return t < num_tiles && tile_m0(t) + quarter * 32 < p.M && tile_n0(t) + col_base + u * CPS < p.N;
    };
    auto prefetch_residual = [&](int t, int u, int b) {
      while (u >= CHUNKS) { u -= CHUNKS; t += tile_step; }   // chunk u of tile t, counted on into my next tiles
      if (lane == 0 && box_live(t, u)) {
        // the store that last read this buffer is done: with NBUF > 1 that is the store of the PREVIOUS chunk (the box goes
        // DIST chunks ahead into the buffer the chunk before the current one used), so one store may stay outstanding
        if (NBUF > 1) tma_store_wait_read<1>();
        else tma_store_wait_read<0>();
        mbar_expect_tx(&res_bar[ew * 3 + b], STAGING_BYTES);
        tma_load_2d(my_staging + b * STAGING_BYTES, &tm_res, &res_bar[ew * 3 + b],
                    tile_n0(t) + col_base + u * CPS, tile_m0(t) + quarter * 32);
      }
    };
    if (GATE && NBUF > 1) {
#pragma unroll
      for (int d = 0; d < DIST; ++d) prefetch_residual(first_tile, d, d);
    }

    int it = 0;
    for (int tile = first_tile; tile < num_tiles; tile += tile_step, ++it) {
      const int m0 = tile_m0(tile), n0 = tile_n0(tile);
      const int as = it & 1; const uint32_t aphase = (it >> 1) & 1;
      if (!RAW) {
        named_bar_sync(1, EPI_WARPS * 32);                   // everyone finished reading the previous tile's params
        for (int i = etid; i < BN; i += EPI_WARPS * 32) {
          const int n = n0 + i; const bool ok = n < p.N;
          float dw = ok ? p.delta_w[n] : 0.f;
          float bs = (ok && p.bias) ? load_bias(p.bias, p.bias_dtype, n) : 0.f;
          if (GATE && ok && p.gate) {                        // (y*dw + b)*g == y*(dw*g) + b*g: fold the gate in
            const float g = p.gate[n];
            dw *= g; bs *= g;
          }
          s_dw[i] = dw;
          s_bias[i] = bs;
          s_zp[i] = (int16_t)((ok && p.zp_w) ? __float2int_rn(p.zp_w[n]) + p.zp_offset : p.zp_offset);
        }
        named_bar_sync(1, EPI_WARPS * 32);
      }
      const int row = m0 + quarter * 32 + lane;
      const float da = (!RAW && row < p.M) ? p.delta_a[row] : 0.f;
      const int rs = (!RAW && p.rowsum_a && row < p.M) ? p.rowsum_a[row] : 0;   // only read when a zero point is in play
      const uint64_t da2 = pack_f32x2(da, da);

      mbar_wait(&tmem_full_bar[as], aphase);
      tcgen05_fence_after();
      const uint32_t t_row = tmem_base + as * BN + col_base + ((uint32_t)(quarter * 32) << 16);

#pragma unroll 1
      for (int u = 0; u < CHUNKS; ++u) {
        uint8_t* sbuf_ptr = my_staging + sb * STAGING_BYTES;
        const bool live = box_live(tile, u);
        if (GATE && NBUF == 1) prefetch_residual(tile, u, 0);   // long-K variant: no spare buffer, load in place
        if (GATE && live) {                                  // residual box prefetched DIST chunks ago has landed
          mbar_wait(&res_bar[ew * 3 + sb], (res_phase >> sb) & 1u);
          res_phase ^= 1u << sb;
        }
        uint4 o[CPS / ELEMS];                                // this thread's 128-byte output row, packed
#pragma unroll
        for (int h = 0; h < CPS / 32; ++h) {
          uint32_t v[32];
          tmem_ld_32x32(t_row + u * CPS + h * 32, v);
          tmem_ld_wait();
          if (u == CHUNKS - 1 && h == CPS / 32 - 1) {        // accumulator fully read: hand TMEM back to the MMA warp
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (MMA2 && cta_rank != 0) mbar_arrive_remote(&tmem_empty_bar[as], 0);   // the leader owns the MMA schedule
              else mbar_arrive(&tmem_empty_bar[as]);
            }
          }
#pragma unroll
          for (int c = 0; c < 32 / ELEMS; ++c) {             // 16-byte chunks of this thread's row
            const int col = col_base + u * CPS + h * 32 + c * ELEMS;   // column inside the tile (multiple of 4)
            const int j = (h * 32) / ELEMS + c;              // chunk index inside the 128B staging row
            if (RAW) {
              o[j] = make_uint4(v[c * 4], v[c * 4 + 1], v[c * 4 + 2], v[c * 4 + 3]);
            } else {
              float y[ELEMS];
#pragma unroll
              for (int g4 = 0; g4 < ELEMS / 4; ++g4) {
                const float4 dw4 = *reinterpret_cast<const float4*>(&s_dw[col + 4 * g4]);
                const float4 b4 = *reinterpret_cast<const float4*>(&s_bias[col + 4 * g4]);
                const short4 z4 = *reinterpret_cast<const short4*>(&s_zp[col + 4 * g4]);
                const float dwv[4] = {dw4.x, dw4.y, dw4.z, dw4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
                const int zv[4] = {z4.x, z4.y, z4.z, z4.w};
#pragma unroll
                for (int e = 0; e < 4; e += 2) {                 // two outputs per packed fp32x2 operation
                  const int acc0 = (int)v[c * ELEMS + 4 * g4 + e] + zv[e] * rs;
                  const int acc1 = (int)v[c * ELEMS + 4 * g4 + e + 1] + zv[e + 1] * rs;
                  const uint64_t sc2 = mul_f32x2(pack_f32x2(dwv[e], dwv[e + 1]), da2);
                  uint64_t t2 = fma_f32x2(pack_f32x2((float)acc0, (float)acc1), sc2, pack_f32x2(bv[e], bv[e + 1]));
                  if (EPI == B200Q_EPI_GELU_TANH) t2 = gelu_tanh2(t2);
                  unpack_f32x2(t2, y[4 * g4 + e], y[4 * g4 + e + 1]);
                }
              }
              if (GATE) {                                    // residual (fp32 x4) sits where the output goes
                const uint4 r4 = *reinterpret_cast<const uint4*>(sbuf_ptr + lane * 128 + ((j ^ (lane & 7)) << 4));
                y[0] += __uint_as_float(r4.x); y[1] += __uint_as_float(r4.y);
                y[2] += __uint_as_float(r4.z); y[3] += __uint_as_float(r4.w);
              }
              o[j] = pack_chunk<OutT>(y);
            }
          }
        }
        if (!GATE) {                                         // single staging buffer: the previous store must have read it
          if (lane == 0) tma_store_wait_read<0>();
          __syncwarp();
        }
#pragma unroll
        for (int j = 0; j < CPS / ELEMS; ++j)
          *reinterpret_cast<uint4*>(sbuf_ptr + lane * 128 + ((j ^ (lane & 7)) << 4)) = o[j];
        fence_proxy_async_smem();                            // generic-proxy smem writes -> visible to TMA
        __syncwarp();
        if (lane == 0) {
          if (live) tma_store_2d(&tm_out, sbuf_ptr, n0 + col_base + u * CPS, m0 + quarter * 32);
          tma_store_commit();
        }
        if (GATE && NBUF > 1) {                              // the box DIST chunks ahead (this tile or my next ones) goes into
          const int pb = sb == 0 ? NBUF - 1 : sb - 1;        // the buffer the previous chunk used
          prefetch_residual(tile, u + DIST, pb);
          if (++sb == NBUF) sb = 0;
        }
      }
    }
    if (lane == 0) tma_store_wait<0>();
  }

  tcgen05_fence_before();
  __syncthreads();
  if (CL == 2) cluster_sync();                                 // the peer may still signal this CTA's barriers / fill its smem
  tcgen05_fence_after();
  if (warp == 2) { if (MMA2) tmem_dealloc_2sm<TMEM_COLS>(tmem_base); else tmem_dealloc<TMEM_COLS>(tmem_base); }
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

template <typename OutT, int EPI, bool W4, int NSTAGE, int PAIR>
static int launch_gemm_cl(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, const CUtensorMap& tr,
                          const GemmParams& p, cudaStream_t st) {
  constexpr int CL = PAIR ? 2 : 1;
  auto kern = gemm_i8_kernel<OutT, EPI, W4, NSTAGE, PAIR>;
  static bool configured = false;    // cudaFuncSetAttribute once per instantiation, not per call
  const int smem_bytes = GemmSmem<W4, NSTAGE, PAIR>::total;
  if (!configured) {
    B200Q_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    configured = true;
  }
  const int tiles_m = (p.M + BM - 1) / BM, tiles_n = (p.N + BN - 1) / BN;
  const int cluster_tiles = ((tiles_m + CL - 1) / CL) * tiles_n;
  int grid = cluster_tiles * CL;
  const int cap = (sm_count() / CL) * CL;                      // persistent: one CTA per SM, whole clusters only
  if (grid > cap) grid = cap;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)GemmSmem<W4, NSTAGE, PAIR>::threads);
  cfg.dynamicSmemBytes = (size_t)smem_bytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  B200Q_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, ta, tb, to, tr, p));
  B200Q_CHECK_LAUNCH();
  return B200Q_OK;
}

// tb1: B tensor map with a full [256 x 128B] box (CL = 1); tb2: half box [128 x 128B] for the multicast path (CL = 2)
struct GemmMaps { CUtensorMap a, b1, b2, out, res; };

// scheduling knob (b200q_gemm_set_cluster): 0 = auto, 1 = single-CTA tiles, 2 = 2-CTA clusters with multicast B,
// 3 = 2-CTA clusters with tcgen05.mma.cta_group::2
static int g_force_cluster = 0;
// measured on B200 (tools/probe_gemm_modes.py): cta_group::2 pairs win for W8A8 (+3..15 %); the W4A8 converter warps
// lose from the cross-CTA coupling, so W4A8 stays on single-CTA tiles
static constexpr int kAutoPairW8 = 2, kAutoPairW4 = 0;

template <typename OutT, int EPI, bool W4, int NSTAGE = 4>
static int launch_gemm(const GemmMaps& m, const GemmParams& p, cudaStream_t st) {
  const int tiles_m = (p.M + BM - 1) / BM, tiles_n = (p.N + BN - 1) / BN;
  // pairs only pay off when there is more than one wave of tiles and at least two row blocks
  int pair = (tiles_m >= 2 && tiles_m * tiles_n >= 2 * sm_count()) ? (W4 ? kAutoPairW4 : kAutoPairW8) : 0;
  if (g_force_cluster == 1) pair = 0;
  if (g_force_cluster >= 2) pair = tiles_m >= 2 ? g_force_cluster - 1 : 0;
  if (pair == 2) return launch_gemm_cl<OutT, EPI, W4, NSTAGE, 2>(m.a, m.b2, m.out, m.res, p, st);
  if (pair == 1) return launch_gemm_cl<OutT, EPI, W4, NSTAGE, 1>(m.a, m.b2, m.out, m.res, p, st);
  return launch_gemm_cl<OutT, EPI, W4, NSTAGE, 0>(m.a, m.b1, m.out, m.res, p, st);
}

template <typename OutT, bool W4>
static int dispatch_epi(int epi, const GemmMaps& m, const GemmParams& p, cudaStream_t st) {
  switch (epi) {
    case B200Q_EPI_NONE: return launch_gemm<OutT, B200Q_EPI_NONE, W4>(m, p, st);
    case B200Q_EPI_GELU_TANH: return launch_gemm<OutT, B200Q_EPI_GELU_TANH, W4>(m, p, st);
  }
  set_error("gemm: unsupported epilogue %d for this out_dtype", epi);
  return B200Q_ERR_BAD_ARG;
}

template <bool W4>
int gemm_i8_common(const int8_t* qa, int64_t lda, const void* qw, int64_t ldw, const float* delta_a,
                   const float* delta_w, const float* zp_w, const int32_t* rowsum_a, const void* bias, int bias_dtype,
                   void* out, int out_dtype, int64_t ldo, int64_t M, int64_t N, int64_t K, int epilogue,
                   const float* residual, int64_t ldr, const float* gate, cudaStream_t st, int zp_offset = 0) {
  B200Q_REQUIRE(M >= 0 && N >= 0 && K >= 0, B200Q_ERR_BAD_ARG, "gemm: negative shape");
  if (M == 0 || N == 0) return B200Q_OK;
  B200Q_REQUIRE(K > 0, B200Q_ERR_BAD_ARG, "gemm: K must be > 0");
  B200Q_REQUIRE(M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31), B200Q_ERR_UNSUPPORTED, "gemm: dimension >= 2^31");
  B200Q_REQUIRE(qa && qw && out, B200Q_ERR_BAD_ARG, "gemm: null operand pointer");
  const int64_t kw_bytes = W4 ? ((K + 7) / 8) * 4 : K;      // bytes of one weight row
  B200Q_REQUIRE(lda >= K && ldw >= kw_bytes && ldo >= N, B200Q_ERR_BAD_ARG, "gemm: leading dimension too small");
  B200Q_REQUIRE(!(W4 || zp_offset != 0) || rowsum_a != nullptr || out_dtype == B200Q_I32, B200Q_ERR_BAD_ARG,
                "gemm_w4a8: rowsum_a is required (the unsigned-nibble bias is folded through the zero-point term)");
  B200Q_REQUIRE(lda % 16 == 0 && ldw % 16 == 0 && aligned(qa, 16) && aligned(qw, 16), B200Q_ERR_BAD_ARG,
                "gemm: qa/qw must be 16-byte aligned with lda, ldw (bytes) multiples of 16 (TMA global-stride rule)");
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

  GemmMaps mp;
  CUtensorMap &ta = mp.a, &to = mp.out, &tr = mp.res;
  int rc;
  if ((rc = make_tmap_2d(&ta, qa, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, M, K, lda, BM, BK, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  for (int half = 0; half < 2; ++half) {
    CUtensorMap* tb = half ? &mp.b2 : &mp.b1;
    const int box_rows = half ? BN / 2 : BN;
    if (W4) {
      if ((rc = make_tmap_2d(tb, qw, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, N, kw_bytes, ldw, box_rows, BK / 2, CU_TENSOR_MAP_SWIZZLE_NONE))) return rc;
    } else {
      if ((rc = make_tmap_2d(tb, qw, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, N, K, ldw, box_rows, BK, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    }
  }
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
  p.zp_offset = W4 ? -8 : zp_offset;

  if (raw) return launch_gemm<int32_t, B200Q_EPI_NONE, W4>(mp, p, st);
  if (epilogue == B200Q_EPI_GATE_RESIDUAL) {
    B200Q_REQUIRE(out_dtype == B200Q_F32, B200Q_ERR_BAD_ARG, "gemm: gate-residual epilogue writes the fp32 residual stream");
    B200Q_REQUIRE(residual != nullptr, B200Q_ERR_BAD_ARG, "gemm: residual required");
    B200Q_REQUIRE(ldr >= N && aligned(residual, 16) && (ldr * 4) % 16 == 0, B200Q_ERR_BAD_ARG, "gemm: bad residual layout");
    if ((rc = make_tmap_2d(&tr, residual, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, M, N, ldr, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    // short K: the epilogue (8 B/element of HBM traffic) is the bottleneck -> 3 ring stages + prefetched residual boxes;
    // long K: the main loop dominates -> keep the 4-stage ring
    if (K < 4096) return launch_gemm<float, B200Q_EPI_GATE_RESIDUAL, W4, 3>(mp, p, st);
    return launch_gemm<float, B200Q_EPI_GATE_RESIDUAL, W4, 4>(mp, p, st);
  }
  switch (out_dtype) {
    case B200Q_BF16: return dispatch_epi<__nv_bfloat16, W4>(epilogue, mp, p, st);
    case B200Q_F16: return dispatch_epi<__half, W4>(epilogue, mp, p, st);
    case B200Q_F32: return dispatch_epi<float, W4>(epilogue, mp, p, st);
  }
  return B200Q_ERR_BAD_ARG;
}

}  // namespace b200q

using namespace b200q;

extern "C" int b200q_gemm_set_cluster(int mode) {
  if (mode < 0 || mode > 3) return B200Q_ERR_BAD_ARG;
  g_force_cluster = mode;
  return B200Q_OK;
}

extern "C" int b200q_gemm_w8a8(const int8_t* qa, int64_t lda, const int8_t* qw, int64_t ldw, const float* delta_a,
                               const float* delta_w, const float* zp_w, const int32_t* rowsum_a, const void* bias,
                               int bias_dtype, void* out, int out_dtype, int64_t ldo, int64_t M, int64_t N, int64_t K,
                               int epilogue, const float* residual, int64_t ldr, const float* gate,
                               b200q_stream_t stream) {
  clear_error();
  return gemm_i8_common<false>(qa, lda, qw, ldw, delta_a, delta_w, zp_w, rowsum_a, bias, bias_dtype, out, out_dtype, ldo,
                               M, N, K, epilogue, residual, ldr, gate, (cudaStream_t)stream);
}

namespace b200q {
int gemm_w4a8_impl(const int8_t* qa, int64_t lda, const uint8_t* qw4, int64_t ldw4, const float* delta_a,
                   const float* delta_w, const float* zp_w, const int32_t* rowsum_a, const void* bias, int bias_dtype,
                   void* out, int out_dtype, int64_t ldo, int64_t M, int64_t N, int64_t K, int epilogue,
                   const float* residual, int64_t ldr, const float* gate, cudaStream_t st) {
  return gemm_i8_common<true>(qa, lda, qw4, ldw4, delta_a, delta_w, zp_w, rowsum_a, bias, bias_dtype, out, out_dtype, ldo,
                              M, N, K, epilogue, residual, ldr, gate, st);
}
// W4A8 through an expanded copy of the weights: `qw_u8` holds the unsigned nibbles (code + 8) as bytes, so this is the
// W8A8 kernel with the nibble bias folded through the zero-point term (zp_eff = zp_w - 8), bit-identical accumulators.
int gemm_w4a8_expanded_impl(const int8_t* qa, int64_t lda, const int8_t* qw_u8, int64_t ldu, const float* delta_a,
                            const float* delta_w, const float* zp_w, const int32_t* rowsum_a, const void* bias, int bias_dtype,
                            void* out, int out_dtype, int64_t ldo, int64_t M, int64_t N, int64_t K, int epilogue,
                            const float* residual, int64_t ldr, const float* gate, cudaStream_t st) {
  return gemm_i8_common<false>(qa, lda, qw_u8, ldu, delta_a, delta_w, zp_w, rowsum_a, bias, bias_dtype, out, out_dtype, ldo,
                               M, N, K, epilogue, residual, ldr, gate, st, -8);
}
}  // namespace b200q
