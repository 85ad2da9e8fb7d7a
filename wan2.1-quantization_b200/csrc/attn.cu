// (c) Quantized attention: int8 Q.K^T and u8 P.V on the 5th-gen tensor cores (tcgen05.mma.kind::i8, int32 accumulators
// in TMEM), operands streamed by TMA, softmax in registers between the two products.
//
// Replaces the reference's materialised fake-quant attention
// (ViDiT-Q/examples/Wan2.1/models/quant_opensora.py:430-478: q/k/v DynamicQuantizers, `q*scale @ k^T`, fp32 softmax,
// QuantizedAttentionMapOpenSORA, `attn @ v`), which builds S and P as [H, L, L] tensors (51 GB at L = 32,760) and so cannot
// run at the BASELINE shapes at all.
//
//   Q, K  int8 codes, one fp32 scale per (token, head)        = DynamicQuantizer on [B*H*L, hd] rows (:430-435)
//   V     int8 codes, one fp32 scale per (head, channel)      = DynamicQuantizer on [B*H*hd, L] rows (:440-442); stored
//         TRANSPOSED [H*hd, Lk] by b200q_quant_vt so that a key block is a K-major B operand
//   S     = Qq.Kq^T exact in int32;  x[i,j] = S[i,j] * dq[i]*dk[j]*sm_scale                 (:456-457)
//   P~    = exp(x - rowmax(x)) in (0,1], quantized to UNSIGNED 8 bit with the fixed step 1/255
//         (= the unsigned [0, 2^b-1] grid of DynamicQuantizer.forward_with_quant_params, base_quantizer.py:197-199, with one
//         step per query row: delta_i = rowmax(P_i)/255, because P_i = P~_i / l_i and rowmax(P~_i) = 1)
//   O     = (P~q . Vq) exact in int32, * dv[c] / (255 * l_i),  l_i = sum_j P~[i,j] in fp32
//
// Two passes over the keys, so that every P~ code of a row refers to the SAME (final) row maximum and the whole P.V
// product accumulates in one int32 TMEM accumulator (no per-block rescale of integer accumulators):
//   pass 1: S tiles only -> m_i = max_j S[i,j]*dk[j]                      (tensor pipe + 2.5 issue slots / element)
//   pass 2: S tiles again -> P~ codes -> shared memory (SWIZZLE_128B, the A operand of P.V) -> O += P~q.Vq
// The reference's own attention-map grouping ('row' = one scale per KEY column over all queries, quant_attn.py:168-174)
// needs the complete [L, L] map before the first code can be produced; it is kept as the small-L parity path
// (wan_b200/attention_q.py) and this kernel is the fast mode, reported as such in DESIGN.md.
//
// CTA = 256 queries of one head (two 128-row Q tiles), 384 threads, persistent over (head, q-tile-pair) work items:
//   warps 0-3 / 4-7  softmax warpgroup of Q tile 0 / 1: thread = one query row = one TMEM lane
//   warp 8           TMA producer (Q tiles, K ring, V^T ring)
//   warp 9           MMA issuer (one lane)
//   warp 10          per-key scale loader (dk -> shared memory ring) and TMEM allocator
// TMEM (512 columns): [S0 | O0 | S1 | O1], 128 columns each; in pass 1 the O columns serve as a second S buffer.
// The two warpgroups ping-pong: while one is in its softmax the tensor pipe produces the other's S tile / consumes its P.
#include "common.cuh"
#include "ptx.cuh"

namespace b200q {
using namespace ptx;

int make_tmap_2d(CUtensorMap* out, const void* base, CUtensorMapDataType dt, int elem_bytes, int64_t rows, int64_t cols,
                 int64_t ld, int box_rows, int box_cols, CUtensorMapSwizzle swz);

namespace attn {

constexpr int BQ = 128;            // queries per Q tile (one TMEM lane each)
constexpr int BKEY = 128;          // keys per block
constexpr int HD = 128;            // head dim == bytes per int8 row == one 128B swizzle row
constexpr int TILE = 128 * 128;    // every operand tile is 16 KB
constexpr int KS = 4, VS = 3, SCS = 4;
// NW = softmax warpgroups per Q tile.  NW = 1: one thread per query row handles all 128 keys of a block (352 threads).
// NW = 2: two warpgroups share a row, 64 keys each (608 threads): four compute warps per scheduler instead of two, which
// hides the fixed-latency dependency stalls the profile of the NW = 1 kernel shows (24 % of warp samples).
template <int NW> struct Cfg {
  static constexpr int compute_warps = 8 * NW;
  static constexpr int threads = (compute_warps + 3) * 32;
  static constexpr int cols = 128 / NW;           // key columns of a block per warpgroup
  static constexpr int cw = 32 / NW;              // columns per tcgen05.ld / per inner chunk (4 chunks per block per WG)
};
constexpr uint32_t MAGIC_I = 0x4B400000u;    // bit pattern of 1.5*2^23: as_float(MAGIC_I + s) == 12582912 + s for |s| < 2^22
constexpr float MAGIC_F = 12582912.0f;

struct Smem {
  static constexpr int q = 0;                          // 2 tiles
  static constexpr int k = q + 2 * TILE;               // KS tiles
  static constexpr int v = k + KS * TILE;              // VS tiles
  static constexpr int p = v + VS * TILE;              // 2 tiles x 2 buffers
  static constexpr int sc = p + 4 * TILE;              // SCS x 128 x (c, d)
  static constexpr int xch = sc + SCS * 1024;          // row-statistics exchange between the warpgroups of a tile (NW = 2)
  static constexpr int bar = xch + 2 * 2 * 2 * 128 * 4;
  static constexpr int total = bar + 512;
};
static_assert(Smem::total <= 232448, "dynamic smem budget (227 KB) exceeded");

struct Params {
  int Lq, Lk, H;
  const float* dq; long long dq_st, dq_sh;
  const float* dk; long long dk_st, dk_sh;
  const float* dv;
  __nv_bfloat16* out; long long ldo;
  float scale_log2e;
  int n_items, n_qt;
  int p1_two;                    // pass 1 hands two key blocks (256 S columns) to the warpgroup per barrier round trip
  float* m_out; float* l_out;
  uint8_t* p_out; long long ldp;
  int32_t* acc_out; long long ldacc;
};

// kind::i8 instruction descriptor, D = s32, K-major operands; A unsigned (P codes) or signed (Q codes), B signed
__host__ __device__ constexpr uint32_t idesc_i8(int M, int N, bool a_signed) {
  return (2u << 4) | ((a_signed ? 1u : 0u) << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// S accumulators are PRE-INITIALISED to MAGIC_I by the warpgroup that releases them (tcgen05.st) and every Q.K^T MMA
// accumulates on top: the tensor core hands back MAGIC_I + S, whose bit pattern read as fp32 is 12582912 + S exactly, so
// the int32 -> fp32 conversion costs no instruction (it is folded into the per-key FMA constant d = -12582912*dk).
__device__ __forceinline__ void fill_magic(uint32_t taddr, int ncols) {
  for (int c = 0; c < ncols; c += 16) tmem_st_32x16_splat(taddr + c, MAGIC_I);
  tmem_st_wait();
}

// exp2 of two non-positive arguments on the FMA/ALU pipes (no MUFU): Cody-Waite split x = n + r, r in [-0.5, 0.5], degree-4
// minimax polynomial for 2^r (max relative error 7e-6 = 0.002 code steps), 2^n by integer addition into the exponent
// field.  The softmax of pass 2 is bound by the 16 MUFU.EX2 per cycle per SM; routing a quarter of the exponentials through
// this path balances the MUFU against the FMA and ALU pipes (the technique FlashAttention-4 uses on the same hardware).
__device__ __forceinline__ uint64_t exp2_poly2(float x0, float x1) {
  const uint64_t magic2 = pack_f32x2(MAGIC_F, MAGIC_F), nmagic2 = pack_f32x2(-MAGIC_F, -MAGIC_F);
  const uint64_t xc = pack_f32x2(fmaxf(x0, -126.f), fmaxf(x1, -126.f));
  const uint64_t t = add_f32x2(xc, magic2);                      // integer part n = rne(x) in the low mantissa bits
  const uint64_t r = fma_f32x2(add_f32x2(t, nmagic2), pack_f32x2(-1.f, -1.f), xc);   // r = x - n
  uint64_t q = fma_f32x2(r, pack_f32x2(0.009670767933130264f, 0.009670767933130264f), pack_f32x2(0.05587553605437279f, 0.05587553605437279f));
  q = fma_f32x2(q, r, pack_f32x2(0.24022211134433746f, 0.24022211134433746f));
  q = fma_f32x2(q, r, pack_f32x2(0.6931272745132446f, 0.6931272745132446f));
  q = fma_f32x2(q, r, pack_f32x2(1.0f, 1.0f));
  uint32_t q0, q1, t0, t1;
  unpack_u32x2(q, q0, q1);
  unpack_u32x2(t, t0, t1);
  return pack_u32x2(q0 + (t0 << 23), q1 + (t1 << 23));          // (MAGIC_I + n) << 23 == n << 23 (mod 2^32)
}

template <int CW>
__device__ __forceinline__ void tmem_ld_chunk(uint32_t taddr, uint32_t* v) {
  if constexpr (CW == 32) tmem_ld_32x32(taddr, v); else tmem_ld_32x16(taddr, v);
}

template <bool PREMAGIC, bool POLY, int NW>
__global__ void __launch_bounds__(Cfg<NW>::threads, 1)
attn_i8_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
               const __grid_constant__ CUtensorMap tm_v, const Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) {
    printf("b200q: dynamic smem base not 1024-byte aligned\n");
    __trap();
  }
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Smem::bar);
  uint64_t* q_full = bars;                 // 1
  uint64_t* q_empty = q_full + 1;          // 1
  uint64_t* k_full = q_empty + 1;          // KS
  uint64_t* k_empty = k_full + KS;         // KS
  uint64_t* v_full = k_empty + KS;         // VS
  uint64_t* v_empty = v_full + VS;         // VS
  uint64_t* sc_full = v_empty + VS;        // SCS
  uint64_t* sc_empty = sc_full + SCS;      // SCS
  uint64_t* s_full = sc_empty + SCS;       // [tile]: S region written (pass 1: 256 columns, pass 2: 128)
  uint64_t* s_free = s_full + 2;           // [tile]: S region drained and re-initialised
  uint64_t* o_full = s_free + 2;           // [tile]: O accumulator complete
  uint64_t* o_free = o_full + 2;           // [tile]: O columns drained and re-initialised
  uint64_t* p_full = o_free + 2;           // [tile][pbuf] = 4
  uint64_t* p_free = p_full + 4;           // 4
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(p_free + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nb = (p.Lk + BKEY - 1) / BKEY;

  if (threadIdx.x == 0) {
    mbar_init(q_full, 1); mbar_init(q_empty, 1);
    for (int i = 0; i < KS; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); }
    for (int i = 0; i < VS; ++i) { mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1); }
    for (int i = 0; i < SCS; ++i) { mbar_init(&sc_full[i], 1); mbar_init(&sc_empty[i], 8 * NW); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1); mbar_init(&s_free[i], 4 * NW);
      mbar_init(&o_full[i], 1); mbar_init(&o_free[i], 4 * NW);
    }
    for (int i = 0; i < 4; ++i) { mbar_init(&p_full[i], 4 * NW); mbar_init(&p_free[i], 1); }
    fence_barrier_init();
  }
  constexpr int W_TMA = 8 * NW, W_MMA = 8 * NW + 1, W_SC = 8 * NW + 2;
  constexpr int COLS = Cfg<NW>::cols, CW = Cfg<NW>::cw;
  if (warp == W_TMA && lane == 0) { prefetch_tmap(&tm_q); prefetch_tmap(&tm_k); prefetch_tmap(&tm_v); }
  if (warp == W_SC) tmem_alloc<512>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == W_TMA) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int ks = 0, vs = 0; uint32_t kph = 0, vph = 0; int it = 0;
      auto load_k = [&](int h, int j) {
        mbar_wait(&k_empty[ks], kph ^ 1);
        mbar_expect_tx(&k_full[ks], TILE);
        tma_load_2d(smem + Smem::k + ks * TILE, &tm_k, &k_full[ks], h * HD, j * BKEY);
        if (++ks == KS) { ks = 0; kph ^= 1; }
      };
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++it) {
        const int h = item / p.n_qt, q0 = (item % p.n_qt) * (2 * BQ);
        mbar_wait(q_empty, (it & 1) ^ 1);
        mbar_expect_tx(q_full, 2 * TILE);
        tma_load_2d(smem + Smem::q, &tm_q, q_full, h * HD, q0);
        tma_load_2d(smem + Smem::q + TILE, &tm_q, q_full, h * HD, q0 + BQ);
        for (int j = 0; j < nb; ++j) load_k(h, j);                       // pass 1
        load_k(h, 0);                                                    // pass 2: K(0), then K(j+1), V(j)
        for (int j = 0; j < nb; ++j) {
          if (j + 1 < nb) load_k(h, j + 1);
          mbar_wait(&v_empty[vs], vph ^ 1);
          mbar_expect_tx(&v_full[vs], TILE);
          tma_load_2d(smem + Smem::v + vs * TILE, &tm_v, &v_full[vs], j * BKEY, h * HD);
          if (++vs == VS) { vs = 0; vph ^= 1; }
        }
      }
    }
  } else if (warp == W_MMA) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc_qk = idesc_i8(BQ, BKEY, true);
      constexpr uint32_t idesc_pv = idesc_i8(BQ, HD, false);
      uint32_t useS[2] = {0, 0};                     // S-region uses so far per Q tile
      uint32_t pcnt[2] = {0, 0};                     // P tiles consumed so far per Q tile
      int ks = 0, vs = 0; uint32_t kph = 0, vph = 0; int it = 0;
      const uint32_t q_addr = smem_u32(smem + Smem::q);
      // S[t][col .. col+127] += Q_t . K(stage)^T   (accumulates on the MAGIC_I pre-initialised columns)
      auto qk = [&](int t, int col, int stage) {
        const uint64_t adesc = make_kmajor_sw128_desc(q_addr + t * TILE);
        const uint64_t bdesc = make_kmajor_sw128_desc(smem_u32(smem + Smem::k + stage * TILE));
        const uint32_t d = tmem_base + t * 256 + col;
#pragma unroll
        for (int k = 0; k < HD / 32; ++k)
          mma_i8_ss(d, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc_qk, (PREMAGIC || k != 0) ? 1u : 0u);
      };
      auto release_k = [&]() {
        mma_commit(&k_empty[ks]);
        if (++ks == KS) { ks = 0; kph ^= 1; }
      };
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++it) {
        mbar_wait(q_full, it & 1);
        // ---- pass 1: two key blocks (256 S columns) per hand-off ----
        for (int j = 0; j < nb; j += (p.p1_two ? 2 : 1)) {
          const bool two = j + 1 < nb && p.p1_two;
          const int ks2 = (ks + 1 == KS) ? 0 : ks + 1;
          mbar_wait(&k_full[ks], kph);
          if (two) mbar_wait(&k_full[ks2], (ks + 1 == KS) ? (kph ^ 1) : kph);
          for (int t = 0; t < 2; ++t) {
            mbar_wait(&s_free[t], useS[t] & 1);
            if (j == 0) mbar_wait(&o_free[t], it & 1);                   // the O columns double as S columns in pass 1
            tcgen05_fence_after();
            qk(t, 0, ks);
            if (two) qk(t, 128, ks2);
            mma_commit(&s_full[t]);
            ++useS[t];
          }
          release_k();
          if (two) release_k();
        }
        // ---- pass 2 ----
        mbar_wait(&k_full[ks], kph);
        for (int t = 0; t < 2; ++t) {
          mbar_wait(&s_free[t], useS[t] & 1);
          tcgen05_fence_after();
          qk(t, 0, ks);
          mma_commit(&s_full[t]);
          ++useS[t];
        }
        release_k();
        if (nb == 1) mma_commit(q_empty);
        for (int j = 0; j < nb; ++j) {
          const bool more = j + 1 < nb;
          if (more) mbar_wait(&k_full[ks], kph);
          for (int t = 0; t < 2; ++t) {
            if (more) {                                                  // S(j+1) as soon as S(j) has been read
              mbar_wait(&s_free[t], useS[t] & 1);
              tcgen05_fence_after();
              qk(t, 0, ks);
              mma_commit(&s_full[t]);
              ++useS[t];
            }
            if (more && t == 1) {
              release_k();
              if (j + 2 == nb) mma_commit(q_empty);                      // last read of the Q tiles
            }
            if (t == 0) mbar_wait(&v_full[vs], vph);
            const int pb = pcnt[t] & 1;
            mbar_wait(&p_full[t * 2 + pb], (pcnt[t] >> 1) & 1);
            tcgen05_fence_after();
            const uint64_t adesc = make_kmajor_sw128_desc(smem_u32(smem + Smem::p + (t * 2 + pb) * TILE));
            const uint64_t bdesc = make_kmajor_sw128_desc(smem_u32(smem + Smem::v + vs * TILE));
            const uint32_t d = tmem_base + t * 256 + 128;
#pragma unroll
            for (int k = 0; k < BKEY / 32; ++k)
              mma_i8_ss(d, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc_pv, (j | k) != 0 ? 1u : 0u);
            mma_commit(&p_free[t * 2 + pb]);
            ++pcnt[t];
            if (!more) mma_commit(&o_full[t]);                           // O complete
          }
          mma_commit(&v_empty[vs]);
          if (++vs == VS) { vs = 0; vph ^= 1; }
        }
      }
    }
  } else if (warp == W_SC) {
    // ===================== per-key scale loader =====================
    int scs = 0; uint32_t scph = 0;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      const int h = item / p.n_qt;
      for (int pass = 0; pass < 2; ++pass) {
        for (int j = 0; j < nb; ++j) {
          mbar_wait(&sc_empty[scs], scph ^ 1);
          float c[4], d[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int key = j * BKEY + lane * 4 + i;
            const bool ok = key < p.Lk;
            c[i] = ok ? __ldg(p.dk + (long long)key * p.dk_st + (long long)h * p.dk_sh) : 0.f;
            d[i] = ok ? -MAGIC_F * c[i] : -INFINITY;                     // masked key: x = -inf -> P~ = 0
          }
          float4* s4 = reinterpret_cast<float4*>(smem + Smem::sc + scs * 1024);
          s4[lane * 2] = make_float4(c[0], c[1], d[0], d[1]);
          s4[lane * 2 + 1] = make_float4(c[2], c[3], d[2], d[3]);
          __syncwarp();
          if (lane == 0) mbar_arrive(&sc_full[scs]);
          if (++scs == SCS) { scs = 0; scph ^= 1; }
        }
      }
    }
  } else if (warp < 8 * NW) {
    // ===================== softmax warpgroups =====================
    const int t = warp / (4 * NW), half = (warp >> 2) % NW, quarter = warp & 3;
    const int r = quarter * 32 + lane;                                   // row inside the Q tile == TMEM lane
    const int c0 = half * COLS;                                          // first key column (of a block) / O column of this WG
    const uint32_t t_lane = tmem_base + t * 256 + ((uint32_t)(quarter * 32) << 16);
    uint32_t useS = 0, pcnt = 0, itn = 0;
    int scs = 0; uint32_t scph = 0;
    constexpr uint32_t MG = PREMAGIC ? 0u : MAGIC_I;                      // int -> fp32 bias added here unless pre-initialised
    const uint64_t magic2 = pack_f32x2(MAGIC_F, MAGIC_F), c255 = pack_f32x2(255.f, 255.f);
    float* xch_m = reinterpret_cast<float*>(smem + Smem::xch) + (t * 2) * 128;          // [half][row]
    float* xch_l = reinterpret_cast<float*>(smem + Smem::xch) + 512 + (t * 2) * 128;
    auto release = [&](uint64_t* bar) {                                  // TMEM reads/writes of this warp done -> MMA issuer
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar);
    };
    if (PREMAGIC) { fill_magic(t_lane + c0, COLS); fill_magic(t_lane + 128 + c0, COLS); }
    release(&s_free[t]);
    release(&o_free[t]);
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++itn) {
      const int h = item / p.n_qt, q0 = (item % p.n_qt) * (2 * BQ) + t * BQ;
      const int row = q0 + r;
      const bool row_ok = row < p.Lq;
      const float a = row_ok ? __ldg(p.dq + (long long)row * p.dq_st + (long long)h * p.dq_sh) * p.scale_log2e : 1.f;

      // ---- pass 1: m = max_j S[i,j]*dk[j] ----
      float m = -INFINITY;
      for (int j = 0; j < nb; j += (p.p1_two ? 2 : 1)) {
        const int nblk = (j + 1 < nb && p.p1_two) ? 2 : 1;
        mbar_wait(&s_full[t], useS & 1);
        tcgen05_fence_after();
        for (int hb = 0; hb < nblk; ++hb) {
          mbar_wait(&sc_full[scs], scph);
          const float4* s4 = reinterpret_cast<const float4*>(smem + Smem::sc + scs * 1024);
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) {
            uint32_t v[CW];
            tmem_ld_chunk<CW>(t_lane + hb * 128 + c0 + ch * CW, v);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < CW; e += 2) {
              const float4 cd = s4[(c0 + ch * CW + e) >> 1];
              const uint64_t t2 = fma_f32x2(pack_u32x2(v[e] + MG, v[e + 1] + MG), pack_f32x2(cd.x, cd.y), pack_f32x2(cd.z, cd.w));
              float t0, t1;
              unpack_f32x2(t2, t0, t1);
              m = fmaxf(m, fmaxf(t0, t1));
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&sc_empty[scs]);
          if (++scs == SCS) { scs = 0; scph ^= 1; }
        }
        if (PREMAGIC) { fill_magic(t_lane + c0, COLS); if (nblk == 2) fill_magic(t_lane + 128 + c0, COLS); }
        release(&s_free[t]);
        ++useS;
      }
      if constexpr (NW == 2) {                                           // the two warpgroups of a tile saw 64 keys per block each
        xch_m[half * 128 + r] = m;
        named_bar_sync(1 + t, 256);
        m = fmaxf(m, xch_m[(half ^ 1) * 128 + r]);
      }

      // ---- pass 2: P~ = exp2((S*dk - m)*a) -> u8 codes -> smem ; l = sum P~ ----
      const uint64_t a2 = pack_f32x2(a, a);
      const float nma = -m * a;
      const uint64_t nma2 = pack_f32x2(nma, nma);
      uint64_t sum2 = pack_f32x2(0.f, 0.f);
      for (int j = 0; j < nb; ++j) {
        const int pb = pcnt & 1;
        mbar_wait(&sc_full[scs], scph);
        mbar_wait(&p_free[t * 2 + pb], ((pcnt >> 1) & 1) ^ 1);           // P.V of two blocks ago has consumed this buffer
        mbar_wait(&s_full[t], useS & 1);
        tcgen05_fence_after();
        const float4* s4 = reinterpret_cast<const float4*>(smem + Smem::sc + scs * 1024);
        uint8_t* prow = smem + Smem::p + (t * 2 + pb) * TILE + r * 128;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          uint32_t v[CW];
          tmem_ld_chunk<CW>(t_lane + c0 + ch * CW, v);
          tmem_ld_wait();
          if (ch == 3) {                                                 // S(j) is in registers: hand the columns back
            if (PREMAGIC) fill_magic(t_lane + c0, COLS);
            release(&s_free[t]);
          }
          uint32_t w[CW / 4];
#pragma unroll
          for (int e = 0; e < CW; e += 4) {
            const float4 cd0 = s4[(c0 + ch * CW + e) >> 1], cd1 = s4[((c0 + ch * CW + e) >> 1) + 1];
            uint64_t t01 = fma_f32x2(pack_u32x2(v[e] + MG, v[e + 1] + MG), pack_f32x2(cd0.x, cd0.y), pack_f32x2(cd0.z, cd0.w));
            uint64_t t23 = fma_f32x2(pack_u32x2(v[e + 2] + MG, v[e + 3] + MG), pack_f32x2(cd1.x, cd1.y), pack_f32x2(cd1.z, cd1.w));
            t01 = fma_f32x2(t01, a2, nma2);
            t23 = fma_f32x2(t23, a2, nma2);
            float x0, x1, x2, x3;
            unpack_f32x2(t01, x0, x1);
            unpack_f32x2(t23, x2, x3);
            const uint64_t p01 = pack_f32x2(ex2_approx(x0), ex2_approx(x1));
            const uint64_t p23 = (POLY && (e & 4)) ? exp2_poly2(x2, x3) : pack_f32x2(ex2_approx(x2), ex2_approx(x3));
            sum2 = add_f32x2(sum2, add_f32x2(p01, p23));
            uint32_t u0, u1, u2, u3;
            unpack_u32x2(fma_f32x2(p01, c255, magic2), u0, u1);          // rne(P~*255) in the low mantissa byte
            unpack_u32x2(fma_f32x2(p23, c255, magic2), u2, u3);
            w[e >> 2] = __byte_perm(__byte_perm(u0, u1, 0x0040), __byte_perm(u2, u3, 0x0040), 0x5410);
          }
          // CW consecutive keys of this row = CW/16 16-byte chunks of the SWIZZLE_128B row
          const int k16 = (c0 + ch * CW) >> 4;
          *reinterpret_cast<uint4*>(prow + ((k16 ^ (r & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
          if constexpr (CW == 32)
            *reinterpret_cast<uint4*>(prow + (((k16 + 1) ^ (r & 7)) << 4)) = make_uint4(w[CW / 4 - 4], w[CW / 4 - 3], w[CW / 4 - 2], w[CW / 4 - 1]);
          if (p.p_out != nullptr && row_ok) {
            uint8_t* g = p.p_out + ((long long)h * p.Lq + row) * p.ldp + j * BKEY + c0 + ch * CW;
            *reinterpret_cast<uint4*>(g) = make_uint4(w[0], w[1], w[2], w[3]);
            if constexpr (CW == 32)
              *reinterpret_cast<uint4*>(g + 16) = make_uint4(w[CW / 4 - 4], w[CW / 4 - 3], w[CW / 4 - 2], w[CW / 4 - 1]);
          }
        }
        fence_proxy_async_smem();                                        // generic-proxy writes -> visible to the tensor core
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&p_full[t * 2 + pb]);
          mbar_arrive(&sc_empty[scs]);
        }
        if (++scs == SCS) { scs = 0; scph ^= 1; }
        ++useS; ++pcnt;
      }

      // ---- read-out: O = acc * dv[c] / (255 * l) ----
      float s0, s1;
      unpack_f32x2(sum2, s0, s1);
      float l = s0 + s1;
      if constexpr (NW == 2) {
        xch_l[half * 128 + r] = l;
        named_bar_sync(1 + t, 256);
        l += xch_l[(half ^ 1) * 128 + r];
      }
      const float inv = 1.0f / (255.0f * l);
      mbar_wait(&o_full[t], itn & 1);
      tcgen05_fence_after();
      if (half == 0) {
        if (row_ok && p.m_out != nullptr) p.m_out[(long long)h * p.Lq + row] = m * a;
        if (row_ok && p.l_out != nullptr) p.l_out[(long long)h * p.Lq + row] = l;
      }
      const float4* dv4 = reinterpret_cast<const float4*>(p.dv + h * HD);
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        uint32_t v[CW];
        tmem_ld_chunk<CW>(t_lane + 128 + c0 + ch * CW, v);
        tmem_ld_wait();
        if (ch == 3) {
          if (PREMAGIC) fill_magic(t_lane + 128 + c0, COLS);
          release(&o_free[t]);
        }
        if (row_ok) {
          const int col = c0 + ch * CW;
          if (p.acc_out != nullptr) {
            int32_t* g = p.acc_out + (long long)row * p.ldacc + h * HD + col;
#pragma unroll
            for (int e = 0; e < CW; e += 4) *reinterpret_cast<uint4*>(g + e) = make_uint4(v[e], v[e + 1], v[e + 2], v[e + 3]);
          }
          __nv_bfloat16* g = p.out + (long long)row * p.ldo + h * HD + col;
#pragma unroll
          for (int e = 0; e < CW; e += 8) {
            const float4 d0 = __ldg(dv4 + ((col + e) >> 2)), d1 = __ldg(dv4 + ((col + e) >> 2) + 1);
            __nv_bfloat162 o0 = __floats2bfloat162_rn((float)(int)v[e] * (d0.x * inv), (float)(int)v[e + 1] * (d0.y * inv));
            __nv_bfloat162 o1 = __floats2bfloat162_rn((float)(int)v[e + 2] * (d0.z * inv), (float)(int)v[e + 3] * (d0.w * inv));
            __nv_bfloat162 o2 = __floats2bfloat162_rn((float)(int)v[e + 4] * (d1.x * inv), (float)(int)v[e + 5] * (d1.y * inv));
            __nv_bfloat162 o3 = __floats2bfloat162_rn((float)(int)v[e + 6] * (d1.z * inv), (float)(int)v[e + 7] * (d1.w * inv));
            uint4 o;
            o.x = *reinterpret_cast<uint32_t*>(&o0); o.y = *reinterpret_cast<uint32_t*>(&o1);
            o.z = *reinterpret_cast<uint32_t*>(&o2); o.w = *reinterpret_cast<uint32_t*>(&o3);
            *reinterpret_cast<uint4*>(g + e) = o;
          }
        }
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  if (warp == W_SC) tmem_dealloc<512>(tmem_base);
}

// ---- V^T quantizer: per-(head, channel) scale over all tokens, codes written transposed ---------------------------------
// v [Lk, C] (row pitch ldv) -> vt [C, Lk] int8 (row pitch ldvt >= Lk, multiple of 16), delta[c] = max(amax[c]/127, 1e-6)
// (DynamicQuantizer sym on V^T rows, base_quantizer.py:110-129,151-157; quant_opensora.py:440-442).
// amax[c] comes from calib_kernel (one read of v); this kernel is the second read.  Tile = 128 tokens x 64 channels.
template <typename T>
__global__ void __launch_bounds__(256) quant_vt_kernel(const T* __restrict__ v, long long ldv, int Lk, int C,
                                                       const float* __restrict__ amax, int n_levels, int8_t* __restrict__ vt,
                                                       long long ldvt, float* __restrict__ delta) {
  __shared__ uint32_t tile[64][33];                  // [channel][token quad], +1 word pad: conflict-free both ways
  const int t0 = blockIdx.x * 128, c0 = blockIdx.y * 64;
  const int g = threadIdx.x >> 5, tq = threadIdx.x & 31;   // warp = 8-channel group, lane = 4 consecutive tokens
  constexpr int N = Vec16<T>::N;                     // 8 (16-bit types) or 4 (fp32)
  static_assert(N == 8 || N == 4, "");
  float d[8], rcp[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = c0 + g * 8 + i;
    float dl = (c < C ? amax[c] : 0.f) / (float)n_levels;
    if (dl < 1e-6f) dl = 1e-6f;                      // base_quantizer.py:122-128
    d[i] = dl; rcp[i] = 1.0f / dl;
    if (blockIdx.x == 0 && tq == 0 && c < C) delta[c] = dl;
  }
  uint32_t wv[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int tok = t0 + tq * 4 + k;
    float f[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (tok < Lk) {
      const T* src = v + (long long)tok * ldv + c0 + g * 8;
      if (c0 + g * 8 + 8 <= C) {
        if (N == 8) {
          Vec16<T>::unpack(*reinterpret_cast<const uint4*>(src), f);
        } else {
          Vec16<T>::unpack(*reinterpret_cast<const uint4*>(src), f);
          Vec16<T>::unpack(*reinterpret_cast<const uint4*>(src + 4), f + 4);
        }
      } else {
        for (int i = 0; i < 8; ++i) if (c0 + g * 8 + i < C) f[i] = to_f32(src[i]);
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int q = rne_to_int(div_rn_hoisted(f[i], d[i], rcp[i]));
      q = max(-n_levels - 1, min(n_levels, q));
      wv[i] |= ((uint32_t)q & 0xffu) << (8 * k);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) tile[g * 8 + i][tq] = wv[i];
  __syncthreads();
  // 64 rows x 128 B: 4 threads per row, 32 B each
  const int rowi = threadIdx.x >> 2, part = threadIdx.x & 3;
  const int c = c0 + rowi;
  if (c < C) {
    int8_t* dst = vt + (long long)c * ldvt + t0 + part * 32;
#pragma unroll
    for (int hlf = 0; hlf < 2; ++hlf) {
      if (t0 + part * 32 + hlf * 16 < ldvt) {
        const uint32_t* s = &tile[rowi][part * 8 + hlf * 4];
        *reinterpret_cast<uint4*>(dst + hlf * 16) = make_uint4(s[0], s[1], s[2], s[3]);
      }
    }
  }
}

}  // namespace attn
}  // namespace b200q

using namespace b200q;

extern "C" int b200q_quant_vt(const void* v, int v_dtype, int64_t Lk, int64_t C, int64_t ldv, int n_bits,
                              float* absmax_ws, int8_t* vt, int64_t ldvt, float* delta, b200q_stream_t stream) {
  clear_error();
  cudaStream_t st = (cudaStream_t)stream;
  B200Q_REQUIRE(v && absmax_ws && vt && delta, B200Q_ERR_BAD_ARG, "quant_vt: null pointer");
  B200Q_REQUIRE(Lk > 0 && C > 0 && Lk < (1ll << 31) && C < (1ll << 31), B200Q_ERR_BAD_ARG, "quant_vt: bad shape");
  B200Q_REQUIRE(n_bits >= 2 && n_bits <= 8, B200Q_ERR_BAD_ARG, "quant_vt: n_bits must be in [2, 8]");
  B200Q_REQUIRE(ldvt >= Lk && ldvt % 16 == 0 && aligned(vt, 16), B200Q_ERR_BAD_ARG,
                "quant_vt: vt must be 16-byte aligned with a row pitch that is a multiple of 16 and >= Lk");
  const int esz = v_dtype == B200Q_F32 ? 4 : 2;
  B200Q_REQUIRE(v_dtype >= B200Q_F32 && v_dtype <= B200Q_F16, B200Q_ERR_BAD_ARG, "quant_vt: bad dtype");
  B200Q_REQUIRE(ldv >= C && aligned(v, 16) && (ldv * esz) % 16 == 0, B200Q_ERR_BAD_ARG,
                "quant_vt: v must be 16-byte aligned with a 16-byte-multiple row pitch");
  B200Q_CUDA_OK(cudaMemsetAsync(absmax_ws, 0, (size_t)C * sizeof(float), st));
  int rc = b200q_calib_absmax_minmax(v, v_dtype, Lk, C, ldv, absmax_ws, nullptr, nullptr, stream);
  if (rc != B200Q_OK) return rc;
  const dim3 grid((unsigned)((Lk + 127) / 128), (unsigned)((C + 63) / 64));
  const int n_levels = (1 << (n_bits - 1)) - 1;
  switch (v_dtype) {
    case B200Q_F32:
      attn::quant_vt_kernel<float><<<grid, 256, 0, st>>>((const float*)v, ldv, (int)Lk, (int)C, absmax_ws, n_levels, vt, ldvt, delta);
      break;
    case B200Q_BF16:
      attn::quant_vt_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)v, ldv, (int)Lk, (int)C, absmax_ws, n_levels, vt, ldvt, delta);
      break;
    default:
      attn::quant_vt_kernel<__half><<<grid, 256, 0, st>>>((const __half*)v, ldv, (int)Lk, (int)C, absmax_ws, n_levels, vt, ldvt, delta);
  }
  B200Q_CHECK_LAUNCH();
  return B200Q_OK;
}

// scheduling knob (b200q_attn_set_mode): bit 0 = S accumulators pre-initialised with the int->fp32 bias (tcgen05.st),
// bit 1 = pass 1 hands two key blocks per barrier round trip, bit 2 = a quarter of the exponentials of pass 2 evaluated by
// a polynomial on the FMA/ALU pipes instead of the MUFU, bit 3 = two softmax warpgroups per Q tile (64 keys of a block
// each, 608 threads).  Bits 0, 1, 3: results identical (bit 3 up to the fp32 summation order of l); bit 2: P~ within 7e-6.
static int g_attn_mode = 2;   // measured on B200 (tools/probe_attn_i8.py, H=12 L=32760): mode 0 7.99 ms, 1 8.77, 2 7.57, 3 8.55
extern "C" int b200q_attn_set_mode(int mode) {
  if (mode < 0 || mode > 15) return B200Q_ERR_BAD_ARG;
  g_attn_mode = mode;
  return B200Q_OK;
}

extern "C" int b200q_attn_i8(const int8_t* qq, int64_t ldq, const float* dq, int64_t dq_tok_stride, int64_t dq_head_stride,
                             const int8_t* kq, int64_t ldk, const float* dk, int64_t dk_tok_stride, int64_t dk_head_stride,
                             const int8_t* vtq, int64_t ldvt, const float* dv, int64_t Lq, int64_t Lk, int num_heads,
                             int head_dim, float sm_scale, void* out, int out_dtype, int64_t ldo, float* m_out,
                             float* l_out, uint8_t* p_out, int64_t ldp, int32_t* acc_out, int64_t ldacc,
                             b200q_stream_t stream) {
  clear_error();
  using namespace attn;
  B200Q_REQUIRE(qq && dq && kq && dk && vtq && dv && out, B200Q_ERR_BAD_ARG, "attn_i8: null pointer");
  B200Q_REQUIRE(Lq > 0 && Lk > 0 && num_heads > 0, B200Q_ERR_BAD_ARG, "attn_i8: bad shape");
  B200Q_REQUIRE(head_dim == HD, B200Q_ERR_UNSUPPORTED, "attn_i8: head_dim must be 128 (Wan2.1: 1536/12 = 5120/40 = 128)");
  B200Q_REQUIRE(out_dtype == B200Q_BF16, B200Q_ERR_UNSUPPORTED, "attn_i8: output is bf16");
  B200Q_REQUIRE(Lq < (1ll << 31) - 512 && Lk < (1ll << 31) - 512, B200Q_ERR_UNSUPPORTED, "attn_i8: sequence too long");
  // int32 accumulator of P.V: Lk * 255 * 127 must stay below 2^31
  B200Q_REQUIRE(Lk <= 66000, B200Q_ERR_UNSUPPORTED, "attn_i8: Lk > 66000 could overflow the int32 P.V accumulator; split the keys");
  const int64_t D = (int64_t)num_heads * HD;
  B200Q_REQUIRE(ldq >= D && ldk >= D && ldq % 16 == 0 && ldk % 16 == 0 && aligned(qq, 16) && aligned(kq, 16),
                B200Q_ERR_BAD_ARG, "attn_i8: qq/kq must be 16-byte aligned [L, H*128] with pitches that are multiples of 16");
  B200Q_REQUIRE(ldvt >= Lk && ldvt % 16 == 0 && aligned(vtq, 16), B200Q_ERR_BAD_ARG,
                "attn_i8: vtq must be 16-byte aligned [H*128, Lk] with a pitch that is a multiple of 16");
  B200Q_REQUIRE(ldo >= D && ldo % 8 == 0 && aligned(out, 16), B200Q_ERR_BAD_ARG, "attn_i8: bad output layout");
  B200Q_REQUIRE(aligned(dv, 16), B200Q_ERR_BAD_ARG, "attn_i8: dv must be 16-byte aligned");
  if (p_out) B200Q_REQUIRE(ldp % 16 == 0 && ldp >= ((Lk + 127) / 128) * 128 && aligned(p_out, 16), B200Q_ERR_BAD_ARG,
                           "attn_i8: p_out pitch must be a multiple of 16 and cover whole 128-key blocks");
  if (acc_out) B200Q_REQUIRE(ldacc >= D && ldacc % 4 == 0 && aligned(acc_out, 16), B200Q_ERR_BAD_ARG, "attn_i8: bad acc_out layout");

  CUtensorMap tq, tk, tv;
  int rc;
  if ((rc = make_tmap_2d(&tq, qq, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, Lq, D, ldq, BQ, HD, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  if ((rc = make_tmap_2d(&tk, kq, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, Lk, D, ldk, BKEY, HD, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  if ((rc = make_tmap_2d(&tv, vtq, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, D, Lk, ldvt, HD, BKEY, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;

  Params p{};
  p.Lq = (int)Lq; p.Lk = (int)Lk; p.H = num_heads;
  p.dq = dq; p.dq_st = dq_tok_stride; p.dq_sh = dq_head_stride;
  p.dk = dk; p.dk_st = dk_tok_stride; p.dk_sh = dk_head_stride;
  p.dv = dv;
  p.out = (__nv_bfloat16*)out; p.ldo = ldo;
  p.scale_log2e = sm_scale * 1.4426950408889634f;
  p.n_qt = (int)((Lq + 2 * BQ - 1) / (2 * BQ));
  p.n_items = p.n_qt * num_heads;
  p.m_out = m_out; p.l_out = l_out; p.p_out = p_out; p.ldp = ldp; p.acc_out = acc_out; p.ldacc = ldacc;

  static bool configured = false;
  if (!configured) {
#define B200Q_ATTN_CFG(A, B, C) \
    B200Q_CUDA_OK(cudaFuncSetAttribute(attn_i8_kernel<A, B, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem::total))
    B200Q_ATTN_CFG(false, false, 1); B200Q_ATTN_CFG(true, false, 1); B200Q_ATTN_CFG(false, true, 1); B200Q_ATTN_CFG(true, true, 1);
    B200Q_ATTN_CFG(false, false, 2); B200Q_ATTN_CFG(true, false, 2); B200Q_ATTN_CFG(false, true, 2); B200Q_ATTN_CFG(true, true, 2);
#undef B200Q_ATTN_CFG
    configured = true;
  }
  p.p1_two = (g_attn_mode & 2) ? 1 : 0;
  int grid = p.n_items < sm_count() ? p.n_items : sm_count();
  cudaStream_t st = (cudaStream_t)stream;
#define B200Q_ATTN_LAUNCH(A, B, C) attn_i8_kernel<A, B, C><<<grid, Cfg<C>::threads, Smem::total, st>>>(tq, tk, tv, p)
  switch (g_attn_mode & 13) {
    case 0: B200Q_ATTN_LAUNCH(false, false, 1); break;
    case 1: B200Q_ATTN_LAUNCH(true, false, 1); break;
    case 4: B200Q_ATTN_LAUNCH(false, true, 1); break;
    case 5: B200Q_ATTN_LAUNCH(true, true, 1); break;
    case 8: B200Q_ATTN_LAUNCH(false, false, 2); break;
    case 9: B200Q_ATTN_LAUNCH(true, false, 2); break;
    case 12: B200Q_ATTN_LAUNCH(false, true, 2); break;
    default: B200Q_ATTN_LAUNCH(true, true, 2); break;
  }
#undef B200Q_ATTN_LAUNCH
  B200Q_CHECK_LAUNCH();
  return B200Q_OK;
}
