// bf16 flash attention on the 5th-gen tensor cores: the attention core of the W8A8 DiT step (BASELINE configs[1]/[3]).
//
// Replaces the reference's flash-attn call (ViDiT-Q/examples/Wan2.1/wan/modules/attention.py:94-127 =
// flash_attn_varlen_func on bf16 q,k,v with softmax_scale = head_dim^-0.5, no mask, no dropout; the SDPA fallback
// :171-178 is the same function): O = softmax(Q.K^T * scale) . V per head, head_dim = 128.
//
//   S   = Q.K^T          tcgen05.mma.kind::f16 (bf16 x bf16 -> fp32), A = Q tile and B = K tile from shared memory
//                        (K-major, SWIZZLE_128B, written by TMA), D = 128x128 fp32 in TMEM
//   P   = 2^(S*c - m)    softmax warpgroup, thread = query row = TMEM lane: tcgen05.ld S -> row max -> FFMA2 -> MUFU.EX2
//                        -> bf16x2 -> tcgen05.st back into the S columns (P aliases S)
//   O  += P.V            tcgen05.mma.kind::f16 with A = P read from TENSOR MEMORY and B = the V tile as it lies in global
//                        memory ([keys, head_dim] = MN-major, SWIZZLE_128B): no transposed copy of V, no P round trip through
//                        shared memory
//   out = O / l          after the last key block
//
// Lazy rescaling: the running maximum m only moves (and O, l are only rescaled) when a block's row maximum exceeds it by
// more than 2^8; until then P = 2^(x - m) simply exceeds 1 (bf16 has the exponent range, the fp32 accumulators the
// headroom), so in steady state no thread touches O between the first and the last key block.  O is rescaled by the row's
// own softmax thread (TMEM lane = row): legal exactly when P.V(j-1) has completed and P.V(j) has not been issued, which
// the issue order below guarantees (S(j) complete implies P.V(j-1) complete: one in-order tensor pipe).
//
// Two kernels, chosen per head on the device (b200q_attn_bf16: Cauchy-Schwarz score bound from the row-norm maxima):
//   attn_bf16_kp_kernel   bounded heads (every head of the RMS-normed Wan workload): max-free softmax, pipelined over KEY
//                         BLOCKS with three S/P buffers on CTA pairs - described in front of the kernel further down;
//   attn_bf16_kernel      the two-tile kernel described here: online softmax for unbounded heads (FAST = false), and the
//                         max-free softmax in the same structure (FAST = true, b200q_attn_bf16_set_variant(0)).
//
// Two-tile kernel: CTA = 256 queries of one head (two 128-row Q tiles), persistent over (head, query-tile-pair) items; 18 warps:
//   warps 0-7 / 8-15  softmax warps of Q tile 0 / 1: TWO threads per query row (TMEM lane), 64 key columns each - the
//                     per-tile chain S -> softmax -> P.V -> next S is the critical path, so the softmax of a tile is spread
//                     over 8 warps; the two half-row maxima / sums meet through shared memory and a named barrier
//   warp 16           TMA producer (Q tiles, K ring, V ring) + TMEM allocator
//   warp 17           MMA issuer (whole warp walks the schedule, one elected lane issues);
//                     issue order per key block j:  S0(j)  P1.V(j-1)  S1(j)  P0.V(j)
// so the tensor pipe works on one tile while the other tile's warps are in their softmax.
// TMEM (512 columns): [S0/P0 | S1/P1 | O0 | O1], 128 columns each.
#include "common.cuh"
#include "ptx.cuh"

namespace b200q {
using namespace ptx;

int make_tmap_2d(CUtensorMap* out, const void* base, CUtensorMapDataType dt, int elem_bytes, int64_t rows, int64_t cols,
                 int64_t ld, int box_rows, int box_cols, CUtensorMapSwizzle swz);

namespace fa {

constexpr int BQ = 128, BKEY = 128, HD = 128;
constexpr int HALF = 128 * 64 * 2;          // one TMA box: 128 rows x 64 bf16 = 16 KB
constexpr int TILE = 2 * HALF;              // a 128 x 128 bf16 operand tile = two boxes (head_dim halves)
constexpr int KS = 2, VS = 2;
constexpr int THREADS = 18 * 32;             // 16 softmax warps + TMA producer + MMA issuer
constexpr float RESCALE_THRESHOLD = 8.0f;   // log2 units

// CL = CTAs per cluster.  CL == 2: a CTA pair works on 512 queries of one head with tcgen05.mma.cta_group::2 (M = 256: 128
// query rows from each CTA); each CTA stages only HALF of every K tile (64 of the 128 keys) and HALF of every V tile (64 of
// the 128 head_dim columns), which halves the TMA / L2 traffic of K and V and cuts the shared-memory operand reads of a
// Q.K^T instruction from 8 KB to 6 KB per CTA (at M = N = 128 an SS instruction reads shared memory at the 128 B/clk port
// limit) and those of P.V from 4 KB to 2 KB.  The ring stages are half as large, so the pair runs four of them.
template <int CL>
struct SmemT {
  static constexpr int ks = CL == 2 ? 4 : KS, vs = CL == 2 ? 4 : VS;
  static constexpr int ktile = TILE / CL, vtile = TILE / CL;   // bytes per ring stage in this CTA
  static constexpr int khalf = HALF / CL;                      // one K box: (128 / CL) keys x 64 bf16
  static constexpr int q = 0;                       // 2 tiles
  static constexpr int k = q + 2 * TILE;            // ks stages
  static constexpr int v = k + ks * ktile;          // vs stages
  static constexpr int xch = v + vs * vtile;         // half-row statistics exchange: [tile][parity][half][row] fp32
  static constexpr int bar = xch + 2 * 2 * 2 * 128 * 4;
  static constexpr int total = bar + 256;
};
using Smem = SmemT<1>;
static_assert(SmemT<1>::total <= 232448 && SmemT<2>::total <= 232448, "dynamic smem budget (227 KB) exceeded");

struct Params {
  int Lq, Lk, H;
  __nv_bfloat16* out; long long ldo;
  float scale_log2e;
  int n_items, n_qt;
  int n_splits, bps;             // key splits per (head, query-tile pair) item and key blocks per split
  long long split_stride;        // elements between the partial outputs of consecutive splits (n_splits > 1)
  int mode;                      // scheduling knob (b200q_attn_bf16_set_mode)
  float* lse_out;                // optional [H, Lq]: log2(sum_j 2^(x_j)) per row, for key-split merges
  const float* qk_norm;          // optional [2, H]: max over rows of |q_i|^2 / |k_j|^2 of the head's slices (attn_qk_norm_kernel)
};

// Bounded-score heads: |S_ij * scale * log2(e)| <= |q_i| |k_j| scale log2(e) <= B (Cauchy-Schwarz on the per-head maxima of
// the row norms).  For B <= FAST_BOUND the exponentials 2^x need no running maximum at all: P = 2^x lies in
// [2^-80, 2^80], the fp32 row sum below 2^(80+31) and the fp32 O accumulator below 2^(80+31) |v|_max - bf16 and fp32 have
// the exponent range, and the relative precision of a float does not depend on its magnitude.  Such heads take the
// max-free kernel (no row maximum, no exchange between the two threads of a row, no rescaling, S consumed in two
// streamed halves); every other head takes the online-softmax kernel.  Both kernels walk the same item list and skip the
// heads of the other class.
constexpr float FAST_BOUND = 80.0f;
__device__ __forceinline__ bool head_is_bounded(const Params& p, int h) {
  if (p.qk_norm == nullptr) return false;
  const float b = sqrtf(p.qk_norm[h] * p.qk_norm[p.H + h]) * fabsf(p.scale_log2e);   // qk_norm holds squared norms
  return b <= FAST_BOUND;                                                 // false for NaN / inf
}
// No head of the wanted class in this launch (the usual case for one of the two kernels): every thread of every CTA sees
// the same answer, so the kernel can return before it touches barriers, tensor memory or the cluster.
__device__ __forceinline__ bool no_head_of_class(const Params& p, bool bounded) {
  for (int h = 0; h < p.H; ++h)
    if (head_is_bounded(p, h) == bounded) return false;
  return true;
}

// kind::f16 instruction descriptor: D = fp32, A = B = bf16, A K-major; B K-major (Q.K^T) or MN-major (P.V)
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N, bool b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((b_mn_major ? 1u : 0u) << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// MN-major SWIZZLE_128B operand: 64-element (128 B) rows along MN, 8-row groups (1024 B) along K, MN atoms LBO apart
// (cute/atom/mma_traits_sm100.hpp make_umma_desc<Major::MN>: ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units)
__device__ __forceinline__ uint64_t make_mnmajor_sw128_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// D[tmem] (+)= A[smem] * B[smem]^T, bf16 -> fp32
__device__ __forceinline__ void mma_bf16_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem], bf16 -> fp32
__device__ __forceinline__ void mma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// cta_group::2 forms: D = 256 x N over both CTAs' TMEM (128 lanes each), A = 128 rows from each CTA (shared memory at the
// same CTA-relative address, or each CTA's own TMEM columns), B = N/2 rows from each CTA's shared memory
__device__ __forceinline__ void mma_bf16_ss_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_bf16_ts_2sm(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier given by its shared::cluster address (mapa): the softmax warps of both CTAs of a pair report to the
// LEADER's barriers, whose MMA warp issues the pair's instructions.  RELAXED on purpose: what the arrival publishes lives in
// tensor memory and is ordered by tcgen05.wait::st / ::ld + tcgen05.fence::before_thread_sync on this side and
// tcgen05.fence::after_thread_sync on the waiting side; a release at cluster scope would put MEMBAR.ALL.GPU + ERRBAR in
// front of every arrival (and an acquire.cluster wait a CCTL.IVALL behind every wait) on the critical S -> P -> P.V chain.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
      "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
      "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]),
      "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]),
      "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
      "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

// Bounded mbarrier wait, written as one PTX loop: a waiting warp is woken by every barrier event of the CTA (ncu: 4-13
// wake-ups per wait), so the instructions per failed attempt are what a wait costs - try_wait, one add, one compare, one
// branch.  A protocol bug still surfaces as a launch failure: after 2^20 failed attempts (each suspends up to 20 us when
// nothing happens at all) the warp traps.
__device__ __forceinline__ void bar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1, P2;\n"
      ".reg .u32 cnt;\n"
      "mov.u32 cnt, 0;\n"
      "B200Q_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n"
      "@P1 bra.uni B200Q_DONE_%=;\n"
      "add.u32 cnt, cnt, 1;\n"
      "setp.lt.u32 P2, cnt, 0x100000;\n"
      "@P2 bra.uni B200Q_WAIT_%=;\n"
      "trap;\n"
      "B200Q_DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity), "r"(20000u)
      : "memory");
}

// Critical-path waits (S ready / P ready): selectable polling flavour, b200q_attn_bf16_set_mode bits 64 / 128.
//   0: try_wait with a 20 us suspend hint (fewest issue slots)   64: try_wait, default time limit   128: test_wait spin
__device__ __forceinline__ void bar_wait_crit(uint64_t* bar, uint32_t parity, int mode) {
  if ((mode & 192) == 0) { bar_wait(bar, parity); return; }
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0, spins = 0;
  while (true) {
    if (mode & 128)
      asm volatile("{\n.reg .pred P1;\nmbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\nselp.u32 %0, 1, 0, P1;\n}\n" : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    else
      asm volatile("{\n.reg .pred P1;\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\nselp.u32 %0, 1, 0, P1;\n}\n" : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    if (done) return;
    if (++spins > 400000000u) __trap();
  }
}

// Rare path of the lazy rescaling: multiply this thread's 64 columns of the O accumulator row by alpha (rolled loop: runs
// a handful of times per row).
__device__ __forceinline__ void rescale_o(uint32_t t_o, float alpha) {
  const uint64_t a2 = pack_f32x2(alpha, alpha);
#pragma unroll 1
  for (int oc = 0; oc < 4; ++oc) {
    uint32_t o[16];
    tmem_ld_32x16(t_o + oc * 16, o);
    tmem_ld_wait();
#pragma unroll
    for (int e = 0; e < 16; e += 2) {
      const uint64_t s2 = mul_f32x2(pack_u32x2(o[e], o[e + 1]), a2);
      unpack_u32x2(s2, o[e], o[e + 1]);
    }
    tmem_st_32x16(t_o + oc * 16, o);
  }
}

// 2^x for two lanes on the FMA / ALU pipes (no MUFU): Cody-Waite split x = n + r, r in [-0.5, 0.5], degree-4 minimax
// polynomial for 2^r (max relative error 7e-6, far below the bf16 rounding of P), 2^n by integer addition into the
// exponent field.  The softmax is bound by the 16 MUFU.EX2 per clock per SM; routing a quarter of the exponentials through
// this path takes the MUFU off the critical path (the technique FlashAttention-4 uses on the same hardware).
__device__ __forceinline__ uint64_t exp2_poly2(uint64_t x2) {
  const uint64_t magic2 = pack_f32x2(12582912.0f, 12582912.0f), nmagic2 = pack_f32x2(-12582912.0f, -12582912.0f);
  float x0, x1;
  unpack_f32x2(x2, x0, x1);
  const uint64_t xc = pack_f32x2(fmaxf(x0, -126.f), fmaxf(x1, -126.f));
  const uint64_t t = add_f32x2(xc, magic2);                              // integer part n = rne(x) in the low mantissa bits
  const uint64_t r = add_f32x2(xc, mul_f32x2(add_f32x2(t, nmagic2), pack_f32x2(-1.f, -1.f)));   // r = x - n
  uint64_t q = fma_f32x2(r, pack_f32x2(0.009670767933130264f, 0.009670767933130264f), pack_f32x2(0.05587553605437279f, 0.05587553605437279f));
  q = fma_f32x2(q, r, pack_f32x2(0.24022211134433746f, 0.24022211134433746f));
  q = fma_f32x2(q, r, pack_f32x2(0.6931272745132446f, 0.6931272745132446f));
  q = fma_f32x2(q, r, pack_f32x2(1.0f, 1.0f));
  uint32_t q0, q1, t0, t1;
  unpack_u32x2(q, q0, q1);
  unpack_u32x2(t, t0, t1);
  return pack_u32x2(q0 + (t0 << 23), q1 + (t1 << 23));                   // (magic + n) << 23 == n << 23 (mod 2^32)
}

// P = 2^(S*c - m) for one 16-key chunk: bf16x2 codes into pk[0..7], fp32 row sum into sum2.  POLY pairs of the 8 go through
// the polynomial, the rest through MUFU.EX2.
template <int POLY>
__device__ __forceinline__ void exp_chunk(const uint32_t (&sc)[16], uint32_t* pk, uint64_t c2, uint64_t nm2, uint64_t& sum2) {
#pragma unroll
  for (int e = 0; e < 16; e += 2) {
    const uint64_t x2 = fma_f32x2(pack_u32x2(sc[e], sc[e + 1]), c2, nm2);
    uint64_t p2;
    if ((e >> 1) < POLY) {
      p2 = exp2_poly2(x2);
    } else {
      float x0, x1;
      unpack_f32x2(x2, x0, x1);
      p2 = pack_f32x2(ex2f(x0), ex2f(x1));
    }
    sum2 = add_f32x2(sum2, p2);
    float p0, p1;
    unpack_f32x2(p2, p0, p1);
    pk[e >> 1] = pack_bf16x2(p0, p1);
  }
}

// Max-free variant (bounded heads): P = 2^(S*c) for one 16-key chunk, |S*c| <= FAST_BOUND.  POLY of the 8 pairs take a
// degree-3 polynomial on the FMA pipe (max relative error 7.5e-5, bf16 rounds P to 3.9e-3): t = S*c + 1.5*2^23 puts
// n = rne(S*c) into the low mantissa bits, r = S*c - n by a second FFMA2, 2^n by an integer multiply-add into the exponent
// field - 8 issue slots per pair against 3 (FFMA2 + 2 MUFU) on the MUFU path, and no clamp because the range is known.
template <int POLY>
__device__ __forceinline__ void exp_chunk_fast(const uint32_t (&sc)[16], uint32_t* pk, uint64_t c2, uint64_t& sum2) {
  const uint64_t magic2 = pack_f32x2(12582912.0f, 12582912.0f);
#pragma unroll
  for (int e = 0; e < 16; e += 2) {
    const uint64_t s2 = pack_u32x2(sc[e], sc[e + 1]);
    uint64_t p2;
    // the polynomial pairs are spread over the chunk so that neither pipe sees a burst
    constexpr uint32_t kPolyMask[9] = {0x00, 0x01, 0x11, 0x49, 0x55, 0x57, 0x77, 0x7F, 0xFF};
    if ((kPolyMask[POLY] >> (e >> 1)) & 1u) {
      const uint64_t t = fma_f32x2(s2, c2, magic2);
      const uint64_t nn = fma_f32x2(t, pack_f32x2(-1.f, -1.f), magic2);                      // -n
      const uint64_t r = fma_f32x2(s2, c2, nn);                                              // r = S*c - n in [-0.5, 0.5]
      uint64_t q = fma_f32x2(r, pack_f32x2(0.0551716685295105f, 0.0551716685295105f), pack_f32x2(0.2426111251115799f, 0.2426111251115799f));
      q = fma_f32x2(q, r, pack_f32x2(0.6932609677314758f, 0.6932609677314758f));
      q = fma_f32x2(q, r, pack_f32x2(0.9999280571937561f, 0.9999280571937561f));
      uint32_t q0, q1, t0, t1;
      unpack_u32x2(q, q0, q1);
      unpack_u32x2(t, t0, t1);
      p2 = pack_u32x2(q0 + (t0 << 23), q1 + (t1 << 23));
    } else {
      const uint64_t x2 = mul_f32x2(s2, c2);
      float x0, x1;
      unpack_f32x2(x2, x0, x1);
      p2 = pack_f32x2(ex2f(x0), ex2f(x1));
    }
    sum2 = add_f32x2(sum2, p2);
    float p0, p1;
    unpack_f32x2(p2, p0, p1);
    pk[e >> 1] = pack_bf16x2(p0, p1);
  }
}

__device__ __forceinline__ float max3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));      // FMNMX3: one issue slot for two comparisons
  return d;
}
__device__ __forceinline__ float max16(const uint32_t (&sc)[16], float m) {
  float m0 = m, m1 = m;
#pragma unroll
  for (int e = 0; e < 16; e += 4) {
    m0 = max3(m0, __uint_as_float(sc[e]), __uint_as_float(sc[e + 1]));
    m1 = max3(m1, __uint_as_float(sc[e + 2]), __uint_as_float(sc[e + 3]));
  }
  return fmaxf(m0, m1);
}

// Diagnostic build (-DB200Q_FA_TRACE): CTA 0 prints, per warp role, the cycles spent in each phase of its first item.
#ifdef B200Q_FA_TRACE
__device__ __forceinline__ long long tr_clock() { long long c; asm volatile("mov.u64 %0, %%clock64;" : "=l"(c)::"memory"); return c; }
#define TR_DECL long long tr_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}
#define TR(i) do { if (j == 100) tr_acc[i] = tr_clock(); } while (0)     /* absolute time stamps of key block 100 */
#else
#define TR_DECL
#define TR(i)
#endif

template <int POLY, bool FAST, int CL>
__global__ void __launch_bounds__(THREADS, 1)
attn_bf16_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                 const __grid_constant__ CUtensorMap tm_v, const Params p) {
  using Smem = SmemT<CL>;
  constexpr int KS = Smem::ks, VS = Smem::vs;
  extern __shared__ __align__(1024) uint8_t smem[];
  if (no_head_of_class(p, FAST)) return;
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
  uint64_t* s_full = v_empty + VS;         // [tile]: S(j) complete in TMEM
  uint64_t* p_full = s_full + 2;           // [tile]: P(j) written to TMEM (and S(j) consumed)
  uint64_t* o_full = p_full + 2;           // [tile]: O accumulator of the item complete
  uint64_t* o_free = o_full + 2;           // [tile]: O read out by the warpgroup
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_free + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nb_all = (p.Lk + BKEY - 1) / BKEY;
  // item -> (head, key split, group of 2 * CL query tiles); a split covers key blocks [kb0, kb0 + nb).  Items are walked by
  // clusters; CTA `rank` of a pair owns query rows [rank * 256, rank * 256 + 256) of the item's 512.
  const int per_head = p.n_qt * p.n_splits;
  const int rank = (CL == 2) ? (int)cluster_ctarank() : 0;
  const int item0 = (int)blockIdx.x / CL, item_step = (int)gridDim.x / CL;
  constexpr uint16_t kPair = 3;

  if (threadIdx.x == 0) {
    mbar_init(q_full, 1); mbar_init(q_empty, 1);
    for (int i = 0; i < KS; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); }
    for (int i = 0; i < VS; ++i) { mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {                      // CL == 2: both CTAs' softmax warps report to the leader
      mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 8 * CL);
      mbar_init(&o_full[i], 1); mbar_init(&o_free[i], 8 * CL);
    }
    fence_barrier_init();
  }
  constexpr int W_TMA = 16, W_MMA = 17;
  if (warp == W_TMA && lane == 0) { prefetch_tmap(&tm_q); prefetch_tmap(&tm_k); prefetch_tmap(&tm_v); }
  if (warp == W_TMA) { if (CL == 2) tmem_alloc_2sm<512>(tmem_slot); else tmem_alloc<512>(tmem_slot); }
  tcgen05_fence_before();
  __syncthreads();
  if (CL == 2) cluster_sync();                           // the peer's barriers exist before anyone signals them
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == W_TMA) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int ks = 0, vs = 0; uint32_t kph = 0, vph = 0; int it = 0;
      for (int item = item0; item < p.n_items; item += item_step) {
        const int h = item / per_head, sp = (item % per_head) / p.n_qt, q0 = (item % p.n_qt) * (2 * BQ * CL) + rank * (2 * BQ);
        if (head_is_bounded(p, h) != FAST) continue;                      // the other kernel's head
        const int kb0 = sp * p.bps, nb = min(p.bps, nb_all - kb0);
        bar_wait(q_empty, (it & 1) ^ 1);
        ++it;
        if (CL == 2) {
          // the leader's MMA warp reads both CTAs' operands: every byte of the pair is credited to the LEADER's barrier
          const uint32_t lead = mapa_u32(q_full, 0);
          if (rank == 0) mbar_expect_tx(q_full, 4 * TILE);
#pragma unroll
          for (int t = 0; t < 2; ++t)
#pragma unroll
            for (int hf = 0; hf < 2; ++hf)
              tma_load_2d_2sm(smem + Smem::q + t * TILE + hf * HALF, &tm_q, lead, h * HD + hf * 64, q0 + t * BQ);
        } else {
          mbar_expect_tx(q_full, 2 * TILE);
#pragma unroll
          for (int t = 0; t < 2; ++t)
#pragma unroll
            for (int hf = 0; hf < 2; ++hf)
              tma_load_2d(smem + Smem::q + t * TILE + hf * HALF, &tm_q, q_full, h * HD + hf * 64, q0 + t * BQ);
        }
        for (int j = 0; j < nb; ++j) {
          bar_wait(&k_empty[ks], kph ^ 1);
          if (CL == 2) {
            // my half of the key block: keys [rank * 64, rank * 64 + 64), all 128 head_dim columns (two 64 x 64 boxes)
            const uint32_t lead = mapa_u32(&k_full[ks], 0);
            if (rank == 0) mbar_expect_tx(&k_full[ks], TILE);
            tma_load_2d_2sm(smem + Smem::k + ks * Smem::ktile, &tm_k, lead, h * HD, (kb0 + j) * BKEY + rank * 64);
            tma_load_2d_2sm(smem + Smem::k + ks * Smem::ktile + Smem::khalf, &tm_k, lead, h * HD + 64, (kb0 + j) * BKEY + rank * 64);
          } else {
            mbar_expect_tx(&k_full[ks], TILE);
            tma_load_2d(smem + Smem::k + ks * TILE, &tm_k, &k_full[ks], h * HD, (kb0 + j) * BKEY);
            tma_load_2d(smem + Smem::k + ks * TILE + HALF, &tm_k, &k_full[ks], h * HD + 64, (kb0 + j) * BKEY);
          }
          if (++ks == KS) { ks = 0; kph ^= 1; }
          bar_wait(&v_empty[vs], vph ^ 1);
          if (CL == 2) {
            // my half of the value block: all 128 keys, head_dim columns [rank * 64, rank * 64 + 64) (one 128 x 64 box)
            const uint32_t lead = mapa_u32(&v_full[vs], 0);
            if (rank == 0) mbar_expect_tx(&v_full[vs], TILE);
            tma_load_2d_2sm(smem + Smem::v + vs * Smem::vtile, &tm_v, lead, h * HD + rank * 64, (kb0 + j) * BKEY);
          } else {
            mbar_expect_tx(&v_full[vs], TILE);
            tma_load_2d(smem + Smem::v + vs * TILE, &tm_v, &v_full[vs], h * HD, (kb0 + j) * BKEY);
            tma_load_2d(smem + Smem::v + vs * TILE + HALF, &tm_v, &v_full[vs], h * HD + 64, (kb0 + j) * BKEY);
          }
          if (++vs == VS) { vs = 0; vph ^= 1; }
        }
      }
    }
  } else if (warp == W_MMA) {
    if (CL == 1 || rank == 0) {
    // ===================== MMA issuer (CL == 2: the leader CTA issues for the pair) =====================
    // The whole warp walks the schedule (uniform control flow keeps descriptors and addresses in uniform registers); one
    // elected lane issues the tcgen05 instructions.
    constexpr uint32_t idesc_qk = idesc_bf16(BQ * CL, BKEY, false);
    constexpr uint32_t idesc_pv = idesc_bf16(BQ * CL, HD, true);
    int ks = 0, vs = 0; uint32_t kph = 0, vph = 0; int it = 0;
    uint32_t pcnt0 = 0, pcnt1 = 0;                    // P tiles consumed so far per Q tile
    const uint64_t qd0 = make_kmajor_sw128_desc(smem_u32(smem + Smem::q));
    const uint64_t qd1 = make_kmajor_sw128_desc(smem_u32(smem + Smem::q + TILE));
    const uint64_t kd = make_kmajor_sw128_desc(smem_u32(smem + Smem::k));
    const uint64_t vd = make_mnmajor_sw128_desc(smem_u32(smem + Smem::v), HALF, 1024);
    // S[t] = Q_t . K(stage)^T : 8 x (M128, N128, K16); k-step kk: head_dim half kk / 4, 32 bytes per step inside the 128 B row
    auto qk = [&](int t, int stage) {
      const uint64_t a0 = t ? qd1 : qd0;
      const uint64_t b0 = kd + (uint64_t)(stage * (Smem::ktile >> 4));
      const uint32_t d = tmem_base + t * 128;
      if (elect_one()) {
        if (!(p.mode & 16)) {
#pragma unroll
          for (int kk = 0; kk < HD / 16; ++kk) {
            const uint64_t off = (uint64_t)((kk >> 2) * (HALF >> 4) + (kk & 3) * 2);
            const uint64_t boff = (uint64_t)((kk >> 2) * (Smem::khalf >> 4) + (kk & 3) * 2);
            if (CL == 2) mma_bf16_ss_2sm(d, a0 + off, b0 + boff, idesc_qk, kk != 0 ? 1u : 0u);
            else mma_bf16_ss(d, a0 + off, b0 + boff, idesc_qk, kk != 0 ? 1u : 0u);
          }
        }
        if (CL == 2) mma_commit_2sm_mc(&s_full[t], kPair); else mma_commit(&s_full[t]);
      }
      __syncwarp();
    };
    // O[t] (+)= P_t (TMEM, bf16 packed in the S columns) . V(stage) : 8 x (M128, N128, K16); 16 keys = 2048 B of the V tile
    auto pv = [&](int t, int stage, bool first) {
      if (t == 0) { bar_wait_crit(&p_full[0], pcnt0 & 1, p.mode); ++pcnt0; } else { bar_wait_crit(&p_full[1], pcnt1 & 1, p.mode); ++pcnt1; }
      tcgen05_fence_after();
      const uint64_t b0 = vd + (uint64_t)(stage * (Smem::vtile >> 4));
      const uint32_t d = tmem_base + 256 + t * 128;
      const uint32_t pa = tmem_base + t * 128;
      if (elect_one() && !(p.mode & 16)) {
        // P columns: the online-softmax kernel packs the row's 128 keys into S columns 0..63; the max-free kernel leaves
        // each half-row thread's 64 keys in the first 32 of ITS OWN 64 S columns (no cross-thread hazard, no barrier)
#pragma unroll
        for (int kk = 0; kk < BKEY / 16; ++kk) {
          const uint32_t a = pa + (FAST ? (kk >> 2) * 64 + (kk & 3) * 8 : kk * 8);
          if (CL == 2) mma_bf16_ts_2sm(d, a, b0 + (uint64_t)(kk * (2048 >> 4)), idesc_pv, (first && kk == 0) ? 0u : 1u);
          else mma_bf16_ts(d, a, b0 + (uint64_t)(kk * (2048 >> 4)), idesc_pv, (first && kk == 0) ? 0u : 1u);
        }
      }
      __syncwarp();
    };
    auto commit = [&](uint64_t* bar) {
      if (elect_one()) { if (CL == 2) mma_commit_2sm_mc(bar, kPair); else mma_commit(bar); }
      __syncwarp();
    };
    auto wait_o_free = [&](int t, uint32_t parity) { bar_wait(&o_free[t], parity); };
    auto next_k = [&]() { commit(&k_empty[ks]); if (++ks == KS) { ks = 0; kph ^= 1; } };
    auto next_v = [&]() { commit(&v_empty[vs]); if (++vs == VS) { vs = 0; vph ^= 1; } };
    for (int item = item0; item < p.n_items; item += item_step, ++it) {
      if (head_is_bounded(p, item / per_head) != FAST) { --it; continue; }
      const int nb = min(p.bps, nb_all - ((item % per_head) / p.n_qt) * p.bps);
      TR_DECL;
      bar_wait(q_full, it & 1);
      bar_wait(&k_full[ks], kph);
      tcgen05_fence_after();
      qk(0, ks);
      qk(1, ks);
      next_k();
      wait_o_free(0, it & 1);                               // previous item's O0 has been read out
      bar_wait(&v_full[vs], vph);
      pv(0, vs, true);
      for (int j = 1; j < nb; ++j) {
        TR(0);
        bar_wait(&k_full[ks], kph);
        TR(1);
        tcgen05_fence_after();
        qk(0, ks);                                           // S0(j): in order behind P0.V(j-1), which read P0 = S0's columns
        if (j == 1) wait_o_free(1, it & 1);
        TR(2);
        pv(1, vs, j == 1);                                   // P1.V(j-1)
        TR(3);
        next_v();
        qk(1, ks);                                           // S1(j)
        next_k();
        TR(4);
        bar_wait(&v_full[vs], vph);
        TR(5);
        pv(0, vs, false);                                    // P0.V(j)
        TR(6);
      }
#ifdef B200Q_FA_TRACE
      if (blockIdx.x == 0 && it == 0 && lane == 0)
        printf("MMA  j=100 stamps: top %lld  k_ready %lld  S0_issued %lld  PV1_issued %lld  S1_issued %lld  v_ready %lld  PV0_issued %lld\n",
               tr_acc[0] % 10000000, tr_acc[1] % 10000000, tr_acc[2] % 10000000, tr_acc[3] % 10000000, tr_acc[4] % 10000000, tr_acc[5] % 10000000, tr_acc[6] % 10000000);
#endif
      commit(&o_full[0]);
      commit(q_empty);                                       // last read of the Q tiles was S1(nb-1)
      if (nb == 1) wait_o_free(1, it & 1);
      pv(1, vs, nb == 1);                                    // P1.V(nb-1)
      next_v();
      commit(&o_full[1]);
    }
    }
  } else {
    // ===================== softmax warps: thread = (query row, half of the key block) =====================
    const int t = warp >> 3, half = (warp >> 2) & 1, quarter = warp & 3;
    const int r = quarter * 32 + lane;                                   // row inside the Q tile == TMEM lane
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const uint32_t t_s = tmem_base + t * 128 + half * 64 + lane_off;     // my 64 S columns
    const uint32_t t_p = tmem_base + t * 128 + half * 32 + lane_off;     // my 32 P columns (bf16x2; P aliases S columns 0..63)
    const uint32_t t_o = tmem_base + 256 + t * 128 + half * 64 + lane_off;   // my 64 O columns
    float* xch = reinterpret_cast<float*>(smem + Smem::xch) + t * 512;   // [parity][half][row]
    uint32_t scnt = 0, itn = 0, xp = 0;
    const float c = p.scale_log2e;
    const uint64_t c2 = pack_f32x2(c, c);
    const int tail = p.Lk - (nb_all - 1) * BKEY;                         // valid keys in the last block (1..128)
    // where this warp reports "P written" / "O read out": its own CTA's barriers, or (CL == 2) the leader's
    const uint32_t pf_addr = CL == 2 ? mapa_u32(&p_full[t], 0) : 0u, of_addr = CL == 2 ? mapa_u32(&o_free[t], 0) : 0u;
    auto arrive_p = [&]() { if (CL == 2) mbar_arrive_cluster(pf_addr); else mbar_arrive(&p_full[t]); };
    auto arrive_o = [&]() { if (CL == 2) mbar_arrive_cluster(of_addr); else mbar_arrive(&o_free[t]); };
    {                                                                    // O columns start out free
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) arrive_o();
    }
    for (int item = item0; item < p.n_items; item += item_step, ++itn) {
      const int h = item / per_head, sp = (item % per_head) / p.n_qt;
      const int q0 = (item % p.n_qt) * (2 * BQ * CL) + rank * (2 * BQ) + t * BQ;
      if (head_is_bounded(p, h) != FAST) { --itn; continue; }
      const int kb0 = sp * p.bps, nb = min(p.bps, nb_all - kb0);
      const int row = q0 + r;
      const bool row_ok = row < p.Lq;
      float m_ref = FAST ? 0.f : -INFINITY;                              // reference maximum of the exponentials (log2 units)
      uint64_t sum2 = pack_f32x2(0.f, 0.f);
      TR_DECL;
      for (int j = 0; FAST && j < nb; ++j) {
        // ---- max-free softmax of a bounded head: P = 2^(S*c), streamed in two 32-column halves ----
        TR(0);
        bar_wait_crit(&s_full[t], scnt & 1, p.mode);
        TR(1);
        ++scnt;
        tcgen05_fence_after();
        const uint32_t t_pf = t_s;                                       // P goes into the first 32 of my own 64 S columns
        uint32_t sa[2][16], sb[2][16];
        tmem_ld_32x16(t_s, sa[0]);
        tmem_ld_32x16(t_s + 16, sa[1]);
        tmem_ld_wait();
        tmem_ld_32x16(t_s + 32, sb[0]);                                  // in flight during the first half's exponentials
        tmem_ld_32x16(t_s + 48, sb[1]);
        TR(2);
        uint32_t pk[16];
        if (kb0 + j == nb_all - 1 && tail < BKEY) {                      // last, partial key block: -inf -> MUFU path gives P = 0
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            if (half * 64 + e >= tail) sa[0][e] = 0xff800000u;
            if (half * 64 + 16 + e >= tail) sa[1][e] = 0xff800000u;
            if (half * 64 + 32 + e >= tail) sb[0][e] = 0xff800000u;
            if (half * 64 + 48 + e >= tail) sb[1][e] = 0xff800000u;
          }
          exp_chunk_fast<0>(sa[0], pk, c2, sum2);
          exp_chunk_fast<0>(sa[1], pk + 8, c2, sum2);
          tmem_st_32x16(t_pf, pk);
          exp_chunk_fast<0>(sb[0], pk, c2, sum2);
          exp_chunk_fast<0>(sb[1], pk + 8, c2, sum2);
          tmem_st_32x16(t_pf + 16, pk);
        } else {
          exp_chunk_fast<POLY>(sa[0], pk, c2, sum2);
          exp_chunk_fast<POLY>(sa[1], pk + 8, c2, sum2);
          tmem_st_32x16(t_pf, pk);
          tmem_ld_wait();
          TR(3);
          exp_chunk_fast<POLY>(sb[0], pk, c2, sum2);
          exp_chunk_fast<POLY>(sb[1], pk + 8, c2, sum2);
          tmem_st_32x16(t_pf + 16, pk);
        }
        TR(5);
        tmem_st_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) arrive_p();
        TR(6);
      }
      for (int j = 0; !FAST && j < nb; ++j) {
        TR(0);
        bar_wait_crit(&s_full[t], scnt & 1, p.mode);
        TR(1);
        ++scnt;
        tcgen05_fence_after();
        if (p.mode & 8) {                                                // diagnostic: tensor / TMA pipeline alone
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) arrive_p();
          continue;
        }
        uint32_t sr[4][16];                                              // my half of the S row
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) tmem_ld_32x16(t_s + ch * 16, sr[ch]);
        tmem_ld_wait();
        TR(2);
        if (kb0 + j == nb_all - 1 && tail < BKEY) {                      // last, partial key block
#pragma unroll
          for (int ch = 0; ch < 4; ++ch)
#pragma unroll
            for (int e = 0; e < 16; ++e) if (half * 64 + ch * 16 + e >= tail) sr[ch][e] = 0xff800000u;   // -inf: P = 0
        }
        float hmax = __uint_as_float(sr[0][0]);
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) hmax = max16(sr[ch], hmax);
        // the two halves of a row agree on the block maximum; the barrier also orders "every S column of the tile is in
        // registers" before "any P column (which aliases S columns 0..63) is written"
        xch[xp * 256 + half * 128 + r] = hmax;
        TR(3);
        named_bar_sync(1 + t, 256);
        const float bm = fmaxf(hmax, xch[xp * 256 + (half ^ 1) * 128 + r]) * c;   // c > 0
        TR(4);
        xp ^= 1;
        const bool need = bm > m_ref + RESCALE_THRESHOLD;                // first block of an item: m_ref = -inf
        if (__any_sync(0xffffffffu, need)) {
          // ---- rare path: move the reference maximum; l and O accumulated against the old one are rescaled ----
          const float m_new = need ? bm : m_ref;
          const float alpha = ex2f(m_ref - m_new);                       // 0 from -inf, exactly 1 for rows that keep m_ref
          sum2 = mul_f32x2(sum2, pack_f32x2(alpha, alpha));
          if (j > 0) rescale_o(t_o, alpha);                              // S(j) complete implies P.V(j-1) complete
          m_ref = m_new;
        }
        const float nm = -m_ref;
        const uint64_t nm2 = pack_f32x2(nm, nm);
        if (p.mode & 32) {                                               // diagnostic: TMEM load + row maximum + exchange only
          if (nm == 12345.f) tmem_st_32x16(t_p, sr[0]);
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) arrive_p();
          continue;
        }
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          uint32_t pk[16];                                               // bf16x2 P of 32 keys = 16 TMEM columns
          exp_chunk<POLY>(sr[2 * g], pk, c2, nm2, sum2);
          exp_chunk<POLY>(sr[2 * g + 1], pk + 8, c2, nm2, sum2);
          tmem_st_32x16(t_p + g * 16, pk);
        }
        TR(5);
        tmem_st_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) arrive_p();
        TR(6);
      }
#ifdef B200Q_FA_TRACE
      if (blockIdx.x == 0 && itn == 0 && lane == 0 && (warp & 3) == 0)
        printf("SM w%02d j=100 stamps: top %lld  S_ready %lld  ld_done %lld  t3 %lld  t4 %lld  exp_issued %lld  arrived %lld\n",
               warp, tr_acc[0] % 10000000, tr_acc[1] % 10000000, tr_acc[2] % 10000000, tr_acc[3] % 10000000, tr_acc[4] % 10000000, tr_acc[5] % 10000000, tr_acc[6] % 10000000);
#endif

      // ---- read-out: out = O / l ----
      float s0, s1;
      unpack_f32x2(sum2, s0, s1);
      xch[xp * 256 + half * 128 + r] = s0 + s1;
      named_bar_sync(1 + t, 256);
      const float l = (s0 + s1) + xch[xp * 256 + (half ^ 1) * 128 + r];
      xp ^= 1;
      const float inv = 1.0f / l;
      bar_wait(&o_full[t], itn & 1);
      tcgen05_fence_after();
      if (half == 0 && row_ok && p.lse_out != nullptr) p.lse_out[((long long)sp * p.H + h) * p.Lq + row] = m_ref + log2f(l);
      __nv_bfloat16* g = p.out + sp * p.split_stride + (long long)(row_ok ? row : 0) * p.ldo + h * HD + half * 64;
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        uint32_t o[16];
        tmem_ld_32x16(t_o + ch * 16, o);
        tmem_ld_wait();
        if (ch == 3) {                                                   // O is in registers: hand the columns back
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) arrive_o();
        }
        if (row_ok) {
#pragma unroll
          for (int e = 0; e < 16; e += 8) {
            uint4 w;
            w.x = pack_bf16x2(__uint_as_float(o[e]) * inv, __uint_as_float(o[e + 1]) * inv);
            w.y = pack_bf16x2(__uint_as_float(o[e + 2]) * inv, __uint_as_float(o[e + 3]) * inv);
            w.z = pack_bf16x2(__uint_as_float(o[e + 4]) * inv, __uint_as_float(o[e + 5]) * inv);
            w.w = pack_bf16x2(__uint_as_float(o[e + 6]) * inv, __uint_as_float(o[e + 7]) * inv);
            *reinterpret_cast<uint4*>(g + ch * 16 + e) = w;
          }
        }
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (CL == 2) cluster_sync();                           // the peer may still signal this CTA's barriers / read its operands
  tcgen05_fence_after();
  if (warp == W_TMA) { if (CL == 2) tmem_dealloc_2sm<512>(tmem_base); else tmem_dealloc<512>(tmem_base); }
}

// ---------------------------------------------------------------------------------------------------------------------
// Key-pipelined kernel for bounded heads (max-free softmax), CTA pairs.
//
// The two-tile kernel above runs each Q tile through the chain  S(j) -> softmax(j) -> P.V(j) -> S(j+1): P aliases S, so the
// next Q.K^T of a tile cannot be issued before the tile's P.V, and the tensor pipe idles whenever softmax + handshakes of
// one tile outlast the other tile's two products (ncu: tensor pipe 78 % active; softmax warps issue 47 % of the cycles:
// neither side is saturated, each waits for the other).  Without a running maximum nothing in the softmax of key block
// j+1 depends on block j, so this kernel pipelines over KEY BLOCKS instead of Q tiles and keeps THREE of them in flight:
//   * one 128-row Q tile per CTA (256 queries per pair), three S/P buffers in TMEM, key block j lives in buffer j mod 3;
//   * the MMA warp issues  S(0) S(1) S(2) | P.V(0) S(3) | P.V(1) S(4) | ... : P.V(j) is issued two Q.K^T periods after S(j)
//     completed, so a softmax group has two full (Q.K^T + P.V) periods for a block before the tensor pipe could wait for it;
//   * softmax group g (8 warps, two threads per row) takes the key blocks j = g mod 2, whatever buffer they are in;
//   * each CTA stages its own Q tile, half of every K tile (64 keys) and half of every V tile (64 head_dim columns).
// TMEM columns: [S0/P0 | S1/P1 | S2/P2 | O], 128 each.
// ---------------------------------------------------------------------------------------------------------------------
struct SmemKP {
  static constexpr int ks = 5, vs = 5;
  static constexpr int ktile = TILE / 2, vtile = TILE / 2, khalf = HALF / 2;
  static constexpr int q = 0;                          // one 128 x 128 tile
  static constexpr int k = q + TILE;
  static constexpr int v = k + ks * ktile;
  static constexpr int xch = v + vs * vtile;           // [parity][group * 2 + half][row] fp32 partial row sums
  static constexpr int bar = xch + 2 * 4 * 128 * 4;
  static constexpr int total = bar + 512;
};
static_assert(SmemKP::total <= 232448, "dynamic smem budget (227 KB) exceeded");

template <int POLY>
__global__ void __launch_bounds__(THREADS, 1)
attn_bf16_kp_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                    const __grid_constant__ CUtensorMap tm_v, const Params p) {
  using Smem = SmemKP;
  constexpr int KS = Smem::ks, VS = Smem::vs;
  extern __shared__ __align__(1024) uint8_t smem[];
  if (no_head_of_class(p, true)) return;
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) {
    printf("b200q: dynamic smem base not 1024-byte aligned\n");
    __trap();
  }
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Smem::bar);
  uint64_t* k_full = bars;                 // KS   (leader: bytes of both CTAs' halves)
  uint64_t* k_empty = k_full + KS;         // KS   (multicast commit)
  uint64_t* v_full = k_empty + KS;         // VS
  uint64_t* v_empty = v_full + VS;         // VS
  uint64_t* s_full = v_empty + VS;         // [buffer]: S(j) complete in TMEM (multicast commit)
  uint64_t* p_full = s_full + 3;           // [buffer]: P(j) written by both CTAs' softmax group (leader)
  uint64_t* q_full = p_full + 3;           // Q tiles of the item landed in both CTAs (leader)
  uint64_t* q_empty = q_full + 1;          // last Q.K^T of the item complete (multicast commit)
  uint64_t* o_full = q_empty + 1;          // O accumulator of the item complete (multicast commit)
  uint64_t* o_free = o_full + 1;           // O read out by both CTAs (leader)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_free + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nb_all = (p.Lk + BKEY - 1) / BKEY;
  const int per_head = p.n_qt * p.n_splits;              // item -> (head, key split, 256-query group); rank -> its 128 rows
  const int rank = (int)cluster_ctarank();
  const int item0 = (int)blockIdx.x / 2, item_step = (int)gridDim.x / 2;
  constexpr uint16_t kPair = 3;

  if (threadIdx.x == 0) {
    for (int i = 0; i < KS; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); }
    for (int i = 0; i < VS; ++i) { mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1); }
    for (int i = 0; i < 3; ++i) { mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 16); }
    mbar_init(q_full, 1); mbar_init(q_empty, 1);
    mbar_init(o_full, 1); mbar_init(o_free, 32);
    fence_barrier_init();
  }
  constexpr int W_TMA = 16, W_MMA = 17;
  if (warp == W_TMA && lane == 0) { prefetch_tmap(&tm_q); prefetch_tmap(&tm_k); prefetch_tmap(&tm_v); }
  if (warp == W_TMA) tmem_alloc_2sm<512>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t COL_O = 384;

  if (warp == W_TMA) {
    // ===================== TMA producer: my Q tile, my half of every K tile (64 keys) and V tile (64 head_dim columns) ==========
    if (lane == 0) {
      int ks = 0, vs = 0; uint32_t kph = 0, vph = 0; int it = 0;
      for (int item = item0; item < p.n_items; item += item_step) {
        const int h = item / per_head, sp = (item % per_head) / p.n_qt, q0 = (item % p.n_qt) * (2 * BQ) + rank * BQ;
        if (!head_is_bounded(p, h)) continue;                             // the online-softmax kernel's head
        const int kb0 = sp * p.bps, nb = min(p.bps, nb_all - kb0);
        bar_wait(q_empty, (it & 1) ^ 1);
        ++it;
        {
          const uint32_t lead = mapa_u32(q_full, 0);
          if (rank == 0) mbar_expect_tx(q_full, 2 * TILE);
          tma_load_2d_2sm(smem + Smem::q, &tm_q, lead, h * HD, q0);
          tma_load_2d_2sm(smem + Smem::q + HALF, &tm_q, lead, h * HD + 64, q0);
        }
        for (int j = 0; j < nb; ++j) {
          bar_wait(&k_empty[ks], kph ^ 1);
          {
            const uint32_t lead = mapa_u32(&k_full[ks], 0);
            if (rank == 0) mbar_expect_tx(&k_full[ks], TILE);
            tma_load_2d_2sm(smem + Smem::k + ks * Smem::ktile, &tm_k, lead, h * HD, (kb0 + j) * BKEY + rank * 64);
            tma_load_2d_2sm(smem + Smem::k + ks * Smem::ktile + Smem::khalf, &tm_k, lead, h * HD + 64, (kb0 + j) * BKEY + rank * 64);
          }
          if (++ks == KS) { ks = 0; kph ^= 1; }
          bar_wait(&v_empty[vs], vph ^ 1);
          {
            const uint32_t lead = mapa_u32(&v_full[vs], 0);
            if (rank == 0) mbar_expect_tx(&v_full[vs], TILE);
            tma_load_2d_2sm(smem + Smem::v + vs * Smem::vtile, &tm_v, lead, h * HD + rank * 64, (kb0 + j) * BKEY);
          }
          if (++vs == VS) { vs = 0; vph ^= 1; }
        }
      }
    }
  } else if (warp == W_MMA) {
    if (rank == 0) {
      // ===================== MMA issuer (leader CTA, uniform control flow, one elected lane issues) =====================
      constexpr uint32_t idesc_qk = idesc_bf16(2 * BQ, BKEY, false);
      constexpr uint32_t idesc_pv = idesc_bf16(2 * BQ, HD, true);
      int ks = 0, vs = 0; uint32_t kph = 0, vph = 0; int it = 0;
      uint32_t ppar = 0;                                     // bit b: parity of p_full[b]'s next phase
      const uint64_t qd = make_kmajor_sw128_desc(smem_u32(smem + Smem::q));
      const uint64_t kd = make_kmajor_sw128_desc(smem_u32(smem + Smem::k));
      const uint64_t vd = make_mnmajor_sw128_desc(smem_u32(smem + Smem::v), HALF, 1024);
      auto commit = [&](uint64_t* bar) {
        if (elect_one()) mma_commit_2sm_mc(bar, kPair);
        __syncwarp();
      };
      // S[buf] = Q . K(stage)^T : 8 x (M256, N128, K16)
      auto qk = [&](int buf) {
        bar_wait(&k_full[ks], kph);
        tcgen05_fence_after();
        const uint64_t b0 = kd + (uint64_t)(ks * (Smem::ktile >> 4));
        const uint32_t d = tmem_base + buf * 128;
        if (elect_one()) {
          if (!(p.mode & 16)) {
#pragma unroll
            for (int kk = 0; kk < HD / 16; ++kk) {
              const uint64_t aoff = (uint64_t)((kk >> 2) * (HALF >> 4) + (kk & 3) * 2);
              const uint64_t boff = (uint64_t)((kk >> 2) * (Smem::khalf >> 4) + (kk & 3) * 2);
              mma_bf16_ss_2sm(d, qd + aoff, b0 + boff, idesc_qk, kk != 0 ? 1u : 0u);
            }
          }
          mma_commit_2sm_mc(&s_full[buf], kPair);
          mma_commit_2sm_mc(&k_empty[ks], kPair);
        }
        __syncwarp();
        if (++ks == KS) { ks = 0; kph ^= 1; }
      };
      // O (+)= P[buf] (TMEM; each half-row thread's 64 keys in the first 32 of its own 64 S columns) . V(stage)
      auto pv = [&](int buf, bool first) {
        bar_wait_crit(&p_full[buf], (ppar >> buf) & 1u, p.mode);
        ppar ^= 1u << buf;
        bar_wait(&v_full[vs], vph);
        tcgen05_fence_after();
        const uint64_t b0 = vd + (uint64_t)(vs * (Smem::vtile >> 4));
        const uint32_t d = tmem_base + COL_O, pa = tmem_base + buf * 128;
        if (elect_one()) {
          if (!(p.mode & 16)) {
#pragma unroll
            for (int kk = 0; kk < BKEY / 16; ++kk)
              mma_bf16_ts_2sm(d, pa + (kk >> 2) * 64 + (kk & 3) * 8, b0 + (uint64_t)(kk * (2048 >> 4)), idesc_pv, (first && kk == 0) ? 0u : 1u);
          }
          mma_commit_2sm_mc(&v_empty[vs], kPair);
        }
        __syncwarp();
        if (++vs == VS) { vs = 0; vph ^= 1; }
      };
      for (int item = item0; item < p.n_items; item += item_step, ++it) {
        if (!head_is_bounded(p, item / per_head)) { --it; continue; }
        const int nb = min(p.bps, nb_all - ((item % per_head) / p.n_qt) * p.bps);
        bar_wait(q_full, it & 1);
        qk(0);
        if (nb > 1) qk(1);
        if (nb > 2) qk(2);
        if (nb <= 3) commit(q_empty);
        bar_wait(o_free, it & 1);                              // previous item's O has been read out
        int buf = 0;
        for (int j = 0; j < nb; ++j) {
          pv(buf, j == 0);
          if (j + 3 < nb) {
            qk(buf);                                           // in order behind P.V(j), which read this buffer's P
            if (j + 3 == nb - 1) commit(q_empty);
          }
          if (++buf == 3) buf = 0;
        }
        commit(o_full);
      }
    }
  } else {
    // ===================== softmax warps: group g takes key blocks j = g mod 2; thread = (query row, half of the key block) =====
    const int g = warp >> 3, half = (warp >> 2) & 1, quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const uint32_t t_s0 = tmem_base + half * 64 + lane_off;                // my 64 S columns of buffer 0; P goes into the first 32
    const uint32_t t_o = tmem_base + COL_O + (g * 2 + half) * 32 + lane_off;   // my 32 O columns at read-out
    float* xch = reinterpret_cast<float*>(smem + Smem::xch);
    const uint32_t pf_addr0 = mapa_u32(&p_full[0], 0), of_addr = mapa_u32(o_free, 0);
    uint32_t itn = 0, xp = 0, bpar = 0;                                    // bpar bit b: parity of s_full[b] at the item's start
    const float c = p.scale_log2e;
    const uint64_t c2 = pack_f32x2(c, c);
    const int tail = p.Lk - (nb_all - 1) * BKEY;
    {                                                                      // O columns start out free
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(of_addr);
    }
    for (int item = item0; item < p.n_items; item += item_step, ++itn) {
      const int h = item / per_head, sp = (item % per_head) / p.n_qt;
      if (!head_is_bounded(p, h)) { --itn; continue; }
      const int kb0 = sp * p.bps, nb = min(p.bps, nb_all - kb0);
      const int row = (item % p.n_qt) * (2 * BQ) + rank * BQ + r;
      const bool row_ok = row < p.Lq;
      uint64_t sum2 = pack_f32x2(0.f, 0.f);
      int buf = g, use = 0;                                                // key block j lives in buffer j mod 3, its (j / 3)-th use
      for (int j = g; j < nb; j += 2) {
        bar_wait_crit(&s_full[buf], ((bpar >> buf) ^ (uint32_t)use) & 1u, p.mode);
        tcgen05_fence_after();
        const uint32_t t_s = t_s0 + buf * 128;
        const uint32_t pf_addr = pf_addr0 + buf * 8;
        buf += 2; if (buf >= 3) { buf -= 3; ++use; }                       // (j + 2) mod 3, (j + 2) / 3
        if (p.mode & 8) {                                                  // diagnostic: tensor / TMA pipeline alone
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(pf_addr);
          continue;
        }
        uint32_t sa[2][16], sb[2][16];
        tmem_ld_32x16(t_s, sa[0]);
        tmem_ld_32x16(t_s + 16, sa[1]);
        tmem_ld_wait();
        tmem_ld_32x16(t_s + 32, sb[0]);                                    // in flight during the first half's exponentials
        tmem_ld_32x16(t_s + 48, sb[1]);
        uint32_t pk[16];
        if (kb0 + j == nb_all - 1 && tail < BKEY) {                        // last, partial key block: -inf -> MUFU path gives P = 0
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            if (half * 64 + e >= tail) sa[0][e] = 0xff800000u;
            if (half * 64 + 16 + e >= tail) sa[1][e] = 0xff800000u;
            if (half * 64 + 32 + e >= tail) sb[0][e] = 0xff800000u;
            if (half * 64 + 48 + e >= tail) sb[1][e] = 0xff800000u;
          }
          exp_chunk_fast<0>(sa[0], pk, c2, sum2);
          exp_chunk_fast<0>(sa[1], pk + 8, c2, sum2);
          tmem_st_32x16(t_s, pk);
          exp_chunk_fast<0>(sb[0], pk, c2, sum2);
          exp_chunk_fast<0>(sb[1], pk + 8, c2, sum2);
          tmem_st_32x16(t_s + 16, pk);
        } else {
          exp_chunk_fast<POLY>(sa[0], pk, c2, sum2);
          exp_chunk_fast<POLY>(sa[1], pk + 8, c2, sum2);
          tmem_st_32x16(t_s, pk);
          tmem_ld_wait();
          exp_chunk_fast<POLY>(sb[0], pk, c2, sum2);
          exp_chunk_fast<POLY>(sb[1], pk + 8, c2, sum2);
          tmem_st_32x16(t_s + 16, pk);
        }
        tmem_st_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(pf_addr);
      }
      // buffer b was used (nb + 2 - b) / 3 times by this item
      bpar ^= (uint32_t)(((nb + 2) / 3) & 1) | (uint32_t)((((nb + 1) / 3) & 1) << 1) | (uint32_t)(((nb / 3) & 1) << 2);
      // ---- read-out: out = O / l, l = the four partial sums of the row (two groups x two halves) ----
      float s0, s1;
      unpack_f32x2(sum2, s0, s1);
      xch[xp * 512 + (g * 2 + half) * 128 + r] = s0 + s1;
      named_bar_sync(1, 512);
      const float* xr = xch + xp * 512 + r;
      const float l = (xr[0] + xr[128]) + (xr[256] + xr[384]);
      xp ^= 1;
      const float inv = 1.0f / l;
      bar_wait(o_full, itn & 1);
      tcgen05_fence_after();
      if (g == 0 && half == 0 && row_ok && p.lse_out != nullptr) p.lse_out[((long long)sp * p.H + h) * p.Lq + row] = log2f(l);
      uint32_t o[32];
      tmem_ld_32x16(t_o, o);
      tmem_ld_32x16(t_o + 16, o + 16);
      tmem_ld_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(of_addr);                         // O is in registers: hand the columns back
      if (row_ok) {
        __nv_bfloat16* dst = p.out + sp * p.split_stride + (long long)row * p.ldo + h * HD + (g * 2 + half) * 32;
#pragma unroll
        for (int e = 0; e < 32; e += 8) {
          uint4 w;
          w.x = pack_bf16x2(__uint_as_float(o[e]) * inv, __uint_as_float(o[e + 1]) * inv);
          w.y = pack_bf16x2(__uint_as_float(o[e + 2]) * inv, __uint_as_float(o[e + 3]) * inv);
          w.z = pack_bf16x2(__uint_as_float(o[e + 4]) * inv, __uint_as_float(o[e + 5]) * inv);
          w.w = pack_bf16x2(__uint_as_float(o[e + 6]) * inv, __uint_as_float(o[e + 7]) * inv);
          *reinterpret_cast<uint4*>(dst + e) = w;
        }
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  cluster_sync();
  tcgen05_fence_after();
  if (warp == W_TMA) tmem_dealloc_2sm<512>(tmem_base);
}

// Per-head maxima of the squared row norms: norms[h] = max_i |q_i,h|^2 (blockIdx.y = h), norms[H + h] = max_j |k_j,h|^2
// (blockIdx.y = H + h).  16 lanes read one 256-byte head slice of a row (16 B each); one atomicMax per CTA (the values
// are non-negative, so the integer order of the bit patterns is the float order).  norms must be zeroed before.
__global__ void __launch_bounds__(256) attn_qk_norm_kernel(const __nv_bfloat16* __restrict__ q, long long ldq, int Lq,
                                                           const __nv_bfloat16* __restrict__ k, long long ldk, int Lk, int H,
                                                           float* __restrict__ norms) {
  const bool is_k = (int)blockIdx.y >= H;
  const int h = is_k ? (int)blockIdx.y - H : (int)blockIdx.y;
  const __nv_bfloat16* base = (is_k ? k : q) + h * HD;
  const long long ld = is_k ? ldk : ldq;
  const int L = is_k ? Lk : Lq;
  const int grp = threadIdx.x >> 4, l16 = threadIdx.x & 15;
  const uint32_t hmask = 0xffffu << (threadIdx.x & 16);                   // the two row groups of a warp may leave the loop apart
  float best = 0.f;
  for (int row = blockIdx.x * 16 + grp; row < L; row += gridDim.x * 16) {
    const uint4 v = ldg_stream16(base + (long long)row * ld + l16 * 8);
    const uint32_t u[4] = {v.x, v.y, v.z, v.w};
    float ss = 0.f;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float a = __uint_as_float(u[e] << 16), b = __uint_as_float(u[e] & 0xffff0000u);
      ss = fmaf(a, a, fmaf(b, b, ss));
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) ss += __shfl_xor_sync(hmask, ss, o, 16);
    best = fmaxf(best, ss);                                               // NaN rows: see below
    if (ss != ss) best = INFINITY;                                        // NaN input -> unbounded -> online-softmax kernel
  }
  __shared__ float s_best[16];
  if (l16 == 0) s_best[grp] = best;
  __syncthreads();
  if (threadIdx.x == 0) {
    float m = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) m = fmaxf(m, s_best[i]);
    atomicMax(reinterpret_cast<int*>(norms + blockIdx.y), __float_as_int(m));
  }
}

// Key-split merge: out[row, c] = sum_s w_s * part[s][row, c] / sum_s w_s,  w_s = 2^(lse[s][h][row] - max_s lse) - the
// flash-attention split-K identity.  Thread = 8 columns of one row.
__global__ void __launch_bounds__(256) attn_merge_kernel(const __nv_bfloat16* __restrict__ part, long long split_stride, int ld_part,
                                                         const float* __restrict__ lse, int n_splits, int Lq, int H,
                                                         __nv_bfloat16* __restrict__ out, long long ldo) {
  const int vec_per_row = H * HD / 8;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)Lq * vec_per_row) return;
  const int row = (int)(i / vec_per_row), c = (int)(i % vec_per_row) * 8, h = c / HD;
  float w[8], m = -INFINITY, tot = 0.f;
  for (int s = 0; s < n_splits; ++s) { w[s] = lse[((long long)s * H + h) * Lq + row]; m = fmaxf(m, w[s]); }
  for (int s = 0; s < n_splits; ++s) { w[s] = ex2f(w[s] - m); tot += w[s]; }
  const float inv = 1.0f / tot;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int s = 0; s < n_splits; ++s) {
    const uint4 v = *reinterpret_cast<const uint4*>(part + s * split_stride + (long long)row * ld_part + c);
    const uint32_t u[4] = {v.x, v.y, v.z, v.w};
    const float ws = w[s] * inv;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      acc[2 * e] = fmaf(__uint_as_float(u[e] << 16), ws, acc[2 * e]);
      acc[2 * e + 1] = fmaf(__uint_as_float(u[e] & 0xffff0000u), ws, acc[2 * e + 1]);
    }
  }
  uint4 o;
  o.x = pack_bf16x2(acc[0], acc[1]); o.y = pack_bf16x2(acc[2], acc[3]); o.z = pack_bf16x2(acc[4], acc[5]); o.w = pack_bf16x2(acc[6], acc[7]);
  *reinterpret_cast<uint4*>(out + (long long)row * ldo + c) = o;
}

}  // namespace fa
}  // namespace b200q

using namespace b200q;

// scheduling knob: how many of every 8 element pairs take the polynomial exp2 instead of MUFU.EX2 (0..3; default 2 = 25 %).
// Diagnostic bits (results are then meaningless; timing probes only): +8 = softmax warps hand S straight back (the tensor /
// TMA / shared-memory pipeline alone), +16 = the MMA warp walks its schedule without issuing (the softmax warps alone),
// +32 = softmax warps stop after the TMEM load, row maximum and exchange.
static int g_fa_mode = 2;
// max-free kernels: polynomial pairs of every 8 (0..5); -1 = never use a max-free kernel; -2 = not set: the measured best of
// each kernel (two-tile: 3, key-pipelined: 1 - with three key blocks in flight the MUFU is no longer on a critical chain)
static int g_fa_fast_poly = -2;
static int g_fa_cl = 2;             // CTAs per cluster: 2 = CTA pairs with tcgen05.mma.cta_group::2 (default), 1 = single CTAs
extern "C" int b200q_attn_bf16_set_mode(int mode) {
  if (mode < 0 || mode > 255 || (mode & 4)) return B200Q_ERR_BAD_ARG;
  g_fa_mode = mode;
  return B200Q_OK;
}
extern "C" int b200q_attn_bf16_set_fast(int poly_pairs) {
  if (poly_pairs < -2 || poly_pairs > 5) return B200Q_ERR_BAD_ARG;
  g_fa_fast_poly = poly_pairs;
  return B200Q_OK;
}
static int g_fa_kp = 1;             // bounded heads: 1 = key-pipelined kernel (three key blocks in flight), 0 = two-tile kernel
extern "C" int b200q_attn_bf16_set_cluster(int ctas) {
  if (ctas != 1 && ctas != 2) return B200Q_ERR_BAD_ARG;
  g_fa_cl = ctas;
  return B200Q_OK;
}
extern "C" int b200q_attn_bf16_set_variant(int variant) {
  if (variant != 0 && variant != 1) return B200Q_ERR_BAD_ARG;
  g_fa_kp = variant;
  return B200Q_OK;
}

// How many key splits b200q_attn_bf16 should use for this shape (1 = none): (head, query-tile pair) items are walked by one
// persistent CTA per SM, so a grid that is not a multiple of the SM count leaves a partial last wave; splitting the keys
// multiplies the item count and shortens each item.  cost(S) = waves(S) * (key blocks per split + fixed per-item overhead).
extern "C" int b200q_attn_bf16_splits(int64_t Lq, int64_t Lk, int num_heads) {
  using namespace fa;
  if (Lq <= 0 || Lk <= 0 || num_heads <= 0) return 1;
  // items are walked by clusters of cl CTAs; the key-pipelined kernel (the one Wan's RMS-normed heads take) works on
  // 256-query items in CTA pairs, the two-tile kernel on 256 * cl
  const bool kp = g_fa_kp && g_fa_fast_poly != -1;
  const int cl = kp ? 2 : g_fa_cl;
  const int rows = kp ? 2 * BQ : 2 * BQ * cl;
  const long long items = ((Lq + rows - 1) / rows) * num_heads;
  const int nb = (int)((Lk + BKEY - 1) / BKEY), sms = sm_count() / cl;
  int best = 1;
  double best_cost = 0;
  for (int S = 1; S <= 4; ++S) {
    if (S > 1 && nb / S < 8) break;                                   // keep splits long enough to amortise prologue / read-out
    const int bps = (nb + S - 1) / S, s_eff = (nb + bps - 1) / bps;
    const double cost = (double)((items * s_eff + sms - 1) / sms) * (bps + 3.0) + (S > 1 ? 0.02 * (nb + 3.0) * ((items + sms - 1) / sms) : 0.0);
    if (S == 1 || cost < best_cost * 0.97) { if (S == 1 || cost < best_cost) { best = s_eff; best_cost = cost; } }
  }
  return best;
}

// One launch of the persistent kernel: one CTA per SM, whole clusters only.
template <int POLY, bool FAST, int CL>
static int launch_fa(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const fa::Params& p, cudaStream_t st) {
  using namespace fa;
  static bool configured = false;
  if (!configured) {
    B200Q_CUDA_OK(cudaFuncSetAttribute(attn_bf16_kernel<POLY, FAST, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, SmemT<CL>::total));
    configured = true;
  }
  const int cap = (sm_count() / CL) * CL;
  const int grid = p.n_items * CL < cap ? p.n_items * CL : cap;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = SmemT<CL>::total; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  B200Q_CUDA_OK(cudaLaunchKernelEx(&cfg, attn_bf16_kernel<POLY, FAST, CL>, tq, tk, tv, p));
  return B200Q_OK;
}
template <bool FAST, int CL>
static int launch_fa_poly(int poly, const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const fa::Params& p, cudaStream_t st) {
  switch (poly) {
    case 0: return launch_fa<0, FAST, CL>(tq, tk, tv, p, st);
    case 1: return launch_fa<1, FAST, CL>(tq, tk, tv, p, st);
    case 2: return launch_fa<2, FAST, CL>(tq, tk, tv, p, st);
    case 4: return launch_fa<FAST ? 4 : 3, FAST, CL>(tq, tk, tv, p, st);      // 4 and 5 exist for the max-free kernel only
    case 5: return launch_fa<FAST ? 5 : 3, FAST, CL>(tq, tk, tv, p, st);
    default: return launch_fa<3, FAST, CL>(tq, tk, tv, p, st);
  }
}
template <int POLY>
static int launch_fa_kp(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const fa::Params& p, cudaStream_t st) {
  using namespace fa;
  static bool configured = false;
  if (!configured) {
    B200Q_CUDA_OK(cudaFuncSetAttribute(attn_bf16_kp_kernel<POLY>, cudaFuncAttributeMaxDynamicSharedMemorySize, SmemKP::total));
    configured = true;
  }
  const int cap = (sm_count() / 2) * 2;
  const int grid = p.n_items * 2 < cap ? p.n_items * 2 : cap;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = SmemKP::total; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  B200Q_CUDA_OK(cudaLaunchKernelEx(&cfg, attn_bf16_kp_kernel<POLY>, tq, tk, tv, p));
  return B200Q_OK;
}
static int launch_fa_kp_poly(int poly, const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const fa::Params& p, cudaStream_t st) {
  switch (poly) {
    case 0: return launch_fa_kp<0>(tq, tk, tv, p, st);
    case 1: return launch_fa_kp<1>(tq, tk, tv, p, st);
    case 2: return launch_fa_kp<2>(tq, tk, tv, p, st);
    case 4: return launch_fa_kp<4>(tq, tk, tv, p, st);
    case 5: return launch_fa_kp<5>(tq, tk, tv, p, st);
    default: return launch_fa_kp<3>(tq, tk, tv, p, st);
  }
}
static int launch_fa_any(bool fast, int poly, int cl, const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv,
                         const fa::Params& p, cudaStream_t st) {
  if (cl == 2) return fast ? launch_fa_poly<true, 2>(poly, tq, tk, tv, p, st) : launch_fa_poly<false, 2>(poly, tq, tk, tv, p, st);
  return fast ? launch_fa_poly<true, 1>(poly, tq, tk, tv, p, st) : launch_fa_poly<false, 1>(poly, tq, tk, tv, p, st);
}

static int attn_bf16_impl(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, int64_t Lq, int64_t Lk,
                          int num_heads, int head_dim, float sm_scale, void* out, int64_t ldo, float* lse_out, int n_splits,
                          void* part_ws, float* lse_ws, float* qk_norm_ws, bool norms_given, b200q_stream_t stream);

extern "C" int b200q_attn_bf16(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                               int64_t Lq, int64_t Lk, int num_heads, int head_dim, float sm_scale, void* out, int64_t ldo,
                               float* lse_out, int n_splits, void* part_ws, float* lse_ws, float* qk_norm_ws,
                               b200q_stream_t stream) {
  return attn_bf16_impl(q, ldq, k, ldk, v, ldv, Lq, Lk, num_heads, head_dim, sm_scale, out, ldo, lse_out, n_splits, part_ws, lse_ws,
                        qk_norm_ws, false, stream);
}

extern "C" int b200q_attn_bf16_prenorm(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                                       int64_t Lq, int64_t Lk, int num_heads, int head_dim, float sm_scale, void* out, int64_t ldo,
                                       float* lse_out, int n_splits, void* part_ws, float* lse_ws, const float* qk_sq_max,
                                       b200q_stream_t stream) {
  clear_error();
  B200Q_REQUIRE(qk_sq_max != nullptr, B200Q_ERR_BAD_ARG, "attn_bf16_prenorm: qk_sq_max [2, heads] required");
  return attn_bf16_impl(q, ldq, k, ldk, v, ldv, Lq, Lk, num_heads, head_dim, sm_scale, out, ldo, lse_out, n_splits, part_ws, lse_ws,
                        const_cast<float*>(qk_sq_max), true, stream);
}

static int attn_bf16_impl(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, int64_t Lq, int64_t Lk,
                          int num_heads, int head_dim, float sm_scale, void* out, int64_t ldo, float* lse_out, int n_splits,
                          void* part_ws, float* lse_ws, float* qk_norm_ws, bool norms_given, b200q_stream_t stream) {
  clear_error();
  using namespace fa;
  B200Q_REQUIRE(q && k && v && out, B200Q_ERR_BAD_ARG, "attn_bf16: null pointer");
  B200Q_REQUIRE(Lq > 0 && Lk > 0 && num_heads > 0, B200Q_ERR_BAD_ARG, "attn_bf16: bad shape");
  B200Q_REQUIRE(head_dim == HD, B200Q_ERR_UNSUPPORTED, "attn_bf16: head_dim must be 128 (Wan2.1: 1536/12 = 5120/40 = 128)");
  B200Q_REQUIRE(Lq < (1ll << 31) - 512 && Lk < (1ll << 31) - 512, B200Q_ERR_UNSUPPORTED, "attn_bf16: sequence too long");
  const int64_t D = (int64_t)num_heads * HD;
  B200Q_REQUIRE(ldq >= D && ldk >= D && ldv >= D && ldo >= D, B200Q_ERR_BAD_ARG, "attn_bf16: leading dimension < heads*128");
  B200Q_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0 && aligned(q, 16) && aligned(k, 16) && aligned(v, 16) &&
                    aligned(out, 16),
                B200Q_ERR_BAD_ARG, "attn_bf16: q/k/v/out must be 16-byte aligned bf16 with row pitches that are multiples of 8");
  const int cl = g_fa_cl;
  CUtensorMap tq, tk, tv;
  int rc;
  if ((rc = make_tmap_2d(&tq, q, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, Lq, D, ldq, BQ, 64, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  if ((rc = make_tmap_2d(&tk, k, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, Lk, D, ldk, BKEY / cl, 64, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  if ((rc = make_tmap_2d(&tv, v, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, Lk, D, ldv, BKEY, 64, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  Params p{};
  p.Lq = (int)Lq; p.Lk = (int)Lk; p.H = num_heads;
  p.out = (__nv_bfloat16*)out; p.ldo = ldo;
  p.scale_log2e = sm_scale * 1.4426950408889634f;
  p.n_qt = (int)((Lq + 2 * BQ * cl - 1) / (2 * BQ * cl));
  const int nb_all = (int)((Lk + BKEY - 1) / BKEY);
  if (n_splits < 1) n_splits = 1;
  if (n_splits > 8) n_splits = 8;
  p.bps = (nb_all + n_splits - 1) / n_splits;
  p.n_splits = (nb_all + p.bps - 1) / p.bps;                          // every split non-empty
  if (p.n_splits > 1) {
    B200Q_REQUIRE(part_ws && lse_ws && aligned(part_ws, 16), B200Q_ERR_BAD_ARG,
                  "attn_bf16: n_splits > 1 needs part_ws (bf16 [n_splits, Lq, heads*128]) and lse_ws (fp32 [n_splits, heads, Lq])");
    B200Q_REQUIRE(lse_out == nullptr, B200Q_ERR_UNSUPPORTED, "attn_bf16: lse_out is only available without key splits");
    p.out = (__nv_bfloat16*)part_ws; p.ldo = D; p.split_stride = Lq * D; p.lse_out = lse_ws;
  } else {
    p.split_stride = 0; p.lse_out = lse_out;
  }
  p.n_items = p.n_qt * num_heads * p.n_splits;
  p.mode = g_fa_mode;
  p.qk_norm = (g_fa_fast_poly != -1) ? qk_norm_ws : nullptr;
  cudaStream_t st = (cudaStream_t)stream;
  if (p.qk_norm != nullptr) {
    // classify the heads (bounded scores -> max-free kernel), then run both kernels over the item list: each skips the
    // other's heads, so a launch whose class is empty costs a few microseconds
    if (!norms_given) {
      B200Q_CUDA_OK(cudaMemsetAsync(qk_norm_ws, 0, sizeof(float) * 2 * num_heads, st));
      attn_qk_norm_kernel<<<dim3(64, 2 * num_heads), 256, 0, st>>>((const __nv_bfloat16*)q, ldq, (int)Lq, (const __nv_bfloat16*)k, ldk,
                                                                   (int)Lk, num_heads, qk_norm_ws);
      B200Q_CHECK_LAUNCH();
    }
    if (g_fa_kp) {
      // bounded heads on the key-pipelined kernel: 256-query items on CTA pairs, K boxes of 64 keys
      Params pq = p;
      pq.n_qt = (int)((Lq + 2 * BQ - 1) / (2 * BQ));
      pq.n_items = pq.n_qt * num_heads * p.n_splits;
      CUtensorMap tk2 = tk;
      if (cl != 2 && (rc = make_tmap_2d(&tk2, k, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, Lk, D, ldk, BKEY / 2, 64, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
      if ((rc = launch_fa_kp_poly(g_fa_fast_poly == -2 ? 1 : g_fa_fast_poly, tq, tk2, tv, pq, st))) return rc;
    } else if ((rc = launch_fa_any(true, g_fa_fast_poly == -2 ? 3 : g_fa_fast_poly, cl, tq, tk, tv, p, st))) return rc;
  }
  if ((rc = launch_fa_any(false, g_fa_mode & 3, cl, tq, tk, tv, p, st))) return rc;
  if (p.n_splits > 1) {
    const long long vecs = Lq * (D / 8);
    attn_merge_kernel<<<(unsigned)((vecs + 255) / 256), 256, 0, st>>>((const __nv_bfloat16*)part_ws, p.split_stride, (int)D, lse_ws,
                                                                    p.n_splits, (int)Lq, num_heads, (__nv_bfloat16*)out, ldo);
    B200Q_CHECK_LAUNCH();
  }
  return B200Q_OK;
}
