/*
 * b200q.h — C ABI of libb200q.so: the B200 (sm_100a) native kernels for the quantized
 * hot path of the Wan2.1 DiT denoising loop.
 *
 * This is the drop-in boundary that replaces the reference's pybind11 torch-extension
 * modules `viditq_extension.fused` / `viditq_extension.qgemm`
 * (ViDiT-Q/kernels/csrc/fused/pybind.cpp:56-99, ViDiT-Q/kernels/csrc/qgemm/pybind.cpp:5-12,
 * declarations ViDiT-Q/kernels/csrc/qgemm/gemm_cuda.h:17-23) and, above them, the
 * elementwise torch arithmetic of the fake-quant path
 * (ViDiT-Q/quant_utils/qdiff/base/base_quantizer.py:58-162, quant_layer.py:57-74).
 *
 * Conventions
 *  - plain pointers and sizes only: every pointer is a DEVICE pointer unless stated,
 *    every tensor is row-major with an explicit leading dimension in ELEMENTS,
 *  - the caller allocates all outputs; the library never allocates or frees device memory
 *    and holds no reference to a caller buffer after the call returns (TMA descriptors are
 *    built per call on the host stack and passed by value to the kernel),
 *  - every call is asynchronous on `stream` (a cudaStream_t passed as void*), reentrant,
 *    and performs no host synchronisation,
 *  - return value: 0 = ok, negative = error (see enum); b200q_last_error() returns a
 *    thread-local human-readable message for the last failing call on this thread,
 *  - no shape-divisibility requirements: ragged M/N/K tails are predicated in-kernel
 *    (the reference asserts M%128, N%128, K%64: w8a8_gemm_cuda.cu:678-680),
 *  - numerics follow the reference fake-quant path, NOT the reference's fast-math kernels:
 *    IEEE fp32 division by delta and round-half-to-even (base_quantizer.py:119,155).
 */
#ifndef B200Q_H_
#define B200Q_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum b200q_status {
  B200Q_OK = 0,
  B200Q_ERR_BAD_ARG = -1,      /* null pointer, negative size, misaligned pointer, bad enum */
  B200Q_ERR_UNSUPPORTED = -2,  /* valid request this build cannot serve (e.g. row too long)  */
  B200Q_ERR_CUDA = -3          /* a CUDA runtime / driver call or the launch failed          */
};

enum b200q_dtype { B200Q_F32 = 0, B200Q_BF16 = 1, B200Q_F16 = 2, B200Q_I32 = 3 };

/* GEMM epilogues (b200q_gemm_w8a8 / b200q_gemm_w4a8) */
enum b200q_epilogue {
  B200Q_EPI_NONE = 0,          /* out = deq(acc) + bias                                              */
  B200Q_EPI_GELU_TANH = 1,     /* out = gelu_tanh(deq(acc) + bias)          (ffn.0 -> GELU, model.py:286-288) */
  B200Q_EPI_GATE_RESIDUAL = 2  /* out = residual + (deq(acc)+bias)*gate[n]  (x + y*e, model.py:337,362)        */
};

typedef void* b200q_stream_t;

#if defined(__GNUC__)
#define B200Q_API __attribute__((visibility("default")))
#else
#define B200Q_API
#endif

/* ---- library ---------------------------------------------------------------------- */
B200Q_API int b200q_version(void);                 /* (major<<16)|(minor<<8)|patch */
B200Q_API const char* b200q_last_error(void);      /* thread-local; "" if none    */
/* sm count / compute capability of the current device; errors if it is not sm_100. */
B200Q_API int b200q_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ---- (a) per-row quantizer ---------------------------------------------------------
 * One (delta, zero_point) per row; rows = tokens for activations, out-channels for weights.
 * Replaces DynamicQuantizer.quantize (base_quantizer.py:110-157), StaticQuantizer.
 * init_quant_params+quantize (base_quantizer.py:58-99) and the reference kernel
 * fused.quant_sum (kernels/csrc/fused/fused.cu:30-131, 524-645).
 *   sym : n_levels = 2^(b-1)-1 ; delta = amax/n_levels ; (dynamic: delta<1e-6 -> 1e-6) ; zp = 0
 *   asym: n_levels = 2^b ; delta = (max(rowmax,0)-min(rowmin,0))/(n_levels-1) ;
 *         zp = rne(xmin/delta) + n_levels/2
 *   q = rne(x/delta) - zp, stored as int8 (saturated to [-128,127]; the reference clamp
 *   [-n_levels-1, n_levels] never binds).  rowsum[r] = sum_c q[r,c] (int32) feeds the
 *   zero-point term of the GEMM epilogue.  delta / zero_point / rowsum: [rows]; rowsum may be NULL.
 *   x_dtype: B200Q_F32 | B200Q_BF16 | B200Q_F16 (up-cast to fp32, arithmetic in fp32).
 *   stat_max / stat_min (optional, [rows]): the row statistics the parameters were derived from
 *   (sym: |x| max in stat_max; asym: max(rowmax,0), min(rowmin,0)) = the quantizers' x_absmax /
 *   x_max / x_min attributes (base_quantizer.py:74-75,81-86,116-117,132-138).
 */
B200Q_API int b200q_quant_rows(const void* x, int x_dtype, int64_t rows, int64_t cols, int64_t ldx,
                     int n_bits, int sym, int dynamic,
                     int8_t* q, int64_t ldq, float* delta, float* zero_point, int32_t* rowsum,
                     float* stat_max, float* stat_min, b200q_stream_t stream);

/* Quantize with precomputed per-row parameters (StaticQuantizer.quantize after init_done,
 * base_quantizer.py:63-68; reference kernel fused.quant_sum_static). Codes clamp to
 * [max(-n_levels-1,-128), min(n_levels,127)]. */
B200Q_API int b200q_quant_rows_static(const void* x, int x_dtype, int64_t rows, int64_t cols, int64_t ldx,
                            int n_bits, int sym, const float* delta, const float* zero_point,
                            int8_t* q, int64_t ldq, int32_t* rowsum, b200q_stream_t stream);

/* out[r,c] = (q[r,c] + zero_point[r]) * delta[r]   (base_quantizer.py:159-162) */
B200Q_API int b200q_dequant_rows(const int8_t* q, int64_t ldq, int64_t rows, int64_t cols,
                       const float* delta, const float* zero_point,
                       void* out, int out_dtype, int64_t ldo, b200q_stream_t stream);

/* ---- (d) calibration reduction -------------------------------------------------------
 * Single read of x[rows, cols]; running per-channel statistics
 *   absmax_io[c] = max(absmax_io[c], max_r |x[r,c]|)   (get_calib_data_wanx.py:262-263;
 *   min_io[c]    = min(min_io[c],    min_r  x[r,c])     merge :443-468 + ptq_wanx.py:336
 *   max_io[c]    = max(max_io[c],    max_r  x[r,c])     == running elementwise max)
 * Any of the three may be NULL.  Caller initialises absmax_io to 0, min_io to +inf,
 * max_io to -inf once; the buffers then accumulate over calls (and over ranks with an
 * NCCL allreduce(MAX/MIN) on the same buffers).  NaNs are ignored.
 */
B200Q_API int b200q_calib_absmax_minmax(const void* x, int x_dtype, int64_t rows, int64_t cols, int64_t ldx,
                              float* absmax_io, float* min_io, float* max_io,
                              b200q_stream_t stream);

/* ---- (b) quantized linear ------------------------------------------------------------
 * out[m,n] = epi( delta_a[m]*delta_w[n] * ( sum_k qa[m,k]*qw[n,k] + zp_w[n]*rowsum_a[m] ) + bias[n] )
 * Replaces F.linear on the two dequantised operands (quant_layer.py:70) and the reference
 * kernels qgemm.w8a8_of16_bias_weight_asym / _sym / w8a8_o32
 * (kernels/csrc/qgemm/w8a8/w8a8_gemm_cuda.cu:14-838, epilogue :416-441).
 *   qa [M,K] int8 (lda), qw [N,K] int8 (ldw): both K-major; lda, ldw multiples of 16 and
 *   base pointers 16-byte aligned (TMA global-stride rule); M, N, K arbitrary (>0).
 *   delta_a [M], delta_w [N] fp32; zp_w [N] fp32 integers or NULL (symmetric weights);
 *   rowsum_a [M] int32, required iff zp_w != NULL; bias [N] of bias_dtype or NULL.
 *   out_dtype B200Q_BF16 | B200Q_F16 | B200Q_F32: dequantised output;
 *   out_dtype B200Q_I32: raw int32 accumulators (delta/zp/bias/epilogue ignored).
 *   epilogue B200Q_EPI_GATE_RESIDUAL: residual [M,N] fp32 (ldr), gate [N] fp32;
 *   `out` may alias `residual` when out_dtype == B200Q_F32 and ldo == ldr.
 */
B200Q_API int b200q_gemm_w8a8(const int8_t* qa, int64_t lda, const int8_t* qw, int64_t ldw,
                    const float* delta_a, const float* delta_w, const float* zp_w,
                    const int32_t* rowsum_a, const void* bias, int bias_dtype,
                    void* out, int out_dtype, int64_t ldo,
                    int64_t M, int64_t N, int64_t K,
                    int epilogue, const float* residual, int64_t ldr, const float* gate,
                    b200q_stream_t stream);

/* Tile-scheduling knob for both GEMMs (debug / benchmarking): 0 = automatic (default), 1 = single-CTA tiles only,
 * 2 = 2-CTA clusters with TMA-multicast B tiles whenever the problem has >= 2 row blocks, 3 = 2-CTA clusters issuing
 * tcgen05.mma.cta_group::2 (M = 256 across the pair).  Results are identical. */
B200Q_API int b200q_gemm_set_cluster(int mode);

/* codes int8 [N,K] in [-8,7] -> packed uint8 [N, ceil(K/8)*4].  Format (consumed by b200q_gemm_w4a8's in-smem
 * unpacker): K is split in groups of 8 codes; byte i (i=0..3) of the group's 32-bit word holds (code[i]+8) in bits 0-3
 * and (code[4+i]+8) in bits 4-7.  Nibbles are unsigned (the +8 bias is folded into the GEMM's zero-point term, the
 * algebra of the reference's QServe kernel, w4a8_per_channel_gemm_cuda_qserve.cu:290-297,585-586), so the unpack
 * is one AND and one SHIFT+AND per four codes.  ldp (bytes) must be a multiple of 4; use a multiple of 16 for the GEMM. */
B200Q_API int b200q_pack_w4(const int8_t* codes, int64_t ld, int64_t N, int64_t K,
                  uint8_t* packed, int64_t ldp, b200q_stream_t stream);

/* Same contract as b200q_gemm_w8a8 with 4-bit weights packed by b200q_pack_w4
 * (replaces qgemm.w4a8_of16_nobias_weight_asym_qserve,
 *  kernels/csrc/qgemm/w4a8/w4a8_per_channel_gemm_cuda_qserve.cu:304-656).
 * qw4 [N, ceil(K/8)*4] uint8 (ldw4 bytes, multiple of 16).  The packed tile travels by TMA; four converter warps expand
 * it to an int8 SWIZZLE_128B tile in shared memory before the MMA.  rowsum_a is REQUIRED (nibble bias); with
 * out_dtype B200Q_I32 the raw accumulators are sum_k qa*(code+8).
 * expand_ws (optional caller scratch, N * ceil(K/32)*32 bytes, 16-byte aligned): when given and M >= 1024 the packed
 * matrix is expanded once per call to one byte per code (microseconds: the weights are N*K/2 bytes) and the product runs
 * on the W8A8 kernel - in the DiT every one of the M/128 row tiles would otherwise repeat the same expansion in shared
 * memory.  Same accumulators, same epilogue algebra (zp_eff = zp_w - 8); the weights stay 4-bit in HBM and in checkpoints. */
B200Q_API int b200q_gemm_w4a8(const int8_t* qa, int64_t lda, const uint8_t* qw4, int64_t ldw4,
                    const float* delta_a, const float* delta_w, const float* zp_w,
                    const int32_t* rowsum_a, const void* bias, int bias_dtype,
                    void* out, int out_dtype, int64_t ldo,
                    int64_t M, int64_t N, int64_t K,
                    int epilogue, const float* residual, int64_t ldr, const float* gate, void* expand_ws,
                    b200q_stream_t stream);

/* ---- fused token-local operators (next-row f-1) ----------------------------------------
 * LayerNorm (fp32 two-pass statistics, eps) -> optional affine (ln_w, ln_b: [cols] or NULL)
 * -> optional adaLN modulate  y = ln*(1+scale[c]) + shift[c]  (scale/shift [cols] fp32 or NULL)
 * -> per-token symmetric n_bits quantisation (same contract as b200q_quant_rows sym dynamic).
 * Replaces WanLayerNorm + modulation + a_quantizer (wan/modules/model.py:92-102,327,359;
 * reference kernel fused.layernorm_nobias_t2i_quant_sum_fuse, fused.cu:234-380, 708-915,
 * which cannot launch for hidden > 4096).  y_out (optional, dtype y_dtype) receives the
 * un-quantised normalised activations (FP layers / debugging).
 */
B200Q_API int b200q_ln_mod_quant(const void* x, int x_dtype, int64_t rows, int64_t cols, int64_t ldx,
                       const float* ln_w, const float* ln_b, float eps,
                       const float* shift, const float* scale, int n_bits,
                       int8_t* q, int64_t ldq, float* delta, int32_t* rowsum,
                       void* y_out, int y_dtype, int64_t ldy, b200q_stream_t stream);

/* out = residual + y*gate[c]  (fp32 residual stream; y of y_dtype).  Replaces
 * `x = x + y * e[2]` (model.py:337,362) / reference fused.gate_residual_fuse (fused.cu:382-483).
 * gate may be NULL (plain residual add, model.py:352). out may alias residual. */
B200Q_API int b200q_gate_residual(const void* y, int y_dtype, int64_t ldy, const float* gate,
                        const float* residual, int64_t ldr, float* out, int64_t ldo,
                        int64_t rows, int64_t cols, b200q_stream_t stream);

/* y = RMSNorm_D(x) (fp32 statistics, cast back to x's dtype, * weight[cols]) then, if cos/sin are given, the 3-axis RoPE
 * rotation of adjacent channel pairs of every head: (a,b) -> (a*cos - b*sin, a*sin + b*cos) with cos/sin
 * [rows, head_dim/2] fp32 (host-built in float64 from the token's (f,h,w) position, rank-offset under sequence
 * parallelism).  x bf16|fp16 [rows, cols] -> out bf16 [rows, cols].  Replaces WanRMSNorm + rope_apply
 * (wan/modules/model.py:43-89; xdit_context_parallel.py:25-63). */
B200Q_API int b200q_rmsnorm_rope(const void* x, int x_dtype, int64_t rows, int64_t cols, int64_t ldx,
                       const float* weight, float eps, const float* cos_t, const float* sin_t, int head_dim,
                       void* out, int64_t ldo, b200q_stream_t stream);

/* Same, with the attention Q/K quantizer fused in: the rotated fp32 values are quantized per (token, head) — symmetric,
 * delta = amax/n_levels with the 1e-6 floor, codes rne(y/delta) — exactly DynamicQuantizer on the [tokens*heads, head_dim]
 * view (quant_opensora.py:430-435; base_quantizer.py:110-129,151-157).  q_out int8 [rows, cols] (ldq), dq_out fp32
 * [rows, cols/head_dim].  head_dim must be 128 (pass head_dim = 128 even without RoPE tables, e.g. cross-attention q/k).
 * `out` (bf16) may be NULL when only the codes are wanted. */
/* b200q_rmsnorm_rope that also leaves, per 128-column head, the maximum over the rows of the squared norm of its bf16
 * output in head_sq_max fp32 [cols / 128] (merged with atomicMax: the caller zeroes it; several calls may accumulate into
 * it).  Input of b200q_attn_bf16_prenorm: the attention's bounded-head classification without another pass over q and k. */
B200Q_API int b200q_rmsnorm_rope_stats(const void* x, int x_dtype, int64_t rows, int64_t cols, int64_t ldx,
                             const float* weight, float eps, const float* cos_t, const float* sin_t, int head_dim,
                             void* out, int64_t ldo, float* head_sq_max, b200q_stream_t stream);
B200Q_API int b200q_rmsnorm_rope_quant(const void* x, int x_dtype, int64_t rows, int64_t cols, int64_t ldx,
                             const float* weight, float eps, const float* cos_t, const float* sin_t, int head_dim,
                             void* out, int64_t ldo, int8_t* q_out, int64_t ldq, float* dq_out, int n_bits,
                             b200q_stream_t stream);

/* ---- (f-2) smooth scale + randomized Hadamard rotation fused into the per-token quantizer ----
 * y = (x * colscale) . (H_K (x) H_{2^log2_width});  q = rne(y/delta), delta = max|y| / n_levels (floor 1e-6): the
 * activation side of ViDiTQuantizedLinear / QuarotQuantizedLinear / SQQuantizedLinear.forward
 * (quant_utils/qdiff/viditq/viditq_quant_layer.py:58-66, quarot/quarot_quant_layer.py:55-62,
 * smooth_quant/sq_quant_layer.py:55-58), where the reference multiplies by the channel mask and then by a dense fp64
 * [n, n] rotation matrix.  A rotation R = diag(s) . H_n / sqrt(n) (random_hadamard_matrix, quarot/quarot_utils.py:186-192)
 * is passed in factored form: colscale[c] = channel_mask[c] * s[c] / sqrt(n) (fp32 [cols], NULL = ones), hadK = the
 * order-K base block (+-1 entries, fp32 [K*K] row-major, out[i] = sum_j hadK[i][j] * segment_j; NULL iff K == 1),
 * cols == K << log2_width with log2_width in [5, 8]; log2_width == 0 and K == 1 applies colscale only (SmoothQuant).
 * Same operator as matmul_hadU (quarot_utils.py:158-179) evaluated in fp32.  cols % 128 == 0, K <= 32.
 * rowsum (int32 [rows]) and y_out (fp32 [rows, cols], the rotated activations) are optional. */
B200Q_API int b200q_had_quant_rows(const void* x, int x_dtype, int64_t rows, int64_t cols, int64_t ldx,
                         const float* colscale, const float* hadK, int K, int log2_width, int n_bits,
                         int8_t* q, int64_t ldq, float* delta, int32_t* rowsum, float* y_out, int64_t ldy,
                         b200q_stream_t stream);
/* Debug / benchmarking knob: 1 (default) = rows of n = K*128 (K = 8, 12) and n = 20*256 channels take the register-resident
 * kernels (one / two warps per row), 0 = always the shared-memory tile kernel, 2 = like 1 with the three-CTAs-per-SM build
 * of the two-warp kernel (slower: it spills).  Same function; the kernels differ in fp32 rounding order only. */
B200Q_API int b200q_had_set_mode(int warp_kernel);

/* ---- (c) quantized attention ------------------------------------------------------------
 * Replaces the reference's materialised fake-quant attention (examples/Wan2.1/models/quant_opensora.py:430-478:
 * q/k/v DynamicQuantizers, `q*scale @ k^T`, fp32 softmax, attention-map quantizer, `attn @ v`), which builds
 * S and P as [H, L, L] tensors and therefore cannot run at L = 32,760 / 75,600.
 *
 * b200q_quant_vt: V quantizer.  v [Lk, C] (fp32|bf16|fp16, row pitch ldv, C = heads*head_dim) -> vt int8 [C, Lk]
 * (TRANSPOSED, row pitch ldvt: multiple of 16, >= Lk) and delta[C]: one symmetric scale per (head, channel) over all
 * tokens = DynamicQuantizer on `v.permute(0,1,3,2).reshape(-1, N_token)` rows (quant_opensora.py:440-442;
 * base_quantizer.py:110-129,151-157: delta = amax/n_levels, floor 1e-6, codes rne(x/delta) clamped).  absmax_ws [C] fp32
 * is caller-owned scratch (the per-channel |v| maximum, computed by the calibration reduction kernel).
 * The transposed layout makes a 128-key block of V a K-major B operand of the P.V product. */
B200Q_API int b200q_quant_vt(const void* v, int v_dtype, int64_t Lk, int64_t C, int64_t ldv, int n_bits,
                   float* absmax_ws, int8_t* vt, int64_t ldvt, float* delta, b200q_stream_t stream);

/* b200q_attn_bf16: bf16 flash attention, head_dim = 128 - the attention core of the W8A8 step.  Replaces the reference's
 * flash-attn call (examples/Wan2.1/wan/modules/attention.py:94-127: flash_attn_varlen_func(q, k, v, softmax_scale =
 * head_dim^-0.5), no mask, no dropout; SDPA fallback :171-178): out = softmax(q.k^T * sm_scale) . v per head.
 *   q bf16 [Lq, H*128] (ldq), k, v bf16 [Lk, H*128] (ldk, ldv): row pitches in elements, multiples of 8, 16-byte aligned
 *   bases - column slices of a fused q|k|v GEMM output are fine.  out bf16 [Lq, H*128] (ldo).
 *   lse_out (optional, fp32 [H, Lq]): log2(sum_j 2^(x_ij)), x = q.k^T * sm_scale * log2(e) (n_splits == 1 only).
 *   n_splits > 1: every (head, 512-query) work item is split along the keys into n_splits items (one persistent CTA per
 *   SM walks the items, so more and shorter items fill the last wave; b200q_attn_bf16_splits proposes the count); the
 *   partial outputs go to part_ws (bf16 [n_splits, Lq, H*128]) with their log-sum-exp in lse_ws (fp32 [n_splits, H, Lq]),
 *   both caller-owned scratch, and a second launch merges them with the weights 2^(lse_s - lse).
 *   qk_norm_ws (optional, fp32 [2, H] caller-owned scratch): when given, a first launch takes the per-head maxima of the
 *   row norms of q and k; heads whose scores are bounded by Cauchy-Schwarz, max|q_i| max|k_j| sm_scale log2(e) <= 80, run a
 *   max-free softmax (P = 2^x without a running maximum: bf16 / fp32 have the exponent range, and a float's relative
 *   precision does not depend on its magnitude - same function, fewer instructions, no rescaling); all other heads run the
 *   online-softmax kernel.  null = online softmax for every head.
 *   Q.K^T and P.V run as tcgen05.mma.kind::f16 with fp32 accumulators in TMEM, P is handed to the second product through
 *   tensor memory, V is consumed in its natural [keys, head_dim] layout; fp32 softmax statistics. */
B200Q_API int b200q_attn_bf16(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                    int64_t Lq, int64_t Lk, int num_heads, int head_dim, float sm_scale, void* out, int64_t ldo,
                    float* lse_out, int n_splits, void* part_ws, float* lse_ws, float* qk_norm_ws, b200q_stream_t stream);
/* b200q_attn_bf16 with the head classification input precomputed: qk_sq_max fp32 [2, H] = per head, the maximum over the
 * rows of |q_i|^2 (first H entries) and |k_j|^2 (next H) of the bf16 operands - what b200q_rmsnorm_rope_stats leaves behind -
 * so the launch that re-reads q and k is skipped.  Values larger than the true maxima are safe (a head is then merely
 * classified unbounded more often); inf / NaN classify the head as unbounded. */
B200Q_API int b200q_attn_bf16_prenorm(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                            int64_t Lq, int64_t Lk, int num_heads, int head_dim, float sm_scale, void* out, int64_t ldo,
                            float* lse_out, int n_splits, void* part_ws, float* lse_ws, const float* qk_sq_max,
                            b200q_stream_t stream);
/* Key-split count b200q_attn_bf16 should be called with for this shape on the current device (1 = no split). */
B200Q_API int b200q_attn_bf16_splits(int64_t Lq, int64_t Lk, int num_heads);

/* Scheduling knob of b200q_attn_bf16 (debug / benchmarking): how many of every 8 element pairs of the softmax take the
 * degree-4 polynomial exp2 on the FMA pipe instead of MUFU.EX2 (0..3, default 2 = 25 %; P within 7e-6 relative, far below
 * its bf16 rounding). */
B200Q_API int b200q_attn_bf16_set_mode(int mode);
/* Max-free kernels of b200q_attn_bf16 (bounded heads): polynomial pairs of every 8 (0..5); -2 (default) = the measured
 * best of the kernel in use (key-pipelined: 1, two-tile: 3); -1 disables the max-free kernels (every head takes the online
 * softmax even when qk_norm_ws is given). */
B200Q_API int b200q_attn_bf16_set_fast(int poly_pairs);
/* CTAs per work item of b200q_attn_bf16: 2 (default) = CTA pairs, 512 queries per item, tcgen05.mma.cta_group::2 with each
 * CTA staging half of every K and V tile; 1 = single CTAs, 256 queries per item.  Same function either way. */
B200Q_API int b200q_attn_bf16_set_cluster(int ctas);
/* Kernel the bounded (max-free) heads of b200q_attn_bf16 take: 1 (default) = key-pipelined kernel - one 128-row Q tile per
 * CTA, three S/P buffers in tensor memory, three key blocks in flight, CTA pairs; 0 = the two-tile kernel that also
 * serves the online-softmax heads.  Same function either way. */
B200Q_API int b200q_attn_bf16_set_variant(int variant);

/* b200q_scatter_rows: the data movement of the sequence-parallel attention exchange over NVLink peer memory.  Replaces the
 * Ulysses all-to-alls of examples/Wan2.1/wan/distributed/xdit_context_parallel.py:149-192 (xfuser SeqAllToAll4D: four
 * all_to_all_single calls per block with a permuting copy on either side).  n messages (<= 48) in ONE launch: message i
 * copies `rows` rows of `row_bytes` bytes from src[i] (row pitch src_pitch_bytes[i]) to dst[i] (row pitch dst_pitch_bytes);
 * dst[i] is typically a peer GPU's buffer (a CUDA peer mapping), src[i] a column slice of this rank's q|k|v GEMM output
 * or attention output.  src / dst / src_pitch_bytes are HOST arrays; everything 16-byte aligned, row_bytes and pitches
 * multiples of 16.  Ordering against the consumers on the other GPUs is the caller's (one barrier after the launch). */
B200Q_API int b200q_scatter_rows(const void* const* src, void* const* dst, int n, int64_t rows, int64_t row_bytes,
                       const int64_t* src_pitch_bytes, int64_t dst_pitch_bytes, b200q_stream_t stream);

/* b200q_attn_i8: fused int8 attention, head_dim = 128.
 *   qq int8 [Lq, H*128] (ldq), kq int8 [Lk, H*128] (ldk): per-(token, head) symmetric codes (b200q_quant_rows on the
 *   [L*H, 128] view); dq/dk: their fp32 scales, element (token, head) at dq[token*dq_tok_stride + head*dq_head_stride];
 *   vtq int8 [H*128, Lk] (ldvt), dv fp32 [H*128]: from b200q_quant_vt.
 *   S = qq.kq^T exact in int32 (tcgen05.mma.kind::i8); x = S*dq*dk*sm_scale; P~ = exp(x - rowmax x) quantized to unsigned
 *   8 bit with step 1/255 (the [0, 2^b-1] grid of forward_with_quant_params, base_quantizer.py:197-199, one step per
 *   query row); O = (P~q.vtq^T) exact in int32 * dv / (255 * sum_j P~).  out bf16 [Lq, H*128] (ldo).
 *   Deviation from the reference's attention-map grouping ('row': one scale per KEY column over all queries,
 *   quant_attn.py:168-174), which needs the whole [L, L] map first: see DESIGN.md; the parity path for that grouping is
 *   wan_b200/attention_q.py (small L).
 *   Optional debug/parity outputs (NULL to skip): m_out, l_out fp32 [H, Lq] (row maximum in log2 units, row sum of P~);
 *   p_out uint8 [H, Lq, ldp] (P~ codes; ldp multiple of 16 and >= ceil(Lk/128)*128); acc_out int32 [Lq, ldacc]
 *   (raw P.V accumulators).  Lk <= 66,000 (int32 accumulator bound). */
B200Q_API int b200q_attn_i8(const int8_t* qq, int64_t ldq, const float* dq, int64_t dq_tok_stride, int64_t dq_head_stride,
                  const int8_t* kq, int64_t ldk, const float* dk, int64_t dk_tok_stride, int64_t dk_head_stride,
                  const int8_t* vtq, int64_t ldvt, const float* dv,
                  int64_t Lq, int64_t Lk, int num_heads, int head_dim, float sm_scale,
                  void* out, int out_dtype, int64_t ldo,
                  float* m_out, float* l_out, uint8_t* p_out, int64_t ldp, int32_t* acc_out, int64_t ldacc,
                  b200q_stream_t stream);

/* Scheduling knob of b200q_attn_i8 (debug / benchmarking; results identical): bit 0 = S accumulators pre-initialised
 * with the int->fp32 conversion bias by tcgen05.st, bit 1 = pass 1 hands two key blocks per barrier round trip,
 * bit 2 = a quarter of the softmax exponentials evaluated by a degree-4 polynomial on the FMA/ALU pipes instead of the
 * MUFU (P~ within 7e-6 relative of the MUFU path), bit 3 = two softmax warpgroups per query tile (608 threads).
 * Bits 0, 1, 3 leave codes and accumulators bit-identical.  Default 2; measurements in DESIGN.md section 4c. */
B200Q_API int b200q_attn_set_mode(int mode);

#ifdef __cplusplus
}
#endif
#endif /* B200Q_H_ */
