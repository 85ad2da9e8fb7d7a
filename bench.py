#!/usr/bin/env python
"""bench.py — W8A8 Wan2.1 DiT-step benchmark (BASELINE.json metric / configs[1]).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU fake-quant path on the host cores

One "step" = one WanModel.forward (one CFG branch) of Wan2.1-T2V-1.3B, 30 blocks, all ten linears of every block
W8A8 (per-out-channel asymmetric weights, per-token symmetric activations), 832x480x81 synthetic latent =
32,760 tokens, random-init weights.  Prints ONE JSON line (see the contract in the task statement):
  value    : ms per step with inputs resident in HBM (CUDA events, max over ranks)
  e2e      : same through the public API with HOST (pinned) inputs, H2D + D2H inside the timed region
  roofline : dominant kernel = the tcgen05 int8 GEMM; achieved TOPS from CUDA events around every launch
  cpu_baseline : the oracle port of the reference fake-quant block timed on the host cores (bounded sample)
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "wan2.1-quantization_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

METRIC = "W8A8 DiT-step ms (Wan2.1-T2V-1.3B, 30 blocks, 32760 tokens)"
WORKLOAD = ("Wan2.1-T2V-1.3B full 30-block DiT W8A8 denoising step on B200, 832x480x81 synthetic latent "
            "(16x21x60x104 -> 32760 tokens), one CFG branch; BASELINE.json configs[1]")
LATENT_SHAPE = (16, 21, 60, 104)
TEXT_TOKENS, TEXT_DIM = 512, 4096


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--layers", type=int, default=None, help="debug only: fewer blocks (result is then marked invalid)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--model", default="1.3B", choices=["1.3B", "14B"],
                    help="1.3B @ 832x480x81 = BASELINE configs[1] (default, the headline); 14B @ 1280x720x81 = configs[3]")
    ap.add_argument("--attn", default="bf16", choices=["bf16", "int8"],
                    help="attention core: library bf16 flash attention (configs[1]) or the fused int8 Q.K^T/P.V kernel (configs[4])")
    ap.add_argument("--ffn-bits", type=int, default=8, choices=[4, 8], help="4 = W4A8 FFN weights (configs[4])")
    ap.add_argument("--no-variants", action="store_true", help="skip the extra int8-attention timing at N=1")
    ap.add_argument("--max-seconds", type=float, default=600.0, help="watchdog: hard-exit after this wall-clock time")
    ap.add_argument("--pipeline-chunks", type=int, default=0,
                    help="N>1: exchange/attend the heads of a rank's head group in this many chunks (exchange overlaps attention); "
                         "0 = auto: 3 at N=2 (measured 128.1 -> 120.7 ms on 2 B200), 1 elsewhere (not yet measured at N=4/8)")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from Python instead of replaying the CUDA graph")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------
# CPU reference arm / cpu_baseline: the oracle port of the reference fake-quant block on the host cores
# ------------------------------------------------------------------------------------------------------------
def cpu_reference_step_ms(L=32760, rows_sample=2048, q_sample=512, repeats=3, layers=30):
    """Bounded sample of configs[1] on the CPU (reference arithmetic: fp32 fake-quant, quant_layer.py:57-74 through the
    oracle port).  Token-local stages (LN/modulate, ten fake-quant linears, GELU, residuals, cross-attention) are
    timed on `rows_sample` of the L tokens and scaled by L/rows_sample; self-attention is timed for `q_sample`
    queries against all L keys and scaled by L/q_sample; one block x `layers`."""
    import torch.nn.functional as F
    from oracle import fakequant_oracle as O
    torch.set_num_threads(os.cpu_count())
    D, Fd, H = 1536, 8960, 12
    p = O.make_block_params(D, Fd, seed=0)
    blk = O.WanBlockOracle(p, D, Fd, H)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(rows_sample, D, generator=g)
    e = torch.randn(6, D, generator=g) * 0.1
    ctx = torch.randn(TEXT_TOKENS, D, generator=g)
    grid = (21, 30, 52)

    def token_local():
        # WanBlockOracle.forward with the self-attention core replaced by identity on v (timed separately below)
        blk.attention = lambda q, k, v: v.flatten(1)
        return blk.forward(x, e, grid, ctx)

    qh = torch.randn(1, H, q_sample, D // H, generator=g)
    kh = torch.randn(1, H, L, D // H, generator=g)
    vh = torch.randn(1, H, L, D // H, generator=g)

    def attn():
        return F.scaled_dot_product_attention(qh, kh, vh)

    def med(fn):
        fn()
        ts = []
        for _ in range(repeats):
            t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
        ts.sort()
        return ts[len(ts) // 2]

    with torch.no_grad():
        t_local, t_attn = med(token_local), med(attn)
    block_ms = 1e3 * (t_local * L / rows_sample + t_attn * L / q_sample)
    sample = (f"1 of {layers} blocks: token-local stages on {rows_sample}/{L} tokens ({t_local * 1e3:.0f} ms) + self-attention "
              f"on {q_sample}/{L} queries x {L} keys ({t_attn * 1e3:.0f} ms), each scaled linearly to L, x{layers} blocks; "
              f"fp32 torch CPU, median of {repeats}")
    return block_ms * layers, sample


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 3))
    ms, sample = cpu_reference_step_ms(repeats=max(1, steps))
    out = {
        "impl": "reference", "metric": METRIC, "value": ms, "unit": "ms", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32 (CPU fake-quant)", "data": "synthetic", "config": {"workload": WORKLOAD, "parallelism": "cpu"},
        "cpu_baseline": {"value": ms, "unit": "ms", "cores": os.cpu_count(), "kind": "port", "sample": sample},
        "e2e": {"value": ms, "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out))


# ------------------------------------------------------------------------------------------------------------
# HBM-bound kernels of the path (quantizer, calibration reduction): achieved GB/s against the measured copy bandwidth
# ------------------------------------------------------------------------------------------------------------
def hbm_kernel_rates(dev, peak_gbs):
    """Raw C-ABI launches back to back (no Python wrapper work in the timed region), inputs larger than the 126 MB L2,
    CUDA events around 20 launches.  Algorithmic bytes per SURVEY §8d: quantizer rows*cols*(sizeof(in)+1) + 12*rows;
    calibration rows*cols*sizeof(in) + 12*cols."""
    import ctypes
    import b200q
    lib = b200q.load()
    P = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    res = {"peak_gbs": peak_gbs, "peak_source": "MEASURED_PEAKS.json hbm_gbs (copy)" if peak_gbs else "unavailable"}

    def timed(fn, n=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(n):
            fn()
        e.record()
        torch.cuda.synchronize()
        return s.elapsed_time(e) / n

    for name, M, K, dt, code in (("quantizer_32760x8960_bf16", 32760, 8960, torch.bfloat16, 1),
                                 ("quantizer_32760x1536_f32", 32760, 1536, torch.float32, 0)):
        x = torch.randn(M, K, device=dev, dtype=dt)
        q = torch.empty(M, K, dtype=torch.int8, device=dev)
        d = torch.empty(M, device=dev); z = torch.empty(M, device=dev); rs = torch.empty(M, dtype=torch.int32, device=dev)
        ms = timed(lambda: lib.b200q_quant_rows(P(x), code, M, K, K, 8, 1, 1, P(q), K, P(d), P(z), P(rs), None, None, st))
        gbs = (M * K * (x.element_size() + 1) + 12 * M) / ms / 1e6
        res[name] = {"gbs": gbs, "frac": gbs / peak_gbs if peak_gbs else None, "us": ms * 1e3}
        if code == 0:
            stat = torch.zeros(K, device=dev)
            ms = timed(lambda: lib.b200q_calib_absmax_minmax(P(x), 0, M, K, K, P(stat), None, None, st))
            gbs = (M * K * 4 + 12 * K) / ms / 1e6
            res["calibration_32760x1536_f32"] = {"gbs": gbs, "frac": gbs / peak_gbs if peak_gbs else None, "us": ms * 1e3}
        del x, q
    return res


# ------------------------------------------------------------------------------------------------------------
# this repo's arm
# ------------------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch.distributed as dist
    import b200q
    from wan_b200 import model as M
    from wan_b200.parallel import SequenceParallel, exchange_bytes_per_rank

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    b200q.load()
    rc = b200q.load().b200q_device_info(None, None, None)
    if rc != 0:
        raise SystemExit("libb200q: " + b200q.load().b200q_last_error().decode())

    global LATENT_SHAPE
    cfg = M.WAN_1_3B if args.model == "1.3B" else M.WAN_14B
    if args.model == "14B":
        LATENT_SHAPE = (16, 21, 90, 160)                        # 1280x720x81 -> 75,600 tokens (BASELINE configs[3])
    default_cfg = args.model == "1.3B" and args.attn == "bf16" and args.ffn_bits == 8
    if args.pipeline_chunks == 0:
        args.pipeline_chunks = 3 if (world == 2 and (cfg.num_heads // 2) % 3 == 0) else 1
    sp = SequenceParallel(pipeline_chunks=args.pipeline_chunks) if world > 1 else None
    dit = M.WanDiTQ.random(cfg, seed=0, sp=sp, num_layers=args.layers, attn_quant=(args.attn == "int8"),
                           ffn_bits=args.ffn_bits)
    L = (LATENT_SHAPE[1] // 1) * (LATENT_SHAPE[2] // 2) * (LATENT_SHAPE[3] // 2)

    g = torch.Generator().manual_seed(0)
    lat_h = torch.randn(*LATENT_SHAPE, generator=g).pin_memory()
    ctx_h = torch.randn(TEXT_TOKENS, TEXT_DIM, generator=g).pin_memory()
    t_h = torch.tensor([500.0]).pin_memory()
    out_h = torch.empty(*LATENT_SHAPE).pin_memory()
    lat_d, ctx_d, t_d = lat_h.to(dev), ctx_h.to(dev), t_h.to(dev)

    # --- instrument the dominant kernel: CUDA events around every quantized-GEMM launch --------------------------
    gemm_events = []
    orig_qlinear = M.qlinear

    def timed_qlinear(qa, da, rowsum, w, *a, **kw):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        y = orig_qlinear(qa, da, rowsum, w, *a, **kw)
        e.record()
        gemm_events.append((s, e, 2.0 * qa.shape[0] * w.N * w.K, (qa.shape[0], w.N, w.K)))
        return y

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # the public step: CUDA-graph replay of WanDiTQ.forward (wan_b200.model.GraphedDiT); --no-graph launches eagerly
    use_graph = not args.no_graph
    runner = M.GraphedDiT(dit) if use_graph else dit.forward

    def step_resident():
        return runner(lat_d, t_d, ctx_d)

    def step_e2e():
        lat = lat_h.to(dev, non_blocking=True)
        ctx = ctx_h.to(dev, non_blocking=True)
        t = t_h.to(dev, non_blocking=True)
        y = runner(lat, t, ctx)
        out_h.copy_(y, non_blocking=True)
        return y

    for _ in range(max(args.warmup, 3)):
        step_resident()
    sync_all()

    # e2e first (it doubles as extra warm-up of allocator pools / NCCL channels for the device-resident timing below):
    # host buffers, H2D + D2H inside the timed region
    for _ in range(2):
        step_e2e()
    sync_all()
    ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev2.record()
    for _ in range(args.steps):
        step_e2e()
    ev3.record()
    sync_all()
    e2e_ms = ev2.elapsed_time(ev3) / args.steps

    # one sampler per job (rank 0's GPU): eight concurrent nvidia-smi pollers serialise on the driver and slow every rank
    clocks = ClockSampler(local) if rank == 0 else None
    if clocks is not None:
        clocks.start()
    sync_all()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step_resident()
    ev1.record()
    sync_all()
    ms = ev0.elapsed_time(ev1) / args.steps

    # per-kernel pass.  Preferred: the same K steps replayed from a second CUDA graph that carries an external timing
    # event (cudaEventRecordExternal) before and after every quantized-GEMM launch, so the per-launch durations are free
    # of host launch gaps.  Fallback (if that capture is refused): eager launches with ordinary events.
    per_kernel_mode = "cuda-graph replay with external event-record nodes around every GEMM launch"
    pool, cursor = [], [0]

    def timed_qlinear_ext(qa, da, rowsum, w, *a, **kw):
        i = cursor[0]
        cursor[0] += 1
        if i == len(pool):
            pool.append((torch.cuda.Event(enable_timing=True, external=True), torch.cuda.Event(enable_timing=True, external=True),
                         2.0 * qa.shape[0] * w.N * w.K, (qa.shape[0], w.N, w.K)))
        s, e, _, _ = pool[i]
        s.record()
        y = orig_qlinear(qa, da, rowsum, w, *a, **kw)
        e.record()
        return y

    class _Counted:
        def forward(self, *a):
            cursor[0] = 0
            return dit.forward(*a)

    launches0 = b200q.launch_count
    dit.forward(lat_d, t_d, ctx_d)
    launches = (b200q.launch_count - launches0) * args.steps       # this library's launches per step x K
    eager_ms = None
    try:
        if not use_graph:
            raise RuntimeError("--no-graph")
        if world > 1:
            # validated on one GPU only; a capture that failed on a single rank would leave the others inside NCCL
            raise RuntimeError("multi-rank run: per-kernel pass launched eagerly")
        M.qlinear = timed_qlinear_ext
        inst = M.GraphedDiT(_Counted())
        inst(lat_d, t_d, ctx_d)
        sync_all()
        if inst.failed is not None:
            raise RuntimeError(inst.failed)
        for _ in range(args.steps):
            inst(lat_d, t_d, ctx_d)
            torch.cuda.synchronize()
            for s_, e_, o_, shp_ in pool:
                gemm_events.append((s_.elapsed_time(e_), o_, shp_))
        M.qlinear = orig_qlinear
        del inst
    except Exception as ex:  # noqa: BLE001
        per_kernel_mode = "eager launches, CUDA events around every GEMM launch (graph instrumentation failed: %r)" % (ex,)
        gemm_events.clear()
        M.qlinear = timed_qlinear
        sync_all()
        ev4, ev5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev4.record()
        for _ in range(args.steps):
            dit.forward(lat_d, t_d, ctx_d)
        ev5.record()
        sync_all()
        M.qlinear = orig_qlinear
        eager_ms = ev4.elapsed_time(ev5) / args.steps
        gemm_events[:] = [(s_.elapsed_time(e_), o_, shp_) for s_, e_, o_, shp_ in gemm_events]

    # dominant-kernel accounting: all quantized-GEMM launches, and the single heaviest shape (roofline object)
    g_ms = sum(t_ for t_, _, _ in gemm_events)
    g_ops = sum(o for _, o, _ in gemm_events)
    by_shape = {}
    for t_, o, shp in gemm_events:
        a = by_shape.setdefault("x".join(map(str, shp)), [0.0, 0.0, 0])
        a[0] += t_; a[1] += o; a[2] += 1
    gemm_events.clear()

    clk = clocks.stop() if clocks is not None else None

    # extra timing at N=1 on the headline config: the same step with the fused int8 attention kernel (configs[4] semantics)
    variants = {}
    if world == 1 and default_cfg and not args.no_variants and args.layers is None:
        try:
            for blk in dit.blocks:
                blk.attn_quant = True
            vrun = M.GraphedDiT(dit) if use_graph else dit.forward
            for _ in range(2):
                vrun(lat_d, t_d, ctx_d)
            torch.cuda.synchronize()
            v0, v1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            v0.record()
            for _ in range(2):
                vrun(lat_d, t_d, ctx_d)
            v1.record()
            torch.cuda.synchronize()
            variants["w8a8_linears_int8_attention_ms"] = v0.elapsed_time(v1) / 2
        except Exception as ex:  # noqa: BLE001
            variants["w8a8_linears_int8_attention_ms"] = "failed: " + repr(ex)
        finally:
            for blk in dit.blocks:
                blk.attn_quant = False

    if world > 1:
        tmax = torch.tensor([ms, e2e_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms, e2e_ms = float(tmax[0]), float(tmax[1])

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:  # noqa: BLE001
            pass
        bf16_sus = peaks.get("bf16_tflops_sustained")
        if bf16_sus:
            peak, peak_src = 2.0 * bf16_sus, "2 x MEASURED_PEAKS.json bf16_tflops_sustained (kind::i8 issues at 2x the bf16 rate)"
        else:
            peak, peak_src = 2.0 * 1400.0, "fallback: 2 x 1.4 PFLOP/s sustained bf16 (B200_PROFILING.md)"
        all_tops = g_ops / (g_ms * 1e-3) / 1e12 if g_ms > 0 else 0.0
        # the dominant launch: the shape with the largest share of GEMM time
        dom = max(by_shape, key=lambda k: by_shape[k][0]) if by_shape else None
        dom_ms = by_shape[dom][0] / by_shape[dom][2] if dom else 0.0
        dom_ops = by_shape[dom][1] / by_shape[dom][2] if dom else 0.0
        achieved = dom_ops / (dom_ms * 1e-3) / 1e12 if dom_ms > 0 else 0.0
        traffic = None
        try:                        # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu capture
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("gemm", {}).get(dom)
        except Exception:  # noqa: BLE001
            pass
        metric = METRIC if args.model == "1.3B" else "W8A8 DiT-step ms (Wan2.1-T2V-14B, 40 blocks, 75600 tokens)"
        workload = WORKLOAD if args.model == "1.3B" else (
            "Wan2.1-T2V-14B 40-block DiT W8A8 denoising step, 1280x720x81 synthetic latent (16x21x90x160 -> 75600 tokens), "
            "one CFG branch; BASELINE.json configs[3]")
        if not default_cfg:
            workload += f" [attention={args.attn}, ffn weights {args.ffn_bits}-bit]"
        out = {
            "metric": metric, "value": ms, "unit": "ms", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms, "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
            "dtype": "int8 (s32 accumulate, %s attention core)" % ("bf16 library" if args.attn == "bf16" else "int8 tcgen05"),
            "data": "synthetic",
            "config": {"workload": workload, "parallelism": f"ulysses{world}" if world > 1 else "single-gpu",
                       "weights": "random-init, replicated", "l2": "per-step working set (>=200 MB per stage) >> 126 MB L2; no flush needed",
                       "cfg_branches_per_step": 1},
            "clocks": clk,
            "e2e": {"value": e2e_ms, "unit": "ms", "h2d_bytes_per_step": lat_h.numel() * 4 + ctx_h.numel() * 4 + 4,
                    "d2h_bytes_per_step": out_h.numel() * 4},
            "gpu_launches": launches,
            "launch_mode": "cuda-graph replay (wan_b200.model.GraphedDiT)" if use_graph else "eager",
            "eager_ms_per_step": eager_ms,
            "roofline": {"bound": "tensor", "kernel": f"gemm_i8_kernel M,N,K={dom} (tcgen05.mma.cta_group::2.kind::i8, TMA, TMEM)",
                         "achieved": achieved, "peak": peak, "unit": "TOP/s", "frac": achieved / peak if peak else None,
                         "traffic": traffic, "algorithmic_ops_per_launch": dom_ops, "avg_launch_ms": dom_ms,
                         "peak_source": peak_src, "frac_of_nominal_4500": achieved / 4500.0,
                         "all_gemms_tops": all_tops, "all_gemms_frac": all_tops / peak if peak else None,
                         "gemm_share_of_step": g_ms / ((eager_ms or ms) * args.steps),
                         "timing": per_kernel_mode,
                         "by_shape_tops": {k: v[1] / (v[0] * 1e-3) / 1e12 for k, v in by_shape.items() if v[0] > 0}},
        }
        if variants:
            out["variants"] = variants
        try:
            out["hbm_kernels"] = hbm_kernel_rates(dev, peaks.get("hbm_gbs"))
        except Exception as ex:  # noqa: BLE001
            out["hbm_kernels"] = {"error": repr(ex)}
        if world > 1:
            b, pu, pr = exchange_bytes_per_rank(L, cfg.dim, world, cfg.num_heads)
            out["config"]["exchange_bytes_per_rank_per_block"] = b
            out["config"]["head_plan"] = f"Pu={pu} x Pr={pr}"
            out["config"]["pipeline_chunks"] = args.pipeline_chunks
        if args.layers is not None and args.layers != cfg.num_layers:
            out["invalid"] = f"debug run with {args.layers} of {cfg.num_layers} blocks"
        if world == 1 and not args.no_cpu_baseline:
            try:
                cms, sample = cpu_reference_step_ms()
                out["cpu_baseline"] = {"value": cms, "unit": "ms", "cores": os.cpu_count(), "kind": "port", "sample": sample}
            except Exception as ex:  # noqa: BLE001
                out["cpu_baseline"] = {"value": None, "unit": "ms", "cores": os.cpu_count(), "kind": "port",
                                       "sample": "failed: " + repr(ex)}
        print(json.dumps(out), flush=True)
    # No collective and no process-group teardown after the numbers are out: destroying an NCCL communicator whose
    # send/recvs live inside captured CUDA graphs was observed to hang on B200 (torch 2.11 / NCCL 2.28); every rank has
    # finished its collectives at this point, so it leaves immediately.
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0)


def _watchdog(seconds):
    """A benchmark must never hold a GPU box hostage: hard-exit if the run exceeds its wall-clock budget."""
    def fire():
        sys.stderr.write(f"bench.py: watchdog fired after {seconds} s - aborting\n")
        sys.stderr.flush()
        os._exit(3)
    t = threading.Timer(seconds, fire)
    t.daemon = True
    t.start()


if __name__ == "__main__":
    a = parse()
    _watchdog(a.max_seconds)
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
