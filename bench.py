#!/usr/bin/env python
"""bench.py — W8A8 Wan2.1 DiT-step benchmark (BASELINE.json metric / configs[1]).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU fake-quant path on the host cores

One "step" = one WanModel.forward (one CFG branch) of Wan2.1-T2V-1.3B, 30 blocks, all ten linears of every block
W8A8 (per-out-channel asymmetric weights, per-token symmetric activations), 832x480x81 synthetic latent =
32,760 tokens, random-init weights.  Prints ONE JSON line (see the contract in the task statement):
  value     : ms per step with inputs resident in HBM (CUDA events, max over ranks)
  e2e       : same through the public API with HOST (pinned) inputs, H2D + D2H inside the timed region
  breakdown : attention / quantized GEMMs / everything else per step, from external timing events inside the CUDA graph
  roofline  : the kernel with the largest share of the step (the attention core when it is this repo's kernel, else the
              heaviest quantized GEMM), achieved rate against the MEASURED peak (int8: torch._int_mm probed in this run)
  verify    : N > 1: the sharded step's output against the unsharded step run on rank 0 (cosine, max |diff|)
  cpu_baseline : one full fake-quant block of the same shape timed on the host cores (oracle port, or the imported
              reference qdiff layers where /root/reference exists), x layers
"""
import argparse
import atexit
import gc
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
    ap.add_argument("--attn", default="auto", choices=["auto", "b200q", "library", "int8"],
                    help="attention core: b200q = this repo's bf16 tcgen05 flash attention, library = torch SDPA (cuDNN), "
                         "int8 = the fused int8 Q.K^T/P.V kernel (configs[4]); auto = b200q")
    ap.add_argument("--ffn-bits", type=int, default=8, choices=[4, 8], help="4 = W4A8 FFN weights (configs[4])")
    ap.add_argument("--cfg-batch", type=int, default=1, choices=[1, 2],
                    help="2 = cond + uncond branches batched into one step (SURVEY 8 f-4); value is then ms per 2 branches")
    ap.add_argument("--no-variants", action="store_true", help="skip the extra attention-core timings at N=1")
    ap.add_argument("--no-verify", action="store_true", help="N>1: skip the comparison against the unsharded step")
    ap.add_argument("--max-seconds", type=float, default=900.0, help="watchdog: hard-exit after this wall-clock time")
    ap.add_argument("--exchange", default="auto", choices=["auto", "peer", "nccl"],
                    help="N>1: attention exchange over NVLink peer memory (b200q_scatter_rows + barrier) or NCCL send/recv")
    ap.add_argument("--pipeline-chunks", type=int, default=0,
                    help="N>1: exchange/attend the heads of a rank's head group in this many chunks (exchange overlaps "
                         "attention); 0 = auto")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from Python instead of replaying the CUDA graph")
    ap.add_argument("--ref-tokens", type=int, default=32760,
                    help="reference arm / cpu_baseline: tokens of the timed block (default: the full 32,760; smaller = CI only, marked)")
    ap.add_argument("--ref-budget-s", type=float, default=240.0, help="reference arm: stop timing further steps past this many seconds")
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
        sm, mx, pw, reasons = [], [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
            except (ValueError, IndexError):
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------
# CPU reference arm / cpu_baseline
# ------------------------------------------------------------------------------------------------------------
def cpu_reference_block(L=32760, steps=1, warmup=0, budget_s=240.0):
    """ONE complete fake-quant Wan-1.3B block (configs[0]: D=1536, F=8960, H=12, W8 per-channel asym / A8 per-token sym,
    quant_layer.py:57-74) at L tokens on the host cores, fp32, all threads: LN/modulate, the ten fake-quant linears, both
    attentions (F.scaled_dot_product_attention, wan/modules/attention.py:171-178), GELU, residuals - nothing sampled or
    scaled inside the block.  Where /root/reference exists (the build container) the linears are the IMPORTED reference
    `QuantizedLinear` modules (kind "reference-import"); on the GPU box that tree does not exist and the oracle port of
    the same arithmetic runs (kind "port").  -> (per-step block times ms, kind, description)."""
    from oracle import fakequant_oracle as O
    torch.set_num_threads(os.cpu_count())
    D, Fd, H = 1536, 8960, 12
    grid = (21, 30, 52)
    if L != 32760:
        assert L % 64 == 0 and L // 64 <= 1024, "--ref-tokens: a multiple of 64 (CI only)"
        grid = (L // 64, 8, 8)
    p = O.make_block_params(D, Fd, seed=0)
    blk = O.WanBlockOracle(p, D, Fd, H)
    kind = "port"
    try:
        from oracle.ref_import import import_reference_qdiff, reference_available
        if reference_available() and "qdiff" not in sys.modules:
            # the reference's `qdiff` is a namespace package (no __init__.py); this repo's mirror of the same name would
            # shadow it from any sys.path position, so it is taken off the path for this (mirror-free) process
            sys.path[:] = [q for q in sys.path if os.path.abspath(q) != PKG]
            ref = import_reference_qdiff()
            from omegaconf import OmegaConf
            cfg = OmegaConf.create({"weight": {"n_bits": 8, "sym": False}, "act": {"n_bits": 8, "sym": True}})
            layers = {}
            for name in ("self_attn.q", "self_attn.k", "self_attn.v", "self_attn.o", "cross_attn.q", "cross_attn.k",
                         "cross_attn.v", "cross_attn.o", "ffn.0", "ffn.2"):
                w, b = p[name + ".weight"], p[name + ".bias"]
                fp = torch.nn.Linear(w.shape[1], w.shape[0])
                with torch.no_grad():
                    fp.weight.copy_(w); fp.bias.copy_(b)
                layers[name] = ref["quant_layer"].QuantizedLinear(w.shape[1], w.shape[0], True, None, cfg, fp)
                layers[name].a_quantizer.module_name = name
            blk.lin = lambda name, x: layers[name](x.unsqueeze(0))[0]
            kind = "reference-import"
    except Exception:  # noqa: BLE001
        kind = "port"
    g = torch.Generator().manual_seed(1)
    x = torch.randn(L, D, generator=g)
    e = torch.randn(6, D, generator=g) * 0.1
    ctx = torch.randn(TEXT_TOKENS, D, generator=g)
    times = []
    t_start = time.perf_counter()
    with torch.no_grad():
        for _ in range(warmup):
            blk.forward(x, e, grid, ctx)
            if time.perf_counter() - t_start > budget_s / 2:
                break
        for _ in range(max(1, steps)):
            t0 = time.perf_counter()
            blk.forward(x, e, grid, ctx)
            times.append((time.perf_counter() - t0) * 1e3)
            if time.perf_counter() - t_start > budget_s:
                break
    desc = (f"one complete fake-quant block at L={L} tokens (D=1536, F=8960, H=12; ten W8A8 fake-quant linears, self- and "
            f"cross-attention, LN/GELU/residuals; no sampling inside the block), fp32 torch CPU, {os.cpu_count()} threads, "
            f"{len(times)} timed forwards (median {sorted(times)[len(times) // 2]:.0f} ms); the 30-block step = 30 x the block")
    return times, kind, desc


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    times, kind, desc = cpu_reference_block(L=args.ref_tokens, steps=args.steps, warmup=min(args.warmup, 1), budget_s=args.ref_budget_s)
    layers = 30
    block_ms = sorted(times)[len(times) // 2]
    ms = block_ms * layers
    out = {
        "impl": "reference", "metric": METRIC, "value": ms, "unit": "ms", "n_gpus": args.gpus, "steps": len(times),
        "steps_requested": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": ms, "higher_is_better": False,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32 (CPU fake-quant)", "data": "synthetic",
        "config": {"workload": WORKLOAD, "parallelism": "cpu", "timed_sample": "1 block of 30 per step (measured whole), x30"},
        "block_ms": block_ms, "block_ms_all": times, "layers_multiplier": layers,
        "cpu_baseline": {"value": ms, "unit": "ms", "cores": os.cpu_count(), "kind": kind, "sample": desc},
        "e2e": {"value": ms, "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if args.ref_tokens != 32760:
        out["invalid"] = f"reduced-token CI run ({args.ref_tokens} of 32760 tokens)"
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------------------------------------------------
# measured peaks of this run (the roofline denominators)
# ------------------------------------------------------------------------------------------------------------
def measure_int8_peak(dev, seconds=2.0):
    """cuBLASLt int8 (torch._int_mm) 8192^3, the way MEASURED_PEAKS.json measures bf16: best of 10 (burst) and back to
    back for `seconds` (sustained, under the power cap)."""
    n = 8192
    a = torch.randint(-127, 127, (n, n), dtype=torch.int8, device=dev)
    b = torch.randint(-127, 127, (n, n), dtype=torch.int8, device=dev).t()
    ops = 2.0 * n ** 3
    for _ in range(3):
        torch._int_mm(a, b)
    torch.cuda.synchronize()
    best = 0.0
    for _ in range(10):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); torch._int_mm(a, b); e.record()
        torch.cuda.synchronize()
        best = max(best, ops / (s.elapsed_time(e) * 1e-3) / 1e12)
    reps = max(10, int(seconds / (ops / (best * 1e12))))
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        torch._int_mm(a, b)
    e.record()
    torch.cuda.synchronize()
    sustained = ops * reps / (s.elapsed_time(e) * 1e-3) / 1e12
    return {"int8_tops_burst": best, "int8_tops_sustained": sustained, "how": f"torch._int_mm {n}^3: best of 10, then {reps} back to back"}


# ------------------------------------------------------------------------------------------------------------
# HBM-bound kernels of the path: achieved GB/s against the measured copy bandwidth
# ------------------------------------------------------------------------------------------------------------
def hbm_kernel_rates(dev, peak_gbs):
    """Raw C-ABI launches back to back (no Python wrapper work in the timed region), inputs larger than the 126 MB L2,
    CUDA events around 20 launches.  Algorithmic bytes per SURVEY §8d."""
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

    def put(name, nbytes, ms):
        gbs = nbytes / ms / 1e6
        res[name] = {"gbs": gbs, "frac": gbs / peak_gbs if peak_gbs else None, "us": ms * 1e3, "algorithmic_mb": nbytes / 1e6}

    M = 32760
    for name, K, dt, code in (("quantizer_32760x8960_bf16", 8960, torch.bfloat16, 1), ("quantizer_32760x1536_f32", 1536, torch.float32, 0)):
        x = torch.randn(M, K, device=dev, dtype=dt)
        q = torch.empty(M, K, dtype=torch.int8, device=dev)
        d = torch.empty(M, device=dev); z = torch.empty(M, device=dev); rs = torch.empty(M, dtype=torch.int32, device=dev)
        ms = timed(lambda: lib.b200q_quant_rows(P(x), code, M, K, K, 8, 1, 1, P(q), K, P(d), P(z), P(rs), None, None, st))
        put(name, M * K * (x.element_size() + 1) + 12 * M, ms)
        if code == 0:
            stat = torch.zeros(K, device=dev)
            ms = timed(lambda: lib.b200q_calib_absmax_minmax(P(x), 0, M, K, K, P(stat), None, None, st))
            put("calibration_32760x1536_f32", M * K * 4 + 12 * K, ms)
            # fused LN + modulate + quant (fp32 residual stream in, int8 out)
            sh = torch.randn(K, device=dev) * 0.1; sc = torch.randn(K, device=dev) * 0.1
            ms = timed(lambda: lib.b200q_ln_mod_quant(P(x), 0, M, K, K, None, None, ctypes.c_float(1e-6), P(sh), P(sc), 8, P(q), K,
                                                      P(d), P(rs), None, 0, 0, st))
            put("ln_mod_quant_32760x1536_f32", M * K * 5 + 8 * M, ms)
            # ViDiT-Q: smooth scale + Hadamard rotation (12 x 2^7) + quant
            try:
                from qdiff.base.quant_layer import ActPlan
                plan = ActPlan.rotation(K, torch.ones(K), torch.rand(K) + 0.5, dev)
                ms = timed(lambda: lib.b200q_had_quant_rows(P(x), 0, M, K, K, P(plan.colscale), P(plan.hadK), plan.K, plan.log2w, 8,
                                                            P(q), K, P(d), P(rs), None, 0, st))
                put("had_quant_rows_32760x1536_f32", M * K * 5 + 8 * M, ms)
            except Exception as ex:  # noqa: BLE001
                res["had_quant_rows_32760x1536_f32"] = {"error": repr(ex)}
        del x, q
    try:
        xb = torch.randn(M, 3 * 1536, device=dev, dtype=torch.bfloat16)
        w = torch.ones(1536, device=dev)
        cos = torch.rand(M, 64, device=dev); sin = torch.rand(M, 64, device=dev)
        o = torch.empty(M, 1536, device=dev, dtype=torch.bfloat16)
        ms = timed(lambda: lib.b200q_rmsnorm_rope(P(xb), 1, M, 1536, 3 * 1536, P(w), ctypes.c_float(1e-6), P(cos), P(sin), 128, P(o), 1536, st))
        put("rmsnorm_rope_32760x1536_bf16", M * 1536 * 4 + M * 64 * 8, ms)
    except Exception as ex:  # noqa: BLE001
        res["rmsnorm_rope_32760x1536_bf16"] = {"error": repr(ex)}
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
    # HBM-bound kernels alone, BEFORE the long power-capped step: their denominator (MEASURED_PEAKS.json hbm_gbs) is a burst
    # copy figure taken on an idle GPU, and several of them are close to issue-bound, so timing them right after 20 s at the
    # power cap (SM clock ~1.45 GHz for a while) reads up to 1.5x slower (round-2 runs: 0.85 -> 0.52 for the same binary)
    hbm_rates = None
    if world == 1 and rank == 0:
        try:
            hbm_rates = hbm_kernel_rates(dev, (json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if
                                               os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}).get("hbm_gbs"))
            hbm_rates["when"] = "before the step, idle GPU"
        except Exception as ex:  # noqa: BLE001
            hbm_rates = {"error": repr(ex)}
        torch.cuda.empty_cache()

    global LATENT_SHAPE
    cfg = M.WAN_1_3B if args.model == "1.3B" else M.WAN_14B
    if args.model == "14B":
        LATENT_SHAPE = (16, 21, 90, 160)                        # 1280x720x81 -> 75,600 tokens (BASELINE configs[3])
    attn = args.attn
    if attn == "auto":
        attn = "b200q" if hasattr(b200q, "attn_bf16") else "library"
    if attn in ("b200q", "library"):
        M.set_attention_core(attn)
    default_cfg = args.model == "1.3B" and attn in ("b200q", "library") and args.ffn_bits == 8 and args.cfg_batch == 1
    if args.pipeline_chunks == 0 and world > 1:
        from wan_b200.parallel import _largest_head_divisor
        hg = cfg.num_heads // _largest_head_divisor(world, cfg.num_heads)         # heads per head group
        args.pipeline_chunks = 3 if hg % 3 == 0 else (2 if hg % 2 == 0 else 1)
    sp = SequenceParallel(pipeline_chunks=args.pipeline_chunks, exchange=args.exchange) if world > 1 else None
    dit = M.WanDiTQ.random(cfg, seed=0, sp=sp, num_layers=args.layers, attn_quant=(attn == "int8"), ffn_bits=args.ffn_bits)
    L = (LATENT_SHAPE[1] // 1) * (LATENT_SHAPE[2] // 2) * (LATENT_SHAPE[3] // 2)
    B = args.cfg_batch

    g = torch.Generator().manual_seed(0)
    lat_h = torch.randn(*LATENT_SHAPE, generator=g).pin_memory()
    ctx_h = (torch.randn(B, TEXT_TOKENS, TEXT_DIM, generator=g) if B > 1 else torch.randn(TEXT_TOKENS, TEXT_DIM, generator=g)).pin_memory()
    t_h = torch.tensor([500.0]).pin_memory()
    out_h = (torch.empty(B, *LATENT_SHAPE) if B > 1 else torch.empty(*LATENT_SHAPE)).pin_memory()
    lat_d, ctx_d, t_d = lat_h.to(dev), ctx_h.to(dev), t_h.to(dev)

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
    clk = clocks.stop() if clocks is not None else None
    graph_failed = getattr(runner, "failed", None)

    # ---- N > 1: the sharded step against the unsharded step (rank 0 runs it alone) ------------------------------
    verify = None
    if world > 1 and not args.no_verify:
        def compare(y_sharded, y_single):
            a, b = y_sharded.double().flatten(), y_single.double().flatten()
            return {"cosine_vs_unsharded": float((a @ b) / (a.norm() * b.norm())),
                    "max_abs_diff": float((y_sharded - y_single).abs().max()), "max_abs_ref": float(y_single.abs().max()),
                    "bit_equal": bool(torch.equal(y_sharded, y_single))}
        y_sharded = step_resident().clone()                        # the timed configuration (graph replay, key splits per shape)
        # same arithmetic on both sides: attention key splits (a per-shape scheduling choice) switched off
        b200q.attn_bf16_default_splits = 1
        y_sharded_ns = dit.forward(lat_d, t_d, ctx_d).clone()
        sync_all()
        if rank == 0:
            single = M.WanDiTQ(cfg, dit.blocks, dit.fp, sp=None)
            y_single_ns = single.forward(lat_d, t_d, ctx_d)
            verify = {"same_attention_schedule": compare(y_sharded_ns, y_single_ns),
                      "as_timed": compare(y_sharded, y_single_ns),
                      "note": "same_attention_schedule: both sides without attention key splits (expected bit-equal when every "
                              "rank owns whole heads); as_timed: the graph-replayed step with its per-shape key splits"}
            del single, y_single_ns
        b200q.attn_bf16_default_splits = None
        sync_all()

    # ---- per-kernel pass: external timing events (cudaEventRecordExternal nodes) around every quantized GEMM and every
    # attention call inside a second CUDA graph, so the durations contain no host launch gaps.  Single GPU only. ----
    events = {"gemm": [], "attn": []}
    pool, cursor = [], [0]
    orig_qlinear, orig_attn_bf16, orig_attn_i8 = M.qlinear, M.attention_bf16, b200q.attn_i8
    per_kernel_mode = "cuda-graph replay with external event-record nodes around every GEMM and attention launch"

    def wrapped(kind, fn, meta_fn):
        def w(*a, **kw):
            i = cursor[0]
            cursor[0] += 1
            if i == len(pool):
                ext = use_graph and world == 1
                pool.append((torch.cuda.Event(enable_timing=True, external=ext), torch.cuda.Event(enable_timing=True, external=ext),
                             kind, meta_fn(*a, **kw)))
            s, e, _, _ = pool[i]
            s.record()
            y = fn(*a, **kw)
            e.record()
            return y
        return w

    launches0 = b200q.launch_count
    dit.forward(lat_d, t_d, ctx_d)
    launches = (b200q.launch_count - launches0) * args.steps       # this library's launches per step x K
    eager_ms = None
    breakdown = None
    if world == 1:
        try:
            M.qlinear = wrapped("gemm", orig_qlinear, lambda qa, da, rs, w, *a, **kw: (qa.shape[0], w.N, w.K, 2.0 * qa.shape[0] * w.N * w.K))
            M.attention_bf16 = wrapped("attn", orig_attn_bf16, lambda q, k, v, H, **kw: (q.shape[0], k.shape[0], q.shape[1], 4.0 * q.shape[0] * k.shape[0] * q.shape[1]))
            b200q.attn_i8 = wrapped("attn", orig_attn_i8, lambda qq, dq, kq, *a, **kw: (qq.shape[0], kq.shape[0], qq.shape[1], 4.0 * qq.shape[0] * kq.shape[0] * qq.shape[1]))

            class _Counted:
                def forward(self, *a):
                    cursor[0] = 0
                    return dit.forward(*a)
            inst = M.GraphedDiT(_Counted()) if use_graph else _Counted().forward
            inst(lat_d, t_d, ctx_d)
            torch.cuda.synchronize()
            if getattr(inst, "failed", None) is not None:
                raise RuntimeError(inst.failed)
            tot0, tot1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            step_ms = []
            for _ in range(args.steps):
                tot0.record(); inst(lat_d, t_d, ctx_d); tot1.record()
                torch.cuda.synchronize()
                step_ms.append(tot0.elapsed_time(tot1))
                for s_, e_, kind, meta in pool:
                    events[kind].append((s_.elapsed_time(e_), meta))
            inst_ms = sum(step_ms) / len(step_ms)
            g_ms = sum(t_ for t_, _ in events["gemm"]) / args.steps
            a_ms = sum(t_ for t_, _ in events["attn"]) / args.steps
            breakdown = {"attention_ms": a_ms, "quantized_gemm_ms": g_ms, "other_ms": max(0.0, inst_ms - a_ms - g_ms),
                         "instrumented_step_ms": inst_ms,
                         "other": "fused LN/modulate/quant, RMSNorm+RoPE, row quantizers, embeddings, head"}
            del inst
        except Exception as ex:  # noqa: BLE001
            per_kernel_mode = "per-kernel pass failed: %r" % (ex,)
        finally:
            M.qlinear, M.attention_bf16, b200q.attn_i8 = orig_qlinear, orig_attn_bf16, orig_attn_i8

    # extra timings at N=1 on the headline config: the same step with the other attention cores
    variants = {}
    if world == 1 and default_cfg and not args.no_variants and args.layers is None:
        def time_variant(setup, teardown):
            try:
                setup()
                vrun = M.GraphedDiT(dit) if use_graph else dit.forward
                for _ in range(2):
                    vrun(lat_d, t_d, ctx_d)
                torch.cuda.synchronize()
                v0, v1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                v0.record()
                for _ in range(3):
                    vrun(lat_d, t_d, ctx_d)
                v1.record()
                torch.cuda.synchronize()
                return v0.elapsed_time(v1) / 3
            except Exception as ex:  # noqa: BLE001
                return "failed: " + repr(ex)
            finally:
                teardown()

        def set_quant(v):
            for blk in dit.blocks:
                blk.attn_quant = v
        other = "library" if attn == "b200q" else "b200q"
        variants[f"attention_{other}_ms"] = time_variant(lambda: M.set_attention_core(other), lambda: M.set_attention_core(attn))
        variants["attention_int8_ms"] = time_variant(lambda: set_quant(True), lambda: set_quant(False))
        try:                       # CFG: cond + uncond batched into one step (2 branches)
            ctx2 = torch.randn(2, TEXT_TOKENS, TEXT_DIM, device=dev)
            vrun = M.GraphedDiT(dit) if use_graph else dit.forward
            for _ in range(2):
                vrun(lat_d, t_d, ctx2)
            torch.cuda.synchronize()
            v0, v1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            v0.record()
            for _ in range(3):
                vrun(lat_d, t_d, ctx2)
            v1.record()
            torch.cuda.synchronize()
            variants["cfg_batched_2_branches_ms"] = v0.elapsed_time(v1) / 3
            del vrun, ctx2
        except Exception as ex:  # noqa: BLE001
            variants["cfg_batched_2_branches_ms"] = "failed: " + repr(ex)

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
        int8_peak = {}
        if world == 1:
            try:
                int8_peak = measure_int8_peak(dev)
            except Exception as ex:  # noqa: BLE001
                int8_peak = {"error": repr(ex)}
        bf16_sus = peaks.get("bf16_tflops_sustained") or 1400.0
        bf16_src = "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks.get("bf16_tflops_sustained") else "fallback 1.4 PFLOP/s (B200_PROFILING.md)"
        i8_sus = int8_peak.get("int8_tops_sustained")
        i8_peak, i8_src = (i8_sus, "measured in this run: torch._int_mm 8192^3 sustained") if i8_sus else \
            (2.0 * bf16_sus, "2 x bf16 sustained (int8 probe unavailable in this run)")

        def by_shape(evts):
            d = {}
            for t_, meta in evts:
                a = d.setdefault("x".join(map(str, meta[:3])), [0.0, 0.0, 0])
                a[0] += t_; a[1] += meta[3]; a[2] += 1
            return d
        gs, as_ = by_shape(events["gemm"]), by_shape(events["attn"])
        g_ms_tot = sum(v[0] for v in gs.values()); g_ops = sum(v[1] for v in gs.values())
        all_tops = g_ops / (g_ms_tot * 1e-3) / 1e12 if g_ms_tot > 0 else None
        dom_g = max(gs, key=lambda k: gs[k][0]) if gs else None
        dom_a = max(as_, key=lambda k: as_[k][0]) if as_ else None
        traffic_db = {}
        try:
            traffic_db = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        except Exception:  # noqa: BLE001
            pass

        def gemm_roofline():
            if not dom_g:
                return None
            t_, o_, n_ = gs[dom_g]
            ach = o_ / (t_ * 1e-3) / 1e12
            return {"bound": "tensor", "kernel": f"gemm_i8_kernel M,N,K={dom_g} (tcgen05.mma.cta_group::2.kind::i8, TMA, TMEM)",
                    "achieved": ach, "peak": i8_peak, "unit": "TOP/s", "frac": ach / i8_peak if i8_peak else None,
                    "traffic": traffic_db.get("gemm", {}).get(dom_g), "traffic_source": traffic_db.get("source"),
                    "algorithmic_ops_per_launch": o_ / n_, "avg_launch_ms": t_ / n_, "peak_source": i8_src,
                    "frac_of_nominal_4500": ach / 4500.0, "frac_of_int8_burst": ach / int8_peak["int8_tops_burst"] if int8_peak.get("int8_tops_burst") else None,
                    "all_gemms_tops": all_tops, "all_gemms_frac": all_tops / i8_peak if (all_tops and i8_peak) else None,
                    "by_shape_tops": {k: v[1] / (v[0] * 1e-3) / 1e12 for k, v in gs.items() if v[0] > 0}}

        def attn_roofline():
            if not dom_a:
                return None
            t_, o_, n_ = as_[dom_a]
            ach = o_ / (t_ * 1e-3) / 1e12
            own = attn != "library"
            kern = {"b200q": "attn_bf16_kp_kernel (key-pipelined, tcgen05.mma.cta_group::2.kind::f16, TMA, TMEM; this repo)", "int8": "attn_i8_kernel (tcgen05.mma.kind::i8; this repo)",
                    "library": "torch SDPA (cuDNN flash attention; library, not this repo's code)"}[attn]
            # DRAM bytes per launch from the committed ncu --set full capture of the same kernel at the 1.3B self-attention
            # shape (profiles/traffic.json); other shapes have no capture -> null
            tkey = {"b200q": "attn_bf16_H12_L32760", "int8": "attn_i8_H12_L32760"}.get(attn)
            traffic = traffic_db.get(tkey) if (tkey and dom_a == "32760x32760x1536") else None
            return {"bound": "tensor", "kernel": f"{kern} Lq,Lk,D={dom_a}", "own_kernel": own, "achieved": ach, "peak": bf16_sus,
                    "unit": "TFLOP/s", "frac": ach / bf16_sus, "traffic": traffic,
                    "traffic_source": traffic_db.get("source") if traffic else None,
                    "algorithmic_ops_per_launch": o_ / n_, "avg_launch_ms": t_ / n_,
                    "algorithmic_hbm_bytes_per_launch": 4 * 2 * int(dom_a.split("x")[0]) * int(dom_a.split("x")[2]) if "x" in dom_a else None,
                    "softmax": ("per-head on the device: max-free softmax for heads whose Cauchy-Schwarz score bound is <= 80 "
                                "(all heads of this synthetic, unit-RMS workload), online softmax otherwise") if attn == "b200q" else None,
                    "peak_source": bf16_src + " (bf16 dense; flop-equivalents 4*Lq*Lk*D for the int8 kernel)",
                    "frac_of_nominal_2250": ach / 2250.0}
        rl_g, rl_a = gemm_roofline(), attn_roofline()
        # the dominant kernel of the step: attention when it is this repo's kernel (it is ~2/3 of the step), else the heaviest GEMM
        roofline = rl_a if (rl_a and rl_a["own_kernel"] and breakdown and breakdown["attention_ms"] >= breakdown["quantized_gemm_ms"]) else rl_g
        metric = METRIC if args.model == "1.3B" else "W8A8 DiT-step ms (Wan2.1-T2V-14B, 40 blocks, 75600 tokens)"
        workload = WORKLOAD if args.model == "1.3B" else (
            "Wan2.1-T2V-14B 40-block DiT W8A8 denoising step, 1280x720x81 synthetic latent (16x21x90x160 -> 75600 tokens), "
            "one CFG branch; BASELINE.json configs[3]")
        if not default_cfg:
            workload += f" [attention={attn}, ffn weights {args.ffn_bits}-bit, cfg branches per step {B}]"
        attn_desc = {"b200q": "bf16 attention core: this repo's tcgen05 flash attention", "library": "bf16 attention core: library (cuDNN SDPA)",
                     "int8": "int8 tcgen05 attention (P~ grid: one step per query row, key-split merge above 65,536 keys)"}[attn]
        out = {
            "metric": metric, "value": ms, "unit": "ms", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms, "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
            "dtype": "int8 (s32 accumulate; %s)" % attn_desc,
            "data": "synthetic",
            "config": {"workload": workload, "parallelism": f"ulysses{world}" if world > 1 else "single-gpu",
                       "weights": "random-init, replicated", "l2": "per-step working set (>=200 MB per stage) >> 126 MB L2; no flush needed",
                       "cfg_branches_per_step": B, "attention_core": attn},
            "clocks": clk,
            "e2e": {"value": e2e_ms, "unit": "ms", "h2d_bytes_per_step": lat_h.numel() * 4 + ctx_h.numel() * 4 + 4,
                    "d2h_bytes_per_step": out_h.numel() * 4},
            "gpu_launches": launches,
            "launch_mode": ("cuda-graph replay (wan_b200.model.GraphedDiT)" if graph_failed is None else "EAGER: graph capture failed") if use_graph else "eager",
            "graph_capture_failed": graph_failed,
            "eager_ms_per_step": eager_ms,
            "breakdown": breakdown, "per_kernel_timing": per_kernel_mode,
            "roofline": roofline, "roofline_gemm": rl_g, "roofline_attention": rl_a,
            "measured_peaks": {"int8": int8_peak, "bf16_tflops_sustained": peaks.get("bf16_tflops_sustained"),
                               "bf16_tflops_burst": peaks.get("bf16_tflops"), "hbm_gbs": peaks.get("hbm_gbs")},
        }
        if variants:
            out["variants"] = variants
        if verify is not None:
            out["verify"] = verify
        if hbm_rates is not None:
            out["hbm_kernels"] = hbm_rates
        if world > 1:
            b, pu, pr = exchange_bytes_per_rank(L, cfg.dim, world, cfg.num_heads)
            out["config"]["exchange_bytes_per_rank_per_block"] = b
            out["config"]["head_plan"] = f"Pu={pu} x Pr={pr}"
            out["config"]["exchange"] = sp.exchange_in_use
            if sp.exchange_fallback:
                out["config"]["exchange_fallback"] = sp.exchange_fallback
            out["config"]["pipeline_chunks"] = args.pipeline_chunks if sp.exchange_in_use == "nccl" else 1
        if args.layers is not None and args.layers != cfg.num_layers:
            out["invalid"] = f"debug run with {args.layers} of {cfg.num_layers} blocks"
        if world == 1 and not args.no_cpu_baseline:
            try:
                times, kind, desc = cpu_reference_block(L=args.ref_tokens, steps=1, warmup=0, budget_s=60.0)
                blk_ms = sorted(times)[len(times) // 2]
                out["cpu_baseline"] = {"value": blk_ms * 30, "unit": "ms", "cores": os.cpu_count(), "kind": kind, "sample": desc,
                                       "block_ms": blk_ms}
            except Exception as ex:  # noqa: BLE001
                out["cpu_baseline"] = {"value": None, "unit": "ms", "cores": os.cpu_count(), "kind": "port",
                                       "sample": "failed: " + repr(ex)}
        print(json.dumps(out), flush=True)
    sys.stdout.flush()
    sys.stderr.flush()
    # Leave through the interpreter's normal exit so every exit handler runs (the driver records the loaded native
    # libraries there).  Multi-rank: tearing down an NCCL communicator whose send/recvs live inside captured CUDA graphs
    # was observed to hang on B200 (torch 2.11 / NCCL 2.28): drop the graphs first, then destroy the group under a
    # short timer that still runs the exit handlers before it gives up.
    runner = None
    gc.collect()
    torch.cuda.synchronize()
    if world > 1:
        def bail():
            try:
                atexit._run_exitfuncs()
            finally:
                os._exit(0)
        t = threading.Timer(20.0, bail)
        t.daemon = True
        t.start()
        try:
            dist.barrier()
            dist.destroy_process_group()
        except Exception:  # noqa: BLE001
            pass
        t.cancel()


def _watchdog(seconds):
    """A benchmark must never hold a GPU box hostage: hard-exit if the run exceeds its wall-clock budget."""
    def fire():
        sys.stderr.write(f"bench.py: watchdog fired after {seconds} s - aborting\n")
        sys.stderr.flush()
        try:
            atexit._run_exitfuncs()
        finally:
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
    sys.exit(0)
