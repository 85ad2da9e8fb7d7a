"""BASELINE configs[2]: PTQ calibration statistics pass of Wan2.1-T2V-1.3B — per-input-channel abs-max of all 300 linear
inputs over 30 timesteps x 2 CFG passes = 60 hook calls per layer (get_calib_data_wanx.py:219-275, merge :443-473),
tokens sharded L/P over the ranks, ONE allreduce(MAX) on the flat statistics buffer at the end.

    python tools/bench_calibration.py [--calls 60]                                   # 1 GPU
    torchrun --nproc-per-node P tools/bench_calibration.py                          # P GPUs (sequence-sharded)

Only the statistics kernels + the merge are timed (the FP forward that produces the activations is the reference's own
model and is out of scope); inputs are synthetic buffers of the real shapes and dtypes (fp32 LayerNorm outputs,
bf16 attention / GELU outputs), each larger than L2 except the 512-token text context.  Prints one JSON line."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "wan2.1-quantization_b200")]
import b200q  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--calls", type=int, default=60)
    ap.add_argument("--blocks", type=int, default=30)
    ap.add_argument("--no-graph", action="store_true", help="launch every statistics kernel from Python")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    L, D, F, T = 32760, 1536, 8960, 512
    Lr = L // world
    g = torch.Generator(device=dev).manual_seed(1234)           # same data on every rank; each takes its token slice
    full = {"ln": torch.randn(L, D, device=dev, generator=g), "attn": torch.randn(L, D, device=dev, generator=g).bfloat16(),
            "gelu": torch.randn(L, F, device=dev, generator=g).bfloat16(), "ctx": torch.randn(T, D, device=dev, generator=g)}
    mine = {k: (v if k == "ctx" else v[rank * Lr:(rank + 1) * Lr]) for k, v in full.items()}
    # the ten linears of a block and the tensor each one's hook sees (wan/modules/model.py:293-370)
    layers = [("self_attn.q", "ln"), ("self_attn.k", "ln"), ("self_attn.v", "ln"), ("self_attn.o", "attn"),
              ("cross_attn.q", "ln"), ("cross_attn.k", "ctx"), ("cross_attn.v", "ctx"), ("cross_attn.o", "attn"),
              ("ffn.0", "ln"), ("ffn.2", "gelu")]
    widths = [mine[src].shape[1] for _, src in layers]
    total = a.blocks * sum(widths)
    stats = torch.zeros(total, device=dev)
    offs, o = [], 0
    for _ in range(a.blocks):
        for w in widths:
            offs.append(o); o += w

    def one_timestep():
        i = 0
        for _ in range(a.blocks):
            for (_, src), w in zip(layers, widths):
                b200q.calib_update(mine[src], stats[offs[i]:offs[i] + w])
                i += 1

    one_timestep()
    torch.cuda.synchronize()
    # one hook call per linear = 300 launches per timestep: at 8 ranks a launch covers 4,095 tokens (~2 us of HBM time), so
    # the pass is replayed from a CUDA graph instead of being bounded by Python launch latency
    graph = None
    if not a.no_graph:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            one_timestep()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            one_timestep()
    step = graph.replay if graph is not None else one_timestep
    step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    if world > 1:                        # NCCL channel set-up is not part of the pass
        dist.all_reduce(torch.zeros(total, device=dev), op=dist.ReduceOp.MAX)
        torch.cuda.synchronize()
        dist.barrier()
    s, m, e = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    s.record()
    for _ in range(a.calls):
        step()
    m.record()
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
    e.record()
    torch.cuda.synchronize()
    ms, merge_ms = s.elapsed_time(e), m.elapsed_time(e)
    if world > 1:
        t = torch.tensor([ms, merge_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, merge_ms = float(t[0]), float(t[1])
    # correctness: the merged statistic equals the single-device statistic of the unsharded tensors, bit for bit
    ref = torch.cat([full[src].float().abs().amax(dim=0) for _, src in layers])
    ok = bool(torch.equal(stats[:ref.numel()], ref))
    bytes_per_rank = a.calls * a.blocks * sum(mine[src].numel() * mine[src].element_size() + 12 * mine[src].shape[1] for _, src in layers)
    if rank == 0:
        peak = None
        try:
            peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs")
        except Exception:  # noqa: BLE001
            pass
        gbs = bytes_per_rank / ms / 1e6
        print(json.dumps({"metric": "PTQ calibration statistics pass (Wan2.1-T2V-1.3B, 300 linears x %d calls)" % a.calls,
                          "value": ms, "unit": "ms", "n_gpus": world, "scaling": "strong",
                          "bytes_per_rank": bytes_per_rank, "achieved_gbs_per_gpu": gbs, "peak_gbs": peak,
                          "frac": gbs / peak if peak else None, "merged_equals_single_gpu_statistic": ok,
                          "allreduce_max_ms": merge_ms if world > 1 else 0.0,
                          "launches_per_rank": a.calls * a.blocks * len(layers),
                          "launch_mode": "eager" if a.no_graph else "cuda-graph replay of one timestep (300 launches)", "stats_floats": total,
                          "merge": "one allreduce(MAX) over %d floats" % total if world > 1 else "none (1 GPU)"}), flush=True)
    sys.stdout.flush()
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
