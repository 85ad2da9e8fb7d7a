"""Per-phase timing of one sequence-parallel attention (wan_b200.parallel.SequenceParallel.attention) on N ranks:
staging copies, q|k|v exchange, attention core, o exchange.  torchrun --nproc-per-node N tools/probe_exchange.py [1.3B|14B]"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "wan2.1-quantization_b200")]
import b200q  # noqa: E402
from wan_b200 import model as M  # noqa: E402
from wan_b200.parallel import SequenceParallel  # noqa: E402


def main():
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    big = len(sys.argv) > 1 and sys.argv[1] == "14B"
    H, L = (40, 75600) if big else (12, 32760)
    D = H * 128
    Lr = L // world
    M.set_attention_core("b200q")
    sp = SequenceParallel()
    g = torch.Generator(device="cuda").manual_seed(rank)
    q, k = (torch.randn(Lr, D, device=dev, generator=g).to(torch.bfloat16) for _ in range(2))
    qkv = torch.randn(Lr, 3 * D, device=dev, generator=g).to(torch.bfloat16)
    v = qkv[:, 2 * D:]
    Pu, Pr, gidx, h = sp.plan(H)
    W = D // Pu

    def timed(fn, n=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize(); dist.barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(n):
            fn()
        e.record()
        torch.cuda.synchronize()
        t = torch.tensor([s.elapsed_time(e) / n], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    res = {"world": world, "H": H, "L": L, "plan": f"Pu={Pu} Pr={Pr}"}
    res["full_attention_ms"] = timed(lambda: sp.attention(q, k, v, H))
    gm = lambda x: x.view(Lr, Pu, W).permute(1, 0, 2).contiguous()
    res["staging_3_copies_ms"] = timed(lambda: (gm(q), gm(k), gm(v)))
    Q = torch.randn(Pu * Lr, W, device=dev, generator=g).to(torch.bfloat16)
    K = torch.randn(world * Lr, W, device=dev, generator=g).to(torch.bfloat16)
    V = torch.randn(world * Lr, W, device=dev, generator=g).to(torch.bfloat16)
    res["core_ms"] = timed(lambda: M.attention_bf16(Q, K, V, H // Pu))
    # raw exchange of the same byte volume: all_to_all_single of [P, Lr, 3W] (q|k|v) and [Pu.., Lr, W] (o)
    a = torch.empty(world, Lr, 3 * W, device=dev, dtype=torch.bfloat16)
    b = torch.empty_like(a)
    res["a2a_single_qkv_ms"] = timed(lambda: dist.all_to_all_single(b, a))
    a1 = torch.empty(world, Lr, W, device=dev, dtype=torch.bfloat16)
    b1 = torch.empty_like(a1)
    res["a2a_single_o_ms"] = timed(lambda: dist.all_to_all_single(b1, a1))
    res["qkv_bytes_per_rank"] = a.numel() * 2 * (world - 1) // world
    if rank == 0:
        print(json.dumps(res), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
