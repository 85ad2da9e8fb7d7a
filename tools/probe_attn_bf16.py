"""Timing probe: b200q.attn_bf16 vs torch SDPA (cuDNN) at the Wan shapes.  python tools/probe_attn_bf16.py [--big]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "wan2.1-quantization_b200")]
import b200q  # noqa: E402
from wan_b200 import model as M  # noqa: E402


def timed(fn, n):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n


def main():
    dev = torch.device("cuda")
    res = {}
    shapes = [("1.3B self H12 L32760", 12, 32760, 32760, 10), ("1.3B cross H12 L32760 x 512", 12, 32760, 512, 20),
              ("1.3B/8 ranks Pu4xPr2: H3 Lq16380", 3, 16380, 32760, 20), ("ulysses4 H3 L32760", 3, 32760, 32760, 10),
              ("ulysses2 H6 L32760", 6, 32760, 32760, 10)]
    if "--big" in sys.argv:
        shapes.append(("14B self H40 L75600", 40, 75600, 75600, 2))
    for name, H, Lq, Lk, n in shapes:
        g = torch.Generator(device="cuda").manual_seed(0)
        q, k, v = (torch.randn(L, H * 128, device=dev, generator=g).to(torch.bfloat16) for L in (Lq, Lk, Lk))
        flops = 4.0 * Lq * Lk * 128 * H
        # clocks drift under sustained load (power cap): interleave the variants and keep the minimum of three rounds
        modes, lib = {}, 1e9
        for _ in range(3):
            for md in (2, 0, 1, 3):
                b200q.attn_bf16_set_mode(md)
                modes[md] = min(modes.get(md, 1e9), timed(lambda: b200q.attn_bf16(q, k, v, H), n))
            lib = min(lib, timed(lambda: M.sdpa(q, k, v, H), n))
        b200q.attn_bf16_set_mode(2)
        own = modes[2]
        nosplit = timed(lambda: b200q.attn_bf16(q, k, v, H, n_splits=1), n)
        splits = b200q.load().b200q_attn_bf16_splits(Lq, Lk, H)
        o1, o2 = b200q.attn_bf16(q, k, v, H).float(), M.sdpa(q, k, v, H).float()
        res[name] = {"b200q_ms": own, "n_splits": splits, "unsplit_ms": nosplit, "by_poly_pairs_ms": modes, "library_ms": lib, "b200q_tflops": flops / own / 1e9, "library_tflops": flops / lib / 1e9,
                     "max_abs_diff": float((o1 - o2).abs().max())}
        print(name, json.dumps(res[name]), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "probe_attn_bf16.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
