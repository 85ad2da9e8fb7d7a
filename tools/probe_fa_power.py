"""Clock / power under sustained attention launches: b200q online softmax, b200q max-free, library SDPA.  Samples nvidia-smi."""
import os, subprocess, sys, threading, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "wan2.1-quantization_b200")]
import b200q
from wan_b200 import model as M

H, L = 12, 32760
g = torch.Generator(device="cuda").manual_seed(0)
q, k, v = (torch.randn(L, H * 128, device="cuda", generator=g).to(torch.bfloat16) for _ in range(3))


def sample(stop, out):
    while not stop.is_set():
        r = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,temperature.gpu", "--format=csv,noheader,nounits", "-i", "0"],
                           capture_output=True, text=True).stdout.strip().split(",")
        try:
            out.append((float(r[0]), float(r[1]), float(r[2])))
        except Exception:
            pass
        time.sleep(0.05)


def run(name, fn, secs=2.5):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    stop, out = threading.Event(), []
    th = threading.Thread(target=sample, args=(stop, out)); th.start()
    n, t0 = 0, time.time()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    while time.time() - t0 < secs:
        for _ in range(20):
            fn()
        n += 20
        torch.cuda.synchronize()
    e.record(); torch.cuda.synchronize()
    stop.set(); th.join()
    ms = s.elapsed_time(e) / n
    out = out[len(out) // 3:]
    clk = sorted(o[0] for o in out)[len(out) // 2]; pw = sorted(o[1] for o in out)[len(out) // 2]
    print(f"{name:28s} {ms:7.3f} ms  {4.0 * L * L * 128 * H / ms / 1e9:7.1f} TFLOP/s   sm {clk:.0f} MHz  {pw:.0f} W  T {out[-1][2]:.0f}C  ({len(out)} samples)", flush=True)


b200q.attn_bf16_set_fast(-1)
run("b200q online softmax poly 2/8", lambda: b200q.attn_bf16(q, k, v, H), 2.0)
for pp in (2, 3, 4):
    b200q.attn_bf16_set_fast(pp)
    run(f"b200q max-free poly {pp}/8", lambda: b200q.attn_bf16(q, k, v, H))
run("library SDPA (cuDNN)", lambda: M.sdpa(q, k, v, H))
b200q.attn_bf16_set_fast(-1); b200q.attn_bf16_set_mode(10)
run("diag: tensor pipeline alone", lambda: b200q.attn_bf16(q, k, v, H))
b200q.attn_bf16_set_mode(2); b200q.attn_bf16_set_fast(3)
