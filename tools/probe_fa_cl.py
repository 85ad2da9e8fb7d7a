"""CTA-pair (tcgen05.mma.cta_group::2) vs single-CTA attention kernel: correctness against fp32 softmax attention on small
shapes, then sustained timing at H=12, L=32760 with clocks / power.  python tools/probe_fa_cl.py [--quick]"""
import os, subprocess, sys, threading, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "wan2.1-quantization_b200")]
import b200q
from wan_b200 import model as M


def ref(q, k, v, H):
    Lq, D = q.shape
    hd = D // H
    qh, kh, vh = (t.float().view(t.shape[0], H, hd).permute(1, 0, 2) for t in (q, k, v))
    p = torch.softmax((qh @ kh.transpose(1, 2)) * hd ** -0.5, dim=-1)
    return (p @ vh).permute(1, 0, 2).reshape(Lq, D)


CONFIGS = [("two-tile cl=1", 0, 1), ("two-tile cl=2", 0, 2), ("key-pipelined", 1, 2)]
if "--new" in sys.argv:
    CONFIGS = CONFIGS[1:]
if "--kp" in sys.argv:
    CONFIGS = CONFIGS[2:]
ok = True
for name, variant, cl in ([] if "--nocheck" in sys.argv else CONFIGS):
    b200q.attn_bf16_set_cluster(cl)
    b200q.attn_bf16_set_variant(variant)
    for fast in ((3,) if variant else (3, -1)):
        b200q.attn_bf16_set_fast(fast)
        for (H, Lq, Lk) in [(1, 128, 128), (2, 256, 128), (2, 200, 300), (1, 1, 1), (3, 513, 129), (12, 1000, 517), (2, 512, 1024),
                            (1, 128, 4096), (4, 3000, 512), (2, 77, 2000), (3, 5000, 3000)]:
            g = torch.Generator(device="cuda").manual_seed(H * 1000 + Lq + Lk)
            q, k, v = (torch.randn(n, H * 128, device="cuda", generator=g).to(torch.bfloat16) for n in (Lq, Lk, Lk))
            r = ref(q, k, v, H)
            for S in (1, 2):
                if S == 2 and Lk < 2048:
                    continue
                o = b200q.attn_bf16(q, k, v, H, n_splits=S).float()
                torch.cuda.synchronize()
                err = float((o - r).abs().max() / r.abs().max())
                good = err <= 2e-2
                ok &= good
                print(f"{name} fast={fast:2d} H={H:2d} Lq={Lq:5d} Lk={Lk:5d} splits={S}  rel err {err:.2e}  {'ok' if good else 'FAIL'}", flush=True)
b200q.attn_bf16_set_fast(3)
if not ok:
    sys.exit("correctness FAILED")
if "--check" in sys.argv:
    sys.exit(0)

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


def burst(name, fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        time.sleep(0.3)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e))
    print(f"{name:34s} {best:7.3f} ms isolated launch (best of {n})", flush=True)


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
    print(f"{name:34s} {ms:7.3f} ms  {4.0 * L * L * 128 * H / ms / 1e9:7.1f} TFLOP/s   sm {clk:.0f} MHz  {pw:.0f} W  T {out[-1][2]:.0f}C", flush=True)


quick = "--quick" in sys.argv
for name, variant, cl in CONFIGS:
    b200q.attn_bf16_set_cluster(cl)
    b200q.attn_bf16_set_variant(variant)
    for pp in ((3,) if quick else ((0, 1, 2, 3) if "--lowpoly" in sys.argv else (2, 3, 4))):
        b200q.attn_bf16_set_fast(pp)
        burst(f"{name} max-free poly {pp}/8", lambda: b200q.attn_bf16(q, k, v, H))
        run(f"{name} max-free poly {pp}/8", lambda: b200q.attn_bf16(q, k, v, H))
    if not variant:
        b200q.attn_bf16_set_fast(-1)
        run(f"{name} online softmax", lambda: b200q.attn_bf16(q, k, v, H), 2.0)
    b200q.attn_bf16_set_mode(10)
    burst(f"{name} diag: tensor pipeline alone", lambda: b200q.attn_bf16(q, k, v, H))
    run(f"{name} diag: tensor pipeline alone", lambda: b200q.attn_bf16(q, k, v, H), 2.0)
    b200q.attn_bf16_set_mode(2); b200q.attn_bf16_set_fast(3)
burst("library SDPA (cuDNN)", lambda: M.sdpa(q, k, v, H))
run("library SDPA (cuDNN)", lambda: M.sdpa(q, k, v, H))
