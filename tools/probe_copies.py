"""Where do the elementwise copy kernels of a DiT step come from?  torch.profiler with stacks over a 2-block step."""
import os, sys, collections, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "wan2.1-quantization_b200")]
import b200q
from wan_b200 import model as M
import dataclasses
cfg = M.WAN_1_3B
NL = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dit = M.WanDiTQ.random(cfg, seed=0, num_layers=NL)
M.set_attention_core("b200q")
g = torch.Generator(device="cuda").manual_seed(0)
lat = torch.randn(16, 21, 60, 104, device="cuda", generator=g)
ctx = torch.randn(512, cfg.text_dim, device="cuda", generator=g)
t = torch.tensor([500.0], device="cuda")
for _ in range(2):
    dit.forward(lat, t, ctx)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True) as prof:
    dit.forward(lat, t, ctx)
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
kern = collections.defaultdict(lambda: [0, 0.0])
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        kern[ev.name[:90]][0] += 1; kern[ev.name[:90]][1] += ev.device_time_total
for ev in prof.events():
    if ev.device_time_total > 0 and ev.name.startswith("aten::") and ev.name in ("aten::copy_", "aten::add", "aten::mul", "aten::cat", "aten::_to_copy", "aten::contiguous", "aten::clone", "aten::add_", "aten::index", "aten::fill_", "aten::zero_"):
        st = [s for s in (ev.stack or []) if "wan_b200" in s or "b200q/__init__" in s]
        key = (ev.name, st[0].split("/")[-1] if st else "?")
        agg[key][0] += 1; agg[key][1] += ev.device_time_total
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
    print(f"{v[1]:9.1f} us x{v[0]:3d}  {k[0]:18s} {k[1]}")
print(f"--- device kernels of a {NL}-block forward (top 25 by time)")
for k, v in sorted(kern.items(), key=lambda kv: -kv[1][1])[:25]:
    print(f"{v[1]:9.1f} us x{v[0]:3d}  {k}")
