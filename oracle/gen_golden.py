"""Generate tests/golden/*.pt from the IMPORTED, UNMODIFIED reference (run in the build
container only: `python oracle/gen_golden.py`).  The reference ships no golden vectors
(SURVEY §4), so these known-answer files are produced by executing its own
fake-quant code on seeded inputs; inputs are stored with the outputs so nothing
depends on RNG reproducibility.

Each record: dict(inputs..., outputs...) of CPU tensors, saved with torch.save.
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle.ref_import import import_reference_qdiff  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def main():
    ref = import_reference_qdiff()
    from omegaconf import OmegaConf
    bq, ql, mp, qa = ref["base_quantizer"], ref["quant_layer"], ref["mixed_precision"], ref["quant_attn"]
    os.makedirs(OUT, exist_ok=True)
    g = torch.Generator().manual_seed(1234)

    def act(rows, cols, outliers=True):
        x = torch.randn(rows, cols, generator=g) * (torch.rand(rows, 1, generator=g) * 4 + 0.05)
        if outliers:
            idx = torch.randint(0, cols, (max(1, cols // 100),), generator=g)
            x[:, idx] *= 50.0
        return x

    # ---- dynamic quantizers (base_quantizer.py:101-162) -----------------------------
    dyn = {}
    for name, bits, sym in [("sym8", 8, True), ("asym8", 8, False), ("sym4", 4, True), ("asym4", 4, False)]:
        x = act(48, 200)
        if sym:
            x[5] = 0.0                       # exercises the 1e-6 delta floor (:122-128)
            x[6] = x[6] * 1e-9
        x[7] = x[7].abs()                    # one-sided rows (asym: xmin clamps to 0)
        x[8] = -x[8].abs()
        qz = bq.DynamicQuantizer(OmegaConf.create({"n_bits": bits, "sym": sym}))
        qz.module_name = "golden"
        codes = qz.quantize(x.clone())
        deq = qz.forward(x.clone())
        dyn[name] = dict(x=x, n_bits=bits, sym=sym, codes=codes, delta=qz.delta.clone(),
                         zero_point=qz.zero_point.clone(), dequant=deq)
    # ties: values exactly on k+0.5 steps -> pins round-half-to-even
    x = torch.zeros(4, 256)
    x[0] = torch.arange(-128, 128).float() * 0.5
    x[0, 0] = -63.5
    x[0, -1] = 63.5                              # amax 63.5 -> delta 0.5 exactly -> codes k/1 with .5 ties
    x[1] = torch.arange(256).float() * 0.25 - 31.75
    x[2] = torch.linspace(-1, 1, 256)
    x[3] = torch.arange(256).float() * 1.5 - 190.5
    qz = bq.DynamicQuantizer(OmegaConf.create({"n_bits": 8, "sym": True}))
    qz.module_name = "golden"
    dyn["ties_sym8"] = dict(x=x, n_bits=8, sym=True, codes=qz.quantize(x.clone()), delta=qz.delta.clone(),
                            zero_point=qz.zero_point.clone(), dequant=qz.forward(x.clone()))
    torch.save(dyn, os.path.join(OUT, "dynamic_quantizer.pt"))

    # ---- static weight quantizers (base_quantizer.py:43-99), asym only on CPU (:75) ----
    sta = {}
    for name, bits in [("asym8", 8), ("asym4", 4)]:
        w = (torch.rand(40, 136, generator=g) * 2 - 1) * 0.06
        w[3] = w[3].abs()
        w[4] = -w[4].abs()
        w[5, 17] = 0.9
        qz = bq.StaticQuantizer(OmegaConf.create({"n_bits": bits, "sym": False}))
        deq = qz.forward(w.clone())
        qz.init_done = True
        sta[name] = dict(w=w, n_bits=bits, sym=False, codes=qz.quantize(w.clone()), delta=qz.delta.clone(),
                         zero_point=qz.zero_point.clone(), dequant=deq)
    torch.save(sta, os.path.join(OUT, "static_quantizer.pt"))

    # ---- QuantizedLinear (quant_layer.py:8-74), shipped W8-asym / A8-sym ---------------
    lin = {}
    for name, wb in [("w8a8", 8), ("w4a8", 4)]:
        fp = torch.nn.Linear(136, 40)
        with torch.no_grad():
            fp.weight.copy_((torch.rand(40, 136, generator=g) * 2 - 1) * 0.12)
            fp.bias.copy_(torch.randn(40, generator=g) * 0.02)
        cfg = OmegaConf.create({"weight": {"n_bits": wb, "sym": False}, "act": {"n_bits": 8, "sym": True}})
        w0 = fp.weight.detach().clone()
        layer = ql.QuantizedLinear(136, 40, True, None, cfg, fp)
        layer.a_quantizer.module_name = "golden"
        x = act(2 * 24, 136).reshape(2, 24, 136)
        with torch.no_grad():
            y = layer(x.clone())
        lin[name] = dict(x=x, weight=w0, bias=fp.bias.detach().clone(), w_bits=wb, y=y,
                         w_delta=layer.w_quantizer.delta.detach().clone(),
                         w_zero_point=layer.w_quantizer.zero_point.detach().clone(),
                         w_dequant=layer.weight.detach().clone(),
                         a_delta=layer.a_quantizer.delta.detach().clone())
    torch.save(lin, os.path.join(OUT, "quantized_linear.pt"))

    # ---- mixed precision (mixed_precision_quantizer.py:56-186) ---------------------------
    w = (torch.rand(24, 96, generator=g) * 2 - 1) * 0.1
    rec = dict(w=w)
    for i in (0, 1):
        qz = mp.MixedPrecisionStaticQuantizer(
            OmegaConf.create({"n_bits": [4, 8], "sym": False, "i_bitwidth": i}))
        deq = qz.forward(w.clone())
        rec[f"static_i{i}"] = dict(dequant=deq, delta=qz.delta.clone(), zero_point=qz.zero_point.clone(),
                                   delta_list=qz.delta_list.clone(), zero_point_list=qz.zero_point_list.clone())
        qz.init_done = True
        qz.bitwidth_refactor(1 - i)
        rec[f"static_i{i}_refactored"] = dict(dequant=qz.forward(w.clone()), delta=qz.delta.clone())
    x = act(16, 96)
    for i in (0, 1):
        qz = mp.MixedPrecisionDynamicQuantizer(
            OmegaConf.create({"n_bits": [4, 8], "sym": True, "i_bitwidth": i}))
        rec[f"dynamic_i{i}"] = dict(x=x, dequant=qz.forward(x.clone()), delta=qz.delta.clone())
    torch.save(rec, os.path.join(OUT, "mixed_precision.pt"))

    # ---- calibration hook (examples/Wan2.1/get_calib_data_wanx.py:262-263,443-468; ptq_wanx.py:336)
    calls = [act(40, 72).reshape(1, 40, 72) for _ in range(3)]
    per_call = [c.reshape([-1, 72]).abs().max(dim=0)[0] for c in calls]          # the hook body
    merged = torch.cat([torch.stack(per_call[:2], 0), torch.stack(per_call[2:], 0)], 0).max(dim=0)[0]
    torch.save(dict(calls=calls, per_call=per_call, merged=merged), os.path.join(OUT, "calibration.pt"))

    # ---- quantized attention: glue restated from examples/Wan2.1/models/quant_opensora.py:430-478,
    #      quantizers are the imported DynamicQuantizer / QuantizedAttentionMapOpenSORA('row') ----
    B, H, L, hd = 1, 2, 40, 16
    q = torch.randn(B, H, L, hd, generator=g)
    k = torch.randn(B, H, L, hd, generator=g)
    v = torch.randn(B, H, L, hd, generator=g)
    acfg = OmegaConf.create({"attn": {"qk": {"n_bits": 8, "sym": True, "reorder_file_path": None},
                                      "v": {"n_bits": 8, "sym": True},
                                      "attn_map": {"n_bits": 8, "sym": False, "group": "row"}}})
    qq_, kq_, vq_ = (bq.DynamicQuantizer(acfg.attn.qk), bq.DynamicQuantizer(acfg.attn.qk),
                     bq.DynamicQuantizer(acfg.attn.v))
    for z in (qq_, kq_, vq_):
        z.module_name = "golden"
    pm = qa.QuantizedAttentionMapOpenSORA(acfg)
    pm.attn_map_quantizer.module_name = "golden"
    qd = qq_(q.reshape([-1, hd])).reshape([B, H, L, hd])
    kd = kq_(k.reshape([-1, hd])).reshape([B, H, L, hd])
    vd = vq_(v.permute([0, 1, 3, 2]).reshape([-1, L])).reshape([B, H, hd, L]).permute([0, 1, 3, 2])
    attn = ((qd * hd ** -0.5) @ kd.transpose(-2, -1)).to(torch.float32).softmax(dim=-1)
    attn_q = pm(attn.clone())
    out = attn_q @ vd
    torch.save(dict(q=q, k=k, v=v, q_dequant=qd, k_dequant=kd, v_dequant=vd, attn=attn, attn_quant=attn_q,
                    out=out, q_delta=qq_.delta.clone(), k_delta=kq_.delta.clone(), v_delta=vq_.delta.clone(),
                    p_delta=pm.attn_map_quantizer.delta.clone(), p_zero_point=pm.attn_map_quantizer.zero_point.clone()),
               os.path.join(OUT, "quant_attention.pt"))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
