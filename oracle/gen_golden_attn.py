"""Golden vectors for the fused-attention ("row-step") semantics, produced by the IMPORTED, UNMODIFIED reference
quantizers (run in the build container only: `python oracle/gen_golden_attn.py`): q/k/v through DynamicQuantizer
(base_quantizer.py:101-162) with the reshapes of quant_opensora.py:430-442, the attention map through
DynamicQuantizer.forward_with_quant_params (base_quantizer.py:164-206) with delta = row maximum.
Writes tests/golden/quant_attention_rowstep.pt (inputs stored with the outputs)."""
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
    bq = ref["base_quantizer"]
    g = torch.Generator().manual_seed(4321)
    B, H, Lq, Lk, hd = 1, 2, 48, 72, 128
    q = torch.randn(B, H, Lq, hd, generator=g) * 2.0
    k = torch.randn(B, H, Lk, hd, generator=g) * 1.5
    v = torch.randn(B, H, Lk, hd, generator=g)
    cfg = OmegaConf.create({"n_bits": 8, "sym": True})
    zq, zk, zv, zp = (bq.DynamicQuantizer(cfg) for _ in range(4))
    for z in (zq, zk, zv, zp):
        z.module_name = "golden"
    qd = zq(q.reshape([-1, hd])).reshape([B, H, Lq, hd])
    kd = zk(k.reshape([-1, hd])).reshape([B, H, Lk, hd])
    vd = zv(v.permute([0, 1, 3, 2]).reshape([-1, Lk])).reshape([B, H, hd, Lk]).permute([0, 1, 3, 2])
    attn = ((qd * hd ** -0.5) @ kd.transpose(-2, -1)).to(torch.float32).softmax(dim=-1)
    pmax = attn.max(dim=-1, keepdim=True)[0].expand_as(attn).clone()
    attn_q = zp.forward_with_quant_params(attn.clone(), pmax.clone())
    out = attn_q @ vd
    torch.save(dict(q=q, k=k, v=v, q_delta=zq.delta.clone(), k_delta=zk.delta.clone(), v_delta=zv.delta.clone(),
                    attn=attn, attn_quant=attn_q, out=out), os.path.join(OUT, "quant_attention_rowstep.pt"))
    print("quant_attention_rowstep.pt", os.path.getsize(os.path.join(OUT, "quant_attention_rowstep.pt")))


if __name__ == "__main__":
    main()
