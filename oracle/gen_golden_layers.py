"""Golden vectors for the SmoothQuant / QuaRot / ViDiT-Q layer variants, produced by the IMPORTED, UNMODIFIED reference
layers (run in the build container only: `python oracle/gen_golden_layers.py`).
  ViDiT-Q/quant_utils/qdiff/smooth_quant/sq_quant_layer.py, quarot/quarot_quant_layer.py, viditq/viditq_quant_layer.py
The reference's get_rotation_matrix hard-codes "cuda" (quarot_quant_layer.py:28), so the rotation matrix is assigned
directly (any diag(+-1).H/sqrt(n) is valid and it is regenerated on load anyway, quant_model.py:145-152); everything
else — get_channel_mask, update_quantized_weight_*, forward — is the reference's own code on CPU, fp32.
Writes tests/golden/variant_layers.pt (inputs stored with the outputs)."""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle.ref_import import import_reference_qdiff  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def rotation(n, g):
    """diag(+-1) . H_n / sqrt(n) for n = 12 * 2^m (Paley order 12, Sylvester doubling) - built here, not imported."""
    q = 11
    res = {(x * x) % q for x in range(1, q)}
    chi = torch.tensor([0] + [1 if x in res else -1 for x in range(1, q)], dtype=torch.float64)
    Q = chi[(torch.arange(q).view(-1, 1) - torch.arange(q).view(1, -1)) % q]
    S = torch.zeros(12, 12, dtype=torch.float64)
    S[0, 1:] = 1; S[1:, 0] = -1; S[1:, 1:] = Q
    H = torch.eye(12, dtype=torch.float64) + S
    while H.shape[0] < n:
        H = torch.kron(torch.tensor([[1., 1.], [1., -1.]], dtype=torch.float64), H)
    assert H.shape[0] == n and torch.equal(H @ H.t(), n * torch.eye(n, dtype=torch.float64))
    s = torch.randint(0, 2, (n,), generator=g).double() * 2 - 1
    return torch.diag(s) @ H / n ** 0.5


def main():
    import_reference_qdiff()
    from omegaconf import OmegaConf
    import torch.nn as nn
    from qdiff.smooth_quant.sq_quant_layer import SQQuantizedLinear
    from qdiff.quarot.quarot_quant_layer import QuarotQuantizedLinear
    from qdiff.viditq.viditq_quant_layer import ViDiTQuantizedLinear
    g = torch.Generator().manual_seed(77)
    cin, cout, B, N = 96, 40, 2, 12
    rec = {}
    for name, cls, extra in (("smooth_quant", SQQuantizedLinear, {"smooth_quant": {"alpha": 0.6}}),
                             ("quarot", QuarotQuantizedLinear, {"quarot": {}}),
                             ("viditq", ViDiTQuantizedLinear, {"viditq": {"alpha": 0.75}})):
        for w_bits in (8, 4):
            cfg = OmegaConf.create({"weight": {"n_bits": w_bits, "sym": False}, "act": {"n_bits": 8, "sym": True}, **extra})
            fp = nn.Linear(cin, cout)
            with torch.no_grad():
                fp.weight.copy_((torch.rand(cout, cin, generator=g) * 2 - 1) * 0.2)
                fp.weight[:, 5] *= 8                                  # an outlier input channel
                fp.bias.copy_(torch.randn(cout, generator=g) * 0.05)
            layer = cls(cin, cout, True, None, cfg, fp)
            layer.w_quantizer.module_name = layer.a_quantizer.module_name = "golden"
            x = torch.randn(B, N, cin, generator=g)
            x[..., 5] *= 20
            act_mask = x.reshape(-1, cin).abs().max(dim=0)[0].clamp_min(1e-3)     # ptq_wanx.py:336-341
            r = dict(weight=fp.weight.detach().clone(), bias=fp.bias.detach().clone(), x=x, act_mask=act_mask,
                     w_bits=w_bits, cfg=OmegaConf.to_container(cfg) if hasattr(OmegaConf, "to_container") else dict(cfg))
            if hasattr(layer, "channel_mask"):
                layer.get_channel_mask(act_mask)
                r["channel_mask"] = layer.channel_mask.clone()
            if hasattr(layer, "rotation_matrix"):
                layer.rotation_matrix = rotation(cin, g)
                r["rotation_matrix"] = layer.rotation_matrix.clone()
            if name == "smooth_quant":
                layer.update_quantized_weight_scaled()
            elif name == "quarot":
                # quarot_quant_layer.py:30-45 with its hard-coded `.to("cuda")` (:33) left out: same three statements
                layer.w_quantizer.init_done = False
                layer.weight.data = layer.w_quantizer(torch.matmul(layer.fp_module.weight.data.double(), layer.rotation_matrix).float())
                layer.w_quantizer.init_done = True
            else:
                layer.update_quantized_weight_rotated_and_scaled()
            with torch.no_grad():
                y = layer(x)
            r.update(fq_weight=layer.weight.detach().clone(), w_delta=layer.w_quantizer.delta.clone(),
                     w_zero_point=layer.w_quantizer.zero_point.clone(), y=y.detach().clone())
            rec[f"{name}_w{w_bits}"] = r
    torch.save(rec, os.path.join(OUT, "variant_layers.pt"))
    print("variant_layers.pt", os.path.getsize(os.path.join(OUT, "variant_layers.pt")), sorted(rec))


if __name__ == "__main__":
    main()
