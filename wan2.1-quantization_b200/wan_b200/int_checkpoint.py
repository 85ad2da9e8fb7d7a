"""Integer-weight checkpoint: export from a PTQ'd `qdiff` model, load into the B200 runtime.

Reference flow (examples/Wan2.1): ptq_wanx.py:230-258 calls `QuantWanModel.quantize_and_save_weight(path)`
(wan/quant_wanx.py:137-185 + `quantize_and_save_weight_`, wan/quant_wanx_cuda.py:39-55): every QuantizedLinear's
`weight` becomes int8 codes, `w_quantizer.delta` / `.zero_point` are renamed `scale_weight` / `zp_weight`, `fp_weight` /
`fp_module` entries are dropped, `blocks.i.norm{1,2}.weight` = ones are added; quant_generate.py:381-392 then loads the
file through `hardware_forward_refactor` (quant_wanx.py:188-228).

Same key schema here, with three differences (all recorded in the `__b200q_meta__` entry):
  * scales stay fp32 (the reference rounds delta/zero_point AND the FP weight to fp16 before re-deriving the codes,
    so its int codes differ from its own fake-quant path; here codes == the fake-quant codes, bit for bit),
  * 4-bit layers (mixed-precision configs) are stored packed (`<layer>.weight_packed`, b200q_pack_w4 format) with
    `<layer>.weight_bits = 4`; the reference's exporter is W8A8 only (quant_wanx_cuda.py:51),
  * FSDP / DDP wrapper prefixes are stripped from the keys (the reference's loader trips over them).
"""
from __future__ import annotations

import math

import torch

import b200q
from .model import QWeight, WanBlockQ, WanConfig, WanDiTQ

META_KEY = "__b200q_meta__"
_PREFIXES = ("_fsdp_wrapped_module.", "module.", "_orig_mod.")


def strip_wrapper_prefixes(name: str) -> str:
    parts = name
    changed = True
    while changed:
        changed = False
        for p in _PREFIXES:
            if p in parts:
                parts = parts.replace(p, "")
                changed = True
    return parts


def export_int_state_dict(model, device=None) -> dict:
    """state_dict of `model` with every qdiff QuantizedLinear replaced by its integer form (keys above)."""
    from qdiff.base.quant_layer import QuantizedLinear
    dev = device or torch.device("cuda", torch.cuda.current_device())
    sd = {}
    qnames = {}
    for name, mod in model.named_modules():
        if isinstance(mod, QuantizedLinear) and mod.w_quantizer is not None:
            qnames[strip_wrapper_prefixes(name)] = mod
    for k, v in model.state_dict().items():
        k = strip_wrapper_prefixes(k)
        if "fp_weight" in k or "fp_module" in k or ".w_quantizer." in k or ".a_quantizer." in k:
            continue
        layer = k.rsplit(".", 1)[0]
        if layer in qnames and k.endswith(".weight"):
            continue                                   # replaced below
        sd[k] = v.detach().cpu()
    bits = {}
    for name, mod in qnames.items():
        st = mod.int_weight_state(dev)
        if st["packed"] is not None:
            sd[f"{name}.weight_packed"] = st["packed"].contiguous().cpu()
        else:
            sd[f"{name}.weight"] = st["codes"].contiguous().cpu()
        sd[f"{name}.scale_weight"] = st["delta"].cpu()
        if st["zp"] is not None:
            sd[f"{name}.zp_weight"] = st["zp"].cpu()
        bits[name] = int(st["n_bits"])
        variant = {}
        if getattr(mod, "channel_mask", None) is not None:
            variant["channel_mask"] = mod.channel_mask.detach().float().cpu()
        if getattr(mod, "rotation_matrix", None) is not None:
            variant["rotation_matrix"] = mod.rotation_matrix.detach().cpu()
        for kk, vv in variant.items():
            sd[f"{name}.{kk}"] = vv
    # the reference's general LayerNorm kernel wants explicit unit weights (quant_wanx.py:170-175)
    blocks = sorted({int(k.split(".")[1]) for k in sd if k.startswith("blocks.") and k.split(".")[1].isdigit()})
    for i in blocks:
        dim = sd[f"blocks.{i}.modulation"].shape[-1] if f"blocks.{i}.modulation" in sd else None
        if dim is not None:
            sd.setdefault(f"blocks.{i}.norm1.weight", torch.ones(dim))
            sd.setdefault(f"blocks.{i}.norm2.weight", torch.ones(dim))
    sd[META_KEY] = {"format": "b200q-int-weight", "version": 1, "scale_dtype": "float32", "weight_bits": bits,
                    "w4_packing": "b200q_pack_w4"}
    return sd


def save_int_checkpoint(model, path, device=None):
    sd = export_int_state_dict(model, device)
    torch.save(sd, path)
    return sd


def _qweight(sd, name, device):
    """<name>.{weight|weight_packed, scale_weight, zp_weight, bias} -> QWeight on `device`."""
    meta = sd.get(META_KEY, {})
    n_bits = meta.get("weight_bits", {}).get(name, 8)
    delta = sd[f"{name}.scale_weight"].to(device=device, dtype=torch.float32).reshape(-1).contiguous()
    zp = sd.get(f"{name}.zp_weight")
    zp = None if zp is None else zp.to(device=device, dtype=torch.float32).reshape(-1).contiguous()
    bias = sd.get(f"{name}.bias")
    bias = None if bias is None else bias.to(device=device, dtype=torch.float32).contiguous()
    if f"{name}.weight_packed" in sd:
        raise NotImplementedError("packed 4-bit layers are loaded through their int8 codes: export with codes present")
    codes = sd[f"{name}.weight"]
    if codes.dtype != torch.int8:
        raise b200q.B200QError(f"{name}.weight is {codes.dtype}, expected int8 codes (is this an FP checkpoint?)")
    codes = codes.to(device)
    K = codes.shape[1]
    if K % 16 != 0:
        padded = torch.zeros((codes.shape[0], (K + 15) // 16 * 16), dtype=torch.int8, device=device)
        padded[:, :K] = codes
        codes = padded[:, :K]
    packed = b200q.pack_w4(codes) if n_bits <= 4 else None
    return QWeight(codes, delta, zp, bias, n_bits, packed)


def dit_from_int_state_dict(cfg: WanConfig, sd: dict, device=None, sp=None, attn_quant=False) -> WanDiTQ:
    """Build the integer runtime (wan_b200.model.WanDiTQ) from an int-weight state dict: the counterpart of
    hardware_forward_refactor's step (3) (quant_wanx.py:221-228).  FP parts (patch/text/time embeddings, head) are taken
    as stored (remain_fp_regex keeps them FP, quant_configs/config.yaml:9)."""
    dev = device or torch.device("cuda", torch.cuda.current_device())
    sd = {strip_wrapper_prefixes(k): v for k, v in sd.items()}
    f32 = lambda k: sd[k].to(device=dev, dtype=torch.float32).contiguous()
    bf = lambda k: sd[k].to(device=dev, dtype=torch.bfloat16).contiguous()
    blocks = []
    for i in range(cfg.num_layers):
        b = f"blocks.{i}."
        q = lambda n: _qweight(sd, b + n, dev)
        w = {
            "self_attn.qkv": QWeight.cat([q("self_attn.q"), q("self_attn.k"), q("self_attn.v")]),
            "self_attn.o": q("self_attn.o"),
            "cross_attn.q": q("cross_attn.q"),
            "cross_attn.kv": QWeight.cat([q("cross_attn.k"), q("cross_attn.v")]),
            "cross_attn.o": q("cross_attn.o"),
            "ffn.0": q("ffn.0"), "ffn.2": q("ffn.2"),
            "modulation": f32(b + "modulation"),
        }
        for k in ("self_attn.norm_q.weight", "self_attn.norm_k.weight", "cross_attn.norm_q.weight",
                  "cross_attn.norm_k.weight", "norm3.weight", "norm3.bias"):
            if b + k in sd:
                w[k] = f32(b + k)
        blocks.append(WanBlockQ(cfg, w, attn_quant=attn_quant))
    pdim = cfg.in_dim * math.prod(cfg.patch_size)
    fp = {
        "patch": (bf("patch_embedding.weight").reshape(cfg.dim, pdim), bf("patch_embedding.bias")),
        "text0": (bf("text_embedding.0.weight"), bf("text_embedding.0.bias")),
        "text2": (bf("text_embedding.2.weight"), bf("text_embedding.2.bias")),
        "time0": (bf("time_embedding.0.weight"), bf("time_embedding.0.bias")),
        "time2": (bf("time_embedding.2.weight"), bf("time_embedding.2.bias")),
        "timeproj": (bf("time_projection.1.weight"), bf("time_projection.1.bias")),
        "head": (bf("head.head.weight"), bf("head.head.bias")),
        "head_mod": f32("head.modulation").reshape(2, cfg.dim),
    }
    return WanDiTQ(cfg, blocks, fp, sp)


def load_int_checkpoint(cfg: WanConfig, path, device=None, sp=None, attn_quant=False) -> WanDiTQ:
    sd = torch.load(path, map_location="cpu", weights_only=False)
    return dit_from_int_state_dict(cfg, sd, device, sp, attn_quant)
