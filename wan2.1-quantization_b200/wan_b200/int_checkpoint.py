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
from .model import LINEARS, DenseRotation, FPWeight, QWeight, WanBlockQ, WanConfig, WanDiTQ

META_KEY = "__b200q_meta__"
_WRAPPERS = {"_fsdp_wrapped_module", "module", "_orig_mod"}


def strip_wrapper_prefixes(name: str) -> str:
    """Drop FSDP / DDP / torch.compile wrapper COMPONENTS of a dotted key (whole path components only: `fp_module.` is
    not `module.`)."""
    return ".".join(p for p in name.split(".") if p not in _WRAPPERS)


def export_int_state_dict(model, device=None) -> dict:
    """state_dict of `model` with every qdiff QuantizedLinear replaced by its integer form (keys above).  Layers the
    config keeps in floating point (remain_fp_regex: never wrapped; mixed-precision index 0: `quant_mode` False) keep
    their FP weight.  SmoothQuant / QuaRot / ViDiT-Q layers also store what they apply to their activations:
    `<layer>.channel_mask` (fp32 [C_in]) and `<layer>.rotation_sign` (int8 [C_in], R = diag(sign).H_n/sqrt(n)) or, for a
    rotation that is not of that form, the dense `<layer>.rotation_matrix` - the codes are those of the scaled / rotated
    weight, so the runtime must apply the same transform (reference defect B-4, SURVEY appendix B: its exporter drops
    it)."""
    from qdiff.base.quant_layer import QuantizedLinear
    dev = device or torch.device("cuda", torch.cuda.current_device())
    sd = {}
    qnames, fpnames = {}, {}
    for name, mod in model.named_modules():
        if isinstance(mod, QuantizedLinear):
            if mod.w_quantizer is not None and mod.a_quantizer is not None and mod.quant_mode:
                qnames[strip_wrapper_prefixes(name)] = mod
            else:
                fpnames[strip_wrapper_prefixes(name)] = mod
    for k, v in model.state_dict().items():
        if "fp_weight" in k or "fp_module" in k or ".w_quantizer." in k or ".a_quantizer." in k:
            continue
        k = strip_wrapper_prefixes(k)
        layer = k.rsplit(".", 1)[0]
        if (layer in qnames or layer in fpnames) and k.endswith(".weight"):
            continue                                   # replaced below
        sd[k] = v.detach().cpu()
    for name, mod in fpnames.items():
        sd[f"{name}.weight"] = mod.fp_module.weight.detach().cpu()
    bits, infeat = {}, {}
    for name, mod in qnames.items():
        infeat[name] = int(mod.in_features)
        st = mod.int_weight_state(dev)
        if st["packed"] is not None:
            sd[f"{name}.weight_packed"] = st["packed"].contiguous().cpu()
        else:
            sd[f"{name}.weight"] = st["codes"].contiguous().cpu()
        sd[f"{name}.scale_weight"] = st["delta"].cpu()
        if st["zp"] is not None:
            sd[f"{name}.zp_weight"] = st["zp"].cpu()
        bits[name] = int(st["n_bits"])
        if getattr(mod, "channel_mask", None) is not None:
            sd[f"{name}.channel_mask"] = mod.channel_mask.detach().float().cpu()
        if getattr(mod, "rotation_matrix", None) is not None:
            sign = mod.rotation_sign() if hasattr(mod, "rotation_sign") else None
            if sign is not None:
                sd[f"{name}.rotation_sign"] = sign.detach().cpu().to(torch.int8)
            else:
                sd[f"{name}.rotation_matrix"] = mod.rotation_matrix.detach().float().cpu()
    # the reference's general LayerNorm kernel wants explicit unit weights (quant_wanx.py:170-175)
    blocks = sorted({int(k.split(".")[1]) for k in sd if k.startswith("blocks.") and k.split(".")[1].isdigit()})
    for i in blocks:
        dim = sd[f"blocks.{i}.modulation"].shape[-1] if f"blocks.{i}.modulation" in sd else None
        if dim is not None:
            sd.setdefault(f"blocks.{i}.norm1.weight", torch.ones(dim))
            sd.setdefault(f"blocks.{i}.norm2.weight", torch.ones(dim))
    sd[META_KEY] = {"format": "b200q-int-weight", "version": 2, "scale_dtype": "float32", "weight_bits": bits,
                    "in_features": infeat, "w4_packing": "b200q_pack_w4"}
    return sd


def save_int_checkpoint(model, path, device=None):
    sd = export_int_state_dict(model, device)
    torch.save(sd, path)
    return sd


def _act_transform(sd, name, n, device):
    """<name>.{channel_mask, rotation_sign | rotation_matrix} -> the activation transform the runtime applies in front
    of the layer's quantizer (fused kernel plan where the size allows, dense fp32 product otherwise), or None."""
    from qdiff.base.quant_layer import ActPlan
    mask = sd.get(f"{name}.channel_mask")
    sign = sd.get(f"{name}.rotation_sign")
    R = sd.get(f"{name}.rotation_matrix")
    if mask is None and sign is None and R is None:
        return None
    if R is not None:
        return DenseRotation(R, mask, device)
    if sign is not None:
        plan = ActPlan.rotation(n, sign.float(), mask, device)
        if plan is not None:
            return plan
        from qdiff.quarot.quarot_utils import matmul_hadU
        return DenseRotation(matmul_hadU(torch.diag(sign.double())), mask, device)
    plan = ActPlan.scale_only(mask, device)
    return plan if plan is not None else DenseRotation(None, mask, device)


def _layer(sd, name, device):
    """<name>.{weight | weight_packed, scale_weight, zp_weight, bias, ...} -> QWeight, or FPWeight for a layer the
    checkpoint keeps in floating point."""
    meta = sd.get(META_KEY, {})
    bias = sd.get(f"{name}.bias")
    if f"{name}.scale_weight" not in sd:
        w = sd.get(f"{name}.weight")
        if w is None or not w.dtype.is_floating_point:
            raise b200q.B200QError(f"{name}: neither an integer layer (scale_weight missing) nor an FP weight in the checkpoint")
        return FPWeight(w, bias)
    n_bits = meta.get("weight_bits", {}).get(name, 8)
    delta = sd[f"{name}.scale_weight"].to(device=device, dtype=torch.float32).reshape(-1).contiguous()
    zp = sd.get(f"{name}.zp_weight")
    zp = None if zp is None else zp.to(device=device, dtype=torch.float32).reshape(-1).contiguous()
    bias = None if bias is None else bias.to(device=device, dtype=torch.float32).contiguous()
    if f"{name}.weight_packed" in sd:
        packed = sd[f"{name}.weight_packed"]
        if packed.dtype != torch.uint8:
            raise b200q.B200QError(f"{name}.weight_packed is {packed.dtype}, expected uint8 (b200q_pack_w4 format)")
        nbytes = packed.shape[1]
        K = meta.get("in_features", {}).get(name, nbytes * 2)           # 8 codes per 4 bytes
        pitch = (nbytes + 15) // 16 * 16                                 # TMA global stride: multiple of 16 bytes
        buf = torch.zeros((packed.shape[0], pitch), dtype=torch.uint8, device=device)
        buf[:, :nbytes] = packed.to(device)
        return QWeight(None, delta, zp, bias, n_bits, buf[:, :nbytes], K=K, pre=_act_transform(sd, name, K, device))
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
    return QWeight(codes, delta, zp, bias, n_bits, packed, pre=_act_transform(sd, name, K, device))


def dit_from_int_state_dict(cfg: WanConfig, sd: dict, device=None, sp=None, attn_quant=False) -> WanDiTQ:
    """Build the integer runtime (wan_b200.model.WanDiTQ) from an int-weight state dict: the counterpart of
    hardware_forward_refactor's step (3) (quant_wanx.py:221-228).  FP parts (patch/text/time embeddings, head) are taken
    as stored (remain_fp_regex keeps them FP, quant_configs/config.yaml:9); block linears are integer or FP layer by
    layer, as the checkpoint has them."""
    dev = device or torch.device("cuda", torch.cuda.current_device())
    sd = {(k if k == META_KEY else strip_wrapper_prefixes(k)): v for k, v in sd.items()}
    f32 = lambda k: sd[k].to(device=dev, dtype=torch.float32).contiguous()
    bf = lambda k: sd[k].to(device=dev, dtype=torch.bfloat16).contiguous()
    blocks = []
    for i in range(cfg.num_layers):
        b = f"blocks.{i}."
        w = {n: _layer(sd, b + n, dev) for n in LINEARS}
        w["modulation"] = f32(b + "modulation")
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
