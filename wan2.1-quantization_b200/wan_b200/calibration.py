"""Calibration pass: per-input-channel abs-max of every nn.Linear input, accumulated ON DEVICE by the single-read
reduction kernel and merged across sequence-parallel ranks with one allreduce(MAX).

Replaces ViDiT-Q/examples/Wan2.1/get_calib_data_wanx.py:219-275 (SaveActivationHook: `abs().max(dim=0)` = two extra
full passes per linear per call, one [C] tensor appended per call) and :443-473 (stack -> .cpu() -> pickled
all_gather_object -> cat).  Its only consumer takes `.max(dim=0)[0]` over calls (ptq_wanx.py:336), so a running
elementwise max is the same statistic bit-for-bit (max is exact and order independent).

`state_dict()` returns the reference's file format {layer_name: tensor[n_rows, C_in]} with n_rows == 1 (the running
max), which `calib_data[full_name].max(dim=0)[0]` consumes unchanged."""
from __future__ import annotations

import torch
import torch.nn as nn

import b200q
from qdiff.utils import apply_func_to_submodules


class CalibrationCollector:
    def __init__(self, model: nn.Module, device=None, with_minmax=False):
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.names, self.offsets, self.handles = [], {}, []
        total = 0
        layers = apply_func_to_submodules(model, nn.Linear, lambda m, full_name: m, return_d={}, full_name=None)
        for name, mod in layers.items():
            clean = name.replace("_fsdp_wrapped_module.", "")        # get_calib_data_wanx.py:447
            self.offsets[clean] = (total, mod.in_features)
            self.names.append(clean)
            total += mod.in_features
            self.handles.append(mod.register_forward_hook(self._make_hook(clean)))
        # one flat buffer for all layers: a single allreduce merges every statistic (1.3B: 2.7 MB, 14B: 9.6 MB)
        self.absmax = torch.zeros(total, dtype=torch.float32, device=self.device)
        self.with_minmax = with_minmax
        self.xmin = torch.full((total,), float("inf"), device=self.device) if with_minmax else None
        self.xmax = torch.full((total,), float("-inf"), device=self.device) if with_minmax else None
        self.n_calls = {n: 0 for n in self.names}

    def _make_hook(self, name):
        def hook(module, module_in, module_out):
            x = module_in[0]
            off, C = self.offsets[name]
            x2 = x.reshape(-1, C)
            b200q.calib_update(x2, self.absmax[off:off + C],
                               self.xmin[off:off + C] if self.with_minmax else None,
                               self.xmax[off:off + C] if self.with_minmax else None)
            self.n_calls[name] += 1
        return hook

    def update(self, name, x2d):
        """Direct (hook-less) accumulation for runtimes that do not go through nn.Linear modules."""
        off, C = self.offsets[name]
        b200q.calib_update(x2d, self.absmax[off:off + C])
        self.n_calls[name] += 1

    def merge(self, sp=None):
        """Cross-rank merge == allreduce(MAX) on the flat buffer (and MIN for xmin)."""
        import torch.distributed as dist
        if sp is not None and sp.world_size > 1:
            sp.allreduce_max(self.absmax)
            if self.with_minmax:
                dist.all_reduce(self.xmax, op=dist.ReduceOp.MAX, group=sp.group)
                dist.all_reduce(self.xmin, op=dist.ReduceOp.MIN, group=sp.group)
        return self

    def remove_hooks(self):
        for h in self.handles:
            h.remove()
        self.handles = []

    def state_dict(self):
        return {n: self.absmax[o:o + c].detach().cpu().unsqueeze(0) for n, (o, c) in self.offsets.items()}

    def save(self, path, rank=0):
        if rank == 0:
            torch.save(self.state_dict(), path)
