"""Module-tree helpers mirrored from ViDiT-Q/quant_utils/qdiff/utils.py (host logic only)."""
import logging.config
import os
import random

import numpy as np
import torch
import torch.nn as nn


class StraightThrough(nn.Module):
    """utils.py:8-13 — identity placeholder."""

    def __init__(self, channel_num: int = 1):
        super().__init__()

    def forward(self, input):
        return input


def apply_func_to_submodules(module, class_type, function, parent_name="", return_d=None, **kwargs):
    """Depth-first walk over `named_children`; calls `function(child, **kwargs)` for every child that is
    an instance of `class_type`.  The keys 'name', 'full_name', 'parent_module' are filled in per child when
    (and only when) the caller passed them (utils.py:15-50).  Dotted names follow the module tree, e.g.
    `blocks.7.self_attn.q`.  With `return_d` the results are collected by full name and returned."""
    injectable = [k for k in ("name", "full_name", "parent_module") if k in kwargs]
    for child_name, child in list(module.named_children()):
        dotted = child_name if not parent_name else parent_name + "." + child_name
        values = {"name": child_name, "full_name": dotted, "parent_module": module}
        for k in injectable:
            kwargs[k] = values[k]
        if isinstance(child, class_type):
            result = function(child, **kwargs)
            if return_d is not None:
                return_d[dotted] = result
        apply_func_to_submodules(child, class_type, function, dotted, return_d, **kwargs)
    return return_d


def seed_everything(seed=42):
    """utils.py:52-60"""
    random.seed(seed)
    os.environ["PYTHONHASHSEED"] = str(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.benchmark = False


def setup_logging(log_file):
    """utils.py:62-94 — root logger to stdout and to `log_file` (append)."""
    fmt = {"standard": {"format": "%(asctime)s - %(name)s - %(levelname)s - %(message)s"}}
    handlers = {
        "console": {"class": "logging.StreamHandler", "level": "DEBUG", "formatter": "standard",
                    "stream": "ext://sys.stdout"},
        "file": {"class": "logging.FileHandler", "level": "DEBUG", "formatter": "standard",
                 "filename": log_file, "mode": "a"},
    }
    logging.config.dictConfig({"version": 1, "disable_existing_loggers": False, "formatters": fmt,
                               "handlers": handlers,
                               "loggers": {"": {"handlers": ["console", "file"], "level": "DEBUG", "propagate": True}}})
