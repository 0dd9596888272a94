"""Writes tests/golden/lr_sched.npz: the reference's util/lr_sched.py:adjust_learning_rate evaluated on a grid of fractional
epochs for the fine-tune and the pre-train defaults (FSC_finetune_cross.py:58-66, FSC_pretrain.py:58-66)."""
import importlib.util
import os
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("ref_lr_sched", "/root/reference/util/lr_sched.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

out = {}
for name, (lr, min_lr, warm, epochs) in {"finetune": (1e-5, 0.0, 10, 1000), "pretrain": (1.5e-4 * 256 / 256, 1e-6, 10, 300)}.items():
    args = types.SimpleNamespace(lr=lr, min_lr=min_lr, warmup_epochs=warm, epochs=epochs)
    grid = np.concatenate([np.linspace(0, warm, 23), np.linspace(warm, epochs, 41)])
    opt = types.SimpleNamespace(param_groups=[{}])
    out[name + "_args"] = np.array([lr, min_lr, warm, epochs], dtype=np.float64)
    out[name + "_epochs"] = grid
    out[name + "_lr"] = np.array([ref.adjust_learning_rate(opt, float(e), args) for e in grid], dtype=np.float64)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "lr_sched.npz"), **out)
print({k: v.shape for k, v in out.items()})
