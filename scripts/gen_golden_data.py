"""Writes tests/golden/density_synth.npz: outputs of the density-map oracle (oracle/data_oracle.py = the reference's numpy /
scipy calls, util/FSC147.py:262-273, 326-331) on seeded dot annotations, as a fixture that travels to the GPU box."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import data_oracle as D

rng = np.random.default_rng(7)
out = {}
H, W = 480, 640
dots = rng.random((57, 2)) * np.array([W, H])
dots[5] = dots[4]                        # coincident annotations count once
dots[6] = [W - 1e-9, H - 1e-9]           # lands on the clamped last row / column
dots[7] = [0.0, 0.0]
out["dots"] = dots
out["hw"] = np.array([H, W])
v = D.val_density(dots, H, W)
out["val_crop"] = v[:48, :48].copy()
out["val_sum"] = np.float64(v.astype(np.float64).sum())
out["val_corner"] = v[-8:, -8:].copy()
new_H, new_W, start = 384, 512, 77
t = D.train_density(dots, H, W, new_H, new_W, start)
out["train_meta"] = np.array([new_H, new_W, start])
out["train_crop"] = t[100:148, 200:248].copy()
out["train_sum"] = np.float64(t.astype(np.float64).sum())
np.savez_compressed(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "density_synth.npz"), **out)
print({k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()})
