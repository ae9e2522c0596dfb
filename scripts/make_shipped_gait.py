"""Fits the reference's shipped gait table (ETGRL/gait_action_list_ETG_exp.npy; a byte copy is kept under tests/golden/) back to ETG
weights and stores them in paddlerobotics_b200/data/etg_shipped_gait.npz (w [3,20], b [3], max |regenerated - table|)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from paddlerobotics_b200 import etg as E  # noqa: E402

src = "/root/reference/QuadrupedalRobots/ETGRL/gait_action_list_ETG_exp.npy"
if not os.path.exists(src):
    src = os.path.join(ROOT, "tests", "golden", "gait_action_list_ETG_exp.npy")
table = np.load(src)
w, b = E.fit_etg_from_table(table, t0=0.026)
regen = E.etg_act_table(w, b, table.shape[0], t0=0.026)
err = float(np.abs(regen - table).max())
print("fit residual (joint offsets, rad):", err, " b =", b)
assert err < 1e-9
np.savez(os.path.join(ROOT, "paddlerobotics_b200", "data", "etg_shipped_gait.npz"), w=w, b=b, residual=err)
