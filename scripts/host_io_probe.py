"""Diagnostic: e2e host-buffer step time for B2Q_HOST_IO = 0 / 1 / 2 (set in the environment before the handle is created)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from paddlerobotics_b200.env import VecQuadrupedalEnv
from paddlerobotics_b200.etg import ETG_layer, Opt_with_points

layer = ETG_layer(0.5, 0.026, 20, 0.04, np.array([-np.pi / 2, 0]), 0.2, 0.5)
w, b, _ = Opt_with_points(ETG=layer, ETG_T=0.5, Footheight=0.1, Steplength=0.05)
n = 4096
a_np = (np.random.default_rng(0).random((n, 12)) * 0.6 - 0.3).astype(np.float32)
ref = None
for mode in (0, 1, 2):
    os.environ["B2Q_HOST_IO"] = str(mode)
    env = VecQuadrupedalEnv(n, auto_reset=True)
    env.reset(w, b)
    outs = []
    for i in range(30):
        o, r, d = env.step_host(a_np)
        if i in (0, 29): outs.append((o.copy(), r.copy(), d.copy()))
    if ref is None: ref = outs
    same = all(np.array_equal(x, y) for A, Bb in zip(ref, outs) for x, y in zip(A, Bb))
    t0 = time.perf_counter()
    for _ in range(300): env.step_host(a_np)
    t1 = time.perf_counter()
    print(json.dumps({"B2Q_HOST_IO": mode, "us_per_step": (t1 - t0) / 300 * 1e6, "env_steps_per_s": n * 300 / (t1 - t0), "identical_to_mode0": bool(same)}), flush=True)
    env.close()
