"""Step-kernel time vs batch size (device events, L2 not flushed — the state is tiny).  Diagnostic only."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from paddlerobotics_b200.env import VecQuadrupedalEnv
from paddlerobotics_b200.etg import ETG_layer, Opt_with_points

layer = ETG_layer(0.5, 0.026, 20, 0.04, np.array([-np.pi / 2, 0]), 0.2, 0.5)
stable = os.environ.get("GAIT") == "stable"   # coherent gait phases, no falls: every warp's swing feet coincide
w, b, _ = Opt_with_points(ETG=layer, ETG_T=0.5, Footheight=0.03 if stable else 0.1, Steplength=0.02 if stable else 0.05)
for n in [int(x) for x in (sys.argv[1:] or "296 592 1184 2048 2368 4096 4736 8192 16384 65536".split())]:
    env = VecQuadrupedalEnv(n, auto_reset=True)
    env.reset(w, b)
    a = torch.zeros(n, 12, device="cuda") if stable else torch.rand(n, 12, device="cuda") * 0.6 - 0.3
    for _ in range(20): env.step(a)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200): env.step(a)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 200
    print(json.dumps({"gait": "stable" if stable else "default+random residual", "done_frac": float(env.done.float().mean()), "envs": n, "us_per_step": ms * 1e3, "env_steps_per_s": n / ms * 1e3}), flush=True)
    env.close()
