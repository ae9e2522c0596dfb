"""What does the float64 validation build of the step kernel cost on this GPU? (the f32 product kernel vs full f64)"""
import json, sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import etg_weights
from paddlerobotics_b200.env import VecQuadrupedalEnv
w, b = etg_weights()
for prec in ("f32", "f64"):
    for n in (512, 4096):
        env = VecQuadrupedalEnv(n, precision=prec, auto_reset=True); env.reset(w, b)
        a = (torch.rand(n, 12, device="cuda", dtype=env.dtype) * 0.6 - 0.3)
        for _ in range(20): env.step(a)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(100): env.step(a)
        e1.record(); torch.cuda.synchronize()
        print(json.dumps({"precision": prec, "envs": n, "us_per_step": e0.elapsed_time(e1) * 10, "env_steps_per_s": n / (e0.elapsed_time(e1) / 100 * 1e-3)}), flush=True)
        env.close()
