"""Diagnostic: host-buffer step through (a) b2q_step_host (pinned memcpy in/out) vs (b) the step kernel reading the
actions from / writing obs|reward|done to pinned host memory directly (zero-copy over PCIe)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from paddlerobotics_b200.env import VecQuadrupedalEnv
from paddlerobotics_b200.etg import ETG_layer, Opt_with_points

layer = ETG_layer(0.5, 0.026, 20, 0.04, np.array([-np.pi / 2, 0]), 0.2, 0.5)
w, b, _ = Opt_with_points(ETG=layer, ETG_T=0.5, Footheight=0.1, Steplength=0.05)
n = 4096
env = VecQuadrupedalEnv(n, auto_reset=True)
env.reset(w, b)
a_np = (np.random.rand(n, 12) * 0.6 - 0.3).astype(np.float32)
for _ in range(20): env.step_host(a_np)
t0 = time.perf_counter()
for _ in range(300): env.step_host(a_np)
t1 = time.perf_counter()
print(json.dumps({"mode": "step_host (memcpy)", "us_per_step": (t1 - t0) / 300 * 1e6}))

h_act = torch.empty(n, 12, pin_memory=True); h_obs = torch.empty(n, 49, pin_memory=True); h_rew = torch.empty(n, pin_memory=True)
h_done = torch.empty(n, dtype=torch.uint8, pin_memory=True)
h_act.copy_(torch.from_numpy(a_np))
s = torch.cuda.current_stream().cuda_stream
def zc():
    h_act.numpy()[:] = a_np
    rc = env.lib.b2q_step(env.h, h_act.data_ptr(), 0, h_obs.data_ptr(), h_rew.data_ptr(), h_done.data_ptr(), env.info.data_ptr(), s)
    assert rc == 0
    torch.cuda.current_stream().synchronize()
for _ in range(20): zc()
t0 = time.perf_counter()
for _ in range(300): zc()
t1 = time.perf_counter()
print(json.dumps({"mode": "zero-copy (kernel reads/writes pinned host memory)", "us_per_step": (t1 - t0) / 300 * 1e6}))
# mixed: actions zero-copy, outputs staged on device then one D2H
d_out = torch.empty(n * 50 * 4 + n, dtype=torch.uint8, device="cuda"); h_out = torch.empty(n * 50 * 4 + n, dtype=torch.uint8, pin_memory=True)
def mixed():
    h_act.numpy()[:] = a_np
    p = d_out.data_ptr()
    rc = env.lib.b2q_step(env.h, h_act.data_ptr(), 0, p, p + n * 49 * 4, p + n * 50 * 4, env.info.data_ptr(), s)
    assert rc == 0
    h_out.copy_(d_out, non_blocking=True)
    torch.cuda.current_stream().synchronize()
for _ in range(20): mixed()
t0 = time.perf_counter()
for _ in range(300): mixed()
t1 = time.perf_counter()
print(json.dumps({"mode": "actions zero-copy, outputs one D2H", "us_per_step": (t1 - t0) / 300 * 1e6}))
