"""Where the host-buffer step (b2q_step_host through VecQuadrupedalEnv.step_host) spends its wall time at 4096 envs."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from paddlerobotics_b200.env import VecQuadrupedalEnv
from bench import etg_weights
n = 4096
env = VecQuadrupedalEnv(num_envs=n, auto_reset=True)
w, b = etg_weights()
env.reset(w, b)
acts = np.random.default_rng(0).uniform(-0.3, 0.3, (16, n, 12)).astype(np.float32)
for k in range(20):
    env.step_host(acts[k % 16], info=True)
K = 300
def timed(f):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for k in range(K):
        f(k)
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / K * 1e6
full = timed(lambda k: env.step_host(acts[k % 16], info=True))
noinfo = timed(lambda k: env.step_host(acts[k % 16], info=False))
copy = timed(lambda k: np.copyto(env._np_act, acts[k % 16]))
lib, h = env.lib, env.h
ptrs = (env._h_act.data_ptr(), 0, env._h_obs.data_ptr(), env._h_rew.data_ptr(), env._h_done.data_ptr(), env._h_info.data_ptr(), env._stream())
ccall = timed(lambda k: lib.b2q_step_host(h, *ptrs))
a_dev = torch.as_tensor(acts[0], device="cuda")
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); ev0.record()
for k in range(K):
    env.step(a_dev)
ev1.record(); torch.cuda.synchronize()
print({"step_host_info_us": full, "step_host_noinfo_us": noinfo, "np_copyto_us": copy, "c_call_only_us": ccall, "device_step_back_to_back_us": ev0.elapsed_time(ev1) / K * 1e3})
