"""Latency of the reference-style N=1 env object (make_env(...).step(action) -> numpy obs, float reward, bool done, info dict)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from paddlerobotics_b200.env import make_env
env = make_env("Quadrupedal", task="ground", render=False, ETG=1, ETG_T=0.5, reward_p=5, vel_d=0.5)
from paddlerobotics_b200.etg import shipped_gait
w, b = shipped_gait()                      # the reference's walking gait: the robot stays up (a fallen robot lying on its knees takes the general 36-row solve)
obs, info = env.reset(ETG_w=w, ETG_b=b)
a = np.zeros(12)
for _ in range(50): env.step(a)
t0 = time.perf_counter()
for _ in range(1000):
    obs, r, d, info = env.step(a)
    if d: env.reset(ETG_w=w, ETG_b=b)
dt = (time.perf_counter() - t0) / 1000
print(json.dumps({"what": "make_env N=1 env.step (numpy in, numpy obs + info dict out)", "us_per_step": dt * 1e6, "steps_per_s": 1 / dt}))
# the previous implementation of QuadrupedalEnv.step, for comparison: device tensors + four separate device->host reads
from paddlerobotics_b200.env import info_dict
def old_step(action):
    o, r, d, inf = env.vec.step(np.asarray(action, dtype=np.float64).reshape(1, 12), False)
    return o[0].double().cpu().numpy(), float(r[0]), bool(d[0]), info_dict(inf[0].double().cpu().numpy())
for _ in range(50): old_step(a)
t0 = time.perf_counter()
for _ in range(1000): old_step(a)
dt = (time.perf_counter() - t0) / 1000
print(json.dumps({"what": "same through device tensors + 4 D2H reads (previous wrapper)", "us_per_step": dt * 1e6, "steps_per_s": 1 / dt}))
from paddlerobotics_b200.agent import MujocoAgent
agent = MujocoAgent(49, 12, seed=0)
for _ in range(50): agent.sample(obs)
t0 = time.perf_counter()
for _ in range(1000): act = agent.sample(obs)
dt = (time.perf_counter() - t0) / 1000
print(json.dumps({"what": "agent.sample(obs) numpy [49] -> numpy [12] (fused tcgen05 MLP on pinned buffers)", "us_per_call": dt * 1e6}))
