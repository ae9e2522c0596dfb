"""Small invocation of every kernel in libb2q.so for compute-sanitizer (memcheck / racecheck / initcheck):
    compute-sanitizer --tool memcheck  python scripts/sanitize.py
    compute-sanitizer --tool racecheck python scripts/sanitize.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from paddlerobotics_b200.env import VecQuadrupedalEnv
from paddlerobotics_b200.etg import ETG_layer, Opt_with_points
from paddlerobotics_b200.agent import MujocoAgent, SACLearner
from paddlerobotics_b200.replay import ReplayMemory
from paddlerobotics_b200.es import PopulationEvaluator

layer = ETG_layer(0.5, 0.026, 20, 0.04, np.array([-np.pi / 2, 0]), 0.2, 0.5)
w, b, _ = Opt_with_points(ETG=layer, ETG_T=0.5, Footheight=0.1, Steplength=0.05)
rng = np.random.default_rng(0)
hf = (rng.random((32, 32)) * 0.03).astype(np.float64)
for kw in (dict(num_envs=13), dict(num_envs=16, precision="f64"), dict(num_envs=24, auto_reset=True, action_filter=1, max_episode_steps=3),
           dict(num_envs=8, heightfield=(hf, -1.0, -1.0, 0.1)), dict(num_envs=40, threads_per_block=128),
           # round-2 paths: reduced sensor layout + noise + stuck / body-collision epilogue; the FEAT variant (joint-limit rows in shared
           # memory, TORQUE mode, base push, damping) in f32 and f64
           dict(num_envs=13, sensor_motor=2, sensor_imu=2, obs_normal=0, noise_stdev=(0.01, 0.05, 0.1, 0.02, 0.04), stuck_termination=1, body_collisions=1, auto_reset=True),
           dict(num_envs=11, joint_limits=1, knee_contacts=1, external_force=1, base_damping=(0.04, 0.02, 0.04, 0.01)), dict(num_envs=9, joint_limits=1, precision="f64"),
           dict(num_envs=8, motor_mode=1), dict(num_envs=8, motor_mode=2)):
    env = VecQuadrupedalEnv(**kw)
    n = env.num_envs
    env.reset(w, b)
    a = (rng.random((n, 12)) * 0.6 - 0.3)
    if kw.get("motor_mode") == 2:                                # HYBRID: (q*, kp, qd*, kd, tau_ff) per motor
        a5 = np.zeros((n, 12, 5)); a5[:, :, 0] = np.array([0.0, 0.9, -1.8] * 4) + a; a5[:, :, 1] = 100.0; a5[:, :, 3] = 1.5; a = a5.reshape(n, 60)
    if kw.get("joint_limits"):
        a[:, 2::3] = 1.2; a[:, 0::3] = 0.9                       # into the stops: the shared-memory 24-row solve runs
    if kw.get("external_force"):
        env.set_external_force(rng.uniform(-10, 10, (n, 3)))
    if kw.get("noise_stdev"):
        env.reset(w, b, x_offset=rng.uniform(-0.1, 0.1, n))
    for _ in range(4):
        env.step(torch.as_tensor(a, device="cuda", dtype=env.dtype))
    env.step_host(a.astype(np.float32 if env.dtype == torch.float32 else np.float64))
    s = env.get_state(); env.set_state(s)
    env.reset(w, b, env_mask=(np.arange(n) % 2 == 0))
    torch.cuda.synchronize(); env.close()

agent = MujocoAgent(49, 12, seed=0)
obs = torch.randn(200, 49, device="cuda")
agent.predict(np.zeros(49, np.float32)); agent.sample(np.zeros(49, np.float32))
learner = SACLearner(agent, 256)
learner.actor.forward(obs, mode=1, seed=3)
rpm = ReplayMemory(4096, 49, 12)
for _ in range(3):
    rpm.append(torch.randn(512, 49, device="cuda"), torch.rand(512, 12, device="cuda") * 2 - 1, torch.randn(512, device="cuda"), torch.randn(512, 49, device="cuda"), torch.ones(512, device="cuda"))
for _ in range(2):
    learner.learn(*rpm.sample_batch(256), graph=False)
flat = SACLearner(agent, 256, sync="flat")
flat.learn(*rpm.sample_batch(256), graph=False)
# explicit-noise path, the graph path on the learner's static inputs (counter-RNG noise), behaviour cloning (given-target critic head, plain Adam)
learner.learn(*rpm.sample_batch(256), eps_next=torch.randn(256, 12, device="cuda"), eps_cur=torch.randn(256, 12, device="cuda"), graph=False)
g2 = SACLearner(MujocoAgent(49, 12, seed=1), 256)
for _ in range(2):
    g2.learn(*rpm.sample_batch(256, out=g2.static_batch()), graph=True, pull=False)
expert = MujocoAgent(49, 12, seed=2)
learner.bc_learn(torch.randn(256, 49, device="cuda"), torch.randn(256, 49, device="cuda"), expert)
ev = PopulationEvaluator(4, 2, max_steps=5)
ev.evaluate(np.repeat(w[None], 4, 0), np.repeat(b[None], 4, 0))
torch.cuda.synchronize()
print("sanitizer script done")
