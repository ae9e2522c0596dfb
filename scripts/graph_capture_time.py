"""Diagnostic: wall time of the first (capturing) and later graph-replayed SAC learn calls."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from paddlerobotics_b200.agent import MujocoAgent, SACLearner
B = 4096
agent = MujocoAgent(49, 12, seed=0)
learner = SACLearner(agent, B)
g = torch.Generator(device="cuda").manual_seed(0)
obs = torch.randn(B, 49, device="cuda", generator=g); act = torch.rand(B, 12, device="cuda", generator=g) * 2 - 1
rew = torch.randn(B, device="cuda", generator=g); nobs = torch.randn(B, 49, device="cuda", generator=g); term = torch.ones(B, device="cuda")
for k in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    learner.learn(obs, act, rew, nobs, term, graph=True, pull=False)
    torch.cuda.synchronize(); print("graph learn call", k, "%.3f ms" % ((time.perf_counter() - t0) * 1e3), flush=True)
for k in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    learner.learn(obs, act, rew, nobs, term, graph=False, pull=False)
    torch.cuda.synchronize(); print("eager learn call", k, "%.3f ms" % ((time.perf_counter() - t0) * 1e3), flush=True)
