"""A few launches of the tensor-core kernels at batch 8192 (fused MLP forward with dumps + the backward GEMMs) for an `ncu --set full` capture."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from paddlerobotics_b200.agent import MujocoAgent, SACLearner
B = 8192
agent = MujocoAgent(49, 12, seed=0)
L = SACLearner(agent, B)
g = torch.Generator(device="cuda").manual_seed(0)
obs = torch.randn(B, 49, device="cuda", generator=g); act = torch.rand(B, 12, device="cuda", generator=g) * 2 - 1
rew = torch.randn(B, device="cuda", generator=g); nobs = torch.randn(B, 49, device="cuda", generator=g); term = torch.ones(B, device="cuda")
for _ in range(2):
    L.learn(obs, act, rew, nobs, term, graph=False, pull=False)
torch.cuda.synchronize()
