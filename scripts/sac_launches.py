"""One SAC learn (batch from argv, default 8192) repeated 4x without graph, for an ncu launch list."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from paddlerobotics_b200.agent import MujocoAgent, SACLearner
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
agent = MujocoAgent(49, 12, seed=0)
learner = SACLearner(agent, B)
g = torch.Generator(device="cuda").manual_seed(0)
obs = torch.randn(B, 49, device="cuda", generator=g); act = torch.rand(B, 12, device="cuda", generator=g) * 2 - 1
rew = torch.randn(B, device="cuda", generator=g); nobs = torch.randn(B, 49, device="cuda", generator=g); term = torch.ones(B, device="cuda")
for _ in range(4):
    learner.learn(obs, act, rew, nobs, term, graph=False, pull=False)
torch.cuda.synchronize()
# production path: the batch is gathered straight into the learner's static graph inputs, noise comes from the counter RNG in the kernels
for x, sx in zip((obs, act, rew, nobs, term), learner.static_batch()):
    sx.copy_(x)
obs, act, rew, nobs, term = learner.static_batch()
for _ in range(3):
    learner.learn(obs, act, rew, nobs, term, graph=True, pull=False)      # capture + warm-up outside the timed region
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    learner.learn(obs, act, rew, nobs, term, graph=True, pull=False)
e1.record(); torch.cuda.synchronize()
print("batch", B, "graph learn ms", e0.elapsed_time(e1) / 20)
