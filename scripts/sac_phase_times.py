"""Per-phase time of the SAC learner: each b2q_sac_phase captured alone in a CUDA graph and replayed (batch from argv, default 8192)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from paddlerobotics_b200.agent import MujocoAgent, SACLearner
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
agent = MujocoAgent(49, 12, seed=0)
L = SACLearner(agent, B)
g = torch.Generator(device="cuda").manual_seed(0)
obs = torch.randn(B, 49, device="cuda", generator=g); act = torch.rand(B, 12, device="cuda", generator=g) * 2 - 1
rew = torch.randn(B, device="cuda", generator=g); nobs = torch.randn(B, 49, device="cuda", generator=g); term = torch.ones(B, device="cuda")
e1 = torch.randn(B, 12, device="cuda", generator=g); e2 = torch.randn(B, 12, device="cuda", generator=g)
for _ in range(3):
    L.learn(obs, act, rew, nobs, term, eps_next=e1, eps_cur=e2, graph=False)
torch.cuda.synchronize()
args = (obs.data_ptr(), act.data_ptr(), rew.data_ptr(), nobs.data_ptr(), term.data_ptr(), e1.data_ptr(), e2.data_ptr(), 1)
stream = torch.cuda.Stream()
tot = 0.0
import ctypes as C
noeps = args[:5] + (None, None, 1)
def run(phases):
    if phases == "learn":        # the fused entry point (critics' Adam beside the actor forward), explicit noise
        assert L.lib.b2q_sac_learn(L.h, *args[:7], C.c_uint64(1), L.losses.data_ptr(), L._stream()) == 0
    elif phases == "learn, counter-RNG noise":
        assert L.lib.b2q_sac_learn(L.h, *noeps[:7], C.c_uint64(1), L.losses.data_ptr(), L._stream()) == 0
    else:
        for ph in phases:
            assert L.lib.b2q_sac_phase(L.h, ph, *args, L._stream()) == 0
for phases in ((0,), (1,), (2,), (3,), (0, 1, 2, 3), "learn", "learn, counter-RNG noise"):
    with torch.cuda.stream(stream):
        run(phases)
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=stream):
            run(phases)
        for _ in range(5):
            gr.replay()
        torch.cuda.synchronize()
        e0, e1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            gr.replay()
        e1_.record(); torch.cuda.synchronize()
    print("batch", B, "phases", phases, "us per replay %.1f" % (e0.elapsed_time(e1_) / 50 * 1000))
