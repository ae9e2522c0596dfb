"""Multi-GPU path: env shards per rank, no data-path collective; ES fitness all-gather over NCCL.
Run under torchrun with >= 2 GPUs; skipped otherwise."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_WORKER = r'''
import os, sys
sys.path.insert(0, %r)
import numpy as np, torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from paddlerobotics_b200.es import PopulationEvaluator, SimpleGA, solutions_to_etg
g = np.load(os.path.join(%r, "tests", "golden", "reference_vectors.npz"))
np.random.seed(0)
ga = SimpleGA(12, sigma_init=0.02, sigma_decay=0.99, sigma_limit=0.005, elite_ratio=0.25, weight_decay=0.005, popsize=8, param=np.zeros(12))
sol = ga.ask()
w, b = solutions_to_etg(sol, g["opt_points"], g["opt_w0"], g["opt_b0"])
ev = PopulationEvaluator(8, 4, max_steps=40, rank=rank, world=world, device=local)
fit, mlen = ev.evaluate(w, b)
assert fit.shape == (8,)
# every rank must hold the identical full fitness vector, equal to a single-GPU evaluation of the whole population
gathered = [torch.zeros_like(fit) for _ in range(world)]
dist.all_gather(gathered, fit)
assert all(torch.equal(gathered[0], x) for x in gathered)
if rank == 0:
    ev1 = PopulationEvaluator(8, 4, max_steps=40, rank=0, world=1, device=local)
    fit1, _ = ev1.evaluate(w, b)
    assert torch.equal(fit1, fit), (fit1, fit)
ga.tell(fit.cpu().numpy())
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_es_population_sharded_over_two_gpus_nccl(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    script = tmp_path / "w.py"
    script.write_text(_WORKER % (ROOT, ROOT))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29541", str(script)], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert out.stdout.count("ok") == 2


_SAC_WORKER = r'''
import os, sys
sys.path.insert(0, %r)
import numpy as np, torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from paddlerobotics_b200.agent import MujocoAgent, SACLearner, flatten_params
B = 512                      # global batch; each rank learns on its B/world shard
torch.manual_seed(0)         # same data on every rank, then sliced
dev = torch.device("cuda", local)
obs, nobs = torch.randn(B, 49, device=dev), torch.randn(B, 49, device=dev)
act = torch.rand(B, 12, device=dev) * 2 - 1
rew, term = torch.randn(B, device=dev), (torch.rand(B, device=dev) > 0.1).float()
e1, e2 = torch.randn(B, 12, device=dev), torch.randn(B, 12, device=dev)
ag = MujocoAgent(49, 12, device=local, seed=3)                 # identical init on every rank
L = SACLearner(ag, B // world, world=world)
sl = slice(rank * B // world, (rank + 1) * B // world)
L.learn(obs[sl], act[sl], rew[sl], nobs[sl], term[sl], eps_next=e1[sl], eps_cur=e2[sl])
a, c = flatten_params(ag.params)
# every rank ends with identical parameters
for t in (a, c):
    g = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(g, t)
    assert all(torch.equal(g[0], x) for x in g)
# and they match a single-GPU learner on the whole batch (mean of shard gradients == full-batch gradient)
if rank == 0:
    ag1 = MujocoAgent(49, 12, device=local, seed=3)
    L1 = SACLearner(ag1, B, world=1)
    L1.learn(obs, act, rew, nobs, term, eps_next=e1, eps_cur=e2)
    a1, c1 = flatten_params(ag1.params)
    a0, c0 = flatten_params(MujocoAgent(49, 12, device=local, seed=3).params)
    for d, r, z in ((a, a1, a0), (c, c1, c0)):
        dd, rr = d - z, r - z
        cos = float(torch.dot(dd, rr) / (dd.norm() * rr.norm()))
        assert cos > 0.98, cos
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_sac_data_parallel_grad_allreduce_two_gpus(tmp_path):
    """SURVEY §8e collective 2: per-GPU gradient buckets all-reduced (NCCL) between the gradient and Adam phases."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    script = tmp_path / "w2.py"
    script.write_text(_SAC_WORKER % ROOT)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29543", str(script)], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert out.stdout.count("ok") == 2
