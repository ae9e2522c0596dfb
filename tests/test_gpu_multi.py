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
