"""Host side of the ES evaluator: GA parity with the reference's sequences, shard arithmetic, and the N>1 gather path
on world_size=2 gloo (CPU)."""
import os
import subprocess
import sys

import numpy as np

from paddlerobotics_b200.es import SimpleGA, shard_range, solutions_to_etg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_simplega_reproduces_reference_sequences(golden):
    np.random.seed(0)
    ga = SimpleGA(12, sigma_init=0.02, sigma_decay=0.99, sigma_limit=0.005, elite_ratio=0.1, weight_decay=0.005, popsize=40, param=np.zeros(12))  # train.py:288-295
    s1 = ga.ask()
    assert np.array_equal(s1, golden["ga_s1"])
    ga.tell(golden["ga_fit"].copy())
    s2 = ga.ask()
    assert np.array_equal(s2, golden["ga_s2"])
    assert np.array_equal(ga.best_param, golden["ga_best"]) and np.isclose(ga.sigma, float(golden["ga_sigma"]))


def test_shard_ranges_cover_and_are_contiguous():
    for n, g in ((256, 8), (40, 4), (4096, 3)):
        r = [shard_range(n, k, g) for k in range(g)]
        assert r[0][0] == 0 and r[-1][1] == n and all(r[i][1] == r[i + 1][0] for i in range(g - 1))


def test_solutions_to_etg_matches_reference_fit(golden):
    w, b = solutions_to_etg([golden["opt_sol"]], golden["opt_points"], golden["opt_w0"], golden["opt_b0"])
    assert np.allclose(w[0], golden["opt_w1"], atol=1e-14) and np.allclose(b[0], golden["opt_b1"], atol=1e-15)


_WORKER = r'''
import os, sys
sys.path.insert(0, %r)
import numpy as np, torch, torch.distributed as dist
from paddlerobotics_b200.es import SimpleGA, shard_range, all_gather_concat
dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
rank, world = dist.get_rank(), dist.get_world_size()
np.random.seed(0)
ga = SimpleGA(12, sigma_init=0.02, sigma_decay=0.99, sigma_limit=0.005, elite_ratio=0.1, weight_decay=0.005, popsize=16, param=np.zeros(12))
for gen in range(2):
    sol = ga.ask()                                        # identical on every rank (same seed)
    lo, hi = shard_range(16, rank, world)
    local = torch.tensor([-np.sum(s * s) + 0.01 * i for i, s in zip(range(lo, hi), sol[lo:hi])], dtype=torch.float64)   # stand-in fitness of my shard
    fit = all_gather_concat(local, world, rank).numpy()
    ref = np.array([-np.sum(s * s) + 0.01 * i for i, s in enumerate(sol)])
    assert np.array_equal(fit, ref), (fit, ref)
    # the evaluator's packed form: [fitness | mean length] of the shard in ONE collective
    from paddlerobotics_b200.es import gather_fitness_and_length
    fl = torch.stack([local, torch.arange(lo, hi, dtype=torch.float64) + 100.0])
    f2, l2 = gather_fitness_and_length(fl, world, rank)
    assert np.array_equal(f2.numpy(), ref) and np.array_equal(l2.numpy(), np.arange(16) + 100.0)
    ga.tell(fit)
out = torch.tensor(ga.best_param)
gathered = [torch.zeros_like(out) for _ in range(world)]
dist.all_gather(gathered, out)
assert all(torch.equal(gathered[0], g) for g in gathered)   # identical tell() result on every rank
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_population_gather_world2_gloo(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_WORKER % ROOT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29531", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE, stderr=subprocess.STDOUT) for r in range(2)]
    outs = [p.communicate(timeout=180)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert all("ok" in o for o in outs)


def test_bc_replay_and_batched_noise_on_cpu_tensors():
    """Host logic of bc.py on CPU tensors: ring wrap-around keeps (student, expert) rows paired; obs2noise_batch touches only
    the rpy / drpy / q / qd slices with the sigmas of BCtrain.py:55-58 (already divided by the sensor normalisers)."""
    import torch
    from paddlerobotics_b200 import bc
    m = bc.BCReplayMemory(10, 46, 49, device="cpu")
    for k in range(4):                                   # 4 x 4 rows into a ring of 10
        ref = torch.full((4, 49), float(k)) + torch.arange(4).reshape(4, 1) * 0.1
        m.append(ref[:, 3:].clone(), ref)
    assert m.size() == 10
    o, r = m.sample_batch_by_index(torch.arange(10))
    assert torch.equal(o, r[:, 3:])                      # pairs stay aligned after the wrap
    assert set(float(x) for x in r[:, 0].round().unique()) == {1.0, 2.0, 3.0}   # oldest batch (k=0) overwritten except by wrap rows
    g = torch.Generator().manual_seed(0)
    x = torch.zeros(20000, 49)
    n = bc.obs2noise_batch(x, g)
    assert torch.equal(n[:, :7], x[:, :7]) and torch.equal(n[:, 37:], x[:, 37:])
    for lo, hi, sig in bc.NOISE:
        assert abs(float(n[:, lo:hi].std()) - sig) < 0.03 * sig and abs(float(n[:, lo:hi].mean())) < 0.02 * sig
    a = bc.cal_agent_obs(x, sensor_noise=False)
    assert a.shape == (20000, 46)
