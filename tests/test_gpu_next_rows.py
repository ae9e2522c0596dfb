"""SURVEY §8f "next" rows: on-device ETG fit (f-1) and device replay memory feeding the SAC kernels (f-2)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_etg_fit_on_device_matches_reference_host_fit(golden):
    """b2q_etg_fit == Opt_with_points(points=prior+solution, w0, b0) of train.py:81-110,405-407 (float64)."""
    from paddlerobotics_b200.es import SimpleGA, solutions_to_etg, solutions_to_etg_device
    np.random.seed(0)
    ga = SimpleGA(12, sigma_init=0.02, sigma_decay=0.99, sigma_limit=0.005, elite_ratio=0.1, weight_decay=0.005, popsize=64, param=np.zeros(12))
    sol = ga.ask()
    w_h, b_h = solutions_to_etg(sol, golden["opt_points"], golden["opt_w0"], golden["opt_b0"])
    w_d, b_d = solutions_to_etg_device(sol, golden["opt_points"], golden["opt_w0"], golden["opt_b0"])
    assert np.abs(w_d.cpu().numpy() - w_h).max() < 1e-9 and np.abs(b_d.cpu().numpy() - b_h).max() < 1e-12
    # and against the vector made by the reference's own function
    w1, b1 = solutions_to_etg_device(golden["opt_sol"][None, :], golden["opt_points"], golden["opt_w0"], golden["opt_b0"])
    assert np.abs(w1[0].cpu().numpy() - golden["opt_w1"]).max() < 1e-9 and np.abs(b1[0].cpu().numpy() - golden["opt_b1"]).max() < 1e-12


def test_device_replay_memory_append_wrap_and_sample():
    import torch
    from paddlerobotics_b200.replay import ReplayMemory
    rpm = ReplayMemory(1000, 49, 12)
    ref = {k: [] for k in "o a r n t".split()}
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    for step in range(5):                       # 5 x 256 = 1280 > capacity: wraps
        o, n = torch.randn(256, 49, device="cuda", generator=g), torch.randn(256, 49, device="cuda", generator=g)
        a, r, t = torch.rand(256, 12, device="cuda", generator=g), torch.randn(256, device="cuda", generator=g), (torch.rand(256, device="cuda", generator=g) > 0.5).float()
        rpm.append(o, a, r, n, t)
        for k, v in zip("o a r n t".split(), (o, a, r, n, t)):
            ref[k].append(v)
    assert rpm.size() == 1000
    allo = torch.cat(ref["o"]); alln = torch.cat(ref["n"]); allr = torch.cat(ref["r"])
    # ring content: slot s holds transition index i with i % 1000 == s, the latest such i
    for s in (0, 279, 280, 999):
        i = s + 1000 if s + 1000 < 1280 else s
        assert torch.equal(rpm.obs[s], allo[i]) and torch.equal(rpm.next_obs[s], alln[i]) and rpm.reward[s] == allr[i]
    o, a, r, n, t = rpm.sample_batch(4096, seed=3)
    # every sampled row is a stored row, consistent across the five arrays
    idx = (o[:, None, 0] == rpm.obs[None, :, 0]).float().argmax(1)
    assert torch.equal(o, rpm.obs[idx]) and torch.equal(n, rpm.next_obs[idx]) and torch.equal(a, rpm.action[idx]) and torch.equal(r, rpm.reward[idx]) and torch.equal(t, rpm.terminal[idx])
    # roughly uniform over the ring
    hist = torch.bincount(idx, minlength=1000).float()
    assert hist.max() < 20 and (hist > 0).float().mean() > 0.95
    o2 = rpm.sample_batch(4096, seed=3)[0]
    assert torch.equal(o, o2)


def test_on_device_training_loop_smoke(etg_default):
    """rollout (sampled policy) -> device replay -> SAC learn from a CUDA graph, nothing crosses PCIe inside the loop
    (the reference's run_train_episode, train.py:129-178, batched)."""
    import torch
    from paddlerobotics_b200.agent import MujocoAgent, SACLearner
    from paddlerobotics_b200.env import VecQuadrupedalEnv
    from paddlerobotics_b200.replay import ReplayMemory
    w, b = etg_default
    n, B = 512, 256
    env = VecQuadrupedalEnv(n, auto_reset=True)
    ag = MujocoAgent(49, 12, seed=0)
    L = SACLearner(ag, B)
    rpm = ReplayMemory(20000, 49, 12)
    obs = env.reset(w, b).clone()
    losses = []
    for k in range(30):
        act, _ = ag.sample_batch(obs)
        nobs, rew, done, info = env.step(act * 0.3)
        rpm.append(obs, act, rew, nobs, 1.0 - done.float())
        obs = nobs.clone()
        if rpm.size() >= 2048:
            losses.append(L.learn(*rpm.sample_batch(B), graph=True).clone())
    assert len(losses) > 20 and all(torch.isfinite(l).all() for l in losses)
    env.close()


def test_dynamics_identification_evaluator_vs_oracle(golden):
    """f-4: 48-parameter dynamics individuals replay a recorded gait table (ETG off) and are scored with the reference's
    loss_func; GPU rewards == oracle rollouts + numpy restatement; the loss itself is pinned to the reference function."""
    import os
    import torch
    from oracle import oracle as O
    from paddlerobotics_b200.es import DynamicsEvaluator
    from paddlerobotics_b200.etg import dynamic_dict_to_row, param2dynamic_dict

    def loss_np(drpy, motor, md, key):          # numpy restatement of Dynamic_parallel_model.py:29-41
        lm = np.max(np.mean((motor - md[key + "_motor_mean"]) ** 2 / md[key + "_motor_std"] ** 2, axis=0))
        ld = np.max(np.mean((drpy - md[key + "_drpy_mean"]) ** 2 / md[key + "_drpy_std"] ** 2, axis=0))
        return (ld + lm) / 2.0
    md_g = {k[len("dynloss_"):]: golden[k] for k in golden.files if k.startswith("dynloss_exp")}
    assert np.isclose(loss_np(golden["dynloss_drpy"], golden["dynloss_motor"], md_g, "exp"), float(golden["dynloss_value"]), rtol=1e-12)

    T = 30
    gait_tab = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "gait_action_list_CPG_stairstair7_12_3.npy")) + np.array([0, 0.9, -1.8] * 4)
    gait = {"exp": gait_tab[:T] , "ori": gait_tab[100:100 + T]}
    rng = np.random.default_rng(0)
    md = {}
    for k in ("exp", "ori"):
        md[k + "_motor_mean"] = gait[k] + rng.normal(0, 0.02, (T, 12)); md[k + "_motor_std"] = rng.uniform(0.05, 0.1, (T, 12))
        md[k + "_drpy_mean"] = rng.normal(0, 0.2, (T, 3)); md[k + "_drpy_std"] = rng.uniform(0.3, 0.6, (T, 3))
    sols = rng.uniform(-0.25, 0.25, (4, 48))
    ev = DynamicsEvaluator(4, gait, md, steps=T, precision="f64", ring_depth=5)
    rew = ev.evaluate(sols).cpu().numpy()
    ref = np.zeros(4)
    pose = np.array([0, 0.9, -1.8] * 4)
    for i in range(4):
        row = dynamic_dict_to_row(param2dynamic_dict(sols[i]))
        for k in ("exp", "ori"):
            o = O.OracleEnv(O.default_config(etg_enabled=0), row); o.reset()
            motor, drpy = [], []
            for t in range(T):
                _, _, _, info = o.step(gait[k][t] - pose)
                motor.append(info[42:54]); drpy.append(info[39:42])
            ref[i] += (30 - loss_np(np.array(drpy), np.array(motor), md, k)) / 2.0
    assert np.abs(rew - ref).max() < 1e-6, (rew, ref)
    ev.env.close()


def test_train_loop_runs_and_reward_improves():
    """The batched ETG-RL loop (paddlerobotics_b200/train.py ~ ETGRL/train.py:252-449): SAC + one ES phase; the mean step
    reward under the learned residual policy must beat the random-action warm-up phase."""
    from paddlerobotics_b200 import train
    log = train.main(["--num_envs", "1024", "--max_steps", "1300000", "--batch", "1024", "--warmup_steps", "20480", "--log_every", "100",
                      "--es_every_steps", "700000", "--es_train_steps", "1", "--popsize", "10", "--es_rollouts", "2", "--e_step", "200"])
    assert len(log) >= 10
    early = np.mean([r["mean_step_reward"] for r in log[:2]])
    late = np.mean([r["mean_step_reward"] for r in log[-3:]])
    assert np.isfinite(late) and late > early + 0.5, (early, late)


def test_bc_loop_clones_expert():
    """f-3 end to end (paddlerobotics_b200/bc.py ~ ETGRL/BCtrain.py:86-146): a 46-dim noisy-observation student acting in the
    batched env is cloned from a 49-dim expert; the BC actor loss (negative log-likelihood of the expert action) must fall."""
    import torch
    from paddlerobotics_b200 import bc
    from paddlerobotics_b200.agent import MujocoAgent
    from paddlerobotics_b200.env import VecQuadrupedalEnv
    from paddlerobotics_b200.etg import ETG_layer, Opt_with_points
    layer = ETG_layer(0.5, 0.026, 20, 0.04, np.array([-np.pi / 2, 0]), 0.2, 0.5)
    w, b, _ = Opt_with_points(ETG=layer, ETG_T=0.5, Footheight=0.03, Steplength=0.02)
    env = VecQuadrupedalEnv(512, auto_reset=True, max_episode_steps=100)
    expert, student = MujocoAgent(49, 12, seed=1), MujocoAgent(46, 12, seed=2)
    g = torch.Generator(device="cuda").manual_seed(0)
    o = torch.randn(64, 49, device="cuda", generator=g)
    n1 = bc.obs2noise_batch(o, g)
    assert torch.equal(n1[:, :7], o[:, :7]) and torch.equal(n1[:, 37:], o[:, 37:]) and not torch.equal(n1[:, 7:37], o[:, 7:37])
    losses = bc.run_bc(env, student, expert, w, b, iters=60, batch=512, train_every=1)
    a = np.array([l[1] for l in losses])
    assert len(a) >= 50 and np.isfinite(a).all() and a[-10:].mean() < a[:10].mean() - 0.05, (a[:10].mean(), a[-10:].mean())
    env.close()


def test_train_loop_checkpoints_on_the_reference_cadence_and_restores(tmp_path):
    """agent.save + np.savez(w, b, param) (ETGRL/train.py:386-390) and the --load path (:332-333, 439-441): the loop writes itr_<steps>.pt
    with the reference's state-dict keys and itr_<steps>.npz{w,b,param}; a second run restores both."""
    import torch
    from paddlerobotics_b200 import train
    args = ["--num_envs", "256", "--batch", "256", "--warmup_steps", "2048", "--log_every", "20", "--ES", "0", "--task_mode", "ground",
            "--outdir", str(tmp_path), "--suffix", "t", "--eval_every_steps", "10240"]
    train.main(args + ["--max_steps", "25600"])
    files = sorted(os.listdir(tmp_path / "t"))
    pts = [f for f in files if f.endswith(".pt")]
    assert len(pts) >= 2 and all(f[:-3] + ".npz" in files for f in pts), files
    sd = torch.load(tmp_path / "t" / pts[-1])
    assert {"actor_model.l1.weight", "actor_model.mean_linear.bias", "critic_model.l6.weight"} <= set(sd) and sd["actor_model.l1.weight"].shape == (256, 49)
    z = np.load(tmp_path / "t" / (pts[-1][:-3] + ".npz"))
    assert z["w"].shape == (3, 20) and z["b"].shape == (3,) and z["param"].reshape(-1).shape == (12,)
    log = train.main(args + ["--max_steps", "5120", "--load", str(tmp_path / "t" / pts[-1])])
    assert len(log) >= 1 and np.isfinite(log[-1]["mean_step_reward"])


def test_replay_memory_device_cursor_matches_host_cursor_and_replays_from_a_graph():
    """b2q_rpm_*_cursor: ring position / fill level / sample counter in device memory.  Same contents as the host-cursor ring, across the wrap;
    captured ONCE in a CUDA graph, every replay appends at the next position and draws a new sample."""
    import torch
    from paddlerobotics_b200.replay import ReplayMemory
    dev = torch.device("cuda")
    a, b = ReplayMemory(1000, 49, 12), ReplayMemory(1000, 49, 12, device_cursor=True)
    g = torch.Generator(device="cuda"); g.manual_seed(3)
    obs, act, rew, nobs, term = (torch.empty(300, 49, device=dev), torch.empty(300, 12, device=dev), torch.empty(300, device=dev),
                                 torch.empty(300, 49, device=dev), torch.empty(300, device=dev))
    def fill(k):
        obs.fill_(float(k)); nobs.fill_(float(k) + 0.5); act.fill_(float(-k)); rew.copy_(torch.arange(300, device=dev) + 1000.0 * k); term.fill_(float(k % 2))
    fill(0); a.append(obs, act, rew, nobs, term); b.append(obs, act, rew, nobs, term)      # eager call of the cursor path
    torch.cuda.synchronize()
    out = tuple(torch.zeros_like(x[:64]) for x in (obs, act, rew, nobs, term))
    gr = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    mirrors = (b._curr_pos, b._curr_size, b._samples)
    with torch.cuda.graph(gr, stream=side):
        b.append(obs, act, rew, nobs, term)
        b.sample_batch(64, out=out)
    b._curr_pos, b._curr_size, b._samples = mirrors
    torch.cuda.current_stream().wait_stream(side)
    samples = []
    for k in range(1, 5):                                                                  # 4 more blocks of 300: wraps the 1000-row ring
        fill(k); a.append(obs, act, rew, nobs, term)
        gr.replay(); b.advance(300)
        torch.cuda.synchronize()
        samples.append(out[2].clone())
        assert bool(((out[0][:, 0] + 0.5) == out[3][:, 0]).all())                          # rows stay consistent (obs / next_obs of the same transition)
    for x, y in ((a.obs, b.obs), (a.action, b.action), (a.reward, b.reward), (a.next_obs, b.next_obs), (a.terminal, b.terminal)):
        assert torch.equal(x, y)
    assert (a._curr_pos, a._curr_size) == (b._curr_pos, b._curr_size) == (500, 1000)
    assert b.cursor.tolist() == [500, 1000, 4]
    assert not torch.equal(samples[0], samples[1]) and not torch.equal(samples[2], samples[3])   # a new draw per replay
