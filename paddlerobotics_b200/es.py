"""ES population evaluator (SURVEY §8 a15, §8e collective 1) and the host-side GA driver.

* `SimpleGA` keeps the reference's ask/tell contract and draws from NumPy's global RNG in the same order as
  ETGRL/alg/es.py:257-314, so that identically seeded runs produce identical populations on every rank (the reference's
  distributed variant relies on the same property: Dynamic_parallel_model.py:152-182).  The GA arithmetic stays on the
  host (12–48 parameters); what moves to the GPU is the rollout of the whole population.
* `PopulationEvaluator` replaces the serial loop `for solution in solutions: ... run_EStrain_episode(...)`
  (train.py:404-413): individual i owns `rollouts` consecutive envs of this rank's shard; all envs step in lock step
  inside the CUDA step kernel; per-env returns are frozen at the first `done`; fitness = mean over rollouts
  (`b2q_es_fitness`); shards are concatenated with ONE all-gather (NCCL on GPUs, gloo in the CPU tests).
"""
import ctypes as C

import numpy as np

from .etg import ETG_layer, Opt_with_points


def compute_weight_decay(weight_decay, model_param_list):
    grid = np.array(model_param_list)
    return -weight_decay * np.mean(grid * grid, axis=1)


class SimpleGA:
    """Elitist GA with Gaussian mutation; same constructor/ask/tell/reset semantics as the reference class."""

    def __init__(self, num_params, sigma_init=0.1, sigma_decay=0.999, sigma_limit=0.01, popsize=256, elite_ratio=0.1,
                 forget_best=False, weight_decay=0.01, param=None):
        self.num_params, self.popsize = num_params, int(popsize)
        self.sigma_init, self.sigma_decay, self.sigma_limit = sigma_init, sigma_decay, sigma_limit
        self.elite_ratio = elite_ratio
        self.elite_popsize = int(self.popsize * self.elite_ratio)
        self.sigma = sigma_init
        self.elite_params = np.zeros((self.elite_popsize, num_params))
        self.elite_rewards = np.zeros(self.elite_popsize)
        self.best_param = np.zeros(num_params) if param is None else param
        self.curr_best_param = self.best_param
        self.best_reward = 0
        self.first_iteration = True
        self.forget_best, self.weight_decay = forget_best, weight_decay

    def reset(self, param):
        self.best_param = np.copy(param)
        self.curr_best_param = np.copy(param)
        self.first_iteration = True

    def rms_stdev(self):
        return self.sigma

    def ask(self):
        # RNG draw order (global NumPy state): one randn block, then per child two parent picks and, after the first
        # generation, one uniform crossover mask.
        self.epsilon = np.random.randn(self.popsize, self.num_params) * self.sigma
        children = np.empty((self.popsize, self.num_params))
        parents = range(self.elite_popsize)
        for i in range(self.popsize):
            ia, ib = np.random.choice(parents), np.random.choice(parents)
            if self.first_iteration:
                base = self.best_param
            else:
                base = np.copy(self.elite_params[ia])
                take_b = np.where(np.random.rand(base.size) > 0.5)
                base[take_b] = self.elite_params[ib][take_b]
            children[i] = base + self.epsilon[i]
        self.solutions = children
        return children

    def tell(self, reward_table_result):
        assert len(reward_table_result) == self.popsize, "Inconsistent reward_table size reported."
        table = np.array(reward_table_result, dtype=np.float64)
        if self.weight_decay > 0:
            table += compute_weight_decay(self.weight_decay, self.solutions)
        if self.forget_best or self.first_iteration:
            reward, solution = table, self.solutions
        else:
            reward, solution = np.concatenate([table, self.elite_rewards]), np.concatenate([self.solutions, self.elite_params])
        idx = np.argsort(reward)[::-1][0:self.elite_popsize]
        self.elite_rewards, self.elite_params = reward[idx], solution[idx]
        self.curr_best_reward = self.elite_rewards[0]
        self.curr_best_param = np.copy(self.elite_params[0])
        if self.first_iteration or (self.curr_best_reward > self.best_reward):
            self.first_iteration = False
            self.best_reward = self.elite_rewards[0]
            self.best_param = np.copy(self.elite_params[0])
        if self.sigma > self.sigma_limit:
            self.sigma *= self.sigma_decay

    def current_param(self):
        return self.curr_best_param

    def set_mu(self, mu):
        pass

    def best_param_(self):
        return self.best_param

    def result(self):
        return (self.best_param, self.best_reward, self.curr_best_reward, self.sigma)


def shard_range(n, rank, world):
    """Contiguous shard [lo, hi) of n units owned by `rank` (GPU g owns [g·N/G,(g+1)·N/G))."""
    return (n * rank) // world, (n * (rank + 1)) // world


def solutions_to_etg(solutions, prior_points, w0, b0, ETG_T=0.5, etg_layer=None):
    """train.py:404-407: control-point deltas -> (w,b) per individual through Opt_with_points (host LS fit)."""
    layer = etg_layer or ETG_layer(ETG_T, 0.026, 20, 0.04, np.array([-np.pi / 2, 0]), 0.2, ETG_T)
    ws, bs = [], []
    for sol in solutions:
        pts = prior_points + np.asarray(sol).reshape(-1, 2)
        w, b, _ = Opt_with_points(ETG=layer, ETG_T=ETG_T, w0=w0, b0=b0, points=pts)
        ws.append(w); bs.append(b)
    return np.array(ws), np.array(bs)


def solutions_to_etg_device(solutions, prior_points, w0, b0, ETG_T=0.5, device=0, lamb=0.5, precision=1e-4):
    """Same as solutions_to_etg but batched on the GPU (b2q_etg_fit): one thread per individual, float64."""
    import torch
    from . import _lib
    lib = _lib.load()
    layer = ETG_layer(ETG_T, 0.026, 20, 0.04, np.array([-np.pi / 2, 0]), 0.2, ETG_T)
    ts = [0.5 * ETG_T + 0.1, 0, 0.05, 0.1, 0.15, 0.2]
    obs = np.array([layer.update(t) for t in ts]).reshape(6, 20)
    dev = torch.device("cuda", int(device))
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64), device=dev)
    sol = t(np.asarray(solutions).reshape(-1, 12))
    pop = sol.shape[0]
    o, pp, w0t, b0t = t(obs), t(np.asarray(prior_points).reshape(6, 2)), t(np.asarray(w0).reshape(3, 20)), t(np.asarray(b0).reshape(3))
    w = torch.empty(pop, 3, 20, dtype=torch.float64, device=dev)
    b = torch.empty(pop, 3, dtype=torch.float64, device=dev)
    rc = lib.b2q_etg_fit(o.data_ptr(), pp.data_ptr(), sol.data_ptr(), w0t.data_ptr(), b0t.data_ptr(), float(lamb), float(precision), w.data_ptr(), b.data_ptr(), pop,
                         C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    assert rc == 0, rc
    return w, b


def all_gather_concat(local, world, rank, group=None):
    """One all-gather of equally sized shards; returns the concatenation on every rank (torch.distributed)."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return local
    out = torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    return out


def gather_fitness_and_length(fl, world, rank, group=None):
    """fl [2, pop_local] = this rank's [fitness | mean episode length] -> (fitness [pop], length [pop]) on every rank with ONE all-gather
    (SURVEY §8e collective 1: the two vectors travel packed)."""
    if world == 1:
        return fl[0], fl[1]
    g = all_gather_concat(fl.reshape(1, 2, fl.shape[1]), world, rank, group)     # [world, 2, pop_local]
    return g[:, 0, :].reshape(-1), g[:, 1, :].reshape(-1)


class PopulationEvaluator:
    """Evaluates this rank's shard of an ES population on its GPU and all-gathers the fitness vector."""

    def __init__(self, popsize, rollouts, max_steps=400, rank=0, world=1, device=0, policy=None, act_bound=0.3, precision="f32", **env_cfg):
        import torch
        from . import _lib
        from .env import VecQuadrupedalEnv
        if popsize % world != 0:
            raise ValueError("popsize must be divisible by the number of ranks (individuals are assigned whole to a GPU)")
        self.popsize, self.rollouts, self.max_steps, self.rank, self.world = popsize, rollouts, max_steps, rank, world
        self.lo, self.hi = shard_range(popsize, rank, world)
        self.pop_local = self.hi - self.lo
        self.n = self.pop_local * rollouts
        self.env = VecQuadrupedalEnv(self.n, device=device, precision=precision, auto_reset=False, **env_cfg)
        self.lib = _lib.load()
        dev, dt = self.env.device, self.env.dtype
        self.alive = torch.ones(self.n, dtype=torch.uint8, device=dev)
        self.ret = torch.zeros(self.n, dtype=dt, device=dev)
        self.len = torch.zeros(self.n, dtype=torch.int32, device=dev)
        self._fl = torch.zeros(2, self.pop_local, dtype=dt, device=dev)      # [fitness | mean length] packed: ONE all-gather per generation
        self.fitness, self.mean_len = self._fl[0], self._fl[1]
        self.policy, self.act_bound = policy, act_bound
        self.zero_act = torch.zeros(self.n, 12, dtype=dt, device=dev)
        self.es_launches = 0

    def evaluate(self, etg_w, etg_b, residual_noise=None):
        """etg_w [pop,3,20], etg_b [pop,3] for the WHOLE population (identical on every rank); returns fitness[pop]
        (identical on every rank) and mean episode length[pop]."""
        import torch
        w = np.repeat(np.asarray(etg_w)[self.lo:self.hi], self.rollouts, axis=0)
        b = np.repeat(np.asarray(etg_b)[self.lo:self.hi], self.rollouts, axis=0)
        env = self.env
        obs = env.reset(w, b)
        self.alive.fill_(1); self.ret.zero_(); self.len.zero_()
        es = env.obs.element_size()
        stream = env._stream()
        for k in range(self.max_steps):
            if self.policy is not None:
                act = self.policy(obs) * self.act_bound            # agent.predict(obs) * action_bound, train.py:226-228
            else:
                act = self.zero_act
            if residual_noise is not None:
                act = act + residual_noise[k]
            obs, rew, done, _ = env.step(act, donef=(k + 1 > self.max_steps))
            rc = self.lib.b2q_es_accumulate(rew.data_ptr(), done.data_ptr(), self.alive.data_ptr(), self.ret.data_ptr(), self.len.data_ptr(),
                                            self.n, es, stream)
            assert rc == 0
            self.es_launches += 1
        rc = self.lib.b2q_es_fitness(self.ret.data_ptr(), self.len.data_ptr(), self.fitness.data_ptr(), self.mean_len.data_ptr(),
                                     self.pop_local, self.rollouts, es, stream)
        assert rc == 0
        self.es_launches += 1
        return gather_fitness_and_length(self._fl, self.world, self.rank)


class DynamicsEvaluator:
    """Sim-to-real dynamics identification on the GPU: the population of 48-vectors is mapped through
    param2dynamic_dict to per-env dynamics rows; every individual replays the recorded gait tables (ETG off, action =
    table - pose_ori) for `steps` control steps and is scored against the recorded real-robot statistics with the
    reference's loss (RemoteESAgent.sample_episode / batch_sample_episodes, Dynamic_parallel_model.py:53-77).  Individuals
    are sharded over ranks exactly like the xparl actors (`solutions[i*K:(i+1)*K]`, :157-159) and the rewards all-gathered."""

    def __init__(self, popsize, gait, mean_dict, keys=("exp", "ori"), steps=100, rank=0, world=1, device=0, precision="f32", ring_depth=4, **env_cfg):
        import torch
        from . import _lib
        from .env import VecQuadrupedalEnv
        if popsize % world != 0:
            raise ValueError("popsize must be divisible by the number of ranks")
        self.popsize, self.keys, self.steps, self.rank, self.world = popsize, tuple(keys), steps, rank, world
        self.lo, self.hi = shard_range(popsize, rank, world)
        self.pop_local = self.hi - self.lo
        self.n = self.pop_local * len(self.keys)          # env index = key * pop_local + individual
        self.env = VecQuadrupedalEnv(self.n, device=device, precision=precision, etg_enabled=0, ring_depth=ring_depth, **env_cfg)
        self.lib = _lib.load()
        dev, dt = self.env.device, self.env.dtype
        pose = np.array([0, 0.9, -1.8] * 4)
        # per-step action table [steps, n, 12] and statistics [steps, keys, 15]
        act = np.stack([np.repeat((np.asarray(gait[k])[:steps] - pose)[:, None, :], self.pop_local, axis=1) for k in self.keys], axis=1).reshape(steps, self.n, 12)
        self.actions = torch.as_tensor(act, dtype=dt, device=dev)
        self.mean = [torch.as_tensor(np.concatenate([mean_dict[k + "_motor_mean"][:steps], mean_dict[k + "_drpy_mean"][:steps]], 1), dtype=dt, device=dev) for k in self.keys]
        self.std = [torch.as_tensor(np.concatenate([mean_dict[k + "_motor_std"][:steps], mean_dict[k + "_drpy_std"][:steps]], 1), dtype=dt, device=dev) for k in self.keys]
        self.acc = torch.zeros(self.n, 15, dtype=dt, device=dev)
        self.reward = torch.zeros(self.n, dtype=dt, device=dev)

    def evaluate(self, solutions):
        """solutions [pop,48] in [-1,1] (identical on every rank) -> reward [pop] = mean over the gait keys (identical on every rank)."""
        import torch
        from .etg import dynamic_dict_to_row, param2dynamic_dict
        rows = np.array([dynamic_dict_to_row(param2dynamic_dict(np.asarray(s))) for s in np.asarray(solutions)[self.lo:self.hi]])
        env = self.env
        env.set_dynamics(np.tile(rows, (len(self.keys), 1)))          # env.reset(hardset=False, dynamic_param=...) :55
        env.reset()
        self.acc.zero_()
        es, stream, pl = env.obs.element_size(), env._stream(), self.pop_local
        for t in range(self.steps):
            _, _, _, info = env.step(self.actions[t], donef=False)
            for ki in range(len(self.keys)):
                sl = slice(ki * pl, (ki + 1) * pl)
                rc = self.lib.b2q_dyn_accumulate(info[sl].data_ptr(), self.mean[ki][t].data_ptr(), self.std[ki][t].data_ptr(), self.acc[sl].data_ptr(), pl, es, stream)
                assert rc == 0
        assert self.lib.b2q_dyn_finish(self.acc.data_ptr(), self.steps, self.reward.data_ptr(), self.n, es, stream) == 0
        local = self.reward.reshape(len(self.keys), pl).mean(0)        # (reward1 + reward2) / 2, :73
        return all_gather_concat(local.contiguous(), self.world, self.rank)
