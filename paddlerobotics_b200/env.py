"""Gym-style env surface of the reference (rlschool.make_env('Quadrupedal', ...): ETGRL/train.py:305-309) on top of
the CUDA engine.  Two classes:

* VecQuadrupedalEnv — N envs on one GPU.  Device API (torch tensors in/out, zero host traffic) for rollouts; host API
  (numpy in/out through pinned buffers) for callers that live on the CPU like the reference's train.py.
* QuadrupedalEnv    — N=1 wrapper with the reference's exact call shapes:
      obs, info = env.reset(ETG_w=w, ETG_b=b, x_noise=0)            (train.py:131)
      obs, reward, done, info = env.step(action, donef=False)         (train.py:147)
  `info` is a dict with the keys train.py consumes (velx, torso, feet, up, tau, ..., ETG_act, real_action,
  joint_angle, obs-IMU).

PyTorch is used only for device buffers / streams.  Every step runs in csrc/libb2q.so; nothing here computes physics.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._config import ACT_DIM, DYN_DIM, ETG_H, INFO, INFO_DIM, OBS_DIM, STATE_DIM, B2QConfig
from .etg import ETG_layer, Opt_with_points, dynamic_dict_to_row

_CFG_KEYS = {f[0] for f in B2QConfig._fields_}


def _check(lib, h, rc, what):
    if rc != 0:
        msg = lib.b2q_last_error(h)
        raise RuntimeError("%s failed (%d): %s" % (what, rc, msg.decode() if msg else "?"))


class VecQuadrupedalEnv:
    def __init__(self, num_envs, device=0, precision="f32", auto_reset=False, heightfield=None, **cfg):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("paddlerobotics_b200 needs a CUDA device: the hot path has no CPU fallback")
        self.device = torch.device("cuda", int(device))
        self.num_envs = int(num_envs)
        self.dtype = torch.float32 if precision in ("f32", 0) else torch.float64
        c = B2QConfig()
        self.lib.b2q_default_config(C.byref(c))
        c.num_envs, c.device, c.precision, c.auto_reset = self.num_envs, int(device), 0 if self.dtype == torch.float32 else 1, int(auto_reset)
        self._hf_keep = None
        if heightfield is not None:
            hf, x0, y0, cell = heightfield
            hf = np.ascontiguousarray(hf, dtype=np.float64)
            self._hf_keep = hf
            c.terrain_type, c.hf_ny, c.hf_nx = 1, hf.shape[0], hf.shape[1]
            c.hf_x0, c.hf_y0, c.hf_cell = float(x0), float(y0), float(cell)
            c.hf_host = hf.ctypes.data_as(C.POINTER(C.c_double))
        for k, v in cfg.items():
            if k not in _CFG_KEYS:
                raise TypeError("unknown config key %r" % k)
            setattr(c, k, v)
        self.cfg = c
        self.h = C.c_void_p()
        rc = self.lib.b2q_create(C.byref(c), C.byref(self.h))
        if rc != 0:
            raise RuntimeError("b2q_create failed (%d): %s" % (rc, self.lib.b2q_last_error(None).decode()))
        n, dev, dt = self.num_envs, self.device, self.dtype
        self.obs = torch.zeros(n, OBS_DIM, device=dev, dtype=dt)
        self.reward = torch.zeros(n, device=dev, dtype=dt)
        self.done = torch.zeros(n, device=dev, dtype=torch.uint8)
        self.info = torch.zeros(n, INFO_DIM, device=dev, dtype=dt)
        self.control_dt = c.sim_dt * c.action_repeat
        self.observation_dim, self.action_dim = OBS_DIM, ACT_DIM
        # host API staging (pinned) — allocated lazily
        self._h_act = self._h_obs = self._h_rew = self._h_done = self._d_act = self._h_info = None

    # ------------------------------------------------------------------ device API
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _t(self, x, shape):
        t = torch.as_tensor(x, dtype=self.dtype, device=self.device)
        return t.expand(shape).contiguous() if tuple(t.shape) != tuple(shape) else t.contiguous()

    def set_dynamics(self, dyn=None, env_mask=None):
        """dyn: [N,48] rows (engine layout, see etg.dynamic_dict_to_row) or None for defaults; re-settles masked envs."""
        d = None if dyn is None else self._t(dyn, (self.num_envs, DYN_DIM))
        m = None if env_mask is None else torch.as_tensor(env_mask, dtype=torch.uint8, device=self.device).contiguous()
        _check(self.lib, self.h, self.lib.b2q_set_dynamics(self.h, None if m is None else m.data_ptr(), None if d is None else d.data_ptr(), self._stream()), "b2q_set_dynamics")

    def reset(self, ETG_w=None, ETG_b=None, env_mask=None, **_ignored):
        w = None if ETG_w is None else self._t(torch.as_tensor(ETG_w, dtype=self.dtype).reshape(-1, 3, ETG_H), (self.num_envs, 3, ETG_H))
        b = None if ETG_b is None else self._t(torch.as_tensor(ETG_b, dtype=self.dtype).reshape(-1, 3), (self.num_envs, 3))
        m = None if env_mask is None else torch.as_tensor(env_mask, dtype=torch.uint8, device=self.device).contiguous()
        p = lambda t: None if t is None else t.data_ptr()
        _check(self.lib, self.h, self.lib.b2q_reset(self.h, p(m), p(w), p(b), self.obs.data_ptr(), self._stream()), "b2q_reset")
        return self.obs

    def step(self, action, donef=False):
        """action: [N,12] device tensor (joint-space residual, already scaled by act_bound)."""
        a = action if (isinstance(action, torch.Tensor) and action.dtype == self.dtype and action.device == self.device and action.is_contiguous()) \
            else self._t(action, (self.num_envs, ACT_DIM))
        rc = self.lib.b2q_step(self.h, a.data_ptr(), int(bool(donef)), self.obs.data_ptr(), self.reward.data_ptr(), self.done.data_ptr(),
                               self.info.data_ptr(), self._stream())
        if rc != 0:
            _check(self.lib, self.h, rc, "b2q_step")
        return self.obs, self.reward, self.done, self.info

    def get_state(self):
        s = torch.empty(self.num_envs, STATE_DIM, device=self.device, dtype=self.dtype)
        _check(self.lib, self.h, self.lib.b2q_get_state(self.h, s.data_ptr(), self._stream()), "b2q_get_state")
        return s

    def set_state(self, s):
        s = self._t(s, (self.num_envs, STATE_DIM))
        _check(self.lib, self.h, self.lib.b2q_set_state(self.h, s.data_ptr(), self._stream()), "b2q_set_state")

    def launch_count(self):
        return int(self.lib.b2q_launch_count(self.h))

    # ------------------------------------------------------------------ host API (numpy in / numpy out)
    def _host_bufs(self):
        if self._h_act is None:
            n, npdt = self.num_envs, self.dtype
            self._h_act = torch.empty(n, ACT_DIM, dtype=npdt).pin_memory()
            es = self.obs.element_size()
            self._h_out = torch.empty(n * (OBS_DIM + 1) * es + n, dtype=torch.uint8).pin_memory()     # obs | rew | done contiguous: one D2H
            self._h_obs = self._h_out[: n * OBS_DIM * es].view(npdt).reshape(n, OBS_DIM)
            self._h_rew = self._h_out[n * OBS_DIM * es: n * (OBS_DIM + 1) * es].view(npdt)
            self._h_done = self._h_out[n * (OBS_DIM + 1) * es:]
            self._np_act, self._np_obs, self._np_rew, self._np_done = self._h_act.numpy(), self._h_obs.numpy(), self._h_rew.numpy(), self._h_done.numpy()

    def step_host(self, action_np, donef=False, info=False):
        """The reference-facing call with HOST buffers (numpy in / numpy out), one C call and one stream sync: the step kernel
        reads the actions from and stores obs / reward / done to pinned host memory (b2q_step_host).  info=True also returns
        the [N,56] info rows.  The returned arrays are views of the pinned buffers (overwritten by the next call)."""
        self._host_bufs()
        np.copyto(self._np_act, np.asarray(action_np).reshape(self.num_envs, ACT_DIM), casting="same_kind")
        if info and self._h_info is None:
            self._h_info = torch.empty(self.num_envs, INFO_DIM, dtype=self.dtype).pin_memory()
            self._np_info = self._h_info.numpy()
        rc = self.lib.b2q_step_host(self.h, self._h_act.data_ptr(), int(bool(donef)), self._h_obs.data_ptr(), self._h_rew.data_ptr(),
                                    self._h_done.data_ptr(), self._h_info.data_ptr() if info else None, self._stream())
        if rc != 0:
            _check(self.lib, self.h, rc, "b2q_step_host")
        if info:
            return self._np_obs, self._np_rew, self._np_done, self._np_info
        return self._np_obs, self._np_rew, self._np_done

    def h2d_bytes_per_step(self):
        return self.num_envs * ACT_DIM * self.obs.element_size()

    def d2h_bytes_per_step(self):
        return self.num_envs * (OBS_DIM * self.obs.element_size() + self.obs.element_size() + 1)

    def close(self):
        if getattr(self, "h", None):
            self.lib.b2q_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _Space:
    def __init__(self, n):
        self.shape = (n,)


def info_dict(row):
    """[INFO_DIM] row -> the reference's info dict keys."""
    row = np.asarray(row, dtype=np.float64)
    d = {k: float(row[v]) for k, v in INFO.items() if isinstance(v, int)}
    d["ETG_act"] = row[INFO["ETG_act"]].copy()
    d["real_action"] = row[INFO["real_action"]].copy()
    d["joint_angle"] = row[INFO["joint_angle"]].copy()
    d["obs-IMU"] = np.concatenate([row[INFO["rpy"]], row[INFO["drpy"]]])
    return d


class QuadrupedalEnv:
    """N=1 mirror of the reference env object (numpy in/out)."""

    def __init__(self, task="ground", motor_control_mode=None, render=False, sensor_mode=None, normal=1, dynamic_param=None,
                 reward_param=None, ETG=1, ETG_T=0.5, reward_p=5, ETG_path="None", random_param=None, ETG_H=20, vel_d=0.5,
                 step_y=0.05, enable_action_filter=0, device=0, precision="f32", **engine_cfg):
        if render:
            raise ValueError("render=True is not supported (no GUI on the GPU path)")
        if ETG_H != ETG_H_CONST:
            raise ValueError("ETG_H must be %d" % ETG_H_CONST)
        if task not in ("ground", "plane"):
            raise ValueError("task %r: only flat terrain is wired into make_env; pass heightfield= to VecQuadrupedalEnv for others" % (task,))
        cfg = dict(etg_T=float(ETG_T), etg_T2=float(ETG_T), reward_p=float(reward_p), vel_d=float(vel_d), etg_enabled=int(bool(ETG)),
                   action_filter=int(bool(enable_action_filter)))
        for k_ref, k_cfg in (("torso", "w_torso"), ("feet", "w_feet"), ("up", "w_up"), ("tau", "w_tau"), ("stand", "w_stand"),
                             ("badfoot", "w_badfoot"), ("footcontact", "w_footcontact"), ("done", "w_done")):
            if reward_param and k_ref in reward_param:
                cfg[k_cfg] = float(reward_param[k_ref])
        cfg.update(engine_cfg)
        self.vec = VecQuadrupedalEnv(1, device=device, precision=precision, **cfg)
        if dynamic_param:
            self.vec.set_dynamics(dynamic_dict_to_row(dynamic_param)[None, :])
        self.observation_space, self.action_space = _Space(OBS_DIM), _Space(ACT_DIM)
        layer = ETG_layer(ETG_T, 0.026, ETG_H, 0.04, np.array([-np.pi / 2, 0]), 0.2, ETG_T)
        if ETG_path not in (None, "None", "") and str(ETG_path).endswith(".npz"):
            z = np.load(ETG_path)
            self._w, self._b = z["w"], z["b"]
        else:
            self._w, self._b, _ = Opt_with_points(ETG=layer, ETG_T=ETG_T, Footheight=0.1, Steplength=0.05)  # train.py:298-299 defaults

    def reset(self, ETG_w=None, ETG_b=None, x_noise=0, hardset=None, dynamic_param=None, **kw):
        if dynamic_param is not None:
            self.vec.set_dynamics(dynamic_dict_to_row(dynamic_param)[None, :])
        if ETG_w is not None:
            self._w = np.asarray(ETG_w)
        if ETG_b is not None:
            self._b = np.asarray(ETG_b)
        obs = self.vec.reset(self._w, self._b)
        info = {"ETG_act": np.zeros(12)}
        return obs[0].double().cpu().numpy(), info

    def step(self, action, donef=False, **kw):
        # one C call + one stream sync: actions in, obs / reward / done / info out through pinned host buffers
        obs, rew, done, info = self.vec.step_host(np.asarray(action).reshape(1, ACT_DIM), donef, info=True)
        return obs[0].astype(np.float64), float(rew[0]), bool(done[0]), info_dict(info[0])


ETG_H_CONST = ETG_H


def make_env(name, **kwargs):
    """rlschool.make_env('Quadrupedal', task=..., motor_control_mode=..., ...) — ETGRL/train.py:305-309."""
    if name != "Quadrupedal":
        raise ValueError("only 'Quadrupedal' is provided")
    return QuadrupedalEnv(**kwargs)
