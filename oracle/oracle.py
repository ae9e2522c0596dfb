"""ctypes binding of the CPU float64 oracle (oracle/b2q_oracle.c).

TEST INFRASTRUCTURE ONLY — imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (paddlerobotics_b200) never imports it.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libb2q_oracle.so")

NJ, HIST, OBS_DIM, INFO_DIM, NPARAM, ETG_H, HIST_W = 12, 128, 49, 56, 48, 20, 43


def build(force=False):
    src = [os.path.join(_HERE, f) for f in ("b2q_oracle.c", "b2q_oracle.h")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src if os.path.exists(s)):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


class Config(C.Structure):
    _fields_ = [
        ("sim_dt", C.c_double), ("action_repeat", C.c_int), ("solver_iters", C.c_int),
        ("erp", C.c_double), ("warmstart", C.c_double), ("contact_margin", C.c_double),
        ("action_interp", C.c_int), ("torque_limit", C.c_double), ("settle_steps", C.c_int),
        ("max_episode_steps", C.c_int), ("etg_enabled", C.c_int), ("action_filter", C.c_int), ("filter_highcut", C.c_double),
        ("etg_T", C.c_double), ("etg_T2", C.c_double), ("etg_sigma_sq", C.c_double), ("etg_amp", C.c_double),
        ("etg_phase", C.c_double * 2),
        ("w_torso", C.c_double), ("w_feet", C.c_double), ("w_up", C.c_double), ("w_tau", C.c_double),
        ("w_stand", C.c_double), ("w_badfoot", C.c_double), ("w_footcontact", C.c_double), ("w_done", C.c_double),
        ("reward_p", C.c_double), ("vel_d", C.c_double), ("foot_radius", C.c_double),
        ("terrain_type", C.c_int), ("hf_nx", C.c_int), ("hf_ny", C.c_int),
        ("hf_x0", C.c_double), ("hf_y0", C.c_double), ("hf_cell", C.c_double), ("hf", C.POINTER(C.c_double)),
        ("clip_motor_commands", C.c_int), ("max_angle_change", C.c_double),
        ("sensor_dis", C.c_int), ("sensor_contact", C.c_int), ("sensor_imu", C.c_int), ("sensor_motor", C.c_int), ("sensor_etg", C.c_int), ("obs_normal", C.c_int),
        ("noise_stdev", C.c_double * 5), ("noise_seed", C.c_ulonglong),
        ("stuck_termination", C.c_int), ("body_collisions", C.c_int), ("motor_mode", C.c_int), ("joint_limits", C.c_int), ("external_force", C.c_int),
        ("base_damping", C.c_double * 4), ("etg_foot_y_inset", C.c_double), ("knee_contacts", C.c_int),
    ]


class Env(C.Structure):
    _fields_ = [
        ("pos", C.c_double * 3), ("quat", C.c_double * 4), ("vlin", C.c_double * 3), ("vang", C.c_double * 3),
        ("q", C.c_double * 12), ("qd", C.c_double * 12),
        ("last_action", C.c_double * 12), ("has_last", C.c_int),
        ("lam_warm", C.c_double * 4), ("step_count", C.c_int), ("rpy0", C.c_double * 3),
        ("etg_act", C.c_double * 12), ("etg_w", (C.c_double * ETG_H) * 3), ("etg_b", C.c_double * 3),
        ("param", C.c_double * NPARAM),
        ("hist", (C.c_double * HIST_W) * HIST), ("hist_len", C.c_int), ("hist_head", C.c_int),
        ("contact", C.c_int * 4), ("last_tau", C.c_double * 12),
        ("fx1", C.c_double * 12), ("fx2", C.c_double * 12), ("fy1", C.c_double * 12), ("fy2", C.c_double * 12),
        ("snap", C.c_double * 37), ("snap_obs", C.c_double * HIST_W), ("snap_lam", C.c_double * 4),
        ("pos_hist", (C.c_double * 3) * 10), ("ext_force", C.c_double * 3), ("lam_lim", C.c_double * 12), ("env_id", C.c_int), ("hyb", (C.c_double * 12) * 4),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        assert _lib.orc_sizeof_env() == C.sizeof(Env), (_lib.orc_sizeof_env(), C.sizeof(Env))
        dp = C.POINTER(C.c_double)
        _lib.orc_energy.restype = C.c_double
        _lib.orc_energy.argtypes = [C.POINTER(Config), C.POINTER(Env), dp, dp]
    return _lib


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(C.POINTER(C.c_double))


def default_config(**kw):
    c = Config()
    lib().orc_default_config(C.byref(c))
    for k, v in kw.items():
        if k == "etg_phase":
            c.etg_phase[0], c.etg_phase[1] = v
        elif k in ("noise_stdev", "base_damping"):
            arr = getattr(c, k)
            for i, x in enumerate(v):
                arr[i] = float(x)
        else:
            setattr(c, k, v)
    return c


def default_param():
    p = np.zeros(NPARAM)
    lib().orc_default_param(p.ctypes.data_as(C.POINTER(C.c_double)))
    return p


def set_heightfield(cfg, hf, x0, y0, cell):
    hf = np.ascontiguousarray(hf, dtype=np.float64)
    cfg._hf_keep = hf
    cfg.terrain_type = 1
    cfg.hf_ny, cfg.hf_nx = hf.shape
    cfg.hf_x0, cfg.hf_y0, cfg.hf_cell = x0, y0, cell
    cfg.hf = hf.ctypes.data_as(C.POINTER(C.c_double))


# ---- pure pieces ---------------------------------------------------------------------------------
def etg_features(cfg, t):
    r = np.zeros(ETG_H)
    lib().orc_etg_features(C.byref(cfg), C.c_double(t), r.ctypes.data_as(C.POINTER(C.c_double)))
    return r


def etg_act(cfg, w, b, t):
    w, wp = _d(np.asarray(w).reshape(3, ETG_H))
    b, bp = _d(np.asarray(b).reshape(3))
    act = np.zeros(12)
    foot = np.zeros(12)
    lib().orc_etg_act(C.byref(cfg), wp, bp, C.c_double(t), act.ctypes.data_as(C.POINTER(C.c_double)),
                      foot.ctypes.data_as(C.POINTER(C.c_double)))
    return act, foot.reshape(4, 3)


def ik_leg(foot, sign):
    f, fp = _d(foot)
    a = np.zeros(3)
    lib().orc_ik_leg(fp, C.c_int(sign), a.ctypes.data_as(C.POINTER(C.c_double)))
    return a


def fk_leg(ang, sign):
    f, fp = _d(ang)
    a = np.zeros(3)
    lib().orc_fk_leg(fp, C.c_int(sign), a.ctypes.data_as(C.POINTER(C.c_double)))
    return a


def leg_jacobian(ang, leg_id):
    f, fp = _d(ang)
    J = np.zeros(9)
    lib().orc_leg_jacobian(fp, C.c_int(leg_id), J.ctypes.data_as(C.POINTER(C.c_double)))
    return J.reshape(3, 3)


def motor_torque(kp, kd, target, q, qd, limit=0.0):
    (kp, a), (kd, b), (target, c), (q, d), (qd, e) = map(_d, (kp, kd, target, q, qd))
    tau = np.zeros(12)
    lib().orc_motor_torque(a, b, c, d, e, C.c_double(limit), tau.ctypes.data_as(C.POINTER(C.c_double)))
    return tau


def butter2(highcut, fs):
    b, a = np.zeros(3), np.zeros(3)
    dp = C.POINTER(C.c_double)
    lib().orc_butter2(C.c_double(highcut), C.c_double(fs), b.ctypes.data_as(dp), a.ctypes.data_as(dp))
    return b, a


def filter_sequence(b, a, xs, init):
    """ActionFilter.init_history(init) then filter(x) for every row of xs (one joint per column)."""
    xs = np.asarray(xs, dtype=np.float64)
    ys = np.zeros_like(xs)
    dp = C.POINTER(C.c_double)
    b, a = np.ascontiguousarray(b, dtype=np.float64), np.ascontiguousarray(a, dtype=np.float64)
    for j in range(xs.shape[1]):
        h = [C.c_double(float(init[j])) for _ in range(4)]
        for k in range(xs.shape[0]):
            y = C.c_double()
            lib().orc_filter_step(b.ctypes.data_as(dp), a.ctypes.data_as(dp), C.c_double(float(xs[k, j])), C.byref(h[0]), C.byref(h[1]), C.byref(h[2]), C.byref(h[3]), C.byref(y))
            ys[k, j] = y.value
    return ys


def quat_to_rpy(q):
    q, qp = _d(q)
    r = np.zeros(3)
    lib().orc_quat_to_rpy(qp, r.ctypes.data_as(C.POINTER(C.c_double)))
    return r


# ---- env wrapper -----------------------------------------------------------------------------------
class OracleEnv:
    """One float64 oracle environment (N=1 mirror of the reference env)."""

    def __init__(self, cfg=None, param=None, settle=True):
        self.cfg = cfg if cfg is not None else default_config()
        self.e = Env()
        p = default_param() if param is None else np.ascontiguousarray(param, dtype=np.float64)
        lib().orc_env_init(C.byref(self.cfg), C.byref(self.e), p.ctypes.data_as(C.POINTER(C.c_double)))
        if settle:
            lib().orc_env_settle(C.byref(self.cfg), C.byref(self.e))

    def obs_dim(self):
        return int(lib().orc_obs_dim(C.byref(self.cfg)))

    def set_force(self, f=None):
        for k in range(3):
            self.e.ext_force[k] = 0.0 if f is None else float(f[k])

    def reset(self, etg_w=None, etg_b=None, x_offset=0.0):
        obs = np.zeros(self.obs_dim())
        wp = bp = None
        if etg_w is not None:
            w, wp = _d(np.asarray(etg_w).reshape(3, ETG_H))
        if etg_b is not None:
            b, bp = _d(np.asarray(etg_b).reshape(3))
        lib().orc_env_reset_ex(C.byref(self.cfg), C.byref(self.e), wp, bp, C.c_double(float(x_offset)), obs.ctypes.data_as(C.POINTER(C.c_double)))
        return obs

    def step(self, action, donef=False):
        a, ap = _d(action)
        obs = np.zeros(self.obs_dim())
        info = np.zeros(INFO_DIM)
        rew = C.c_double()
        done = C.c_int()
        lib().orc_env_step(C.byref(self.cfg), C.byref(self.e), ap, C.c_int(int(donef)),
                           obs.ctypes.data_as(C.POINTER(C.c_double)), C.byref(rew), C.byref(done),
                           info.ctypes.data_as(C.POINTER(C.c_double)))
        return obs, rew.value, bool(done.value), info

    def substep(self, target):
        t, tp = _d(target)
        lib().orc_substep(C.byref(self.cfg), C.byref(self.e), tp)

    # state access: [pos3 quat4 vlin3 vang3 q12 qd12] (37)
    def get_state(self):
        e = self.e
        return np.concatenate([np.array(e.pos), np.array(e.quat), np.array(e.vlin), np.array(e.vang), np.array(e.q), np.array(e.qd)])

    def set_state(self, s):
        s = np.asarray(s, dtype=np.float64)
        e = self.e
        for name, lo, hi in (("pos", 0, 3), ("quat", 3, 7), ("vlin", 7, 10), ("vang", 10, 13), ("q", 13, 25), ("qd", 25, 37)):
            arr = getattr(e, name)
            for i in range(hi - lo):
                arr[i] = float(s[lo + i])

    def forward_dynamics(self, tau):
        t, tp = _d(tau)
        qdd, wd, vd = np.zeros(12), np.zeros(3), np.zeros(3)
        dp = C.POINTER(C.c_double)
        lib().orc_forward_dynamics(C.byref(self.cfg), C.byref(self.e), tp, qdd.ctypes.data_as(dp), wd.ctypes.data_as(dp), vd.ctypes.data_as(dp))
        return qdd, wd, vd

    def mass_matrix(self):
        M = np.zeros((18, 18))
        lib().orc_mass_matrix(C.byref(self.cfg), C.byref(self.e), M.ctypes.data_as(C.POINTER(C.c_double)))
        return M

    def energy(self):
        k, p = C.c_double(), C.c_double()
        tot = lib().orc_energy(C.byref(self.cfg), C.byref(self.e), C.byref(k), C.byref(p))
        return tot, k.value, p.value

    def foot_world(self):
        f = np.zeros((4, 3))
        lib().orc_foot_world(C.byref(self.e), f.ctypes.data_as(C.POINTER(C.c_double)))
        return f


class OracleBatch:
    """N oracle envs stepped with pthreads — the CPU baseline (kind: "port")."""

    def __init__(self, n, cfg=None, params=None, etg_w=None, etg_b=None):
        self.n = n
        self.cfg = cfg if cfg is not None else default_config()
        self.envs = (Env * n)()
        proto = OracleEnv(self.cfg, None if params is None else params[0])
        for i in range(n):
            if params is None or i == 0 or np.array_equal(params[i], params[0]):
                C.memmove(C.byref(self.envs[i]), C.byref(proto.e), C.sizeof(Env))
            else:
                o = OracleEnv(self.cfg, params[i])
                C.memmove(C.byref(self.envs[i]), C.byref(o.e), C.sizeof(Env))
        for i in range(n):
            self.envs[i].env_id = i
        self.obs = np.zeros((n, OBS_DIM))
        dp = C.POINTER(C.c_double)
        for i in range(n):
            wp = bp = None
            if etg_w is not None:
                w, wp = _d(np.asarray(etg_w[i] if np.ndim(etg_w) == 3 else etg_w).reshape(3, ETG_H))
            if etg_b is not None:
                b, bp = _d(np.asarray(etg_b[i] if np.ndim(etg_b) == 2 else etg_b).reshape(3))
            lib().orc_env_reset(C.byref(self.cfg), C.byref(self.envs[i]), wp, bp, self.obs[i].ctypes.data_as(dp))
        self.rew = np.zeros(n)
        self.done = np.zeros(n, dtype=np.int32)
        self.info = np.zeros((n, INFO_DIM))

    def step(self, actions, donef=False, auto_reset=True, nthreads=1):
        a, ap = _d(np.asarray(actions).reshape(self.n, 60 if self.cfg.motor_mode == 2 else 12))
        dp = C.POINTER(C.c_double)
        lib().orc_batch_step(C.byref(self.cfg), self.envs, C.c_int(self.n), ap, C.c_int(int(donef)), C.c_int(int(auto_reset)),
                             self.obs.ctypes.data_as(dp), self.rew.ctypes.data_as(dp),
                             self.done.ctypes.data_as(C.POINTER(C.c_int)), self.info.ctypes.data_as(dp), C.c_int(nthreads))
        return self.obs, self.rew, self.done, self.info

    def rollout(self, actions, auto_reset=True, nthreads=1):
        """actions [K,n,12]: K control steps per env, envs distributed over pthreads (no per-step barrier)."""
        a, ap = _d(np.asarray(actions).reshape(-1, self.n, 60 if self.cfg.motor_mode == 2 else 12))
        K = a.shape[0]
        ret = np.zeros(self.n)
        nd = np.zeros(self.n, dtype=np.int32)
        dp = C.POINTER(C.c_double)
        lib().orc_batch_rollout(C.byref(self.cfg), self.envs, C.c_int(self.n), ap, C.c_int(K), C.c_int(int(auto_reset)),
                                self.obs.ctypes.data_as(dp), ret.ctypes.data_as(dp), nd.ctypes.data_as(C.POINTER(C.c_int)), C.c_int(nthreads))
        return ret, nd
