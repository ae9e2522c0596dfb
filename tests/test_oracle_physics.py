"""Physical invariants of the float64 dynamics oracle (the part of the path that has no reference golden data)."""
import numpy as np

from oracle import oracle as O


def _free_env(dt, gravity=(0, 0, -10.0)):
    cfg = O.default_config(sim_dt=dt)
    p = O.default_param(); p[:24] = 0; p[26:29] = gravity       # no PD torques
    env = O.OracleEnv(cfg, p, settle=False)
    r = np.random.default_rng(1)
    s = np.zeros(37); s[2] = 10.0
    s[3:7] = [0.1, 0.2, 0.3, 0.9]; s[3:7] /= np.linalg.norm(s[3:7])
    s[7:10] = r.uniform(-1, 1, 3); s[10:13] = r.uniform(-2, 2, 3)
    s[13:25] = np.array([0, 0.9, -1.8] * 4) + r.uniform(-0.3, 0.3, 12); s[25:37] = r.uniform(-3, 3, 12)
    env.set_state(s)
    return env


def test_energy_conservation_first_order():
    """Free flight: energy error shrinks linearly with dt (semi-implicit Euler) => Coriolis/gravity terms consistent."""
    errs = []
    for dt in (1e-3, 1e-4):
        env = _free_env(dt)
        e0 = env.energy()[0]
        for _ in range(int(round(0.05 / dt))):
            env.substep(np.zeros(12))
        errs.append(abs(env.energy()[0] - e0))
    assert errs[1] < errs[0] * 0.15
    assert errs[1] < 5e-3


def test_free_fall_com_acceleration():
    env = _free_env(1e-4)
    M = env.mass_matrix()
    assert np.allclose(M, M.T, atol=1e-12) and np.linalg.eigvalsh(M).min() > 0
    assert np.isclose(M[3, 3], 12.453, atol=1e-9)               # total mass on the base-linear diagonal
    qdd, wd, vd = env.forward_dynamics(np.zeros(12))
    # total linear momentum rate = m g: check through two substeps of the COM is overkill; check dynamics symmetry instead:
    env2 = _free_env(1e-4, gravity=(0, 0, 0))
    qdd0, wd0, vd0 = env2.forward_dynamics(np.zeros(12))
    assert np.allclose(qdd, qdd0, atol=1e-10) and np.allclose(wd, wd0, atol=1e-10)   # uniform gravity does not bend joints
    assert np.allclose(vd - vd0, [0, 0, -10.0], atol=1e-10)


def test_static_equilibrium_on_flat_ground():
    env = O.OracleEnv()
    pose = np.array([0, 0.9, -1.8] * 4)
    for _ in range(1500):
        env.substep(pose)
    s = env.get_state()
    assert np.abs(s[7:13]).max() < 2e-3 and np.abs(s[25:37]).max() < 5e-3   # at rest
    assert 0.24 < s[2] < 0.28
    lam = np.array(env.e.lam_warm)
    assert np.isclose(lam.sum(), 12.453 * 10 * 0.002, rtol=2e-3)           # normal impulses carry the weight
    assert all(env.e.contact[k] == 1 for k in range(4))


def test_friction_cone_and_no_penetration_growth():
    env = O.OracleEnv()
    s = env.get_state(); s[7] = 0.3
    env.set_state(s)
    pose = np.array([0, 0.9, -1.8] * 4)
    for _ in range(300):
        env.substep(pose)
        feet = env.foot_world()
        assert feet[:, 2].min() - 0.02 > -2e-3                              # ERP keeps penetration tiny
    assert abs(env.get_state()[7]) < 0.3                                    # friction removed forward momentum


def test_heightfield_plane_equivalence():
    """A flat height field at z=0 gives the same trajectory as the analytic plane."""
    a = O.OracleEnv()
    cfg = O.default_config()
    O.set_heightfield(cfg, np.zeros((64, 64)), -1.6, -1.6, 0.05)
    b = O.OracleEnv(cfg)
    assert np.allclose(a.get_state(), b.get_state(), atol=1e-12)


def test_slope_heightfield_contact_normal():
    cfg = O.default_config()
    xs = -1.6 + 0.05 * np.arange(64)
    hf = np.tile(0.1 * xs[None, :], (64, 1))                               # 10% slope along x
    O.set_heightfield(cfg, hf, -1.6, -1.6, 0.05)
    env = O.OracleEnv(cfg)
    s = env.get_state()
    assert np.all(np.isfinite(s)) and 0.15 < s[2] < 0.35


def test_shipped_gait_walks_in_this_physics(etg_shipped):
    """Physics regression pinned to reference-held data: the gait table the reference ships (gait_action_list_ETG_exp.npy, fitted back to
    W,b) must walk — forward speed in [0.3, 0.6] m/s, no fall in 600 control steps (the table's length) — in the oracle AND in the
    float32 device code (CPU emulation); the hand-picked gaits of round 1 are not the long-horizon workload any more."""
    import os, sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))
    import emu
    w, b = etg_shipped
    o = O.OracleEnv(); o.reset(w, b)
    e = emu.EmuEnv(1, 0); e.reset(w, b)
    worst_q = 0.0
    for k in range(600):
        ob, r, d, info = o.step(np.zeros(12))
        ob2, r2, d2, info2 = e.step(np.zeros(12))
        assert not d and not d2[0], k
        worst_q = max(worst_q, np.abs(e.get_state()[0][13:25] - o.get_state()[13:25]).max())
    v = o.get_state()[0] / (600 * 0.026)
    assert 0.3 < v < 0.6, v
    assert o.get_state()[2] > 0.2 and abs(info[36]) < 0.3 and abs(info[37]) < 0.3
    assert worst_q <= 1e-4, worst_q          # f32 device code, free running on the walking gait (BASELINE: joint-state drift <= 1e-4)
    e.close()
