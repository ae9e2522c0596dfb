"""Randomised feature combinations: the DEVICE code in CPU emulation (tests/emu, float64) against the float64 oracle, a few control steps each.
The single-feature tests (test_emu_features.py) pin every switch on its own; this one pins their interplay — sensor layout x noise x motor
mode x joint limits / knee contacts x terrain x action interpolation / filter / command clip x control latency from randomised dynamics."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))
import emu  # noqa: E402
from oracle import oracle as O  # noqa: E402

_OKEYS = {f[0] for f in O.Config._fields_}


from conftest import draw_feature_combo as _draw  # noqa: E402


@pytest.mark.parametrize("case", range(24))
def test_random_feature_combination(case, etg_stable):
    w, b = etg_stable
    rng = np.random.default_rng(1000 + case)
    kw, hf = _draw(rng)
    ocfg = O.default_config(**{k: v for k, v in kw.items() if k in _OKEYS})
    ekw = dict(kw)
    if hf is not None:
        O.set_heightfield(ocfg, *hf); ekw["heightfield"] = hf
    from paddlerobotics_b200.etg import dynamic_dict_to_row, param2dynamic_dict
    row = dynamic_dict_to_row(param2dynamic_dict(rng.uniform(-0.3, 0.3, 48)))   # randomised masses / inertias / gains / friction / gravity / control latency (train.py:112-126,253)
    e = emu.EmuEnv(1, 1, **ekw)
    e.set_dynamics(row[None])
    o = O.OracleEnv(ocfg, row)
    xo = float(rng.uniform(-0.1, 0.1))
    oe = e.reset(w, b, x_offset=[xo]); oo = o.reset(w, b, x_offset=xo)
    assert oe.shape[1] == oo.shape[0] and np.abs(oe[0] - oo).max() < 1e-8, kw
    if kw["external_force"]:
        f = rng.uniform(-15, 15, 3); e.set_force(f[None]); o.set_force(f)
    pose = np.array([0.0, 0.9, -1.8] * 4)
    for k in range(6):
        if kw["motor_mode"] == 2:
            a = np.zeros((12, 5)); a[:, 0] = pose + rng.uniform(-0.2, 0.2, 12); a[:, 1] = rng.uniform(60, 140, 12); a[:, 2] = rng.uniform(-1, 1, 12)
            a[:, 3] = rng.uniform(0.5, 3, 12); a[:, 4] = rng.uniform(-2, 2, 12); a = a.reshape(60)
        elif kw["motor_mode"] == 1:
            a = np.array([0.0, 1.0, -6.0] * 4) + rng.uniform(-1, 1, 12)
        else:
            a = rng.uniform(-0.3, 0.3, 12)
        ob, rw, dn, inf = o.step(a); ob2, rw2, dn2, inf2 = e.step(a)
        tol = 1e-6      # two different float64 formulations (link-coordinate ABA + DoF-space PGS vs composite inertia + contact-space PGS) on rough random dynamics
        assert np.abs(ob2[0] - ob).max() < tol and abs(rw2[0] - rw) < tol and bool(dn2[0]) == dn, (kw, k)
        assert np.abs(inf2[0] - inf).max() < tol, (kw, k)
        if dn:
            break
    e.close()


def test_diverged_state_on_a_height_field_ends_the_episode(etg_stable):
    """A non-finite state (reachable with the reference's own dynamics randomisation: zero friction, zero kd, tilted gravity) must end the
    episode in both implementations — the terrain lookup clamps NaN coordinates instead of indexing with them."""
    w, b = etg_stable
    z = np.zeros((20, 20)); hf = (z, -0.5, -0.5, 0.05)
    ocfg = O.default_config(); O.set_heightfield(ocfg, *hf)
    e = emu.EmuEnv(1, 1, heightfield=hf); o = O.OracleEnv(ocfg)
    e.reset(w, b); o.reset(w, b)
    s = o.get_state(); s[0] = np.nan
    o.set_state(s); e.set_state(s[None])
    _, _, dn, _ = o.step(np.zeros(12)); _, _, dn2, _ = e.step(np.zeros(12))
    assert dn and bool(dn2[0])
    e.close()
