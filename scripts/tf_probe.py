"""CPU probe: teacher-forced config 1 (default gait, residual x0.3, 1000 steps incl. resets): per-step f32 error statistics."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
import emu
from oracle import oracle as O
from paddlerobotics_b200 import etg as E
layer = E.ETG_layer(0.5, 0.026, 20, 0.04, np.array([-np.pi / 2, 0]), 0.2, 0.5)
w, b, _ = E.Opt_with_points(ETG=layer, ETG_T=0.5, Footheight=0.1, Steplength=0.05)
e = emu.EmuEnv(1, 0); o = O.OracleEnv(); e.reset(w, b); o.reset(w, b)
acts = np.random.default_rng(0).uniform(-1, 1, (1000, 12)) * 0.3
eq, eqd, er, ep, eqd_rel = [], [], [], [], []
nres = 0
for k in range(1000):
    e.set_state(o.get_state()[None, :])
    ob, rw, dn, inf = e.step(acts[k]); oo, ro, do, io = o.step(acts[k])
    st, so = e.get_state()[0].astype(np.float64), o.get_state()
    eq.append(np.abs(st[13:25] - so[13:25]).max()); ep.append(np.abs(st[:7] - so[:7]).max())
    d = np.abs(st[25:37] - so[25:37]).max(); eqd.append(d); eqd_rel.append(d / max(1.0, np.abs(so[25:37]).max()))
    er.append(abs(float(rw[0]) - ro) / max(1.0, abs(ro)))
    assert bool(dn[0]) == do and np.array_equal(ob[0][3:7].astype(np.float64), oo[3:7]), k
    if do:
        o.reset(); e.reset(); nres += 1
for n, v in (("q", eq), ("pose", ep), ("qd abs", eqd), ("qd rel", eqd_rel), ("rew rel", er)):
    v = np.array(v); print("%-8s max %.3g  p99 %.3g  median %.3g  argmax %d" % (n, v.max(), np.percentile(v, 99), np.median(v), v.argmax()))
print("resets", nres)
