"""CPU probe: f32 device code (emulation) vs f64 oracle, free running — which gait / residual scale survives, and the drift."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "emu")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import emu
from oracle import oracle as O
from paddlerobotics_b200 import etg as E

def fit_shipped():
    table = np.load(os.path.join(ROOT, "tests", "golden", "gait_action_list_ETG_exp.npy"))
    cfg = O.default_config(); K = table.shape[0]; ts = 0.026 + 0.026 * np.arange(K)
    pose = np.array([0, .9, -1.8] * 4)
    A, Y = [], []
    for k in range(K):
        q = table[k] + pose
        for leg in (0, 1):
            foot = O.fk_leg(q[3 * leg:3 * leg + 3], (-1) ** (leg + 1)) + E.HIP_OFFSETS[leg]
            tt = ts[k] if leg == 0 else ts[k] + 0.25
            A.append(np.concatenate([O.etg_features(cfg, tt), [1.0]])); Y.append(foot - E.BASE_FOOT[leg])
    sol = np.linalg.lstsq(np.array(A), np.array(Y), rcond=None)[0]
    return sol[:20].T.copy(), sol[20].copy()

def run(w, b, scale, steps=1000, prec=0, seed=0):
    e = emu.EmuEnv(1, prec); o = O.OracleEnv()
    e.reset(w, b); o.reset(w, b)
    acts = np.random.default_rng(seed).uniform(-1, 1, (steps, 12)) * scale
    wq = wp = wr = 0.0; mism = 0; x0 = o.get_state()[0]
    for k in range(steps):
        ob, rw, dn, inf = e.step(acts[k]); oo, ro, do, io = o.step(acts[k])
        st, so = e.get_state()[0].astype(np.float64), o.get_state()
        wq = max(wq, np.abs(st[13:25] - so[13:25]).max()); wp = max(wp, np.abs(st[:3] - so[:3]).max()); wr = max(wr, abs(float(rw[0]) - ro))
        mism += int(not np.array_equal(ob[0][3:7].astype(np.float64), oo[3:7]))
        if do or dn[0]:
            print("  done at step", k, "oracle", do, "emu", bool(dn[0])); break
    print("  steps %d  q %.3g  pos %.3g  rew %.3g  contact-mism %d  speed %.3f m/s" % (k + 1, wq, wp, wr, mism, (so[0] - x0) / (0.026 * (k + 1))))
    e.close()

if __name__ == "__main__":
    ws, bs = fit_shipped()
    layer = E.ETG_layer(0.5, 0.026, 20, 0.04, np.array([-np.pi / 2, 0]), 0.2, 0.5)
    wd, bd, _ = E.Opt_with_points(ETG=layer, ETG_T=0.5, Footheight=0.1, Steplength=0.05)
    wst, bst, _ = E.Opt_with_points(ETG=layer, ETG_T=0.5, Footheight=0.03, Steplength=0.02)
    for name, (w, b) in (("shipped", (ws, bs)), ("stable", (wst, bst)), ("default", (wd, bd))):
        for sc in (0.0, 0.1, 0.3):
            print(name, "residual x", sc); run(w, b, sc, steps=int(sys.argv[1]) if len(sys.argv) > 1 else 1000)
