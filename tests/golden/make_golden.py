"""Generates tests/golden/*.npz from the UNMODIFIED reference imported in this container (oracle/ref_shim.py).

Run here (the GPU box has no /root/reference):  python tests/golden/make_golden.py
The two .npy gait tables and the .pt checkpoint are the reference's own golden artefacts, copied byte-for-byte
(data, not source): SHA-256 in SURVEY.md App. A.
"""
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref_shim  # noqa: E402


def main():
    ns = ref_shim.load()
    a1, lm, rc, af = ns.a1, ns.laikago_motor, ns.robot_config, ns.action_filter
    rng = np.random.default_rng(20240924)
    out = {}
    # --- kinematics (a1.py:97-173)
    ang = rng.uniform([-0.6, -0.5, -2.6], [0.6, 2.5, -1.0], size=(64, 3))
    out["kin_angles"] = ang
    out["fk_right"] = np.array([a1.foot_position_in_hip_frame(a, -1) for a in ang])
    out["fk_left"] = np.array([a1.foot_position_in_hip_frame(a, 1) for a in ang])
    out["ik_right"] = np.array([a1.foot_position_in_hip_frame_to_joint_angle(f, -1) for f in out["fk_right"]])
    out["ik_left"] = np.array([a1.foot_position_in_hip_frame_to_joint_angle(f, 1) for f in out["fk_left"]])
    out["jac"] = np.array([[a1.analytical_leg_jacobian(a, leg) for leg in range(4)] for a in ang])
    q12 = rng.uniform(-0.4, 0.4, size=(32, 12)) + np.array([0, 0.9, -1.8] * 4)
    out["q12"] = q12
    out["feet_base"] = np.array([a1.foot_positions_in_base_frame(q) for q in q12])
    out["hip_offsets"] = a1.HIP_OFFSETS
    out["init_motor_angles"] = a1.INIT_MOTOR_ANGLES
    # --- motor model (laikago_motor.py:103-175)
    kp = np.array([a1.ABDUCTION_P_GAIN, a1.HIP_P_GAIN, a1.KNEE_P_GAIN] * 4)
    kd = np.array([a1.ABDUCTION_D_GAIN, a1.HIP_D_GAIN, a1.KNEE_D_GAIN] * 4)
    mm = lm.LaikagoMotorModel(kp=kp, kd=kd, motor_control_mode=rc.MotorControlMode.POSITION)
    cmd, q, qd = rng.uniform(-1, 1, (16, 12)), rng.uniform(-1, 1, (16, 12)), rng.uniform(-5, 5, (16, 12))
    out["motor_cmd"], out["motor_q"], out["motor_qd"], out["motor_kp"], out["motor_kd"] = cmd, q, qd, kp, kd
    out["motor_tau"] = np.array([mm.convert_to_torque(c, a, b, b, rc.MotorControlMode.POSITION)[0] for c, a, b in zip(cmd, q, qd)])
    mm2 = lm.LaikagoMotorModel(kp=kp, kd=kd, torque_limits=33.5, motor_control_mode=rc.MotorControlMode.POSITION)
    out["motor_tau_clip"] = np.array([mm2.convert_to_torque(c, a, b, b, rc.MotorControlMode.POSITION)[0] for c, a, b in zip(cmd, q, qd)])
    # --- MapToMinusPiToPi (minitaur.py:67-83)
    ang_w = rng.uniform(-12, 12, 64)
    out["wrap_in"], out["wrap_out"] = ang_w, np.array(ns.minitaur.MapToMinusPiToPi(list(ang_w)))
    # --- Butterworth action filter coefficients + a filtered sequence (action_filter.py:111-216)
    f = af.ActionFilterButter(sampling_rate=1 / 0.026, num_joints=12)
    out["butter_a"], out["butter_b"] = np.array(f.a), np.array(f.b)
    f.init_history(np.array([0, 0.9, -1.8] * 4))
    xs = np.array([0, 0.9, -1.8] * 4) + rng.uniform(-0.3, 0.3, (20, 12))
    out["butter_x"], out["butter_y"] = xs, np.array([f.filter(x) for x in xs])
    # --- train.py helpers on the restated ETG layer (train.py:59-126)
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from paddlerobotics_b200.etg import ETG_layer

    layer = ETG_layer(0.5, 0.026, 20, 0.04, np.array([-np.pi / 2, 0]), 0.2, 0.5)
    w0, b0, pts = ns.Opt_with_points(ETG=layer, ETG_T=0.5, Footheight=0.1, Steplength=0.05)
    out["opt_w0"], out["opt_b0"], out["opt_points"] = w0, b0, pts
    sol = rng.normal(0, 0.02, 12)
    w1, b1, _ = ns.Opt_with_points(ETG=layer, ETG_T=0.5, w0=w0, b0=b0, points=pts + sol.reshape(-1, 2))
    out["opt_sol"], out["opt_w1"], out["opt_b1"] = sol, w1, b1
    p48 = rng.uniform(-1.2, 1.2, 48)
    dd = ns.param2dynamic_dict(p48)
    out["dyn_in"] = p48
    for k, v in dd.items():
        out["dyn_" + k] = np.asarray(v)
    # --- SimpleGA ask/tell (es.py:257-314) with the global RNG seeded
    np.random.seed(0)
    ga = ns.es.SimpleGA(12, sigma_init=0.02, sigma_decay=0.99, sigma_limit=0.005, elite_ratio=0.1, weight_decay=0.005, popsize=40, param=np.zeros(12))
    s1 = ga.ask()
    fit = np.sin(np.arange(40) * 0.7) + s1[:, 0] * 10
    ga.tell(fit.copy())
    s2 = ga.ask()
    out["ga_s1"], out["ga_fit"], out["ga_s2"], out["ga_best"], out["ga_sigma"] = s1, fit, s2, ga.best_param, np.array(ga.sigma)
    # --- MLP known answers from the shipped checkpoint (mujoco_model.py:44-89)
    import torch

    ck = torch.load(ns.REF + "/deployment/exp/stairstair/StairStair3_BC1_itr_500383.pt", map_location="cpu")
    model = ns.MujocoModel(46, 12)
    model.load_state_dict(ck)
    obs = torch.tensor(rng.normal(0, 1, (16, 46)), dtype=torch.float32)
    act = torch.tensor(rng.uniform(-1, 1, (16, 12)), dtype=torch.float32)
    with torch.no_grad():
        mean, logstd = model.policy(obs)
        q1, q2 = model.value(obs, act)
    out["mlp_obs"], out["mlp_act"] = obs.numpy(), act.numpy()
    out["mlp_mean"], out["mlp_logstd"], out["mlp_q1"], out["mlp_q2"] = mean.numpy(), logstd.numpy(), q1.numpy(), q2.numpy()
    # --- dynamics-identification loss (model/Dynamic_parallel_model.py:29-41), ast-extracted (the module imports rlschool)
    import ast
    src = open(ns.REF + "/model/Dynamic_parallel_model.py").read()
    env = {"np": np}
    for node in ast.parse(src).body:
        if isinstance(node, ast.FunctionDef) and node.name == "loss_func":
            exec(compile(ast.Module([node], []), "Dynamic_parallel_model.py", "exec"), env)
    T = 100
    md = {"exp_motor_mean": rng.normal(0, 0.3, (T, 12)), "exp_motor_std": rng.uniform(0.05, 0.2, (T, 12)),
          "exp_drpy_mean": rng.normal(0, 0.5, (T, 3)), "exp_drpy_std": rng.uniform(0.2, 0.6, (T, 3))}
    drpy, motor = rng.normal(0, 0.6, (T, 3)), rng.normal(0, 0.35, (T, 12))
    out["dynloss_drpy"], out["dynloss_motor"] = drpy, motor
    for k, v in md.items():
        out["dynloss_" + k] = v
    out["dynloss_value"] = np.array(env["loss_func"](drpy, motor, md, "exp"))
    # --- BC student observation noise (BCtrain.py:53-59), ast-extracted (the module imports rlschool); global NumPy RNG seeded
    src = open(ns.REF + "/BCtrain.py").read()
    env = {"np": np, "copy": __import__("copy").copy}
    for node in ast.parse(src).body:
        if isinstance(node, ast.FunctionDef) and node.name == "obs2noise":
            exec(compile(ast.Module([node], []), "BCtrain.py", "exec"), env)
    obs49 = rng.normal(0, 1, (8, 49))
    np.random.seed(7)
    out["noise_obs_in"], out["noise_obs_out"] = obs49, np.array([env["obs2noise"](o) for o in obs49])
    # --- observation history stack (deployment/envs/EnvWrapper.py:195-241), class ast-extracted, driven by a fake env
    src = open(ns.REF + "/deployment/envs/EnvWrapper.py").read()
    env = {"np": np, "copy": __import__("copy").copy}
    for node in ast.parse(src).body:
        if isinstance(node, ast.ClassDef) and node.name == "ObservationWrapper":
            exec(compile(ast.Module([node], []), "EnvWrapper.py", "exec"), env)
    seq = rng.normal(0, 1, (12, 7))                       # reset obs + 11 step obs of a 7-dim sensor

    class _Fake:
        def __init__(self): self.k = 0
        def get_obs_dim(self): return 7
        def get_observation(self): self.k += 1; return seq[self.k].copy(), {}
        def reset(self, **kw): self.k = 0; return seq[0].copy(), {}
    out["hist_seq"] = seq
    for mode, T, I in (("stack", 3, 1), ("stack", 2, 2), ("GRU", 3, 1)):
        wr = env["ObservationWrapper"](_Fake(), None, {"RNN": {"time_steps": T, "time_interval": I, "mode": mode}})
        rows = [np.asarray(wr.reset()[0])] + [np.asarray(wr.get_observation()[0]) for _ in range(11)]
        out["hist_%s_%d_%d" % (mode, T, I)] = np.array(rows)
    np.savez_compressed(os.path.join(HERE, "reference_vectors.npz"), **out)
    # reference's own golden artefacts (data)
    shutil.copy(ns.REF + "/gait_action_list_ETG_exp.npy", os.path.join(HERE, "gait_action_list_ETG_exp.npy"))
    shutil.copy(ns.REF + "/deployment/exp/stairstair/gait_action_list_CPG_stairstair7_12_3.npy", os.path.join(HERE, "gait_action_list_CPG_stairstair7_12_3.npy"))
    shutil.copy(ns.REF + "/deployment/exp/stairstair/StairStair3_BC1_itr_500383.pt", os.path.join(HERE, "StairStair3_BC1_itr_500383.pt"))
    for fn in ("gait_action_list_ETG_exp.npy", "gait_action_list_CPG_stairstair7_12_3.npy", "StairStair3_BC1_itr_500383.pt"):
        os.chmod(os.path.join(HERE, fn), 0o644)
    print("wrote", sorted(out.keys()))


if __name__ == "__main__":
    main()
