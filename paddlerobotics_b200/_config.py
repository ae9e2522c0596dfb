"""ctypes mirror of include/b2q.h B2QConfig (no CUDA needed to import)."""
import ctypes as C

ACT_DIM, OBS_DIM, INFO_DIM, STATE_DIM, DYN_DIM, ETG_H = 12, 49, 56, 37, 48, 20

INFO = dict(velx=0, torso=1, feet=2, up=3, tau=4, stand=5, badfoot=6, footcontact=7, done=8, nan=9, energy=10, base_z=11,
            ETG_act=slice(12, 24), real_action=slice(24, 36), rpy=slice(36, 39), drpy=slice(39, 42), joint_angle=slice(42, 54),
            fall=54, step=55)


class B2QConfig(C.Structure):
    _fields_ = [
        ("num_envs", C.c_int32), ("device", C.c_int32), ("precision", C.c_int32), ("threads_per_block", C.c_int32),
        ("sim_dt", C.c_double), ("action_repeat", C.c_int32), ("solver_iters", C.c_int32),
        ("erp", C.c_double), ("warmstart", C.c_double), ("contact_margin", C.c_double),
        ("action_interp", C.c_int32), ("torque_limit", C.c_double), ("settle_steps", C.c_int32),
        ("max_episode_steps", C.c_int32), ("etg_enabled", C.c_int32), ("action_filter", C.c_int32), ("filter_highcut", C.c_double),
        ("etg_T", C.c_double), ("etg_T2", C.c_double), ("etg_sigma_sq", C.c_double), ("etg_amp", C.c_double),
        ("etg_phase0", C.c_double), ("etg_phase1", C.c_double),
        ("w_torso", C.c_double), ("w_feet", C.c_double), ("w_up", C.c_double), ("w_tau", C.c_double), ("w_stand", C.c_double),
        ("w_badfoot", C.c_double), ("w_footcontact", C.c_double), ("w_done", C.c_double), ("reward_p", C.c_double), ("vel_d", C.c_double),
        ("foot_radius", C.c_double), ("ring_depth", C.c_int32), ("auto_reset", C.c_int32), ("terrain_type", C.c_int32),
        ("hf_nx", C.c_int32), ("hf_ny", C.c_int32), ("hf_x0", C.c_double), ("hf_y0", C.c_double), ("hf_cell", C.c_double),
        ("hf_host", C.POINTER(C.c_double)),
        ("clip_motor_commands", C.c_int32), ("max_angle_change", C.c_double),
        ("sensor_dis", C.c_int32), ("sensor_contact", C.c_int32), ("sensor_imu", C.c_int32), ("sensor_motor", C.c_int32), ("sensor_etg", C.c_int32),
        ("obs_normal", C.c_int32), ("noise_stdev", C.c_double * 5), ("noise_seed", C.c_uint64),
        ("stuck_termination", C.c_int32), ("body_collisions", C.c_int32), ("motor_mode", C.c_int32), ("joint_limits", C.c_int32),
        ("external_force", C.c_int32), ("base_damping", C.c_double * 4), ("etg_foot_y_inset", C.c_double), ("knee_contacts", C.c_int32),
    ]
