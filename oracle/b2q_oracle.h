/* b2q_oracle.h — CPU float64 ORACLE for the ETGRL A1 per-step hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under paddlerobotics_b200/ may include, link
 * or call this.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, and only as the checker / CPU baseline.
 *
 * PARITY STATUS
 *   pinned   : ETG spline + closed-form A1 IK/FK (reference golden .npy files and the
 *              importable in-tree functions a1.py:97-173), PD motor model
 *              (laikago_motor.py:103-175), latency lerp (minitaur.py:1172-1193),
 *              action interpolation (minitaur.py:1384-1401), obs packing constants
 *              (EnvWrapper.py:50-109).  See tests/test_oracle_golden.py.
 *   UNPINNED : rigid-body dynamics + contact (pybullet/Bullet is a third-party
 *              dependency that is absent from /root/reference, version unpinned via
 *              rlschool>=1.0.2, ETGRL/README.md:9-12; no reference test or golden
 *              trajectory exists).  The dynamics below restate Bullet's published
 *              btMultiBody algorithm family (Featherstone ABA in link coordinates,
 *              semi-implicit Euler, PGS/sequential-impulse contact rows with pyramid
 *              friction, ERP, warm start); they are validated by physical invariants
 *              (tests/test_oracle_physics.py), not against pybullet output.
 *              Reward/termination formulas are this repo's own definition (rlschool
 *              source absent) — see DESIGN.md §3.
 */
#ifndef B2Q_ORACLE_H
#define B2Q_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

#define ORC_NJ 12
#define ORC_HIST 128          /* >= reference deque(maxlen=100), minitaur.py:175 */
#define ORC_OBS_DIM 49
#define ORC_INFO_DIM 56
#define ORC_NPARAM 48
#define ORC_ETG_H 20
#define ORC_HIST_W 43         /* q12 qd12 tau12 quat4 omega_body3, minitaur.py:1142-1149 */

typedef struct {
  double sim_dt;            /* 0.002 */
  int action_repeat;        /* 13 */
  int solver_iters;         /* int(300/13)=23 */
  double erp;               /* 0.2 */
  double warmstart;         /* 0.85 */
  double contact_margin;    /* 0.02 */
  int action_interp;        /* minitaur.py:1384-1401 */
  double torque_limit;      /* <=0: off (laikago_motor.py:168-173) */
  int settle_steps;         /* a1.py:294-297: 500 */
  int max_episode_steps;    /* >0: done also when the env's own step counter reaches it (donef=(steps>max_step), train.py:147) */
  int etg_enabled;          /* make_env(ETG=0): no open-loop reference, action = joint offsets (Dynamic_parallel_model.py:49,59-60) */
  int action_filter;        /* Butterworth low-pass on the joint targets, minitaur.py:248-251,1403-1422 */
  double filter_highcut;    /* 4 Hz, action_filter.py:42-44 */
  double etg_T, etg_T2, etg_sigma_sq, etg_amp, etg_phase[2];
  double w_torso, w_feet, w_up, w_tau, w_stand, w_badfoot, w_footcontact, w_done;
  double reward_p, vel_d;
  double foot_radius;       /* 0.02 */
  int terrain_type;         /* 0 plane, 1 height field */
  int hf_nx, hf_ny; double hf_x0, hf_y0, hf_cell; const double* hf; /* row-major [ny][nx] */
  int clip_motor_commands;  /* A1.ApplyAction -> _ClipMotorCommands (a1.py:428-458; enable_clip_motor_commands, default False a1.py:229) */
  double max_angle_change;  /* MAX_MOTOR_ANGLE_CHANGE_PER_STEP = 0.2, a1.py:62 */
  /* sensor_mode / normal of SimpleEnv.get_observation (deployment/envs/EnvWrapper.py:60-109; train.py:259-277,306) */
  int sensor_dis, sensor_contact, sensor_imu /*1: rpy+drpy, 2: drpy*/, sensor_motor /*1: q+qd, 2: q*/, sensor_etg, obs_normal;
  double noise_stdev[5];    /* Minitaur._AddSensorNoise: motor angle, velocity, torque, rpy, rpy rate (minitaur.py:59,1206-1211) */
  unsigned long long noise_seed;
  int stuck_termination;    /* rlschool [EXT]: base position spread over the last 10 control steps <= 2e-4 after step 10 */
  int body_collisions;      /* badfoot counts non-toe link / trunk-corner ground contacts */
  int motor_mode;           /* 0 POSITION, 1 TORQUE (laikago_motor.py:131-134) */
  int joint_limits;         /* URDF joint limits a1.py:186-223 as unilateral solver rows (one slot per leg) */
  int external_force;       /* base push (random_param['random_force'], train.py:254) */
  double base_damping[4];   /* Bullet btMultiBody base damping lin k1,k2 ang k1,k2 [EXT] */
  double etg_foot_y_inset;  /* make_env(step_y=): nominal footholds pulled towards the midline */
  int knee_contacts;        /* knee spheres (calf-joint origin, r 0.02) collide with the terrain */
} OrcConfig;

typedef struct {
  /* rigid state, pybullet conventions: world-frame base velocity, quat xyzw */
  double pos[3], quat[4], vlin[3], vang[3], q[12], qd[12];
  double last_action[12]; int has_last;
  double lam_warm[4];
  int step_count;
  double rpy0[3];
  double etg_act[12];
  double etg_w[3][ORC_ETG_H], etg_b[3];
  double param[ORC_NPARAM];
  double hist[ORC_HIST][ORC_HIST_W]; int hist_len, hist_head; /* hist_head = most recent */
  int contact[4];
  double last_tau[12];
  double fx1[12], fx2[12], fy1[12], fy2[12]; /* action-filter history (order 2) */
  /* snapshot for reset */
  double snap[37]; double snap_obs[ORC_HIST_W]; double snap_lam[4];
  double pos_hist[10][3];      /* stuck termination */
  double ext_force[3];         /* world-frame push at the base COM */
  double lam_lim[12];          /* warm starts of the joint-limit rows */
  int env_id;                  /* index of this env in its batch: the sensor-noise counter */
  double hyb[4][12];           /* HYBRID command of the current control step: kp | qd* | kd | tau_ff (laikago_motor.py:152-161) */
} OrcEnv;

void orc_default_config(OrcConfig* c);
void orc_default_param(double* p48);
/* pieces (each cites the reference line it restates in the .c file) */
void orc_etg_features(const OrcConfig* c, double t, double* r20);
void orc_etg_act(const OrcConfig* c, const double w[3][ORC_ETG_H], const double b[3], double t, double* act12, double* foot12);
void orc_ik_leg(const double foot[3], int l_hip_sign, double ang[3]);
void orc_fk_leg(const double ang[3], int l_hip_sign, double foot[3]);
void orc_leg_jacobian(const double ang[3], int leg_id, double J[9]);
void orc_motor_torque(const double* kp, const double* kd, const double* target, const double* q, const double* qd, double limit, double* tau);
void orc_quat_to_rpy(const double q[4], double rpy[3]);
void orc_butter2(double highcut, double fs, double b[3], double a[3]);   /* scipy.signal.butter(2, highcut/(fs/2)) closed form */
void orc_filter_step(const double b[3], const double a[3], double x, double* x1, double* x2, double* y1, double* y2, double* y);
/* dynamics */
void orc_forward_dynamics(const OrcConfig* c, const OrcEnv* e, const double tau[12], double qdd[12], double wdot_w[3], double vdot_w[3]);
void orc_mass_matrix(const OrcConfig* c, const OrcEnv* e, double M[18*18]);   /* via unit-response of ABA; for tests */
double orc_energy(const OrcConfig* c, const OrcEnv* e, double* kinetic, double* potential);
void orc_foot_world(const OrcEnv* e, double feet[4][3]);
/* env */
void orc_env_init(const OrcConfig* c, OrcEnv* e, const double* param48);
void orc_env_settle(const OrcConfig* c, OrcEnv* e);     /* reset pose + settle + snapshot */
void orc_env_reset(const OrcConfig* c, OrcEnv* e, const double* etg_w /*3x20 or NULL*/, const double* etg_b, double* obs);
void orc_env_reset_ex(const OrcConfig* c, OrcEnv* e, const double* etg_w, const double* etg_b, double x_offset, double* obs);
int orc_obs_dim(const OrcConfig* c);
void orc_normal4(unsigned long long seed, unsigned c0, unsigned c1, unsigned c2, double n[4]);   /* Philox4x32-10 + Box-Muller */
void orc_substep(const OrcConfig* c, OrcEnv* e, const double target[12]);
/* action: 12 values, or 60 = [12 motors][q*, kp, qd*, kd, tau_ff] when c->motor_mode == 2 (HYBRID) */
void orc_env_step(const OrcConfig* c, OrcEnv* e, const double* action, int donef,
                  double* obs, double* reward, int* done, double* info);
/* batch helpers (cpu baseline): nthreads pthreads over envs */
void orc_batch_step(const OrcConfig* c, OrcEnv* envs, int n, const double* actions, int donef, int auto_reset,
                    double* obs, double* reward, int* done, double* info, int nthreads);
void orc_batch_rollout(const OrcConfig* c, OrcEnv* envs, int n, const double* actions, int K, int auto_reset,
                       double* obs, double* ret, int* ndone, int nthreads);
int orc_sizeof_env(void);
#ifdef __cplusplus
}
#endif
#endif
