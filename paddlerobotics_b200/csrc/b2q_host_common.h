// b2q_host_common.h — host-side logic shared by the CUDA C-ABI (b2q_api.cu) and the CPU SIMT-emulation harness
// (tests/emu): config -> kernel Cfg conversion and the per-env repack helpers (also usable as device code).
#pragma once
#include "../../include/b2q.h"
#include "b2q_sim.cuh"
#include "b2q_model_host.h"

namespace b2q {

template <typename T>
inline Cfg<T> make_cfg(const B2QConfig& c, const T* hf_dev) {
  Cfg<T> k;
  k.dt = (T)c.sim_dt; k.R = c.action_repeat; k.iters = c.solver_iters; k.erp = (T)c.erp; k.warm = (T)c.warmstart; k.margin = (T)c.contact_margin;
  k.interp = c.action_interp; k.tau_limit = (T)c.torque_limit; k.settle_steps = c.settle_steps;
  k.filter = c.action_filter; k.etg = c.etg_enabled; k.max_steps = c.max_episode_steps;
  k.clip_cmd = c.clip_motor_commands; k.max_dq = (T)c.max_angle_change;
  k.noise_on = 0; for (int i = 0; i < 5; i++) { k.noise[i] = (T)c.noise_stdev[i]; if (c.noise_stdev[i] > 0) k.noise_on = 1; }
  k.noise_seed = c.noise_seed; k.stuck = c.stuck_termination; k.body_coll = c.body_collisions;
  k.motor_mode = c.motor_mode; k.jlim = c.joint_limits; k.extf = c.external_force; k.knee = c.knee_contacts;
  for (int i = 0; i < 4; i++) k.damp[i] = (T)c.base_damping[i];
  {  // scipy.signal.butter(2, highcut / (fs/2)) in closed form (bilinear transform), fs = 1 / control period
    const double PI = 3.14159265358979323846, fs = 1.0 / (c.sim_dt * c.action_repeat);
    const double K = std::tan(PI * c.filter_highcut / fs), n = 1.0 / (1.0 + std::sqrt(2.0) * K + K * K);
    k.fb0 = (T)(K * K * n); k.fb1 = (T)(2.0 * K * K * n); k.fb2 = (T)(K * K * n);
    k.fa1 = (T)(2.0 * (K * K - 1.0) * n); k.fa2 = (T)((1.0 - std::sqrt(2.0) * K + K * K) * n);
  }
  k.etg_T = (T)c.etg_T; k.etg_T2 = (T)c.etg_T2; k.etg_sigma_sq = (T)c.etg_sigma_sq; k.etg_amp = (T)c.etg_amp; k.etg_ph0 = (T)c.etg_phase0; k.etg_ph1 = (T)c.etg_phase1;
  k.w_torso = (T)c.w_torso; k.w_feet = (T)c.w_feet; k.w_up = (T)c.w_up; k.w_tau = (T)c.w_tau; k.w_stand = (T)c.w_stand; k.w_badfoot = (T)c.w_badfoot;
  k.w_footcontact = (T)c.w_footcontact; k.w_done = (T)c.w_done; k.reward_p = (T)c.reward_p; k.vel_d = (T)c.vel_d;
  k.terrain = c.terrain_type; k.hf_nx = c.hf_nx; k.hf_ny = c.hf_ny; k.hf_x0 = (T)c.hf_x0; k.hf_y0 = (T)c.hf_y0; k.hf_cell = (T)c.hf_cell; k.hf_icell = (T)(c.hf_cell > 0 ? 1.0 / c.hf_cell : 0.0); k.idt = (T)(1.0 / c.sim_dt); k.hf = hf_dev;
  return k;
}

inline void default_config(B2QConfig* c) {
  std::memset(c, 0, sizeof(*c));
  c->num_envs = 1; c->device = 0; c->precision = 0; c->threads_per_block = 0;
  c->sim_dt = 0.002; c->action_repeat = 13; c->solver_iters = 23; c->erp = 0.2; c->warmstart = 0.85; c->contact_margin = 0.02;
  c->action_interp = 0; c->torque_limit = 0; c->settle_steps = 500; c->action_filter = 0; c->filter_highcut = 4.0; c->etg_enabled = 1;
  c->etg_T = 0.5; c->etg_T2 = 0.5; c->etg_sigma_sq = 0.04; c->etg_amp = 0.2; c->etg_phase0 = -3.14159265358979323846 / 2; c->etg_phase1 = 0;
  c->w_torso = 1.5; c->w_feet = 0.3; c->w_up = 0.6; c->w_tau = 0.07; c->w_stand = 0; c->w_badfoot = 0.1; c->w_footcontact = 0.1; c->w_done = 1;
  c->reward_p = 5; c->vel_d = 0.5; c->foot_radius = 0.02; c->ring_depth = 4; c->auto_reset = 0; c->terrain_type = 0; c->clip_motor_commands = 0; c->max_angle_change = 0.2;
  c->sensor_dis = 1; c->sensor_contact = 1; c->sensor_imu = 1; c->sensor_motor = 1; c->sensor_etg = 1; c->obs_normal = 1;   // train.py:494-500 defaults
}

// which kernel instantiation a config needs: 0 = the lean default body, 1 = the variant with TORQUE mode / joint-limit rows / base push / damping
inline int config_feat(const B2QConfig& c) {
  return (c.motor_mode || c.joint_limits || c.external_force || c.knee_contacts || c.base_damping[0] != 0 || c.base_damping[1] != 0 || c.base_damping[2] != 0 || c.base_damping[3] != 0) ? 1 : 0;
}
// observation width selected by the sensor flags (SimpleEnv.get_observation, deployment/envs/EnvWrapper.py:60-109)
inline int config_obs_dim(const B2QConfig& c) {
  return (c.sensor_dis ? 3 : 0) + (c.sensor_contact ? 4 : 0) + (c.sensor_imu == 1 ? 6 : c.sensor_imu == 2 ? 3 : 0) +
         (c.sensor_motor == 1 ? 24 : c.sensor_motor == 2 ? 12 : 0) + (c.sensor_etg ? 12 : 0);
}
template <typename T>
inline void build_obs_map(Model<T>& M, const B2QConfig& c) {
  int n = 0;
  auto put = [&](int src, double scale, double shift) { M.obs_src[n] = src; M.obs_scale[n] = (T)scale; M.obs_shift[n] = (T)shift; n++; };
  const bool nrm = c.obs_normal != 0;
  if (c.sensor_dis) for (int i = 0; i < 3; i++) put(i, 1, 0);                                          // BaseDisplacement
  if (c.sensor_contact) for (int i = 0; i < 4; i++) put(3 + i, 1, 0);                                  // FootContactSensor
  if (c.sensor_imu == 1) for (int i = 0; i < 3; i++) put(7 + i, nrm ? 1 : 0.1, 0);                     // IMU: rpy (/0.1) ...
  if (c.sensor_imu) for (int i = 0; i < 3; i++) put(10 + i, nrm ? 1 : 0.5, 0);                         // ... drpy (/0.5)
  if (c.sensor_motor) for (int i = 0; i < 12; i++) put(13 + i, nrm ? 1 : 0.1, nrm ? 0 : (double)M.pose_ori[i % 3]);   // MotorAngle
  if (c.sensor_motor == 1) for (int i = 0; i < 12; i++) put(25 + i, 1, 0);                             // + velocities (MotorAngleAcc)
  if (c.sensor_etg) for (int i = 0; i < 12; i++) put(37 + i, nrm ? 1 : (double)M.etg_std[i], nrm ? 0 : (double)M.etg_mean[i]);
  M.obs_dim = n;
  M.obs_identity = (n == OBS_DIM && nrm) ? 1 : 0;
  for (; n < OBS_DIM; n++) { M.obs_src[n] = 0; M.obs_scale[n] = 0; M.obs_shift[n] = 0; }
}

inline const char* validate_config(const B2QConfig& c) {
  if (c.num_envs < 1) return "num_envs must be >= 1";
  if (c.precision != 0 && c.precision != 1) return "precision must be 0 (f32) or 1 (f64)";
  if (c.action_repeat < 2 || c.action_repeat > 64) return "action_repeat out of range [2,64]";
  if (c.solver_iters < 1 || c.solver_iters > 1000) return "solver_iters out of range";
  if (!(c.sim_dt > 0)) return "sim_dt must be > 0";
  if (c.action_filter && !(c.filter_highcut > 0)) return "filter_highcut must be > 0";
  if (c.clip_motor_commands && !(c.max_angle_change > 0)) return "max_angle_change must be > 0";
  if (c.ring_depth < 1 || c.ring_depth > 16) return "ring_depth out of range [1,16]";
  if (c.terrain_type == 1 && (c.hf_nx < 2 || c.hf_ny < 2 || !c.hf_host || !(c.hf_cell > 0))) return "height field needs hf_nx,hf_ny>=2, hf_cell>0 and hf_host";
  if (c.terrain_type != 0 && c.terrain_type != 1) return "terrain_type must be 0 or 1";
  if (c.sensor_imu < 0 || c.sensor_imu > 2 || c.sensor_motor < 0 || c.sensor_motor > 2) return "sensor_imu / sensor_motor must be 0, 1 or 2";
  if (config_obs_dim(c) < 1) return "sensor flags select an empty observation";
  if (c.motor_mode < 0 || c.motor_mode > 2) return "motor_mode must be 0 (POSITION), 1 (TORQUE) or 2 (HYBRID)";
  for (int i = 0; i < 5; i++) if (!(c.noise_stdev[i] >= 0)) return "noise_stdev must be >= 0";
  for (int i = 0; i < 4; i++) if (!(c.base_damping[i] >= 0)) return "base_damping must be >= 0";
  if (c.threads_per_block != 0 && (c.threads_per_block % 32 != 0 || c.threads_per_block > 128)) return "threads_per_block must be a multiple of 32, at most 128 (kernels are __launch_bounds__(128))";
  return nullptr;
}

// dyn row [48] -> param packs (one thread per env)
template <typename T>
B2Q_HD void pack_param_env(const T* dyn /*[N][48] or null*/, const T* def48, P4<T>* param, int N, int env) {
  const T* p = dyn ? dyn + (size_t)env * 48 : def48;
  for (int k = 0; k < 4; k++) {
    stp(param, 0 + k, N, env, p[3 * k], p[3 * k + 1], p[3 * k + 2], p[24]);
    stp(param, 4 + k, N, env, p[12 + 3 * k], p[12 + 3 * k + 1], p[12 + 3 * k + 2], p[25]);
    stp(param, 8 + k, N, env, p[36 + 3 * k], p[36 + 3 * k + 1], p[36 + 3 * k + 2], T(0));
  }
  stp(param, 12, N, env, p[33], p[34], p[35], p[29]);
  stp(param, 13, N, env, p[30], p[31], p[32], T(0));
  stp(param, 14, N, env, p[26], p[27], p[28], T(0));
}
// etg_w [N][3][20], etg_b [N][3] -> ETG packs
template <typename T>
B2Q_HD void pack_etg_env(const T* w, const T* b, P4<T>* etg, int N, int env) {
  T* e = reinterpret_cast<T*>(etg);
  if (w) for (int i = 0; i < 60; i++) e[((size_t)(i >> 2) * N + env) * 4 + (i & 3)] = w[(size_t)env * 60 + i];
  if (b) for (int i = 0; i < 3; i++) { int idx = 60 + i; e[((size_t)(idx >> 2) * N + env) * 4 + (idx & 3)] = b[(size_t)env * 3 + i]; }
}
// state packs <-> [N][37] rows
template <typename T>
B2Q_HD void get_state_env(const P4<T>* st, T* out, int N, int env) {
  T* o = out + (size_t)env * 37;
  P4<T> a = ldp(st, 0, N, env), b = ldp(st, 1, N, env), c = ldp(st, 2, N, env), d = ldp(st, 3, N, env);
  o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = b.x; o[4] = b.y; o[5] = b.z; o[6] = b.w; o[7] = c.x; o[8] = c.y; o[9] = c.z; o[10] = d.x; o[11] = d.y; o[12] = d.z;
  for (int k = 0; k < 4; k++) {
    P4<T> q = ldp(st, 4 + k, N, env), qd = ldp(st, 8 + k, N, env);
    o[13 + 3 * k] = q.x; o[14 + 3 * k] = q.y; o[15 + 3 * k] = q.z; o[25 + 3 * k] = qd.x; o[26 + 3 * k] = qd.y; o[27 + 3 * k] = qd.z;
  }
}
template <typename T>
B2Q_HD void set_state_env(P4<T>* st, const T* in, int N, int env) {
  const T* o = in + (size_t)env * 37;
  P4<T> c = ldp(st, 2, N, env);
  stp(st, 0, N, env, o[0], o[1], o[2], T(0)); stp(st, 1, N, env, o[3], o[4], o[5], o[6]);
  stp(st, 2, N, env, o[7], o[8], o[9], c.w); stp(st, 3, N, env, o[10], o[11], o[12], T(0));
  for (int k = 0; k < 4; k++) {
    P4<T> q = ldp(st, 4 + k, N, env), qd = ldp(st, 8 + k, N, env);
    stp(st, 4 + k, N, env, o[13 + 3 * k], o[14 + 3 * k], o[15 + 3 * k], q.w);
    stp(st, 8 + k, N, env, o[25 + 3 * k], o[26 + 3 * k], o[27 + 3 * k], qd.w);
  }
}

}  // namespace b2q
