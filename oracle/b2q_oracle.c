/* b2q_oracle.c — CPU float64 oracle (TEST INFRASTRUCTURE; see b2q_oracle.h header).
 *
 * Restates, in plain C, the per-step hot path of PaddleRobotics QuadrupedalRobots/ETGRL:
 *   ETG spline + IK      rlschool ETG_layer/ETG_model (absent; restated and verified against the
 *                        reference's golden .npy) + a1.py:97-110,464-497
 *   FK / Jacobian        a1.py:113-173
 *   action interpolation minitaur.py:1384-1401
 *   PD motor model       laikago_motor.py:103-175, minitaur.py:904-947
 *   stepSimulation       Bullet btMultiBody (third-party, absent): Featherstone ABA in link
 *                        coordinates + PGS contact rows in DoF space  — PARITY UNPINNED
 *   observation history  minitaur.py:1142-1193
 *   observation packing  EnvWrapper.py:50-109
 * The algorithmic formulation here (ABA recursion, DoF-space PGS with unit-impulse responses) is
 * deliberately different from the CUDA product path (composite-inertia + Schur complement,
 * contact-space PGS), so agreement between the two is evidence for both.
 */
#include "b2q_oracle.h"
#include <math.h>
#include <string.h>
#include <stdlib.h>
#include <pthread.h>

/* ------------------------------------------------------------------ constants */
static const double POSE_ORI[12] = {0, 0.9, -1.8, 0, 0.9, -1.8, 0, 0.9, -1.8, 0, 0.9, -1.8}; /* a1.py:83 */
static const double COM_OFFSET[3] = {-0.012731, -0.002186, -0.000515};                       /* a1.py:70 */
static const double HIP_XY[4][2] = {{0.183, -0.047}, {0.183, 0.047}, {-0.183, -0.047}, {-0.183, 0.047}}; /* a1.py:71-72 */
static const double BASE_FOOT[4][3] = {{0.18, -0.15, -0.23}, {0.18, 0.148, -0.23}, {-0.18, -0.14, -0.23}, {-0.18, 0.135, -0.23}};
static const double ETG_MEAN[12] = {2.1505982e-02, 3.6674485e-02, -6.0444288e-02, 2.4625482e-02, 1.5869144e-02, -3.2513142e-02,
                                    2.1506395e-02, 3.1869926e-02, -6.0140789e-02, 2.4625063e-02, 1.1628972e-02, -3.2163858e-02}; /* EnvWrapper.py:50-53 */
static const double ETG_STD[12] = {4.5967497e-02, 2.0340437e-01, 3.7410179e-01, 4.6187632e-02, 1.9441207e-01, 3.9488649e-01,
                                   4.5966785e-02, 2.0323379e-01, 3.7382501e-01, 4.6188373e-02, 1.9457331e-01, 3.9302582e-01}; /* EnvWrapper.py:54-55 */
#define L_UP 0.2
#define L_LOW 0.2
#define L_HIP 0.08505

/* A1 URDF inertials (pybullet_data a1/a1.urdf = unitree a1_description const.xacro; recalled,
 * SURVEY App. B.3, UNVERIFIED against the file which is absent). */
static const double TRUNK_M = 4.713;
static const double TRUNK_I[6] = {0.01683993, 8.3902e-05, 0.000597679, 0.056579028, 2.5134e-05, 0.064713601}; /* xx xy xz yy yz zz */
static const double HIP_M = 0.696, HIP_C[3] = {-0.003311, 0.000635, 3.1e-05};
static const double HIP_I[6] = {0.000469246, -9.409e-06, -3.42e-07, 0.00080749, -4.66e-07, 0.000552929};
static const double THIGH_M = 1.013, THIGH_C[3] = {-0.003237, -0.022327, -0.027326};
static const double THIGH_I[6] = {0.005529065, 4.825e-06, 0.000343869, 0.005139339, 2.2448e-05, 0.001367788};
static const double CALF_M = 0.166, CALF_C[3] = {0.006435, 0.0, -0.107388};
static const double CALF_I[6] = {0.002997972, 0.0, -0.000141163, 0.003014022, 0.0, 3.2426e-05};
static const double TOE_M = 0.06, TOE_I = 9.6e-06;

/* ------------------------------------------------------------------ small math */
static void v3cross(const double a[3], const double b[3], double o[3]) {
  double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  o[0] = x; o[1] = y; o[2] = z;
}
static double v3dot(const double a[3], const double b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static void m3v(const double M[9], const double v[3], double o[3]) {
  double x = M[0] * v[0] + M[1] * v[1] + M[2] * v[2], y = M[3] * v[0] + M[4] * v[1] + M[5] * v[2], z = M[6] * v[0] + M[7] * v[1] + M[8] * v[2];
  o[0] = x; o[1] = y; o[2] = z;
}
static void m3tv(const double M[9], const double v[3], double o[3]) {
  double x = M[0] * v[0] + M[3] * v[1] + M[6] * v[2], y = M[1] * v[0] + M[4] * v[1] + M[7] * v[2], z = M[2] * v[0] + M[5] * v[1] + M[8] * v[2];
  o[0] = x; o[1] = y; o[2] = z;
}
static void m3m(const double A[9], const double B[9], double O[9]) {
  double T[9];
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { double s = 0; for (int k = 0; k < 3; k++) s += A[3 * i + k] * B[3 * k + j]; T[3 * i + j] = s; }
  memcpy(O, T, sizeof T);
}
static void quat_to_mat(const double q[4], double R[9]) { /* xyzw, body->world */
  double x = q[0], y = q[1], z = q[2], w = q[3];
  R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - z * w); R[2] = 2 * (x * z + y * w);
  R[3] = 2 * (x * y + z * w); R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - x * w);
  R[6] = 2 * (x * z - y * w); R[7] = 2 * (y * z + x * w); R[8] = 1 - 2 * (x * x + y * y);
}
void orc_quat_to_rpy(const double q[4], double rpy[3]) { /* pybullet getEulerFromQuaternion, minitaur.py:620 */
  double x = q[0], y = q[1], z = q[2], w = q[3];
  rpy[0] = atan2(2 * (w * x + y * z), 1 - 2 * (x * x + y * y));
  double s = 2 * (w * y - z * x); if (s > 1) s = 1; if (s < -1) s = -1;
  rpy[1] = asin(s);
  rpy[2] = atan2(2 * (w * z + x * y), 1 - 2 * (y * y + z * z));
}
static void rot_axis(int axis, double q, double R[9]) { /* link->parent rotation */
  double c = cos(q), s = sin(q);
  if (axis == 0) { double T[9] = {1, 0, 0, 0, c, -s, 0, s, c}; memcpy(R, T, sizeof T); }
  else { double T[9] = {c, 0, s, 0, 1, 0, -s, 0, c}; memcpy(R, T, sizeof T); }
}
static void sym6_to_m3(const double s[6], double M[9]) { M[0] = s[0]; M[1] = s[1]; M[2] = s[2]; M[3] = s[1]; M[4] = s[3]; M[5] = s[4]; M[6] = s[2]; M[7] = s[4]; M[8] = s[5]; }
static double map_pi(double a) { /* MapToMinusPiToPi minitaur.py:67-83 */
  double m = fmod(a, 2 * M_PI);
  if (m >= M_PI) m -= 2 * M_PI; else if (m < -M_PI) m += 2 * M_PI;
  return m;
}

/* ------------------------------------------------------------------ config */
void orc_default_config(OrcConfig* c) {
  memset(c, 0, sizeof *c);
  c->sim_dt = 0.002; c->action_repeat = 13; c->solver_iters = 23; c->erp = 0.2; c->warmstart = 0.85; c->contact_margin = 0.02;
  c->action_interp = 0; c->torque_limit = 0; c->settle_steps = 500; c->action_filter = 0; c->filter_highcut = 4.0; c->etg_enabled = 1;
  c->clip_motor_commands = 0; c->max_angle_change = 0.2;
  c->etg_T = 0.5; c->etg_T2 = 0.5; c->etg_sigma_sq = 0.04; c->etg_amp = 0.2; c->etg_phase[0] = -M_PI / 2; c->etg_phase[1] = 0; /* train.py:296-297 */
  c->w_torso = 1.5; c->w_feet = 0.3; c->w_up = 0.6; c->w_tau = 0.07; c->w_stand = 0; c->w_badfoot = 0.1; c->w_footcontact = 0.1; c->w_done = 1; /* train.py:478-484 */
  c->reward_p = 5; c->vel_d = 0.5; c->foot_radius = 0.02; c->terrain_type = 0;
  c->sensor_dis = 1; c->sensor_contact = 1; c->sensor_imu = 1; c->sensor_motor = 1; c->sensor_etg = 1; c->obs_normal = 1; /* train.py:494-500 */
}
int orc_obs_dim(const OrcConfig* c) {
  return (c->sensor_dis ? 3 : 0) + (c->sensor_contact ? 4 : 0) + (c->sensor_imu == 1 ? 6 : c->sensor_imu == 2 ? 3 : 0) +
         (c->sensor_motor == 1 ? 24 : c->sensor_motor == 2 ? 12 : 0) + (c->sensor_etg ? 12 : 0);
}
/* counter-based Gaussian shared with the device code (integer part bit-identical): Philox4x32-10, Box-Muller */
void orc_normal4(unsigned long long seed, unsigned c0, unsigned c1, unsigned c2, double n[4]) {
  unsigned c3 = 0, k0 = (unsigned)seed, k1 = (unsigned)(seed >> 32);
  for (int i = 0; i < 10; i++) {
    unsigned long long p0 = 0xD2511F53ull * c0, p1 = 0xCD9E8D57ull * c2;
    unsigned h0 = (unsigned)(p0 >> 32), l0 = (unsigned)p0, h1 = (unsigned)(p1 >> 32), l1 = (unsigned)p1;
    unsigned n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
    c0 = n0; c1 = l1; c2 = n2; c3 = l0; k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  const double sc = 1.0 / 16777216.0;
  double u0 = ((double)(c0 >> 8) + 0.5) * sc, u1 = ((double)(c1 >> 8) + 0.5) * sc, u2 = ((double)(c2 >> 8) + 0.5) * sc, u3 = ((double)(c3 >> 8) + 0.5) * sc;
  double ra = sqrt(-2.0 * log(u0)), rb = sqrt(-2.0 * log(u2));
  n[0] = ra * cos(2 * M_PI * u1); n[1] = ra * sin(2 * M_PI * u1); n[2] = rb * cos(2 * M_PI * u3); n[3] = rb * sin(2 * M_PI * u3);
}
void orc_default_param(double* p) {
  for (int i = 0; i < 12; i++) { p[i] = 100.0; p[12 + i] = (i % 3 == 0) ? 1.0 : 2.0; } /* a1.py:75-80 */
  p[24] = 1.0; p[25] = 0.002; /* a1.py:233 */
  p[26] = 0; p[27] = 0; p[28] = -10.0; /* train.py:125 */
  for (int i = 29; i < 48; i++) p[i] = 1.0;
}

/* ------------------------------------------------------------------ ETG (SURVEY App. A restatement) */
static void etg_forward(const OrcConfig* c, double t, double o[2]) {
  double om = 2 * M_PI / c->etg_T;
  o[0] = c->etg_amp * sin(c->etg_phase[0] + om * t);
  o[1] = c->etg_amp * sin(c->etg_phase[1] + om * t);
}
void orc_etg_features(const OrcConfig* c, double t, double* r) {
  double x[2], u[2];
  etg_forward(c, t, x);
  for (int h = 0; h < ORC_ETG_H; h++) {
    etg_forward(c, h * c->etg_T / (ORC_ETG_H - 0.9), u); /* centres: note the H-0.9 denominator (App. A) */
    double dx = x[0] - u[0], dy = x[1] - u[1];
    r[h] = exp(-(dx * dx + dy * dy) / c->etg_sigma_sq);
  }
}
void orc_ik_leg(const double foot[3], int l_hip_sign, double ang[3]) { /* a1.py:97-110 */
  double l_hip = L_HIP * l_hip_sign, x = foot[0], y = foot[1], z = foot[2];
  double tk = -acos((x * x + y * y + z * z - l_hip * l_hip - L_LOW * L_LOW - L_UP * L_UP) / (2 * L_LOW * L_UP));
  double l = sqrt(L_UP * L_UP + L_LOW * L_LOW + 2 * L_UP * L_LOW * cos(tk));
  double th = asin(-x / l) - tk / 2;
  double c1 = l_hip * y - l * cos(th + tk / 2) * z;
  double s1 = l * cos(th + tk / 2) * y + l_hip * z;
  ang[0] = atan2(s1, c1); ang[1] = th; ang[2] = tk;
}
void orc_fk_leg(const double a[3], int l_hip_sign, double foot[3]) { /* a1.py:113-129 */
  double l_hip = L_HIP * l_hip_sign;
  double ld = sqrt(L_UP * L_UP + L_LOW * L_LOW + 2 * L_UP * L_LOW * cos(a[2]));
  double eff = a[1] + a[2] / 2;
  double ox = -ld * sin(eff), oz = -ld * cos(eff), oy = l_hip;
  foot[0] = ox; foot[1] = cos(a[0]) * oy - sin(a[0]) * oz; foot[2] = sin(a[0]) * oy + cos(a[0]) * oz;
}
void orc_leg_jacobian(const double a[3], int leg_id, double J[9]) { /* a1.py:132-159 */
  double l_hip = L_HIP * ((leg_id % 2) ? 1.0 : -1.0);
  double t1 = a[0], t2 = a[1], t3 = a[2];
  double le = sqrt(L_UP * L_UP + L_LOW * L_LOW + 2 * L_UP * L_LOW * cos(t3)), te = t2 + t3 / 2;
  J[0] = 0; J[1] = -le * cos(te); J[2] = L_LOW * L_UP * sin(t3) * sin(te) / le - le * cos(te) / 2;
  J[3] = -l_hip * sin(t1) + le * cos(t1) * cos(te); J[4] = -le * sin(t1) * sin(te);
  J[5] = -L_LOW * L_UP * sin(t1) * sin(t3) * cos(te) / le - le * sin(t1) * sin(te) / 2;
  J[6] = l_hip * cos(t1) + le * sin(t1) * cos(te); J[7] = le * sin(te) * cos(t1);
  J[8] = L_LOW * L_UP * sin(t3) * cos(t1) * cos(te) / le + le * sin(te) * cos(t1) / 2;
}
void orc_etg_act(const OrcConfig* c, const double w[3][ORC_ETG_H], const double b[3], double t, double* act, double* foot_out) {
  if (!c->etg_enabled) { for (int j = 0; j < 12; j++) { act[j] = 0; if (foot_out) foot_out[j] = 0; } return; }
  double r1[ORC_ETG_H], r2[ORC_ETG_H];
  orc_etg_features(c, t, r1);
  orc_etg_features(c, t + 0.5 * c->etg_T2, r2);
  for (int leg = 0; leg < 4; leg++) {
    const double* r = (leg == 0 || leg == 3) ? r1 : r2; /* FR,RL in phase; FL,RR half a period later (trot) */
    double d[3];
    for (int a = 0; a < 3; a++) { double s = b[a]; for (int h = 0; h < ORC_ETG_H; h++) s += w[a][h] * r[h]; d[a] = s; }
    double ang[3];
    for (int tries = 0; tries < 200; tries++) { /* act_clip: shrink delta until IK is finite [EXT] */
      double f[3];
      for (int a = 0; a < 3; a++) f[a] = BASE_FOOT[leg][a] - (a == 1 ? (BASE_FOOT[leg][1] > 0 ? c->etg_foot_y_inset : -c->etg_foot_y_inset) : 0.0) + d[a] - ((a < 2 ? HIP_XY[leg][a] : 0.0) + COM_OFFSET[a]);
      orc_ik_leg(f, (leg % 2) ? 1 : -1, ang);
      if (!(isnan(ang[0]) || isnan(ang[1]) || isnan(ang[2]))) break;
      for (int a = 0; a < 3; a++) d[a] *= 0.95;
    }
    for (int a = 0; a < 3; a++) {
      act[3 * leg + a] = ang[a] - POSE_ORI[3 * leg + a];
      if (foot_out) foot_out[3 * leg + a] = BASE_FOOT[leg][a] + d[a];
    }
  }
}
void orc_motor_torque(const double* kp, const double* kd, const double* target, const double* q, const double* qd, double limit, double* tau) {
  for (int j = 0; j < 12; j++) { /* laikago_motor.py:165-173, POSITION mode, strength ratio 1 */
    double t = -1 * (kp[j] * (q[j] - target[j])) - kd[j] * (qd[j] - 0.0);
    if (limit > 0) { if (t > limit) t = limit; if (t < -limit) t = -limit; }
    tau[j] = t;
  }
}

/* 2nd-order Butterworth low-pass through the bilinear transform (what scipy.signal.butter(2, Wn) returns), and one step
 * of ActionFilter.filter (action_filter.py:111-120): y = b0 x + b1 x1 + b2 x2 - a1 y1 - a2 y2, then shift the histories */
void orc_butter2(double highcut, double fs, double b[3], double a[3]) {
  double K = tan(M_PI * highcut / fs), n = 1.0 / (1.0 + sqrt(2.0) * K + K * K);
  b[0] = K * K * n; b[1] = 2 * b[0]; b[2] = b[0];
  a[0] = 1.0; a[1] = 2.0 * (K * K - 1.0) * n; a[2] = (1.0 - sqrt(2.0) * K + K * K) * n;
}
void orc_filter_step(const double b[3], const double a[3], double x, double* x1, double* x2, double* y1, double* y2, double* y) {
  double yy = b[0] * x + b[1] * *x1 + b[2] * *x2 - a[1] * *y1 - a[2] * *y2;
  *x2 = *x1; *x1 = x; *y2 = *y1; *y1 = yy; *y = yy;
}

/* ------------------------------------------------------------------ spatial algebra ([ang;lin]) */
typedef struct { double E[9]; double r[3]; } Xf; /* v_child = E (v_parent_lin + w x r ...) */
static void xf_motion(const Xf* X, const double v[6], double o[6]) {
  double w[3], l[3], t[3];
  m3v(X->E, v, w);
  v3cross(X->r, v, t); /* r x w */
  double vl[3] = {v[3] - t[0], v[4] - t[1], v[5] - t[2]};
  m3v(X->E, vl, l);
  o[0] = w[0]; o[1] = w[1]; o[2] = w[2]; o[3] = l[0]; o[4] = l[1]; o[5] = l[2];
}
static void xf_force_T(const Xf* X, const double f[6], double o[6]) { /* child-frame force -> parent frame */
  double n[3], fl[3], t[3];
  m3tv(X->E, f, n); m3tv(X->E, f + 3, fl);
  v3cross(X->r, fl, t);
  o[0] = n[0] + t[0]; o[1] = n[1] + t[1]; o[2] = n[2] + t[2]; o[3] = fl[0]; o[4] = fl[1]; o[5] = fl[2];
}
static void xf_dense(const Xf* X, double M[36]) {
  /* [[E,0],[-E rx, E]] */
  double rx[9] = {0, -X->r[2], X->r[1], X->r[2], 0, -X->r[0], -X->r[1], X->r[0], 0}, Erx[9];
  m3m(X->E, rx, Erx);
  memset(M, 0, 36 * sizeof(double));
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { M[6 * i + j] = X->E[3 * i + j]; M[6 * (i + 3) + j + 3] = X->E[3 * i + j]; M[6 * (i + 3) + j] = -Erx[3 * i + j]; }
}
static void crm(const double v[6], const double u[6], double o[6]) { /* v x u (motion) */
  double a[3], b[3], c[3];
  v3cross(v, u, a); v3cross(v, u + 3, b); v3cross(v + 3, u, c);
  o[0] = a[0]; o[1] = a[1]; o[2] = a[2]; o[3] = b[0] + c[0]; o[4] = b[1] + c[1]; o[5] = b[2] + c[2];
}
static void crf(const double v[6], const double f[6], double o[6]) { /* v x* f */
  double a[3], b[3], c[3];
  v3cross(v, f, a); v3cross(v + 3, f + 3, b); v3cross(v, f + 3, c);
  o[0] = a[0] + b[0]; o[1] = a[1] + b[1]; o[2] = a[2] + b[2]; o[3] = c[0]; o[4] = c[1]; o[5] = c[2];
}
static void m6v(const double M[36], const double v[6], double o[6]) {
  double t[6];
  for (int i = 0; i < 6; i++) { double s = 0; for (int j = 0; j < 6; j++) s += M[6 * i + j] * v[j]; t[i] = s; }
  memcpy(o, t, sizeof t);
}
static void spatial_inertia(double m, const double c[3], const double Ic[9], double I[36]) {
  double cx[9] = {0, -c[2], c[1], c[2], 0, -c[0], -c[1], c[0], 0}, cc[9];
  m3m(cx, cx, cc);
  memset(I, 0, 36 * sizeof(double));
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
    I[6 * i + j] = Ic[3 * i + j] - m * cc[3 * i + j];
    I[6 * i + j + 3] = m * cx[3 * i + j];
    I[6 * (i + 3) + j] = -m * cx[3 * i + j];
  }
  I[6 * 3 + 3] = I[6 * 4 + 4] = I[6 * 5 + 5] = m;
}
static int solve6(const double A[36], const double b[6], double x[6]) { /* Gaussian elimination w/ partial pivoting */
  double M[6][7];
  for (int i = 0; i < 6; i++) { for (int j = 0; j < 6; j++) M[i][j] = A[6 * i + j]; M[i][6] = b[i]; }
  for (int k = 0; k < 6; k++) {
    int p = k; for (int i = k + 1; i < 6; i++) if (fabs(M[i][k]) > fabs(M[p][k])) p = i;
    if (p != k) for (int j = 0; j < 7; j++) { double t = M[k][j]; M[k][j] = M[p][j]; M[p][j] = t; }
    if (M[k][k] == 0) return -1;
    for (int i = k + 1; i < 6; i++) { double f = M[i][k] / M[k][k]; for (int j = k; j < 7; j++) M[i][j] -= f * M[k][j]; }
  }
  for (int i = 5; i >= 0; i--) { double s = M[i][6]; for (int j = i + 1; j < 6; j++) s -= M[i][j] * x[j]; x[i] = s / M[i][i]; }
  return 0;
}

/* ------------------------------------------------------------------ model (per env, after scaling) */
typedef struct {
  double I0[36];          /* base spatial inertia, base frame (origin at trunk COM) */
  double Il[12][36];      /* link spatial inertias, link frames */
  double jr[12][3];       /* joint origin in parent frame */
  int axis[12];
  double mass[13]; double com[13][3]; /* for energy */
} Model;

static void build_model(const double* p, Model* M) {
  double Ic[9], z3[3] = {0, 0, 0};
  /* base: scale mass p[29]; inertia diag scale p[30..32] as sqrt(si sj) */
  sym6_to_m3(TRUNK_I, Ic);
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) Ic[3 * i + j] *= sqrt(p[30 + i] * p[30 + j]);
  M->mass[0] = TRUNK_M * p[29]; memcpy(M->com[0], z3, sizeof z3);
  spatial_inertia(M->mass[0], z3, Ic, M->I0);
  for (int leg = 0; leg < 4; leg++) {
    double mirror = (leg % 2) ? 1.0 : -1.0, fh = (leg < 2) ? 1.0 : -1.0;
    for (int l = 0; l < 3; l++) {
      int i = 3 * leg + l;
      double m, c[3], s6[6];
      if (l == 0) {
        m = HIP_M; c[0] = HIP_C[0] * fh; c[1] = HIP_C[1] * mirror; c[2] = HIP_C[2];
        memcpy(s6, HIP_I, sizeof s6); s6[1] *= mirror * fh; s6[2] *= fh; s6[4] *= mirror;
        M->jr[i][0] = HIP_XY[leg][0] + COM_OFFSET[0]; M->jr[i][1] = HIP_XY[leg][1] + COM_OFFSET[1]; M->jr[i][2] = COM_OFFSET[2];
        M->axis[i] = 0;
        sym6_to_m3(s6, Ic);
      } else if (l == 1) {
        m = THIGH_M; c[0] = THIGH_C[0]; c[1] = THIGH_C[1] * mirror; c[2] = THIGH_C[2];
        memcpy(s6, THIGH_I, sizeof s6); s6[1] *= mirror; s6[4] *= mirror;
        M->jr[i][0] = 0; M->jr[i][1] = L_HIP * mirror; M->jr[i][2] = 0;
        M->axis[i] = 1;
        sym6_to_m3(s6, Ic);
      } else {
        /* calf + fixed toe merged (merge fixed children) */
        double mt = CALF_M + TOE_M, toe[3] = {0, 0, -L_LOW};
        for (int a = 0; a < 3; a++) c[a] = (CALF_M * CALF_C[a] + TOE_M * toe[a]) / mt;
        double Ia[9]; sym6_to_m3(CALF_I, Ia);
        double It[9] = {TOE_I, 0, 0, 0, TOE_I, 0, 0, 0, TOE_I};
        for (int b = 0; b < 2; b++) {
          const double* cb = b ? toe : CALF_C; double mb = b ? TOE_M : CALF_M; double* Ib = b ? It : Ia;
          double d[3] = {cb[0] - c[0], cb[1] - c[1], cb[2] - c[2]}, dd = v3dot(d, d);
          for (int r = 0; r < 3; r++) for (int s = 0; s < 3; s++) Ib[3 * r + s] += mb * ((r == s ? dd : 0) - d[r] * d[s]);
        }
        for (int k = 0; k < 9; k++) Ic[k] = Ia[k] + It[k];
        m = mt;
        M->jr[i][0] = 0; M->jr[i][1] = 0; M->jr[i][2] = -L_UP;
        M->axis[i] = 1;
      }
      double ms = p[33 + l], is = p[36 + i];
      m *= ms; for (int k = 0; k < 9; k++) Ic[k] *= is;
      M->mass[1 + i] = m; memcpy(M->com[1 + i], c, sizeof c);
      spatial_inertia(m, c, Ic, M->Il[i]);
    }
  }
}

/* ------------------------------------------------------------------ ABA */
typedef struct {
  Model mdl;
  Xf X[12];
  double v[13][6], c[12][6], IA[13][36], pA[13][6], U[12][6], d[12], u[12], a[13][6];
  double R0[9];                  /* base->world */
  double Rw[12][9], pw[12][3];   /* link->world rotation, link origin in world */
  double damp[4], fext_b[3];     /* base damping coefficients, external push in base coordinates */
} Dyn;

static void dyn_kinematics(const OrcEnv* e, Dyn* D) {
  quat_to_mat(e->quat, D->R0);
  for (int i = 0; i < 12; i++) {
    int par = (i % 3 == 0) ? -1 : i - 1;
    double Rl[9]; rot_axis(D->mdl.axis[i], e->q[i], Rl);
    for (int r = 0; r < 3; r++) for (int s = 0; s < 3; s++) D->X[i].E[3 * r + s] = Rl[3 * s + r];
    memcpy(D->X[i].r, D->mdl.jr[i], 3 * sizeof(double));
    const double* Rp = par < 0 ? D->R0 : D->Rw[par];
    const double* pp = par < 0 ? e->pos : D->pw[par];
    double t[3]; m3v(Rp, D->mdl.jr[i], t);
    D->pw[i][0] = pp[0] + t[0]; D->pw[i][1] = pp[1] + t[1]; D->pw[i][2] = pp[2] + t[2];
    m3m(Rp, Rl, D->Rw[i]);
  }
  memset(D->damp, 0, sizeof D->damp); memset(D->fext_b, 0, sizeof D->fext_b);
}

static void dyn_aba(const OrcEnv* e, Dyn* D, const double tau[12], double qdd[12], double a0[6]) {
  /* pass 1 */
  double w[3], vl[3];
  m3tv(D->R0, e->vang, w); m3tv(D->R0, e->vlin, vl);
  for (int k = 0; k < 3; k++) { D->v[0][k] = w[k]; D->v[0][3 + k] = vl[k]; }
  memcpy(D->IA[0], D->mdl.I0, sizeof D->IA[0]);
  { double Iv[6]; m6v(D->mdl.I0, D->v[0], Iv); crf(D->v[0], Iv, D->pA[0]);
    /* Bullet base damping (force = m v (k1 + k2|v|), torque = I w (k1 + k2|w|)) and the external push, both as bias forces of the base */
    double lv = sqrt(v3dot(D->v[0] + 3, D->v[0] + 3)), lw = sqrt(v3dot(D->v[0], D->v[0]));
    for (int k = 0; k < 3; k++) {
      D->pA[0][k] += Iv[k] * (D->damp[2] + D->damp[3] * lw);
      D->pA[0][3 + k] += Iv[3 + k] * (D->damp[0] + D->damp[1] * lv);
      D->pA[0][3 + k] -= D->fext_b[k];
    }
  }
  for (int i = 0; i < 12; i++) {
    int pb = (i % 3 == 0) ? 0 : i; /* body index of parent: base=0, link j -> body j+1 */
    double vj[6] = {0, 0, 0, 0, 0, 0}; vj[D->mdl.axis[i]] = e->qd[i];
    xf_motion(&D->X[i], D->v[pb], D->v[i + 1]);
    for (int k = 0; k < 6; k++) D->v[i + 1][k] += vj[k];
    crm(D->v[i + 1], vj, D->c[i]);
    memcpy(D->IA[i + 1], D->mdl.Il[i], sizeof D->IA[0]);
    double Iv[6]; m6v(D->mdl.Il[i], D->v[i + 1], Iv); crf(D->v[i + 1], Iv, D->pA[i + 1]);
  }
  /* pass 2 */
  for (int i = 11; i >= 0; i--) {
    int pb = (i % 3 == 0) ? 0 : i, ax = D->mdl.axis[i];
    double* IA = D->IA[i + 1]; double* pA = D->pA[i + 1];
    for (int k = 0; k < 6; k++) D->U[i][k] = IA[6 * k + ax];
    D->d[i] = D->U[i][ax];
    D->u[i] = tau[i] - pA[ax];
    double Ia[36], pa[6], Iac[6];
    for (int r = 0; r < 6; r++) for (int s = 0; s < 6; s++) Ia[6 * r + s] = IA[6 * r + s] - D->U[i][r] * D->U[i][s] / D->d[i];
    m6v(Ia, D->c[i], Iac);
    for (int k = 0; k < 6; k++) pa[k] = pA[k] + Iac[k] + D->U[i][k] * D->u[i] / D->d[i];
    double Xd[36], T[36], pap[6];
    xf_dense(&D->X[i], Xd);
    for (int r = 0; r < 6; r++) for (int s = 0; s < 6; s++) { double acc = 0; for (int k = 0; k < 6; k++) acc += Ia[6 * r + k] * Xd[6 * k + s]; T[6 * r + s] = acc; }
    for (int r = 0; r < 6; r++) for (int s = 0; s < 6; s++) { double acc = 0; for (int k = 0; k < 6; k++) acc += Xd[6 * k + r] * T[6 * k + s]; D->IA[pb][6 * r + s] += acc; }
    xf_force_T(&D->X[i], pa, pap);
    for (int k = 0; k < 6; k++) D->pA[pb][k] += pap[k];
  }
  double rhs[6]; for (int k = 0; k < 6; k++) rhs[k] = -D->pA[0][k];
  solve6(D->IA[0], rhs, D->a[0]);
  /* pass 3 */
  for (int i = 0; i < 12; i++) {
    int pb = (i % 3 == 0) ? 0 : i, ax = D->mdl.axis[i];
    double ap[6]; xf_motion(&D->X[i], D->a[pb], ap);
    for (int k = 0; k < 6; k++) ap[k] += D->c[i][k];
    double s = 0; for (int k = 0; k < 6; k++) s += D->U[i][k] * ap[k];
    qdd[i] = (D->u[i] - s) / D->d[i];
    ap[ax] += qdd[i];
    memcpy(D->a[i + 1], ap, sizeof ap);
  }
  memcpy(a0, D->a[0], 6 * sizeof(double));
}

/* response of the accelerations (or velocities) to a spatial force/impulse f (link coords) applied on body
 * `body` (0 = base, i+1 = link i) and generalized joint forces tj (may be NULL): out = M^-1 [..] (18) */
static void dyn_delta(const Dyn* D, int body, const double f[6], const double* tj, double out[18]) {
  double pa[13][6]; memset(pa, 0, sizeof pa);
  double u[12];
  if (f) for (int k = 0; k < 6; k++) pa[body][k] = -f[k];
  for (int i = 11; i >= 0; i--) {
    int pb = (i % 3 == 0) ? 0 : i, ax = D->mdl.axis[i];
    u[i] = (tj ? tj[i] : 0.0) - pa[i + 1][ax];
    double pp[6], pap[6];
    for (int k = 0; k < 6; k++) pp[k] = pa[i + 1][k] + D->U[i][k] * u[i] / D->d[i];
    xf_force_T(&D->X[i], pp, pap);
    for (int k = 0; k < 6; k++) pa[pb][k] += pap[k];
  }
  double rhs[6], a[13][6]; for (int k = 0; k < 6; k++) rhs[k] = -pa[0][k];
  solve6(D->IA[0], rhs, a[0]);
  for (int k = 0; k < 6; k++) out[k] = a[0][k];
  for (int i = 0; i < 12; i++) {
    int pb = (i % 3 == 0) ? 0 : i, ax = D->mdl.axis[i];
    double ap[6]; xf_motion(&D->X[i], a[pb], ap);
    double s = 0; for (int k = 0; k < 6; k++) s += D->U[i][k] * ap[k];
    double qdd = (u[i] - s) / D->d[i];
    ap[ax] += qdd; memcpy(a[i + 1], ap, sizeof ap);
    out[6 + i] = qdd;
  }
}

void orc_forward_dynamics(const OrcConfig* c, const OrcEnv* e, const double tau[12], double qdd[12], double wdot_w[3], double vdot_w[3]) {
  (void)c;
  Dyn Dstack; Dyn* D = &Dstack;
  build_model(e->param, &D->mdl);
  dyn_kinematics(e, D);
  double a0[6]; dyn_aba(e, D, tau, qdd, a0);
  double gB[3]; m3tv(D->R0, e->param + 26, gB);
  double wxv[3]; v3cross(D->v[0], D->v[0] + 3, wxv);
  double al[3] = {a0[3] + gB[0] + wxv[0], a0[4] + gB[1] + wxv[1], a0[5] + gB[2] + wxv[2]};
  m3v(D->R0, a0, wdot_w); m3v(D->R0, al, vdot_w);
}

void orc_mass_matrix(const OrcConfig* c, const OrcEnv* e, double M[18 * 18]) {
  (void)c;
  /* columns of M^-1 via dyn_delta, then invert numerically (tests only) */
  Dyn Dstack; Dyn* D = &Dstack;
  build_model(e->param, &D->mdl);
  dyn_kinematics(e, D);
  double tau[12] = {0}, qdd[12], a0[6]; dyn_aba(e, D, tau, qdd, a0);
  double Minv[18][18];
  for (int k = 0; k < 18; k++) {
    double f[6] = {0}, tj[12] = {0}, out[18];
    if (k < 6) { f[k] = 1; dyn_delta(D, 0, f, NULL, out); } else { tj[k - 6] = 1; dyn_delta(D, 0, NULL, tj, out); }
    for (int r = 0; r < 18; r++) Minv[r][k] = out[r];
  }
  /* Gauss-Jordan inverse */
  double A[18][36];
  for (int i = 0; i < 18; i++) for (int j = 0; j < 18; j++) { A[i][j] = Minv[i][j]; A[i][18 + j] = (i == j); }
  for (int k = 0; k < 18; k++) {
    int p = k; for (int i = k + 1; i < 18; i++) if (fabs(A[i][k]) > fabs(A[p][k])) p = i;
    if (p != k) for (int j = 0; j < 36; j++) { double t = A[k][j]; A[k][j] = A[p][j]; A[p][j] = t; }
    double piv = A[k][k]; for (int j = 0; j < 36; j++) A[k][j] /= piv;
    for (int i = 0; i < 18; i++) if (i != k) { double f = A[i][k]; if (f != 0) for (int j = 0; j < 36; j++) A[i][j] -= f * A[k][j]; }
  }
  for (int i = 0; i < 18; i++) for (int j = 0; j < 18; j++) M[18 * i + j] = A[i][18 + j];
}

double orc_energy(const OrcConfig* c, const OrcEnv* e, double* kin, double* pot) {
  (void)c;
  Dyn Dstack; Dyn* D = &Dstack;
  build_model(e->param, &D->mdl);
  dyn_kinematics(e, D);
  double tau[12] = {0}, qdd[12], a0[6]; dyn_aba(e, D, tau, qdd, a0); /* fills v[] */
  double K = 0, P = 0;
  for (int b = 0; b < 13; b++) {
    const double* I = b ? D->mdl.Il[b - 1] : D->mdl.I0;
    double Iv[6]; m6v(I, D->v[b], Iv);
    for (int k = 0; k < 6; k++) K += 0.5 * D->v[b][k] * Iv[k];
    const double* R = b ? D->Rw[b - 1] : D->R0; const double* p = b ? D->pw[b - 1] : e->pos;
    double cw[3]; m3v(R, D->mdl.com[b], cw);
    for (int k = 0; k < 3; k++) P -= D->mdl.mass[b] * e->param[26 + k] * (p[k] + cw[k]);
  }
  if (kin) *kin = K; if (pot) *pot = P;
  return K + P;
}

/* ------------------------------------------------------------------ terrain */
static double terrain_height(const OrcConfig* c, double x, double y, double n[3]) {
  if (c->terrain_type == 0 || !c->hf) { n[0] = 0; n[1] = 0; n[2] = 1; return 0.0; }
  double fx = (x - c->hf_x0) / c->hf_cell, fy = (y - c->hf_y0) / c->hf_cell;
  if (!(fx >= 0)) fx = 0; if (!(fy >= 0)) fy = 0;   /* also catches NaN positions (a diverged state must end the episode, not index out of bounds) */
  if (fx > c->hf_nx - 1.000001) fx = c->hf_nx - 1.000001; if (fy > c->hf_ny - 1.000001) fy = c->hf_ny - 1.000001;
  int ix = (int)fx, iy = (int)fy; double tx = fx - ix, ty = fy - iy;
  const double* h = c->hf; int nx = c->hf_nx;
  double h00 = h[iy * nx + ix], h10 = h[iy * nx + ix + 1], h01 = h[(iy + 1) * nx + ix], h11 = h[(iy + 1) * nx + ix + 1];
  double hh = (1 - tx) * (1 - ty) * h00 + tx * (1 - ty) * h10 + (1 - tx) * ty * h01 + tx * ty * h11;
  double dhdx = ((1 - ty) * (h10 - h00) + ty * (h11 - h01)) / c->hf_cell;
  double dhdy = ((1 - tx) * (h01 - h00) + tx * (h11 - h10)) / c->hf_cell;
  double inv = 1.0 / sqrt(dhdx * dhdx + dhdy * dhdy + 1.0);
  n[0] = -dhdx * inv; n[1] = -dhdy * inv; n[2] = inv;
  return hh;
}

static void toe_world(const Dyn* D, int leg, double o[3]) {
  int i = 3 * leg + 2; double t[3] = {0, 0, -L_LOW}, tw[3];
  m3v(D->Rw[i], t, tw);
  o[0] = D->pw[i][0] + tw[0]; o[1] = D->pw[i][1] + tw[1]; o[2] = D->pw[i][2] + tw[2];
}
void orc_foot_world(const OrcEnv* e, double feet[4][3]) {
  Dyn Dstack; Dyn* D = &Dstack;
  build_model(e->param, &D->mdl); dyn_kinematics(e, D);
  for (int k = 0; k < 4; k++) toe_world(D, k, feet[k]);
}

/* ------------------------------------------------------------------ history (minitaur.py:1142-1193) */
static void hist_push(OrcEnv* e, const double o[ORC_HIST_W]) {
  e->hist_head = (e->hist_head + ORC_HIST - 1) % ORC_HIST; /* appendleft */
  memcpy(e->hist[e->hist_head], o, sizeof(double) * ORC_HIST_W);
  if (e->hist_len < 100) e->hist_len++;
}
static const double* hist_at(const OrcEnv* e, int k) { return e->hist[(e->hist_head + k) % ORC_HIST]; }
static void delayed_obs(const OrcConfig* c, const OrcEnv* e, double latency, double o[ORC_HIST_W]) {
  if (latency <= 0 || e->hist_len == 1) { memcpy(o, hist_at(e, 0), sizeof(double) * ORC_HIST_W); return; }
  int n = (int)(latency / c->sim_dt);
  if (n + 1 >= e->hist_len) { memcpy(o, hist_at(e, e->hist_len - 1), sizeof(double) * ORC_HIST_W); return; }
  double rem = latency - n * c->sim_dt, al = rem / c->sim_dt;
  const double* a = hist_at(e, n); const double* b = hist_at(e, n + 1);
  for (int k = 0; k < ORC_HIST_W; k++) o[k] = (1.0 - al) * a[k] + al * b[k];
}
static void true_obs(const OrcEnv* e, double o[ORC_HIST_W]) {
  double R[9], wb[3];
  memcpy(o, e->q, 12 * sizeof(double)); memcpy(o + 12, e->qd, 12 * sizeof(double)); memcpy(o + 24, e->last_tau, 12 * sizeof(double));
  memcpy(o + 36, e->quat, 4 * sizeof(double));
  quat_to_mat(e->quat, R); m3tv(R, e->vang, wb); /* TransformAngularVelocityToLocalFrame minitaur.py:849-870 */
  memcpy(o + 40, wb, 3 * sizeof(double));
}

/* ------------------------------------------------------------------ one physics substep */
void orc_substep(const OrcConfig* c, OrcEnv* e, const double target[12]) {
  const double dt = c->sim_dt;
  Dyn Dstack; Dyn* D = &Dstack;
  build_model(e->param, &D->mdl);
  dyn_kinematics(e, D);
  /* ApplyAction: PD on the *current* observation (pd_latency = 0, minitaur.py:100,1195-1199) */
  double tau[12], cmd[12];
  for (int j = 0; j < 12; j++) {   /* A1._ClipMotorCommands, a1.py:440-458: np.clip(cmd, q - max_change, q + max_change) on the current angles */
    cmd[j] = target[j];
    if (c->clip_motor_commands) cmd[j] = fmin(fmax(cmd[j], e->q[j] - c->max_angle_change), e->q[j] + c->max_angle_change);
  }
  orc_motor_torque(e->param, e->param + 12, cmd, e->q, e->qd, c->torque_limit, tau);
  if (c->motor_mode == 1) for (int j = 0; j < 12; j++) { /* MotorControlMode.TORQUE: the command is the torque, laikago_motor.py:131-134 */
    double t = target[j]; if (c->torque_limit > 0) { if (t > c->torque_limit) t = c->torque_limit; if (t < -c->torque_limit) t = -c->torque_limit; }
    tau[j] = t;
  }
  if (c->motor_mode == 2) for (int j = 0; j < 12; j++) { /* MotorControlMode.HYBRID, laikago_motor.py:162-164 (+ the torque clip :166-172) */
    double t = -1.0 * (e->hyb[0][j] * (e->q[j] - cmd[j])) - e->hyb[2][j] * (e->qd[j] - e->hyb[1][j]) + e->hyb[3][j];
    if (c->torque_limit > 0) { if (t > c->torque_limit) t = c->torque_limit; if (t < -c->torque_limit) t = -c->torque_limit; }
    tau[j] = t;
  }
  memcpy(e->last_tau, tau, sizeof tau);
  memcpy(D->damp, c->base_damping, sizeof D->damp);
  if (c->external_force) m3tv(D->R0, e->ext_force, D->fext_b);
  double qdd[12], a0[6];
  dyn_aba(e, D, tau, qdd, a0);
  /* unconstrained velocities, world-frame base velocity as Bullet stores it */
  double gB[3]; m3tv(D->R0, e->param + 26, gB);
  double wxv[3]; v3cross(D->v[0], D->v[0] + 3, wxv);
  double al[3] = {a0[3] + gB[0] + wxv[0], a0[4] + gB[1] + wxv[1], a0[5] + gB[2] + wxv[2]}, wd[3], vd[3];
  m3v(D->R0, a0, wd); m3v(D->R0, al, vd);
  double vang[3], vlin[3], qd[12];
  for (int k = 0; k < 3; k++) { vang[k] = e->vang[k] + dt * wd[k]; vlin[k] = e->vlin[k] + dt * vd[k]; }
  for (int j = 0; j < 12; j++) qd[j] = e->qd[j] + dt * qdd[j];
  /* generalized velocity in body coordinates */
  double nu[18]; m3tv(D->R0, vang, nu); m3tv(D->R0, vlin, nu + 3); memcpy(nu + 6, qd, sizeof qd);

  /* contact rows: 4 toe spheres; rows [n0..n3 | (t1,t2) x 4] */
  int act[4]; double dist[4]; double J[12][18], MJ[12][18], A[12], lam[12], targ[12];
  double mu = e->param[24];
  memset(J, 0, sizeof J);
  for (int leg = 0; leg < 4; leg++) {
    double tw[3], n[3]; toe_world(D, leg, tw);
    double h = terrain_height(c, tw[0], tw[1], n);
    dist[leg] = tw[2] - h - c->foot_radius;
    act[leg] = dist[leg] < c->contact_margin;
    /* tangent basis: world x projected on the tangent plane, t2 = n x t1 */
    double t1[3] = {1 - n[0] * n[0], -n[0] * n[1], -n[0] * n[2]}, nn = sqrt(v3dot(t1, t1)), t2[3];
    for (int k = 0; k < 3; k++) t1[k] /= nn;
    v3cross(n, t1, t2);
    double x[3] = {tw[0] - c->foot_radius * n[0], tw[1] - c->foot_radius * n[1], tw[2] - c->foot_radius * n[2]};
    const double* dirs[3] = {n, t1, t2};
    for (int dd = 0; dd < 3; dd++) {
      int row = dd == 0 ? leg : 4 + 2 * leg + (dd - 1);
      const double* ew = dirs[dd];
      double xb_w[3] = {x[0] - e->pos[0], x[1] - e->pos[1], x[2] - e->pos[2]}, xb[3], eb[3], xe[3];
      m3tv(D->R0, xb_w, xb); m3tv(D->R0, ew, eb); v3cross(xb, eb, xe);
      for (int k = 0; k < 3; k++) { J[row][k] = xe[k]; J[row][3 + k] = eb[k]; }
      for (int l = 0; l < 3; l++) {
        int i = 3 * leg + l; double aw[3], ax[3] = {0, 0, 0}, rr[3], cr[3];
        ax[D->mdl.axis[i]] = 1; m3v(D->Rw[i], ax, aw);
        for (int k = 0; k < 3; k++) rr[k] = x[k] - D->pw[i][k];
        v3cross(aw, rr, cr);
        J[row][6 + i] = v3dot(ew, cr);
      }
      /* unit impulse response: force e at x on the calf link, in calf coordinates */
      int ic = 3 * leg + 2; double xl_w[3] = {x[0] - D->pw[ic][0], x[1] - D->pw[ic][1], x[2] - D->pw[ic][2]}, xl[3], el[3], f[6];
      m3tv(D->Rw[ic], xl_w, xl); m3tv(D->Rw[ic], ew, el); v3cross(xl, el, f);
      f[3] = el[0]; f[4] = el[1]; f[5] = el[2];
      dyn_delta(D, ic + 1, f, NULL, MJ[row]);
      double s = 0; for (int k = 0; k < 18; k++) s += J[row][k] * MJ[row][k];
      A[row] = s;
      targ[row] = 0;
    }
    targ[leg] = dist[leg] > 0 ? -dist[leg] / dt : c->erp * (-dist[leg]) / dt;
  }
  /* knee spheres (calf-joint origin, radius 0.02) against the terrain: rows [kn0..kn3 | (kt1,kt2) x 4], built like the toe rows; no warm start */
  int kact[4] = {0, 0, 0, 0}; double kJ[12][18], kMJ[12][18], kA[12], klam[12], ktarg[12];
  memset(kJ, 0, sizeof kJ); memset(klam, 0, sizeof klam); memset(ktarg, 0, sizeof ktarg);
  if (c->knee_contacts) for (int leg = 0; leg < 4; leg++) {
    const double kr = 0.02; int ic = 3 * leg + 2; double n[3];
    double h = terrain_height(c, D->pw[ic][0], D->pw[ic][1], n);
    double kd = D->pw[ic][2] - h - kr;
    kact[leg] = kd < c->contact_margin;
    if (!kact[leg]) continue;
    double t1[3] = {1 - n[0] * n[0], -n[0] * n[1], -n[0] * n[2]}, nn = sqrt(v3dot(t1, t1)), t2[3];
    for (int k = 0; k < 3; k++) t1[k] /= nn;
    v3cross(n, t1, t2);
    double x[3] = {D->pw[ic][0] - kr * n[0], D->pw[ic][1] - kr * n[1], D->pw[ic][2] - kr * n[2]};
    const double* dirs[3] = {n, t1, t2};
    for (int dd = 0; dd < 3; dd++) {
      int row = dd == 0 ? leg : 4 + 2 * leg + (dd - 1);
      const double* ew = dirs[dd];
      double xb_w[3] = {x[0] - e->pos[0], x[1] - e->pos[1], x[2] - e->pos[2]}, xb[3], eb[3], xe[3];
      m3tv(D->R0, xb_w, xb); m3tv(D->R0, ew, eb); v3cross(xb, eb, xe);
      for (int k = 0; k < 3; k++) { kJ[row][k] = xe[k]; kJ[row][3 + k] = eb[k]; }
      for (int l = 0; l < 3; l++) {
        int i = 3 * leg + l; double aw[3], ax[3] = {0, 0, 0}, rr[3], cr[3];
        ax[D->mdl.axis[i]] = 1; m3v(D->Rw[i], ax, aw);
        for (int k = 0; k < 3; k++) rr[k] = x[k] - D->pw[i][k];
        v3cross(aw, rr, cr);
        kJ[row][6 + i] = v3dot(ew, cr);
      }
      double xl_w[3] = {x[0] - D->pw[ic][0], x[1] - D->pw[ic][1], x[2] - D->pw[ic][2]}, xl[3], el[3], f[6];
      m3tv(D->Rw[ic], xl_w, xl); m3tv(D->Rw[ic], ew, el); v3cross(xl, el, f);
      f[3] = el[0]; f[4] = el[1]; f[5] = el[2];
      dyn_delta(D, ic + 1, f, NULL, kMJ[row]);
      double ss = 0; for (int k = 0; k < 18; k++) ss += kJ[row][k] * kMJ[row][k];
      kA[row] = ss;
    }
    ktarg[leg] = kd > 0 ? -kd / dt : c->erp * (-kd) / dt;
  }
  double dnu[18]; memset(dnu, 0, sizeof dnu);
  for (int r = 0; r < 12; r++) lam[r] = 0;
  for (int leg = 0; leg < 4; leg++) {
    lam[leg] = act[leg] ? c->warmstart * e->lam_warm[leg] : 0.0;
    for (int k = 0; k < 18; k++) dnu[k] += MJ[leg][k] * lam[leg];
  }
  /* URDF joint limits (a1.py:186-223) as unilateral rows, Bullet btMultiBodyJointLimitConstraint style: one row per joint towards the
   * nearer stop (the far-side row of Bullet can never be active), same target-velocity rule as a contact (approach allowed up to gap/dt,
   * ERP on violation), solved BEFORE the contact rows in every iteration (non-contact multibody constraints come first in
   * btMultiBodyConstraintSolver::solveSingleIteration). */
  static const double QLO[3] = {-0.802851455917, -1.0471975512, -2.69653369433}, QHI[3] = {0.802851455917, 4.18879020479, -0.916297857297};
  double ls[12], lMJ[12][18], lA[12], ltarg[12], llam[12]; int lact[12];
  for (int i = 0; i < 12; i++) { lact[i] = 0; llam[i] = 0; }
  if (c->joint_limits) for (int i = 0; i < 12; i++) {
    double glo = e->q[i] - QLO[i % 3], ghi = QHI[i % 3] - e->q[i], gap = glo; ls[i] = 1.0;
    if (ghi < glo) { gap = ghi; ls[i] = -1.0; }
    lact[i] = gap < 0.06;   /* further from a stop than 0.06 rad (target velocity beyond -30 rad/s) the row cannot bind: dropped, as in the kernel */
    if (!lact[i]) continue;
    double tj[12]; memset(tj, 0, sizeof tj); tj[i] = ls[i];
    dyn_delta(D, 0, NULL, tj, lMJ[i]);
    lA[i] = ls[i] * lMJ[i][6 + i];
    ltarg[i] = gap > 0 ? -gap / dt : c->erp * (-gap) / dt;
    llam[i] = c->warmstart * e->lam_lim[i];
    for (int k = 0; k < 18; k++) dnu[k] += lMJ[i][k] * llam[i];
  }
  for (int it = 0; it < c->solver_iters; it++) {
    if (c->joint_limits) for (int i = 0; i < 12; i++) if (lact[i]) {
      double u = ls[i] * (nu[6 + i] + dnu[6 + i]);
      double ln = llam[i] + (ltarg[i] - u) / lA[i]; if (ln < 0) ln = 0;
      double dl = ln - llam[i]; llam[i] = ln;
      for (int k = 0; k < 18; k++) dnu[k] += lMJ[i][k] * dl;
    }
    for (int leg = 0; leg < 4; leg++) if (act[leg]) {
      int r = leg; double u = 0; for (int k = 0; k < 18; k++) u += J[r][k] * (nu[k] + dnu[k]);
      double ln = lam[r] + (targ[r] - u) / A[r]; if (ln < 0) ln = 0;
      double dl = ln - lam[r]; lam[r] = ln;
      for (int k = 0; k < 18; k++) dnu[k] += MJ[r][k] * dl;
    }
    for (int leg = 0; leg < 4; leg++) if (kact[leg]) {          /* knee normals after the toe normals */
      int r = leg; double u = 0; for (int k = 0; k < 18; k++) u += kJ[r][k] * (nu[k] + dnu[k]);
      double ln = klam[r] + (ktarg[r] - u) / kA[r]; if (ln < 0) ln = 0;
      double dl = ln - klam[r]; klam[r] = ln;
      for (int k = 0; k < 18; k++) dnu[k] += kMJ[r][k] * dl;
    }
    for (int leg = 0; leg < 4; leg++) if (act[leg]) for (int tdir = 0; tdir < 2; tdir++) {
      int r = 4 + 2 * leg + tdir; double u = 0; for (int k = 0; k < 18; k++) u += J[r][k] * (nu[k] + dnu[k]);
      double lim = mu * lam[leg], ln = lam[r] + (targ[r] - u) / A[r];
      if (ln > lim) ln = lim; if (ln < -lim) ln = -lim;
      double dl = ln - lam[r]; lam[r] = ln;
      for (int k = 0; k < 18; k++) dnu[k] += MJ[r][k] * dl;
    }
    for (int leg = 0; leg < 4; leg++) if (kact[leg]) for (int tdir = 0; tdir < 2; tdir++) {   /* knee friction after the toe friction */
      int r = 4 + 2 * leg + tdir; double u = 0; for (int k = 0; k < 18; k++) u += kJ[r][k] * (nu[k] + dnu[k]);
      double lim = mu * klam[leg], ln = klam[r] + (0.0 - u) / kA[r];
      if (ln > lim) ln = lim; if (ln < -lim) ln = -lim;
      double dl = ln - klam[r]; klam[r] = ln;
      for (int k = 0; k < 18; k++) dnu[k] += kMJ[r][k] * dl;
    }
  }
  for (int leg = 0; leg < 4; leg++) { e->lam_warm[leg] = lam[leg]; e->contact[leg] = lam[leg] > 0; }
  for (int i = 0; i < 12; i++) e->lam_lim[i] = c->joint_limits ? llam[i] : 0.0;
  for (int k = 0; k < 18; k++) nu[k] += dnu[k];
  /* integrate (semi-implicit Euler) */
  m3v(D->R0, nu, e->vang); m3v(D->R0, nu + 3, e->vlin);
  for (int j = 0; j < 12; j++) { e->qd[j] = nu[6 + j]; e->q[j] += dt * e->qd[j]; }
  for (int k = 0; k < 3; k++) e->pos[k] += dt * e->vlin[k];
  {
    double wx = e->vang[0], wy = e->vang[1], wz = e->vang[2], th = sqrt(wx * wx + wy * wy + wz * wz) * dt;
    double kk = th < 1e-4 ? 0.5 - th * th / 48.0 : sin(0.5 * th) / th, cw = cos(0.5 * th);
    double dq[4] = {wx * dt * kk, wy * dt * kk, wz * dt * kk, cw}, *q = e->quat, o[4];
    o[0] = dq[3] * q[0] + dq[0] * q[3] + dq[1] * q[2] - dq[2] * q[1];
    o[1] = dq[3] * q[1] - dq[0] * q[2] + dq[1] * q[3] + dq[2] * q[0];
    o[2] = dq[3] * q[2] + dq[0] * q[1] - dq[1] * q[0] + dq[2] * q[3];
    o[3] = dq[3] * q[3] - dq[0] * q[0] - dq[1] * q[1] - dq[2] * q[2];
    double nn = 1.0 / sqrt(o[0] * o[0] + o[1] * o[1] + o[2] * o[2] + o[3] * o[3]);
    for (int k = 0; k < 4; k++) q[k] = o[k] * nn;
  }
  /* ReceiveObservation */
  double ob[ORC_HIST_W]; true_obs(e, ob); hist_push(e, ob);
}

/* ------------------------------------------------------------------ env */
void orc_env_init(const OrcConfig* c, OrcEnv* e, const double* p) {
  memset(e, 0, sizeof *e);
  if (p) memcpy(e->param, p, sizeof e->param); else orc_default_param(e->param);
  /* default ETG weights: zero (caller sets through reset) */
  (void)c;
}

static void pack_obs(const OrcConfig* c, const OrcEnv* e, const double start_pos[3], double* obs_out, double* ctrl_out, int noisy) {
  double obs[ORC_OBS_DIM];
  double ctrl[ORC_HIST_W]; delayed_obs(c, e, e->param[25], ctrl);
  double nrpy[3] = {0, 0, 0}, ndrpy[3] = {0, 0, 0};
  int noise_on = 0; for (int i = 0; i < 5; i++) if (c->noise_stdev[i] > 0) noise_on = 1;
  if (noisy && noise_on) { /* Minitaur._AddSensorNoise (minitaur.py:1206-1211) on the motor angle / velocity / torque and IMU getters */
    double n4[4];
    for (int leg = 0; leg < 4; leg++) for (int qty = 0; qty < 3; qty++) {
      orc_normal4(c->noise_seed, (unsigned)e->env_id, (unsigned)e->step_count, (unsigned)(16 * leg + qty), n4);
      for (int j = 0; j < 3; j++) ctrl[12 * qty + 3 * leg + j] += c->noise_stdev[qty] * n4[j];
    }
    orc_normal4(c->noise_seed, (unsigned)e->env_id, (unsigned)e->step_count, 64u, n4); for (int k = 0; k < 3; k++) nrpy[k] = c->noise_stdev[3] * n4[k];
    orc_normal4(c->noise_seed, (unsigned)e->env_id, (unsigned)e->step_count, 65u, n4); for (int k = 0; k < 3; k++) ndrpy[k] = c->noise_stdev[4] * n4[k];
  }
  double dtc = c->sim_dt * c->action_repeat;
  for (int k = 0; k < 3; k++) obs[k] = (e->pos[k] - start_pos[k]) / dtc;
  for (int k = 0; k < 4; k++) obs[3 + k] = e->contact[k] ? 1.0 : 0.0;
  double rpy[3]; orc_quat_to_rpy(e->quat, rpy);
  double R[9], wb[3]; quat_to_mat(e->quat, R); m3tv(R, e->vang, wb);
  for (int k = 0; k < 3; k++) { obs[7 + k] = (rpy[k] + nrpy[k] - e->rpy0[k]) / 0.1; obs[10 + k] = (wb[k] + ndrpy[k]) / 0.5; }     /* EnvWrapper.py:79-88 */
  for (int j = 0; j < 12; j++) { obs[13 + j] = (map_pi(ctrl[j]) - POSE_ORI[j]) / 0.1; obs[25 + j] = ctrl[12 + j] / 1.0; } /* :64-70 */
  for (int j = 0; j < 12; j++) obs[37 + j] = (e->etg_act[j] - ETG_MEAN[j]) / ETG_STD[j];                     /* :103-106 */
  if (ctrl_out) memcpy(ctrl_out, ctrl, sizeof ctrl);
  /* sensor_mode selection in sorted-key order (BaseDisplacement, FootContactSensor, IMU, MotorAngle[Acc]) then ETG; normal=0 -> raw units */
  int n = 0; const int nr = c->obs_normal;
  if (c->sensor_dis) for (int i = 0; i < 3; i++) obs_out[n++] = obs[i];
  if (c->sensor_contact) for (int i = 0; i < 4; i++) obs_out[n++] = obs[3 + i];
  if (c->sensor_imu == 1) for (int i = 0; i < 3; i++) obs_out[n++] = nr ? obs[7 + i] : obs[7 + i] * 0.1;
  if (c->sensor_imu) for (int i = 0; i < 3; i++) obs_out[n++] = nr ? obs[10 + i] : obs[10 + i] * 0.5;
  if (c->sensor_motor) for (int i = 0; i < 12; i++) obs_out[n++] = nr ? obs[13 + i] : obs[13 + i] * 0.1 + POSE_ORI[i];
  if (c->sensor_motor == 1) for (int i = 0; i < 12; i++) obs_out[n++] = obs[25 + i];
  if (c->sensor_etg) for (int i = 0; i < 12; i++) obs_out[n++] = nr ? obs[37 + i] : obs[37 + i] * ETG_STD[i] + ETG_MEAN[i];
}

void orc_env_settle(const OrcConfig* c, OrcEnv* e) {
  /* Reset / ResetPose / _SettleDownForReset: minitaur.py:403-445, a1.py:289-304,326-349 */
  e->pos[0] = 0; e->pos[1] = 0; e->pos[2] = 0.32; /* a1.py:52 */
  e->quat[0] = e->quat[1] = e->quat[2] = 0; e->quat[3] = 1;
  memset(e->vlin, 0, sizeof e->vlin); memset(e->vang, 0, sizeof e->vang);
  memcpy(e->q, POSE_ORI, sizeof e->q); memset(e->qd, 0, sizeof e->qd);
  memset(e->lam_warm, 0, sizeof e->lam_warm); memset(e->last_tau, 0, sizeof e->last_tau);
  e->hist_len = 0; e->hist_head = 0; e->has_last = 0; e->step_count = 0;
  double ob[ORC_HIST_W]; true_obs(e, ob); hist_push(e, ob);
  OrcConfig cs = *c; cs.motor_mode = 0; /* the reset pose is held by the POSITION controller (a1.py:289-304) whatever the policy's motor mode */
  for (int i = 0; i < c->settle_steps; i++) orc_substep(&cs, e, POSE_ORI);
  memcpy(e->snap, e->pos, 3 * sizeof(double)); memcpy(e->snap + 3, e->quat, 4 * sizeof(double));
  memcpy(e->snap + 7, e->vlin, 3 * sizeof(double)); memcpy(e->snap + 10, e->vang, 3 * sizeof(double));
  memcpy(e->snap + 13, e->q, 12 * sizeof(double)); memcpy(e->snap + 25, e->qd, 12 * sizeof(double));
  true_obs(e, e->snap_obs); memcpy(e->snap_lam, e->lam_warm, sizeof e->snap_lam);
}

void orc_env_reset(const OrcConfig* c, OrcEnv* e, const double* w, const double* b, double* obs) { orc_env_reset_ex(c, e, w, b, 0.0, obs); }
void orc_env_reset_ex(const OrcConfig* c, OrcEnv* e, const double* w, const double* b, double x_offset, double* obs) {
  /* K2 semantics: masked copy of the pre-settled snapshot; history filled with the settled observation */
  memcpy(e->pos, e->snap, 3 * sizeof(double)); memcpy(e->quat, e->snap + 3, 4 * sizeof(double));
  memcpy(e->vlin, e->snap + 7, 3 * sizeof(double)); memcpy(e->vang, e->snap + 10, 3 * sizeof(double));
  memcpy(e->q, e->snap + 13, 12 * sizeof(double)); memcpy(e->qd, e->snap + 25, 12 * sizeof(double));
  e->pos[0] += x_offset;   /* env.reset(x_noise=...): the episode starts displaced along x (train.py:131,505) */
  memcpy(e->lam_warm, e->snap_lam, sizeof e->snap_lam); memset(e->lam_lim, 0, sizeof e->lam_lim);
  memcpy(e->last_tau, e->snap_obs + 24, 12 * sizeof(double));
  for (int k = 0; k < ORC_HIST; k++) memcpy(e->hist[k], e->snap_obs, sizeof e->snap_obs);
  e->hist_len = 100; e->hist_head = 0;
  e->has_last = 0; e->step_count = 0;
  for (int j = 0; j < 12; j++) e->fx1[j] = e->fx2[j] = e->fy1[j] = e->fy2[j] = e->snap_obs[j]; /* init_history(GetMotorAngles()) at step 0 */
  for (int k = 0; k < 4; k++) e->contact[k] = e->lam_warm[k] > 0;
  if (w) memcpy(e->etg_w, w, sizeof e->etg_w);
  if (b) memcpy(e->etg_b, b, sizeof e->etg_b);
  orc_quat_to_rpy(e->quat, e->rpy0);
  orc_etg_act(c, e->etg_w, e->etg_b, 0.0, e->etg_act, NULL);
  if (obs) pack_obs(c, e, e->pos, obs, NULL, 0);
}

static double c_prec(double v, double t, double m) { double w = (v - t) * atanh(sqrt(0.95)) / m; return tanh(w * w); }

void orc_env_step(const OrcConfig* c, OrcEnv* e, const double* action, int donef,
                  double* obs, double* reward, int* done, double* info) {
  const int R = c->action_repeat; const double dtc = c->sim_dt * R;
  double target[12], start_pos[3], feet0[4][3], feet1[4][3];
  for (int j = 0; j < 12; j++) target[j] = POSE_ORI[j] + e->etg_act[j] + action[j]; /* deployment/test.py:95-99 */
  if (c->motor_mode == 1) memcpy(target, action, sizeof target);                     /* TORQUE mode: the action is the torque */
  if (c->motor_mode == 2) for (int j = 0; j < 12; j++) {   /* HYBRID: per motor (q*, kp, qd*, kd, tau_ff), laikago_motor.py:27-33,152-161; taken as commanded */
    target[j] = action[5 * j]; e->hyb[0][j] = action[5 * j + 1]; e->hyb[1][j] = action[5 * j + 2]; e->hyb[2][j] = action[5 * j + 3]; e->hyb[3][j] = action[5 * j + 4];
  }
  if (c->action_filter && c->motor_mode != 2) { /* Minitaur.Step: action = _FilterAction(action), minitaur.py:250-251 */
    double fb[3], fa[3]; orc_butter2(c->filter_highcut, 1.0 / dtc, fb, fa);
    for (int j = 0; j < 12; j++) orc_filter_step(fb, fa, target[j], &e->fx1[j], &e->fx2[j], &e->fy1[j], &e->fy2[j], &target[j]);
  }
  memcpy(start_pos, e->pos, sizeof start_pos);
  orc_foot_world(e, feet0);
  for (int i = 0; i < R; i++) { /* minitaur.py:248-260 */
    double proc[12];
    if (c->action_interp && e->has_last && c->motor_mode != 2) { double lerp = (double)(i + 1) / R; for (int j = 0; j < 12; j++) proc[j] = e->last_action[j] + lerp * (target[j] - e->last_action[j]); }
    else memcpy(proc, target, sizeof proc);
    orc_substep(c, e, proc);
  }
  memcpy(e->last_action, target, sizeof target); e->has_last = 1;
  e->step_count++;
  orc_etg_act(c, e->etg_w, e->etg_b, e->step_count * dtc, e->etg_act, NULL);
  double ctrl[ORC_HIST_W];
  pack_obs(c, e, start_pos, obs, ctrl, 1);
  /* reward (this repo's definition; rlschool RewardShaping absent) */
  orc_foot_world(e, feet1);
  double velx = (e->pos[0] - start_pos[0]) / dtc;
  double torso = velx < c->vel_d ? velx : c->vel_d;
  double feet = 0; for (int k = 0; k < 4; k++) { double fv = (feet1[k][0] - feet0[k][0]) / dtc; feet += (fv < c->vel_d ? fv : c->vel_d) / 4.0; }
  double rpy[3]; orc_quat_to_rpy(e->quat, rpy);
  double up = 1.0 - 0.5 * (c_prec(rpy[0], 0, 0.25) + c_prec(rpy[1], 0, 0.25));
  double pw = 0; for (int j = 0; j < 12; j++) pw += ctrl[24 + j] * ctrl[12 + j];
  double energy = fabs(pw) * c->sim_dt * R; /* minitaur.py:810-818 */
  double Rm[9]; quat_to_mat(e->quat, Rm);
  /* knees (calf joint origins) */
  int bad = 0, nofoot = 0; double meanz = 0; int above = 0;
  {
    Dyn Dstack; Dyn* D = &Dstack; build_model(e->param, &D->mdl); dyn_kinematics(e, D);
    for (int k = 0; k < 4; k++) {
      double nrm[3]; double h = terrain_height(c, D->pw[3 * k + 2][0], D->pw[3 * k + 2][1], nrm);
      if (!c->body_collisions) { if (D->pw[3 * k + 2][2] - h < 0.03) bad++; }
      else { /* knee sphere r 0.02, hip joint cylinder r 0.046, two trunk-box corners per leg quadrant (a1 URDF shapes [EXT]) */
        if (D->pw[3 * k + 2][2] - h < 0.02) bad++;
        double hh = terrain_height(c, D->pw[3 * k + 1][0], D->pw[3 * k + 1][1], nrm);
        if (D->pw[3 * k + 1][2] - hh < 0.046) bad++;
        double cx = (k < 2) ? 0.1335 : -0.1335, cy = (k & 1) ? 0.097 : -0.097;
        for (int zz = 0; zz < 2; zz++) {
          double cb[3] = {cx + COM_OFFSET[0], cy + COM_OFFSET[1], (zz ? 0.057 : -0.057) + COM_OFFSET[2]}, cw[3];
          m3v(Rm, cb, cw);
          double ch = terrain_height(c, e->pos[0] + cw[0], e->pos[1] + cw[1], nrm);
          if (e->pos[2] + cw[2] - ch < 0) bad++;
        }
      }
      if (!e->contact[k]) nofoot++;
      double fb_w[3] = {feet1[k][0] - e->pos[0], feet1[k][1] - e->pos[1], feet1[k][2] - e->pos[2]}, fb[3];
      m3tv(Rm, fb_w, fb); meanz += fb[2] / 4.0; if (fb[2] > 0) above = 1;
    }
    }
  int nanf = 0;
  for (int k = 0; k < 3; k++) if (!isfinite(e->pos[k]) || !isfinite(e->vlin[k])) nanf = 1;
  for (int j = 0; j < 12; j++) if (!isfinite(e->q[j]) || !isfinite(e->qd[j])) nanf = 1;
  int fall = (Rm[8] < 0.5) || (meanz > -0.1) || above || nanf;
  double r_torso = c->w_torso * torso, r_feet = c->w_feet * feet, r_up = c->w_up * up, r_tau = -c->w_tau * energy;
  double r_bad = -c->w_badfoot * bad, r_fc = -c->w_footcontact * (nofoot > 2 ? nofoot - 2 : 0), r_done = fall ? -c->w_done : 0.0;
  *reward = c->reward_p * (r_torso + r_feet + r_up + r_tau + r_bad + r_fc + r_done);
  int stuck = 0;
  if (c->stuck_termination) {
    memcpy(e->pos_hist[(e->step_count - 1) % 10], e->pos, 3 * sizeof(double));
    if (e->step_count > 10) {
      double m[3] = {0, 0, 0}, v = 0;
      for (int h = 0; h < 10; h++) for (int k = 0; k < 3; k++) m[k] += (e->pos_hist[h][k] - e->pos[k]) / 10.0;
      for (int h = 0; h < 10; h++) for (int k = 0; k < 3; k++) { double d = e->pos_hist[h][k] - e->pos[k] - m[k]; v += d * d / 10.0; }
      stuck = v <= 2e-4 * 2e-4;
    }
  }
  *done = fall || donef || (c->max_episode_steps > 0 && e->step_count >= c->max_episode_steps) || stuck;
  if (info) {
    memset(info, 0, sizeof(double) * ORC_INFO_DIM);
    info[0] = velx; info[1] = r_torso; info[2] = r_feet; info[3] = r_up; info[4] = r_tau; info[5] = 0; info[6] = r_bad; info[7] = r_fc; info[8] = r_done;
    info[9] = nanf; info[10] = energy; info[11] = e->pos[2];
    for (int j = 0; j < 12; j++) { info[12 + j] = e->etg_act[j]; info[24 + j] = target[j]; info[42 + j] = e->q[j]; }
    double wb[3]; m3tv(Rm, e->vang, wb);
    for (int k = 0; k < 3; k++) { info[36 + k] = rpy[k]; info[39 + k] = wb[k]; }
    info[54] = fall; info[55] = e->step_count;
  }
}

/* ------------------------------------------------------------------ batch (cpu baseline) */
typedef struct { const OrcConfig* c; OrcEnv* envs; int lo, hi; const double* act; int donef, auto_reset; double* obs; double* rew; int* done; double* info; } Job;
static void* job_run(void* p) {
  Job* j = (Job*)p;
  for (int i = j->lo; i < j->hi; i++) {
    double info[ORC_INFO_DIM];
    orc_env_step(j->c, &j->envs[i], j->act + (size_t)(j->c->motor_mode == 2 ? 60 : 12) * i, j->donef, j->obs + ORC_OBS_DIM * i, j->rew + i, j->done + i, j->info ? j->info + ORC_INFO_DIM * i : info);
    if (j->auto_reset && j->done[i]) orc_env_reset(j->c, &j->envs[i], NULL, NULL, j->obs + ORC_OBS_DIM * i);
  }
  return NULL;
}
void orc_batch_step(const OrcConfig* c, OrcEnv* envs, int n, const double* actions, int donef, int auto_reset,
                    double* obs, double* reward, int* done, double* info, int nthreads) {
  if (nthreads < 1) nthreads = 1; if (nthreads > 256) nthreads = 256; if (nthreads > n) nthreads = n;
  pthread_t th[256]; Job jobs[256];
  for (int t = 0; t < nthreads; t++) {
    jobs[t] = (Job){c, envs, (int)((long)n * t / nthreads), (int)((long)n * (t + 1) / nthreads), actions, donef, auto_reset, obs, reward, done, info};
    if (t > 0) pthread_create(&th[t], NULL, job_run, &jobs[t]);
  }
  job_run(&jobs[0]);
  for (int t = 1; t < nthreads; t++) pthread_join(th[t], NULL);
}
/* K consecutive control steps per env inside each thread (envs are independent: mirrors one-env-per-process
 * actors, Dynamic_parallel_model.py:96-99).  actions: [K][n][12]; returns the sum of rewards in ret[n]. */
typedef struct { const OrcConfig* c; OrcEnv* envs; int lo, hi, n, K; const double* act; int auto_reset; double* obs; double* ret; int* ndone; } RJob;
static void* rjob_run(void* p) {
  RJob* j = (RJob*)p;
  for (int i = j->lo; i < j->hi; i++) {
    double info[ORC_INFO_DIM], rew; int done; j->ret[i] = 0; j->ndone[i] = 0;
    for (int k = 0; k < j->K; k++) {
      orc_env_step(j->c, &j->envs[i], j->act + ((size_t)k * j->n + i) * (j->c->motor_mode == 2 ? 60 : 12), 0, j->obs + ORC_OBS_DIM * i, &rew, &done, info);
      j->ret[i] += rew; j->ndone[i] += done;
      if (j->auto_reset && done) orc_env_reset(j->c, &j->envs[i], NULL, NULL, j->obs + ORC_OBS_DIM * i);
    }
  }
  return NULL;
}
void orc_batch_rollout(const OrcConfig* c, OrcEnv* envs, int n, const double* actions, int K, int auto_reset,
                       double* obs, double* ret, int* ndone, int nthreads) {
  if (nthreads < 1) nthreads = 1; if (nthreads > 256) nthreads = 256; if (nthreads > n) nthreads = n;
  pthread_t th[256]; RJob jobs[256];
  for (int t = 0; t < nthreads; t++) {
    jobs[t] = (RJob){c, envs, (int)((long)n * t / nthreads), (int)((long)n * (t + 1) / nthreads), n, K, actions, auto_reset, obs, ret, ndone};
    if (t > 0) pthread_create(&th[t], NULL, rjob_run, &jobs[t]);
  }
  rjob_run(&jobs[0]);
  for (int t = 1; t < nthreads; t++) pthread_join(th[t], NULL);
}
int orc_sizeof_env(void) { return (int)sizeof(OrcEnv); }
