// b2q_model_host.h — host-side construction of the A1 model constants consumed by the kernels.
// Numbers: link lengths / hip offsets / init pose from the reference (a1.py:52,70-73,83,98-100); link inertials
// from pybullet_data a1/a1.urdf as recalled in SURVEY.md App. B.3 (the URDF is absent from /root/reference —
// UNVERIFIED); ETG constants from train.py:296-297 and EnvWrapper.py:50-55.
#pragma once
#include <cmath>
#include <cstring>
#include "b2q_sim.cuh"

namespace b2q {

struct A1Nominal {
  static constexpr double l_up = 0.2, l_low = 0.2, l_hip = 0.08505;
};

template <typename T>
inline void build_model_host(Model<T>& M, double foot_radius, double etg_T, double etg_amp, double ph0, double ph1, double foot_y_inset = 0.0) {
  const double COM_OFF[3] = {-0.012731, -0.002186, -0.000515};
  const double HIP_XY[4][2] = {{0.183, -0.047}, {0.183, 0.047}, {-0.183, -0.047}, {-0.183, 0.047}};
  const double BASE_FOOT[4][3] = {{0.18, -0.15, -0.23}, {0.18, 0.148, -0.23}, {-0.18, -0.14, -0.23}, {-0.18, 0.135, -0.23}};
  const double ETG_MEAN[12] = {2.1505982e-02, 3.6674485e-02, -6.0444288e-02, 2.4625482e-02, 1.5869144e-02, -3.2513142e-02,
                               2.1506395e-02, 3.1869926e-02, -6.0140789e-02, 2.4625063e-02, 1.1628972e-02, -3.2163858e-02};
  const double ETG_STD[12] = {4.5967497e-02, 2.0340437e-01, 3.7410179e-01, 4.6187632e-02, 1.9441207e-01, 3.9488649e-01,
                              4.5966785e-02, 2.0323379e-01, 3.7382501e-01, 4.6188373e-02, 1.9457331e-01, 3.9302582e-01};
  const double TRUNK_M = 4.713, TRUNK_I[6] = {0.01683993, 8.3902e-05, 0.000597679, 0.056579028, 2.5134e-05, 0.064713601};
  const double HIP_M = 0.696, HIP_C[3] = {-0.003311, 0.000635, 3.1e-05};
  const double HIP_I[6] = {0.000469246, -9.409e-06, -3.42e-07, 0.00080749, -4.66e-07, 0.000552929};
  const double THIGH_M = 1.013, THIGH_C[3] = {-0.003237, -0.022327, -0.027326};
  const double THIGH_I[6] = {0.005529065, 4.825e-06, 0.000343869, 0.005139339, 2.2448e-05, 0.001367788};
  const double CALF_M = 0.166, CALF_C[3] = {0.006435, 0.0, -0.107388};
  const double CALF_I[6] = {0.002997972, 0.0, -0.000141163, 0.003014022, 0.0, 3.2426e-05};
  const double TOE_M = 0.06, TOE_I = 9.6e-06;
  const double PI = 3.14159265358979323846;

  std::memset(&M, 0, sizeof(M));
  M.m0 = (T)TRUNK_M;
  for (int i = 0; i < 6; i++) M.I0[i] = (T)TRUNK_I[i];
  M.foot_r = (T)foot_radius; M.l_up = (T)A1Nominal::l_up; M.l_low = (T)A1Nominal::l_low;
  M.pose_ori[0] = (T)0.0; M.pose_ori[1] = (T)0.9; M.pose_ori[2] = (T)-1.8;
  M.qlo[0] = (T)-0.802851455917; M.qhi[0] = (T)0.802851455917;      // a1.py:186-223 (same bounds on the four legs)
  M.qlo[1] = (T)-1.0471975512; M.qhi[1] = (T)4.18879020479;
  M.qlo[2] = (T)-2.69653369433; M.qhi[2] = (T)-0.916297857297;
  M.knee_r = (T)0.02;
  M.obs_dim = OBS_DIM; M.obs_identity = 1;
  for (int j = 0; j < OBS_DIM; j++) { M.obs_src[j] = j; M.obs_scale[j] = (T)1; M.obs_shift[j] = (T)0; }
  for (int j = 0; j < 12; j++) { M.etg_mean[j] = (T)ETG_MEAN[j]; M.etg_std[j] = (T)ETG_STD[j]; M.etg_istd[j] = (T)(1.0 / ETG_STD[j]); }
  for (int h = 0; h < ETG_H; h++) {  // RBF centres: forward(h*T/(H-0.9)), SURVEY App. A
    double t = h * etg_T / (ETG_H - 0.9), om = 2 * PI / etg_T;
    M.etg_u[h][0] = (T)(etg_amp * std::sin(ph0 + om * t));
    M.etg_u[h][1] = (T)(etg_amp * std::sin(ph1 + om * t));
  }
  for (int leg = 0; leg < 4; leg++) {
    LegModel<T>& L = M.leg[leg];
    double mirror = (leg % 2) ? 1.0 : -1.0, fh = (leg < 2) ? 1.0 : -1.0;
    L.p1[0] = (T)(HIP_XY[leg][0] + COM_OFF[0]); L.p1[1] = (T)(HIP_XY[leg][1] + COM_OFF[1]); L.p1[2] = (T)COM_OFF[2];
    L.lhip = (T)(A1Nominal::l_hip * mirror);
    for (int a = 0; a < 3; a++) M.base_foot[leg][a] = (T)(BASE_FOOT[leg][a] - (a == 1 ? (BASE_FOOT[leg][1] > 0 ? foot_y_inset : -foot_y_inset) : 0.0));
    // hip
    L.m[0] = (T)HIP_M;
    L.com[0][0] = (T)(HIP_C[0] * fh); L.com[0][1] = (T)(HIP_C[1] * mirror); L.com[0][2] = (T)HIP_C[2];
    { double s[6]; std::memcpy(s, HIP_I, sizeof s); s[1] *= mirror * fh; s[2] *= fh; s[4] *= mirror; for (int i = 0; i < 6; i++) L.I[0][i] = (T)s[i]; }
    // thigh
    L.m[1] = (T)THIGH_M;
    L.com[1][0] = (T)THIGH_C[0]; L.com[1][1] = (T)(THIGH_C[1] * mirror); L.com[1][2] = (T)THIGH_C[2];
    { double s[6]; std::memcpy(s, THIGH_I, sizeof s); s[1] *= mirror; s[4] *= mirror; for (int i = 0; i < 6; i++) L.I[1][i] = (T)s[i]; }
    // calf + fixed toe merged into one body (composite mass / COM / inertia)
    {
      double mt = CALF_M + TOE_M, toe[3] = {0, 0, -A1Nominal::l_low}, c[3];
      for (int a = 0; a < 3; a++) c[a] = (CALF_M * CALF_C[a] + TOE_M * toe[a]) / mt;
      double I[3][3] = {{CALF_I[0], CALF_I[1], CALF_I[2]}, {CALF_I[1], CALF_I[3], CALF_I[4]}, {CALF_I[2], CALF_I[4], CALF_I[5]}};
      for (int a = 0; a < 3; a++) I[a][a] += TOE_I;
      for (int b = 0; b < 2; b++) {
        const double* cb = b ? toe : CALF_C; double mb = b ? TOE_M : CALF_M;
        double d[3] = {cb[0] - c[0], cb[1] - c[1], cb[2] - c[2]}, dd = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
        for (int r = 0; r < 3; r++) for (int s = 0; s < 3; s++) I[r][s] += mb * ((r == s ? dd : 0.0) - d[r] * d[s]);
      }
      L.m[2] = (T)mt;
      for (int a = 0; a < 3; a++) L.com[2][a] = (T)c[a];
      L.I[2][0] = (T)I[0][0]; L.I[2][1] = (T)I[0][1]; L.I[2][2] = (T)I[0][2]; L.I[2][3] = (T)I[1][1]; L.I[2][4] = (T)I[1][2]; L.I[2][5] = (T)I[2][2];
    }
  }
}

// default per-env dynamics row (B2Q_DYN_DIM = 48): a1.py:75-80,233, train.py:125
inline void default_dyn_row(double* p) {
  for (int i = 0; i < 12; i++) { p[i] = 100.0; p[12 + i] = (i % 3 == 0) ? 1.0 : 2.0; }
  p[24] = 1.0; p[25] = 0.002; p[26] = 0; p[27] = 0; p[28] = -10.0;
  for (int i = 29; i < 48; i++) p[i] = 1.0;
}

}  // namespace b2q
