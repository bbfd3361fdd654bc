// ccp_pack.h — host-side packing of the public model description (include/ccp.h) into the
// device image (ccp_core.h).  Pure setup code: runs once per handle, like
// PandaModel::initModel (panda_rbdl.cpp:73-148) and grasping_point() (grasping_point.cpp:5-33).
#pragma once

#include <math.h>

#include "ccp.h"
#include "ccp_core.h"

// rotation matrix (row-major) -> unit quaternion (w,x,y,z); largest-pivot branch selection
static inline void ccp_pack_rot_to_quat(const double* R, double* q) {
  const double tr = R[0] + R[4] + R[8];
  if (tr > 0.0) {
    double s = sqrt(tr + 1.0) * 2.0;
    q[0] = 0.25 * s;
    q[1] = (R[7] - R[5]) / s;
    q[2] = (R[2] - R[6]) / s;
    q[3] = (R[3] - R[1]) / s;
  } else if (R[0] > R[4] && R[0] > R[8]) {
    double s = sqrt(1.0 + R[0] - R[4] - R[8]) * 2.0;
    q[0] = (R[7] - R[5]) / s;
    q[1] = 0.25 * s;
    q[2] = (R[1] + R[3]) / s;
    q[3] = (R[2] + R[6]) / s;
  } else if (R[4] > R[8]) {
    double s = sqrt(1.0 + R[4] - R[0] - R[8]) * 2.0;
    q[0] = (R[2] - R[6]) / s;
    q[1] = (R[1] + R[3]) / s;
    q[2] = 0.25 * s;
    q[3] = (R[5] + R[7]) / s;
  } else {
    double s = sqrt(1.0 + R[8] - R[0] - R[4]) * 2.0;
    q[0] = (R[3] - R[1]) / s;
    q[1] = (R[2] + R[6]) / s;
    q[2] = (R[5] + R[7]) / s;
    q[3] = 0.25 * s;
  }
  double nrm = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  for (int k = 0; k < 4; ++k) q[k] /= nrm;
}

// Returns 0 on success, -1 on an invalid description.
static inline int ccp_pack_model(const ccp_model_desc* d, ccp_model* M) {
  if (!d || !M) return -1;
  if (d->n_arms < 2 || d->n_arms > CCP_MAX_ARMS) return -1;
  memset(M, 0, sizeof *M);
  M->n_arms = d->n_arms;
  M->max_iter = 250;  // ConstraintFunction.h:26
  ccp_model_set_tolerance(M, 1e-3, 5e-3);  // ConstrainedPlanningCommon.cpp:120-121
  M->step = 0.30;     // ConstraintFunction.h:71
  for (int i = 0; i < CCP_DOF; ++i) {
    M->lb[i] = d->lb[i];
    M->ub[i] = d->ub[i];
  }
  ccp_model_set_margin(M, 1e-3);  // ConstraintFunction.h:45
  // Stock Panda alpha pattern on every arm (bitwise: no alpha calibration) -> structured link code.
  static const double kPi2 = 1.57079632679489661923;
  static const double kPandaAlpha[7] = {0.0, -1.0 * kPi2, kPi2, kPi2, -1.0 * kPi2, kPi2, kPi2};
  int panda = 1;
  for (int a = 0; a < d->n_arms; ++a)
    for (int i = 0; i < CCP_DOF; ++i)
      if (d->arm[a].dh_alpha[i] != kPandaAlpha[i]) panda = 0;
  M->panda_alpha = panda;
  for (int a = 0; a < d->n_arms; ++a) {
    const ccp_arm_desc& s = d->arm[a];
    ccp_arm& A = M->arm[a];
    for (int i = 0; i < CCP_DOF; ++i) {
      ccp_link& L = A.link[i];
      const double al = s.dh_alpha[i];
      if (panda) {
        // exact quarter turns: sin = +-1, cos = 0 (the reference's cos(M_PI_2) = 6.1e-17 is rounding noise)
        L.sa = (i == 0) ? 0.0 : ((al < 0.0) ? -1.0 : 1.0);
        L.ca = (i == 0) ? 1.0 : 0.0;
      } else {
        L.sa = sin(al);
        L.ca = cos(al);
      }
      L.sha = sin(0.5 * al);
      L.cha = cos(0.5 * al);
      L.tx = s.dh_a[i];              // panda_rbdl.cpp:159: (a, -sin(alpha) d, cos(alpha) d)
      L.ty = -1.0 * L.sa * s.dh_d[i];
      L.tz = L.ca * s.dh_d[i];
      L.qoff = s.dh_theta_offset[i];
      L.hqoff = 0.5 * s.dh_theta_offset[i];
    }
    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < 3; ++c) A.Rwb[3 * r + c] = s.t_wb[4 * r + c];
      A.pwb[r] = s.t_wb[4 * r + 3];
    }
    ccp_pack_rot_to_quat(A.Rwb, A.qwb);
    A.fl = s.flange;
    A.sphi = sin(s.ee_yaw);
    A.cphi = cos(s.ee_yaw);
    A.shphi = sin(0.5 * s.ee_yaw);
    A.chphi = cos(0.5 * s.ee_yaw);
    A.hq7 = A.link[6].hqoff + 0.5 * s.ee_yaw;
    A.r6[0] = A.link[6].tx;
    A.r6[1] = -1.0 * A.link[6].sa * A.fl + A.link[6].ty;
    A.r6[2] = A.link[6].ca * A.fl + A.link[6].tz;
  }
  // base 0 seen from base a:  Rrel = Rwb_a^T Rwb_0,  prel = Rwb_a^T (pwb_0 - pwb_a),  qrel = conj(qwb_a) qwb_0
  for (int a = 1; a < d->n_arms; ++a) {
    ccp_arm& A = M->arm[a];
    const ccp_arm& Z = M->arm[0];
    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < 3; ++c)
        A.Rrel[3 * r + c] = A.Rwb[0 + r] * Z.Rwb[0 + c] + A.Rwb[3 + r] * Z.Rwb[3 + c] + A.Rwb[6 + r] * Z.Rwb[6 + c];
      A.prel[r] = A.Rwb[0 + r] * (Z.pwb[0] - A.pwb[0]) + A.Rwb[3 + r] * (Z.pwb[1] - A.pwb[1]) +
                  A.Rwb[6 + r] * (Z.pwb[2] - A.pwb[2]);
    }
    ccp_qmul_conj_left(A.qwb, Z.qwb, A.qrel);
    for (int k = 0; k < 4; ++k) A.qrel_scaled[k] = A.qrel[k] * CCP_PANDA_QSCALE;
  }
  for (int p = 0; p < CCPC_MAX_ARMS - 1; ++p) {
    M->ref[p].q0[0] = 1.0;
  }
  // "stock" (ccp_core.h, PANDA = 2): the terms the specialised code skips must be exact zeros / unit diagonals
  int stock = panda && !(d->flags & CCP_MODEL_NO_STOCK);
  static const int kZeroA[7] = {1, 1, 1, 0, 0, 1, 0}, kZeroD[7] = {0, 1, 0, 1, 0, 1, 1};
  for (int a = 0; a < d->n_arms && stock; ++a) {
    for (int i = 0; i < CCP_DOF; ++i) {
      if (d->arm[a].dh_theta_offset[i] != 0.0) stock = 0;
      if (kZeroA[i] && d->arm[a].dh_a[i] != 0.0) stock = 0;
      if (kZeroD[i] && d->arm[a].dh_d[i] != 0.0) stock = 0;
    }
    if (a >= 1)
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
          const double v = M->arm[a].Rrel[3 * r + c];
          if (r == c ? (v != 1.0 && v != -1.0) : (v != 0.0)) stock = 0;
        }
  }
  M->stock = stock;
  return 0;
}

// The reference's stock constants (panda_rbdl.cpp:97-99,125,31; ConstraintFunction.h:27-28;
// grasping_point.cpp:11-16).
static inline int ccp_fill_default_model(int32_t n_arms, const int32_t* arm_index, ccp_model_desc* d) {
  if (!d || !arm_index || n_arms < 2 || n_arms > CCP_MAX_ARMS) return -1;
  static const double kPi2 = 1.57079632679489661923;  // M_PI_2
  static const double al[7] = {0.0, -1.0 * kPi2, kPi2, kPi2, -1.0 * kPi2, kPi2, kPi2};
  static const double aa[7] = {0.0, 0.0, 0.0, 0.0825, -0.0825, 0.0, 0.088};
  static const double dd[7] = {0.333, 0.0, 0.316, 0.0, 0.384, 0.0, 0.0};
  static const double lb[7] = {-2.8973, -1.7628, -2.8973, -3.0718, -2.8973, -0.0175, -2.8973};
  static const double ub[7] = {2.8973, 1.7628, 2.8973, -0.0698, 2.8973, 3.7525, 2.8973};
  // base frames: 0 = left, 1 = right, 2 = top (top is turned half a revolution about z)
  static const double twb[3][12] = {
      {1, 0, 0, 0.0, 0, 1, 0, 0.3, 0, 0, 1, 1.006},
      {1, 0, 0, 0.0, 0, 1, 0, -0.3, 0, 0, 1, 1.006},
      {-1, 0, 0, 1.35, 0, -1, 0, 0.3, 0, 0, 1, 1.006},
  };
  memset(d, 0, sizeof *d);
  d->n_arms = n_arms;
  for (int i = 0; i < 7; ++i) {
    d->lb[i] = lb[i];
    d->ub[i] = ub[i];
  }
  for (int a = 0; a < n_arms; ++a) {
    const int idx = arm_index[a];
    if (idx < 0 || idx > 2) return -1;
    ccp_arm_desc& A = d->arm[a];
    for (int i = 0; i < 7; ++i) {
      A.dh_a[i] = aa[i];
      A.dh_d[i] = dd[i];
      A.dh_alpha[i] = al[i];
      A.dh_theta_offset[i] = 0.0;
    }
    for (int k = 0; k < 12; ++k) A.t_wb[k] = twb[idx][k];
    A.flange = 0.107;
    A.ee_yaw = -3.14159265358979323846 / 4.0;
  }
  return 0;
}
