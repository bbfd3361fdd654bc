// ccp_ik.h — batched single-arm pose IK (damped least squares with joint-limit clamping), __host__ __device__.
//
// Replaces, for a whole batch of (target, seed) pairs, what the reference does one call at a time through TRAC-IK
// (third party, not vendored): IKTask::solve / random_solve (src/base/constraints/ik_task.cpp:16-49) ->
// panda_ik::solve / TrackIKAdaptor::randomSolve (src/kinematics/panda_tracik.cpp:62-88,140-158), as used by the goal
// sampler (jy_ConstrainedValidStateSampler.h:63-189: one seeded solve, then up to 14 random restarts drawn
// N(nominal, 0.3) clipped to the limits, the solution nearest to the reference configuration wins).
//
// TRAC-IK's arithmetic cannot be reproduced (KDL Newton-Raphson with random restarts racing an SQP solver against a
// wall-clock timeout), and does not need to be: an IK answer is correct iff FK(q) hits the target within the
// tolerance (TRAC-IK's default eps = 1e-5 on every twist component) inside the joint limits, which is what the
// tests check with the reference-faithful FK of oracle A.  The iteration here is KDL's ChainIkSolverPos_NR_JL idea
// (Newton step on the 6-D pose error, clamp to the limits) with a damped normal-equation solve:
//     dq = J^T (J J^T + lambda^2 I)^-1 e,   e = (p_t - p, 2 sgn(w) vec(quat(R_t R^T))),   q <- clamp(q + dq).
// FK and the geometric Jacobian are ccp_arm_fk_t (ccp_core.h), i.e. PandaModel::getTransform / getJacobianMatrix, in the
// model's link-code mode (the stock alpha pattern turns R Rx(alpha) into a column permutation).
#pragma once

#include "ccp_core.h"

struct ccp_ik_opt {
  int32_t max_iter;  // Newton steps per solve
  int32_t pad;
  double eps_p;      // |p_t - p|_inf  <= eps_p  (m)
  double eps_r;      // |e_rot|_inf    <= eps_r  (rad)
  double lambda2;    // damping lambda^2 added to the diagonal of J J^T
  double margin;     // a solution must stay this far inside [lb, ub] (TrackIKAdaptor::isValid, panda_tracik.cpp:99-108)
};

// rotation error vector of R_t R^T (row-major 3x3 inputs): 2 sgn(w) vec(q), |.| = 2 sin(theta/2), in the base frame
CCP_HD void ccp_rot_error(const double* Rt, const double* R, double* er) {
  double E[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      E[3 * i + j] = CCP_FMA(Rt[3 * i], R[3 * j], CCP_FMA(Rt[3 * i + 1], R[3 * j + 1], Rt[3 * i + 2] * R[3 * j + 2]));
  // Shepperd: pick the largest of (trace, E00, E11, E22) as pivot
  const double tr = E[0] + E[4] + E[8];
  double w, x, y, z;
  // (one reciprocal per branch and three products: an FP64 division is ~14 instructions)
  if (tr > 0.0) {
    const double s = sqrt(tr + 1.0) * 2.0;  // 4 w
    const double is = 1.0 / s;
    w = 0.25 * s;
    x = (E[7] - E[5]) * is;
    y = (E[2] - E[6]) * is;
    z = (E[3] - E[1]) * is;
  } else if (E[0] > E[4] && E[0] > E[8]) {
    const double s = sqrt(1.0 + E[0] - E[4] - E[8]) * 2.0;  // 4 x
    const double is = 1.0 / s;
    w = (E[7] - E[5]) * is;
    x = 0.25 * s;
    y = (E[1] + E[3]) * is;
    z = (E[2] + E[6]) * is;
  } else if (E[4] > E[8]) {
    const double s = sqrt(1.0 + E[4] - E[0] - E[8]) * 2.0;  // 4 y
    const double is = 1.0 / s;
    w = (E[2] - E[6]) * is;
    x = (E[1] + E[3]) * is;
    y = 0.25 * s;
    z = (E[5] + E[7]) * is;
  } else {
    const double s = sqrt(1.0 + E[8] - E[0] - E[4]) * 2.0;  // 4 z
    const double is = 1.0 / s;
    w = (E[3] - E[1]) * is;
    x = (E[2] + E[6]) * is;
    y = (E[5] + E[7]) * is;
    z = 0.25 * s;
  }
  const double sg = (w < 0.0) ? -2.0 : 2.0;
  er[0] = sg * x;
  er[1] = sg * y;
  er[2] = sg * z;
}

// One TRIP of a solve: evaluate the pose error at q; if it is under the tolerance or the iteration budget is spent the
// solve is finished (returns true, q untouched); otherwise take one damped Newton step, clamp, count it, return false.
// The kernels run one trip per loop pass (lane refill); ccp_ik_solve_one below is the same trips in a plain loop.
// Tt: target EE pose in the arm's base frame, row-major 3x4 [R|p] (the frame getTransform returns).
// PANDA: link-code mode of the FK (0 generic, 1 structured alpha, 2 stock table: ccp_arm_fk_t).
template <int PANDA>
CCP_HD bool ccp_ik_trip(const ccp_arm& A, const double* lb, const double* ub, const double* Tt, double* q,
                        const ccp_ik_opt& O, int32_t& it, bool& conv, double& ep_inf, double& er_inf) {
  double Rt[9];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) Rt[3 * r + c] = Tt[4 * r + c];
  {
    double T[12], J[42];
    ccp_arm_fk_t<PANDA>(A, q, T, J);
    double R[9], e[6];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
      for (int c = 0; c < 3; ++c) R[3 * r + c] = T[4 * r + c];
      e[r] = Tt[4 * r + 3] - T[4 * r + 3];
    }
    ccp_rot_error(Rt, R, e + 3);
    ep_inf = fmax(fabs(e[0]), fmax(fabs(e[1]), fabs(e[2])));
    er_inf = fmax(fabs(e[3]), fmax(fabs(e[4]), fabs(e[5])));
    conv = (ep_inf <= O.eps_p) && (er_inf <= O.eps_r);
    if (conv || it >= O.max_iter) return true;
    ++it;
    // G = J J^T + lambda^2 I (lower triangle), G = L D L^T with unit-diagonal L (no square roots: an FP64 sqrt is ~15
    // instructions and a Cholesky needs six), solve G y = e
    double L[6][6], dinv[6];
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
      for (int j = 0; j <= i; ++j) {
        double acc = (i == j) ? O.lambda2 : 0.0;
#pragma unroll
        for (int k = 0; k < CCPC_DOF; ++k) acc = CCP_FMA(J[7 * i + k], J[7 * j + k], acc);
        L[i][j] = acc;
      }
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      // column j.  Below the diagonal L[i][k] (k < j) holds l_ik and, above it, L[k][i] (k < i) holds l_ik d_k
      double d = L[j][j];
#pragma unroll
      for (int k = 0; k < j; ++k) d = CCP_FMA(-L[j][k], L[k][j], d);
      d = fmax(d, 1e-300);  // G is positive definite by construction (lambda^2 > 0); guards a zero Jacobian with lambda = 0
      const double inv = 1.0 / d;
      dinv[j] = inv;
#pragma unroll
      for (int i = j + 1; i < 6; ++i) {
        double v = L[i][j];
#pragma unroll
        for (int k = 0; k < j; ++k) v = CCP_FMA(-L[i][k], L[k][j], v);
        L[j][i] = v;         // l_ij d_j
        L[i][j] = v * inv;   // l_ij
      }
    }
    double y[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {  // L z = e
      double v = e[i];
#pragma unroll
      for (int k = 0; k < i; ++k) v = CCP_FMA(-L[i][k], y[k], v);
      y[i] = v;
    }
#pragma unroll
    for (int i = 5; i >= 0; --i) {  // L^T y = D^-1 z
      double v = y[i] * dinv[i];
#pragma unroll
      for (int k = i + 1; k < 6; ++k) v = CCP_FMA(-L[k][i], y[k], v);
      y[i] = v;
    }
    // q <- clamp(q + J^T y)      (KDL ChainIkSolverPos_NR_JL clamps to the limits after every step)
#pragma unroll
    for (int k = 0; k < CCPC_DOF; ++k) {
      double dq = 0.0;
#pragma unroll
      for (int i = 0; i < 6; ++i) dq = CCP_FMA(J[7 * i + k], y[i], dq);
      double v = q[k] + dq;
      v = (v < lb[k]) ? lb[k] : v;
      v = (v > ub[k]) ? ub[k] : v;
      q[k] = v;
    }
  }
  return false;
}

// success test of a finished solve: converged, and inside the limits by the margin (TrackIKAdaptor::isValid)
CCP_HD bool ccp_ik_accept(const double* lb, const double* ub, const double* q, const ccp_ik_opt& O, bool conv) {
  bool inside = true;
#pragma unroll
  for (int k = 0; k < CCPC_DOF; ++k) inside = inside && !(q[k] < lb[k] + O.margin) && !(q[k] > ub[k] - O.margin);
  return conv && inside;
}

// One solve.  q: seed in, last iterate out.  err[0] = |p_t - p|_inf, err[1] = |e_rot|_inf at exit.
template <int PANDA>
CCP_HD void ccp_ik_solve_one(const ccp_arm& A, const double* lb, const double* ub, const double* Tt, double* q,
                             const ccp_ik_opt& O, int32_t* iters, bool* ok, double* err) {
  int32_t it = 0;
  bool conv = false;
  double ep_inf = 0.0, er_inf = 0.0;
  while (!ccp_ik_trip<PANDA>(A, lb, ub, Tt, q, O, it, conv, ep_inf, er_inf)) {
  }
  *iters = it;
  *ok = ccp_ik_accept(lb, ub, q, O, conv);
  if (err) {
    err[0] = ep_inf;
    err[1] = er_inf;
  }
}

// TrackIKAdaptor::getRandomConfig (panda_tracik.cpp:62-79): N(nominal, sigma) per joint, clipped to the limits;
// nominal = mid-range (panda_tracik.cpp:131-134).  z: a standard normal deviate supplied by the caller.
CCP_HD double ccp_ik_random_joint(double lb, double ub, double sigma, double z) {
  double v = CCP_FMA(z, sigma, 0.5 * (lb + ub));
  v = (v < lb) ? lb : v;
  v = (v > ub) ? ub : v;
  return v;
}
