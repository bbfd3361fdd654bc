/*
 * oracle_a.c — ORACLE-A: reference-faithful CPU restatement of the closed-chain projection.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (closed_chain_motion_planner_b200/, include/)
 * links or calls this file.  It is used by tests/, __graft_entry__.smoke() and the cpu_baseline /
 * --impl reference legs of bench.py.
 *
 * PARITY STATUS: the reference has no tests and cannot be compiled here (needs OMPL,
 * RBDL, Eigen, Boost, ROS, yaml-cpp: none are installed, no network), so nothing is pinned by
 * reference TESTS.  It IS pinned by reference OUTPUTS: debug/dumbbell_path.txt and
 * debug/Wine_Bottle_path.txt (tests/golden/) are path.interpolate() dumps whose interior rows are
 * states of discreteGeodesic between consecutive roadmap vertices, i.e. real outputs of the
 * reference's project().  oa_discrete_geodesic re-walks those edges and reproduces all 25 rows to
 * the 6 printed digits (tests/test_geodesic_golden.py).  Further pins: row 0 = start_joint, the
 * tolerance band of every interior row, the EE poses quoted in config/ *.yaml comments, Franka's
 * published flange pose.
 *
 * It restates, in plain C with libm, the arithmetic of (paths relative to the reference root):
 *   - PandaModel::initModel / transformDH        src/kinematics/panda_rbdl.cpp:73-161
 *   - PandaModel::getTranslation/getRotation/getTransform/getJacobianMatrix
 *                                                 src/kinematics/panda_rbdl.cpp:9-42
 *     including what the THIRD-PARTY library RBDL (unpinned, CMakeLists.txt:33; not vendored)
 *     does for a chain of revolute joints about arbitrary axes: X_J = Xrot(q, axis),
 *     X_lambda = X_J * Xtrans(r), X_base = X_lambda * X_base[parent]
 *     (published algorithm: Featherstone spatial transforms as in rbdl/Kinematics.cc
 *     UpdateKinematicsCustom, CalcBodyToBaseCoordinates, CalcBodyWorldOrientation,
 *     CalcPointJacobian6D).
 *   - KinematicChainConstraint::{setInitialPosition,jointValid,project,function,isSatisfied}
 *                                                 include/.../base/constraints/ConstraintFunction.h:31-120
 *   - ompl::base::Constraint::jacobian (THIRD-PARTY OMPL, unpinned, CMakeLists.txt:35): the default
 *     central-difference stencil, h = sqrt(eps) max(1,|x_j|), 1.5 m1 - 0.6 m2 + 0.1 m3.
 *   - Eigen (THIRD-PARTY, unpinned, CMakeLists.txt:32): Quaterniond(Matrix3d) (Shoemake branches),
 *     angularDistance = 2 atan2(|vec|, |w|) of a * conj(b), Isometry3d inverse/product, and
 *     JacobiSVD(ThinU|ThinV).solve = minimum-norm least squares with rank threshold
 *     min(rows,cols) * eps * sigma_max.  The SVD here is a one-sided Jacobi (Hestenes) SVD.
 *   - grasping_point base frames                  src/kinematics/grasping_point.cpp:11-16
 *   - KinematicChainSpace::enforceBounds          include/.../kinematics/KinematicChain.h:118-130
 *
 * The 3-arm (21-DoF) case has no reference implementation (ConstraintFunction.h:24,135 hard-code two
 * arms); its definition here — chains (arm0,arm1) and (arm0,arm2), co-dimension 4 — IS the spec.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define OA_DOF 7
#define OA_MAX_ARMS 3

typedef struct oa_arm {
  double axis[OA_DOF][3];  /* joint axes at the zero configuration, base frame (panda_rbdl.cpp:120) */
  double jpos[OA_DOF][3];  /* parent->joint translations (panda_rbdl.cpp:128-130) */
  double rot_ee[9];        /* frame-7 orientation at q = 0 (panda_rbdl.cpp:124) */
  double ee_pos[3];        /* rot_ee * (0,0,0.107) (panda_rbdl.cpp:125-126) */
  double M_ee[9];          /* rot_ee * AngleAxis(-pi/4, z) (panda_rbdl.cpp:31) */
  double twb_R[9], twb_p[3];
} oa_arm;

typedef struct oa_model {
  int n_arms;
  int max_iter; /* 250, ConstraintFunction.h:26 */
  double tol1, tol2;
  double step;   /* 0.30 */
  double margin; /* 1e-3 */
  double lb[OA_DOF], ub[OA_DOF];
  oa_arm arm[OA_MAX_ARMS];
  double init_R[OA_MAX_ARMS - 1][9]; /* init_chain_ (ConstraintFunction.h:39) */
  double init_t[OA_MAX_ARMS - 1][3];
} oa_model;

/* ---------- tiny linear algebra, row-major 3x3 ---------- */
static void m3_mul(const double* A, const double* B, double* C) {
  double T[9];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) T[3 * r + c] = A[3 * r] * B[c] + A[3 * r + 1] * B[3 + c] + A[3 * r + 2] * B[6 + c];
  memcpy(C, T, sizeof T);
}
static void m3_tmul(const double* A, const double* B, double* C) { /* A^T B */
  double T[9];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) T[3 * r + c] = A[r] * B[c] + A[3 + r] * B[3 + c] + A[6 + r] * B[6 + c];
  memcpy(C, T, sizeof T);
}
static void m3_vec(const double* A, const double* v, double* o) {
  double t[3];
  for (int r = 0; r < 3; ++r) t[r] = A[3 * r] * v[0] + A[3 * r + 1] * v[1] + A[3 * r + 2] * v[2];
  memcpy(o, t, sizeof t);
}
static void m3_tvec(const double* A, const double* v, double* o) {
  double t[3];
  for (int r = 0; r < 3; ++r) t[r] = A[r] * v[0] + A[3 + r] * v[1] + A[6 + r] * v[2];
  memcpy(o, t, sizeof t);
}

/* transformDH, panda_rbdl.cpp:150-161 */
static void transform_dh(double a, double d, double alpha, double theta, double* R, double* p) {
  double st = sin(theta), ct = cos(theta);
  double sa = sin(alpha), ca = cos(alpha);
  R[0] = ct;      R[1] = -1 * st; R[2] = 0.0;
  R[3] = st * ca; R[4] = ct * ca; R[5] = -1 * sa;
  R[6] = st * sa; R[7] = ct * sa; R[8] = ca;
  p[0] = a; p[1] = -1 * sa * d; p[2] = ca * d;
}

/* PandaModel::initModel(dh), panda_rbdl.cpp:73-148.  dh: 7x4 offsets (a, d, theta, alpha), row-major,
 * or NULL for zeros (the shipped configuration, panda_rbdl.cpp:66-71). */
void oa_init_arm(oa_arm* A, const double* dh, const double* twb12) {
  static const double dh_al[7] = {0.0, -1.0 * M_PI_2, M_PI_2, M_PI_2, -1.0 * M_PI_2, M_PI_2, M_PI_2};
  static const double dh_a[7] = {0.0, 0.0, 0.0, 0.0825, -0.0825, 0.0, 0.088};
  static const double dh_d[7] = {0.333, 0.0, 0.316, 0.0, 0.384, 0.0, 0.0};
  double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, p[3] = {0, 0, 0};
  double gpos[OA_DOF][3];
  for (int i = 0; i < OA_DOF; ++i) {
    double ao = dh ? dh[4 * i + 0] : 0.0, dof = dh ? dh[4 * i + 1] : 0.0;
    double qo = dh ? dh[4 * i + 2] : 0.0, alo = dh ? dh[4 * i + 3] : 0.0;
    double Rl[9], pl[3], Rp[3];
    transform_dh(dh_a[i] + ao, dh_d[i] + dof, dh_al[i] + alo, qo, Rl, pl);
    m3_vec(R, pl, Rp); /* T = T * T_l */
    for (int k = 0; k < 3; ++k) p[k] += Rp[k];
    m3_mul(R, Rl, R);
    for (int k = 0; k < 3; ++k) {
      A->axis[i][k] = R[3 * k + 2];
      gpos[i][k] = p[k];
    }
  }
  memcpy(A->rot_ee, R, sizeof R);
  double e[3] = {0.0, 0.0, 0.107};
  m3_vec(A->rot_ee, e, A->ee_pos);
  for (int k = 0; k < 3; ++k) A->jpos[0][k] = gpos[0][k];
  for (int i = 1; i < OA_DOF; ++i)
    for (int k = 0; k < 3; ++k) A->jpos[i][k] = gpos[i][k] - gpos[i - 1][k];
  /* Eigen::AngleAxisd(-M_PI/4., UnitZ()).toRotationMatrix() */
  double ang = -M_PI / 4., c = cos(ang), s = sin(ang);
  double Rz[9] = {c, -s, 0, s, c, 0, 0, 0, 1};
  m3_mul(A->rot_ee, Rz, A->M_ee);
  for (int r = 0; r < 3; ++r) {
    for (int cc = 0; cc < 3; ++cc) A->twb_R[3 * r + cc] = twb12[4 * r + cc];
    A->twb_p[r] = twb12[4 * r + 3];
  }
}

/* RBDL Xrot(angle, axis).E — the coordinate-transform (transposed Rodrigues) matrix */
static void xrot(double angle, const double* ax, double* E) {
  double s = sin(angle), c = cos(angle), v = 1.0 - c;
  E[0] = ax[0] * ax[0] * v + c;         E[1] = ax[1] * ax[0] * v + ax[2] * s; E[2] = ax[0] * ax[2] * v - ax[1] * s;
  E[3] = ax[0] * ax[1] * v - ax[2] * s; E[4] = ax[1] * ax[1] * v + c;         E[5] = ax[1] * ax[2] * v + ax[0] * s;
  E[6] = ax[0] * ax[2] * v + ax[1] * s; E[7] = ax[1] * ax[2] * v - ax[0] * s; E[8] = ax[2] * ax[2] * v + c;
}

/* UpdateKinematicsCustom for the 7-body chain: E[i] world->body, r[i] body origin in base coords */
static void update_kinematics(const oa_arm* A, const double* q, double E[OA_DOF][9], double r[OA_DOF][3]) {
  double Ep[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, rp[3] = {0, 0, 0};
  for (int i = 0; i < OA_DOF; ++i) {
    double EJ[9], t[3];
    xrot(q[i], A->axis[i], EJ);
    m3_tvec(Ep, A->jpos[i], t); /* r_i = r_parent + E_parent^T * jpos */
    for (int k = 0; k < 3; ++k) r[i][k] = rp[k] + t[k];
    m3_mul(EJ, Ep, E[i]);       /* E_i = E_J * E_parent */
    memcpy(Ep, E[i], sizeof Ep);
    memcpy(rp, r[i], sizeof rp);
  }
}

/* getTransform, panda_rbdl.cpp:35-42 (arm base frame) */
void oa_get_transform(const oa_arm* A, const double* q, double* R, double* p) {
  double E[OA_DOF][9], r[OA_DOF][3], t[3];
  update_kinematics(A, q, E, r);
  m3_tvec(E[6], A->ee_pos, t); /* CalcBodyToBaseCoordinates */
  for (int k = 0; k < 3; ++k) p[k] = r[6][k] + t[k];
  m3_tmul(E[6], A->M_ee, R);   /* CalcBodyWorldOrientation(...).transpose() * M */
}

/* getJacobianMatrix, panda_rbdl.cpp:9-22: 6x7 row-major, rows [linear(3); angular(3)] */
void oa_get_jacobian(const oa_arm* A, const double* q, double* J) {
  double E[OA_DOF][9], r[OA_DOF][3], t[3], p[3];
  update_kinematics(A, q, E, r);
  m3_tvec(E[6], A->ee_pos, t);
  for (int k = 0; k < 3; ++k) p[k] = r[6][k] + t[k];
  for (int i = 0; i < OA_DOF; ++i) {
    double w[3];
    m3_tvec(E[i], A->axis[i], w); /* joint axis in base coordinates */
    double l[3] = {p[0] - r[i][0], p[1] - r[i][1], p[2] - r[i][2]};
    J[0 * 7 + i] = w[1] * l[2] - w[2] * l[1];
    J[1 * 7 + i] = w[2] * l[0] - w[0] * l[2];
    J[2 * 7 + i] = w[0] * l[1] - w[1] * l[0];
    J[3 * 7 + i] = w[0];
    J[4 * 7 + i] = w[1];
    J[5 * 7 + i] = w[2];
  }
}

/* Eigen::Quaterniond(Matrix3d): (w,x,y,z) */
static void quat_from_mat(const double* m, double* q) {
  double t = m[0] + m[4] + m[8];
  if (t > 0) {
    t = sqrt(t + 1.0);
    q[0] = 0.5 * t;
    t = 0.5 / t;
    q[1] = (m[7] - m[5]) * t;
    q[2] = (m[2] - m[6]) * t;
    q[3] = (m[3] - m[1]) * t;
  } else {
    int i = 0;
    if (m[4] > m[0]) i = 1;
    if (m[8] > m[4 * i]) i = 2;
    int j = (i + 1) % 3, k = (j + 1) % 3;
    t = sqrt(m[4 * i] - m[4 * j] - m[4 * k] + 1.0);
    q[1 + i] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (m[3 * k + j] - m[3 * j + k]) * t;
    q[1 + j] = (m[3 * j + i] + m[3 * i + j]) * t;
    q[1 + k] = (m[3 * k + i] + m[3 * i + k]) * t;
  }
}

/* a.angularDistance(b), Eigen >= 3.3 */
static double angular_distance(const double* a, const double* b) {
  /* d = a * conj(b) */
  double bw = b[0], bx = -b[1], by = -b[2], bz = -b[3];
  double dw = a[0] * bw - a[1] * bx - a[2] * by - a[3] * bz;
  double dx = a[0] * bx + a[1] * bw + a[2] * bz - a[3] * by;
  double dy = a[0] * by + a[2] * bw + a[3] * bx - a[1] * bz;
  double dz = a[0] * bz + a[3] * bw + a[1] * by - a[2] * bx;
  return 2.0 * atan2(sqrt(dx * dx + dy * dy + dz * dz), fabs(dw));
}

/* t_wb * getTransform(q), ConstraintFunction.h:89-90 */
static void world_transform(const oa_arm* A, const double* q, double* R, double* p) {
  double Rb[9], pb[3], t[3];
  oa_get_transform(A, q, Rb, pb);
  m3_mul(A->twb_R, Rb, R);
  m3_vec(A->twb_R, pb, t);
  for (int k = 0; k < 3; ++k) p[k] = t[k] + A->twb_p[k];
}

/* t_wa.inverse() * t_w0 */
static void chain(const oa_model* M, const double* x, int a, double* Rc, double* tc) {
  double R0[9], p0[3], Ra[9], pa[3];
  world_transform(&M->arm[0], x, R0, p0);
  world_transform(&M->arm[a], x + OA_DOF * a, Ra, pa);
  m3_tmul(Ra, R0, Rc);
  double d[3] = {p0[0] - pa[0], p0[1] - pa[1], p0[2] - pa[2]};
  m3_tvec(Ra, d, tc);
}

/* setInitialPosition, ConstraintFunction.h:31-40 */
void oa_set_initial_position(oa_model* M, const double* q_start) {
  for (int a = 1; a < M->n_arms; ++a) chain(M, q_start, a, M->init_R[a - 1], M->init_t[a - 1]);
}

/* function, ConstraintFunction.h:84-102; out has 2*(n_arms-1) entries: (err_p, err_r) per chain */
void oa_function(const oa_model* M, const double* x, double* out) {
  for (int a = 1; a < M->n_arms; ++a) {
    double Rc[9], tc[3], qc[4], q0[4];
    chain(M, x, a, Rc, tc);
    quat_from_mat(Rc, qc);
    quat_from_mat(M->init_R[a - 1], q0);
    double err_r = angular_distance(qc, q0);
    const double* t0 = M->init_t[a - 1];
    double e[3] = {tc[0] - t0[0], tc[1] - t0[1], tc[2] - t0[2]};
    double err_p = sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]);
    out[2 * (a - 1)] = err_p;
    out[2 * (a - 1) + 1] = err_r;
  }
}

/* ompl::base::Constraint::jacobian (default): J is m x n ROW-major here */
void oa_jacobian_fd(const oa_model* M, const double* x, double* J) {
  const int n = OA_DOF * M->n_arms, m = 2 * (M->n_arms - 1);
  double y1[OA_DOF * OA_MAX_ARMS], y2[OA_DOF * OA_MAX_ARMS], t1[4], t2[4], m1[4], m2[4], m3[4];
  memcpy(y1, x, n * sizeof(double));
  memcpy(y2, x, n * sizeof(double));
  for (int j = 0; j < n; ++j) {
    const double ax = fabs(x[j]);
    const double h = sqrt(2.220446049250313e-16) * (ax >= 1 ? ax : 1);
    y1[j] += h; y2[j] -= h;
    oa_function(M, y1, t1); oa_function(M, y2, t2);
    for (int k = 0; k < m; ++k) m1[k] = (t1[k] - t2[k]) / (y1[j] - y2[j]);
    y1[j] += h; y2[j] -= h;
    oa_function(M, y1, t1); oa_function(M, y2, t2);
    for (int k = 0; k < m; ++k) m2[k] = (t1[k] - t2[k]) / (y1[j] - y2[j]);
    y1[j] += h; y2[j] -= h;
    oa_function(M, y1, t1); oa_function(M, y2, t2);
    for (int k = 0; k < m; ++k) m3[k] = (t1[k] - t2[k]) / (y1[j] - y2[j]);
    for (int k = 0; k < m; ++k) J[k * n + j] = 1.5 * m1[k] - 0.6 * m2[k] + 0.1 * m3[k];
    y1[j] = y2[j] = x[j];
  }
}

/* Independent analytic Jacobian from the geometric arm Jacobians (for tests of the stencil and of
 * the engine's jacobian(); never used by oa_project).  Row-major m x n. */
void oa_jacobian_analytic(const oa_model* M, const double* x, double* J) {
  const int n = OA_DOF * M->n_arms, m = 2 * (M->n_arms - 1);
  memset(J, 0, sizeof(double) * m * n);
  double R0[9], p0[3];
  world_transform(&M->arm[0], x, R0, p0);
  for (int a = 1; a < M->n_arms; ++a) {
    double Ra[9], pa[3], Rc[9], tc[3], qc[4], q0[4];
    world_transform(&M->arm[a], x + OA_DOF * a, Ra, pa);
    chain(M, x, a, Rc, tc);
    quat_from_mat(Rc, qc);
    quat_from_mat(M->init_R[a - 1], q0);
    double bw = q0[0], bx = -q0[1], by = -q0[2], bz = -q0[3];
    double dw = qc[0] * bw - qc[1] * bx - qc[2] * by - qc[3] * bz;
    double dv[3] = {qc[0] * bx + qc[1] * bw + qc[2] * bz - qc[3] * by,
                    qc[0] * by + qc[2] * bw + qc[3] * bx - qc[1] * bz,
                    qc[0] * bz + qc[3] * bw + qc[1] * by - qc[2] * bx};
    const double* t0 = M->init_t[a - 1];
    double e[3] = {tc[0] - t0[0], tc[1] - t0[1], tc[2] - t0[2]};
    double f0 = sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]);
    double sv = sqrt(dv[0] * dv[0] + dv[1] * dv[1] + dv[2] * dv[2]);
    double u[3] = {0, 0, 0}, nn[3] = {0, 0, 0}, w[3], mm[3];
    if (f0 > 0) for (int k = 0; k < 3; ++k) u[k] = e[k] / f0;
    if (sv > 0) for (int k = 0; k < 3; ++k) nn[k] = (dw < 0 ? -1.0 : 1.0) * dv[k] / sv;
    m3_vec(Ra, u, w);
    m3_vec(Ra, nn, mm);
    /* geometric Jacobians in the world frame, lever arm to p0 for BOTH arms */
    for (int which = 0; which < 2; ++which) {
      const int arm = which == 0 ? 0 : a;
      const oa_arm* A = &M->arm[arm];
      double E[OA_DOF][9], r[OA_DOF][3];
      update_kinematics(A, x + OA_DOF * arm, E, r);
      for (int i = 0; i < OA_DOF; ++i) {
        double zb[3], z[3], ob[3], o[3];
        m3_tvec(E[i], A->axis[i], zb);
        m3_vec(A->twb_R, zb, z);
        m3_vec(A->twb_R, r[i], ob);
        for (int k = 0; k < 3; ++k) o[k] = ob[k] + A->twb_p[k];
        double l[3] = {p0[0] - o[0], p0[1] - o[1], p0[2] - o[2]};
        double c[3] = {z[1] * l[2] - z[2] * l[1], z[2] * l[0] - z[0] * l[2], z[0] * l[1] - z[1] * l[0]};
        double sgn = which == 0 ? 1.0 : -1.0;
        J[(2 * (a - 1)) * n + OA_DOF * arm + i] = sgn * (w[0] * c[0] + w[1] * c[1] + w[2] * c[2]);
        J[(2 * (a - 1) + 1) * n + OA_DOF * arm + i] = sgn * (mm[0] * z[0] + mm[1] * z[1] + mm[2] * z[2]);
      }
    }
  }
}

/* j.jacobiSvd(ThinU|ThinV).solve(f): minimum-norm least-squares solution of J dx = f.
 * One-sided Jacobi on A = J^T (n x m): A V = U S.  dx = U S^-1 V^T f over s_i > m*eps*s_max. */
void oa_svd_solve(const double* J, int m, int n, const double* f, double* dx) {
  double A[OA_DOF * OA_MAX_ARMS][4], V[4][4];
  for (int i = 0; i < n; ++i)
    for (int k = 0; k < m; ++k) A[i][k] = J[k * n + i];
  for (int a = 0; a < m; ++a)
    for (int b = 0; b < m; ++b) V[a][b] = (a == b) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 60; ++sweep) {
    int rotated = 0;
    for (int p = 0; p < m - 1; ++p)
      for (int q = p + 1; q < m; ++q) {
        double alpha = 0, beta = 0, gamma = 0;
        for (int i = 0; i < n; ++i) {
          alpha += A[i][p] * A[i][p];
          beta += A[i][q] * A[i][q];
          gamma += A[i][p] * A[i][q];
        }
        if (gamma == 0.0 || fabs(gamma) <= 2.220446049250313e-16 * sqrt(alpha * beta)) continue;
        rotated = 1;
        double zeta = (beta - alpha) / (2.0 * gamma);
        double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
        for (int i = 0; i < n; ++i) {
          double ap = A[i][p], aq = A[i][q];
          A[i][p] = c * ap - s * aq;
          A[i][q] = s * ap + c * aq;
        }
        for (int i = 0; i < m; ++i) {
          double vp = V[i][p], vq = V[i][q];
          V[i][p] = c * vp - s * vq;
          V[i][q] = s * vp + c * vq;
        }
      }
    if (!rotated) break;
  }
  double sig[4], smax = 0;
  for (int k = 0; k < m; ++k) {
    double s2 = 0;
    for (int i = 0; i < n; ++i) s2 += A[i][k] * A[i][k];
    sig[k] = sqrt(s2);
    if (sig[k] > smax) smax = sig[k];
  }
  const double thr = (double)m * 2.220446049250313e-16 * smax;
  for (int i = 0; i < n; ++i) dx[i] = 0.0;
  for (int k = 0; k < m; ++k) {
    if (!(sig[k] > thr) || sig[k] == 0.0) continue;
    double vf = 0;
    for (int a = 0; a < m; ++a) vf += V[a][k] * f[a];
    double coef = vf / (sig[k] * sig[k]); /* U[:,k] = A[:,k]/sig -> dx += A[:,k] * vf / sig^2 */
    for (int i = 0; i < n; ++i) dx[i] += A[i][k] * coef;
  }
}

/* jointValid, ConstraintFunction.h:43-55 */
int oa_joint_valid(const oa_model* M, const double* q) {
  double eps = 0.001;
  (void)eps;
  for (int arm = 0; arm < M->n_arms; arm++)
    for (int i = 0; i < OA_DOF; i++) {
      if (q[arm * 7 + i] < M->lb[i] + M->margin) return 0;
      if (q[arm * 7 + i] > M->ub[i] - M->margin) return 0;
    }
  return 1;
}

/* isSatisfied, ConstraintFunction.h:114-120 */
int oa_is_satisfied(const oa_model* M, const double* x) {
  double f[4];
  oa_function(M, x, f);
  for (int a = 0; a < M->n_arms - 1; ++a) {
    if (!isfinite(f[2 * a]) || !isfinite(f[2 * a + 1])) return 0;
    if (!(f[2 * a] <= M->tol1 && f[2 * a + 1] <= M->tol2)) return 0;
  }
  return 1;
}

/* project, ConstraintFunction.h:57-82.  use_fd = 1 is the reference (OMPL stencil); use_fd = 0 swaps in
 * oa_jacobian_analytic (for experiments only).  Keeps the reference's norm1/norm2 bookkeeping:
 * `norm1 = f[0] > tol1` is a 0/1 flag, `norm2 = f[1]` is only assigned when the first clause is false.
 * For n_arms = 3 the loop continues while ANY chain violates and succeeds when ALL chains pass.
 * Returns project()'s bool; *iters = Newton steps taken; *conv = residual test alone.            */
int oa_project(const oa_model* M, double* x, int use_fd, int* iters, int* conv, double* f_out) {
  const int n = OA_DOF * M->n_arms, m = 2 * (M->n_arms - 1);
  unsigned int iter = 0;
  int steps = 0;
  double norm1[2] = {0, 0}, norm2[2] = {0, 0};
  double f[4], J[4 * OA_DOF * OA_MAX_ARMS], dx[OA_DOF * OA_MAX_ARMS];
  oa_function(M, x, f);
  for (;;) {
    int cont = 0;
    for (int a = 0; a < M->n_arms - 1; ++a) {
      /* ((norm1 = f[0] > tolerance1_) || (norm2 = f[1]) > tolerance2_) */
      int c;
      if ((norm1[a] = (double)(f[2 * a] > M->tol1)) != 0.0) c = 1;
      else c = ((norm2[a] = f[2 * a + 1]) > M->tol2);
      cont = cont || c;
    }
    if (!(cont && iter++ < (unsigned int)M->max_iter)) break;
    if (use_fd) oa_jacobian_fd(M, x, J);
    else oa_jacobian_analytic(M, x, J);
    oa_svd_solve(J, m, n, f, dx);
    for (int i = 0; i < n; ++i) x[i] -= M->step * dx[i]; /* x -= 0.30 * solve(f) */
    oa_function(M, x, f);
    ++steps;
  }
  int residual_ok = 1;
  for (int a = 0; a < M->n_arms - 1; ++a)
    residual_ok = residual_ok && (norm1[a] < M->tol1) && (norm2[a] < M->tol2);
  if (iters) *iters = steps;
  if (conv) *conv = residual_ok;
  if (f_out) memcpy(f_out, f, sizeof(double) * m);
  return (oa_joint_valid(M, x) && residual_ok) ? 1 : 0;
}

/* KinematicChainSpace::enforceBounds, KinematicChain.h:118-130 */
void oa_enforce_bounds(double* x, int n) {
  const double pi = 3.14159265358979323846;
  for (int i = 0; i < n; ++i) {
    double v = fmod(x[i], 2.0 * pi);
    if (v < -pi) v += 2.0 * pi;
    else if (v >= pi) v -= 2.0 * pi;
    x[i] = v;
  }
}

/* ---------- model assembly from the stock constants ---------- */
/* arm_index: grasping_point::t_wb order, 0 = left, 1 = right, 2 = top (grasping_point.cpp:11-20);
 * dh_offsets: n_arms x 7 x 4 or NULL. */
void oa_default_model(oa_model* M, int n_arms, const int* arm_index, const double* dh_offsets) {
  static const double lb[7] = {-2.8973, -1.7628, -2.8973, -3.0718, -2.8973, -0.0175, -2.8973};
  static const double ub[7] = {2.8973, 1.7628, 2.8973, -0.0698, 2.8973, 3.7525, 2.8973};
  static const double twb[3][12] = {
      {1, 0, 0, 0.0, 0, 1, 0, 0.3, 0, 0, 1, 1.006},
      {1, 0, 0, 0.0, 0, 1, 0, -0.3, 0, 0, 1, 1.006},
      {-1, 0, 0, 1.35, 0, -1, 0, 0.3, 0, 0, 1, 1.006},
  };
  memset(M, 0, sizeof *M);
  M->n_arms = n_arms;
  M->max_iter = 250;
  M->tol1 = 0.001;
  M->tol2 = 0.005;
  M->step = 0.30;
  M->margin = 0.001;
  memcpy(M->lb, lb, sizeof lb);
  memcpy(M->ub, ub, sizeof ub);
  for (int a = 0; a < n_arms; ++a)
    oa_init_arm(&M->arm[a], dh_offsets ? dh_offsets + 28 * a : NULL, twb[arm_index[a]]);
  for (int p = 0; p < OA_MAX_ARMS - 1; ++p) {
    M->init_R[p][0] = M->init_R[p][4] = M->init_R[p][8] = 1.0;
  }
}
void oa_set_arm_base(oa_model* M, int arm, const double* twb12) {
  oa_arm* A = &M->arm[arm];
  for (int r = 0; r < 3; ++r) {
    for (int cc = 0; cc < 3; ++cc) A->twb_R[3 * r + cc] = twb12[4 * r + cc];
    A->twb_p[r] = twb12[4 * r + 3];
  }
}
int oa_model_size(void) { return (int)sizeof(oa_model); }
void oa_set_tolerance(oa_model* M, double t1, double t2) { M->tol1 = t1; M->tol2 = t2; }
void oa_set_options(oa_model* M, double step, int max_iter, double margin) {
  M->step = step; M->max_iter = max_iter; M->margin = margin;
}
void oa_get_init_chain(const oa_model* M, int pair, double* R9, double* t3) {
  memcpy(R9, M->init_R[pair], 9 * sizeof(double));
  memcpy(t3, M->init_t[pair], 3 * sizeof(double));
}
void oa_arm_transform(const oa_model* M, int arm, const double* q, double* T12) {
  double R[9], p[3];
  oa_get_transform(&M->arm[arm], q, R, p);
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) T12[4 * r + c] = R[3 * r + c];
    T12[4 * r + 3] = p[r];
  }
}
void oa_arm_jacobian(const oa_model* M, int arm, const double* q, double* J42) {
  oa_get_jacobian(&M->arm[arm], q, J42);
}

/* ---------- the synthetic seed stream (Seeds-U): same counter-based generator the
 * engine's batched sampler uses, restated here so the CPU baseline consumes identical seeds ---- */
static uint64_t mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
static double uniform01(uint64_t seed, uint64_t sample, uint32_t lane) {
  uint64_t h = mix64(seed ^ 0xD1B54A32D192ED03ull);
  h = mix64(h + sample * 0x9E3779B97F4A7C15ull);
  h = mix64(h + (uint64_t)lane);
  return (double)(h >> 11) * (1.0 / 9007199254740992.0);
}
/* x: AOS count x n */
void oa_seeds_uniform(const oa_model* M, uint64_t seed, int64_t first, int64_t count, double* x) {
  const int n = OA_DOF * M->n_arms;
  for (int64_t s = 0; s < count; ++s)
    for (int j = 0; j < n; ++j) {
      int i = j % OA_DOF;
      x[s * n + j] = fma(uniform01(seed, (uint64_t)(first + s), (uint32_t)j), M->ub[i] - M->lb[i], M->lb[i]);
    }
}

/* ---------- batched drivers (OpenMP over disjoint ranges), AOS count x n ---------- */
void oa_function_batch(const oa_model* M, const double* x, int64_t count, double* f, int nthreads) {
  const int n = OA_DOF * M->n_arms, m = 2 * (M->n_arms - 1);
  (void)nthreads;
#pragma omp parallel for schedule(dynamic, 64) num_threads(nthreads > 0 ? nthreads : 1)
  for (int64_t s = 0; s < count; ++s) oa_function(M, x + s * n, f + s * m);
}
void oa_jacobian_batch(const oa_model* M, const double* x, int64_t count, double* J, int use_fd, int nthreads) {
  const int n = OA_DOF * M->n_arms, m = 2 * (M->n_arms - 1);
  (void)nthreads;
#pragma omp parallel for schedule(dynamic, 16) num_threads(nthreads > 0 ? nthreads : 1)
  for (int64_t s = 0; s < count; ++s) {
    if (use_fd) oa_jacobian_fd(M, x + s * n, J + s * m * n);
    else oa_jacobian_analytic(M, x + s * n, J + s * m * n);
  }
}
/* x is updated in place (like the reference); ok/conv are bytes; iters int32; resid count x m */
void oa_project_batch(const oa_model* M, double* x, int64_t count, int use_fd, uint8_t* ok, uint8_t* conv,
                      int32_t* iters, double* resid, int nthreads) {
  const int n = OA_DOF * M->n_arms, m = 2 * (M->n_arms - 1);
  (void)nthreads;
#pragma omp parallel for schedule(dynamic, 4) num_threads(nthreads > 0 ? nthreads : 1)
  for (int64_t s = 0; s < count; ++s) {
    int it = 0, cv = 0;
    double f[4];
    int r = oa_project(M, x + s * n, use_fd, &it, &cv, f);
    if (ok) ok[s] = (uint8_t)r;
    if (conv) conv[s] = (uint8_t)cv;
    if (iters) iters[s] = it;
    if (resid) memcpy(resid + s * m, f, sizeof(double) * m);
  }
}
int oa_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ---------- manifold traversal ---------- */
/* KinematicChainSpace::interpolate, KinematicChain.h:145-171 */
void oa_interpolate(const double* from, const double* to, double t, double* state, int n) {
  const double pi = 3.14159265358979323846;
  for (int i = 0; i < n; ++i) {
    double diff = to[i] - from[i];
    if (fabs(diff) <= pi) state[i] = from[i] + diff * t;
    else {
      if (diff > 0.0) diff = 2.0 * pi - diff;
      else diff = -2.0 * pi - diff;
      state[i] = from[i] - diff * t;
      if (state[i] > pi) state[i] -= 2.0 * pi;
      else if (state[i] < -pi) state[i] += 2.0 * pi;
    }
  }
}
/* RealVectorStateSpace::distance */
static double oa_distance(const double* a, const double* b, int n) {
  double dist = 0.0;
  for (int i = 0; i < n; ++i) {
    double diff = a[i] - b[i];
    dist += diff * diff;
  }
  return sqrt(dist);
}
/* jy_ProjectedStateSpace::discreteGeodesic(from, to, interpolate = true, &geodesic), jy_ProjectedStateSpace.cpp:32-96.
 * states: max_states x n, states[0] = from.  Returns the reference's bool; *n_states = geodesic->size().
 * (With interpolate = true the validity checker is never consulted, :66.)                                   */
int oa_discrete_geodesic(const oa_model* M, const double* from, const double* to, double delta_, double lambda_,
                         int use_fd, int max_states, double* states, int* n_states) {
  const int n = OA_DOF * M->n_arms;
  int ns = 0;
  memcpy(states, from, sizeof(double) * n);
  ns = 1;
  const double tolerance = delta_;
  double dist, step, total = 0;
  if ((dist = oa_distance(from, to, n)) <= tolerance) { *n_states = ns; return 1; }
  const double max = dist * lambda_;
  double previous[OA_DOF * OA_MAX_ARMS], scratch[OA_DOF * OA_MAX_ARMS];
  memcpy(previous, from, sizeof(double) * n);
  do {
    oa_interpolate(previous, to, delta_ / dist, scratch, n);
    if (!oa_project(M, scratch, use_fd, NULL, NULL, NULL) || (step = oa_distance(previous, scratch, n)) > lambda_ * delta_)
      break;
    total += step;
    if (total > max) break;
    const double newDist = oa_distance(scratch, to, n);
    if (newDist >= dist) break;
    dist = newDist;
    memcpy(previous, scratch, sizeof(double) * n);
    if (ns >= max_states) break; /* out of room (not in the reference; mirrors the engine's cap) */
    memcpy(states + (size_t)ns * n, scratch, sizeof(double) * n);
    ++ns;
  } while (dist >= tolerance);
  *n_states = ns;
  return dist <= tolerance;
}
void oa_discrete_geodesic_batch(const oa_model* M, const double* from, const double* to, int64_t edges, double delta_,
                                double lambda_, int use_fd, int max_states, double* states, int32_t* n_states,
                                uint8_t* reached, int nthreads) {
  const int n = OA_DOF * M->n_arms;
  (void)nthreads;
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads > 0 ? nthreads : 1)
  for (int64_t e = 0; e < edges; ++e) {
    int ns = 0;
    int r = oa_discrete_geodesic(M, from + e * n, to + e * n, delta_, lambda_, use_fd, max_states,
                                 states + (size_t)e * max_states * n, &ns);
    n_states[e] = ns;
    reached[e] = (uint8_t)r;
  }
}
