// ConstraintFunction.hpp — header-only C++ host mirror of the reference's constraint interface over the C ABI
// (include/ccp.h, libccp.so).  Same class and method names, argument meaning and error behaviour as
//   include/closed_chain_motion_planner/base/constraints/ConstraintFunction.h:21-137   (KinematicChainConstraint)
//   include/closed_chain_motion_planner/kinematics/panda_model.h:7-23                   (ArmModel)
//   include/closed_chain_motion_planner/kinematics/panda_rbdl.h:8-77                    (PandaModel)
// so the reference's call sites (ConstrainedPlanningCommon.cpp:126-129, jy_ProjectedStateSpace.cpp:13,20,27,65)
// compile against it after the `Eigen::Ref<VectorXd>` -> `double*` change shown in INTEGRATION.md (Eigen and OMPL are
// not installed in this image; with them present, the adaptor in INTEGRATION.md restores the exact signatures).
// All arithmetic runs on the GPU; a single-state call is a batch of one.  No CPU fallback.
#pragma once

#include <array>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <new>
#include <stdexcept>
#include <string>
#include <vector>

#include "ccp.h"

namespace ccp {

struct Exception : std::runtime_error {  // stands in for ompl::Exception
  using std::runtime_error::runtime_error;
};

// panda_model.h:7-23 — only what the constraint reads.  t_wb is row-major 3x4 [R|p].
struct ArmModel {
  std::string name;
  int index = 0;
  std::array<double, 12> t_wb{{1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0}};
  std::array<double, 28> dh_offsets{};  // 7x4 (a, d, theta, alpha) calibration offsets, panda_rbdl.cpp:92-95
};
typedef std::shared_ptr<ArmModel> ArmModelPtr;

// grasping_point.cpp:11-16
inline std::array<double, 12> base_frame(int index) {
  static const double twb[3][12] = {{1, 0, 0, 0.0, 0, 1, 0, 0.3, 0, 0, 1, 1.006},
                                    {1, 0, 0, 0.0, 0, 1, 0, -0.3, 0, 0, 1, 1.006},
                                    {-1, 0, 0, 1.35, 0, -1, 0, 0.3, 0, 0, 1, 1.006}};
  std::array<double, 12> t;
  std::memcpy(t.data(), twb[index], sizeof twb[index]);
  return t;
}

// std::allocator over page-locked host memory (ccp_host_alloc): buffers handed to the host-buffer entry points should
// live in it, so that their copies overlap the kernels.  Pinning is slow — keep such vectors alive across calls.
template <class T>
struct PinnedAllocator {
  using value_type = T;
  PinnedAllocator() = default;
  template <class U>
  PinnedAllocator(const PinnedAllocator<U>&) {}
  // small blocks stay pageable: pinning costs more than it saves below the size where the host path chunks a batch
  static constexpr std::size_t kPinFrom = 1u << 20;
  T* allocate(std::size_t count) {
    void* p = nullptr;
    if (count * sizeof(T) < kPinFrom) {
      p = std::malloc(count ? count * sizeof(T) : 1);
      if (!p) throw std::bad_alloc();
    } else if (ccp_host_alloc(&p, count * sizeof(T)) != CCP_OK || !p) {
      throw std::bad_alloc();
    }
    return static_cast<T*>(p);
  }
  void deallocate(T* p, std::size_t count) {
    if (count * sizeof(T) < kPinFrom) std::free(p);
    else ccp_host_free(p);
  }
  template <class U>
  bool operator==(const PinnedAllocator<U>&) const { return true; }
  template <class U>
  bool operator!=(const PinnedAllocator<U>&) const { return false; }
};
template <class T>
using PinnedVector = std::vector<T, PinnedAllocator<T>>;

struct ProjectBatchResult {
  PinnedVector<double> x;         // count x n, AOS
  PinnedVector<uint8_t> ok;       // project()'s return value per state
  PinnedVector<uint8_t> converged;
  PinnedVector<int32_t> iters;
  PinnedVector<double> resid;     // count x m
};

class KinematicChainConstraint {
 public:
  // ConstraintFunction.h:24 — links = ambient dimension (14, or 21 for the three-arm extension)
  explicit KinematicChainConstraint(unsigned int links, int device = 0) : n_(links), device_(device) {
    if (links != 14 && links != 21) throw Exception("KinematicChainConstraint: links must be 14 or 21");
  }
  ~KinematicChainConstraint() { ccp_destroy(h_); }
  KinematicChainConstraint(const KinematicChainConstraint&) = delete;
  KinematicChainConstraint& operator=(const KinematicChainConstraint&) = delete;

  unsigned int getAmbientDimension() const { return n_; }
  unsigned int getCoDimension() const { return 2 * (n_ / 7 - 1); }

  // ConstraintFunction.h:122-126
  void setArmModels(const ArmModelPtr& arm1, const ArmModelPtr& arm2, const ArmModelPtr& arm3 = nullptr) {
    const ArmModelPtr arms[3] = {arm1, arm2, arm3};
    const int k = (int)n_ / 7;
    if (!arm1 || !arm2 || (k == 3) != (arm3 != nullptr)) throw Exception("setArmModels: wrong number of arms");
    ccp_model_desc d;
    int32_t idx[3] = {0, 0, 0};
    if (ccp_default_model(k, idx, &d) != CCP_OK) throw Exception("ccp_default_model failed");
    for (int a = 0; a < k; ++a) {
      for (int i = 0; i < 7; ++i) {
        d.arm[a].dh_a[i] += arms[a]->dh_offsets[4 * i + 0];
        d.arm[a].dh_d[i] += arms[a]->dh_offsets[4 * i + 1];
        d.arm[a].dh_theta_offset[i] = arms[a]->dh_offsets[4 * i + 2];
        d.arm[a].dh_alpha[i] += arms[a]->dh_offsets[4 * i + 3];
      }
      std::memcpy(d.arm[a].t_wb, arms[a]->t_wb.data(), sizeof d.arm[a].t_wb);
    }
    ccp_destroy(h_);
    h_ = nullptr;
    if (ccp_create(&d, device_, &h_) != CCP_OK) throw Exception(std::string("ccp_create: ") + ccp_last_error(nullptr));
  }

  // ConstraintFunction.h:31-40
  void setInitialPosition(const double* init_joint) { check(ccp_set_reference(need(), init_joint)); }

  // ConstraintFunction.h:104-112 (throws like the reference's ompl::Exception)
  void setTolerance(const double tolerance1, const double tolerance2) {
    if (tolerance1 <= 0 || tolerance2 <= 0)
      throw Exception("ompl::base::Constraint::setProjectionTolerance(): tolerance must be positive.");
    check(ccp_set_tolerance(need(), tolerance1, tolerance2));
  }
  // The reference's setMaxIterations(1000) (ConstrainedPlanningCommon.cpp:129) sets an OMPL base member the loop
  // never reads (the loop uses the private 250, ConstraintFunction.h:26): kept as a no-op.
  void setMaxIterations(unsigned int) {}
  // damping (lambda^2 of a damped-least-squares step) and clamp (clamp every iterate to the limits) are opt-in modes
  // the reference does not have; parity runs keep them at 0
  void setOptions(double step, int max_iter, double joint_margin, double damping = 0.0, bool clamp = false) {
    ccp_options o{step, max_iter, clamp ? 1 : 0, joint_margin, damping};
    check(ccp_set_options(need(), &o));
  }

  // ConstraintFunction.h:84-102
  void function(const double* x, double* out) const { check(ccp_function_batch_host(need(), x, 1, out)); }
  // ompl::base::Constraint::jacobian (called at ConstraintFunction.h:70): out is m x n, row-major
  void jacobian(const double* x, double* out) const { check(ccp_jacobian_batch_host(need(), x, 1, out)); }
  // ConstraintFunction.h:57-82: x is updated in place (also on failure); returns converged && jointValid
  bool project(double* x) const {
    uint8_t ok = 0;
    check(ccp_project_batch_host(need(), x, 1, x, &ok, nullptr, nullptr, nullptr));
    return ok != 0;
  }
  // ConstraintFunction.h:114-120
  bool isSatisfied(const double* x) const {
    double f[4];
    function(x, f);
    double t1, t2;
    check(ccp_get_options(need(), nullptr, &t1, &t2));
    for (unsigned k = 0; k < getCoDimension(); k += 2)
      if (!std::isfinite(f[k]) || !std::isfinite(f[k + 1]) || !(f[k] <= t1) || !(f[k + 1] <= t2)) return false;
    return true;
  }
  // ConstraintFunction.h:43-55 (pure comparisons; no arithmetic to offload)
  bool jointValid(const double* q) const {
    static const double lb[7] = {-2.8973, -1.7628, -2.8973, -3.0718, -2.8973, -0.0175, -2.8973};
    static const double ub[7] = {2.8973, 1.7628, 2.8973, -0.0698, 2.8973, 3.7525, 2.8973};
    ccp_options o;
    check(ccp_get_options(need(), &o, nullptr, nullptr));
    for (unsigned j = 0; j < n_; ++j) {
      if (q[j] < lb[j % 7] + o.joint_margin) return false;
      if (q[j] > ub[j % 7] - o.joint_margin) return false;
    }
    return true;
  }

  // ---- batched entry points (north star) ----
  // Host states (count x n, AOS): copies in, projects on the GPU, copies out.  The result lives in page-locked
  // memory; a caller with batch after batch passes the previous result back in (second form) so it is pinned once.
  ProjectBatchResult projectBatch(const double* states, int64_t count) const {
    ProjectBatchResult r;
    projectBatch(states, count, r);
    return r;
  }
  void projectBatch(const double* states, int64_t count, ProjectBatchResult& r) const {
    const unsigned m = getCoDimension();
    r.x.resize((size_t)count * n_);
    r.ok.resize(count);
    r.converged.resize(count);
    r.iters.resize(count);
    r.resid.resize((size_t)count * m);
    check(ccp_project_batch_host(need(), states, count, r.x.data(), r.ok.data(), r.converged.data(), r.iters.data(),
                                 r.resid.data()));
  }
  // Streaming form for a caller with batch after batch of host states (caller-owned buffers, page-locked for full
  // overlap; any output but x_out may be null): submit batch k + 1 before waiting for batch k and the GPU never idles
  // on a batch's stragglers or copies.  At most two tickets are outstanding.
  int64_t submitBatch(const double* states, int64_t count, double* x_out, uint8_t* ok, uint8_t* converged = nullptr,
                      int32_t* iters = nullptr, double* resid = nullptr) const {
    int64_t ticket = 0;
    check(ccp_project_batch_host_submit(need(), states, count, x_out, ok, converged, iters, resid, &ticket));
    return ticket;
  }
  void waitBatch(int64_t ticket) const { check(ccp_project_batch_host_wait(need(), ticket)); }
  // Device states, asynchronous on `stream` (cudaStream_t as void*); any output may be null.
  void projectBatchDevice(const double* seeds_dev, int64_t count, ccp_layout layout, double* x_out_dev, uint8_t* ok_dev,
                          uint8_t* converged_dev, int32_t* iters_dev, double* resid_dev, double* compact_dev,
                          int64_t* n_ok_dev, void* stream) const {
    check(ccp_project_batch(need(), seeds_dev, count, layout, x_out_dev, ok_dev, converged_dev, iters_dev, resid_dev,
                            compact_dev, n_ok_dev, stream));
  }

  // Pipelined form for a caller that projects batch after batch on one stream (the batched sampler): the samples
  // still iterating when a batch runs dry are carried into the next launch instead of idling the GPU on them.
  // Per-seed outputs are complete after the next non-pipelined projection or flushProjections().
  void projectBatchDevicePipelined(const double* seeds_dev, int64_t count, ccp_layout layout, double* x_out_dev,
                                   uint8_t* ok_dev, uint8_t* converged_dev, int32_t* iters_dev, double* resid_dev,
                                   double* compact_dev, int64_t* n_ok_dev, void* stream) const {
    check(ccp_project_batch_pipelined(need(), seeds_dev, count, layout, x_out_dev, ok_dev, converged_dev, iters_dev,
                                      resid_dev, compact_dev, n_ok_dev, stream));
  }
  void flushProjections(double* compact_dev, int64_t* n_ok_dev, void* stream) const {
    check(ccp_project_flush(need(), compact_dev, n_ok_dev, stream));
  }
  // Multi-GPU sampler: make the projection kernel store every converged state into all ranks' pools
  // (peer-mapped device memory) — the all-gather of the states fused into the kernel.  world = 0 switches it off.
  void setGatherPeers(int world, int rank, const uint64_t* pool_dev_ptrs, int64_t capacity) const {
    check(ccp_set_gather_peers(need(), world, rank, pool_dev_ptrs, capacity));
  }

  ccp_handle* handle() const { return h_; }

 private:
  ccp_handle* need() const {
    if (!h_) throw Exception("KinematicChainConstraint: setArmModels() first");
    return h_;
  }
  void check(int rc) const {
    if (rc != CCP_OK) throw Exception(std::string("ccp error: ") + ccp_last_error(h_));
  }
  unsigned int n_;
  int device_;
  ccp_handle* h_ = nullptr;
};

typedef std::shared_ptr<KinematicChainConstraint> ChainConstraintPtr;

// panda_rbdl.h:43-77 — FK and geometric Jacobian of one arm in its base frame, on the GPU.
class PandaModel {
 public:
  static constexpr int kDof = 7;
  explicit PandaModel(int device = 0) : c_(14, device) { initModel(nullptr); }
  void initModel(const double* dh /* 7x4 row-major or null */) {
    auto a = std::make_shared<ArmModel>();
    auto b = std::make_shared<ArmModel>();
    if (dh) {
      std::memcpy(a->dh_offsets.data(), dh, sizeof(double) * 28);
      std::memcpy(b->dh_offsets.data(), dh, sizeof(double) * 28);
    }
    c_.setArmModels(a, b);
  }
  int getDof() { return kDof; }
  // panda_rbdl.cpp:35-42: row-major 3x4 [R|p] of the EE frame in the arm's base frame
  std::array<double, 12> getTransform(const double* q) const {
    std::array<double, 12> T;
    if (ccp_arm_fk_batch_host(c_.handle(), 0, q, 1, T.data(), nullptr) != CCP_OK)
      throw Exception(std::string("getTransform: ") + ccp_last_error(c_.handle()));
    return T;
  }
  std::array<double, 9> getRotation(const double* q) const {
    auto T = getTransform(q);
    return {{T[0], T[1], T[2], T[4], T[5], T[6], T[8], T[9], T[10]}};
  }
  std::array<double, 3> getTranslation(const double* q) const {
    auto T = getTransform(q);
    return {{T[3], T[7], T[11]}};
  }
  // panda_rbdl.cpp:9-22: 6x7 row-major, rows [linear(3); angular(3)]
  std::array<double, 42> getJacobianMatrix(const double* q) const {
    std::array<double, 42> J;
    if (ccp_arm_fk_batch_host(c_.handle(), 0, q, 1, nullptr, J.data()) != CCP_OK)
      throw Exception(std::string("getJacobianMatrix: ") + ccp_last_error(c_.handle()));
    return J;
  }
  // 7x2 row-major (low, high), panda_rbdl.cpp:44-55
  std::array<double, 14> getJointLimit() const {
    return {{-2.8973, 2.8973, -1.7628, 1.7628, -2.8973, 2.8973, -3.0718, -0.0698, -2.8973, 2.8973, -0.0175, 3.7525,
             -2.8973, 2.8973}};
  }

 private:
  KinematicChainConstraint c_;
};

}  // namespace ccp
