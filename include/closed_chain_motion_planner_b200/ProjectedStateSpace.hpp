// ProjectedStateSpace.hpp — header-only C++ host mirror of the reference's state-space seam over the C ABI
// (include/ccp.h, libccp.so), with the reference's class and method names:
//   include/closed_chain_motion_planner/kinematics/KinematicChain.h:69-174   KinematicChainSpace (bounds, enforceBounds,
//                                                                            distance, interpolate)
//   include/closed_chain_motion_planner/base/jy_ProjectedStateSpace.h:17-76  jy_ProjectedStateSampler, jy_ProjectedStateSpace
//   src/base/jy_ProjectedStateSpace.cpp:10-29                                sampleUniform / sampleUniformNear / sampleGaussian
//   src/base/jy_ProjectedStateSpace.cpp:32-96                                discreteGeodesic
//   include/closed_chain_motion_planner/base/jy_ConstrainedValidStateSampler.h:63-189  the goal sampler's per-arm IK loop
// States are plain `double*` of length n = 7K (the reference's states are Eigen::Map views of exactly that,
// KinematicChain.cpp:97); with OMPL present the adaptor in INTEGRATION.md restores the ob::State signatures.
// Every projection runs on the GPU through libccp.so; nothing here includes a CUDA header.
#pragma once

#include <cmath>
#include <cstdint>
#include <memory>
#include <vector>

#include "ccp.h"
#include "closed_chain_motion_planner_b200/ConstraintFunction.hpp"

namespace ccp {

// KinematicChain.h:69-174 — RealVectorStateSpace with the Panda limits per arm and the reference's wrap semantics.
class KinematicChainSpace {
 public:
  explicit KinematicChainSpace(unsigned int numLinks) : n_(numLinks) {
    if (numLinks % 7 != 0 || numLinks < 14) throw Exception("KinematicChainSpace: numLinks must be 14 or 21");
    static const double lb[7] = {-2.8973, -1.7628, -2.8973, -3.0718, -2.8973, -0.0175, -2.8973};
    static const double ub[7] = {2.8973, 1.7628, 2.8973, -0.0698, 2.8973, 3.7525, 2.8973};
    for (unsigned i = 0; i < n_; ++i) {  // KinematicChain.h:77-99
      low_.push_back(lb[i % 7]);
      high_.push_back(ub[i % 7]);
    }
  }
  unsigned int getDimension() const { return n_; }
  const std::vector<double>& low() const { return low_; }
  const std::vector<double>& high() const { return high_; }

  // KinematicChain.h:118-130: fmod wrap into [-pi, pi) — NOT a clamp; joint 6 values in (pi, 3.7525] become negative
  void enforceBounds(double* state) const {
    const double pi = 3.14159265358979323846;
    for (unsigned i = 0; i < n_; ++i) {
      double v = std::fmod(state[i], 2.0 * pi);
      if (v < -pi) v += 2.0 * pi;
      else if (v >= pi) v -= 2.0 * pi;
      state[i] = v;
    }
  }
  bool equalStates(const double* a, const double* b) const {  // KinematicChain.h:132-143
    for (unsigned i = 0; i < n_; ++i)
      if (std::fabs(a[i] - b[i]) > 1e-10) return false;
    return true;
  }
  double distance(const double* a, const double* b) const {  // RealVectorStateSpace::distance (inherited)
    double acc = 0.0;
    for (unsigned i = 0; i < n_; ++i) acc = std::fma(a[i] - b[i], a[i] - b[i], acc);
    return std::sqrt(acc);
  }
  // KinematicChain.h:145-171: the short way round when |to - from| > pi, wrapped back into [-pi, pi]
  void interpolate(const double* from, const double* to, double t, double* state) const {
    const double pi = 3.14159265358979323846;
    for (unsigned i = 0; i < n_; ++i) {
      double diff = to[i] - from[i];
      if (std::fabs(diff) <= pi) {
        state[i] = std::fma(diff, t, from[i]);
      } else {
        diff = (diff > 0.0) ? 2.0 * pi - diff : -2.0 * pi - diff;
        double v = std::fma(-diff, t, from[i]);
        if (v > pi) v -= 2.0 * pi;
        else if (v < -pi) v += 2.0 * pi;
        state[i] = v;
      }
    }
  }

 private:
  unsigned int n_;
  std::vector<double> low_, high_;
};
typedef std::shared_ptr<KinematicChainSpace> KinematicChainSpacePtr;

struct GeodesicBatchResult {
  std::vector<uint8_t> reached;   // discreteGeodesic's return value per edge
  std::vector<int32_t> n_states;  // valid states per edge (>= 1; states[e][0] = from)
  std::vector<double> states;     // edges x max_states x n
  std::vector<int32_t> iters;     // Newton iterations spent per edge
  int max_states = 0;
};

class jy_ProjectedStateSpace;

// jy_ProjectedStateSpace.h:17-39 + .cpp:10-29.  sampleUniform pops from a pool of PROJECTED states that one
// ccp_sample_project_batch call refills (seed kernel -> projection -> enforceBounds wrap -> compaction).  The
// reference ignores project()'s return value (.cpp:13) and hands failed projections to the planner; the pool only
// holds states with project() == true.
//   setReturnFailed(true) reproduces the reference's distribution instead: every projected seed is returned, failed
//   projections included (their last iterate, wrapped), for planners that rely on rejecting them later.
//   setPrefetch(true): the NEXT pool is projected on the GPU while the planner consumes the current one
//   (ccp_host_batch_submit with sampler arguments: seeds generated on the device, only the ok states come back).
class jy_ProjectedStateSampler {
 public:
  jy_ProjectedStateSampler(const jy_ProjectedStateSpace* space, int64_t pool_size, uint64_t rng_seed);
  ~jy_ProjectedStateSampler();
  void sampleUniform(double* state) {  // .cpp:10-15
    while (pos_ >= count_) refill();
    std::memcpy(state, &pool_[(size_t)pos_++ * n_], sizeof(double) * n_);
  }
  void setReturnFailed(bool on) { return_failed_ = on; }
  void setPrefetch(bool on) { prefetch_ = on; }
  // .cpp:17-22 / :24-29: a small batch around `near`; the lowest-numbered draw that projected wins, else the wrapped
  // last iterate of the first draw (what the reference would have returned).  Returns whether a draw projected.
  bool sampleUniformNear(double* state, const double* near, double distance, int tries = 32) {
    return first_success(1, near, distance, state, tries);
  }
  bool sampleGaussian(double* state, const double* mean, double stdDev, int tries = 32) {
    return first_success(2, mean, stdDev, state, tries);
  }
  int64_t refills() const { return refills_; }

 private:
  void refill();
  void submit_next();
  bool first_success(int mode, const double* center, double spread, double* state, int tries);
  const jy_ProjectedStateSpace* space_;
  unsigned int n_;
  int64_t pool_size_, count_ = 0, pos_ = 0, next_index_ = 0, refills_ = 0;
  uint64_t rng_seed_;
  std::vector<double> pool_;
  bool return_failed_ = false, prefetch_ = false;
  double* next_pool_ = nullptr;  // page-locked (ccp_host_alloc): the pool being projected ahead
  int64_t next_ticket_ = 0;
  bool next_failed_ = false;     // the batch in flight was submitted with return_failed_ set
};
typedef std::shared_ptr<jy_ProjectedStateSampler> jy_ProjectedStateSamplerPtr;

// jy_ProjectedStateSpace.h:41-76 (ompl::base::ProjectedStateSpace with the reference's traversal)
class jy_ProjectedStateSpace {
 public:
  jy_ProjectedStateSpace(const KinematicChainSpacePtr& ambientSpace, const ChainConstraintPtr& constraint)
      : space_(ambientSpace), constraint_(constraint) {
    if (!ambientSpace || !constraint || ambientSpace->getDimension() != constraint->getAmbientDimension())
      throw Exception("jy_ProjectedStateSpace: space and constraint dimensions differ");
  }
  void setDelta(double delta) {  // ConstrainedPlanningCommon.cpp:118
    if (!(delta > 0)) throw Exception("ompl::base::ConstrainedStateSpace::setDelta(): delta must be positive.");
    delta_ = delta;
  }
  void setLambda(double lambda) {  // ConstrainedPlanningCommon.cpp:119
    if (!(lambda > 1)) throw Exception("ompl::base::ConstrainedStateSpace::setLambda(): lambda must be > 1.");
    lambda_ = lambda;
  }
  double getDelta() const { return delta_; }
  double getLambda() const { return lambda_; }
  const ChainConstraintPtr& getConstraint() const { return constraint_; }
  const KinematicChainSpacePtr& getSpace() const { return space_; }
  double distance(const double* a, const double* b) const { return space_->distance(a, b); }

  // jy_ProjectedStateSpace.h:41-50
  jy_ProjectedStateSamplerPtr allocStateSampler(int64_t pool_size = 65536, uint64_t rng_seed = 0) const {
    return std::make_shared<jy_ProjectedStateSampler>(this, pool_size, rng_seed);
  }
  jy_ProjectedStateSamplerPtr allocDefaultStateSampler() const { return allocStateSampler(); }

  // discreteGeodesic(from, to, interpolate = true, &geodesic) for MANY edges in one kernel launch (.cpp:32-96).
  // The reference's state-validity call (MoveIt collision, :66) is not part of it: validate the returned states.
  GeodesicBatchResult discreteGeodesicBatch(const double* from, const double* to, int64_t edges, int max_states = 64) const {
    GeodesicBatchResult r;
    const unsigned n = space_->getDimension();
    r.max_states = max_states;
    r.reached.resize(edges);
    r.n_states.resize(edges);
    r.iters.resize(edges);
    r.states.resize((size_t)edges * max_states * n);
    int rc = ccp_geodesic_batch_host(constraint_->handle(), from, to, edges, delta_, lambda_, max_states, r.states.data(),
                                     r.n_states.data(), r.reached.data(), r.iters.data());
    if (rc != CCP_OK) throw Exception(std::string("discreteGeodesicBatch: ") + ccp_last_error(constraint_->handle()));
    return r;
  }
  // One edge, reference signature (.cpp:32): returns whether `to` was reached; *geodesic gets the states incl. `from`.
  bool discreteGeodesic(const double* from, const double* to, bool interpolate = true,
                        std::vector<std::vector<double>>* geodesic = nullptr, int max_states = 256) const {
    if (!interpolate)
      throw Exception("discreteGeodesic(interpolate = false) needs the host collision checker: validate the returned states");
    GeodesicBatchResult r = discreteGeodesicBatch(from, to, 1, max_states);
    if (geodesic) {
      const unsigned n = space_->getDimension();
      geodesic->clear();
      for (int i = 0; i < r.n_states[0]; ++i)
        geodesic->emplace_back(r.states.begin() + (size_t)i * n, r.states.begin() + (size_t)(i + 1) * n);
    }
    return r.reached[0] != 0;
  }

 private:
  KinematicChainSpacePtr space_;
  ChainConstraintPtr constraint_;
  double delta_ = 0.25, lambda_ = 2.0;  // ConstrainedPlanningCommon.cpp:118-119
};
typedef std::shared_ptr<jy_ProjectedStateSpace> jy_ProjectedStateSpacePtr;

inline jy_ProjectedStateSampler::jy_ProjectedStateSampler(const jy_ProjectedStateSpace* space, int64_t pool_size,
                                                          uint64_t rng_seed)
    : space_(space), n_(space->getSpace()->getDimension()), pool_size_(pool_size), rng_seed_(rng_seed) {
  if (pool_size < 1) throw Exception("jy_ProjectedStateSampler: pool_size must be positive");
  pool_.resize((size_t)pool_size * n_);
}

inline jy_ProjectedStateSampler::~jy_ProjectedStateSampler() {
  if (next_ticket_) ccp_host_batch_wait(space_->getConstraint()->handle(), next_ticket_, nullptr);
  ccp_host_free(next_pool_);
}

// prefetch: enqueue the projection of the next pool_size_ seeds of the stream; returns at once
inline void jy_ProjectedStateSampler::submit_next() {
  ccp_handle* h = space_->getConstraint()->handle();
  if (!next_pool_ && ccp_host_alloc((void**)&next_pool_, sizeof(double) * (size_t)pool_size_ * n_) != CCP_OK)
    throw Exception("jy_ProjectedStateSampler: ccp_host_alloc failed");
  ccp_sampler_args a;
  a.rng_seed = rng_seed_;
  a.first_index = next_index_;
  a.mode = 0;
  a.wrap_bounds = 1;
  a.distance = 0.0;
  a.near_host = nullptr;
  next_index_ += pool_size_;
  ccp_host_batch b;
  std::memset(&b, 0, sizeof b);
  b.sampler = &a;
  b.count = pool_size_;
  next_failed_ = return_failed_;
  if (return_failed_) b.x_out_host = next_pool_;
  else {
    b.compact_host = next_pool_;
    b.compact_capacity = pool_size_;
  }
  if (ccp_host_batch_submit(h, &b, &next_ticket_) != CCP_OK)
    throw Exception(std::string("jy_ProjectedStateSampler: ") + ccp_last_error(h));
}

inline void jy_ProjectedStateSampler::refill() {
  ccp_handle* h = space_->getConstraint()->handle();
  if (prefetch_ || next_ticket_) {  // (a pool still in flight after setPrefetch(false) is consumed first)
    if (!next_ticket_) submit_next();
    int64_t n_ok = -1;
    if (ccp_host_batch_wait(h, next_ticket_, &n_ok) != CCP_OK)
      throw Exception(std::string("jy_ProjectedStateSampler: ") + ccp_last_error(h));
    next_ticket_ = 0;
    count_ = next_failed_ ? pool_size_ : n_ok;  // as submitted, whatever setReturnFailed says by now
    std::memcpy(pool_.data(), next_pool_, sizeof(double) * (size_t)count_ * n_);
    pos_ = 0;
    ++refills_;
    if (prefetch_) submit_next();  // the GPU works on the next pool while the planner consumes this one
    return;
  }
  ccp_sampler_args a;
  a.rng_seed = rng_seed_;
  a.first_index = next_index_;
  a.mode = 0;
  a.wrap_bounds = 1;  // space_->enforceBounds(state), .cpp:14
  a.distance = 0.0;
  a.near_host = nullptr;
  next_index_ += pool_size_;
  int64_t n_ok = 0;
  const int rc = return_failed_
                     ? ccp_sample_project_batch_host(h, &a, pool_size_, pool_.data(), nullptr, nullptr, nullptr, nullptr)
                     : ccp_sample_project_batch_host(h, &a, pool_size_, nullptr, nullptr, nullptr, pool_.data(), &n_ok);
  if (rc != CCP_OK) throw Exception(std::string("jy_ProjectedStateSampler: ") + ccp_last_error(h));
  count_ = return_failed_ ? pool_size_ : n_ok;
  pos_ = 0;
  ++refills_;
}

inline bool jy_ProjectedStateSampler::first_success(int mode, const double* center, double spread, double* state, int tries) {
  ccp_sampler_args a;
  a.rng_seed = rng_seed_;
  a.first_index = next_index_;
  a.mode = mode;
  a.wrap_bounds = 1;
  a.distance = spread;
  a.near_host = center;
  next_index_ += tries;
  std::vector<double> x((size_t)tries * n_);
  std::vector<uint8_t> ok(tries);
  ccp_handle* h = space_->getConstraint()->handle();
  if (ccp_sample_project_batch_host(h, &a, tries, x.data(), ok.data(), nullptr, nullptr, nullptr) != CCP_OK)
    throw Exception(std::string("jy_ProjectedStateSampler: ") + ccp_last_error(h));
  int first = 0;
  bool any = false;
  for (int i = 0; i < tries; ++i)
    if (ok[i]) {
      first = i;
      any = true;
      break;
    }
  std::memcpy(state, &x[(size_t)first * n_], sizeof(double) * n_);
  return any;
}

// Pool refill over EVERY GPU of the box from one C++ process (the reference planner is one process, main.cpp:27-63):
// one constraint per device with the same model, a peer group over their handles; sample(total) shards the seed stream
// over the devices, the projection kernels gather the converged states into every device's pool over NVLink, and the
// rows come back from device 0.  Returns the number of projected (ok, wrapped) states appended to `out`.
class MultiGpuProjectedSampler {
 public:
  MultiGpuProjectedSampler(const std::vector<ChainConstraintPtr>& constraints, int64_t seeds_per_refill, uint64_t rng_seed)
      : constraints_(constraints), total_(seeds_per_refill), rng_seed_(rng_seed) {
    if (constraints.empty()) throw Exception("MultiGpuProjectedSampler: no constraints");
    std::vector<ccp_handle*> hs;
    for (auto& c : constraints) hs.push_back(c->handle());
    n_ = constraints[0]->getAmbientDimension();
    const int64_t per_rank = (total_ + (int64_t)hs.size() - 1) / (int64_t)hs.size();
    capacity_ = per_rank * 2 / 5 + 64;  // uniform seeds succeed on ~21-23 %; overflow is reported, not silent
    if (capacity_ > per_rank) capacity_ = per_rank;
    if (ccp_peer_group_create(hs.data(), (int32_t)hs.size(), capacity_, &group_) != CCP_OK)
      throw Exception(std::string("MultiGpuProjectedSampler: ") + ccp_peer_group_last_error(nullptr));
  }
  ~MultiGpuProjectedSampler() { ccp_peer_group_destroy(group_); }
  MultiGpuProjectedSampler(const MultiGpuProjectedSampler&) = delete;
  MultiGpuProjectedSampler& operator=(const MultiGpuProjectedSampler&) = delete;
  int world() const { return ccp_peer_group_world(group_); }
  int64_t sample(std::vector<double>* out, std::vector<int64_t>* counts = nullptr) {
    ccp_sampler_args a;
    a.rng_seed = rng_seed_;
    a.first_index = next_index_;
    a.mode = 0;
    a.wrap_bounds = 1;
    a.distance = 0.0;
    a.near_host = nullptr;
    next_index_ += total_;
    std::vector<int64_t> cnt(world());
    if (ccp_peer_group_sample_project(group_, &a, total_, cnt.data()) != CCP_OK)
      throw Exception(std::string("MultiGpuProjectedSampler: ") + ccp_peer_group_last_error(group_));
    int64_t rows = 0;
    for (int64_t v : cnt) rows += v;
    const size_t at = out->size();
    out->resize(at + (size_t)rows * n_);
    int64_t got = 0;
    if (ccp_peer_group_gather_host(group_, 0, out->data() + at, rows, cnt.data(), &got) != CCP_OK)
      throw Exception(std::string("MultiGpuProjectedSampler: ") + ccp_peer_group_last_error(group_));
    if (counts) *counts = cnt;
    return got;
  }
  ccp_peer_group* group() const { return group_; }

 private:
  std::vector<ChainConstraintPtr> constraints_;
  ccp_peer_group* group_ = nullptr;
  unsigned int n_ = 0;
  int64_t total_, capacity_ = 0, next_index_ = 0;
  uint64_t rng_seed_;
};

// The goal sampler's per-arm IK loop (jy_ConstrainedValidStateSampler.h:63-189) for a batch of targets: `restarts`
// solves per target (restart 0 from q_ref if given, the rest from N(mid-range, 0.3) clipped to the limits); the seeded
// solution wins, else the successful restart nearest to q_ref (without q_ref: the lowest-numbered success; restarts that
// can no longer win are abandoned, n_success counts those that finished).  targets: n_targets x 12 (row-major 3x4 EE pose
// in the arm's base frame); q_best: n_targets x 7.
inline void ikSampleBatch(const KinematicChainConstraint& c, int arm, const double* targets, int64_t n_targets, int restarts,
                          uint64_t rng_seed, const double* q_ref, double* q_best, uint8_t* ok, int32_t* n_success = nullptr,
                          double sigma = 0.3) {
  if (ccp_ik_sample_batch_host(c.handle(), arm, targets, n_targets, restarts, rng_seed, sigma, q_ref, nullptr, q_best, ok,
                               n_success) != CCP_OK)
    throw Exception(std::string("ikSampleBatch: ") + ccp_last_error(c.handle()));
}

// jy_ValidStateSampler::sampleCalibGoal / sampleRandomGoal (jy_ConstrainedValidStateSampler.h:63-189) for a batch of object
// poses: T_obj n x 12 (row-major 3x4, world), t_o7 K x 12 (the grasp frames, ConstrainedPlanningCommon.cpp:105-111),
// q_ref n x 7K or nullptr; q_out n x 7K, ok n.  A pose is ok when every arm found an IK solution; the row is then a
// closed-chain goal configuration.  The reference's IKValid collision check stays with the caller.
inline void sampleGoalBatch(const KinematicChainConstraint& c, const double* T_obj, int64_t n, const double* t_o7,
                            const double* q_ref, double* q_out, uint8_t* ok, int restarts = 15, uint64_t rng_seed = 0,
                            double sigma = 0.3) {
  if (ccp_goal_sample_batch_host(c.handle(), T_obj, n, t_o7, q_ref, restarts, rng_seed, sigma, nullptr, q_out, ok) != CCP_OK)
    throw Exception(std::string("sampleGoalBatch: ") + ccp_last_error(c.handle()));
}

}  // namespace ccp
