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
class jy_ProjectedStateSampler {
 public:
  jy_ProjectedStateSampler(const jy_ProjectedStateSpace* space, int64_t pool_size, uint64_t rng_seed);
  void sampleUniform(double* state) {  // .cpp:10-15
    while (pos_ >= count_) refill();
    std::memcpy(state, &pool_[(size_t)pos_++ * n_], sizeof(double) * n_);
  }
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
  bool first_success(int mode, const double* center, double spread, double* state, int tries);
  const jy_ProjectedStateSpace* space_;
  unsigned int n_;
  int64_t pool_size_, count_ = 0, pos_ = 0, next_index_ = 0, refills_ = 0;
  uint64_t rng_seed_;
  std::vector<double> pool_;
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

inline void jy_ProjectedStateSampler::refill() {
  ccp_sampler_args a;
  a.rng_seed = rng_seed_;
  a.first_index = next_index_;
  a.mode = 0;
  a.wrap_bounds = 1;  // space_->enforceBounds(state), .cpp:14
  a.distance = 0.0;
  a.near_host = nullptr;
  next_index_ += pool_size_;
  int64_t n_ok = 0;
  ccp_handle* h = space_->getConstraint()->handle();
  if (ccp_sample_project_batch_host(h, &a, pool_size_, nullptr, nullptr, nullptr, pool_.data(), &n_ok) != CCP_OK)
    throw Exception(std::string("jy_ProjectedStateSampler: ") + ccp_last_error(h));
  count_ = n_ok;
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

// The goal sampler's per-arm IK loop (jy_ConstrainedValidStateSampler.h:63-189) for a batch of targets: `restarts`
// solves per target (restart 0 from q_ref if given, the rest from N(mid-range, 0.3) clipped to the limits); the seeded
// solution wins, else the successful restart nearest to q_ref.  targets: n_targets x 12 (row-major 3x4 EE pose in the
// arm's base frame); q_best: n_targets x 7.
inline void ikSampleBatch(const KinematicChainConstraint& c, int arm, const double* targets, int64_t n_targets, int restarts,
                          uint64_t rng_seed, const double* q_ref, double* q_best, uint8_t* ok, int32_t* n_success = nullptr,
                          double sigma = 0.3) {
  if (ccp_ik_sample_batch_host(c.handle(), arm, targets, n_targets, restarts, rng_seed, sigma, q_ref, nullptr, q_best, ok,
                               n_success) != CCP_OK)
    throw Exception(std::string("ikSampleBatch: ") + ccp_last_error(c.handle()));
}

}  // namespace ccp
