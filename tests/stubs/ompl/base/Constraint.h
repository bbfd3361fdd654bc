// MINIMAL stand-in of ompl/base/Constraint.h (OMPL >= 1.4 constrained-planning API): the four virtuals the reference's
// KinematicChainConstraint overrides or inherits (ConstraintFunction.h:57,84,114 and the jacobian() called at :70), and
// the State* wrapper the samplers call.  Test infrastructure only.
#pragma once
#include <Eigen/Core>
#include <ompl/base/spaces/constraint/ConstrainedStateSpace.h>
#include <ompl/util/Exception.h>
namespace ompl {
namespace base {
class Constraint {
 public:
  Constraint(unsigned ambientDim, unsigned coDim, double tolerance = 1e-4)
      : n_(ambientDim), k_(coDim), tolerance_(tolerance), maxIterations_(50) {}
  virtual ~Constraint() = default;
  virtual void function(const Eigen::Ref<const Eigen::VectorXd>& x, Eigen::Ref<Eigen::VectorXd> out) const = 0;
  virtual void jacobian(const Eigen::Ref<const Eigen::VectorXd>& x, Eigen::Ref<Eigen::MatrixXd> out) const = 0;
  virtual bool project(Eigen::Ref<Eigen::VectorXd> x) const = 0;
  bool project(State* state) const { return project(*state->as<ConstrainedStateSpace::StateType>()); }
  virtual bool isSatisfied(const Eigen::Ref<const Eigen::VectorXd>& x) const = 0;
  unsigned getAmbientDimension() const { return n_; }
  unsigned getCoDimension() const { return k_; }
  void setMaxIterations(unsigned m) { maxIterations_ = m; }
 protected:
  const unsigned n_, k_;
  double tolerance_;
  unsigned maxIterations_;
};
}  // namespace base
}  // namespace ompl
