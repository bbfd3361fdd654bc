// MINIMAL stand-in of OMPL's ConstrainedStateSpace::StateType: an OMPL state that is also an Eigen::Map of its values
// (test infrastructure only).
#pragma once
#include <Eigen/Core>
#include <vector>
namespace ompl {
namespace base {
class State {
 public:
  virtual ~State() = default;
  template <class T>
  T* as() {
    return static_cast<T*>(this);
  }
};
class ConstrainedStateSpace {
 public:
  class StateType : public State, public Eigen::Map<Eigen::VectorXd> {
   public:
    explicit StateType(unsigned n) : Eigen::Map<Eigen::VectorXd>(nullptr, n), store_(n) { p_ = store_.data(); }
   private:
    std::vector<double> store_;
  };
};
}  // namespace base
}  // namespace ompl
