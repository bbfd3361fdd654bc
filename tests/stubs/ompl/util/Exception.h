// MINIMAL stand-in of ompl/util/Exception.h (test infrastructure only).
#pragma once
#include <stdexcept>
#include <string>
namespace ompl {
class Exception : public std::runtime_error {
 public:
  explicit Exception(const std::string& what) : std::runtime_error(what) {}
};
}  // namespace ompl
