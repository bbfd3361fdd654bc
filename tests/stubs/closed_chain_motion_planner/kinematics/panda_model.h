// MINIMAL stand-in of the reference's kinematics/panda_model.h:7-23 (only the members the adaptor reads).
#pragma once
#include <Eigen/Core>
#include <memory>
#include <string>
struct ArmModel {
  std::string name;
  int index = 0;
  Eigen::Isometry3d t_wb;
};
typedef std::shared_ptr<ArmModel> ArmModelPtr;
