// Compiles the OMPL adaptor (include/closed_chain_motion_planner_b200/ompl_adaptor/ConstraintFunction.h) against
// stand-in OMPL / Eigen headers (tests/stubs/) and drives it the way the reference does: through a pointer to
// ompl::base::Constraint (function / jacobian / project / isSatisfied), project(State *) as the samplers call it
// (jy_ProjectedStateSpace.cpp:13), and the new projectBatch on individually allocated states.
// usage: test_ompl_adaptor start.bin seeds.bin out.bin
#include <cstdio>
#include <vector>

#include "closed_chain_motion_planner_b200/ompl_adaptor/ConstraintFunction.h"

int main(int argc, char** argv) {
  if (argc < 4) return 2;
  double start[14];
  FILE* f = fopen(argv[1], "rb");
  if (!f || fread(start, 8, 14, f) != 14) return 3;
  fclose(f);
  f = fopen(argv[2], "rb");
  int64_t count = 0;
  if (!f || fread(&count, 8, 1, f) != 1) return 4;
  std::vector<double> seeds((size_t)count * 14);
  if (fread(seeds.data(), 8, seeds.size(), f) != seeds.size()) return 4;
  fclose(f);

  auto left = std::make_shared<ArmModel>();
  left->name = "panda_left";
  left->index = 0;
  left->t_wb.at(1, 3) = 0.3;  // grasping_point.cpp:11-16
  left->t_wb.at(2, 3) = 1.006;
  auto top = std::make_shared<ArmModel>();
  top->name = "panda_top";
  top->index = 2;
  top->t_wb.at(0, 0) = -1.0;
  top->t_wb.at(1, 1) = -1.0;
  top->t_wb.at(0, 3) = 1.35;
  top->t_wb.at(1, 3) = 0.3;
  top->t_wb.at(2, 3) = 1.006;

  ChainConstraintPtr constraint = std::make_shared<KinematicChainConstraint>(14);  // main.cpp:41
  constraint->setArmModels(left, top);                                             // ConstrainedPlanningCommon.cpp:126
  Eigen::VectorXd q0(14);
  for (int i = 0; i < 14; ++i) q0[i] = start[i];
  constraint->setInitialPosition(q0);   // :127
  constraint->setTolerance(0.001, 0.005);  // :128
  constraint->setMaxIterations(1000);      // :129 (the OMPL base member the loop never reads)
  bool threw = false;
  try {
    constraint->setTolerance(0.0, 1.0);
  } catch (const ompl::Exception&) {
    threw = true;
  }
  if (!threw) return 5;

  const ompl::base::Constraint* base = constraint.get();  // OMPL only ever sees the base class
  Eigen::VectorXd x(14), fx(2);
  for (int i = 0; i < 14; ++i) x[i] = seeds[i];
  Eigen::MatrixXd J(2, 14);
  base->function(x, fx);
  base->jacobian(x, J);
  const bool ok0 = base->project(x);
  const bool sat0 = base->isSatisfied(x), jv0 = constraint->jointValid(x);

  // project(State *) one state at a time, as jy_ProjectedStateSampler does, on the first 64 seeds
  const int S = 64;
  std::vector<double> single((size_t)S * 14);
  std::vector<uint8_t> single_ok(S);
  for (int s = 0; s < S; ++s) {
    ompl::base::ConstrainedStateSpace::StateType st(14);
    for (int i = 0; i < 14; ++i) st[i] = seeds[(size_t)s * 14 + i];
    single_ok[s] = base->project(&st);
    for (int i = 0; i < 14; ++i) single[(size_t)s * 14 + i] = st[i];
  }
  // projectBatch on individually allocated states
  std::vector<std::unique_ptr<ompl::base::ConstrainedStateSpace::StateType>> owned;
  std::vector<ompl::base::State*> states;
  for (int64_t s = 0; s < count; ++s) {
    owned.emplace_back(new ompl::base::ConstrainedStateSpace::StateType(14));
    for (int i = 0; i < 14; ++i) (*owned.back())[i] = seeds[(size_t)s * 14 + i];
    states.push_back(owned.back().get());
  }
  std::vector<uint8_t> ok;
  constraint->projectBatch(states, ok);

  FILE* o = fopen(argv[3], "wb");
  if (!o) return 7;
  fwrite(fx.data(), 8, 2, o);
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 14; ++j) fwrite(&J(i, j), 8, 1, o);
  fwrite(x.data(), 8, 14, o);
  uint8_t flags[3] = {(uint8_t)ok0, (uint8_t)sat0, (uint8_t)jv0};
  fwrite(flags, 1, 3, o);
  fwrite(single.data(), 8, single.size(), o);
  fwrite(single_ok.data(), 1, S, o);
  for (int64_t s = 0; s < count; ++s) fwrite(owned[s]->data(), 8, 14, o);
  fwrite(ok.data(), 1, ok.size(), o);
  fclose(o);
  return 0;
}
