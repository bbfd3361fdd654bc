// C++ host-side smoke of the reference-shaped interface (include/closed_chain_motion_planner_b200/ConstraintFunction.hpp).
// Mirrors ConstrainedProblem::setConstrainedOptions (ConstrainedPlanningCommon.cpp:116-132) and the sampler's
// project call (jy_ProjectedStateSpace.cpp:13).  usage: test_constraint seeds.bin out.bin  (built and checked by
// tests/test_cpp_host_gpu.py, which compares out.bin with the CPU oracle bit for bit).
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <vector>

#include "closed_chain_motion_planner_b200/ConstraintFunction.hpp"

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 3;
  int64_t count = 0;
  double start[14];
  if (fread(&count, sizeof count, 1, f) != 1 || fread(start, sizeof(double), 14, f) != 14) return 4;
  std::vector<double> seeds((size_t)count * 14);
  if (fread(seeds.data(), sizeof(double), seeds.size(), f) != seeds.size()) return 4;
  fclose(f);

  auto left = std::make_shared<ccp::ArmModel>();
  left->name = "panda_left";
  left->index = 0;
  left->t_wb = ccp::base_frame(0);
  auto top = std::make_shared<ccp::ArmModel>();
  top->name = "panda_top";
  top->index = 2;
  top->t_wb = ccp::base_frame(2);

  auto constraint = std::make_shared<ccp::KinematicChainConstraint>(14);
  bool threw = false;
  try {
    double z[14] = {0};
    constraint->project(z);  // before setArmModels
  } catch (const ccp::Exception&) {
    threw = true;
  }
  if (!threw) return 5;
  constraint->setArmModels(left, top);
  constraint->setInitialPosition(start);
  constraint->setTolerance(0.001, 0.005);
  constraint->setMaxIterations(1000);
  threw = false;
  try {
    constraint->setTolerance(0.0, 0.005);
  } catch (const ccp::Exception&) {
    threw = true;
  }
  if (!threw) return 6;

  // single-state calls, reference style
  std::vector<double> x0(seeds.begin(), seeds.begin() + 14);
  double fx[2], J[28];
  constraint->function(x0.data(), fx);
  constraint->jacobian(x0.data(), J);
  bool ok0 = constraint->project(x0.data());
  bool sat0 = constraint->isSatisfied(x0.data());
  bool jv0 = constraint->jointValid(x0.data());

  // batched call
  ccp::ProjectBatchResult r = constraint->projectBatch(seeds.data(), count);

  // streaming form: the two halves of the batch in flight together, same results
  {
    const int64_t half = count / 2, rest = count - half;
    std::vector<double> sx((size_t)count * 14);
    std::vector<uint8_t> sok(count);
    std::vector<int32_t> sit(count);
    int64_t t0 = constraint->submitBatch(seeds.data(), half, sx.data(), sok.data(), nullptr, sit.data());
    int64_t t1 = constraint->submitBatch(seeds.data() + half * 14, rest, sx.data() + half * 14, sok.data() + half, nullptr,
                                         sit.data() + half);
    constraint->waitBatch(t0);
    constraint->waitBatch(t1);
    if (memcmp(sx.data(), r.x.data(), sizeof(double) * sx.size()) != 0) return 8;
    if (memcmp(sok.data(), r.ok.data(), sok.size()) != 0 || memcmp(sit.data(), r.iters.data(), 4 * sit.size()) != 0) return 9;
  }

  ccp::PandaModel pm;
  double q0[7] = {0, 0, 0, 0, 0, 0, 0};
  auto T = pm.getTransform(q0);
  auto Jg = pm.getJacobianMatrix(start);

  FILE* o = fopen(argv[2], "wb");
  if (!o) return 7;
  fwrite(fx, sizeof(double), 2, o);
  fwrite(J, sizeof(double), 28, o);
  fwrite(x0.data(), sizeof(double), 14, o);
  uint8_t flags[3] = {(uint8_t)ok0, (uint8_t)sat0, (uint8_t)jv0};
  fwrite(flags, 1, 3, o);
  fwrite(r.x.data(), sizeof(double), r.x.size(), o);
  fwrite(r.ok.data(), 1, r.ok.size(), o);
  fwrite(r.iters.data(), sizeof(int32_t), r.iters.size(), o);
  fwrite(T.data(), sizeof(double), 12, o);
  fwrite(Jg.data(), sizeof(double), 42, o);
  fclose(o);
  printf("cpp host test: %lld states, first ok=%d\n", (long long)count, (int)ok0);
  return 0;
}
