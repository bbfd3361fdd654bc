// C++ host-side test of the state-space seam (include/closed_chain_motion_planner_b200/ProjectedStateSpace.hpp), written
// the way the reference uses it: allocStateSampler -> sampleUniform (jy_ProjectedStateSpace.cpp:10-15), sampleUniformNear,
// discreteGeodesic between sampled states (stefanBiPRM.cpp:315), the goal sampler's IK loop.
// usage: test_state_space start.bin out.bin   (built and checked by tests/test_cpp_state_space_gpu.py)
#include <cstdio>
#include <vector>

#include "closed_chain_motion_planner_b200/ProjectedStateSpace.hpp"

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 3;
  double start[14];
  if (fread(start, sizeof(double), 14, f) != 14) return 4;
  fclose(f);

  auto left = std::make_shared<ccp::ArmModel>();
  left->name = "panda_left";
  left->t_wb = ccp::base_frame(0);
  auto top = std::make_shared<ccp::ArmModel>();
  top->name = "panda_top";
  top->t_wb = ccp::base_frame(2);
  auto constraint = std::make_shared<ccp::KinematicChainConstraint>(14);
  constraint->setArmModels(left, top);
  constraint->setInitialPosition(start);
  constraint->setTolerance(0.001, 0.005);

  auto ambient = std::make_shared<ccp::KinematicChainSpace>(14);
  auto space = std::make_shared<ccp::jy_ProjectedStateSpace>(ambient, constraint);
  space->setDelta(0.25);   // ConstrainedPlanningCommon.cpp:118
  space->setLambda(2.0);   // :119
  bool threw = false;
  try {
    space->setDelta(0.0);
  } catch (const ccp::Exception&) {
    threw = true;
  }
  if (!threw) return 5;

  // ---- sampler: 300 states popped one by one from a pool of 2000 seeds per refill ----
  auto sampler = space->allocStateSampler(2000, /*rng_seed*/ 11);
  const int S = 300;
  std::vector<double> samples((size_t)S * 14);
  for (int i = 0; i < S; ++i) sampler->sampleUniform(&samples[(size_t)i * 14]);
  std::vector<uint8_t> sat(S), jv(S);
  for (int i = 0; i < S; ++i) {
    sat[i] = constraint->isSatisfied(&samples[(size_t)i * 14]);
    jv[i] = constraint->jointValid(&samples[(size_t)i * 14]);
  }
  double near_state[14];
  const bool near_ok = sampler->sampleUniformNear(near_state, start, 0.2);
  double gauss_state[14];
  const bool gauss_ok = sampler->sampleGaussian(gauss_state, start, 0.1);

  // ---- geodesics: from the start configuration towards each of the first 64 samples, in one batch ----
  const int E = 64, MS = 48;
  std::vector<double> from((size_t)E * 14), to(samples.begin(), samples.begin() + (size_t)E * 14);
  for (int e = 0; e < E; ++e) std::copy(start, start + 14, &from[(size_t)e * 14]);
  ccp::GeodesicBatchResult g = space->discreteGeodesicBatch(from.data(), to.data(), E, MS);
  std::vector<std::vector<double>> one;
  const bool reached0 = space->discreteGeodesic(from.data(), to.data(), true, &one, MS);

  // ---- goal IK for arm 0: targets = FK of the first 32 samples' left-arm joints ----
  ccp::PandaModel pm;
  const int NT = 32;
  std::vector<double> targets((size_t)NT * 12), qref((size_t)NT * 7), qbest((size_t)NT * 7, 0.0);
  for (int i = 0; i < NT; ++i) {
    auto T = pm.getTransform(&samples[(size_t)i * 14]);
    std::copy(T.begin(), T.end(), &targets[(size_t)i * 12]);
    std::copy(start, start + 7, &qref[(size_t)i * 7]);
  }
  std::vector<uint8_t> ikok(NT);
  std::vector<int32_t> iksucc(NT);
  // ikSampleBatch uses the handle's arm 0 in ITS base frame; pm's arms sit in the identity frame like the targets
  ccp::KinematicChainConstraint ikc(14);
  ikc.setArmModels(std::make_shared<ccp::ArmModel>(), std::make_shared<ccp::ArmModel>());
  ccp::ikSampleBatch(ikc, 0, targets.data(), NT, 15, 5, qref.data(), qbest.data(), ikok.data(), iksucc.data());

  // ---- the same stream through a PREFETCHING sampler (the next pool is projected while this one is consumed) and
  // through one that returns failed projections too, as the reference's sampler does (jy_ProjectedStateSpace.cpp:13) ----
  auto ahead = space->allocStateSampler(2000, /*rng_seed*/ 11);
  ahead->setPrefetch(true);
  const int SA = 700;  // spans more than one pool
  std::vector<double> ahead_samples((size_t)SA * 14);
  for (int i = 0; i < SA; ++i) {
    ahead->sampleUniform(&ahead_samples[(size_t)i * 14]);
    if (i == 100) {
      constraint->isSatisfied(&ahead_samples[0]);  // other calls on the handle while a pool is in flight
      ahead->setPrefetch(false);                   // the pool in flight is still consumed next, nothing after it is prefetched
    }
  }
  int64_t ahead_refills = ahead->refills();
  auto all_states = space->allocStateSampler(500, /*rng_seed*/ 11);
  all_states->setReturnFailed(true);
  const int SF = 120;
  std::vector<double> failed_too((size_t)SF * 14);
  for (int i = 0; i < SF; ++i) all_states->sampleUniform(&failed_too[(size_t)i * 14]);

  // ---- goal sampling for the whole chain (sampleCalibGoal): the object frame is taken as the world frame, so
  // t_o7_a = t_wb_a * FK_a(start_a) and the object pose "identity" asks both arms back to their start poses ----
  double t_o7[2][12], T_obj[2][12] = {{1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0}, {1, 0, 0, 0.01, 0, 1, 0, 0, 0, 0, 1, 0}};
  const int arm_frames[2] = {0, 2};
  for (int a = 0; a < 2; ++a) {
    auto Tb = pm.getTransform(start + 7 * a);  // 3x4 in the arm's base frame
    auto W = ccp::base_frame(arm_frames[a]);   // 3x4 base frame in the world
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 4; ++c) {
        double acc = (c == 3) ? W[4 * r + 3] : 0.0;
        for (int k = 0; k < 3; ++k) acc += W[4 * r + k] * Tb[4 * k + c];
        t_o7[a][4 * r + c] = acc;
      }
  }
  double goal_q[2][14];
  uint8_t goal_ok[2];
  double goal_ref[2][14];
  for (int i = 0; i < 2; ++i) std::copy(start, start + 14, goal_ref[i]);
  ccp::sampleGoalBatch(*constraint, &T_obj[0][0], 2, &t_o7[0][0], &goal_ref[0][0], &goal_q[0][0], goal_ok, 15, 3);
  uint8_t goal_sat[2] = {(uint8_t)constraint->isSatisfied(goal_q[0]), (uint8_t)constraint->isSatisfied(goal_q[1])};

  FILE* o = fopen(argv[2], "wb");
  if (!o) return 7;
  fwrite(goal_q, sizeof(double), 28, o);
  fwrite(goal_ok, 1, 2, o);
  fwrite(goal_sat, 1, 2, o);
  fwrite(ahead_samples.data(), sizeof(double), ahead_samples.size(), o);
  fwrite(&ahead_refills, sizeof ahead_refills, 1, o);
  fwrite(failed_too.data(), sizeof(double), failed_too.size(), o);
  fwrite(samples.data(), sizeof(double), samples.size(), o);
  fwrite(sat.data(), 1, S, o);
  fwrite(jv.data(), 1, S, o);
  int64_t refills = sampler->refills();
  fwrite(&refills, sizeof refills, 1, o);
  uint8_t flags[3] = {(uint8_t)near_ok, (uint8_t)gauss_ok, (uint8_t)reached0};
  fwrite(flags, 1, 3, o);
  fwrite(near_state, sizeof(double), 14, o);
  fwrite(gauss_state, sizeof(double), 14, o);
  fwrite(g.reached.data(), 1, E, o);
  fwrite(g.n_states.data(), sizeof(int32_t), E, o);
  fwrite(g.states.data(), sizeof(double), g.states.size(), o);
  int32_t n_one = (int32_t)one.size();
  fwrite(&n_one, sizeof n_one, 1, o);
  fwrite(targets.data(), sizeof(double), targets.size(), o);
  fwrite(qbest.data(), sizeof(double), qbest.size(), o);
  fwrite(ikok.data(), 1, NT, o);
  fclose(o);
  return 0;
}
