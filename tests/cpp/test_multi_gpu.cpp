// C++ host-side test of the multi-GPU entry points (include/ccp.h "multi-GPU for a C++ host"), the way a single-process
// C++ planner would use them: one KinematicChainConstraint per GPU, a peer group over their handles, pool refills
// sharded over the devices with the gather fused into the projection kernels.  No CUDA header is included.
// usage: test_multi_gpu start.bin out.bin [world]     exit code 77 = fewer than two GPUs (skip)
// (built and checked by tests/test_multi_gpu_cpp.py)
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "closed_chain_motion_planner_b200/ProjectedStateSpace.hpp"

static ccp::ChainConstraintPtr make_constraint(int device, const double* start) {
  auto left = std::make_shared<ccp::ArmModel>();
  left->name = "panda_left";
  left->t_wb = ccp::base_frame(0);
  auto top = std::make_shared<ccp::ArmModel>();
  top->name = "panda_top";
  top->t_wb = ccp::base_frame(2);
  auto c = std::make_shared<ccp::KinematicChainConstraint>(14, device);
  c->setArmModels(left, top);
  c->setInitialPosition(start);
  c->setTolerance(0.001, 0.005);
  return c;
}

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  const int ndev = ccp_device_count();
  int world = argc > 3 ? atoi(argv[3]) : ndev;
  if (world > ndev) world = ndev;
  if (world > 8) world = 8;
  if (world < 2) {
    printf("SKIP: %d GPU(s) visible\n", ndev);
    return 77;
  }
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 3;
  double start[14];
  if (fread(start, sizeof(double), 14, f) != 14) return 4;
  fclose(f);

  std::vector<ccp::ChainConstraintPtr> cs;
  for (int d = 0; d < world; ++d) cs.push_back(make_constraint(d, start));

  // ---- two refills of 200 001 seeds over all devices ----
  const int64_t total = 200001;
  ccp::MultiGpuProjectedSampler sampler(cs, total, /*rng_seed*/ 21);
  std::vector<double> rows;
  std::vector<int64_t> counts1, counts2;
  const int64_t got1 = sampler.sample(&rows, &counts1);
  const int64_t got2 = sampler.sample(&rows, &counts2);

  // every device holds the same gathered pool: compare device world-1's copy of the second refill with device 0's
  std::vector<double> last((size_t)got2 * 14);
  std::vector<int64_t> cl(world);
  int64_t rows_last = 0;
  if (ccp_peer_group_gather_host(sampler.group(), world - 1, last.data(), got2, cl.data(), &rows_last) != CCP_OK) return 5;
  const bool same_pool = rows_last == got2 && std::equal(last.begin(), last.end(), rows.begin() + (size_t)got1 * 14) &&
                         std::equal(cl.begin(), cl.end(), counts2.begin());

  // ---- the same two slices of the stream on ONE device through the ordinary host entry point ----
  std::vector<double> single((size_t)2 * total * 14);
  int64_t n1 = 0, n2 = 0;
  ccp_sampler_args a;
  a.rng_seed = 21;
  a.first_index = 0;
  a.mode = 0;
  a.wrap_bounds = 1;
  a.distance = 0.0;
  a.near_host = nullptr;
  if (ccp_sample_project_batch_host(cs[0]->handle(), &a, total, nullptr, nullptr, nullptr, single.data(), &n1) != CCP_OK) return 6;
  a.first_index = total;
  if (ccp_sample_project_batch_host(cs[0]->handle(), &a, total, nullptr, nullptr, nullptr, single.data() + (size_t)n1 * 14, &n2) !=
      CCP_OK)
    return 6;

  // ---- error paths ----
  ccp_peer_group* bad = nullptr;
  ccp_handle* dup[2] = {cs[0]->handle(), cs[0]->handle()};
  const int rc_dup = ccp_peer_group_create(dup, 2, 16, &bad);  // the same device twice
  ccp_handle* one[1] = {cs[0]->handle()};
  ccp_peer_group* tiny = nullptr;
  int rc_overflow = ccp_peer_group_create(one, 1, 8, &tiny);  // capacity 8 overflows at once
  int64_t c1 = 0;
  if (rc_overflow == CCP_OK) rc_overflow = ccp_peer_group_sample_project(tiny, &a, 5000, &c1);
  ccp_peer_group_destroy(tiny);
  // the handles stay usable on their own after the group's calls
  double x[14];
  std::copy(start, start + 14, x);
  x[0] += 0.05;
  const bool single_ok = cs[world - 1]->project(x);

  FILE* o = fopen(argv[2], "wb");
  if (!o) return 7;
  int64_t hdr[8] = {world, got1, got2, n1, n2, (int64_t)same_pool, (int64_t)rc_dup, (int64_t)rc_overflow};
  fwrite(hdr, sizeof(int64_t), 8, o);
  int64_t flag = single_ok;
  fwrite(&flag, sizeof flag, 1, o);
  fwrite(counts1.data(), sizeof(int64_t), world, o);
  fwrite(counts2.data(), sizeof(int64_t), world, o);
  fwrite(rows.data(), sizeof(double), rows.size(), o);
  fwrite(single.data(), sizeof(double), (size_t)(n1 + n2) * 14, o);
  fclose(o);
  return 0;
}
