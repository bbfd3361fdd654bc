// The NCCL fallback of the multi-GPU C ABI (ccp_allgather_converged) from ONE process driving W devices with
// ncclCommInitAll communicators — the host owns NCCL, libccp.so finds its entry points at run time.
// usage: test_multi_nccl start.bin out.bin [world]     exit code 77 = fewer than two GPUs (skip)
#include <cuda_runtime.h>
#include <nccl.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ccp.h"

#define CK(x)                                                  \
  do {                                                         \
    if ((x) != 0) {                                            \
      fprintf(stderr, "failed: %s (line %d)\n", #x, __LINE__); \
      return 10;                                               \
    }                                                          \
  } while (0)

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  int ndev = ccp_device_count();
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
  const int64_t per = 50000, cap = 20064;
  std::vector<ccp_handle*> h(world);
  std::vector<double*> compact(world), pool(world);
  std::vector<int64_t*> n_ok(world), counts(world);
  std::vector<cudaStream_t> st(world);
  std::vector<ncclComm_t> comm(world);
  std::vector<int> devs(world);
  for (int d = 0; d < world; ++d) devs[d] = d;
  CK(ncclCommInitAll(comm.data(), world, devs.data()));
  int32_t idx[2] = {0, 2};
  ccp_model_desc md;
  CK(ccp_default_model(2, idx, &md));
  for (int d = 0; d < world; ++d) {
    CK(cudaSetDevice(d));
    CK(ccp_create(&md, d, &h[d]));
    CK(ccp_set_reference(h[d], start));
    CK(cudaMalloc(&compact[d], sizeof(double) * 14 * per));
    CK(cudaMalloc(&pool[d], sizeof(double) * 14 * cap * world));
    CK(cudaMalloc(&n_ok[d], 8));
    CK(cudaMalloc(&counts[d], 8 * world));
    CK(cudaMemset(n_ok[d], 0, 8));
    CK(cudaStreamCreate(&st[d]));
    ccp_sampler_args a = {33, d * per, 0, 0, 0.0, nullptr};
    CK(ccp_sample_project_batch(h[d], &a, per, CCP_LAYOUT_AOS, nullptr, nullptr, nullptr, compact[d], n_ok[d], st[d]));
  }
  CK(ncclGroupStart());
  for (int d = 0; d < world; ++d) {
    cudaSetDevice(d);
    CK(ccp_allgather_converged(h[d], comm[d], world, compact[d], n_ok[d], cap, pool[d], counts[d], st[d]));
  }
  CK(ncclGroupEnd());
  std::vector<int64_t> c0(world), cl(world), own(world);
  std::vector<double> p0((size_t)world * cap * 14), pl((size_t)world * cap * 14);
  for (int d = 0; d < world; ++d) {
    cudaSetDevice(d);
    CK(cudaStreamSynchronize(st[d]));
    CK(cudaMemcpy(&own[d], n_ok[d], 8, cudaMemcpyDeviceToHost));
  }
  cudaSetDevice(0);
  CK(cudaMemcpy(c0.data(), counts[0], 8 * world, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(p0.data(), pool[0], sizeof(double) * p0.size(), cudaMemcpyDeviceToHost));
  cudaSetDevice(world - 1);
  CK(cudaMemcpy(cl.data(), counts[world - 1], 8 * world, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(pl.data(), pool[world - 1], sizeof(double) * pl.size(), cudaMemcpyDeviceToHost));
  int64_t same = 1;
  for (int d = 0; d < world; ++d) {
    same = same && c0[d] == cl[d] && c0[d] == own[d] && c0[d] <= cap;
    for (int64_t i = 0; i < c0[d] * 14 && same; ++i) same = p0[(size_t)d * cap * 14 + i] == pl[(size_t)d * cap * 14 + i];
  }
  FILE* o = fopen(argv[2], "wb");
  if (!o) return 7;
  int64_t hdr[3] = {world, same, cap};
  fwrite(hdr, 8, 3, o);
  fwrite(c0.data(), 8, world, o);
  for (int d = 0; d < world; ++d) fwrite(&p0[(size_t)d * cap * 14], sizeof(double), (size_t)c0[d] * 14, o);
  fclose(o);
  for (int d = 0; d < world; ++d) {
    ncclCommDestroy(comm[d]);
    ccp_destroy(h[d]);
  }
  return 0;
}
