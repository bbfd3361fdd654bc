// fp64_probe.cu — DFMA issue/latency characterisation on the box (not part of the product).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_probe fp64_probe.cu && ./fp64_probe
// For W warps per SM sub-partition and C independent dependent-chains per thread, prints the achieved
// DFMA warp-instructions per cycle per sub-partition (peak 0.5) => exposes the dependent-issue latency.
#include <cstdio>
#include <cuda_runtime.h>

template <int C>
__global__ void chains(double* sink, int inner, double a, double b) {
  double r[C];
#pragma unroll
  for (int c = 0; c < C; ++c) r[c] = threadIdx.x * 1e-3 + c;
#pragma unroll 1
  for (int i = 0; i < inner; ++i) {
#pragma unroll
    for (int u = 0; u < 32; ++u) {
#pragma unroll
      for (int c = 0; c < C; ++c) r[c] = fma(r[c], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int c = 0; c < C; ++c) s += r[c];
  if (s == 123.456) sink[0] = s;
}

template <int C>
double run(int warps_per_smsp, int sms, double* sink, double ghz) {
  int block = warps_per_smsp * 4 * 32;
  int inner = 2048 / C;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  chains<C><<<sms, block>>>(sink, 8, 1.0000001, 1e-9);
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    chains<C><<<sms, block>>>(sink, inner, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  double warp_instr_per_smsp = (double)warps_per_smsp * inner * 32.0 * C;
  double cycles = best * 1e-3 * ghz * 1e9;
  return warp_instr_per_smsp / cycles;
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  double ghz = khz * 1e-6;
  double* sink;
  cudaMalloc(&sink, 8);
  printf("SMs %d, clock %.3f GHz (attribute; assumes the GPU runs at it)\n", sms, ghz);
  printf("warps/SMSP  chains  DFMA warp-instr/cycle/SMSP (peak 0.5)  => cycles between dependent DFMAs ~ W*C/rate\n");
  int ws[] = {1, 2, 3, 4, 6, 8, 16};
  for (int w : ws) {
    double r1 = run<1>(w, sms, sink, ghz), r2 = run<2>(w, sms, sink, ghz), r4 = run<4>(w, sms, sink, ghz),
           r8 = run<8>(w, sms, sink, ghz);
    printf("%2d   C=1 %.3f (lat %.1f)   C=2 %.3f   C=4 %.3f   C=8 %.3f\n", w, r1, w / r1, r2, r4, r8);
  }
  return 0;
}
