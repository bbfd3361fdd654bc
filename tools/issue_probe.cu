// issue_probe.cu — does a non-FP64 instruction cost issue time next to FP64 work on B200? (diagnostic)
//   nvcc -O3 --fmad=false -gencode arch=compute_100a,code=sm_100a -o issue_probe issue_probe.cu && ./issue_probe
// W warps per SM sub-partition run 8 independent DFMA chains each; variant I adds I independent integer ops (LOP3/IADD
// on their own registers) per DFMA.  If the integer ops hide in the FP64 pipe's second cycle the time does not change.
#include <cstdio>
#include <cuda_runtime.h>

template <int I>
__global__ void k(double* sink, unsigned* isink, int inner, double a, double b, unsigned m) {
  double r[8];
  unsigned u[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) { r[c] = threadIdx.x * 1e-3 + c; u[c] = threadIdx.x + c; }
#pragma unroll 1
  for (int i = 0; i < inner; ++i) {
#pragma unroll
    for (int rep = 0; rep < 8; ++rep) {
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        r[c] = fma(r[c], a, b);
        if (I >= 1) u[c] = (u[c] ^ m) + 0x9E3779B9u;          // LOP3 + IADD -> counts as ~2 ALU ops
        if (I >= 2) u[(c + 3) & 7] = (u[(c + 3) & 7] << 1) | (u[c] >> 31);
      }
    }
  }
  double s = 0; unsigned t = 0;
#pragma unroll
  for (int c = 0; c < 8; ++c) { s += r[c]; t ^= u[c]; }
  if (s == 123.456) sink[0] = s;
  if (t == 0x12345u) isink[0] = t;
}

template <int I>
float run(int warps, int sms, double* sink, unsigned* isink) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int block = warps * 4 * 32;
  k<I><<<sms, block>>>(sink, isink, 4, 1.0000001, 1e-9, 0x5bd1e995u);
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    k<I><<<sms, block>>>(sink, isink, 512, 1.0000001, 1e-9, 0x5bd1e995u);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  return best;
}

int main() {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double* sink; unsigned* isink; cudaMalloc(&sink, 8); cudaMalloc(&isink, 4);
  const double dfma_per_warp = 512.0 * 8 * 8;
  for (int w : {1, 2, 3, 4}) {
    float t0 = run<0>(w, sms, sink, isink), t1 = run<1>(w, sms, sink, isink), t2 = run<2>(w, sms, sink, isink);
    double clk0 = t0 * 1e-3 * 1.965e9 / (dfma_per_warp * w);
    printf("%d warps/SMSP: DFMA only %.3f ms (%.2f clk per DFMA per SMSP)   +2 int/DFMA %.3f ms (x%.2f)   +4 int/DFMA %.3f ms (x%.2f)\n",
           w, t0, clk0, t1, t1 / t0, t2, t2 / t0);
  }
  return 0;
}
