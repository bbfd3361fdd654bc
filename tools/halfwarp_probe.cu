// halfwarp_probe.cu — two questions about divergent paths of one warp on B200 (diagnostic):
//  (1) does a DFMA issued for 16 active lanes occupy the FP64 pipe for half the time of a 32-lane one?
//  (2) does the scheduler interleave the two divergent paths of a warp, so that one path's dependent-issue
//      latency is filled with the other path's instructions?
//   nvcc -O3 --fmad=false -gencode arch=compute_100a,code=sm_100a -o halfwarp_probe halfwarp_probe.cu && ./halfwarp_probe
// MODE 0: all 32 lanes run CH independent DFMA chains.  MODE 1: lanes 16-31 leave at once (half the lane-work).
// MODE 2: lanes 0-15 and lanes 16-31 run two different copies of the loop (divergent until the end; same lane-work as 0).
#include <cstdio>
#include <cuda_runtime.h>

template <int CH>
__device__ __forceinline__ void body(double (&r)[CH], int inner, double a, double b) {
#pragma unroll 1
  for (int i = 0; i < inner; ++i) {
#pragma unroll
    for (int rep = 0; rep < 16; ++rep) {
#pragma unroll
      for (int c = 0; c < CH; ++c) r[c] = fma(r[c], a, b);
    }
  }
}
template <int CH>
__device__ __noinline__ void body2(double (&r)[CH], int inner, double a, double b) {
#pragma unroll 1
  for (int i = 0; i < inner; ++i) {
#pragma unroll
    for (int rep = 0; rep < 16; ++rep) {
#pragma unroll
      for (int c = 0; c < CH; ++c) r[c] = fma(r[c], b, a);  // a different instruction stream at a different address
    }
  }
}

template <int CH, int MODE>
__global__ void k(double* sink, int inner, double a, double b) {
  double r[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) r[c] = threadIdx.x * 1e-3 + c;
  const int lane = threadIdx.x & 31;
  if (MODE == 1 && lane >= 16) return;
  if (MODE == 2 && lane >= 16) body2<CH>(r, inner, a, b);
  else body<CH>(r, inner, a, b);
  double s = 0;
#pragma unroll
  for (int c = 0; c < CH; ++c) s += r[c];
  if (s == 123.456) sink[0] = s;
}

template <int CH, int MODE>
float run(int warps, int sms, double* sink) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int block = warps * 4 * 32;
  k<CH, MODE><<<sms, block>>>(sink, 4, 1.0000001, 1e-9);
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    k<CH, MODE><<<sms, block>>>(sink, 2048 / CH, 1.0000001, 1e-9);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  return best;
}

template <int CH>
void report(int sms, double* sink) {
  for (int w : {1, 3}) {
    float t0 = run<CH, 0>(w, sms, sink), t1 = run<CH, 1>(w, sms, sink), t2 = run<CH, 2>(w, sms, sink);
    const double dfma = (2048.0 / CH) * 16 * CH;  // per lane
    printf("%d chains/lane, %d warps/SMSP: 32 lanes %.3f ms (%.2f clk per warp-DFMA)   lanes 0-15 only %.3f ms (x%.2f)   "
           "two divergent halves %.3f ms (x%.2f)\n",
           CH, w, t0, t0 * 1e-3 * 1.965e9 / (dfma * w), t1, t1 / t0, t2, t2 / t0);
  }
}

int main() {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double* sink; cudaMalloc(&sink, 8);
  report<8>(sms, sink);   // throughput-bound
  report<1>(sms, sink);   // latency-bound: one dependent chain
  report<2>(sms, sink);
  return 0;
}
