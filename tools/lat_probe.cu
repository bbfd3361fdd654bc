// lat_probe.cu — dependent-issue latencies of the FP64 instructions the projection kernel is made of
// (diagnostic, not part of the product).  One warp, one long dependent chain per test, clock64() around it.
//   nvcc -O3 --fmad=false -gencode arch=compute_100a,code=sm_100a -o lat_probe lat_probe.cu && ./lat_probe
#include <cstdio>
#include <cuda_runtime.h>

#define N 512
struct params { double a, b, c; };

template <int T>
__global__ void probe(const __grid_constant__ params P, double* out, long long* cyc, double seed, int sel) {
  double r = seed + threadIdx.x * 1e-9, s = seed * 0.5, t = seed * 0.25, u = seed * 0.125;
  const double a = out[1], b = out[2];  // register operands unknown at compile time
  long long t0 = clock64();
#pragma unroll
  for (int i = 0; i < N; ++i) {
    if (T == 0) r = fma(r, a, b);                     // DFMA, register operands
    if (T == 1) r = r * a;                            // DMUL
    if (T == 2) r = r + b;                            // DADD
    if (T == 3) r = fma(r, P.a, P.b);                 // DFMA, constant-bank operands
    if (T == 4) r = fma(r, 1.0000001, 1e-9);          // DFMA, immediates
    if (T == 5) { r = r * a; r = r + b; }             // DMUL -> DADD
    if (T == 6) { r = fma(r, a, b); r = (sel & 1) ? r : -r; }   // DFMA -> select/negate (ALU on FP64 halves)
    if (T == 7) { r = fma(r, a, b); s = fma(s, a, b); }         // 2 chains
    if (T == 8) { r = fma(r, a, b); s = fma(s, a, b); t = fma(t, a, b); u = fma(u, a, b); }  // 4 chains
    if (T == 9) { r = fma(r, a, b); r = (r > 1e300) ? s : r; }  // DFMA -> DSETP -> FSEL x2
    if (T == 10) { r = fma(r, a, b); s = r * a; t = r + b; r = fma(s, t, r); }  // small diamond
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[T] = t1 - t0;
  out[8 + T * 32 + threadIdx.x] = r + s + t + u;
}

int main() {
  double* out;
  long long* cyc;
  cudaMalloc(&out, 8 * 4096);
  cudaMalloc(&cyc, 8 * 64);
  double h[4] = {0, 1.0000001, 1e-9, 0};
  cudaMemcpy(out, h, 32, cudaMemcpyHostToDevice);
  params P{1.0000001, 1e-9, 0.5};
  const char* names[] = {"DFMA reg", "DMUL", "DADD", "DFMA c[]", "DFMA imm", "DMUL->DADD (2 ops)", "DFMA->neg-select",
                         "DFMA x2 chains (2 ops)", "DFMA x4 chains (4 ops)", "DFMA->DSETP->FSEL", "diamond (4 ops)"};
  for (int rep = 0; rep < 2; ++rep) {
    probe<0><<<1, 32>>>(P, out, cyc, 1.0, 1); probe<1><<<1, 32>>>(P, out, cyc, 1.0, 1); probe<2><<<1, 32>>>(P, out, cyc, 1.0, 1);
    probe<3><<<1, 32>>>(P, out, cyc, 1.0, 1); probe<4><<<1, 32>>>(P, out, cyc, 1.0, 1); probe<5><<<1, 32>>>(P, out, cyc, 1.0, 1);
    probe<6><<<1, 32>>>(P, out, cyc, 1.0, 1); probe<7><<<1, 32>>>(P, out, cyc, 1.0, 1); probe<8><<<1, 32>>>(P, out, cyc, 1.0, 1);
    probe<9><<<1, 32>>>(P, out, cyc, 1.0, 1); probe<10><<<1, 32>>>(P, out, cyc, 1.0, 1);
    cudaDeviceSynchronize();
  }
  long long hc[16];
  cudaMemcpy(hc, cyc, sizeof(long long) * 11, cudaMemcpyDeviceToHost);
  for (int t = 0; t < 11; ++t) printf("%-28s %8lld clk / %d steps = %6.2f clk per step\n", names[t], hc[t], N, (double)hc[t] / N);
  return 0;
}
