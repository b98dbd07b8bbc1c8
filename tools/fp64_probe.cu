// Latency / issue rate of the fp64 pipe on this GPU (the BA solve and accumulate are chains of dependent DFMAs).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/fp64_probe tools/fp64_probe.cu && tools/bin/fp64_probe
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS>
__global__ void dfma_chain(double* out, long long* cyc, double a, double b, int n) {
  double x[CHAINS];
#pragma unroll
  for (int c = 0; c < CHAINS; c++) x[c] = (double)threadIdx.x + c;
  long long t0 = clock64();
  for (int i = 0; i < n; i++) {
#pragma unroll
    for (int c = 0; c < CHAINS; c++) x[c] = fma(x[c], a, b);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int c = 0; c < CHAINS; c++) s += x[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

__global__ void ffma_chain(float* out, long long* cyc, float a, float b, int n) {
  float x = (float)threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < n; i++) x = fmaf(x, a, b);
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = x;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

__global__ void rcp_chain(double* out, long long* cyc, int n) {
  double x = 1.5 + threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < n; i++) {
    double r = (double)(1.0f / (float)x);
    r = r * (2.0 - x * r);
    r = r * (2.0 - x * r);
    x = r + 1.25;
  }
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

__global__ void rcp64h_chain(double* out, long long* cyc, int n) {
  double x = 1.5 + threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < n; i++) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(r, fma(-x, r, 1.0), r);
    r = fma(r, fma(-x, r, 1.0), r);
    x = r + 1.25;
  }
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

__global__ void rcp_error(double* out) {
  // worst relative error of the rcp.approx + 2 Newton reciprocal over a sweep of magnitudes
  double worst = 0.0;
  for (int i = 0; i < 4096; i++) {
    const double x = (1.0 + (threadIdx.x * 4096 + i) * 7.450580596923828e-9) * exp2((double)((int)(threadIdx.x % 61) - 30));
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(r, fma(-x, r, 1.0), r);
    r = fma(r, fma(-x, r, 1.0), r);
    const double e = fabs(r * x - 1.0);
    worst = e > worst ? e : worst;
  }
  out[threadIdx.x] = worst;
}

__global__ void div_chain(double* out, long long* cyc, int n) {
  double x = 1.5 + threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < n; i++) x = 1.0 / x + 1.25;
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

int main() {
  double* out; long long* cyc; float* outf;
  cudaMalloc(&out, 1 << 20); cudaMalloc(&outf, 1 << 20); cudaMalloc(&cyc, 64);
  long long h;
  const int n = 4096;
  int threads[] = {32, 128, 512};
  for (int t : threads) {
    dfma_chain<1><<<1, t>>>(out, cyc, 1.0000001, 1e-9, n); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("DFMA 1 chain/thread,  %3d threads/SM: %.1f cycles per dependent DFMA\n", t, (double)h / n);
    dfma_chain<4><<<1, t>>>(out, cyc, 1.0000001, 1e-9, n); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("DFMA 4 chains/thread, %3d threads/SM: %.1f cycles per round of 4 (=> %.2f DFMA/clk/SM)\n", t, (double)h / n, 4.0 * t / ((double)h / n));
    dfma_chain<8><<<1, t>>>(out, cyc, 1.0000001, 1e-9, n); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("DFMA 8 chains/thread, %3d threads/SM: %.1f cycles per round of 8 (=> %.2f DFMA/clk/SM)\n", t, (double)h / n, 8.0 * t / ((double)h / n));
  }
  ffma_chain<<<1, 32>>>(outf, cyc, 1.0000001f, 1e-9f, n); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("FFMA dependent: %.1f cycles\n", (double)h / n);
  rcp_chain<<<1, 32>>>(out, cyc, n); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("fp64 reciprocal (f32 seed + 2 Newton) + DADD, dependent: %.1f cycles\n", (double)h / n);
  rcp64h_chain<<<1, 32>>>(out, cyc, n); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("fp64 reciprocal (rcp.approx.ftz.f64 + 2 Newton) + DADD, dependent: %.1f cycles\n", (double)h / n);
  {
    rcp_error<<<1, 256>>>(out);
    double he[256], w = 0;
    cudaMemcpy(he, out, sizeof(he), cudaMemcpyDeviceToHost);
    for (double v : he) w = v > w ? v : w;
    printf("  its worst |r*x - 1| over 1M samples, 2^-30..2^30: %.3g\n", w);
  }
  div_chain<<<1, 32>>>(out, cyc, n); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("fp64 IEEE division + DADD, dependent: %.1f cycles\n", (double)h / n);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
