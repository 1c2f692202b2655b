// Probe: fragment layout and throughput of mma.sync.m8n8k4.f64 on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_probe dmma_probe.cu && ./dmma_probe
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b, double c0, double c1) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};"
               : "=d"(d0), "=d"(d1) : "d"(a), "d"(b), "d"(c0), "d"(c1));
}

__global__ void layout_check(const double* A /*8x4 row-major*/, const double* B /*4x8 row-major [k][n]*/, double* D /*8x8*/) {
  const int l = threadIdx.x;
  const double a = A[(l >> 2) * 4 + (l & 3)];          // row = l / 4, col = l % 4
  const double b = B[(l & 3) * 8 + (l >> 2)];          // row (k) = l % 4, col (n) = l / 4
  double d0, d1;
  dmma(d0, d1, a, b, 0.0, 0.0);
  D[(l >> 2) * 8 + 2 * (l & 3)] = d0;                  // row = l / 4, cols 2 (l % 4), +1
  D[(l >> 2) * 8 + 2 * (l & 3) + 1] = d1;
}

__global__ void dmma_rate(double* out, int iters) {
  double c[8][2];
  for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
  const double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) dmma(c[i][0], c[i][1], a, b, c[i][0], c[i][1]);
  }
  double s = 0;
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void dfma_rate(double* out, int iters) {
  double c[16];
  for (int i = 0; i < 16; ++i) c[i] = threadIdx.x + i;
  const double a = 1.0 + 1e-9 * threadIdx.x, b = 1e-9 * threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = fma(c[i], a, b);
  }
  double s = 0;
  for (int i = 0; i < 16; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  double hA[32], hB[32], hD[64], *dA, *dB, *dD;
  for (int i = 0; i < 32; ++i) { hA[i] = 1 + i * 0.5; hB[i] = 2 - i * 0.25; }
  cudaMalloc(&dA, sizeof hA); cudaMalloc(&dB, sizeof hB); cudaMalloc(&dD, sizeof hD);
  cudaMemcpy(dA, hA, sizeof hA, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, sizeof hB, cudaMemcpyHostToDevice);
  layout_check<<<1, 32>>>(dA, dB, dD);
  cudaMemcpy(hD, dD, sizeof hD, cudaMemcpyDeviceToHost);
  double worst = 0;
  for (int i = 0; i < 8; ++i) for (int n = 0; n < 8; ++n) {
    double r = 0; for (int k = 0; k < 4; ++k) r += hA[i * 4 + k] * hB[k * 8 + n];
    double e = r - hD[i * 8 + n]; if (e < 0) e = -e; if (e > worst) worst = e;
  }
  printf("{\"layout_max_abs_error\": %g", worst);
  int dev = 0, sms = 0, khz = 0; cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  double* out; cudaMalloc(&out, sizeof(double) * sms * 8 * 512);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  for (int warps = 4; warps <= 16; warps *= 2) {
    float ms;
    dmma_rate<<<sms, warps * 32>>>(out, 100);
    cudaEventRecord(e0); dmma_rate<<<sms, warps * 32>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    const double dm = double(iters) * 8 * warps * 256 * sms / (ms * 1e-3);
    dfma_rate<<<sms, warps * 32>>>(out, 100);
    cudaEventRecord(e0); dfma_rate<<<sms, warps * 32>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    const double df = double(iters) * 16 * warps * 32 * sms / (ms * 1e-3);
    printf(", \"warps_%d\": {\"dmma_tflops\": %.2f, \"dfma_tflops\": %.2f}", warps, 2 * dm / 1e12, 2 * df / 1e12);
  }
  printf(", \"sms\": %d, \"clock_mhz\": %d}\n", sms, khz / 1000);
  return 0;
}
