// Probe: fragment layout and throughput of mma.sync.m16n8k8.tf32 (fp32 accumulate) on sm_100a, next to FFMA.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tf32_probe tf32_probe.cu && ./tf32_probe
#include <cstdio>
#include <cstdint>
#include <cmath>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t to_tf32(float x) { uint32_t r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x)); return r; }

__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// A 16x8 row-major, B 8x8 [k][n] row-major, D 16x8
__global__ void layout_check(const float* A, const float* B, float* D) {
  const int l = threadIdx.x, g = l >> 2, t = l & 3;
  uint32_t a[4] = {to_tf32(A[g * 8 + t]), to_tf32(A[(g + 8) * 8 + t]), to_tf32(A[g * 8 + t + 4]), to_tf32(A[(g + 8) * 8 + t + 4])};
  uint32_t b[2] = {to_tf32(B[t * 8 + g]), to_tf32(B[(t + 4) * 8 + g])};
  float c[4] = {0, 0, 0, 0};
  mma_tf32(c, a, b);
  D[g * 8 + 2 * t] = c[0]; D[g * 8 + 2 * t + 1] = c[1]; D[(g + 8) * 8 + 2 * t] = c[2]; D[(g + 8) * 8 + 2 * t + 1] = c[3];
}

__global__ void mma_rate(float* out, int iters) {
  float c[8][4];
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) c[i][j] = threadIdx.x + i + j;
  uint32_t a[4] = {to_tf32(1.0f + 1e-3f * threadIdx.x), to_tf32(0.5f), to_tf32(0.25f), to_tf32(1.5f)};
  uint32_t b[2] = {to_tf32(1.0f - 1e-3f * threadIdx.x), to_tf32(0.75f)};
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) mma_tf32(c[i], a, b);
  }
  float s = 0;
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void ffma_rate(float* out, int iters) {
  float c[16];
  for (int i = 0; i < 16; ++i) c[i] = threadIdx.x + i;
  const float a = 1.0f + 1e-6f * threadIdx.x, b = 1e-6f * threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = fmaf(c[i], a, b);
  }
  float s = 0;
  for (int i = 0; i < 16; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  float hA[128], hB[64], hD[128], *dA, *dB, *dD;
  for (int i = 0; i < 128; ++i) hA[i] = 1 + (i % 17) * 0.5f;
  for (int i = 0; i < 64; ++i) hB[i] = 2 - (i % 13) * 0.25f;
  cudaMalloc(&dA, sizeof hA); cudaMalloc(&dB, sizeof hB); cudaMalloc(&dD, sizeof hD);
  cudaMemcpy(dA, hA, sizeof hA, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, sizeof hB, cudaMemcpyHostToDevice);
  layout_check<<<1, 32>>>(dA, dB, dD);
  cudaMemcpy(hD, dD, sizeof hD, cudaMemcpyDeviceToHost);
  double worst = 0;
  for (int i = 0; i < 16; ++i) for (int n = 0; n < 8; ++n) {
    double r = 0; for (int k = 0; k < 8; ++k) r += double(hA[i * 8 + k]) * hB[k * 8 + n];
    worst = fmax(worst, fabs(r - hD[i * 8 + n]));
  }
  printf("{\"layout_max_abs_error\": %g", worst);
  int dev = 0, sms = 0, khz = 0; cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  float* out; cudaMalloc(&out, sizeof(float) * sms * 8 * 1024);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  for (int warps = 4; warps <= 32; warps *= 2) {
    float ms;
    mma_rate<<<sms, warps * 32>>>(out, 100);
    cudaEventRecord(e0); mma_rate<<<sms, warps * 32>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    const double mm = double(iters) * 8 * warps * 1024 * sms / (ms * 1e-3);
    ffma_rate<<<sms, warps * 32>>>(out, 100);
    cudaEventRecord(e0); ffma_rate<<<sms, warps * 32>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    const double ff = double(iters) * 16 * warps * 32 * sms / (ms * 1e-3);
    printf(", \"warps_%d\": {\"mma_tf32_tflops\": %.1f, \"ffma_tflops\": %.1f}", warps, 2 * mm / 1e12, 2 * ff / 1e12);
  }
  printf(", \"sms\": %d, \"clock_mhz\": %d}\n", sms, khz / 1000);
  return 0;
}
