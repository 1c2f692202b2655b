// VAE wind-field decoder (reset path): generative/vae.py:134-186, env/generative_wind_field.py:52-62.
// 64 latents -> 3 x (Dense 1000 + ReLU) -> Dense 4410 as four cuBLASLt GEMMs (bias / bias+ReLU fused in
// the GEMM epilogue), then one kernel for  reshape (7,7,90) -> linear resize to (23,23,90) -> central
// differences -> crop -> [21,21,10,9,2].
// Included from inside `namespace ble` of ble_engine.cu.
#pragma once

constexpr int kDecLatents = 64, kDecHidden = 1000, kDecFlowW = 7, kDecFlows = 90;
constexpr int kDecOut = kDecFlowW * kDecFlowW * kDecFlows;      // 4410
constexpr int kDecResized = kNX + 2;                            // 23

// jax.image.resize(..., method='linear') when up-sampling 7 -> 23: half-pixel centres, triangle kernel,
// taps outside the input dropped and the remaining weights renormalised (== edge clamp).
struct ResizeTaps { int p0[kDecResized], p1[kDecResized]; float w0[kDecResized], w1[kDecResized]; };

inline ResizeTaps make_resize_taps() {
  ResizeTaps t;
  const double scale = double(kDecResized) / double(kDecFlowW);
  for (int a = 0; a < kDecResized; ++a) {
    const double pos = (a + 0.5) / scale - 0.5;
    double w[kDecFlowW], sum = 0.0;
    for (int i = 0; i < kDecFlowW; ++i) { w[i] = fmax(0.0, 1.0 - fabs(pos - i)); sum += w[i]; }
    int first = -1, second = -1;
    for (int i = 0; i < kDecFlowW; ++i) if (w[i] > 0.0) { if (first < 0) first = i; else second = i; }
    t.p0[a] = first; t.w0[a] = float(w[first] / sum);
    t.p1[a] = second < 0 ? first : second; t.w1[a] = second < 0 ? 0.f : float(w[second] / sum);
  }
  return t;
}

// flow: [F][7][7][90] (Dense_3 output, row-major); out: native [F][21][21][10][9][2].
// One thread per (field, i, j, c): u = dPsi/d(axis 0), v = -dPsi/d(axis 1) at resized (i + 1, j + 1).
__global__ void __launch_bounds__(256)
k_decode_epilogue(const float* __restrict__ flow, float* __restrict__ out, ResizeTaps taps, int64_t n_fields) {
  const int64_t t = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  const int64_t per_field = int64_t(kNX) * kNY * kDecFlows;
  if (t >= n_fields * per_field) return;
  const int64_t f = t / per_field;
  int r = int(t - f * per_field);
  const int c = r % kDecFlows; r /= kDecFlows;
  const int j = r % kNY;
  const int i = r / kNY;
  const float* psi = flow + f * kDecOut + c;
  auto resized = [&](int a, int b) {                       // jax.image.resize: axis 0 first, then axis 1
    const float ta = taps.w0[a] * psi[(taps.p0[a] * kDecFlowW + taps.p0[b]) * kDecFlows] +
                     taps.w1[a] * psi[(taps.p1[a] * kDecFlowW + taps.p0[b]) * kDecFlows];
    const float tb = taps.w0[a] * psi[(taps.p0[a] * kDecFlowW + taps.p1[b]) * kDecFlows] +
                     taps.w1[a] * psi[(taps.p1[a] * kDecFlowW + taps.p1[b]) * kDecFlows];
    return taps.w0[b] * ta + taps.w1[b] * tb;
  };
  const float u = (resized(i + 2, j + 1) - resized(i, j + 1)) * 0.5f;      // vae.py:171-173, crop [1:-1]
  const float v = -(resized(i + 1, j + 2) - resized(i + 1, j)) * 0.5f;     // vae.py:174-176, :183
  // c = pressure * 9 + time (reshape of the 90 flow fields into (10, 9), vae.py:182)
  reinterpret_cast<float2*>(out)[(f * kNX * kNY + int64_t(i) * kNY + j) * kDecFlows + c] = make_float2(u, v);
}
