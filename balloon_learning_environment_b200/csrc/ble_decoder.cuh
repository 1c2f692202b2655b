// VAE wind-field decoder (reset path): generative/vae.py:134-186, env/generative_wind_field.py:52-62.
// 64 latents -> 3 x (Dense 1000 + ReLU) -> Dense 4410 as four cuBLASLt GEMMs (bias / bias+ReLU fused in
// the GEMM epilogue; TF32 tensor cores by default), then ONE kernel for  reshape (7,7,90) -> linear resize to
// (23,23,90) -> central differences -> crop -> [21,21,10,9,2]: k_flow_to_windows writes the lookup windows of the
// field bank directly (ble_generate_fields), k_decode_epilogue the native layout (ble_decode_fields).
// Included from inside `namespace ble` of ble_engine.cu.
#pragma once

constexpr int kDecLatents = 64, kDecHidden = 1000, kDecFlowW = 7, kDecFlows = 90;
constexpr int kDecOut = kDecFlowW * kDecFlowW * kDecFlows;      // 4410
constexpr int kDecOutPad = 4416;   // leading dimension of Dense_3's weights / output: a multiple of 32 floats, so that cuBLASLt
                                   // may pick its TMA-fed sm_100 kernels (4410 is only 8-byte aligned) and rows are 16 B aligned
constexpr int kDecResized = kNX + 2;                            // 23

// jax.image.resize(..., method='linear') when up-sampling 7 -> 23: half-pixel centres, triangle kernel,
// taps outside the input dropped and the remaining weights renormalised (== edge clamp).
// resize_tap(a) is constexpr so that the fused writer can unroll over `a` with the taps as immediates; the table the
// other kernels index at run time is filled from the same function.
struct ResizeTap { int p0, p1; float w0, w1; };
__host__ __device__ constexpr ResizeTap resize_tap(int a) {
  const double pos = (a + 0.5) * double(kDecFlowW) / double(kDecResized) - 0.5;
  double w[kDecFlowW] = {}, sum = 0.0;
  for (int i = 0; i < kDecFlowW; ++i) {
    const double d = pos > i ? pos - i : i - pos;
    w[i] = d < 1.0 ? 1.0 - d : 0.0;
    sum += w[i];
  }
  int first = -1, second = -1;
  for (int i = 0; i < kDecFlowW; ++i) if (w[i] > 0.0) { if (first < 0) first = i; else second = i; }
  ResizeTap t{first, second < 0 ? first : second, float(w[first] / sum), second < 0 ? 0.f : float(w[second] / sum)};
  return t;
}
struct ResizeTaps { int p0[kDecResized], p1[kDecResized]; float w0[kDecResized], w1[kDecResized]; };

inline ResizeTaps make_resize_taps() {
  ResizeTaps t;
  for (int a = 0; a < kDecResized; ++a) {
    const ResizeTap k = resize_tap(a);
    t.p0[a] = k.p0; t.p1[a] = k.p1; t.w0[a] = k.w0; t.w1[a] = k.w1;
  }
  return t;
}

// One resized stream-function sample: jax.image.resize applies axis 0 first, then axis 1.  psi points at
// [7][7][stride] for one flow field c; explicit fma so that every kernel that calls this rounds identically.
__device__ __forceinline__ float resized_psi(const float* psi, int stride, const ResizeTaps& t, int a, int b) {
  const float ta = fmaf(t.w0[a], psi[(t.p0[a] * kDecFlowW + t.p0[b]) * stride],
                        t.w1[a] * psi[(t.p1[a] * kDecFlowW + t.p0[b]) * stride]);
  const float tb = fmaf(t.w0[a], psi[(t.p0[a] * kDecFlowW + t.p1[b]) * stride],
                        t.w1[a] * psi[(t.p1[a] * kDecFlowW + t.p1[b]) * stride]);
  return fmaf(t.w0[b], ta, t.w1[b] * tb);
}

// flow: [F][4416] rows holding [7][7][90] (Dense_3 output, row-major, padded); out: native [F][21][21][10][9][2].
// One thread per (field, i, j, c): u = dPsi/d(axis 0), v = -dPsi/d(axis 1) at resized (i + 1, j + 1).
__global__ void __launch_bounds__(256)
k_decode_epilogue(const float* __restrict__ flow, float* __restrict__ out, const __grid_constant__ ResizeTaps taps,
                  int64_t n_fields) {
  const int64_t t = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  const int64_t per_field = int64_t(kNX) * kNY * kDecFlows;
  if (t >= n_fields * per_field) return;
  const int64_t f = t / per_field;
  int r = int(t - f * per_field);
  const int c = r % kDecFlows; r /= kDecFlows;
  const int j = r % kNY;
  const int i = r / kNY;
  const float* psi = flow + f * kDecOutPad + c;
  const float u = (resized_psi(psi, kDecFlows, taps, i + 2, j + 1) - resized_psi(psi, kDecFlows, taps, i, j + 1)) * 0.5f;      // vae.py:171-173, crop [1:-1]
  const float v = -(resized_psi(psi, kDecFlows, taps, i + 1, j + 2) - resized_psi(psi, kDecFlows, taps, i + 1, j)) * 0.5f;    // vae.py:174-176, :183
  // c = pressure * 9 + time (reshape of the 90 flow fields into (10, 9), vae.py:182)
  reinterpret_cast<float2*>(out)[(f * kNX * kNY + int64_t(i) * kNY + j) * kDecFlows + c] = make_float2(u, v);
}

// The reset path's field writer: Dense_3 output -> lookup windows in ONE pass, so that the 1.9 MB (X64) / 3.7 MB
// (X128) of windows per field are the only HBM traffic besides the 17.6 KB of flow values (the round-1 path wrote
// the native field, read it back with 15 KB-strided gathers and then wrote the windows: 1.27 TB/s).
// One CTA per (field, y-cell iy).  The windows of that cell row need the native grid rows y = iy, iy + 1, i.e.
// the resized stream function on rows b = iy .. iy + 3:
//   stage 1: R[4][23][90]      resized rows, from the field's 7x7x90 flow values staged in shared memory
//   stage 2: S[2][21][90] (u,v) central differences = the two native rows (aliases the flow staging buffer)
//   stage 3: 72 (p-cell, t-cell) rows x blocks_per_row 64-byte blocks, written as consecutive float4 per thread
//            (a warp store covers 512 contiguous bytes).
// Values are bit-identical to k_decode_epilogue + k_fields_to_windows (same resized_psi, same differences).
constexpr int kFlowRows = 4;
constexpr int kFlowThreads = 384;                               // 360 = 90 flows x 4 resized rows do stage 1
constexpr int kFlowSmemFloats = kFlowRows * kDecResized * kDecFlows + 2 * kNX * kDecFlows * 2;   // R + S = 8,280 + 7,560

template <int A>
__device__ __forceinline__ void flow_resize_row(const float (&col0)[kDecFlowW], const float (&col1)[kDecFlowW], float w0b, float w1b,
                                                float* __restrict__ out /* R + (bb * 23) * 90 + c */) {
  if constexpr (A < kDecResized) {
    constexpr ResizeTap t = resize_tap(A);
    const float ta = fmaf(t.w0, col0[t.p0], t.w1 * col0[t.p1]);          // == resized_psi(.., A, b)
    const float tb = fmaf(t.w0, col1[t.p0], t.w1 * col1[t.p1]);
    out[A * kDecFlows] = fmaf(w0b, ta, w1b * tb);
    flow_resize_row<A + 1>(col0, col1, w0b, w1b, out);
  }
}

template <bool kX128>
__global__ void __launch_bounds__(kFlowThreads)
k_flow_to_windows(const float* __restrict__ flow, float* __restrict__ cells, FieldLayout layout,
                  const __grid_constant__ ResizeTaps taps, int64_t first_field, int64_t n_fields,
                  const int32_t* __restrict__ dst_index /* nullptr: first_field + f */) {
  extern __shared__ float smem[];
  float* R = smem;                                              // [4][23][90]
  float* S = smem + kFlowRows * kDecResized * kDecFlows;        // [2][21][90][2]; first holds psi [7][7][90]
  const int64_t f = blockIdx.x / kYC;
  const int iy = int(blockIdx.x - f * kYC);
  if (f >= n_fields) return;
  const int tid = threadIdx.x;

  {   // stage 0: the field's 4,410 flow values (2,205 float2), all loads in flight before the first store
    const float2* src = reinterpret_cast<const float2*>(flow + f * kDecOutPad);
    constexpr int kPairs = kDecOut / 2, kIter = (kPairs + kFlowThreads - 1) / kFlowThreads;      // 2205, 6
    float2 v[kIter];
#pragma unroll
    for (int k = 0; k < kIter; ++k) {
      const int i = tid + k * kFlowThreads;
      v[k] = i < kPairs ? __ldg(src + i) : make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int k = 0; k < kIter; ++k) {
      const int i = tid + k * kFlowThreads;
      if (i < kPairs) reinterpret_cast<float2*>(S)[i] = v[k];
    }
  }
  __syncthreads();
  // stage 1: thread (c, bb) keeps the two psi columns its resized row blends (7 + 7 values) in registers and walks
  // the 23 resized positions with compile-time taps
  if (tid < kDecFlows * kFlowRows) {
    const int c = tid % kDecFlows, bb = tid / kDecFlows;
    const int b = iy + bb;
    const int p0b = taps.p0[b], p1b = taps.p1[b];
    const float w0b = taps.w0[b], w1b = taps.w1[b];
    float col0[kDecFlowW], col1[kDecFlowW];
#pragma unroll
    for (int p = 0; p < kDecFlowW; ++p) {
      col0[p] = S[(p * kDecFlowW + p0b) * kDecFlows + c];
      col1[p] = S[(p * kDecFlowW + p1b) * kDecFlows + c];
    }
    flow_resize_row<0>(col0, col1, w0b, w1b, R + bb * kDecResized * kDecFlows + c);
  }
  __syncthreads();
  // stage 2 (overwrites the flow staging): thread (c, dy, half of the x range)
  float2* S2 = reinterpret_cast<float2*>(S);
  if (tid < kDecFlows * 4) {
    const int c = tid % kDecFlows, dy = (tid / kDecFlows) & 1, half = tid / (2 * kDecFlows);
    const int x0 = half ? 11 : 0, x1 = half ? kNX : 11;
    const float* r0 = R + ((dy + 0) * kDecResized) * kDecFlows + c;
    const float* r1 = R + ((dy + 1) * kDecResized) * kDecFlows + c;
    const float* r2 = R + ((dy + 2) * kDecResized) * kDecFlows + c;
    float lo = r1[x0 * kDecFlows], mid = r1[(x0 + 1) * kDecFlows];
#pragma unroll 1
    for (int x = x0; x < x1; ++x) {
      const float hi = r1[(x + 2) * kDecFlows];
      const float u = (hi - lo) * 0.5f;
      const float v = -(r2[(x + 1) * kDecFlows] - r0[(x + 1) * kDecFlows]) * 0.5f;
      S2[(dy * kNX + x) * kDecFlows + c] = make_float2(u, v);
      lo = mid; mid = hi;
    }
  }
  __syncthreads();
  // stage 3: the 72 (p-cell, t-cell) window rows of this y-cell, kRows rows per pass.  A thread keeps its 16-byte chunk
  // position within the row (block b, chunk ch -> grid column ix, dy, dp) for the whole loop; only the row advances, so
  // the shared-memory pointer moves by kRows flow indices per pass (+1 more when t wraps into the next pressure cell:
  // c = p * 9 + t) and the destination by kRows whole rows.  A warp store covers 512 contiguous bytes.
  constexpr int kCpr = kX128 ? (kNX - 1) * 8 : kNX * 4;         // 16-byte chunks per row: 160 / 84
  constexpr int kRows = kFlowThreads / kCpr;                    // rows per pass: 2 / 4
  static_assert((kPC * kTC) % kRows == 0 && kTC % kRows == 0, "rows per pass must divide the t-cells");
  if (tid < kRows * kCpr) {
    const int r0 = tid / kCpr, within = tid - r0 * kCpr;
    const int b = within >> 2, ch = within & 3;
    const int ix = kX128 ? (b >> 1) + (b & 1) : b;
    const int64_t dst_field = dst_index != nullptr ? int64_t(dst_index[f]) : first_field + f;
    float4* dst = reinterpret_cast<float4*>(cells + dst_field * layout.field_floats + int64_t(iy) * kPC * kTC * layout.row_floats) + tid;
    const float2* s = S2 + ((ch >> 1) * kNX + ix) * kDecFlows + (ch & 1) * kNT + r0;      // row r0: pc = 0, tc = r0
    int tc = r0;
#pragma unroll 2
    for (int pass = 0; pass < kPC * kTC / kRows; ++pass) {
      const float2 t0 = s[0], t1 = s[1];
      __stcs(dst, make_float4(t0.x, t0.y, t1.x, t1.y));
      dst += kRows * kCpr;
      tc += kRows; s += kRows;
      if (tc >= kTC) { tc -= kTC; s += 1; }
    }
  }
}
