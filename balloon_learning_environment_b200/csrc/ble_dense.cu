// Dense layers of the QR-DQN network (agents/networks.py:63-98: 8 x Dense(600) + ReLU) on the 5th-generation tensor
// cores: ONE hand-written GEMM kernel serves the forward pass, the input gradient and the weight (+ bias) gradient of
// every layer (learner.py DenseStack is the caller).
//
//   * D[M, N] = A[M, K] . B[N, K]^T with both operands K-contiguous in HBM (modes 0-3: forward, input gradient; the
//     caller keeps a transposed copy of each weight matrix for the latter), or D = A^T . B for row-major A [K, M],
//     B [K, N] (mode 4, MN-major tensor-core operands: the weight gradient straight from activations and output
//     gradients, no transposed copies);
//   * operands: fp32 in HBM, read as TF32 by the tensor core (`tcgen05.mma.cta_group::1.kind::tf32`, fp32 accumulate
//     in tensor memory) -- the precision jax's default matmul has on the reference's GPU path;
//   * tile 128 x 160 x 32 (N = 160 covers the 600-wide layers in 4 tiles, so the 8,192-sample products are 256 CTAs =
//     ONE wave at 2 CTAs per SM, and the 153 logits in one), operands brought by TMA tensor copies
//     (`cp.async.bulk.tensor.2d`; K-major: 128-byte swizzle, one box per operand; MN-major: 128-byte swizzle with
//     32-byte atom, boxes of 32 x 32; out-of-range rows / columns zero-filled) into a 3-stage mbarrier ring, 4 MMAs
//     (K = 8) per stage issued by one elected thread, accumulator = 128 lanes x 160 TMEM columns;
//   * warp roles: warp 0 = TMA producer, warp 1 = TMEM allocation + MMA issue, warps 2-5 = epilogue (each reads its
//     32 TMEM lanes with `tcgen05.ld.32x32b.x32`: a thread holds 32 consecutive columns of one output row);
//   * fused epilogues: + bias; + bias and ReLU, which can also leave the ReLU mask packed 32 columns per word; x mask
//     for the input gradient (from that packed mask, or from the forward activation itself); split-K atomic
//     accumulation into the TRANSPOSED result (a register column is 32 consecutive rows across the lanes, so a warp's
//     32 reductions fall on 32 consecutive floats), the last row of A^T -- the caller's column of ones -- going to the
//     bias gradient.  The row-major tile leaves through swizzled shared-memory staging and TMA stores; any mode can
//     also write the transposed tile directly.
//
// Two CTAs per SM (109 KB of shared memory and 256 TMEM columns each) so that one tile's epilogue overlaps the other's
// main loop.  The forward product of the 8,192 x 600 x 600 layers draws 94 % of what the L2 delivers chip-wide for its
// operand tiles; rows of every buffer should start on 128-byte lines (27 -> 23 us).  mbarrier waits
// are bounded (trap after ~2 s) so that a protocol error surfaces as a CUDA error instead of a hung device.
#include <cuda.h>           // CUtensorMap types only: the encoder is fetched with cudaGetDriverEntryPoint (no -lcuda)
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>

#include "../../include/ble_b200.h"

namespace ble {
namespace {

constexpr int kBM = 128, kBN = 160, kBK = 32;           // kBK floats = 128 B = the swizzle span
constexpr int kUmmaK = 8;                               // tf32: 32 B of K per MMA
constexpr int kBBytes = kBN * kBK * 4;
constexpr int kThreads = 192;
// kAcc = 1 (default): one 128 x 160 accumulator per CTA, 3 stages, 2 CTAs per SM.  kAcc = 2 (BLE_DENSE_ROWS=256, A/B): TWO
// accumulators (a 256 x 160 tile whose halves share every B tile in shared memory: a third less operand traffic per FLOP),
// 4 stages, the whole TMEM, 1 CTA per SM -- measured slower (see ble_dense_tf32).
template <int kAcc> struct Tile {
  static constexpr int kRows = kBM * kAcc;
  static constexpr int kABytes = kRows * kBK * 4;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = kAcc == 1 ? 3 : 4;
  static constexpr int kTmemCols = kAcc == 1 ? 256 : 512;          // power of two >= kAcc * kBN
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*alignment slack*/ + 256 /*barriers*/;
};

struct DenseArgs {
  int64_t m, n, k;
  int k_blocks_per_split;
  const float* aux; int64_t ld_aux;                     // bias [N] (modes 0, 1) or forward activation [M, N] (mode 2)
  float* d; int64_t ldd;                                // row-major output (may be null when only dt is wanted)
  float* dt; int64_t ldt;                               // transposed output [N, M] or null
  int tma_store;                                        // 1: the row-major tile leaves through shared memory + TMA stores
  uint32_t* relu_bits; int64_t ld_bits;                 // packed ReLU mask [M, ld_bits words]: written by mode 1, read by mode 2
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  const long long t0 = clock64();
  for (;;) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) return;
    if (clock64() - t0 > 4000000000ll) __trap();        // ~2 s: a pipeline protocol error, not a slow copy
  }
}

__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint32_t bar, uint32_t dst, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}

// Shared-memory matrix descriptor of a K-major tile stored as rows of 128 B with the 128-byte swizzle: start address,
// stride between 8-row groups = 1024 B, descriptor version 1 (sm_100), layout type 2 = SWIZZLE_128B.
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
  return uint64_t((smem_addr & 0x3FFFFu) >> 4) | (uint64_t(1) << 16) | (uint64_t(1024 >> 4) << 32) | (uint64_t(1) << 46) |
         (uint64_t(2) << 61);
}

// The same for an MN-major tile (mode 4): the operand arrives as rows of K holding 32 consecutive M (or N) elements --
// TMA boxes {32 elements of MN, 32 rows of K} of 4096 B.  For 32-bit MN-major operands the tensor core accepts ONE layout,
// the 128-byte swizzle with a 32-byte atom (descriptor layout type 1, TMA CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B; with the plain
// 128-byte swizzle the MMA returned zeros): the swizzle pattern spans 4 rows of K, so consecutive K are 128 B apart,
// groups of 4 K are 512 B apart (stride byte offset; one K = 8 MMA reads two of them), and the next 32 elements of MN
// are one box = 4096 B further (leading byte offset).
__device__ __forceinline__ uint64_t umma_desc_mn(uint32_t smem_addr) {
  return uint64_t((smem_addr & 0x3FFFFu) >> 4) | (uint64_t(4096 >> 4) << 16) | (uint64_t(512 >> 4) << 32) | (uint64_t(1) << 46) |
         (uint64_t(1) << 61);
}

// Instruction descriptor: D = fp32, A = B = TF32, both K-major, N = kBN, M = kBM.
constexpr uint32_t kInstrDesc = (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(kBN >> 3) << 17) | (uint32_t(kBM >> 4) << 24);

constexpr uint32_t kInstrDescMN = kInstrDesc | (1u << 15) | (1u << 16);       // A and B MN-major

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t accumulate,
                                          uint32_t idesc = kInstrDesc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}

// Row-major output: 32 x 32 boxes staged in shared memory with the 128-byte swizzle, written by TMA (clipped at the
// tensor's edges), so that the HBM writes are whole lines instead of 32 half-filled sectors per store instruction.
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
               ::"l"(map), "r"(c0), "r"(c1), "r"(src) : "memory");
}

__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_load_32(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// kMode: 0 = + bias, 1 = + bias then ReLU, 2 = x ReLU mask, 3 = atomic accumulation (split-K), 4 = 3 with MN-major operands
template <int kMode, int kAcc>
__global__ void __launch_bounds__(kThreads, kAcc == 1 ? 2 : 1)
k_dense_tf32(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
             const __grid_constant__ CUtensorMap map_d, DenseArgs args) {
  constexpr bool kAccumulate = kMode >= 3;              // split-K epilogue
  constexpr bool kMnMajor = kMode == 4;                 // operands row-major [K, M] / [K, N] instead of [M, K] / [N, K]
  constexpr int kStages = Tile<kAcc>::kStages, kStageBytes = Tile<kAcc>::kStageBytes, kABytes = Tile<kAcc>::kABytes;
  constexpr int kTmemCols = Tile<kAcc>::kTmemCols, kRows = Tile<kAcc>::kRows;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;            // swizzled tiles need 1024-byte alignment
  const uint32_t bars = base + kStages * kStageBytes;                      // full[kStages], empty[kStages], tmem_full
  const uint32_t tmem_slot = bars + (2 * kStages + 1) * 8;
  uint8_t* generic_base = smem_raw + (base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(generic_base + kStages * kStageBytes + (2 * kStages + 1) * 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * kBN, m0 = blockIdx.y * kRows;
  const int total_kb = int((args.k + kBK - 1) / kBK);
  const int kb0 = blockIdx.z * args.k_blocks_per_split;
  const int kb1 = min(total_kb, kb0 + args.k_blocks_per_split);
  const int num_kb = kb1 - kb0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(bars + 8 * s, 1); mbar_init(bars + 8 * (kStages + s), 1); }
    mbar_init(bars + 8 * 2 * kStages, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    if (lane == 0) {                                                       // ---- TMA producer ----
      for (int i = 0; i < num_kb; ++i) {
        const int s = i % kStages;
        const uint32_t parity = uint32_t(i / kStages) & 1u;
        mbar_wait(bars + 8 * (kStages + s), parity ^ 1u);                  // slot free (passes at once the first round)
        const uint32_t full = bars + 8 * s;
        mbar_expect_tx(full, kStageBytes);
        const uint32_t sa = base + s * kStageBytes;
        if (kMnMajor) {
#pragma unroll
          for (int c = 0; c < kRows / 32; ++c) tma_load_2d(&map_a, full, sa + c * 4096, m0 + 32 * c, (kb0 + i) * kBK);
#pragma unroll
          for (int c = 0; c < kBN / 32; ++c) tma_load_2d(&map_b, full, sa + kABytes + c * 4096, n0 + 32 * c, (kb0 + i) * kBK);
        } else {
          tma_load_2d(&map_a, full, sa, (kb0 + i) * kBK, m0);
          tma_load_2d(&map_b, full, sa + kABytes, (kb0 + i) * kBK, n0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {                                                       // ---- MMA issue ----
      for (int i = 0; i < num_kb; ++i) {
        const int s = i % kStages;
        const uint32_t parity = uint32_t(i / kStages) & 1u;
        mbar_wait(bars + 8 * s, parity);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t sa = base + s * kStageBytes;
#pragma unroll
        for (int kk = 0; kk < kBK / kUmmaK; ++kk) {
#pragma unroll
          for (int acc = 0; acc < kAcc; ++acc) {            // the accumulators of a 256-row tile share the B tile
            const uint32_t td = tmem_base + uint32_t(acc * kBN), sa_acc = sa + acc * (kBM * kBK * 4);
            if (kMnMajor) {
              umma_tf32(td, umma_desc_mn(sa_acc + kk * 1024), umma_desc_mn(sa + kABytes + kk * 1024),
                        (i > 0 || kk > 0) ? 1u : 0u, kInstrDescMN);
            } else {
              umma_tf32(td, umma_desc(sa_acc + kk * kUmmaK * 4), umma_desc(sa + kABytes + kk * kUmmaK * 4),
                        (i > 0 || kk > 0) ? 1u : 0u);
            }
          }
        }
        umma_commit(bars + 8 * (kStages + s));                             // slot reusable once these MMAs have read it
      }
      umma_commit(bars + 8 * 2 * kStages);                                 // accumulator complete
    }
  } else {                                                                 // ---- epilogue: warps 2..5 ----
    const int q = warp & 3;                                                // the TMEM lane quarter this warp may read
    if (num_kb > 0) {
      mbar_wait(bars + 8 * 2 * kStages, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
#pragma unroll 1
    for (int ac = 0; ac < kAcc * (kBN / 32); ++ac) {
      const int acc = ac / (kBN / 32), c = ac - acc * (kBN / 32);
      const int64_t m = int64_t(m0) + acc * kBM + q * 32 + lane;
      float v[32];
      if (num_kb > 0) {
        tmem_load_32(tmem_base + (uint32_t(q * 32) << 16) + uint32_t(acc * kBN + c * 32), v);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0.f;
      }
      const int64_t nc = int64_t(n0) + c * 32;
      if (nc >= args.n) continue;
      const bool row_ok = m < args.m;
      const bool full_chunk = nc + 32 <= args.n;
      if (kMode == 0 || kMode == 1) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float b = (nc + j < args.n) ? __ldg(args.aux + nc + j) : 0.f;
          v[j] += b;
          if (kMode == 1) v[j] = fmaxf(v[j], 0.f);
        }
        if (kMode == 1 && args.relu_bits != nullptr && row_ok) {
          // the ReLU mask of this row's 32 columns as ONE word: what the input gradient of the backward pass reads back
          // instead of 32 activations (tiles start at multiples of 160 columns, so the chunk is word-aligned)
          uint32_t word = 0;
#pragma unroll
          for (int j = 0; j < 32; ++j) word |= (v[j] > 0.f ? 1u : 0u) << j;
          args.relu_bits[m * args.ld_bits + (nc >> 5)] = word;
        }
      } else if (kMode == 2 && args.relu_bits != nullptr) {
        if (row_ok) {
          const uint32_t word = __ldg(args.relu_bits + m * args.ld_bits + (nc >> 5));
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = ((word >> j) & 1u) ? v[j] : 0.f;
        }
      } else if (kMode == 2) {
        if (row_ok) {
          const float* h = args.aux + m * args.ld_aux + nc;
          if (full_chunk && (args.ld_aux & 3) == 0) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 hv = __ldg(reinterpret_cast<const float4*>(h + j));
              v[j] = hv.x > 0.f ? v[j] : 0.f; v[j + 1] = hv.y > 0.f ? v[j + 1] : 0.f;
              v[j + 2] = hv.z > 0.f ? v[j + 2] : 0.f; v[j + 3] = hv.w > 0.f ? v[j + 3] : 0.f;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = (nc + j < args.n && __ldg(h + j) > 0.f) ? v[j] : 0.f;
          }
        }
      }
      if (kAccumulate) {                                // split-K: accumulate into the TRANSPOSED result, where the 32 lanes
        if (row_ok) {                                   // of a reduction are 32 consecutive floats
          // aux != null: the LAST row of A is the caller's row of ones, so its products are the column sums of B^T --
          // the bias gradient -- and go to aux[n] instead of the (m x n) weight-gradient block
          const bool bias_row = args.aux != nullptr && m == args.m - 1;
          float* p = bias_row ? const_cast<float*>(args.aux) + nc : args.dt + nc * args.ldt + m;
          const int64_t step = bias_row ? 1 : args.ldt;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (nc + j < args.n) atomicAdd(p, v[j]);
            p += step;
          }
        }
        continue;
      }
      if (args.tma_store) {
        // the operand ring is idle once the accumulator is complete: warp q stages its 32 x 32 box of chunk c there
        const uint32_t row = base + uint32_t(acc) * (kBM * kBN * 4) + uint32_t(q) * (kBN / 32) * 4096u + uint32_t(c) * 4096u +
                             uint32_t(lane) * 128u;
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(row + (uint32_t(g ^ (lane & 7)) << 4)),
                       "f"(v[4 * g]), "f"(v[4 * g + 1]), "f"(v[4 * g + 2]), "f"(v[4 * g + 3]) : "memory");
        }
      } else if (args.d != nullptr && row_ok) {
        float* out = args.d + m * args.ldd + nc;
        if (full_chunk && (args.ldd & 3) == 0) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(out + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (nc + j < args.n) out[j] = v[j];
        }
      }
      if (args.dt != nullptr && row_ok) {
        float* pt = args.dt + nc * args.ldt + m;                         // lanes = consecutive m: coalesced
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          if (nc + j < args.n) *pt = v[j];
          pt += args.ldt;
        }
      }
    }
  }
  if (!kAccumulate && args.tma_store && warp >= 2) {
    const int q = warp & 3;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");          // generic-proxy writes -> visible to the TMA
    __syncwarp();
    if (lane == 0) {
      for (int acc = 0; acc < kAcc; ++acc) {
        if (int64_t(m0) + acc * kBM + q * 32 >= args.m) break;
        for (int c = 0; c < kBN / 32; ++c) {
          if (int64_t(n0) + c * 32 >= args.n) break;
          tma_store_2d(&map_d, base + uint32_t(acc) * (kBM * kBN * 4) + uint32_t(q) * (kBN / 32) * 4096u + uint32_t(c) * 4096u,
                       n0 + c * 32, m0 + acc * kBM + q * 32);
        }
      }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");      // shared memory must outlive the reads
    }
  }
  __syncwarp();                                                            // the single-lane roles reconverge their warps
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols) : "memory");
  }
}

// ---- small companions ---------------------------------------------------------------------------------------
// dst[c, r] = src[r, c] (32 x 32 tiles through shared memory, both sides coalesced)
__global__ void __launch_bounds__(256)
k_transpose(const float* __restrict__ src, int64_t lds, int64_t rows, int64_t cols, float* __restrict__ dst, int64_t ldd) {
  __shared__ float tile[32][33];
  const int64_t r0 = int64_t(blockIdx.y) * 32, c0 = int64_t(blockIdx.x) * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8)
    tile[i][tx] = (r0 + i < rows && c0 + tx < cols) ? src[(r0 + i) * lds + c0 + tx] : 0.f;
  __syncthreads();
  for (int i = ty; i < 32; i += 8)
    if (c0 + i < cols && r0 + tx < rows) dst[(c0 + i) * ldd + r0 + tx] = tile[tx][i];
}

// out[r] (+)= sum_c src[r, c]: one CTA per row, 16-byte loads (bias gradient = row sums of the transposed output
// gradient; rows start 16-byte aligned because the pitch is a multiple of 4 floats)
__global__ void __launch_bounds__(256)
k_row_sum(const float* __restrict__ src, int64_t lds, int64_t rows, int64_t cols, float* __restrict__ out, int accumulate) {
  __shared__ float part[8];
  const int64_t r = blockIdx.x;
  const float* p = src + r * lds;
  float s = 0.f;
  if ((lds & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
    const int64_t vec = cols >> 2;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int64_t c = threadIdx.x; c < vec; c += 256) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(p) + c);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    s = (acc.x + acc.y) + (acc.z + acc.w);
    for (int64_t c = (vec << 2) + threadIdx.x; c < cols; c += 256) s += p[c];
  } else {
    for (int64_t c = threadIdx.x; c < cols; c += 256) s += p[c];
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += part[i];
    out[r] = accumulate ? out[r] + t : t;
  }
}

typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiled tensor_map_encoder() {
  static EncodeTiled cached = []() -> EncodeTiled {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult found;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &found) != cudaSuccess ||
        found != cudaDriverEntryPointSuccess) return nullptr;
    return reinterpret_cast<EncodeTiled>(fn);
  }();
  return cached;
}

// [rows, k] fp32, row pitch ld floats -> box {32 floats, box_rows rows}, 128-byte swizzle, zero fill out of range
bool operand_map(CUtensorMap* map, const float* p, int64_t rows, int64_t k, int64_t ld, int box_rows,
                 CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
  if (p == nullptr) { *map = CUtensorMap{}; return true; }
  EncodeTiled enc = tensor_map_encoder();
  if (enc == nullptr) return false;
  const cuuint64_t dims[2] = {cuuint64_t(k), cuuint64_t(rows)};
  const cuuint64_t strides[1] = {cuuint64_t(ld) * 4};
  const cuuint32_t box[2] = {cuuint32_t(kBK), cuuint32_t(box_rows)};
  const cuuint32_t elem[2] = {1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(p), dims, strides, box, elem,
             CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int kMode, int kAcc>
int launch_dense(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& md, const DenseArgs& args, dim3 grid,
                 cudaStream_t s) {
  constexpr int kSmemBytes = Tile<kAcc>::kSmemBytes;
  static bool configured[16] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return BLE_ERR_CUDA;
  if (!configured[dev]) {
    if (cudaFuncSetAttribute(k_dense_tf32<kMode, kAcc>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes) != cudaSuccess)
      return BLE_ERR_CUDA;
    configured[dev] = true;
  }
  k_dense_tf32<kMode, kAcc><<<grid, kThreads, kSmemBytes, s>>>(ma, mb, md, args);
  return cudaGetLastError() == cudaSuccess ? BLE_OK : BLE_ERR_CUDA;
}

}  // namespace
}  // namespace ble

extern "C" {

int ble_dense_tf32(const float* a, int64_t lda, const float* b, int64_t ldb, int64_t m, int64_t n, int64_t k, int32_t mode,
                   const float* aux, int64_t ld_aux, float* d, int64_t ldd, float* dt, int64_t ldt, int32_t split_k,
                   uint32_t* relu_bits, int64_t ld_bits, void* stream) {
  using namespace ble;
  const bool mn = mode == 4;                            // operands [k, m] / [k, n] row-major
  const bool acc = mode >= 3;
  if (a == nullptr || b == nullptr || m <= 0 || n <= 0 || k <= 0 || mode < 0 || mode > 4 || lda < (mn ? m : k) || ldb < (mn ? n : k) ||
      (lda & 3) != 0 || (ldb & 3) != 0 || (reinterpret_cast<uintptr_t>(a) & 15) != 0 || (reinterpret_cast<uintptr_t>(b) & 15) != 0 ||
      (d == nullptr && dt == nullptr) || (d != nullptr && ldd < n) || (dt != nullptr && ldt < (acc && aux != nullptr ? m - 1 : m)) ||
      (mode <= 1 && aux == nullptr) || (mode == 2 && aux == nullptr && relu_bits == nullptr) ||
      (relu_bits != nullptr && (mode == 0 || acc || ld_bits < (n + 31) / 32)) || (mode == 2 && aux != nullptr && ld_aux < n) || (acc && (dt == nullptr || d != nullptr)) ||
      split_k < 1 || (!acc && split_k != 1)) {
    return BLE_ERR_INVALID_ARGUMENT;
  }
  // row-major output through TMA stores when its rows can be a tensor map (16-byte pitch and base) and end on a 16-byte
  // boundary (the TMA writes whole 16-byte granules: measured, a 257-column tensor had columns 257..259 overwritten);
  // else direct stores
  const bool tma_store = !acc && d != nullptr && (ldd & 3) == 0 && (n & 3) == 0 && (reinterpret_cast<uintptr_t>(d) & 15) == 0;
  CUtensorMap ma, mb, md;
  // BLE_DENSE_ROWS=256 (A/B): 256-row tiles, two accumulators sharing every B tile -- a third less operand traffic per FLOP,
  // but one CTA per SM (no epilogue / main-loop overlap between CTAs) and 128 CTAs for 148 SMs: measured 0.872 ms per
  // 8,192-sample SGD step against 0.781 ms with the 128-row tiles, so the 128-row kernel is the default
  static const bool allow_256 = [] { const char* e = std::getenv("BLE_DENSE_ROWS"); return e != nullptr && std::atoi(e) == 256; }();
  const bool rows256 = allow_256 && !acc && ((m + 255) / 256) * ((n + kBN - 1) / kBN) >= 120;
  const int tile_rows = rows256 ? 256 : kBM;
  const bool maps_ok = mn ? operand_map(&ma, a, k, m, lda, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B) &&          // boxes {32 of MN, 32 of K}
                            operand_map(&mb, b, k, n, ldb, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)
                          : operand_map(&ma, a, m, k, lda, tile_rows) && operand_map(&mb, b, n, k, ldb, kBN);   // boxes {32 of K, tile rows}
  if (!maps_ok || !operand_map(&md, tma_store ? d : nullptr, m, n, ldd, 32)) return BLE_ERR_CUDA;
  const int total_kb = int((k + kBK - 1) / kBK);
  const int splits = split_k > total_kb ? total_kb : split_k;
  DenseArgs args{m, n, k, (total_kb + splits - 1) / splits, aux, ld_aux, d, ldd, dt, ldt, tma_store ? 1 : 0, relu_bits, ld_bits};
  const dim3 grid(unsigned((n + kBN - 1) / kBN), unsigned((m + tile_rows - 1) / tile_rows), unsigned(splits));
  cudaStream_t s = cudaStream_t(stream);
  switch (mode) {
    case 0: return rows256 ? launch_dense<0, 2>(ma, mb, md, args, grid, s) : launch_dense<0, 1>(ma, mb, md, args, grid, s);
    case 1: return rows256 ? launch_dense<1, 2>(ma, mb, md, args, grid, s) : launch_dense<1, 1>(ma, mb, md, args, grid, s);
    case 2: return rows256 ? launch_dense<2, 2>(ma, mb, md, args, grid, s) : launch_dense<2, 1>(ma, mb, md, args, grid, s);
    case 3: return launch_dense<3, 1>(ma, mb, md, args, grid, s);
    default: return launch_dense<4, 1>(ma, mb, md, args, grid, s);
  }
}

int ble_transpose_f32(const float* src, int64_t ld_src, int64_t rows, int64_t cols, float* dst, int64_t ld_dst, void* stream) {
  if (src == nullptr || dst == nullptr || rows <= 0 || cols <= 0 || ld_src < cols || ld_dst < rows) return BLE_ERR_INVALID_ARGUMENT;
  const dim3 grid(unsigned((cols + 31) / 32), unsigned((rows + 31) / 32));
  ble::k_transpose<<<grid, 256, 0, cudaStream_t(stream)>>>(src, ld_src, rows, cols, dst, ld_dst);
  return cudaGetLastError() == cudaSuccess ? BLE_OK : BLE_ERR_CUDA;
}

int ble_row_sum_f32(const float* src, int64_t ld_src, int64_t rows, int64_t cols, float* out, int32_t accumulate, void* stream) {
  if (src == nullptr || out == nullptr || rows <= 0 || cols <= 0 || ld_src < cols) return BLE_ERR_INVALID_ARGUMENT;
  ble::k_row_sum<<<unsigned(rows), 256, 0, cudaStream_t(stream)>>>(src, ld_src, rows, cols, out, accumulate);
  return cudaGetLastError() == cudaSuccess ? BLE_OK : BLE_ERR_CUDA;
}

}  // extern "C"
