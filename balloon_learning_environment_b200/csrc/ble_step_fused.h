// Host-side interface of the fused step kernel (ble_step_fused.cu), used by the engine (ble_engine.cu).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

#include "ble_devstate.cuh"

namespace ble {

struct FusedOut {
  float* reward;            // [steps][n]
  uint8_t* done;            // [steps][n]
  float2* wind_uv;          // [n] (state after the last step) or nullptr
  uint8_t* status;          // [n] or nullptr: BalloonStatus after the step (info: out_of_power / burst / zeropressure)
  int32_t* time_elapsed;    // [n] or nullptr: seconds since reset (info['time_elapsed'])
  uint8_t* sim_error;       // [n] or nullptr: sticky "atmosphere query out of range / search failed" flag
};

// Kernel shapes: 0 = k_step_warp (throughput shape: one warp carries 32 balloons through the whole step, 4 warps / CTA);
// 4 / 8 / 14 = k_step_roles<kW> (latency shape: kW warps share the 32 balloons of a CTA).
constexpr int kFusedShapes[4] = {0, 4, 8, 14};

// Opts every shape into its dynamic shared memory and reports how many of its CTAs one SM holds (same order).
cudaError_t fused_setup(int blocks_per_sm[4]);
// noise_mode: 0 = no noise, 1 = evaluate the 10 harmonics in the kernel, 2 = read d.noise_partial (k_noise ran ahead)
// returns the number of kernel launches issued (a rollout at the throughput shape is n_steps dependent launches)
int fused_launch(int shape, const DevState<float>& d, const int32_t* actions, const FusedOut& out, int noise_mode,
                 int n_steps, cudaStream_t stream);

}  // namespace ble
