// QR-DQN learner surface for vectorised rollouts (BASELINE configs[4]).
//
// The reference trains a quantile network with Acme's DQN builder (acme_utils.py:217-277,
// train_acme_qrdqn.py:43-81) or Dopamine's JaxQuantileAgent (agents/quantile_agent.py:37-160) on
// 1099-float Perciatelli observations.  Here the N parallel balloons write straight into a device
// replay ring and the kernels below do everything around the dense layers:
//   k_qr_greedy        argmax_a mean_j theta[a, j]                      (behaviour / eval policy, acme_utils.py:250-268)
//   k_qr_target        r + discount * theta'[a*, :]                     (dopamine target_distribution)
//   k_qr_loss          quantile-Huber loss + its gradient in one pass   (dopamine train / rlax.quantile_q_learning)
//   k_replay_sample    n-step transition assembly + observation gather  (n_step = 5, acme_utils.py:224)
//   k_adam             optax.adam on a flat parameter buffer            (acme_utils.py:225,233)
//   k_marco_polo       MarcoPoloExploration + RandomWalkAgent per balloon (agents/marco_polo_exploration.py:36-93)
// The dense layers themselves are plain library GEMMs (cuBLAS through torch.nn.functional.linear).
// All entry points are stateless: caller-owned device pointers, sizes and a stream.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdint>

#include "../../include/ble_b200.h"

namespace ble {

constexpr double kPi = 3.14159265358979323846;
#include "ble_rng.cuh"

namespace {

constexpr int kMaxAtoms = 64;          // two atoms per lane
constexpr int kWarpsPerBlock = 4;
constexpr int kNumFeatures = 1099;     // env/features.py:291

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// One warp per sample: q[a] = mean over atoms, first maximum wins (jnp.argmax).
__device__ __forceinline__ int greedy_of(const float* __restrict__ logits, int a_count, int n_atoms, int lane,
                                         float* __restrict__ q_out) {
  int best = 0;
  float best_q = 0.f;
  for (int a = 0; a < a_count; ++a) {
    float s = 0.f;
    for (int j = lane; j < n_atoms; j += 32) s += logits[a * n_atoms + j];
    const float q = warp_sum(s) / float(n_atoms);
    if (q_out != nullptr && lane == 0) q_out[a] = q;
    if (a == 0 || q > best_q) { best_q = q; best = a; }
  }
  return best;
}

__global__ void __launch_bounds__(32 * kWarpsPerBlock)
k_qr_greedy(const float* __restrict__ logits, int64_t b, int a_count, int n_atoms, int32_t* __restrict__ actions,
            float* __restrict__ q_values) {
  const int64_t s = blockIdx.x * int64_t(kWarpsPerBlock) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (s >= b) return;
  const int best = greedy_of(logits + s * a_count * n_atoms, a_count, n_atoms, lane,
                             q_values != nullptr ? q_values + s * a_count : nullptr);
  if (lane == 0) actions[s] = best;
}

__global__ void __launch_bounds__(32 * kWarpsPerBlock)
k_qr_target(const float* __restrict__ next_logits, const float* __restrict__ reward, const float* __restrict__ discount,
            int64_t b, int a_count, int n_atoms, float* __restrict__ target) {
  const int64_t s = blockIdx.x * int64_t(kWarpsPerBlock) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (s >= b) return;
  const float* row = next_logits + s * a_count * n_atoms;
  const int best = greedy_of(row, a_count, n_atoms, lane, nullptr);
  const float r = reward[s], g = discount[s];
  for (int j = lane; j < n_atoms; j += 32) target[s * n_atoms + j] = r + g * row[best * n_atoms + j];
}

// One warp per sample.  Lane l owns source quantiles i = l and i = l + 32; the N targets are held
// two per lane and broadcast by shuffle, so the N x N pairwise terms never touch memory.
__global__ void __launch_bounds__(32 * kWarpsPerBlock)
k_qr_loss(const float* __restrict__ logits, const int32_t* __restrict__ actions, const float* __restrict__ target,
          const float* __restrict__ weight, float kappa, int64_t b, int a_count, int n_atoms, float grad_scale,
          float* __restrict__ loss, float* __restrict__ grad) {
  const int64_t s = blockIdx.x * int64_t(kWarpsPerBlock) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (s >= b) return;
  int act = actions[s];
  act = act < 0 ? 0 : (act >= a_count ? a_count - 1 : act);
  const float* theta = logits + (s * a_count + act) * n_atoms;
  const float* tgt = target + s * n_atoms;
  const float t0 = lane < n_atoms ? tgt[lane] : 0.f;
  const float t1 = lane + 32 < n_atoms ? tgt[lane + 32] : 0.f;
  const bool has0 = lane < n_atoms, has1 = lane + 32 < n_atoms;
  const float th0 = has0 ? theta[lane] : 0.f, th1 = has1 ? theta[lane + 32] : 0.f;
  const float tau0 = (float(lane) + 0.5f) / float(n_atoms), tau1 = (float(lane + 32) + 0.5f) / float(n_atoms);
  float l0 = 0.f, l1 = 0.f, g0 = 0.f, g1 = 0.f;
  for (int j = 0; j < n_atoms; ++j) {
    const float tj = __shfl_sync(0xffffffffu, j < 32 ? t0 : t1, j & 31);
    {
      const float d = tj - th0, ad = fabsf(d);
      const float w = fabsf(tau0 - (d < 0.f ? 1.f : 0.f));
      l0 += w * (ad <= kappa ? 0.5f * d * d : kappa * (ad - 0.5f * kappa));
      g0 += w * (ad <= kappa ? d : copysignf(kappa, d));
    }
    {
      const float d = tj - th1, ad = fabsf(d);
      const float w = fabsf(tau1 - (d < 0.f ? 1.f : 0.f));
      l1 += w * (ad <= kappa ? 0.5f * d * d : kappa * (ad - 0.5f * kappa));
      g1 += w * (ad <= kappa ? d : copysignf(kappa, d));
    }
  }
  const float wgt = weight != nullptr ? weight[s] : 1.f;
  const float total = warp_sum((has0 ? l0 : 0.f) + (has1 ? l1 : 0.f)) / float(n_atoms);
  if (lane == 0) loss[s] = total;
  if (grad != nullptr) {
    float* grow = grad + s * a_count * n_atoms;
    const float scale = -grad_scale * wgt / float(n_atoms);
    for (int a = 0; a < a_count; ++a) {
      if (has0) grow[a * n_atoms + lane] = a == act ? scale * g0 : 0.f;
      if (has1) grow[a * n_atoms + lane + 32] = a == act ? scale * g1 : 0.f;
    }
  }
}

// ---- replay ---------------------------------------------------------------------------------
struct ReplayView {
  const float* obs;            // [capacity][E][F]
  const int32_t* action;       // [capacity][E]
  const float* reward;         // [capacity][E]
  const uint8_t* terminal;     // [capacity][E]  env ended the episode after this step (bootstrap cut)
  const uint8_t* truncated;    // [capacity][E]  step limit ended the episode after this step
  int64_t capacity, envs, count;
  int n_step;
  float gamma;
};

// Returns n_used (> 0) or 0 when (t, e) cannot be sampled; see oracle/qrdqn.py:nstep_transition.
__device__ int nstep_at(const ReplayView& r, int64_t t, int64_t e, float* ret, float* disc, int64_t* t_next) {
  const int64_t oldest = r.count > r.capacity ? r.count - r.capacity : 0;
  if (t < oldest || t >= r.count) return 0;
  float acc = 0.f, g = 1.f;
  for (int k = 0; k < r.n_step; ++k) {
    if (t + k >= r.count) return 0;
    const int64_t idx = ((t + k) % r.capacity) * r.envs + e;
    acc += g * r.reward[idx];
    g *= r.gamma;
    if (r.terminal[idx]) {
      *ret = acc; *disc = 0.f;
      *t_next = t + k + 1 < r.count ? t + k + 1 : t;
      return k + 1;
    }
    if (r.truncated[idx]) return 0;
  }
  if (t + r.n_step >= r.count) return 0;
  *ret = acc; *disc = g; *t_next = t + r.n_step;
  return r.n_step;
}

// One warp per sample: lane 0 draws (t, e) until the window is valid, then the warp copies the two
// observation rows (coalesced 4-byte lanes; rows are 4,396 B so only 4-byte alignment is guaranteed).
__global__ void __launch_bounds__(32 * kWarpsPerBlock)
k_replay_sample(ReplayView r, const int64_t* __restrict__ forced /*[B,2] (t, e) or nullptr*/, uint64_t seed, int64_t b,
                int features, int pitch, float* __restrict__ state, float* __restrict__ next_state, int32_t* __restrict__ action,
                float* __restrict__ ret_out, float* __restrict__ disc_out, uint8_t* __restrict__ valid,
                int64_t* __restrict__ picked /*[B,2] or nullptr*/) {
  const int64_t s = blockIdx.x * int64_t(kWarpsPerBlock) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (s >= b) return;
  int64_t t = 0, e = 0, t_next = 0;
  float ret = 0.f, disc = 0.f;
  int n_used = 0;
  if (lane == 0) {
    if (forced != nullptr) {
      t = forced[2 * s]; e = forced[2 * s + 1];
      if (e >= 0 && e < r.envs) n_used = nstep_at(r, t, e, &ret, &disc, &t_next);
    } else if (r.count > 0) {
      Philox rng;
      rng.init(seed, uint64_t(s));
      const int64_t oldest = r.count > r.capacity ? r.count - r.capacity : 0;
      const int64_t span = r.count - oldest;
      for (int attempt = 0; attempt < 32 && n_used == 0; ++attempt) {
        t = oldest + int64_t(rng.uniform() * double(span));
        e = int64_t(rng.uniform() * double(r.envs));
        t = t >= r.count ? r.count - 1 : t;
        e = e >= r.envs ? r.envs - 1 : e;
        n_used = nstep_at(r, t, e, &ret, &disc, &t_next);
      }
    }
    if (n_used == 0) { t_next = t = 0; e = 0; ret = 0.f; disc = 0.f; }
  }
  t = __shfl_sync(0xffffffffu, t, 0); e = __shfl_sync(0xffffffffu, e, 0);
  t_next = __shfl_sync(0xffffffffu, t_next, 0);
  n_used = __shfl_sync(0xffffffffu, n_used, 0);
  const float* src0 = r.obs + ((t % r.capacity) * r.envs + e) * features;
  const float* src1 = r.obs + ((t_next % r.capacity) * r.envs + e) * features;
  float* dst0 = state + s * pitch;
  float* dst1 = next_state + s * pitch;
  for (int f = lane; f < features; f += 32) { dst0[f] = src0[f]; dst1[f] = src1[f]; }
  if (lane == 0) {
    action[s] = r.action[(t % r.capacity) * r.envs + e];
    ret_out[s] = ret; disc_out[s] = disc;
    valid[s] = n_used > 0 ? 1 : 0;
    if (picked != nullptr) { picked[2 * s] = t; picked[2 * s + 1] = e; }
  }
}

// ---- optimiser -------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_adam(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, int64_t count,
       float lr, float b1, float b2, float omb1, float omb2, float eps, float inv_c1, float inv_c2, float grad_scale) {
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < count; i += stride) {
    const float gi = g[i] * grad_scale;
    const float mi = b1 * m[i] + omb1 * gi;            // 1 - beta formed in double on the host (1.f - 0.999f is 4.7e-5 off)
    const float vi = b2 * v[i] + omb2 * gi * gi;
    m[i] = mi; v[i] = vi;
    p[i] -= lr * (mi * inv_c1) / (sqrtf(vi * inv_c2) + eps);
  }
}

// ---- exploration -----------------------------------------------------------------------------
// state rows (int32 [4][E]): exploratory_episode, exploratory_phase, phase_elapsed_s, walk_elapsed_s.
__global__ void __launch_bounds__(128)
k_marco_polo(const float* __restrict__ obs, const int32_t* __restrict__ rl_actions, const uint8_t* __restrict__ begin,
             int64_t n, int32_t* __restrict__ st, double* __restrict__ walk_target, const uint64_t* __restrict__ seeds,
             int64_t step_index, float probability, int32_t* __restrict__ actions) {
  const int64_t e = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (e >= n) return;
  Philox rng;
  rng.init(seeds[e], (uint64_t(2) << 32) + uint64_t(step_index));
  int episode = st[e], phase = st[n + e], phase_s = st[2 * n + e], walk_s = st[3 * n + e];
  double target = walk_target[e];
  if (begin != nullptr && begin[e]) {                  // CombinedActor.observe_first (acme_utils.py:172-174)
    walk_s = 0;
    target = 6500.0 + (11400.0 - 6500.0) * rng.uniform();          // random_walk_agent.py:57-60,75-78
    phase_s = 0;
    episode = rng.uniform() <= double(probability) ? 1 : 0;        // marco_polo_exploration.py:57-59
    phase = 0;
  }
  phase_s += 180;                                      // _update_phase (:75-84), AGENT_TIME_STEP = 3 min
  if (episode) {
    const int limit = phase ? 2 * 3600 : 4 * 3600;     // :35-36
    if (phase_s >= limit) { phase ^= 1; phase_s = 0; }
  }
  int action = rl_actions[e];
  if (phase) {                                         // RandomWalkAgent.step (:80-91) + _select_action (:62-73)
    walk_s += 180;
    target += double(walk_s) * 0.1666 * rng.normal();
    const double p = double(obs[e * int64_t(kNumFeatures)]) * (14000.0 - 5000.0) + 5000.0;
    action = p - 100.0 > target ? 2 : (p + 100.0 < target ? 0 : 1);
  }
  st[e] = episode; st[n + e] = phase; st[2 * n + e] = phase_s; st[3 * n + e] = walk_s;
  walk_target[e] = target;
  actions[e] = action;
}

inline unsigned warp_grid(int64_t samples) { return unsigned((samples + kWarpsPerBlock - 1) / kWarpsPerBlock); }

inline int finish() { return cudaGetLastError() == cudaSuccess ? BLE_OK : BLE_ERR_CUDA; }

}  // namespace
}  // namespace ble

extern "C" {

int ble_qr_greedy(const float* logits, int64_t batch, int32_t num_actions, int32_t num_atoms, int32_t* actions,
                  float* q_values, void* stream) {
  if (logits == nullptr || actions == nullptr || batch < 0 || num_actions <= 0 || num_atoms <= 0) return BLE_ERR_INVALID_ARGUMENT;
  if (batch == 0) return BLE_OK;
  ble::k_qr_greedy<<<ble::warp_grid(batch), 32 * ble::kWarpsPerBlock, 0, cudaStream_t(stream)>>>(
      logits, batch, num_actions, num_atoms, actions, q_values);
  return ble::finish();
}

int ble_qr_target(const float* next_logits, const float* reward, const float* discount, int64_t batch, int32_t num_actions,
                  int32_t num_atoms, float* target, void* stream) {
  if (next_logits == nullptr || reward == nullptr || discount == nullptr || target == nullptr || batch < 0 ||
      num_actions <= 0 || num_atoms <= 0) return BLE_ERR_INVALID_ARGUMENT;
  if (batch == 0) return BLE_OK;
  ble::k_qr_target<<<ble::warp_grid(batch), 32 * ble::kWarpsPerBlock, 0, cudaStream_t(stream)>>>(
      next_logits, reward, discount, batch, num_actions, num_atoms, target);
  return ble::finish();
}

int ble_qr_loss(const float* logits, const int32_t* actions, const float* target, const float* weight, float kappa,
                int64_t batch, int32_t num_actions, int32_t num_atoms, float grad_scale, float* loss, float* grad_logits,
                void* stream) {
  if (logits == nullptr || actions == nullptr || target == nullptr || loss == nullptr || batch < 0 || num_actions <= 0 ||
      num_atoms <= 0 || num_atoms > ble::kMaxAtoms || !(kappa > 0.f)) return BLE_ERR_INVALID_ARGUMENT;
  if (batch == 0) return BLE_OK;
  ble::k_qr_loss<<<ble::warp_grid(batch), 32 * ble::kWarpsPerBlock, 0, cudaStream_t(stream)>>>(
      logits, actions, target, weight, kappa, batch, num_actions, num_atoms, grad_scale, loss, grad_logits);
  return ble::finish();
}

int ble_replay_sample(const ble_replay_view* view, const int64_t* forced_indices, uint64_t seed, int64_t batch,
                      float* state, float* next_state, int32_t* action, float* n_step_return, float* discount,
                      uint8_t* valid, int64_t* picked, void* stream) {
  if (view == nullptr || view->obs == nullptr || view->action == nullptr || view->reward == nullptr ||
      view->terminal == nullptr || view->truncated == nullptr || view->capacity <= 0 || view->num_envs <= 0 ||
      view->count < 0 || view->n_step <= 0 || view->num_features <= 0 ||
      (view->out_pitch != 0 && view->out_pitch < view->num_features) || state == nullptr || next_state == nullptr ||
      action == nullptr || n_step_return == nullptr || discount == nullptr || valid == nullptr || batch < 0) {
    return BLE_ERR_INVALID_ARGUMENT;
  }
  if (batch == 0) return BLE_OK;
  ble::ReplayView r{view->obs, view->action, view->reward, view->terminal, view->truncated,
                    view->capacity, view->num_envs, view->count, view->n_step, view->gamma};
  ble::k_replay_sample<<<ble::warp_grid(batch), 32 * ble::kWarpsPerBlock, 0, cudaStream_t(stream)>>>(
      r, forced_indices, seed, batch, view->num_features, view->out_pitch > 0 ? view->out_pitch : view->num_features, state,
      next_state, action, n_step_return, discount, valid, picked);
  return ble::finish();
}

int ble_adam_step(float* params, const float* grads, float* m, float* v, int64_t count, double learning_rate, double beta1,
                  double beta2, double eps, int64_t step, float grad_scale, void* stream) {
  if (params == nullptr || grads == nullptr || m == nullptr || v == nullptr || count < 0 || step < 1) return BLE_ERR_INVALID_ARGUMENT;
  if (count == 0) return BLE_OK;
  const double c1 = std::pow(beta1, double(step)), c2 = std::pow(beta2, double(step));
  const float inv_c1 = float(1.0 / (1.0 - c1)), inv_c2 = float(1.0 / (1.0 - c2));
  const unsigned grid = unsigned(std::min<int64_t>((count + 255) / 256, 148 * 8));
  ble::k_adam<<<grid, 256, 0, cudaStream_t(stream)>>>(params, grads, m, v, count, float(learning_rate), float(beta1), float(beta2),
                                                      float(1.0 - beta1), float(1.0 - beta2), float(eps),
                                                      inv_c1, inv_c2, grad_scale);
  return ble::finish();
}

int ble_marco_polo_step(const float* obs, const int32_t* rl_actions, const uint8_t* begin, int64_t num_envs,
                        int32_t* state, double* walk_target, const uint64_t* seeds, int64_t step_index,
                        float exploratory_episode_probability, int32_t* actions, void* stream) {
  if (obs == nullptr || rl_actions == nullptr || state == nullptr || walk_target == nullptr || seeds == nullptr ||
      actions == nullptr || num_envs < 0 || step_index < 0) return BLE_ERR_INVALID_ARGUMENT;
  if (num_envs == 0) return BLE_OK;
  ble::k_marco_polo<<<unsigned((num_envs + 127) / 128), 128, 0, cudaStream_t(stream)>>>(
      obs, rl_actions, begin, num_envs, state, walk_target, seeds, step_index, exploratory_episode_probability, actions);
  return ble::finish();
}

}  // extern "C"
