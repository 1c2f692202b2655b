/* ble_b200.h -- C ABI of the B200-native batched Balloon Learning Environment transition function.
 *
 * Drop-in boundary for the reference's simulator injection point:
 *   BalloonEnv(arena: BalloonArenaInterface)            env/balloon_env.py:113,144-148
 *   BalloonArenaInterface.{reset, step, get/set_simulator_state, get/set_balloon_state,
 *                          get_measurements}             env/balloon_arena.py:42-120
 * (paths relative to /root/reference/balloon_learning_environment/).  The reference is pure
 * Python and has no FFI; INTEGRATION.md shows the ctypes binding a maintainer would add and the
 * `CudaBalloonArena(BalloonArenaInterface)` adaptor that sits on top of these entry points.
 *
 * Conventions
 *   - every function returns 0 on success or a negative ble_status; nothing throws or aborts
 *     across the ABI; ble_last_error() returns a static/handle-owned message.
 *   - all array arguments are caller-owned DEVICE pointers unless the name ends in `_host`;
 *     they must stay valid until the work queued on `stream` completes.  Calls are stream-ordered
 *     and asynchronous (no hidden synchronisation) except the `_host` variants, which return
 *     after the results have landed in host memory.
 *   - `stream` is a cudaStream_t passed as void* (0 = default stream).
 *   - one handle per GPU; a handle is not thread-safe.
 *   - N = number of balloons (environments) of the handle.  Stepping a balloon whose status is
 *     not OK is a no-op with reward 0 and done = 1 (the reference asserts instead,
 *     env/balloon/balloon.py:288).
 */
#ifndef BLE_B200_H_
#define BLE_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ble_handle ble_handle;

typedef enum {
  BLE_OK = 0,
  BLE_ERR_INVALID_ARGUMENT = -1,
  BLE_ERR_CUDA = -2,
  BLE_ERR_NOT_READY = -3,      /* e.g. stepping before fields / state were provided
                                  (reference: RuntimeError, env/grid_based_wind_field.py:86-87) */
  BLE_ERR_OUT_OF_MEMORY = -4,
  BLE_ERR_UNSUPPORTED = -5
} ble_status;

/* Arithmetic of the physics kernels. */
#define BLE_PRECISION_FP32 0   /* production: fp32 math, fp64 only for Julian time / height secant */
#define BLE_PRECISION_FP64 1   /* audit: follows the reference's fp64 arithmetic */

/* Forecast wind model (env/wind_field.py). */
#define BLE_WIND_GRID 0        /* GridBasedWindField: per-balloon [21,21,10,9,2] grids             */
#define BLE_WIND_SIMPLE_STATIC 1  /* SimpleStaticWindField (env/wind_field.py:149-184)             */

/* Device layout of the wind grids: every lookup reads one contiguous 128-byte window (its 16
 * corners x {u,v}).  X64 packs overlapping windows every 64 B (1,935,360 B per grid, ~1.5 HBM lines
 * per random lookup); X128 stores each window on its own 128-byte line (3,686,400 B per grid, exactly
 * one line per lookup). */
#define BLE_LAYOUT_X64 0
#define BLE_LAYOUT_X128 1

typedef struct {
  int32_t precision;           /* BLE_PRECISION_*                                                   */
  int32_t wind_model;          /* BLE_WIND_*                                                        */
  int32_t enable_noise;        /* 1: ground truth = forecast + simplex noise (wind_field.py:125-145) */
  int32_t field_layout;        /* BLE_LAYOUT_*                                                      */
  int32_t enable_features;     /* 1: keep the WindGP history and allow ble_features_perciatelli      */
  int32_t decoder_fp32;        /* 0 (default): the decoder GEMMs run on the tensor cores (TF32 inputs, fp32
                                * accumulate -- jax's default matmul precision on the reference's GPU
                                * path); 1: fp32 FMA, what the reference computes on a CPU            */
  int32_t auto_reset;          /* 1: a balloon whose step returned done = 1 (terminal status, or stepped while already
                                * finished) starts a new episode before ble_step returns control of the stream: reward /
                                * done / info describe the step that ended the episode, the state (and the WindGP history)
                                * is the new episode's.  Its seed is splitmix64 of the seed of the episode that ended (the
                                * chain starts at the seed given to ble_reset), so runs are reproducible.  The balloon keeps
                                * its wind field; a caller that wants a new field per episode regenerates it with
                                * ble_generate_fields_at for the balloons whose done flag it saw.  Not available inside
                                * ble_rollout.  0 (default): a finished balloon is frozen (stepping it is a no-op with
                                * reward 0, done 1; the reference asserts, env/balloon/balloon.py:288).              */
  int32_t reserved[1];         /* must be 0                                                         */
} ble_config;

/* State exchange: two row-major device matrices, one row per field, N columns.
 * f64 rows (BLE_F_*) mirror the float fields of BalloonState (env/balloon/balloon.py:73-208),
 * i64 rows (BLE_I_*) the discrete ones plus the three safety layers' internal state. */
enum {
  BLE_F_X = 0, BLE_F_Y, BLE_F_PRESSURE, BLE_F_AMBIENT_TEMPERATURE, BLE_F_INTERNAL_TEMPERATURE,
  BLE_F_ENVELOPE_VOLUME, BLE_F_SUPERPRESSURE, BLE_F_MOLS_AIR, BLE_F_MOLS_LIFT_GAS,
  BLE_F_BATTERY_CHARGE /* Wh */, BLE_F_ACS_POWER /* W */, BLE_F_ACS_MASS_FLOW, BLE_F_SOLAR_CHARGING,
  BLE_F_POWER_LOAD, BLE_F_CENTER_LAT /* rad */, BLE_F_CENTER_LNG /* rad */, BLE_F_UPWELLING_INFRARED,
  BLE_F_ATMOSPHERE_ALPHA /* lapse blend, env/balloon/standard_atmosphere.py:82-84 */,
  BLE_NUM_F
};
enum {
  BLE_I_DATE_TIME = 0 /* UNIX s */, BLE_I_TIME_ELAPSED /* s */, BLE_I_LAST_COMMAND, BLE_I_STATUS,
  BLE_I_ENVELOPE_STATE, BLE_I_ALTITUDE_STATE, BLE_I_POWER_PAUSED,
  BLE_I_SUNRISE_H /* PowerSafetyLayer._sunrise_with_hysteresis, UNIX s */, BLE_I_SUNSET,
  BLE_I_POWER_SAFETY_ENABLED,
  BLE_NUM_I
};
typedef struct {
  double* f64;    /* [BLE_NUM_F][N] */
  int64_t* i64;   /* [BLE_NUM_I][N] */
} ble_state_soa;

/* Lifetime. Replaces BalloonArena.__init__ (env/balloon_arena.py:126-159). */
int ble_create(int device, int64_t n_envs, const ble_config* config, ble_handle** out);
int ble_destroy(ble_handle* h);
const char* ble_last_error(const ble_handle* h /* may be NULL: error of the last failed ble_create */);
int64_t ble_num_envs(const ble_handle* h);

/* Wind fields: F grids in the reference's native layout float32 [F,21,21,10,9,2]
 * (GridWindFieldSampler.sample_field, env/grid_wind_field_sampler.py:33-41) and the balloon ->
 * grid map int32 [N] (may be NULL: keep the current map).  The handle keeps its own re-laid-out copy. */
int ble_upload_fields(ble_handle* h, const float* fields, int64_t n_fields,
                      const int32_t* env_to_field, void* stream);

/* The same in pieces, so that a large bank of grids never has to be resident twice:
 * allocate room for n_fields grids, convert `count` native-layout grids into slots
 * [first_field, first_field + count), set the balloon -> grid map int32 [N]. */
int ble_alloc_fields(ble_handle* h, int64_t n_fields, void* stream);
int ble_write_fields(ble_handle* h, const float* fields, int64_t first_field, int64_t count, void* stream);
int ble_set_field_map(ble_handle* h, const int32_t* env_to_field, void* stream);

/* VAE wind-field generator, reset path (generative/vae.py:134-186, env/generative_wind_field.py:52-62).
 * ble_set_decoder: the four Dense layers of vae.Decoder in flax layout -- kernels[i] float32 [in, out]
 *   row-major with (in, out) = (64,1000), (1000,1000), (1000,1000), (1000,4410); biases[i] float32 [out]
 *   (the arrays of models/offlineskies22_decoder.msgpack).  `kernels` / `biases` are HOST arrays of four
 *   DEVICE pointers; the handle keeps its own copy.
 * ble_decode_fields: latents float32 [F, 64] (the reference draws jax.random.normal) -> fields float32
 *   [F,21,21,10,9,2] in the native layout, ready for ble_write_fields.  Four cuBLASLt fp32 GEMMs with
 *   fused bias(+ReLU) epilogues, then resize 7->23 / central differences / crop in one kernel. */
int ble_set_decoder(ble_handle* h, const float* const* kernels, const float* const* biases, void* stream);
int ble_decode_fields(ble_handle* h, const float* latents, int64_t n_fields, float* fields, void* stream);

/* Simplex noise parameters, SimplexWindNoise.reset_wind_noise (env/wind_field.py:196-207,
 * env/simplex_wind_noise.py:98-114): seeds int64 [N,2,5] (component u/v, harmonic),
 * offsets float32 [N,2,5,4].  Builds the 256-entry permutation tables on the device. */
int ble_set_noise(ble_handle* h, const int64_t* seeds, const float* offsets, void* stream);

/* get/set_balloon_state (env/balloon_arena.py:213-220) for all N balloons. */
int ble_state_upload(ble_handle* h, const ble_state_soa* state, void* stream);
int ble_state_download(ble_handle* h, ble_state_soa* state, void* stream);

/* BalloonArena.reset (env/balloon_arena.py:161-182) for balloons with mask[i] != 0 (NULL = all):
 * samples atmosphere, start time, position, pressure, upwelling IR from seeds[i] (distributions of
 * utils/sampling.py:37-152 and env/balloon_arena.py:228-268), runs the power-safety sunrise/sunset
 * search and the stable-init solve (env/balloon/stable_init.py:132-157), and re-seeds the noise. */
int ble_reset(ble_handle* h, const uint64_t* seeds, const uint8_t* mask, void* stream);

/* Deterministic part of reset only: recompute sunrise/sunset + (optionally) stable-init from the
 * currently uploaded x, y, pressure, date_time, ... (BalloonState.__post_init__,
 * env/balloon/balloon.py:210-215; stable_init.cold_start_to_stable_params). */
int ble_init_derived(ble_handle* h, int32_t run_stable_init, void* stream);

/* BalloonEnv.step for all balloons (env/balloon_env.py:157-190 -> env/balloon_arena.py:184-202 ->
 * env/balloon/balloon.py:263-328): actions int32 [N] in {0 DOWN, 1 STAY, 2 UP};
 * reward float32 [N] (perciatelli_reward_function), done uint8 [N],
 * wind_uv float32 [N,2] = ground-truth wind applied during the step (may be NULL). */
int ble_step(ble_handle* h, const int32_t* actions, float* reward, uint8_t* done, float* wind_uv,
             void* stream);

/* ble_step with the `info` dictionary of BalloonEnv.step (env/balloon_env.py:186-190,280-290) written by the
 * step kernel itself, so a caller never needs a second pass over the state:
 *   status       uint8 [N]  BalloonStatus after the step: 0 OK, 1 OUT_OF_POWER, 2 BURST, 3 ZEROPRESSURE
 *                           (info['out_of_power'], info['envelope_burst'], info['zeropressure'])
 *   time_elapsed int32 [N]  seconds since reset (info['time_elapsed'])
 *   sim_error    uint8 [N]  1 once an atmosphere query of this balloon fell outside the table or a sunrise /
 *                           sunset search failed since the last reset (the reference asserts / raises there:
 *                           env/balloon/standard_atmosphere.py:95-96,126-127, env/balloon/solar.py:318-322); sticky.
 * reward and done are required, every other pointer may be NULL. */
typedef struct {
  float* reward;           /* [N]   */
  uint8_t* done;           /* [N]   */
  float* wind_uv;          /* [N,2] */
  uint8_t* status;         /* [N]   */
  int32_t* time_elapsed;   /* [N]   */
  uint8_t* sim_error;      /* [N]   */
  void* reserved[2];       /* must be NULL */
} ble_step_out;
int ble_step_ex(ble_handle* h, const int32_t* actions, const ble_step_out* out, void* stream);

/* n_steps consecutive BalloonEnv.step calls in ONE kernel launch, for open-loop action sequences (a random
 * agent, a replayed episode): actions int32 [n_steps][N]; out->reward / out->done are [n_steps][N], the other
 * outputs describe the state after the last step.  Balloons do not interact (env/balloon_arena.py:184-202),
 * so a CTA carries its 32 balloons through all the steps without any grid-wide synchronisation.  Not available
 * while the WindGP measurement history is being tracked (ble_features_track). */
int ble_rollout(ble_handle* h, const int32_t* actions, int32_t n_steps, const ble_step_out* out, void* stream);

/* Same, with HOST buffers (pageable or pinned): copies actions in, steps, copies reward/done out
 * and waits.  This is the call a Python/NumPy user of the reference makes per step.  Before it
 * returns it queues, on `stream`, the wind-noise kernel of the NEXT step (the wind at the post-step
 * state is the next step's pre-step wind, env/balloon_arena.py:184-202); any call that changes the
 * state discards that result.  Use one stream per handle for consecutive calls. */
int ble_step_host(ble_handle* h, const int32_t* actions_host, float* reward_host,
                  uint8_t* done_host, void* stream);

/* BalloonArena.get_measurements' wind_at_balloon (env/balloon_arena.py:222-226,270-275):
 * ground-truth wind at every balloon's CURRENT state, float32 [N,2]. */
int ble_wind_at_balloon(ble_handle* h, float* wind_uv, void* stream);

/* GridBasedWindField.get_forecast for M arbitrary points (env/grid_based_wind_field.py:70-94):
 * xyzt float32 [M,4] = (x km, y km, pressure Pa, elapsed hours), field_idx int32 [M] (grid index,
 * NOT balloon index), uv float32 [M,2].  Clip + time boomerang + fp32 point rounding as the
 * reference, 16-corner multilinear interpolation. */
int ble_wind_gather(ble_handle* h, const float* xyzt, const int32_t* field_idx, float* uv,
                    int64_t m, void* stream);

/* SimulatorState.wind_field for reference consumers (env/simulator_data.py:25-34): WindField.get_forecast
 * (with_noise = 0; env/wind_field.py:69-87) or get_ground_truth (with_noise = 1; :125-145) at M arbitrary points of
 * a BALLOON's own wind grid and noise generators.  xyzt float64 [M,4] = (x m, y m, pressure Pa, elapsed seconds) as
 * the reference passes them, env_idx int32 [M] = balloon index, uv float32 [M,2]. */
int ble_wind_query(ble_handle* h, const double* xyzt, const int32_t* env_idx, int32_t with_noise, float* uv,
                   int64_t m, void* stream);

/* SimulatorState.atmosphere: Atmosphere.at_pressure (which = 0, q = pressure Pa) / at_height (which = 1, q = metres)
 * of balloon env_idx[i]'s atmosphere (env/balloon/standard_atmosphere.py:89-154).  out float64 [M,4] = (height m,
 * temperature K, pressure Pa, density kg/m^3); a query outside the table, where the reference asserts (:95-96,
 * :126-127), yields a NaN row. */
int ble_atmosphere_query(ble_handle* h, int32_t which, const double* q, const int32_t* env_idx, double* out,
                         int64_t m, void* stream);

/* Observation surface: PerciatelliFeatureConstructor (env/features.py:269-581) for all balloons.
 * With enable_features = 1, ble_reset and ble_step end with the constructor's observe() of the new
 * state (env/balloon_arena.py:179-182, 201), i.e. the WindGP measurement history (env/wind_gp.py:98-119,
 * last 6 h = 120 measurements) is maintained by the handle.
 *   ble_features_perciatelli: get_features() -> obs float32 [N, 1099] (16 ambient features + 361 x 3
 *     wind-column features from the GP posterior, the forecast column and the reachable pressure range).
 *   ble_features_observe: observe() of the current state (only needed after ble_state_upload).
 *   ble_features_clear: forget every balloon's measurement history.
 *   ble_features_track: on = 0 stops ble_reset / ble_step from appending measurements (a rollout that does not
 *     read the observation then skips that kernel); on = 1 (default) resumes.  The history is only meaningful
 *     if tracking was on for every step since the last reset or clear. */
int ble_features_perciatelli(ble_handle* h, float* obs, void* stream);
int ble_features_observe(ble_handle* h, void* stream);
int ble_features_clear(ble_handle* h, void* stream);
int ble_features_track(ble_handle* h, int32_t on);

/* Derived properties of BalloonState at the CURRENT state (env/balloon/balloon.py:217-250), all
 * balloons, float64 [BLE_NUM_D][N]. */
enum {
  BLE_D_LAT = 0 /* rad, BalloonState.latlng */, BLE_D_LNG, BLE_D_SOLAR_ELEVATION /* deg */, BLE_D_SOLAR_FLUX,
  BLE_D_EXCESS_ENERGY /* 0/1 */, BLE_D_NAVIGATION_IS_PAUSED /* 0/1 */, BLE_D_PRESSURE_RATIO, BLE_D_BATTERY_SOC,
  BLE_D_ALTITUDE /* m, Atmosphere.at_pressure(pressure).height */,
  BLE_NUM_D
};
int ble_derived(ble_handle* h, double* out, void* stream);

/* ---- evaluation surface ------------------------------------------------
 * ble_generate_fields: GenerativeWindFieldSampler.sample_field (env/generative_wind_field.py:52-62) for
 *   `count` seeds: z ~ N(0, I_64) from Philox(seed) -> ble_set_decoder's network -> fields
 *   [first_field, first_field + count) of the bank allocated with ble_alloc_fields.  seeds: device
 *   uint64 [count].
 * ble_agent_station_seeker: StationSeekerAgent.pick_action (agents/station_seeker_agent.py:72-113) on
 *   obs float32 [N, 1099] (device) -> actions int32 [N] (0 DOWN, 1 STAY, 2 UP) and, if best_level is not
 *   NULL, the chosen level of the 361-level column (-1 where no level is valid; the reference asserts).
 * ble_agent_random_walk: RandomWalkAgent (agents/random_walk_agent.py:35-94).  step_index 0 is
 *   begin_episode (draws the target pressure), k >= 1 the k-th agent.step; the target pressures live in
 *   the handle, the random stream is Philox(seeds[e], step_index) (jax.random in the reference; its tests
 *   forbid depending on the stream).  seeds: device uint64 [N].
 * ble_eval_begin / ble_eval_accumulate / ble_eval_results: eval_lib.eval_agent's bookkeeping
 *   (eval/eval_lib.py:123-211) for N simultaneous flights.  begin: zero the sums, every balloon whose
 *   status is OK starts flying.  accumulate (after every ble_step, with that step's reward): reward sum,
 *   steps within the 50 km radius, step count, stop at a terminal status; flight_path (NULL or float32
 *   [6][N]: x km, y km, pressure, superpressure, elapsed seconds, battery soc -- SimpleBalloonState,
 *   eval_lib.py:58-80) receives this step's sample, NaN for balloons that already finished.
 *   results: float64 [BLE_NUM_E][N], the fields of EvaluationResult (eval_lib.py:84-116). */
enum {
  BLE_E_CUMULATIVE_REWARD = 0, BLE_E_TIME_WITHIN_RADIUS, BLE_E_OUT_OF_POWER, BLE_E_ENVELOPE_BURST, BLE_E_ZEROPRESSURE,
  BLE_E_FINAL_TIMESTEP, BLE_E_ACTIVE /* 1 = still flying when queried */,
  BLE_NUM_E
};
int ble_generate_fields(ble_handle* h, const uint64_t* seeds, int64_t first_field, int64_t count, void* stream);
/* Same for a scattered set of fields (the balloons whose episode just ended): field_index is device
 * int32 [count], each entry in [0, n_fields) and distinct. */
int ble_generate_fields_at(ble_handle* h, const uint64_t* seeds, const int32_t* field_index, int64_t count,
                           void* stream);
/* The latents ble_generate_fields decodes for these seeds: device float32 [count, 64], z ~ N(0, I_64) per seed
 * (env/generative_wind_field.py:57-58 draws jax.random.normal(key, (64,)); here Philox4x32-10 keyed by the seed).
 * ble_decode_fields(latents) is the native-layout view of what ble_generate_fields writes into the bank. */
int ble_sample_latents(ble_handle* h, const uint64_t* seeds, int64_t count, float* latents, void* stream);
int ble_agent_station_seeker(ble_handle* h, const float* obs, int32_t* actions, int32_t* best_level, void* stream);
int ble_agent_random_walk(ble_handle* h, const float* obs, const uint64_t* seeds, int32_t step_index,
                          int32_t* actions, void* stream);
int ble_eval_begin(ble_handle* h, void* stream);
int ble_eval_accumulate(ble_handle* h, const float* reward, float* flight_path, void* stream);
int ble_eval_results(ble_handle* h, double* out, void* stream);

/* ---- QR-DQN learner surface (BASELINE configs[4]) ----------------------
 * Stateless kernels around the quantile network the reference trains with Acme's DQN builder
 * (acme_utils.py:217-277: 8 layers x 600 units, 3 actions x 51 atoms, QrDqn(huber_param=1), n_step 5,
 * discount 0.993, Adam 2e-6 / eps 2e-5, target period 25 learner steps) or Dopamine's JaxQuantileAgent
 * (agents/quantile_agent.py:37-160, agents/configs/quantile.gin).  Every pointer is caller-owned DEVICE
 * memory; no handle is needed.  The dense layers are library GEMMs on the caller's side.
 *   ble_qr_greedy: actions[b] = argmax_a mean_j logits[b, a, j] (first maximum), the epsilon = 0
 *     behaviour / eval policy of acme_utils.py:250-268; q_values float32 [B, A] or NULL.
 *   ble_qr_target: target[b, j] = reward[b] + discount[b] * next_logits[b, a*, j] with a* the greedy action
 *     of next_logits (dopamine quantile_agent.target_distribution; discount = gamma^n * (1 - terminal)).
 *   ble_qr_loss: loss[b] = (1/N) sum_i sum_j |tau_i - 1{d_ij < 0}| huber_kappa(d_ij), d_ij = target[b, j] -
 *     logits[b, actions[b], i], tau_i = (i + 0.5) / N (dopamine quantile_agent.train == rlax
 *     quantile_q_learning for kappa = 1).  If grad_logits is not NULL it receives
 *     grad_scale * weight[b] * d loss[b] / d logits[b, :, :] (zero rows for the other actions); weight may be
 *     NULL (= 1).  num_atoms <= 64.
 *   ble_replay_sample: draws `batch` n-step transitions from a time-major ring of N-balloon steps.
 *     A window that reaches a terminal step is cut there (discount 0); one that crosses a step-limit
 *     truncation, the write cursor or the ring's oldest step is redrawn (up to 32 times; valid[b] = 0 if
 *     none was found).  forced_indices (int64 [B, 2] = absolute step, balloon) replaces the random draw
 *     when not NULL; picked (int64 [B, 2] or NULL) reports the indices used.
 *   ble_adam_step: optax.adam on a flat buffer: m, v moments, step counted from 1, grads multiplied by
 *     grad_scale first (1 / world_size after a summing all-reduce).
 *   ble_marco_polo_step: MarcoPoloExploration (agents/marco_polo_exploration.py:36-93) wrapped around
 *     RandomWalkAgent (agents/random_walk_agent.py:35-94) the way acme_utils.CombinedActor does
 *     (acme_utils.py:161-183): begin[e] != 0 marks the first observation of an episode.  state: int32
 *     [4][N] (exploratory_episode, exploratory_phase, phase_elapsed_s, walk_elapsed_s); walk_target:
 *     float64 [N]; random stream Philox(seeds[e], step_index). */
typedef struct ble_replay_view {
  const float* obs;            /* [capacity][N][num_features] observation the action was chosen on */
  const int32_t* action;       /* [capacity][N] */
  const float* reward;         /* [capacity][N] */
  const uint8_t* terminal;     /* [capacity][N] 1: the environment ended the episode after this step */
  const uint8_t* truncated;    /* [capacity][N] 1: the step limit ended the episode after this step  */
  int64_t capacity;            /* ring length in steps                                               */
  int64_t num_envs;            /* N                                                                  */
  int64_t count;               /* steps written so far (monotonic; slot = step % capacity)           */
  int32_t n_step;              /* update horizon (5)                                                 */
  int32_t num_features;        /* 1099                                                               */
  float gamma;                 /* 0.993                                                              */
  int32_t out_pitch;           /* row pitch (floats) of the state / next_state outputs; 0 = num_features */
} ble_replay_view;
int ble_qr_greedy(const float* logits, int64_t batch, int32_t num_actions, int32_t num_atoms, int32_t* actions,
                  float* q_values, void* stream);
int ble_qr_target(const float* next_logits, const float* reward, const float* discount, int64_t batch,
                  int32_t num_actions, int32_t num_atoms, float* target, void* stream);
int ble_qr_loss(const float* logits, const int32_t* actions, const float* target, const float* weight, float kappa,
                int64_t batch, int32_t num_actions, int32_t num_atoms, float grad_scale, float* loss,
                float* grad_logits, void* stream);
int ble_replay_sample(const ble_replay_view* view, const int64_t* forced_indices, uint64_t seed, int64_t batch,
                      float* state, float* next_state, int32_t* action, float* n_step_return, float* discount,
                      uint8_t* valid, int64_t* picked, void* stream);
int ble_adam_step(float* params, const float* grads, float* m, float* v, int64_t count, double learning_rate,
                  double beta1, double beta2, double eps, int64_t step, float grad_scale, void* stream);
int ble_marco_polo_step(const float* obs, const int32_t* rl_actions, const uint8_t* begin, int64_t num_envs,
                        int32_t* state, double* walk_target, const uint64_t* seeds, int64_t step_index,
                        float exploratory_episode_probability, int32_t* actions, void* stream);

/* Dense layers of the QuantileNetwork (agents/networks.py:63-98; the reference evaluates them as flax nn.Dense inside
 * dopamine's / acme's jitted train step, agents/quantile_agent.py:122-139, acme_utils.py:217-277) on the tcgen05
 * tensor cores (TF32 operands, fp32 accumulation).  Stateless; every pointer is caller-owned device memory.
 *   ble_dense_tf32: D[m, n] = A[m, k] . B[n, k]^T, A and B row-major with K contiguous (pitches lda, ldb in floats,
 *     multiples of 4; base addresses 16-byte aligned).  mode 0: D += aux[n] (bias); 1: bias then ReLU; 2: D *= (aux[m, n]
 *     > 0) with aux row-major of pitch ld_aux (the ReLU mask of the backward pass); 3: the TRANSPOSED result dt is ACCUMULATED
 *     atomically and the K range is split over split_k CTAs per tile (weight gradient; the caller zeroes dt; if aux is not
 *     NULL the LAST row of A is taken to be a row of ones appended by the caller and its results -- the column sums of
 *     B^T, i.e. the bias gradient -- are accumulated into aux[n] instead of dt).  d (pitch
 *     ldd) and / or the transposed result dt [n, m] (pitch ldt) are written (mode 3: dt only).
 *     mode 4 = mode 3 with the operands given the other way round in memory: a = [k, m] and b = [k, n] row-major
 *     (pitches lda >= m, ldb >= n), i.e. D = A^T . B for row-major A [k, m], B [k, n] -- the weight gradient straight from
 *     the row-major activations and output gradients (MN-major tensor-core operands), no transposed copies.
 *     relu_bits (or NULL): the ReLU mask packed 32 columns per word, [m, ld_bits] -- mode 1 WRITES it (bit j of word w of
 *     row r = D[r, 32 w + j] > 0), mode 2 READS it instead of aux when it is not NULL.
 *   ble_transpose_f32: dst[c, r] = src[r, c].   ble_row_sum_f32: out[r] (+)= sum_c src[r, c] (bias gradient). */
int ble_dense_tf32(const float* a, int64_t lda, const float* b, int64_t ldb, int64_t m, int64_t n, int64_t k, int32_t mode,
                   const float* aux, int64_t ld_aux, float* d, int64_t ldd, float* dt, int64_t ldt, int32_t split_k,
                   uint32_t* relu_bits, int64_t ld_bits, void* stream);
int ble_transpose_f32(const float* src, int64_t ld_src, int64_t rows, int64_t cols, float* dst, int64_t ld_dst, void* stream);
int ble_row_sum_f32(const float* src, int64_t ld_src, int64_t rows, int64_t cols, float* out, int32_t accumulate, void* stream);

/* Number of kernel launches issued by this handle so far (bench.py's gpu_launches). */
int64_t ble_launch_count(const ble_handle* h);

#ifdef __cplusplus
}
#endif
#endif  /* BLE_B200_H_ */
