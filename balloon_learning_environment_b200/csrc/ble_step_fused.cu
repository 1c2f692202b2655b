// The one-launch step kernels of the production (fp32) build and their launcher.
//
// Reference path: env/balloon_env.py:157-190 -> env/balloon_arena.py:184-202 -> WindField.get_ground_truth
// (env/wind_field.py:125-145: grid interpolation + 10 simplex-noise harmonics) -> Balloon.simulate_step
// (env/balloon/balloon.py:263-328: three safety layers, 18 Euler sub-steps) -> reward / terminal.
//
// Two shapes of the same computation, both one launch per BalloonEnv.step (or per ble_rollout of many steps):
//
// k_step_roles<kW>  LATENCY shape, for batches of at most one wave (the 8-GPU split of 65,536 balloons).
//   32 balloons per CTA, kW warps; lane l of every warp = balloon 32 * block + l.
//   phase 1a  13 independent TASKS dealt round-robin to the kW warps:
//               N0..N9  one simplex-noise harmonic each; the 32 balloons' permutation tables of that harmonic
//                       (8 KB, contiguous in HBM) arrive in the warp's staging buffer by ONE TMA bulk copy
//                       (cp.async.bulk + mbarrier) issued before the lattice coordinates are computed;
//               ST0..1  the time-only half of the NOAA solar calculator at t0 and t0 + 180 s;
//               G       forecast gather (8 x LDG.128 of the lookup's 128-byte window), power / envelope /
//                       altitude safety layers -> effective action, and X = (p/P_i)^k for role P.
//   barrier   (whole CTA)
//   phase 1b  wind = forecast + blended noise; warps 0..2 finish the sun at the three track points
//             (position-dependent half: great-circle offset + hour angle), which needs the wind.
//   barrier   (4 role warps; the other warps are done)
//   phase 2   18 Euler sub-steps, warp w = role w (P pressure/atmosphere, T thermal body, E envelope + ACS,
//             S sun + power) on the branch-free role functions of ble_step_roles.cuh.  Every role runs ITS OWN
//             loop (so the register allocation is the maximum over the roles, not their sum); one named barrier
//             and one double-buffered shared-memory exchange per sub-step; each role stores the rows it owns.
//
// k_step_warp       THROUGHPUT shape, for batches of several waves (65,536 balloons on one GPU).
//   One WARP carries 32 balloons through the whole step: no CTA barrier, no shared-memory exchange, no duplicated
//   work; the four roles run back to back in every thread and the compiler interleaves their independent chains.
//   The only shared memory is the warp's 8 KB permutation-table staging buffer (same TMA bulk copy).
//
// n_steps > 1 (ble_rollout): the same CTA runs the steps back to back on actions[step][N]; balloons do not
// interact, so there is no grid-wide synchronisation and no launch gap between steps.
#include "ble_step_fused.h"
#include <algorithm>
#include <cstdlib>

#include "ble_step_roles.cuh"

namespace ble {

constexpr int kFusedTasks = 13;
constexpr int kPermStageBytes = 32 * 256;

// Named barrier of the four role warps.  `bar.sync` is an ALIGNED barrier: every thread of a warp has to execute it
// together, and after an `if (live) { ... }` whose condition differs between lanes nothing makes the lanes reconverge by
// themselves (compute-sanitizer --tool synccheck: "divergent thread(s) in block") -- hence the __syncwarp() first.
__device__ __forceinline__ void role_barrier() {
  __syncwarp();
  asm volatile("bar.sync 1, 128;" ::: "memory");
}

__device__ __forceinline__ void mbar_init(uint32_t bar) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}

// One TMA bulk copy: the permutation tables of harmonic h10 for the `count` balloons starting at e0.
__device__ __forceinline__ void stage_perm_tables(const DevState<float>& d, int h10, int64_t e0, int count, uint8_t* stage,
                                                  uint32_t bar) {
  const uint32_t bytes = uint32_t(count) * 256u;
  const uint8_t* src = d.perm + (int64_t(h10) * d.n + e0) * 256;
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
  // the tables stream through once per step (168 MB at 65,536 balloons, more than the L2): evict-first keeps them from
  // pushing the 26 MB of state rows, which every step re-reads, out of the L2 (81.0 -> 77.8 us at 32,768 balloons,
  // 109.1 -> 108.3 us at 65,536; profiles/r02_step_timing_evict_first.jsonl)
  uint64_t policy;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
               ::"r"(smem_u32(stage)), "l"(src), "r"(bytes), "r"(bar), "l"(policy) : "memory");
}

// NoisyWindHarmonic.get_noise (simplex_wind_noise.py:116-146) for harmonic h10 of balloon ec; the warp's staging
// buffer must hold the tables (lane l's table at stage + 256 l, rotated by 4 l bytes).
struct NoiseOffsets { float x, y, p, t; };
__device__ __forceinline__ NoiseOffsets load_noise_offsets(const DevState<float>& d, int h10, int64_t ec) {
  NoiseOffsets o;
  o.x = d.offsets[(int64_t(h10) * 4 + 0) * d.n + ec]; o.y = d.offsets[(int64_t(h10) * 4 + 1) * d.n + ec];
  o.p = d.offsets[(int64_t(h10) * 4 + 2) * d.n + ec]; o.t = d.offsets[(int64_t(h10) * 4 + 3) * d.n + ec];
  return o;
}
// `tables_done()` runs (in every lane, converged) right after the last table read: the caller releases the staging
// buffer there and starts the next copy.  Lanes beyond the batch evaluate on whatever their slot of the buffer holds
// (the result is discarded) so that the warp stays converged through the callback.
template <typename Done>
__device__ __forceinline__ float noise_harmonic(int h10, const NoiseOffsets& off, int lane, double x, double y,
                                                double p, int32_t t_elapsed, const uint8_t* stage, uint32_t bar,
                                                uint32_t parity, bool valid, Done tables_done) {
  const double* hp = kHarmonicsInvDev[h10];
  const double X = fma(x, hp[0], double(off.x));
  const double Y = fma(y, hp[1], double(off.y));
  const double Z = fma(p, hp[2], double(off.p));
  const double W = fma(double(t_elapsed), hp[3], double(off.t));
  mbar_wait(bar, parity);
  RotatedPerm perm{stage + lane * 256, lane * 4};
  const float v = simplex_noise4_v2<float>(perm, X, Y, Z, W, tables_done);
  return valid ? float(kNoiseMagnitude) * v : 0.f;
}

// The three safety layers, once per agent step (balloon.py:305-313), and X = (p/P_i)^k at the pre-step pressure.
struct SafetyOut {
  int eff;
  uint32_t flags_base;     // flags word without the status bits
  double x0;               // 0 = not available (pressure outside layers 0..2)
};
__device__ __forceinline__ SafetyOut safety_layers(const DevState<float>& d, int64_t ec, int64_t e, bool stepped, uint32_t fl,
                                                   int action, double p_pre, int64_t ts_pre) {
  SafetyOut o;
  const Atmosphere atm = load_atmosphere(d, ec);
  int64_t sunrise_h = d.l[int64_t(L_SUNRISE_H) * d.n + ec], sunset = d.l[int64_t(L_SUNSET) * d.n + ec];
  int env_state = int((fl >> 4) & 7), alt_state = int((fl >> 7) & 3), paused = int((fl >> 9) & 1);
  const int psl = int((fl >> 10) & 1);
  int eff = action;
  if (psl) eff = power_safety<double>(eff, ts_pre, DD(d, D_CHARGE, ec), &sunrise_h, &sunset, &paused);
  eff = envelope_safety<double>(eff, DD(d, D_SP, ec), &env_state);
  roles::PressureRole pr;
  pr.select_layer(atm, p_pre);
  double altitude, x0 = 0.0;
  bool ok = true;
  if (pr.layer >= 0) {
    x0 = pr.x_from_scratch(p_pre);
    altitude = (x0 - 1.0) * pr.tl + atm_h(pr.layer);
  } else {
    Atmosphere a2 = atm;
    double t_unused;
    a2.at_pressure(p_pre, &altitude, &t_unused);
    ok = a2.ok;
  }
  eff = altitude_safety<double>(eff, altitude, &alt_state);
  o.eff = eff;
  o.x0 = x0;
  const uint32_t err = ((fl >> 11) & 1u) | (ok ? 0u : 1u);
  o.flags_base = pack_flags(0, action, env_state, alt_state, paused, psl, int(err));
  if (stepped) {
    d.l[int64_t(L_SUNRISE_H) * d.n + e] = sunrise_h;
    d.l[int64_t(L_SUNSET) * d.n + e] = sunset;
  }
  return o;
}

// info outputs of BalloonEnv.step + the flags word, written by whoever owns the balloon's discrete state
__device__ __forceinline__ void store_discrete(const DevState<float>& d, const FusedOut& out, int64_t e, bool valid, bool stepped,
                                               uint32_t fl, uint32_t flags_base, int status, bool atm_ok, int32_t t_new,
                                               int64_t ts_new, bool last, float uf, float vf) {
  if (stepped) {
    d.l[int64_t(L_DATE_TIME) * d.n + e] = ts_new;                          // balloon.py:546-547
    d.t_elapsed[e] = t_new;
    const uint32_t nf = flags_base | uint32_t(status) | (atm_ok ? 0u : (1u << 11));
    d.flags[e] = nf;
    if (out.sim_error != nullptr) out.sim_error[e] = uint8_t((nf >> 11) & 1u);
  } else if (valid && out.sim_error != nullptr) {
    out.sim_error[e] = uint8_t((fl >> 11) & 1u);
  }
  if (valid) {
    if (out.wind_uv != nullptr && last) out.wind_uv[e] = stepped ? make_float2(uf, vf) : make_float2(0.f, 0.f);
    if (out.status != nullptr) out.status[e] = uint8_t(stepped ? status : int(fl & 3u));
    if (out.time_elapsed != nullptr) out.time_elapsed[e] = t_new;
  }
}

// =====================================================================================================================
// Latency shape
// =====================================================================================================================
template <int kW>
struct FusedSmem {
  // phase 1 results
  float noise[10][32];
  float fu[32], fv[32];
  double phase[2][32];                 // SolarTimeFast at t0 and t0 + 180 s
  float sdecl[2][32], cdecl[2][32], sflux[2][32];
  float cz[3][32];
  int eff[32];
  uint32_t flags_base[32];            // flags word without the status bits, after the safety layers
  double x0[32];                      // (p/P_i)^k at the pre-step pressure (0 = not available)
  // phase 2 exchange, double-buffered by sub-step parity
  double p[2][32], tamb[2][32], vol[2][32], sp[2][32], mols[2][32], charge[2][32];
  float dtb[2][32], dts[2][32], cv[2][32];
  int st_env[2][32], st_pwr[2][32];
  alignas(8) uint64_t bar[kW];
};

// What every role needs to know about the step it is in.
struct RoleCtx {
  int64_t e, ec;
  int lane;
  bool valid, stepped, last;
  int action, eff;
  double x_pre, y_pre, p_pre, u, v;
  float uf, vf;
  int32_t t_pre;
  int64_t ts_pre;
  uint32_t fl;
  int64_t o;               // index of this (step, balloon) in reward / done
};

// status of sub-step b as both E (envelope) and S (power) saw it; later assignment wins (balloon.py:541-542)
template <typename Smem>
__device__ __forceinline__ int substep_status(const Smem& sm, int b, int lane) {
  return sm.st_pwr[b][lane] ? int(kOutOfPower) : sm.st_env[b][lane];
}

template <typename Smem>
__device__ __forceinline__ void run_role_pressure(const DevState<float>& d, const FusedOut& out, Smem& sm, const RoleCtx& c) {
  const int lane = c.lane;
  roles::PressureRole pr;
  pr.init(load_atmosphere(d, c.ec), double(RR(d, R_MOLS_GAS, c.ec)), c.p_pre);
  {
    const double x0 = sm.x0[lane];
    if (pr.layer >= 0 && x0 > 0.0) pr.seed_x(c.p_pre, x0);
  }
  double p = c.p_pre, t_amb = DD(d, D_TAMB, c.ec), vol = DD(d, D_VOL, c.ec), mols = DD(d, D_MOLS_AIR, c.ec);
  float cv = fm::cbrtf_pos(float(vol));
  bool live = c.stepped;
  int n_done = 0, status = kOk;
#pragma unroll 1
  for (int k = 0; k < kSubSteps; ++k) {
    const int b = k & 1;
    if (live) {
      double np, nt;
      pr.step(p, t_amb, vol, mols, cv, [&d, &c]() { return load_atmosphere(d, c.ec); }, &np, &nt);
      sm.p[b][lane] = np; sm.tamb[b][lane] = nt;
      p = np; t_amb = nt;
    }
    role_barrier();
    if (live) {
      vol = sm.vol[b][lane]; mols = sm.mols[b][lane]; cv = sm.cv[b][lane];
      ++n_done;
      status = substep_status(sm, b, lane);
      if (status != kOk) live = false;                                     // break (balloon.py:327-328)
    }
  }
  if (c.stepped) {
    DD(d, D_X, c.e) = c.x_pre + c.u * double(kStrideS) * double(n_done);   // balloon.py:394-395
    DD(d, D_Y, c.e) = c.y_pre + c.v * double(kStrideS) * double(n_done);
    DD(d, D_P, c.e) = p; DD(d, D_TAMB, c.e) = t_amb;
  }
  store_discrete(d, out, c.e, c.valid, c.stepped, c.fl, sm.flags_base[lane], status, pr.atm_ok,
                 c.t_pre + kStrideS * n_done, c.ts_pre + int64_t(kStrideS) * n_done, c.last, c.uf, c.vf);
}

template <typename Smem>
__device__ __forceinline__ void run_role_thermal(const DevState<float>& d, Smem& sm, const RoleCtx& c) {
  const int lane = c.lane;
  const float earth_per_area = earth_heat_per_area<float>(RR(d, R_IR, c.ec));
  double p = c.p_pre, t_amb = DD(d, D_TAMB, c.ec), t_int = DD(d, D_TINT, c.ec);
  float cv = fm::cbrtf_pos(float(DD(d, D_VOL, c.ec)));
  bool live = c.stepped;
#pragma unroll 1
  for (int k = 0; k < kSubSteps; ++k) {
    const int b = k & 1;
    float dtb = 0.f;
    if (live) {
      dtb = roles::thermal_body(cv, t_int, t_amb, p, earth_per_area);
      sm.dtb[b][lane] = dtb;
    }
    role_barrier();
    if (live) {
      p = sm.p[b][lane]; t_amb = sm.tamb[b][lane]; cv = sm.cv[b][lane];
      t_int = t_int + double(dtb + sm.dts[b][lane]) * double(kStrideS);    // balloon.py:462-467
      if (substep_status(sm, b, lane) != kOk) live = false;
    }
  }
  if (c.stepped) DD(d, D_TINT, c.e) = t_int;
}

template <typename Smem>
__device__ __forceinline__ void run_role_envelope(const DevState<float>& d, Smem& sm, const RoleCtx& c) {
  const int lane = c.lane;
  const double mols_gas = double(RR(d, R_MOLS_GAS, c.ec));
  double p = c.p_pre, t_int = DD(d, D_TINT, c.ec), sp = DD(d, D_SP, c.ec), mols = DD(d, D_MOLS_AIR, c.ec);
  double vol = DD(d, D_VOL, c.ec);
  float acs_power = RR(d, R_ACS_W, c.ec), acs_flow = RR(d, R_ACS_FLOW, c.ec);
  bool live = c.stepped;
#pragma unroll 1
  for (int k = 0; k < kSubSteps; ++k) {
    const int b = k & 1;
    if (live) {
      const roles::EnvelopeOut eo = roles::envelope_acs(mols_gas, mols, t_int, p, sp, c.eff);
      sm.vol[b][lane] = eo.volume; sm.sp[b][lane] = eo.superpressure; sm.mols[b][lane] = eo.mols_air;
      sm.cv[b][lane] = eo.cv; sm.st_env[b][lane] = eo.status;
      acs_power = eo.acs_power; acs_flow = eo.flow; vol = eo.volume; sp = eo.superpressure; mols = eo.mols_air;
    }
    role_barrier();
    if (live) {
      p = sm.p[b][lane];
      t_int = t_int + double(sm.dtb[b][lane] + sm.dts[b][lane]) * double(kStrideS);
      if (substep_status(sm, b, lane) != kOk) live = false;
    }
  }
  if (c.stepped) {
    DD(d, D_VOL, c.e) = vol; DD(d, D_SP, c.e) = sp; DD(d, D_MOLS_AIR, c.e) = mols;
    RR(d, R_ACS_W, c.e) = acs_power; RR(d, R_ACS_FLOW, c.e) = acs_flow;
  }
}

template <typename Smem>
__device__ __forceinline__ void run_role_sun(const DevState<float>& d, const FusedOut& out, Smem& sm, const RoleCtx& c) {
  const int lane = c.lane;
  SunTrack<float> sun;
  sun.c0 = sm.cz[0][lane]; sun.c1 = sm.cz[1][lane]; sun.c2 = sm.cz[2][lane];
  sun.f0 = sm.sflux[0][lane]; sun.f2 = sm.sflux[1][lane];
  double p = c.p_pre, sp = DD(d, D_SP, c.ec), charge = DD(d, D_CHARGE, c.ec);
  float cv = fm::cbrtf_pos(float(DD(d, D_VOL, c.ec)));
  float solar_w = RR(d, R_SOLAR_W, c.ec), load_w = RR(d, R_LOAD_W, c.ec), acs_power = RR(d, R_ACS_W, c.ec);
  bool live = c.stepped;
  int n_done = 0, status = kOk;
#pragma unroll 1
  for (int k = 0; k < kSubSteps; ++k) {
    const int b = k & 1;
    if (live) {
      float cz, flux;
      roles::sun_track_at(sun, k, &cz, &flux);
      const roles::SunOut so = roles::sun_power(roles::sun_angles_fast(cz), flux, cv, p, sp, charge, c.eff);
      sm.dts[b][lane] = so.d_t_solar; sm.charge[b][lane] = so.charge; sm.st_pwr[b][lane] = so.out_of_power;
      solar_w = so.solar_w; load_w = so.load_w; acs_power = so.acs_power; charge = so.charge;
    }
    role_barrier();
    if (live) {
      p = sm.p[b][lane]; sp = sm.sp[b][lane]; cv = sm.cv[b][lane];
      ++n_done;
      status = substep_status(sm, b, lane);
      if (status != kOk) live = false;
    }
  }
  if (c.stepped) {
    DD(d, D_CHARGE, c.e) = charge;
    RR(d, R_SOLAR_W, c.e) = solar_w; RR(d, R_LOAD_W, c.e) = load_w;
    // reward on the post-step state (env/balloon_env.py:44-102)
    BalloonState<float> s;
    s.x = c.x_pre + c.u * double(kStrideS) * double(n_done);
    s.y = c.y_pre + c.v * double(kStrideS) * double(n_done);
    s.pressure = p; s.charge = charge; s.acs_power = acs_power;
    float el = 0.f;
    if (c.action == kDown) {                               // excess_energy's sun (balloon.py:231-238)
      float cz, flux;
      roles::sun_track_at(sun, n_done, &cz, &flux);
      el = roles::sun_angles_fast(cz).el;
    }
    out.reward[c.o] = perciatelli_reward<float>(s, c.action, el);
    out.done[c.o] = (status != kOk) ? 1 : 0;
  } else if (c.valid) {                                    // finished balloon: no-op (documented divergence)
    out.reward[c.o] = 0.f;
    out.done[c.o] = 1;
  }
}

// noise_mode: 0 = no noise, 1 = evaluate the harmonics here, 2 = read d.noise_partial (k_noise ran ahead of time:
// the observation path needs the same values for WindGP.observe, and ble_step_host queues them behind its copies)
template <int kW>
__global__ void __launch_bounds__(32 * kW, (kW <= 4 ? 4 : (kW <= 8 ? 2 : 1)))
k_step_roles(DevState<float> d, const int32_t* __restrict__ actions, FusedOut out, int noise_mode, int n_steps) {
  extern __shared__ __align__(128) uint8_t fused_dyn[];
  FusedSmem<kW>& sm = *reinterpret_cast<FusedSmem<kW>*>(fused_dyn);
  uint8_t* stage = fused_dyn + ((sizeof(FusedSmem<kW>) + 127) & ~size_t(127)) + (threadIdx.x >> 5) * kPermStageBytes;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t e0 = int64_t(blockIdx.x) * 32;
  RoleCtx c;
  c.lane = lane;
  c.e = e0 + lane;
  c.valid = c.e < d.n;
  c.ec = c.valid ? c.e : d.n - 1;                           // clamp: every thread stays for the barriers
  const int count = int(min(int64_t(32), d.n - e0));
  const uint32_t bar = smem_u32(&sm.bar[warp]);
  uint32_t bar_parity = 0;
  if (lane == 0) mbar_init(bar);
  __syncwarp();
  asm volatile("griddepcontrol.wait;" ::: "memory");        // the previous step (launch_step: PDL) has completed and flushed
  asm volatile("griddepcontrol.launch_dependents;");        // the next step's grid may be placed from now on (it parks at its wait)

  for (int step = 0; step < n_steps; ++step) {
    c.fl = d.flags[c.ec];
    c.stepped = c.valid && (c.fl & 3u) == uint32_t(kOk);
    c.x_pre = DD(d, D_X, c.ec); c.y_pre = DD(d, D_Y, c.ec); c.p_pre = DD(d, D_P, c.ec);
    c.t_pre = d.t_elapsed[c.ec];
    c.ts_pre = d.l[int64_t(L_DATE_TIME) * d.n + c.ec];
    c.last = step == n_steps - 1;
    c.o = int64_t(step) * d.n + c.e;
    int action = actions[int64_t(step) * d.n + c.ec];
    c.action = action < 0 ? 0 : (action > 2 ? 2 : action);

    // ------------------------------------------------------------------ phase 1a: tasks
#pragma unroll 1
    for (int task = warp; task < kFusedTasks; task += kW) {
      if (task < 10) {
        if (noise_mode == 1) {
          if (task == warp && lane == 0) stage_perm_tables(d, task, e0, count, stage, bar);   // this warp's first harmonic
          sm.noise[task][lane] = noise_harmonic(
              task, load_noise_offsets(d, task, c.ec), lane, c.x_pre, c.y_pre, c.p_pre, c.t_pre, stage, bar, bar_parity, c.valid,
              [&]() {                                       // tables read: the warp's next harmonic may overwrite them
                __syncwarp();
                if (task + kW < 10 && lane == 0) stage_perm_tables(d, task + kW, e0, count, stage, bar);
              });
          bar_parity ^= 1u;
        } else {
          sm.noise[task][lane] = noise_mode == 2 ? d.noise_partial[int64_t(task) * d.n + c.ec] : 0.f;
        }
      } else if (task < 12) {
        const int j = task - 10;
        const roles::SolarTimeFast st = roles::solar_time_fast(c.ts_pre + int64_t(180 * j));
        sm.phase[j][lane] = st.phase;
        sm.sdecl[j][lane] = st.sin_decl; sm.cdecl[j][lane] = st.cos_decl; sm.sflux[j][lane] = st.flux;
      } else {
        // forecast at the PRE-step state (GridBasedWindField.get_forecast, grid_based_wind_field.py:70-94)
        float fu, fv;
        forecast_at<float, DevState<float>>(d, c.ec, c.x_pre, c.y_pre, c.p_pre, c.t_pre, &fu, &fv);
        sm.fu[lane] = fu; sm.fv[lane] = fv;
        const SafetyOut so = safety_layers(d, c.ec, c.e, c.stepped, c.fl, c.action, c.p_pre, c.ts_pre);
        sm.eff[lane] = so.eff; sm.x0[lane] = so.x0; sm.flags_base[lane] = so.flags_base;
      }
    }
    __syncthreads();

    // ------------------------------------------------------------------ phase 1b: wind, sun track
    c.uf = sm.fu[lane]; c.vf = sm.fv[lane];
    if (noise_mode != 0) {                                  // NoisyWindComponent.get_noise (:180-211)
      float nu = 0.f, nv = 0.f;
#pragma unroll
      for (int h = 0; h < 5; ++h) {
        nu += sm.noise[h][lane] * kBlendU[h];
        nv += sm.noise[5 + h][lane] * kBlendV[h];
      }
      c.uf += nu * kBlendScaleU;
      c.vf += nv * kBlendScaleV;
    }
    c.u = double(c.uf); c.v = double(c.vf);
    if (warp < 3) {                                         // one point of the quadratic sun track each (SunTrack)
      const int j = warp;
      const roles::StationTrig stn = roles::station_trig(RR(d, R_LAT0, c.ec), RR(d, R_LNG0, c.ec));
      const double pa = sm.phase[0][lane];
      double pb = sm.phase[1][lane];
      if (pb < pa - kPi) pb += 2.0 * kPi;                  // fraction_of_day wrapped at midnight
      const double w1 = 0.5 * j, w0 = 1.0 - w1;            // the time-only terms move linearly within the step
      const float sd = float(w0) * sm.sdecl[0][lane] + float(w1) * sm.sdecl[1][lane];
      const float cd = float(w0) * sm.cdecl[0][lane] + float(w1) * sm.cdecl[1][lane];
      const double dt_s = 90.0 * j;
      sm.cz[j][lane] = roles::cos_zenith_fast(stn, w0 * pa + w1 * pb, sd, cd, float(c.x_pre + c.u * dt_s),
                                              float(c.y_pre + c.v * dt_s));
    }
    if (warp >= 4) {
      if (n_steps == 1) return;
      __syncthreads();                                      // end-of-step barrier of the role warps
      continue;
    }
    role_barrier();
    c.eff = sm.eff[lane];

    // ------------------------------------------------------------------ phase 2: 18 sub-steps, one role per warp
    if (warp == 0) run_role_pressure(d, out, sm, c);
    else if (warp == 1) run_role_thermal(d, sm, c);
    else if (warp == 2) run_role_envelope(d, sm, c);
    else run_role_sun(d, out, sm, c);
    if (n_steps > 1) __syncthreads();
  }
}

// =====================================================================================================================
// Throughput shape
// =====================================================================================================================
// Warps per CTA.  The warps of a CTA share nothing, so the CTA is only the unit in which the hardware deals work to the
// SMs: with one warp per CTA the 2,048 warps of a 65,536-balloon batch spread as 13 or 14 per SM instead of 12 or 16
// (measured 117 vs 120 us per step at 65,536 and 88 vs 92 us at 32,768, gpurun_out r02p -> profiles/r02_step_timing_cta.jsonl).
#ifndef BLE_WARPS_PER_CTA
#define BLE_WARPS_PER_CTA 1
#endif
constexpr int kWarpsPerCta = BLE_WARPS_PER_CTA;

#ifndef BLE_WARP_MIN_BLOCKS
#define BLE_WARP_MIN_BLOCKS (16 / BLE_WARPS_PER_CTA)
#endif
#ifdef BLE_WARP_MAXNREG
__global__ void __maxnreg__(BLE_WARP_MAXNREG)
#else
__global__ void __launch_bounds__(32 * kWarpsPerCta, BLE_WARP_MIN_BLOCKS)
#endif
k_step_warp(DevState<float> d, const int32_t* __restrict__ actions, FusedOut out, int noise_mode, int n_steps) {
  extern __shared__ __align__(128) uint8_t warp_dyn[];
  __shared__ alignas(8) uint64_t s_bar[kWarpsPerCta];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* stage = warp_dyn + warp * kPermStageBytes;
  const int64_t e0 = (int64_t(blockIdx.x) * kWarpsPerCta + warp) * 32;
  if (e0 >= d.n) return;                                    // whole warp out of range (no CTA-wide barrier anywhere)
  const int64_t e = e0 + lane;
  const bool valid = e < d.n;
  const int64_t ec = valid ? e : d.n - 1;
  const int count = int(min(int64_t(32), d.n - e0));
  const uint32_t bar = smem_u32(&s_bar[warp]);
  uint32_t bar_parity = 0;
  if (lane == 0) mbar_init(bar);
  __syncwarp();

  asm volatile("griddepcontrol.wait;" ::: "memory");        // the previous step (launch_step: PDL) has completed and flushed
  asm volatile("griddepcontrol.launch_dependents;");        // the next step's grid may be placed from now on (it parks at its wait)
  // the tables of harmonic 0 are requested before anything else; inside a rollout the request for the NEXT step goes out
  // as soon as harmonic 9 has read its tables, so that the copy runs under the 18 sub-steps
  if (noise_mode == 1 && lane == 0) stage_perm_tables(d, 0, e0, count, stage, bar);
  for (int step = 0; step < n_steps; ++step) {
    const uint32_t fl = d.flags[ec];
    const bool stepped = valid && (fl & 3u) == uint32_t(kOk);
    const double x_pre = DD(d, D_X, ec), y_pre = DD(d, D_Y, ec), p_pre = DD(d, D_P, ec);
    const int32_t t_pre = d.t_elapsed[ec];
    const int64_t ts_pre = d.l[int64_t(L_DATE_TIME) * d.n + ec];
    const int64_t o = int64_t(step) * d.n + e;
    int action = actions[int64_t(step) * d.n + ec];
    action = action < 0 ? 0 : (action > 2 ? 2 : action);

    // ---- wind at the PRE-step state: forecast window + 10 noise harmonics ----
    float uf, vf;
    forecast_at<float, DevState<float>>(d, ec, x_pre, y_pre, p_pre, t_pre, &uf, &vf);
    if (noise_mode != 0) {
      float nu = 0.f, nv = 0.f;
      NoiseOffsets off = load_noise_offsets(d, 0, ec);
#pragma unroll 1
      for (int h = 0; h < 10; ++h) {
        float nh;
        if (noise_mode == 1) {
          const NoiseOffsets cur = off;
          if (h + 1 < 10) off = load_noise_offsets(d, h + 1, ec);      // in flight while this harmonic is evaluated
          nh = noise_harmonic(h, cur, lane, x_pre, y_pre, p_pre, t_pre, stage, bar, bar_parity, valid, [&]() {
            __syncwarp();                                   // every lane has read its table: the next copy may overwrite it
            const bool more = h + 1 < 10 || step + 1 < n_steps;
            if (more && lane == 0) stage_perm_tables(d, h + 1 < 10 ? h + 1 : 0, e0, count, stage, bar);
          });
          bar_parity ^= 1u;
        } else {
          nh = d.noise_partial[int64_t(h) * d.n + ec];
        }
        if (h < 5) nu += nh * kBlendU[h]; else nv += nh * kBlendV[h - 5];
      }
      uf += nu * kBlendScaleU;
      vf += nv * kBlendScaleV;
    }

    // ---- safety layers, sun track ----
    const SafetyOut so = safety_layers(d, ec, e, stepped, fl, action, p_pre, ts_pre);
    const SunTrack<float> sun = roles::sun_track_fast(roles::solar_time_fast(ts_pre), roles::solar_time_fast(ts_pre + 180),
                                                      RR(d, R_LAT0, ec), RR(d, R_LNG0, ec), x_pre, y_pre, double(uf), double(vf));

    // ---- 18 sub-steps, the four roles back to back ----
    roles::PressureRole pr;
    pr.init(load_atmosphere(d, ec), double(RR(d, R_MOLS_GAS, ec)), p_pre);
    if (pr.layer >= 0 && so.x0 > 0.0) pr.seed_x(p_pre, so.x0);
    const double mols_gas = double(RR(d, R_MOLS_GAS, ec));
    const float earth_per_area = earth_heat_per_area<float>(RR(d, R_IR, ec));
    double p = p_pre, t_amb = DD(d, D_TAMB, ec), t_int = DD(d, D_TINT, ec), vol = DD(d, D_VOL, ec);
    double sp = DD(d, D_SP, ec), mols = DD(d, D_MOLS_AIR, ec), charge = DD(d, D_CHARGE, ec);
    float cv = fm::cbrtf_pos(float(vol));
    float acs_power = RR(d, R_ACS_W, ec), acs_flow = RR(d, R_ACS_FLOW, ec);
    float solar_w = RR(d, R_SOLAR_W, ec), load_w = RR(d, R_LOAD_W, ec);
    int n_done = 0, status = kOk;
    if (stepped) {
#pragma unroll 1
      for (int k = 0; k < kSubSteps; ++k) {
        float cz, flux;
        roles::sun_track_at(sun, k, &cz, &flux);
        double np, nt;
        pr.step(p, t_amb, vol, mols, cv, [&d, ec]() { return load_atmosphere(d, ec); }, &np, &nt);
        const float dtb = roles::thermal_body(cv, t_int, t_amb, p, earth_per_area);
        const roles::EnvelopeOut eo = roles::envelope_acs(mols_gas, mols, t_int, p, sp, so.eff);
        const roles::SunOut po = roles::sun_power(roles::sun_angles_fast(cz), flux, cv, p, sp, charge, so.eff);
        p = np; t_amb = nt;
        t_int = t_int + double(dtb + po.d_t_solar) * double(kStrideS);     // balloon.py:462-467
        vol = eo.volume; sp = eo.superpressure; mols = eo.mols_air; cv = eo.cv;
        charge = po.charge;
        acs_power = eo.acs_power; acs_flow = eo.flow; solar_w = po.solar_w; load_w = po.load_w;
        ++n_done;
        status = po.out_of_power ? int(kOutOfPower) : eo.status;           // later assignment wins (:541-542)
        if (status != kOk) break;                                          // :327-328
      }
    }

    // ---- epilogue ----
    // (x, y, time are re-read here rather than kept in registers across the sub-step loop: they are L1 hits)
    const int32_t t_old = d.t_elapsed[ec];
    const int64_t ts_old = d.l[int64_t(L_DATE_TIME) * d.n + ec];
    if (stepped) {
      const double travelled = double(kStrideS) * double(n_done);
      const double x_new = DD(d, D_X, e) + double(uf) * travelled;         // balloon.py:394-395
      const double y_new = DD(d, D_Y, e) + double(vf) * travelled;
      DD(d, D_X, e) = x_new; DD(d, D_Y, e) = y_new; DD(d, D_P, e) = p;
      DD(d, D_TAMB, e) = t_amb; DD(d, D_TINT, e) = t_int; DD(d, D_VOL, e) = vol;
      DD(d, D_SP, e) = sp; DD(d, D_MOLS_AIR, e) = mols; DD(d, D_CHARGE, e) = charge;
      RR(d, R_ACS_W, e) = acs_power; RR(d, R_ACS_FLOW, e) = acs_flow;
      RR(d, R_SOLAR_W, e) = solar_w; RR(d, R_LOAD_W, e) = load_w;
      BalloonState<float> s;                                               // reward (env/balloon_env.py:44-102)
      s.x = x_new; s.y = y_new; s.pressure = p; s.charge = charge; s.acs_power = acs_power;
      float el = 0.f;
      if (action == kDown) {                                               // excess_energy's sun (balloon.py:231-238)
        float cz, flux;
        roles::sun_track_at(sun, n_done, &cz, &flux);
        el = roles::sun_angles_fast(cz).el;
      }
      out.reward[o] = perciatelli_reward<float>(s, action, el);
      out.done[o] = (status != kOk) ? 1 : 0;
    } else if (valid) {                                                    // finished balloon: no-op
      out.reward[o] = 0.f;
      out.done[o] = 1;
    }
    store_discrete(d, out, e, valid, stepped, fl, so.flags_base, status, pr.atm_ok, t_old + kStrideS * n_done,
                   ts_old + int64_t(kStrideS) * n_done, step == n_steps - 1, uf, vf);
    if (n_steps > 1) __syncwarp();
  }
}

// =====================================================================================================================
// Launcher
// =====================================================================================================================
template <int kW> static size_t roles_smem() { return ((sizeof(FusedSmem<kW>) + 127) & ~size_t(127)) + size_t(kW) * kPermStageBytes; }
static size_t warp_smem() { return size_t(kWarpsPerCta) * kPermStageBytes; }

template <typename Kernel> static cudaError_t setup_kernel(Kernel kernel, int threads, size_t smem, int* blocks) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
  if (e != cudaSuccess) return e;
  // Shared-memory carveout: exactly what the resident CTAs need, the rest of the unified storage stays L1.  The kernels'
  // global traffic is streaming, but their register spills (152 B per thread in the sub-step loop) are not: with the
  // whole storage carved out as shared memory the spill set (78 KB per SM) missed the 28 KB that was left of L1 on every
  // reload; sized to fit, the step went from 117 to 111 us at 65,536 balloons (profiles/r02_step_timing_carveout.jsonl).
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, int(cudaSharedmemCarveoutMaxShared));
  if (e != cudaSuccess) return e;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks, kernel, threads, smem);
  if (e != cudaSuccess) return e;
  int carveout = -2;
  if (const char* env = std::getenv("BLE_STEP_CARVEOUT")) carveout = std::atoi(env);   // percent, A/B override
  if (carveout == -2) {
    constexpr size_t kUnified = 228 * 1024, kPerCtaReserve = 1024;
    const size_t need = size_t(*blocks) * (smem + kPerCtaReserve + 128);
    carveout = int(std::min<size_t>(100, (need * 100 + kUnified - 1) / kUnified));
  }
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, carveout);
  if (e != cudaSuccess) return e;
  int fitted = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&fitted, kernel, threads, smem);
  if (e != cudaSuccess) return e;
  if (fitted < *blocks)                                     // the hint cost residency: take the whole storage again
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, int(cudaSharedmemCarveoutMaxShared));
  return e;
}

cudaError_t fused_setup(int blocks_per_sm[4]) {
  cudaError_t e = setup_kernel(k_step_warp, 32 * kWarpsPerCta, warp_smem(), &blocks_per_sm[0]);
  if (e == cudaSuccess) e = setup_kernel(k_step_roles<4>, 32 * 4, roles_smem<4>(), &blocks_per_sm[1]);
  if (e == cudaSuccess) e = setup_kernel(k_step_roles<8>, 32 * 8, roles_smem<8>(), &blocks_per_sm[2]);
  if (e == cudaSuccess) e = setup_kernel(k_step_roles<14>, 32 * 14, roles_smem<14>(), &blocks_per_sm[3]);
  return e;
}

// Programmatic dependent launch: consecutive steps are consecutive launches of the same kernel on one stream.  With the
// attribute set, the CTAs of step t + 1 are placed on an SM as soon as there is room (every CTA of step t signals
// `griddepcontrol.launch_dependents` right after its own wait) and park at `griddepcontrol.wait` -- the first thing a step
// kernel does before it touches global memory -- until step t has completed and flushed: the launch latency and the ramp
// of the next grid run under the previous step.  Measured: 36.5 -> 34.7 us per step at 4,096 balloons, 41.6 -> 40.4 at
// 8,192, 50.2 -> 48.8 at 16,384, unchanged at 65,536 (profiles/r02_step_timing_pdl.jsonl).  Rollouts are launched the
// classic way: they are long, there is nothing to hide, and CTAs parked early unbalanced the SMs (29.2 -> 31.5 us).
// BLE_STEP_PDL=0 launches the classic way (A/B).
template <typename Kernel>
static void launch_step(Kernel kernel, unsigned grid, unsigned block, size_t smem, cudaStream_t s, const DevState<float>& d,
                        const int32_t* actions, const FusedOut& out, int noise_mode, int n_steps) {
  static const bool pdl = [] { const char* e = std::getenv("BLE_STEP_PDL"); return e == nullptr || std::atoi(e) != 0; }();
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute attr{};
  attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr.val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = &attr; cfg.numAttrs = (pdl && n_steps == 1) ? 1 : 0;   // rollouts are long: nothing to hide, and CTAs parked early unbalance the SMs
  cudaLaunchKernelEx(&cfg, kernel, d, actions, out, noise_mode, n_steps);
}

int fused_launch(int shape, const DevState<float>& d, const int32_t* actions, const FusedOut& out, int noise_mode,
                 int n_steps, cudaStream_t s) {
  const unsigned grid = unsigned((d.n + 31) / 32);
  if (shape == 0 && n_steps > 1) {
    // Throughput shape (the batch fills the GPU): K launches of one step beat one launch of K steps -- inside a long
    // launch the warps of an SM drift apart and the 130 KB of SASS stop sharing instruction-cache lines (101 us per
    // step against 134 us at 65,536 balloons, 72 against 86 us at 32,768; profiles/r02b_step_timing.jsonl) -- so a
    // rollout is issued as K programmatic-dependent launches.  Bit-identical by construction.
    static const bool split = [] { const char* e = std::getenv("BLE_ROLLOUT_SPLIT"); return e == nullptr || std::atoi(e) != 0; }();
    if (split) {
      for (int k = 0; k < n_steps; ++k) {
        FusedOut o = out;
        o.reward += int64_t(k) * d.n; o.done += int64_t(k) * d.n;
        if (k + 1 < n_steps) { o.wind_uv = nullptr; o.status = nullptr; o.time_elapsed = nullptr; o.sim_error = nullptr; }
        launch_step(k_step_warp, unsigned((d.n + 32 * kWarpsPerCta - 1) / (32 * kWarpsPerCta)), 32 * kWarpsPerCta, warp_smem(), s,
                    d, actions + int64_t(k) * d.n, o, noise_mode, 1);
      }
      return n_steps;
    }
  }
  switch (shape) {
    case 14: launch_step(k_step_roles<14>, grid, 32 * 14, roles_smem<14>(), s, d, actions, out, noise_mode, n_steps); break;
    case 8: launch_step(k_step_roles<8>, grid, 32 * 8, roles_smem<8>(), s, d, actions, out, noise_mode, n_steps); break;
    case 4: launch_step(k_step_roles<4>, grid, 32 * 4, roles_smem<4>(), s, d, actions, out, noise_mode, n_steps); break;
    default:
      launch_step(k_step_warp, unsigned((d.n + 32 * kWarpsPerCta - 1) / (32 * kWarpsPerCta)), 32 * kWarpsPerCta, warp_smem(), s,
                  d, actions, out, noise_mode, n_steps);
      break;
  }
  return 1;
}

}  // namespace ble
