// The fused one-launch step kernel of the production (fp32) build and its launcher.
#include "ble_step_fused.h"

#include "ble_step_roles.cuh"

namespace ble {

// k_step_fused: BalloonEnv.step for 32 balloons per CTA in ONE launch (production fp32 build).
//
// Reference path: env/balloon_env.py:157-190 -> env/balloon_arena.py:184-202 -> WindField.get_ground_truth
// (env/wind_field.py:125-145: grid interpolation + 10 simplex-noise harmonics) -> Balloon.simulate_step
// (env/balloon/balloon.py:263-328: three safety layers, 18 Euler sub-steps) -> reward / terminal.
//
// Work decomposition (kW warps per CTA, lane l of every warp = balloon 32 * block + l):
//   phase 1a  14 independent TASKS dealt round-robin to the kW warps:
//               N0..N9  one simplex-noise harmonic each; the 32 balloons' permutation tables of that harmonic
//                       (8 KB, contiguous in HBM) arrive in the warp's staging buffer by ONE TMA bulk copy
//                       (cp.async.bulk + mbarrier) issued before the lattice coordinates are computed;
//               ST0..2  the time-only half of the NOAA solar calculator at t0, t0 + 90 s, t0 + 180 s;
//               G       forecast gather (8 x LDG.128 of the lookup's 128-byte window), power / envelope /
//                       altitude safety layers -> effective action, and X = (p/P_i)^k for role P.
//   barrier   (whole CTA)
//   phase 1b  wind = forecast + blended noise; warps 0..2 finish the sun at the three track points
//             (position-dependent half: great-circle offset + hour angle), which needs the wind.
//   barrier   (4 role warps; the other warps are done)
//   phase 2   18 Euler sub-steps, warp w = role w (P pressure/atmosphere, T thermal body, E envelope + ACS,
//             S sun + power) on the branch-free role functions of ble_step_roles.cuh; one named barrier and
//             one double-buffered shared-memory exchange per sub-step.
//   epilogue  every role stores the rows it owns; role S evaluates the reward.
// kW = 4 is the throughput shape (65,536 balloons: 3.5 waves of 4 CTAs / SM); kW = 14 gives every task its
// own warp and is the latency shape for batches below one wave (8,192 balloons per GPU in the 8-GPU split).
//
// kSteps > 1 (ble_rollout): the same CTA runs kSteps agent steps back to back on actions[step][N]; balloons do
// not interact, so there is no grid-wide synchronisation and no launch gap between steps.

template <int kW>
struct FusedSmem {
  // phase 1 results
  float noise[10][32];
  float fu[32], fv[32];
  double fod[3][32], eqt[3][32];
  float sdecl[3][32], cdecl[3][32], sflux[3][32];
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

constexpr int kFusedTasks = 14;
constexpr int kPermStageBytes = 32 * 256;

__device__ __forceinline__ void role_barrier() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}

// noise_mode: 0 = no noise, 1 = evaluate the harmonics here, 2 = read d.noise_partial (k_noise ran ahead of time:
// the observation path needs the same values for WindGP.observe, and ble_step_host queues them behind its copies)
template <int kW>
__global__ void __launch_bounds__(32 * kW, (kW <= 4 ? 4 : (kW <= 8 ? 2 : 1)))
k_step_fused(DevState<float> d, const int32_t* __restrict__ actions, FusedOut out, int noise_mode, int n_steps) {
  using Real = float;
  extern __shared__ __align__(128) uint8_t fused_dyn[];
  FusedSmem<kW>& sm = *reinterpret_cast<FusedSmem<kW>*>(fused_dyn);
  uint8_t* stage = fused_dyn + ((sizeof(FusedSmem<kW>) + 127) & ~size_t(127));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t e0 = int64_t(blockIdx.x) * 32;
  const int64_t e = e0 + lane;
  const bool valid = e < d.n;
  const int64_t ec = valid ? e : d.n - 1;                   // clamp: every thread stays for the barriers
  const int count = int(min(int64_t(32), d.n - e0));
  const uint32_t bar = smem_u32(&sm.bar[warp]);
  uint32_t bar_parity = 0;
  if (lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();

  for (int step = 0; step < n_steps; ++step) {
    const int32_t* act_row = actions + int64_t(step) * d.n;
    const uint32_t fl = d.flags[ec];
    const bool stepped = valid && (fl & 3u) == uint32_t(kOk);
    const double x_pre = DD(d, D_X, ec), y_pre = DD(d, D_Y, ec), p_pre = DD(d, D_P, ec);
    const int32_t t_pre = d.t_elapsed[ec];
    const int64_t ts_pre = d.l[int64_t(L_DATE_TIME) * d.n + ec];
    int action = act_row[ec];
    action = action < 0 ? 0 : (action > 2 ? 2 : action);

    // ------------------------------------------------------------------ phase 1a: tasks
    for (int task = warp; task < kFusedTasks; task += kW) {
      if (task < 10) {
        if (noise_mode == 1) {
          const int h10 = task;
          if (lane == 0) {
            const uint32_t bytes = uint32_t(count) * 256u;
            const uint8_t* src = d.perm + (int64_t(h10) * d.n + e0) * 256;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(stage + warp * kPermStageBytes)), "l"(src), "r"(bytes), "r"(bar) : "memory");
          }
          // lattice coordinates while the copy is in flight (NoisyWindHarmonic.get_noise, simplex_wind_noise.py:116-146)
          const double* hp = kHarmonicsInvDev[h10];
          const double X = fma(x_pre, hp[0], double(d.offsets[(int64_t(h10) * 4 + 0) * d.n + ec]));
          const double Y = fma(y_pre, hp[1], double(d.offsets[(int64_t(h10) * 4 + 1) * d.n + ec]));
          const double Z = fma(p_pre, hp[2], double(d.offsets[(int64_t(h10) * 4 + 2) * d.n + ec]));
          const double Wc = fma(double(t_pre), hp[3], double(d.offsets[(int64_t(h10) * 4 + 3) * d.n + ec]));
          mbar_wait(bar, bar_parity);
          bar_parity ^= 1u;
          float v = 0.f;
          if (valid) {
            RotatedPerm perm{stage + warp * kPermStageBytes + lane * 256, lane * 4};
            v = simplex_noise4<Real>(perm, X, Y, Z, Wc);
          }
          sm.noise[h10][lane] = float(kNoiseMagnitude) * v;
          __syncwarp();                                     // every lane is done with the staging buffer
        } else {
          sm.noise[task][lane] = noise_mode == 2 ? d.noise_partial[int64_t(task) * d.n + ec] : 0.f;
        }
      } else if (task < 13) {
        const int j = task - 10;
        const SolarTime<Real> st = solar_time<Real>(ts_pre + int64_t(90 * j));
        sm.fod[j][lane] = st.fod; sm.eqt[j][lane] = st.eq_time_deg;
        sm.sdecl[j][lane] = st.sin_decl; sm.cdecl[j][lane] = st.cos_decl; sm.sflux[j][lane] = st.flux;
      } else {
        // forecast at the PRE-step state (GridBasedWindField.get_forecast, grid_based_wind_field.py:70-94)
        Real fu, fv;
        forecast_at<Real, DevState<Real>>(d, ec, x_pre, y_pre, p_pre, t_pre, &fu, &fv);
        sm.fu[lane] = fu; sm.fv[lane] = fv;
        // safety layers, once per agent step (balloon.py:305-313)
        Atmosphere atm = load_atmosphere(d, ec);
        int64_t sunrise_h = d.l[int64_t(L_SUNRISE_H) * d.n + ec], sunset = d.l[int64_t(L_SUNSET) * d.n + ec];
        int env_state = int((fl >> 4) & 7), alt_state = int((fl >> 7) & 3), paused = int((fl >> 9) & 1);
        const int psl = int((fl >> 10) & 1);
        int eff = action;
        if (psl) eff = power_safety<double>(eff, ts_pre, DD(d, D_CHARGE, ec), &sunrise_h, &sunset, &paused);
        eff = envelope_safety<double>(eff, DD(d, D_SP, ec), &env_state);
        roles::PressureRole pr;
        pr.atm = atm;
        pr.select_layer(p_pre);
        double altitude, x0 = 0.0;
        if (pr.layer >= 0) {
          x0 = pr.x_from_scratch(p_pre);
          altitude = (x0 - 1.0) * pr.tl + atm_h(pr.layer);
        } else {
          double t_unused;
          atm.at_pressure(p_pre, &altitude, &t_unused);
        }
        eff = altitude_safety<double>(eff, altitude, &alt_state);
        sm.eff[lane] = eff;
        sm.x0[lane] = x0;
        const uint32_t err = ((fl >> 11) & 1u) | (atm.ok ? 0u : 1u);
        sm.flags_base[lane] = pack_flags(0, action, env_state, alt_state, paused, psl, int(err));
        if (stepped) {
          d.l[int64_t(L_SUNRISE_H) * d.n + e] = sunrise_h;
          d.l[int64_t(L_SUNSET) * d.n + e] = sunset;
        }
      }
    }
    __syncthreads();

    // ------------------------------------------------------------------ phase 1b: wind, sun track
    float uf = sm.fu[lane], vf = sm.fv[lane];
    if (noise_mode != 0) {                                  // NoisyWindComponent.get_noise (:180-211)
      float nu = 0.f, nv = 0.f;
#pragma unroll
      for (int h = 0; h < 5; ++h) {
        nu += sm.noise[h][lane] * kBlendU[h];
        nv += sm.noise[5 + h][lane] * kBlendV[h];
      }
      uf += nu * kBlendScaleU;
      vf += nv * kBlendScaleV;
    }
    const double u = double(uf), v = double(vf);
    if (warp < 3) {
      const int j = warp;
      const double dt_s = 90.0 * j;
      Real lat, lng;
      latlng_from_offset<Real>(RR(d, R_LAT0, ec), RR(d, R_LNG0, ec), Real(x_pre + u * dt_s), Real(y_pre + v * dt_s), &lat, &lng);
      SolarTime<Real> st;
      st.fod = sm.fod[j][lane]; st.eq_time_deg = sm.eqt[j][lane];
      st.sin_decl = sm.sdecl[j][lane]; st.cos_decl = sm.cdecl[j][lane]; st.flux = sm.sflux[j][lane];
      sm.cz[j][lane] = solar_cos_zenith<Real>(st, lat, lng);
    }
    if (warp >= 4) {
      if (n_steps == 1) return;
      __syncthreads();                                      // end-of-step barrier of the role warps
      continue;
    }
    role_barrier();

    // ------------------------------------------------------------------ phase 2: 18 sub-steps, one role per warp
    const int eff = sm.eff[lane];
    SunTrack<Real> sun;
    sun.c0 = sm.cz[0][lane]; sun.c1 = sm.cz[1][lane]; sun.c2 = sm.cz[2][lane];
    sun.f0 = sm.sflux[0][lane]; sun.f2 = sm.sflux[2][lane];
    bool live = stepped;
    int n_done = 0, status = kOk;
    // state every role carries
    double p = p_pre;
    double t_int = DD(d, D_TINT, ec);
    float cv = 0.f;
    // role-private state
    roles::PressureRole pr;                                  // P
    double t_amb = 0, vol = 0, mols = 0;                     // P (t_amb also T)
    float earth_per_area = 0.f;                              // T
    double sp = 0, mols_gas = 0;                             // E, S (sp)
    double charge = 0;                                       // S
    float acs_power = 0.f, acs_flow = 0.f, solar_w = 0.f, load_w = 0.f;
    {
      const double vol0 = DD(d, D_VOL, ec);
      cv = fm::cbrtf_pos(float(vol0));
      if (warp == 0) {
        pr.init(load_atmosphere(d, ec), double(RR(d, R_MOLS_GAS, ec)), p_pre);
        const double x0 = sm.x0[lane];
        if (pr.layer >= 0 && x0 > 0.0) pr.seed_x(p_pre, x0);
        t_amb = DD(d, D_TAMB, ec); vol = vol0; mols = DD(d, D_MOLS_AIR, ec);
      } else if (warp == 1) {
        t_amb = DD(d, D_TAMB, ec);
        earth_per_area = earth_heat_per_area<Real>(RR(d, R_IR, ec));
      } else if (warp == 2) {
        sp = DD(d, D_SP, ec); mols = DD(d, D_MOLS_AIR, ec); mols_gas = double(RR(d, R_MOLS_GAS, ec));
        acs_power = RR(d, R_ACS_W, ec); acs_flow = RR(d, R_ACS_FLOW, ec);
      } else {
        sp = DD(d, D_SP, ec); charge = DD(d, D_CHARGE, ec);
        solar_w = RR(d, R_SOLAR_W, ec); load_w = RR(d, R_LOAD_W, ec);
      }
    }
#pragma unroll 1
    for (int k = 0; k < kSubSteps; ++k) {
      const int b = k & 1;
      if (live) {
        if (warp == 0) {
          double np, nt;
          pr.step(p, t_amb, vol, mols, cv, &np, &nt);
          sm.p[b][lane] = np; sm.tamb[b][lane] = nt;
        } else if (warp == 1) {
          sm.dtb[b][lane] = roles::thermal_body(cv, t_int, t_amb, p, earth_per_area);
        } else if (warp == 2) {
          const roles::EnvelopeOut eo = roles::envelope_acs(mols_gas, mols, t_int, p, sp, eff);
          sm.vol[b][lane] = eo.volume; sm.sp[b][lane] = eo.superpressure; sm.mols[b][lane] = eo.mols_air;
          sm.cv[b][lane] = eo.cv; sm.st_env[b][lane] = eo.status;
          acs_power = eo.acs_power; acs_flow = eo.flow; vol = eo.volume;
        } else {
          float cz, flux;
          roles::sun_track_at(sun, k, &cz, &flux);
          const roles::SunOut so = roles::sun_power(roles::sun_angles_fast(cz), flux, cv, p, sp, charge, eff);
          sm.dts[b][lane] = so.d_t_solar; sm.charge[b][lane] = so.charge; sm.st_pwr[b][lane] = so.out_of_power;
          solar_w = so.solar_w; load_w = so.load_w; acs_power = so.acs_power;
        }
      }
      role_barrier();
      if (live) {
        p = sm.p[b][lane];
        t_int = t_int + double(sm.dtb[b][lane] + sm.dts[b][lane]) * double(kStrideS);     // balloon.py:462-467
        cv = sm.cv[b][lane];
        if (warp == 0) { t_amb = sm.tamb[b][lane]; vol = sm.vol[b][lane]; mols = sm.mols[b][lane]; }
        else if (warp == 1) { t_amb = sm.tamb[b][lane]; }
        else if (warp == 2) { sp = sm.sp[b][lane]; mols = sm.mols[b][lane]; }
        else { sp = sm.sp[b][lane]; charge = sm.charge[b][lane]; }
        ++n_done;
        status = sm.st_pwr[b][lane] ? int(kOutOfPower) : sm.st_env[b][lane];     // later assignment wins (:541-542)
        if (status != kOk) live = false;                                         // break (:327-328)
      }
    }

    // ------------------------------------------------------------------ epilogue: each role stores what it owns
    const bool last = step == n_steps - 1;
    const int64_t o = int64_t(step) * d.n + e;
    if (warp == 0) {
      if (stepped) {
        DD(d, D_X, e) = x_pre + u * double(kStrideS) * double(n_done);     // balloon.py:394-395
        DD(d, D_Y, e) = y_pre + v * double(kStrideS) * double(n_done);
        DD(d, D_P, e) = p; DD(d, D_TAMB, e) = t_amb;
        d.l[int64_t(L_DATE_TIME) * d.n + e] = ts_pre + int64_t(kStrideS) * n_done;     // :546-547
        d.t_elapsed[e] = t_pre + kStrideS * n_done;
        const uint32_t nf = sm.flags_base[lane] | uint32_t(status) | (pr.atm.ok ? 0u : (1u << 11));
        d.flags[e] = nf;
        if (out.sim_error != nullptr) out.sim_error[e] = uint8_t((nf >> 11) & 1u);
      } else if (valid && out.sim_error != nullptr) {
        out.sim_error[e] = uint8_t((fl >> 11) & 1u);
      }
      if (valid) {
        if (out.wind_uv != nullptr && last) out.wind_uv[e] = stepped ? make_float2(uf, vf) : make_float2(0.f, 0.f);
        if (out.status != nullptr) out.status[e] = uint8_t(stepped ? status : int(fl & 3u));
        if (out.time_elapsed != nullptr) out.time_elapsed[e] = stepped ? t_pre + kStrideS * n_done : t_pre;
      }
    } else if (warp == 1) {
      if (stepped) DD(d, D_TINT, e) = t_int;
    } else if (warp == 2) {
      if (stepped) {
        DD(d, D_VOL, e) = vol;
        DD(d, D_SP, e) = sp; DD(d, D_MOLS_AIR, e) = mols;
        RR(d, R_ACS_W, e) = acs_power; RR(d, R_ACS_FLOW, e) = acs_flow;
      }
    } else {
      if (stepped) {
        DD(d, D_CHARGE, e) = charge;
        RR(d, R_SOLAR_W, e) = solar_w; RR(d, R_LOAD_W, e) = load_w;
        // reward on the post-step state (env/balloon_env.py:44-102)
        BalloonState<Real> s;
        s.x = x_pre + u * double(kStrideS) * double(n_done);
        s.y = y_pre + v * double(kStrideS) * double(n_done);
        s.pressure = p; s.charge = charge; s.acs_power = acs_power;
        float el = 0.f;
        if (action == kDown) {                             // excess_energy's sun (balloon.py:231-238)
          float cz, flux;
          roles::sun_track_at(sun, n_done, &cz, &flux);
          el = roles::sun_angles_fast(cz).el;
        }
        out.reward[o] = perciatelli_reward<Real>(s, action, el);
        out.done[o] = (status != kOk) ? 1 : 0;
      } else if (valid) {                                  // finished balloon: no-op (documented divergence)
        out.reward[o] = 0.f;
        out.done[o] = 1;
      }
    }
    if (n_steps > 1) __syncthreads();
  }
}

template <int kW> static size_t fused_smem() { return ((sizeof(FusedSmem<kW>) + 127) & ~size_t(127)) + size_t(kW) * kPermStageBytes; }

template <int kW> static cudaError_t setup_one(int* blocks) {
  cudaError_t e = cudaFuncSetAttribute(k_step_fused<kW>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(fused_smem<kW>()));
  if (e != cudaSuccess) return e;
  return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks, k_step_fused<kW>, 32 * kW, fused_smem<kW>());
}

cudaError_t fused_setup(int blocks_per_sm[4]) {
  cudaError_t e = setup_one<4>(&blocks_per_sm[0]);
  if (e == cudaSuccess) e = setup_one<8>(&blocks_per_sm[1]);
  if (e == cudaSuccess) e = setup_one<10>(&blocks_per_sm[2]);
  if (e == cudaSuccess) e = setup_one<14>(&blocks_per_sm[3]);
  return e;
}

void fused_launch(int warps, const DevState<float>& d, const int32_t* actions, const FusedOut& out, int noise_mode,
                  int n_steps, cudaStream_t s) {
  const unsigned grid = unsigned((d.n + 31) / 32);
  switch (warps) {
    case 14: k_step_fused<14><<<grid, 32 * 14, fused_smem<14>(), s>>>(d, actions, out, noise_mode, n_steps); break;
    case 10: k_step_fused<10><<<grid, 32 * 10, fused_smem<10>(), s>>>(d, actions, out, noise_mode, n_steps); break;
    case 8: k_step_fused<8><<<grid, 32 * 8, fused_smem<8>(), s>>>(d, actions, out, noise_mode, n_steps); break;
    default: k_step_fused<4><<<grid, 32 * 4, fused_smem<4>(), s>>>(d, actions, out, noise_mode, n_steps); break;
  }
}

}  // namespace ble
