"""Samples of the reference's reset distributions (this container only) -> tests/golden/reset_samples.npz.

    python -m tests.golden.tier0.make_reset_samples

BalloonArena.reset (env/balloon_arena.py:160-182) is run N times on the UNMODIFIED reference under the Tier-0 stubs;
recorded per reset: the atmosphere's alpha (standard_atmosphere.py:76-87), the start time (utils/sampling.py:53-72),
x / y (Beta(1.2, 2) * 200 km at a uniform angle, balloon_arena.py:243-249), the centre latitude / longitude
(sampling.py:37-50), the pressure (sampling.py:75-97: U(6500, p(MIN_ALTITUDE))) and the upwelling infrared
(sampling.py:100-152).  The GPU test compares the device reset's draws against these with two-sample
Kolmogorov-Smirnov tests (the streams differ -- jax threefry there, Philox here -- the distributions must not).
"""
import os

import tests.golden.tier0.boot as boot  # noqa: F401

import numpy as np
from balloon_learning_environment.env import balloon_arena
from balloon_learning_environment.env import wind_field

from tests.golden.tier0 import make_golden as mg

OUT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N = 3000


def main():
  arena = balloon_arena.BalloonArena(mg._NullFeatures, wind_field.SimpleStaticWindField(), seed=0)
  rows = []
  for k in range(N):
    arena.reset(1000003 * k + 17)
    st = arena.get_balloon_state()
    atm = arena._atmosphere
    lo, hi, lapse = atm._LAPSE_RATES_LOW, atm._LAPSE_RATES_HIGH, atm._lapse_rates
    j = int(np.argmax(np.abs(np.asarray(hi) - np.asarray(lo))))
    alpha = float((lapse[j] - lo[j]) / (hi[j] - lo[j]))
    rows.append([alpha, mg.ts_of(st.date_time), st.x.m, st.y.m, st.center_latlng.lat().degrees,
                 st.center_latlng.lng().degrees, st.pressure, st.upwelling_infrared, st.battery_charge.watt_hours,
                 float(atm.at_height(__import__('balloon_learning_environment.env.balloon.altitude_safety',
                                                fromlist=['x']).MIN_ALTITUDE).pressure)])
  a = np.asarray(rows, np.float64)
  np.savez_compressed(os.path.join(OUT, 'reset_samples.npz'),
                      columns=np.array(['alpha', 'date_time', 'x', 'y', 'lat_deg', 'lng_deg', 'pressure', 'upwelling_infrared',
                                        'battery_charge', 'max_pressure']), samples=a.astype(np.float64))
  print('wrote', a.shape, 'means', a.mean(0))


if __name__ == '__main__':
  main()
