"""Runs the reference's OWN unit tests for the hot path, unmodified, under the Tier-0 stubs.

Fidelity evidence for the stub set: if the stubs distorted the
transition arithmetic, these tests (known-answer tables for solar, atmosphere, ACS,
safety layers, grid interpolation, reward ...) would fail.

    python -m tests.golden.tier0.run_reference_tests
"""
import sys

import tests.golden.tier0.boot as boot  # noqa: F401  (must precede reference imports)
import pytest

_R = boot.REFERENCE_ROOT + '/balloon_learning_environment/'
FILES = [
    'env/balloon/solar_test.py', 'env/balloon/standard_atmosphere_test.py',
    'env/balloon/acs_test.py', 'env/balloon/power_table_test.py',
    'env/balloon/altitude_safety_test.py', 'env/balloon/envelope_safety_test.py',
    'env/balloon/power_safety_test.py', 'env/balloon/balloon_test.py',
    'env/balloon/stable_init_test.py', 'env/balloon/pressure_range_builder_test.py',
    'env/grid_based_wind_field_test.py', 'env/wind_field_test.py', 'env/wind_gp_test.py',
    'utils/spherical_geometry_test.py', 'utils/transforms_test.py', 'utils/units_test.py',
    'env/features_test.py', 'env/balloon_arena_test.py', 'env/balloon_env_test.py',
]

if __name__ == '__main__':
  sys.exit(pytest.main(['-q', '-p', 'no:cacheprovider', '--rootdir', '/tmp',
                        '-W', 'ignore'] + [_R + f for f in FILES] + sys.argv[1:]))
