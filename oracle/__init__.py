"""CPU oracle for the BLE transition function (TEST INFRASTRUCTURE ONLY).

This package is a from-scratch fp64 NumPy restatement of the reference's
`BalloonEnv.step` hot path.  Every function cites the
reference file:line it follows (paths relative to
/root/reference/balloon_learning_environment/).

It is the *checker*: only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import it.  The
product package `balloon_learning_environment_b200` never imports it and has
no CPU fallback.

Parity pinning:
  * pinned  : atmosphere, solar, thermal, ACS, superpressure, safety layers,
              simulate_step trajectories, reward, grid interpolation, stable
              init -- against golden vectors dumped from the reference's own
              code (tests/golden/tier0/make_golden.py) and the known-answer
              tables in the reference's unit tests.
  * UNPINNED: OpenSimplex noise values (third-party `opensimplex==0.3` is not
              vendored in the reference and not installed here;
              `oracle/opensimplex4.py` restates the published algorithm).
"""
