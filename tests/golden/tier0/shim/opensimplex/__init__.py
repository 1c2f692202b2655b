"""Tier-0 stand-in for the un-vendored third-party `opensimplex==0.3` package.

Delegates to the repo's restatement of the published algorithm
(oracle/opensimplex4.py) so that reference trajectories dumped by
make_golden.py and the oracle use the *same* noise.  Noise VALUES are
therefore not pinned against the real package (see oracle/opensimplex4.py).
"""
from oracle.opensimplex4 import OpenSimplex  # noqa: F401
