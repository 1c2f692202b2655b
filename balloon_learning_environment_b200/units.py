"""Unit value types with the attribute names the reference's `units` module exposes.

Only what crosses the arena boundary is provided (utils/units.py:26-373): code written against
`BalloonState.x.km`, `.acs_power.watts`, `.battery_charge.watt_hours`, `WindVector.u.mps` keeps
working on the states returned by CudaBalloonArena.
"""
import dataclasses

_METERS_PER_FOOT = 0.3048


@dataclasses.dataclass(frozen=True)
class Distance:
  m: float = 0.0

  @classmethod
  def of(cls, *, m=0.0, meters=0.0, km=0.0, kilometers=0.0, feet=0.0):
    return cls(m + meters + (km + kilometers) * 1000.0 + feet * _METERS_PER_FOOT)

  meters = property(lambda self: self.m)
  km = property(lambda self: self.m / 1000.0)
  kilometers = property(lambda self: self.m / 1000.0)
  feet = property(lambda self: self.m / _METERS_PER_FOOT)


@dataclasses.dataclass(frozen=True)
class Velocity:
  mps: float = 0.0
  meters_per_second = property(lambda self: self.mps)
  kmph = property(lambda self: self.mps * 3.6)


@dataclasses.dataclass(frozen=True)
class Power:
  watts: float = 0.0


@dataclasses.dataclass(frozen=True)
class Energy:
  watt_hours: float = 0.0
