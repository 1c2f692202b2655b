"""Spherical offset -> lat/lng (ORACLE / test infrastructure).

Follows utils/spherical_geometry.py:44-76 and `s2sphere.LatLng.normalized()`
(lat clamped to +-pi/2, lng = IEEE remainder(lng, 2 pi)).
"""
import numpy as np

from oracle import constants as C


def latlng_from_offset(center_lat, center_lng, x_m, y_m):
  """All radians / metres -> (lat, lng) radians."""
  x_m = np.asarray(x_m, np.float64); y_m = np.asarray(y_m, np.float64)
  heading = np.arctan2(x_m / 1000.0, y_m / 1000.0)                        # :61
  angle = np.sqrt(x_m * x_m + y_m * y_m) / C.EARTH_RADIUS_M               # :62
  cos_a, sin_a = np.cos(angle), np.sin(angle)
  sin_from, cos_from = np.sin(center_lat), np.cos(center_lat)
  sin_lat = cos_a * sin_from + sin_a * cos_from * np.cos(heading)         # :69-70
  d_lng = np.arctan2(sin_a * cos_from * np.sin(heading), cos_a - sin_from * sin_lat)
  new_lat = np.arcsin(sin_lat)
  new_lat = np.minimum(np.maximum(new_lat, -np.pi / 2.0), np.pi / 2.0)
  new_lng = center_lng + d_lng
  new_lng = new_lng - (2 * np.pi) * np.rint(new_lng / (2 * np.pi))        # IEEE remainder
  return new_lat, new_lng
