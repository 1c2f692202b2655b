import math
class Angle:
    def __init__(self, r): self.radians = r
    @property
    def degrees(self): return math.degrees(self.radians)
class LatLng:
    def __init__(self, lat_r, lng_r): self._lat, self._lng = lat_r, lng_r
    @classmethod
    def from_degrees(cls, lat, lng): return cls(math.radians(lat), math.radians(lng))
    @classmethod
    def from_radians(cls, lat, lng): return cls(lat, lng)
    def lat(self): return Angle(self._lat)
    def lng(self): return Angle(self._lng)
    @property
    def is_valid(self): return abs(self._lat) <= math.pi / 2 and abs(self._lng) <= math.pi
    def normalized(self):
        return LatLng(max(-math.pi/2, min(math.pi/2, self._lat)), math.remainder(self._lng, 2*math.pi))
    def __eq__(self, o): return self._lat == o._lat and self._lng == o._lng
