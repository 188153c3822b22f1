"""Planet / Ellipsoid stand-ins: only what the 'planet' port reads (Topozero.py:560-569, Geo2rdr.py:457-466)."""
from __future__ import annotations

# components/isceobj/Constants + Planet/AstronomicalHandbook.py (WGS-84)
EarthMajorSemiAxis = 6378137.0
EarthEccentricitySquared = 0.0066943799901


class Ellipsoid:
    def __init__(self, a=EarthMajorSemiAxis, e2=EarthEccentricitySquared):
        self.a = a
        self.e2 = e2

    def get_a(self): return self.a
    def get_e2(self): return self.e2


class Planet:
    def __init__(self, pname="Earth"):
        if pname != "Earth":
            raise ValueError("only Earth is tabulated here; pass an object with get_elp() for other bodies")
        self.name = pname
        self.ellipsoid = Ellipsoid()

    def get_elp(self): return self.ellipsoid
