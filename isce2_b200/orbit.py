"""Host-side orbit helpers: the hand-off of isceobj.Orbit.Orbit.exportToC (components/isceobj/Orbit/Orbit.py:1060-1081)
and getENUHeading (Orbit.py:804-831), for our own light-weight Orbit and for a real ISCE Orbit object alike."""
from __future__ import annotations

import datetime
import math

import numpy as np

from .planet import EarthEccentricitySquared, EarthMajorSemiAxis


class StateVector:
    def __init__(self, time=None, position=None, velocity=None):
        self._time = time
        self._position = list(position) if position is not None else None
        self._velocity = list(velocity) if velocity is not None else None

    def setTime(self, t): self._time = t
    def getTime(self): return self._time
    def setPosition(self, p): self._position = list(p)
    def getPosition(self): return self._position
    def setVelocity(self, v): self._velocity = list(v)
    def getVelocity(self): return self._velocity


class Orbit:
    """Ordered list of ECEF state vectors (datetime, m, m/s)."""

    def __init__(self):
        self._stateVectors = []

    def addStateVector(self, sv):
        self._stateVectors.append(sv)
        self._stateVectors.sort(key=lambda s: s.getTime())

    def __iter__(self):
        return iter(self._stateVectors)

    def __len__(self):
        return len(self._stateVectors)

    @property
    def minTime(self): return self._stateVectors[0].getTime()
    @property
    def maxTime(self): return self._stateVectors[-1].getTime()

    @classmethod
    def from_arrays(cls, day, t, pos, vel):
        """day: datetime at 00:00 of the reference day; t seconds of day."""
        o = cls()
        for ti, p, v in zip(t, pos, vel):
            o._stateVectors.append(StateVector(day + datetime.timedelta(seconds=float(ti)), p, v))
        return o

    def getENUHeading(self, time=None, planet=None):
        return enu_heading_deg(self, time)


def state_vectors(orbit):
    if hasattr(orbit, "_stateVectors"):
        return list(orbit._stateVectors)
    if hasattr(orbit, "stateVectors"):
        return list(orbit.stateVectors)
    return list(iter(orbit))


def export_rows(orbit, reference):
    """Orbit.exportToC: rows [t, x, y, z, vx, vy, vz] with t in seconds since 00:00 of `reference`'s day."""
    ref_epoch = reference.replace(hour=0, minute=0, second=0, microsecond=0)
    rows = []
    for sv in state_vectors(orbit):
        tim = (sv.getTime() - ref_epoch).total_seconds()
        rows.append([tim] + list(sv.getPosition()) + list(sv.getVelocity()))
    a = np.array(rows, dtype=np.float64)
    return np.ascontiguousarray(a[:, 0]), np.ascontiguousarray(a[:, 1:4]), np.ascontiguousarray(a[:, 4:7])


def seconds_since_midnight(dt):
    """components/iscesys/DateTimeUtil/DateTimeUtil.py:50-57"""
    return dt.hour * 3600.0 + dt.minute * 60.0 + dt.second + dt.microsecond * 1e-6


def _hermite4(t, pos, vel, tq):
    """Same 4-point Hermite scheme as orbitHermite.c:4-94, window chosen as in orbit.c:203-211 (host helper for the
    single evaluation behind the default peg heading)."""
    n = len(t)
    i = 0
    while i < n and t[i] < tq:
        i += 1
    i = min(max(i - 2, 0), n - 4)
    tt, x, v = t[i:i + 4], pos[i:i + 4], vel[i:i + 4]
    xx, vv = np.zeros(3), np.zeros(3)
    for a in range(4):
        s = sum(1.0 / (tt[a] - tt[j]) for j in range(4) if j != a)
        f0 = 1.0 - 2.0 * (tq - tt[a]) * s
        f1 = tq - tt[a]
        h = 1.0
        for k in range(4):
            if k != a:
                h *= (tq - tt[k]) / (tt[a] - tt[k])
        hdot = 0.0
        for j in range(4):
            if j == a:
                continue
            pr = 1.0
            for k in range(4):
                if k != a and k != j:
                    pr *= (tq - tt[k]) / (tt[a] - tt[k])
            hdot += pr / (tt[a] - tt[j])
        g1 = h + 2.0 * (tq - tt[a]) * hdot
        g0 = 2.0 * (f0 * hdot - h * s)
        xx += (x[a] * f0 + v[a] * f1) * h * h
        vv += (x[a] * g0 + v[a] * g1) * h
    return xx, vv


def enu_heading_deg(orbit, time, a=EarthMajorSemiAxis, e2=EarthEccentricitySquared):
    """Orbit.getENUHeading (Orbit.py:804-831): heading of the Hermite-interpolated velocity in the local ENU frame."""
    svs = state_vectors(orbit)
    ref = svs[0].getTime()
    t = np.array([(s.getTime() - ref).total_seconds() for s in svs])
    pos = np.array([s.getPosition() for s in svs], dtype=np.float64)
    vel = np.array([s.getVelocity() for s in svs], dtype=np.float64)
    if time is None:
        tq = 0.5 * (t[0] + t[-1])
    else:
        tq = (time - ref).total_seconds()
    p, v = _hermite4(t, pos, vel, tq)
    from .synth import xyz_to_llh
    lat, lon, _ = xyz_to_llh(p, a, e2)
    lat, lon = math.radians(float(lat)), math.radians(float(lon))
    east = np.array([-math.sin(lon), math.cos(lon), 0.0])
    north = np.array([-math.sin(lat) * math.cos(lon), -math.sin(lat) * math.sin(lon), math.cos(lat)])
    return math.degrees(math.atan2(float(np.dot(east, v)), float(np.dot(north, v))))
