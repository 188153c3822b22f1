#!/usr/bin/env python
"""Generate tests/golden/ref_python_vectors.json by running the REFERENCE's own Python code.

Run only in the build container (needs /root/reference); the JSON it writes is committed so
that the tests can run where the reference tree is absent (the GPU box).

The reference's pure-Python geometry (isceobj.Planet.Ellipsoid, isceobj.Orbit.Orbit incl. its
single-point rdr2geo / geo2rdr, components/isceobj/Orbit/Orbit.py:834-916,1000-1057) imports
cleanly once two compiled-module imports are stubbed (``isce`` top-level package and
``iscesys.StdOEL.StdOEL``); none of the stubbed code is on the arithmetic path.
"""
import datetime
import json
import logging
import os
import sys
import types

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def import_reference():
    sys.path.insert(0, os.path.join(REF, "components"))
    isce = types.ModuleType("isce")
    isce.logging = logging
    isce.__version__ = "2.6.3"
    sys.modules["isce"] = isce
    import iscesys.StdOEL
    st = types.ModuleType("iscesys.StdOEL.StdOEL")
    sys.modules["iscesys.StdOEL.StdOEL"] = st
    iscesys.StdOEL.StdOEL = st
    logging.disable(logging.CRITICAL)
    # Orbit._hermiteOrbitInterpolation dlopens <Orbit dir>/orbitHermite.so (C wrapper + Fortran in the official
    # build).  The reference also ships the same routine in pure C (components/isceobj/Orbit/src/
    # orbitHermiteInC.c); oracle/Makefile compiles that file unchanged into oracle/_ref/orbitHermite.so and the
    # loader is redirected to it (the reference tree is read-only).
    import ctypes
    real_load = ctypes.cdll.LoadLibrary
    redirect = os.path.join(HERE, "..", "..", "oracle", "_ref", "orbitHermite.so")
    ctypes.cdll.LoadLibrary = lambda name: real_load(redirect if str(name).endswith("orbitHermite.so") else name)
    from isceobj.Orbit.Orbit import Orbit, StateVector
    from isceobj.Planet.Planet import Planet
    # Orbit.geo2rdr imports isceobj.Util.Poly2D (an installed-tree alias of Util/Library/python/Poly2D.py that
    # needs the compiled combinedlibmodule) before looking at its arguments; we always pass an explicit
    # zero-Doppler callable, so an empty stand-in module is enough.
    import isceobj.Util
    p2 = types.ModuleType("isceobj.Util.Poly2D")
    p2.Poly2D = type("Poly2D", (), {})
    sys.modules["isceobj.Util.Poly2D"] = p2
    return Orbit, StateVector, Planet


def load_rsc_orbit():
    rows = []
    with open(os.path.join(REF, "components/isceobj/Util/Library/orbit/test/hdr_WGS84.rsc")) as f:
        for line in f:
            v = [float(x) for x in line.split()]
            if len(v) == 7:
                rows.append(v)
    return rows


def main():
    Orbit, StateVector, Planet = import_reference()
    planet = Planet(pname="Earth")
    elp = planet.ellipsoid
    day = datetime.datetime(2010, 1, 1)
    rows = load_rsc_orbit()
    orb = Orbit()
    orb.configure()
    for r in rows:
        sv = StateVector()
        sv.configure()
        sv.setTime(day + datetime.timedelta(seconds=r[0]))
        sv.setPosition(r[1:4])
        sv.setVelocity(r[4:7])
        orb.addStateVector(sv)

    out = {"generator": "tests/golden/make_golden.py (reference Python imported from /root/reference)",
           "ellipsoid": {"a": elp.a, "e2": elp.e2}, "orbit_rsc": rows, "day": day.isoformat()}

    # --- ellipsoid ---
    pts = [[40.15, -104.97, 2119.0], [0.0, 0.0, 0.0], [-33.3, 151.2, 55.5], [71.0, -156.8, 12.0],
           [36.87, -113.97, 1000.14], [-89.0, 10.0, 3000.0], [12.5, 179.5, -30.0]]
    out["llh_to_xyz"] = [{"llh": p, "xyz": list(map(float, elp.llh_to_xyz(p)))} for p in pts]
    out["xyz_to_llh"] = [{"xyz": e["xyz"], "llh": list(map(float, elp.xyz_to_llh(e["xyz"])))} for e in out["llh_to_xyz"]]
    out["radii"] = [{"lat_deg": la, "east": float(elp.eastRadiusOfCurvature([la, 0.0, 0.0])),
                     "north": float(elp.northRadiusOfCurvature([la, 0.0, 0.0])),
                     "hdg_deg": hd, "dir": float(elp.radiusOfCurvature([la, 0.0, 0.0], hdg=hd))}
                    for la, hd in [(40.0, 90.0), (40.0, 0.0), (40.0, 60.0), (33.5340581084, -166.483356977), (-12.0, 193.0)]]

    # --- orbit interpolation ---
    times = [59030.0 + x for x in (0.0, 3.7, 15.0, 55.0, 61.25, 70.0, 99.999, 127.5, 140.0)]
    interp = []
    for tq in times:
        dt = day + datetime.timedelta(seconds=tq)
        e = {"t": tq}
        for m in ("hermite", "legendre"):
            try:
                sv = orb.interpolateOrbit(dt, method=m)
            except Exception:  # the Python Legendre refuses epochs without 4/5 bracketing vectors
                e[m] = None
                continue
            e[m] = None if sv is None else {"pos": list(map(float, sv.getPosition())),
                                            "vel": list(map(float, sv.getVelocity()))}
        e["enu_heading_deg"] = float(orb.getENUHeading(dt)) if e["hermite"] is not None else None
        interp.append(e)
    out["interp"] = interp

    # --- rdr2geo (constant-height solve) and geo2rdr ---
    r2g = []
    for tq, rng, h, side in [(59085.0, 850000.0, 0.0, -1), (59085.0, 850000.0, 1500.0, -1), (59072.5, 900000.0, 250.0, -1),
                             (59100.25, 830000.0, -20.0, 1), (59061.0, 1000000.0, 3000.0, 1), (59120.0, 870000.5, 800.0, -1)]:
        dt = day + datetime.timedelta(seconds=tq)
        llh = orb.rdr2geo(dt, rng, height=h, side=side)
        hdg = float(orb.getENUHeading(dt))
        e = {"t": tq, "rng": rng, "height": h, "side": side, "hdg_deg": hdg, "llh": list(map(float, llh))}
        tg, rg = orb.geo2rdr([float(llh[0]), float(llh[1]), h], side=side, doppler=lambda t, r: 0.0, wvl=0.0)
        e["geo2rdr_t"] = (tg - day).total_seconds()
        e["geo2rdr_rng"] = float(rg)
        r2g.append(e)
    out["rdr2geo"] = r2g

    with open(os.path.join(HERE, "ref_python_vectors.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", os.path.join(HERE, "ref_python_vectors.json"))


if __name__ == "__main__":
    main()
