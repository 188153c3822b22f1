"""ctypes doors onto the reference's own C++ restatements of the path -- TEST INFRASTRUCTURE ONLY.

oracle/_ref/libisce2_cpp{topo,geo,resamp}_ref.so are built by `make -C oracle ref` from the reference's sources,
unchanged, where they lie under /root/reference (components/zerodop/GPUtopozero/src, GPUgeo2rdr/src, GPUresampslc/src;
CPU branches, GPU_ACC_ENABLED undefined) plus the extern "C" shims oracle/ref_cpp_*_shim.cpp.  tests/ use them to hold
oracle/zerodop_oracle.c against reference-authored code (tests/test_oracle_cpp_pins.py).  Never imported by isce2_b200/.
"""
from __future__ import annotations

import ctypes as C
import os
import re
import sys
import tempfile

import numpy as np

from . import oracle as orc

_dp = C.POINTER(C.c_double)
_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int)
_D, _I = C.c_double, C.c_int
_libs = {}


def available():
    orc.build()
    return all(os.path.exists(p) for p in orc.REF_CPP_LIBS.values())


def _lib(kind):
    if kind not in _libs:
        orc.build()
        L = C.CDLL(orc.REF_CPP_LIBS[kind])
        if kind == "topo":
            L.ref_cpp_interp_dem.argtypes = [_I, _fp, _I, _I, C.c_long, _ip, _ip, _dp, _dp, _fp, _fp]
            L.ref_cpp_sinc_table.argtypes = [_fp]
            L.ref_cpp_latlon.argtypes = [_D, _D, _dp, _dp, _I]
            for f in (L.ref_cpp_reast, L.ref_cpp_rnorth):
                f.restype = _D
                f.argtypes = [_D] * 3
            L.ref_cpp_rdir.restype = _D
            L.ref_cpp_rdir.argtypes = [_D] * 4
            L.ref_cpp_tcnbasis.argtypes = [_dp, _dp, _D, _D, _dp, _dp, _dp]
            L.ref_cpp_enubasis.argtypes = [_D, _D, _dp]
            L.ref_cpp_radar_to_xyz.restype = _D
            L.ref_cpp_radar_to_xyz.argtypes = [_D] * 5 + [_dp, _dp]
            L.ref_cpp_convert_sch.argtypes = [_D] * 5 + [_dp, _dp, _I]
            L.ref_cpp_interp_orbit.restype = _I
            L.ref_cpp_interp_orbit.argtypes = [_I, _dp, _dp, _dp, _I, _D, _dp, _dp]
            L.ref_cpp_eval_poly2d.restype = _D
            L.ref_cpp_eval_poly2d.argtypes = [_I, _I, _D, _D, _D, _D, _dp, _D, _D]
            L.ref_cpp_insertion_sort.argtypes = [_dp, _I]
            L.ref_cpp_binary_search.restype = _I
            L.ref_cpp_binary_search.argtypes = [_dp, _I, _D]
            L.ref_cpp_topo.argtypes = [_D] * 11 + [_I] * 11 + [_I, _dp, _dp, _dp, _fp, _dp, _dp, _dp, _dp, _dp, _fp, _fp, _dp]
        elif kind == "geo":
            L.ref_cpp_geo2rdr.argtypes = [_D] * 7 + [_I] * 9 + [_dp, _dp, _dp, _I, _D, _D, _dp] + [_dp] * 7
        elif kind == "resamp":
            L.ref_cpp_sinc_coef.argtypes = [_D, _D, _I, _D, _I, _dp]
            L.ref_cpp_resamp_slc.restype = _I
            L.ref_cpp_resamp_slc.argtypes = [_I] * 4 + [_D] * 6 + [_I] + [_dp] * 5 + [_fp, _dp, _dp, _fp]
        _libs[kind] = L
    return _libs[kind]


def _d(a):
    return a.ctypes.data_as(_dp)


def _f(a):
    return a.ctypes.data_as(_fp)


class _capture_stdout:
    """The reference prints its progress with printf: route fd 1 into a file for the duration of a call."""

    def __enter__(self):
        sys.stdout.flush()
        self._tmp = tempfile.TemporaryFile(mode="w+b")
        self._saved = os.dup(1)
        os.dup2(self._tmp.fileno(), 1)
        self.text = ""
        return self

    def __exit__(self, *exc):
        C.CDLL(None).fflush(None)
        os.dup2(self._saved, 1)
        os.close(self._saved)
        self._tmp.seek(0)
        self.text = self._tmp.read().decode(errors="replace")
        self._tmp.close()
        return False


def interp_dem(method, dem, ix, iy, fx, fy, sinc_table=None):
    """TopoMethods::interpolate over arrays (TopoMethods.cpp:72-84); dem [ny][nx] float32, ix / iy 1-based."""
    dem = np.ascontiguousarray(dem, np.float32)
    ny, nx = dem.shape
    ix = np.ascontiguousarray(ix, np.int32)
    iy = np.ascontiguousarray(iy, np.int32)
    fx = np.ascontiguousarray(fx, np.float64)
    fy = np.ascontiguousarray(fy, np.float64)
    out = np.empty(ix.size, np.float32)
    tab = None if sinc_table is None else np.ascontiguousarray(sinc_table, np.float32)
    with _capture_stdout():
        _lib("topo").ref_cpp_interp_dem(orc.DEM_METHODS[method.upper()], _f(dem), nx, ny, ix.size, ix.ctypes.data_as(_ip),
                                        iy.ctypes.data_as(_ip), _d(fx), _d(fy), _f(out), _f(tab) if tab is not None else None)
    return out


def topo_sinc_table():
    """fintp of TopoMethods::prepareMethods(SINC) -- built by the pre-2021 UniformInterp::sinc_coef."""
    out = np.empty(8192 * 8, np.float32)
    with _capture_stdout():
        _lib("topo").ref_cpp_sinc_table(_f(out))
    return out


def sinc_coef(beta, relfiltlen, decfactor, pedestal, weight):
    """Interpolator::sinc_coef (GPUresampslc/src/Interpolator.cpp:119-136), the post-2021 form."""
    n = int(round(relfiltlen / beta)) * decfactor
    out = np.zeros(n, np.float64)
    _lib("resamp").ref_cpp_sinc_coef(beta, relfiltlen, decfactor, pedestal, weight, _d(out))
    return out


def _poly_block(p):
    """oracle.Poly2D -> [azimuthOrder, rangeOrder, azimuthMean, rangeMean, azimuthNorm, rangeNorm, coeffs ...] (or None)."""
    if p is None:
        return None
    c = p.c
    return np.ascontiguousarray(np.concatenate([[c.azimuth_order, c.range_order, c.mean_azimuth, c.mean_range, c.norm_azimuth,
                                                 c.norm_range], p.coeffs.ravel()]), np.float64)


def resamp_slc(*, slc, out_shape, wvl=0.056, slr=2.3, r0=0.0, ref_wvl=None, ref_r0=None, ref_slr=None, flatten=False,
               rg_carrier=None, az_carrier=None, rg_offsets=None, az_offsets=None, doppler=None, resid_az=None, resid_rg=None):
    """ResampSlc::_resamp_cpu (GPUresampslc/src/ResampSlc.cpp:164-383) on one complex64 image: same keywords as
    oracle.resamp_slc.  Returns (out, printed text)."""
    slc = np.ascontiguousarray(slc, np.complex64)
    inlength, inwidth = slc.shape
    outlength, outwidth = out_shape
    blocks = [_poly_block(orc._poly2d_or_none(q)) for q in (rg_carrier, az_carrier, rg_offsets, az_offsets, doppler)]
    ra = np.ascontiguousarray(resid_az, np.float64) if resid_az is not None else None
    rr = np.ascontiguousarray(resid_rg, np.float64) if resid_rg is not None else None
    out = np.zeros((outlength, outwidth), np.complex64)
    with _capture_stdout() as cap:
        _lib("resamp").ref_cpp_resamp_slc(inwidth, inlength, outwidth, outlength, wvl, slr, r0, wvl if ref_wvl is None else ref_wvl,
                                          slr if ref_slr is None else ref_slr, r0 if ref_r0 is None else ref_r0, int(bool(flatten)),
                                          *[(_d(b) if b is not None else None) for b in blocks], _f(slc.view(np.float32)),
                                          _d(ra) if ra is not None else None, _d(rr) if rr is not None else None,
                                          _f(out.view(np.float32)))
    return out, cap.text


def latlon(a, e2, vec, to_xyz):
    """Ellipsoid::latlon (Ellipsoid.cpp:34-75): LLH (rad, rad, m) -> XYZ when to_xyz, else XYZ -> LLH."""
    v = np.ascontiguousarray(vec, np.float64).copy()
    o = np.zeros(3)
    if to_xyz:
        _lib("topo").ref_cpp_latlon(a, e2, _d(o), _d(v), 1)
    else:
        _lib("topo").ref_cpp_latlon(a, e2, _d(v), _d(o), 2)
    return o


def reast(a, e2, lat):
    return _lib("topo").ref_cpp_reast(a, e2, lat)


def rnorth(a, e2, lat):
    return _lib("topo").ref_cpp_rnorth(a, e2, lat)


def rdir(a, e2, hdg, lat):
    return _lib("topo").ref_cpp_rdir(a, e2, hdg, lat)


def tcnbasis(pos, vel, a, e2):
    pos = np.ascontiguousarray(pos, np.float64)
    vel = np.ascontiguousarray(vel, np.float64)
    t, c, n = np.zeros(3), np.zeros(3), np.zeros(3)
    _lib("topo").ref_cpp_tcnbasis(_d(pos), _d(vel), a, e2, _d(t), _d(c), _d(n))
    return t, c, n


def enubasis(lat, lon):
    m = np.zeros(9)
    _lib("topo").ref_cpp_enubasis(lat, lon, _d(m))
    return m


def radar_to_xyz(a, e2, lat, lon, hdg):
    m, ov = np.zeros(9), np.zeros(3)
    r = _lib("topo").ref_cpp_radar_to_xyz(a, e2, lat, lon, hdg, _d(m), _d(ov))
    return m, ov, r


def xyz_to_sch(a, e2, lat, lon, hdg, xyz):
    x = np.ascontiguousarray(xyz, np.float64).copy()
    s = np.zeros(3)
    _lib("topo").ref_cpp_convert_sch(a, e2, lat, lon, hdg, _d(s), _d(x), 1)
    return s


def interp_orbit(t, pos, vel, method, tq):
    """Orbit::interpolateOrbit (GPUtopozero/src/Orbit.cpp).  Its SCH branch exits the process outside the span."""
    t = np.ascontiguousarray(t, np.float64)
    pos = np.ascontiguousarray(pos, np.float64)
    vel = np.ascontiguousarray(vel, np.float64)
    p, v = np.zeros(3), np.zeros(3)
    stat = _lib("topo").ref_cpp_interp_orbit(len(t), _d(t), _d(pos), _d(vel), orc.ORBIT_METHODS[method.upper()], float(tq),
                                             _d(p), _d(v))
    return stat, p, v


def eval_poly2d(coeffs, azi, rng, mean_range=0.0, mean_azimuth=0.0, norm_range=1.0, norm_azimuth=1.0):
    c = np.ascontiguousarray(np.atleast_2d(np.asarray(coeffs, np.float64)))
    return _lib("topo").ref_cpp_eval_poly2d(c.shape[1] - 1, c.shape[0] - 1, mean_range, mean_azimuth, norm_range, norm_azimuth,
                                            _d(c), float(azi), float(rng))


def insertion_sort(a):
    v = np.ascontiguousarray(a, np.float64).copy()
    _lib("topo").ref_cpp_insertion_sort(_d(v), len(v))
    return v


def binary_search(a, val):
    """LinAlg::binarySearch (LinAlg.cpp): 0-based index."""
    v = np.ascontiguousarray(a, np.float64)
    return _lib("topo").ref_cpp_binary_search(_d(v), len(v), float(val))


def topo(*, dem, first_lat, first_lon, delta_lat, delta_lon, orbit_t, orbit_pos, orbit_vel, length, width, r0, dr, prf, t0,
         wvl, side, peg_heading, doppler_coeffs=((0.0,),), a=6378137.0, e2=0.0066943799901, dem_method="BILINEAR",
         orbit_method="HERMITE", numiter=25, extraiter=10, thresh=0.05, want_mask=True):
    """The CPU branch of Topo::topo (Topo.cpp:127-365, 600-950) on buffers.  Returns the layers in the oracle's layout
    (los / inc as [length][2][width]) plus the DEM crop the reference printed ('crop': width, length, first line/pixel)."""
    dem = np.ascontiguousarray(dem, np.float32)
    ny, nx = dem.shape
    dop2 = orc.Poly2D(doppler_coeffs)
    slr = orc.Poly2D([[r0, dr]])
    # what the Poly2d-backed accessors deliver line by line (Poly2dInterpolator.cpp:5-36), through the reference's own evalPoly2d
    dop = np.array([[dop2(l, c) for c in range(width)] for l in range(length if dop2.coeffs.shape[0] > 1 else 1)])
    if dop.shape[0] == 1:
        dop = np.repeat(dop, length, 0)
    rho = np.repeat(np.array([[slr(0, c) for c in range(width)]]), length, 0)
    dop, rho = np.ascontiguousarray(dop), np.ascontiguousarray(rho)
    lat, lon, hgt = np.zeros((length, width)), np.zeros((length, width)), np.zeros((length, width))
    los, inc = np.zeros((length, width, 2), np.float32), np.zeros((length, width, 2), np.float32)
    mask = np.zeros((length, width)) if want_mask else None
    t = np.ascontiguousarray(orbit_t, np.float64)
    pos = np.ascontiguousarray(orbit_pos, np.float64)
    vel = np.ascontiguousarray(orbit_vel, np.float64)
    import time
    t_call = time.perf_counter()
    with _capture_stdout() as cap:
        _lib("topo").ref_cpp_topo(first_lat, first_lon, delta_lat, delta_lon, a, e2, peg_heading, prf, t0, wvl, thresh, numiter,
                                  extraiter, nx, ny, side, length, width, 1, 1, orc.DEM_METHODS[dem_method.upper()],
                                  orc.ORBIT_METHODS[orbit_method.upper()], len(t), _d(t), _d(pos), _d(vel), _f(dem), _d(dop),
                                  _d(rho), _d(lat), _d(lon), _d(hgt), _f(los), _f(inc), _d(mask) if want_mask else None)
    t_call = time.perf_counter() - t_call
    m = re.search(r"Actual DEM bounds used:\s*Dimensions: (\d+) (\d+).*?Lines: (\d+) (\d+)\s*Pixels: (\d+) (\d+)", cap.text, re.S)
    crop = dict(zip(("width", "length", "line0", "line1", "pixel0", "pixel1"), map(int, m.groups()))) if m else None
    m = re.search(r"Total convergence: (\d+) out of", cap.text)
    return dict(lat=lat, lon=lon, hgt=hgt, los=np.ascontiguousarray(np.moveaxis(los, 2, 1)),
                inc=np.ascontiguousarray(np.moveaxis(inc, 2, 1)), mask=None if mask is None else mask.astype(np.int8), crop=crop,
                totalconv=int(m.group(1)) if m else None, log=cap.text, seconds=t_call)


def geo2rdr(*, lat, lon, hgt, orbit_t, orbit_pos, orbit_vel, length, width, r0, dr, prf, t0, wvl, side=-1,
            doppler_coeffs=(0.0,), doppler_mean=0.0, doppler_norm=1.0, a=6378137.0, e2=0.0066943799901,
            orbit_method="HERMITE", bistatic=False):
    """The CPU branch of Geo2rdr::geo2rdr (Geo2rdr.cpp:84-175, 395-499) on buffers; same keywords as oracle.geo2rdr
    (side is unused by the reference as well: geo2rdr.f90 never reads ilrl)."""
    lat = np.ascontiguousarray(lat, np.float64)
    lon = np.ascontiguousarray(lon, np.float64)
    hgt = np.ascontiguousarray(hgt, np.float64)
    demlength, demwidth = lat.shape
    t = np.ascontiguousarray(orbit_t, np.float64)
    pos = np.ascontiguousarray(orbit_pos, np.float64)
    vel = np.ascontiguousarray(orbit_vel, np.float64)
    dc = np.ascontiguousarray(np.asarray(doppler_coeffs, np.float64).ravel())
    out = {k: np.zeros((demlength, demwidth)) for k in ("azt", "rgm", "azoff", "rgoff")}
    import time
    t_call = time.perf_counter()
    with _capture_stdout() as cap:
        _lib("geo").ref_cpp_geo2rdr(a, e2, dr, r0, wvl, t0, prf, length, width, demlength, demwidth, 1, 1, int(bool(bistatic)),
                                    orc.ORBIT_METHODS[orbit_method.upper()], len(t), _d(t), _d(pos), _d(vel), len(dc) - 1,
                                    float(doppler_mean), float(doppler_norm), _d(dc), _d(lat), _d(lon), _d(hgt),
                                    _d(out["azt"]), _d(out["rgm"]), _d(out["azoff"]), _d(out["rgoff"]))
    out["seconds"] = time.perf_counter() - t_call
    for key, pat in (("num_outside", r"outside the image: (\d+)"), ("num_valid", r"with valid data:\s+(\d+)"),
                     ("num_conv", r"that converged:\s+(\d+)")):
        m = re.search(pat, cap.text)
        out[key] = int(m.group(1)) if m else None
    out["log"] = cap.text
    return out
