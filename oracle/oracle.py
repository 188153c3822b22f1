"""ctypes front-end of the CPU oracle -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product (isce2_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libzerodop_oracle.so")
REF_LIB_PATH = os.path.join(HERE, "_ref", "libisce2_c_ref.so")
# the reference's C++ restatements of the path, compiled unchanged (oracle/Makefile target ref; doors in ref_cpp.py)
REF_CPP_LIBS = {k: os.path.join(HERE, "_ref", f"libisce2_cpp{k}_ref.so") for k in ("topo", "geo", "resamp")}

DEM_METHODS = {"SINC": 0, "BILINEAR": 1, "BICUBIC": 2, "NEAREST": 3, "AKIMA": 4, "BIQUINTIC": 5}
ORBIT_METHODS = {"HERMITE": 0, "SCH": 1, "LEGENDRE": 2}

_dp = C.POINTER(C.c_double)
_fp = C.POINTER(C.c_float)


class OrcOrbit(C.Structure):
    _fields_ = [("nvec", C.c_int), ("t", _dp), ("pos", _dp), ("vel", _dp)]


class OrcPoly2d(C.Structure):
    _fields_ = [("range_order", C.c_int), ("azimuth_order", C.c_int), ("mean_range", C.c_double),
                ("mean_azimuth", C.c_double), ("norm_range", C.c_double), ("norm_azimuth", C.c_double),
                ("coeffs", _dp)]


class OrcPoly1d(C.Structure):
    _fields_ = [("order", C.c_int), ("mean", C.c_double), ("norm", C.c_double), ("coeffs", _dp)]


class OrcTopoParams(C.Structure):
    _fields_ = [("numiter", C.c_int), ("extraiter", C.c_int), ("thresh", C.c_double),
                ("idemwidth", C.c_int), ("idemlength", C.c_int),
                ("firstlat", C.c_double), ("firstlon", C.c_double), ("deltalat", C.c_double), ("deltalon", C.c_double),
                ("major", C.c_double), ("e2", C.c_double),
                ("length", C.c_int), ("width", C.c_int), ("nrnglooks", C.c_int), ("nazlooks", C.c_int),
                ("peghdg", C.c_double), ("prf", C.c_double), ("t0", C.c_double), ("wvl", C.c_double),
                ("ilrl", C.c_int), ("method", C.c_int), ("orbitmethod", C.c_int)]


class OrcTopoResult(C.Structure):
    _fields_ = [("min_lat", C.c_double), ("max_lat", C.c_double), ("min_lon", C.c_double), ("max_lon", C.c_double),
                ("totalconv", C.c_longlong), ("total_iters", C.c_longlong),
                ("ustartx", C.c_int), ("ustarty", C.c_int), ("udemwidth", C.c_int), ("udemlength", C.c_int),
                ("ufirstlat", C.c_double), ("ufirstlon", C.c_double), ("demmax", C.c_float)]


class OrcGeoParams(C.Structure):
    _fields_ = [("major", C.c_double), ("e2", C.c_double), ("drho", C.c_double), ("rho0", C.c_double),
                ("wvl", C.c_double), ("t0", C.c_double), ("prf", C.c_double),
                ("length", C.c_int), ("width", C.c_int), ("ilrl", C.c_int),
                ("nrnglooks", C.c_int), ("nazlooks", C.c_int), ("demwidth", C.c_int), ("demlength", C.c_int),
                ("bistatic", C.c_int), ("orbitmethod", C.c_int)]


class OrcGeoResult(C.Structure):
    _fields_ = [("num_outside", C.c_longlong), ("num_valid", C.c_longlong), ("num_conv", C.c_longlong),
                ("total_iters", C.c_longlong)]


class OrcGeozeroParams(C.Structure):
    _fields_ = [("major", C.c_double), ("e2", C.c_double), ("min_lat", C.c_double), ("min_lon", C.c_double),
                ("max_lat", C.c_double), ("max_lon", C.c_double), ("drho", C.c_double), ("rho0", C.c_double),
                ("wvl", C.c_double), ("t0", C.c_double), ("prf", C.c_double), ("length", C.c_int), ("width", C.c_int),
                ("nrnglooks", C.c_int), ("nazlooks", C.c_int), ("lat_first", C.c_double), ("lon_first", C.c_double),
                ("dlat", C.c_double), ("dlon", C.c_double), ("demwidth", C.c_int), ("demlength", C.c_int)]


class OrcGeozeroResult(C.Structure):
    _fields_ = [("geowidth", C.c_int), ("geolength", C.c_int), ("geomin_lat", C.c_double), ("geomax_lat", C.c_double),
                ("geomin_lon", C.c_double), ("geomax_lon", C.c_double), ("num_outside_dem", C.c_longlong),
                ("num_outside_image", C.c_longlong), ("num_valid", C.c_longlong), ("total_iters", C.c_longlong)]


class OrcResampParams(C.Structure):
    _fields_ = [("inwidth", C.c_int), ("inlength", C.c_int), ("outwidth", C.c_int), ("outlength", C.c_int),
                ("wvl", C.c_double), ("slr", C.c_double), ("r0", C.c_double), ("refwvl", C.c_double), ("refr0", C.c_double),
                ("refslr", C.c_double), ("flatten", C.c_int)]


def build(force=False):
    """Compile the oracle (and oracle/_ref when /root/reference is mounted)."""
    src = os.path.join(HERE, "zerodop_oracle.c")
    stale = (not os.path.exists(LIB_PATH)) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src)
    if force or stale:
        subprocess.check_call(["make", "-C", HERE, "libzerodop_oracle.so"], stdout=subprocess.DEVNULL)
    ref_libs = [REF_LIB_PATH] + list(REF_CPP_LIBS.values())
    if os.path.isdir("/root/reference") and (force or not all(os.path.exists(q) for q in ref_libs)):
        subprocess.check_call(["make", "-C", HERE, "ref"], stdout=subprocess.DEVNULL)


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        L.orc_latlon.argtypes = [C.c_double, C.c_double, _dp, _dp, C.c_int]
        for f in (L.orc_reast, L.orc_rnorth):
            f.restype = C.c_double
            f.argtypes = [C.c_double] * 3
        L.orc_rdir.restype = C.c_double
        L.orc_rdir.argtypes = [C.c_double] * 4
        L.orc_tcnbasis.argtypes = [_dp, _dp, C.c_double, C.c_double, _dp, _dp, _dp]
        L.orc_enubasis.argtypes = [C.c_double, C.c_double, _dp]
        L.orc_radar_to_xyz.restype = C.c_double
        L.orc_radar_to_xyz.argtypes = [C.c_double] * 5 + [_dp, _dp]
        L.orc_xyz_to_sch.argtypes = [_dp, _dp, C.c_double, _dp, _dp]
        for f in (L.orc_interp_hermite, L.orc_interp_legendre, L.orc_interp_sch):
            f.restype = C.c_int
            f.argtypes = [C.POINTER(OrcOrbit), C.c_double, _dp, _dp]
        L.orc_compute_acceleration.restype = C.c_int
        L.orc_compute_acceleration.argtypes = [C.POINTER(OrcOrbit), C.c_double, _dp]
        L.orc_eval_poly2d.restype = C.c_double
        L.orc_eval_poly2d.argtypes = [C.POINTER(OrcPoly2d), C.c_double, C.c_double]
        L.orc_eval_poly1d.restype = C.c_double
        L.orc_eval_poly1d.argtypes = [C.POINTER(OrcPoly1d), C.c_double]
        L.orc_interp_dem.restype = C.c_float
        L.orc_interp_dem.argtypes = [C.c_int, _fp, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int]
        _ip = C.POINTER(C.c_int)
        L.orc_interp_dem_batch.argtypes = [C.c_int, _fp, C.c_int, C.c_int, C.c_long, _ip, _ip, _dp, _dp, _fp]
        L.orc_test_set_cpp_quirks.argtypes = [C.c_int]
        L.orc_test_set_sinc_table.argtypes = [_fp]
        L.orc_sinc_table.argtypes = [_fp]
        L.orc_insertion_sort.argtypes = [_dp, _dp, _dp, C.c_int]
        L.orc_binarysearch.restype = C.c_int
        L.orc_binarysearch.argtypes = [_dp, C.c_int, C.c_double]
        L.orc_topo.restype = C.c_int
        L.orc_topo.argtypes = [C.POINTER(OrcTopoParams), _fp, C.POINTER(OrcOrbit), C.POINTER(OrcPoly2d),
                               C.POINTER(OrcPoly2d), _dp, C.c_int, C.c_int, _dp, _dp, _dp, _fp, _fp,
                               C.POINTER(C.c_int8), C.POINTER(OrcTopoResult), C.c_int]
        L.orc_geo2rdr.restype = C.c_int
        L.orc_geo2rdr.argtypes = [C.POINTER(OrcGeoParams), _dp, _dp, _dp, C.POINTER(OrcOrbit), C.POINTER(OrcPoly1d),
                                  C.c_int, C.c_int, _dp, _dp, _dp, _dp, C.POINTER(OrcGeoResult), C.c_int]
        L.orc_geozero_grid.argtypes = [C.POINTER(OrcGeozeroParams)] + [C.POINTER(C.c_int)] * 6
        L.orc_geozero.restype = C.c_int
        L.orc_geozero.argtypes = [C.POINTER(OrcGeozeroParams), _fp, C.POINTER(OrcOrbit), C.POINTER(OrcPoly1d), _fp, C.c_int,
                                  C.c_int, C.c_int, _fp, C.POINTER(C.c_int16), _dp, _dp, C.POINTER(OrcGeozeroResult), C.c_int]
        L.orc_resamp_sinc_table.argtypes = [_fp]
        L.orc_resamp_slc.restype = C.c_int
        L.orc_resamp_slc.argtypes = [C.POINTER(OrcResampParams)] + [C.POINTER(OrcPoly2d)] * 5 + [_fp, _dp, _dp, _fp, C.c_int]
        _lib = L
    return _lib


def _d(a):
    return a.ctypes.data_as(_dp)


def _f(a):
    return a.ctypes.data_as(_fp)


class Orbit:
    def __init__(self, t, pos, vel):
        self.t = np.ascontiguousarray(t, np.float64)
        self.pos = np.ascontiguousarray(pos, np.float64).reshape(-1, 3)
        self.vel = np.ascontiguousarray(vel, np.float64).reshape(-1, 3)
        self.c = OrcOrbit(len(self.t), _d(self.t), _d(self.pos), _d(self.vel))

    def interp(self, tq, method="HERMITE"):
        p = np.zeros(3)
        v = np.zeros(3)
        fn = {"HERMITE": lib().orc_interp_hermite, "LEGENDRE": lib().orc_interp_legendre,
              "SCH": lib().orc_interp_sch}[method.upper()]
        stat = fn(C.byref(self.c), float(tq), _d(p), _d(v))
        return stat, p, v

    def acceleration(self, tq):
        a = np.zeros(3)
        stat = lib().orc_compute_acceleration(C.byref(self.c), float(tq), _d(a))
        return stat, a


class Poly2D:
    def __init__(self, coeffs, mean_range=0.0, mean_azimuth=0.0, norm_range=1.0, norm_azimuth=1.0):
        self.coeffs = np.ascontiguousarray(np.atleast_2d(np.asarray(coeffs, np.float64)))
        az, rg = self.coeffs.shape
        self.c = OrcPoly2d(rg - 1, az - 1, mean_range, mean_azimuth, norm_range, norm_azimuth, _d(self.coeffs))

    def __call__(self, azi, rng):
        return lib().orc_eval_poly2d(C.byref(self.c), float(azi), float(rng))


class Poly1D:
    def __init__(self, coeffs, mean=0.0, norm=1.0):
        self.coeffs = np.ascontiguousarray(np.asarray(coeffs, np.float64).ravel())
        self.c = OrcPoly1d(len(self.coeffs) - 1, mean, norm, _d(self.coeffs))

    def __call__(self, x):
        return lib().orc_eval_poly1d(C.byref(self.c), float(x))


def latlon_to_xyz(llh_rad, a, e2):
    llh = np.ascontiguousarray(llh_rad, np.float64)
    xyz = np.zeros(3)
    lib().orc_latlon(a, e2, _d(xyz), _d(llh), 1)
    return xyz


def xyz_to_latlon(xyz, a, e2):
    xyz = np.ascontiguousarray(xyz, np.float64)
    llh = np.zeros(3)
    lib().orc_latlon(a, e2, _d(xyz), _d(llh), 2)
    return llh


def interp_dem(method, dem, ix, iy, fx, fy):
    dem = np.ascontiguousarray(dem, np.float32)
    ny, nx = dem.shape
    return float(lib().orc_interp_dem(DEM_METHODS[method.upper()], _f(dem), int(ix), int(iy), float(fx), float(fy), nx, ny))


def interp_dem_batch(method, dem, ix, iy, fx, fy, cpp_quirks=0):
    """orc_interp_dem over arrays.  cpp_quirks: test hook of tests/test_oracle_cpp_pins.py (see zerodop_oracle.c)."""
    dem = np.ascontiguousarray(dem, np.float32)
    ny, nx = dem.shape
    ix = np.ascontiguousarray(ix, np.int32)
    iy = np.ascontiguousarray(iy, np.int32)
    fx = np.ascontiguousarray(fx, np.float64)
    fy = np.ascontiguousarray(fy, np.float64)
    out = np.empty(ix.size, np.float32)
    ip = C.POINTER(C.c_int)
    lib().orc_test_set_cpp_quirks(int(cpp_quirks))
    try:
        lib().orc_interp_dem_batch(DEM_METHODS[method.upper()], _f(dem), nx, ny, ix.size, ix.ctypes.data_as(ip),
                                   iy.ctypes.data_as(ip), _d(fx), _d(fy), _f(out))
    finally:
        lib().orc_test_set_cpp_quirks(0)
    return out


def sinc_table():
    out = np.empty(8192 * 8, np.float32)
    lib().orc_sinc_table(_f(out))
    return out


def topo(*, dem, first_lat, first_lon, delta_lat, delta_lon, orbit_t, orbit_pos, orbit_vel, length, width,
         r0, dr, prf, t0, wvl, side, peg_heading, doppler_coeffs=((0.0,),), a=6378137.0, e2=0.0066943799901,
         dem_method="BILINEAR", orbit_method="HERMITE", numiter=25, extraiter=10, thresh=0.05,
         nrnglooks=1, nazlooks=1, want_inc=True, want_mask=True, want_los=True, line0=0, nlines=-1,
         rho_image=None, nthreads=0):
    """Run the CPU oracle topo.  Returns dict of arrays + result struct fields."""
    dem = np.ascontiguousarray(dem, np.float32)
    orb = Orbit(orbit_t, orbit_pos, orbit_vel)
    dop = Poly2D(doppler_coeffs)
    slr = Poly2D([[r0, dr * nrnglooks]])
    p = OrcTopoParams(numiter, extraiter, thresh, dem.shape[1], dem.shape[0], first_lat, first_lon, delta_lat,
                      delta_lon, a, e2, length, width, nrnglooks, nazlooks, peg_heading, prf, t0, wvl, side,
                      DEM_METHODS[dem_method.upper()], ORBIT_METHODS[orbit_method.upper()])
    n = length - line0 if nlines < 0 else nlines
    out = dict(lat=np.empty((n, width)), lon=np.empty((n, width)), hgt=np.empty((n, width)))
    out["los"] = np.empty((n, 2, width), np.float32) if want_los else None
    out["inc"] = np.empty((n, 2, width), np.float32) if want_inc else None
    out["mask"] = np.empty((n, width), np.int8) if want_mask else None
    res = OrcTopoResult()
    rimg = None
    if rho_image is not None:
        rimg = np.ascontiguousarray(rho_image, np.float64)
    rc = lib().orc_topo(C.byref(p), _f(dem), C.byref(orb.c), C.byref(dop.c), C.byref(slr.c),
                        _d(rimg) if rimg is not None else None, line0, n,
                        _d(out["lat"]), _d(out["lon"]), _d(out["hgt"]),
                        _f(out["los"]) if want_los else None, _f(out["inc"]) if want_inc else None,
                        out["mask"].ctypes.data_as(C.POINTER(C.c_int8)) if want_mask else None,
                        C.byref(res), nthreads)
    if rc != 0:
        raise RuntimeError(f"orc_topo failed rc={rc}")
    for k, _ in OrcTopoResult._fields_:
        out[k] = getattr(res, k)
    out["mean_iters"] = res.total_iters / float(n * width)
    return out


def geo2rdr(*, lat, lon, hgt, orbit_t, orbit_pos, orbit_vel, length, width, r0, dr, prf, t0, wvl, side=-1,
            doppler_coeffs=(0.0,), doppler_mean=0.0, doppler_norm=1.0, a=6378137.0, e2=0.0066943799901,
            orbit_method="HERMITE", bistatic=False, nrnglooks=1, nazlooks=1, line0=0, nlines=-1, nthreads=0,
            want=("azt", "rgm", "azoff", "rgoff")):
    lat = np.ascontiguousarray(lat, np.float64)
    lon = np.ascontiguousarray(lon, np.float64)
    hgt = np.ascontiguousarray(hgt, np.float64)
    demlength, demwidth = lat.shape
    orb = Orbit(orbit_t, orbit_pos, orbit_vel)
    dop = Poly1D(doppler_coeffs, doppler_mean, doppler_norm)
    p = OrcGeoParams(a, e2, dr, r0, wvl, t0, prf, length, width, side, nrnglooks, nazlooks, demwidth, demlength,
                     int(bool(bistatic)), ORBIT_METHODS[orbit_method.upper()])
    n = demlength - line0 if nlines < 0 else nlines
    out = {k: (np.empty((n, demwidth)) if k in want else None) for k in ("azt", "rgm", "azoff", "rgoff")}
    res = OrcGeoResult()
    rc = lib().orc_geo2rdr(C.byref(p), _d(lat), _d(lon), _d(hgt), C.byref(orb.c), C.byref(dop.c), line0, n,
                           *[(_d(out[k]) if out[k] is not None else None) for k in ("azt", "rgm", "azoff", "rgoff")],
                           C.byref(res), nthreads)
    if rc != 0:
        raise RuntimeError(f"orc_geo2rdr failed rc={rc}")
    for k, _ in OrcGeoResult._fields_:
        out[k] = getattr(res, k)
    out["mean_iters"] = res.total_iters / float(n * demwidth)
    return out


def scene_topo_kwargs(sc, **over):
    """Map an isce2_b200.synth.Scene onto topo() keyword arguments."""
    kw = dict(dem=sc.dem, first_lat=sc.first_lat, first_lon=sc.first_lon, delta_lat=sc.delta_lat,
              delta_lon=sc.delta_lon, orbit_t=sc.orbit_t, orbit_pos=sc.orbit_pos, orbit_vel=sc.orbit_vel,
              length=sc.length, width=sc.width, r0=sc.r0, dr=sc.dr, prf=sc.prf, t0=sc.t0, wvl=sc.wvl,
              side=sc.side, peg_heading=sc.peg_heading, doppler_coeffs=sc.doppler_coeffs, a=sc.a, e2=sc.e2)
    kw.update(over)
    return kw


GEOZERO_METHODS = {"SINC": 0, "BILINEAR": 1, "BICUBIC": 2, "NEAREST": 3}


def geozero_params(*, dem_shape, first_lat, first_lon, delta_lat, delta_lon, snwe, length, width, r0, dr, prf, t0, wvl,
                   a=6378137.0, e2=0.0066943799901, nrnglooks=1, nazlooks=1):
    return OrcGeozeroParams(a, e2, snwe[0], snwe[2], snwe[1], snwe[3], dr, r0, wvl, t0, prf, length, width, nrnglooks,
                            nazlooks, first_lat, first_lon, delta_lat, delta_lon, dem_shape[1], dem_shape[0])


def geozero_grid(p):
    v = [C.c_int() for _ in range(6)]
    lib().orc_geozero_grid(C.byref(p), *[C.byref(x) for x in v])
    return dict(zip(("geo_len", "geo_wid", "min_lat_idx", "max_lat_idx", "min_lon_idx", "max_lon_idx"), [x.value for x in v]))


def geozero(*, dem, image, orbit_t, orbit_pos, orbit_vel, method="BILINEAR", side=-1, doppler_coeffs=(0.0,),
            doppler_mean=0.0, doppler_norm=1.0, nthreads=0, want_indices=True, **grid_kw):
    """One band through geozero.  image: [length][width] float32 or complex64; returns dict(geo, dem_crop, az_idx, rng_idx, ...)."""
    dem = np.ascontiguousarray(dem, np.float32)
    image = np.ascontiguousarray(image)
    iscomplex = np.iscomplexobj(image)
    image = image.astype(np.complex64 if iscomplex else np.float32, copy=False)
    length, width = image.shape
    p = geozero_params(dem_shape=dem.shape, length=length, width=width, **grid_kw)
    g = geozero_grid(p)
    shape = (g["geo_len"], g["geo_wid"])
    out = np.zeros(shape, image.dtype)
    crop = np.zeros(shape, np.int16)
    az = np.full(shape, np.nan) if want_indices else None
    rg = np.full(shape, np.nan) if want_indices else None
    orb = Orbit(orbit_t, orbit_pos, orbit_vel)
    dop = Poly1D(doppler_coeffs, doppler_mean, doppler_norm)
    res = OrcGeozeroResult()
    rc = lib().orc_geozero(C.byref(p), _f(dem), C.byref(orb.c), C.byref(dop.c), image.ctypes.data_as(_fp), int(iscomplex),
                           GEOZERO_METHODS[method.upper()], int(side), out.ctypes.data_as(_fp),
                           crop.ctypes.data_as(C.POINTER(C.c_int16)), _d(az) if az is not None else None,
                           _d(rg) if rg is not None else None, C.byref(res), nthreads)
    if rc != 0:
        raise RuntimeError(f"orc_geozero failed rc={rc}")
    r = dict(geo=out, dem_crop=crop, az_idx=az, rng_idx=rg, grid=g)
    for k, _ in OrcGeozeroResult._fields_:
        r[k] = getattr(res, k)
    return r


def _poly2d_or_none(p):
    """p: None, a Poly2D of this module, or (coeffs[, mean_range, mean_azimuth, norm_range, norm_azimuth])."""
    if p is None or isinstance(p, Poly2D):
        return p
    if isinstance(p, (list, tuple)) and len(p) and not np.isscalar(p[0]) and np.ndim(p[0]) == 2:
        return Poly2D(*p)
    return Poly2D(p)


def resamp_slc(*, slc, out_shape, wvl=0.056, slr=2.3, r0=0.0, ref_wvl=None, ref_r0=None, ref_slr=None, flatten=False,
               rg_carrier=None, az_carrier=None, rg_offsets=None, az_offsets=None, doppler=None, resid_az=None, resid_rg=None,
               nthreads=0, cpp_positions=False):
    """resamp_slc.f90 on one complex64 image; polynomials as oracle.Poly2D / coefficient lists / None (zero).
    cpp_positions (tests only): evaluate the Doppler / carrier polynomials where the reference's C++ restatement does."""
    slc = np.ascontiguousarray(slc, np.complex64)
    inlength, inwidth = slc.shape
    outlength, outwidth = out_shape
    p = OrcResampParams(inwidth, inlength, outwidth, outlength, wvl, slr, r0, wvl if ref_wvl is None else ref_wvl,
                        r0 if ref_r0 is None else ref_r0, slr if ref_slr is None else ref_slr, int(bool(flatten)))
    polys = [_poly2d_or_none(q) for q in (rg_carrier, az_carrier, rg_offsets, az_offsets, doppler)]
    ra = np.ascontiguousarray(resid_az, np.float64) if resid_az is not None else None
    rr = np.ascontiguousarray(resid_rg, np.float64) if resid_rg is not None else None
    out = np.zeros((outlength, outwidth), np.complex64)
    lib().orc_test_set_cpp_quirks(4 if cpp_positions else 0)
    try:
        rc = lib().orc_resamp_slc(C.byref(p), *[(C.byref(q.c) if q is not None else None) for q in polys],
                                  slc.ctypes.data_as(_fp), _d(ra) if ra is not None else None, _d(rr) if rr is not None else None,
                                  out.ctypes.data_as(_fp), nthreads)
    finally:
        lib().orc_test_set_cpp_quirks(0)
    if rc != 0:
        raise RuntimeError(f"orc_resamp_slc failed rc={rc}")
    return out


# ---- multilooking / mask projection (other consumers): plain numpy restatements ----
def _to_line_band_sample(a, scheme):
    """View of an image stored in `scheme` order as [line][band][sample]."""
    if a.ndim == 2:
        return a[:, None, :]
    return {"BIL": a, "BIP": np.moveaxis(a, 2, 1), "BSQ": np.moveaxis(a, 0, 1)}[scheme.upper()]


def _from_line_band_sample(o, scheme, ndim):
    if ndim == 2:
        return np.ascontiguousarray(o[:, 0, :])
    return np.ascontiguousarray({"BIL": o, "BIP": np.moveaxis(o, 1, 2), "BSQ": np.moveaxis(o, 1, 0)}[scheme.upper()])


def looks(image, down_looks, across_looks, scheme="BIL", method="AVERAGE"):
    """takeLooks<T> / takeLookscpx<T> (components/mroipac/looks/bindings/looksmodule.cpp:130-200, :204-275) or, for
    method 'NEAREST', the gdal.Translate -outsize decimation of runMultilook (contrib/stack/stripmapStack/topo.py:411-424).
    The accumulation order is the reference's: the `down` lines are added one after the other into a double line
    buffer (:160-177), then the `across` neighbours one after the other (:179-190), / double(down*across), cast to T."""
    a = np.asarray(image)
    v = _to_line_band_sample(a, scheme)
    nd, _, na = v.shape
    ld, la = int(down_looks), int(across_looks)
    ol, ow = nd // ld, na // la
    if method.upper() in ("NEAREST", "GDAL"):
        o = v[ld // 2:ol * ld:ld, :, la // 2:ow * la:la][:ol, :, :ow]  # source index floor((i + 0.5) * looks)
        return _from_line_band_sample(o, scheme, a.ndim)
    acc_t = np.complex128 if np.iscomplexobj(a) else np.float64
    o = np.zeros((ol, v.shape[1], ow), a.dtype)
    norm = float(ld * la)
    for line in range(ol):
        bdbl = np.zeros(v.shape[1:], acc_t)
        for i in range(ld):  # bdbl[j] += ain[j]
            bdbl = bdbl + v[line * ld + i].astype(acc_t)
        s = np.zeros((v.shape[1], ow), acc_t)
        for k in range(la):  # sum += bdbl[(j + k) * bands + b]
            s = s + bdbl[:, k:ow * la:la][:, :ow]
        if np.iscomplexobj(a):  # complex<T>(static_cast<T>(sum.real() / norm), static_cast<T>(sum.imag() / norm))
            o[line] = (s.real / norm).astype(a.real.dtype) + 1j * (s.imag / norm).astype(a.real.dtype)
            continue
        q = s / norm
        if np.issubdtype(a.dtype, np.integer):
            o[line] = np.trunc(q).astype(a.dtype)  # static_cast<T>(double): toward zero
        else:
            o[line] = q.astype(a.dtype)
    return _from_line_band_sample(o, scheme, a.ndim)


def mask_to_radar(mask, start_lat, delta_lat, start_lon, delta_lon, lat, lon):
    """SWBDStitcher.toRadar (contrib/demUtils/swbdstitcher/SWBDStitcher.py:107-131), the arithmetic only (lines :123-126)."""
    lati = np.clip(((lat - start_lat) / delta_lat).astype(int), 0, mask.shape[0] - 1)
    loni = np.clip(((lon - start_lon) / delta_lon).astype(int), 0, mask.shape[1] - 1)
    return (mask[lati, loni] + 1).astype(mask.dtype)
