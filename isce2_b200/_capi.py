"""Thin ctypes layer over the C ABI of libb200geom.so (include/b200geom.h).

numpy arrays in, numpy arrays out; the GIL is released for the duration of every native call
(ctypes.CDLL does that).  There is no CPU fallback here: if the CUDA library is missing or no
device is visible, the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# B200GEOM_LIB lets a developer A/B an experimental build of the same ABI (never a different backend)
LIB_PATH = os.environ.get("B200GEOM_LIB") or os.path.join(_HERE, "libb200geom.so")

B200_OK = 0
ERRORS = {-1: "EINVAL", -2: "ENODEVICE", -3: "ECUDA", -4: "EORBIT", -5: "EDEM", -6: "ENOMEM"}
DEM_METHODS = {"SINC": 0, "BILINEAR": 1, "BICUBIC": 2, "NEAREST": 3, "AKIMA": 4, "BIQUINTIC": 5}
ORBIT_METHODS = {"HERMITE": 0, "SCH": 1, "LEGENDRE": 2}

_dp = C.POINTER(C.c_double)
_fp = C.POINTER(C.c_float)


class B200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libb200geom: {ERRORS.get(code, code)}: {msg}")
        self.code = code


class Orbit(C.Structure):
    _fields_ = [("nvec", C.c_int), ("t", _dp), ("pos", _dp), ("vel", _dp)]


class Poly2d(C.Structure):
    _fields_ = [("range_order", C.c_int), ("azimuth_order", C.c_int), ("mean_range", C.c_double),
                ("mean_azimuth", C.c_double), ("norm_range", C.c_double), ("norm_azimuth", C.c_double),
                ("coeffs", _dp)]


class Poly1d(C.Structure):
    _fields_ = [("order", C.c_int), ("mean", C.c_double), ("norm", C.c_double), ("coeffs", _dp)]


class TopoParams(C.Structure):
    _fields_ = [("numiter", C.c_int), ("extraiter", C.c_int), ("thresh", C.c_double),
                ("dem_width", C.c_int), ("dem_length", C.c_int),
                ("first_lat", C.c_double), ("first_lon", C.c_double), ("delta_lat", C.c_double),
                ("delta_lon", C.c_double), ("major", C.c_double), ("e2", C.c_double),
                ("length", C.c_int), ("width", C.c_int), ("nrnglooks", C.c_int), ("nazlooks", C.c_int),
                ("peg_heading", C.c_double), ("prf", C.c_double), ("t0", C.c_double), ("wvl", C.c_double),
                ("look_side", C.c_int), ("dem_method", C.c_int), ("orbit_method", C.c_int),
                ("line0", C.c_int), ("nlines", C.c_int), ("device", C.c_int)]


class TopoOutputs(C.Structure):
    _fields_ = [("lat", _dp), ("lon", _dp), ("hgt", _dp), ("los", _fp), ("inc", _fp), ("mask", C.POINTER(C.c_int8))]


class TopoResult(C.Structure):
    _fields_ = [("min_lat", C.c_double), ("max_lat", C.c_double), ("min_lon", C.c_double), ("max_lon", C.c_double),
                ("converged", C.c_longlong), ("iterations", C.c_longlong),
                ("dem_x0", C.c_int), ("dem_y0", C.c_int), ("dem_nx", C.c_int), ("dem_ny", C.c_int),
                ("dem_max", C.c_float), ("ms_setup", C.c_float), ("ms_kernels", C.c_float), ("ms_pixels", C.c_float),
                ("ms_solve", C.c_float), ("ms_mask", C.c_float), ("ms_total", C.c_float),
                ("gpu_launches", C.c_int)]


class GeoParams(C.Structure):
    _fields_ = [("major", C.c_double), ("e2", C.c_double), ("drho", C.c_double), ("rho0", C.c_double),
                ("wvl", C.c_double), ("t0", C.c_double), ("prf", C.c_double),
                ("length", C.c_int), ("width", C.c_int), ("look_side", C.c_int),
                ("nrnglooks", C.c_int), ("nazlooks", C.c_int), ("dem_width", C.c_int), ("dem_length", C.c_int),
                ("bistatic", C.c_int), ("orbit_method", C.c_int),
                ("line0", C.c_int), ("nlines", C.c_int), ("device", C.c_int), ("out_f32", C.c_int)]


class GeoOutputs(C.Structure):
    _fields_ = [("azt", C.c_void_p), ("rgm", C.c_void_p), ("azoff", C.c_void_p), ("rgoff", C.c_void_p)]


class GeoResult(C.Structure):
    _fields_ = [("num_outside", C.c_longlong), ("num_valid", C.c_longlong), ("num_converged", C.c_longlong),
                ("iterations", C.c_longlong), ("ms_setup", C.c_float), ("ms_kernels", C.c_float),
                ("ms_total", C.c_float), ("gpu_launches", C.c_int)]


class GeoJob(C.Structure):
    _fields_ = [("p", C.POINTER(GeoParams)), ("orbit", C.POINTER(Orbit)), ("dop", C.POINTER(Poly1d)),
                ("out", C.POINTER(GeoOutputs)), ("res", C.POINTER(GeoResult))]


class GeozeroParams(C.Structure):
    _fields_ = [("major", C.c_double), ("e2", C.c_double), ("min_lat", C.c_double), ("max_lat", C.c_double),
                ("min_lon", C.c_double), ("max_lon", C.c_double), ("drho", C.c_double), ("rho0", C.c_double),
                ("wvl", C.c_double), ("t0", C.c_double), ("prf", C.c_double), ("length", C.c_int), ("width", C.c_int),
                ("look_side", C.c_int), ("nrnglooks", C.c_int), ("nazlooks", C.c_int), ("first_lat", C.c_double),
                ("first_lon", C.c_double), ("delta_lat", C.c_double), ("delta_lon", C.c_double), ("dem_width", C.c_int),
                ("dem_length", C.c_int), ("device", C.c_int)]


class GeozeroResult(C.Structure):
    _fields_ = [("geo_width", C.c_int), ("geo_length", C.c_int), ("geo_min_lat", C.c_double), ("geo_max_lat", C.c_double),
                ("geo_min_lon", C.c_double), ("geo_max_lon", C.c_double), ("num_outside_dem", C.c_longlong),
                ("num_outside_image", C.c_longlong), ("num_valid", C.c_longlong), ("iterations", C.c_longlong),
                ("ms_setup", C.c_float), ("ms_solve", C.c_float), ("ms_kernels", C.c_float), ("ms_total", C.c_float),
                ("gpu_launches", C.c_int)]


class ResampParams(C.Structure):
    _fields_ = [("in_width", C.c_int), ("in_length", C.c_int), ("out_width", C.c_int), ("out_length", C.c_int),
                ("wvl", C.c_double), ("slr", C.c_double), ("r0", C.c_double), ("ref_wvl", C.c_double), ("ref_r0", C.c_double),
                ("ref_slr", C.c_double), ("flatten", C.c_int), ("device", C.c_int)]


class ResampResult(C.Structure):
    _fields_ = [("num_valid", C.c_longlong), ("ms_kernels", C.c_float), ("ms_total", C.c_float), ("gpu_launches", C.c_int)]


GEOZERO_METHODS = {"SINC": 0, "BILINEAR": 1, "BICUBIC": 2, "NEAREST": 3}
SCHEMES = {"BIL": 0, "BIP": 1, "BSQ": 2}

class LooksResult(C.Structure):
    _fields_ = [("out_length", C.c_int), ("out_width", C.c_int), ("ms_kernels", C.c_float), ("ms_total", C.c_float),
                ("gpu_launches", C.c_int)]


class MaskResult(C.Structure):
    _fields_ = [("ms_kernels", C.c_float), ("ms_total", C.c_float), ("gpu_launches", C.c_int)]


EXPORTS = ["b200_topo_run", "b200_topo_plan_create", "b200_topo_plan_execute", "b200_topo_plan_fetch",
           "b200_topo_plan_device_layers", "b200_topo_plan_destroy", "b200_geo2rdr_run", "b200_geo_plan_create",
           "b200_geo_plan_create_from_topo", "b200_geo_plan_execute", "b200_geo_plan_fetch", "b200_geo_plan_destroy",
           "b200_abi_version", "b200_release_cached_memory", "b200_device_count", "b200_device_name", "b200_alloc_pinned", "b200_free_pinned",
           "b200_fp64_peak", "b200_device_primitive", "b200_geozero_grid", "b200_geozero_plan_create",
           "b200_geozero_plan_geocode", "b200_geozero_plan_fetch", "b200_geozero_plan_destroy", "b200_geozero_run",
           "b200_resamp_slc_run", "b200_resamp_slc_from_geo_plan", "b200_topo_geo2rdr_run",
           "b200_looks_run", "b200_mask_to_radar_run", "b200_topo_plan_looks", "b200_geo_plan_freeze_geometry",
           "b200_d2h_floor", "b200_host_file_register", "b200_host_file_unregister", "b200_host_file_bytes",
           "b200_host_file_bytes_read"]

_lib = None


def lib():
    """Load libb200geom.so; raises (loudly) when the CUDA extension was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "or `make -C isce2_b200/csrc` (there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    err = [C.c_char_p, C.c_size_t]
    L.b200_topo_run.argtypes = [C.POINTER(TopoParams), C.c_void_p, C.c_int, C.POINTER(Orbit), C.POINTER(Poly2d),
                                C.POINTER(Poly2d), _dp, C.POINTER(TopoOutputs), C.POINTER(TopoResult)] + err
    L.b200_topo_geo2rdr_run.argtypes = [C.POINTER(TopoParams), C.c_void_p, C.c_int, C.POINTER(Orbit), C.POINTER(Poly2d),
                                        C.POINTER(Poly2d), _dp, C.POINTER(TopoOutputs), C.POINTER(TopoResult), C.c_int,
                                        C.POINTER(GeoJob)] + err
    L.b200_topo_plan_create.argtypes = [C.POINTER(TopoParams), C.c_void_p, C.c_int, C.POINTER(Orbit), C.POINTER(Poly2d),
                                        C.POINTER(Poly2d), _dp, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)] + err
    L.b200_topo_plan_execute.argtypes = [C.c_void_p, _fp] + err
    L.b200_topo_plan_fetch.argtypes = [C.c_void_p, C.POINTER(TopoOutputs), C.POINTER(TopoResult)] + err
    L.b200_topo_plan_device_layers.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
    L.b200_topo_plan_destroy.argtypes = [C.c_void_p]
    L.b200_topo_plan_destroy.restype = None
    L.b200_geo2rdr_run.argtypes = [C.POINTER(GeoParams), _dp, _dp, _dp, C.POINTER(Orbit), C.POINTER(Poly1d),
                                   C.POINTER(GeoOutputs), C.POINTER(GeoResult)] + err
    L.b200_geo_plan_create.argtypes = [C.POINTER(GeoParams), _dp, _dp, _dp, C.POINTER(C.c_void_p)] + err
    L.b200_geo_plan_create_from_topo.argtypes = [C.POINTER(GeoParams), C.c_void_p, C.POINTER(C.c_void_p)] + err
    L.b200_geo_plan_execute.argtypes = [C.c_void_p, C.POINTER(GeoParams), C.POINTER(Orbit), C.POINTER(Poly1d),
                                        C.c_int, C.c_int, C.c_int, C.c_int, _fp] + err
    L.b200_geo_plan_freeze_geometry.argtypes = [C.c_void_p, C.c_double, C.c_double] + err
    L.b200_geo_plan_fetch.argtypes = [C.c_void_p, C.POINTER(GeoOutputs), C.POINTER(GeoResult)] + err
    L.b200_geo_plan_destroy.argtypes = [C.c_void_p]
    L.b200_geo_plan_destroy.restype = None
    L.b200_release_cached_memory.restype = None
    L.b200_device_name.argtypes = [C.c_int, C.c_char_p, C.c_size_t]
    L.b200_alloc_pinned.argtypes = [C.c_size_t]
    L.b200_alloc_pinned.restype = C.c_void_p
    L.b200_free_pinned.argtypes = [C.c_void_p]
    L.b200_free_pinned.restype = None
    L.b200_fp64_peak.argtypes = [C.c_int, _dp] + err
    L.b200_d2h_floor.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_size_t, _fp] + err
    L.b200_host_file_register.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_longlong] + err
    L.b200_host_file_unregister.argtypes = [C.c_void_p]
    L.b200_host_file_bytes.restype = C.c_ulonglong
    L.b200_host_file_bytes.argtypes = []
    L.b200_host_file_bytes_read.restype = C.c_ulonglong
    L.b200_host_file_bytes_read.argtypes = []
    L.b200_device_primitive.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, C.POINTER(Orbit), _dp, _dp] + err
    L.b200_geozero_grid.argtypes = [C.POINTER(GeozeroParams), C.POINTER(C.c_int), C.POINTER(C.c_int)] + err
    L.b200_geozero_plan_create.argtypes = [C.POINTER(GeozeroParams), C.c_void_p, C.c_int, C.POINTER(Orbit), C.POINTER(Poly1d),
                                           C.POINTER(C.c_void_p)] + err
    L.b200_geozero_plan_geocode.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, _fp] + err
    L.b200_geozero_plan_fetch.argtypes = [C.c_void_p, C.POINTER(C.c_int16), _dp, _dp, C.POINTER(GeozeroResult)] + err
    L.b200_geozero_plan_destroy.argtypes = [C.c_void_p]
    L.b200_geozero_plan_destroy.restype = None
    L.b200_geozero_run.argtypes = [C.POINTER(GeozeroParams), C.c_void_p, C.c_int, C.POINTER(Orbit), C.POINTER(Poly1d), C.c_void_p,
                                   C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_int16),
                                   C.POINTER(GeozeroResult)] + err
    L.b200_resamp_slc_run.argtypes = [C.POINTER(ResampParams)] + [C.POINTER(Poly2d)] * 5 + [C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_int, C.c_void_p, C.POINTER(ResampResult)] + err
    L.b200_resamp_slc_from_geo_plan.argtypes = [C.POINTER(ResampParams), C.c_void_p] + [C.POINTER(Poly2d)] * 5 + [
        C.c_void_p, C.c_void_p, C.POINTER(ResampResult)] + err
    L.b200_looks_run.argtypes = [C.c_void_p, C.c_void_p] + [C.c_int] * 9 + [C.POINTER(LooksResult)] + err
    L.b200_topo_plan_looks.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(LooksResult)] + err
    L.b200_mask_to_radar_run.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double,
                                         C.c_void_p, C.c_void_p, C.c_int, C.c_size_t, C.c_void_p, C.c_int,
                                         C.POINTER(MaskResult)] + err
    _lib = L
    return L


def _check(rc, errbuf):
    if rc != B200_OK:
        raise B200Error(rc, errbuf.value.decode(errors="replace"))


def _errbuf():
    return C.create_string_buffer(512)


def device_count():
    return int(lib().b200_device_count())


def device_name(device=0):
    b = C.create_string_buffer(256)
    lib().b200_device_name(device, b, 256)
    return b.value.decode()


def fp64_peak(device=0):
    v = C.c_double()
    e = _errbuf()
    _check(lib().b200_fp64_peak(device, C.byref(v), e, 512), e)
    return v.value


def d2h_floor(host, nbytes=None, chunk_bytes=0, device=0):
    """Milliseconds (CUDA events) of a device -> host copy of nbytes into the page-locked array `host`: the floor of
    any call that has to deliver that many bytes of results over PCIe."""
    host = np.asarray(host)
    nbytes = host.nbytes if nbytes is None else int(nbytes)
    if nbytes > host.nbytes:
        raise ValueError("host buffer is smaller than the requested copy")
    ms = C.c_float()
    e = _errbuf()
    _check(lib().b200_d2h_floor(device, host.ctypes.data_as(C.c_void_p), nbytes, int(chunk_bytes), C.byref(ms), e, 512), e)
    return float(ms.value)


def host_file_register(address, nbytes, fd, file_offset):
    """Tell the library that host addresses [address, address + nbytes) are a shared writable mapping of the file open as
    `fd`, starting at file_offset: results bound for them are written with pwrite (include/b200geom.h)."""
    e = _errbuf()
    _check(lib().b200_host_file_register(C.c_void_p(address), int(nbytes), int(fd), int(file_offset), e, 512), e)


def host_file_bytes():
    """Bytes the copier threads have written with pwrite since the library was loaded."""
    return int(lib().b200_host_file_bytes())


def host_file_bytes_read():
    """Bytes the copier threads have read with pread since the library was loaded."""
    return int(lib().b200_host_file_bytes_read())


def host_file_unregister(address):
    return lib().b200_host_file_unregister(C.c_void_p(address)) == B200_OK


def pinned_empty(shape, dtype):
    """numpy array over page-locked host memory allocated by the library (freed with the array)."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    p = lib().b200_alloc_pinned(max(n, 1))
    if not p:
        raise MemoryError("b200_alloc_pinned failed")
    buf = (C.c_char * max(n, 1)).from_address(p)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    class _Owner:
        def __init__(self, ptr):
            self.ptr = ptr

        def __del__(self):
            try:
                lib().b200_free_pinned(self.ptr)
            except Exception:
                pass

    owner = _Owner(p)
    # keep the owner alive as long as any view of the buffer lives
    buf._b200_owner = owner
    return arr


class _Keep:
    """Holds numpy arrays alive while their pointers sit in ctypes structs."""

    def __init__(self):
        self.refs = []

    def d(self, a):
        a = np.ascontiguousarray(a, np.float64)
        self.refs.append(a)
        return a.ctypes.data_as(_dp)


def make_orbit(keep, t, pos, vel):
    t = np.ascontiguousarray(t, np.float64)
    return Orbit(len(t), keep.d(t), keep.d(np.asarray(pos).reshape(-1, 3)), keep.d(np.asarray(vel).reshape(-1, 3)))


def make_poly2d(keep, coeffs, mean_range=0.0, mean_azimuth=0.0, norm_range=1.0, norm_azimuth=1.0):
    c = np.atleast_2d(np.asarray(coeffs, np.float64))
    return Poly2d(c.shape[1] - 1, c.shape[0] - 1, mean_range, mean_azimuth, norm_range, norm_azimuth, keep.d(c))


def make_poly1d(keep, coeffs, mean=0.0, norm=1.0):
    c = np.asarray(coeffs, np.float64).ravel()
    return Poly1d(len(c) - 1, mean, norm, keep.d(c))


def _dem_arg(dem):
    if dem.dtype == np.float32:
        code = 0
    elif dem.dtype == np.int16:
        code = 1
    else:
        raise TypeError("DEM must be float32 or int16 (cast it first)")
    if not dem.flags["C_CONTIGUOUS"]:
        dem = np.ascontiguousarray(dem)
    return dem, code


def topo_params(*, dem_shape, first_lat, first_lon, delta_lat, delta_lon, length, width, prf, t0, wvl, side,
                peg_heading, a=6378137.0, e2=0.0066943799901, dem_method="BILINEAR", orbit_method="HERMITE",
                numiter=25, extraiter=10, thresh=0.05, nrnglooks=1, nazlooks=1, line0=0, nlines=-1, device=0):
    return TopoParams(numiter, extraiter, thresh, dem_shape[1], dem_shape[0], first_lat, first_lon, delta_lat, delta_lon,
                      a, e2, length, width, nrnglooks, nazlooks, peg_heading, prf, t0, wvl, side,
                      DEM_METHODS[dem_method.upper()], ORBIT_METHODS[orbit_method.upper()], line0, nlines, device)


def _result_dict(res):
    return {k: getattr(res, k) for k, _ in res._fields_}


class TopoPlan:
    """Device-resident topo (b200_topo_plan_*)."""

    def __init__(self, params, dem, orbit_t, orbit_pos, orbit_vel, doppler_coeffs, slrng_coeffs=None, rho_image=None,
                 want_los=True, want_inc=False, want_mask=False, doppler_poly=None, slrng_poly=None):
        L = lib()
        self.keep = _Keep()
        self.params = params
        dem, code = _dem_arg(dem)
        self.keep.refs.append(dem)
        orb = make_orbit(self.keep, orbit_t, orbit_pos, orbit_vel)
        dop = doppler_poly if doppler_poly is not None else make_poly2d(self.keep, doppler_coeffs)
        slr = slrng_poly if slrng_poly is not None else (make_poly2d(self.keep, slrng_coeffs) if slrng_coeffs is not None else None)
        rimg = self.keep.d(rho_image) if rho_image is not None else None
        self.handle = C.c_void_p()
        e = _errbuf()
        _check(L.b200_topo_plan_create(C.byref(params), dem.ctypes.data_as(C.c_void_p), code, C.byref(orb), C.byref(dop),
                                       C.byref(slr) if slr is not None else None, rimg, int(want_los), int(want_inc),
                                       int(want_mask), C.byref(self.handle), e, 512), e)
        self.want = (want_los, want_inc, want_mask)
        n = params.length - max(params.line0, 0) if params.nlines < 0 else params.nlines
        self.nlines = n
        self.width = params.width

    def execute(self):
        ms = C.c_float()
        e = _errbuf()
        _check(lib().b200_topo_plan_execute(self.handle, C.byref(ms), e, 512), e)
        return ms.value

    def fetch(self, out=None):
        n, w = self.nlines, self.width
        if out is None:
            out = dict(lat=np.empty((n, w)), lon=np.empty((n, w)), hgt=np.empty((n, w)),
                       los=np.empty((n, 2, w), np.float32) if self.want[0] else None,
                       inc=np.empty((n, 2, w), np.float32) if self.want[1] else None,
                       mask=np.empty((n, w), np.int8) if self.want[2] else None)
        o = TopoOutputs(out["lat"].ctypes.data_as(_dp), out["lon"].ctypes.data_as(_dp), out["hgt"].ctypes.data_as(_dp),
                        out["los"].ctypes.data_as(_fp) if out.get("los") is not None else None,
                        out["inc"].ctypes.data_as(_fp) if out.get("inc") is not None else None,
                        out["mask"].ctypes.data_as(C.POINTER(C.c_int8)) if out.get("mask") is not None else None)
        res = TopoResult()
        e = _errbuf()
        _check(lib().b200_topo_plan_fetch(self.handle, C.byref(o), C.byref(res), e, 512), e)
        out = dict(out)
        out.update(_result_dict(res))
        return out

    LAYERS = {"lat": (0, np.float64, 1), "lon": (1, np.float64, 1), "hgt": (2, np.float64, 1), "los": (3, np.float32, 2),
              "inc": (4, np.float32, 2), "mask": (5, np.int8, 1)}

    def looks(self, layer, down_looks, across_looks, method="AVERAGE"):
        """Multilooked copy of one resident layer (b200_topo_plan_looks).  Returns (array, result dict)."""
        code, dt, bands = self.LAYERS[layer]
        ol = self.nlines // down_looks if down_looks > 0 else 0
        ow = self.width // across_looks if across_looks > 0 else 0
        out = np.zeros((ol, ow) if bands == 1 else (ol, bands, ow), dt)
        res = LooksResult()
        e = _errbuf()
        _check(lib().b200_topo_plan_looks(self.handle, code, int(down_looks), int(across_looks), LOOKS_METHODS[method.upper()],
                                          out.ctypes.data_as(C.c_void_p), C.byref(res), e, 512), e)
        return out, _result_dict(res)

    def close(self):
        if self.handle:
            lib().b200_topo_plan_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def topo_run(params, dem, orbit_t, orbit_pos, orbit_vel, doppler_coeffs, slrng_coeffs=None, rho_image=None,
             want_los=True, want_inc=False, want_mask=False, out=None, doppler_poly=None, slrng_poly=None):
    """One-shot b200_topo_run with host buffers.  Returns dict of arrays + result fields."""
    L = lib()
    keep = _Keep()
    dem, code = _dem_arg(dem)
    orb = make_orbit(keep, orbit_t, orbit_pos, orbit_vel)
    dop = doppler_poly if doppler_poly is not None else make_poly2d(keep, doppler_coeffs)
    slr = slrng_poly if slrng_poly is not None else (make_poly2d(keep, slrng_coeffs) if slrng_coeffs is not None else None)
    rimg = keep.d(rho_image) if rho_image is not None else None
    n = params.length - max(params.line0, 0) if params.nlines < 0 else params.nlines
    w = params.width
    if out is None:
        out = dict(lat=np.empty((n, w)), lon=np.empty((n, w)), hgt=np.empty((n, w)),
                   los=np.empty((n, 2, w), np.float32) if want_los else None,
                   inc=np.empty((n, 2, w), np.float32) if want_inc else None,
                   mask=np.empty((n, w), np.int8) if want_mask else None)
    o = TopoOutputs(out["lat"].ctypes.data_as(_dp), out["lon"].ctypes.data_as(_dp), out["hgt"].ctypes.data_as(_dp),
                    out["los"].ctypes.data_as(_fp) if out.get("los") is not None else None,
                    out["inc"].ctypes.data_as(_fp) if out.get("inc") is not None else None,
                    out["mask"].ctypes.data_as(C.POINTER(C.c_int8)) if out.get("mask") is not None else None)
    res = TopoResult()
    e = _errbuf()
    _check(L.b200_topo_run(C.byref(params), dem.ctypes.data_as(C.c_void_p), code, C.byref(orb), C.byref(dop),
                           C.byref(slr) if slr is not None else None, rimg, C.byref(o), C.byref(res), e, 512), e)
    out = dict(out)
    out.update(_result_dict(res))
    return out


def geo_params(*, length, width, dem_shape, r0, dr, prf, t0, wvl, side=-1, a=6378137.0, e2=0.0066943799901,
               orbit_method="HERMITE", bistatic=False, nrnglooks=1, nazlooks=1, line0=0, nlines=-1, device=0,
               out_f32=False):
    return GeoParams(a, e2, dr, r0, wvl, t0, prf, length, width, side, nrnglooks, nazlooks, dem_shape[1], dem_shape[0],
                     int(bool(bistatic)), ORBIT_METHODS[orbit_method.upper()], line0, nlines, device, int(bool(out_f32)))


_GEO_KEYS = ("azt", "rgm", "azoff", "rgoff")


def _geo_out(params, want, out):
    n = params.dem_length - max(params.line0, 0) if params.nlines < 0 else params.nlines
    dt = np.float32 if params.out_f32 else np.float64
    if out is None:
        out = {k: (np.empty((n, params.dem_width), dt) if k in want else None) for k in _GEO_KEYS}
    o = GeoOutputs(*[(out[k].ctypes.data_as(C.c_void_p) if out.get(k) is not None else None) for k in _GEO_KEYS])
    return out, o


def geo2rdr_run(params, lat, lon, hgt, orbit_t, orbit_pos, orbit_vel, doppler_coeffs=(0.0,), doppler_mean=0.0,
                doppler_norm=1.0, want=_GEO_KEYS, out=None, block_rows=False):
    """b200_geo2rdr_run.  lat / lon / hgt are the whole [dem_length][dem_width] images; with block_rows=True they hold
    only the rows [line0, line0 + nlines) of the block (a rank of a sharded run that owns nothing else): the library
    reads no other row, so the image base it is given is the block's address moved back by line0 rows."""
    L = lib()
    keep = _Keep()
    orb = make_orbit(keep, orbit_t, orbit_pos, orbit_vel)
    dop = make_poly1d(keep, doppler_coeffs, doppler_mean, doppler_norm)
    out, o = _geo_out(params, want, out)
    res = GeoResult()
    e = _errbuf()
    ptrs = [keep.d(lat), keep.d(lon), keep.d(hgt)]
    if block_rows:
        n = params.dem_length - max(params.line0, 0) if params.nlines < 0 else params.nlines
        for a in (lat, lon, hgt):
            if a.shape != (n, params.dem_width):
                raise ValueError(f"block_rows: expected the block's {n} x {params.dem_width} rows, got {a.shape}")
        back = max(params.line0, 0) * params.dem_width * 8
        ptrs = [C.cast(C.c_void_p(C.cast(q, C.c_void_p).value - back), _dp) for q in ptrs]
    _check(L.b200_geo2rdr_run(C.byref(params), ptrs[0], ptrs[1], ptrs[2], C.byref(orb), C.byref(dop),
                              C.byref(o), C.byref(res), e, 512), e)
    out = dict(out)
    out.update(_result_dict(res))
    return out


def topo_geo2rdr_run(params, dem, orbit_t, orbit_pos, orbit_vel, doppler_coeffs, geo_jobs, slrng_coeffs=None, rho_image=None,
                     want_los=True, want_inc=False, want_mask=False, out=None, doppler_poly=None, slrng_poly=None):
    """Fused b200_topo_geo2rdr_run: topo over the block of `params`, then every job of `geo_jobs` on the lat / lon / hgt still
    resident in HBM.  Each job is a dict(params=GeoParams, orbit=(t, pos, vel), doppler=(coeffs, mean, norm),
    want=(...keys of azt/rgm/azoff/rgoff), out=None | dict of host arrays holding the block's rows).
    Returns (topo dict, [geo dicts])."""
    L = lib()
    keep = _Keep()
    dem, code = _dem_arg(dem)
    orb = make_orbit(keep, orbit_t, orbit_pos, orbit_vel)
    dop = doppler_poly if doppler_poly is not None else make_poly2d(keep, doppler_coeffs)
    slr = slrng_poly if slrng_poly is not None else (make_poly2d(keep, slrng_coeffs) if slrng_coeffs is not None else None)
    rimg = keep.d(rho_image) if rho_image is not None else None
    n = params.length - max(params.line0, 0) if params.nlines < 0 else params.nlines
    w = params.width
    if out is None:
        out = dict(lat=np.empty((n, w)), lon=np.empty((n, w)), hgt=np.empty((n, w)),
                   los=np.empty((n, 2, w), np.float32) if want_los else None,
                   inc=np.empty((n, 2, w), np.float32) if want_inc else None,
                   mask=np.empty((n, w), np.int8) if want_mask else None)
    o = TopoOutputs(out["lat"].ctypes.data_as(_dp), out["lon"].ctypes.data_as(_dp), out["hgt"].ctypes.data_as(_dp),
                    out["los"].ctypes.data_as(_fp) if out.get("los") is not None else None,
                    out["inc"].ctypes.data_as(_fp) if out.get("inc") is not None else None,
                    out["mask"].ctypes.data_as(C.POINTER(C.c_int8)) if out.get("mask") is not None else None)
    jobs = (GeoJob * max(len(geo_jobs), 1))()
    gouts, gres = [], []
    for j, jb in enumerate(geo_jobs):
        gp = jb["params"]
        # the block is the topo block: the output buffers are sized from it
        gp_blk = GeoParams.from_buffer_copy(gp)
        gp_blk.line0, gp_blk.nlines = max(params.line0, 0), n
        jo, jos = _geo_out(gp_blk, tuple(jb.get("want", _GEO_KEYS)), jb.get("out"))
        jorb = make_orbit(keep, *jb["orbit"])
        dc = jb.get("doppler", ((0.0,), 0.0, 1.0))
        jdop = make_poly1d(keep, *dc)
        r = GeoResult()
        keep.refs += [gp_blk, jos, jorb, jdop, r]
        jobs[j] = GeoJob(C.pointer(gp_blk), C.pointer(jorb), C.pointer(jdop), C.pointer(jos), C.pointer(r))
        gouts.append(jo)
        gres.append(r)
    res = TopoResult()
    e = _errbuf()
    _check(L.b200_topo_geo2rdr_run(C.byref(params), dem.ctypes.data_as(C.c_void_p), code, C.byref(orb), C.byref(dop),
                                   C.byref(slr) if slr is not None else None, rimg, C.byref(o), C.byref(res), len(geo_jobs), jobs,
                                   e, 512), e)
    out = dict(out)
    out.update(_result_dict(res))
    geos = []
    for jo, r in zip(gouts, gres):
        d = dict(jo)
        d.update(_result_dict(r))
        geos.append(d)
    return out, geos


class GeoPlan:
    """Device-resident geo2rdr: reference geometry uploaded once, many secondary orbits run against it."""

    def __init__(self, params, lat=None, lon=None, hgt=None, topo_plan=None):
        L = lib()
        self.keep = _Keep()
        self.params = params
        self.handle = C.c_void_p()
        e = _errbuf()
        if topo_plan is not None:
            self.topo_plan = topo_plan  # keep the layers alive
            _check(L.b200_geo_plan_create_from_topo(C.byref(params), topo_plan.handle, C.byref(self.handle), e, 512), e)
            self.nlines = topo_plan.nlines
        else:
            _check(L.b200_geo_plan_create(C.byref(params), self.keep.d(lat), self.keep.d(lon), self.keep.d(hgt),
                                          C.byref(self.handle), e, 512), e)
            self.nlines = params.dem_length - max(params.line0, 0) if params.nlines < 0 else params.nlines

    def execute(self, params, orbit_t, orbit_pos, orbit_vel, doppler_coeffs=(0.0,), doppler_mean=0.0, doppler_norm=1.0,
                want=_GEO_KEYS):
        keep = _Keep()
        orb = make_orbit(keep, orbit_t, orbit_pos, orbit_vel)
        dop = make_poly1d(keep, doppler_coeffs, doppler_mean, doppler_norm)
        ms = C.c_float()
        e = _errbuf()
        self.last_params = params
        self.last_want = tuple(want)
        _check(lib().b200_geo_plan_execute(self.handle, C.byref(params), C.byref(orb), C.byref(dop),
                                           *[int(k in want) for k in _GEO_KEYS], C.byref(ms), e, 512), e)
        return ms.value

    def freeze_geometry(self, a=6378137.0, e2=0.0066943799901):
        """Stack shape: convert the (now fixed) lat / lon / hgt to ECEF once for all secondary dates
        (b200_geo_plan_freeze_geometry)."""
        e = _errbuf()
        _check(lib().b200_geo_plan_freeze_geometry(self.handle, float(a), float(e2), e, 512), e)

    def fetch(self, out=None):
        p = self.last_params
        dt = np.float32 if p.out_f32 else np.float64
        if out is None:
            out = {k: (np.empty((self.nlines, p.dem_width), dt) if k in self.last_want else None) for k in _GEO_KEYS}
        o = GeoOutputs(*[(out[k].ctypes.data_as(C.c_void_p) if out.get(k) is not None else None) for k in _GEO_KEYS])
        res = GeoResult()
        e = _errbuf()
        _check(lib().b200_geo_plan_fetch(self.handle, C.byref(o), C.byref(res), e, 512), e)
        out = dict(out)
        out.update(_result_dict(res))
        return out

    def close(self):
        if self.handle:
            lib().b200_geo_plan_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def device_primitive(what, vec3_or_t, a=6378137.0, e2=0.0066943799901, orbit=None, device=0):
    keep = _Keep()
    inp = np.zeros(3)
    v = np.atleast_1d(np.asarray(vec3_or_t, np.float64))
    inp[: len(v)] = v
    out = np.zeros(7)
    orb = make_orbit(keep, *orbit) if orbit is not None else None
    e = _errbuf()
    _check(lib().b200_device_primitive(device, what, a, e2, C.byref(orb) if orb is not None else None,
                                       inp.ctypes.data_as(_dp), out.ctypes.data_as(_dp), e, 512), e)
    return out


# ---------------------------------------------------------------------------------------------------------------
# geozero
# ---------------------------------------------------------------------------------------------------------------
def geozero_params(*, dem_shape, first_lat, first_lon, delta_lat, delta_lon, snwe, length, width, r0, dr, prf, t0, wvl,
                   side=-1, a=6378137.0, e2=0.0066943799901, nrnglooks=1, nazlooks=1, device=0):
    """snwe = (min_lat, max_lat, min_lon, max_lon) in degrees, as Geocode.snwe."""
    return GeozeroParams(a, e2, snwe[0], snwe[1], snwe[2], snwe[3], dr, r0, wvl, t0, prf, length, width, side, nrnglooks,
                         nazlooks, first_lat, first_lon, delta_lat, delta_lon, dem_shape[1], dem_shape[0], device)


def geozero_grid(params, require_nonempty=False):
    w, l = C.c_int(), C.c_int()
    e = _errbuf()
    _check(lib().b200_geozero_grid(C.byref(params), C.byref(w), C.byref(l), e, 512), e)
    if require_nonempty and (l.value < 1 or w.value < 1):
        raise B200Error(-1, f"empty output grid ({l.value} lines x {w.value} samples): check the bounding box")
    return l.value, w.value


def _image_arg(image, width, length, nbands, scheme):
    """(contiguous array, is_complex): float32 or complex64 samples in the file's own interleaving."""
    a = np.asarray(image)
    is_complex = np.iscomplexobj(a)
    a = np.ascontiguousarray(a, np.complex64 if is_complex else np.float32)
    if a.size != width * length * nbands:
        raise ValueError(f"image has {a.size} samples, expected {nbands} x {length} x {width}")
    return a, is_complex


class GeozeroPlan:
    """Geometry of one geocoding grid, solved once on the GPU; geocode() any number of images / bands with it."""

    def __init__(self, params, dem, orbit_t, orbit_pos, orbit_vel, doppler_coeffs=(0.0,), doppler_mean=0.0, doppler_norm=1.0):
        self.params = params
        keep = _Keep()
        dem, code = _dem_arg(dem)
        orb = make_orbit(keep, orbit_t, orbit_pos, orbit_vel)
        dop = make_poly1d(keep, doppler_coeffs, doppler_mean, doppler_norm)
        self.handle = C.c_void_p()
        e = _errbuf()
        _check(lib().b200_geozero_plan_create(C.byref(params), dem.ctypes.data_as(C.c_void_p), code, C.byref(orb), C.byref(dop),
                                              C.byref(self.handle), e, 512), e)
        self.geo_length, self.geo_width = geozero_grid(params)

    def geocode(self, image, method="BILINEAR", nbands=1, scheme="BIL", out=None):
        p = self.params
        a, is_complex = _image_arg(image, p.width, p.length, nbands, scheme)
        sch = scheme.upper()
        shape = ((self.geo_length, self.geo_width) if nbands == 1 else
                 (self.geo_length, nbands, self.geo_width) if sch == "BIL" else
                 (self.geo_length, self.geo_width, nbands) if sch == "BIP" else (nbands, self.geo_length, self.geo_width))
        if out is None:
            out = np.empty(shape, a.dtype)
        if out.dtype != a.dtype or out.size != int(np.prod(shape)) or not out.flags["C_CONTIGUOUS"]:
            raise ValueError("out must be a C-contiguous array of the image's sample type on the geocoded grid")
        ms = C.c_float()
        e = _errbuf()
        _check(lib().b200_geozero_plan_geocode(self.handle, a.ctypes.data_as(C.c_void_p), int(is_complex), int(nbands),
                                               SCHEMES[sch], GEOZERO_METHODS[method.upper()], out.ctypes.data_as(C.c_void_p),
                                               C.byref(ms), e, 512), e)
        self.ms_kernels = ms.value
        return out

    def fetch(self, want_indices=False, dem_crop=None):
        n = (self.geo_length, self.geo_width)
        crop = dem_crop if dem_crop is not None else np.empty(n, np.int16)
        az = np.empty(n) if want_indices else None
        rg = np.empty(n) if want_indices else None
        res = GeozeroResult()
        e = _errbuf()
        _check(lib().b200_geozero_plan_fetch(self.handle, crop.ctypes.data_as(C.POINTER(C.c_int16)),
                                             az.ctypes.data_as(_dp) if az is not None else None,
                                             rg.ctypes.data_as(_dp) if rg is not None else None, C.byref(res), e, 512), e)
        r = dict(dem_crop=crop, az_idx=az, rng_idx=rg)
        r.update(_result_dict(res))
        return r

    def close(self):
        if self.handle:
            lib().b200_geozero_plan_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def geozero_run(params, dem, orbit_t, orbit_pos, orbit_vel, image, method="BILINEAR", nbands=1, scheme="BIL",
                doppler_coeffs=(0.0,), doppler_mean=0.0, doppler_norm=1.0, out=None, dem_crop=None):
    """One call of the reference verb geozero_Py, all bands of one image."""
    keep = _Keep()
    dem, code = _dem_arg(dem)
    orb = make_orbit(keep, orbit_t, orbit_pos, orbit_vel)
    dop = make_poly1d(keep, doppler_coeffs, doppler_mean, doppler_norm)
    a, is_complex = _image_arg(image, params.width, params.length, nbands, scheme)
    gl, gw = geozero_grid(params, require_nonempty=True)
    sch = scheme.upper()
    shape = ((gl, gw) if nbands == 1 else (gl, nbands, gw) if sch == "BIL" else (gl, gw, nbands) if sch == "BIP" else (nbands, gl, gw))
    if out is None:
        out = np.empty(shape, a.dtype)
    if dem_crop is None:
        dem_crop = np.empty((gl, gw), np.int16)
    res = GeozeroResult()
    e = _errbuf()
    _check(lib().b200_geozero_run(C.byref(params), dem.ctypes.data_as(C.c_void_p), code, C.byref(orb), C.byref(dop),
                                  a.ctypes.data_as(C.c_void_p), int(is_complex), int(nbands), SCHEMES[sch],
                                  GEOZERO_METHODS[method.upper()], out.ctypes.data_as(C.c_void_p),
                                  dem_crop.ctypes.data_as(C.POINTER(C.c_int16)), C.byref(res), e, 512), e)
    r = dict(geo=out, dem_crop=dem_crop)
    r.update(_result_dict(res))
    return r


# ---------------------------------------------------------------------------------------------------------------
# resamp_slc
# ---------------------------------------------------------------------------------------------------------------
def _poly2d_arg(keep, p):
    """None, a (coeffs, mean_range, mean_azimuth, norm_range, norm_azimuth) tuple as poly.poly2d_fields returns, or a
    plain 2-D coefficient list."""
    if p is None:
        return None
    if isinstance(p, tuple) and len(p) == 5:
        return make_poly2d(keep, *p)
    return make_poly2d(keep, p)


def resamp_slc_run(slc, out_shape, *, wvl=0.056, slr=2.3, r0=0.0, ref_wvl=None, ref_r0=None, ref_slr=None, flatten=False,
                   rg_carrier=None, az_carrier=None, rg_offsets=None, az_offsets=None, doppler=None, resid_az=None,
                   resid_rg=None, out=None, device=0):
    """One call of the reference verb resamp_slc_Py.  slc: complex64 [in_length][in_width]; resid_*: float64 or float32
    [out_length][out_width] (both of one type) or None; returns dict(slc=..., num_valid=..., ms_kernels=...)."""
    keep = _Keep()
    a = np.ascontiguousarray(slc, np.complex64)
    if a.ndim != 2:
        raise ValueError("slc must be a single-band image")
    ol, ow = int(out_shape[0]), int(out_shape[1])
    p = ResampParams(a.shape[1], a.shape[0], ow, ol, wvl, slr, r0, wvl if ref_wvl is None else ref_wvl,
                     r0 if ref_r0 is None else ref_r0, slr if ref_slr is None else ref_slr, int(bool(flatten)), device)
    polys = [_poly2d_arg(keep, q) for q in (rg_carrier, az_carrier, rg_offsets, az_offsets, doppler)]
    rdt = None
    res_ptr = []
    for r in (resid_az, resid_rg):
        if r is None:
            res_ptr.append(None)
            continue
        r = np.asarray(r)
        dt = np.float32 if r.dtype == np.float32 else np.float64
        if rdt is not None and dt != rdt:
            raise ValueError("residual azimuth and range offsets must have one data type")
        rdt = dt
        r = np.ascontiguousarray(r, dt)
        if r.shape != (ol, ow):
            raise ValueError(f"residual offsets have shape {r.shape}, expected {(ol, ow)}")
        keep.refs.append(r)
        res_ptr.append(r.ctypes.data_as(C.c_void_p))
    if out is None:
        out = np.empty((ol, ow), np.complex64)
    if out.dtype != np.complex64 or out.shape != (ol, ow) or not out.flags["C_CONTIGUOUS"]:
        raise ValueError("out must be a C-contiguous complex64 array of the output size")
    res = ResampResult()
    e = _errbuf()
    _check(lib().b200_resamp_slc_run(C.byref(p), *[(C.byref(q) if q is not None else None) for q in polys],
                                     a.ctypes.data_as(C.c_void_p), res_ptr[0], res_ptr[1], 1 if rdt == np.float32 else 0,
                                     out.ctypes.data_as(C.c_void_p), C.byref(res), e, 512), e)
    r = dict(slc=out)
    r.update(_result_dict(res))
    return r


def resamp_slc_from_geo_plan(geo_plan, slc, *, wvl=0.056, slr=2.3, r0=0.0, ref_wvl=None, ref_r0=None, ref_slr=None,
                             flatten=False, rg_carrier=None, az_carrier=None, rg_offsets=None, az_offsets=None, doppler=None,
                             out=None):
    """geo2rdr -> resamp_slc without leaving the GPU: the offsets of the executed GeoPlan are the residual images."""
    keep = _Keep()
    a = np.ascontiguousarray(slc, np.complex64)
    gp = geo_plan.last_params
    ol, ow = geo_plan.nlines, gp.dem_width
    p = ResampParams(a.shape[1], a.shape[0], ow, ol, wvl, slr, r0, wvl if ref_wvl is None else ref_wvl,
                     r0 if ref_r0 is None else ref_r0, slr if ref_slr is None else ref_slr, int(bool(flatten)), gp.device)
    polys = [_poly2d_arg(keep, q) for q in (rg_carrier, az_carrier, rg_offsets, az_offsets, doppler)]
    if out is None:
        out = np.empty((ol, ow), np.complex64)
    res = ResampResult()
    e = _errbuf()
    _check(lib().b200_resamp_slc_from_geo_plan(C.byref(p), geo_plan.handle, *[(C.byref(q) if q is not None else None) for q in polys],
                                               a.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), C.byref(res), e, 512), e)
    r = dict(slc=out)
    r.update(_result_dict(res))
    return r


# ---- multilooking of the geometry layers, mask projection (other consumers) ----
# numpy dtype -> element type of looks_C (looksmodule.cpp:80-127)
LOOKS_TYPES = {np.dtype(np.int8): 0, np.dtype(np.int16): 1, np.dtype(np.int32): 2, np.dtype(np.int64): 3,
               np.dtype(np.float32): 4, np.dtype(np.float64): 5, np.dtype(np.complex64): 6}
SCHEMES = {"BIL": 0, "BIP": 1, "BSQ": 2}
LOOKS_METHODS = {"AVERAGE": 0, "ISCE": 0, "NEAREST": 1, "GDAL": 1}


def _band_axes(scheme):
    # which quantity (0 line, 1 band, 2 sample) runs along each axis of an array stored in the given interleaving
    return {"BIL": (0, 1, 2), "BIP": (0, 2, 1), "BSQ": (1, 0, 2)}[scheme]


def looks_run(image, down_looks, across_looks, *, scheme="BIL", method="AVERAGE", out=None, device=0):
    """b200_looks_run.  `image`: 2-D [length][width] or 3-D in the storage order of `scheme` (BIL [line][band][sample],
    BIP [line][sample][band], BSQ [band][line][sample]).  Returns (out array in the same layout, result dict)."""
    a = np.ascontiguousarray(image)
    if a.dtype == np.uint8:
        a = a.view(np.int8)  # BYTE is 'i1' in ISCE (Image.py:62)
    if a.dtype not in LOOKS_TYPES:
        raise TypeError(f"Error. Unrecognized data type {a.dtype}")
    scheme = scheme.upper()
    if a.ndim == 2:
        length, width, bands = a.shape[0], a.shape[1], 1
        oshape = (length // down_looks if down_looks > 0 else 0, width // across_looks if across_looks > 0 else 0)
    else:
        ax = _band_axes(scheme)
        length, bands, width = a.shape[ax.index(0)], a.shape[ax.index(1)], a.shape[ax.index(2)]
        dims = {0: length // down_looks if down_looks > 0 else 0, 1: bands, 2: width // across_looks if across_looks > 0 else 0}
        oshape = tuple(dims[k] for k in ax)
    if out is None:
        out = np.zeros(oshape, a.dtype)
    elif out.shape != oshape or out.dtype != a.dtype or not out.flags["C_CONTIGUOUS"]:
        raise ValueError(f"out must be a C-contiguous {a.dtype} array of shape {oshape}")
    res = LooksResult()
    e = _errbuf()
    _check(lib().b200_looks_run(a.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), LOOKS_TYPES[a.dtype], length, width,
                                bands, SCHEMES[scheme], int(down_looks), int(across_looks), LOOKS_METHODS[method.upper()],
                                int(device), C.byref(res), e, 512), e)
    return out, _result_dict(res)


def mask_to_radar_run(mask, start_lat, delta_lat, start_lon, delta_lon, lat, lon, *, out=None, device=0):
    """b200_mask_to_radar_run: `mask` [mask_length][mask_width] (int8 / int16 / int32 / float32) sampled at the radar pixels
    whose latitude / longitude are `lat` / `lon` (float64 or float32, any shape)."""
    m = np.ascontiguousarray(mask)
    if m.dtype == np.uint8:
        m = m.view(np.int8)
    codes = {np.dtype(np.int8): 0, np.dtype(np.int16): 1, np.dtype(np.int32): 2, np.dtype(np.float32): 4}
    if m.dtype not in codes or m.ndim != 2:
        raise TypeError("mask must be a 2-D int8 / int16 / int32 / float32 array")
    lat = np.ascontiguousarray(lat)
    lon = np.ascontiguousarray(lon)
    if lat.dtype != lon.dtype or lat.dtype not in (np.float32, np.float64) or lat.shape != lon.shape:
        raise TypeError("lat / lon must be float32 or float64 arrays of one shape")
    if out is None:
        out = np.zeros(lat.shape, m.dtype)
    elif out.shape != lat.shape or out.dtype != m.dtype or not out.flags["C_CONTIGUOUS"]:
        raise ValueError("out must be a C-contiguous array of the mask's dtype and lat's shape")
    res = MaskResult()
    e = _errbuf()
    _check(lib().b200_mask_to_radar_run(m.ctypes.data_as(C.c_void_p), codes[m.dtype], m.shape[0], m.shape[1], float(start_lat),
                                        float(delta_lat), float(start_lon), float(delta_lon), lat.ctypes.data_as(C.c_void_p),
                                        lon.ctypes.data_as(C.c_void_p), int(lat.dtype == np.float32), lat.size,
                                        out.ctypes.data_as(C.c_void_p), int(device), C.byref(res), e, 512), e)
    return out, _result_dict(res)
