"""TEST-ONLY stand-in for ``isce2_b200._capi`` that lets tests/test_multirank_cpu.py drive the whole control flow of
``bench.run_b200`` (sharding, barriers, reductions, the two end-to-end arms and their consistency report, the JSON line) with
world_size 2 on a machine without GPUs.  It computes nothing real: layers and offsets are cheap closed-form functions of
the pixel index, written so that -- like on the device -- a block described with a re-based sensing start differs from
the absolute description in the last bits only.  Never imported by the product or by bench.py."""
import numpy as np

from isce2_b200 import _capi as real

TopoResult, GeoResult, topo_params, geo_params, _check = real.TopoResult, real.GeoResult, real.topo_params, real.geo_params, real._check
B200Error = real.B200Error


def device_count(): return 1
def device_name(device=0): return "fake device (CPU test of the bench control flow)"
def fp64_peak(device=0): return 33.9
def pinned_empty(shape, dtype): return np.empty(shape, dtype)
def d2h_floor(host, nbytes=None, chunk_bytes=0, device=0): return 1e-3 * np.asarray(host).nbytes / 50e6


def _block(p):
    n = p.length - max(p.line0, 0) if p.nlines < 0 else p.nlines
    return max(p.line0, 0), n


def _fill_topo(p, out):
    l0, n = _block(p)
    y = (l0 + np.arange(n, dtype=np.float64))[:, None]
    x = np.arange(p.width, dtype=np.float64)[None, :]
    out["lat"][...] = 35.0 + 1e-5 * y + 1e-7 * x
    out["lon"][...] = -118.0 + 2e-5 * x
    out["hgt"][...] = 100.0 + 0.01 * x + 0.02 * y
    for k in ("los", "inc", "mask"):
        if out.get(k) is not None:
            out[k][...] = 0
    return dict(min_lat=35.0, max_lat=35.1, min_lon=-118.0, max_lon=-117.5, converged=n * p.width, iterations=6 * n * p.width,
                dem_x0=1, dem_y0=1, dem_nx=64, dem_ny=64, dem_max=1000.0, ms_setup=0.1, ms_kernels=1.0, ms_pixels=0.8, ms_solve=0.5,
                ms_mask=0.2, ms_total=1.5, gpu_launches=4)


def _fill_geo(p, lat, out):
    # "azimuth time" from the latitude; the offset is formed as the device forms it: (t - t0) * prf - line
    l0, n = _block(p) if p.nlines >= 0 else (max(p.line0, 0), p.dem_length - max(p.line0, 0))
    t = 21600.0 + (np.asarray(lat) - 35.0) / 1e-5 / 486.486
    line = (l0 + np.arange(lat.shape[0], dtype=np.float64))[:, None]
    azoff = (t - p.t0) * p.prf - line
    rgoff = 1e-3 * np.arange(p.dem_width, dtype=np.float64)[None, :] + 0 * line
    bad = np.zeros(azoff.shape, bool)
    bad[:, :3] = True
    dt = np.float32 if p.out_f32 else np.float64
    for k, v in (("azoff", azoff), ("rgoff", rgoff)):
        if out.get(k) is not None:
            out[k][...] = np.where(bad, -999999.0, v).astype(dt)
    npx = azoff.size
    return dict(num_outside=int(bad.sum()), num_valid=int(npx - bad.sum()), num_converged=int(npx - bad.sum()), iterations=3 * npx,
                ms_setup=0.1, ms_kernels=0.3, ms_total=0.5, gpu_launches=2)


class _Lib:
    def b200_topo_plan_fetch(self, handle, out, res_ref, e, n):
        for k, v in handle.value_dict.items():
            setattr(res_ref._obj, k, v)
        return 0

    def b200_geo_plan_fetch(self, handle, out, res_ref, e, n):
        for k, v in handle.value_dict.items():
            setattr(res_ref._obj, k, v)
        return 0


def lib(): return _Lib()


class _Handle:
    def __init__(self): self.value_dict = {}


class TopoPlan:
    def __init__(self, params, dem, *a, **kw):
        self.params, self.handle = params, _Handle()
        l0, n = _block(params)
        self.nlines, self.width = n, params.width
        self.layers = dict(lat=np.empty((n, params.width)), lon=np.empty((n, params.width)), hgt=np.empty((n, params.width)))

    def execute(self):
        self.handle.value_dict = _fill_topo(self.params, self.layers)
        return 1.0

    def close(self): pass


class GeoPlan:
    def __init__(self, params, lat=None, lon=None, hgt=None, topo_plan=None):
        self.handle, self.topo = _Handle(), topo_plan

    def freeze_geometry(self, *a): pass

    def execute(self, params, t, pos, vel, want=("azoff", "rgoff"), **kw):
        n, w = self.topo.layers["lat"].shape
        out = dict(azoff=np.empty((n, w), np.float32), rgoff=np.empty((n, w), np.float32))
        q = real.GeoParams.from_buffer_copy(params)
        q.line0, q.nlines = max(self.topo.params.line0, 0), n
        self.handle.value_dict = _fill_geo(q, self.topo.layers["lat"], out)
        return 0.3

    def close(self): pass


def topo_run(params, dem, t, pos, vel, dop, slr=None, out=None, **kw):
    r = _fill_topo(params, out)
    d = dict(out)
    d.update(r)
    return d


def geo2rdr_run(params, lat, lon, hgt, t, pos, vel, *a, want=("azoff", "rgoff"), out=None, **kw):
    q = real.GeoParams.from_buffer_copy(params)
    q.line0, q.nlines = max(params.line0, 0), lat.shape[0]
    r = _fill_geo(q, lat, out)
    d = dict(out)
    d.update(r)
    return d


def topo_geo2rdr_run(params, dem, t, pos, vel, dop, geo_jobs, slr=None, out=None, **kw):
    tr = topo_run(params, dem, t, pos, vel, dop, slr, out=out)
    geos = []
    l0, n = _block(params)
    for jb in geo_jobs:
        q = real.GeoParams.from_buffer_copy(jb["params"])
        q.line0, q.nlines = l0, n
        r = _fill_geo(q, out["lat"], jb["out"])
        d = dict(jb["out"])
        d.update(r)
        geos.append(d)
    return tr, geos
