"""Callers on either side of the hot path:

* ``run_topo_merged``  -- what ``runTopoGPU`` of components/isceobj/TopsProc/runTopo.py:114-359 does: ONE topo over the
  union grid of all swaths / bursts of a TOPS acquisition, then per-burst ``.vrt`` windows into the merged layers
  (``buildVRT``, :362-423), so the per-burst launch and I/O overhead disappears.
* ``Geo2rdrStack``     -- the topsStack shape (contrib/stack/topsStack/geo2rdr.py:233-302, Stack.py:805-827: one
  ``geo2rdr.py`` process per secondary date, one call per burst): each burst's reference geometry is uploaded to
  a GPU once and every secondary date runs against the resident copy.

Both sit on the public Component / C-ABI surface of this package (``createTopozero``, ``_capi.GeoPlan``); frames,
bursts and orbits are duck-typed exactly like the objects the reference passes around (``burst.sensingStart``,
``burst.startingRange``, ``burst.numberOfLines``, ``frame.farRange`` ...).
"""
from __future__ import annotations

import datetime
import math
import os
import threading
from dataclasses import dataclass

import numpy as np

from . import _capi, image as IF
from .orbit import Orbit, export_rows, seconds_since_midnight, state_vectors
from .planet import EarthEccentricitySquared, EarthMajorSemiAxis
from .poly import Poly2D


# ---------------------------------------------------------------------------------------------------------------
# union grid of the swaths (runTopo.py:159-172) and burst windows (:316-319)
# ---------------------------------------------------------------------------------------------------------------
@dataclass
class SwathGrid:
    r0: float                    # near range of the left-most swath [m]
    dr: float                    # range pixel size [m]
    t0: datetime.datetime        # sensing start of the top-most swath
    dt: float                    # azimuth time interval [s]
    wvl: float
    width: int
    length: int

    def window(self, burst):
        """(top, bottom, left, right) of a burst inside the merged grid (runTopo.py:316-319)."""
        top = int(np.rint((burst.sensingStart - self.t0).total_seconds() / self.dt))
        bottom = top + burst.numberOfLines
        left = int(np.rint((burst.startingRange - self.r0) / self.dr))
        right = left + burst.numberOfSamples
        return top, bottom, left, right


def _frame_attr(frame, name):
    """TOPSSwathSLCProduct exposes sensingStart / sensingStop / startingRange / farRange as properties over its
    bursts (components/isceobj/Sensor/TOPS/TOPSSwathSLCProduct.py); plain containers of bursts get the same values."""
    if hasattr(frame, name):
        return getattr(frame, name)
    b = frame.bursts
    if name == "sensingStart":
        return min(x.sensingStart for x in b)
    if name == "sensingStop":
        return max(x.sensingStop for x in b)
    if name == "startingRange":
        return min(x.startingRange for x in b)
    if name == "farRange":
        return max(x.farRange for x in b)
    raise AttributeError(name)


def union_grid(frames):
    top = min(frames, key=lambda x: _frame_attr(x, "sensingStart"))
    left = min(frames, key=lambda x: _frame_attr(x, "startingRange"))
    bottom = max(frames, key=lambda x: _frame_attr(x, "sensingStop"))
    right = max(frames, key=lambda x: _frame_attr(x, "farRange"))
    b0 = frames[0].bursts[0]
    r0 = _frame_attr(left, "startingRange")
    rmax = _frame_attr(right, "farRange")
    dr = b0.rangePixelSize
    t0 = _frame_attr(top, "sensingStart")
    tmax = _frame_attr(bottom, "sensingStop")
    dt = b0.azimuthTimeInterval
    width = int(np.round((rmax - r0) / dr) + 1)
    length = int(np.round((tmax - t0).total_seconds() / dt) + 1)
    return SwathGrid(r0=r0, dr=dr, t0=t0, dt=dt, wvl=b0.radarWavelength, width=width, length=length)


def merged_orbit(frames):
    """TopsProc.getMergedOrbit (components/isceobj/TopsProc/TopsProc.py:488-510): the first burst's state vectors plus
    every state vector of the other bursts that falls outside the span collected so far."""
    orb = Orbit()
    for sv in state_vectors(frames[0].bursts[0].orbit):
        orb.addStateVector(sv)
    for pp in frames:
        for bb in pp.bursts:
            for sv in state_vectors(bb.orbit):
                if sv.getTime() < orb.minTime or sv.getTime() > orb.maxTime:
                    orb.addStateVector(sv)
    return orb


# ---------------------------------------------------------------------------------------------------------------
# burst views (runTopo.py:362-423)
# ---------------------------------------------------------------------------------------------------------------
_VRT_HEADER = '<VRTDataset rasterXSize="{width}" rasterYSize="{lgth}">'
_VRT_BAND = '''    <VRTRasterBand dataType="{dtype}" band="{band}">
        <NoDataValue>0.0</NoDataValue>
        <SimpleSource>
            <SourceFilename relativeToVRT="1">{relpath}</SourceFilename>
            <SourceBand>{band}</SourceBand>
            <SourceProperties RasterXSize="{gwidth}" RasterYSize="{glgth}" DataType="{dtype}"/>
            <SrcRect xOff="{left}" yOff="{top}" xSize="{width}" ySize="{lgth}"/>
            <DstRect xOff="0" yOff="0" xSize="{width}" ySize="{lgth}"/>
        </SimpleSource>
    </VRTRasterBand>
'''
_VRT_TAIL = "</VRTDataset>"


def build_vrt(srcname, dstname, dims, bbox, bands=1, dtype="FLOAT"):
    """Write ``dstname.xml`` + ``dstname.vrt`` describing the window bbox = [top, bottom, left, right] of the merged
    raster ``srcname`` (dims = [width, length] of the parent); no pixel is copied."""
    width = bbox[3] - bbox[2]
    lgth = bbox[1] - bbox[0]
    odtype = dtype
    try:
        gdt = {"FLOAT": "Float32", "DOUBLE": "Float64", "BYTE": "UInt8"}[dtype.upper()]
    except KeyError:
        raise Exception("Unsupported type {0}".format(dtype))
    relpath = os.path.relpath(srcname + ".vrt", os.path.dirname(dstname))
    img = IF.createImage()
    img.bands = bands
    img.scheme = "BIL"
    img.setWidth(width)
    img.setLength(lgth)
    img.dataType = odtype
    img.filename = dstname
    img.setAccessMode("READ")
    img.renderHdr()
    with open(dstname + ".vrt", "w") as fid:
        fid.write(_VRT_HEADER.format(width=width, lgth=lgth) + "\n")
        for bnd in range(bands):
            fid.write(_VRT_BAND.format(width=width, lgth=lgth, gwidth=dims[0], glgth=dims[1], left=bbox[2], top=bbox[0],
                                       relpath=relpath, dtype=gdt, band=bnd + 1))
        fid.write(_VRT_TAIL + "\n")


# ---------------------------------------------------------------------------------------------------------------
# merged topo
# ---------------------------------------------------------------------------------------------------------------
def run_topo_merged(frames, dem_image, dirname, *, swaths=None, swath_starts=None, orbit=None, look_side=-1,
                    dem_method="BIQUINTIC", orbit_method="HERMITE", inc=False, mask=False, devices=None):
    """One topo over the union grid of `frames` (each with ``.bursts``), layers ``lat/lon/hgt/los.rdr`` (+ ``incLocal.rdr``,
    ``shadowMask.rdr`` on request) under `dirname`, and ``IW<n>/lat_%02d.rdr`` ... views for every burst.

    Settings are those of runTopoGPU (runTopo.py:224-262): zero Doppler, slant range r0 + dr*pixel, 25 + 10 iterations,
    threshold 0.05 m, peg heading = ENU heading of the merged orbit at the first line, BIQUINTIC DEM interpolation,
    HERMITE orbit.  Returns a dict with the grid, the Topo object (bounding box, timings) and the burst windows."""
    from .topozero import createTopozero

    swaths = list(swaths) if swaths is not None else [i + 1 for i in range(len(frames))]
    swath_starts = list(swath_starts) if swath_starts is not None else [0] * len(frames)
    if len(frames) == 0:
        raise Exception("There is no common region between the two dates to process")
    g = union_grid(frames)
    orb = orbit if orbit is not None else merged_orbit(frames)
    os.makedirs(dirname, exist_ok=True)

    poly = Poly2D(name="topsApp_dopplerPoly")
    poly.setWidth(g.width)
    poly.setLength(g.length)
    poly.setNormRange(1.0)
    poly.setNormAzimuth(1.0)
    poly.setMeanRange(0.0)
    poly.setMeanAzimuth(0.0)
    poly.initPoly(rangeOrder=0, azimuthOrder=0, coeffs=[[0.0]])

    topo = createTopozero()
    topo.slantRangePixelSpacing = g.dr
    topo.prf = 1.0 / g.dt
    topo.radarWavelength = g.wvl
    topo.orbit = orb
    topo.width = g.width
    topo.length = g.length
    topo.lookSide = look_side
    topo.sensingStart = g.t0
    topo.rangeFirstSample = g.r0
    topo.numberRangeLooks = 1
    topo.numberAzimuthLooks = 1
    topo.polyDoppler = poly
    topo.demInterpolationMethod = dem_method
    topo.orbitInterpolationMethod = orbit_method
    topo.numberIterations = 25
    topo.secondaryIterations = 10
    topo.threshold = 0.05
    topo.pegHeading = math.radians(orb.getENUHeading(g.t0))
    topo.ellipsoidMajorSemiAxis = EarthMajorSemiAxis
    topo.ellipsoidEccentricitySquared = EarthEccentricitySquared
    topo.latFilename = os.path.join(dirname, "lat.rdr")
    topo.lonFilename = os.path.join(dirname, "lon.rdr")
    topo.heightFilename = os.path.join(dirname, "hgt.rdr")
    topo.losFilename = os.path.join(dirname, "los.rdr")
    if inc:
        topo.incFilename = os.path.join(dirname, "incLocal.rdr")
    if mask:
        topo.maskFilename = os.path.join(dirname, "shadowMask.rdr")
    if devices is not None:
        topo.gpuDevices = list(devices)
    topo.topo(demImage=dem_image)

    layers = [("lat", 1, "DOUBLE"), ("lon", 1, "DOUBLE"), ("hgt", 1, "DOUBLE"), ("los", 2, "FLOAT")]
    if inc:
        layers.append(("incLocal", 2, "FLOAT"))
    if mask:
        layers.append(("shadowMask", 1, "BYTE"))
    windows = {}
    for swath, frame, istart in zip(swaths, frames, swath_starts):
        outname = os.path.join(dirname, "IW{0}".format(swath))
        os.makedirs(outname, exist_ok=True)
        for ind, burst in enumerate(frame.bursts):
            box = list(g.window(burst))
            windows[(swath, ind + istart + 1)] = box
            for name, bands, dtype in layers:
                build_vrt(os.path.join(dirname, name + ".rdr"), os.path.join(outname, "%s_%02d.rdr" % (name, ind + istart + 1)),
                          [g.width, g.length], box, bands=bands, dtype=dtype)
    return dict(grid=g, topo=topo, windows=windows, orbit=orb,
                bbox=[topo.minimumLatitude, topo.maximumLatitude, topo.minimumLongitude, topo.maximumLongitude])


# ---------------------------------------------------------------------------------------------------------------
# stack geo2rdr: one resident reference geometry x N secondary dates
# ---------------------------------------------------------------------------------------------------------------
class Geo2rdrStack:
    """Batched form of contrib/stack/topsStack/geo2rdr.py ``runGeo2rdrCPU`` (:51-104).

        st = Geo2rdrStack(devices=[0, 1, ...])
        st.add_geometry("IW1/01", lat="geom_reference/IW1/lat_01.rdr", lon=..., hgt=...)     # once per burst
        st.add_job("IW1/01", info=secondary_burst, rangeOffName=..., azOffName=..., misreg_az=0.0, misreg_rg=0.0)
        st.run()

    Geometries are dealt round-robin to the devices and stay in HBM; every job of a geometry runs on the device
    that holds it (one host thread per device, no inter-GPU exchange).  `info` carries the attributes the reference
    reads from the secondary burst: rangePixelSize, azimuthTimeInterval, radarWavelength, orbit, numberOfSamples,
    numberOfLines, sensingStart, startingRange.  Outputs: FLOAT ``.off`` rasters with .xml / .vrt, invalid = -999999."""

    def __init__(self, devices=None, look_side=-1, a=EarthMajorSemiAxis, e2=EarthEccentricitySquared,
                 output_precision="single", orbit_method="HERMITE"):
        self.devices = list(devices) if devices else [0]
        self.look_side = look_side
        self.a, self.e2 = a, e2
        self.output_precision = output_precision
        self.orbit_method = orbit_method
        self._geoms = {}   # key -> dict(lat, lon, hgt arrays, device index)
        self._jobs = []
        self.results = []

    @staticmethod
    def _load(x):
        if isinstance(x, np.ndarray):
            return np.ascontiguousarray(x, dtype=np.float64)
        if isinstance(x, str):
            return np.ascontiguousarray(IF.read_view(x), dtype=np.float64)
        return np.ascontiguousarray(IF.read_raster(x), dtype=np.float64)

    def add_geometry(self, key, lat, lon, hgt):
        la, lo, h = self._load(lat), self._load(lon), self._load(hgt)
        if not (la.shape == lo.shape == h.shape) or la.ndim != 2:
            raise Exception("lat/lon/hgt of geometry {0} must be single-band images of one shape".format(key))
        self._geoms[key] = dict(lat=la, lon=lo, hgt=h, slot=len(self._geoms) % len(self.devices))

    def add_job(self, key, info, rangeOffName, azOffName, misreg_az=0.0, misreg_rg=0.0, doppler=(0.0,)):
        if key not in self._geoms:
            raise KeyError("geometry {0} has not been added".format(key))
        self._jobs.append(dict(key=key, info=info, rg=rangeOffName, az=azOffName, misreg_az=misreg_az, misreg_rg=misreg_rg,
                               doppler=tuple(doppler)))

    def _out_image(self, filename, shape):
        img = IF.createImage()
        img.setFilename(filename)
        img.setAccessMode("write")
        if self.output_precision.upper() == "SINGLE":
            img.setDataType("FLOAT")
            img.setCaster("write", "DOUBLE")
        else:
            img.setDataType("DOUBLE")
        img.setWidth(shape[1])
        img.setLength(shape[0])
        img.createImage()
        return img

    def _params(self, job, shape, device):
        info = job["info"]
        # contrib/stack/topsStack/geo2rdr.py:66-91
        delta = datetime.timedelta(seconds=job["misreg_az"] * info.azimuthTimeInterval)
        start = info.sensingStart - delta
        p = _capi.geo_params(length=int(info.numberOfLines), width=int(info.numberOfSamples), dem_shape=shape,
                             r0=float(info.startingRange - job["misreg_rg"]), dr=float(info.rangePixelSize),
                             prf=1.0 / float(info.azimuthTimeInterval), t0=seconds_since_midnight(start),
                             wvl=float(info.radarWavelength), side=int(self.look_side), a=self.a, e2=self.e2,
                             orbit_method=self.orbit_method, device=device,
                             out_f32=self.output_precision.upper() == "SINGLE")
        return p, start

    def run(self):
        nd = len(self.devices)
        by_slot = [[] for _ in range(nd)]
        for j in self._jobs:
            by_slot[self._geoms[j["key"]]["slot"]].append(j)
        results, errors = [[] for _ in range(nd)], [None] * nd

        def work(s):
            try:
                dev = self.devices[s]
                plans = {}
                for job in by_slot[s]:
                    gm = self._geoms[job["key"]]
                    shape = gm["lat"].shape
                    p, start = self._params(job, shape, dev)
                    if job["key"] not in plans:  # upload once, reuse for every date
                        plans[job["key"]] = _capi.GeoPlan(p, lat=gm["lat"], lon=gm["lon"], hgt=gm["hgt"])
                        if sum(1 for j in by_slot[s] if j["key"] == job["key"]) > 1:
                            # several dates on this geometry: LLH -> ECEF once instead of once per date
                            plans[job["key"]].freeze_geometry(self.a, self.e2)
                    plan = plans[job["key"]]
                    t, pos, vel = export_rows(job["info"].orbit, start)
                    ms = plan.execute(p, t, pos, vel, doppler_coeffs=job["doppler"], want=("azoff", "rgoff"))
                    rg = self._out_image(job["rg"], shape)
                    az = self._out_image(job["az"], shape)
                    azm, rgm = az.memMap(), rg.memMap()
                    with IF.file_backed([azm, rgm]):  # the .off rasters of this date are written with pwrite
                        r = plan.fetch(out=dict(azt=None, rgm=None, azoff=azm, rgoff=rgm))
                    for img in (rg, az):
                        img.finalizeImage()
                        img.renderHdr()
                    r = dict(r)
                    r.update(key=job["key"], device=dev, ms_kernels=ms, rangeOffName=job["rg"], azOffName=job["az"])
                    results[s].append(r)
                for plan in plans.values():
                    plan.close()
            except Exception as e:  # surfaced in the caller's thread
                errors[s] = e

        # the geometry rasters are mappings of the files topo wrote: declared once, for every slot's uploads (pread)
        with IF.file_backed([], inputs=[a for g in self._geoms.values() for a in (g["lat"], g["lon"], g["hgt"])]):
            if nd == 1:
                work(0)
            else:
                th = [threading.Thread(target=work, args=(s,)) for s in range(nd)]
                for x in th:
                    x.start()
                for x in th:
                    x.join()
        for e in errors:
            if e is not None:
                raise e
        self.results = [r for rs in results for r in rs]
        self._jobs = []
        return self.results
