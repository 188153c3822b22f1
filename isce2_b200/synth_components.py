"""Drive the drop-in Components on a synthetic scene the way the reference's applications do: createTopozero().topo()
(components/isceobj/StripmapProc/runTopo.py:66-103, TopsProc/runTopo.py:69-88) with createGeo2rdr() for a secondary
acquisition (StripmapProc/runGeo2rdr.py:57-110) chained onto it, writing the .rdr / .off rasters with their .xml / .vrt.
Used by bench.py's e2e_component arm and by the Component tests; scenes come from isce2_b200.synth."""
import datetime
import os

import numpy as np

from . import image as IF
from .orbit import Orbit
from .planet import Planet
from .poly import Poly2D


def prepare_dem(sc, path, as_int16=False):
    """Write sc.dem as an ISCE DEM raster (+ XML) and return the loaded image object (what the 'dem' port is wired to)."""
    dem = IF.createDemImage()
    arr = np.round(sc.dem).astype(np.int16) if as_int16 else np.ascontiguousarray(sc.dem, np.float32)
    arr.tofile(path)
    dem.initImage(path, "read", sc.dem.shape[1], "SHORT" if as_int16 else "FLOAT")
    dem.setLength(sc.dem.shape[0])
    dem.firstLatitude, dem.firstLongitude = sc.first_lat, sc.first_lon
    dem.deltaLatitude, dem.deltaLongitude = sc.delta_lat, sc.delta_lon
    dem.renderHdr()
    return IF.createDemImage().load(path + ".xml")


def _day(sc):
    return sc.sensing_start.replace(hour=0, minute=0, second=0, microsecond=0)


def make_topo(sc, dem_img, outdir, *, dem_method="BIQUINTIC", orbit_method="HERMITE", inc=True, mask=True, devices=None):
    from . import createTopozero
    topo = createTopozero()
    topo.slantRangePixelSpacing, topo.prf, topo.radarWavelength = sc.dr, sc.prf, sc.wvl
    topo.orbit = Orbit.from_arrays(_day(sc), sc.orbit_t, sc.orbit_pos, sc.orbit_vel)
    topo.width, topo.length = sc.width, sc.length
    topo.wireInputPort(name="dem", object=dem_img)
    topo.wireInputPort(name="planet", object=Planet(pname="Earth"))
    topo.numberRangeLooks = topo.numberAzimuthLooks = 1
    topo.lookSide = sc.side
    topo.sensingStart = sc.sensing_start
    topo.rangeFirstSample = sc.r0
    topo.demInterpolationMethod = dem_method
    topo.orbitInterpolationMethod = orbit_method
    if any(c != 0.0 for row in sc.doppler_coeffs for c in row):  # native Doppler (StripmapProc/runTopo.py:91-101)
        dop = Poly2D()
        dop.setWidth(sc.width)
        dop.setLength(sc.length)
        dop.initPoly(rangeOrder=len(sc.doppler_coeffs[0]) - 1, azimuthOrder=len(sc.doppler_coeffs) - 1, coeffs=sc.doppler_coeffs)
        topo.polyDoppler = dop
    topo.latFilename = os.path.join(outdir, "lat.rdr")
    topo.lonFilename = os.path.join(outdir, "lon.rdr")
    topo.heightFilename = os.path.join(outdir, "hgt.rdr")
    topo.losFilename = os.path.join(outdir, "los.rdr")
    if inc:
        topo.incFilename = os.path.join(outdir, "incLocal.rdr")
    if mask:
        topo.maskFilename = os.path.join(outdir, "shadowMask.rdr")
    if devices is not None:
        topo.gpuDevices = list(devices)
    return topo


def make_geo2rdr(sc, sec, outdir, *, t0, r0, orbit_method="HERMITE", doppler_cycles_per_prf=(0.0,), double=False):
    """Geo2rdr for the secondary acquisition `sec` on the reference grid of `sc` (sensing start / starting range already
    misregistered by the caller, contrib/stack/topsStack/geo2rdr.py:90-91)."""
    from . import createGeo2rdr
    g = createGeo2rdr()
    g.configure()
    g.slantRangePixelSpacing, g.prf, g.radarWavelength = sc.dr, sc.prf, sc.wvl
    g.orbit = Orbit.from_arrays(_day(sc), sec.orbit_t, sec.orbit_pos, sec.orbit_vel)
    g.width, g.length = sc.width, sc.length
    g.wireInputPort(name="planet", object=Planet(pname="Earth"))
    g.lookSide = sc.side
    g.setSensingStart(_day(sc) + datetime.timedelta(seconds=t0))
    g.rangeFirstSample = r0
    g.numberRangeLooks = g.numberAzimuthLooks = 1
    g.dopplerCentroidCoeffs = list(doppler_cycles_per_prf)
    g.fmrateCoeffs = [0.0]
    g.orbitInterpolationMethod = orbit_method
    g.rangeOffsetImageName = os.path.join(outdir, "range.off")
    g.azimuthOffsetImageName = os.path.join(outdir, "azimuth.off")
    if double:
        g.outputPrecision = "DOUBLE"
    return g


def run_components(sc, sec, dem_img, outdir, *, dem_method="BIQUINTIC", orbit_method="HERMITE", inc=True, mask=True, devices=None,
                   misreg_az=0.013, misreg_rg=1.7):
    """topo() with the secondary's geo2rdr chained onto it; returns what was written."""
    os.makedirs(outdir, exist_ok=True)
    topo = make_topo(sc, dem_img, outdir, dem_method=dem_method, orbit_method=orbit_method, inc=inc, mask=mask, devices=devices)
    grdr = make_geo2rdr(sc, sec, outdir, t0=sc.t0 - misreg_az, r0=sc.r0 - misreg_rg, orbit_method=orbit_method,
                        doppler_cycles_per_prf=[c / sc.prf for c in sc.doppler_coeffs[0]])
    topo.chainGeo2rdr(grdr)
    topo.topo()
    files = [f for f in os.listdir(outdir) if f.endswith((".rdr", ".off"))]
    return dict(files=sorted(files), bytes_written=sum(os.path.getsize(os.path.join(outdir, f)) for f in files),
                snwe=topo.snwe, num_valid=getattr(grdr, "numValid", None), gpu_timings=getattr(topo, "gpuTimings", None))


def run_components_separately(sc, sec, dem_img, outdir, *, dem_method="BIQUINTIC", orbit_method="HERMITE", inc=True, mask=True,
                              devices=None, misreg_az=0.013, misreg_rg=1.7):
    """The reference's own sequence (TopsProc/runTopo.py, then runGeo2rdr.py in a later step): topo() writes its rasters,
    geo2rdr() reads lat / lon / hgt back from them.  Returns the seconds of the two calls and what was written."""
    import time
    from . import image as IF
    os.makedirs(outdir, exist_ok=True)
    topo = make_topo(sc, dem_img, outdir, dem_method=dem_method, orbit_method=orbit_method, inc=inc, mask=mask, devices=devices)
    t0 = time.perf_counter()
    topo.topo()
    t1 = time.perf_counter()
    grdr = make_geo2rdr(sc, sec, outdir, t0=sc.t0 - misreg_az, r0=sc.r0 - misreg_rg, orbit_method=orbit_method,
                        doppler_cycles_per_prf=[c / sc.prf for c in sc.doppler_coeffs[0]])
    if devices is not None:
        grdr.gpuDevices = list(devices)
    imgs = {}
    for key, name in (("lat", "lat.rdr"), ("lon", "lon.rdr"), ("hgt", "hgt.rdr")):
        img = IF.createImage()
        img.load(os.path.join(outdir, name + ".xml"))
        img.setAccessMode("READ")
        imgs[key] = img
    grdr.geo2rdr(latImage=imgs["lat"], lonImage=imgs["lon"], demImage=imgs["hgt"])
    t2 = time.perf_counter()
    files = [f for f in os.listdir(outdir) if f.endswith((".rdr", ".off"))]
    return dict(seconds_topo=t1 - t0, seconds_geo2rdr=t2 - t1, files=sorted(files),
                bytes_written=sum(os.path.getsize(os.path.join(outdir, f)) for f in files), num_valid=getattr(grdr, "numValid", None))
