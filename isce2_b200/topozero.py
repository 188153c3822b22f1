"""Drop-in for ``zerodop.topozero``: ``createTopozero()`` -> ``Topo`` with the attribute / port / method surface of
components/zerodop/topozero/Topozero.py (class Topo :44-751), driving the B200 CUDA library instead of the
Fortran extension module.

Differences visible to a caller: none in the interface; the arithmetic runs on the GPU(s) named by
``self.gpuDevices`` (default: device 0; set e.g. ``[0,1,2,3]`` or the environment variable B200_DEVICES=0,1,2,3 to
shard the azimuth lines over several GPUs of the box -- one host thread per GPU, no inter-GPU exchange), and a
missing GPU / unbuilt library raises instead of silently computing on the CPU.
"""
from __future__ import annotations

import datetime
import os
import threading

import numpy as np

from . import _capi, image as IF
from .component import Component
from .orbit import enu_heading_deg, export_rows, seconds_since_midnight
from .planet import EarthEccentricitySquared, EarthMajorSemiAxis
from .poly import Poly2D, poly2d_fields


def _devices(attr):
    if attr:
        return [int(d) for d in attr]
    env = os.environ.get("B200_DEVICES", "")
    if env.strip():
        return [int(d) for d in env.split(",") if d.strip() != ""]
    return [0]


class Topo(Component):
    family = "topo"
    logging_name = "isce.zerodop.topozero"

    # Topozero.py:46-55
    interpolationMethods = {"SINC": 0, "BILINEAR": 1, "BICUBIC": 2, "NEAREST": 3, "AKIMA": 4, "BIQUINTIC": 5}
    orbitInterpolationMethods = {"HERMITE": 0, "SCH": 1, "LEGENDRE": 2}

    # ---- Topozero.py:59-69 ----
    @property
    def snwe(self):
        return (self.minimumLatitude, self.maximumLatitude, self.minimumLongitude, self.maximumLongitude)

    @snwe.setter
    def snwe(self, snwe):
        (self.minimumLatitude, self.maximumLatitude, self.minimumLongitude, self.maximumLongitude) = snwe

    # ---- Topozero.py:72-131 ----
    def topo(self, demImage=None, intImage=None):
        for port in self._inputPorts:
            port()
        if demImage is not None:
            self.demImage = demImage
        if intImage is not None:
            self.intImage = intImage
            if self.width is None:
                self.width = self.intImage.getWidth()
            if self.length is None:
                self.length = self.intImage.getLength()
        self.setDefaults()
        self.createImages()
        self.updateDefaults()
        self._run()
        self.destroyImages()
        return None

    def _run(self):
        dem = IF.read_raster(self.demImage)
        if dem.dtype not in (np.float32, np.int16):
            dem = np.asarray(dem).astype(np.float32)  # the reference's 'read' FLOAT caster (Topozero.py:380)
        if dem.ndim != 2:
            raise Exception("DEM must be a single-band image")
        t, pos, vel = export_rows(self.orbit, self.sensingStart)  # Orbit.exportToC(reference=sensingStart) :123
        dcoef, dmr, dma, dnr, dna = poly2d_fields(self.polyDoppler)
        keep = _capi._Keep()
        dop = _capi.make_poly2d(keep, dcoef, dmr, dma, dnr, dna)
        rho_image = None
        slr = None
        if isinstance(self.slantRangeImage, Poly2D) or hasattr(self.slantRangeImage, "getCoeffs"):
            scoef, smr, sma, snr, sna = poly2d_fields(self.slantRangeImage)
            slr = _capi.make_poly2d(keep, scoef, smr, sma, snr, sna)
        else:
            rho_image = np.ascontiguousarray(IF.read_raster(self.slantRangeImage), dtype=np.float64)
        devices = _devices(self.gpuDevices)
        n = len(devices)
        length, width = int(self.length), int(self.width)
        # output layers: this package's images, or images the caller handed in (Topozero.py:274-302), mapped writable
        outs = dict(lat=IF.output_memmap(self.latImage, length, width), lon=IF.output_memmap(self.lonImage, length, width),
                    hgt=IF.output_memmap(self.heightImage, length, width),
                    los=IF.output_memmap(self.losImage, length, width, 2) if self.losImage else None,
                    inc=IF.output_memmap(self.incImage, length, width, 2) if self.incImage else None,
                    mask=IF.output_memmap(self.maskImage, length, width) if self.maskImage else None)
        results, errors = [None] * n, [None] * n
        chained = [g._chain_prepare(self) for g in self.chainedGeo2rdr]
        chained_results = [[None] * n for _ in chained]

        def work(i):
            a = (length * i) // n
            b = (length * (i + 1)) // n
            if b <= a:
                return
            p = _capi.topo_params(dem_shape=dem.shape, first_lat=float(self.firstLatitude), first_lon=float(self.firstLongitude),
                                  delta_lat=float(self.deltaLatitude), delta_lon=float(self.deltaLongitude), length=length,
                                  width=width, prf=float(self.prf), t0=seconds_since_midnight(self.sensingStart),
                                  wvl=float(self.radarWavelength), side=int(self.lookSide), peg_heading=float(self.pegHeading),
                                  a=float(self.ellipsoidMajorSemiAxis), e2=float(self.ellipsoidEccentricitySquared),
                                  dem_method=self.demInterpolationMethod, orbit_method=self.orbitInterpolationMethod,
                                  numiter=int(self.numberIterations), extraiter=int(self.secondaryIterations),
                                  thresh=float(self.threshold), nrnglooks=int(self.numberRangeLooks),
                                  nazlooks=int(self.numberAzimuthLooks), line0=a, nlines=b - a, device=devices[i])
            blk = {k: (v[a:b] if v is not None else None) for k, v in outs.items()}
            try:
                if chained:
                    # fused verb: geo2rdr of every chained component on the block's layers while they are resident
                    jobs = [dict(params=c["params"](a, b - a, devices[i]), orbit=c["orbit"], doppler=c["doppler"], want=c["want"],
                                 out={k: (v[a:b] if v is not None else None) for k, v in c["outs"].items()}) for c in chained]
                    results[i], geos = _capi.topo_geo2rdr_run(p, dem, t, pos, vel, None, jobs, rho_image=rho_image,
                                                              want_los=blk["los"] is not None, want_inc=blk["inc"] is not None,
                                                              want_mask=blk["mask"] is not None, out=blk, doppler_poly=dop,
                                                              slrng_poly=slr)
                    for j, gres in enumerate(geos):
                        chained_results[j][i] = gres
                else:
                    results[i] = _capi.topo_run(p, dem, t, pos, vel, None, None, rho_image=rho_image,
                                                want_los=blk["los"] is not None, want_inc=blk["inc"] is not None,
                                                want_mask=blk["mask"] is not None, out=blk, doppler_poly=dop, slrng_poly=slr)
            except Exception as e:  # surfaced below, in the caller's thread
                errors[i] = e

        # the rasters being written are file mappings: say so, and the library writes them with pwrite (image.file_backed)
        with IF.file_backed(list(outs.values()) + [v for c in chained for v in c["outs"].values()]):
            if n == 1:
                work(0)
            else:
                th = [threading.Thread(target=work, args=(i,)) for i in range(n)]
                for x in th:
                    x.start()
                for x in th:
                    x.join()
        for e in errors:
            if e is not None:
                raise e
        res = [r for r in results if r is not None]
        # getState (Topozero.py:541-546): bbox over all blocks, reduced on the host (no inter-GPU exchange)
        self.minimumLatitude = min(r["min_lat"] for r in res)
        self.maximumLatitude = max(r["max_lat"] for r in res)
        self.minimumLongitude = min(r["min_lon"] for r in res)
        self.maximumLongitude = max(r["max_lon"] for r in res)
        self.totalConverged = sum(r["converged"] for r in res)
        self.gpuTimings = [{k: r[k] for k in ("ms_setup", "ms_kernels", "ms_pixels", "ms_mask", "ms_total", "gpu_launches")}
                           for r in res]
        self.logger.info("Total convergence: %d out of %d", self.totalConverged, length * width)
        for g, rs in zip(self.chainedGeo2rdr, chained_results):
            g._chain_finish([r for r in rs if r is not None])

    def chainGeo2rdr(self, grdr):
        """B200 extension (not in the reference): run `grdr` (a configured Geo2rdr component: orbit, sensingStart, output
        file names ...) inside this component's topo() call, on the lat / lon / hgt layers of every block of lines while
        they are still in GPU memory (b200_topo_geo2rdr_run).  Equivalent to calling
        ``grdr.geo2rdr(latImage=, lonImage=, demImage=)`` on this component's output files afterwards -- same .off / .rdr
        rasters, same XML -- without reading them back from disk and uploading them again."""
        self.chainedGeo2rdr.append(grdr)
        return grdr

    # ---- Topozero.py:133-205 ----
    def setDefaults(self):
        if self.ellipsoidMajorSemiAxis is None:
            self.ellipsoidMajorSemiAxis = EarthMajorSemiAxis
        if self.ellipsoidEccentricitySquared is None:
            self.ellipsoidEccentricitySquared = EarthEccentricitySquared
        if self.numberIterations is None:
            self.numberIterations = 25
        if self.secondaryIterations is None:
            self.secondaryIterations = 10
        if self.threshold is None:
            self.threshold = 0.05
        if self.heightFilename == '':
            self.heightFilename = 'z.rdr'
            self.logger.warning('The real height file has been given the default name %s' % (self.heightFilename))
        if self.latFilename == '':
            self.latFilename = 'lat.rdr'
            self.logger.warning('The latitude file has been given the default name %s' % (self.latFilename))
        if self.lonFilename == '':
            self.lonFilename = 'lon.rdr'
            self.logger.warning('The longitude file has been given the default name %s' % (self.lonFilename))
        if self.losFilename == '':
            self.losFilename = 'los.rdr'
            self.logger.warning('The los file has been given the default name %s' % (self.losFilename))
        if self.numberRangeLooks is None:
            self.numberRangeLooks = 1
        if self.numberAzimuthLooks is None:
            self.numberAzimuthLooks = 1
        if self.lookSide is None:
            self.lookSide = -1
        if self.pegHeading is None:
            tbef = self.sensingStart + datetime.timedelta(seconds=(0.5 * self.length / self.prf))
            if hasattr(self.orbit, "getENUHeading"):
                hdg = self.orbit.getENUHeading(tbef)
            else:
                hdg = enu_heading_deg(self.orbit, tbef)
            self.pegHeading = np.radians(hdg)
            self.logger.warning('Default Peg heading set to: ' + str(self.pegHeading))
        if self.polyDoppler is None:
            self.polyDoppler = Poly2D(name=self.name + '_dopplerPoly')
            self.polyDoppler.setWidth(self.width)
            self.polyDoppler.setLength(self.length)
            self.polyDoppler.setNormRange(1.0)
            self.polyDoppler.setNormAzimuth(1.0)
            self.polyDoppler.setMeanRange(0.0)
            self.polyDoppler.setMeanAzimuth(0.0)
            self.polyDoppler.initPoly(rangeOrder=0, azimuthOrder=0, coeffs=[[0.0]])
        else:
            if self.polyDoppler.getWidth() != self.width:
                raise Exception('Doppler Centroid object does not have the same width as input image')
            if self.polyDoppler.getLength() != self.length:
                raise Exception('Doppler Centroid object does not have the same length as input image')
        if self.demInterpolationMethod is None:
            self.demInterpolationMethod = 'BILINEAR'
        else:
            if self.demInterpolationMethod.upper() not in list(self.interpolationMethods.keys()):
                raise Exception('Interpolation method must be one of ' + str(list(self.interpolationMethods.keys())))
        if self.orbitInterpolationMethod is None:
            self.orbitInterpolationMethod = 'HERMITE'
        else:
            if self.orbitInterpolationMethod.upper() not in list(self.orbitInterpolationMethods.keys()):
                raise Exception('Orbit interpolation method must be one of ' + str(list(self.orbitInterpolationMethods.keys())))
        if self.slantRangeFilename in ['', None] and self.slantRangeImage is None:
            if self.slantRangePixelSpacing is None:
                raise Exception('No slant range file provided. slantRangePixelSpacing cannot be None')
            if self.rangeFirstSample is None:
                raise Exception('No slant range file provided. rangeFirstSample cannot be None')

    def updateDefaults(self):
        if self.demLength is None:
            self.demLength = self.demImage.getLength()
        if self.demWidth is None:
            self.demWidth = self.demImage.getWidth()
        for attr, getter in (("firstLatitude", "getFirstLatitude"), ("firstLongitude", "getFirstLongitude"),
                             ("deltaLatitude", "getDeltaLatitude"), ("deltaLongitude", "getDeltaLongitude")):
            if getattr(self, attr) is None and hasattr(self.demImage, getter):
                setattr(self, attr, getattr(self.demImage, getter)())

    # ---- Topozero.py:214-259 ----
    def destroyImages(self):
        self.latImage.addDescription('Pixel-by-pixel latitude in degrees.')
        self.latImage.finalizeImage()
        self.latImage.renderHdr()
        self.lonImage.addDescription('Pixel-by-pixel longitude in degrees.')
        self.lonImage.finalizeImage()
        self.lonImage.renderHdr()
        self.heightImage.addDescription('Pixel-by-pixel height in meters.')
        self.heightImage.finalizeImage()
        self.heightImage.renderHdr()
        descr = '''Two channel Line-Of-Sight geometry image (all angles in degrees). Represents vector drawn from target to platform.
                Channel 1: Incidence angle measured from vertical at target (always +ve).
                Channel 2: Azimuth angle measured from North in Anti-clockwise direction.'''
        self.losImage.setImageType('bil')
        self.losImage.addDescription(descr)
        self.losImage.finalizeImage()
        self.losImage.renderHdr()
        if hasattr(self.demImage, "finalizeImage"):
            self.demImage.finalizeImage()
        if self.incImage:
            descr = '''Two channel angle file.
                    Channel 1: Angle between ray to target and the vertical at the sensor
                    Channel 2: Local incidence angle accounting for DEM slope at target'''
            self.incImage.addDescription(descr)
            self.incImage.finalizeImage()
            self.incImage.renderHdr()
        if self.maskImage:
            descr = 'Radar shadow-layover mask. 1 - Radar Shadow. 2 - Radar Layover. 3 - Both.'
            self.maskImage.addDescription(descr)
            self.maskImage.finalizeImage()
            self.maskImage.renderHdr()
        if self.slantRangeImage:
            try:
                self.slantRangeImage.finalizeImage()
            except Exception:
                pass
        return

    # ---- Topozero.py:261-393 ----
    def _new_image(self, filename, dataType, bands=1, scheme='BIP'):
        img = IF.createImage()
        img.initImage(filename, 'write', self.width, dataType, bands=bands, scheme=scheme)
        img.setLength(self.length)
        return img

    def createImages(self):
        if self.demImage is None and not self.demFilename == '':
            self.demImage = IF.createDemImage()
            self.demImage.load(self.demFilename + '.xml') if os.path.exists(self.demFilename + '.xml') else \
                self.demImage.initImage(self.demFilename, 'read', self.demWidth)
        elif self.demImage is None:
            self.logger.error('Must either pass the demImage in the call or set self.demFilename.')
            raise Exception
        if self.latImage is None and not self.latFilename == '':
            self.latImage = self._new_image(self.latFilename, 'DOUBLE')
        elif self.latImage is None:
            self.logger.error('Must either pass the latImage in the call or set self.latFilename.')
            raise Exception
        if self.lonImage is None and not self.lonFilename == '':
            self.lonImage = self._new_image(self.lonFilename, 'DOUBLE')
        elif self.lonImage is None:
            self.logger.error('Must either pass the lonImage in the call or set self.lonFilename.')
            raise Exception
        if self.heightImage is None and not self.heightFilename == '':
            self.heightImage = self._new_image(self.heightFilename, 'DOUBLE')
        elif self.heightImage is None:
            self.logger.error('Must either pass the heightImage in the call or set self.heightFilename.')
            raise Exception

        if self.slantRangeImage is None and not self.slantRangeFilename == '':
            if self.rangeFirstSample:
                raise Exception('Cannot provide both slant range image and range first sample as input')
            if self.slantRangePixelSpacing:
                raise Exception('Cannot provide both slant range image and slant range pixel spacing as input')
            self.slantRangeImage = IF.createImage()
            self.slantRangeImage.load(self.slantRangeFilename + '.xml')
            self.slantRangeImage.setAccessMode('READ')
            if self.slantRangeImage.width != self.width:
                raise Exception('Slant Range Image width {0} does not match input width {1}'.format(self.slantRangeImage.width, self.width))
            if self.slantRangeImage.length != self.length:
                raise Exception('Slant Range Image length {0} does not match input length {1}'.format(self.slantRangeImage.length, self.length))
            self.slantRangeImage.createImage()
            self.rangeFirstSample = 0.0
            self.slantRangePixelSpacing = 0.0
        elif self.slantRangeImage is not None:
            if self.slantRangeImage.width != self.width:
                raise Exception('Slant Range Image width {0} does not match input width {1}'.format(self.slantRangeImage.width, self.width))
            if self.slantRangeImage.length != self.length:
                raise Exception('Slant Range Image length {0} does not match input length {1}'.format(self.slantRangeImage.length, self.length))
        else:
            r0 = self.rangeFirstSample
            dr = self.slantRangePixelSpacing * self.numberRangeLooks
            self.slantRangeImage = Poly2D()
            self.slantRangeImage.setWidth(self.width)
            self.slantRangeImage.setLength(self.length)
            self.slantRangeImage.setNormRange(1.0)
            self.slantRangeImage.setNormAzimuth(1.0)
            self.slantRangeImage.setMeanRange(0.0)
            self.slantRangeImage.setMeanAzimuth(0.0)
            self.slantRangeImage.initPoly(rangeOrder=1, azimuthOrder=0, coeffs=[[r0, dr]])

        if self.losImage is None and not self.losFilename == '':
            self.losImage = self._new_image(self.losFilename, 'FLOAT', bands=2, scheme='BIL')
        if self.incImage is None and not self.incFilename == '':
            self.incImage = self._new_image(self.incFilename, 'FLOAT', bands=2, scheme='BIL')
        if self.maskImage is None and not self.maskFilename == '':
            self.maskImage = self._new_image(self.maskFilename, 'BYTE', bands=1, scheme='BIL')

        if hasattr(self.demImage, "setCaster"):
            self.demImage.setCaster('read', 'FLOAT')
        for img in (self.latImage, self.lonImage, self.heightImage, self.losImage, self.incImage, self.maskImage):
            if img is not None:
                if getattr(img, "length", None) is None:
                    img.setLength(self.length)
                img.createImage()
        return

    # ---- setters Topozero.py:433-539 ----
    def setNumberIterations(self, var): self.numberIterations = int(var)
    def setSecondaryIterations(self, var): self.secondaryIterations = int(var)
    def setThreshold(self, var): self.threshold = float(var)
    def setDemWidth(self, var): self.demWidth = int(var)
    def setDemLength(self, var): self.demLength = int(var)
    def setOrbit(self, var): self.orbit = var
    def setFirstLatitude(self, var): self.firstLatitude = float(var)
    def setFirstLongitude(self, var): self.firstLongitude = float(var)
    def setDeltaLatitude(self, var): self.deltaLatitude = float(var)
    def setDeltaLongitude(self, var): self.deltaLongitude = float(var)
    def setEllipsoidMajorSemiAxis(self, var): self.ellipsoidMajorSemiAxis = float(var)
    def setEllipsoidEccentricitySquared(self, var): self.ellipsoidEccentricitySquared = float(var)
    def setLength(self, var): self.length = int(var)
    def setWidth(self, var): self.width = int(var)
    def setRangePixelSpacing(self, var): self.slantRangePixelSpacing = float(var)
    def setRangeFirstSample(self, var): self.rangeFirstSample = float(var)
    def setNumberRangeLooks(self, var): self.numberRangeLooks = int(var)
    def setNumberAzimuthLooks(self, var): self.numberAzimuthLooks = int(var)
    def setPegHeading(self, var): self.pegHeading = float(var)
    def setPRF(self, var): self.prf = float(var)
    def setRadarWavelength(self, var): self.radarWavelength = float(var)
    def setLosFilename(self, var): self.losFilename = var
    def setLatFilename(self, var): self.latFilename = var
    def setLonFilename(self, var): self.lonFilename = var
    def setHeightFilename(self, var): self.heightFilename = var
    def setIncidenceFilename(self, var): self.incFilename = var
    def setMaskFilename(self, var): self.maskFilename = var
    def setLookSide(self, var): self.lookSide = int(var)
    def setPolyDoppler(self, var): self.polyDoppler = var.copy()
    def getMinimumLatitude(self): return self.minimumLatitude
    def getMinimumLongitude(self): return self.minimumLongitude
    def getMaximumLatitude(self): return self.maximumLatitude
    def getMaximumLongitude(self): return self.maximumLongitude

    # ---- ports Topozero.py:560-609 ----
    def addPlanet(self):
        planet = self._inputPorts.getPort(name='planet').getObject()
        if (planet):
            try:
                ellipsoid = planet.get_elp()
                self.ellipsoidMajorSemiAxis = ellipsoid.get_a()
                self.ellipsoidEccentricitySquared = ellipsoid.get_e2()
            except AttributeError as strerr:
                self.logger.error(strerr)
                raise AttributeError

    def addFrame(self):
        frame = self._inputPorts.getPort(name='frame').getObject()
        if (frame):
            try:
                instrument = frame.getInstrument()
                self.slantRangePixelSpacing = instrument.getRangePixelSize()
                self.prf = instrument.getPulseRepetitionFrequency()
                self.radarWavelength = instrument.getRadarWavelength()
                self.orbit = frame.getOrbit()
            except AttributeError as strerr:
                self.logger.error(strerr)
                raise AttributeError

    def addDEM(self):
        dem = self._inputPorts.getPort(name='dem').getObject()
        if (dem):
            try:
                self.demImage = dem
                self.demWidth = dem.getWidth()
                self.demLength = dem.getLength()
                self.firstLatitude = dem.getFirstLatitude()
                self.firstLongitude = dem.getFirstLongitude()
                self.deltaLatitude = dem.getDeltaLatitude()
                self.deltaLongitude = dem.getDeltaLongitude()
            except AttributeError as strerr:
                self.logger.error(strerr)
                raise AttributeError

    def addInterferogram(self):
        ifg = self._inputPorts.getPort(name='interferogram').getObject()
        if (ifg):
            try:
                self.intImage = ifg
                self.width = ifg.getWidth()
                self.length = ifg.getLength()
            except AttributeError as strerr:
                self.logger.error(strerr)
                raise AttributeError

    # ---- Topozero.py:613-704 ----
    def __init__(self):
        super(Topo, self).__init__()
        self.numberIterations = None
        self.secondaryIterations = None
        self.threshold = None
        self.demWidth = None
        self.demLength = None
        self.orbit = None
        self.sensingStart = None
        self.firstLatitude = None
        self.firstLongitude = None
        self.deltaLatitude = None
        self.deltaLongitude = None
        self.ellipsoidMajorSemiAxis = None
        self.ellipsoidEccentricitySquared = None
        self.length = None
        self.width = None
        self.slantRangePixelSpacing = None
        self.rangeFirstSample = None
        self.numberRangeLooks = None
        self.numberAzimuthLooks = None
        self.pegHeading = None
        self.prf = None
        self.radarWavelength = None
        self.demFilename = ''
        self.latFilename = ''
        self.lonFilename = ''
        self.heightFilename = ''
        self.losFilename = ''
        self.incFilename = ''
        self.maskFilename = ''
        self.slantRangeFilename = ''
        self.demImage = None
        self.latImage = None
        self.lonImage = None
        self.heightImage = None
        self.losImage = None
        self.incImage = None
        self.maskImage = None
        self.slantRangeImage = None
        self.intImage = None
        self.demAccessor = None
        self.latAccessor = None
        self.lonAccessor = None
        self.heightAccessor = None
        self.losAccessor = None
        self.incAccessor = None
        self.maskAccessor = None
        self.slantRangeAccessor = None
        self.minimumLatitude = None
        self.minimumLongitude = None
        self.maximumLatitude = None
        self.maximumLongitude = None
        self.lookSide = None
        self.polyDoppler = None
        self.polyDopplerAccessor = None
        self.demInterpolationMethod = None
        self.orbitInterpolationMethod = None
        self.gpuDevices = None  # B200 extension: list of CUDA device ordinals to shard the azimuth lines over
        self.gpuTimings = None
        self.totalConverged = None
        self.chainedGeo2rdr = []  # B200 extension: see chainGeo2rdr()
        self.dictionaryOfVariables = {
            'NUMBER_ITERATIONS': ['numberIterations', 'int', 'optional'],
            'DEM_WIDTH': ['demWidth', 'int', 'mandatory'],
            'DEM_LENGTH': ['demLength', 'int', 'mandatory'],
            'FIRST_LATITUDE': ['firstLatitude', 'float', 'mandatory'],
            'FIRST_LONGITUDE': ['firstLongitude', 'float', 'mandatory'],
            'DELTA_LATITUDE': ['deltaLatitude', 'float', 'mandatory'],
            'DELTA_LONGITUDE': ['deltaLongitude', 'float', 'mandatory'],
            'ELLIPSOID_MAJOR_SEMIAXIS': ['ellipsoidMajorSemiAxis', 'float', 'optional'],
            'ELLIPSOID_ECCENTRICITY_SQUARED': ['ellipsoidEccentricitySquared', 'float', 'optional'],
            'LENGTH': ['length', 'int', 'mandatory'],
            'WIDTH': ['width', 'int', 'mandatory'],
            'SLANT_RANGE_PIXEL_SPACING': ['slantRangePixelSpacing', 'float', 'mandatory'],
            'RANGE_FIRST_SAMPLE': ['rangeFirstSample', 'float', 'mandatory'],
            'NUMBER_RANGE_LOOKS': ['numberRangeLooks', 'int', 'mandatory'],
            'NUMBER_AZIMUTH_LOOKS': ['numberAzimuthLooks', 'int', 'mandatory'],
            'PEG_HEADING': ['pegHeading', 'float', 'mandatory'],
            'PRF': ['prf', 'float', 'mandatory'],
            'RADAR_WAVELENGTH': ['radarWavelength', 'float', 'mandatory'],
            'LAT_ACCESSOR': ['latAccessor', 'int', 'optional'],
            'LON_ACCESSOR': ['lonAccessor', 'int', 'optional'],
            'HEIGHT_R_ACCESSOR': ['heightAccessor', 'int', 'optional'],
        }
        self.dictionaryOfOutputVariables = {
            'MINIMUM_LATITUDE': 'minimumLatitude',
            'MINIMUM_LONGITUDE': 'minimumLongitude',
            'MAXIMUM_LATITUDE': 'maximumLatitude',
            'MAXIMUM_LONGITUDE': 'maximumLongitude',
        }
        self.descriptionOfVariables = {}
        self.mandatoryVariables = []
        self.optionalVariables = []
        self.initOptionalAndMandatoryLists()
        return None

    def createPorts(self):
        self.inputPorts['frame'] = self.addFrame
        self.inputPorts['planet'] = self.addPlanet
        self.inputPorts['dem'] = self.addDEM
        self.inputPorts['interferogram'] = self.addInterferogram
        return None


def createTopozero():
    """components/zerodop/topozero/__init__.py:33-35"""
    return Topo()
