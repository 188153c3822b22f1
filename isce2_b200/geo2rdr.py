"""Drop-in for ``zerodop.geo2rdr``: ``createGeo2rdr()`` -> ``Geo2rdr`` with the parameter / port / method surface of
components/zerodop/geo2rdr/Geo2rdr.py (class Geo2rdr :163-562), driving the B200 CUDA library.
"""
from __future__ import annotations

import sys
import threading

import numpy as np

from . import _capi, image as IF
from .component import Component
from .orbit import export_rows, seconds_since_midnight
from .planet import EarthEccentricitySquared, EarthMajorSemiAxis
from .poly import Poly1D, Poly2D, poly1d_fields
from .topozero import _devices

P = Component.Parameter

# Geo2rdr.py:45-161
ELLIPSOID_MAJOR_SEMIAXIS = P('ellipsoidMajorSemiAxis', public_name='ELLIPSOID_MAJOR_SEMIAXIS', default=EarthMajorSemiAxis,
                             type=float, mandatory=True, doc='Ellipsoid Major Semi Axis of planet for geocoding')
ELLIPSOID_ECCENTRICITY_SQUARED = P('ellipsoidEccentricitySquared', public_name='ELLIPSOID_ECCENTRICITY_SQUARED',
                                   default=EarthEccentricitySquared, type=float, mandatory=True,
                                   doc='Ellipsoid Eccentricity Squared of planet for geocoding')
SLANT_RANGE_PIXEL_SPACING = P('slantRangePixelSpacing', public_name='SLANT_RANGE_PIXEL_SPACING', default=None, type=float,
                              mandatory=True, doc='Slant Range Pixel Spacing (single look) in meters')
RANGE_FIRST_SAMPLE = P('rangeFirstSample', public_name='RANGE_FIRST_SAMPLE', default=None, type=float, mandatory=True,
                       doc='Range to first sample')
PRF = P('prf', public_name='PRF', default=None, type=float, mandatory=True, doc='Pulse repetition frequency')
RADAR_WAVELENGTH = P('radarWavelength', public_name='RADAR_WAVELENGTH', default=None, type=float, mandatory=True,
                     doc='Radar wavelength')
SENSING_START = P('sensingStart', public_name='SENSING_START', default=None, type=float,
                  doc='Sensing start time for the first line')
NUMBER_RANGE_LOOKS = P('numberRangeLooks', public_name='NUMBER_RANGE_LOOKS', default=None, type=int, mandatory=True,
                       doc='Number of range looks used to generate radar image')
NUMBER_AZIMUTH_LOOKS = P('numberAzimuthLooks', public_name='NUMBER_AZIMUTH_LOOKS', default=None, type=int, mandatory=True,
                         doc='Number of azimuth looks used to generate radar image')
RANGE_FILENAME = P('rangeFilename', public_name='RANGE_FILENAME', default=None, type=str, mandatory=True,
                   doc='Filename of the output range in meters')
AZIMUTH_FILENAME = P('azimuthFilename', public_name='AZIMUTH_FILENAME', default=None, type=str, mandatory=True,
                     doc='Filename of the output azimuth in seconds')
RANGE_OFFSET_FILENAME = P('rangeOffFilename', public_name='RANGE_OFFSET_FILENAME', default=None, type=str, mandatory=True,
                          doc='Filename of the output range offsets for use with resamp')
AZIMUTH_OFFSET_FILENAME = P('azimuthOffFilename', public_name='AZIMUTH_OFFSET_FILENAME', default=None, type=str,
                            mandatory=True, doc='Filename of the output azimuth offsets for use with resamp')
LOOK_SIDE = P('lookSide', public_name='LOOK_SIDE', default=None, type=int, mandatory=True,
              doc='Right (-1) / Left (1) . Look direction of the radar platform')
BISTATIC_DELAY_CORRECTION_FLAG = P('bistaticDelayCorrectionFlag', public_name='BISTATIC_DELAY_CORRECTION_FLAG', default=None,
                                   type=bool, mandatory=True, doc='Include bistatic delay correction term. E.g: ASAR / ALOS-1')
OUTPUT_PRECISION = P('outputPrecision', public_name='OUTPUT_PRECISION', default='single', type=bool, mandatory=True,
                     doc='Set to double for double precision offsets / coordinates. Angles are always single precision.')
ORBIT_INTERPOLATION_METHOD = P('orbitInterpolationMethod', public_name="orbit interpolation method", default=None, type=str,
                               mandatory=True, doc='Set to HERMITE/ SCH / LEGENDRE')


class Geo2rdr(Component):
    family = 'geo2rdr'
    logging_name = 'isce.zerodop.geo2rdr'

    parameter_list = (RANGE_FILENAME, AZIMUTH_FILENAME, RANGE_OFFSET_FILENAME, AZIMUTH_OFFSET_FILENAME,
                      SLANT_RANGE_PIXEL_SPACING, ELLIPSOID_ECCENTRICITY_SQUARED, ELLIPSOID_MAJOR_SEMIAXIS, RANGE_FIRST_SAMPLE,
                      SENSING_START, NUMBER_RANGE_LOOKS, NUMBER_AZIMUTH_LOOKS, PRF, RADAR_WAVELENGTH, LOOK_SIDE,
                      BISTATIC_DELAY_CORRECTION_FLAG, OUTPUT_PRECISION, ORBIT_INTERPOLATION_METHOD)

    orbitMethods = {'HERMITE': 0, 'SCH': 1, 'LEGENDRE': 2}

    # ---- Geo2rdr.py:192-262 ----
    def geo2rdr(self, latImage=None, lonImage=None, demImage=None):
        self.activateInputPorts()
        if latImage is not None:
            self.latImage = latImage
        if lonImage is not None:
            self.lonImage = lonImage
        if demImage is not None:
            self.demImage = demImage
        if self.orbit is None:
            raise Exception('No orbit provided for geocoding')
        self.setDefaults()
        self.createImages()
        self._run()
        self.destroyImages()
        return None

    def _layer(self, img, rows, cols):
        """lat/lon may be images or Poly2D objects evaluated at 0-based (row, col) (Geo2rdr.py:216-226)."""
        if isinstance(img, Poly2D) or (hasattr(img, "getCoeffs") and hasattr(img, "getMeanAzimuth")):
            az = np.arange(rows, dtype=np.float64)[:, None]
            rg = np.arange(cols, dtype=np.float64)[None, :]
            y = (az - img.getMeanAzimuth()) / img.getNormAzimuth()
            x = (rg - img.getMeanRange()) / img.getNormRange()
            out = np.zeros((rows, cols))
            sy = np.ones_like(y)
            for row in img.getCoeffs():
                sx = np.ones_like(x)
                for c in row:
                    out += sx * sy * c
                    sx = sx * x
                sy = sy * y
            return out
        return np.ascontiguousarray(IF.read_raster(img), dtype=np.float64)  # the 'read' DOUBLE caster (:208)

    def _run(self):
        rows, cols = int(self.demLength), int(self.demWidth)
        hgt = self._layer(self.demImage, rows, cols)
        lat = self._layer(self.latImage, rows, cols)
        lon = self._layer(self.lonImage, rows, cols)
        t, pos, vel = export_rows(self.orbit, self.sensingStart)
        coeffs, mean, norm = poly1d_fields(self.polyDoppler)
        single = self.outputPrecision.upper() == 'SINGLE'
        imgs = dict(azt=self.azimuthImage, rgm=self.rangeImage, azoff=self.azimuthOffsetImage, rgoff=self.rangeOffsetImage)
        outs = {k: (IF.output_memmap(v, rows, cols) if v is not None else None) for k, v in imgs.items()}
        want = tuple(k for k, v in outs.items() if v is not None)
        devices = _devices(self.gpuDevices)
        n = len(devices)
        results, errors = [None] * n, [None] * n

        def work(i):
            a = (rows * i) // n
            b = (rows * (i + 1)) // n
            if b <= a:
                return
            p = _capi.geo_params(length=int(self.length), width=int(self.width), dem_shape=(rows, cols),
                                 r0=float(self.rangeFirstSample), dr=float(self.slantRangePixelSpacing), prf=float(self.prf),
                                 t0=seconds_since_midnight(self.sensingStart), wvl=float(self.radarWavelength),
                                 side=int(self.lookSide), a=float(self.ellipsoidMajorSemiAxis),
                                 e2=float(self.ellipsoidEccentricitySquared), orbit_method=self.orbitInterpolationMethod,
                                 bistatic=bool(self.bistaticDelayCorrectionFlag), nrnglooks=int(self.numberRangeLooks),
                                 nazlooks=int(self.numberAzimuthLooks), line0=a, nlines=b - a, device=devices[i], out_f32=single)
            blk = {k: (v[a:b] if v is not None else None) for k, v in outs.items()}
            try:
                results[i] = _capi.geo2rdr_run(p, lat, lon, hgt, t, pos, vel, coeffs, mean, norm, want=want, out=blk)
            except Exception as e:
                errors[i] = e

        # the .off / .rdr rasters being written, and the lat / lon / hgt rasters being read, are file mappings (image.file_backed)
        with IF.file_backed(list(outs.values()), inputs=[lat, lon, hgt]):
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
        self.numOutsideImage = sum(r["num_outside"] for r in res)
        self.numValid = sum(r["num_valid"] for r in res)
        self.numConverged = sum(r["num_converged"] for r in res)
        self.gpuTimings = [{k: r[k] for k in ("ms_setup", "ms_kernels", "ms_total", "gpu_launches")} for r in res]
        self.logger.info("Number of pixels outside the image: %d; with valid data: %d", self.numOutsideImage, self.numValid)

    # ---- B200 extension: this component as a job fused behind Topo.topo() (Topo.chainGeo2rdr) ----
    def _chain_prepare(self, topo):
        """Everything geo2rdr() does before the verb, with the Topo component's output images as lat / lon / hgt."""
        self.activateInputPorts()
        self.latImage, self.lonImage, self.demImage = topo.latImage, topo.lonImage, topo.heightImage
        if self.orbit is None:
            raise Exception('No orbit provided for geocoding')
        self.setDefaults()
        self.createImages()
        rows, cols = int(self.demLength), int(self.demWidth)
        t, pos, vel = export_rows(self.orbit, self.sensingStart)
        single = self.outputPrecision.upper() == 'SINGLE'
        imgs = dict(azt=self.azimuthImage, rgm=self.rangeImage, azoff=self.azimuthOffsetImage, rgoff=self.rangeOffsetImage)
        outs = {k: (IF.output_memmap(v, rows, cols) if v is not None else None) for k, v in imgs.items()}

        def params(line0, nlines, device):
            return _capi.geo_params(length=int(self.length), width=int(self.width), dem_shape=(rows, cols),
                                    r0=float(self.rangeFirstSample), dr=float(self.slantRangePixelSpacing), prf=float(self.prf),
                                    t0=seconds_since_midnight(self.sensingStart), wvl=float(self.radarWavelength),
                                    side=int(self.lookSide), a=float(self.ellipsoidMajorSemiAxis),
                                    e2=float(self.ellipsoidEccentricitySquared), orbit_method=self.orbitInterpolationMethod,
                                    bistatic=bool(self.bistaticDelayCorrectionFlag), nrnglooks=int(self.numberRangeLooks),
                                    nazlooks=int(self.numberAzimuthLooks), line0=line0, nlines=nlines, device=device,
                                    out_f32=single)

        return dict(params=params, orbit=(t, pos, vel), doppler=poly1d_fields(self.polyDoppler), outs=outs,
                    want=tuple(k for k, v in outs.items() if v is not None))

    def _chain_finish(self, res):
        self.numOutsideImage = sum(r["num_outside"] for r in res)
        self.numValid = sum(r["num_valid"] for r in res)
        self.numConverged = sum(r["num_converged"] for r in res)
        self.gpuTimings = [{k: r[k] for k in ("ms_setup", "ms_kernels", "ms_total", "gpu_launches")} for r in res]
        self.logger.info("Number of pixels outside the image: %d; with valid data: %d", self.numOutsideImage, self.numValid)
        # the lat / lon / hgt images belong to the Topo component, which finalizes them itself
        for outfile in [self.rangeImage, self.azimuthImage, self.rangeOffsetImage, self.azimuthOffsetImage]:
            if outfile is not None:
                outfile.finalizeImage()
                outfile.renderHdr()
        self.polyDopplerAccessor = None

    # ---- Geo2rdr.py:264-297 ----
    def setDefaults(self):
        if self.polyDoppler is None:
            self.polyDoppler = Poly1D(name=self.name + '_geo2rdrPoly')
            self.polyDoppler.setMean(0.0)
            self.polyDoppler.initPoly(order=len(self.dopplerCentroidCoeffs) - 1, coeffs=self.dopplerCentroidCoeffs)
        if all(v is None for v in [self.rangeImageName, self.azimuthImageName, self.rangeOffsetImageName,
                                   self.azimuthOffsetImageName]):
            print('No outputs requested from geo2rdr. Check again.')
            sys.exit(0)
        if self.demWidth is None:
            self.demWidth = self.demImage.width
        if self.demLength is None:
            self.demLength = self.demImage.length
        if any(v != self.demWidth for v in [self.demImage.width, self.latImage.width, self.lonImage.width]):
            print('Input lat, lon, z images should all have the same width')
            sys.exit(0)
        if any(v != self.demLength for v in [self.demImage.length, self.latImage.length, self.lonImage.length]):
            print('Input lat, lon, z images should all have the same length')
            sys.exit(0)
        if self.bistaticDelayCorrectionFlag is None:
            self.bistaticDelayCorrectionFlag = False
            print('Turning off bistatic delay correction term by default.')
        if self.orbitInterpolationMethod is None:
            self.orbitInterpolationMethod = 'HERMITE'
        if self.numberRangeLooks is None:
            self.numberRangeLooks = 1
        if self.numberAzimuthLooks is None:
            self.numberAzimuthLooks = 1
        if self.lookSide is None:
            self.lookSide = -1
        if self.outputPrecision is None:
            self.outputPrecision = 'single'

    # ---- Geo2rdr.py:299-318 ----
    def destroyImages(self):
        for outfile in [self.rangeImage, self.azimuthImage, self.rangeOffsetImage, self.azimuthOffsetImage]:
            if outfile is not None:
                outfile.finalizeImage()
                outfile.renderHdr()
        self.polyDopplerAccessor = None
        for img in (self.latImage, self.lonImage, self.demImage):
            if not isinstance(img, Poly2D) and hasattr(img, "finalizeImage"):
                img.finalizeImage()

    # ---- Geo2rdr.py:320-387 ----
    def _out_image(self, filename, what):
        img = IF.createImage()
        img.setFilename(filename)
        img.setAccessMode('write')
        if self.outputPrecision.upper() == 'SINGLE':
            img.setDataType('FLOAT')
            img.setCaster('write', 'DOUBLE')
        elif self.outputPrecision.upper() == 'DOUBLE':
            img.setDataType('DOUBLE')
        else:
            raise Exception('Undefined output precision for {0} image in geo2rdr.'.format(what))
        img.setWidth(self.demWidth)
        img.setLength(self.demLength)
        img.createImage()
        return img

    def createImages(self):
        if self.rangeImageName:
            self.rangeImage = self._out_image(self.rangeImageName, 'range')
        if self.rangeOffsetImageName:
            self.rangeOffsetImage = self._out_image(self.rangeOffsetImageName, 'range offset')
        if self.azimuthImageName:
            self.azimuthImage = self._out_image(self.azimuthImageName, 'azimuth')
        if self.azimuthOffsetImageName:
            self.azimuthOffsetImage = self._out_image(self.azimuthOffsetImageName, 'azimuth offset')
        self.polyDopplerAccessor = 0

    # ---- setters Geo2rdr.py:408-455 ----
    def setEllipsoidMajorSemiAxis(self, var): self.ellipsoidMajorSemiAxis = float(var)
    def setEllipsoidEccentricitySquared(self, var): self.ellipsoidEccentricitySquared = float(var)
    def setRangePixelSpacing(self, var): self.slantRangePixelSpacing = float(var)
    def setRangeFirstSample(self, var): self.rangeFirstSample = float(var)
    def setPRF(self, var): self.prf = float(var)
    def setRadarWavelength(self, var): self.radarWavelength = float(var)
    def setSensingStart(self, var): self.sensingStart = var
    def setLength(self, var): self.length = int(var)
    def setWidth(self, var): self.width = int(var)
    def setNumberRangeLooks(self, var): self.numberRangeLooks = int(var)
    def setNumberAzimuthLooks(self, var): self.numberAzimuthLooks = int(var)
    def setDemWidth(self, var): self.demWidth = int(var)
    def setDemLength(self, var): self.demLength = int(var)
    def setLookSide(self, var): self.lookSide = int(var)
    def setOrbit(self, var): self.orbit = var
    def setPolyDoppler(self, var): self.polyDoppler = var

    # ---- ports Geo2rdr.py:457-503 ----
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
                self.lookSide = instrument.getPlatform().pointingDirection
                self.slantRangePixelSpacing = instrument.getRangePixelSize()
                self.prf = instrument.getPulseRepetitionFrequency()
                self.radarWavelength = instrument.getRadarWavelength()
            except AttributeError as strerr:
                self.logger.error(strerr)
                raise AttributeError

    def addDem(self):
        dem = self._inputPorts.getPort(name='dem').getObject()
        if (dem):
            try:
                self.demImage = dem
                self.demWidth = dem.getWidth()
                self.demLength = dem.getLength()
            except AttributeError as strerr:
                self.logger.error(strerr)
                raise AttributeError

    def addRadarImage(self):
        ifg = self._inputPorts.getPort(name='radarImage').getObject()
        if (ifg):
            try:
                self.inputImage = ifg
                self.width = ifg.getWidth()
                self.length = ifg.getLength()
            except AttributeError as strerr:
                self.logger.error(strerr)
                raise AttributeError

    # ---- Geo2rdr.py:506-557 ----
    def __init__(self, name=''):
        super(Geo2rdr, self).__init__(self.__class__.family, name)
        self.latImage = None
        self.lonImage = None
        self.demImage = None
        self.demWidth = None
        self.demLength = None
        self.rangeImageName = None
        self.rangeImage = None
        self.azimuthImageName = None
        self.azimuthImage = None
        self.rangeOffsetImageName = None
        self.rangeOffsetImage = None
        self.azimuthOffsetImageName = None
        self.azimuthOffsetImage = None
        self.length = None
        self.width = None
        self.polyDoppler = None
        self.polyDopplerAccessor = None
        self.dopplerCentroidCoeffs = None
        self.fmrateCoeffs = None
        self.orbit = None
        self.bistaticDelayCorrectionFlag = None
        self.dictionaryOfOutputVariables = {}
        self.gpuDevices = None  # B200 extension: CUDA device ordinals to shard the lat/lon/hgt lines over
        self.gpuTimings = None
        return None

    def createPorts(self):
        from .component import Port
        self._inputPorts.add(Port(name='frame', method=self.addFrame))
        self._inputPorts.add(Port(name='planet', method=self.addPlanet))
        self._inputPorts.add(Port(name='dem', method=self.addDem))
        self._inputPorts.add(Port(name='radarImage', method=self.addRadarImage))
        return None


def createGeo2rdr(name=''):
    """components/zerodop/geo2rdr/__init__.py:3-5"""
    return Geo2rdr(name=name)
