"""Drop-in for ``zerodop.geozero``: ``createGeozero()`` -> ``Geocode`` with the parameter / port / method surface of
components/zerodop/geozero/Geozero.py (class Geocode :157-630), driving the B200 CUDA library.

Visible differences: none in the interface.  Inside, the range-Doppler solve of the output grid runs once per
``geocode()`` call and is shared by all bands (the reference re-solves every band, Geozero.py:216-241); a caller
that geocodes several products onto the same grid (TopsProc/runGeocode.py loops over ~6 files) can keep the solved
grid on the GPU between files with ``_capi.GeozeroPlan``.
"""
from __future__ import annotations

import math
import os

import numpy as np

from . import _capi, image as IF
from .component import Component, Port
from .orbit import export_rows, seconds_since_midnight
from .planet import EarthEccentricitySquared, EarthMajorSemiAxis
from .poly import Poly1D, poly1d_fields

P = Component.Parameter

# Geozero.py:44-155
INTERPOLATION_METHOD = P('method', public_name='INTERPOLATION_METHOD', default=None, type=str, mandatory=True,
                         doc='Interpolation method. Can be sinc/ bilinear/ bicubic/ nearest')
MINIMUM_LATITUDE = P('minimumLatitude', public_name='MINIMUM_LATITUDE', default=None, type=float, mandatory=True,
                     doc='Minimum Latitude to geocode')
MAXIMUM_LATITUDE = P('maximumLatitude', public_name='MAXIMUM_LATITUDE', default=None, type=float, mandatory=True,
                     doc='Maximum Latitude to geocode')
MINIMUM_LONGITUDE = P('minimumLongitude', public_name='MINIMUM_LONGITUDE', default=None, type=float, mandatory=True,
                      doc='Minimum Longitude to geocode')
MAXIMUM_LONGITUDE = P('maximumLongitude', public_name='MAXIMUM_LONGITUDE', default=None, type=float, mandatory=True,
                      doc='Maximum Longitude to geocode')
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
DEM_CROP_FILENAME = P('demCropFilename', public_name='DEM_CROP_FILENAME', default=None, type=str, mandatory=True,
                      doc='Filename for the cropped DEM output')
GEO_FILENAME = P('geoFilename', public_name='GEO_FILENAME', default=None, type=str, mandatory=True,
                 doc='Output geocoded file name')
LOOK_SIDE = P('lookSide', public_name='LOOK_SIDE', default=None, type=int, mandatory=True,
              doc='Right (-1) / Left (1) . Look direction of the radar platform')


def _as_samples(img):
    """The raster of `img` in its own interleaving as float32 / complex64 samples (the 'read' FLOAT caster of
    Geozero.py:199-200 for every other data type), plus what is needed to write the result back."""
    arr = IF.read_raster(img)
    dtype = str(img.dataType).upper()
    is_complex = dtype.startswith('C')
    want = np.complex64 if is_complex else np.float32
    if arr.dtype != want:
        arr = np.asarray(arr).astype(want)
    return arr, is_complex


class Geocode(Component):
    interp_methods = {'sinc': 0, 'bilinear': 1, 'bicubic': 2, 'nearest': 3}

    family = 'geocode'
    logging_name = 'isce.zerodop.geocode'

    parameter_list = (INTERPOLATION_METHOD, MINIMUM_LATITUDE, MAXIMUM_LATITUDE, MINIMUM_LONGITUDE, MAXIMUM_LONGITUDE,
                      SLANT_RANGE_PIXEL_SPACING, ELLIPSOID_ECCENTRICITY_SQUARED, ELLIPSOID_MAJOR_SEMIAXIS, RANGE_FIRST_SAMPLE,
                      SENSING_START, NUMBER_RANGE_LOOKS, NUMBER_AZIMUTH_LOOKS, PRF, RADAR_WAVELENGTH, DEM_CROP_FILENAME,
                      GEO_FILENAME, LOOK_SIDE)

    # ---- Geozero.py:184-266 ----
    def geocode(self, demImage=None, inputImage=None, method=None):
        self.activateInputPorts()
        if demImage is not None:
            self.demImage = demImage
        if inputImage is not None:
            self.inputImage = inputImage
        if method is not None:
            self.method = method
        if self.orbit is None:
            raise Exception('No orbit provided for geocoding')
        self.setDefaults()
        if self.method is None or str(self.method).lower() not in self.interp_methods:
            raise KeyError(self.method)
        params = self._params()
        self.createImages(params)
        self._run(params)
        self.destroyImages()
        self.geoImage.setWidth(self.geoWidth)
        self.geoImage.trueDataType = self.geoImage.getDataType()
        self.geoImage.coord2.coordDescription = 'Latitude'
        self.geoImage.coord2.coordUnits = 'degree'
        self.geoImage.coord2.coordStart = self.maximumGeoLatitude
        self.geoImage.coord2.coordDelta = self.deltaLatitude
        self.geoImage.coord1.coordDescription = 'Longitude'
        self.geoImage.coord1.coordUnits = 'degree'
        self.geoImage.coord1.coordStart = self.minimumGeoLongitude
        self.geoImage.coord1.coordDelta = self.deltaLongitude
        descr = getattr(self.inputImage, 'description', None)
        if descr not in [None, '']:
            self.geoImage.addDescription(descr)
        self.geoImage.renderHdr()
        return None

    def _params(self):
        dem = self.demImage
        if self.demWidth is None:
            self.demWidth, self.demLength = dem.getWidth(), dem.getLength()
        if self.width is None:
            self.width, self.length = self.inputImage.getWidth(), self.inputImage.getLength()
        return _capi.geozero_params(dem_shape=(int(self.demLength), int(self.demWidth)), first_lat=float(self.firstLatitude),
                                    first_lon=float(self.firstLongitude), delta_lat=float(self.deltaLatitude),
                                    delta_lon=float(self.deltaLongitude),
                                    snwe=(float(self.minimumLatitude), float(self.maximumLatitude), float(self.minimumLongitude),
                                          float(self.maximumLongitude)),
                                    length=int(self.length), width=int(self.width), r0=float(self.rangeFirstSample),
                                    dr=float(self.slantRangePixelSpacing), prf=float(self.prf),
                                    t0=seconds_since_midnight(self.sensingStart), wvl=float(self.radarWavelength),
                                    side=int(self.lookSide), a=float(self.ellipsoidMajorSemiAxis),
                                    e2=float(self.ellipsoidEccentricitySquared), nrnglooks=int(self.numberRangeLooks),
                                    nazlooks=int(self.numberAzimuthLooks), device=int(self.gpuDevice or 0))

    def _run(self, params):
        dem = IF.read_raster(self.demImage)
        if dem.dtype not in (np.float32, np.int16):
            dem = np.asarray(dem).astype(np.float32)  # the 'read' FLOAT caster (Geozero.py:204)
        if dem.ndim != 2:
            raise Exception('DEM must be a single-band image')
        t, pos, vel = export_rows(self.orbit, self.sensingStart)  # Orbit.exportToC(reference=sensingStart) :211
        coeffs, mean, norm = poly1d_fields(self.polyDoppler)
        image, is_complex = _as_samples(self.inputImage)
        nbands = int(self.inputImage.getBands())
        scheme = str(self.inputImage.scheme).upper() if nbands > 1 else 'BIL'
        plan = _capi.GeozeroPlan(params, dem, t, pos, vel, coeffs, mean, norm)
        try:
            out_mm = self.geoImage.memMap()
            if out_mm.dtype == image.dtype and out_mm.flags['C_CONTIGUOUS']:
                with IF.file_backed([out_mm], inputs=[image]):  # both are mappings of rasters (image.file_backed)
                    plan.geocode(image, method=self.method, nbands=nbands, scheme=scheme, out=out_mm)
            else:  # the 'write' FLOAT caster (Geozero.py:313-314): the file keeps the input's data type
                tmp = plan.geocode(image, method=self.method, nbands=nbands, scheme=scheme)
                out_mm[...] = tmp.reshape(out_mm.shape).astype(out_mm.dtype)
            crop = self.demCropImage.memMap() if self.demCropImage is not None else None
            r = plan.fetch(dem_crop=crop if (crop is not None and crop.dtype == np.int16) else None)
            if crop is not None and crop.dtype != np.int16:
                crop[...] = r['dem_crop']
        finally:
            plan.close()
        # getState (Geozero.py:446-452)
        self.geoWidth = r['geo_width']
        self.geoLength = r['geo_length']
        self.minimumGeoLatitude = r['geo_min_lat']
        self.minimumGeoLongitude = r['geo_min_lon']
        self.maximumGeoLatitude = r['geo_max_lat']
        self.maximumGeoLongitude = r['geo_max_lon']
        self.numOutsideDEM = r['num_outside_dem']
        self.numOutsideImage = r['num_outside_image']
        self.numValid = r['num_valid']
        self.gpuTimings = {k: r[k] for k in ('ms_setup', 'ms_kernels', 'gpu_launches')}
        self.logger.info('Number of pixels with outside DEM: %d; outside the image: %d; with valid data: %d',
                         self.numOutsideDEM, self.numOutsideImage, self.numValid)

    # ---- Geozero.py:268-275 ----
    def setDefaults(self):
        if self.polyDoppler is None:
            self.polyDoppler = Poly1D(name=self.name + '_geozeroPoly')
            self.polyDoppler.setMean(0.0)
            self.polyDoppler.initPoly(order=len(self.dopplerCentroidCoeffs) - 1, coeffs=self.dopplerCentroidCoeffs)

    # ---- Geozero.py:277-289 ----
    def destroyImages(self):
        if self.demCropImage is not None:
            self.demCropImage.renderHdr()
            self.demCropImage.finalizeImage()
        self.geoImage.finalizeImage()
        self.polyDopplerAccessor = None

    # ---- Geozero.py:291-326 ----
    def createImages(self, params=None):
        geo_length, geo_width = _capi.geozero_grid(params if params is not None else self._params())
        if geo_length < 1 or geo_width < 1:
            raise ValueError('Empty geocoding grid: check the bounding box against the DEM')
        if self.demCropFilename:
            self.demCropImage = IF.createDemImage()
            self.demCropImage.initImage(self.demCropFilename, 'write', geo_width)
            self.demCropImage.setLength(geo_length)
            self.demCropImage.createImage()
            self.demCropAccessor = 0
        else:
            self.demCropImage = None
            self.demCropAccessor = 0
        if self.geoFilename is None:
            raise ValueError('Output geoFilename not specified')
        # the geocoded file has the format of the input: bands, interleaving, data type (IU.copyAttributes, :308-311)
        src = self.inputImage
        self.geoImage = IF.createImage()
        self.geoImage.bands = int(src.getBands())
        self.geoImage.scheme = str(src.scheme).upper()
        self.geoImage.dataType = str(src.dataType).upper()
        self.geoImage.imageType = src.imageType
        self.geoImage.byteOrder = getattr(src, 'byteOrder', 'l')
        self.geoImage.setFilename(self.geoFilename)
        self.geoImage.setAccessMode('write')
        self.geoImage.setWidth(geo_width)
        self.geoImage.setLength(geo_length)
        self.geoImage.createImage()
        self.geoAccessor = 0
        self.polyDopplerAccessor = 0

    # ---- Geozero.py:328-338 ----
    def computeGeoImageWidth(self):
        deg2rad = math.pi / 180.0
        dlon = self.deltaLongitude * deg2rad
        lon_first = self.firstLongitude * deg2rad
        min_lon = deg2rad * self.minimumLongitude
        max_lon = deg2rad * self.maximumLongitude
        min_lon_idx = int((min_lon - lon_first) / dlon)
        max_lon_idx = int((max_lon - lon_first) / dlon)
        geo_wid = max_lon_idx - min_lon_idx + 1
        return geo_wid

    # ---- setters / getters Geozero.py:366-476 ----
    def setMinimumLatitude(self, var): self.minimumLatitude = float(var)
    def setMinimumLongitude(self, var): self.minimumLongitude = float(var)
    def setMaximumLatitude(self, var): self.maximumLatitude = float(var)
    def setMaximumLongitude(self, var): self.maximumLongitude = float(var)
    def setEllipsoidMajorSemiAxis(self, var): self.ellipsoidMajorSemiAxis = float(var)
    def setEllipsoidEccentricitySquared(self, var): self.ellipsoidEccentricitySquared = float(var)
    def setRangePixelSpacing(self, var): self.slantRangePixelSpacing = float(var)
    def setRangeFirstSample(self, var): self.rangeFirstSample = float(var)
    def setPRF(self, var): self.prf = float(var)
    def setRadarWavelength(self, var): self.radarWavelength = float(var)
    def setSensingStart(self, var): self.sensingStart = var
    def setFirstLatitude(self, var): self.firstLatitude = float(var)
    def setFirstLongitude(self, var): self.firstLongitude = float(var)
    def setDeltaLatitude(self, var): self.deltaLatitude = float(var)
    def setDeltaLongitude(self, var): self.deltaLongitude = float(var)
    def setLength(self, var): self.length = int(var)
    def setWidth(self, var): self.width = int(var)
    def setNumberRangeLooks(self, var): self.numberRangeLooks = int(var)
    def setNumberAzimuthLooks(self, var): self.numberAzimuthLooks = int(var)
    def setDemWidth(self, var): self.demWidth = int(var)
    def setDemLength(self, var): self.demLength = int(var)
    def setLookSide(self, var): self.lookSide = int(var)
    def setOrbit(self, var): self.orbit = var
    def setDemCropFilename(self, var): self.demCropFilename = var
    def setPolyDoppler(self, var): self.polyDoppler = var
    def setGeocodeFilename(self, var): self.geoFilename = var
    def getGeoWidth(self): return self.geoWidth
    def getGeoLength(self): return self.geoLength
    def getLatitudeSpacing(self): return self.latitudeSpacing
    def getLongitudeSpacing(self): return self.longitudeSpacing
    def getMinimumGeoLatitude(self): return self.minimumGeoLatitude
    def getMinimumGeoLongitude(self): return self.minimumGeoLongitude
    def getMaximumGeoLatitude(self): return self.maximumGeoLatitude
    def getMaximumGeoLongitude(self): return self.maximumGeoLongitude

    # ---- ports Geozero.py:478-548 ----
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

    def addReferenceSlc(self):
        formslc = self._inputPorts.getPort(name='referenceslc').getObject()
        if (formslc):
            try:
                self.rangeFirstSample = formslc.startingRange
            except AttributeError as strerr:
                self.logger.error(strerr)
                raise AttributeError
            self.dopplerCentroidCoeffs = formslc.dopplerCentroidCoefficients

    def addDem(self):
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

    def addRadarImage(self):
        ifg = self._inputPorts.getPort(name='tobegeocoded').getObject()
        if (ifg):
            try:
                self.inputImage = ifg
                self.width = ifg.getWidth()
                self.length = ifg.getLength()
                inName = ifg.getFilename()
                self.geoFilename = os.path.join(os.path.dirname(inName), os.path.basename(inName) + '.geo')
                print('Output: ', self.geoFilename)
            except AttributeError as strerr:
                self.logger.error(strerr)
                raise AttributeError

    # ---- Geozero.py:553-566 ----
    @property
    def snwe(self):
        return (self.minimumLatitude, self.maximumLatitude, self.minimumLongitude, self.maximumLongitude)

    @snwe.setter
    def snwe(self, snwe):
        (self.minimumLatitude, self.maximumLatitude, self.minimumLongitude, self.maximumLongitude) = snwe

    # ---- Geozero.py:571-613 ----
    def __init__(self, name=''):
        super(Geocode, self).__init__(self.__class__.family, name)
        self.demImage = None
        self.demWidth = None
        self.demLength = None
        self.firstLatitude = None
        self.firstLongitude = None
        self.deltaLatitude = None
        self.deltaLongitude = None
        self.inputImage = None
        self.length = None
        self.width = None
        self.demCropImage = None
        self.demCropAccessor = None
        self.polyDoppler = None
        self.polyDopplerAccessor = None
        self.dopplerCentroidCoeffs = None
        self.geoImage = None
        self.geoAccessor = None
        self.geoWidth = None
        self.geoLength = None
        self.orbit = None
        self.latitudeSpacing = None
        self.longitudeSpacing = None
        self.minimumGeoLatitude = None
        self.minimumGeoLongitude = None
        self.maximumGeoLatitude = None
        self.maximumGeoLongitude = None
        self.dictionaryOfOutputVariables = {
            'GEO_WIDTH': 'self.geoWidth', 'GEO_LENGTH': 'self.geoLength', 'LATITUDE_SPACING': 'self.latitudeSpacing',
            'LONGITUDE_SPACING': 'self.longitudeSpacing', 'MINIMUM_GEO_LATITUDE': 'self.minimumGeoLatitude',
            'MINIMUM_GEO_LONGITUDE': 'self.minimumGeoLongitude', 'MAXIMUM_GEO_LATITUDE': 'self.maximumGeoLatitude',
            'MAXIMUM_GEO_LONGITUDE': 'self.maximumGeoLongitude'}
        self.gpuDevice = None  # B200 extension: CUDA device ordinal (default 0)
        self.gpuTimings = None
        return None

    # ---- Geozero.py:616-629 ----
    def createPorts(self):
        self._inputPorts.add(Port(name='frame', method=self.addFrame))
        self._inputPorts.add(Port(name='planet', method=self.addPlanet))
        self._inputPorts.add(Port(name='dem', method=self.addDem))
        self._inputPorts.add(Port(name='tobegeocoded', method=self.addRadarImage))
        self._inputPorts.add(Port(name='referenceslc', method=self.addReferenceSlc))
        return None


def createGeozero():
    """components/zerodop/geozero/__init__.py:3-5"""
    return Geocode()
