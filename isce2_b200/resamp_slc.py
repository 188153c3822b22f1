"""Drop-in for ``stdproc.createResamp_slc()``: ``Resamp_slc`` with the attribute / port / method surface of
components/stdproc/stdproc/resamp_slc/Resamp_slc.py (:47-397), driving the B200 CUDA library (the
consumer of geo2rdr's ``range.off`` / ``azimuth.off`` rasters, which arrive here as the residual offset images).

Complex SLCs with sinc interpolation, the only branch the reference implements (resamp_slc.f90:69-74, :270-274).
"""
from __future__ import annotations

import logging

import numpy as np

from . import _capi, image as IF
from .component import Component, Port
from .poly import Poly2D, poly2d_fields


class Resamp_slc(Component):
    interpolationMethods = {'SINC': 0, 'BILINEAR': 1, 'BICUBIC': 2, 'NEAREST': 3, 'AKIMA': 4, 'BIQUINTIC': 5}

    # ---- Resamp_slc.py:56-84 ----
    def resamp_slc(self, imageIn=None, imageOut=None):
        for port in self.inputPorts:
            port()
        if imageIn is not None:
            self.imageIn = imageIn
        if self.imageIn is None:
            self.logger.error("Input slc image not set.")
            raise Exception
        if imageOut is not None:
            self.imageOut = imageOut
        if self.imageOut is None:
            self.logger.error("Output slc image not set.")
            raise Exception
        self.setDefaults()
        self.createImages()
        self._run()
        self.destroyImages()
        return

    def _poly(self, p):
        if p is None:  # the zero polynomial of Resamp_slc.py:86-140
            return None
        return poly2d_fields(p)

    def _run(self):
        if not self.isComplex:
            raise Exception('Real data interpolation not implemented yet.')  # resamp_slc.f90:272
        if self.method.upper() != 'SINC':
            # resamp_slc.f90:69-74
            self.logger.warning('Currently Only Sinc interpolation is available for complex data. '
                                'Setting interpolation method to sinc')
            self.method = 'SINC'
        slc = IF.read_raster(self.imageIn)
        if slc.dtype != np.complex64:
            slc = np.asarray(slc).astype(np.complex64)
        ol, ow = int(self.outputLines), int(self.outputWidth)

        def resid(img):
            if img is None:
                return None
            r = IF.read_raster(img)  # FLOAT .off rasters stay float32 on the way to the GPU: the widening to double
            if r.dtype not in (np.float32, np.float64):  # (the 'read' DOUBLE caster, :120-133) happens in the kernel
                r = np.asarray(r).astype(np.float64)
            return r[:ol]

        ra, rr = resid(self.residualAzimuthImage), resid(self.residualRangeImage)
        if ra is not None and rr is not None and ra.dtype != rr.dtype:
            ra, rr = np.asarray(ra, np.float64), np.asarray(rr, np.float64)
        out = IF.output_memmap(self.imageOut, ol, ow)  # ours, or an image object the caller handed in
        direct = out.dtype == np.complex64 and out.flags['C_CONTIGUOUS'] and out.shape == (ol, ow)
        with IF.file_backed([out] if direct else [], inputs=[slc, ra, rr]):  # rasters in, raster out (image.file_backed)
            r = _capi.resamp_slc_run(slc[:int(self.inputLines)], (ol, ow), wvl=float(self.radarWavelength),
                                     slr=float(self.slantRangePixelSpacing), r0=float(self.startingRange),
                                     ref_wvl=float(self.referenceWavelength), ref_r0=float(self.referenceStartingRange),
                                     ref_slr=float(self.referenceSlantRangePixelSpacing), flatten=bool(self.flatten),
                                     rg_carrier=self._poly(self.rangeCarrierPoly), az_carrier=self._poly(self.azimuthCarrierPoly),
                                     rg_offsets=self._poly(self.rangeOffsetsPoly), az_offsets=self._poly(self.azimuthOffsetsPoly),
                                     doppler=self._poly(self.dopplerPoly), resid_az=ra, resid_rg=rr, out=out if direct else None,
                                     device=int(self.gpuDevice or 0))
        if not direct:
            out[...] = r['slc'].reshape(out.shape)
        self.numValid = r['num_valid']
        self.gpuTimings = {k: r[k] for k in ('ms_kernels', 'ms_total', 'gpu_launches')}

    # ---- Resamp_slc.py:86-160 ----
    def createImages(self):
        if getattr(self.imageIn, '_mmap', None) is None and hasattr(self.imageIn, 'createImage'):
            self.imageIn.createImage()
        if getattr(self.imageOut, 'length', None) in (None, 0):
            self.imageOut.setLength(int(self.outputLines))
        if getattr(self.imageOut, '_mmap', None) is None:
            self.imageOut.createImage()
        for name, poly in (('Range Carrier', self.rangeCarrierPoly), ('Azimuth Carrier', self.azimuthCarrierPoly)):
            if poly is None:
                print('No {0} provided.'.format(name))
                print('Assuming zero {0}.'.format(name.lower()))
        if self.rangeOffsetsPoly is None:
            print('No range offset polynomial provided')
        if self.azimuthOffsetsPoly is None:
            print('No azimuth offset polynomial provided')
        for img in (self.residualRangeImage, self.residualAzimuthImage):
            if img is not None and getattr(img, '_mmap', None) is None and hasattr(img, 'createImage'):
                if hasattr(img, 'setCaster'):
                    img.setCaster('read', 'DOUBLE')
                img.createImage()
        if self.dopplerPoly is None:
            print('No doppler polynomial provided')
            print('Assuming zero doppler centroid')

    # ---- Resamp_slc.py:162-176 ----
    def destroyImages(self):
        if self.residualRangeImage is not None:
            self.residualRangeImage.finalizeImage()
        if self.residualAzimuthImage is not None:
            self.residualAzimuthImage.finalizeImage()
        self.imageIn.finalizeImage()
        self.imageOut.finalizeImage()
        return

    # ---- Resamp_slc.py:178-233 ----
    def setDefaults(self):
        if self.inputLines is None:
            self.inputLines = self.imageIn.getLength()
            self.logger.warning('The variable INPUT_LINES has been set to the default value %d which is the number of lines in the slc image.' % (self.inputLines))
        if self.inputWidth is None:
            self.inputWidth = self.imageIn.getWidth()
            self.logger.warning('The variable INPUT_WIDTH has been set to the default value %d which is the width of the slc image.' % (self.inputWidth))
        if self.inputWidth != self.imageIn.getWidth():
            raise Exception('Width of input image {0} does not match specified width {1}'.format(self.imageIn.getWidth(), self.inputWidth))
        if self.startingRange is None:
            self.startingRange = 0.0
        if self.referenceStartingRange is None:
            self.referenceStartingRange = self.startingRange
        if self.referenceSlantRangePixelSpacing is None:
            self.referenceSlantRangePixelSpacing = self.slantRangePixelSpacing
        if self.referenceWavelength is None:
            self.referenceWavelength = self.radarWavelength
        if self.outputLines is None:
            self.outputLines = self.imageOut.getLength()
            self.logger.warning('The variable OUTPUT_LINES has been set to the default value %d which is the number of lines in the slc image.' % (self.outputLines))
        if self.outputWidth is None:
            self.outputWidth = self.imageOut.getWidth()
            self.logger.warning('The variable OUTPUT_WIDTH has been set to the default value %d which is the width of the slc image.' % (self.outputWidth))
        if (self.outputWidth != self.imageOut.getWidth()):
            raise Exception('Width of output image {0} does not match specified width {1}'.format(self.imageOut.getWidth(), self.outputWidth))
        if self.imageIn.dataType.upper().startswith('C'):
            self.isComplex = True
        else:
            self.isComplex = False
        if self.imageIn.getBands() > 1:
            raise Exception('The code currently is setup to resample single band images only')
        if self.method is None:
            if self.isComplex:
                self.method = 'SINC'
            else:
                self.method = 'BILINEAR'
        if self.flatten is None:
            self.logger.warning('No flattening requested')
            self.flatten = False
        return

    # ---- setters Resamp_slc.py:258-280 ----
    def setInputWidth(self, var): self.inputWidth = int(var)
    def setInputLines(self, var): self.inputLines = int(var)
    def setOutputWidth(self, var): self.outputWidth = int(var)
    def setOutputLines(self, var): self.outputLines = int(var)
    def setRadarWavelength(self, var): self.radarWavelength = float(var)
    def setSlantRangePixelSpacing(self, var): self.slantRangePixelSpacing = float(var)

    def __getstate__(self):
        d = dict(self.__dict__)
        del d['logger']
        return d

    def __setstate__(self, d):
        self.__dict__.update(d)
        self.logger = logging.getLogger('isce.stdproc.resamp_slc')
        return

    # ---- ports Resamp_slc.py:293-332 ----
    def addOffsets(self):
        offsets = self._inputPorts['offsets']
        if offsets:
            polys = offsets.getFitPolynomials()
            self.azimuthOffsetsPoly = polys[0]
            self.rangeOffsetsPoly = polys[1]

    def addSlc(self):
        SPEED_OF_LIGHT = 299792458.0
        formslc = self._inputPorts['slc']
        if (formslc):
            coeffs = []
            coeffs.append([2 * np.pi * val for val in formslc.dopplerCentroidCoefficients])
            self.dopplerPoly = Poly2D()
            self.dopplerPoly.initPoly(rangeOrder=len(formslc.dopplerCentroidCoefficients) - 1, azimuthOrder=0, coeffs=coeffs)
            delr = 0.5 * SPEED_OF_LIGHT / formslc.rangeSamplingRate
            self.slantRangePixelSpacing = delr
            self.radarWavelength = formslc.radarWavelength
            src = formslc.slcImage
            img = IF.createImage()
            for a in ('filename', 'width', 'length', 'bands', 'dataType', 'scheme', 'byteOrder', 'imageType'):
                if hasattr(src, a):
                    setattr(img, a, getattr(src, a))
            img.setAccessMode('read')
            self.imageIn = img

    def addReferenceImage(self):
        refImg = self._inputPorts['reference']
        if (refImg):
            self.outputWidth = refImg.getWidth()
            self.outputLines = refImg.getLength()

    # ---- Resamp_slc.py:334-397 ----
    def __init__(self):
        Component.__init__(self)
        self.inputWidth = None
        self.inputLines = None
        self.outputWidth = None
        self.outputLines = None
        self.radarWavelength = None
        self.slantRangePixelSpacing = None
        self.azimuthOffsetsPoly = None
        self.azimuthOffsetsAccessor = None
        self.rangeOffsetsPoly = None
        self.rangeOffsetsAccessor = None
        self.rangeCarrierPoly = None
        self.rangeCarrierAccessor = None
        self.azimuthCarrierPoly = None
        self.azimuthCarrierAccessor = None
        self.residualRangeImage = None
        self.residualAzimuthImage = None
        self.residualRangeAccessor = None
        self.residualAzimuthAccessor = None
        self.dopplerPoly = None
        self.dopplerAccessor = None
        self.isComplex = None
        self.method = None
        self.flatten = None
        self.startingRange = None
        self.referenceWavelength = None
        self.referenceStartingRange = None
        self.referenceSlantRangePixelSpacing = None
        self.imageIn = None
        self.imageOut = None
        self.logger = logging.getLogger('isce.stdproc.resamp_slc')
        self._inputPorts.add(Port(name='offsets', method=self.addOffsets))
        self._inputPorts.add(Port(name='slc', method=self.addSlc))
        self._inputPorts.add(Port(name='reference', method=self.addReferenceImage))
        self.dictionaryOfVariables = {
            'INPUT_WIDTH': ['self.inputWidth', 'int', 'mandatory'],
            'INPUT_LINES': ['self.inputLines', 'int', 'optional'],
            'OUTPUT_LINES': ['self.outputLines', 'int', 'optional'],
            'OUTPUT_WIDTH': ['self.outputWidth', 'int', 'optional'],
            'RADAR_WAVELENGTH': ['self.radarWavelength', 'float', 'mandatory'],
            'SLANT_RANGE_PIXEL_SPACING': ['self.slantRangePixelSpacing', 'float', 'mandatory'],
        }
        self.dictionaryOfOutputVariables = {}
        self.gpuDevice = None  # B200 extension: CUDA device ordinal (default 0)
        self.gpuTimings = None
        self.numValid = None
        return


def createResamp_slc():
    """components/stdproc/stdproc/resamp_slc/__init__.py"""
    return Resamp_slc()
