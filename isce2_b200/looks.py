"""Multilooking of the geometry layers on the GPU -- SURVEY 8(f) row N4, "runMultilook of geometry layers".

``Looks`` mirrors components/mroipac/looks/Looks.py (class Looks :36-113: setDownLooks / setAcrossLooks / setInputImage /
setOutputFilename / looks()) and drives b200_looks_run instead of the ``looks`` extension module;
``runMultilook`` mirrors contrib/stack/stripmapStack/topo.py:365-441 (same arguments, same files written), with
method='isce' -> box mean of mroipac.looks, method='gdal' -> nearest-neighbour decimation of gdal.Translate -outsize.
No CPU fallback: without a CUDA device the calls raise.
"""
from __future__ import annotations

import os
import shutil

import numpy as np

from . import _capi, image as IF
from .component import Component


class Looks(Component):
    family = "looks"
    logging_name = "isce.mroipac.looks"

    def __init__(self):
        super().__init__()
        self.acrossLooks = None
        self.downLooks = None
        self.inputImage = None
        self.outputFilename = None
        self.method = "AVERAGE"  # B200 extension: 'NEAREST' gives the gdal.Translate decimation of runMultilook
        self.gpuDevice = 0
        self.gpuTimings = None

    def setInputImage(self, var): self.inputImage = var
    def setAcrossLooks(self, var): self.acrossLooks = int(var)
    def setDownLooks(self, var): self.downLooks = int(var)
    def setOutputFilename(self, var): self.outputFilename = str(var)

    # ---- Looks.py:36-85 ----
    def looks(self):
        inImage = self.inputImage.clone()
        inImage.setAccessMode('READ')
        inImage.createImage()
        outWidth = inImage.getWidth() // self.acrossLooks
        outLength = inImage.getLength() // self.downLooks

        outImage = self.inputImage.clone()
        # if the image is not a geo the part below does not matter (Looks.py:47-58)
        try:
            outImage.coord1.coordDelta = self.inputImage.coord1.coordDelta * self.acrossLooks
            outImage.coord2.coordDelta = self.inputImage.coord2.coordDelta * self.downLooks
            outImage.coord1.coordStart = self.inputImage.coord1.coordStart + \
                0.5 * (self.acrossLooks - 1) * self.inputImage.coord1.coordDelta
            outImage.coord2.coordStart = self.inputImage.coord2.coordStart + \
                0.5 * (self.downLooks - 1) * self.inputImage.coord2.coordDelta
        except Exception:
            pass
        outImage.setWidth(outWidth)
        outImage.setLength(outLength)
        outImage.setFilename(self.outputFilename)
        outImage.setAccessMode('WRITE')
        if outLength > 0 and outWidth > 0:
            out = outImage.createImage()
            src = np.ascontiguousarray(inImage.memMap())
            res_arr, res = _capi.looks_run(src, self.downLooks, self.acrossLooks, scheme=inImage.scheme, method=self.method,
                                           device=self.gpuDevice)
            out[...] = res_arr.view(out.dtype)
            self.gpuTimings = {k: res[k] for k in ("ms_kernels", "ms_total", "gpu_launches")}
        else:  # fewer lines / samples than looks: the reference's loops do not execute and leave an empty raster
            os.makedirs(os.path.dirname(os.path.abspath(self.outputFilename)), exist_ok=True)
            open(self.outputFilename, 'wb').close()
        inImage.finalizeImage()
        outImage.finalizeImage()
        outImage.renderHdr()
        if hasattr(outImage, "renderVRT"):
            outImage.renderVRT()
        return outImage


def runMultilook(in_dir, out_dir, alks, rlks, in_ext='.rdr', out_ext='.rdr', method='gdal',
                 fbase_list=['hgt', 'incLocal', 'lat', 'lon', 'los', 'shadowMask', 'waterMask'], device=0):
    """contrib/stack/stripmapStack/topo.py:365-441."""
    msg = 'generate multilooked geometry files with alks={} and rlks={}'.format(alks, rlks)
    if method == 'isce':
        msg += ' using the box mean of mroipac.looks.Looks() on the GPU ...'
    elif method == 'gdal':
        msg += ' using the nearest-neighbour decimation of gdal.Translate() on the GPU ...'
    else:
        raise ValueError('un-supported multilook method: {}'.format(method))
    print('-' * 50 + '\n' + msg)
    os.makedirs(out_dir, exist_ok=True)
    for fbase in fbase_list:
        in_file = os.path.join(in_dir, '{}{}'.format(fbase, in_ext))
        out_file = os.path.join(out_dir, '{}{}'.format(fbase, out_ext))
        if all(os.path.isfile(in_file + ext) for ext in ['', '.vrt', '.xml']):
            print('multilook {}'.format(in_file))
            inImage = IF.createImage()
            inImage.load(in_file + '.xml')
            inImage.filename = in_file
            lkObj = Looks()
            lkObj.setDownLooks(alks)
            lkObj.setAcrossLooks(rlks)
            lkObj.setInputImage(inImage)
            lkObj.setOutputFilename(out_file)
            lkObj.method = 'AVERAGE' if method == 'isce' else 'NEAREST'
            lkObj.gpuDevice = device
            lkObj.looks()
            # the full-resolution xml / vrt beside the multilooked file, to recover the number of looks (:434-439)
            if in_file != out_file + '.full':
                shutil.copy(in_file + '.xml', out_file + '.full.xml')
                shutil.copy(in_file + '.vrt', out_file + '.full.vrt')
    return out_dir
