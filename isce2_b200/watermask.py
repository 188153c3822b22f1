"""Projection of a geocoded (water body) mask into radar coordinates on the GPU -- SURVEY 8(f) row N4,
"water-mask projection".

``toRadar`` mirrors ``SWBDStitcher.toRadar`` (contrib/demUtils/swbdstitcher/SWBDStitcher.py:107-131), ``geo2radar`` the
wrapper of contrib/stack/stripmapStack/createWaterMask.py:66-71: same file arguments, same output raster + XML.  The
lat.rdr / lon.rdr rasters are the ones topozero wrote.  No CPU fallback.
"""
from __future__ import annotations

import numpy as np

from . import _capi, image as IF


def toRadar(maskin, latin, lonin, output, device=0):
    maskim = IF.createImage()
    maskim.load(maskin + '.xml')
    latim = IF.createImage()
    latim.load(latin + '.xml')
    lonim = IF.createImage()
    lonim.load(lonin + '.xml')
    mask = np.fromfile(maskin, maskim.toNumpyDataType())
    lat = np.fromfile(latin, latim.toNumpyDataType())
    lon = np.fromfile(lonin, lonim.toNumpyDataType())
    mask = np.reshape(mask, [maskim.coord2.coordSize, maskim.coord1.coordSize])
    startLat = maskim.coord2.coordStart
    deltaLat = maskim.coord2.coordDelta
    startLon = maskim.coord1.coordStart
    deltaLon = maskim.coord1.coordDelta
    # the mask starts from the top left corner: deltaLat < 0
    cropped, res = _capi.mask_to_radar_run(mask, startLat, deltaLat, startLon, deltaLon, lat, lon, device=device)
    cropped = np.reshape(cropped, (latim.coord2.coordSize, latim.coord1.coordSize))
    cropped.tofile(output)
    croppedim = IF.createImage()
    croppedim.initImage(output, 'read', cropped.shape[1], maskim.dataType)
    croppedim.setLength(cropped.shape[0])
    croppedim.renderHdr()
    if hasattr(croppedim, "renderVRT"):
        croppedim.renderVRT()
    return res


def geo2radar(geo_file, rdr_file, lat_file, lon_file, device=0):
    """stripmapStack/createWaterMask.py:66-71"""
    print('converting water mask file to radar coordinates ...')
    toRadar(geo_file, lat_file, lon_file, rdr_file, device=device)
    return rdr_file
