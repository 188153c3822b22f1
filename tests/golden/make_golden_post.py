#!/usr/bin/env python
"""Generate tests/golden/ref_toradar.npz by running the REFERENCE's own ``SWBDStitcher.toRadar``
(contrib/demUtils/swbdstitcher/SWBDStitcher.py:107-131) on a small synthetic mask / lat / lon triple.

Run only in the build container (needs /root/reference).  The function is taken from the reference file as it is (its
source text is extracted with ``ast`` and executed; nothing is copied into the repo); the only stand-in is
``createImage``, for which this repo's metadata class is used -- toRadar touches nothing of it but load(), the
coordinate records, toNumpyDataType(), initImage() and renderHdr().
"""
import ast
import os
import sys
import tempfile

import numpy as np

REF = "/root/reference/contrib/demUtils/swbdstitcher/SWBDStitcher.py"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from isce2_b200 import image as IF  # noqa: E402


def reference_toRadar():
    src = open(REF).read()
    tree = ast.parse(src)
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "SWBDStitcher")
    fn = next(n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == "toRadar")
    fn.decorator_list = []
    mod = ast.Module(body=[fn], type_ignores=[])
    ns = {"np": np, "createImage": IF.createImage}
    exec(compile(mod, REF, "exec"), ns)
    return ns["toRadar"]


def write(path, arr, dtype, start=(0.0, 0.0), delta=(1.0, 1.0)):
    arr.tofile(path)
    im = IF.createImage()
    im.initImage(path, "read", arr.shape[1], dtype)
    im.setLength(arr.shape[0])
    im.coord1.coordStart, im.coord1.coordDelta = start[1], delta[1]
    im.coord2.coordStart, im.coord2.coordDelta = start[0], delta[0]
    im.coord1.coordSize, im.coord2.coordSize = arr.shape[1], arr.shape[0]
    im.renderHdr()


def main():
    rng = np.random.default_rng(20261017)
    ml, mw = 97, 131
    mask = rng.integers(-1, 2, (ml, mw)).astype(np.int8)  # SWBD convention: -1 water, 0 land (+ a few 1's)
    start_lat, start_lon, d = 35.4, -118.3, 1.0 / 3600.0
    L, W = 40, 57
    # radar pixels over the mask, some of them outside it on every side (clipped by the reference)
    lat = start_lat - d * rng.uniform(-6.0, ml + 6.0, (L, W))
    lon = start_lon + d * rng.uniform(-6.0, mw + 6.0, (L, W))
    toRadar = reference_toRadar()
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        for tag, dt, isce_dt in (("f64", np.float64, "DOUBLE"), ("f32", np.float32, "FLOAT")):
            write(os.path.join(tmp, "mask.msk"), mask, "BYTE", (start_lat, start_lon), (-d, d))
            write(os.path.join(tmp, "lat.rdr"), lat.astype(dt), isce_dt)
            write(os.path.join(tmp, "lon.rdr"), lon.astype(dt), isce_dt)
            toRadar(os.path.join(tmp, "mask.msk"), os.path.join(tmp, "lat.rdr"), os.path.join(tmp, "lon.rdr"),
                    os.path.join(tmp, "waterMask.rdr"))
            out["out_" + tag] = np.fromfile(os.path.join(tmp, "waterMask.rdr"), np.int8).reshape(L, W)
    np.savez_compressed(os.path.join(HERE, "ref_toradar.npz"), mask=mask, lat=lat, lon=lon, start_lat=start_lat,
                        start_lon=start_lon, delta=d, numpy_version=np.__version__, **out)
    print("wrote ref_toradar.npz", {k: v.shape for k, v in out.items()}, "numpy", np.__version__)


if __name__ == "__main__":
    main()
