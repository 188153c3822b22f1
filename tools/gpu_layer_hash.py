#!/usr/bin/env python
"""SHA-1 of every topo layer of a rough test scene (run on the GPU box): A/B check that an experimental build
(B200GEOM_LIB=...) leaves the outputs bit-identical."""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import parity_util as pu  # noqa: E402

out = {"lib": os.environ.get("B200GEOM_LIB", "default")}
for method in ("BIQUINTIC", "BILINEAR"):
    # the DEM edge cuts through the scene's far range: exercises the edge fall-backs of the interpolators
    sc = pu.rough_scene(192, 8192)
    g = pu.gpu_topo(sc, dem_method=method)
    out[method] = {k: hashlib.sha1(g[k].tobytes()).hexdigest()[:12] for k in ("lat", "lon", "hgt", "los", "inc", "mask")}
print(json.dumps(out))
