"""The per-pixel device functions (isce2_b200/csrc/*.cuh), compiled for the host by g++, must reproduce the
oracle bit for bit: same operations in the same order, no fused multiply-adds (the CUDA build uses
-fmad=false).  This is a development aid for a container without a GPU; the product never loads it."""
import numpy as np
import pytest

from isce2_b200 import synth
from oracle import oracle as orc
from tests import parity_util as pu
from tests.emu import emu


@pytest.mark.parametrize("name,mid", [("BILINEAR", 1), ("BICUBIC", 2), ("NEAREST", 3), ("BIQUINTIC", 5)])
def test_pixel_functions_bit_exact(name, mid):
    sc = pu.rough_scene(12, 2048, dem_spacing_arcsec=1.0)
    o = orc.topo(**orc.scene_topo_kwargs(sc, dem_method=name, want_mask=False))
    e = emu.topo(sc, o, dem_method=mid)
    assert o["total_iters"] == e["iters"]
    for k in ("lat", "lon", "hgt", "los", "inc"):
        assert np.array_equal(o[k], e[k]), k


def test_native_doppler_and_left_looking_bit_exact():
    sc = synth.make_scene(8, 2048, sensor="nisar")
    o = orc.topo(**orc.scene_topo_kwargs(sc, dem_method="BIQUINTIC", orbit_method="LEGENDRE", want_mask=False))
    e = emu.topo(sc, o, dem_method=5, orbit_method=2)
    for k in ("lat", "lon", "hgt", "los", "inc"):
        assert np.array_equal(o[k], e[k]), k
