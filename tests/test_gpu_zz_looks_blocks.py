"""Multilooking of images that span several pipeline blocks in the band-sequential and pixel-interleaved layouts (one
host<->device piece per band plane and block).  Kept in a file that sorts last: the case was added after the round's last
full GPU pass (it has since been run on a B200 on its own), so under `pytest -x` a surprise here cannot hide the rest."""
import numpy as np
import pytest

from isce2_b200 import _capi
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape,scheme", [((2, 3000, 3000), "BSQ"), ((3000, 3000, 2), "BIP"), ((3000, 2, 3000), "BIL")])
def test_looks_over_several_blocks(shape, scheme):
    rng = np.random.default_rng(5)
    b = rng.normal(size=shape).astype(np.float32)
    gb, rb = _capi.looks_run(b, 3, 5, scheme=scheme)
    assert rb["gpu_launches"] > 1
    assert np.array_equal(gb, orc.looks(b, 3, 5, scheme=scheme)), scheme
    gn, _ = _capi.looks_run(b, 3, 5, scheme=scheme, method="NEAREST")
    assert np.array_equal(gn, orc.looks(b, 3, 5, scheme=scheme, method="NEAREST")), scheme
