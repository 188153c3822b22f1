"""resamp_slc on the GPU against the CPU oracle, raw C-ABI calls and the drop-in Component fed with the
.off rasters geo2rdr writes.

Tolerance: the arithmetic is single-precision COMPLEX with double-precision phases reduced to float32 sines / cosines,
so the GPU and the oracle agree bit for bit except where the last bit of a double sin / cos (device libm vs glibc) flips a
float32 rounding: at most 1e-3 of the samples may differ, and then by no more than 4 float32 ulps of the sample's
magnitude.  Validity (zero fill outside the input image) is bit-exact."""
import datetime
import os

import numpy as np
import pytest

import isce2_b200
from isce2_b200 import _capi, image as IF, synth
from isce2_b200.orbit import Orbit
from isce2_b200.poly import Poly2D
from oracle import oracle as orc
from tests import parity_util as pu

pytestmark = pytest.mark.gpu


def _slc(L, W, seed=0):
    rng = np.random.default_rng(seed)
    return ((rng.normal(size=(L, W)) + 1j * rng.normal(size=(L, W))) * 37.0).astype(np.complex64)


def _compare(g, o):
    assert g.shape == o.shape and g.dtype == o.dtype == np.complex64
    assert np.array_equal(g == 0, o == 0)
    d = np.abs(g.astype(np.complex128) - o.astype(np.complex128))
    mag = np.maximum(np.abs(o.astype(np.complex128)), 1e-30)
    nbad = int((d != 0).sum())
    assert nbad <= max(2, int(1e-3 * d.size)), (nbad, d.size)
    assert float((d / mag).max()) <= 4 * 1.2e-7, float((d / mag).max())
    return nbad


def test_identity_shift_and_invalid_offsets():
    z = _slc(300, 700)
    r = _capi.resamp_slc_run(z, z.shape)
    assert np.array_equal(r["slc"], orc.resamp_slc(slc=z, out_shape=z.shape))
    assert np.array_equal(r["slc"][4:-5, 4:-5], z[4:-5, 4:-5]) and r["num_valid"] == (300 - 9) * (700 - 9)
    assert r["gpu_launches"] == 1  # zero carriers: the up-front pass is the identity and is skipped
    rr = np.zeros(z.shape, np.float32)
    rr[100:110] = -999999.0
    q = _capi.resamp_slc_run(z, z.shape, resid_rg=rr)
    assert not q["slc"][100:110].any() and np.array_equal(q["slc"], orc.resamp_slc(slc=z, out_shape=z.shape, resid_rg=rr))


@pytest.mark.parametrize("resid_dtype", [np.float32, np.float64])
def test_resamp_parity_offsets_carriers_doppler_flatten(resid_dtype):
    L, W, OL, OW = 500, 1800, 460, 1700
    z = _slc(L, W, 1)
    y, x = np.mgrid[0:OL, 0:OW].astype(np.float64)
    ra = (3.3 + 0.002 * y + 0.5 * np.sin(x / 130.0)).astype(resid_dtype)
    rr = (-7.6 + 0.001 * x + 0.4 * np.cos(y / 90.0)).astype(resid_dtype)
    kw = dict(out_shape=(OL, OW), wvl=0.0555, slr=2.33, r0=800e3, ref_wvl=0.0555, ref_r0=800.1e3, ref_slr=2.33, flatten=True,
              rg_carrier=[[0.0, 0.0003]], az_carrier=[[0.0, 0.9], [2.3, 1e-4], [0.021, 0.0]],
              rg_offsets=([[0.2, 0.004], [0.001, 0.0]], 5.0, 2.0, 100.0, 50.0), az_offsets=[[1.5, -0.0005]], doppler=[[1.7, 0.004]],
              resid_az=ra, resid_rg=rr)
    g = _capi.resamp_slc_run(z, **kw)
    okw = dict(kw)
    okw["rg_offsets"] = orc.Poly2D([[0.2, 0.004], [0.001, 0.0]], 5.0, 2.0, 100.0, 50.0)
    o = orc.resamp_slc(slc=z, **okw)
    assert g["gpu_launches"] == 2 and g["num_valid"] == int((o != 0).sum()) > 0.9 * OL * OW
    _compare(g["slc"], o)


def test_resamp_component_with_geo2rdr_offset_rasters(tmp_path):
    # reference geometry -> geo2rdr against a perturbed secondary orbit -> range.off / azimuth.off (FLOAT) -> resample the
    # secondary SLC onto the reference grid: the call sequence of contrib/stack/topsStack/resamp_withCarrier.py:57-104
    sc = pu.rough_scene(260, 2000)
    c = pu.cpu_topo(sc, dem_method="BILINEAR", want_inc=False, want_mask=False)
    sec = synth.config_c1_secondary(length=sc.length, width=sc.width)
    kw = pu.secondary_kwargs(sc, sec, recenter=0.37)
    off = tmp_path / "offsets"
    grdr = isce2_b200.createGeo2rdr()
    grdr.configure()
    grdr.slantRangePixelSpacing, grdr.prf, grdr.radarWavelength = sc.dr, sc.prf, sc.wvl
    day = sc.sensing_start.replace(hour=0, minute=0, second=0, microsecond=0)
    grdr.orbit = Orbit.from_arrays(day, sec.orbit_t, sec.orbit_pos, sec.orbit_vel)
    grdr.width, grdr.length = sc.width, sc.length
    grdr.lookSide = sc.side
    grdr.setSensingStart(day + datetime.timedelta(seconds=kw["t0"]))
    grdr.rangeFirstSample = kw["r0"]
    grdr.numberRangeLooks = grdr.numberAzimuthLooks = 1
    grdr.dopplerCentroidCoeffs = [0.]
    grdr.rangeOffsetImageName = str(off / "range.off")
    grdr.azimuthOffsetImageName = str(off / "azimuth.off")
    imgs = {}
    for name, arr in (("lat", c["lat"]), ("lon", c["lon"]), ("hgt", c["hgt"])):
        path = str(tmp_path / (name + ".rdr"))
        arr.tofile(path)
        img = IF.createImage()
        img.initImage(path, "read", sc.width, "DOUBLE")
        img.setLength(sc.length)
        img.renderHdr()
        imgs[name] = img
    grdr.demImage, grdr.latImage, grdr.lonImage = imgs["hgt"], imgs["lat"], imgs["lon"]
    grdr.geo2rdr()
    rg = np.fromfile(off / "range.off", np.float32).reshape(sc.length, sc.width)
    az = np.fromfile(off / "azimuth.off", np.float32).reshape(sc.length, sc.width)
    assert (rg != np.float32(-999999.0)).mean() > 0.5

    slc = _slc(sc.length, sc.width, 9)
    slc_path = str(tmp_path / "secondary.slc")
    slc.tofile(slc_path)
    inimg = IF.createImage()
    inimg.initImage(slc_path, "read", sc.width, "CFLOAT")
    inimg.setLength(sc.length)
    inimg.renderHdr()
    azcarr = Poly2D()
    azcarr.initPoly(rangeOrder=1, azimuthOrder=2, coeffs=[[0.0, 1e-4], [0.3, 0.0], [2e-4, 0.0]])
    rObj = isce2_b200.createResamp_slc()
    rObj.slantRangePixelSpacing = sc.dr
    rObj.radarWavelength = sc.wvl
    rObj.azimuthCarrierPoly = azcarr
    rObj.imageIn = IF.createImage().load(slc_path + ".xml")
    rngImg = IF.createImage().load(str(off / "range.off.xml"))
    aziImg = IF.createImage().load(str(off / "azimuth.off.xml"))
    imgOut = IF.createImage()
    imgOut.setWidth(sc.width)
    imgOut.filename = str(tmp_path / "coreg.slc")
    imgOut.dataType = "CFLOAT"
    imgOut.setAccessMode("write")
    rObj.outputWidth, rObj.outputLines = sc.width, sc.length
    rObj.residualRangeImage, rObj.residualAzimuthImage = rngImg, aziImg
    rObj.flatten = True
    rObj.resamp_slc(imageOut=imgOut)
    imgOut.renderHdr()
    out = np.fromfile(tmp_path / "coreg.slc", np.complex64).reshape(sc.length, sc.width)
    o = orc.resamp_slc(slc=slc, out_shape=(sc.length, sc.width), wvl=sc.wvl, slr=sc.dr, flatten=True,
                       az_carrier=[[0.0, 1e-4], [0.3, 0.0], [2e-4, 0.0]], resid_az=az.astype(np.float64),
                       resid_rg=rg.astype(np.float64))
    _compare(out, o)
    assert rObj.numValid == int((o != 0).sum()) > 0.3 * o.size
    hdr = IF.createImage().load(str(tmp_path / "coreg.slc.xml"))
    assert (hdr.width, hdr.length, hdr.dataType) == (sc.width, sc.length, "CFLOAT")
    # errors of the reference's setDefaults are kept
    bad = isce2_b200.createResamp_slc()
    bad.imageIn = IF.createImage().load(slc_path + ".xml")
    bad.inputWidth = sc.width + 1
    with pytest.raises(Exception, match="does not match specified width"):
        bad.resamp_slc(imageOut=imgOut)


def test_fused_geo2rdr_to_resamp_equals_the_path_through_the_offset_rasters():
    sc = pu.rough_scene(200, 1500)
    c = pu.cpu_topo(sc, dem_method="BILINEAR", want_inc=False, want_mask=False)
    sec = synth.config_c1_secondary(length=sc.length, width=sc.width)
    kw = pu.secondary_kwargs(sc, sec, recenter=0.37)
    slc = _slc(sc.length, sc.width, 4)
    rkw = dict(wvl=sc.wvl, slr=sc.dr, flatten=True, az_carrier=[[0.0, 1e-4], [0.3, 0.0], [2e-4, 0.0]], doppler=[[0.02, 1e-6]])
    for f32 in (True, False):
        gp = _capi.geo_params(length=kw["length"], width=kw["width"], dem_shape=c["lat"].shape, r0=kw["r0"], dr=kw["dr"],
                              prf=kw["prf"], t0=kw["t0"], wvl=kw["wvl"], side=kw["side"], out_f32=f32)
        plan = _capi.GeoPlan(gp, lat=c["lat"], lon=c["lon"], hgt=c["hgt"])
        plan.execute(gp, kw["orbit_t"], kw["orbit_pos"], kw["orbit_vel"], want=("azoff", "rgoff"))
        fused = _capi.resamp_slc_from_geo_plan(plan, slc, **rkw)
        offs = plan.fetch()
        plan.close()
        host = _capi.resamp_slc_run(slc, slc.shape, resid_az=offs["azoff"], resid_rg=offs["rgoff"], **rkw)
        assert np.array_equal(fused["slc"], host["slc"]) and fused["num_valid"] == host["num_valid"] > 0.3 * slc.size
        # invalid geo2rdr pixels (-999999) resample to zero
        assert not fused["slc"][offs["rgoff"] == -999999.0].any()
    plan = _capi.GeoPlan(gp, lat=c["lat"], lon=c["lon"], hgt=c["hgt"])
    with pytest.raises(_capi.B200Error, match="executed"):
        _capi.lib()  # keep the library loaded
        plan.last_params = gp
        _capi.resamp_slc_from_geo_plan(plan, slc, **rkw)
    plan.close()
