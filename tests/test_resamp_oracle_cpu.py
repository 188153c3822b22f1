"""Pins of the resamp_slc oracle (oracle/zerodop_oracle.c: orc_resamp_slc) that need no GPU.

The reference's own test for this module (components/stdproc/stdproc/resamp_slc/test/testResamp_slc.py) drives an API
that no longer exists and ships no data, and the Fortran cannot be built here; the restatement is therefore pinned by
the polynomial evaluator it shares with the rest of the path (bit-identical to the reference's poly2d.c,
tests/test_oracle_pins.py) and by the behaviour the algorithm must have: identity and whole-pixel shifts are exact,
band-limited signals shift analytically, declared carriers / Doppler / flattening phases come back as written."""
import numpy as np
import pytest

from oracle import oracle as orc


def _noise(L, W, seed=0):
    rng = np.random.default_rng(seed)
    return (rng.normal(size=(L, W)) + 1j * rng.normal(size=(L, W))).astype(np.complex64)


def test_sinc_table_is_normalised_per_phase_and_differs_from_the_geometry_table():
    import ctypes as C
    t = np.zeros((8192, 8), np.float32)
    orc.lib().orc_resamp_sinc_table(t.ctypes.data_as(C.POINTER(C.c_float)))
    g = np.zeros((8192, 8), np.float32)
    orc.lib().orc_sinc_table(g.ctypes.data_as(C.POINTER(C.c_float)))
    assert np.abs(t.astype(np.float64).sum(axis=1) - 1.0).max() < 4e-7
    assert np.abs(g.astype(np.float64).sum(axis=1) - 1.0).max() > 1e-4  # the topozero / geozero table is not normalised
    # phase 0 is a unit impulse on one tap: resampling with zero offsets is the identity
    assert np.sort(np.abs(t[0]))[-1] == 1.0 and np.sort(np.abs(t[0]))[-2] < 1e-7


def test_identity_and_whole_pixel_shifts_are_exact_and_borders_are_zero():
    L, W = 70, 90
    z = _noise(L, W)
    o = orc.resamp_slc(slc=z, out_shape=(L, W))
    # k = i must satisfy 4 < k < inwidth - 4 (1-based), same for lines (resamp_slc.f90:197-207)
    assert np.array_equal(o[4:L - 5, 4:W - 5], z[4:L - 5, 4:W - 5])
    assert not o[:4].any() and not o[L - 5:].any() and not o[:, :4].any() and not o[:, W - 5:].any()
    # +3 samples in range through the polynomial, -2 lines in azimuth through the residual image
    ra = np.full((L, W), -2.0)
    s = orc.resamp_slc(slc=z, out_shape=(L, W), rg_offsets=[[3.0]], resid_az=ra)
    assert np.array_equal(s[8:L - 8, 8:W - 12], z[6:L - 10, 11:W - 9])
    # a smaller output grid and the invalid-offset marker of geo2rdr (-999999): zeros, no crash
    rr = np.zeros((20, 30))
    rr[5, 7] = -999999.0
    q = orc.resamp_slc(slc=z, out_shape=(20, 30), resid_rg=rr)
    assert q.shape == (20, 30) and q[5, 7] == 0 and q[5, 8] == z[5, 8]


def test_band_limited_signal_shifts_analytically():
    L, W = 80, 120
    y, x = np.mgrid[1:L + 1, 1:W + 1].astype(np.float64)
    a, b = 0.31, -0.23  # rad / sample, well inside the band of the 8-tap windowed sinc
    z = np.exp(1j * (a * x + b * y)).astype(np.complex64)
    dx, dy = 0.37, -0.41
    o = orc.resamp_slc(slc=z, out_shape=(L, W), rg_offsets=[[dx]], az_offsets=[[dy]])
    want = np.exp(1j * (a * (x + dx) + b * (y + dy)))
    inner = np.s_[8:L - 8, 8:W - 8]
    assert np.abs(o[inner] - want[inner]).max() < 5e-3
    # offsets that vary over the image: polynomial in range + residual ramp in azimuth
    rg_poly = [[0.2, 0.004]]           # 0.2 + 0.004 * range pixel
    ra = 0.003 * (y - 1)               # grows with the line
    o2 = orc.resamp_slc(slc=z, out_shape=(L, W), rg_offsets=rg_poly, resid_az=ra)
    want2 = np.exp(1j * (a * (x + 0.2 + 0.004 * x) + b * (y + ra)))
    assert np.abs(o2[inner] - want2[inner]).max() < 5e-3


def test_carriers_and_doppler_are_removed_before_and_restored_after_the_interpolation():
    L, W = 90, 100
    y, x = np.mgrid[1:L + 1, 1:W + 1].astype(np.float64)
    env = (1.0 + 0.3 * np.cos(0.11 * x) * np.sin(0.07 * y))  # slowly varying amplitude
    dx, dy = 0.45, 0.28
    envs = (1.0 + 0.3 * np.cos(0.11 * (x + dx)) * np.sin(0.07 * (y + dy)))
    inner = np.s_[10:L - 10, 10:W - 10]
    # an azimuth carrier far outside the interpolator's band (TOPS-like quadratic phase): c0 + c1*az + c2*az^2 + 0.9*rng
    az_car = [[0.0, 0.9], [2.3, 0.0], [0.021, 0.0]]  # rows = azimuth powers, columns = range powers
    ph = lambda yy, xx: 0.9 * xx + 2.3 * yy + 0.021 * yy * yy
    z = (env * np.exp(1j * ph(y, x))).astype(np.complex64)
    o = orc.resamp_slc(slc=z, out_shape=(L, W), rg_offsets=[[dx]], az_offsets=[[dy]], az_carrier=az_car)
    want = envs * np.exp(1j * ph(y + dy, x + dx))
    assert np.abs(o[inner] - want[inner]).max() < 5e-3
    # without declaring the carrier the same data cannot be interpolated
    bad = orc.resamp_slc(slc=z, out_shape=(L, W), rg_offsets=[[dx]], az_offsets=[[dy]])
    assert np.abs(bad[inner] - want[inner]).max() > 0.3
    # Doppler centroid in radians per line (Resamp_slc.py addSlc: 2 pi fd / prf), here range dependent
    dop = [[1.7, 0.004]]
    zd = (env * np.exp(1j * (1.7 + 0.004 * x) * y)).astype(np.complex64)
    od = orc.resamp_slc(slc=zd, out_shape=(L, W), az_offsets=[[dy]], doppler=dop)
    envd = (1.0 + 0.3 * np.cos(0.11 * x) * np.sin(0.07 * (y + dy)))
    wantd = envd * np.exp(1j * (1.7 + 0.004 * x) * (y + dy))
    assert np.abs(od[inner] - wantd[inner]).max() < 5e-3


def test_flattening_phase_is_the_formula_of_the_reference():
    L, W = 40, 60
    z = _noise(L, W, 3)
    kw = dict(slc=z, out_shape=(L, W), rg_offsets=[[0.3, 0.001]], wvl=0.0555, slr=2.33, r0=800e3, ref_wvl=0.0557, ref_r0=800.7e3,
              ref_slr=2.31)
    a = orc.resamp_slc(flatten=False, **kw)
    b = orc.resamp_slc(flatten=True, **kw)
    i = np.arange(1, W + 1, dtype=np.float64)[None, :]
    r_ro = 0.3 + 0.001 * i
    ph = (4 * np.pi / 0.0555) * ((800e3 - 800.7e3) + (i - 1.0) * (2.33 - 2.31) + r_ro * 2.33) + \
         (4 * np.pi * (800.7e3 + (i - 1.0) * 2.31)) * (1.0 / 0.0557 - 1.0 / 0.0555)
    v = np.abs(a) > 0.1
    ratio = (b[v] / a[v])
    want = np.exp(1j * np.broadcast_to(ph, a.shape)[v])
    assert np.abs(ratio - want).max() < 2e-3  # float32 samples, phases of ~1e8 rad reduced modulo 2 pi in double
