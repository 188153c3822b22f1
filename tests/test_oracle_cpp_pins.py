"""Hold the CPU oracle (oracle/zerodop_oracle.c) against REFERENCE-AUTHORED code for the Fortran-only arithmetic.

The Fortran of the path cannot be compiled here (no gfortran).  The reference ships its own C++ restatement of it --
components/zerodop/GPUtopozero/src/*.cpp, GPUgeo2rdr/src/*.cpp and GPUresampslc/src/*.cpp (CPU branches) --
which compiles unchanged with g++ (oracle/Makefile target ref -> oracle/_ref/libisce2_cpp*_ref.so, doors in
oracle/ref_cpp.py).  Where that C++ agrees with the Fortran as written the oracle must equal it BIT FOR BIT; where the
C++ itself departs from the Fortran the departure is named here with both file:line's, and either emulated through the
oracle's test hook (so that every other line of the routine is still compared bit for bit) or the comparison is left out:

 interpolator   C++ vs Fortran                                                   what is asserted here
 BILINEAR       identical (UniformInterp.cpp:24-50 = uniform_interp.f90:13-44)    bit-exact, 1e6 draws incl. window edges
 NEAREST        identical (TopoMethods.cpp:168-185 = topozeroMethods.f:200-220)   bit-exact
 BIQUINTIC      identical (UniformInterp.cpp:204-265 = spline.f:5-117)            bit-exact
 SINC           (a) table: UniformInterp::sinc_coef (UniformInterp.cpp:150-168) predates the 2021 change of
                uniform_interp.f90:319-363 (soff, wgthgt, floor) -- the current formula is Interpolator::sinc_coef
                (GPUresampslc/src/Interpolator.cpp:119-136): oracle table == that, bit-exact.
                (b) evaluation: UniformInterp.cpp:186 forms the products in double, uniform_interp.f90:424-425 in
                real*4 -- bit-exact with the table injected and the hook's arithmetic; <= 4 float32 ulps without.
 AKIMA          AkimaLib.cpp:70-76 stores the Y slope into slpx (akima_reg.F:95-100: slpy); Constants.h:13 AKI_EPS is an
                int (0); AkimaLib.cpp:107,121-134 subtracts the float32 corner samples in float (akima_reg.F:166-169,
                254-256 in double) -- bit-exact with the hook on a DEM whose sample differences are exact in float32
                (any integer / 1/8 m DEM, i.e. every real SRTM tile); <= 1 float32 ulp on full-mantissa samples.
 BICUBIC        not comparable: UniformInterp::bicubic builds its weight table as vector(16) + 16 push_back's
                (UniformInterp.cpp:87-92), so wt[0..15] are empty vectors and the first wt[i][j] dereferences null;
                both C++ tables also carry the `0.0-9.0` typo in row 4 (UniformInterp.cpp:72, Interpolator.cpp:61).
                The oracle's bicubic stays pinned by polynomial reproduction as written (tests/test_oracle_pins.py).

Whole images: Topo::topo (Topo.cpp) and Geo2rdr::geo2rdr (Geo2rdr.cpp) run here on the scenes of the parity tests
(BILINEAR, BIQUINTIC, NEAREST as they are; AKIMA and SINC with the departures above switched on in the oracle and, for
SINC, Topo.cpp's own table handed to it: orc_test_set_cpp_quirks / orc_test_set_sinc_table).
Known departures of Topo.cpp from topozero.f90, none of which the comparison below depends on except as stated:
MAX_H = -1000 (Constants.h:49; topozeroState.f:74: 9000) and 0-based crop indices clamped to 1 (Topo.cpp:321-333) --
the DEM handed over is therefore smaller than either bounding box, so both crop to the same array; inc channel 1 is the
look angle (Topo.cpp:804; topozero.f90:694-700: psi) -- skipped; the mask's co-sorts are three independent sorts
(Topo.cpp:871-873, 904-906; topozero.f90:735,787 co-sort) so its layover bit is never set -- only the shadow bit is
compared; pow(r,3) (Ellipsoid.cpp:46) against r**3 changes the last bit of ~1e-4 of the XYZ -> LLH results.
"""
import math

import numpy as np
import pytest

from isce2_b200 import synth
from oracle import oracle as orc
from oracle import ref_cpp
from tests import parity_util as pu

pytestmark = pytest.mark.skipif(not ref_cpp.available(), reason="oracle/_ref C++ reference libraries not built "
                                                                "(/root/reference never mounted here)")
A = 6378137.0
E2 = 0.0066943799901


# ------------------------------------------------------------------ DEM interpolators
def _draws(n, nx, ny, seed):
    rng = np.random.default_rng(seed)
    ix = rng.integers(-1, nx + 3, n).astype(np.int32)  # beyond every method's window on both sides
    iy = rng.integers(-1, ny + 3, n).astype(np.int32)
    fx = rng.random(n).astype(np.float32).astype(np.float64)  # the callers' fractions are float32 differences
    fy = rng.random(n).astype(np.float32).astype(np.float64)
    fx[:2000] = 0.0
    fy[1000:3000] = 0.0
    # every edge index of every window, both axes
    k = 3000
    for e in (0, 1, 2, 3, 4, nx - 4, nx - 3, nx - 2, nx - 1, nx, nx + 1):
        ix[k:k + 500] = e
        k += 500
    for e in (0, 1, 2, 3, 4, ny - 4, ny - 3, ny - 2, ny - 1, ny, ny + 1):
        iy[k:k + 500] = e
        k += 500
    return ix, iy, fx, fy


def _dem(nx, ny, seed, quantum=None):
    rng = np.random.default_rng(seed)
    z = rng.standard_normal((ny, nx)).cumsum(0).cumsum(1) * 3 + 500
    if quantum:
        z = np.round(z / quantum) * quantum
    return z.astype(np.float32)


@pytest.mark.parametrize("method", ["BILINEAR", "NEAREST", "BIQUINTIC"])
def test_interpolators_bit_exact(method):
    nx, ny, n = 161, 147, 1_000_000
    dem = _dem(nx, ny, 1)
    ix, iy, fx, fy = _draws(n, nx, ny, 2)
    ref = ref_cpp.interp_dem(method, dem, ix, iy, fx, fy)
    got = orc.interp_dem_batch(method, dem, ix, iy, fx, fy)
    assert np.array_equal(ref, got), int((ref != got).sum())
    nbad = int((ref == -1000.0).sum())
    assert 0 < nbad < n // 4  # the out-of-window branch (BADVALUE) is part of the comparison


def test_sinc_table_is_the_current_reference_formula():
    f = ref_cpp.sinc_coef(1.0, 8.0, 8192, 0.0, 1)  # r_filter of prepareMethods (topozeroMethods.f:55)
    tab = np.empty(8192 * 8, np.float32)
    for i in range(8):  # fintp(i + j*sinc_len) = r_filter(j + i*sinc_sub), topozeroMethods.f:57-61
        tab[i::8] = f[i * 8192:(i + 1) * 8192].astype(np.float32)
    assert np.array_equal(tab, orc.sinc_table())
    # and the table TopoMethods::prepareMethods builds is the older variant -- hence the injection below
    old = ref_cpp.topo_sinc_table()
    assert not np.array_equal(old, tab) and np.abs(old - tab).max() < 0.2


def test_sinc_evaluation():
    nx, ny, n = 161, 147, 1_000_000
    dem = _dem(nx, ny, 3)
    ix, iy, fx, fy = _draws(n, nx, ny, 4)
    tab = orc.sinc_table()
    ref = ref_cpp.interp_dem("SINC", dem, ix, iy, fx, fy, sinc_table=tab)
    assert np.array_equal(ref, orc.interp_dem_batch("SINC", dem, ix, iy, fx, fy, cpp_quirks=2))
    got = orc.interp_dem_batch("SINC", dem, ix, iy, fx, fy)  # as the Fortran rounds: real*4 products and sums
    assert np.array_equal(ref == -1000.0, got == -1000.0)
    ulp = float(np.spacing(np.float32(np.abs(dem).max())))  # sums of terms of the samples' magnitude
    assert np.abs(ref.astype(np.float64) - got).max() <= 6 * ulp


def test_akima():
    nx, ny, n = 161, 147, 1_000_000
    ix, iy, fx, fy = _draws(n, nx, ny, 6)
    dem = _dem(nx, ny, 5, quantum=0.125)
    ref = ref_cpp.interp_dem("AKIMA", dem, ix, iy, fx, fy)
    assert np.array_equal(ref, orc.interp_dem_batch("AKIMA", dem, ix, iy, fx, fy, cpp_quirks=1))
    # flat patches exercise the equal-slope branches (aki_almostEqual) and the 0/0 guards of the weights
    dem2 = dem.copy()
    dem2[40:60, 50:90] = 321.0
    dem2[80:120:2, :] = dem2[81:121:2, :]
    ref = ref_cpp.interp_dem("AKIMA", dem2, ix, iy, fx, fy)
    assert np.array_equal(ref, orc.interp_dem_batch("AKIMA", dem2, ix, iy, fx, fy, cpp_quirks=1))
    # full-mantissa samples: the C++ float subtraction of the corner samples costs at most one float32 ulp
    dem3 = _dem(nx, ny, 7)
    ref = ref_cpp.interp_dem("AKIMA", dem3, ix, iy, fx, fy)
    got = orc.interp_dem_batch("AKIMA", dem3, ix, iy, fx, fy, cpp_quirks=1)
    ulp = float(np.spacing(np.float32(np.abs(dem3).max())))
    assert np.abs(ref.astype(np.float64) - got).max() <= 2 * ulp
    # without the hook the oracle follows akima_reg.F (slpy stored): it must differ from the C++ on rough terrain
    assert not np.array_equal(ref, orc.interp_dem_batch("AKIMA", dem3, ix, iy, fx, fy))


# ------------------------------------------------------------------ geometry primitives
def test_geometry_primitives_bit_exact():
    L = orc.lib()
    rng = np.random.default_rng(3)
    n = 20000
    n_llh_diff = 0
    for _ in range(n):
        lat, lon, hdg = rng.uniform(-1.5, 1.5), rng.uniform(-3.1, 3.1), rng.uniform(-3.1, 3.1)
        h = rng.uniform(-500.0, 800e3)
        xyz = orc.latlon_to_xyz([lat, lon, h], A, E2)
        assert np.array_equal(xyz, ref_cpp.latlon(A, E2, [lat, lon, h], True))  # Ellipsoid.cpp:35-41 = latlon.F:48-54
        l1, l2 = orc.xyz_to_latlon(xyz, A, E2), ref_cpp.latlon(A, E2, xyz, False)  # Ellipsoid.cpp:42-57 = latlon.F:56-71
        if not np.array_equal(l1, l2):  # pow(r,3) vs r**3: last bit only
            n_llh_diff += 1
            assert abs(l1[0] - l2[0]) < 1e-15 and abs(l1[1] - l2[1]) < 1e-15 and abs(l1[2] - l2[2]) < 1e-8
        assert L.orc_reast(A, E2, lat) == ref_cpp.reast(A, E2, lat)  # curvature.F:34
        assert L.orc_rnorth(A, E2, lat) == ref_cpp.rnorth(A, E2, lat)  # curvature.F:45
        assert L.orc_rdir(A, E2, hdg, lat) == ref_cpp.rdir(A, E2, hdg, lat)  # curvature.F:62
        vel = rng.standard_normal(3) * 7000.0
        t, c, nn = np.zeros(3), np.zeros(3), np.zeros(3)
        L.orc_tcnbasis(orc._d(xyz), orc._d(vel), A, E2, orc._d(t), orc._d(c), orc._d(nn))  # tcnbasis.F:26-39
        rt, rc, rn = ref_cpp.tcnbasis(xyz, vel, A, E2)
        assert np.array_equal(t, rt) and np.array_equal(c, rc) and np.array_equal(nn, rn)
        m = np.zeros(9)
        L.orc_enubasis(lat, lon, orc._d(m))  # enubasis.F:39-60
        assert np.array_equal(m, ref_cpp.enubasis(lat, lon))
        ov = np.zeros(3)
        r = L.orc_radar_to_xyz(A, E2, lat, lon, hdg, orc._d(m), orc._d(ov))  # radar_to_xyz.F:49-92
        rm, rov, rr = ref_cpp.radar_to_xyz(A, E2, lat, lon, hdg)
        assert np.array_equal(m, rm) and np.array_equal(ov, rov) and r == rr
        tgt = orc.latlon_to_xyz([lat + rng.uniform(-.05, .05), lon + rng.uniform(-.05, .05), rng.uniform(-500, 9000)], A, E2)
        sch = np.zeros(3)
        L.orc_xyz_to_sch(orc._d(m), orc._d(ov), r, orc._d(tgt), orc._d(sch))  # convert_sch_to_xyz.F:63-72
        assert np.array_equal(sch, ref_cpp.xyz_to_sch(A, E2, lat, lon, hdg, tgt))
    assert n_llh_diff < n // 1000


def test_orbit_and_poly2d_bit_exact():
    sc = synth.make_scene(64, 256, sensor="nisar", dem=False)
    o = orc.Orbit(sc.orbit_t, sc.orbit_pos, sc.orbit_vel)
    rng = np.random.default_rng(0)
    # (the C++ SCH interpolator assigns instead of accumulating -- Orbit.cpp:144-145 `=` where orbit.c:166-167 has `+=` --
    # and is left out; the oracle's is bit-identical to orbit.c itself, tests/test_oracle_pins.py)
    for method in ("HERMITE", "LEGENDRE"):
        for tq in rng.uniform(sc.orbit_t[0] - 5.0, sc.orbit_t[-1] + 5.0, 1500):  # beyond the span too: stat == 1
            s1, p1, v1 = o.interp(tq, method)
            s2, p2, v2 = ref_cpp.interp_orbit(o.t, o.pos, o.vel, method, tq)
            assert s1 == s2 and np.array_equal(p1, p2) and np.array_equal(v1, v2), (method, tq)
    p = orc.Poly2D([[-120.0, 2e-2, -1.5e-6, 3e-11], [0.3, -1e-4, 2e-9, 0.0]], 10.0, 5.0, 1000.0, 300.0)
    for _ in range(2000):
        az, rg = rng.uniform(0, 60000), rng.uniform(0, 25000)
        assert p(az, rg) == ref_cpp.eval_poly2d(p.coeffs, az, rg, 10.0, 5.0, 1000.0, 300.0)


def test_sort_building_block_of_the_mask():
    """The key ordering of the mask's co-sort.  (LinAlg::binarySearch, LinAlg.cpp:113-129, is a nearest-neighbour search
    with other return conventions than the Fortran's binarysearch, topozero.f90:933-963: not comparable.)"""
    rng = np.random.default_rng(5)
    for _ in range(300):
        n = int(rng.integers(2, 400))
        a = np.sort(rng.standard_normal(n)) + rng.standard_normal(n) * rng.choice([0.0, 0.05])  # nearly sorted, like ctrack
        a[rng.integers(0, n, 3)] = a[rng.integers(0, n, 3)]  # ties
        s, b, c = a.copy(), np.arange(n, dtype=np.float64), -np.arange(n, dtype=np.float64)
        orc.lib().orc_insertion_sort(orc._d(s), orc._d(b), orc._d(c), n)
        assert np.array_equal(s, ref_cpp.insertion_sort(a))  # LinAlg.cpp:98-111 == topozero.f90:910-930 on the key
        assert np.array_equal(a[b.astype(int)], s) and np.array_equal(b, -c)  # the companions travel with the key
        assert np.all(np.diff(b)[np.diff(s) == 0] > 0)  # stable: equal keys keep their order (moves only on strict >)


# ------------------------------------------------------------------ whole images
def _clamped_dem(sc, pad_deg=0.02):
    """Window of sc.dem that (a) covers every pixel of the scene for terrain between -500 and 3000 m plus pad_deg and
    (b) lies inside the scene's bounding box +- 0.15 deg for BOTH height pairs ({-500, 9000}: topozeroState.f:74,
    {-500, -1000}: Constants.h:48-49), so that the Fortran-faithful crop and Topo.cpp's crop are both clamped to it."""
    dur = (sc.length - 1) / sc.prf
    lats, lons = [], []
    for tq in (sc.t0, sc.t0 + dur):
        p, v = synth.hermite_point(sc.orbit_t, sc.orbit_pos, sc.orbit_vel, tq)
        for rg in (sc.r0, sc.r0 + (sc.width - 1) * sc.dr):
            for h in (-500.0, 3000.0):
                la, lo, _ = synth.xyz_to_llh(synth._ground_point(p, v, rg, h, sc.side))
                lats.append(float(la))
                lons.append(float(lo))
    i0 = int(math.floor((max(lats) + pad_deg - sc.first_lat) / sc.delta_lat))
    i1 = int(math.ceil((min(lats) - pad_deg - sc.first_lat) / sc.delta_lat))
    j0 = int(math.floor((min(lons) - pad_deg - sc.first_lon) / sc.delta_lon))
    j1 = int(math.ceil((max(lons) + pad_deg - sc.first_lon) / sc.delta_lon))
    assert 0 < i0 < i1 < sc.dem.shape[0] and 0 < j0 < j1 < sc.dem.shape[1]
    return np.ascontiguousarray(sc.dem[i0:i1 + 1, j0:j1 + 1]), sc.first_lat + i0 * sc.delta_lat, sc.first_lon + j0 * sc.delta_lon


@pytest.mark.parametrize("sensor,dem_method,orbit_method", [("s1", "BILINEAR", "HERMITE"), ("s1", "BIQUINTIC", "HERMITE"),
                                                            ("s1", "NEAREST", "HERMITE"), ("nisar", "BIQUINTIC", "LEGENDRE"),
                                                            ("nisar", "BILINEAR", "HERMITE")])
def test_whole_image_topo_against_reference_cpp(sensor, dem_method, orbit_method):
    """T2 + T4 + T5 + T6 + the shadow half of T8 of SURVEY 8a in one comparison: lat / lon / hgt, both LOS channels, the
    local incidence angle and the shadow bit of every pixel, native Doppler and left looking included."""
    length, width = 48, 2048
    sc = pu.rough_scene(length, width, sensor=sensor)
    dem, flat, flon = _clamped_dem(sc)
    kw = orc.scene_topo_kwargs(sc, dem_method=dem_method, orbit_method=orbit_method)
    ref = ref_cpp.topo(**{**kw, "dem": dem, "first_lat": flat, "first_lon": flon})
    ny, nx = dem.shape
    # Topo.cpp:321-333 with every side clamped: rows / columns 1 .. n-1 of what it was given, origin one post in
    assert ref["crop"] == dict(width=nx - 1, length=ny - 1, line0=1, line1=ny - 1, pixel0=1, pixel1=nx - 1), ref["crop"]
    got = orc.topo(**{**kw, "dem": np.ascontiguousarray(dem[1:, 1:]), "first_lat": flat + sc.delta_lat,
                      "first_lon": flon + sc.delta_lon})
    assert (got["ustartx"], got["ustarty"], got["udemwidth"], got["udemlength"]) == (1, 1, nx - 1, ny - 1)
    n = length * width
    # rough terrain: some pixels never converge, so the secondary-iteration averaging (topozero.f90:572-593) is compared too
    assert got["totalconv"] == ref["totalconv"] and 0.5 * n < got["totalconv"] < n  # (nearest neighbour: steps in the terrain)
    for k, tol in (("lat", 1e-13), ("lon", 1e-13), ("hgt", 1e-8)):
        d = np.abs(got[k] - ref[k])
        assert d.max() <= tol, (k, d.max())  # the pow(r,3) last bit, nothing else
        assert np.mean(d == 0) > 0.999, (k, np.mean(d == 0))
    assert np.array_equal(got["los"], ref["los"])
    assert np.array_equal(got["inc"][:, 1, :], ref["inc"][:, 1, :])  # local incidence angle (topozero.f90:662-692)
    shadow_ref, shadow_got = ref["mask"] & 1, got["mask"] & 1
    assert np.array_equal(shadow_ref, shadow_got) and shadow_got.sum() > 100  # topozero.f90:789-809
    assert (got["mask"] & 2).sum() > 100 and (ref["mask"] & 2).sum() == 0  # Topo.cpp never sets the layover bit (see header)
    # inc channel 1: the C++ writes the look angle, the Fortran psi -- they must differ
    assert np.abs(got["inc"][:, 0, :] - ref["inc"][:, 0, :]).max() > 1.0


@pytest.mark.parametrize("case", ["zero_doppler", "doppler_poly", "bistatic", "legendre_native_doppler", "outside"])
def test_whole_image_geo2rdr_against_reference_cpp(case):
    """G1 + G2 + G3 of SURVEY 8a: every output of every pixel, validity included, bit for bit (config 1 geometry:
    perturbed secondary orbit, misregistered start time / range)."""
    length, width = 64, 2048
    sensor = "nisar" if case == "legendre_native_doppler" else "s1"
    sc = pu.rough_scene(length, width, sensor=sensor)
    c = pu.cpu_topo(sc, dem_method="BILINEAR", want_mask=False, want_inc=False)
    if sensor == "s1":
        sec = synth.config_c1_secondary(length=length, width=width)
        kw = pu.secondary_kwargs(sc, sec, recenter=0.37 + (0.12 if case == "outside" else 0.0))
    else:
        kw = dict(orbit_t=sc.orbit_t, orbit_pos=sc.orbit_pos, orbit_vel=sc.orbit_vel, length=length, width=width,
                  r0=sc.r0 - 1.7, dr=sc.dr, prf=sc.prf, t0=sc.t0 - 0.013, wvl=sc.wvl, side=sc.side)
    opts = dict(zero_doppler={}, outside={}, bistatic=dict(bistatic=True),
                doppler_poly=dict(doppler_coeffs=(0.02, 1e-6, -2e-11), doppler_mean=100.0, doppler_norm=2.0),
                legendre_native_doppler=dict(orbit_method="LEGENDRE",
                                             doppler_coeffs=tuple(x / sc.prf for x in sc.doppler_coeffs[0])))[case]
    got = orc.geo2rdr(lat=c["lat"], lon=c["lon"], hgt=c["hgt"], **kw, **opts)
    ref = ref_cpp.geo2rdr(lat=c["lat"], lon=c["lon"], hgt=c["hgt"], **kw, **opts)
    for k in ("azt", "rgm", "azoff", "rgoff"):
        assert np.array_equal(got[k], ref[k]), (k, int((got[k] != ref[k]).sum()))
    assert got["num_valid"] == ref["num_valid"] and got["num_outside"] == ref["num_outside"]
    assert got["num_conv"] == ref["num_conv"]
    n = length * width
    if case == "outside":
        assert got["num_valid"] == 0
    else:
        assert 0.3 * n < got["num_valid"] < n  # the scene is cut by the secondary's window: both branches are exercised


# ------------------------------------------------------------------ resamp_slc against ResampSlc::_resamp_cpu
# Departures of GPUresampslc/src/ResampSlc.cpp from resamp_slc.f90, each handled as stated:
#  (1) the sinc table is used as sinc_coef delivers it (ResampMethods.cpp:31-39) and sinc_eval_2d neither normalises the
#      weights nor divides by their sum (Interpolator.cpp:163-176); resamp_slcMethods.f:66-73 normalises every sub-sample
#      phase of the table and uniform_interp.f90:456-484 divides by the weight sum again.  The C++ result is therefore the
#      Fortran's times (sum of the 8 azimuth taps) x (sum of the 8 range taps) of ITS table, 1 .. 1.003: applied here from
#      the reference's own table, per pixel.
#  (2) Doppler at the output pixel (ResampSlc.cpp:293) against resamp_slc.f90:211 (secondary's coordinate, 12-AUG-2020),
#      carriers added back at 0-based coordinates (ResampSlc.cpp:311) against the 1-based :233-235: the oracle's test hook
#      evaluates them where the C++ does; everything else in the routine is the path under test.
#  (3) the C++ admits one more line / sample at the far edge (ResampSlc.cpp:288,291: k >= n-4 on 0-based k; resamp_slc.f90:
#      198,204: k >= n-4 on 1-based k): pixels the C++ fills and the Fortran skips must be exactly those.
#  (4) accumulation: complex<float> products summed in float, azimuth outer (Interpolator.cpp:169-172) against real*8
#      weights and range outer: float32 rounding, bounded below.
#  (5) pixels outside the bounds keep whatever the previous lines left in the output line buffer (the C++ never clears
#      it); the Fortran writes zeros.  Left out of the comparison.
RESAMP_CASES = {
    "identity":              dict(shape=(96, 128), out=(96, 128)),
    "offsets+residuals":     dict(shape=(160, 220), out=(150, 200), resid=True, offsets=True),
    "carriers+doppler":      dict(shape=(160, 220), out=(150, 200), resid=True, offsets=True, carriers=True, doppler=True),
    "flatten":               dict(shape=(160, 220), out=(150, 200), resid=True, offsets=True, carriers=True, doppler=True, flatten=True),
    "two tiles (1000 + 80)": dict(shape=(1100, 72), out=(1080, 64), resid=True, offsets=True, carriers=True, doppler=True, flatten=True),
}


def _resamp_case(c, seed=5):
    rng = np.random.default_rng(seed)
    L, W = c["shape"]
    oL, oW = c["out"]
    P = orc.Poly2D
    kw = dict(slc=(rng.standard_normal((L, W)) + 1j * rng.standard_normal((L, W))).astype(np.complex64), out_shape=(oL, oW),
              wvl=0.0555, slr=2.33, r0=800000.0, ref_wvl=0.0556, ref_r0=800010.0, ref_slr=2.33, flatten=bool(c.get("flatten")))
    if c.get("offsets"):
        kw["rg_offsets"] = P([[1.5, 0.002], [0.001, 0.0]])
        kw["az_offsets"] = P([[-0.5, 0.0005], [0.003, 0.0]])
    if c.get("resid"):
        kw["resid_az"] = 0.3 * rng.standard_normal((oL, oW)) + 2.3
        kw["resid_rg"] = 0.3 * rng.standard_normal((oL, oW)) - 1.7
    if c.get("carriers"):
        kw["rg_carrier"] = P([[0.1, 0.002], [0.001, 0.0]])
        kw["az_carrier"] = P([[0.2, 0.0], [0.03, 1e-5]])
    if c.get("doppler"):
        kw["doppler"] = P([[0.05, 0.0004], [0.0003, 0.0]])
    return kw


def _np_poly(p, az, rg):
    out = np.zeros(az.shape)
    if p is not None:
        for m in range(p.coeffs.shape[0]):
            for n in range(p.coeffs.shape[1]):
                out += p.coeffs[m, n] * az ** m * rg ** n
    return out


@pytest.mark.parametrize("case", list(RESAMP_CASES))
def test_whole_image_resamp_slc_against_reference_cpp(case):
    kw = _resamp_case(RESAMP_CASES[case])
    ours = orc.resamp_slc(**kw, cpp_positions=True)
    ref, text = ref_cpp.resamp_slc(**kw)
    assert "Interpolating" in text
    L, W = kw["slc"].shape
    oL, oW = kw["out_shape"]
    # (1): the reference's own unnormalised table -> per-pixel product of its tap sums
    tab = ref_cpp.sinc_coef(1.0, 8.0, 8192, 0.0, 1).reshape(8, 8192).T.astype(np.float32)
    S = tab.astype(np.float64).sum(axis=1)
    ii, jj = np.meshgrid(np.arange(oL, dtype=np.float64), np.arange(oW, dtype=np.float64), indexing="ij")
    ao = _np_poly(kw.get("az_offsets"), ii + 1, jj + 1) + (kw.get("resid_az") if kw.get("resid_az") is not None else 0.0)
    ro = _np_poly(kw.get("rg_offsets"), ii + 1, jj + 1) + (kw.get("resid_rg") if kw.get("resid_rg") is not None else 0.0)
    fa, ka = np.modf(ii + ao)
    fr, kr = np.modf(jj + ro)
    scale = S[np.clip((fa * 8192).astype(int), 0, 8191)] * S[np.clip((fr * 8192).astype(int), 0, 8191)]
    # (3): validity from the two bounds rules; the reference fills its extra far-edge line / sample
    inside = (ka >= 4) & (kr >= 4)
    v_f = inside & (ka < L - 5) & (kr < W - 5)
    v_cpp = inside & (ka < L - 4) & (kr < W - 4)
    assert ((ours != 0) == v_f).all()
    assert (ref[v_cpp] != 0).all()
    # (5): outside its bounds the C++ `continue`s without clearing imgOut (ResampSlc.cpp:171,288-291), so such pixels repeat
    # the last value written to their column; resamp_slc.f90:176 zeroes the line first.  Not compared.
    both = v_f
    assert both.sum() > 0.8 * (oL - 12) * (oW - 12)
    # (4): float32 accumulation in a different order -- a few ulps of the largest of the 64 terms
    err = np.abs(ours.astype(np.complex128) * scale - ref)[both]
    amp = np.abs(ref)[both]
    assert (err <= 4e-6 * np.maximum(amp, 1.0)).all(), (err / np.maximum(amp, 1.0)).max()
    assert np.median(err / amp) < 5e-7
    # and the hook matters exactly when the polynomials vary: without it the same comparison fails by orders of magnitude
    if RESAMP_CASES[case].get("doppler"):
        plain = orc.resamp_slc(**kw)
        assert np.median(np.abs(plain.astype(np.complex128) * scale - ref)[both] / amp) > 1e-3


def test_whole_image_topo_akima_against_reference_cpp():
    """The AKIMA interpolator through the whole path: with the three departures of AkimaLib.cpp switched on in the oracle
    (slope stored into slpx, integer AKI_EPS, float32 corner differences: the header of this file) and a DEM of whole metres
    (every SRTM tile; the float32 differences are then exact), Topo::topo and the oracle agree as for the other methods."""
    length, width = 32, 1536
    sc = pu.rough_scene(length, width)
    dem, flat, flon = _clamped_dem(sc)
    dem = np.round(dem).astype(np.float32)
    kw = orc.scene_topo_kwargs(sc, dem_method="AKIMA", orbit_method="HERMITE")
    ref = ref_cpp.topo(**{**kw, "dem": dem, "first_lat": flat, "first_lon": flon})
    orc.lib().orc_test_set_cpp_quirks(1)
    try:
        got = orc.topo(**{**kw, "dem": np.ascontiguousarray(dem[1:, 1:]), "first_lat": flat + sc.delta_lat,
                          "first_lon": flon + sc.delta_lon})
    finally:
        orc.lib().orc_test_set_cpp_quirks(0)
    plain = orc.topo(**{**kw, "dem": np.ascontiguousarray(dem[1:, 1:]), "first_lat": flat + sc.delta_lat,
                        "first_lon": flon + sc.delta_lon})
    n = length * width
    assert got["totalconv"] == ref["totalconv"] and 0.5 * n < got["totalconv"] <= n
    for k, tol in (("lat", 1e-13), ("lon", 1e-13), ("hgt", 1e-8)):
        d = np.abs(got[k] - ref[k])
        assert d.max() <= tol, (k, d.max())
        assert np.mean(d == 0) > 0.999, (k, np.mean(d == 0))
    assert np.array_equal(got["los"], ref["los"])
    assert np.array_equal(got["inc"][:, 1, :], ref["inc"][:, 1, :])
    assert np.array_equal(ref["mask"] & 1, got["mask"] & 1)
    # ... and the departures are real: as the Fortran is written (akima_reg.F:95-100) the heights differ
    assert np.abs(plain["hgt"] - ref["hgt"]).max() > 1e-3


def test_whole_image_topo_sinc_against_reference_cpp():
    """The SINC interpolator through the whole path.  Topo.cpp builds its table with the older UniformInterp::sinc_coef and
    forms the products in double (header of this file): with that table handed to the oracle and the C++'s arithmetic
    switched on, every other line of the sinc path -- window placement, the one-cell shift, index clamping, edge
    fall-backs -- is compared through Topo::topo."""
    length, width = 32, 1536
    sc = pu.rough_scene(length, width)
    dem, flat, flon = _clamped_dem(sc)
    kw = orc.scene_topo_kwargs(sc, dem_method="SINC", orbit_method="HERMITE")
    ref = ref_cpp.topo(**{**kw, "dem": dem, "first_lat": flat, "first_lon": flon})
    table = np.ascontiguousarray(ref_cpp.topo_sinc_table(), np.float32)
    assert not np.array_equal(table, orc.sinc_table())  # the two formulas do differ
    args = {**kw, "dem": np.ascontiguousarray(dem[1:, 1:]), "first_lat": flat + sc.delta_lat, "first_lon": flon + sc.delta_lon}
    orc.lib().orc_test_set_sinc_table(table.ctypes.data_as(orc._fp))
    orc.lib().orc_test_set_cpp_quirks(2)
    try:
        got = orc.topo(**args)
    finally:
        orc.lib().orc_test_set_cpp_quirks(0)
        orc.lib().orc_test_set_sinc_table(None)
    assert np.array_equal(np.ascontiguousarray(orc.sinc_table()), orc.sinc_table()) and not np.array_equal(table, orc.sinc_table())
    # on this terrain most pixels do not reach the 5 cm threshold with SINC and go through all primary and secondary
    # iterations -- in both implementations alike
    assert got["totalconv"] == ref["totalconv"] and 0 < got["totalconv"] <= length * width
    for k, tol in (("lat", 1e-13), ("lon", 1e-13), ("hgt", 1e-8)):
        d = np.abs(got[k] - ref[k])
        assert d.max() <= tol, (k, d.max())
        assert np.mean(d == 0) > 0.999, (k, np.mean(d == 0))
    assert np.array_equal(got["los"], ref["los"])
    assert np.array_equal(got["inc"][:, 1, :], ref["inc"][:, 1, :])
    assert np.array_equal(ref["mask"] & 1, got["mask"] & 1)
