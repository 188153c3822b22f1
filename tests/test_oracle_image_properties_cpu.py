"""Whole-image pin of the topozero / geo2rdr restatement (the reference asserts no image output anywhere, so
"self-consistency becomes the pin").  Every pixel of an oracle run is checked against the DEFINING equations of the
range-Doppler geometry, evaluated here in plain numpy, independently of the restatement's internals:

  (1) |P - S(t_line)| = slant range of the pixel             (range sphere, topozero.f90:495-516)
  (2) (P - S) . V = lambda/2 * f_dop * range                  (Doppler cone; zero-Doppler and native-Doppler scenes)
  (3) the point lies on the side of the track the look direction says
  (4) its ellipsoidal height is the DEM height at its own (lat, lon) to the iteration threshold
  (5) geo2rdr of the layers with the same orbit returns the pixel's own line / sample (offsets ~ 0, 1e-3 px)
  (6) los channel 1 is the angle between the line of sight and the local vertical

with S, V from the reference-identical Hermite / Legendre interpolators (bit-identical to the reference C, see
test_oracle_pins.py) and P = WGS-84 LLH -> ECEF in numpy."""
import numpy as np
import pytest

from isce2_b200 import synth
from oracle import oracle as orc
from tests import parity_util as pu


def _ecef(lat_deg, lon_deg, h, a, e2):
    la, lo = np.radians(lat_deg), np.radians(lon_deg)
    n = a / np.sqrt(1.0 - e2 * np.sin(la) ** 2)
    return np.stack([(n + h) * np.cos(la) * np.cos(lo), (n + h) * np.cos(la) * np.sin(lo), (n * (1.0 - e2) + h) * np.sin(la)], -1)


def _bilinear(dem, y, x):
    iy, ix = np.floor(y).astype(int), np.floor(x).astype(int)
    fy, fx = y - iy, x - ix
    d = dem.astype(np.float64)
    return ((1 - fy) * (1 - fx) * d[iy, ix] + (1 - fy) * fx * d[iy, ix + 1] + fy * (1 - fx) * d[iy + 1, ix] + fy * fx * d[iy + 1, ix + 1])


@pytest.mark.parametrize("sensor,orbit_method,dem_method,dop2d", [
    ("s1", "HERMITE", "BILINEAR", None), ("s1", "HERMITE", "BIQUINTIC", None), ("nisar", "LEGENDRE", "BICUBIC", None),
    ("s1", "SCH", "BILINEAR", None),
    # azimuth-varying (2-D) Doppler, Hz against (line, range pixel): Topozero.py:305-334
    ("s1", "HERMITE", "BILINEAR", [[40.0, 1.0e-2, -2.0e-7], [0.8, 1.0e-5, 0.0]])])
def test_every_pixel_satisfies_the_range_doppler_equations(sensor, orbit_method, dem_method, dop2d):
    sc = synth.make_scene(20, 1536, sensor=sensor)  # ordinary terrain: (nearly) every pixel converges
    if dop2d is not None:
        sc.doppler_coeffs = dop2d
    o = orc.topo(**orc.scene_topo_kwargs(sc, dem_method=dem_method, orbit_method=orbit_method, want_mask=False))
    # (1)-(3), (5), (6) hold for every pixel whether or not its iteration met the threshold: the final pass always puts
    # the point on the pixel's range sphere and Doppler cone (topozero.f90:618-644); (4) is a statement about converged ones
    assert o["totalconv"] >= 0.999 * sc.pixels
    orb = orc.Orbit(sc.orbit_t, sc.orbit_pos, sc.orbit_vel)
    P = _ecef(o["lat"], o["lon"], o["hgt"], sc.a, sc.e2)
    rng = sc.r0 + sc.dr * np.arange(sc.width)
    dop = np.array([[orc.Poly2D(sc.doppler_coeffs)(line, pix) for pix in (0, sc.width // 2, sc.width - 1)] for line in (0, sc.length - 1)])
    cols = [0, sc.width // 2, sc.width - 1]
    for line in range(sc.length):
        stat, S, V = orb.interp(sc.t0 + line / sc.prf, orbit_method)
        assert stat == 0
        d = P[line] - S
        r = np.linalg.norm(d, axis=1)
        assert np.abs(r - rng).max() < 2e-6, (line, np.abs(r - rng).max())                       # (1), metres
        # (2): checked on the columns where the Doppler polynomial was evaluated through the reference-identical evaluator
        if line in (0, sc.length - 1):
            fd = dop[0 if line == 0 else 1]
            lhs = (d[cols] @ V)
            rhs = 0.5 * sc.wvl * fd * rng[cols]
            assert np.abs(lhs - rhs).max() < 2e-5 * np.linalg.norm(V), (line, lhs - rhs)          # metres x speed
        # (3): sign of the cross-track component, c = unit(n x v) with n the inward vertical at the satellite
        nhat = -S / np.linalg.norm(S)
        chat = np.cross(nhat, V)
        assert np.all(np.sign(d @ chat) == -sc.side)
        # (6)
        up = _ecef(o["lat"][line], o["lon"][line], o["hgt"][line] + 1.0, sc.a, sc.e2) - P[line]
        cosi = np.abs(np.sum(-d * up, axis=1)) / r
        assert np.abs(np.degrees(np.arccos(np.clip(cosi, -1, 1))) - o["los"][line, 0]).max() < 2e-5
        # los channel 2: direction of the target -> platform vector in the local horizontal plane, from North, anticlockwise
        la, lo = np.radians(o["lat"][line]), np.radians(o["lon"][line])
        east = np.stack([-np.sin(lo), np.cos(lo), 0 * lo], -1)
        north = np.stack([-np.sin(la) * np.cos(lo), -np.sin(la) * np.sin(lo), np.cos(la)], -1)
        az = np.degrees(np.arctan2(-np.sum(d * north, 1), -np.sum(d * east, 1)) - 0.5 * np.pi)
        dz = np.abs(az - o["los"][line, 1])
        assert np.minimum(dz, np.abs(dz - 360.0)).max() < 2e-5
    # (4): plain bilinear DEM at the pixel's own position against its height; the iteration stops at 0.05 m in slant range
    y = (o["lat"] - sc.first_lat) / sc.delta_lat
    x = (o["lon"] - sc.first_lon) / sc.delta_lon
    dh = np.abs(_bilinear(sc.dem, y, x) - o["hgt"])
    tol = 0.2 if dem_method == "BILINEAR" else 1.5  # the other interpolators differ from bilinear by their own overshoot
    assert np.quantile(dh, 0.999) < tol, np.quantile(dh, 0.999)
    # (5) -- geo2rdr knows the Doppler as a function of range only (Geo2rdr.py:264-270)
    if dop2d is not None:
        return
    g = orc.geo2rdr(lat=o["lat"], lon=o["lon"], hgt=o["hgt"], orbit_t=sc.orbit_t, orbit_pos=sc.orbit_pos, orbit_vel=sc.orbit_vel,
                    length=sc.length, width=sc.width, r0=sc.r0, dr=sc.dr, prf=sc.prf, t0=sc.t0, wvl=sc.wvl, side=sc.side,
                    orbit_method=orbit_method,
                    # geo2rdr takes the centroid in cycles / PRF against the range pixel (StripmapProc/runGeo2rdr.py:77-80)
                    doppler_coeffs=tuple(c / sc.prf for c in sc.doppler_coeffs[0]))
    inner = np.s_[1:-1, 1:-1]  # the first / last line and sample sit on the acquisition bounds geo2rdr tests against
    assert (g["azoff"][inner] != -999999.0).all()
    assert np.abs(g["azoff"][inner]).max() < pu.TOL_OFFSET_PX and np.abs(g["rgoff"][inner]).max() < pu.TOL_OFFSET_PX


def test_shadow_bit_against_its_geometric_definition_and_mask_on_simple_terrain():
    """Shadow (mask bit 1, topozero.f90:791-809): the look angle at the satellite -- the angle between the direction to the
    centre of the local approximating sphere and the direction to the target -- must grow with range; where it does not
    (forward scan against the running maximum, backward scan against the running minimum) the pixel is shadowed.  The
    angle is rebuilt here from the output layers alone (law of cosines replaced by the actual vectors)."""
    sc = pu.rough_scene(16, 4096)
    o = orc.topo(**orc.scene_topo_kwargs(sc, dem_method="BILINEAR", want_mask=True))
    m = o["mask"]
    assert set(np.unique(m).tolist()) <= {0, 1, 2, 3} and (m == 1).any() and (m >= 2).any()
    orb = orc.Orbit(sc.orbit_t, sc.orbit_pos, sc.orbit_vel)
    P = _ecef(o["lat"], o["lon"], o["hgt"], sc.a, sc.e2)
    mism = 0
    for line in range(sc.length):
        _, S, V = orb.interp(sc.t0 + line / sc.prf, "HERMITE")
        llh = orc.xyz_to_latlon(S, sc.a, sc.e2)  # radians, metres
        slat = np.sin(llh[0])
        re = sc.a / np.sqrt(1.0 - sc.e2 * slat ** 2)
        rn = sc.a * (1.0 - sc.e2) / (1.0 - sc.e2 * slat ** 2) ** 1.5
        rcurv = re * rn / (re * np.cos(sc.peg_heading) ** 2 + rn * np.sin(sc.peg_heading) ** 2)  # curvature.F:26-64
        up = np.array([np.cos(llh[0]) * np.cos(llh[1]), np.cos(llh[0]) * np.sin(llh[1]), np.sin(llh[0])])
        centre = _ecef(np.degrees(llh[0]), np.degrees(llh[1]), 0.0, sc.a, sc.e2) - rcurv * up
        d = P[line] - S
        c = centre - S
        ang = np.degrees(np.arccos(np.clip((d @ c) / (np.linalg.norm(d, axis=1) * np.linalg.norm(c)), -1, 1))).astype(np.float32)
        sh = np.zeros(sc.width, bool)
        run = ang[0]
        for i in range(1, sc.width):
            if ang[i] <= run:
                sh[i] = True
            else:
                run = ang[i]
        run = ang[-1]
        for i in range(sc.width - 2, -1, -1):
            if ang[i] >= run:
                sh[i] = True
            else:
                run = ang[i]
        mism += int((sh != ((m[line] & 1) == 1)).sum())
    # measured: 0 of 65536 (995 shadowed); a last-bit difference of the two float32 angle sequences could move an edge
    assert mism <= 8, mism
    # no relief, no fold-over, no shadow
    flat = synth.make_scene(6, 1024)
    flat.dem = np.full_like(flat.dem, 250.0)
    assert not orc.topo(**orc.scene_topo_kwargs(flat, dem_method="BILINEAR", want_mask=True))["mask"].any()


def _sphere_centre(sc, S):
    llh = orc.xyz_to_latlon(S, sc.a, sc.e2)
    slat = np.sin(llh[0])
    re = sc.a / np.sqrt(1.0 - sc.e2 * slat ** 2)
    rn = sc.a * (1.0 - sc.e2) / (1.0 - sc.e2 * slat ** 2) ** 1.5
    rcurv = re * rn / (re * np.cos(sc.peg_heading) ** 2 + rn * np.sin(sc.peg_heading) ** 2)
    up = np.array([np.cos(llh[0]) * np.cos(llh[1]), np.cos(llh[0]) * np.sin(llh[1]), np.sin(llh[0])])
    return _ecef(np.degrees(llh[0]), np.degrees(llh[1]), 0.0, sc.a, sc.e2) - rcurv * up


def test_layover_bit_against_an_independent_reconstruction():
    """Layover (mask bit 2, topozero.f90:729-865): the terrain under the line is sampled on a regular cross-track grid; where
    the slant range of those samples is not monotonic in the cross-track position two ground points share a range, and the
    radar pixels at those ranges are flagged.  Rebuilt here from the output layers with numpy's own interpolation, an own
    grid (the imaged extent only, no extrapolated margin) and a plain bilinear DEM: the flagged pixels must overlap the
    oracle's to IoU > 0.9 (measured 0.958; the remaining difference is the grid / margin / search-rounding detail of the
    reference that the oracle restates and this reconstruction does not)."""
    sc = pu.rough_scene(16, 4096)
    o = orc.topo(**orc.scene_topo_kwargs(sc, dem_method="BILINEAR", want_mask=True))
    m = o["mask"]
    orb = orc.Orbit(sc.orbit_t, sc.orbit_pos, sc.orbit_vel)
    P = _ecef(o["lat"], o["lon"], o["hgt"], sc.a, sc.e2)
    rng = sc.r0 + sc.dr * np.arange(sc.width)
    w, ow = sc.width, 2 * sc.width + 1
    inter = union = n_ref = 0
    for line in range(sc.length):
        _, S, V = orb.interp(sc.t0 + line / sc.prf, "HERMITE")
        d, c = P[line] - S, _sphere_centre(sc, S) - S
        cth = (d @ c) / (np.linalg.norm(d, axis=1) * np.linalg.norm(c))
        ct = rng * np.sqrt(1.0 - cth ** 2)  # cross-track position of every pixel (topozero.f90:662)
        order = np.argsort(ct, kind="stable")
        cs = ct[order]
        grid = np.linspace(cs[0], cs[-1], ow)
        glat, glon = np.interp(grid, cs, o["lat"][line][order]), np.interp(grid, cs, o["lon"][line][order])
        h = _bilinear(sc.dem, (glat - sc.first_lat) / sc.delta_lat, (glon - sc.first_lon) / sc.delta_lon)
        orng = np.linalg.norm(_ecef(glat, glon, h, sc.a, sc.e2) - S, axis=1)
        ro = np.argsort(orng, kind="stable")
        ctr, osr = grid[ro], orng[ro]
        flag = np.zeros(ow, bool)
        flag[1:w] |= ctr[1:w] <= np.maximum.accumulate(ctr)[:w - 1]  # forward scan, bounded by width as in the reference (:835)
        run = ctr[-1]
        for i in range(ow - 2, -1, -1):  # backward scan; forward-flagged samples reset it (:847)
            if ctr[i] >= run and not flag[i]:
                flag[i] = True
            else:
                run = ctr[i]
        lay = np.zeros(w, bool)
        lay[np.clip(np.searchsorted(rng, osr[flag], side="right"), 1, w - 1) - 1] = True
        ref = (m[line] & 2) == 2
        inner = np.s_[2:-2]  # the reference maps out-of-swath samples of its margin onto the edge pixels
        inter += int((lay[inner] & ref[inner]).sum())
        union += int((lay[inner] | ref[inner]).sum())
        n_ref += int(ref[inner].sum())
    assert n_ref > 1000 and inter / union > 0.9, (n_ref, inter / union)


def test_geo2rdr_on_a_secondary_orbit_against_an_independent_root_finder():
    """configs[1]: offsets against a perturbed secondary orbit.  For sampled pixels the zero-Doppler time is found again by
    bracketing + bisection on (P - S(t)) . V(t) = 0 (no Newton step, no derivative), the range from the interpolated state
    at that time; the oracle's azimuth time must agree to 1e-8 s and its range to 1e-5 m (1e-3 px = 2e-6 s / 2.3e-3 m)."""
    sc = synth.make_scene(40, 2048)
    o = orc.topo(**orc.scene_topo_kwargs(sc, dem_method="BILINEAR", want_mask=False))
    sec = synth.config_c1_secondary(length=sc.length, width=sc.width)
    kw = pu.secondary_kwargs(sc, sec, recenter=0.37)
    g = orc.geo2rdr(lat=o["lat"], lon=o["lon"], hgt=o["hgt"], **kw)
    gb = orc.geo2rdr(lat=o["lat"], lon=o["lon"], hgt=o["hgt"], bistatic=True, **kw)
    orb = orc.Orbit(sec.orbit_t, sec.orbit_pos, sec.orbit_vel)
    P = _ecef(o["lat"], o["lon"], o["hgt"], sc.a, sc.e2)

    def f(t, p):
        _, S, V = orb.interp(t, "HERMITE")
        return float((p - S) @ V)

    rng = np.random.default_rng(11)
    checked = 0
    for line, pix in zip(rng.integers(0, sc.length, 120), rng.integers(0, sc.width, 120)):
        if g["azt"][line, pix] == -999999.0:
            continue
        p = P[line, pix]
        lo, hi = g["azt"][line, pix] - 0.5, g["azt"][line, pix] + 0.5
        assert f(lo, p) > 0 > f(hi, p)  # the satellite approaches, then recedes
        for _ in range(60):
            mid = 0.5 * (lo + hi)
            if f(mid, p) > 0:
                lo = mid
            else:
                hi = mid
        t = 0.5 * (lo + hi)
        _, S, _ = orb.interp(t, "HERMITE")
        assert abs(t - g["azt"][line, pix]) < 1e-8, (line, pix, t - g["azt"][line, pix])
        assert abs(np.linalg.norm(p - S) - g["rgm"][line, pix]) < 1e-5
        # offsets are relative to the 0-based line / sample (geo2rdr.f90:374-375)
        assert abs((t - kw["t0"]) * kw["prf"] - line - g["azoff"][line, pix]) < 1e-4
        assert abs((np.linalg.norm(p - S) - kw["r0"]) / kw["dr"] - pix - g["rgoff"][line, pix]) < 1e-4
        # bistatic delay correction (geo2rdr.f90:331-368): the echo is attributed to the time shifted by the two-way
        # travel time, the range is that of the state interpolated there
        if gb["azt"][line, pix] != -999999.0:
            tb = t + 2.0 * np.linalg.norm(p - S) / 299792458.0
            _, Sb, _ = orb.interp(tb, "HERMITE")
            assert abs(tb - gb["azt"][line, pix]) < 1e-8 and abs(np.linalg.norm(p - Sb) - gb["rgm"][line, pix]) < 1e-5
        checked += 1
    assert checked > 80


def test_dem_interpolators_against_independent_implementations():
    """BIQUINTIC (spline.f:15-117 as called at topozeroMethods.f:196) is a separable natural cubic spline over the 6 x 6
    window floor-1 .. floor+4 evaluated between its 2nd and 3rd node: scipy's CubicSpline(bc_type='natural') gives the
    same float32 value on every draw.  BILINEAR against the textbook formula, NEAREST against rounding; outside their
    windows all return -1000 (topozeroMethods.f:112)."""
    from scipy.interpolate import CubicSpline
    rng = np.random.default_rng(0)
    dem = (rng.normal(size=(40, 50)) * 100).astype(np.float32)
    for _ in range(1500):
        ix, iy = int(rng.integers(3, 47)), int(rng.integers(3, 37))  # 1-based cell indices
        fx, fy = rng.random(2)
        rows = [CubicSpline(np.arange(6), dem[r, ix - 2:ix + 4].astype(np.float64), bc_type="natural")(1.0 + fx)
                for r in range(iy - 2, iy + 4)]
        want = np.float32(CubicSpline(np.arange(6), np.array(rows), bc_type="natural")(1.0 + fy))
        assert orc.interp_dem("BIQUINTIC", dem, ix, iy, fx, fy) == want
        d = dem.astype(np.float64)
        bl = ((1 - fy) * ((1 - fx) * d[iy - 1, ix - 1] + fx * d[iy - 1, ix]) + fy * ((1 - fx) * d[iy, ix - 1] + fx * d[iy, ix]))
        assert abs(orc.interp_dem("BILINEAR", dem, ix, iy, fx, fy) - bl) <= 2e-5 * max(1.0, abs(bl))
        assert orc.interp_dem("NEAREST", dem, ix, iy, fx, fy) == dem[iy - 1 + int(round(fy)), ix - 1 + int(round(fx))]
    for method, bad in (("BIQUINTIC", (2, 10)), ("BIQUINTIC", (48, 10)), ("BILINEAR", (50, 10)), ("BILINEAR", (0, 10))):
        assert orc.interp_dem(method, dem, bad[0], bad[1], 0.5, 0.5) == -1000.0


def test_polynomial_reproduction_of_the_six_dem_interpolators_as_written():
    """What each interpolator of the reference does to constants and planes (the restatement keeps every quirk, DESIGN
    section 8): constants come back exactly from all but SINC (its eight-tap table is not renormalised per phase: ~0.6 %
    ripple); planes come back to float32 rounding from BILINEAR, AKIMA and BIQUINTIC and, along x only, from BICUBIC, whose
    dzdy column typo (uniform_interp.f90:152-154) shows as an error proportional to the y slope; NEAREST and the
    one-cell-shifted SINC do not reproduce planes."""
    ny, nx = 40, 50
    yy, xx = np.mgrid[0:ny, 0:nx].astype(np.float64)
    rng = np.random.default_rng(1)
    pts = [(int(rng.integers(6, nx - 6)), int(rng.integers(6, ny - 6)), rng.random(), rng.random()) for _ in range(200)]

    def worst(method, field, truth):
        dem = field.astype(np.float32)
        return max(abs(orc.interp_dem(method, dem, ix, iy, fx, fy) - truth(ix - 1 + fx, iy - 1 + fy)) for ix, iy, fx, fy in pts)

    const = (np.full((ny, nx), 123.5), lambda x, y: 123.5)
    lin_x = (10 + 2.0 * xx, lambda x, y: 10 + 2 * x)
    lin_y = (10 - 3.0 * yy, lambda x, y: 10 - 3 * y)
    for m in ("BILINEAR", "BICUBIC", "NEAREST", "AKIMA", "BIQUINTIC"):
        assert worst(m, *const) == 0.0, m
    assert 0.1 < worst("SINC", *const) < 1.5
    for m in ("BILINEAR", "AKIMA", "BIQUINTIC"):
        assert worst(m, *lin_x) < 1e-5 and worst(m, *lin_y) < 1e-5, m
    assert worst("BICUBIC", *lin_x) < 1e-5 and 0.05 < worst("BICUBIC", *lin_y) < 0.5
    assert worst("NEAREST", *lin_x) > 0.5 and worst("SINC", *lin_x) > 0.5
