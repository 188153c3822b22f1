"""Deterministic synthetic inputs for the zero-Doppler geometry path.

Keplerian orbit -> ECEF state vectors (the hand-off format of
``isceobj.Orbit.Orbit.exportToC``, components/isceobj/Orbit/Orbit.py:1060-1081),
fractal 1-arcsec float32 DEM (ISCE DEM georeferencing, components/isceobj/Image/
Image.py:711-733), and the sensor parameters of the BASELINE.json configs
(SURVEY.md section 8d).  Pure numpy; used by tests, bench.py and the smoke test.
"""
from __future__ import annotations

import dataclasses
import datetime
import math

import numpy as np

GM = 3.986004418e14
OMEGA_E = 7.292115e-5
WGS84_A = 6378137.0
WGS84_E2 = 0.0066943799901


# ----------------------------------------------------------------------------
# orbit
# ----------------------------------------------------------------------------
def keplerian_state_vectors(t_mid, t_lo, t_hi, *, step=10.0, a=7071e3, inc_deg=98.18,
                            sublat_deg=36.0, lon0_deg=-118.0, da=0.0, d_cross=0.0, d_along_s=0.0,
                            round_pos=1e-6):
    """Circular Keplerian orbit sampled every ``step`` s on [t_lo, t_hi] (seconds of day).

    Returns (t[n], pos[n,3], vel[n,3]) in ECEF.  ``da`` = radial offset (m), ``d_cross`` =
    cross-track offset (m) realised as a small node rotation, ``d_along_s`` = along-track
    phase offset (s): the "perturbed secondary" of BASELINE config 1.
    """
    a = a + da
    n = math.sqrt(GM / a ** 3)
    inc = math.radians(inc_deg)
    u0 = math.asin(math.sin(math.radians(sublat_deg)) / math.sin(inc))
    # node chosen so that the sub-satellite longitude at t_mid is lon0
    x0 = math.cos(u0)
    y0 = math.sin(u0) * math.cos(inc)
    node = math.radians(lon0_deg) - math.atan2(y0, x0) + d_cross / a
    k0 = math.ceil(t_lo / step)
    k1 = math.floor(t_hi / step)
    t = np.arange(k0, k1 + 1, dtype=np.float64) * step
    u = u0 + n * (t - t_mid + d_along_s)
    cu, su = np.cos(u), np.sin(u)
    cO, sO, ci, si = math.cos(node), math.sin(node), math.cos(inc), math.sin(inc)
    r_in = a * np.stack([cu * cO - su * ci * sO, cu * sO + su * ci * cO, su * si], axis=1)
    v_in = a * n * np.stack([-su * cO - cu * ci * sO, -su * sO + cu * ci * cO, cu * si], axis=1)
    th = OMEGA_E * (t - t_mid)
    c, s = np.cos(th), np.sin(th)
    pos = np.stack([c * r_in[:, 0] + s * r_in[:, 1], -s * r_in[:, 0] + c * r_in[:, 1], r_in[:, 2]], axis=1)
    vrot = np.stack([c * v_in[:, 0] + s * v_in[:, 1], -s * v_in[:, 0] + c * v_in[:, 1], v_in[:, 2]], axis=1)
    vel = vrot - np.stack([-OMEGA_E * pos[:, 1], OMEGA_E * pos[:, 0], np.zeros_like(t)], axis=1)
    if round_pos:
        pos = np.round(pos / round_pos) * round_pos
        vel = np.round(vel / round_pos) * round_pos
    return np.ascontiguousarray(t), np.ascontiguousarray(pos), np.ascontiguousarray(vel)


# ----------------------------------------------------------------------------
# minimal geodesy (host-side helpers for scene construction only)
# ----------------------------------------------------------------------------
def llh_to_xyz(lat_deg, lon_deg, h, a=WGS84_A, e2=WGS84_E2):
    lat, lon = np.radians(lat_deg), np.radians(lon_deg)
    re = a / np.sqrt(1.0 - e2 * np.sin(lat) ** 2)
    return np.stack([(re + h) * np.cos(lat) * np.cos(lon), (re + h) * np.cos(lat) * np.sin(lon),
                     (re * (1.0 - e2) + h) * np.sin(lat)], axis=-1)


def xyz_to_llh(xyz, a=WGS84_A, e2=WGS84_E2):
    """Closed-form ECEF -> geodetic (same formula family as latlon.F:51-71), degrees."""
    x, y, z = xyz[..., 0], xyz[..., 1], xyz[..., 2]
    q2 = x * x + y * y
    a2, e4 = a * a, e2 * e2
    p = q2 / a2
    q = (1.0 - e2) * z * z / a2
    r = (p + q - e4) / 6.0
    s = e4 * p * q / (4.0 * r ** 3)
    t = np.cbrt(1.0 + s + np.sqrt(s * (2.0 + s)))
    u = r * (1.0 + t + 1.0 / t)
    rv = np.sqrt(u * u + e4 * q)
    w = e2 * (u + rv - q) / (2.0 * rv)
    k = np.sqrt(u + rv + w * w) - w
    d = k * np.sqrt(q2) / (k + e2)
    lat = np.arctan2(z, d)
    lon = np.arctan2(y, x)
    h = (k + e2 - 1.0) * np.sqrt(d * d + z * z) / k
    return np.degrees(lat), np.degrees(lon), h


def _ground_point(pos, vel, rng, h, side, a=WGS84_A, e2=WGS84_E2):
    """Zero-Doppler ground point at slant range rng and ellipsoid height h (bisection on look angle)."""
    nhat = -pos / np.linalg.norm(pos)
    chat = np.cross(nhat, vel)
    chat /= np.linalg.norm(chat)
    vhat = vel / np.linalg.norm(vel)
    down = nhat - np.dot(nhat, vhat) * vhat
    down /= np.linalg.norm(down)
    lo, hi = 0.0, math.radians(80.0)
    for _ in range(80):
        th = 0.5 * (lo + hi)
        look = math.cos(th) * down - side * math.sin(th) * chat
        _, _, hh = xyz_to_llh(pos + rng * look, a, e2)
        if hh < h:
            lo = th  # point is below the surface: look angle too small
        else:
            hi = th
    return pos + rng * (math.cos(th) * down - side * math.sin(th) * chat)


def hermite_point(t, pos, vel, tq):
    """Cubic-Hermite orbit sample at tq using the two bracketing vectors (scene construction only)."""
    i = int(np.clip(np.searchsorted(t, tq) - 1, 0, len(t) - 2))
    h = t[i + 1] - t[i]
    s = (tq - t[i]) / h
    h00, h10 = 2 * s ** 3 - 3 * s ** 2 + 1, s ** 3 - 2 * s ** 2 + s
    h01, h11 = -2 * s ** 3 + 3 * s ** 2, s ** 3 - s ** 2
    p = h00 * pos[i] + h10 * h * vel[i] + h01 * pos[i + 1] + h11 * h * vel[i + 1]
    d00, d10 = (6 * s ** 2 - 6 * s) / h, 3 * s ** 2 - 4 * s + 1
    d01, d11 = (-6 * s ** 2 + 6 * s) / h, 3 * s ** 2 - 2 * s
    v = d00 * pos[i] + d10 * vel[i] + d01 * pos[i + 1] + d11 * vel[i + 1]
    return p, v


def enu_heading_rad(pos, vel, a=WGS84_A, e2=WGS84_E2):
    """Heading of the velocity in the local ENU frame (Orbit.getENUHeading, Orbit.py:804-831), radians."""
    lat, lon, _ = xyz_to_llh(np.asarray(pos), a, e2)
    lat, lon = math.radians(float(lat)), math.radians(float(lon))
    east = np.array([-math.sin(lon), math.cos(lon), 0.0])
    north = np.array([-math.sin(lat) * math.cos(lon), -math.sin(lat) * math.sin(lon), math.cos(lat)])
    return math.atan2(float(np.dot(east, vel)), float(np.dot(north, vel)))


# ----------------------------------------------------------------------------
# DEM
# ----------------------------------------------------------------------------
def fractal_tile(n=2048, beta=1.9, seed=20261017):
    """Periodic n x n fractal surface (2-D spectral synthesis, amplitude ~ f^-beta), scaled to [0,1]."""
    rng = np.random.default_rng(seed)
    fy = np.fft.fftfreq(n)[:, None]
    fx = np.fft.rfftfreq(n)[None, :]
    f = np.sqrt(fx * fx + fy * fy)
    f[0, 0] = 1.0
    amp = f ** (-beta)
    amp[0, 0] = 0.0
    phase = rng.uniform(0.0, 2.0 * np.pi, size=amp.shape)
    spec = amp * np.exp(1j * phase)
    z = np.fft.irfft2(spec, s=(n, n))
    z -= z.min()
    z /= z.max()
    return z


def fractal_dem(nlat, nlon, *, hmin=-100.0, hmax=2000.0, beta=1.9, seed=20261017, tile=2048):
    """float32 DEM [nlat][nlon] built by wrapping a periodic fractal tile (seamless)."""
    z = fractal_tile(tile, beta, seed)
    iy = np.arange(nlat) % tile
    ix = np.arange(nlon) % tile
    dem = z[np.ix_(iy, ix)]
    return np.ascontiguousarray((hmin + (hmax - hmin) * dem).astype(np.float32))


# ----------------------------------------------------------------------------
# scenes
# ----------------------------------------------------------------------------
@dataclasses.dataclass
class Scene:
    name: str
    length: int
    width: int
    wvl: float
    dr: float
    r0: float
    prf: float
    side: int  # ISCE lookSide: -1 right, +1 left
    t0: float
    sensing_start: datetime.datetime
    orbit_t: np.ndarray
    orbit_pos: np.ndarray
    orbit_vel: np.ndarray
    peg_heading: float
    doppler_coeffs: list  # Poly2D coeffs [az][rg] (Hz vs range pixel)
    dem: np.ndarray
    first_lat: float
    first_lon: float
    delta_lat: float
    delta_lon: float
    a: float = WGS84_A
    e2: float = WGS84_E2
    nrnglooks: int = 1
    nazlooks: int = 1

    @property
    def pixels(self):
        return self.length * self.width


SENSORS = {
    # Sentinel-1 IW-like
    "s1": dict(wvl=0.05546576, dr=2.329562, r0=800e3, prf=486.486, side=-1, a=7071e3, inc_deg=98.18,
               doppler=[[0.0]]),
    # NISAR-like L-band stripmap with native Doppler
    "nisar": dict(wvl=0.238, dr=6.25, r0=900e3, prf=1650.0, side=+1, a=7125e3, inc_deg=98.4,
                  doppler=[[-120.0, 2.0e-2, -1.5e-6, 3e-11]]),
}


def make_scene(length, width, *, sensor="s1", dem_spacing_arcsec=1.0, beta=1.9, hmin=-100.0, hmax=2000.0,
               seed=20261017, sv_step=10.0, sv_margin=65.0, name=None, perturb=None, t0=21600.0,
               dem=True):
    """Build a synthetic scene.  ``perturb`` = dict(da=, d_cross=, d_along_s=) gives a secondary orbit."""
    s = SENSORS[sensor]
    dur = (length - 1) / s["prf"]
    t_mid = t0 + 0.5 * dur
    pert = perturb or {}
    # reference orbit (unperturbed) defines the footprint / DEM; the perturbed one is what is returned
    t, pos, vel = keplerian_state_vectors(t_mid, t0 - sv_margin, t0 + dur + sv_margin, step=sv_step, a=s["a"],
                                          inc_deg=s["inc_deg"], **pert)
    tr, posr, velr = keplerian_state_vectors(t_mid, t0 - sv_margin, t0 + dur + sv_margin, step=sv_step, a=s["a"],
                                             inc_deg=s["inc_deg"])
    pm, vm = hermite_point(t, pos, vel, t0 + 0.5 * length / s["prf"])
    peg = enu_heading_rad(pm, vm)
    first_lat = first_lon = 0.0
    dlat = -dem_spacing_arcsec / 3600.0
    dlon = dem_spacing_arcsec / 3600.0
    demarr = np.zeros((2, 2), np.float32)
    if dem:
        lats, lons = [], []
        for tq in (t0, t0 + dur):
            p, v = hermite_point(tr, posr, velr, tq)
            for rg in (s["r0"], s["r0"] + (width - 1) * s["dr"]):
                for h in (-500.0, 9000.0):
                    la, lo, _ = xyz_to_llh(_ground_point(p, v, rg, h, s["side"]))
                    lats.append(float(la))
                    lons.append(float(lo))
        pad = 0.2
        # snap the DEM origin onto the posting grid (like a real SRTM mosaic)
        first_lat = math.ceil((max(lats) + pad) / abs(dlat)) * abs(dlat)
        first_lon = math.floor((min(lons) - pad) / dlon) * dlon
        nlat = int(math.ceil((first_lat - (min(lats) - pad)) / abs(dlat))) + 1
        nlon = int(math.ceil(((max(lons) + pad) - first_lon) / dlon)) + 1
        demarr = fractal_dem(nlat, nlon, hmin=hmin, hmax=hmax, beta=beta, seed=seed)
    start = datetime.datetime(2026, 10, 17) + datetime.timedelta(seconds=t0)
    return Scene(name=name or f"{sensor}_{length}x{width}", length=length, width=width, wvl=s["wvl"], dr=s["dr"],
                 r0=s["r0"], prf=s["prf"], side=s["side"], t0=t0, sensing_start=start, orbit_t=t, orbit_pos=pos,
                 orbit_vel=vel, peg_heading=peg, doppler_coeffs=[list(r) for r in s["doppler"]], dem=demarr,
                 first_lat=first_lat, first_lon=first_lon, delta_lat=dlat, delta_lon=dlon)


# BASELINE.json configs.  Full sizes; tests pass smaller length/width.
def config_c0(length=1500, width=21000, **kw):
    return make_scene(length, width, sensor="s1", name="C0_s1_burst_topo", **kw)


def config_c1_secondary(length=1500, width=21000, seed=7, **kw):
    """Secondary acquisition of C1: perturbed orbit; misregistration applied by the caller as in
    contrib/stack/topsStack/geo2rdr.py:90-91 (sensingStart - misreg_az, startingRange - misreg_rg)."""
    rng = np.random.default_rng(seed)
    j = rng.uniform(-1.0, 1.0, 3)
    pert = dict(da=120.0 + 5.0 * j[0], d_cross=80.0 + 5.0 * j[1], d_along_s=0.37 + 0.01 * j[2])
    return make_scene(length, width, sensor="s1", name=f"C1_secondary_seed{seed}", perturb=pert, dem=False, **kw)


def config_c2(length=13500, width=25000, **kw):
    return make_scene(length, width, sensor="s1", name="C2_s1_swath", **kw)


def config_c3(length=60000, width=25000, **kw):
    return make_scene(length, width, sensor="nisar", name="C3_nisar_frame", **kw)


def make_tops_acquisition(n_swaths=2, n_bursts=3, burst_lines=60, burst_samples=900, overlap_lines=8, overlap_samples=60,
                          swath_lag_lines=5, sv_per_burst=None, **scene_kw):
    """A TOPS-like acquisition cut out of one synthetic scene: `n_swaths` sub-swaths (each later in time and farther in
    range than the previous one, overlapping by `overlap_samples`), `n_bursts` bursts each (overlapping by
    `overlap_lines`).  Returns (scene, frames); a frame has ``.bursts`` and every burst the attributes the reference reads
    from a ``BurstSLC`` (components/isceobj/Sensor/TOPS/BurstSLC.py): sensingStart/Stop, startingRange, farRange,
    numberOfLines/Samples, rangePixelSize, azimuthTimeInterval, radarWavelength, orbit.  The union grid of the frames
    (TopsProc/runTopo.py:159-172) is exactly the scene's grid."""
    from types import SimpleNamespace

    from .orbit import Orbit, StateVector

    az_step = burst_lines - overlap_lines
    rg_step = burst_samples - overlap_samples
    length = swath_lag_lines * (n_swaths - 1) + az_step * (n_bursts - 1) + burst_lines
    width = rg_step * (n_swaths - 1) + burst_samples
    sc = make_scene(length, width, **scene_kw)
    day = sc.sensing_start.replace(hour=0, minute=0, second=0, microsecond=0)
    dt = 1.0 / sc.prf
    svs = [StateVector(day + datetime.timedelta(seconds=float(t)), p, v) for t, p, v in zip(sc.orbit_t, sc.orbit_pos, sc.orbit_vel)]
    frames = []
    k = 0
    for s in range(n_swaths):
        bursts = []
        for b in range(n_bursts):
            top = s * swath_lag_lines + b * az_step
            left = s * rg_step
            orb = Orbit()
            if sv_per_burst:  # annotation-style orbits: every burst carries only part of the state vectors
                lo = (k * 2) % max(1, len(svs) - sv_per_burst + 1)
                for sv in svs[lo:lo + sv_per_burst]:
                    orb.addStateVector(sv)
            else:
                for sv in svs:
                    orb.addStateVector(sv)
            k += 1
            start = sc.sensing_start + datetime.timedelta(seconds=top * dt)
            bursts.append(SimpleNamespace(
                sensingStart=start, sensingStop=start + datetime.timedelta(seconds=(burst_lines - 1) * dt),
                startingRange=sc.r0 + left * sc.dr, farRange=sc.r0 + (left + burst_samples - 1) * sc.dr,
                numberOfLines=burst_lines, numberOfSamples=burst_samples, rangePixelSize=sc.dr, azimuthTimeInterval=dt,
                radarWavelength=sc.wvl, orbit=orb, window=(top, top + burst_lines, left, left + burst_samples)))
        frames.append(SimpleNamespace(bursts=bursts, numberOfBursts=n_bursts,
                                      sensingStart=bursts[0].sensingStart, sensingStop=bursts[-1].sensingStop,
                                      startingRange=bursts[0].startingRange, farRange=bursts[0].farRange))
    return sc, frames
