"""Merged-swath topo + per-burst views + stack geo2rdr on the GPU, checked against the CPU oracle.
Call shapes: components/isceobj/TopsProc/runTopo.py:114-359 (runTopoGPU) and contrib/stack/topsStack/geo2rdr.py:51-104."""
import datetime
import os
from types import SimpleNamespace

import numpy as np
import pytest

from isce2_b200 import image as IF, swath, synth
from isce2_b200.orbit import Orbit
from oracle import oracle as orc
from tests import parity_util as pu
from tests.test_gpu_components import _write_dem

pytestmark = pytest.mark.gpu


def test_merged_topo_views_and_stack_geo2rdr(tmp_path):
    sc, frames = synth.make_tops_acquisition(n_swaths=2, n_bursts=3, burst_lines=40, burst_samples=1500, overlap_lines=6,
                                             overlap_samples=100, swath_lag_lines=4, beta=1.6, hmax=2500.0)
    dem, demf = _write_dem(sc, str(tmp_path / "dem.dem"))
    geom = str(tmp_path / "geom_reference")
    out = swath.run_topo_merged(frames, dem, geom, swaths=[1, 2], swath_starts=[0, 2], inc=True, mask=True)
    g = out["grid"]
    assert (g.length, g.width) == (sc.length, sc.width)

    # the merged layers equal the oracle run on the union grid with runTopoGPU's settings (peg heading at the first line)
    peg = np.radians(out["orbit"].getENUHeading(g.t0))
    sc.dem = demf
    sc.peg_heading = peg
    c = pu.cpu_topo(sc, dem_method="BIQUINTIC")
    lat = np.fromfile(os.path.join(geom, "lat.rdr")).reshape(sc.length, sc.width)
    lon = np.fromfile(os.path.join(geom, "lon.rdr")).reshape(sc.length, sc.width)
    hgt = np.fromfile(os.path.join(geom, "hgt.rdr")).reshape(sc.length, sc.width)
    msk = np.fromfile(os.path.join(geom, "shadowMask.rdr"), np.int8).reshape(sc.length, sc.width)
    n = lat.size
    assert np.abs(lat - c["lat"]).max() < 1e-7 and (np.abs(lat - c["lat"]) > pu.TOL_LATLON_DEG).sum() <= max(2, n // 50000)
    assert np.abs(lon - c["lon"]).max() < 1e-7 and (np.abs(lon - c["lon"]) > pu.TOL_LATLON_DEG).sum() <= max(2, n // 50000)
    assert np.abs(hgt - c["hgt"]).max() < pu.TOL_HGT_M
    assert np.array_equal(msk, c["mask"])
    assert abs(out["bbox"][0] - c["min_lat"]) < 1e-9 and abs(out["bbox"][3] - c["max_lon"]) < 1e-9

    # burst views: numbering follows swath_starts, every view is the window of the merged layer
    assert sorted(out["windows"]) == [(1, 1), (1, 2), (1, 3), (2, 3), (2, 4), (2, 5)]
    los = np.fromfile(os.path.join(geom, "los.rdr"), np.float32).reshape(sc.length, 2, sc.width)
    for (sw, num), (top, bottom, left, right) in out["windows"].items():
        d = os.path.join(geom, "IW%d" % sw)
        assert np.array_equal(np.asarray(IF.read_view(os.path.join(d, "lat_%02d.rdr" % num))), lat[top:bottom, left:right])
        assert np.array_equal(np.asarray(IF.read_view(os.path.join(d, "hgt_%02d.rdr" % num))), hgt[top:bottom, left:right])
        assert np.array_equal(np.asarray(IF.read_view(os.path.join(d, "los_%02d.rdr" % num))), los[top:bottom, :, left:right])
        assert np.array_equal(np.asarray(IF.read_view(os.path.join(d, "shadowMask_%02d.rdr" % num))), msk[top:bottom, left:right])
        hdr = IF.createImage().load(os.path.join(d, "incLocal_%02d.rdr.xml" % num))
        assert (hdr.width, hdr.length, hdr.bands, hdr.dataType) == (right - left, bottom - top, 2, "FLOAT")

    # ---- stack: 3 secondary dates x 6 bursts against the resident per-burst geometry ----
    st = swath.Geo2rdrStack(devices=[0])
    keys = {}
    for (sw, num) in out["windows"]:
        d = os.path.join(geom, "IW%d" % sw)
        key = "IW%d/%02d" % (sw, num)
        st.add_geometry(key, os.path.join(d, "lat_%02d.rdr" % num), os.path.join(d, "lon_%02d.rdr" % num),
                        os.path.join(d, "hgt_%02d.rdr" % num))
        keys[(sw, num)] = key
    day = sc.sensing_start.replace(hour=0, minute=0, second=0, microsecond=0)
    jobs = []
    for date in range(3):
        sec = synth.config_c1_secondary(length=sc.length, width=sc.width, seed=date + 1, beta=1.6, hmax=2500.0)
        orb = Orbit.from_arrays(day, sec.orbit_t, sec.orbit_pos, sec.orbit_vel)
        misreg_az, misreg_rg = 0.3 * (date + 1), 0.5 * date
        for f, sw, istart in zip(frames, [1, 2], [0, 2]):
            for ind, b in enumerate(f.bursts):
                # the secondary burst images the same ground 0.37 s "earlier" on its own clock (synthetic along-track shift)
                info = SimpleNamespace(rangePixelSize=b.rangePixelSize, azimuthTimeInterval=b.azimuthTimeInterval,
                                       radarWavelength=b.radarWavelength, orbit=orb, numberOfSamples=b.numberOfSamples,
                                       numberOfLines=b.numberOfLines, startingRange=b.startingRange,
                                       sensingStart=b.sensingStart - datetime.timedelta(seconds=0.37))
                num = ind + istart + 1
                od = tmp_path / ("coreg_secondarys/date%d/IW%d" % (date, sw))
                os.makedirs(od, exist_ok=True)
                job = dict(key=keys[(sw, num)], info=info, rg=str(od / ("range_%02d.off" % num)),
                           az=str(od / ("azimuth_%02d.off" % num)), misreg_az=misreg_az, misreg_rg=misreg_rg, sec=sec,
                           window=out["windows"][(sw, num)])
                st.add_job(job["key"], info, job["rg"], job["az"], misreg_az=misreg_az, misreg_rg=misreg_rg)
                jobs.append(job)
    res = st.run()
    assert len(res) == len(jobs) == 18
    nvalid = 0
    for job in jobs:
        top, bottom, left, right = job["window"]
        info, sec = job["info"], job["sec"]
        start = info.sensingStart - datetime.timedelta(seconds=job["misreg_az"] * info.azimuthTimeInterval)
        t0 = (start - day).total_seconds()
        o = orc.geo2rdr(lat=np.ascontiguousarray(lat[top:bottom, left:right]), lon=np.ascontiguousarray(lon[top:bottom, left:right]),
                        hgt=np.ascontiguousarray(hgt[top:bottom, left:right]), orbit_t=sec.orbit_t, orbit_pos=sec.orbit_pos,
                        orbit_vel=sec.orbit_vel, length=info.numberOfLines, width=info.numberOfSamples,
                        r0=info.startingRange - job["misreg_rg"], dr=info.rangePixelSize, prf=1.0 / info.azimuthTimeInterval,
                        t0=t0, wvl=info.radarWavelength, side=-1)
        rg = np.fromfile(job["rg"], np.float32).reshape(bottom - top, right - left)
        az = np.fromfile(job["az"], np.float32).reshape(bottom - top, right - left)
        bad = np.float32(-999999.0)
        assert np.array_equal(rg == bad, o["rgoff"] == -999999.0)
        v = rg != bad
        nvalid += int(v.sum())
        if v.any():
            assert np.abs(rg[v] - o["rgoff"][v]).max() < pu.TOL_OFFSET_PX
            assert np.abs(az[v] - o["azoff"][v]).max() < pu.TOL_OFFSET_PX
        hdr = IF.createImage().load(job["rg"] + ".xml")
        assert (hdr.dataType, hdr.width, hdr.length) == ("FLOAT", right - left, bottom - top)
        assert os.path.exists(job["az"] + ".vrt")
    assert nvalid > 0.3 * sum((j["window"][1] - j["window"][0]) * (j["window"][3] - j["window"][2]) for j in jobs)
