"""CPU-side checks of the drop-in boundary: the C-ABI library builds for sm_100a, loads, exports every
symbol include/b200geom.h declares, and refuses to compute without a device (no CPU fallback)."""
import os
import re

import numpy as np
import pytest

from isce2_b200 import _capi, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "b200geom.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", hdr)))


def test_header_symbols_are_exported():
    lib = _capi.lib()
    declared = _declared_symbols()
    assert len(declared) >= 19
    for sym in declared:
        assert hasattr(lib, sym), f"{sym} declared in include/b200geom.h but not exported by libb200geom.so"
    assert sorted(_capi.EXPORTS) == declared
    assert lib.b200_abi_version() == 7  # 2: geozero + resamp_slc; 3: fused topo+geo2rdr; 4: looks + mask projection; 5: frozen stack geometry; 6: d2h floor, pageable sinks; 7: file-backed destinations


def test_struct_layouts_match_header():
    """ctypes mirrors of the ABI structs must have the field order of the header (sizes are checked natively by
    passing them through; here we pin the field names)."""
    hdr = open(os.path.join(ROOT, "include", "b200geom.h")).read()

    def fields(struct_name):
        chunk = [c for c in hdr.split("typedef struct {") if ("} " + struct_name + ";") in c][0]
        body = chunk.split("} " + struct_name + ";")[0]
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        names = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            decl = re.sub(r"^(const\s+)?(unsigned\s+)?(long long|double|float|int8_t|int|void|b200_[a-z0-9_]+)\s*", "", decl)
            names += [n.strip().lstrip("*") for n in decl.split(",")]
        return names

    for cname, ctype in (("b200_topo_params", _capi.TopoParams), ("b200_topo_result", _capi.TopoResult),
                         ("b200_geo_params", _capi.GeoParams), ("b200_geo_result", _capi.GeoResult),
                         ("b200_topo_outputs", _capi.TopoOutputs), ("b200_geo_outputs", _capi.GeoOutputs),
                         ("b200_orbit", _capi.Orbit), ("b200_poly2d", _capi.Poly2d), ("b200_poly1d", _capi.Poly1d),
                         ("b200_geozero_params", _capi.GeozeroParams), ("b200_geozero_result", _capi.GeozeroResult),
                         ("b200_resamp_params", _capi.ResampParams), ("b200_resamp_result", _capi.ResampResult),
                         ("b200_geo_job", _capi.GeoJob), ("b200_looks_result", _capi.LooksResult),
                         ("b200_mask_result", _capi.MaskResult)):
        assert fields(cname) == [f[0] for f in ctype._fields_], cname


@pytest.mark.skipif(_capi.device_count() > 0, reason="a CUDA device is present")
def test_no_cpu_fallback():
    sc = synth.make_scene(4, 64, dem_spacing_arcsec=3.0)
    p = _capi.topo_params(dem_shape=sc.dem.shape, first_lat=sc.first_lat, first_lon=sc.first_lon, delta_lat=sc.delta_lat,
                          delta_lon=sc.delta_lon, length=sc.length, width=sc.width, prf=sc.prf, t0=sc.t0, wvl=sc.wvl,
                          side=sc.side, peg_heading=sc.peg_heading)
    with pytest.raises(_capi.B200Error) as ei:
        _capi.topo_run(p, sc.dem, sc.orbit_t, sc.orbit_pos, sc.orbit_vel, sc.doppler_coeffs, [[sc.r0, sc.dr]])
    assert ei.value.code == -2 and "no CPU fallback" in str(ei.value)
    gp = _capi.geo_params(length=4, width=64, dem_shape=(4, 64), r0=sc.r0, dr=sc.dr, prf=sc.prf, t0=sc.t0, wvl=sc.wvl)
    z = np.zeros((4, 64))
    with pytest.raises(_capi.B200Error) as ei:
        _capi.geo2rdr_run(gp, z, z, z, sc.orbit_t, sc.orbit_pos, sc.orbit_vel)
    assert ei.value.code == -2
    with pytest.raises(_capi.B200Error) as ei:  # the fused verb has no host route either
        _capi.topo_geo2rdr_run(p, sc.dem, sc.orbit_t, sc.orbit_pos, sc.orbit_vel, sc.doppler_coeffs,
                               [dict(params=gp, orbit=(sc.orbit_t, sc.orbit_pos, sc.orbit_vel))], [[sc.r0, sc.dr]])
    assert ei.value.code == -2
    zp = _capi.geozero_params(dem_shape=sc.dem.shape, first_lat=sc.first_lat, first_lon=sc.first_lon, delta_lat=sc.delta_lat,
                              delta_lon=sc.delta_lon, snwe=(sc.first_lat - 0.05, sc.first_lat - 0.01, sc.first_lon + 0.01,
                                                            sc.first_lon + 0.05),
                              length=4, width=64, r0=sc.r0, dr=sc.dr, prf=sc.prf, t0=sc.t0, wvl=sc.wvl)
    assert min(_capi.geozero_grid(zp)) > 0  # sizing the output grid needs no device
    with pytest.raises(_capi.B200Error) as ei:
        _capi.geozero_run(zp, sc.dem, sc.orbit_t, sc.orbit_pos, sc.orbit_vel, np.zeros((4, 64), np.float32))
    assert ei.value.code == -2
    with pytest.raises(_capi.B200Error) as ei:
        _capi.resamp_slc_run(np.zeros((16, 16), np.complex64), (16, 16))
    assert ei.value.code == -2


def test_argument_validation_precedes_device_use():
    sc = synth.make_scene(4, 64, dem_spacing_arcsec=3.0)
    p = _capi.topo_params(dem_shape=sc.dem.shape, first_lat=sc.first_lat, first_lon=sc.first_lon, delta_lat=sc.delta_lat,
                          delta_lon=sc.delta_lon, length=sc.length, width=sc.width, prf=sc.prf, t0=sc.t0, wvl=sc.wvl,
                          side=sc.side, peg_heading=sc.peg_heading, orbit_method="LEGENDRE")
    # fewer than 9 state vectors with LEGENDRE: the reference prints and stops (topozero.f90:123-126)
    with pytest.raises(_capi.B200Error) as ei:
        _capi.topo_run(p, sc.dem, sc.orbit_t[:5], sc.orbit_pos[:5], sc.orbit_vel[:5], sc.doppler_coeffs, [[sc.r0, sc.dr]])
    assert ei.value.code == -4 and "9 state vectors" in str(ei.value)
    p.dem_method = 7  # not one of the reference's six methods: "Undefined interpolation method." (topozero.f90:96-99)
    p.orbit_method = 0
    with pytest.raises(_capi.B200Error) as ei:
        _capi.topo_run(p, sc.dem, sc.orbit_t, sc.orbit_pos, sc.orbit_vel, sc.doppler_coeffs, [[sc.r0, sc.dr]])
    assert ei.value.code == -1


def test_file_backed_destination_registry(tmp_path):
    """b200_host_file_register / _unregister keep a table of address ranges that are file mappings (no GPU involved):
    overlaps and bad arguments are refused, image.file_backed registers writable memmaps only and takes everything back."""
    import numpy as np

    from isce2_b200 import _capi
    from isce2_b200 import image as IF
    a = np.memmap(str(tmp_path / "a.bin"), dtype=np.float64, mode="w+", shape=(64, 512))
    b = np.memmap(str(tmp_path / "b.bin"), dtype=np.float32, mode="w+", shape=(64, 2, 512))
    ro = np.memmap(str(tmp_path / "a.bin"), dtype=np.float64, mode="r", shape=(64, 512))
    fd = os.open(str(tmp_path / "a.bin"), os.O_RDWR)
    try:
        _capi.host_file_register(a.ctypes.data, a.nbytes, fd, 0)
        with pytest.raises(_capi.B200Error):  # overlapping range
            _capi.host_file_register(a.ctypes.data + 4096, 4096, fd, 4096)
        with pytest.raises(_capi.B200Error):
            _capi.host_file_register(b.ctypes.data, b.nbytes, -1, 0)
        assert _capi.host_file_unregister(a.ctypes.data)
        assert not _capi.host_file_unregister(a.ctypes.data)
    finally:
        os.close(fd)
    with IF.file_backed([a, None, b, ro, np.zeros(8), a]):
        assert not _capi.host_file_unregister(ro.ctypes.data)  # read-only mappings and plain arrays are not declared
        with pytest.raises(_capi.B200Error):                   # a and b are
            fd = os.open(str(tmp_path / "b.bin"), os.O_RDWR)
            try:
                _capi.host_file_register(b.ctypes.data, b.nbytes, fd, 0)
            finally:
                os.close(fd)
    assert not _capi.host_file_unregister(a.ctypes.data) and not _capi.host_file_unregister(b.ctypes.data)
    # inputs: read-only mappings and plain-ndarray views of mappings (what numpy.ascontiguousarray hands back) are declared
    view = np.ascontiguousarray(ro[8:40])
    assert type(view) is np.ndarray and IF._file_range(view, False) == (view.ctypes.data, view.nbytes, str(tmp_path / "a.bin"), 8 * 512 * 8)
    assert IF._file_range(view, True) is None and IF._file_range(np.zeros(4), False) is None
    assert IF._file_range(ro[:, 3:9], False) is None  # not contiguous
    with IF.file_backed([b], inputs=[view, None]):
        assert _capi.host_file_unregister(view.ctypes.data)
    os.environ["B200_FILE_WRITES"] = "0"
    try:
        with IF.file_backed([a]):  # switched off: nothing is declared
            assert not _capi.host_file_unregister(a.ctypes.data)
    finally:
        del os.environ["B200_FILE_WRITES"]
