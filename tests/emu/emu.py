"""Host emulation of the device per-pixel functions (development aid, tests only)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libemu_kernels.so")


class EmuTopoArgs(C.Structure):
    _fields_ = [("a", C.c_double), ("e2", C.c_double), ("wvl", C.c_double), ("thresh", C.c_double),
                ("ilrl", C.c_int), ("numiter", C.c_int), ("extraiter", C.c_int),
                ("ufirstlat", C.c_double), ("ufirstlon", C.c_double), ("deltalat", C.c_double), ("deltalon", C.c_double),
                ("nx", C.c_int), ("ny", C.c_int), ("method", C.c_int), ("width", C.c_int), ("length", C.c_int),
                ("nazlooks", C.c_int), ("t0", C.c_double), ("prf", C.c_double), ("peghdg", C.c_double),
                ("orbit_method", C.c_int), ("n_orbit", C.c_int), ("dop_range_order", C.c_int),
                ("dop_azimuth_order", C.c_int), ("r0", C.c_double), ("dr", C.c_double),
                ("line0", C.c_int), ("nlines", C.c_int), ("want_inc", C.c_int), ("use_ref", C.c_int)]


def build():
    src = os.path.join(HERE, "emu_kernels.cpp")
    hdr_dir = os.path.join(HERE, "..", "..", "isce2_b200", "csrc")
    newest = max([os.path.getmtime(src)] + [os.path.getmtime(os.path.join(hdr_dir, f)) for f in os.listdir(hdr_dir)
                                            if f.endswith(".cuh")])
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < newest:
        subprocess.check_call(["/usr/bin/g++", "-O2", "-ffp-contract=off", "-mfma", "-fPIC", "-shared", "-std=c++17", "-o", LIB, src])
    return C.CDLL(LIB)


def topo(sc, crop, *, dem_method=1, orbit_method=0, want_inc=True, numiter=25, extraiter=10, thresh=0.05, line0=0, nlines=-1,
         use_ref=True):
    """sc: synth.Scene; crop: dict with ustartx, ustarty, udemwidth, udemlength, ufirstlat, ufirstlon (from the oracle)."""
    L = build()
    x0, y0, nx, ny = crop["ustartx"], crop["ustarty"], crop["udemwidth"], crop["udemlength"]
    dem = np.ascontiguousarray(sc.dem[y0 - 1:y0 - 1 + ny, x0 - 1:x0 - 1 + nx], np.float32)
    n = sc.length - line0 if nlines < 0 else nlines
    dop = np.ascontiguousarray(np.atleast_2d(np.asarray(sc.doppler_coeffs, np.float64)))
    A = EmuTopoArgs(sc.a, sc.e2, sc.wvl, thresh, sc.side, numiter, extraiter, crop["ufirstlat"], crop["ufirstlon"],
                    sc.delta_lat, sc.delta_lon, nx, ny, dem_method, sc.width, sc.length, sc.nazlooks, sc.t0, sc.prf,
                    sc.peg_heading, orbit_method, len(sc.orbit_t), dop.shape[1] - 1, dop.shape[0] - 1, sc.r0,
                    sc.dr * sc.nrnglooks, line0, n, int(want_inc), int(use_ref))
    w = sc.width
    out = dict(lat=np.empty((n, w)), lon=np.empty((n, w)), hgt=np.empty((n, w)), los=np.empty((n, 2, w), np.float32),
               inc=np.empty((n, 2, w), np.float32), ctrack=np.empty((n, w)), elev=np.empty((n, w), np.float32))
    it = C.c_longlong()
    dp = C.POINTER(C.c_double)
    fp = C.POINTER(C.c_float)
    t = np.ascontiguousarray(sc.orbit_t)
    pos = np.ascontiguousarray(sc.orbit_pos)
    vel = np.ascontiguousarray(sc.orbit_vel)
    L.emu_topo(C.byref(A), dem.ctypes.data_as(fp), t.ctypes.data_as(dp), pos.ctypes.data_as(dp), vel.ctypes.data_as(dp),
               dop.ctypes.data_as(dp), out["lat"].ctypes.data_as(dp), out["lon"].ctypes.data_as(dp),
               out["hgt"].ctypes.data_as(dp), out["los"].ctypes.data_as(fp), out["inc"].ctypes.data_as(fp),
               out["ctrack"].ctypes.data_as(dp), out["elev"].ctypes.data_as(fp), C.byref(it))
    out["iters"] = it.value
    return out
