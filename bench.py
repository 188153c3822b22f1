#!/usr/bin/env python
"""bench.py -- headline benchmark of the zero-Doppler geometry hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host CPU cores

Metric (BASELINE.json): topo+geo2rdr Mpixels/s.  Workload (BASELINE.json configs[2]): full 9-burst Sentinel-1 IW
swath, 13500 x 25000 radar pixels, synthetic Keplerian orbit + fractal 1-arcsec DEM, BIQUINTIC DEM interpolation,
incidence layer and layover/shadow mask on, followed by geo2rdr of the resulting lat/lon/hgt against a perturbed
secondary orbit (range/azimuth offsets, float32).  One "step" = one pass of topo + geo2rdr over the whole swath; a
pixel counts once when it has been through both.

  value : device-resident throughput (DEM, orbit and the previous layers already in HBM; CUDA events on the launch
          stream, summed over the kernels of the step; max over ranks).
  e2e   : the same step through the C ABI with HOST buffers (pinned), host<->device copies inside the timed region:
          one b200_topo_geo2rdr_run call (host DEM and orbits in; every topo layer and the offsets out to the host;
          geo2rdr runs on the layers while they are resident in HBM).  `e2e_two_calls` is the reference's own call
          sequence b200_topo_run -> host lat/lon/hgt -> b200_geo2rdr_run, which sends the 24 B/pixel back up.
  N > 1 : one process per GPU (torchrun); the swath is sharded by contiguous azimuth line blocks, no collective on
          the data path (torch.distributed/gloo is used only for the timing barrier and the max over ranks).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from isce2_b200 import synth  # noqa: E402

# unit-weight algorithmic FP64 operations per pixel of the REFERENCE algorithm
W1_TOPO_ITER = {"BILINEAR": 260.0, "BIQUINTIC": 670.0}
W1_TOPO_FINAL = {"BILINEAR": 380.0, "BIQUINTIC": 2020.0}
W1_MASK = {"BILINEAR": 200.0, "BIQUINTIC": 1000.0}
W1_GEO_BASE, W1_GEO_ITER = 35.0, {"HERMITE": 520.0, "LEGENDRE": 215.0}
FP64_NOMINAL_TFLOPS = 37.2  # 148 SMs x 64 DFMA/clk x 2 x 1.965 GHz

WORKLOADS = {
    # BASELINE.json configs[2]
    "c2": dict(length=13500, width=25000, sensor="s1", dem_method="BIQUINTIC", orbit_method="HERMITE", inc=True, mask=True,
               desc="S1 IW 9-burst swath 13500x25000, topo(BIQUINTIC,+inc,+layover/shadow mask) + geo2rdr(perturbed secondary, f32 offsets)"),
    # configs[0] + configs[1] (single burst, the reference's CPU-runnable case)
    "c0c1": dict(length=1500, width=21000, sensor="s1", dem_method="BILINEAR", orbit_method="HERMITE", inc=False, mask=False,
                 desc="S1 IW burst 1500x21000, topo(BILINEAR) + geo2rdr(perturbed secondary, f32 offsets)"),
    # configs[3]
    "c3": dict(length=60000, width=25000, sensor="nisar", dem_method="BIQUINTIC", orbit_method="LEGENDRE", inc=True, mask=True,
               desc="NISAR-like L-band frame 60000x25000, native Doppler, Legendre orbit, topo(BIQUINTIC,+inc,+mask) + geo2rdr"),
}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# --------------------------------------------------------------------------------------------------
# rank plumbing (no torch unless WORLD_SIZE > 1)
# --------------------------------------------------------------------------------------------------
class Ranks:
    def __init__(self):
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.dist = None
        if self.world > 1:
            import torch
            import torch.distributed as dist
            import datetime
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            # a finite timeout: ranks that fall out of step raise instead of waiting for each other forever (long enough
            # for the idle ranks of the --impl reference arm, which wait at the final barrier while rank 0 computes)
            dist.init_process_group(backend="gloo", rank=self.rank, world_size=self.world,
                                    timeout=datetime.timedelta(seconds=1000))
            self.dist = dist
            self.torch = torch

    def barrier(self):
        if self.dist:
            self.dist.barrier()

    def reduce_max(self, x):
        if not self.dist:
            return float(x)
        t = self.torch.tensor([float(x)], dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t[0])

    def reduce_sum(self, x):
        if not self.dist:
            return float(x)
        t = self.torch.tensor([float(x)], dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t[0])

    def close(self, ok=True):
        if self.dist:
            if ok:  # after an error the other ranks are not at this barrier: leave without it
                self.dist.barrier()
            self.dist.destroy_process_group()


# --------------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region: NVML polled every 10 ms from a thread (the timed
    region of a multi-GPU run lasts ~0.1 s, too short for `nvidia-smi -lms`), nvidia-smi as the fallback."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, device):
        self.rows = []   # (time, sm_mhz, max_mhz, power_w, set(reasons))
        self.proc = None
        self.stop_flag = False
        self.source = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates by PCI bus id; CUDA_VISIBLE_DEVICES is not set by torchrun, so ordinals agree
            self.h = pynvml.nvmlDeviceGetHandleByIndex(int(device))
            self.nv = pynvml
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.source = "nvml"
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.source = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-i", str(device), "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                                         text=True)
            self.source = "nvidia-smi"
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _poll(self):
        nv = self.nv
        bits = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        while not self.stop_flag:
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                except Exception:
                    mask = 0
                try:
                    pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1e3
                except Exception:
                    pw = 0.0
                self.rows.append((time.time(), sm, self.max_mhz, pw, {k for k, b in bits.items() if mask & b}))
            except Exception:
                pass
            time.sleep(0.01)

    def _pump(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.strip().split(",")]
            if len(parts) < 7:
                continue
            try:
                self.rows.append((time.time(), float(parts[0]), float(parts[1]), float(parts[2]),
                                  {nm for nm, v in zip(self.NAMES, parts[3:7]) if v.lower().startswith("active")}))
            except ValueError:
                continue

    def stop(self, t_lo, t_hi):
        self.stop_flag = True
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        if self.source is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml / nvidia-smi unavailable"]}
        rows = [r for r in self.rows if t_lo <= r[0] <= t_hi + 0.05]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples inside the timed region"], "source": self.source}
        reasons = set()
        for r in rows:
            reasons |= r[4]
        return {"sm_mhz": float(np.median([r[1] for r in rows])), "sm_max_mhz": float(max(r[2] for r in rows)),
                "power_w_max": float(max(r[3] for r in rows)), "samples": len(rows), "reasons": sorted(reasons),
                "source": self.source}


# --------------------------------------------------------------------------------------------------
# workload
# --------------------------------------------------------------------------------------------------
def build_workload(name, lines_override=None):
    w = dict(WORKLOADS[name])
    if lines_override:
        w["length"] = int(lines_override)
    t0 = time.time()
    sc = synth.make_scene(w["length"], w["width"], sensor=w["sensor"], name=name)
    sec = synth.make_scene(w["length"], w["width"], sensor=w["sensor"], dem=False,
                           perturb=dict(da=120.0, d_cross=80.0, d_along_s=0.37))
    log(f"[bench] scene {name}: {sc.length}x{sc.width}, DEM {sc.dem.shape} ({sc.dem.nbytes / 1e6:.0f} MB) built in {time.time() - t0:.1f}s")
    return w, sc, sec


def secondary_geo_kwargs(sc, sec):
    # contrib/stack/topsStack/geo2rdr.py:90-91 with misreg_az = 0.013 s, misreg_rg = 1.7 m
    return dict(orbit_t=sec.orbit_t, orbit_pos=sec.orbit_pos, orbit_vel=sec.orbit_vel, length=sc.length, width=sc.width,
                r0=sc.r0 - 1.7, dr=sc.dr, prf=sc.prf, t0=sc.t0 - 0.013, wvl=sc.wvl, side=sc.side)


def shard(length, rank, world):
    a = (length * rank) // world
    b = (length * (rank + 1)) // world
    return a, b - a


# --------------------------------------------------------------------------------------------------
# CPU baseline (the oracle == C restatement of the reference Fortran/C path, all host threads)
# --------------------------------------------------------------------------------------------------
def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def use_all_host_cores():
    """torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm is meant to use every host core, so the OpenMP
    thread count is set explicitly (libgomp is the runtime the oracle links against)."""
    import ctypes
    n = host_cores()
    try:
        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(n)
    except OSError:
        pass
    return n


def cpu_sample(w, sc, sec, lines, line0=None):
    from oracle import oracle as orc
    use_all_host_cores()
    if line0 is None:
        line0 = max(0, sc.length // 2 - lines // 2)
    lines = min(lines, sc.length - line0)
    t0 = time.perf_counter()
    out = orc.topo(**orc.scene_topo_kwargs(sc, dem_method=w["dem_method"], orbit_method=w["orbit_method"], want_inc=w["inc"],
                                           want_mask=w["mask"], line0=line0, nlines=lines))
    t1 = time.perf_counter()
    kw = secondary_geo_kwargs(sc, sec)
    # geo2rdr works on the sample's rows only: present them as a `lines`-row image whose first row is radar line line0
    kw["t0"] = kw["t0"] + line0 / sc.prf
    kw["length"] = sc.length - line0
    g = orc.geo2rdr(lat=out["lat"], lon=out["lon"], hgt=out["hgt"], orbit_method=w["orbit_method"], want=("azoff", "rgoff"), **kw)
    t2 = time.perf_counter()
    return dict(lines=lines, pixels=lines * sc.width, t_topo=t1 - t0, t_geo=t2 - t1, K=out["mean_iters"], N=g["mean_iters"])


def cpu_baseline(w, sc, sec, budget_s=15.0):
    cal = cpu_sample(w, sc, sec, 4)
    per_line = (cal["t_topo"] + cal["t_geo"]) / cal["lines"]
    lines = int(max(8, min(512, budget_s / max(per_line, 1e-6))))
    s = cpu_sample(w, sc, sec, lines)
    t = s["t_topo"] + s["t_geo"]
    return {"value": s["pixels"] / t / 1e6, "unit": "Mpixels/s", "cores": host_cores(), "kind": "port",
            "sample": f"{s['lines']} azimuth lines x {sc.width} samples from the middle of the swath "
                      f"(topo {s['t_topo']:.2f}s + geo2rdr {s['t_geo']:.2f}s), OpenMP over pixels as in the reference",
            "note": "C restatement of the ISCE2 Fortran/C reference (oracle/), gfortran is not available in this image",
            "K_topo_iters": s["K"], "N_geo_iters": s["N"]}


def run_reference(args, ranks):
    """--impl reference: the reference algorithm (oracle port) on the host cores, bounded sample per step."""
    if ranks.rank != 0:
        return
    from oracle import oracle as orc
    orc.build()
    w, sc, sec = build_workload(args.workload, args.lines)
    cal = cpu_sample(w, sc, sec, 4)
    per_line = (cal["t_topo"] + cal["t_geo"]) / cal["lines"]
    lines = int(max(4, min(256, args.ref_step_seconds / max(per_line, 1e-6))))
    for _ in range(args.warmup):
        cpu_sample(w, sc, sec, lines)
    t0 = time.perf_counter()
    px = 0
    for _ in range(args.steps):
        s = cpu_sample(w, sc, sec, lines)
        px += s["pixels"]
    dt = time.perf_counter() - t0
    val = px / dt / 1e6
    sample = f"each step = {lines} azimuth lines x {sc.width} samples from the middle of the swath (bounded sample of the workload)"
    line = {"impl": "reference", "metric": "topo+geo2rdr Mpixels/s", "value": val, "unit": "Mpixels/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": w["desc"], "pixels_per_step_full": sc.pixels, "dem_method": w["dem_method"],
                       "orbit_method": w["orbit_method"]},
            "cpu_baseline": {"value": val, "unit": "Mpixels/s", "cores": host_cores(), "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "Mpixels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# this repo's arm
# --------------------------------------------------------------------------------------------------
def alloc_host(shape, dtype, capi):
    try:
        return capi.pinned_empty(shape, dtype), True
    except Exception:
        return np.empty(shape, dtype), False


def bind_to_gpu_local_cpus(dev):
    """Multi-GPU runs: keep the rank (and therefore its page-locked staging buffers, first-touched below) on the CPUs
    NVML reports as local to its GPU, so that host<->device copies do not cross the socket interconnect."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(dev))
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * i + b for i, wd in enumerate(words) for b in range(64) if (int(wd) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus and len(cpus) < len(os.sched_getaffinity(0)):
            os.sched_setaffinity(0, cpus)
            return sorted(cpus)
    except Exception:
        pass
    return None


def run_b200(args, ranks):
    from isce2_b200 import _capi as capi
    if capi.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device visible; this arm has no CPU fallback (use --impl reference for the CPU baseline)")
    dev = ranks.local_rank % capi.device_count()
    if ranks.world > 1:
        bound = bind_to_gpu_local_cpus(dev)
        if bound:
            log(f"[bench] rank {ranks.rank}: bound to the {len(bound)} CPUs local to GPU {dev}")
    w, sc, sec = build_workload(args.workload, args.lines)
    line0, nlines = shard(sc.length, ranks.rank, ranks.world)
    npix_local = nlines * sc.width
    npix_total = sc.length * sc.width
    fp64_peak = capi.fp64_peak(dev)
    log(f"[bench] rank {ranks.rank}/{ranks.world} device {dev} ({capi.device_name(dev)}), lines [{line0},{line0 + nlines}), "
        f"FP64 peak (DFMA microbenchmark) {fp64_peak:.1f} TFLOP/s")

    tparams = capi.topo_params(dem_shape=sc.dem.shape, first_lat=sc.first_lat, first_lon=sc.first_lon, delta_lat=sc.delta_lat,
                               delta_lon=sc.delta_lon, length=sc.length, width=sc.width, prf=sc.prf, t0=sc.t0, wvl=sc.wvl,
                               side=sc.side, peg_heading=sc.peg_heading, a=sc.a, e2=sc.e2, dem_method=w["dem_method"],
                               orbit_method=w["orbit_method"], line0=line0, nlines=nlines, device=dev)
    gk = secondary_geo_kwargs(sc, sec)
    gparams = capi.geo_params(length=gk["length"], width=gk["width"], dem_shape=(sc.length, sc.width), r0=gk["r0"], dr=gk["dr"],
                              prf=gk["prf"], t0=gk["t0"], wvl=gk["wvl"], side=gk["side"], orbit_method=w["orbit_method"],
                              line0=line0, nlines=nlines, device=dev, out_f32=True)
    slr = [[sc.r0, sc.dr * sc.nrnglooks]]

    # ---------------- device-resident arm ----------------
    dem_host, _ = alloc_host(sc.dem.shape, np.float32, capi)
    dem_host[...] = sc.dem
    tplan = capi.TopoPlan(tparams, dem_host, sc.orbit_t, sc.orbit_pos, sc.orbit_vel, sc.doppler_coeffs, slr,
                          want_los=True, want_inc=w["inc"], want_mask=w["mask"])
    tplan.execute()  # layers must exist before geo2rdr borrows them
    gplan = capi.GeoPlan(gparams, topo_plan=tplan)

    def device_step():
        ms_t = tplan.execute()
        ms_g = gplan.execute(gparams, gk["orbit_t"], gk["orbit_pos"], gk["orbit_vel"], want=("azoff", "rgoff"))
        return ms_t, ms_g

    for _ in range(args.warmup):
        device_step()
    ranks.barrier()
    sampler = ClockSampler(dev)
    t_lo = time.time()
    w0 = time.perf_counter()
    ms_topo = ms_geo = 0.0
    for _ in range(args.steps):
        a, b = device_step()
        ms_topo += a
        ms_geo += b
    wall_dev = time.perf_counter() - w0
    t_hi = time.time()
    clocks = sampler.stop(t_lo, t_hi)
    ranks.barrier()
    # per-kernel split and iteration statistics of the last step
    # (the layers stay on the device: fetch only the result structs)
    import ctypes as C
    res = capi.TopoResult()
    e = C.create_string_buffer(512)
    capi._check(capi.lib().b200_topo_plan_fetch(tplan.handle, None, C.byref(res), e, 512), e)
    gres = capi.GeoResult()
    capi._check(capi.lib().b200_geo_plan_fetch(gplan.handle, None, C.byref(gres), e, 512), e)
    K = res.iterations / float(npix_local)
    Ngeo = gres.iterations / float(npix_local)
    ms_step_local = (ms_topo + ms_geo) / args.steps
    ms_step = ranks.reduce_max(ms_step_local)
    value = npix_total / (ms_step * 1e-3) / 1e6
    # the resident layers (52 B/pixel + offsets) go back to the workspace cache before the end-to-end arm allocates its own
    gplan.close()
    tplan.close()

    # ---------------- end-to-end arm: reference-facing C-ABI calls with host buffers ----------------
    outs = {}
    pinned = True
    for k, shp, dt in (("lat", (nlines, sc.width), np.float64), ("lon", (nlines, sc.width), np.float64),
                       ("hgt", (nlines, sc.width), np.float64), ("los", (nlines, 2, sc.width), np.float32),
                       ("inc", (nlines, 2, sc.width), np.float32), ("mask", (nlines, sc.width), np.int8)):
        if (k == "inc" and not w["inc"]) or (k == "mask" and not w["mask"]):
            outs[k] = None
            continue
        outs[k], pin = alloc_host(shp, dt, capi)
        pinned = pinned and pin
    gouts = {"azt": None, "rgm": None}
    for k in ("azoff", "rgoff"):
        gouts[k], pin = alloc_host((nlines, sc.width), np.float32, capi)
        pinned = pinned and pin
    # geo2rdr reads the block's rows of lat/lon/hgt from the host buffers topo just filled
    gparams_e2e = capi.geo_params(length=gk["length"] - line0, width=gk["width"], dem_shape=(nlines, sc.width), r0=gk["r0"],
                                  dr=gk["dr"], prf=gk["prf"], t0=gk["t0"] + line0 / sc.prf, wvl=gk["wvl"], side=gk["side"],
                                  orbit_method=w["orbit_method"], device=dev, out_f32=True)

    def e2e_step_two_calls():
        # the reference's own call sequence: topo verb -> host lat / lon / hgt -> geo2rdr verb
        capi.topo_run(tparams, dem_host, sc.orbit_t, sc.orbit_pos, sc.orbit_vel, sc.doppler_coeffs, slr,
                      want_los=True, want_inc=w["inc"], want_mask=w["mask"], out=outs)
        r = capi.geo2rdr_run(gparams_e2e, outs["lat"], outs["lon"], outs["hgt"], gk["orbit_t"], gk["orbit_pos"], gk["orbit_vel"],
                             want=("azoff", "rgoff"), out=gouts)
        return r

    fused_job = dict(params=gparams, orbit=(gk["orbit_t"], gk["orbit_pos"], gk["orbit_vel"]), want=("azoff", "rgoff"), out=gouts)

    def e2e_step_fused():
        # one call: same host inputs, same host outputs; geo2rdr runs on the layers while they are resident in HBM
        _, geos = capi.topo_geo2rdr_run(tparams, dem_host, sc.orbit_t, sc.orbit_pos, sc.orbit_vel, sc.doppler_coeffs, [fused_job],
                                        slr, want_los=True, want_inc=w["inc"], want_mask=w["mask"], out=outs)
        return geos[0]

    e2e_steps = max(1, min(args.steps, args.e2e_steps))

    def time_e2e(step, tag):
        for _ in range(args.warmup):
            step()
        ranks.barrier()
        times = []
        for _ in range(e2e_steps):
            w0 = time.perf_counter()
            r = step()
            times.append(time.perf_counter() - w0)
        log(f"[bench] rank {ranks.rank} e2e ({tag}) step times (ms): {[round(1e3 * t, 1) for t in times]}")
        wall = ranks.reduce_max(sum(times) / e2e_steps)
        ranks.barrier()
        return wall, r

    wall_two, r = time_e2e(e2e_step_two_calls, "b200_topo_run + b200_geo2rdr_run")
    check = {k: v.copy() for k, v in gouts.items() if v is not None} if npix_local <= 64_000_000 else None
    wall_e2e, r_fused = time_e2e(e2e_step_fused, "b200_topo_geo2rdr_run")
    # The fused call must leave the same offsets in the host buffers as the two calls.  On the block that starts at line
    # 0 they are bit-identical; the two-call arm of the other ranks describes its block with a re-based sensing start
    # and line count (so that geo2rdr can read the block's rows as a whole image), which moves the last bits.  Reported,
    # never fatal: every rank must reach the collectives below.
    e2e_diff, e2e_valid_equal = 0.0, float(r_fused["num_valid"] == r["num_valid"])
    if check is not None:
        for k, v in check.items():
            bad_a, bad_b = v == np.float32(-999999.0), gouts[k] == np.float32(-999999.0)
            e2e_valid_equal = min(e2e_valid_equal, float(np.array_equal(bad_a, bad_b)))
            both = ~bad_a & ~bad_b
            if both.any():
                e2e_diff = max(e2e_diff, float(np.abs(v[both].astype(np.float64) - gouts[k][both].astype(np.float64)).max()))
    e2e_diff = ranks.reduce_max(e2e_diff)
    e2e_valid_equal = -ranks.reduce_max(-e2e_valid_equal)
    e2e_value = npix_total / wall_e2e / 1e6
    small = res.dem_nx * res.dem_ny * 4 + 2 * 7 * 8 * len(sc.orbit_t)
    d2h = sum(v.nbytes for v in outs.values() if v is not None) + sum(v.nbytes for v in gouts.values() if v is not None)
    h2d_two = ranks.reduce_sum(small + 3 * 8 * npix_local)
    h2d = ranks.reduce_sum(small)
    d2h = ranks.reduce_sum(d2h)
    valid_frac = r["num_valid"] / float(npix_local)

    # ---------------- roofline of the dominant kernel ----------------
    # Work model: unit-weight FP64 operations of the REFERENCE algorithm per pixel.  The heavy DEM
    # interpolators run the solve and the final pass as two kernels (k_topo_solve dominant), the light ones fused.
    dm = w["dem_method"]
    split = dm in ("BIQUINTIC", "BICUBIC")
    w1_solve = W1_TOPO_ITER[dm] * K
    w1_final = W1_TOPO_FINAL[dm]
    w1_mask = W1_MASK[dm] if w["mask"] else 0.0
    w1_geo = W1_GEO_BASE + W1_GEO_ITER[w["orbit_method"]] * 9.0  # reference needs N = 9 steps
    w1_step = w1_solve + w1_final + w1_mask + w1_geo
    ms_solve = res.ms_solve if split else res.ms_pixels
    dom_name = f"k_topo_solve<{dm}>" if split else f"k_topo_fused<{dm}>"
    dom_w1 = w1_solve if split else (w1_solve + w1_final)
    achieved = dom_w1 * npix_local / (ms_solve * 1e-3) / 1e12
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    bytes_px = 8 if split else (24 + 8 + (8 if w["inc"] else 0))  # solve kernel: the SCH height it hands over
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        traffic = tj.get(dom_name, {}).get("dram_bytes_per_pixel")
        traffic = traffic * npix_local if traffic is not None else None
    except Exception:
        pass

    def tf(w1, ms):
        return w1 * npix_local / (ms * 1e-3) / 1e12 if ms > 0 else None

    kernels = {dom_name: {"ms": ms_solve, "w1_per_pixel": dom_w1, "tflops_w1": achieved}}
    if split:
        kernels[f"k_topo_final<{dm}>"] = {"ms": res.ms_pixels - res.ms_solve, "w1_per_pixel": w1_final,
                                          "tflops_w1": tf(w1_final, res.ms_pixels - res.ms_solve)}
    if w["mask"]:
        kernels[f"k_topo_mask<{dm}>"] = {"ms": res.ms_mask, "w1_per_pixel": w1_mask, "tflops_w1": tf(w1_mask, res.ms_mask)}
    kernels["k_geo2rdr_poly"] = {"ms": gres.ms_kernels, "w1_per_pixel": w1_geo, "tflops_w1": tf(w1_geo, gres.ms_kernels),
                                 "note": f"solves the reference's equation in N={Ngeo:.2f} true-Newton steps on orbit polynomials; "
                                         "W1 counts the reference's 9 quasi-Newton steps with full Hermite re-interpolation"}
    roofline = {"kernel": dom_name, "bound": "fp64", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                "frac": achieved / fp64_peak, "traffic": traffic,
                "peak_source": "measured live: b200_fp64_peak DFMA microbenchmark (MEASURED_PEAKS.json has no FP64 entry)",
                "frac_of_nominal_37.2": achieved / FP64_NOMINAL_TFLOPS,
                "work_model": f"W1 (reference algorithm, unit-weight ops, SURVEY 8d): {dom_w1 / K:.0f}*K per pixel, K={K:.3f}"
                if split else f"W1: {W1_TOPO_ITER[dm]:.0f}*K + {W1_TOPO_FINAL[dm]:.0f} per pixel, K={K:.3f}",
                "avg_launch_ms": ms_solve, "pixels_per_launch": npix_local,
                "hbm": {"achieved_gbs": bytes_px * npix_local / (ms_solve * 1e-3) / 1e9, "peak_gbs": hbm_peak,
                        "frac": bytes_px * npix_local / (ms_solve * 1e-3) / 1e9 / hbm_peak, "bytes_per_pixel": bytes_px,
                        "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback (B200_PROFILING.md)"},
                "kernels": kernels}

    line = None
    if ranks.rank == 0:
        cb = cpu_baseline(w, sc, sec) if (ranks.world == 1 and not args.no_cpu_baseline) else None
        line = {"metric": "topo+geo2rdr Mpixels/s", "value": value, "unit": "Mpixels/s", "n_gpus": ranks.world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": w["desc"], "pixels_per_step": npix_total, "dem_method": dm,
                           "orbit_method": w["orbit_method"], "sharding": f"{ranks.world} contiguous azimuth line blocks, no collective",
                           "l2": "inputs (DEM crop + previous layers) and outputs exceed the 126 MB L2 between iterations",
                           "K_topo_iters_per_pixel": K, "N_geo_steps_per_pixel": Ngeo, "geo2rdr_valid_fraction": valid_frac,
                           "dem_crop": [res.dem_ny, res.dem_nx]},
                "e2e": {"value": e2e_value, "unit": "Mpixels/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                        "ms_per_step": wall_e2e * 1e3, "steps": e2e_steps, "pinned_host_buffers": bool(pinned),
                        "api": "b200_topo_geo2rdr_run (host DEM + orbits in; host lat/lon/hgt/los/inc/mask + range/azimuth "
                               "offsets out; geo2rdr consumes the layers in HBM)",
                        "vs_two_calls": {"max_abs_offset_diff_px": e2e_diff, "validity_equal": bool(e2e_valid_equal),
                                         "compared": check is not None}},
                "e2e_two_calls": {"value": npix_total / wall_two / 1e6, "unit": "Mpixels/s", "h2d_bytes_per_step": int(h2d_two),
                                  "d2h_bytes_per_step": int(d2h), "ms_per_step": wall_two * 1e3, "steps": e2e_steps,
                                  "api": "b200_topo_run + b200_geo2rdr_run (the reference's call sequence: lat/lon/hgt go "
                                         "back up through the host)"},
                "gpu_launches": int(args.steps * ((2 if split else 1) + (1 if w["mask"] else 0) + 2)),
                "clocks": clocks,
                "roofline": roofline,
                "cpu_baseline": cb,
                "breakdown_ms": {"topo_solve": ms_solve, "topo_pixels": res.ms_pixels, "topo_mask": res.ms_mask, "geo2rdr": gres.ms_kernels,
                                 "topo_step_avg": ms_topo / args.steps, "geo2rdr_step_avg": ms_geo / args.steps,
                                 "wall_device_step": wall_dev / args.steps * 1e3},
                "work_equivalent_tflops": {"value": w1_step * npix_total / (ms_step * 1e-3) / 1e12,
                                           "note": "reference-algorithm W1 ops of the whole step / device time; the CUDA path removes reference work (hoisted setup, constant spline factors), so this can exceed executed FLOP/s"}}
        print(json.dumps(line), flush=True)
    return line


def run_c4(args, ranks):
    """BASELINE configs[4]: topsStack geo2rdr batch -- one reference geometry (the C2 swath's lat/lon/hgt, computed once
    per GPU and kept resident) against 29 perturbed secondary orbits; the 29 jobs are dealt round-robin to the ranks
    (contrib/stack/topsStack/Stack.py:805-827 launches one process per secondary date).  One step = all 29 jobs."""
    from isce2_b200 import _capi as capi
    dev = ranks.local_rank % capi.device_count()
    w = dict(WORKLOADS["c2"])
    if args.lines:
        w["length"] = int(args.lines)
    sc = synth.make_scene(w["length"], w["width"], sensor="s1", name="c4")
    tparams = capi.topo_params(dem_shape=sc.dem.shape, first_lat=sc.first_lat, first_lon=sc.first_lon, delta_lat=sc.delta_lat,
                               delta_lon=sc.delta_lon, length=sc.length, width=sc.width, prf=sc.prf, t0=sc.t0, wvl=sc.wvl,
                               side=sc.side, peg_heading=sc.peg_heading, dem_method="BIQUINTIC", device=dev)
    tplan = capi.TopoPlan(tparams, sc.dem, sc.orbit_t, sc.orbit_pos, sc.orbit_vel, sc.doppler_coeffs, [[sc.r0, sc.dr]],
                          want_los=False, want_inc=False, want_mask=False)
    tplan.execute()
    jobs = [j for j in range(29) if j % ranks.world == ranks.rank]
    secs = {j: synth.config_c1_secondary(length=sc.length, width=sc.width, seed=j + 1) for j in jobs}
    gp = capi.geo_params(length=sc.length, width=sc.width, dem_shape=(sc.length, sc.width), r0=sc.r0 - 1.7, dr=sc.dr, prf=sc.prf,
                         t0=sc.t0 - 0.013, wvl=sc.wvl, side=sc.side, device=dev, out_f32=True)
    gplan = capi.GeoPlan(gp, topo_plan=tplan)
    # the reference geometry is fixed for the whole batch: its ECEF coordinates are formed once (outside the timed
    # region, like the topo run that produced it), not once per secondary date
    gplan.freeze_geometry()

    def step():
        ms = 0.0
        for j in jobs:
            o = secs[j]
            ms += gplan.execute(gp, o.orbit_t, o.orbit_pos, o.orbit_vel, want=("azoff", "rgoff"))
        return ms

    for _ in range(args.warmup):
        step()
    ranks.barrier()
    sampler = ClockSampler(dev)
    t_lo = time.time()
    ms = sum(step() for _ in range(args.steps)) / args.steps
    clocks = sampler.stop(t_lo, time.time())
    ms = ranks.reduce_max(ms)
    npx = 29 * sc.pixels
    if ranks.rank == 0:
        print(json.dumps({"metric": "geo2rdr Mpixels/s (29-orbit stack batch)", "value": npx / (ms * 1e-3) / 1e6, "unit": "Mpixels/s",
                          "n_gpus": ranks.world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
                          "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                          "config": {"workload": "topsStack batch: 29 secondary orbits x (13500 x 25000) geo2rdr on one resident "
                                                 "reference geometry, jobs dealt round-robin to the GPUs",
                                     "pixels_per_step": npx, "jobs_on_rank0": len(jobs)},
                          "gpu_launches": int(args.steps * 2 * len(jobs)), "clocks": clocks}), flush=True)
    gplan.close()
    tplan.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS) + ["c4"])
    ap.add_argument("--lines", type=int, default=None, help="override the number of azimuth lines (debugging)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--ref-step-seconds", type=float, default=6.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--max-seconds", type=float, default=1200.0, help="watchdog: abort the whole process after this long")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        log("[bench] note: W < 3 warm-up steps requested; the timing rules ask for >= 3")
    # watchdog: a bench that is still running after --max-seconds is stuck (e.g. ranks out of step); abort hard so that a
    # hung run cannot hold the GPUs
    killer = threading.Timer(args.max_seconds, lambda: (log(f"[bench] still running after {args.max_seconds:.0f} s: aborting"),
                                                        os._exit(124)))
    killer.daemon = True
    killer.start()
    ranks = Ranks()
    ok = False
    try:
        if args.workload == "c4" and args.impl == "b200":
            run_c4(args, ranks)
        elif args.impl == "reference":
            if args.workload == "c4":
                args.workload = "c2"
            run_reference(args, ranks)
        else:
            run_b200(args, ranks)
        ok = True
    finally:
        ranks.close(ok)
        killer.cancel()


if __name__ == "__main__":
    main()
