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
    line0 = max(0, min(line0, sc.length - 1))
    lines = min(lines, sc.length - line0)
    t0 = time.perf_counter()
    out = orc.topo(**orc.scene_topo_kwargs(sc, dem_method=w["dem_method"], orbit_method=w["orbit_method"], want_inc=w["inc"],
                                           want_mask=w["mask"], line0=line0, nlines=lines))
    t1 = time.perf_counter()
    kw = secondary_geo_kwargs(sc, sec)
    # geo2rdr works on the sample's rows only: present them as a `lines`-row image whose first row is radar line line0
    kw["t0"] = kw["t0"] + line0 / sc.prf
    kw["length"] = sc.length - line0
    g = orc.geo2rdr(lat=out["lat"], lon=out["lon"], hgt=out["hgt"], orbit_method=w["orbit_method"], want=("azoff", "rgoff"),
                    doppler_coeffs=tuple(c / sc.prf for c in sc.doppler_coeffs[0]), **kw)
    t2 = time.perf_counter()
    return dict(lines=lines, line0=line0, pixels=lines * sc.width, t_topo=t1 - t0, t_geo=t2 - t1, K=out["mean_iters"], N=g["mean_iters"])


def strip_start(sc, lines, k):
    """Strip k of the CPU arms: strips rotate over three places of the swath -- its first lines, its middle, its last
    lines -- and every revisit of a place takes the next `lines` lines there, so the strips of a run never repeat."""
    place, visit = k % 3, k // 3
    third = sc.length // 3
    off = (visit * lines) % max(1, third - lines)
    return min(max(0, place * third + off if place < 2 else sc.length - lines - off), max(0, sc.length - lines))


def cpu_baseline(w, sc, sec, budget_s=15.0):
    cal = cpu_sample(w, sc, sec, 4)
    per_line = (cal["t_topo"] + cal["t_geo"]) / cal["lines"]
    lines = int(max(4, min(256, budget_s / 3.0 / max(per_line, 1e-6))))
    parts = [cpu_sample(w, sc, sec, lines, strip_start(sc, lines, k)) for k in range(3)]
    t = sum(s["t_topo"] + s["t_geo"] for s in parts)
    px = sum(s["pixels"] for s in parts)
    return {"value": px / t / 1e6, "unit": "Mpixels/s", "cores": host_cores(), "kind": "port",
            "sample": f"3 strips of {lines} azimuth lines x {sc.width} samples (first lines, middle, last lines of the swath: lines "
                      f"{[s['line0'] for s in parts]}); topo {sum(s['t_topo'] for s in parts):.2f}s + geo2rdr "
                      f"{sum(s['t_geo'] for s in parts):.2f}s, OpenMP over pixels as in the reference",
            "note": "C restatement of the ISCE2 Fortran/C reference (oracle/), pinned bit for bit against the reference's own C / C++ "
                    "(tests/test_oracle_cpp_pins.py); gfortran is not available in this image",
            "K_topo_iters": float(np.mean([s["K"] for s in parts])), "N_geo_iters": float(np.mean([s["N"] for s in parts]))}


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
    for k in range(args.warmup):
        cpu_sample(w, sc, sec, lines, strip_start(sc, lines, k))
    t0 = time.perf_counter()
    px = 0
    starts = []
    for k in range(args.steps):
        s = cpu_sample(w, sc, sec, lines, strip_start(sc, lines, args.warmup + k))
        starts.append(s["line0"])
        px += s["pixels"]
    dt = time.perf_counter() - t0
    val = px / dt / 1e6
    covered = min(1.0, args.steps * lines / float(sc.length))
    sample = (f"each step = {lines} azimuth lines x {sc.width} samples (bounded sample of the workload, ~{args.ref_step_seconds:.0f} s of "
              f"CPU work); the strips rotate over the first lines, the middle and the last lines of the swath and never repeat: the "
              f"{args.steps} timed steps cover {100 * covered:.1f} % of its {sc.length} lines (the cost per pixel is position independent "
              f"to a few per cent: K iterations per pixel vary with the terrain)")
    also = reference_cpp_sample(w, sc, sec, lines)
    line = {"impl": "reference", "metric": "topo+geo2rdr Mpixels/s", "value": val, "unit": "Mpixels/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": w["desc"], "pixels_per_step_full": sc.pixels, "dem_method": w["dem_method"],
                       "orbit_method": w["orbit_method"], "strip_first_lines": starts},
            "cpu_baseline": {"value": val, "unit": "Mpixels/s", "cores": host_cores(), "kind": "port", "sample": sample,
                             "also_timed": also},
            "e2e": {"value": val, "unit": "Mpixels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def reference_cpp_sample(w, sc, sec, lines):
    """The reference's own C++ restatement of the path (components/zerodop/GPUtopozero + GPUgeo2rdr, CPU branches, compiled
    unchanged into oracle/_ref) on one strip of the same workload, next to the port: it is the checker of the port
    (bit-identical whole images, tests/test_oracle_cpp_pins.py), not the timed arm -- it is NOT the Fortran path the
    Components run (no layover bit, other DEM crop rule) and its data structures make it slower."""
    try:
        from oracle import ref_cpp
        if not ref_cpp.available():
            return {"unavailable": "oracle/_ref C++ libraries not built"}
        use_all_host_cores()
        n = max(4, min(lines, 64))
        line0 = max(0, sc.length // 2 - n // 2)
        from oracle import oracle as orc
        kw = orc.scene_topo_kwargs(sc, dem_method=w["dem_method"], orbit_method=w["orbit_method"])
        kw.update(length=n, t0=sc.t0 + line0 / sc.prf)  # the strip as its own short acquisition
        kw.pop("dem_method"), kw.pop("orbit_method")
        r = ref_cpp.topo(dem_method=w["dem_method"], orbit_method=w["orbit_method"], want_mask=w["mask"], **kw)
        gk = secondary_geo_kwargs(sc, sec)
        gk.update(t0=gk["t0"] + line0 / sc.prf, length=sc.length - line0)
        g = ref_cpp.geo2rdr(lat=r["lat"], lon=r["lon"], hgt=r["hgt"], orbit_method=w["orbit_method"],
                            doppler_coeffs=tuple(c / sc.prf for c in sc.doppler_coeffs[0]), **gk)
        t = r["seconds"] + g["seconds"]
        return {"what": "reference C++ restatement (GPUtopozero / GPUgeo2rdr CPU branches, unchanged, oracle/_ref)", "kind": "reference",
                "value": n * sc.width / t / 1e6, "unit": "Mpixels/s", "lines": n, "seconds_topo": r["seconds"], "seconds_geo2rdr": g["seconds"],
                "cores": host_cores()}
    except Exception as e:  # noqa: BLE001
        return {"unavailable": f"{type(e).__name__}: {e}"}


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


def host_mem_available():
    """Bytes of host memory this process may still take: MemAvailable, capped by the cgroup limit if there is one."""
    avail = None
    try:
        for ln in open("/proc/meminfo"):
            if ln.startswith("MemAvailable:"):
                avail = int(ln.split()[1]) * 1024
    except Exception:
        pass
    for f, cur in (("/sys/fs/cgroup/memory.max", "/sys/fs/cgroup/memory.current"),
                   ("/sys/fs/cgroup/memory/memory.limit_in_bytes", "/sys/fs/cgroup/memory/memory.usage_in_bytes")):
        try:
            lim = open(f).read().strip()
            if lim != "max" and int(lim) < (1 << 60):
                left = int(lim) - int(open(cur).read().strip())
                avail = left if avail is None else min(avail, left)
        except Exception:
            pass
    return avail


class Guard:
    """Local work of the secondary configs runs under a guard: a failure is recorded (and reported in the JSON line)
    instead of raised, and every rank still reaches every collective -- they are all issued outside guarded calls."""

    def __init__(self, fatal):
        self.fatal = fatal
        self.err = None

    def __call__(self, fn, *a, **kw):
        if self.err is not None:
            return None
        try:
            return fn(*a, **kw)
        except Exception as e:  # noqa: BLE001
            if self.fatal:
                raise
            import traceback
            self.err = f"{type(e).__name__}: {e}"
            log("[bench] secondary config failed (reported, not fatal):\n" + traceback.format_exc())
            return None


def load_json(path):
    try:
        return json.load(open(path))
    except Exception:
        return {}


def kernel_names(w):
    dm, om = w["dem_method"], w["orbit_method"]
    split = dm in ("BIQUINTIC", "BICUBIC", "SINC", "AKIMA")
    names = [f"k_topo_solve<{dm}>", f"k_topo_final<{dm}>"] if split else [f"k_topo_fused<{dm}>"]
    if w["mask"]:
        names.append(f"k_topo_mask<{dm}>")
    names.append(f"k_geo2rdr_poly<{om}>")
    return split, names


def measure_config(args, ranks, capi, dev, name, *, main, steps, warmup, e2e_steps, two_calls, want_e2e=True, lines=None):
    """One BASELINE config through the device-resident arm and the end-to-end arm(s).  Returns the dict that becomes the
    JSON line (main) or an entry of other_configs; complete on rank 0, None elsewhere."""
    G = Guard(fatal=main)
    st = {}  # local state shared by the guarded closures

    def setup():
        w, sc, sec = build_workload(name, lines)
        line0, nlines = shard(sc.length, ranks.rank, ranks.world)
        st.update(w=w, sc=sc, sec=sec, line0=line0, nlines=nlines, npix_local=nlines * sc.width, npix_total=sc.length * sc.width)
        st["tparams"] = capi.topo_params(dem_shape=sc.dem.shape, first_lat=sc.first_lat, first_lon=sc.first_lon,
                                         delta_lat=sc.delta_lat, delta_lon=sc.delta_lon, length=sc.length, width=sc.width,
                                         prf=sc.prf, t0=sc.t0, wvl=sc.wvl, side=sc.side, peg_heading=sc.peg_heading, a=sc.a,
                                         e2=sc.e2, dem_method=w["dem_method"], orbit_method=w["orbit_method"], line0=line0,
                                         nlines=nlines, device=dev)
        gk = secondary_geo_kwargs(sc, sec)
        st["gk"] = gk
        # native-Doppler sensors: geo2rdr takes the Doppler polynomial in cycles/PRF vs range pixel (runGeo2rdr.py:77-80)
        st["gdop"] = tuple(c / sc.prf for c in sc.doppler_coeffs[0])
        st["gparams"] = capi.geo_params(length=gk["length"], width=gk["width"], dem_shape=(sc.length, sc.width), r0=gk["r0"],
                                        dr=gk["dr"], prf=gk["prf"], t0=gk["t0"], wvl=gk["wvl"], side=gk["side"],
                                        orbit_method=w["orbit_method"], line0=line0, nlines=nlines, device=dev, out_f32=True)
        st["slr"] = [[sc.r0, sc.dr * sc.nrnglooks]]
        dem_host, _ = alloc_host(sc.dem.shape, np.float32, capi)
        dem_host[...] = sc.dem
        st["dem_host"] = dem_host
        log(f"[bench] {name}: rank {ranks.rank}/{ranks.world} device {dev}, lines [{line0},{line0 + nlines})")

    def device_setup():
        sc, w = st["sc"], st["w"]
        st["tplan"] = capi.TopoPlan(st["tparams"], st["dem_host"], sc.orbit_t, sc.orbit_pos, sc.orbit_vel, sc.doppler_coeffs,
                                    st["slr"], want_los=True, want_inc=w["inc"], want_mask=w["mask"])
        st["tplan"].execute()  # layers must exist before geo2rdr borrows them
        st["gplan"] = capi.GeoPlan(st["gparams"], topo_plan=st["tplan"])

    def device_step():
        gk = st["gk"]
        ms_t = st["tplan"].execute()
        ms_g = st["gplan"].execute(st["gparams"], gk["orbit_t"], gk["orbit_pos"], gk["orbit_vel"], doppler_coeffs=st["gdop"],
                                   want=("azoff", "rgoff"))
        return ms_t, ms_g

    def device_warm():
        for _ in range(warmup):
            device_step()

    def device_timed():
        sampler = ClockSampler(dev)
        t_lo = time.time()
        w0 = time.perf_counter()
        ms_topo = ms_geo = 0.0
        for _ in range(steps):
            a, b = device_step()
            ms_topo += a
            ms_geo += b
        st["wall_dev"] = time.perf_counter() - w0
        st["clocks"] = sampler.stop(t_lo, time.time())
        st["ms_topo"], st["ms_geo"] = ms_topo, ms_geo
        # per-kernel split and iteration statistics of the last step (the layers stay on the device)
        import ctypes as C
        res = capi.TopoResult()
        e = C.create_string_buffer(512)
        capi._check(capi.lib().b200_topo_plan_fetch(st["tplan"].handle, None, C.byref(res), e, 512), e)
        gres = capi.GeoResult()
        capi._check(capi.lib().b200_geo_plan_fetch(st["gplan"].handle, None, C.byref(gres), e, 512), e)
        st["res"], st["gres"] = res, gres
        st["ms_step_local"] = (ms_topo + ms_geo) / steps

    def device_close():
        # the resident layers (52 B/pixel + offsets) go back to the workspace cache before the end-to-end arm allocates
        for k in ("gplan", "tplan"):
            if st.get(k) is not None:
                st[k].close()
                st[k] = None

    G(setup)
    G(device_setup)
    G(device_warm)
    ranks.barrier()
    G(device_timed)
    ranks.barrier()
    ms_step = ranks.reduce_max(st.get("ms_step_local", 0.0))
    G(device_close)

    # ---------------- end-to-end arm: reference-facing C-ABI calls with host buffers ----------------
    def e2e_setup():
        sc, w, nlines = st["sc"], st["w"], st["nlines"]
        need = nlines * sc.width * (24 + 8 + (8 if w["inc"] else 0) + (1 if w["mask"] else 0) + 8)
        avail = host_mem_available()
        if avail is not None and need > 0.45 * avail:
            st["e2e_skip"] = (f"needs {need / 1e9:.1f} GB of page-locked host buffers on this rank, "
                              f"{avail / 1e9:.0f} GB of host memory available")
            return
        outs, pinned = {}, True
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
        st.update(outs=outs, gouts=gouts, pinned=pinned)
        st["d2h_local"] = sum(v.nbytes for v in outs.values() if v is not None) + sum(v.nbytes for v in gouts.values() if v is not None)

    def e2e_step_two_calls():
        # the reference's own call sequence: topo verb -> host lat / lon / hgt -> geo2rdr verb on the block's rows
        sc, w, gk, outs = st["sc"], st["w"], st["gk"], st["outs"]
        capi.topo_run(st["tparams"], st["dem_host"], sc.orbit_t, sc.orbit_pos, sc.orbit_vel, sc.doppler_coeffs, st["slr"],
                      want_los=True, want_inc=w["inc"], want_mask=w["mask"], out=outs)
        return capi.geo2rdr_run(st["gparams"], outs["lat"], outs["lon"], outs["hgt"], gk["orbit_t"], gk["orbit_pos"],
                                gk["orbit_vel"], doppler_coeffs=st["gdop"], want=("azoff", "rgoff"), out=st["gouts"], block_rows=True)

    def e2e_step_fused():
        # one call: same host inputs, same host outputs; geo2rdr runs on the layers while they are resident in HBM
        sc, w, gk = st["sc"], st["w"], st["gk"]
        job = dict(params=st["gparams"], orbit=(gk["orbit_t"], gk["orbit_pos"], gk["orbit_vel"]), doppler=(st["gdop"], 0.0, 1.0),
                   want=("azoff", "rgoff"), out=st["gouts"])
        _, geos = capi.topo_geo2rdr_run(st["tparams"], st["dem_host"], sc.orbit_t, sc.orbit_pos, sc.orbit_vel, sc.doppler_coeffs,
                                        [job], st["slr"], want_los=True, want_inc=w["inc"], want_mask=w["mask"], out=st["outs"])
        return geos[0]

    def time_e2e(step, tag, key):
        """Warm-up until the step time has settled on every rank (at least `warmup` calls; then until two consecutive calls
        are within 5 % of their predecessor on all ranks, at most 12 more: on the 8-GPU box the first ~8 calls of half the
        ranks run at half speed), barrier, e2e_steps timed calls; the reported time is the MEDIAN step, max over ranks."""
        live = lambda: not st.get("e2e_skip") and "outs" in st  # noqa: E731

        def one():
            if not live():
                return None
            w0 = time.perf_counter()
            step()
            return time.perf_counter() - w0

        prev, stable, nwarm = None, 0, 0
        for k in range(warmup + 12):
            t = G(one)
            nwarm += 1
            dev_ = abs(t / prev - 1.0) if (t and prev) else (1.0 if t else 0.0)
            dev_ = ranks.reduce_max(dev_)  # a collective per warm-up call, reached by every rank whatever happened locally
            prev = t
            stable = stable + 1 if dev_ < 0.05 else 0
            if k + 1 >= warmup and (stable >= 2 or ranks.world == 1):
                break

        # Multi-rank runs: on the 8-GPU box ranks 0-3 ran their first ~5-8 calls after EVERY barrier at half speed (260 ms
        # instead of 120 for the C2 swath), warm-up or not -- so each timed block is primed with 8 untimed calls issued back
        # to back with the timed ones (no barrier, no collective in between: all ranks keep the fabric loaded throughout)
        nprime = 8 if ranks.world > 1 else 0

        def timed():
            if not live():
                return
            times, r = [], None
            for _ in range(nprime):
                step()
            for _ in range(e2e_steps):
                w0 = time.perf_counter()
                r = step()
                times.append(time.perf_counter() - w0)
            log(f"[bench] {name}: rank {ranks.rank} e2e ({tag}) after {nwarm} warm-up + {nprime} priming calls, step times (ms): {[round(1e3 * t, 1) for t in times]}")
            st[key + "_try"] = dict(times=times, median=float(np.median(times)), r=r, nwarm=nwarm + nprime)

        # a block of e2e_steps calls whose slowest call is > 15 % over its fastest on some rank was disturbed (the boxes are
        # virtual machines: host-side copies occasionally run at ~80 % for a second or two): it is measured once more and
        # the steadier block counts, with the first one reported next to it
        medians = []
        for attempt in range(2):
            ranks.barrier()
            G(timed)
            d = st.pop(key + "_try", None) or {}
            wall_try = ranks.reduce_max(d.get("median", 0.0))
            spread = ranks.reduce_max((max(d["times"]) / min(d["times"])) if d.get("times") else 0.0)
            medians.append(wall_try)
            if d and (key not in st or wall_try <= min(medians[:-1] or [wall_try])):
                st[key] = d
            if spread <= 1.15 or e2e_steps < 3:
                break
        d = st.get(key) or {}
        if d:
            d["medians_of_all_blocks_ms"] = [1e3 * m for m in medians]
        wall = min(medians) if medians else 0.0
        drift = ranks.reduce_max((d["times"][-1] / d["times"][0]) if d.get("times") else 0.0)
        ranks.barrier()
        return wall, drift

    def d2h_floor_local():
        # the same bytes, device -> the same page-locked buffers, nothing else: the PCIe floor of the e2e call
        if st.get("e2e_skip") or "outs" not in st:
            return
        bufs = [v for v in list(st["outs"].values()) + list(st["gouts"].values()) if v is not None]
        ms = 0.0
        for _ in range(2):  # second pass is the measurement (first touches the pages)
            ms = sum(capi.d2h_floor(b, chunk_bytes=64 << 20, device=dev) for b in bufs)
        st["floor_ms"] = ms

    wall_two = wall_e2e = drift = 0.0
    check = None
    if want_e2e:
        G(e2e_setup)
        if two_calls:
            wall_two, _ = time_e2e(e2e_step_two_calls, "b200_topo_run + b200_geo2rdr_run", "two")

            def snapshot():
                # a strip of the block (<= 64 Mpixel): the fused call must leave the same offsets in the host buffers
                if "gouts" not in st:
                    return
                rows = max(1, min(st["nlines"], 64_000_000 // st["sc"].width))
                st["check"] = {k: v[:rows].copy() for k, v in st["gouts"].items() if v is not None}
            G(snapshot)
        wall_e2e, drift = time_e2e(e2e_step_fused, "b200_topo_geo2rdr_run", "fused")
        check = st.get("check")
    e2e_diff, e2e_valid_equal = 0.0, 1.0

    def compare_arms():
        # the fused call must have left the same offsets in the host buffers as the two calls (the floor measurement
        # below overwrites them)
        nonlocal e2e_diff, e2e_valid_equal
        if check is None or not st.get("fused") or not st.get("two"):
            return
        e2e_valid_equal = float(st["fused"]["r"]["num_valid"] == st["two"]["r"]["num_valid"])
        for k, v in check.items():
            cur = st["gouts"][k][:v.shape[0]]
            bad_a, bad_b = v == np.float32(-999999.0), cur == np.float32(-999999.0)
            e2e_valid_equal = min(e2e_valid_equal, float(np.array_equal(bad_a, bad_b)))
            both = ~bad_a & ~bad_b
            if both.any():
                e2e_diff = max(e2e_diff, float(np.abs(v[both].astype(np.float64) - cur[both].astype(np.float64)).max()))

    G(compare_arms)
    if want_e2e:
        ranks.barrier()
        G(d2h_floor_local)
        ranks.barrier()
    floor_ms = ranks.reduce_max(st.get("floor_ms", 0.0))
    e2e_diff = ranks.reduce_max(e2e_diff)
    e2e_valid_equal = -ranks.reduce_max(-e2e_valid_equal)
    res = st.get("res")
    small = (res.dem_nx * res.dem_ny * 4 + 2 * 7 * 8 * len(st["sc"].orbit_t)) if res is not None else 0
    h2d = ranks.reduce_sum(small)
    h2d_two = ranks.reduce_sum(small + 3 * 8 * st.get("npix_local", 0))
    d2h = ranks.reduce_sum(st.get("d2h_local", 0))
    valid = ranks.reduce_sum(st["fused"]["r"]["num_valid"] if st.get("fused") else (st["gres"].num_valid if st.get("gres") else 0))
    iters_t = ranks.reduce_sum(res.iterations if res is not None else 0)
    iters_g = ranks.reduce_sum(st["gres"].iterations if st.get("gres") else 0)
    err_any = ranks.reduce_max(1.0 if G.err else 0.0)
    for k in ("outs", "gouts", "check", "dem_host"):
        st.pop(k, None)
    capi.lib().b200_release_cached_memory() if hasattr(capi.lib(), "b200_release_cached_memory") else None
    if ranks.rank != 0:
        return None
    if G.err or err_any or res is None:
        return {"error": G.err or "a rank failed (see its stderr)", "workload": WORKLOADS[name]["desc"]}

    # ---------------- roofline of the dominant kernel + whole-step FP64 ----------------
    w, sc = st["w"], st["sc"]
    npix_local, npix_total, gres = st["npix_local"], st["npix_total"], st["gres"]
    K = iters_t / float(npix_total)
    Ngeo = iters_g / float(npix_total)
    value = npix_total / (ms_step * 1e-3) / 1e6
    dm = w["dem_method"]
    split, knames = kernel_names(w)
    w1_solve = W1_TOPO_ITER[dm] * K
    w1_final = W1_TOPO_FINAL[dm]
    w1_mask = W1_MASK[dm] if w["mask"] else 0.0
    w1_geo = W1_GEO_BASE + W1_GEO_ITER[w["orbit_method"]] * 9.0  # reference needs N = 9 steps
    w1_step = w1_solve + w1_final + w1_mask + w1_geo
    ms_solve = res.ms_solve if split else res.ms_pixels
    dom_name = knames[0]
    dom_w1 = w1_solve if split else (w1_solve + w1_final)
    achieved = dom_w1 * npix_local / (ms_solve * 1e-3) / 1e12
    peaks = load_json(os.path.join(ROOT, "MEASURED_PEAKS.json"))
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    bytes_px = 8 if split else (24 + 8 + (8 if w["inc"] else 0))  # solve kernel: the SCH height it hands over
    prof_all = load_json(os.path.join(ROOT, "profiles", "kernel_counters.json"))
    prof = dict(prof_all.get("kernels", {}))
    prof.update(prof_all.get("by_workload", {}).get(name, {}))  # counters of this workload's own capture where there is one
    pk = prof.get(dom_name, {})
    traffic = pk.get("dram_bytes_per_pixel")
    traffic = traffic * npix_local if traffic is not None else None

    def tf(w1, ms):
        return w1 * npix_local / (ms * 1e-3) / 1e12 if ms > 0 else None

    kms = {dom_name: ms_solve}
    kw1 = {dom_name: dom_w1}
    if split:
        kms[knames[1]] = res.ms_pixels - res.ms_solve
        kw1[knames[1]] = w1_final
    if w["mask"]:
        kms[f"k_topo_mask<{dm}>"] = res.ms_mask
        kw1[f"k_topo_mask<{dm}>"] = w1_mask
    kms[knames[-1]] = gres.ms_kernels
    kw1[knames[-1]] = w1_geo
    kernels, exec_flops, exec_known = {}, 0.0, True
    for kn, ms in kms.items():
        c = prof.get(kn, {})
        # executed FP64 (dadd + dmul + 2 dfma thread instructions, ncu) per pixel; the iterative kernels scale with K
        fpp = c.get("fp64_flop_per_pixel")
        if fpp is not None and c.get("per_iteration_K") and kn.startswith(("k_topo_solve", "k_topo_fused")):
            fpp = fpp * K / c["per_iteration_K"]
        if fpp is None:
            exec_known = False
        else:
            exec_flops += fpp * npix_local
        kernels[kn] = {"ms": ms, "w1_per_pixel": kw1[kn], "tflops_w1": tf(kw1[kn], ms),
                       "tflops_executed": (fpp * npix_local / (ms * 1e-3) / 1e12) if (fpp is not None and ms > 0) else None,
                       "frac_of_nominal_executed": (fpp * npix_local / (ms * 1e-3) / 1e12 / FP64_NOMINAL_TFLOPS)
                       if (fpp is not None and ms > 0) else None}
    kernels[knames[-1]]["note"] = (f"solves the reference's equation in N={Ngeo:.2f} true-Newton steps on orbit polynomials; "
                                   "W1 counts the reference's 9 quasi-Newton steps with full orbit re-interpolation")
    ms_kernels_local = sum(kms.values())
    fp64_peak = st["fp64_peak"] if "fp64_peak" in st else capi.fp64_peak(dev)
    roofline = {"kernel": dom_name, "bound": "fp64", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                "frac": achieved / fp64_peak, "traffic": traffic,
                "traffic_source": "profiles/kernel_counters.json (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per "
                                  "pixel of the profiled launch) x pixels of this launch" if traffic is not None else None,
                "peak_source": "measured live: b200_fp64_peak DFMA microbenchmark (MEASURED_PEAKS.json has no FP64 entry; its ncu "
                               "record is profiles/r02_ncu_fp64_peak.json)",
                "frac_of_nominal_37.2": achieved / FP64_NOMINAL_TFLOPS,
                "work_model": f"W1 (reference algorithm, unit-weight ops, SURVEY 8d): {dom_w1 / K:.0f}*K per pixel, K={K:.3f}"
                if split else f"W1: {W1_TOPO_ITER[dm]:.0f}*K + {W1_TOPO_FINAL[dm]:.0f} per pixel, K={K:.3f}",
                "avg_launch_ms": ms_solve, "pixels_per_launch": npix_local,
                "hbm": {"achieved_gbs": bytes_px * npix_local / (ms_solve * 1e-3) / 1e9, "peak_gbs": hbm_peak,
                        "frac": bytes_px * npix_local / (ms_solve * 1e-3) / 1e9 / hbm_peak, "bytes_per_pixel": bytes_px,
                        "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback (B200_PROFILING.md)"},
                "whole_step_executed_fp64": {
                    "tflops": exec_flops / (ms_kernels_local * 1e-3) / 1e12 if exec_known and ms_kernels_local > 0 else None,
                    "frac_of_nominal_37.2": exec_flops / (ms_kernels_local * 1e-3) / 1e12 / FP64_NOMINAL_TFLOPS
                    if exec_known and ms_kernels_local > 0 else None,
                    "note": "sum over the step's kernels of ncu's dadd + dmul + 2 dfma thread instructions per pixel "
                            "(profiles/kernel_counters.json) x this rank's pixels / the kernels' device time"},
                "kernels": kernels}
    line = {"metric": "topo+geo2rdr Mpixels/s", "value": value, "unit": "Mpixels/s", "n_gpus": ranks.world, "steps": steps,
            "warmup": warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": w["desc"], "pixels_per_step": npix_total, "dem_method": dm, "orbit_method": w["orbit_method"],
                       "sharding": f"{ranks.world} contiguous azimuth line blocks, no collective",
                       "l2": "inputs (DEM crop + previous layers) and outputs exceed the 126 MB L2 between iterations",
                       "K_topo_iters_per_pixel": K, "N_geo_steps_per_pixel": Ngeo,
                       "geo2rdr_valid_fraction": valid / float(npix_total), "dem_crop": [res.dem_ny, res.dem_nx]},
            "gpu_launches": int(steps * ((2 if split else 1) + (1 if w["mask"] else 0) + 2)),
            "clocks": st.get("clocks"), "roofline": roofline,
            "breakdown_ms": {"topo_solve": ms_solve, "topo_pixels": res.ms_pixels, "topo_mask": res.ms_mask, "geo2rdr": gres.ms_kernels,
                             "topo_step_avg": st["ms_topo"] / steps, "geo2rdr_step_avg": st["ms_geo"] / steps,
                             "wall_device_step": st["wall_dev"] / steps * 1e3},
            "work_equivalent_tflops": {"value": w1_step * npix_total / (ms_step * 1e-3) / 1e12,
                                       "note": "reference-algorithm W1 ops of the whole step / device time; the CUDA path removes "
                                               "reference work (hoisted setup, constant spline factors), so this can exceed executed FLOP/s"}}
    if want_e2e and st.get("e2e_skip"):
        line["e2e"] = {"value": None, "unit": "Mpixels/s", "skipped": st["e2e_skip"]}
    elif want_e2e:
        floor = {"ms": floor_ms, "value": npix_total / (floor_ms * 1e-3) / 1e6 if floor_ms > 0 else None,
                 "what": "the same output bytes copied device -> the same page-locked buffers by all ranks at once, no kernels"}
        line["e2e"] = {"value": npix_total / wall_e2e / 1e6, "unit": "Mpixels/s", "h2d_bytes_per_step": int(h2d),
                       "d2h_bytes_per_step": int(d2h), "ms_per_step": wall_e2e * 1e3, "steps": e2e_steps, "statistic": "median step, max over ranks",
                       "last_over_first_step": drift, "warmup_calls": (st.get("fused") or {}).get("nwarm"),
                       "medians_of_all_blocks_ms": (st.get("fused") or {}).get("medians_of_all_blocks_ms"),
                       "pinned_host_buffers": bool(st.get("pinned")),
                       "api": "b200_topo_geo2rdr_run (host DEM + orbits in; host lat/lon/hgt/los/inc/mask + range/azimuth "
                              "offsets out; geo2rdr consumes the layers in HBM)",
                       "d2h_floor": floor, "frac_of_d2h_floor": (floor_ms * 1e-3 / wall_e2e) if wall_e2e > 0 else None}
        if two_calls:
            line["e2e"]["vs_two_calls"] = {"max_abs_offset_diff_px": e2e_diff, "validity_equal": bool(e2e_valid_equal),
                                           "compared": check is not None,
                                           "rows_compared_per_rank": (next(iter(check.values())).shape[0] if check else 0)}
            line["e2e_two_calls"] = {"value": npix_total / wall_two / 1e6, "unit": "Mpixels/s", "h2d_bytes_per_step": int(h2d_two),
                                     "d2h_bytes_per_step": int(d2h), "ms_per_step": wall_two * 1e3, "steps": e2e_steps,
                                     "api": "b200_topo_run + b200_geo2rdr_run (the reference's call sequence: lat/lon/hgt go "
                                            "back up through the host)"}
    return line


def measure_c4(args, ranks, capi, dev, steps, warmup, lines=None, main=False):
    """BASELINE configs[4]: topsStack geo2rdr batch -- one reference geometry (the C2 swath's lat/lon/hgt, computed once
    per GPU and kept resident) against 29 perturbed secondary orbits; the 29 jobs are dealt round-robin to the ranks
    (contrib/stack/topsStack/Stack.py:805-827 launches one process per secondary date).  One step = all 29 jobs."""
    G = Guard(fatal=main)
    st = {}

    def setup():
        w = dict(WORKLOADS["c2"])
        if lines:
            w["length"] = int(lines)
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
        st.update(sc=sc, tplan=tplan, gplan=gplan, gp=gp, jobs=jobs, secs=secs)

    def step():
        ms = 0.0
        for j in st["jobs"]:
            o = st["secs"][j]
            ms += st["gplan"].execute(st["gp"], o.orbit_t, o.orbit_pos, o.orbit_vel, want=("azoff", "rgoff"))
        return ms

    def warm():
        for _ in range(warmup):
            step()

    def timed():
        sampler = ClockSampler(dev)
        t_lo = time.time()
        st["ms"] = sum(step() for _ in range(steps)) / steps
        st["clocks"] = sampler.stop(t_lo, time.time())

    def close():
        for k in ("gplan", "tplan"):
            if st.get(k) is not None:
                st[k].close()

    G(setup)
    G(warm)
    ranks.barrier()
    G(timed)
    ranks.barrier()
    ms = ranks.reduce_max(st.get("ms", 0.0))
    err_any = ranks.reduce_max(1.0 if G.err else 0.0)
    G(close)
    if ranks.rank != 0:
        return None
    if G.err or err_any:
        return {"error": G.err or "a rank failed (see its stderr)"}
    sc = st["sc"]
    npx = 29 * sc.pixels
    return {"metric": "geo2rdr Mpixels/s (29-orbit stack batch)", "value": npx / (ms * 1e-3) / 1e6, "unit": "Mpixels/s",
            "n_gpus": ranks.world, "steps": steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"topsStack batch: 29 secondary orbits x ({sc.length} x {sc.width}) geo2rdr on one resident "
                                   "reference geometry, jobs dealt round-robin to the GPUs",
                       "pixels_per_step": npx, "jobs_on_rank0": len(st["jobs"])},
            "gpu_launches": int(steps * 2 * len(st["jobs"])), "clocks": st.get("clocks")}


def measure_component(args, ranks, capi, dev, name, lines=None):
    """The call a user of the reference makes: createTopozero().topo() chained with createGeo2rdr().geo2rdr(), writing
    the .rdr / .off rasters with their .xml / .vrt into a directory (tmpfs when there is one) -- rank 0 only, one GPU,
    files included in the timed region (components/isceobj/TopsProc/runTopo.py:69-88, StripmapProc/runGeo2rdr.py:57-110)."""
    import shutil
    import tempfile
    if ranks.rank != 0:
        return None
    from isce2_b200 import synth_components as comp
    w, sc, sec = build_workload(name, lines)
    need = sc.pixels * 49
    base = "/dev/shm" if os.path.isdir("/dev/shm") else None
    try:
        free = shutil.disk_usage(base or tempfile.gettempdir()).free
    except Exception:
        free = None
    avail = host_mem_available()
    if (free is not None and need > 0.6 * free) or (base and avail is not None and need > 0.4 * avail):
        return {"value": None, "skipped": f"needs {need / 1e9:.1f} GB of output files, {0 if free is None else free / 1e9:.0f} GB free"}
    out = {"unit": "Mpixels/s", "api": "createTopozero().topo() with chainGeo2rdr(createGeo2rdr()): DEM array + orbit objects in; "
                                         "lat/lon/hgt/los/inc/mask .rdr and range/azimuth .off rasters + .xml/.vrt written",
           "directory": base or tempfile.gettempdir()}
    import contextlib
    times = []
    demdir = tempfile.mkdtemp(prefix="b200_bench_dem_", dir=base)
    try:
        dem_img = comp.prepare_dem(sc, os.path.join(demdir, "dem.dem"))  # the DEM file exists before the call, as in the apps
        for i in range(3):
            d = tempfile.mkdtemp(prefix="b200_bench_", dir=base)
            try:
                pw0 = capi.host_file_bytes() if hasattr(capi, "host_file_bytes") else 0
                t0 = time.perf_counter()
                with contextlib.redirect_stdout(sys.stderr):  # the Components print like the reference's; stdout carries the JSON line
                    info = comp.run_components(sc, sec, dem_img, d, dem_method=w["dem_method"], orbit_method=w["orbit_method"],
                                               inc=w["inc"], mask=w["mask"], devices=[dev])
                times.append(time.perf_counter() - t0)
                out["bytes_written"] = info.get("bytes_written")
                # of which by pwrite on the rasters' files (image.file_backed) rather than through their mappings
                out["bytes_written_with_pwrite"] = (capi.host_file_bytes() - pw0) if hasattr(capi, "host_file_bytes") else None
                gt = info.get("gpu_timings") or []
                out["library_call_ms"] = [round(float(g["ms_total"]), 1) for g in gt]  # inside b200_topo_geo2rdr_run
            finally:
                shutil.rmtree(d, ignore_errors=True)
        # the reference's own sequence: topo() writes its rasters, a separate geo2rdr() reads lat / lon / hgt back from them
        # (TopsProc/runTopo.py, then runGeo2rdr.py as a later step of the same application)
        two = []
        for i in range(3):
            d = tempfile.mkdtemp(prefix="b200_bench_", dir=base)
            try:
                with contextlib.redirect_stdout(sys.stderr):
                    info = comp.run_components_separately(sc, sec, dem_img, d, dem_method=w["dem_method"], orbit_method=w["orbit_method"],
                                                          inc=w["inc"], mask=w["mask"], devices=[dev])
                two.append((info["seconds_topo"], info["seconds_geo2rdr"]))
            finally:
                shutil.rmtree(d, ignore_errors=True)
        tt, tg = float(np.median([a for a, _ in two[1:]])), float(np.median([b for _, b in two[1:]]))
        out["two_calls"] = {"api": "createTopozero().topo(), then createGeo2rdr().geo2rdr(latImage, lonImage, demImage) on the rasters it wrote",
                            "seconds_topo": round(tt, 3), "seconds_geo2rdr": round(tg, 3), "value": sc.pixels / (tt + tg) / 1e6,
                            "unit": "Mpixels/s", "statistic": "medians of the steps after the first"}
        log(f"[bench] {name}: component path, two calls (s): {[(round(a, 3), round(b, 3)) for a, b in two]}")
    except Exception as e:  # noqa: BLE001
        import traceback
        log(traceback.format_exc())
        return {"error": f"{type(e).__name__}: {e}"}
    finally:
        shutil.rmtree(demdir, ignore_errors=True)
    log(f"[bench] {name}: component path step times (s): {[round(t, 3) for t in times]}")
    # the floor of anything that writes these files: the same bytes stored into fresh memory maps of new files in the same
    # directory by the same number of host threads, no GPU involved (page allocation + zeroing + copy of the page cache)
    try:
        out["file_floor"] = file_write_floor(base, int(out.get("bytes_written") or need))
        out["file_floor_pwrite"] = file_write_floor(base, int(out.get("bytes_written") or need), how="pwrite")
    except Exception as e:  # noqa: BLE001
        out["file_floor"] = {"error": f"{type(e).__name__}: {e}"}
    t = float(np.median(times[1:])) if len(times) > 1 else times[0]
    out.update(value=sc.pixels / t / 1e6, ms_per_step=t * 1e3, steps=len(times) - 1, statistic="median of the steps after the first")
    return out


def file_write_floor(base, nbytes, nfiles=8, threads=None, how="mmap"):
    """Seconds to store nbytes into nfiles fresh files under `base` from host memory with `threads` threads -- through
    numpy.memmap(mode='w+') stores (numpy releases the GIL in the copies; munmap included) or through os.pwrite -- what
    any writer of new rasters pays."""
    import shutil
    import tempfile
    import threading as th
    threads = threads or max(2, min(16, len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 2)))  # = the library's copier pool
    per = (nbytes // nfiles) & ~4095
    src = np.ones(64 << 20, np.uint8)
    d = tempfile.mkdtemp(prefix="b200_floor_", dir=base)
    try:
        t0 = time.perf_counter()
        if how == "mmap":
            maps = [np.memmap(os.path.join(d, f"f{i}.bin"), dtype=np.uint8, mode="w+", shape=(per,)) for i in range(nfiles)]
        else:
            maps = [os.open(os.path.join(d, f"f{i}.bin"), os.O_CREAT | os.O_WRONLY, 0o644) for i in range(nfiles)]
        jobs = [(m, o) for m in maps for o in range(0, per, src.size)]

        def work(k):
            for m, o in jobs[k::threads]:
                n = min(src.size, per - o)
                if how == "mmap":
                    m[o:o + n] = src[:n]
                else:
                    os.pwrite(m, memoryview(src)[:n], o)

        ts = [th.Thread(target=work, args=(k,)) for k in range(threads)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        if how != "mmap":
            for fd in maps:
                os.close(fd)
        del maps, jobs
        dt = time.perf_counter() - t0
    finally:
        shutil.rmtree(d, ignore_errors=True)
    return {"seconds": dt, "GBps": per * nfiles / dt / 1e9, "threads": threads, "bytes": per * nfiles,
            "what": ("numpy stores into fresh memory maps of" if how == "mmap" else "os.pwrite into") +
                    " new files in the same directory, no GPU involved"}


def run_b200(args, ranks):
    t_start = time.time()
    from isce2_b200 import _capi as capi
    if capi.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device visible; this arm has no CPU fallback (use --impl reference for the CPU baseline)")
    dev = ranks.local_rank % capi.device_count()
    if ranks.world > 1:
        bound = bind_to_gpu_local_cpus(dev)
        if bound:
            log(f"[bench] rank {ranks.rank}: bound to the {len(bound)} CPUs local to GPU {dev}")
    log(f"[bench] rank {ranks.rank}/{ranks.world} device {dev} ({capi.device_name(dev)}), FP64 peak (DFMA microbenchmark) "
        f"{capi.fp64_peak(dev):.1f} TFLOP/s")
    e2e_steps = max(1, args.e2e_steps)
    if args.workload == "c4":
        line = measure_c4(args, ranks, capi, dev, args.steps, args.warmup, lines=args.lines, main=True)
        if ranks.rank == 0:
            print(json.dumps(line), flush=True)
        return line
    line = measure_config(args, ranks, capi, dev, args.workload, main=True, steps=args.steps, warmup=args.warmup, e2e_steps=e2e_steps,
                          two_calls=True, lines=args.lines)
    others = {}
    if args.other_configs != "none":
        osteps, owarm = max(1, min(args.steps, 5)), max(1, min(args.warmup, 3))
        for oname in [n for n in ("c0c1", "c3", "c4") if n != args.workload and (args.other_configs == "all" or n in args.other_configs.split(","))]:
            t0 = time.time()
            if oname == "c4":
                r = measure_c4(args, ranks, capi, dev, osteps, owarm, lines=args.other_lines)
            else:
                r = measure_config(args, ranks, capi, dev, oname, main=False, steps=osteps, warmup=owarm, e2e_steps=min(e2e_steps, 5),
                                   two_calls=False, lines=args.other_lines)
            if ranks.rank == 0:
                if r is not None and "error" not in r:  # the sub-lines keep what distinguishes them from the main line
                    r = {k: r[k] for k in ("metric", "value", "unit", "ms_per_step", "steps", "warmup", "config", "e2e", "roofline",
                                           "breakdown_ms", "gpu_launches", "clocks") if k in r}
                    r["bench_seconds"] = round(time.time() - t0, 1)
                others[oname] = r
    comp = None
    if args.component and ranks.world == 1:
        comp = measure_component(args, ranks, capi, dev, args.workload, lines=args.lines)
    ranks.barrier()
    if ranks.rank == 0:
        st_cb = None
        if ranks.world == 1 and not args.no_cpu_baseline:
            w, sc, sec = build_workload(args.workload, args.lines)
            st_cb = cpu_baseline(w, sc, sec)
        line["cpu_baseline"] = st_cb
        line["other_configs"] = others
        if comp is not None:
            line["e2e_component"] = comp
            if comp.get("value") and line.get("e2e", {}).get("value"):
                comp["frac_of_c_abi_e2e"] = comp["value"] / line["e2e"]["value"]
            for key, fl in (("frac_of_file_floor", "file_floor"), ("frac_of_file_floor_pwrite", "file_floor_pwrite")):
                ff = comp.get(fl) or {}
                if comp.get("ms_per_step") and ff.get("seconds"):
                    comp[key] = ff["seconds"] * 1e3 / comp["ms_per_step"]  # > 1: the path beats that way of writing the files
        line["bench_seconds"] = round(time.time() - t_start, 1)  # whole run of this arm, all configurations and baselines
        print(json.dumps(line), flush=True)
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS) + ["c4"])
    ap.add_argument("--lines", type=int, default=None, help="override the number of azimuth lines (debugging)")
    ap.add_argument("--e2e-steps", type=int, default=10, help="timed end-to-end calls per arm (the median is reported)")
    ap.add_argument("--other-configs", default="all",
                    help="BASELINE configs measured next to --workload and embedded as other_configs: all | none | c0c1,c3,c4")
    ap.add_argument("--other-lines", type=int, default=None, help="override the azimuth lines of the other configs (debugging)")
    ap.add_argument("--component", type=int, default=1, help="1: also time the Component path (createTopozero().topo() -> files)")
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
        if args.impl == "reference":
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
