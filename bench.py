#!/usr/bin/env python3
"""bench.py -- the headline benchmark of BASELINE.json on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c2|c3haar|c3db8|c4|c5|...]

A "step" is one pass of the hot path over one batch of synthetic input: forward + inverse of the named
transform.  Default workload = BASELINE.json configs[1] ("c2"): batched MODWT Daubechies4 J=6 forward+inverse on
4,096 signals x 65,536 fp64 samples per GPU (weak scaling: every rank owns its own 4,096 signals, sharded by signal,
no collective on the data path).  `value` = sample-transforms per second summed over all ranks, counting
batch*N samples for the forward and batch*N for the inverse of every step, inputs resident in HBM.

Prints ONE JSON line (see the task contract): metric/value/unit, roofline (the LONGER of the two kernels of a step:
algorithmic bytes 8*(J+2) per sample / CUDA-event duration vs MEASURED_PEAKS.json), cpu_baseline (the oracle's
restatement of the reference's default FFT-convolution MODWT on the host cores, bounded sample), e2e (same metric
through the host-buffer C ABI with pinned host memory, copies inside the timed region), clocks, gpu_launches, and --
on the default workload --
  per_config   the other BASELINE configs (c3haar, c3db8, c4, c5), 3 warm-up + >= 10 timed steps each: forward /
               inverse ms, fraction of the measured HBM peak and of the fp64 FMA rate measured in this run,
               round-trip error, SM clock during that config (N = 1 only)
  c5_strong    BASELINE configs[4] as north_star states it: 8,192 series x 65,536, Daubechies20 J=8, STRONG-sharded by
               series over the N ranks (8192/N series per rank, time = max over ranks)
  multi_device (N > 1; rank 0 alone, the other ranks wait at the barrier) ONE context over all N devices -- the path a
               single JVM would use: split-series transforms bit-identical to the unsplit ones, and the host-buffer
               e2e through that one context
  link_probe   raw pinned-host <-> device copy rates of all ranks at once (what bounds e2e at N > 1)

--impl reference times the reference's own CPU algorithm (oracle port: FFT-convolution MODWT exactly as
MODWTTransform.java:752-837 + FastFourierTransform.java:172-212; no JVM exists here) on all host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    #  name: (kind, wavelet class, levels, batch per GPU, n)
    "c2": ("modwt", "Daubechies4", 6, 4096, 65536),
    "c5": ("modwt", "Daubechies20", 8, 8192, 65536),     # BASELINE configs[4]; `c5_strong` shards these 8,192 series
    "c3haar": ("fwt", "Haar1", 20, 1024, 1 << 20),
    "c3db8": ("fwt", "Daubechies8", 20, 1024, 1 << 20),
    "c4": ("wpt", "Symlet8", 6, 512, 65536),
    # SURVEY section 8f rows (not BASELINE configs): 2-D FWT of 32 matrices 4096 x 4096 at full depth (levels = lvlM = lvlN,
    # n = rows = cols), and the reference's sliding-window shape (512-sample windows, step 64, J = 8) over a 2^24 series
    "fwt2d": ("fwt2d", "Daubechies4", 12, 32, 4096),
    "windows": ("windows", "Daubechies4", 8, ((1 << 24) - 512) // 64 + 1, 512),
    # arbitrary length is the reference MODWT's contract (MODWTInverseTest.java:20-92): a non-2^p length, same filter bank
    "modwt_n100k": ("modwt", "Daubechies4", 6, 2048, 100000),
}
PER_CONFIG = ["c3haar", "c3db8", "c4", "c5"]
WINDOW_HOP = 64
C5_TOTAL_SERIES = 8192


def workload_desc(name, kind, cls, levels, batch, n):
    if kind == "fwt2d":
        return "%s: 2-D FWT %s lvlM=lvlN=%d forward+reverse, %d matrices of %d x %d fp64 per GPU" % (
            name, cls, levels, batch, n, n)
    if kind == "windows":
        return ("%s: sliding-window MODWT %s J=%d, %d windows of %d samples (step %d) of one series per GPU, forward "
                "(windows read in place) + inverse" % (name, cls, levels, batch, n, WINDOW_HOP))
    return "%s: batched %s %s J=%d forward+inverse, %d signals x %d fp64 samples per GPU" % (
        name, kind.upper(), cls, levels, batch, n)


def algorithmic_bytes_per_sample(kind, levels, n=0):
    # MODWT fwd or inv 8*(J+2) B/sample; FWT / WPT 16 B/sample; 2-D = row pass + column pass;
    # windows: every series sample is read once (8*hop per window), J+1 rows written per window
    if kind == "fwt2d":
        return 32
    if kind == "windows":
        return 8.0 * (levels + 1) + 8.0 * WINDOW_HOP / n
    return 8 * (levels + 2) if kind == "modwt" else 16


def flops_per_sample(kind, levels, L):
    """fp64 flop per sample per direction (2 per FMA): MODWT 2 filters x L taps x J levels; WPT L FMA per sample per
    level; FWT the same on a halving prefix."""
    if kind in ("modwt", "windows"):
        return 4.0 * L * levels
    if kind == "wpt":
        return 2.0 * L * levels
    if kind == "fwt":
        return 4.0 * L * (1.0 - 0.5 ** levels)
    if kind == "fwt2d":
        return 2.0 * 4.0 * L * (1.0 - 0.5 ** levels)
    return 0.0


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:  # noqa: BLE001
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi sampling of SM clock / throttle reasons for the whole run; window(t0, t1) summarises the samples whose
    timestamp lies inside one timed region."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.rows = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "20"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:  # noqa: BLE001
            self.p = None

    def _read(self):
        import datetime
        self.f.flush()
        self.f.seek(0)
        rows = []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                ts = datetime.datetime.strptime(parts[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, float(parts[1]), float(parts[2]),
                             [nm for nm, v in zip(names, parts[5:9]) if v.lower().startswith("active")]))
            except ValueError:
                continue
        return rows

    def window(self, t0, t1):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        if self.rows is None:
            time.sleep(0.06)   # let the sample that covers the end of the region land in the file
        rows = self.rows if self.rows is not None else self._read()
        inside = [r for r in rows if t0 - 0.02 <= r[0] <= t1 + 0.02]
        use = inside if inside else rows[-3:]
        if use:
            reasons = set()
            for r in use:
                reasons.update(r[3])
            out.update(sm_mhz=float(np.median([r[1] for r in use])), sm_max_mhz=float(max(r[2] for r in use)),
                       sm_mhz_min=float(min(r[1] for r in use)), reasons=sorted(reasons), samples=len(use),
                       samples_inside_timed_region=len(inside))
        return out

    def stop(self):
        if self.p is None:
            return
        time.sleep(0.05)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.p.kill()
        self.rows = self._read()
        self.p = None
        try:
            os.unlink(self.f.name)
        except OSError:
            pass


def cpu_reference_time(kind, cls, levels, n, nsig, threads, repeats=1, schedule="per_signal"):
    """Time the oracle's restatement of the reference's CPU path on `nsig` signals with `threads` host threads.
    MODWT uses the reference's default FFT convolution (AUTO picks FFT for every BASELINE config).
    schedule = "parallel_wpt": the WPT through the reference's ParallelWaveletPacketTransform schedule (one signal at
    a time, every level's packets forked over the threads) instead of one signal per thread."""
    import jwave_pro_b200 as jw
    from jwave_pro_b200.synth import splitmix_uniform
    from oracle import c_oracle as oracle
    w = jw.wavelets.create(cls)
    s, wv = w.getScalingDeComposition(), w.getWaveletDeComposition()
    X = splitmix_uniform(0x5EED0002, (nsig, n))
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        if kind in ("modwt", "windows"):   # windows: the reference copies each window out and transforms it
            g, h = oracle.modwt_filters(s, wv)
            c = oracle.batch("modwt_fwd_fft", X, levels, g, h, nthreads=threads)
            oracle.batch("modwt_inv_fft", c, levels, g, h, nthreads=threads)
        elif kind == "fwt2d":
            m = int(round(n ** 0.5))
            c = oracle.batch2d("fwt", X.reshape(nsig, m, m), levels, levels, s, wv, nthreads=threads)
            oracle.batch2d("fwt", c, levels, levels, w.getScalingReConstruction(), w.getWaveletReConstruction(),
                           reverse=True, nthreads=threads)
        elif kind == "wpt" and schedule == "parallel_wpt":
            c = oracle.parallel_wpt(X, levels, s, wv, nthreads=threads)
            oracle.parallel_wpt(c, levels, w.getScalingReConstruction(), w.getWaveletReConstruction(), reverse=True,
                                nthreads=threads)
        else:
            c = oracle.batch(kind + "_fwd", X, levels, s, wv, nthreads=threads)
            oracle.batch(kind + "_rev", c, levels, w.getScalingReConstruction(), w.getWaveletReConstruction(),
                         nthreads=threads)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return best


def cpu_baseline_for(kind, cls, levels, unit, batch, cores, budget_s, schedule="per_signal"):
    """Bounded-sample CPU baseline: about `budget_s` seconds of host work."""
    probe_n = max(1, min(batch, cores if schedule == "per_signal" else 2))
    probe = cpu_reference_time(kind, cls, levels, unit, probe_n, cores, schedule=schedule)
    per_sig = probe / probe_n if schedule != "per_signal" else probe / max(1, -(-probe_n // cores))
    if schedule == "per_signal":
        nsig = int(max(cores, min(batch, cores * max(1, round(budget_s / max(per_sig, 1e-4))))))
    else:
        nsig = int(max(2, min(batch, round(budget_s / max(per_sig, 1e-4)))))
    dt = cpu_reference_time(kind, cls, levels, unit, nsig, cores, schedule=schedule)
    what = {"modwt": "FFT-convolution MODWT", "windows": "FFT-convolution MODWT"}.get(kind, kind.upper())
    if schedule == "parallel_wpt":
        how = ("ParallelWaveletPacketTransform schedule (ParallelWaveletPacketTransform.java:79-110,155-158,197-233: one "
               "signal at a time, each level's packets forked over %d threads when packet >= 64 and packets >= 8)" % cores)
    else:
        how = "%d threads (one signal per thread)" % cores
    return {"value": 2.0 * nsig * unit / dt / 1e9, "unit": "Gsamples/s", "cores": cores, "kind": "port",
            "sample": "%d of %d signals x %d samples, forward+inverse, %s; C restatement of the reference's %s" % (
                nsig, batch, unit, how, what)}


def run_reference(args, kind, cls, levels, batch, n, rank, world):
    """--impl reference: rank 0 alone works; other ranks exit 0."""
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    unit = n * n if kind == "fwt2d" else n   # samples per unit of work (signal, window or matrix)
    # bounded sample per step: a few signals per core, so K+W steps finish in minutes
    probe = cpu_reference_time(kind, cls, levels, unit, cores, cores)
    # seconds of host work per step, sized so that the whole --steps K --warmup W run ends within ~2.5 minutes
    target = min(6.0, max(0.5, 150.0 / max(1, args.steps + args.warmup)))
    nsig = int(max(cores, min(batch, cores * max(1, round(target / max(probe, 1e-3))))))
    times = []
    for i in range(args.warmup + args.steps):
        dt = cpu_reference_time(kind, cls, levels, unit, nsig, cores)
        if i >= args.warmup:
            times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    value = 2.0 * nsig * unit / (ms * 1e-3) / 1e9
    sample = "%d signals x %d samples per step (of %d), forward+inverse, %d threads, one signal per thread" % (
        nsig, unit, batch, cores)
    line = {
        "impl": "reference", "metric": "MODWT/FWT/WPT Gsamples/s (forward+inverse sample-transforms per second)",
        "value": value, "unit": "Gsamples/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": workload_desc(args.workload, kind, cls, levels, batch, n),
                   "note": "CPU arm: C restatement (oracle/jwave_oracle.c, -O2 -ffp-contract=off) of the reference's "
                           "default path (FFT-convolution MODWT; no JVM in this image), bounded sample per step"},
        "cpu_baseline": {"value": value, "unit": "Gsamples/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "Gsamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


class DeviceRun:
    """Device-resident measurement of one workload on this process's device slots: synthetic inputs in HBM, forward +
    inverse per step on an explicit stream per slot, CUDA events on that stream."""

    def __init__(self, jw, torch, ctx, devices, rank, name, batch_override=0, flags=0):
        import ctypes
        self.jw, self.torch, self.ctx, self.devices, self.flags = jw, torch, ctx, devices, flags
        self.name = name
        kind, cls, levels, batch, n = WORKLOADS[name]
        if batch_override > 0:
            batch = batch_override
        self.kind, self.cls, self.levels, self.batch, self.n = kind, cls, levels, batch, n
        self.unit = n * n if kind == "fwt2d" else n
        w = jw.wavelets.create(cls)
        self.wavelet = w
        self.L = len(w.getScalingDeComposition())
        if kind in ("modwt", "windows"):
            self.tr = jw.CudaMODWTTransform(w, context=ctx)
        elif kind in ("fwt", "fwt2d"):
            self.tr = jw.CudaFastWaveletTransform(w, context=ctx)
        else:
            self.tr = jw.CudaWaveletPacketTransform(w, context=ctx)
        self.out_rows = levels + 1 if kind in ("modwt", "windows") else 1
        self.series_len = (batch - 1) * WINDOW_HOP + n   # windows workload: one series per GPU, `batch` windows
        self.bufs = []
        for slot, d in enumerate(devices):
            with torch.cuda.device(d):
                gen = torch.Generator(device="cuda:%d" % d)
                gen.manual_seed(0x5EED0002 + rank * 16 + slot)
                xshape = (self.series_len,) if kind == "windows" else (batch, self.unit)
                x = torch.rand(xshape, dtype=torch.float64, device="cuda:%d" % d, generator=gen) * 2.0 - 1.0
                c = torch.empty((batch, self.out_rows * self.unit), dtype=torch.float64, device="cuda:%d" % d)
                xr = torch.empty((batch, self.unit), dtype=torch.float64, device="cuda:%d" % d)
                self.bufs.append((x, c, xr, torch.cuda.Stream(device=d)))   # explicit stream: kernels AND events
        if kind == "windows":
            self._lib = jw._native.load()
            self._g, self._h = (np.ascontiguousarray(v) for v in self.tr._filters())
            self._dp = ctypes.POINTER(ctypes.c_double)
            self._vp = ctypes.c_void_p

    def fwd(self, slot):
        x, c, xr, st = self.bufs[slot]
        tr, kind, batch, n, levels = self.tr, self.kind, self.batch, self.n, self.levels
        if kind == "windows":
            rc = self._lib.jwc_modwt_forward_windows_dev(
                self.ctx.handle, slot, self._vp(st.cuda_stream), self._vp(x.data_ptr()), self._vp(c.data_ptr()),
                self.series_len, n, WINDOW_HOP, levels, self._g.ctypes.data_as(self._dp),
                self._h.ctypes.data_as(self._dp), len(self._g), self.flags)
            assert rc == 0, self._lib.jwc_last_error()
        elif kind == "fwt2d":
            tr.forward2DDevice(x.data_ptr(), c.data_ptr(), batch, n, n, levels, levels, stream=st.cuda_stream,
                               flags=self.flags, slot=slot)
        elif kind == "modwt":
            tr.forwardMODWTDevice(x.data_ptr(), c.data_ptr(), batch, n, levels, stream=st.cuda_stream, flags=self.flags,
                                  slot=slot)
        else:
            tr.forwardDevice(x.data_ptr(), c.data_ptr(), batch, n, levels, stream=st.cuda_stream, flags=self.flags,
                             slot=slot)

    def inv(self, slot):
        x, c, xr, st = self.bufs[slot]
        tr, kind, batch, n, levels = self.tr, self.kind, self.batch, self.n, self.levels
        if kind == "fwt2d":
            tr.reverse2DDevice(c.data_ptr(), xr.data_ptr(), batch, n, n, levels, levels, stream=st.cuda_stream,
                               flags=self.flags, slot=slot)
        elif kind in ("modwt", "windows"):
            tr.inverseMODWTDevice(c.data_ptr(), xr.data_ptr(), batch, n, levels, stream=st.cuda_stream,
                                  flags=self.flags, slot=slot)
        else:
            tr.reverseDevice(c.data_ptr(), xr.data_ptr(), batch, n, levels, stream=st.cuda_stream, flags=self.flags,
                             slot=slot)

    def round_trip_error(self):
        if self.kind == "windows":   # every reconstructed window against its span of the series
            return max(float((b[2] - b[0].unfold(0, self.n, WINDOW_HOP)).abs().max()) for b in self.bufs)
        return max(float((b[2] - b[0]).abs().max()) for b in self.bufs)

    def measure(self, steps, warmup, sync_all, reduce_max):
        torch, devices, bufs = self.torch, self.devices, self.bufs
        for _ in range(warmup):
            for s in range(len(devices)):
                self.fwd(s)
                self.inv(s)
        sync_all()
        pr = self.round_trip_error() if warmup > 0 else 0.0
        ev = []
        launches0 = self.ctx.launch_count()
        sync_all()
        t0 = time.time()
        for s, d in enumerate(devices):
            with torch.cuda.device(d):
                e = [torch.cuda.Event(enable_timing=True) for _ in range(2 * steps + 1)]
                e[0].record(bufs[s][3])
                ev.append(e)
        for k in range(steps):
            for s, d in enumerate(devices):
                with torch.cuda.device(d):
                    self.fwd(s)
                    ev[s][2 * k + 1].record(bufs[s][3])
                    self.inv(s)
                    ev[s][2 * k + 2].record(bufs[s][3])
        sync_all()
        t1 = time.time()
        launches = self.ctx.launch_count() - launches0
        total_ms = max(e[0].elapsed_time(e[-1]) for e in ev)
        fwd_ms = max(sum(e[2 * k].elapsed_time(e[2 * k + 1]) for k in range(steps)) for e in ev) / steps
        inv_ms = max(sum(e[2 * k + 1].elapsed_time(e[2 * k + 2]) for k in range(steps)) for e in ev) / steps
        # the same events, step by step: the first and the last (up to) ten steps of the timed region on their own -- a long
        # region runs into the board's power cap (sw_power_cap) and its steps get slower as the SM clock comes down
        w = min(10, steps)
        head_ms = max(e[0].elapsed_time(e[2 * w]) for e in ev) / w
        tail_ms = max(e[2 * (steps - w)].elapsed_time(e[2 * steps]) for e in ev) / w
        total_ms, fwd_ms, inv_ms, head_ms, tail_ms = reduce_max([total_ms, fwd_ms, inv_ms, head_ms, tail_ms])   # device time = max over ranks
        return {"ms_per_step": total_ms / steps, "fwd_ms": fwd_ms, "inv_ms": inv_ms, "round_trip_max_err": pr,
                "launches": int(launches), "t0": t0, "t1": t1, "steps": steps, "warmup": warmup,
                "ms_first_steps": head_ms, "ms_last_steps": tail_ms, "window_steps": w}

    def fractions(self, res, peak_gbs, fp64_tflops):
        bps = algorithmic_bytes_per_sample(self.kind, self.levels, self.n)
        bps_inv = 8.0 * (self.levels + 2) if self.kind == "windows" else bps   # the inverse reads J+1 rows, writes one
        fl = flops_per_sample(self.kind, self.levels, self.L)
        work = self.batch * self.unit
        out = {}
        for d, ms, b in (("fwd", res["fwd_ms"], bps), ("inv", res["inv_ms"], bps_inv)):
            gbs = b * work / (ms * 1e-3) / 1e9
            tf = fl * work / (ms * 1e-3) / 1e12
            out[d] = {"ms": ms, "gbs": gbs, "hbm_frac": gbs / peak_gbs, "tflops": tf,
                      "fp64_frac": (tf / fp64_tflops) if fp64_tflops else None,
                      "algorithmic_bytes": b * work}
        return out

    def free(self):
        self.bufs = []
        self.ctx.synchronize()
        self.ctx.release_scratch()
        self.torch.cuda.empty_cache()


def link_probe(torch, device, reduce_sum_fn, barrier):
    """Raw pinned-host <-> device copy rates with every rank copying at once: H2D alone, D2H alone, both together.
    Returns per-rank and summed GB/s (this is the ceiling of the host-buffer e2e numbers)."""
    nbytes = 256 << 20
    h_in = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d_a = torch.empty(nbytes, dtype=torch.uint8, device=device)
    d_b = torch.empty(nbytes, dtype=torch.uint8, device=device)
    s1, s2 = torch.cuda.Stream(device=device), torch.cuda.Stream(device=device)

    def run(h2d, d2h, reps=4):
        torch.cuda.synchronize(device)
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            if h2d:
                with torch.cuda.stream(s1):
                    d_a.copy_(h_in, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    h_out.copy_(d_b, non_blocking=True)
        torch.cuda.synchronize(device)
        return nbytes * reps / (time.perf_counter() - t0) / 1e9

    run(True, True, 1)
    h, d, b = run(True, False), run(False, True), run(True, True)
    sh, sd, sb = reduce_sum_fn([h, d, b])
    return {"h2d_gbs_rank0": h, "d2h_gbs_rank0": d, "duplex_gbs_per_direction_rank0": b,
            "h2d_gbs_all_ranks": sh, "d2h_gbs_all_ranks": sd, "duplex_gbs_per_direction_all_ranks": sb,
            "bytes_per_copy": nbytes,
            "note": "cudaMemcpyAsync between pinned host memory and the device, all ranks at the same time"}


def e2e_measure(jw, torch, run, devices, esteps, psteps, batch_per_gpu, barrier, reduce_max, world, with_stream=True):
    """The same metric through the host-buffer C ABI (pinned host memory, copies inside the timed region)."""
    kind, levels, n, unit, out_rows, w = run.kind, run.levels, run.n, run.unit, run.out_rows, run.wavelet
    nd = len(devices)
    ectx = jw.Context(devices)
    if kind == "windows":
        # one series on the host; a bounded number of windows so that pinned staging stays ~1 GB
        eb = min(run.batch, 32768)
        elen = (eb - 1) * WINDOW_HOP + n
        hx = torch.empty(elen, dtype=torch.float64).pin_memory()
        hx.copy_(run.bufs[0][0][:elen].cpu())
        hc = torch.empty((eb, out_rows * unit), dtype=torch.float64).pin_memory()
        hr = torch.empty((eb, unit), dtype=torch.float64).pin_memory()
        et = jw.CudaMODWTTransform(w, context=ectx)
        X, C, R = hx.numpy(), hc.numpy().reshape(eb, out_rows, n), hr.numpy()
        Xref = np.lib.stride_tricks.sliding_window_view(X, n)[::WINDOW_HOP][:eb]
        e_fwd = lambda Cb: et.forwardMODWTWindows(X, n, WINDOW_HOP, levels, out=Cb)  # noqa: E731
        e_inv = lambda Cb: et.inverseMODWTBatch(Cb, out=R)  # noqa: E731
        in_bytes = elen * 8
    else:
        eb = min(run.batch, batch_per_gpu) * nd
        hx = torch.empty((eb, unit), dtype=torch.float64).pin_memory()
        hx.copy_(torch.cat([b[0][:eb // nd].cpu() for b in run.bufs]))
        hc = torch.empty((eb, out_rows * unit), dtype=torch.float64).pin_memory()
        hr = torch.empty((eb, unit), dtype=torch.float64).pin_memory()
        in_bytes = eb * unit * 8
        if kind == "modwt":
            et = jw.CudaMODWTTransform(w, context=ectx)
            X, C, R = hx.numpy(), hc.numpy().reshape(eb, out_rows, n), hr.numpy()
            e_fwd = lambda Cb: et.forwardMODWTBatch(X, levels, out=Cb)  # noqa: E731
            e_inv = lambda Cb: et.inverseMODWTBatch(Cb, out=R)  # noqa: E731
        elif kind == "fwt2d":
            et = jw.CudaFastWaveletTransform(w, context=ectx)
            X, C, R = (t_.numpy().reshape(eb, n, n) for t_ in (hx, hc, hr))
            e_fwd = lambda Cb: et.forward2DBatch(X, levels, levels, out=Cb)  # noqa: E731
            e_inv = lambda Cb: et.reverse2DBatch(Cb, levels, levels, out=R)  # noqa: E731
        else:
            et = (jw.CudaFastWaveletTransform if kind == "fwt" else jw.CudaWaveletPacketTransform)(w, context=ectx)
            X, C, R = hx.numpy(), hc.numpy(), hr.numpy()
            e_fwd = lambda Cb: et.forwardBatch(X, levels, out=Cb)  # noqa: E731
            e_inv = lambda Cb: et.reverseBatch(Cb, levels, out=R)  # noqa: E731
        Xref = X
    step = lambda: (e_fwd(C), e_inv(C))  # noqa: E731
    step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(esteps):
        step()
    serial_ms = (time.perf_counter() - t0) * 1e3 / esteps
    assert float(np.max(np.abs(R - Xref))) <= 1e-10
    e_ms = None
    if with_stream:
        # streaming form of the same work: two host threads on one context, the forward call of batch k+1 runs while
        # the inverse call of batch k does, so the forward's D2H and the inverse's H2D share the link full-duplex
        from concurrent.futures import ThreadPoolExecutor
        hc2 = torch.empty_like(hc).pin_memory()
        Cs = [C, hc2.numpy().reshape(C.shape)]
        f_call = lambda k: e_fwd(Cs[k % 2])  # noqa: E731
        i_call = lambda k: e_inv(Cs[k % 2])  # noqa: E731

        def pipeline(ex, steps):
            ff = ex.submit(f_call, 0)
            for k in range(steps):
                ff.result()
                if k + 1 < steps:
                    ff = ex.submit(f_call, k + 1)   # forward of the next batch ...
                ex.submit(i_call, k).result()       # ... while this batch is inverted

        R[:] = 0.0
        with ThreadPoolExecutor(2) as ex:
            pipeline(ex, 3)   # untimed: second stream lane, staging pool growth, first touch of the second buffer
            barrier()
            t0 = time.perf_counter()
            pipeline(ex, psteps)
            e_ms = (time.perf_counter() - t0) * 1e3 / psteps
        assert float(np.max(np.abs(R - Xref))) <= 1e-10
        e_ms, serial_ms = reduce_max([e_ms, serial_ms])
    else:
        serial_ms = reduce_max([serial_ms])[0]
    c_bytes = eb * out_rows * unit * 8
    h2d = (in_bytes + c_bytes) * world        # forward input + inverse coefficients
    d2h = (c_bytes + eb * unit * 8) * world   # forward coefficients + inverse result
    gs = lambda ms: 2.0 * eb * unit * world / (ms * 1e-3) / 1e9  # noqa: E731
    out = {"value": gs(e_ms if e_ms else serial_ms), "unit": "Gsamples/s", "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": d2h, "ms_per_step": e_ms if e_ms else serial_ms, "serial_value": gs(serial_ms),
           "serial_ms_per_step": serial_ms, "batch_per_gpu": eb // nd,
           "gbs_per_direction": (h2d / ((e_ms if e_ms else serial_ms) * 1e-3)) / 1e9}
    ectx.close()
    return out


def multi_device_check(jw, torch, ngpu):
    """Rank 0 only: ONE context over all `ngpu` devices (the path a single JVM uses).  (i) one long series split over
    the devices -- forward and inverse MODWT, FWT and WPT -- must be bit-identical to the unsplit transform on device 0;
    (ii) the host-buffer e2e (jwc_modwt_forward + jwc_modwt_inverse, batch sharded by signal inside the C layer)."""
    out = {}
    devs = list(range(ngpu))
    ctx = jw.Context(devs)
    one = jw.Context([0])
    try:
        w = jw.wavelets.Daubechies4()
        n, J = 1 << 22, 6
        gen = torch.Generator(device="cuda:0")
        gen.manual_seed(0x5EED0077)
        x = torch.rand(n, dtype=torch.float64, device="cuda:0", generator=gen) * 2.0 - 1.0
        t1 = jw.CudaMODWTTransform(w, context=one)
        tP = jw.CudaMODWTTransform(w, context=ctx)
        c_ref = torch.empty((J + 1, n), dtype=torch.float64, device="cuda:0")
        t1.forwardMODWTDevice(x.data_ptr(), c_ref.data_ptr(), 1, n, J)
        xr_ref = torch.empty(n, dtype=torch.float64, device="cuda:0")
        t1.inverseMODWTDevice(c_ref.data_ptr(), xr_ref.data_ptr(), 1, n, J)
        torch.cuda.synchronize(0)
        bounds = [n * p // ngpu for p in range(ngpu + 1)]
        xs, cs, xrs = [], [], []
        for p in range(ngpu):
            a, b = bounds[p], bounds[p + 1]
            xs.append(x[a:b].to("cuda:%d" % p))
            cs.append(torch.empty((J + 1, b - a), dtype=torch.float64, device="cuda:%d" % p))
            xrs.append(torch.empty(b - a, dtype=torch.float64, device="cuda:%d" % p))
        for p in range(ngpu):
            torch.cuda.synchronize(p)
        tP.forwardMODWTSplitDevice([t.data_ptr() for t in xs], [t.data_ptr() for t in cs], n, J)
        tP.inverseMODWTSplitDevice([t.data_ptr() for t in cs], [t.data_ptr() for t in xrs], n, J)
        ok_f = all(torch.equal(cs[p].to("cuda:0"), c_ref[:, bounds[p]:bounds[p + 1]]) for p in range(ngpu))
        ok_i = all(torch.equal(xrs[p].to("cuda:0"), xr_ref[bounds[p]:bounds[p + 1]]) for p in range(ngpu))
        out["split_bit_identical"] = {"modwt_forward": bool(ok_f), "modwt_inverse": bool(ok_i)}
        del c_ref, cs
        # FWT / WPT of one long series split over the devices: the per-device chunks come back in the local layout
        # (include/jwavecuda.h); assembled band by band they must equal the unsplit transform
        for kind, T in (("fwt", jw.CudaFastWaveletTransform), ("wpt", jw.CudaWaveletPacketTransform)):
            tr1, trP = T(w, context=one), T(w, context=ctx)
            lv = 22 if kind == "fwt" else 6
            y_ref = torch.empty(n, dtype=torch.float64, device="cuda:0")
            z_ref = torch.empty(n, dtype=torch.float64, device="cuda:0")
            tr1.forwardDevice(x.data_ptr(), y_ref.data_ptr(), 1, n, lv)
            tr1.reverseDevice(y_ref.data_ptr(), z_ref.data_ptr(), 1, n, lv)
            torch.cuda.synchronize(0)
            ys = [torch.empty(bounds[p + 1] - bounds[p], dtype=torch.float64, device="cuda:%d" % p) for p in range(ngpu)]
            zs = [torch.empty_like(t) for t in ys]
            trP.forwardSplitDevice([t.data_ptr() for t in xs], [t.data_ptr() for t in ys], n, lv)
            trP.reverseSplitDevice([t.data_ptr() for t in ys], [t.data_ptr() for t in zs], n, lv)
            got = trP.splitLayoutToGlobal([t.cpu().numpy() for t in ys], n, lv)
            yr, zr = y_ref.cpu().numpy(), z_ref.cpu().numpy()
            back = np.concatenate([t.cpu().numpy() for t in zs])
            ls = trP.splitLevels(n, lv)
            scale = float(np.max(np.abs(x.cpu().numpy())))
            out["split_bit_identical"][kind + "_forward"] = bool(np.array_equal(got[n >> ls:], yr[n >> ls:]) if kind == "fwt"
                                                                 else np.array_equal(got, yr))
            out["split_max_err"] = dict(out.get("split_max_err", {}), **{
                kind + "_forward_vs_unsplit": float(np.max(np.abs(got - yr))) / scale,
                kind + "_round_trip": float(np.max(np.abs(back - x.cpu().numpy()))) / scale,
                kind + "_inverse_vs_unsplit": float(np.max(np.abs(back - zr))) / scale})
            out["split_levels_" + kind] = int(ls)
        out["split_series"] = "%d samples, Daubechies4, MODWT J=%d / FWT 22 levels / WPT 6 levels, %d chunks" % (n, J, ngpu)
        # (ii) host-buffer e2e through the single context: C2 shape, 256 signals per device
        B, N2, J2 = 256 * ngpu, 65536, 6
        hx = torch.rand((B, N2), dtype=torch.float64).mul_(2.0).sub_(1.0).pin_memory()
        hc = torch.empty((B, J2 + 1, N2), dtype=torch.float64).pin_memory()
        hr = torch.empty((B, N2), dtype=torch.float64).pin_memory()
        X, C, R = hx.numpy(), hc.numpy(), hr.numpy()
        tP.forwardMODWTBatch(X, J2, out=C)
        tP.inverseMODWTBatch(C, out=R)
        reps = 3
        t0 = time.perf_counter()
        for _ in range(reps):
            tP.forwardMODWTBatch(X, J2, out=C)
            tP.inverseMODWTBatch(C, out=R)
        ms = (time.perf_counter() - t0) * 1e3 / reps
        err = float(np.max(np.abs(R - X)))
        # parity of the sharded result against the single-device context on a subset of signals from every shard
        sel = sorted(set([0, B - 1] + [B * p // ngpu for p in range(ngpu)]))
        c1 = t1.forwardMODWTBatch(X[sel], J2)
        same = bool(np.array_equal(c1, C[sel]))
        out["e2e_single_process"] = {"value": 2.0 * B * N2 / (ms * 1e-3) / 1e9, "unit": "Gsamples/s", "ms_per_step": ms,
                                     "signals": B, "round_trip_max_err": err, "shards_match_single_device": same,
                                     "h2d_bytes_per_step": B * N2 * 8 * (J2 + 2), "d2h_bytes_per_step": B * N2 * 8 * (J2 + 2),
                                     "note": "jwc_modwt_forward + jwc_modwt_inverse back to back on pinned host buffers, "
                                             "one context over %d devices, one host thread per device inside the C "
                                             "layer" % ngpu}
        assert err <= 1e-10
    finally:
        ctx.close()
        one.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20,
                    help="timed steps (20 = the protocol of the round-1 record; --steps 50 runs into the board's power cap and "
                         "gives the sustained figure, see config.ms_per_step_first/last_10_steps)")
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="override signals per GPU (debug; invalidates the headline)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-per-config", action="store_true", help="skip per_config / c5_strong / multi_device / link_probe")
    ap.add_argument("--per-config-steps", type=int, default=10)
    ap.add_argument("--e2e-full-batch", action="store_true", help="also run the N=1 e2e once at the full batch")
    ap.add_argument("--tune", default="", help="comma list key=value passed to jwc_set_tuning")
    ap.add_argument("--flags", type=int, default=0)
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        print("note: warmup < 3 breaks the timing rules; use only for smoke runs", file=sys.stderr)

    kind, cls, levels, batch, n = WORKLOADS[args.workload]
    if args.batch > 0:
        batch = args.batch
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        return run_reference(args, kind, cls, levels, batch, n, rank, world)

    import torch
    import jwave_pro_b200 as jw

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (there is no CPU fallback)")
    distributed = world > 1
    if distributed:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        devices = [local_rank]
    else:
        # plain `python bench.py --gpus N` (no torchrun): one process drives N devices, one stream each
        devices = list(range(max(1, args.gpus)))
        torch.cuda.set_device(devices[0])
    n_gpus = world if distributed else len(devices)

    numa = None
    if distributed and not os.environ.get("JWC_NO_CPU_AFFINITY"):
        # one process per GPU: run (and first-touch the pinned staging buffers) on the CPUs next to this rank's GPU
        try:
            import pynvml
            pynvml.nvmlInit()
            pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local_rank))
            numa = "cpu affinity of rank = NVML ideal CPUs of its GPU (%d cpus)" % len(os.sched_getaffinity(0))
        except Exception as e:  # noqa: BLE001
            numa = "cpu affinity not set (%s)" % type(e).__name__
    ctx = jw.Context(devices)
    for kv in filter(None, args.tune.split(",")):
        k, v = kv.split("=")
        ctx.set_tuning(k, int(v))

    from jwave_pro_b200.sharding import reduce_max as _rmax, reduce_sum as _rsum
    rmax = (lambda v: _rmax(v, device="cuda")) if distributed else (lambda v: [float(x) for x in v])
    rsum = (lambda v: _rsum(v, device="cuda")) if distributed else (lambda v: [float(x) for x in v])

    def barrier():
        if distributed:
            dist.barrier()

    def sync_all():
        for d in devices:
            torch.cuda.synchronize(d)
        if distributed:
            dist.barrier()
            torch.cuda.synchronize()

    clock = ClockSampler(devices[0]) if (rank == 0 and not os.environ.get("JWC_NO_CLOCK_SAMPLER")) else None
    peak, peak_src = measured_peak()
    fp64_peak_cold = ctx.dfma_tflops()   # fp64 FMA rate of this device before anything has heated it (burst clocks)

    # ---- headline workload, device-resident ---------------------------------------------------------------------------
    run = DeviceRun(jw, torch, ctx, devices, rank, args.workload, batch_override=args.batch, flags=args.flags)
    res = run.measure(args.steps, args.warmup, sync_all, rmax)
    clocks = clock.window(res["t0"], res["t1"]) if clock else None
    unit = run.unit
    samples_per_step = 2.0 * run.batch * unit * n_gpus
    value = samples_per_step / (res["ms_per_step"] * 1e-3) / 1e9
    # fp64 FMA rate of this device measured with the library's own kernel, before the run (burst clocks) and right
    # after it (under the power cap the timed steps ran at); fractions are quoted against the LARGER of the two
    fp64_peak_after = ctx.dfma_tflops()
    fp64_peak = max(fp64_peak_cold, fp64_peak_after)
    fr = run.fractions(res, peak, fp64_peak)

    # ---- roofline of the dominant kernel = the longer direction of a step ------------------------------------------
    dom = "fwd" if res["fwd_ms"] >= res["inv_ms"] else "inv"
    oth = "inv" if dom == "fwd" else "fwd"
    traffic, traffic_src = None, None
    try:   # measured DRAM bytes of this kernel from the committed ncu capture, scaled to this launch
        for fn in ("r2_traffic.json", "r1_traffic.json"):
            pth = os.path.join(ROOT, "profiles", fn)
            if not os.path.exists(pth):
                continue
            tj = json.load(open(pth))
            key = "forward_bytes_per_sample" if dom == "fwd" else "inverse_bytes_per_sample"
            if args.workload in tj and key in tj[args.workload]:
                traffic = tj[args.workload][key] * run.batch * unit
                traffic_src = "profiles/%s (ncu --set full capture at a smaller batch, scaled per sample)" % fn
                break
    except Exception:  # noqa: BLE001
        pass
    names = {"fwd": "%s forward" % kind, "inv": "%s inverse" % kind}
    roofline = {"bound": "hbm", "achieved": fr[dom]["gbs"], "peak": peak, "unit": "GB/s", "frac": fr[dom]["hbm_frac"],
                "traffic": traffic, "traffic_source": traffic_src, "kernel": names[dom], "peak_source": peak_src,
                "algorithmic_bytes_per_launch": fr[dom]["algorithmic_bytes"], "avg_ms": fr[dom]["ms"],
                "fp64_frac": fr[dom]["fp64_frac"], "fp64_peak_tflops_measured_in_run": fp64_peak,
                "fp64_peak_tflops_before_and_after_timed_region": [fp64_peak_cold, fp64_peak_after],
                "other_direction": {"kernel": names[oth], "achieved": fr[oth]["gbs"], "frac": fr[oth]["hbm_frac"],
                                    "avg_ms": fr[oth]["ms"], "fp64_frac": fr[oth]["fp64_frac"]}}

    # ---- e2e: the same metric through the host-buffer C ABI (pinned host memory, copies inside the timed region) ----
    e2e = None
    if not args.no_e2e:
        esteps = max(2, min(args.steps, 5))
        psteps = max(20, 4 * esteps)   # long enough that the one-call pipeline fill is < 5 % of the timed region
        e2e = e2e_measure(jw, torch, run, devices, esteps, psteps, 8 if kind == "fwt2d" else 256, barrier, rmax,
                          world if distributed else 1)
        e2e["note"] = ("jwc_*_forward + jwc_*_inverse on pinned host buffers, %d units per GPU per step (bounded so "
                       "pinned staging stays small); PCIe-bound. value: forward of batch k+1 and inverse of batch k "
                       "issued from two host threads (both link directions busy), %d steps incl. pipeline fill; "
                       "serial_value: the two calls one after the other" % (e2e["batch_per_gpu"], psteps))
        if args.e2e_full_batch and n_gpus == 1 and kind != "windows":
            full = e2e_measure(jw, torch, run, devices, 2, 0, run.batch, barrier, rmax, 1, with_stream=False)
            e2e["full_batch"] = {"value": full["serial_value"], "ms_per_step": full["serial_ms_per_step"],
                                 "batch_per_gpu": full["batch_per_gpu"], "gbs_per_direction": full["gbs_per_direction"],
                                 "note": "the whole %d-unit batch of the config, forward call then inverse call" % run.batch}
    run.free()
    del run

    # ---- the other BASELINE configs, the strong-scaling config, the single-context multi-device path --------------
    per_config, c5_strong, multi_device, link = None, None, None, None
    extras = (args.workload == "c2" and args.batch == 0 and not args.no_per_config)
    if extras:
        ksteps = max(10, args.per_config_steps)
        if n_gpus == 1:
            per_config = {}
            for name in PER_CONFIG:
                r = DeviceRun(jw, torch, ctx, devices, rank, name, flags=args.flags)
                m = r.measure(ksteps, 3, sync_all, rmax)
                f = r.fractions(m, peak, fp64_peak)
                per_config[name] = {
                    "workload": workload_desc(name, r.kind, r.cls, r.levels, r.batch, r.n),
                    "value": 2.0 * r.batch * r.unit / (m["ms_per_step"] * 1e-3) / 1e9, "unit": "Gsamples/s",
                    "ms_per_step": m["ms_per_step"], "steps": ksteps, "warmup": 3,
                    "fwd_ms": m["fwd_ms"], "inv_ms": m["inv_ms"],
                    "hbm_frac": {"fwd": f["fwd"]["hbm_frac"], "inv": f["inv"]["hbm_frac"]},
                    "fp64_frac": {"fwd": f["fwd"]["fp64_frac"], "inv": f["inv"]["fp64_frac"]},
                    "achieved_gbs": {"fwd": f["fwd"]["gbs"], "inv": f["inv"]["gbs"]},
                    "achieved_tflops": {"fwd": f["fwd"]["tflops"], "inv": f["inv"]["tflops"]},
                    "bound": "fp64" if (f["fwd"]["fp64_frac"] or 0) > f["fwd"]["hbm_frac"] else "hbm",
                    "round_trip_max_err": m["round_trip_max_err"], "gpu_launches": m["launches"],
                    "clocks": clock.window(m["t0"], m["t1"]) if clock else None}
                if name == "c5":
                    c5_strong = {"series_total": C5_TOTAL_SERIES, "series_per_rank": r.batch, "n_gpus": 1,
                                 "value": per_config[name]["value"], "unit": "Gsamples/s",
                                 "ms_per_step": m["ms_per_step"], "fwd_ms": m["fwd_ms"], "inv_ms": m["inv_ms"],
                                 "scaling": "strong", "round_trip_max_err": m["round_trip_max_err"]}
                r.free()
                del r
        else:
            if C5_TOTAL_SERIES % n_gpus == 0:
                r = DeviceRun(jw, torch, ctx, devices, rank, "c5", batch_override=C5_TOTAL_SERIES // n_gpus,
                              flags=args.flags)
                m = r.measure(ksteps, 3, sync_all, rmax)
                perr = rmax([m["round_trip_max_err"]])[0]
                c5_strong = {"series_total": C5_TOTAL_SERIES, "series_per_rank": r.batch, "n_gpus": n_gpus,
                             "value": 2.0 * C5_TOTAL_SERIES * r.unit / (m["ms_per_step"] * 1e-3) / 1e9,
                             "unit": "Gsamples/s", "ms_per_step": m["ms_per_step"], "fwd_ms": m["fwd_ms"],
                             "inv_ms": m["inv_ms"], "scaling": "strong", "round_trip_max_err": perr,
                             "clocks": clock.window(m["t0"], m["t1"]) if clock else None,
                             "note": "8,192 series x 65,536, Daubechies20 J=8 sharded by series: %d per rank, time = "
                                     "max over ranks, no collective on the data path" % r.batch}
                r.free()
                del r
        if not args.no_e2e:
            link = link_probe(torch, "cuda:%d" % devices[0], rsum, barrier)
        if distributed:
            torch.cuda.empty_cache()
            barrier()
            if rank == 0:
                try:
                    multi_device = multi_device_check(jw, torch, world)
                except Exception as e:  # noqa: BLE001
                    multi_device = {"error": "%s: %s" % (type(e).__name__, e)}
            barrier()

    # ---- CPU baseline beside it (rank 0, N = 1 only) -----------------------------------------------------------
    cpu = None
    if rank == 0 and n_gpus == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        cpu = cpu_baseline_for(kind, cls, levels, unit, batch, cores, 12.0)
        if per_config is not None and "c4" in per_config:
            k4, c4cls, l4, b4, n4 = WORKLOADS["c4"]
            per_config["c4"]["cpu_baseline"] = cpu_baseline_for(k4, c4cls, l4, n4, b4, cores, 4.0)
            try:
                per_config["c4"]["cpu_baseline_parallel_wpt"] = cpu_baseline_for(k4, c4cls, l4, n4, b4, cores, 4.0,
                                                                                 schedule="parallel_wpt")
            except AttributeError:
                pass

    if clock:
        clock.stop()
    if rank == 0:
        line = {
            "metric": "MODWT/FWT/WPT Gsamples/s (forward+inverse sample-transforms per second)",
            "value": value, "unit": "Gsamples/s", "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_desc(args.workload, kind, cls, levels, batch, n),
                       "l2": "inputs larger than L2 (%.1f GiB read per direction vs 126 MB L2)" % (
                           (levels + 1 if kind in ("modwt", "windows") else 1) * batch * unit * 8 / 2 ** 30),
                       "sharding": "by signal, no collective", "numa": numa, "tune": args.tune, "flags": args.flags,
                       "round_trip_max_err": res["round_trip_max_err"],
                       "ms_per_step_first_%d_steps" % res["window_steps"]: res["ms_first_steps"],
                       "ms_per_step_last_%d_steps" % res["window_steps"]: res["ms_last_steps"]},
            "roofline": roofline, "directions": fr, "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks,
            "gpu_launches": res["launches"],
        }
        if per_config is not None:
            line["per_config"] = per_config
        if c5_strong is not None:
            line["c5_strong"] = c5_strong
        if multi_device is not None:
            line["multi_device"] = multi_device
        if link is not None:
            line["link_probe"] = link
        print(json.dumps(line))
    if distributed:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
