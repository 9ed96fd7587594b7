#!/usr/bin/env python3
"""bench.py -- the headline benchmark of BASELINE.json on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c2|c3haar|c3db8|c4|c5]

A "step" is one pass of the hot path over one batch of synthetic input: forward + inverse of the named
transform.  Default workload = BASELINE.json configs[1] ("c2"): batched MODWT Daubechies4 J=6 forward+inverse on
4,096 signals x 65,536 fp64 samples per GPU (weak scaling: every rank owns its own 4,096 signals, sharded by signal,
no collective on the data path).  `value` = sample-transforms per second summed over all ranks, counting
batch*N samples for the forward and batch*N for the inverse of every step, inputs resident in HBM.

Prints ONE JSON line (see the task contract): metric/value/unit, roofline (forward kernel: algorithmic bytes
8*(J+2) per sample / CUDA-event duration vs MEASURED_PEAKS.json), cpu_baseline (the oracle's restatement of the
reference's default FFT-convolution MODWT on the host cores, bounded sample), e2e (same metric through the
host-buffer C ABI with pinned host memory, copies inside the timed region), clocks, gpu_launches.

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
    "c5": ("modwt", "Daubechies20", 8, 8192, 65536),     # total 8,192 series -> per GPU 8192 / N (strong) in BASELINE; here per GPU
    "c3haar": ("fwt", "Haar1", 20, 1024, 1 << 20),
    "c3db8": ("fwt", "Daubechies8", 20, 1024, 1 << 20),
    "c4": ("wpt", "Symlet8", 6, 512, 65536),
    # SURVEY section 8f rows (not BASELINE configs): 2-D FWT of 32 matrices 4096 x 4096 at full depth (levels = lvlM = lvlN,
    # n = rows = cols), and the reference's sliding-window shape (512-sample windows, step 64, J = 8) over a 2^24 series
    "fwt2d": ("fwt2d", "Daubechies4", 12, 32, 4096),
    "windows": ("windows", "Daubechies4", 8, ((1 << 24) - 512) // 64 + 1, 512),
}
WINDOW_HOP = 64


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


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:  # noqa: BLE001
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi sampling of SM clock / throttle reasons.  Started before the warm-up (nvidia-smi takes a few hundred
    ms to produce its first line); stop(t0, t1) keeps the samples whose timestamp lies inside the timed region."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "20"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:  # noqa: BLE001
            self.p = None

    def stop(self, t0=None, t1=None):
        import datetime
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        time.sleep(0.05)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.p.kill()
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
        os.unlink(self.f.name)
        inside = [r for r in rows if t0 is not None and t0 - 0.02 <= r[0] <= t1 + 0.02]
        use = inside if inside else rows[-3:]
        if use:
            reasons = set()
            for r in use:
                reasons.update(r[3])
            out.update(sm_mhz=float(np.median([r[1] for r in use])), sm_max_mhz=float(max(r[2] for r in use)),
                       reasons=sorted(reasons), samples=len(use), samples_inside_timed_region=len(inside))
        return out


def cpu_reference_time(kind, cls, levels, n, nsig, threads, repeats=1):
    """Time the oracle's restatement of the reference's CPU path on `nsig` signals with `threads` host threads.
    MODWT uses the reference's default FFT convolution (AUTO picks FFT for every BASELINE config)."""
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
        else:
            c = oracle.batch(kind + "_fwd", X, levels, s, wv, nthreads=threads)
            oracle.batch(kind + "_rev", c, levels, w.getScalingReConstruction(), w.getWaveletReConstruction(),
                         nthreads=threads)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return best


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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="override signals per GPU (debug; invalidates the headline)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
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
    w = jw.wavelets.create(cls)
    unit = n * n if kind == "fwt2d" else n   # samples per unit of work (signal, window or matrix)
    if kind in ("modwt", "windows"):
        tr = jw.CudaMODWTTransform(w, context=ctx)
    elif kind in ("fwt", "fwt2d"):
        tr = jw.CudaFastWaveletTransform(w, context=ctx)
    else:
        tr = jw.CudaWaveletPacketTransform(w, context=ctx)

    # ---- synthetic inputs, resident in HBM (uniform(-1,1), seeded per rank) ---------------------------------
    out_rows = levels + 1 if kind in ("modwt", "windows") else 1
    series_len = (batch - 1) * WINDOW_HOP + n   # windows workload: one series per GPU, `batch` windows
    bufs = []
    for slot, d in enumerate(devices):
        with torch.cuda.device(d):
            gen = torch.Generator(device="cuda:%d" % d)
            gen.manual_seed(0x5EED0002 + rank * 16 + slot)
            xshape = (series_len,) if kind == "windows" else (batch, unit)
            x = torch.rand(xshape, dtype=torch.float64, device="cuda:%d" % d, generator=gen) * 2.0 - 1.0
            c = torch.empty((batch, out_rows * unit), dtype=torch.float64, device="cuda:%d" % d)
            xr = torch.empty((batch, unit), dtype=torch.float64, device="cuda:%d" % d)
            bufs.append((x, c, xr, torch.cuda.Stream(device=d)))   # explicit stream: kernels AND events go here
    if kind == "windows":
        import ctypes
        _lib = jw._native.load()
        _g, _h = (np.ascontiguousarray(v) for v in tr._filters())
        _dp = ctypes.POINTER(ctypes.c_double)

    def fwd(slot):
        x, c, xr, st = bufs[slot]
        if kind == "windows":
            rc = _lib.jwc_modwt_forward_windows_dev(ctx.handle, slot, ctypes.c_void_p(st.cuda_stream),
                                                    ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(c.data_ptr()),
                                                    series_len, n, WINDOW_HOP, levels, _g.ctypes.data_as(_dp),
                                                    _h.ctypes.data_as(_dp), len(_g), args.flags)
            assert rc == 0, _lib.jwc_last_error()
        elif kind == "fwt2d":
            tr.forward2DDevice(x.data_ptr(), c.data_ptr(), batch, n, n, levels, levels, stream=st.cuda_stream,
                               flags=args.flags, slot=slot)
        elif kind == "modwt":
            tr.forwardMODWTDevice(x.data_ptr(), c.data_ptr(), batch, n, levels, stream=st.cuda_stream, flags=args.flags,
                                  slot=slot)
        else:
            tr.forwardDevice(x.data_ptr(), c.data_ptr(), batch, n, levels, stream=st.cuda_stream, flags=args.flags,
                             slot=slot)

    def inv(slot):
        x, c, xr, st = bufs[slot]
        if kind == "fwt2d":
            tr.reverse2DDevice(c.data_ptr(), xr.data_ptr(), batch, n, n, levels, levels, stream=st.cuda_stream,
                               flags=args.flags, slot=slot)
        elif kind in ("modwt", "windows"):
            tr.inverseMODWTDevice(c.data_ptr(), xr.data_ptr(), batch, n, levels, stream=st.cuda_stream,
                                  flags=args.flags, slot=slot)
        else:
            tr.reverseDevice(c.data_ptr(), xr.data_ptr(), batch, n, levels, stream=st.cuda_stream, flags=args.flags,
                             slot=slot)

    def sync_all():
        for d in devices:
            torch.cuda.synchronize(d)
        if distributed:
            dist.barrier()
            torch.cuda.synchronize()

    clock = ClockSampler(devices[0]) if (rank == 0 and not os.environ.get("JWC_NO_CLOCK_SAMPLER")) else None
    for _ in range(args.warmup):
        for s in range(len(devices)):
            fwd(s)
            inv(s)
    sync_all()
    # correctness of the timed path itself: round trip must hold
    if kind == "windows":   # every reconstructed window against its span of the series
        pr = max(float((b[2] - b[0].unfold(0, n, WINDOW_HOP)).abs().max()) for b in bufs) if args.warmup > 0 else 0.0
    else:
        pr = max(float((b[2] - b[0]).abs().max()) for b in bufs) if args.warmup > 0 else 0.0

    # ---- timed region ---------------------------------------------------------------------------------------------
    ev = []
    launches0 = ctx.launch_count()
    sync_all()
    t_region0 = time.time()
    for s, d in enumerate(devices):
        with torch.cuda.device(d):
            st = bufs[s][3]
            e = [torch.cuda.Event(enable_timing=True) for _ in range(2 * args.steps + 1)]
            e[0].record(st)
            ev.append(e)
    for k in range(args.steps):
        for s, d in enumerate(devices):
            with torch.cuda.device(d):
                fwd(s)
                ev[s][2 * k + 1].record(bufs[s][3])
                inv(s)
                ev[s][2 * k + 2].record(bufs[s][3])
    sync_all()
    t_region1 = time.time()
    launches = ctx.launch_count() - launches0
    clocks = clock.stop(t_region0, t_region1) if clock else None
    total_ms = max(e[0].elapsed_time(e[-1]) for e in ev)
    fwd_ms = max(sum(e[2 * k].elapsed_time(e[2 * k + 1]) for k in range(args.steps)) for e in ev) / args.steps
    inv_ms = max(sum(e[2 * k + 1].elapsed_time(e[2 * k + 2]) for k in range(args.steps)) for e in ev) / args.steps
    if distributed:   # device time = max over ranks
        from jwave_pro_b200.sharding import reduce_max
        total_ms, fwd_ms, inv_ms = reduce_max([total_ms, fwd_ms, inv_ms], device="cuda")
    ms_per_step = total_ms / args.steps
    samples_per_step = 2.0 * batch * unit * n_gpus
    value = samples_per_step / (ms_per_step * 1e-3) / 1e9

    # ---- roofline of the dominant kernel (forward) ------------------------------------------------------------------
    peak, peak_src = measured_peak()
    bps = algorithmic_bytes_per_sample(kind, levels, n)
    bps_inv = 8.0 * (levels + 2) if kind == "windows" else bps   # the inverse reads J+1 rows, writes one, per window
    fwd_gbs = bps * batch * unit / (fwd_ms * 1e-3) / 1e9
    inv_gbs = bps_inv * batch * unit / (inv_ms * 1e-3) / 1e9
    traffic, traffic_src = None, None
    try:   # measured DRAM bytes of this kernel from the committed ncu capture, scaled to this launch
        tj = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))
        if args.workload in tj:
            traffic = tj[args.workload]["forward_bytes_per_sample"] * batch * unit
            traffic_src = "profiles/r1_traffic.json (ncu --set full capture at a smaller batch, scaled per sample)"
    except Exception:  # noqa: BLE001
        pass
    roofline = {"bound": "hbm", "achieved": fwd_gbs, "peak": peak, "unit": "GB/s", "frac": fwd_gbs / peak,
                "traffic": traffic, "traffic_source": traffic_src, "kernel": "%s forward" % kind, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": bps * batch * unit, "avg_ms": fwd_ms,
                "inverse": {"achieved": inv_gbs, "frac": inv_gbs / peak, "avg_ms": inv_ms}}

    # ---- e2e: the same metric through the host-buffer C ABI (pinned host memory, copies inside the timed region) ----
    e2e = None
    if not args.no_e2e:
        nd = len(devices)
        ectx = jw.Context(devices)
        if kind == "windows":
            # one series on the host; a bounded number of windows so that pinned staging stays ~1 GB
            eb = min(batch, 32768)
            elen = (eb - 1) * WINDOW_HOP + n
            hx = torch.empty(elen, dtype=torch.float64).pin_memory()
            hx.copy_(bufs[0][0][:elen].cpu())
            hc = torch.empty((eb, out_rows * unit), dtype=torch.float64).pin_memory()
            hr = torch.empty((eb, unit), dtype=torch.float64).pin_memory()
            et = jw.CudaMODWTTransform(w, context=ectx)
            X, C, R = hx.numpy(), hc.numpy().reshape(eb, out_rows, n), hr.numpy()
            Xref = np.lib.stride_tricks.sliding_window_view(X, n)[::WINDOW_HOP][:eb]
            e_fwd = lambda Cb: et.forwardMODWTWindows(X, n, WINDOW_HOP, levels, out=Cb)  # noqa: E731
            e_inv = lambda Cb: et.inverseMODWTBatch(Cb, out=R)  # noqa: E731
            in_bytes = elen * 8
        else:
            eb = min(batch, 8 if kind == "fwt2d" else 256) * nd
            hx = torch.empty((eb, unit), dtype=torch.float64).pin_memory()
            hx.copy_(torch.cat([b[0][:eb // nd].cpu() for b in bufs]))
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
        esteps = max(2, min(args.steps, 5))
        if distributed:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(esteps):
            step()
        serial_ms = (time.perf_counter() - t0) * 1e3 / esteps
        assert float(np.max(np.abs(R - Xref))) <= 1e-10
        # streaming form of the same work: two host threads on one context, the forward call of batch k+1 runs while
        # the inverse call of batch k does, so the forward's D2H and the inverse's H2D share the link full-duplex
        from concurrent.futures import ThreadPoolExecutor
        hc2 = torch.empty_like(hc).pin_memory()
        Cs = [C, hc2.numpy().reshape(C.shape)]
        f_call = lambda k: e_fwd(Cs[k % 2])  # noqa: E731
        i_call = lambda k: e_inv(Cs[k % 2])  # noqa: E731
        psteps = max(20, 4 * esteps)   # long enough that the one-call pipeline fill is < 5 % of the timed region

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
            if distributed:
                dist.barrier()
            t0 = time.perf_counter()
            pipeline(ex, psteps)
            e_ms = (time.perf_counter() - t0) * 1e3 / psteps
        assert float(np.max(np.abs(R - Xref))) <= 1e-10
        if distributed:
            from jwave_pro_b200.sharding import reduce_max
            e_ms, serial_ms = reduce_max([e_ms, serial_ms], device="cuda")
        wmul = world if distributed else 1
        c_bytes = eb * out_rows * unit * 8
        h2d = (in_bytes + c_bytes) * wmul        # forward input + inverse coefficients
        d2h = (c_bytes + eb * unit * 8) * wmul   # forward coefficients + inverse result
        gs = lambda ms: 2.0 * eb * unit * wmul / (ms * 1e-3) / 1e9  # noqa: E731
        e2e = {"value": gs(e_ms), "unit": "Gsamples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": e_ms, "serial_value": gs(serial_ms), "serial_ms_per_step": serial_ms,
               "batch_per_gpu": eb // len(devices),
               "note": "jwc_*_forward + jwc_*_inverse on pinned host buffers, %d units per GPU per step (bounded so "
                       "pinned staging stays small); PCIe-bound. value: forward of batch k+1 and inverse of batch k "
                       "issued from two host threads (both link directions busy), %d steps incl. pipeline fill; "
                       "serial_value: the two calls one after the other" % (eb // len(devices), psteps)}
        ectx.close()

    # ---- CPU baseline beside it (rank 0, N = 1 only) -----------------------------------------------------------
    cpu = None
    if rank == 0 and n_gpus == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        probe = cpu_reference_time(kind, cls, levels, unit, cores, cores)
        nsig = int(max(cores, min(batch, cores * max(1, round(12.0 / max(probe, 1e-3))))))
        dt = cpu_reference_time(kind, cls, levels, unit, nsig, cores)
        cpu = {"value": 2.0 * nsig * unit / dt / 1e9, "unit": "Gsamples/s", "cores": cores, "kind": "port",
               "sample": "%d of %d signals x %d samples, forward+inverse, %d threads (one signal per thread); C "
                         "restatement of the reference's %s" % (
                             nsig, batch, unit, cores,
                             "FFT-convolution MODWT" if kind in ("modwt", "windows") else kind.upper())}

    if rank == 0:
        line = {
            "metric": "MODWT/FWT/WPT Gsamples/s (forward+inverse sample-transforms per second)",
            "value": value, "unit": "Gsamples/s", "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_desc(args.workload, kind, cls, levels, batch, n),
                       "l2": "inputs larger than L2 (%.1f GiB read per direction vs 126 MB L2)" % (
                           (out_rows if kind in ("modwt", "windows") else 1) * batch * unit * 8 / 2 ** 30),
                       "sharding": "by signal, no collective", "numa": numa, "tune": args.tune, "flags": args.flags,
                       "round_trip_max_err": pr},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks, "gpu_launches": int(launches),
        }
        print(json.dumps(line))
    if distributed:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
