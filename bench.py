#!/usr/bin/env python
"""bench.py -- sampled+folded+dead-timed photon events/s on the bright-burst configuration.

Contract: ``python bench.py --gpus N --steps K --warmup W`` (under torchrun for N > 1) prints ONE
JSON line on rank 0.  A *step* is one bright long GRB (BASELINE.json configs[1]: GBMGRB_CPL
peak_flux=1e-4, duration=300 s, all 14 GBM detectors = 14 (GRB x detector) units) through the whole
hot path: fmax, source thinning, CPL energies, DRM fold, background, combine, GBM dead time.
At N > 1 every rank simulates its own burst (units shard by GRB, weak scaling) and the per-unit
event counts are exchanged with one small NCCL all_gather (the event-offset gather).

``value``  : events/s with the unit records and detector tables already resident in HBM.
``e2e``    : the same through the host-facing call with HOST buffers: unit records + tables are
             copied H2D and all six event lists of every detector are copied D2H into pinned host
             memory inside the timed region.
``--impl reference`` times the reference's own numba kernels and classes (staged under oracle/_ref by
oracle/stage_ref.py; the C port of oracle/ if they are not at hand) on all host cores on a bounded,
stratified sample of the same workload.
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

METRIC = "sampled+folded+dead-timed photon events/s"
UNIT = "events/s"
WORKLOAD = "C2 bright long GBMGRB_CPL: peak_flux=1e-4 duration=300s trise=5 tdecay=60 z=1 alpha=-0.66 " \
           "ep_start=500 ep_tau=2, 14 GBM detectors, synthetic 140x128 DRMs, bkg 500 Hz over [-100,300]s"
BYTES_PER_EVENT = 18      # (f64 time + u8 channel) x (component list + total list)


_REAL_STDOUT = None


def emit(line):
    """the single JSON line, on the process's original stdout"""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def bright_params():
    from cosmogrb_b200 import workloads

    return dict(workloads.BRIGHT_LONG)


def build_inputs(grb_index=0):
    """Host-side inputs of one bright GRB: 14 unit records + 14 detector table blocks."""
    from cosmogrb_b200 import workloads

    return workloads.burst_inputs(bright_params(), grb_index=grb_index)


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler(object):
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [r for r in self.rows if t0 <= r[0] <= t1] or self.rows[-3:]
        for _, line in rows:
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0]))
                smax.append(float(f[1]))
                for n, v in zip(names, f[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle on a bounded stratified sample of the workload
# ------------------------------------------------------------------------------------------------
def cpu_sample(n_windows, window_s, threads, seed=1):
    """Time the CPU oracle on `n_windows` windows of `window_s` seconds spread evenly over the burst
    of detector n0 (source thinned with the full burst's fmax; background window scaled 400/300).
    Returns (events, seconds, n_candidates, n_trials)."""
    from concurrent.futures import ThreadPoolExecutor

    from cosmogrb_b200 import workloads as h
    from oracle import oracle as orc

    p = bright_params()
    rsp = h.response("n0")
    s = orc.source(kind=0, interp_extrap="linear", peak_flux=p["peak_flux"], ep_start=p["ep_start"], ep_tau=p["ep_tau"],
                   alpha=p["alpha"], trise=p["trise"], tdecay=p["tdecay"], z=p["z"], emin=rsp.emin, emax=rsp.emax)
    ea = rsp.effective_area_packed
    fm = orc.fmax(s, ea, 0.0, p["duration"])
    w = h.bkg_weights("n0")
    edges, cum = rsp.energy_edges, rsp.cumulative_matrix
    dur = p["duration"]

    def one(k):
        a = (k + 0.5) * dur / n_windows - 0.5 * window_s
        ba = -100.0 + a * 400.0 / dur
        n, c = orc.process_detector(s, edges, ea, cum, w, a, a + window_s, ba, ba + window_s * 400.0 / dur, 500.0,
                                    seed_numba=seed * 7919 + k, seed_numpy=seed * 104729 + k, fmax=fm)
        return n, int(c[3]), int(c[4])

    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as ex:      # ctypes releases the GIL
        res = list(ex.map(one, range(n_windows)))
    dt = time.perf_counter() - t0
    return sum(r[0] for r in res), dt, sum(r[1] for r in res), sum(r[2] for r in res)


def reference_sampler():
    """(kind, sample(n_windows, window_s, seed) -> (events, seconds, detail)) of the CPU arm: the reference's own numba
    kernels and classes when its files are at hand (/root/reference, or the copy staged under oracle/_ref for the GPU
    box: oracle/stage_ref.py, oracle/ref_numba.py), one forked process per host core; otherwise the C port of
    oracle/, one thread per host core."""
    threads = len(os.sched_getaffinity(0))
    p = bright_params()
    try:
        from oracle import ref_numba

        root, why = ref_numba.available()
        if root is None or os.environ.get("CGRB_REFERENCE_ARM") == "port":
            raise RuntimeError(why or "port forced")
        ref_numba.setup(p)

        def sample(n_windows, window_s, seed=1):
            ev, n_src, dt = ref_numba.run_sample(n_windows, window_s, threads, p["duration"])
            return ev, dt, f"{n_src} source events"

        return "reference", threads, sample, f"numba kernels + classes of the reference ({root})"
    except Exception as e:          # noqa: BLE001
        from oracle import oracle as orc

        orc.build()

        def sample(n_windows, window_s, seed=1):
            ev, dt, nc, ntr = cpu_sample(n_windows, window_s, threads, seed=seed)
            return ev, dt, f"{nc} candidates, {ntr} rejection trials"

        return "port", threads, sample, f"C port of the reference's numba path (oracle/); the reference itself: {e}"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    kind, threads, sample_fn, how = reference_sampler()
    n_windows, window_s = 4 * threads, 0.1          # ~10-20 s of CPU work per step
    for _ in range(max(args.warmup, 0)):
        sample_fn(threads, 0.01)
    ev, secs, detail = 0, 0.0, ""
    for k in range(args.steps):
        e, dt, detail = sample_fn(n_windows, window_s, seed=k + 1)
        ev += e
        secs += dt
    value = ev / secs
    sample = (f"{n_windows} windows x {window_s}s of detector n0 spread evenly over the 300 s burst, thinned with the whole "
              f"burst's fmax (fraction {n_windows * window_s / 300:.4f} of one of the 14 units) per step; {how}; last step: {detail}")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)
    return 0


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
KERNEL_BYTES = {
    "k_sample_energy_env": 9.0,   # reads the photon time (8 B), writes its channel (1 B): the fold is fused
    # algorithmic HBM bytes per EVENT of each kernel's own product
    "k_background": 9.0,        # writes f64 time + u8 channel per background event
    "k_merge_deadtime": 18.0,   # reads 9 B (component lists), writes <= 9 B (survivors) per merged event
    "k_combine": 18.0,          # reads 9 B, writes 9 B per merged event
    "k_deadtime_count": 9.0,    # reads 9 B per merged event
    "k_deadtime_write": 18.0,   # reads 9 B, writes <= 9 B per merged event
    "k_source_thin": 16.0,      # reads the stored running gap sum, writes the accepted time (per source event)
    "k_gap_prefix": 8.0,        # writes one f64 running gap sum per candidate (counted per event here)
    "k_source_gather": 16.0,    # moves one f64 per source event
    "k_sample_energy": 16.0,    # reads time, writes energy per source event
    "k_digitize": 9.0,          # reads energy, writes channel per source event
}


def start_memlog():
    """CGRB_MEMLOG=1: log this process's resident set size to stderr twice a second (debugging aid)"""
    if not os.environ.get("CGRB_MEMLOG"):
        return

    def loop():
        while True:
            try:
                with open("/proc/self/status") as f:
                    rss = [l for l in f if l.startswith("VmRSS")][0].split()[1]
                sys.stderr.write(f"[memlog] {time.time():.1f} rss_kb={rss} leg={_LEG[0]}\n")
                sys.stderr.flush()
            except Exception:
                pass
            time.sleep(0.5)

    threading.Thread(target=loop, daemon=True).start()


_LEG = ["init"]


def run_gpu(args):
    import torch

    start_memlog()

    import __graft_entry__ as ge

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if rank == 0:
        ge.build()
    dist = None
    if world > 1:
        import torch.distributed as dist

        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist.barrier()
    else:
        torch.cuda.set_device(0)
    from cosmogrb_b200 import _capi
    from cosmogrb_b200.engine import Engine, Philox

    dev = torch.device("cuda", local if world > 1 else 0)
    from cosmogrb_b200.engine import get_engine

    eng = get_engine()              # the process-wide engine the public API uses
    assert eng.device == dev, (eng.device, dev)
    units, tabs = build_inputs(grb_index=rank)
    seed = 0x00C0560612B

    def sync_all():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def gather_counts(counts_dev):
        """the event-offset gather: per-unit (n_src, n_bkg, n_kept) of all ranks -> global offsets"""
        if dist is None:
            return None
        c = counts_dev.view(torch.int64).view(len(units), -1)[:, [1, 2, 4]].contiguous()   # n_src, n_bkg, n_kept
        out = torch.empty((world,) + tuple(c.shape), dtype=c.dtype, device=dev)
        dist.all_gather_into_tensor(out, c)
        return torch.cumsum(out.view(-1, 3), dim=0)

    # ---- resident-input arm ---------------------------------------------------------------------
    _LEG[0] = "resident"
    batch = eng.new_batch(units, tabs)              # units + tables resident, fmax on device

    def step_resident(k):
        eng.reset_counts(batch)
        eng.process(batch, Philox(seed + k))
        gather_counts(batch.d_counts)

    for k in range(args.warmup):
        step_resident(k)
    sync_all()
    clocks = ClockSampler(dev.index)
    clocks.start()
    time.sleep(0.3)
    _capi.profile_enable(True)
    l0 = _capi.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    events_total = 0
    sync_all()
    t_wall0 = time.time()
    ev0.record()
    for k in range(args.steps):
        step_resident(args.warmup + k)
    ev1.record()
    sync_all()
    t_wall1 = time.time()
    ms = ev0.elapsed_time(ev1)
    launches = _capi.launch_count() - l0
    prof = _capi.profile_read()
    _capi.profile_enable(False)
    clk = clocks.stop(t_wall0, t_wall1)
    counts = eng.fetch_counts(batch)
    eng.check_status(batch)
    ev_step = int(counts["n_src"].sum() + counts["n_bkg"].sum())      # last step (all steps are ~equal)
    events_total = ev_step * args.steps
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    tot = torch.tensor([events_total], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms_max, events_all = float(t.item()), float(tot.item())
    value = events_all / (ms_max * 1e-3)

    # per-kernel durations (live CUDA events, per launch, averaged over the timed steps)
    per = {}
    for name, m in prof:
        a = per.setdefault(name, [0.0, 0])
        a[0] += m
        a[1] += 1
    kern = {n: {"ms_per_launch": v[0] / v[1], "launches": v[1], "ms_per_step": v[0] / args.steps} for n, v in per.items()}
    total_kernel_ms = sum(v["ms_per_step"] for v in kern.values()) or 1.0
    for v in kern.values():
        v["share"] = v["ms_per_step"] / total_kernel_ms
    top = max(kern, key=lambda n: kern[n]["ms_per_step"])
    n_src, n_bkg = int(counts["n_src"].sum()), int(counts["n_bkg"].sum())
    n_tot = n_src + n_bkg
    units_of = {"k_background": n_bkg, "k_merge_deadtime": n_tot, "k_gap_prefix": n_src, "k_combine": n_tot, "k_deadtime_count": n_tot, "k_deadtime_write": n_tot,
                "k_source_thin": n_src, "k_source_gather": n_src, "k_sample_energy": n_src, "k_digitize": n_src,
                "k_sample_energy_env": n_src}
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    # live denominators measured in this run (torch kernels as a yardstick only): a write-only fill and a copy
    def live_bw(fn, nbytes, reps=5):
        fn()
        torch.cuda.synchronize(dev)
        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 0.0
        for _ in range(reps):
            a.record()
            fn()
            b_.record()
            b_.synchronize()
            best = max(best, nbytes / (a.elapsed_time(b_) * 1e-3) / 1e9)
        return best

    try:
        buf_a = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
        buf_b = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
        peaks_live = {"fill_write_gbs": live_bw(lambda: buf_a.fill_(1), 1 << 30),
                      "copy_read_write_gbs": live_bw(lambda: buf_b.copy_(buf_a), 2 << 30)}
        # bare device->host yardstick of the e2e arm: ONE plain async copy of 1 GiB into pinned host memory, all
        # ranks at the same time (at N > 1 the GPUs of a box share the host's PCIe / memory side)
        pin = torch.empty(1 << 30, dtype=torch.uint8, pin_memory=True)
        pin.copy_(buf_a, non_blocking=True)
        # SUSTAINED and barrier-synchronised: 4 x 1 GiB back to back on every rank at the same moment.  (A best-of-3
        # of single copies, as in round 1, catches moments when the other ranks are idle: at 8 GPUs it read 18.9 GB/s
        # per GPU where the box sustains 11.4 -- scripts/e2e_probe.py, 91 GB/s for the eight together.)
        sync_all()
        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(4):
            pin.copy_(buf_a, non_blocking=True)
        b_.record()
        b_.synchronize()
        d2h_rate = 4 * (1 << 30) / (a.elapsed_time(b_) * 1e-3) / 1e9
        peaks_live["d2h_pinned_gbs_per_gpu_all_ranks_concurrent"] = d2h_rate
        if dist is not None:
            agg = torch.tensor([d2h_rate], dtype=torch.float64, device=dev)
            dist.all_reduce(agg, op=dist.ReduceOp.SUM)
            peaks_live["d2h_pinned_gbs_all_ranks_sum"] = float(agg.item())
        sync_all()
        del buf_a, buf_b, pin
    except Exception as e:          # noqa: BLE001
        peaks_live = {"error": str(e)}
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"

    # DRAM bytes per launch measured by ncu (--set full) on this same workload, kept under profiles/
    traffic, issue_frac, warp_inst, tj = {}, {}, {}, {}
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tj = json.load(f)
        traffic, issue_frac = tj.get("dram_bytes_per_launch", {}), tj.get("issue_slots_busy_frac", {})
        warp_inst = tj.get("warp_inst_per_launch", {})
    except Exception:
        pass

    def hbm_roofline(name):
        key = name.split("<")[0]
        nbytes = KERNEL_BYTES.get(key, 0.0) * units_of.get(key, 0)
        ach = nbytes / (kern[name]["ms_per_launch"] * 1e-3) / 1e9 if name in kern else 0.0
        return {"kernel": name, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": traffic.get(name, traffic.get(key)), "algorithmic_bytes_per_launch": nbytes, "ms_per_launch": kern[name]["ms_per_launch"],
                "share_of_step": kern[name]["share"], "peak_source": peak_src,
                "issue_slots_busy_frac_ncu": issue_frac.get(name, issue_frac.get(key))}

    def issue_roofline(name):
        """The regime that binds an RNG / rejection kernel: warp instructions issued per second against the SMs'
        issue slots (SMs x 4 schedulers x 1 instruction per clock x the SM clock sampled during the timed region).
        The instruction count per launch is ncu's smsp__inst_executed.sum of the tracked capture of this same
        workload (profiles/traffic.json); the duration is this run's live CUDA-event time of the kernel.  The
        instruction count is the builder's choice, so `warp_inst_per_event` is the progress metric, `frac` how
        busy the issue slots are."""
        key = name.split("<")[0]
        wi = warp_inst.get(name, warp_inst.get(key))
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        mhz = clk.get("sm_mhz") or clk.get("sm_max_mhz") or 1965.0
        peak_i = sms * 4 * mhz * 1e6 / 1e9                 # G warp-instructions / s
        ach = (wi / (kern[name]["ms_per_launch"] * 1e-3) / 1e9) if wi else None
        nu = units_of.get(key, 0)
        return {"kernel": name, "bound": "issue", "achieved": ach, "peak": peak_i, "unit": "Gwarp-inst/s",
                "frac": (ach / peak_i) if ach else None, "traffic": traffic.get(name, traffic.get(key)),
                "warp_inst_per_launch_ncu": wi, "warp_inst_per_event": (wi / nu) if (wi and nu) else None,
                "ms_per_launch": kern[name]["ms_per_launch"], "share_of_step": kern[name]["share"],
                "peak_source": f"{sms} SMs x 4 issue slots x {mhz:.0f} MHz (clock sampled in this run)",
                "issue_slots_busy_frac_ncu": issue_frac.get(name, issue_frac.get(key)),
                "hbm": {k: v for k, v in hbm_roofline(name).items() if k in ("achieved", "peak", "frac", "unit",
                                                                           "algorithmic_bytes_per_launch", "peak_source")}}

    ISSUE_BOUND = ("k_sample_energy_env", "k_source_thin", "k_gap_prefix", "k_sample_energy", "k_thin_nodes")
    roofline = issue_roofline(top) if top.split("<")[0] in ISSUE_BOUND else hbm_roofline(top)
    roofline_issue_kernels = [issue_roofline(n) for n in kern if n.split("<")[0] in ISSUE_BOUND and n != top]
    hbm_kernels = [hbm_roofline(n) for n in kern if n.split("<")[0] in ("k_background", "k_merge_deadtime", "k_combine",
                                                                           "k_deadtime_count", "k_deadtime_write",
                                                                           "k_source_gather")]

    # ---- e2e arm: the public API, GBMGRB_CPL(...).go(): host objects in, host arrays out -------------
    # The burst object is built once (like the reference, construction = responses, backgrounds and
    # fmax; go() = the hot call).  Every go() uploads the unit records and the 14 detector tables
    # (the engine's table cache is cleared so the copy really happens) and returns the six event lists
    # of each of the 14 detectors in pinned host memory.
    from cosmogrb_b200 import gbm

    _LEG[0] = "e2e"
    p = bright_params()
    grb = None if args.no_e2e else gbm.GBMGRB_CPL(ra=312.0, dec=-62.0, T0=0.1, name=f"bench_{rank}", **p)

    def step_e2e(k):
        eng._dtab_cache.clear()
        grb.go(seed=seed + 1000 + k, stream_base=16 * rank)
        ev = nb = 0
        for lc in grb.lightcurves.values():
            st = lc.lightcurve_storage
            ev += st.n_counts_source + st.n_counts_background
            nb += 9 * (st.n_counts + st.n_counts_source + st.n_counts_background)
        return {"events": ev, "d2h_bytes": nb + 14 * 64}

    h2d = units.nbytes + sum(t.nbytes for t in tabs)
    e2e_steps = 0 if args.no_e2e else max(1, min(args.steps, 3))
    for k in range(0 if args.no_e2e else max(1, min(args.warmup, 2))):
        step_e2e(k)
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    d2h, ev_e2e = 0, 0
    e0.record()
    for k in range(e2e_steps):
        out = step_e2e(10 + k)
        d2h += out["d2h_bytes"]
        ev_e2e += out["events"]
    e1.record()
    sync_all()
    e2e_steps = max(e2e_steps, 1)
    ms_e2e = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    ev_e2e_t = torch.tensor([ev_e2e], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(ms_e2e, op=dist.ReduceOp.MAX)
        dist.all_reduce(ev_e2e_t, op=dist.ReduceOp.SUM)
    e2e_value = float(ev_e2e_t.item()) / (float(ms_e2e.item()) * 1e-3)
    # the e2e arm against its own roofline: bytes this rank copied device->host per second vs the bare-copy yardstick
    e2e_d2h_gbs = (d2h / (float(ms_e2e.item()) * 1e-3) / 1e9) if d2h else 0.0
    d2h_peak = peaks_live.get("d2h_pinned_gbs_per_gpu_all_ranks_concurrent")

    # ---- population leg (BASELINE.json configs[4]: 100,000 GRBs x 14 detectors, STRONG scaling) ------------
    # GRBs/s through the public API, Universe.go(): the SAME population at every N, its bursts dealt over the
    # ranks (LPT on the sizing pass's costs); unit records built on the host, events stay in HBM, the GBM
    # trigger runs on the device, only the per-unit counters + 32 B per detector come back, then the NCCL
    # event-count gather.  `--population P` (P bursts per GPU, weak scaling) is kept for profiling runs.
    population = None
    _LEG[0] = "population"
    n_pop = args.population * world if args.population > 0 else args.population_total
    if n_pop > 0:
        from cosmogrb_b200.instruments.gbm.gbm_universe import GBM_CPL_Universe
        from cosmogrb_b200.universe import synthetic_population

        strong = args.population <= 0
        # warm-up pass on a 2000-burst-per-GPU population: the steady state of a long run, where the pooled event
        # buffers have reached their final sizes and NCCL has been through its first collective (first-touch costs
        # -- cudaMalloc of tens of GB by the caching allocator -- took between 5 and 260 ms in otherwise identical runs)
        warm = GBM_CPL_Universe(synthetic_population(min(n_pop, 2000 * world), seed=12345), save_path=None)
        warm.go(save=False, seed=seed, trigger=True)
        del warm
        # (1) per-kernel attribution on a sample of the same population model, on ONE compute stream: the timed run
        # below keeps two batches in flight on two streams, whose kernels overlap -- CUDA events around a launch
        # would then charge a kernel with its neighbour's time
        n_prof = min(n_pop, 5000 * world)
        os.environ["COSMOGRB_B200_STREAMS"] = "1"
        prof_uni = GBM_CPL_Universe(synthetic_population(n_prof, seed=12345), save_path=None)
        sync_all()
        _capi.profile_enable(True)
        prof_uni.go(save=False, seed=seed, trigger=True)
        sync_all()
        pprof = _capi.profile_read()
        _capi.profile_enable(False)
        os.environ.pop("COSMOGRB_B200_STREAMS", None)
        pec = np.maximum(prof_uni.event_counts, 0)
        prof_sim = int((np.asarray(prof_uni.event_counts).reshape(prof_uni.n_grbs, -1).min(axis=1) >= 0)[prof_uni.local_grbs].sum())
        n_src_local = int(pec[prof_uni.local_grbs][:, :, 0].sum())
        n_bkg_local = int(pec[prof_uni.local_grbs][:, :, 1].sum())
        n_tot_local = n_src_local + n_bkg_local
        del prof_uni
        # (2) the timed run: the whole population, no per-launch events
        uni = GBM_CPL_Universe(synthetic_population(n_pop, seed=12345), save_path=None)
        sync_all()
        eng.host_times.clear()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()
        uni.go(save=False, seed=seed, trigger=True)      # events stay in HBM; the GBM trigger runs on the device
        p1.record()
        sync_all()
        ms_local = p0.elapsed_time(p1)
        ms_pop = torch.tensor([ms_local], dtype=torch.float64, device=dev)
        ms_all = [ms_pop.clone()]
        if dist is not None:
            ms_all = [torch.zeros_like(ms_pop) for _ in range(world)]
            dist.all_gather(ms_all, ms_pop)
            dist.all_reduce(ms_pop, op=dist.ReduceOp.MAX)
        ms_pop = float(ms_pop.item())
        # simulated bursts only: a skipped burst (runaway lambda(t)) reads -1 and counts for nothing
        simulated = np.asarray(uni.event_counts).reshape(uni.n_grbs, -1).min(axis=1) >= 0
        n_sim = int(simulated.sum())
        ec = np.maximum(uni.event_counts, 0)
        n_ev = int(ec[:, :, 0].sum() + ec[:, :, 1].sum())
        pk = {}
        for name, m in pprof:
            pk[name] = pk.get(name, 0.0) + m
        tot_k = sum(pk.values()) or 1.0
        per1000 = 1000.0 / max(prof_sim, 1)
        # HBM roofline of the streaming kernels on the sample (this rank's share of its events)
        ptraffic = tj.get("population", {}) if isinstance(tj, dict) else {}
        pop_roof = {}
        for name, nb_, nu in (("k_background", 9.0, n_bkg_local), ("k_merge_deadtime", 18.0, n_tot_local),
                              ("k_trig_bin", 9.0, n_tot_local)):
            if name in pk and pk[name] > 0:
                ach = nb_ * nu / (pk[name] * 1e-3) / 1e9
                pop_roof[name] = {"ms_per_1000_bursts": pk[name] * per1000, "achieved_gbs": ach, "frac_of_measured_hbm": ach / peak,
                                  "algorithmic_bytes_per_event": nb_,
                                  "traffic_bytes_per_event_ncu": ptraffic.get("dram_bytes_per_event", {}).get(name),
                                  "issue_slots_busy_frac_ncu": ptraffic.get("issue_slots_busy_frac", {}).get(name)}
        population = {
            "workload": f"C5: {n_pop} GRBs x 14 detectors over {world} GPU(s), synthetic popsynth-style population "
                        "(seed 12345), events stay in HBM, GBM trigger on the device, counters + 32 B per detector read back; bursts whose candidate count "
                        "does not fit the memory budget are skipped and NOT counted (the reference's lambda(t) runs away for low Epeak)",
            "metric": "GRBs/s", "scaling": "strong" if strong else "weak",
            "n_grbs": int(uni.n_grbs), "n_simulated": n_sim, "n_skipped": int(uni.n_grbs) - n_sim,
            "n_skipped_local": len(uni.skipped), "events": n_ev,
            "gbm_trigger": {"on_device": True, "detected_local": int(uni.detected.sum()), "local_grbs": int(len(uni.local_grbs)),
                            "kernel_ms_per_1000_bursts": round((pk.get("k_trig_bin", 0.0) + pk.get("k_trig_eval", 0.0)) * per1000, 3)},
            "ms": ms_pop,
            "grbs_per_s": n_sim / (ms_pop * 1e-3),
            "events_per_s": n_ev / (ms_pop * 1e-3), "batches_per_gpu": int(uni.n_batches),
            "ms_all_ranks": [round(float(x.item()), 2) for x in ms_all],
            "kernel_sample": f"per-kernel times: a separate pass over {n_prof} bursts of the same population model on ONE compute "
                             f"stream ({prof_sim} simulated on this rank); the timed run keeps two batches in flight on two streams",
            "kernel_ms_per_1000_bursts": tot_k * per1000,
            "kernels_ms_per_1000_bursts": {k: round(v * per1000, 4) for k, v in sorted(pk.items(), key=lambda kv: -kv[1])[:24]},
            "hbm_kernels": pop_roof, "host_seconds": {k: round(v, 4) for k, v in uni.timings.items()},
            "host_detail": {k: round(v, 4) for k, v in eng.host_times.items()},
        }

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu:
        # the CPU arm in a process of its own (it forks workers: not from a process that holds a CUDA context)
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                               capture_output=True, text=True, timeout=900)
            cpu_baseline = json.loads(r.stdout.strip().splitlines()[-1])["cpu_baseline"]
        except Exception as e:      # noqa: BLE001
            cpu_baseline = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {e}"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "units_per_step_per_gpu": len(units),
                       "events_per_step_per_gpu": ev_step, "source_events": n_src, "background_events": n_bkg,
                       "candidates": int(counts["n_candidates"].sum()), "rejection_trials": int(counts["n_trials"].sum()),
                       "exact_lambda_evaluations": int(counts["n_exact"].sum()),
                       "kept_after_dead_time": int(counts["n_kept"].sum()), "rng": "philox4x32-10",
                       "l2": "no flush needed: each step streams > 2 GB of event buffers (L2 is 126 MB)",
                       "parallelism": f"{world} x (1 GRB = 14 units per GPU per step)",
                       # second metric of BASELINE.json ("GRBs/s at 1/2/4/8 B200"): the 100k-burst population under
                       # STRONG scaling (same population at every N); details under the top-level "population" key
                       "population": None if population is None else {
                           k: population[k] for k in ("metric", "scaling", "n_grbs", "n_simulated", "n_skipped",
                                                      "grbs_per_s", "events_per_s", "ms", "ms_all_ranks")}},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h // e2e_steps,
                    "steps": e2e_steps, "d2h_gbs_per_gpu": e2e_d2h_gbs, "d2h_yardstick_gbs_per_gpu": d2h_peak,
                    "frac_of_d2h_yardstick": (e2e_d2h_gbs / d2h_peak) if d2h_peak else None,
                    "note": "bound by the device->host copy of the six event lists (18 B per event); the yardstick is a bare "
                            "sustained copy into pinned memory (4 x 1 GiB back to back) issued by all ranks at the same moment: "
                            "the host side of a box is shared by its GPUs"},
            "gpu_launches": int(launches),
            "rates": {"candidates_per_s": float(counts["n_candidates"].sum()) * world / (ms_max / args.steps * 1e-3),
                      "rejection_trials_per_s": float(counts["n_trials"].sum()) * world / (ms_max / args.steps * 1e-3)},
            "peaks_live": peaks_live,
            "roofline": roofline, "roofline_hbm_kernels": hbm_kernels, "roofline_issue_kernels": roofline_issue_kernels,
            "kernels": kern, "cpu_baseline": cpu_baseline, "clocks": clk, "population": population,
        }
        emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the e2e leg (profiling runs)")
    ap.add_argument("--population-total", type=int, default=100000,
                    help="GRBs of the population leg, dealt over ALL ranks (BASELINE configs[4], strong scaling); 0 = skip")
    ap.add_argument("--population", type=int, default=0,
                    help="if > 0: GRBs PER GPU of the population leg instead (weak scaling; profiling runs)")
    args = ap.parse_args()
    # keep stdout clean for the ONE JSON line: libraries (NCCL's version banner, torchrun warnings)
    # write to fd 1 too, so everything else goes to stderr and the line is written to the real stdout
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
