"""Host side of the photon-event engine: device buffers (torch tensors), work lists, launches.

PyTorch is used only for device memory, streams and host<->device copies; all arithmetic of the
hot path happens in the CUDA kernels behind the C ABI (include/cosmogrb_b200.h).  There is no
CPU fallback: without a CUDA device or without the shared library every call raises.

Vocabulary: a *unit* is one (GRB, detector) pair -- one ``LightCurve.process()`` of the reference
(lightcurve/lightcurve.py:167-200).  A *batch* is a set of units processed by one sequence of
launches; each unit owns a slice (``*_off`` .. ``*_off + n``) of the flat event arrays.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import sys
import threading
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from . import _capi
from ._capi import COUNTS_DTYPE, UNIT_DTYPE, WORK_DTYPE, RngDesc

DEFAULT_SEED = 0x00C0560612B
# photons per CTA work item of the energy kernels: amortises the 44 KB table prologue (measured on the bright
# burst: 8192 -> 3.94 ms, 16384 -> 3.80, 32768 -> 3.82, 65536 -> 4.07); CGRB_ENERGY_CHUNK overrides for tuning
ENERGY_CHUNK = int(os.environ.get("CGRB_ENERGY_CHUNK", 16384))
BACKGROUND_SIDE_STREAM = os.environ.get("CGRB_BACKGROUND_SIDE_STREAM", "1") != "0"
SIDE_STREAM_MAX_UNITS = int(os.environ.get("CGRB_SIDE_STREAM_MAX_UNITS", 64))     # a few bursts x 14 detectors
# thinning: units with at most this many candidate slots are cut into short segments (a segment is processed by one
# CTA, 512 candidates at a time: 16 serial rounds for a full 8192-candidate segment, which is what a single small
# burst waits for); both numbers are even multiples of the kernels' 512-candidate tile
SMALL_UNIT_SEGMENTS = (int(os.environ.get("CGRB_SMALL_UNIT_SLOTS", 65536)), int(os.environ.get("CGRB_SMALL_SEG", 1024)))
DIGITIZE_CHUNK = 65536       # photons per CTA of the stand-alone digitize kernel (amortises staging the 18 KB guide)
DEFAULT_MAX_TRIALS = 1 << 22


def _torch():
    import torch

    return torch


def _now():
    import time

    return time.perf_counter()


# ----------------------------------------------------------------------------------------------
# uniform sources
# ----------------------------------------------------------------------------------------------
@dataclass
class Philox:
    """Counter-based Philox4x32-10 stream keyed by (seed, unit stream id, stage, index)."""

    seed: int = DEFAULT_SEED


@dataclass
class Injected:
    """Replay of explicit uniform streams (parity mode).  All arrays are host numpy arrays;
    ``thin``/``bkg_time`` hold (u1, u2) pairs per candidate, ``digitize``/``bkg_chan`` one uniform
    per event, ``energy`` is consumed from per-photon start positions ``energy_off``.

    Each of ``thin``, ``digitize``, ``bkg_time``, ``bkg_chan`` is a list with one array per unit.
    ``energy`` is one flat array and ``energy_off`` a list (one int64 array per unit) of absolute
    start positions in it, one per source photon.
    """

    thin: Optional[list] = None
    energy: Optional[np.ndarray] = None
    energy_off: Optional[list] = None
    digitize: Optional[list] = None
    bkg_time: Optional[list] = None
    bkg_chan: Optional[list] = None


def make_units(n):
    u = np.zeros(n, dtype=UNIT_DTYPE)
    u["ea_scale"] = 1.0
    return u


def fill_gamma_constants(units):
    """Per-GRB constants of a = 2 + alpha used by the device igamc (CPython's math.gamma / math.lgamma, evaluated
    once per distinct alpha: the 14 detectors of a burst share it)."""
    alpha = np.asarray(units["alpha"], dtype="f8")
    uniq, inv = np.unique(alpha, return_inverse=True)
    g = np.full(len(uniq), np.nan)
    lg = np.full(len(uniq), np.nan)
    lg1 = np.full(len(uniq), np.nan)
    for k, al in enumerate(uniq):
        a = 2.0 + float(al)
        if a > 0:
            g[k], lg[k], lg1[k] = math.gamma(a), math.lgamma(a), math.lgamma(1.0 + a)
    units["gamma_a"], units["lgamma_a"], units["lgamma_1pa"] = g[inv], lg[inv], lg1[inv]
    return units


CAPACITY_MAX = 1 << 62


def capacity(rate, span):
    """Candidate capacity: Poisson mean + 8 sigma + slack.  The candidate count of thinning at
    rate ``rate`` over ``span`` seconds is Poisson(rate*span); P(N > mean + 8 sigma) < 1e-15.
    A non-finite or astronomically large mean saturates at CAPACITY_MAX (the engine then refuses the
    unit) instead of wrapping around in the integer cast."""
    with np.errstate(all="ignore"):
        mean = np.asarray(rate, dtype="f8") * span
        mean = np.where(np.isfinite(mean), np.maximum(mean, 0.0), np.inf)
        cap = np.ceil(mean + 8.0 * np.sqrt(mean)) + 64
        cap = np.where(cap < float(CAPACITY_MAX), cap, float(CAPACITY_MAX))
    return cap.astype(np.int64)


def source_work_list(src_cap, chunk, small):
    """The thinning work list with per-unit segment lengths: unit u is cut into segments of ``small[1]`` candidates if
    its capacity ``src_cap[u]`` is at most ``small[0]``, of ``chunk`` candidates otherwise (both even multiples of the
    kernels' 512-candidate tile, at most CGRB_SEG); every unit owns at least one entry.  Same layout as
    ``cgrb_build_work(which=0)``: entries ordered by unit, then by start."""
    cap = np.asarray(src_cap, dtype=np.int64)
    ch = np.where(cap <= small[0], small[1], chunk).astype(np.int64)
    nseg = np.maximum((cap + ch - 1) // ch, 1)
    w = np.zeros(int(nseg.sum()), dtype=_capi.WORK_DTYPE)
    w["unit"] = np.repeat(np.arange(len(cap), dtype=np.int32), nseg)
    first = np.concatenate(([0], np.cumsum(nseg)[:-1]))
    seg = np.arange(len(w), dtype=np.int64) - np.repeat(first, nseg)
    w["seg"] = seg
    w["start"] = seg * np.repeat(ch, nseg)
    return w


def _unit_work_begin(work, n_units):
    return np.searchsorted(work["unit"], np.arange(n_units + 1), side="left").astype(np.int64)


class Batch(object):
    """Device-resident state of one batch of units.  All event arrays are flat torch tensors on
    the engine's device; ``units``/``counts`` are the host copies of the per-unit structs."""

    def __init__(self, engine, units, dtab, interp_extrap):
        self.engine = engine
        self.units = units
        self.n_units = len(units)
        self.dtab = dtab
        self.interp_extrap = interp_extrap
        self.counts = None          # host copy after fetch_counts()
        self.d_units = None
        self.d_counts = None
        self.t = {}                 # name -> torch tensor
        self.work = {}              # which -> (host work, device work, device unit_work_begin)
        self._keep = []             # keeps injected uniform tensors alive

    # -- per-unit views (host numpy) ------------------------------------------------------------
    def view(self, name, ui, kind):
        """Host copy of unit ``ui``'s slice of event array ``name``; kind in src|bkg|tot|kept."""
        u, c = self.units[ui], self.counts[ui]
        off, n = {
            "src": (u["src_off"], c["n_src"]),
            "bkg": (u["bkg_off"], c["n_bkg"]),
            "tot": (u["tot_off"], c["n_total"]),
            "kept": (u["tot_off"], c["n_kept"]),
        }[kind]
        return self.t[name][int(off):int(off) + int(n)].cpu().numpy()


class Engine(object):
    """One engine per process and device."""

    def __init__(self, device=None, seed=DEFAULT_SEED, interp_extrap="linear", energy_sampler="envelope", squeeze=True,
                 deadtime_mode="reference", deadtime_tau=2.6e-6, deadtime_tau_overflow=10e-6, thinning="majorant"):
        torch = _torch()
        self.lib = _capi.load()
        if not torch.cuda.is_available():
            raise _capi.EngineUnavailable(
                "no CUDA device: the cosmogrb_b200 photon-event engine has no CPU fallback"
            )
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = torch.device(device)
        info = (C.c_int * 4)()
        _capi.check(self.lib.cgrb_device_info(self.device.index or 0, info), "cgrb_device_info")
        self.sm_count, self.cc = info[0], (info[1], info[2])
        self.seed = int(seed)
        self.interp_extrap = interp_extrap
        self.energy_sampler = energy_sampler
        self.squeeze = bool(squeeze)
        # candidate proposal of the source thinning under Philox: "majorant" (piecewise-constant bound built
        # from the squeeze table: same point process, far fewer candidates) | "literal" (the reference's single
        # fmax over the whole window).  Injected uniforms always replay the literal loop.
        if thinning not in ("majorant", "literal"):
            raise ValueError("thinning must be 'majorant' or 'literal'")
        self.thinning = thinning
        # GBM dead-time model: "reference" = literal _gbm_dead_time (gbm_lightcurve.py:80-115), "physical" =
        # non-paralyzable window after EVERY kept event
        self.set_deadtime(deadtime_mode, deadtime_tau, deadtime_tau_overflow)
        self.squeeze_ratio, self.squeeze_min_candidates, self.squeeze_max_nodes = 8, 2048, 1 << 14
        self.energy_row_spacing, self.energy_candidates_per_row = 0.02, 256
        # hard ceiling on the thinning candidates of ONE unit (fmax x duration): beyond it the unit's
        # buffers cannot be sized sensibly (the reference's own loop would run for years); raised loudly
        self.max_candidates_per_unit = 1 << 31
        self._call_counter = 0
        self.host_times = {}        # host seconds spent in allocation / uploads / work-list building (diagnostics)
        self._pool = None
        self._dtab_cache = {}
        self.launches = 0           # kernels launched by this engine (bench.py's gpu_launches)

    def set_deadtime(self, mode="reference", tau=2.6e-6, tau_overflow=10e-6):
        if mode not in ("reference", "physical"):
            raise ValueError(f"deadtime mode must be 'reference' or 'physical', got {mode!r}")
        self.deadtime_mode, self.deadtime_tau, self.deadtime_tau_overflow = mode, float(tau), float(tau_overflow)

    def _deadtime_model(self, mode=None):
        mode = self.deadtime_mode if mode is None else mode
        if mode not in ("reference", "physical"):
            raise ValueError(f"deadtime mode must be 'reference' or 'physical', got {mode!r}")
        return _capi.DeadtimeModel(_capi.DEADTIME_PHYSICAL if mode == "physical" else _capi.DEADTIME_REFERENCE, 0,
                                   self.deadtime_tau, self.deadtime_tau_overflow)

    # -- small helpers --------------------------------------------------------------------------
    @property
    def stream(self):
        return _torch().cuda.current_stream(self.device).cuda_stream

    def copy_stream(self):
        """side stream for device->host copies that overlap the next batch's kernels"""
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = _torch().cuda.Stream(device=self.device)
        return self._copy_stream

    def next_seed(self):
        """A fresh Philox seed per sampler call (the reference reseeds from OS entropy before every
        call -- sampler/source.py:129; here the sequence is reproducible for a fixed engine seed)."""
        self._call_counter += 1
        return (self.seed + 0x9E3779B97F4A7C15 * self._call_counter) & 0xFFFFFFFFFFFFFFFF

    def to_device(self, arr):
        """Host array -> device tensor.  The copy runs on a dedicated upload stream so that it is not ordered
        behind kernels already queued on the compute stream (a pageable host->device copy blocks the host
        until the stream reaches it: on the compute stream that would serialise host preparation of the next
        batch with the kernels of the current one); the compute stream then waits for the upload."""
        torch = _torch()
        t0 = _now()
        a = np.ascontiguousarray(arr)
        if a.dtype.fields is not None:
            a = a.view(np.uint8)
        if getattr(self, "_upload_stream", None) is None:
            self._upload_stream = torch.cuda.Stream(device=self.device)
        cur = torch.cuda.current_stream(self.device)
        with torch.cuda.stream(self._upload_stream):
            d = torch.from_numpy(a).to(self.device, non_blocking=False)
        cur.wait_stream(self._upload_stream)
        d.record_stream(cur)
        self.host_times["to_device"] = self.host_times.get("to_device", 0.0) + (_now() - t0)
        return d

    def empty(self, n, dtype):
        """Uninitialised device buffer of at least ``n`` elements.  Inside ``use_pool`` the buffers of a batch
        are recycled from the pool slot in call order (a batch issues the same sequence of requests every
        time), so a long run of batches does not go through the allocator at all."""
        torch = _torch()
        n = max(int(n), 1)
        pool = self._pool
        if pool is None:
            t0 = _now()
            t = torch.empty(n, dtype=dtype, device=self.device)
            self.host_times["alloc"] = self.host_times.get("alloc", 0.0) + (_now() - t0)
            return t
        i = pool["i"]
        pool["i"] = i + 1
        bufs = pool["bufs"]
        if i < len(bufs) and bufs[i].dtype == dtype and bufs[i].numel() >= n:
            return bufs[i]
        t0 = _now()
        if os.environ.get("CGRB_POOL_LOG"):
            had = (str(bufs[i].dtype), bufs[i].numel()) if i < len(bufs) else None
            print(f"[pool miss] slot {i}: want {dtype} x {n}, had {had}", file=sys.stderr, flush=True)
        # head-room: batches vary in size.  Allocated in the DEFAULT stream's context whatever stream the batch runs
        # on: torch's caching allocator keeps one block pool per stream, and with the two batch streams of a population
        # run the outgrown buffers of one stream were no use to the other -- device memory filled up with cached
        # blocks and every further growth went through the allocator's free-everything-and-retry path (0.6 s per
        # 100,000 bursts).  Ordering is the pool's own: a slot is reused only after its previous batch was read back.
        with torch.cuda.stream(torch.cuda.default_stream(self.device)):
            t = torch.empty(n + n // 4 + 1024, dtype=dtype, device=self.device)
        self.host_times["alloc"] = self.host_times.get("alloc", 0.0) + (_now() - t0)
        if i < len(bufs):
            bufs[i] = t
        else:
            bufs.append(t)
        return t

    def use_pool(self, slot):
        """Route ``empty`` through pool slot ``slot`` (0 / 1: two batches may be in flight), or None to go back
        to plain allocations.  The caller guarantees that the slot's previous batch has been read back."""
        if slot is None:
            self._pool = None
            return
        slots = self.__dict__.setdefault("_pool_slots", {})
        pool = slots.setdefault(slot, {"i": 0, "bufs": []})
        pool["i"] = 0
        self._pool = pool

    def upload_tables(self, tables):
        """tables: list of float64 blocks (Response.device_table) -> one device tensor."""
        key = tuple(id(t) for t in tables)
        hit = self._dtab_cache.get(key)
        if hit is not None and all(a is b for a, b in zip(hit[0], tables)):
            return hit[1]
        block = np.concatenate([np.asarray(t, dtype="f8").reshape(-1) for t in tables])
        assert block.size == len(tables) * _capi.DT_SIZE
        d = self.to_device(block)
        if len(self._dtab_cache) > 64:
            self._dtab_cache.clear()
        self._dtab_cache[key] = (list(tables), d)
        return d

    def _extrap_code(self, interp_extrap):
        return {"linear": 0, "clamp": 1}[interp_extrap or self.interp_extrap]

    def _rng_desc(self, rng, batch, kind):
        """Build the C rng descriptor for one stage (kind in thin|energy|digitize|bkg_time|bkg_chan)."""
        d = RngDesc()
        if isinstance(rng, Philox):
            d.mode = 0
            d.seed = rng.seed & 0xFFFFFFFFFFFFFFFF
            return d
        if not isinstance(rng, Injected):
            raise TypeError("rng must be Philox or Injected")
        d.mode = 1
        d.seed = 0
        if kind == "energy":
            if rng.energy is None or rng.energy_off is None:
                raise ValueError("Injected.energy / energy_off required")
            u = self.to_device(np.asarray(rng.energy, dtype="f8"))
            total = int(batch.units["src_off"][-1] + batch.units["src_cap"][-1] + 1)
            off = np.zeros(total, dtype=np.int64)
            for ui in range(batch.n_units):
                o = np.asarray(rng.energy_off[ui], dtype=np.int64)
                s = int(batch.units["src_off"][ui])
                off[s:s + len(o)] = o
            d_off = self.to_device(off)
            d_len = self.to_device(np.array([len(rng.energy)], dtype=np.int64))
        else:
            streams = getattr(rng, kind)
            if streams is None:
                raise ValueError(f"Injected.{kind} required")
            if len(streams) != batch.n_units:
                raise ValueError(f"Injected.{kind} needs one stream per unit")
            lens = np.array([len(s) for s in streams], dtype=np.int64)
            offs = np.concatenate(([0], np.cumsum(lens)[:-1])).astype(np.int64)
            flat = np.concatenate([np.asarray(s, dtype="f8") for s in streams]) if lens.sum() else np.zeros(1)
            u = self.to_device(flat)
            d_off = self.to_device(offs)
            d_len = self.to_device(lens)
        batch._keep += [u, d_off, d_len]
        d.d_u, d.d_off, d.d_len = u.data_ptr(), d_off.data_ptr(), d_len.data_ptr()
        return d

    # -- batch construction ---------------------------------------------------------------------
    def fmax(self, units, tables, interp_extrap=None):
        """Source._get_energy_integrated_max (source.py:88-110) of every unit: one launch, fmax returned."""
        units = np.array(units, dtype=UNIT_DTYPE, copy=True)
        fill_gamma_constants(units)
        dtab = self.upload_tables(tables) if isinstance(tables, (list, tuple)) else tables
        d_units = self.to_device(units)
        _capi.check(self.lib.cgrb_fmax(dtab.data_ptr(), d_units.data_ptr(), len(units),
                                       self._extrap_code(interp_extrap), self.stream), "cgrb_fmax")
        self.launches += 1
        # read back the fmax column only (8 of the record's 224 bytes)
        torch = _torch()
        col = UNIT_DTYPE.fields["fmax"][1] // 8
        return d_units.view(torch.uint8).view(torch.float64).view(len(units), UNIT_DTYPE.itemsize // 8)[:, col].cpu().numpy()

    def new_batch(self, units, tables, interp_extrap=None, compute_fmax=True):
        """Upload units and detector tables, compute fmax on the device (unless given), size the
        per-unit capacities and offsets."""
        units = np.array(units, dtype=UNIT_DTYPE, copy=True)
        fill_gamma_constants(units)
        dtab = self.upload_tables(tables) if isinstance(tables, (list, tuple)) else tables
        b = Batch(self, units, dtab, interp_extrap or self.interp_extrap)
        if compute_fmax:
            units["fmax"] = self.fmax(units, dtab, b.interp_extrap)
        self.finalize_layout(b)
        return b

    def finalize_layout(self, b):
        units = b.units
        no_src = (units["flags"] & _capi.UNIT_NO_SOURCE) != 0
        no_bkg = (units["flags"] & _capi.UNIT_NO_BACKGROUND) != 0
        src_cap = capacity(units["fmax"], units["src_tstop"] - units["src_tstart"])
        bkg_cap = capacity(units["bkg_rate"], units["bkg_tstop"] - units["bkg_tstart"])
        src_cap[no_src] = 0
        bkg_cap[no_bkg] = 0
        worst = int(np.argmax(np.maximum(src_cap, bkg_cap))) if len(units) else 0
        if len(units) and max(int(src_cap[worst]), int(bkg_cap[worst])) > self.max_candidates_per_unit:
            raise _capi.EngineError(
                f"unit {worst} needs {float(max(src_cap[worst], bkg_cap[worst])):.3g} thinning candidates "
                f"(fmax = {float(units['fmax'][worst]):.3g} /s over {float(units['src_tstop'][worst] - units['src_tstart'][worst]):.3g} s), "
                f"more than max_candidates_per_unit = {self.max_candidates_per_unit}")
        keep_src = units["src_cap"] > 0
        keep_bkg = units["bkg_cap"] > 0
        units["src_cap"] = np.where(keep_src, units["src_cap"], src_cap)   # caller may pre-size
        units["bkg_cap"] = np.where(keep_bkg, units["bkg_cap"], bkg_cap)
        units["src_off"] = np.concatenate(([0], np.cumsum(units["src_cap"] + 1)[:-1]))
        units["bkg_off"] = np.concatenate(([0], np.cumsum(units["bkg_cap"] + 1)[:-1]))
        units["tot_off"] = np.concatenate(([0], np.cumsum(units["src_cap"] + units["bkg_cap"] + 2)[:-1]))
        # squeeze table of the thinning acceptance probability: ~1 node per SQUEEZE_RATIO candidates,
        # none for short units where building it would cost as much as it saves
        tab_n = np.where(units["src_cap"] >= self.squeeze_min_candidates,
                         np.clip(units["src_cap"] // self.squeeze_ratio, 64, self.squeeze_max_nodes), 0)
        if not self.squeeze:
            tab_n[:] = 0
        units["tab_n"] = tab_n
        units["tab_off"] = np.concatenate(([0], np.cumsum(tab_n)[:-1]))
        b.n_tab_nodes = int(tab_n.sum())
        # rows of the energy proposal table: a grid of w = (1+z)/ec(t) over the burst, ~2 % apart for long
        # units, fewer for short ones (a row costs 256 exps to build, about as much as 100 photons)
        with np.errstate(all="ignore"):
            span = np.where(units["kind"] == 0,
                            np.abs(np.log1p(units["src_tstop"] / units["ep_tau"])
                                   - np.log1p(units["src_tstart"] / units["ep_tau"])), 0.0)
        span = np.nan_to_num(span, nan=0.0, posinf=0.0)
        need = np.ceil(span / self.energy_row_spacing).astype(np.int64) + 1
        nw = np.clip(np.minimum(need, units["src_cap"] // self.energy_candidates_per_row), 1, _capi.ENV_MAXW)
        units["etab_nw"] = nw
        units["etab_off"] = np.concatenate(([0], np.cumsum(nw)[:-1]))
        b.n_etab_rows = int(nw.sum())
        b.n_src_slots = int(np.sum(units["src_cap"] + 1))
        b.n_bkg_slots = int(np.sum(units["bkg_cap"] + 1))
        b.n_tot_slots = b.n_src_slots + b.n_bkg_slots
        b.d_units = self.to_device(units)
        torch = _torch()
        b.d_counts = torch.zeros(b.n_units * COUNTS_DTYPE.itemsize, dtype=torch.uint8, device=self.device)

    def _work(self, b, which, chunk, small=None):
        """Work list of a batch: (unit, chunk) entries, ``chunk`` elements each.  ``small = (limit, chunk_small)``:
        units whose capacity is at most ``limit`` are cut into chunks of ``chunk_small`` instead -- a function of
        the unit alone, so results do not depend on what else is in the batch."""
        key = (which, chunk, small)
        if key not in b.work:
            t0 = _now()
            if small is None:
                w = _capi.build_work(b.units, which, chunk)
            else:
                assert which == 0
                w = source_work_list(b.units["src_cap"], chunk, small)
            wb = _unit_work_begin(w, b.n_units)
            t1 = _now()
            b.work[key] = (w, self.to_device(w), self.to_device(wb))
            self.host_times["work_build"] = self.host_times.get("work_build", 0.0) + (t1 - t0)
        return b.work[key]

    def set_counts(self, b, **cols):
        """Overwrite per-unit counters from the host (stage-level entry points)."""
        c = np.zeros(b.n_units, dtype=COUNTS_DTYPE)
        if b.counts is not None:
            c[:] = b.counts
        for k, v in cols.items():
            c[k] = v
        b.counts = c
        b.d_counts = self.to_device(c)

    def reset_counts(self, b):
        _capi.check(self.lib.cgrb_counts_reset(b.d_counts.data_ptr(), b.n_units, self.stream), "cgrb_counts_reset")
        self.launches += 1
        b.counts = None

    # the six event lists of LightCurveStorage (light_curve_storage.py:16-80): name, slice kind
    HOST_ARRAYS = (("times_out", "kept"), ("pha_out", "kept"), ("times_src", "src"), ("pha_src", "src"),
                   ("times_bkg", "bkg"), ("pha_bkg", "bkg"))

    def to_host(self, b, pinned=None, extra=()):
        """Copy every unit's event lists to (pinned) host memory, compacted: returns a dict with
        one host array per list plus per-unit offsets.  ``pinned`` is a cache of pinned staging
        buffers that is reused (and grown) across calls."""
        torch = _torch()
        self.fetch_counts(b)
        self.check_status(b)
        out = self._enqueue_event_copies(b, pinned, extra)
        torch.cuda.current_stream(self.device).synchronize()
        return out

    def to_host_async(self, b, ready_event, copy_stream, extra=()):
        """``to_host`` on a side stream: waits for ``ready_event`` (recorded after the batch's kernels),
        reads the unit counters, then enqueues the event-list copies on ``copy_stream`` and returns
        without waiting for them -- the caller keeps launching the next batch on the compute stream and
        synchronises ``copy_stream`` before touching the arrays."""
        torch = _torch()
        # The counters size the copies, so the host has to see them first.  They come back on a stream of their own,
        # ordered behind the batch's kernels only: on the copy stream they would queue behind the PREVIOUS batch's
        # event lists (tens of ms), the host would enqueue this batch's copies only after those had finished, and the
        # copy engine would idle for the host's enqueue latency once per batch.
        if getattr(self, "_count_stream", None) is None:
            self._count_stream = torch.cuda.Stream(device=self.device)
        cnt = self._count_stream
        with torch.cuda.stream(cnt):
            cnt.wait_event(ready_event)
            cbuf = torch.empty(b.d_counts.numel(), dtype=torch.uint8, pin_memory=True)
            cbuf.copy_(b.d_counts, non_blocking=True)
            cnt.synchronize()
        b.counts = cbuf.numpy().view(COUNTS_DTYPE).copy()
        self.check_status(b)
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ready_event)
            return self._enqueue_event_copies(b, None, extra)

    def fetch_results_after(self, b, ready_event, want_trigger=False):
        """Counters (and trigger results) of a batch read back on the copy stream once ``ready_event`` (recorded
        after the batch's last kernel) has fired -- NOT stream-ordered behind whatever the compute stream has
        been given since, so a later batch keeps the GPU busy while this one is read."""
        torch = _torch()
        cs = self.copy_stream()
        with torch.cuda.stream(cs):
            cs.wait_event(ready_event)
            cbuf = torch.empty(b.d_counts.numel(), dtype=torch.uint8, pin_memory=True)
            cbuf.copy_(b.d_counts, non_blocking=True)
            tbuf = None
            if want_trigger:
                nbytes = b.n_units * _capi.TRIGGER_DTYPE.itemsize
                tbuf = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
                tbuf.copy_(b.t["trigger"][:nbytes], non_blocking=True)
            cs.synchronize()
        b.counts = cbuf.numpy().view(COUNTS_DTYPE).copy()
        self.check_status(b)
        if not want_trigger:
            return b.counts, None
        r = tbuf.numpy().view(_capi.TRIGGER_DTYPE).copy()
        if np.any(r["razor"] & 2):
            raise _capi.EngineError("cgrb_gbm_trigger: background_duration + pre_window + longest time scale "
                                    "exceed the kernel's staged window span (1832 bins of 16 ms)")
        return b.counts, r

    def _enqueue_event_copies(self, b, pinned, extra):
        torch = _torch()
        c = b.counts
        # pinned=None: fresh pinned buffers per call (torch's caching host allocator recycles them), so
        # the returned arrays stay valid for as long as the caller keeps them
        n_of = {"kept": c["n_kept"], "src": c["n_src"], "bkg": c["n_bkg"], "tot": c["n_total"]}
        off_of = {"kept": b.units["tot_off"], "src": b.units["src_off"], "bkg": b.units["bkg_off"],
                  "tot": b.units["tot_off"]}
        out = {"counts": c, "d2h_bytes": int(c.nbytes), "events": int(c["n_src"].sum() + c["n_bkg"].sum())}
        for name, kind in tuple(self.HOST_ARRAYS) + tuple(extra):
            src = b.t[name]
            n = n_of[kind].astype(np.int64)
            starts = np.concatenate(([0], np.cumsum(n)))
            total = int(starts[-1])
            buf = pinned.get(name) if pinned is not None else None
            if buf is None or buf.numel() < total or buf.dtype != src.dtype:
                buf = torch.empty(max(int(total * 1.1) + 1024, 1) if pinned is not None else max(total, 1),
                                  dtype=src.dtype, pin_memory=True)
                if pinned is not None:
                    pinned[name] = buf
            for ui in range(b.n_units):
                k = int(n[ui])
                if k:
                    o = int(off_of[kind][ui])
                    buf[int(starts[ui]):int(starts[ui]) + k].copy_(src[o:o + k], non_blocking=True)
            out[name] = buf[:total]
            out[name + "_offsets"] = starts
            out["d2h_bytes"] += total * src.element_size()
        return out

    def fetch_counts(self, b):
        b.counts = b.d_counts.cpu().numpy().view(COUNTS_DTYPE).copy()
        return b.counts

    # -- stages ---------------------------------------------------------------------------------
    def sample_events(self, b, rng, want_candidates=False):
        torch = _torch()
        w, dw, dwb = self._work(b, 0, _capi.SEG, small=SMALL_UNIT_SEGMENTS)
        nw = len(w)
        t = b.t
        t["segsum"] = self.empty(nw, torch.float64)
        t["segstart"] = self.empty(nw, torch.float64)
        t["segcount"] = self.empty(nw, torch.int32)
        t["stage"] = self.empty(b.n_src_slots, torch.float64)
        t["times_src"] = self.empty(b.n_src_slots, torch.float64)
        cand_t = cand_a = None
        if want_candidates:
            t["cand_times"] = cand_t = self.empty(b.n_src_slots, torch.float64)
            t["cand_accept"] = cand_a = torch.zeros(max(b.n_src_slots, 1), dtype=torch.uint8, device=self.device)
        desc = self._rng_desc(rng, b, "thin")
        thin_tab = thin_cum = None
        if b.n_tab_nodes > 0:
            # rebuilt on every call (it is part of the sampling work, ~1/8 of a lambda per candidate)
            w4, dw4, _ = self._work(b, 4, 256)
            t["thin_tab"] = thin_tab = self.empty(2 * b.n_tab_nodes, torch.float64)
            t.pop("thin_cum", None)
            if self.thinning == "majorant" and not isinstance(rng, Injected):
                t["thin_cum"] = thin_cum = self.empty(self.lib.cgrb_thin_cum_bytes(b.n_tab_nodes), torch.uint8)
            _capi.check(self.lib.cgrb_thin_tables(b.dtab.data_ptr(), b.d_units.data_ptr(), b.n_units,
                                                  dw4.data_ptr(), len(w4), thin_tab.data_ptr(),
                                                  thin_cum.data_ptr() if thin_cum is not None else None,
                                                  b.n_tab_nodes, self.stream),
                        "cgrb_thin_tables")
            self.launches += 2 + 2 * (thin_cum is not None)
        _capi.check(
            self.lib.cgrb_sample_events(
                b.dtab.data_ptr(), b.d_units.data_ptr(), b.n_units, dw.data_ptr(), nw, dwb.data_ptr(),
                self._extrap_code(b.interp_extrap), C.byref(desc), t["segsum"].data_ptr(),
                t["segstart"].data_ptr(), t["segcount"].data_ptr(), t["stage"].data_ptr(),
                t["times_src"].data_ptr(), cand_t.data_ptr() if want_candidates else None,
                cand_a.data_ptr() if want_candidates else None,
                thin_tab.data_ptr() if thin_tab is not None else None,
                thin_cum.data_ptr() if thin_cum is not None else None, b.n_tab_nodes,
                b.d_counts.data_ptr(), self.stream),
            "cgrb_sample_events")
        self.launches += 4

    def majorant_totals(self, b):
        """Expected candidates per unit under the majorant proposal (Lambda_T = fmax x integral of the
        piecewise bound), NaN for units without a table; None if the last ``sample_events`` used the literal
        proposal."""
        if "thin_cum" not in b.t:
            return None
        torch = _torch()
        L = b.t["thin_cum"][:8 * b.n_tab_nodes].view(torch.float64)
        out = np.full(b.n_units, np.nan)
        has = b.units["tab_n"] >= 4
        last = (b.units["tab_off"] + b.units["tab_n"] - 1)[has]
        if has.any():
            out[has] = L[torch.as_tensor(last.astype(np.int64), device=self.device)].cpu().numpy()
        return out

    def energy_tables(self, b):
        """The per-unit envelope tables of the tabulated-envelope energy sampler (they depend on the units only), on the
        current stream; ``sample_energy(..., tables_ready=True)`` then samples with them (the caller orders the two)."""
        torch = _torch()
        t = b.t
        if "etab" not in t:
            t["etab"] = self.empty(self.lib.cgrb_energy_workspace_bytes(b.n_units, b.n_etab_rows), torch.uint8)
        _capi.check(
            self.lib.cgrb_energy_envelope_tables(b.dtab.data_ptr(), b.d_units.data_ptr(), b.n_units,
                                                 self._extrap_code(b.interp_extrap), t["etab"].data_ptr(), b.n_etab_rows,
                                                 self.stream),
            "cgrb_energy_envelope_tables")
        self.launches += 2

    def sample_energy(self, b, rng, max_trials=DEFAULT_MAX_TRIALS, want_trials=False, sampler=None,
                      fold=False, want_energy=True, tables_ready=False):
        """``sampler``: "literal" (the reference's rejection loop, trial for trial; always used under
        injected uniforms) or "envelope" (same output distribution from a tabulated envelope, Philox
        only; the default under Philox).  ``fold=True`` (envelope sampler only) also folds every photon
        through the response in the same kernel (``pha_src``, identical to ``digitize_batch`` with the
        same Philox seed); the energies themselves are then written only if ``want_energy``.
        Returns True if the channels were produced."""
        torch = _torch()
        sampler = sampler or self.energy_sampler
        if isinstance(rng, Injected):
            sampler = "literal"
        if sampler not in ("literal", "envelope"):
            raise ValueError("sampler must be 'literal' or 'envelope'")
        fold = bool(fold) and sampler == "envelope"
        w, dw, _ = self._work(b, 3, ENERGY_CHUNK)
        t = b.t
        en = None
        if want_energy or not fold:
            t["energy_src"] = en = self.empty(b.n_src_slots, torch.float64)
        else:
            t.pop("energy_src", None)
        tr = None
        if want_trials:
            t["trials"] = tr = torch.zeros(max(b.n_src_slots, 1), dtype=torch.int32, device=self.device)
        desc = self._rng_desc(rng, b, "energy")
        if sampler == "envelope":
            if "etab" not in t:
                t["etab"] = self.empty(self.lib.cgrb_energy_workspace_bytes(b.n_units, b.n_etab_rows), torch.uint8)
            pha = None
            if fold:
                t["pha_src"] = pha = self.empty(b.n_src_slots, torch.uint8)
            fn = self.lib.cgrb_sample_energy_envelope_prepared if tables_ready else self.lib.cgrb_sample_energy_envelope
            _capi.check(
                fn(b.dtab.data_ptr(), b.d_units.data_ptr(), b.n_units, dw.data_ptr(), len(w),
                   self._extrap_code(b.interp_extrap), C.byref(desc), int(max_trials), t["etab"].data_ptr(),
                   b.n_etab_rows, t["times_src"].data_ptr(), en.data_ptr() if en is not None else None,
                   pha.data_ptr() if pha is not None else None,
                   tr.data_ptr() if want_trials else None, b.d_counts.data_ptr(), self.stream),
                "cgrb_sample_energy_envelope")
            self.launches += 1 if tables_ready else 3
            return fold
        _capi.check(
            self.lib.cgrb_sample_energy(
                b.dtab.data_ptr(), b.d_units.data_ptr(), b.n_units, dw.data_ptr(), len(w),
                self._extrap_code(b.interp_extrap), C.byref(desc), int(max_trials),
                t["times_src"].data_ptr(), t["energy_src"].data_ptr(),
                tr.data_ptr() if want_trials else None, b.d_counts.data_ptr(), self.stream),
            "cgrb_sample_energy")
        self.launches += 1
        return False

    def digitize_batch(self, b, rng, trials=None):
        """``Response.digitize`` of every photon.  ``trials`` (the envelope sampler's per-photon trial counts, Philox
        only) replays the fold fused into ``sample_energy(fold=True)``: same channel uniforms, same channels."""
        torch = _torch()
        w, dw, _ = self._work(b, 3, DIGITIZE_CHUNK)
        t = b.t
        t["pha_src"] = self.empty(b.n_src_slots, torch.uint8)
        desc = self._rng_desc(rng, b, "digitize")
        _capi.check(
            self.lib.cgrb_digitize(
                b.dtab.data_ptr(), b.d_units.data_ptr(), b.n_units, dw.data_ptr(), len(w), C.byref(desc),
                t["energy_src"].data_ptr(), trials.data_ptr() if trials is not None else None,
                t["pha_src"].data_ptr(), b.d_counts.data_ptr(), self.stream),
            "cgrb_digitize")
        self.launches += 1

    def sample_background(self, b, rng):
        torch = _torch()
        u = b.units
        has = (u["flags"] & _capi.UNIT_NO_BACKGROUND) == 0
        n_tiles = int(np.maximum((u["bkg_cap"][has] + _capi.BKG_TILE - 1) // _capi.BKG_TILE, 1).sum())
        t = b.t
        t["times_bkg"] = self.empty(b.n_bkg_slots, torch.float64)
        t["pha_bkg"] = self.empty(b.n_bkg_slots, torch.uint8)
        if n_tiles == 0:
            return
        t["bkg_ws"] = self.empty(self.lib.cgrb_sample_background_workspace_bytes(b.n_units, n_tiles), torch.uint8)
        d_time = self._rng_desc(rng, b, "bkg_time")
        d_chan = self._rng_desc(rng, b, "bkg_chan") if isinstance(rng, Injected) else d_time
        _capi.check(
            self.lib.cgrb_sample_background(
                b.dtab.data_ptr(), b.d_units.data_ptr(), b.n_units, n_tiles, C.byref(d_time), C.byref(d_chan),
                t["bkg_ws"].data_ptr(), t["times_bkg"].data_ptr(), t["pha_bkg"].data_ptr(), b.d_counts.data_ptr(),
                self.stream),
            "cgrb_sample_background")
        self.launches += 3

    def combine(self, b):
        torch = _torch()
        w, dw, _ = self._work(b, 2, _capi.MERGE_TILE)
        t = b.t
        t["times_tot"] = self.empty(b.n_tot_slots, torch.float64)
        t["pha_tot"] = self.empty(b.n_tot_slots, torch.uint8)
        _capi.check(
            self.lib.cgrb_combine(
                b.d_units.data_ptr(), b.n_units, dw.data_ptr(), len(w), t["times_bkg"].data_ptr(),
                t["pha_bkg"].data_ptr(), t["times_src"].data_ptr(), t["pha_src"].data_ptr(),
                t["times_tot"].data_ptr(), t["pha_tot"].data_ptr(), b.d_counts.data_ptr(), self.stream),
            "cgrb_combine")
        self.launches += 1

    def dead_time(self, b, want_selection=False, mode=None):
        torch = _torch()
        model = self._deadtime_model(mode)
        w, dw, dwb = self._work(b, 2, _capi.DT_TILE)
        t = b.t
        t["tile_count"] = self.empty(len(w), torch.int32)
        t["times_out"] = self.empty(b.n_tot_slots, torch.float64)
        t["pha_out"] = self.empty(b.n_tot_slots, torch.uint8)
        sel = None
        if want_selection:
            t["selection"] = sel = torch.zeros(max(b.n_tot_slots, 1), dtype=torch.uint8, device=self.device)
        _capi.check(
            self.lib.cgrb_gbm_dead_time(
                b.d_units.data_ptr(), b.n_units, dw.data_ptr(), len(w), dwb.data_ptr(),
                t["times_tot"].data_ptr(), t["pha_tot"].data_ptr(), t["tile_count"].data_ptr(),
                t["times_out"].data_ptr(), t["pha_out"].data_ptr(),
                sel.data_ptr() if want_selection else None, b.d_counts.data_ptr(), C.byref(model), self.stream),
            "cgrb_gbm_dead_time")
        self.launches += 2

    def merge_dead_time(self, b, mode=None):
        """combine + dead time in one pass: only the survivors reach HBM (``times_out``/``pha_out``);
        bit-identical to ``combine`` followed by ``dead_time``."""
        torch = _torch()
        model = self._deadtime_model(mode)
        w, dw, dwb = self._work(b, 2, _capi.MD_TILE)
        t = b.t
        t["tile_state"] = self.empty(self.lib.cgrb_merge_dead_time_workspace_bytes(len(w)), torch.uint8)
        t["times_out"] = self.empty(b.n_tot_slots, torch.float64)
        t["pha_out"] = self.empty(b.n_tot_slots, torch.uint8)
        _capi.check(
            self.lib.cgrb_merge_dead_time(
                b.d_units.data_ptr(), b.n_units, dw.data_ptr(), len(w), dwb.data_ptr(),
                t["times_bkg"].data_ptr(), t["pha_bkg"].data_ptr(), t["times_src"].data_ptr(),
                t["pha_src"].data_ptr(), t["tile_state"].data_ptr(), t["times_out"].data_ptr(),
                t["pha_out"].data_ptr(), b.d_counts.data_ptr(), C.byref(model), self.stream),
            "cgrb_merge_dead_time")
        self.launches += 2

    def gbm_trigger(self, b, chan_bounds, threshold=4.5, background_duration=17.0, pre_window=4.0,
                    base_timescale=0.016, fetch=True):
        """GBMLightCurveAnalyzer for every unit of a processed batch, on the device (events stay in HBM).
        ``chan_bounds``: int32 [n_detector_tables, 8] (gbm_lightcurve_analyzer.channel_bounds).  Returns the
        per-unit results as a numpy array of ``_capi.TRIGGER_DTYPE`` (``fetch=False``: only enqueue; read
        them later with ``fetch_trigger``)."""
        torch = _torch()
        w, dw, _ = self._work(b, 2, _capi.TRIG_TILE)
        u = b.units
        span = float(np.max(np.maximum(u["bkg_tstop"], u["src_tstop"]) - np.minimum(u["bkg_tstart"], u["src_tstart"])))
        bins_per_unit = (int(math.ceil(span / base_timescale)) + 4 + 3) // 4 * 4     # a multiple of 4: aligned rows
        cb = self.to_device(np.ascontiguousarray(chan_bounds, dtype=np.int32).reshape(-1, 8))
        ws = self.empty(self.lib.cgrb_gbm_trigger_workspace_bytes(b.n_units, bins_per_unit), torch.uint8)
        out = self.empty(b.n_units * _capi.TRIGGER_DTYPE.itemsize, torch.uint8)
        par = _capi.TriggerPar(threshold, base_timescale, background_duration, pre_window)
        _capi.check(
            self.lib.cgrb_gbm_trigger(
                b.d_units.data_ptr(), b.n_units, dw.data_ptr(), len(w), cb.data_ptr(), C.byref(par),
                b.t["times_out"].data_ptr(), b.t["pha_out"].data_ptr(), b.d_counts.data_ptr(), bins_per_unit,
                ws.data_ptr(), out.data_ptr(), self.stream),
            "cgrb_gbm_trigger")
        self.launches += 2
        b.t["trigger"] = out
        b._keep.append(cb)
        return self.fetch_trigger(b) if fetch else None

    def fetch_trigger(self, b):
        n = b.n_units * _capi.TRIGGER_DTYPE.itemsize
        r = b.t["trigger"][:n].cpu().numpy().view(_capi.TRIGGER_DTYPE).copy()
        if np.any(r["razor"] & 2):
            raise _capi.EngineError("cgrb_gbm_trigger: background_duration + pre_window + longest time scale "
                                    "exceed the kernel's staged window span (1832 bins of 16 ms)")
        return r

    def analyze_light_curve(self, times, pha, chan_bounds, n_counts_source, **par):
        """The trigger analysis of ONE given light curve (host arrays: the time-sorted total list)."""
        torch = _torch()
        times = np.ascontiguousarray(times, dtype="f8")
        n = len(times)
        pha8 = np.ascontiguousarray(np.asarray(pha).astype(np.int64).clip(0, 255).astype(np.uint8))
        u = make_units(1)
        u["src_cap"], u["bkg_cap"] = max(n - 1, 0), 0
        if n:
            u["bkg_tstart"], u["bkg_tstop"] = times[0], times[-1]
            u["src_tstart"], u["src_tstop"] = times[0], times[-1]
        b = Batch(self, u, None, self.interp_extrap)
        b.n_src_slots, b.n_bkg_slots, b.n_tot_slots = n + 1, 1, n + 2
        b.d_units = self.to_device(u)
        self.set_counts(b, n_kept=n, n_total=n, n_src=int(n_counts_source))
        b.t["times_out"] = self.to_device(np.concatenate((times, [0.0, 0.0])))
        b.t["pha_out"] = self.to_device(np.concatenate((pha8, np.zeros(2, np.uint8))))
        return self.gbm_trigger(b, np.asarray(chan_bounds).reshape(1, 8), **par)[0]

    def process(self, b, rng, max_trials=DEFAULT_MAX_TRIALS, diagnostics=False, energy_sampler=None,
                keep_merged=False):
        """LightCurve.process() for every unit of the batch (lightcurve.py:167-200): source times,
        energies, channels; background times, channels; combine; dead time."""
        # The background chain (order -> plan -> one-pass kernel) depends on nothing the source chain produces, so it
        # runs on a side stream next to it and is joined before the merge: a single small burst is a chain of ~20
        # dependent launches, each far too small to fill the GPU, and this takes three of them off the critical path
        # Only for small batches: a population batch of thousands of units fills the GPU kernel by kernel, and there
        # the extra streams cost more than they give (measured: 10,200 -> 6,500-8,600 GRB/s).
        torch = _torch()
        main = torch.cuda.current_stream(self.device)
        side = tables_done = None
        # (and never inside a pooled population run: the pool hands buffers out in call order, which the branches change)
        if BACKGROUND_SIDE_STREAM and isinstance(rng, Philox) and b.n_units <= SIDE_STREAM_MAX_UNITS and self._pool is None:
            sides = self.__dict__.setdefault("_side_streams", {})
            side = sides.get(main.cuda_stream)
            if side is None:
                side = sides[main.cuda_stream] = torch.cuda.Stream(device=self.device)
            side.wait_stream(main)          # the batch's uploads, and whatever used the pooled buffers before
            with torch.cuda.stream(side):
                self.sample_background(b, rng)
                bkg_done = torch.cuda.Event()
                bkg_done.record(side)
            # the energy sampler's per-unit tables depend on the units only: a second branch
            if (energy_sampler or self.energy_sampler) == "envelope":
                side2 = sides.get((main.cuda_stream, 2))
                if side2 is None:
                    side2 = sides[(main.cuda_stream, 2)] = torch.cuda.Stream(device=self.device)
                side2.wait_stream(main)
                with torch.cuda.stream(side2):
                    self.energy_tables(b)
                    tables_done = torch.cuda.Event()
                    tables_done.record(side2)
        self.sample_events(b, rng, want_candidates=diagnostics)
        if tables_done is not None:
            main.wait_event(tables_done)
        folded = self.sample_energy(b, rng, max_trials=max_trials, want_trials=diagnostics, sampler=energy_sampler,
                                    fold=True, want_energy=diagnostics, tables_ready=tables_done is not None)
        if not folded:
            self.digitize_batch(b, rng)
        if side is None:
            self.sample_background(b, rng)
        else:
            main.wait_event(bkg_done)
        if diagnostics or keep_merged:      # keeps the merged pre-dead-time list (and the per-event selection)
            self.combine(b)
            self.dead_time(b, want_selection=diagnostics)
        else:
            self.merge_dead_time(b)
        return b

    def check_status(self, b):
        c = b.counts if b.counts is not None else self.fetch_counts(b)
        bad = np.nonzero(c["status"])[0]
        if len(bad):
            names = {1: "source capacity", 2: "background capacity", 4: "injected uniforms exhausted",
                     8: "trial cap", 16: "output capacity", 32: "energy envelope violated",
                     64: "physical dead time: event rate x window too high (chain > 4096 events)",
                     128: "thinning majorant violated", 256: "thinning squeeze margin violated",
                     512: "internal: a bounded wait inside a kernel gave up"}
            ui = int(bad[0])
            st = int(c["status"][ui])
            what = ", ".join(v for k, v in names.items() if st & k)
            raise _capi.EngineError(f"{len(bad)} unit(s) failed, first: unit {ui}: {what}")

    # -- single-array conveniences used by the class shims ---------------------------------------
    def _single_unit(self, response, bkg_weights=None, interp_extrap=None, **fields):
        u = make_units(1)
        for k, v in fields.items():
            u[k] = v
        tab = response.device_table(bkg_weights, interp_extrap or self.interp_extrap)
        return u, [tab]

    def energy_integrated_evolution(self, unit, tables, times, interp_extrap=None):
        torch = _torch()
        units = np.array(unit, dtype=UNIT_DTYPE, copy=True)
        fill_gamma_constants(units)
        dtab = self.upload_tables(tables)
        du = self.to_device(units)
        times = np.atleast_1d(np.asarray(times, dtype="f8"))
        dt = self.to_device(times)
        out = self.empty(len(times), torch.float64)
        _capi.check(
            self.lib.cgrb_energy_integrated_evolution(
                dtab.data_ptr(), du.data_ptr(), 0, self._extrap_code(interp_extrap), dt.data_ptr(),
                len(times), out.data_ptr(), self.stream),
            "cgrb_energy_integrated_evolution")
        self.launches += 1
        return out[:len(times)].cpu().numpy()

    def evolution(self, unit, energy, times):
        torch = _torch()
        units = np.array(unit, dtype=UNIT_DTYPE, copy=True)
        fill_gamma_constants(units)
        du = self.to_device(units)
        energy = np.atleast_1d(np.asarray(energy, dtype="f8"))
        times = np.atleast_1d(np.asarray(times, dtype="f8"))
        de, dt = self.to_device(energy), self.to_device(times)
        out = self.empty(len(times) * len(energy), torch.float64)
        _capi.check(
            self.lib.cgrb_evolution(du.data_ptr(), 0, de.data_ptr(), len(energy), dt.data_ptr(), len(times),
                                    out.data_ptr(), self.stream),
            "cgrb_evolution")
        self.launches += 1
        return out[:len(times) * len(energy)].cpu().numpy().reshape(len(times), len(energy))

    def digitize(self, response, energies, rng=None):
        """Response.digitize for a plain array of photon energies."""
        torch = _torch()
        energies = np.atleast_1d(np.asarray(energies, dtype="f8"))
        n = len(energies)
        if n == 0:
            return np.zeros(0)
        u, tabs = self._single_unit(response, src_cap=n, flags=_capi.UNIT_NO_BACKGROUND)
        b = Batch(self, u, self.upload_tables(tabs), self.interp_extrap)
        u["src_off"], u["bkg_off"], u["tot_off"] = 0, 0, 0
        b.n_src_slots, b.n_bkg_slots, b.n_tot_slots = n + 1, 1, n + 2
        b.d_units = self.to_device(u)
        self.set_counts(b, n_src=n)
        b.t["energy_src"] = self.to_device(np.concatenate((energies, [0.0])))
        self.digitize_batch(b, rng or Philox(self.next_seed()))
        return b.t["pha_src"][:n].cpu().numpy().astype("f8")

    def sample_background_times(self, tstart, tstop, rate):
        """Background.sample_times: homogeneous Poisson arrival times, tstart prepended."""
        u = make_units(1)
        u["bkg_tstart"], u["bkg_tstop"], u["bkg_rate"] = tstart, tstop, rate
        u["flags"] = _capi.UNIT_NO_SOURCE
        b = self.new_batch(u, self._uniform_table(), compute_fmax=False)
        self.sample_background(b, Philox(self.next_seed()))
        self.fetch_counts(b)
        self.check_status(b)
        return b.view("times_bkg", 0, "bkg")

    def _uniform_table(self):
        if getattr(self, "_utab", None) is None:
            t = np.zeros(_capi.DT_SIZE)
            t[_capi.DT_BKG_CDF:_capi.DT_BKG_CDF + _capi.NCHAN] = np.arange(1, _capi.NCHAN + 1) / _capi.NCHAN
            from .response.response import background_guide

            t[_capi.DT_BKG_GUIDE:_capi.DT_BKG_GUIDE + _capi.BKG_GUIDE_N // 8] = background_guide(
                t[_capi.DT_BKG_CDF:_capi.DT_BKG_CDF + _capi.NCHAN]).view("f8")
            self._utab = [t]
        return self._utab

    def sample_background_channels(self, weights, n, rng=None):
        """BackgroundSpectrumTemplate.sample_channel(size=n): int64 channel indices."""
        torch = _torch()
        w = np.asarray(weights)
        if len(w) != _capi.NCHAN:
            raise ValueError(f"the engine is compiled for {_capi.NCHAN} channels")
        t = np.zeros(_capi.DT_SIZE)
        cdf = w.astype("f8").cumsum()
        cdf /= cdf[-1]
        t[_capi.DT_BKG_CDF:_capi.DT_BKG_CDF + _capi.NCHAN] = cdf
        dtab = self.to_device(t)
        out = self.empty(n, torch.uint8)
        rng = rng or Philox(self.next_seed())
        d = RngDesc()
        keep = None
        if isinstance(rng, Philox):
            d.mode, d.seed = 0, rng.seed & 0xFFFFFFFFFFFFFFFF
        else:
            keep = self.to_device(np.asarray(rng.bkg_chan[0], dtype="f8"))
            d.mode, d.d_u = 1, keep.data_ptr()
        _capi.check(self.lib.cgrb_background_channels(dtab.data_ptr(), 0, 0, C.byref(d), int(n), out.data_ptr(),
                                                      self.stream), "cgrb_background_channels")
        self.launches += 1
        return out[:n].cpu().numpy().astype(np.int64)

    def combine_dead_time(self, times_bkg, pha_bkg, times_src, pha_src, fused=True, mode=None):
        """LightCurve._combine + GBMLightCurve._filter_deadtime for one detector's two time-sorted
        event lists; returns (times, pha) of the survivors.  ``fused`` selects the one-pass kernel."""
        torch = _torch()
        nb, ns = len(times_bkg), len(times_src)
        u = make_units(1)
        u["src_cap"], u["bkg_cap"] = max(ns - 1, 0), max(nb - 1, 0)
        b = Batch(self, u, None, self.interp_extrap)
        b.n_src_slots, b.n_bkg_slots = int(u["src_cap"][0]) + 1, int(u["bkg_cap"][0]) + 1
        b.n_tot_slots = b.n_src_slots + b.n_bkg_slots
        b.d_units = self.to_device(u)
        self.set_counts(b, n_src=ns, n_bkg=nb)
        pad = lambda a, n, dt: np.concatenate((np.ascontiguousarray(a).astype(dt), np.zeros(max(n - len(a), 0) + 1, dt)))
        b.t["times_src"] = self.to_device(pad(times_src, b.n_src_slots, "f8"))
        b.t["pha_src"] = self.to_device(pad(pha_src, b.n_src_slots, np.uint8))
        b.t["times_bkg"] = self.to_device(pad(times_bkg, b.n_bkg_slots, "f8"))
        b.t["pha_bkg"] = self.to_device(pad(pha_bkg, b.n_bkg_slots, np.uint8))
        if fused:
            self.merge_dead_time(b, mode=mode)
        else:
            self.combine(b)
            self.dead_time(b, mode=mode)
        c = self.fetch_counts(b)
        self.check_status(b)
        n = int(c["n_kept"][0])
        assert int(c["n_total"][0]) == nb + ns
        return b.t["times_out"][:n].cpu().numpy(), b.t["pha_out"][:n].cpu().numpy()

    def gbm_dead_time(self, times, pha, mode=None):
        """_gbm_dead_time for one time-sorted event list; returns the boolean selection."""
        torch = _torch()
        times = np.ascontiguousarray(times, dtype="f8")
        n = len(times)
        if n == 0:
            return np.zeros(0, dtype=bool)
        pha8 = np.ascontiguousarray(np.asarray(pha).astype(np.int64).clip(0, 255).astype(np.uint8))
        u = make_units(1)
        u["src_cap"], u["bkg_cap"] = n, 0
        u["flags"] = _capi.UNIT_NO_BACKGROUND
        b = Batch(self, u, None, self.interp_extrap)
        b.n_src_slots, b.n_bkg_slots, b.n_tot_slots = n + 1, 1, n + 2
        b.d_units = self.to_device(u)
        self.set_counts(b, n_total=n)
        b.t["times_tot"] = self.to_device(np.concatenate((times, [0.0, 0.0])))
        b.t["pha_tot"] = self.to_device(np.concatenate((pha8, np.zeros(2, np.uint8))))
        self.dead_time(b, want_selection=True, mode=mode)
        self.fetch_counts(b)
        self.check_status(b)
        return b.t["selection"][:n].cpu().numpy().astype(bool)


_engine = None
_engine_lock = threading.Lock()


def get_engine(**kw):
    """Process-wide engine on the current CUDA device."""
    global _engine
    with _engine_lock:
        if _engine is None:
            from .config import cosmogrb_config

            opts = dict(seed=cosmogrb_config["engine"]["seed"],
                        interp_extrap=cosmogrb_config["engine"]["interp_extrap"],
                        energy_sampler=cosmogrb_config["engine"]["energy_sampler"],
                        deadtime_mode=cosmogrb_config["engine"]["deadtime_mode"],
                        deadtime_tau=cosmogrb_config["engine"]["deadtime_tau"],
                        deadtime_tau_overflow=cosmogrb_config["engine"]["deadtime_tau_overflow"],
                        thinning=cosmogrb_config["engine"]["thinning"])
            opts.update(kw)
            _engine = Engine(**opts)
        return _engine


def reset_engine():
    global _engine
    with _engine_lock:
        _engine = None
