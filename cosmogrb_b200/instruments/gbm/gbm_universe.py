"""GBM population runs (cosmogrb/instruments/gbm/gbm_universe.py:7-175).

``GBM_CPL_Universe`` maps the population columns to GRB keyword arguments exactly like the
reference (:19-29: ep -> ep_start, fluxes.latent -> peak_flux, tau -> ep_tau) and simulates the
bursts in batches of (GRB x 14 detector) units.  The 14 detector tables are shared by all bursts
(synthetic DRM = base matrix x viewing-angle factor, carried per unit in ``ea_scale``).
"""
import numpy as np

from ... import _capi
from ...engine import Philox, get_engine, make_units
from ...io.grb_save import write_grb_file
from ...lightcurve.light_curve_storage import LightCurveStorage
from ...utils.logging import setup_logger
from ...universe.universe import (GRBWrapper, ParameterServer, Universe, dist_info, gather_event_counts, gather_rows,
                                  shard_by_cost)
from .gbm_background import gbm_background_template
from .gbm_grb import GBMGRB, GBMGRB_CPL, GBMGRB_CPL_Constant
from .synthetic_response import GBM_DETECTORS, angle_factor, base_response, separation_angles

logger = setup_logger(__name__)

N_DET = len(GBM_DETECTORS)
BYTES_PER_SLOT = 8 + 8 + 8 + 1 + 8 + 1 + 8 + 1 + 1      # stage/times/energy/pha + merged + filtered + flags


class _GBMBatchedUniverse(Universe):
    """Shared batched dispatch of the two GBM universes."""

    _kind = 0

    def _unit_columns(self, idx):
        raise NotImplementedError()

    def base_tables(self, interp_extrap):
        return [base_response(d).device_table(gbm_background_template(d).weights, interp_extrap)
                for d in GBM_DETECTORS]

    def build_units(self, idx, rate_seed=0):
        """Unit records of the GRBs ``idx`` (14 per GRB, GRB-major)."""
        idx = np.asarray(idx, dtype=np.int64)
        n = len(idx)
        u = make_units(n * N_DET)
        cols = self._unit_columns(idx)
        for k, v in cols.items():
            u[k] = np.repeat(np.asarray(v, dtype="f8"), N_DET)
        u["kind"] = self._kind
        u["z"] = np.repeat(self._z[idx], N_DET)
        u["det"] = np.tile(np.arange(N_DET), n)
        u["stream_id"] = (np.repeat(idx, N_DET) * 16 + np.tile(np.arange(N_DET), n)).astype(np.uint32)
        u["src_tstart"] = 0.0
        u["src_tstop"] = np.repeat(self._duration[idx], N_DET)
        u["bkg_tstart"], u["bkg_tstop"] = GBMGRB._background_start, GBMGRB._background_stop
        # Background._background_rate ~ N(500, 10) per detector (background.py:125, gbm_grb.py:120): one table
        # for the whole population, drawn once per seed and indexed by GRB, so that neither the sharding nor
        # the batching changes a burst's rates
        u["bkg_rate"] = self._background_rates(rate_seed)[idx].reshape(-1)
        ang = separation_angles(self._ra[idx], self._dec[idx])              # [n, 14]
        fac = np.stack([angle_factor(d, ang[:, j]) for j, d in enumerate(GBM_DETECTORS)], axis=1)
        u["ea_scale"] = fac.reshape(-1)
        return u, ang

    def _background_rates(self, rate_seed):
        cache = self.__dict__.setdefault("_rate_cache", {})
        if rate_seed not in cache:
            cache.clear()
            cache[rate_seed] = np.random.default_rng([0x6B67, int(rate_seed)]).normal(500.0, 10.0, (self._n_grbs, N_DET))
        return cache[rate_seed]

    def cost_estimate(self):
        """relative cost of each GRB for the LPT deal: background events + a flux-weighted source term"""
        pf = np.asarray(self._local_parameters["peak_flux"])
        return 2e5 * N_DET + 2e10 * pf * self._duration

    def _run_batched(self, grbs_per_batch=None, save=True, seed=None, memory_budget=64e9, trigger=False,
                     trigger_threshold=4.5, simul_trigger_window=0.5, max_n_dets=12, runaway="skip"):
        """Batches of GRBs sized to ``memory_budget`` bytes of event buffers.  A burst whose candidate
        count (fmax x duration) does not fit the budget on its own is SKIPPED with a warning and listed in
        ``self.skipped`` (its event counts read -1): the reference's model makes lambda(t) explode for
        low-Epeak bursts at late times (the normalisation band drifts above the cutoff), and its own
        thinning loop would take years on them."""
        import time

        import torch

        from ...engine import capacity

        tm = self.timings = dict(fmax_pass=0.0, build_units=0.0, new_batch=0.0, launch=0.0, wait=0.0, gather=0.0)
        t_ = time.perf_counter
        eng = get_engine()
        rank, world = dist_info()
        tabs = self.base_tables(eng.interp_extrap)
        dtab = eng.upload_tables(tabs)
        seed = eng.seed if seed is None else seed

        unit_cache = {}      # GRB index -> (its 14 unit records, angles) built in the sizing pass

        def fmax_and_sizes(grbs):
            """fmax of every unit of the bursts ``grbs`` (one launch per 4096 bursts) and the bytes of event
            buffers each burst needs"""
            fm = np.zeros((len(grbs), N_DET))
            sz = np.zeros(len(grbs))
            for p0 in range(0, len(grbs), 4096):
                idx = grbs[p0:p0 + 4096]
                units, ang = self.build_units(idx, rate_seed=seed & 0xFFFFFFFF)
                unit_cache[(int(idx[0]), int(idx[-1]), len(idx))] = (idx, units, ang)
                f = eng.fmax(units, dtab)
                fm[p0:p0 + len(idx)] = f.reshape(-1, N_DET)
                slots = (capacity(f, units["src_tstop"] - units["src_tstart"]).astype("f8")
                         + capacity(units["bkg_rate"], units["bkg_tstop"] - units["bkg_tstart"]) + 2)
                sz[p0:p0 + len(idx)] = slots.reshape(-1, N_DET).sum(axis=1) * BYTES_PER_SLOT
            return fm, sz

        t0 = t_()
        if world > 1:
            # The LPT deal needs the real cost of every burst (costs are heavy-tailed in ways the flux alone does
            # not predict; bursts that will be skipped cost nothing).  Each rank sizes 1 / world of the bursts and
            # the rows (14 fmax + bytes) are exchanged: the sizing head stays constant under weak scaling.
            parts = np.array_split(np.arange(self._n_grbs), world)
            fm_p, sz_p = fmax_and_sizes(parts[rank])
            rows = gather_rows(np.concatenate((fm_p, sz_p[:, None]), axis=1))
            allrows = np.concatenate(rows, axis=0)
            assert allrows.shape == (self._n_grbs, N_DET + 1)
            fm_g, sz_g = allrows[:, :N_DET], allrows[:, N_DET]
            cost = np.where((sz_g > memory_budget) | ~np.isfinite(sz_g), 0.0, sz_g)
            self._shards = shard_by_cost(cost, world)
            mine = self._shards[rank]
            fmax_all, sizes = fm_g[mine], sz_g[mine]
        else:
            self._shards = [np.arange(self._n_grbs)]
            mine = self._shards[0]
            fmax_all, sizes = fmax_and_sizes(mine)
        local_counts = np.zeros((len(mine), N_DET, 3), dtype=np.int64)
        pinned = {}
        self.n_batches = 0
        self.skipped = []
        if trigger:
            from .gbm_lightcurve_analyzer import channel_bounds, simultaneous_trigger

            chan_bounds = np.stack([channel_bounds(base_response(d).channel_edges) for d in GBM_DETECTORS])
            local_trig = np.zeros((len(mine), N_DET), dtype=_capi.TRIGGER_DTYPE)
            local_det = np.zeros(len(mine), dtype=bool)
        tm["fmax_pass"] = t_() - t0
        too_big = (sizes > memory_budget) | ~np.isfinite(sizes)
        if runaway == "raise" and np.any(too_big):
            names = [self._name[int(mine[j])] for j in np.nonzero(too_big)[0][:8]]
            raise _capi.EngineError(f"{int(too_big.sum())} burst(s) do not fit the memory budget ({memory_budget:.3g} B of "
                                    f"event buffers; runaway lambda(t)): {', '.join(names)}{' ...' if too_big.sum() > 8 else ''}")
        for j in np.nonzero(too_big)[0]:
            g = int(mine[j])
            self.skipped.append(g)
            local_counts[j] = -1
            logger.warning(f"{self._name[g]}: {sizes[j] / BYTES_PER_SLOT:.3g} event slots do not fit the memory "
                           f"budget ({memory_budget:.3g} B): skipped")
        todo = np.nonzero(~too_big)[0]

        def finish(item):
            """read back the counters of a launched batch (and its events when saving)"""
            sel, idx, b, ang, ready = item
            t3 = t_()
            tr = None
            if save:
                ready.synchronize()         # the batch ran on its own stream: the copies below are issued on the caller's
                arrays = eng.to_host(b, pinned)
                c = arrays["counts"]
                self._save_chunk(idx, b, arrays, ang)
                if trigger:
                    tr = eng.fetch_trigger(b)
            else:
                # read on the copy stream behind the batch's own event: a plain .cpu() would queue behind the
                # NEXT batch's kernels on the compute stream and serialise the pipeline
                c, tr = eng.fetch_results_after(b, ready, want_trigger=trigger)
            if trigger:
                tr = tr.reshape(-1, N_DET)
                local_trig[sel] = tr
                for j in range(len(sel)):
                    local_det[sel[j]] = simultaneous_trigger(
                        GBM_DETECTORS, ang[j], tr[j]["detected"], tr[j]["time"], tr[j]["time_scale"],
                        simul_trigger_window=simul_trigger_window, max_n_dets=max_n_dets)[0]
            tm["wait"] += t_() - t3
            local_counts[sel, :, 0] = c["n_src"].reshape(-1, N_DET)
            local_counts[sel, :, 1] = c["n_bkg"].reshape(-1, N_DET)
            local_counts[sel, :, 2] = c["n_kept"].reshape(-1, N_DET)

        def cached_units(idx):
            """unit records of the bursts ``idx`` cut from the sizing pass's blocks (rebuilt if not there)"""
            for (cidx, cunits, cang) in unit_cache.values():
                pos = np.searchsorted(cidx, idx)
                if cidx[0] <= idx[0] and idx[-1] <= cidx[-1] and np.all(pos < len(cidx)) and np.array_equal(cidx[pos], idx):
                    rows = (pos[:, None] * N_DET + np.arange(N_DET)[None, :]).reshape(-1)
                    return cunits[rows].copy(), cang[pos]
            return self.build_units(idx, rate_seed=seed & 0xFFFFFFFF)

        # two batches in flight: the host prepares and launches batch k+1 before it waits for batch k
        import os

        # bytes of event buffers per batch (two batches are in flight); COSMOGRB_B200_BATCH_BYTES overrides it for
        # tuning runs without moving the skip threshold above
        budget = float(os.environ.get("COSMOGRB_B200_BATCH_BYTES", memory_budget / 2))
        # The two batches in flight run on two compute streams (even / odd batches; each owns one pool slot, so
        # buffer reuse stays stream-ordered): the last, partly filled wave of CTAs of one batch's kernels and its
        # latency-bound stretches (trigger search, look-backs) are filled with the other batch's work.
        # COSMOGRB_B200_STREAMS=1 puts everything back on the caller's stream.
        caller_stream = torch.cuda.current_stream(eng.device)
        n_streams = int(os.environ.get("COSMOGRB_B200_STREAMS", "2"))
        if n_streams > 1:
            if getattr(eng, "_batch_streams", None) is None:
                eng._batch_streams = [torch.cuda.Stream(device=eng.device) for _ in range(2)]
            batch_streams = eng._batch_streams
            for bs in batch_streams:
                bs.wait_stream(caller_stream)
        else:
            batch_streams = [caller_stream, caller_stream]
        pos, in_flight = 0, None
        while pos < len(todo):
            # grow the batch until the memory budget (or grbs_per_batch) is reached
            acc, end = 0.0, pos
            while end < len(todo) and (end == pos or acc + sizes[todo[end]] <= budget) and \
                    (end - pos) < (grbs_per_batch or 512):
                acc += sizes[todo[end]]
                end += 1
            sel = todo[pos:end]
            idx = mine[sel]
            t0 = t_()
            units, ang = cached_units(idx)
            units["fmax"] = fmax_all[sel].reshape(-1)
            t1 = t_()
            eng.use_pool(self.n_batches % 2)          # event buffers recycled from the slot used two batches ago
            try:
                with torch.cuda.stream(batch_streams[self.n_batches % 2]):
                    b = eng.new_batch(units, dtab, compute_fmax=False)
                    t2 = t_()
                    eng.process(b, Philox(seed))
                    if trigger:
                        eng.gbm_trigger(b, chan_bounds, threshold=trigger_threshold, fetch=False)
                    ready = torch.cuda.Event()
                    ready.record(torch.cuda.current_stream(eng.device))
            finally:
                eng.use_pool(None)
            t3 = t_()
            tm["build_units"] += t1 - t0
            tm["new_batch"] += t2 - t1
            tm["launch"] += t3 - t2
            if in_flight is not None:
                finish(in_flight)
            in_flight = (sel, idx, b, ang, ready)
            pos = end
            self.n_batches += 1
        if in_flight is not None:
            finish(in_flight)
        if n_streams > 1:
            for bs in batch_streams:
                caller_stream.wait_stream(bs)
        # the one exchange of the population run: per-unit event counts of every rank
        t0 = t_()
        gathered = gather_event_counts(local_counts.reshape(-1, 3))
        tm["gather"] = t_() - t0
        shards = self._shards
        full = np.zeros((self._n_grbs, N_DET, 3), dtype=np.int64)
        for r in range(world):
            full[shards[r]] = gathered[r].reshape(-1, N_DET, 3)
        self.event_counts = full
        self.local_grbs = mine
        if trigger:
            # detection flags of the local bursts (the all_gather above carries counts only; a flag per burst
            # travels the same way when a global view is wanted)
            self.trigger_results = local_trig
            self.detected = local_det

    def _save_chunk(self, idx, b, arrays, ang):
        def sl(name, i):
            o = arrays[name + "_offsets"]
            return arrays[name][int(o[i]):int(o[i + 1])].numpy().copy()

        for j, g in enumerate(idx):
            ps = self._parameter_servers[int(g)]
            storages, responses = [], []
            for d, det in enumerate(GBM_DETECTORS):
                i = j * N_DET + d
                rsp = base_response(det)
                storages.append(LightCurveStorage(
                    name=det, tstart=GBMGRB._background_start, tstop=GBMGRB._background_stop, time_adjustment=0.0,
                    pha=sl("pha_out", i), times=sl("times_out", i), pha_source=sl("pha_src", i).astype("f8"),
                    times_source=sl("times_src", i), pha_background=sl("pha_bkg", i).astype(np.int64),
                    times_background=sl("times_bkg", i), channels=rsp.channels, ebounds=rsp.channel_edges, T0=0.0,
                    instrument="GBM", extra_info={"angle": float(np.rad2deg(ang[j, d])),
                                                  "ea_scale": float(b.units["ea_scale"][i])}))
                responses.append(rsp)
            p = dict(ps.parameters)
            source_params = {k: p[k] for k in p if k not in ("name", "ra", "dec", "z", "duration", "T0")}
            write_grb_file(ps.file_path, name=p["name"], T0=p["T0"], z=p["z"], duration=p["duration"], ra=p["ra"],
                           dec=p["dec"], source_params=source_params, extra_info={}, storages=storages,
                           responses=responses)


class GBM_CPL_Universe(_GBMBatchedUniverse):
    _kind = 0

    def __init__(self, population, save_path="."):
        super(GBM_CPL_Universe, self).__init__(population, save_path=save_path)

    def _grb_wrapper(self, parameter_server, serial=False):
        return GBM_CPL_Wrapper(parameter_server, serial=serial)

    def _process_populations(self):
        super(GBM_CPL_Universe, self)._process_populations()
        self._local_parameters["ep_start"] = self._population.ep             # gbm_universe.py:24-29
        self._local_parameters["alpha"] = self._population.alpha
        self._local_parameters["peak_flux"] = self._population.fluxes.latent
        self._local_parameters["trise"] = self._population.trise
        self._local_parameters["tdecay"] = self._population.tdecay
        self._local_parameters["ep_tau"] = self._population.tau

    def _unit_columns(self, idx):
        lp = self._local_parameters
        return {k: np.asarray(lp[k])[idx] for k in ("ep_start", "alpha", "peak_flux", "trise", "tdecay", "ep_tau")}

    def _parameter_server_type(self, **kwargs):
        return GBM_CPL_ParameterServer(**kwargs)


class GBM_CPL_Wrapper(GRBWrapper):
    def _grb_type(self, **kwargs):
        return GBMGRB_CPL(**kwargs)


class GBM_CPL_ParameterServer(ParameterServer):
    def __init__(self, name, ra, dec, z, duration, T0, peak_flux, alpha, ep_start, ep_tau, trise, tdecay):
        super(GBM_CPL_ParameterServer, self).__init__(
            name=name, ra=ra, dec=dec, z=z, duration=duration, T0=T0, peak_flux=peak_flux, alpha=alpha,
            ep_start=ep_start, ep_tau=ep_tau, trise=trise, tdecay=tdecay)


class GBM_CPL_Constant_Universe(_GBMBatchedUniverse):
    _kind = 1

    def __init__(self, population, save_path="."):
        super(GBM_CPL_Constant_Universe, self).__init__(population, save_path=save_path)

    def _grb_wrapper(self, parameter_server, serial=False):
        return GBM_CPL_Constant_Wrapper(parameter_server, serial=serial)

    def _process_populations(self):
        super(GBM_CPL_Constant_Universe, self)._process_populations()
        self._local_parameters["ep"] = self._population.ep
        self._local_parameters["alpha"] = self._population.alpha
        self._local_parameters["peak_flux"] = self._population.fluxes.latent

    def _unit_columns(self, idx):
        lp = self._local_parameters
        n = len(idx)
        return dict(ep_start=np.asarray(lp["ep"])[idx], alpha=np.asarray(lp["alpha"])[idx],
                    peak_flux=np.asarray(lp["peak_flux"])[idx], trise=np.ones(n), tdecay=np.ones(n), ep_tau=np.ones(n))

    def _parameter_server_type(self, **kwargs):
        return GBM_CPL_Constant_ParameterServer(**kwargs)


class GBM_CPL_Constant_Wrapper(GRBWrapper):
    def _grb_type(self, **kwargs):
        return GBMGRB_CPL_Constant(**kwargs)


class GBM_CPL_Constant_ParameterServer(ParameterServer):
    def __init__(self, name, ra, dec, z, duration, T0, peak_flux, alpha, ep):
        super(GBM_CPL_Constant_ParameterServer, self).__init__(
            name=name, ra=ra, dec=dec, z=z, duration=duration, T0=T0, peak_flux=peak_flux, alpha=alpha, ep=ep)


__all__ = ["GBM_CPL_Universe", "GBM_CPL_Constant_Universe", "GBM_CPL_Wrapper", "GBM_CPL_Constant_Wrapper",
           "GBM_CPL_ParameterServer", "GBM_CPL_Constant_ParameterServer"]
