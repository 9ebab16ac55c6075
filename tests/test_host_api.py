"""CPU: host-side mirror of the reference API -- parameter descriptors / metaclass (the reference's
own test/test_grb.py cases), the GRBSave file layout round trip, populations, rank sharding and
the world_size-2 gloo gather of event counts."""
import os
import subprocess
import sys

import numpy as np
import pytest

from cosmogrb_b200.grb import GRB, SourceParameter
from cosmogrb_b200.io import store
from cosmogrb_b200.io.grb_save import GRBSave, write_grb_file
from cosmogrb_b200.lightcurve.light_curve_storage import LightCurveStorage
from cosmogrb_b200.sampler.constant_cpl import ConstantCPL
from cosmogrb_b200.sampler.cpl_source import CPLSourceFunction
from cosmogrb_b200.universe import Population, synthetic_population
from cosmogrb_b200.universe.universe import shard_by_cost
from tests import helpers as h

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class DummyGRB(GRB):
    def __init__(self, **kwargs):
        super(DummyGRB, self).__init__(**kwargs, source_function_class=ConstantCPL)

    def _setup(self):
        pass


class ChildDummyGRB(DummyGRB):
    x = SourceParameter(default=0, vmin=-1, vmax=1)


def test_grb_constructor():                      # reference test/test_grb.py:28-90
    with pytest.raises(TypeError):
        GRB()
    grb = DummyGRB()
    assert grb._source_params == {}
    for rp in ["name", "T0", "z", "ra", "dec", "duration"]:
        assert rp in GRB._required_names and rp in DummyGRB._required_names
    assert (grb.name, grb.T0, grb.z, grb.ra, grb.dec, grb.duration) == ("SynthGRB", 0, 1, 0, 0, 1)
    grb = DummyGRB(name="AnotherDummy")
    assert grb.name == "AnotherDummy" and grb.z == 1
    grb = DummyGRB(name="name", T0=10, z=4, ra=10, dec=4, duration=3.0)
    assert (grb.name, grb.T0, grb.z, grb.ra, grb.dec, grb.duration) == ("name", 10, 4, 10, 4, 3)
    with pytest.raises(AssertionError):
        DummyGRB(z=-1)
    grb = DummyGRB()
    with pytest.raises(AssertionError):
        grb.z = -10
    with pytest.raises(AssertionError):
        grb.z = 1000


def test_inherited_grb():                        # reference test/test_grb.py:93-153
    grb = ChildDummyGRB()
    assert grb._source_params == {}
    assert "x" in ChildDummyGRB._parameter_names and "x" not in DummyGRB._parameter_names
    for rp in ["name", "T0", "z", "ra", "dec", "duration"]:
        assert rp in ChildDummyGRB._required_names
    assert grb.x == 0
    grb = ChildDummyGRB(name="AnotherDummy", x=0.5)
    assert grb.name == "AnotherDummy" and grb.x == 0.5
    grb = ChildDummyGRB(name="name", T0=10, z=4, ra=10, dec=4, duration=3.0, x=-0.2)
    assert grb.x == -0.2 and grb.duration == 3
    with pytest.raises(AssertionError):
        ChildDummyGRB(x=-10)
    grb = ChildDummyGRB()
    with pytest.raises(AssertionError):
        grb.x = -10
    with pytest.raises(AssertionError):
        grb.x = 1000
    # unknown keyword arguments are silently dropped (grb.py:61-80; README's stale `ep=`/`tau=`)
    grb = ChildDummyGRB(not_a_parameter=3)
    assert "not_a_parameter" not in grb._source_params


def test_source_function_asserts_negative_alpha():
    rsp = h.response("n0")
    with pytest.raises(AssertionError):
        CPLSourceFunction(alpha=0.5, response=rsp)
    with pytest.raises(AssertionError):
        ConstantCPL(alpha=0.0, response=rsp)
    sf = CPLSourceFunction(alpha=-0.7, response=rsp)
    assert sf.emin == rsp.emin and sf.emax == rsp.emax and sf.index == -0.7


def _storage(name, n=50, seed=0):
    rs = np.random.RandomState(seed)
    rsp = h.response(name)
    t = np.sort(rs.uniform(-100, 300, n))
    return LightCurveStorage(name=name, tstart=-100, tstop=300, time_adjustment=1.5, pha=rs.randint(0, 128, n),
                             times=t, pha_source=rs.randint(0, 128, 7).astype("f8"), times_source=t[:7],
                             pha_background=rs.randint(0, 128, n - 5), times_background=t[5:],
                             channels=rsp.channels, ebounds=rsp.channel_edges, T0=0.1, instrument="GBM",
                             extra_info={"angle": 12.5}), rsp


@pytest.mark.parametrize("backend", ["npz"] + (["h5"] if store.have_h5py() else []))
def test_grb_file_round_trip(tmp_path, backend):
    pairs = [_storage(d, seed=i) for i, d in enumerate(("n0", "n1", "b0"))]
    fn = tmp_path / "SynthGRB_0_store.h5"
    write_grb_file(fn, name="SynthGRB_0", T0=0.1, z=1.0, duration=2.0, ra=312.0, dec=-62.0,
                   source_params=dict(peak_flux=1e-5, alpha=-0.66, ep_start=500.0), extra_info={"note": "x"},
                   storages=[p[0] for p in pairs], responses=[p[1] for p in pairs], backend=backend)
    g = GRBSave.from_file(fn)
    assert (g.name, g.z, g.ra, g.dec, g.duration, g.T0) == ("SynthGRB_0", 1.0, 312.0, -62.0, 2.0, 0.1)
    assert g.source_params["alpha"] == -0.66 and g.extra_info["note"] == "x"
    assert list(g.keys()) == ["b0", "n0", "n1"]          # name order, as h5py lists a group (either backend)
    for (lc, rsp) in pairs:
        got = g[lc.name]["lightcurve"]
        assert np.array_equal(got.times, lc.times) and np.array_equal(got.pha, lc.pha)
        assert np.array_equal(got.times_source, lc.times_source) and np.array_equal(got.pha_source, lc.pha_source)
        assert np.array_equal(got.times_background, lc.times_background)
        assert got.tstart == -100 and got.tstop == 300 and got.time_adjustment == 1.5 and got.instrument == "GBM"
        assert got.extra_info["angle"] == 12.5
        r2 = g[lc.name]["response"]
        assert np.array_equal(r2.matrix, rsp.matrix) and r2.geometric_area == rsp.geometric_area
        assert np.array_equal(r2.cumulative_matrix, rsp.cumulative_matrix)
    with pytest.raises(RuntimeWarning):
        g["n0"] = 1
    # layout paths of grb/grb.py:205-314
    t = store.open_tree(fn)
    for p in ("detectors/n0/total_signal/times", "detectors/n0/source_signal/pha",
              "detectors/n0/background_signal/times", "detectors/n0/response/matrix", "n0/extra_info/angle",
              "source/peak_flux", "detectors/n0/channels"):
        assert p in t, p
    assert t.attrs("/")["n_lightcurves"] == 3


def test_storage_binning():
    lc, _ = _storage("n0", n=2000)
    txt = repr(lc)                              # light_curve_storage.py:597-640
    assert "n_counts_source" in txt and "time adjustment" in txt and "angle" in txt and "GBM" in txt
    bins, counts = lc.binned_counts(1.0, None, None, tmin=-100, tmax=300)
    assert bins.shape[1] == 2 and counts.sum() <= 2000 and counts.sum() >= 1990
    assert lc.count_spectrum().sum() == 2000
    assert lc.n_counts == 2000 and lc.n_counts_source == 7


def test_population_and_universe_setup(tmp_path):
    pop = synthetic_population(200, seed=12345)
    pop2 = synthetic_population(200, seed=12345)
    assert np.array_equal(pop.fluxes.latent, pop2.fluxes.latent)
    assert np.all((pop.alpha >= -1.5) & (pop.alpha < 0) & (pop.alpha != -1.0))
    assert np.all((pop.fluxes.latent >= 1e-9) & (pop.fluxes.latent <= 1e-4))
    fn = tmp_path / "pop.npz"
    pop.writeto(fn)
    from cosmogrb_b200.instruments.gbm.gbm_universe import GBM_CPL_Universe

    uni = GBM_CPL_Universe(str(fn), save_path=str(tmp_path))
    assert uni.n_grbs == 200
    ps = uni.parameter_servers[3]
    assert set(ps.parameters) == {"name", "ra", "dec", "z", "duration", "T0", "peak_flux", "alpha", "ep_start",
                                  "ep_tau", "trise", "tdecay"}
    assert ps.parameters["ep_start"] == pop.ep[3] and ps.parameters["ep_tau"] == pop.tau[3]     # gbm_universe.py:24-29
    assert ps.parameters["name"] == "SynthGRB_3" and str(ps.file_path).endswith("SynthGRB_3_store.h5")
    units, ang = uni.build_units(np.array([3, 7]))
    assert len(units) == 28 and ang.shape == (2, 14)
    assert np.all(units["stream_id"][:14] == 3 * 16 + np.arange(14))
    assert np.all((units["ea_scale"] > 0) & (units["ea_scale"] <= 1))
    u2, _ = uni.build_units(np.array([7]))                    # unit records do not depend on the batching
    assert np.array_equal(units[14:].tobytes(), u2.tobytes())
    # background rates ~ N(500, 10): one table per seed for the whole population, so neither the batching nor
    # the sharding changes a burst's rates; another seed gives other rates
    u3, _ = uni.build_units(np.arange(200), rate_seed=5)
    u4, _ = uni.build_units(np.array([11, 150]), rate_seed=5)
    assert np.array_equal(u3["bkg_rate"][11 * 14:12 * 14], u4["bkg_rate"][:14])
    assert np.array_equal(u3["bkg_rate"][150 * 14:151 * 14], u4["bkg_rate"][14:])
    assert abs(u3["bkg_rate"].mean() - 500.0) < 1.0 and 8.0 < u3["bkg_rate"].std() < 12.0
    u5, _ = uni.build_units(np.array([11]), rate_seed=6)
    assert not np.array_equal(u5["bkg_rate"], u4["bkg_rate"][:14])
    with pytest.raises(RuntimeError):
        Population(ra=[1.0], dec=[2.0])


def test_gamma_constants_per_distinct_alpha():
    """fill_gamma_constants evaluates math.gamma / lgamma once per distinct alpha: same values as unit by unit"""
    import math

    from cosmogrb_b200.engine import fill_gamma_constants, make_units

    u = make_units(30)
    u["alpha"] = np.repeat([-0.66, -1.5, -1.0 + 1e-9, -2.5, 0.3], 6)
    fill_gamma_constants(u)
    for i in range(30):
        a = 2.0 + float(u["alpha"][i])
        if a > 0:
            assert u["gamma_a"][i] == math.gamma(a) and u["lgamma_a"][i] == math.lgamma(a)
            assert u["lgamma_1pa"][i] == math.lgamma(1.0 + a)
        else:
            assert np.isnan(u["gamma_a"][i]) and np.isnan(u["lgamma_a"][i]) and np.isnan(u["lgamma_1pa"][i])


def test_shard_by_cost():
    rs = np.random.RandomState(0)
    cost = rs.pareto(1.5, 1000) + 1
    for world in (1, 2, 4, 8):
        shards = shard_by_cost(cost, world)
        allidx = np.sort(np.concatenate(shards))
        assert np.array_equal(allidx, np.arange(1000))
        loads = np.array([cost[s].sum() for s in shards])
        assert loads.max() <= loads.mean() * 1.05 + cost.max()
    assert [list(s) for s in shard_by_cost([], 2)] == [[], []]
    # a population-sized deal (heap for the expensive head, snake for the cheap tail): a permutation, balanced to a
    # small fraction of a per cent, skipped bursts (cost 0) included, deterministic, and fast enough to be no head
    cost = rs.pareto(1.2, 100000) + 1
    cost[rs.randint(0, 100000, 1300)] = 0.0
    import time

    for world in (2, 8):
        t0 = time.perf_counter()
        shards = shard_by_cost(cost, world)
        assert time.perf_counter() - t0 < 0.1
        assert np.array_equal(np.sort(np.concatenate(shards)), np.arange(100000))
        loads = np.array([cost[s].sum() for s in shards])
        assert loads.max() <= loads.mean() * 1.002
        assert all(np.array_equal(a, b) for a, b in zip(shards, shard_by_cost(cost, world)))


GLOO_WORKER = r"""
import os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np, torch.distributed as dist
from cosmogrb_b200.universe.universe import gather_event_counts, gather_rows, dist_info, shard_by_cost
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + sys.argv[2], rank=int(sys.argv[3]), world_size=2)
rank, world = dist_info()
assert world == 2
cost = np.arange(1, 12, dtype="f8")
mine = shard_by_cost(cost, world)[rank]
local = np.stack([mine * 10 + 1, mine * 10 + 2, mine * 10 + 3], axis=1)   # ragged: ranks hold different counts
parts = gather_event_counts(local)
full = np.zeros((11, 3), dtype=np.int64)
for r in range(world):
    full[shard_by_cost(cost, world)[r]] = parts[r]
want = np.stack([np.arange(11) * 10 + k for k in (1, 2, 3)], axis=1)
assert np.array_equal(full, want), full
offsets = np.concatenate(([0], np.cumsum(full[:, 2])))[:-1]              # global event offsets
assert offsets[1] == 3
# sizing pass: each rank sizes its slice of the bursts, every rank ends up with every burst's row
parts = np.array_split(np.arange(11), world)
rows = np.stack([parts[rank] + 0.25, np.where(parts[rank] == 3, np.inf, parts[rank] * 2.0)], axis=1)
allrows = np.concatenate(gather_rows(rows), axis=0)
assert allrows.shape == (11, 2) and np.array_equal(allrows[:, 0], np.arange(11) + 0.25)
assert np.isinf(allrows[3, 1]) and allrows[4, 1] == 8.0
dist.barrier(); dist.destroy_process_group()
print("ok", rank)
"""


def test_gather_event_counts_gloo_world2(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(GLOO_WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
        assert "ok" in o


def test_capacity_saturates_instead_of_wrapping():
    """fmax x duration of a runaway unit (1e19 candidates, inf, nan) must not wrap around in the integer cast:
    it saturates and the engine then refuses the unit"""
    from cosmogrb_b200.engine import CAPACITY_MAX, capacity

    c = capacity(np.array([0.0, 500.0, 1e19, np.inf, np.nan, -3.0]), 30.0)
    assert c[0] == 64 and c[5] == 64
    assert 15000 + 8 * np.sqrt(15000) <= c[1] <= 15000 + 8 * np.sqrt(15000) + 66
    assert np.all(c[2:5] == CAPACITY_MAX)
    assert np.all(c >= 0)


def test_pipeline_chunks_cover_every_unit_once():
    """process_batch cuts a large call into chunks of detectors (device->host copy of one chunk overlaps the
    kernels of the next); every unit appears exactly once, in order, and small calls stay in one piece"""
    from cosmogrb_b200 import _capi
    from cosmogrb_b200.lightcurve import lightcurve as lcmod

    u = np.zeros(14, dtype=_capi.UNIT_DTYPE)
    u["src_tstop"], u["bkg_tstart"], u["bkg_tstop"], u["bkg_rate"] = 300.0, -100.0, 300.0, 500.0
    u["fmax"] = 3.6e4                                  # ~1.1e7 candidates per detector: the bright burst
    chunks = lcmod._chunks(u, None)
    assert len(chunks) >= 4
    assert np.array_equal(np.concatenate(chunks), np.arange(14))
    u["fmax"] = 100.0                                  # the README burst: one batch
    assert len(lcmod._chunks(u, None)) == 1
    assert len(lcmod._chunks(u[:1], None)) == 1


def test_universe_index_lists_only_simulated_bursts(tmp_path):
    """Universe.save: string datasets survive both backends (h5py has no conversion for numpy '<U' arrays: they
    are written as variable-length strings), and bursts that were skipped (counts -1) have no file listed"""
    pop = synthetic_population(6, seed=3)
    fn = tmp_path / "pop.npz"
    pop.writeto(fn)
    from cosmogrb_b200.instruments.gbm.gbm_universe import GBM_CPL_Universe

    uni = GBM_CPL_Universe(str(fn), save_path=str(tmp_path))
    ec = np.full((6, 14, 3), 10, dtype=np.int64)
    ec[4] = -1                                                # a skipped burst
    uni.event_counts, uni._is_processed = ec, True
    for backend in ["npz"] + (["h5"] if store.have_h5py() else []):
        out = tmp_path / f"index_{backend}.h5"
        real_write = store.Tree.write
        try:
            store.Tree.write = lambda self, file_name, backend=backend, _w=real_write: _w(self, file_name, backend=backend)
            uni.save(out)
        finally:
            store.Tree.write = real_write
        t = store.open_tree(out)
        files = [str(x) for x in t.get("grb_save_files")]
        assert len(files) == 5 and not any(f.endswith("SynthGRB_4_store.h5") for f in files)
        assert files[0].endswith("SynthGRB_0_store.h5")
        assert np.array_equal(t.get("event_counts"), ec)


def test_h5_backend_string_datasets(tmp_path):
    pytest.importorskip("h5py")
    t = store.Tree()
    t.dataset("names", np.array(["a", "bcd"]))
    t.dataset("extra_info/tags", ["x", "yz"])
    t.dataset("extra_info/note", "hello")
    t.write(tmp_path / "s.h5", backend="h5")
    r = store.open_tree(tmp_path / "s.h5")
    assert list(r.get("names")) == ["a", "bcd"] and list(r.get("extra_info/tags")) == ["x", "yz"]
    assert r.get("extra_info/note") == "hello"


def test_source_work_list_per_unit_segments():
    """engine.source_work_list: the thinning work list with short segments for small units.  With the rule off it is
    cgrb_build_work(which=0); with it on every unit's candidates are covered once, by segments of its own length
    (a function of the unit alone), starts even, segment lengths at most CGRB_SEG."""
    from cosmogrb_b200 import _capi, engine

    rs = np.random.RandomState(4)
    caps = np.concatenate(([0, 1, 1023, 1024, 1025, 65536, 65537, 8192 * 3 + 1], rs.randint(1, 400000, 200)))
    units = np.zeros(len(caps), dtype=_capi.UNIT_DTYPE)
    units["src_cap"] = caps
    ref = _capi.build_work(units, 0, _capi.SEG)
    off = engine.source_work_list(caps, _capi.SEG, (-1, 1024))
    assert np.array_equal(off, ref)
    w = engine.source_work_list(caps, _capi.SEG, (65536, 1024))
    assert np.all(np.diff(w["unit"]) >= 0) and set(w["unit"]) == set(range(len(caps)))
    for u, cap in enumerate(caps):
        mine = w[w["unit"] == u]
        ch = 1024 if cap <= 65536 else _capi.SEG
        assert np.array_equal(mine["seg"], np.arange(len(mine))) and np.array_equal(mine["start"], np.arange(len(mine)) * ch)
        assert len(mine) == max(-(-int(cap) // ch), 1)
        assert np.all(mine["start"] % 2 == 0) and ch <= _capi.SEG and ch % 512 == 0
    # what the kernels derive from it: a segment ends where the unit's next entry starts, the last one at the capacity
    nxt_same = np.concatenate((w["unit"][1:] == w["unit"][:-1], [False]))
    seg_len = np.where(nxt_same, np.concatenate((w["start"][1:], [0])) - w["start"], _capi.SEG)
    n = np.minimum(seg_len, caps[w["unit"]] - w["start"])
    assert np.all(n <= _capi.SEG) and np.all(n[caps[w["unit"]] > 0] > 0)
    assert np.array_equal(np.bincount(w["unit"], weights=np.maximum(n, 0), minlength=len(caps)).astype(np.int64), caps)
