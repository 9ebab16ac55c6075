"""Pin SURVEY.md §8 row f1 (the ``GRB.save`` / ``GRBSave.from_file`` file layout) on the REAL reference code.

TEST INFRASTRUCTURE, authoring container only (needs /root/reference).  Usage:
    python oracle/make_golden_f1.py            (writes tests/golden/grb_save_layout.json)
    python oracle/make_golden_f1.py --check    (checks 1-3 below + the committed listing, writes nothing)

h5py is not installable here, so the reference's UNMODIFIED ``GRB.save`` (grb/grb.py:205-314), ``GRBSave.from_file``
(io/grb_save.py:146-255) and ``recursively_{save,load}_dict_contents`` (utils/hdf5_utils.py) run against
``oracle/h5_standin.py`` registered as ``h5py``.  Three checks, each asserted as the fixture is minted:

  1. reference writer -> layout listing (every group / dataset / attribute path with dtype, shape, compression)
     == the listing of the file ``cosmogrb_b200.io.grb_save.write_grb_file`` writes for the same burst;
  2. a file written by ``cosmogrb_b200`` is read by the reference's ``GRBSave.from_file``: every array and
     attribute equal to what went in;
  3. a file written by the reference's ``GRB.save`` is read by ``cosmogrb_b200.io.grb_save.GRBSave.from_file``.

The listing is committed so the CPU suite can hold ``write_grb_file`` to it without the reference at hand
(tests/test_f1_layout.py); the same test re-runs 1-3 live wherever /root/reference exists.
"""
import importlib
import json
import os
import sys
import tempfile
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden", "grb_save_layout.json")

DETS = ("n0", "b1")


def burst_content(seed=5):
    """A small two-detector burst (NaI + BGO) as plain arrays: what both writers are given."""
    rng = np.random.RandomState(seed)
    dets = {}
    for k, name in enumerate(DETS):
        ns, nb = 40 + 7 * k, 65 + 3 * k
        ts = np.sort(rng.uniform(0.0, 20.0, ns))
        tb = np.sort(rng.uniform(-10.0, 30.0, nb))
        # dtypes as the reference's pipeline hands them to the storage: _digitize returns float64 channels
        # (response/response.py:13), the background template int64, the combined list goes through astype(int)
        ps, pb = rng.randint(0, 128, ns).astype(np.float64), rng.randint(0, 128, nb).astype(np.int64)
        order = np.argsort(np.concatenate([tb, ts]), kind="stable")
        tt, pt = np.concatenate([tb, ts])[order], np.concatenate([pb, ps])[order].astype(int)
        keep = np.ones(len(tt), bool)
        keep[3::17] = False
        matrix = rng.uniform(0.0, 1.0, (140, 128))           # (photon bins, channels) as Response stores it
        dets[name] = dict(
            tstart=-10.0, tstop=30.0, time_adjustment=0.25 * k, instrument="GBM",
            extra_info={"angle": 12.5 + k, "lat_lon": np.array([1.0, 2.0 + k])} if k == 0 else {},
            times=tt[keep], pha=pt[keep], times_source=ts, pha_source=ps, times_background=tb, pha_background=pb,
            matrix=matrix, energy_edges=np.geomspace(5.0, 5e4, 141), channel_edges=np.geomspace(8.0, 1000.0, 129),
            geometric_area=126.0 + k)
    return dict(name="SynthGRB_7", T0=1.5, z=2.25, duration=20.0, ra=312.0, dec=-62.0,
                source_params=dict(peak_flux=5e-7, trise=1.0, tdecay=2.0, ep_start=300.0, ep_tau=1.5, alpha=-0.8,
                                   emin=10.0, emax=1e4),
                extra_info=dict(note="population 3", weights=np.arange(3.0), nested=dict(a=1, b=2.5)),
                dets=dets)


def install_reference():
    """The reference's save / load modules behind the import shims, h5py replaced by the stand-in."""
    from oracle import h5_standin, ref_shims

    ref_shims.install("linear")
    sys.modules["h5py"] = h5_standin
    dd = types.ModuleType("dask.distributed")
    dd.worker_client = None
    sys.modules.setdefault("dask", types.ModuleType("dask"))
    sys.modules["dask.distributed"] = dd
    sys.modules["cosmogrb"].cosmogrb_config = {}
    hdf5_utils = importlib.import_module("cosmogrb.utils.hdf5_utils")
    assert hdf5_utils.h5py is h5_standin
    ns = types.SimpleNamespace(
        h5=h5_standin,
        grb=importlib.import_module("cosmogrb.grb.grb"),
        grb_save=importlib.import_module("cosmogrb.io.grb_save"),
        storage=importlib.import_module("cosmogrb.lightcurve.light_curve_storage"),
        response=importlib.import_module("cosmogrb.response.response"),
    )
    return ns


def reference_save(ref, content, file_name):
    """``GRB.save`` is called unbound on a plain object carrying the attributes the method reads: the reference's
    GRB constructors pull in gbmgeometry / dask, the method itself needs none of that."""
    lcs = {}
    for name, d in content["dets"].items():
        rsp = ref.response.Response(matrix=d["matrix"], geometric_area=d["geometric_area"],
                                    energy_edges=d["energy_edges"], channel_edges=d["channel_edges"])
        st = ref.storage.LightCurveStorage(
            name=name, tstart=d["tstart"], tstop=d["tstop"], time_adjustment=d["time_adjustment"], pha=d["pha"],
            times=d["times"], pha_source=d["pha_source"], times_source=d["times_source"],
            pha_background=d["pha_background"], times_background=d["times_background"], channels=rsp.channels,
            ebounds=rsp.channel_edges, T0=content["T0"], instrument=d["instrument"], extra_info=d["extra_info"])
        lcs[name] = types.SimpleNamespace(lightcurve_storage=st, response=rsp)
    me = types.SimpleNamespace(
        name=content["name"], T0=content["T0"], z=content["z"], duration=content["duration"], ra=content["ra"],
        dec=content["dec"], _source_params=content["source_params"], _extra_info=content["extra_info"],
        _lightcurves=lcs)
    ref.grb.GRB.save(me, file_name)


def package_save(content, file_name, backend="h5"):
    from cosmogrb_b200.io.grb_save import write_grb_file
    from cosmogrb_b200.lightcurve.light_curve_storage import LightCurveStorage
    from cosmogrb_b200.response.response import Response

    storages, responses = [], []
    for name, d in content["dets"].items():
        rsp = Response(matrix=d["matrix"], geometric_area=d["geometric_area"], energy_edges=d["energy_edges"],
                       channel_edges=d["channel_edges"])
        storages.append(LightCurveStorage(
            name=name, tstart=d["tstart"], tstop=d["tstop"], time_adjustment=d["time_adjustment"], pha=d["pha"],
            times=d["times"], pha_source=d["pha_source"], times_source=d["times_source"],
            pha_background=d["pha_background"], times_background=d["times_background"], channels=rsp.channels,
            ebounds=rsp.channel_edges, T0=content["T0"], instrument=d["instrument"], extra_info=d["extra_info"]))
        responses.append(rsp)
    write_grb_file(file_name, content["name"], content["T0"], content["z"], content["duration"], content["ra"],
                   content["dec"], content["source_params"], content["extra_info"], storages, responses,
                   backend=backend)


def _same(a, b, what):
    if isinstance(a, dict):
        assert isinstance(b, dict) and sorted(a) == sorted(b), (what, sorted(a), sorted(b) if isinstance(b, dict) else b)
        for k in a:
            _same(a[k], b[k], f"{what}/{k}")
        return
    if isinstance(b, bytes):
        b = b.decode()
    if isinstance(a, bytes):
        a = a.decode()
    if isinstance(a, str) or isinstance(b, str):
        assert str(a) == str(b), (what, a, b)
        return
    np.testing.assert_array_equal(np.asarray(a), np.asarray(b), err_msg=what)


def check_loaded(g, content, what):
    """A reloaded burst (either package's ``GRBSave``) against what was written."""
    assert g.name == content["name"], what
    for k in ("T0", "z", "duration", "ra", "dec"):
        assert float(getattr(g, k)) == content[k], (what, k)
    _same(content["source_params"], g.source_params, what + ":source")
    _same(content["extra_info"], g.extra_info, what + ":extra_info")
    assert list(g.keys()) == sorted(content["dets"]), (what, list(g.keys()))      # h5py lists groups by name
    for name, d in content["dets"].items():
        lc, rsp = g[name]["lightcurve"], g[name]["response"]
        for k in ("times", "pha", "times_source", "pha_source", "times_background", "pha_background"):
            got = getattr(lc, k)
            np.testing.assert_array_equal(got, d[k], err_msg=f"{what}:{name}:{k}")
            assert np.asarray(got).dtype == d[k].dtype, (what, name, k)
        for k in ("tstart", "tstop", "time_adjustment"):
            assert float(getattr(lc, k)) == d[k], (what, name, k)
        assert lc.instrument == d["instrument"] and float(lc.T0) == content["T0"]
        _same(d["extra_info"], lc.extra_info, f"{what}:{name}:extra_info")
        np.testing.assert_array_equal(rsp.matrix, d["matrix"])
        np.testing.assert_array_equal(rsp.energy_edges, d["energy_edges"])
        np.testing.assert_array_equal(rsp.channel_edges, d["channel_edges"])
        assert float(rsp.geometric_area) == d["geometric_area"]


def cross_check(ref, tmpdir):
    """Checks 1-3 of the module docstring; returns the reference writer's layout listing."""
    from cosmogrb_b200.io.grb_save import GRBSave

    content = burst_content()
    f_ref, f_pkg = os.path.join(tmpdir, "ref_store.h5"), os.path.join(tmpdir, "pkg_store.h5")
    reference_save(ref, content, f_ref)
    package_save(content, f_pkg, backend="h5")
    with ref.h5.File(f_ref, "r") as f:
        layout_ref = ref.h5.listing(f)
    with ref.h5.File(f_pkg, "r") as f:
        layout_pkg = ref.h5.listing(f)
    assert layout_pkg == layout_ref, _diff(layout_ref, layout_pkg)                     # 1
    check_loaded(ref.grb_save.GRBSave.from_file(f_pkg), content, "reference reads package file")     # 2
    check_loaded(GRBSave.from_file(f_ref), content, "package reads reference file")                  # 3
    check_loaded(ref.grb_save.GRBSave.from_file(f_ref), content, "reference reads reference file")
    return layout_ref


def _diff(a, b):
    lines = []
    for p in sorted(set(a) | set(b)):
        if a.get(p) != b.get(p):
            lines.append(f"{p}: reference {a.get(p)} != package {b.get(p)}")
    return "\n".join(lines)


def main():
    ref = install_reference()
    with tempfile.TemporaryDirectory() as tmp:
        layout = cross_check(ref, tmp)
    if "--check" in sys.argv:                    # tests/test_f1_layout.py: the committed listing is still what
        with open(GOLD) as fh:                   # the reference's writer produces
            assert json.load(fh) == json.loads(json.dumps(layout)), "tests/golden/grb_save_layout.json is stale"
        print("f1 cross-check ok:", len(layout), "paths")
        return
    with open(GOLD, "w") as fh:
        json.dump(layout, fh, indent=1, sort_keys=True)
        fh.write("\n")
    print("wrote", GOLD, len(layout), "paths")


if __name__ == "__main__":
    main()
