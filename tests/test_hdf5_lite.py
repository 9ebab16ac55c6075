"""``cosmogrb_b200/utils/hdf5_lite.py``: the engine's own HDF5 reader against the REAL input files of the reference.

The reference reads its HDF5 inputs through h5py (background templates: sampler/background.py:46-60; popsynth
populations: universe/universe.py:45-47).  h5py is absent from the build image, so the files themselves are the answer
key: ``cosmogrb/data/gbm_backgrounds/<det>.h5`` (written by h5py: superblock 0, big-endian float32, contiguous) and
``cosmogrb/data/test_grb_pop.h5`` (written by popsynth: nested old-style groups, scalar / array datasets, LZF-compressed
chunked arrays, h5py bools, variable-length strings, attributes).  ``tests/golden/gbm_background_n0.h5`` is a copy of
the 2.5 KB n0 template for machines without the reference tree.
"""
import os

import numpy as np
import pytest

from cosmogrb_b200.utils import hdf5_lite as h5

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DATA = "/root/reference/cosmogrb/data"
DETS = ("n0", "n1", "n2", "n3", "n4", "n5", "n6", "n7", "n8", "n9", "na", "nb", "b0", "b1")
needs_ref = pytest.mark.skipif(not os.path.isdir(REF_DATA), reason="the reference tree is not on this machine")


def _tables():
    from cosmogrb_b200.instruments.gbm.synthetic_response import gbm_tables

    return gbm_tables()


def test_lzf_decoder_literals_back_references_and_overlap():
    # literal "abc"; back reference of length 6 at distance 3 (overlapping: "abcabc"); literal "Z";
    # long back reference: length 7 + 2 + 2 = 11 at distance 10
    stream = bytes([2]) + b"abc" + bytes([(4 << 5) | 0, 2]) + bytes([0]) + b"Z" + bytes([(7 << 5) | 0, 2, 9])
    assert h5.lzf_decompress(stream, 21) == b"abcabcabcZ" + b"abcabcabcZa"
    with pytest.raises(h5.HDF5Error):
        h5.lzf_decompress(stream, 20)
    with pytest.raises(h5.HDF5Error):
        h5.lzf_decompress(bytes([(4 << 5) | 0, 2]), 6)            # reference before the start of the output


def test_background_template_fixture_and_from_file():
    from cosmogrb_b200.sampler.background import BackgroundSpectrumTemplate

    fn = os.path.join(ROOT, "tests", "golden", "gbm_background_n0.h5")
    with h5.File(fn) as f:
        assert f.keys() == ["counts"] and "counts" in f and "nope" not in f
        d = f["counts"]
        assert d.shape == (128,) and d.dtype == np.float32
        counts = d[()]
    assert np.array_equal(counts, _tables()["bkg_counts"][0])           # the hand-decoded table the engine ships
    t = BackgroundSpectrumTemplate.from_file(fn)                        # background.py:46-60 without h5py
    assert np.array_equal(t.weights, counts / counts.sum()) and t.channels[:3] == [0, 1, 2]
    assert BackgroundSpectrumTemplate.from_file(fn, start_at_one=True).channels[0] == 1


def test_not_hdf5_and_unsupported_features_fail_loudly(tmp_path):
    p = tmp_path / "x.h5"
    p.write_bytes(b"not an hdf5 file at all")
    with pytest.raises(h5.HDF5Error):
        h5.File(p)
    raw = bytearray(open(os.path.join(ROOT, "tests", "golden", "gbm_background_n0.h5"), "rb").read())
    raw[8] = 2                                                          # superblock version 2: not decoded, not guessed
    p.write_bytes(bytes(raw))
    with pytest.raises(h5.HDF5Error, match="superblock version 2"):
        h5.File(p)
    with pytest.raises(h5.HDF5Error):
        h5.File(os.path.join(ROOT, "tests", "golden", "gbm_background_n0.h5"), "w")
    from cosmogrb_b200.universe.population import Population

    with pytest.raises((ValueError, ImportError, RuntimeError)):
        Population.from_file(tmp_path / "x.h5")


@needs_ref
@pytest.mark.parametrize("det", DETS)
def test_all_background_templates(det):
    with h5.File(os.path.join(REF_DATA, "gbm_backgrounds", f"{det}.h5")) as f:
        assert np.array_equal(f["counts"][()], _tables()["bkg_counts"][DETS.index(det)])


@needs_ref
def test_popsynth_population_file():
    """The reference's own test population (cosmogrb/test/conftest.py:101; scripts/gen_test_data.py wrote it)."""
    fn = os.path.join(REF_DATA, "test_grb_pop.h5")
    with h5.File(fn) as f:
        a = f.attrs
        assert a["name"] == "sfr_pareto" and int(a["seed"]) == 1234 and float(a["r_max"]) == 5.0 and bool(a["has_lf"])
        assert {"distances", "fluxes", "flux_obs", "theta", "phi", "selection", "auxiliary_quantities",
                "luminosities", "popsynth", "truth"} <= set(f.keys())
        assert f["selection"][()].dtype == bool and f["selection"][()].all()             # LZF chunk, h5py bool
        assert f["popsynth/auxiliary samplers/trise/type"][()] == b"TruncatedNormalAuxSampler"  # vlen string
        assert float(f["truth/alpha/lower"][()]) == -1.5 and int(f["popsynth/seed"][()]) == 1234
        assert f["unknown_distance_idx"][()].shape == (0,)
        aux = f["auxiliary_quantities"]
        t90, trise = aux["t90/true_values"][()], aux["trise/true_values"][()]
        # the generator's own relations (scripts/gen_test_data.py:20-50) hold on the decoded numbers
        np.testing.assert_allclose(aux["duration/true_values"][()], 1.5 * t90, rtol=1e-15)
        np.testing.assert_allclose(aux["tdecay/true_values"][()],
                                   (10 * t90 + trise + np.sqrt(trise) * np.sqrt(20 * t90 + trise)) / 50.0, rtol=1e-14)
        # flux = L / (4 pi d_L^2): the two bursts' ratio follows their luminosities and distances' ordering
        assert (f["fluxes"][()] > 0).all() and np.array_equal(f["fluxes"][()], f["flux_obs"][()])   # flux_sigma = -1
        names = []
        f.visititems(lambda n, o: names.append(n))
        assert len(names) > 150 and "truth/sfr/r_max" in names

    from cosmogrb_b200.universe.population import Population

    pop = Population.from_file(fn)
    assert pop.n_objects == 2 and pop.selection.all()
    with h5.File(fn) as f:
        np.testing.assert_array_equal(pop.distances, f["distances"][()])
        np.testing.assert_array_equal(pop.fluxes.latent, f["fluxes"][()])
        np.testing.assert_allclose(pop.ra, np.rad2deg(f["phi"][()]))
        np.testing.assert_allclose(pop.dec, 90.0 - np.rad2deg(f["theta"][()]))
    assert ((0 <= pop.ra) & (pop.ra <= 360)).all() and ((-90 <= pop.dec) & (pop.dec <= 90)).all()
    for k in ("ep", "alpha", "trise", "tdecay", "tau", "duration"):     # gbm_universe.py:19-29 reads exactly these
        assert len(getattr(pop, k)) == 2
    assert ((-1.5 <= pop.alpha) & (pop.alpha <= 0)).all() and ((1.5 <= pop.tau) & (pop.tau <= 2.5)).all()


@needs_ref
def test_universe_from_the_reference_test_population(tmp_path):
    """``GBM_CPL_Universe(population_file)`` (universe/universe.py:30-72, gbm_universe.py:10-29) on the file the
    reference's own test-suite uses: two parameter servers with the file's numbers (no GPU needed to build them)."""
    from cosmogrb_b200.instruments.gbm.gbm_universe import GBM_CPL_Universe

    fn = os.path.join(REF_DATA, "test_grb_pop.h5")
    uni = GBM_CPL_Universe(fn, save_path=str(tmp_path))
    ps = uni._parameter_servers
    assert len(ps) == 2 and [p.parameters["name"] for p in ps] == ["SynthGRB_0", "SynthGRB_1"]
    with h5.File(fn) as f:
        aux = f["auxiliary_quantities"]
        for i, p in enumerate(ps):
            q = p.parameters
            assert q["z"] == f["distances"][()][i] and q["peak_flux"] == f["fluxes"][()][i] and q["T0"] == 0.0
            assert q["ep_start"] == aux["ep/true_values"][()][i] and q["ep_tau"] == aux["tau/true_values"][()][i]
            assert q["alpha"] == aux["alpha/true_values"][()][i] and q["duration"] == aux["duration/true_values"][()][i]
            assert q["trise"] == aux["trise/true_values"][()][i] and q["tdecay"] == aux["tdecay/true_values"][()][i]


@needs_ref
def test_store_reads_real_hdf5_without_h5py():
    """``io.store.open_tree`` (behind ``GRBSave.from_file``) on a real HDF5 file when h5py is not importable."""
    from cosmogrb_b200.io import store

    if store.have_h5py():
        pytest.skip("h5py is installed: open_tree takes the h5py route")
    t = store.open_tree(os.path.join(REF_DATA, "test_grb_pop.h5"))
    assert t.attrs("/")["name"] == "sfr_pareto" and int(t.attrs("/")["n_model"]) == 500
    assert t.get("distances").shape == (2,) and "auxiliary_quantities/ep/true_values" in t
    assert t.get("popsynth/auxiliary samplers/ep/type") == "Log10NormalAuxSampler"
    d = t.load_dict("spatial_params")
    assert sorted(d) == ["a", "decay", "peak", "r0", "r_max", "rise"] and float(d["r_max"][0]) == 5.0
