"""the ``GRB.save`` / ``GRBSave.from_file`` layout, pinned on the reference's own writer and reader.

``tests/golden/grb_save_layout.json`` is the listing (every group / dataset / attribute path, dtype, shape, compression)
of the file the reference's UNMODIFIED ``GRB.save`` (grb/grb.py:205-314) writes for the burst of
``oracle/make_golden_f1.py::burst_content`` -- with ``oracle/h5_standin.py`` in place of h5py, which is not
installable here.  The CPU suite holds ``write_grb_file`` to that listing through both backends; where
/root/reference exists the cross-reads run live as well (reference reader on the package's file and the reverse).
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "grb_save_layout.json")


@pytest.fixture
def standin_h5py(monkeypatch):
    from oracle import h5_standin

    monkeypatch.setitem(sys.modules, "h5py", h5_standin)
    return h5_standin


def test_h5_backend_writes_the_reference_layout(tmp_path, standin_h5py):
    from cosmogrb_b200.io.grb_save import GRBSave
    from oracle import make_golden_f1 as f1

    content = f1.burst_content()
    fn = str(tmp_path / "SynthGRB_7_store.h5")
    f1.package_save(content, fn, backend="h5")
    with standin_h5py.File(fn, "r") as f:
        layout = standin_h5py.listing(f)
    with open(GOLD) as fh:
        gold = json.load(fh)
    assert json.loads(json.dumps(layout)) == gold, f1._diff(gold, json.loads(json.dumps(layout)))
    # the quirks of the reference's writer that the listing carries
    assert "/n0/extra_info/angle" in gold and "/detectors/n0/extra_info" not in gold    # grb.py:256-259: at the ROOT
    assert gold["/detectors/n0/source_signal/pha"]["dtype"] == "f8"                     # response.py:13 (float channels)
    assert gold["/detectors/n0/channels"]["compression"] is None
    f1.check_loaded(GRBSave.from_file(fn), content, "package reads its own h5 file")


def test_npz_twin_has_the_same_paths(tmp_path):
    """The NumPy twin (no h5py at all) stores the same tree: paths, attribute names, dtypes and shapes."""
    from cosmogrb_b200.io import store
    from cosmogrb_b200.io.grb_save import GRBSave
    from oracle import make_golden_f1 as f1

    content = f1.burst_content()
    fn = str(tmp_path / "SynthGRB_7_store.h5")
    f1.package_save(content, fn, backend="npz")
    t = store.open_tree(fn)
    with open(GOLD) as fh:
        gold = json.load(fh)
    for path, e in gold.items():
        assert path in t, path
        if e["kind"] == "group":
            assert t.is_group(path) and sorted(t.attrs(path)) == sorted(e["attrs"]), path
            for k, kind in e["attrs"].items():
                v = t.attrs(path)[k]
                assert ("str" if isinstance(v, str) else np.asarray(v).dtype.str.lstrip("<>|=")) == kind, (path, k)
        else:
            v = t.get(path)
            if e["dtype"] == "str":
                assert isinstance(v, str), path
            else:
                assert np.asarray(v).dtype.str.lstrip("<>|=") == e["dtype"] and list(np.shape(v)) == e["shape"], path
    n_paths = len(t._datasets) + len(t._attrs)
    assert n_paths == len(gold), (n_paths, len(gold))
    f1.check_loaded(GRBSave.from_file(fn), content, "package reads its npz twin")


@pytest.mark.skipif(not os.path.isdir("/root/reference/cosmogrb"), reason="the reference tree is not on this machine")
def test_cross_read_with_the_reference_live():
    """Reference ``GRBSave.from_file`` on the package's file, the package's reader on the reference's file, and the
    committed listing against the reference's writer (a process of its own: the import shims replace modules)."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "make_golden_f1.py"), "--check"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "f1 cross-check ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
