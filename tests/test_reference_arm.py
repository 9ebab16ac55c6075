"""CPU: the timed CPU arm of bench.py (``--impl reference``).  The reference's own numba kernels and classes run from
/root/reference or from the copy staged under oracle/_ref (oracle/stage_ref.py, oracle/ref_numba.py); where neither
is at hand the tests skip (the arm then falls back to the C port and says so)."""
import filecmp
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _root():
    from oracle import ref_numba

    root, why = ref_numba.available()
    if root is None:
        pytest.skip(why)
    return root


def test_staging_copies_the_reference_files_unchanged(tmp_path):
    from oracle import stage_ref

    if not os.path.isdir("/root/reference/cosmogrb"):
        pytest.skip("no /root/reference here")
    n = stage_ref.stage("/root/reference", str(tmp_path))
    assert n >= 15 and stage_ref.staged_root(str(tmp_path)) == str(tmp_path)
    for rel in stage_ref.NEEDED:
        src = os.path.join("/root/reference/cosmogrb", rel)
        if os.path.exists(src):
            assert filecmp.cmp(src, os.path.join(str(tmp_path), "cosmogrb", rel), shallow=False)
    # nothing of it is tracked by git
    tracked = subprocess.run(["git", "ls-files", "oracle/_ref"], cwd=ROOT, capture_output=True, text=True).stdout
    assert tracked.strip() == ""


def test_reference_windows_agree_with_the_port():
    """GBMLightCurve.process() of the real reference on windows of the bright burst (whole-burst fmax) against the C
    oracle's process_detector on the same windows: two draws of the same Poisson processes."""
    _root()
    code = r'''
import json, os, sys
sys.path.insert(0, %r)
from oracle import ref_numba
from oracle import oracle as orc
from cosmogrb_b200 import workloads as h
p = dict(h.BRIGHT_LONG)
fm = ref_numba.setup(p)
rsp = h.response("n0")
s = orc.source(kind=0, interp_extrap="linear", peak_flux=p["peak_flux"], ep_start=p["ep_start"], ep_tau=p["ep_tau"],
               alpha=p["alpha"], trise=p["trise"], tdecay=p["tdecay"], z=p["z"], emin=rsp.emin, emax=rsp.emax)
fo = orc.fmax(s, rsp.effective_area_packed, 0.0, p["duration"])
ev, n_src, dt = ref_numba.run_sample(4, 0.004, 2, p["duration"])
tot = src = 0
for k in range(4):
    a = (k + 0.5) * 300.0 / 4 - 0.002
    ba = -100.0 + a * 400.0 / 300.0
    n, c = orc.process_detector(s, rsp.energy_edges, rsp.effective_area_packed, rsp.cumulative_matrix, h.bkg_weights("n0"),
                                a, a + 0.004, ba, ba + 0.004 * 400.0 / 300.0, 500.0, seed_numba=5 + k, seed_numpy=9 + k, fmax=fo)
    tot += n
    src += int(c[0])
print(json.dumps(dict(fm=fm, fo=fo, ev=ev, n_src=n_src, tot=tot, src=src, dt=dt)))
''' % ROOT
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    import json

    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert abs(d["fm"] - d["fo"]) <= 1e-10 * d["fo"]
    assert d["ev"] > 100 and d["dt"] > 0
    assert abs(d["n_src"] - d["src"]) < 6.0 * np.sqrt(d["n_src"] + d["src"] + 1.0)
    assert abs(d["ev"] - d["tot"]) < 6.0 * np.sqrt(d["ev"] + d["tot"] + 1.0)
