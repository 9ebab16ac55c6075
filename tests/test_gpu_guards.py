"""GPU: no kernel writes outside the buffers it was given.  compute-sanitizer is not available on the GPU pool, so
every device buffer of a batch is allocated with a guard zone of a known byte pattern behind it and the zones
are verified after the whole pipeline (production path, diagnostics path and the trigger) has run."""
import numpy as np
import pytest

from tests import helpers as h
from cosmogrb_b200 import _capi
from cosmogrb_b200.engine import Philox

pytestmark = pytest.mark.gpu
GUARD = 4096        # bytes


def _guarded(engine, monkeypatch):
    import torch

    made = []
    orig = type(engine).empty

    def empty(self, n, dtype):
        n = max(int(n), 1)
        item = torch.empty((), dtype=dtype).element_size()
        raw = torch.full((n * item + GUARD,), 0xAB, dtype=torch.uint8, device=self.device)
        made.append((raw, n * item))
        return raw[:n * item].view(dtype)

    monkeypatch.setattr(type(engine), "empty", empty)
    return made, orig


def _check(made):
    import torch

    for raw, nbytes in made:
        tail = raw[nbytes:]
        assert bool(torch.all(tail == 0xAB)), f"guard zone of a {nbytes}-byte buffer was overwritten"


@pytest.mark.parametrize("diagnostics", [False, True])
def test_no_out_of_bounds_writes(engine, monkeypatch, diagnostics):
    from cosmogrb_b200.instruments.gbm.gbm_lightcurve_analyzer import channel_bounds

    made, _ = _guarded(engine, monkeypatch)
    dets = ("n0", "b0", "n7")
    tabs = [h.response(d).device_table(h.bkg_weights(d), "linear") for d in dets]
    units = np.concatenate([
        h.unit(det_index=0, stream_id=1, **h.CONFTEST_BRIGHT),
        h.unit(det_index=1, stream_id=2, **h.README_DEMO),
        h.unit(det_index=2, stream_id=3, flags=_capi.UNIT_NO_SOURCE, bkg=(0.0, 37.3, 811.0), **h.README_DEMO),
        h.unit(det_index=0, stream_id=4, flags=_capi.UNIT_NO_BACKGROUND, **dict(h.CONFTEST_BRIGHT, peak_flux=3e-5)),
        h.unit(kind=1, det_index=2, stream_id=5, **dict(z=1.0, peak_flux=5e-9, alpha=-0.66, ep=500.0, duration=2.0)),
    ])
    b = engine.new_batch(units, tabs)
    engine.process(b, Philox(5), diagnostics=diagnostics)
    cb = np.stack([channel_bounds(h.response(d).channel_edges) for d in dets])
    engine.gbm_trigger(b, cb)
    c = engine.fetch_counts(b)
    engine.check_status(b)
    assert c["n_src"][2] == 0 and c["n_bkg"][3] == 0 and np.all(c["n_kept"] <= c["n_total"])
    assert len(made) > 10
    _check(made)
