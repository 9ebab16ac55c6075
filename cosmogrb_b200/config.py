"""Configuration (mirror of cosmogrb/config.py:13-83 without OmegaConf, which is optional here).

Same keys as the reference (``logging``, ``gbm.orbit``, ``multiprocess``) plus an ``engine``
section for the GPU photon-event engine.  A plain nested dict with attribute access is used; if
``~/.config/cosomogrb/cosmogrb_config.yml`` (sic, the reference's spelling) exists and PyYAML is
importable it is merged over the defaults.  Nothing is written to the home directory.
"""
from __future__ import annotations

from pathlib import Path


class _Node(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:  # pragma: no cover
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def _wrap(d):
    return _Node({k: _wrap(v) if isinstance(v, dict) else v for k, v in d.items()})


_DEFAULTS = {
    "logging": {
        "debug": False,
        "level": "INFO",
        "console": {"on": True, "level": "WARNING"},
        "file": {"on": False, "level": "WARNING"},
    },
    "gbm": {"orbit": {"default_time": 0.0, "use_random_time": True}},
    "multiprocess": {"n_grb_workers": 6, "n_universe_workers": 6},
    "engine": {
        # Philox key of the process-wide engine
        "seed": 0x00C0560612B,
        # semantics of interpolation.interp outside the table: "linear" | "clamp"
        "interp_extrap": "linear",
        # photon-energy sampler under Philox: "envelope" (tight exact envelope) | "literal"
        "energy_sampler": "envelope",
        # source-time proposal under Philox: "majorant" (piecewise bound from the squeeze table) | "literal"
        "thinning": "majorant",
        # upper bound on CPL rejection trials per photon
        "max_trials": 1 << 22,
        # GBM dead time: "reference" (literal _gbm_dead_time, gbm_lightcurve.py:80-115) | "physical"
        # (non-paralyzable: every kept event opens a window of deadtime_tau, or deadtime_tau_overflow for
        # the overflow channel)
        "deadtime_mode": "reference",
        "deadtime_tau": 2.6e-6,
        "deadtime_tau_overflow": 10e-6,
    },
}


def _merge(base, over):
    for k, v in over.items():
        if isinstance(v, dict) and isinstance(base.get(k), dict):
            _merge(base[k], v)
        else:
            base[k] = v


def _load():
    import copy

    cfg = copy.deepcopy(_DEFAULTS)
    path = Path("~/.config/cosomogrb/cosmogrb_config.yml").expanduser()
    if path.is_file():
        try:
            import yaml

            with path.open() as f:
                local = yaml.safe_load(f) or {}
            _merge(cfg, local)
        except Exception:  # a broken user file must not break the engine
            pass
    return _wrap(cfg)


cosmogrb_config = _load()

__all__ = ["cosmogrb_config"]
