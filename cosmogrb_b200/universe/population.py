"""Population input of ``Universe`` runs.

The reference reads a popsynth population file (universe/universe.py:45-47) and uses the columns
``ra, dec, distances (z), duration, ep, alpha, fluxes.latent, trise, tdecay, tau``
(universe.py:74-88, instruments/gbm/gbm_universe.py:19-29).  popsynth cannot be installed offline,
so ``Population`` is a plain column container with an ``.npz`` file format (same column names; a
popsynth file is accepted when popsynth is importable), and ``synthetic_population`` draws a
popsynth-style population as the reference's test-data generator does (scripts/gen_test_data.py:20-126,
SURVEY.md §8d).
"""
from types import SimpleNamespace

import numpy as np

COLUMNS = ("ra", "dec", "distances", "duration", "ep", "alpha", "fluxes_latent", "trise", "tdecay", "tau")


class Population(object):
    def __init__(self, **columns):
        missing = [c for c in ("ra", "dec", "distances", "duration") if c not in columns]
        if missing:
            raise RuntimeError(f"The population must contain {missing}")
        n = len(np.atleast_1d(columns["ra"]))
        self._cols = {}
        for k, v in columns.items():
            v = np.atleast_1d(np.asarray(v, dtype="f8"))
            if len(v) != n:
                raise ValueError(f"column {k} has {len(v)} entries, expected {n}")
            self._cols[k] = v
        self.n_objects = n
        self.selection = np.ones(n, dtype=bool)

    def __getattr__(self, name):
        cols = self.__dict__.get("_cols", {})
        if name in cols:
            return cols[name]
        if name == "fluxes" and "fluxes_latent" in cols:
            return SimpleNamespace(latent=cols["fluxes_latent"])
        raise AttributeError(name)

    def __len__(self):
        return self.n_objects

    def to_sub_population(self):
        return self

    def writeto(self, file_name):
        with open(str(file_name), "wb") as fh:
            np.savez_compressed(fh, **self._cols)

    @classmethod
    def from_file(cls, file_name):
        file_name = str(file_name)
        with open(file_name, "rb") as fh:
            magic = fh.read(4)
        if magic.startswith(b"PK"):
            with np.load(file_name) as z:
                return cls(**{k: z[k] for k in z.files})
        try:
            import popsynth
        except ImportError:
            return cls._from_popsynth_hdf5(file_name)
        pop = popsynth.Population.from_file(file_name).to_sub_population()
        cols = dict(ra=pop.ra, dec=pop.dec, distances=pop.distances, duration=pop.duration)
        for k in ("ep", "alpha", "trise", "tdecay", "tau"):
            if hasattr(pop, k):
                cols[k] = getattr(pop, k)
        cols["fluxes_latent"] = pop.fluxes.latent
        return cls(**cols)

    @classmethod
    def _from_popsynth_hdf5(cls, file_name):
        """A popsynth population file without popsynth (and without h5py: ``utils/hdf5_lite``).  The layout is the one
        ``popsynth.Population.writeto`` produces -- ``distances``, ``fluxes`` (latent), ``theta`` / ``phi`` (sky
        position in radians: ra = deg(phi), dec = 90 - deg(theta)), ``selection`` and one group per auxiliary quantity
        under ``auxiliary_quantities`` with its ``true_values`` -- and, like ``to_sub_population()``
        (universe/universe.py:45-47), only the selected objects are kept."""
        from ..utils import hdf5_lite as h5

        try:
            f = h5.File(file_name)
        except h5.HDF5Error as e:
            raise ValueError(f"{file_name} is neither an .npz population nor a readable popsynth HDF5 file: {e}") from e
        for need in ("distances", "fluxes", "theta", "phi", "selection", "auxiliary_quantities"):
            if need not in f:
                raise RuntimeError(f"{file_name}: not a popsynth population file (no {need!r})")
        sel = np.asarray(f["selection"][()], dtype=bool)
        cols = dict(ra=np.rad2deg(f["phi"][()])[sel], dec=(90.0 - np.rad2deg(f["theta"][()]))[sel],
                    distances=f["distances"][()][sel], fluxes_latent=f["fluxes"][()][sel])
        aux = f["auxiliary_quantities"]
        for k in aux.keys():
            cols[k] = aux[k]["true_values"][()][sel]
        return cls(**cols)


def _truncnorm(rng, mu, sigma, lo, hi, n):
    out = rng.normal(mu, sigma, n)
    bad = (out < lo) | (out > hi)
    while bad.any():
        out[bad] = rng.normal(mu, sigma, bad.sum())
        bad = (out < lo) | (out > hi)
    return out


def _luminosity_distance_cm(z):
    """flat LCDM (H0 = 67.4, Om = 0.315), trapezoid on a fine grid"""
    zg = np.linspace(0.0, max(float(np.max(z)), 1e-3), 4096)
    ez = np.sqrt(0.315 * (1 + zg) ** 3 + 0.685)
    dc = np.concatenate(([0.0], np.cumsum(0.5 * (1 / ez[1:] + 1 / ez[:-1]) * np.diff(zg))))
    c_over_h0_cm = 2.99792458e5 / 67.4 * 3.0856775814913673e24
    return (1 + z) * np.interp(z, zg, dc) * c_over_h0_cm


def synthetic_population(n, seed=12345):
    """popsynth-style GRB population (shapes follow scripts/gen_test_data.py:20-126)."""
    rng = np.random.default_rng(seed)
    # redshift: SFR-like pdf on (0, 5]  ~ (1 + r z) / (1 + (z / p)^d), r = 0.1... rise, peak 1.5, decay 4
    zg = np.linspace(1e-3, 5.0, 2000)
    pdf = (1.0 + 0.1 * zg) / (1.0 + (zg / 1.5) ** 4.0) * zg ** 2 / (1 + zg)
    cdf = np.cumsum(pdf)
    cdf /= cdf[-1]
    z = np.interp(rng.random(n), cdf, zg)
    lum = 1e50 * (1.0 - rng.random(n)) ** (-1.0 / 1.5)                     # Pareto(Lmin = 1e50, alpha = 1.5)
    peak_flux = np.clip(lum / (4 * np.pi * _luminosity_distance_cm(z) ** 2), 1e-9, 1e-4)
    ep = 10 ** rng.normal(2.0, 0.5, n)
    alpha = _truncnorm(rng, -1.0, 0.25, -1.5, 0.0, n)
    alpha[alpha == -1.0] = -1.0001                                         # alpha == -1 is broken in the reference
    tau = _truncnorm(rng, 2.0, 0.25, 1.5, 2.5, n)
    trise = _truncnorm(rng, 1.0, 1.0, 0.01, 5.0, n)
    t90 = 10 ** rng.normal(1.0, 0.25, n)
    tdecay = (10 * t90 + trise + np.sqrt(trise) * np.sqrt(20 * t90 + trise)) / 50.0
    duration = 1.5 * t90
    ra = rng.uniform(0, 360, n)
    dec = np.rad2deg(np.arcsin(rng.uniform(-1, 1, n)))
    return Population(ra=ra, dec=dec, distances=z, duration=duration, ep=ep, alpha=alpha,
                      fluxes_latent=peak_flux, trise=trise, tdecay=tdecay, tau=tau)
