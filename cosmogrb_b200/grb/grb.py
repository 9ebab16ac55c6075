"""``GRB`` (cosmogrb/grb/grb.py:24-368): parameters, detector fan-out and storage.

``go()`` replaces the reference's dask map of ``process_lightcurve`` over the detectors with one
batched GPU run of all light curves (lightcurve.process_batch).  ``client``/``serial`` are accepted
for signature compatibility; there is nothing to distribute inside one GRB any more.
"""
import abc
import collections

import numpy as np

from ..io.grb_save import write_grb_file
from ..lightcurve.lightcurve import process_batch
from ..sampler.source_function import SourceFunction
from ..utils.logging import setup_logger
from ..utils.meta import GRBMeta, RequiredParameter

logger = setup_logger(__name__)


class GRB(object, metaclass=GRBMeta):
    name = RequiredParameter(default="SynthGRB")
    z = RequiredParameter(default=1, vmin=0, vmax=20)
    T0 = RequiredParameter(default=0)
    ra = RequiredParameter(default=0, vmin=0, vmax=360)
    dec = RequiredParameter(default=0, vmin=-90, vmax=90)
    duration = RequiredParameter(default=1, vmin=0)

    def __init__(self, source_function_class=None, **kwargs):
        self._source_params = {}
        self._required_params = {}
        self._lightcurves = collections.OrderedDict()
        self._responses = collections.OrderedDict()
        self._backgrounds = collections.OrderedDict()
        assert issubclass(source_function_class, SourceFunction)
        self._source_function = source_function_class
        # unknown keyword arguments are silently dropped, like the reference (grb.py:61-80)
        for k, v in kwargs.items():
            if k in self._parameter_names:
                self._source_params[k] = v
                getattr(self, k)            # validates the bounds
            elif k in self._required_names:
                self._required_params[k] = v
                getattr(self, k)
        self._extra_info = {}
        self._setup()

    @abc.abstractmethod
    def _setup(self):
        pass

    def _add_background(self, name, background):
        self._backgrounds[name] = background

    def _add_response(self, name, response):
        self._responses[name] = response

    def _add_lightcurve(self, lightcurve):
        self._lightcurves[lightcurve.name] = lightcurve

    @property
    def lightcurves(self):
        return self._lightcurves

    def _check(self):
        for key in self._required_names:
            assert getattr(self, key) is not None, f"you have not set {key}"
        assert self.z > 0, f"z: {self.z} must be greater than zero"
        assert self.duration > 0, f"duration: {self.duration} must be greater than zero"

    def go(self, client=None, serial=False, seed=None, stream_base=0):
        """Simulate every detector of the burst (grb.py:152-203) in one GPU batch."""
        self._check()
        lcs = list(self._lightcurves.values())
        ids = [stream_base + i for i in range(len(lcs))]
        for lc, storage in zip(lcs, process_batch(lcs, seed=seed, stream_ids=ids)):
            lc.set_storage(storage)

    def save(self, file_name, clean_up=False):
        """Write the burst to ``file_name`` (layout of grb.py:205-314; HDF5 when h5py is available,
        otherwise the key-path-identical .npz twin, see io/store.py)."""
        write_grb_file(file_name, name=self.name, T0=self.T0, z=self.z, duration=self.duration, ra=self.ra,
                       dec=self.dec, source_params=self._source_params, extra_info=self._extra_info,
                       storages=[lc.lightcurve_storage for lc in self._lightcurves.values()],
                       responses=[lc.response for lc in self._lightcurves.values()])
        if clean_up:
            self._lightcurves.clear()

    def _output(self):
        d = collections.OrderedDict(name=self.name, z=self.z, ra=self.ra, dec=self.dec, duration=self.duration,
                                    T0=self.T0)
        d.update(self._source_params)
        if self._extra_info:
            d.update(self._extra_info)
        return d

    def __repr__(self):
        return "\n".join(f"{k:<12}{v}" for k, v in self._output().items())

    def info(self):
        print(repr(self))


def process_lightcurve(lc):
    return lc.process()
