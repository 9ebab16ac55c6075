"""Background samplers (cosmogrb/sampler/background.py:47-174) over the CUDA engine."""
import numpy as np

from .. import _capi
from ..engine import Philox, get_engine, make_units
from .sampler import Sampler


class BackgroundSpectrumTemplate(object):
    def __init__(self, counts, start_at_one=False):
        self._counts = np.asarray(counts)
        j = 1 if start_at_one else 0
        self._start = j
        self._channels = [i + j for i in range(len(self._counts))]
        self._normalize_counts()

    def _normalize_counts(self):
        self._weights = self._counts / self._counts.sum()       # background.py:69-73

    @property
    def weights(self):
        return self._weights

    @property
    def channels(self):
        return self._channels

    def sample_channel(self, size=None):
        """``np.random.choice(channels, size, p=weights)`` of the reference (:80-93) on the GPU."""
        n = 1 if size is None else int(size)
        eng = get_engine()
        out = eng.sample_background_channels(self._weights, n) + self._start
        return out[0] if size is None else out

    @classmethod
    def from_file(cls, file_name, start_at_one=False):
        try:
            import h5py as h5
        except ImportError:             # the engine's own reader decodes the reference's template files
            from ..utils import hdf5_lite as h5
        with h5.File(file_name, "r") as f:
            counts = f["counts"][()]
        return cls(counts, start_at_one)


class Background(Sampler):
    def __init__(self, tstart, tstop, average_rate=1000, background_spectrum_template=None):
        self._background_rate = np.random.normal(average_rate, 10)      # background.py:125
        self._background_spectrum_template = background_spectrum_template
        super(Background, self).__init__(tstart=tstart, tstop=tstop)

    @property
    def background_rate(self):
        return self._background_rate

    @property
    def spectrum_template(self):
        return self._background_spectrum_template

    def sample_times(self):
        eng = get_engine()
        return eng.sample_background_times(self._tstart, self._tstop, self._background_rate)

    def sample_channel(self, size=None):
        if self._background_spectrum_template is not None:
            return self._background_spectrum_template.sample_channel(size=size)
        raise NotImplementedError()
