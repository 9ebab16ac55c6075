"""``GRBSave`` -- a saved burst reloaded (cosmogrb/io/grb_save.py:13-305); the writer is ``GRB.save``."""
import collections

from ..lightcurve.light_curve_storage import LightCurveStorage
from ..response.response import Response
from . import store


def write_grb_file(file_name, name, T0, z, duration, ra, dec, source_params, extra_info, storages, responses,
                   backend=None):
    """The on-disk layout of ``GRB.save`` (grb/grb.py:205-314): root attrs, ``source/<param>``,
    optional ``extra_info``, ``detectors/<det>/{total,source,background}_signal/{pha,times}``,
    ``detectors/<det>/response/...`` and the per-detector extra info at the ROOT (``/<det>/extra_info``,
    grb.py:256-259)."""
    import numpy as np

    tree = store.Tree()
    tree.attrs("/").update(grb_name=name, n_lightcurves=len(storages), T0=T0, z=z, duration=duration, ra=ra, dec=dec)
    tree.group("source")
    for k, v in source_params.items():
        tree.dataset(f"source/{k}", v)
    if extra_info:
        for k, v in extra_info.items():
            tree.dataset(f"extra_info/{k}", v)
    tree.group("detectors")
    for lc, rsp in zip(storages, responses):
        g = f"detectors/{lc.name}"
        tree.attrs(g).update(tstart=lc.tstart, tstop=lc.tstop, time_adjustment=lc.time_adjustment,
                             instrument=lc.instrument)
        if lc.extra_info:
            for k, v in lc.extra_info.items():
                tree.dataset(f"{lc.name}/extra_info/{k}", v)
        tree.dataset(f"{g}/channels", np.asarray(lc.channels))
        for grp, pha, times in (("total_signal", lc.pha, lc.times), ("source_signal", lc.pha_source, lc.times_source),
                                ("background_signal", lc.pha_background, lc.times_background)):
            tree.dataset(f"{g}/{grp}/pha", pha, compress=True)
            tree.dataset(f"{g}/{grp}/times", times, compress=True)
        tree.dataset(f"{g}/response/matrix", rsp.matrix, compress=True)
        tree.dataset(f"{g}/response/energy_edges", rsp.energy_edges, compress=True)
        tree.dataset(f"{g}/response/channel_edges", rsp.channel_edges, compress=True)
        tree.attrs(f"{g}/response")["geometric_area"] = rsp.geometric_area
    tree.write(file_name, backend=backend)


class GRBSave(collections.UserDict):
    def __init__(self, grb_name, T0, ra, dec, duration, z, lightcurves, responses, source_params, extra_info=None):
        assert isinstance(lightcurves, dict), "lightcurves must be a dict containing LightCurveStorage objects"
        assert isinstance(responses, dict), "responses must be a dict containing Response objects"
        for k, v in lightcurves.items():
            assert isinstance(v, LightCurveStorage), f"{k} is not of type LightCurveStorage"
        for k, v in responses.items():
            assert isinstance(v, Response), f"{k} is not of type Response"
        for key1, key2 in zip(responses.keys(), lightcurves.keys()):
            assert key1 == key2, f"{key1} is not {key2} so there is a mismatch between the responses and lightcurves"
        self._lock_dict = False
        super(GRBSave, self).__init__(
            {key: dict(lightcurve=lightcurves[key], response=responses[key]) for key in lightcurves})
        self._lock_dict = True
        assert isinstance(source_params, dict)
        self._source_params = source_params
        if extra_info is not None:
            assert isinstance(extra_info, dict)
        self._extra_info = extra_info
        self._name, self._T0, self._ra, self._dec, self._z, self._duration = grb_name, T0, ra, dec, z, duration

    def __setitem__(self, key, value):
        if self._lock_dict:
            raise RuntimeWarning("Cannot modify the internal data!")
        super(GRBSave, self).__setitem__(key, value)

    def __delitem__(self, key):
        if self._lock_dict:
            raise RuntimeWarning("Cannot modify the internal data!")
        super(GRBSave, self).__delitem__(key)

    name = property(lambda self: self._name)
    ra = property(lambda self: self._ra)
    dec = property(lambda self: self._dec)
    z = property(lambda self: self._z)
    duration = property(lambda self: self._duration)
    T0 = property(lambda self: self._T0)
    extra_info = property(lambda self: self._extra_info)
    source_params = property(lambda self: self._source_params)

    @classmethod
    def from_file(cls, file_name):
        t = store.open_tree(file_name)
        root = t.attrs("/")
        source_params = t.load_dict("source") if "source" in t else {}
        extra_info = t.load_dict("extra_info") if "extra_info" in t else None
        lightcurves, responses = {}, {}
        for lc_name in t.children("detectors"):
            g = f"detectors/{lc_name}"
            a = t.attrs(g)
            lc_extra = t.load_dict(f"{lc_name}/extra_info") if f"{lc_name}/extra_info" in t else {}
            rsp = Response(matrix=t.get(f"{g}/response/matrix"),
                           geometric_area=t.attrs(f"{g}/response")["geometric_area"],
                           energy_edges=t.get(f"{g}/response/energy_edges"),
                           channel_edges=t.get(f"{g}/response/channel_edges"))
            responses[lc_name] = rsp
            lightcurves[lc_name] = LightCurveStorage(
                name=lc_name, tstart=a["tstart"], tstop=a["tstop"], time_adjustment=a["time_adjustment"],
                pha=t.get(f"{g}/total_signal/pha"), times=t.get(f"{g}/total_signal/times"),
                pha_source=t.get(f"{g}/source_signal/pha"), times_source=t.get(f"{g}/source_signal/times"),
                pha_background=t.get(f"{g}/background_signal/pha"),
                times_background=t.get(f"{g}/background_signal/times"),
                channels=rsp.channels, ebounds=rsp.channel_edges, T0=root["T0"], instrument=a["instrument"],
                extra_info=lc_extra)
        return cls(grb_name=root["grb_name"], T0=root["T0"], ra=root["ra"], dec=root["dec"], z=root["z"],
                   duration=root["duration"], lightcurves=lightcurves, responses=responses,
                   source_params=source_params, extra_info=extra_info)

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
