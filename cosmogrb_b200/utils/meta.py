"""Parameter descriptors and the GRB metaclass (same semantics as cosmogrb/utils/meta.py:1-171:
values live in the owning object's ``_source_params`` / ``_required_params`` dict, defaults are
materialised on first read, bounds violations raise ``AssertionError`` on get and on set, and
subclasses inherit the parameter names of their bases)."""


class Parameter(object):
    def __init__(self, dict_name, default=None, vmin=None, vmax=None):
        self.name = None
        self._dict_name = dict_name
        self._default = default
        self._vmin = vmin
        self._vmax = vmax
        self._current_value = default

    @property
    def default(self):
        return self._default

    def _check(self, value):
        if isinstance(value, str):
            return
        if self._vmin is not None:
            assert value >= self._vmin, f"trying to set {self.name} to a value below {self._vmin} is not allowed"
        if self._vmax is not None:
            assert value <= self._vmax, f"trying to set {self.name} to a value above {self._vmax} is not allowed"

    def __set_name__(self, owner, name):
        self.name = name

    def __get__(self, obj, objtype=None):
        if obj is None:
            return self
        store = getattr(obj, self._dict_name)
        if self.name not in store:
            store[self.name] = self._default
        value = store[self.name]
        self._check(value)
        return value

    def __set__(self, obj, value):
        self._check(value)
        getattr(obj, self._dict_name)[self.name] = value
        self._current_value = value


class SourceParameter(Parameter):
    def __init__(self, default=None, vmin=None, vmax=None):
        super(SourceParameter, self).__init__("_source_params", default, vmin, vmax)


class RequiredParameter(Parameter):
    def __init__(self, default=None, vmin=None, vmax=None):
        super(RequiredParameter, self).__init__("_required_params", default, vmin, vmax)


class GRBMeta(type):
    """Collects ``SourceParameter`` / ``RequiredParameter`` attributes into ``_parameter_names`` /
    ``_required_names`` (inherited from the bases first, then the class's own, in definition order)."""

    def __new__(mcls, name, bases, attrs):
        pnames, rnames = list(attrs.get("_parameter_names", [])), list(attrs.get("_required_names", []))
        for base in bases:
            pnames.extend(n for n in getattr(base, "_parameter_names", []) if n not in pnames)
            rnames.extend(n for n in getattr(base, "_required_names", []) if n not in rnames)
        for k, v in attrs.items():
            if isinstance(v, SourceParameter):
                v.name = k
                if k not in pnames:
                    pnames.append(k)
            elif isinstance(v, RequiredParameter):
                v.name = k
                if k not in rnames:
                    rnames.append(k)
        attrs["_parameter_names"], attrs["_required_names"] = pnames, rnames
        cls = super().__new__(mcls, name, bases, attrs)
        abstracts = {n for n, v in attrs.items() if getattr(v, "__isabstractmethod__", False)}
        for base in bases:
            for n in getattr(base, "__abstractmethods__", set()):
                if getattr(getattr(cls, n, None), "__isabstractmethod__", False):
                    abstracts.add(n)
        cls.__abstractmethods__ = frozenset(abstracts)
        return cls


__all__ = ["Parameter", "SourceParameter", "RequiredParameter", "GRBMeta"]
