"""Empty stand-in so that ``jaxincell/_plot.py`` imports (plotting is out of scope; TEST INFRASTRUCTURE ONLY)."""
import sys
import types


class _Anything:
    def __getattr__(self, name):
        return _Anything()

    def __call__(self, *a, **k):
        return _Anything()

    def __iter__(self):
        return iter(())


for _name in ("pyplot", "animation", "colors", "gridspec", "cm", "ticker"):
    _m = types.ModuleType(f"matplotlib.{_name}")
    _m.__getattr__ = lambda attr, _n=_name: type(attr, (), {"__init__": lambda self, *a, **k: None}) if attr[:1].isupper() else _Anything()
    sys.modules[f"matplotlib.{_name}"] = _m
    globals()[_name] = _m


def use(*a, **k):
    pass
