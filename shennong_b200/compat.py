"""Run code written for ``shennong`` on this engine without editing its imports

    import shennong_b200.compat
    shennong_b200.compat.install()          # from now on `import shennong...` resolves here
    from shennong.processor.mfcc import MfccProcessor
    from shennong import pipeline

``install()`` registers a meta-path finder that maps ``shennong`` and every
``shennong.<module>`` to ``shennong_b200`` / ``shennong_b200.<module>``.  Modules
of the reference that lie outside the frame-based feature path (alignment,
bottleneck, CREPE, UBM / VTLN training, ...) resolve to stubs whose attributes
raise ``NotImplementedError`` when used, so that a script importing them still
starts.  This is how the reference's own test files were run against the
engine (profiles/r01_reference_tests_probe.txt).

Nothing is installed on import; a real ``shennong`` package already imported
is left alone unless ``force=True``.
"""

import importlib
import importlib.abc
import importlib.util
import sys
import types

_TARGET = 'shennong_b200'
_ALIAS = 'shennong'


def _stub_module(name):
    """Placeholder for an out-of-scope module of the reference"""
    stub = types.ModuleType(name)
    stub.__path__ = []
    stub.__doc__ = f'{name}: not part of shennong_b200 (out of the feature path)'

    class _Missing:
        def __init__(self, *args, **kwargs):
            raise NotImplementedError(stub.__doc__)

        @classmethod
        def load(cls, *args, **kwargs):
            raise NotImplementedError(stub.__doc__)

    def __getattr__(attr):
        if attr.startswith('__'):
            raise AttributeError(attr)
        return _Missing

    stub.__getattr__ = __getattr__
    return stub


class _AliasFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, name, path=None, target=None):
        if name == _ALIAS or name.startswith(_ALIAS + '.'):
            return importlib.util.spec_from_loader(name, self)
        return None

    def create_module(self, spec):
        real = _TARGET + spec.name[len(_ALIAS):]
        try:
            return importlib.import_module(real)
        except ModuleNotFoundError as err:
            if err.name is not None and not err.name.startswith(_TARGET):
                raise          # a genuine missing dependency of an in-scope module
            return _stub_module(spec.name)

    def exec_module(self, module):
        pass


_finder = None


def install(force=False):
    """Makes ``import shennong`` resolve to this engine; returns True if done"""
    global _finder
    if _finder is not None:
        return True
    present = sys.modules.get(_ALIAS)
    if present is not None and getattr(present, '__name__', '') != _TARGET:
        if not force:
            return False
        for key in [k for k in sys.modules
                    if k == _ALIAS or k.startswith(_ALIAS + '.')]:
            del sys.modules[key]
    package = importlib.import_module(_TARGET)
    _finder = _AliasFinder()
    sys.meta_path.insert(0, _finder)
    sys.modules[_ALIAS] = package
    for key, module in list(sys.modules.items()):
        if key.startswith(_TARGET + '.'):
            sys.modules[_ALIAS + key[len(_TARGET):]] = module
    return True


def uninstall():
    """Removes the alias (modules already imported under it stay bound)"""
    global _finder
    if _finder is None:
        return
    sys.meta_path.remove(_finder)
    _finder = None
    for key in [k for k in sys.modules
                if k == _ALIAS or k.startswith(_ALIAS + '.')]:
        del sys.modules[key]
