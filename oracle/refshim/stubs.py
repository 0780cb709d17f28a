"""Permissive placeholders for names the reference imports at module top level but never touches on
the generate -> act path (optax, distrax, orbax, tensorflow, unused flax/jax corners)."""
from __future__ import annotations

import importlib.abc
import importlib.machinery
import sys
import types


class Anything:
    """Absorbs attribute access, calls, subscripts and use as a base class / decorator argument.
    Touching it on the executed path is harmless only if the result is unused; arithmetic on it raises."""

    def __init__(self, label="stub"):
        object.__setattr__(self, "_label", label)

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return Anything(f"{self._label}.{name}")

    def __call__(self, *args, **kwargs):
        return Anything(f"{self._label}()")

    def __getitem__(self, item):
        return Anything(f"{self._label}[]")

    def __mro_entries__(self, bases):
        return (object,)

    def __or__(self, other):
        return self

    __ror__ = __or__

    def __iter__(self):
        return iter(())

    def __repr__(self):
        return f"<refshim stub {self._label}>"


class StubModule(types.ModuleType):
    def __init__(self, name):
        super().__init__(name)
        self.__path__ = []          # a package, so that sub-imports reach the finder

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return Anything(f"{self.__name__}.{name}")


class _Loader(importlib.abc.Loader):
    def create_module(self, spec):
        return StubModule(spec.name)

    def exec_module(self, module):
        pass


class StubFinder(importlib.abc.MetaPathFinder):
    """Serves a StubModule for any not-yet-registered module under the given top-level names."""

    def __init__(self, roots):
        self.roots = set(roots)

    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in self.roots and fullname not in sys.modules:
            return importlib.machinery.ModuleSpec(fullname, _Loader(), is_package=True)
        return None
