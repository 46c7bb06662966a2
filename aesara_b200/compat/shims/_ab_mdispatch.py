"""Minimal multiple-dispatch registry used by the ``cons`` / ``etuples`` /
``unification`` shims.

The reference front-end (``aesara/graph/rewriting/unify.py:17-23, 72-245``)
expects ``_car``, ``_cdr``, ``apply``, ``_unify`` and ``_reify`` to be
dispatcher objects exposing ``.add(signature, fn)`` and ``.register(*types)``.
Those third-party packages are not installed in this image (and there is no
network), so this file provides just enough of that protocol.

Resolution rule: among all registered signatures matching the call's argument
types, keep the ones that are not strictly less specific than another match;
if several remain (an ambiguity), the one registered last wins.
"""

from itertools import product


class Dispatcher:
    def __init__(self, name):
        self.name = name
        self._entries = []  # (signature tuple of types, fn, registration index)
        self._cache = {}

    # -- registration -----------------------------------------------------
    def add(self, signature, fn):
        signature = tuple(signature)
        # a tuple inside a signature is a union: expand into all combinations
        choices = [s if isinstance(s, tuple) else (s,) for s in signature]
        for sig in product(*choices):
            self._entries = [e for e in self._entries if e[0] != sig]
            self._entries.append((sig, fn, len(self._entries)))
        self._entries = [(s, f, i) for i, (s, f, _) in enumerate(self._entries)]
        self._cache.clear()

    def register(self, *signature):
        def deco(fn):
            self.add(signature, fn)
            return fn

        return deco

    # -- lookup -------------------------------------------------------------
    @staticmethod
    def _supersedes(a, b):
        """True when signature ``a`` is at least as specific as ``b``."""
        return len(a) == len(b) and all(issubclass(x, y) for x, y in zip(a, b))

    def dispatch(self, *types):
        try:
            return self._cache[types]
        except KeyError:
            pass
        matches = [
            e
            for e in self._entries
            if len(e[0]) == len(types)
            and all(issubclass(t, s) for t, s in zip(types, e[0]))
        ]
        best = None
        if matches:
            minimal = [
                e
                for e in matches
                if not any(
                    o is not e
                    and self._supersedes(o[0], e[0])
                    and not self._supersedes(e[0], o[0])
                    for o in matches
                )
            ]
            best = max(minimal, key=lambda e: e[2])[1]
        self._cache[types] = best
        return best

    def __call__(self, *args):
        fn = self.dispatch(*[type(a) for a in args])
        if fn is None:
            raise NotImplementedError(
                f"no {self.name} implementation for "
                f"({', '.join(type(a).__name__ for a in args)})"
            )
        return fn(*args)

    def __repr__(self):
        return f"<dispatched {self.name}>"
