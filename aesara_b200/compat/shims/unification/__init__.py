"""In-repo stand-in for the ``unification`` package (not installed here).

API surface fixed by the reference: ``aesara/graph/rewriting/unify.py:17-23``
and ``aesara/graph/rewriting/basic.py:1620-1650``.
"""
from .core import assoc, reify, unify  # noqa: F401
from .variable import Var, isvar, var, variables, vars  # noqa: F401
