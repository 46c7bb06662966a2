"""In-repo stand-in for the ``cons`` package (not installed here).
Only ``cons.core.{ConsError,_car,_cdr}`` are used by the reference
(``aesara/graph/rewriting/unify.py:18``)."""
from .core import ConsError, car, cdr  # noqa: F401
