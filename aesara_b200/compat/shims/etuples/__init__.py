"""In-repo stand-in for the ``etuples`` package (not installed here).
Surface used by the reference: ``etuples.{apply,etuple,etuplize}`` and
``etuples.core.ExpressionTuple(.evaled_obj)``
(``aesara/graph/rewriting/unify.py:19-20``, ``rewriting/basic.py:1620``)."""
from .core import ExpressionTuple, apply, etuple, etuplize, rands, rator  # noqa: F401
