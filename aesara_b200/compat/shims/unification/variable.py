"""Logic variables (``Var``) — see ``aesara/graph/rewriting/unify.py:36-70``
which subclasses ``Var`` and relies on ``Var._id``, ``Var._refs`` and
``.token``."""
from contextlib import contextmanager

_glv = set()  # values temporarily treated as logic variables


class Var:
    __slots__ = ("token", "__weakref__")
    _id = 1
    _refs = {}

    def __new__(cls, token=None, prefix=""):
        if token is None:
            token = f"{prefix}_{Var._id}"
            Var._id += 1
        obj = cls._refs.get(token, None)
        if obj is None:
            obj = object.__new__(cls)
            obj.token = token
            cls._refs[token] = obj
        return obj

    def __str__(self):
        return f"~{self.token}"

    __repr__ = __str__

    def __eq__(self, other):
        if type(self) == type(other):
            return self.token == other.token
        return NotImplemented

    def __hash__(self):
        return hash((type(self), self.token))


def var(*args, **kwargs):
    return Var(*args, **kwargs)


def vars(n, **kwargs):
    return [var(**kwargs) for _ in range(n)]


def isvar(o):
    if isinstance(o, Var):
        return True
    if _glv:
        try:
            return o in _glv
        except TypeError:
            return False
    return False


@contextmanager
def variables(*values):
    old = _glv.copy()
    _glv.update(values)
    try:
        yield
    finally:
        _glv.clear()
        _glv.update(old)
