"""Unification / reification with generator-style handlers.

Handlers registered on ``_unify``/``_reify`` may be ordinary functions or
generators.  A generator yields either another handler call (whose final value
is sent back in) or a plain value; the last plain value yielded is the result.
That is the protocol ``aesara/graph/rewriting/unify.py:141-245`` is written
against.
"""
from collections.abc import Mapping, Sequence, Set
from copy import copy
from types import GeneratorType

from _ab_mdispatch import Dispatcher
from .utils import transitive_get as walk
from .variable import Var, isvar

construction_sentinel = object()


def assoc(s, u, v):
    """Return a copy of substitution ``s`` extended with ``u -> v``."""
    if hasattr(s, "copy"):
        s = s.copy()
    else:
        s = copy(s)
    s[u] = v
    return s


def stream_eval(z):
    """Drive a (possibly nested) generator computation to its final value."""
    if not isinstance(z, GeneratorType):
        return z
    stack = [z]
    send, out = None, None
    while stack:
        g = stack[-1]
        try:
            out = g.send(send)
        except StopIteration:
            stack.pop()
            send = out
            continue
        if isinstance(out, GeneratorType):
            stack.append(out)
            send = None
        else:
            send = out
    return out


# ---------------------------------------------------------------- reify
_reify = Dispatcher("_reify")


@_reify.register(object, Mapping)
def _reify_object(o, s):
    return o


@_reify.register(Var, Mapping)
def _reify_Var(o, s):
    o_w = walk(o, s)
    if o_w is o:
        yield o_w
    else:
        yield _reify(o_w, s)


def _reify_Iterable_ctor(ctor, t, s):
    res = []
    for y in t:
        r = yield _reify(y, s)
        res.append(r)
    yield construction_sentinel
    yield ctor(res)


for _seq, _ctor in ((tuple, tuple), (list, list)):
    _reify.add(
        (_seq, Mapping), lambda t, s, _c=_ctor: _reify_Iterable_ctor(_c, t, s)
    )


def _reify_Mapping(o, s):
    res = {}
    for k, v in o.items():
        res[k] = yield _reify(v, s)
    yield construction_sentinel
    yield type(o)(res)


_reify.add((dict, Mapping), _reify_Mapping)


def reify(e, s):
    if len(s) == 0:
        return e
    return stream_eval(_reify(e, s))


# ---------------------------------------------------------------- unify
_unify = Dispatcher("_unify")


@_unify.register(object, object, Mapping)
def _unify_object(u, v, s):
    return s if u == v else False


@_unify.register(Var, (Var, object), Mapping)
def _unify_Var_object(u, v, s):
    u_w = walk(u, s)
    v_w = walk(v, s) if isvar(v) else v
    if u_w == v_w:
        yield s
    elif isvar(u_w):
        yield assoc(s, u_w, v_w)
    elif isvar(v_w):
        yield assoc(s, v_w, u_w)
    else:
        yield _unify(u_w, v_w, s)


_unify.add((object, Var, Mapping), _unify_Var_object)


def _unify_Sequence(u, v, s):
    if len(u) != len(v):
        yield False
        return
    for uu, vv in zip(u, v):
        s = yield _unify(uu, vv, s)
        if s is False:
            return
    yield s


for _seq in (tuple, list):
    _unify.add((_seq, _seq, Mapping), _unify_Sequence)


def _unify_Mapping(u, v, s):
    if len(u) != len(v):
        yield False
        return
    for key, uval in u.items():
        if key not in v:
            yield False
            return
        s = yield _unify(uval, v[key], s)
        if s is False:
            return
    yield s


_unify.add((dict, dict, Mapping), _unify_Mapping)


def unify(u, v, s=None):
    """Most-general unifier of ``u`` and ``v`` extending ``s``, or ``False``."""
    if s is None:
        s = {}
    if u is v:
        return s
    return stream_eval(_unify(u, v, s))
