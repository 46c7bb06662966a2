"""Expression tuples: ``(operator, *operands)`` tuples that remember (or
lazily compute) the object they evaluate to."""
from collections.abc import Mapping, Sequence

from cons.core import ConsError, _car, _cdr
from unification.core import _reify, _unify, construction_sentinel
from unification.variable import Var, isvar

from _ab_mdispatch import Dispatcher

_NO_VALUE = object()

apply = Dispatcher("apply")


@apply.register(object, object)
def _apply_default(rator, rands):
    if not callable(rator):
        raise NotImplementedError(f"{rator!r} is not callable")
    return rator(*rands)


def rator(x):
    return _car(x)


def rands(x):
    return _cdr(x)


class ExpressionTuple(Sequence):
    __slots__ = ("_evaled_obj", "_tuple", "_parent")

    def __init__(self, seq=None, **kwargs):
        if seq is not None and not (
            isinstance(seq, ExpressionTuple) and "evaled_obj" not in kwargs
        ):
            self._tuple = tuple(seq)
            self._evaled_obj = kwargs.pop("evaled_obj", _NO_VALUE)
        elif isinstance(seq, ExpressionTuple):
            self._tuple = seq._tuple
            self._evaled_obj = seq._evaled_obj
        else:
            self._tuple = ()
            self._evaled_obj = kwargs.pop("evaled_obj", _NO_VALUE)
        self._parent = None

    # evaluation ---------------------------------------------------------
    @property
    def evaled_obj(self):
        if self._evaled_obj is _NO_VALUE:
            if len(self._tuple) == 0:
                raise ValueError("cannot evaluate an empty expression tuple")
            args = [
                a.evaled_obj if isinstance(a, ExpressionTuple) else a
                for a in self._tuple
            ]
            op, operands = args[0], args[1:]
            self._evaled_obj = apply(op, ExpressionTuple(operands))
        return self._evaled_obj

    @evaled_obj.setter
    def evaled_obj(self, obj):
        raise ValueError("Value of evaluated expression cannot be set!")

    # sequence protocol ----------------------------------------------------
    def __getitem__(self, key):
        res = self._tuple[key]
        if isinstance(key, slice):
            et = ExpressionTuple(res)
            if key == slice(None):
                et._evaled_obj = self._evaled_obj
            return et
        return res

    def __len__(self):
        return len(self._tuple)

    def __iter__(self):
        return iter(self._tuple)

    def __contains__(self, item):
        return item in self._tuple

    def __add__(self, x):
        return ExpressionTuple(self._tuple + tuple(x))

    def __radd__(self, x):
        return ExpressionTuple(tuple(x) + self._tuple)

    def __eq__(self, other):
        if isinstance(other, ExpressionTuple):
            return self._tuple == other._tuple
        if isinstance(other, tuple):
            return self._tuple == other
        return NotImplemented

    def __hash__(self):
        return hash(self._tuple)

    def __repr__(self):
        return f"ExpressionTuple({self._tuple!r})"

    def __str__(self):
        return f"e({', '.join(str(i) for i in self._tuple)})"


def etuple(*args, **kwargs):
    return ExpressionTuple(args, **kwargs)


def etuplize(x, shallow=False, return_bad_args=False, convert_ConsPairs=True):
    """Turn an object that has a ``car``/``cdr`` decomposition into an
    ``ExpressionTuple``.  Raises ``TypeError`` when ``x`` has none (unless
    ``return_bad_args``)."""
    if isinstance(x, ExpressionTuple):
        return x
    try:
        op, args = _car(x), _cdr(x)
    except (ConsError, NotImplementedError):
        op, args = None, None
    if not callable(op) or not isinstance(args, (list, tuple, ExpressionTuple)):
        if return_bad_args:
            return x
        raise TypeError(f"x is neither a non-str Sequence nor term: {type(x)}")
    if shallow:
        et_op, et_args = op, args
    else:
        et_op = etuplize(op, return_bad_args=True)
        et_args = tuple(etuplize(a, return_bad_args=True) for a in args)
    return etuple(et_op, *et_args, evaled_obj=x)


# car/cdr of an expression tuple itself
def _car_et(z):
    if len(z) == 0:
        raise ConsError("Not a cons pair")
    return z[0]


def _cdr_et(z):
    if len(z) == 0:
        raise ConsError("Not a cons pair")
    return z[1:]


_car.add((ExpressionTuple,), _car_et)
_cdr.add((ExpressionTuple,), _cdr_et)


# unification hooks ---------------------------------------------------------
def _unify_ExpressionTuple(u, v, s):
    return _unify(getattr(u, "_tuple", u), getattr(v, "_tuple", v), s)


_unify.add((ExpressionTuple, ExpressionTuple, Mapping), _unify_ExpressionTuple)
_unify.add((tuple, ExpressionTuple, Mapping), _unify_ExpressionTuple)
_unify.add((ExpressionTuple, tuple, Mapping), _unify_ExpressionTuple)


def _reify_ExpressionTuple(u, s):
    res = yield _reify(u._tuple, s)
    yield construction_sentinel
    same = len(res) == len(u) and all(
        (a is b)
        or (
            not isinstance(a, (Var, ExpressionTuple))
            and not isinstance(b, (Var, ExpressionTuple))
            and _safe_eq(a, b)
        )
        for a, b in zip(u, res)
    )
    yield u if same else ExpressionTuple(res)


def _safe_eq(a, b):
    try:
        return bool(a == b)
    except Exception:
        return False


_reify.add((ExpressionTuple, Mapping), _reify_ExpressionTuple)
