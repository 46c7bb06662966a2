from collections import OrderedDict
from collections.abc import Sequence
from itertools import islice

from _ab_mdispatch import Dispatcher


class ConsError(ValueError):
    pass


_car = Dispatcher("car")
_cdr = Dispatcher("cdr")


def _car_seq(z):
    if len(z) == 0:
        raise ConsError("Not a cons pair")
    return z[0]


def _cdr_seq(z):
    if len(z) == 0:
        raise ConsError("Not a cons pair")
    return type(z)(list(islice(z, 1, None)))


for _t in (tuple, list):
    _car.add((_t,), _car_seq)
    _cdr.add((_t,), _cdr_seq)


def car(z):
    try:
        return _car(z)
    except NotImplementedError:
        raise ConsError("Not a cons pair")


def cdr(z):
    try:
        return _cdr(z)
    except NotImplementedError:
        raise ConsError("Not a cons pair")
