"""The lowered program: what `B200Linker` hands to the device runtime.

A *program* is the optimised ``FunctionGraph`` (``aesara/graph/fg.py:37``)
re-expressed as plain data — no Aesara objects — so that it can be executed,
serialised to JSON (``tests/golden/*.json``) and replayed on a machine where
the front-end is not installed.  One program node corresponds to exactly one
``Apply`` node of the optimised graph ("whatever survives
``fast_run``").

Layout
------
``Program.vars``      list of ``Var`` (index = variable id)
``Program.inputs``    variable ids, in ``fgraph.inputs`` order
``Program.outputs``   variable ids, in ``fgraph.outputs`` order
``Program.updates``   ``[(output_index, input_index)]`` from ``fgraph.update_mapping``
``Program.nodes``     list of ``Node`` in execution (toposort) order

``Var.kind``: ``"tensor"`` (TensorType), ``"scalar"`` (aesara ``ScalarType``
variables — host-side shape arithmetic), ``"other"`` (slices, None…).

A scalar expression (the body of an ``Elemwise``/``Composite``) is a list of
three-address statements over typed temporaries::

    {"inputs": ["float32", "float32"],          # dtypes of i0, i1, ...
     "stmts":  [{"op": "tanh", "args": ["i0"], "dtype": "float32"},   # -> t0
                {"op": "add",  "args": ["t0", "i1"], "dtype": "float32"}],
     "outputs": ["t1"]}

``args`` entries are ``"iK"`` (input), ``"tK"`` (result of statement K) or a
constant ``{"const": value, "dtype": ...}``.
"""

from __future__ import annotations

import base64
import json
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Tuple

import numpy as np

IR_VERSION = 1


@dataclass
class Var:
    dtype: Optional[str]  # numpy dtype name, None for non-array values
    ndim: int = 0
    kind: str = "tensor"  # "tensor" | "scalar" | "other"
    static_shape: Optional[Tuple[Optional[int], ...]] = None
    name: Optional[str] = None
    const: Optional[np.ndarray] = None  # value for Constant variables
    const_other: Any = None  # non-array constant payload (e.g. None, slice)

    def to_json(self):
        d = {"dtype": self.dtype, "ndim": self.ndim, "kind": self.kind}
        if self.static_shape is not None:
            d["static_shape"] = list(self.static_shape)
        if self.name:
            d["name"] = self.name
        if self.const is not None:
            a = np.asarray(self.const)
            d["const"] = {
                "dtype": a.dtype.name,
                "shape": list(a.shape),
                "b64": base64.b64encode(a.tobytes(order="C")).decode("ascii"),
            }
        if self.const_other is not None:
            d["const_other"] = self.const_other
        return d

    @staticmethod
    def from_json(d):
        const = None
        if "const" in d:
            c = d["const"]
            const = np.frombuffer(
                base64.b64decode(c["b64"]), dtype=np.dtype(c["dtype"])
            ).reshape(tuple(c["shape"])).copy()
        ss = d.get("static_shape")
        return Var(
            dtype=d["dtype"],
            ndim=d["ndim"],
            kind=d.get("kind", "tensor"),
            static_shape=tuple(ss) if ss is not None else None,
            name=d.get("name"),
            const=const,
            const_other=d.get("const_other"),
        )


@dataclass
class Node:
    op: str
    inputs: List[int]
    outputs: List[int]
    params: Dict[str, Any] = field(default_factory=dict)
    # filled by the lowering for error messages (str(apply_node))
    label: Optional[str] = None

    def to_json(self):
        d = {"op": self.op, "inputs": self.inputs, "outputs": self.outputs}
        if self.params:
            d["params"] = _params_to_json(self.params)
        if self.label:
            d["label"] = self.label
        return d

    @staticmethod
    def from_json(d):
        return Node(
            op=d["op"],
            inputs=list(d["inputs"]),
            outputs=list(d["outputs"]),
            params=_params_from_json(d.get("params", {})),
            label=d.get("label"),
        )


def _params_to_json(p):
    out = {}
    for k, v in p.items():
        if isinstance(v, Program):
            out[k] = {"__program__": v.to_json()}
        elif isinstance(v, tuple):
            out[k] = list(v)
        else:
            out[k] = v
    return out


def _params_from_json(p):
    out = {}
    for k, v in p.items():
        if isinstance(v, dict) and "__program__" in v:
            out[k] = Program.from_json(v["__program__"])
        else:
            out[k] = v
    return out


@dataclass
class Program:
    vars: List[Var] = field(default_factory=list)
    inputs: List[int] = field(default_factory=list)
    outputs: List[int] = field(default_factory=list)
    updates: List[Tuple[int, int]] = field(default_factory=list)
    nodes: List[Node] = field(default_factory=list)
    name: Optional[str] = None

    # ------------------------------------------------------------------
    def to_json(self):
        return {
            "ir_version": IR_VERSION,
            "name": self.name,
            "vars": [v.to_json() for v in self.vars],
            "inputs": self.inputs,
            "outputs": self.outputs,
            "updates": [list(u) for u in self.updates],
            "nodes": [n.to_json() for n in self.nodes],
        }

    @staticmethod
    def from_json(d):
        if d.get("ir_version", IR_VERSION) != IR_VERSION:
            raise ValueError(f"unsupported program IR version {d.get('ir_version')}")
        return Program(
            vars=[Var.from_json(v) for v in d["vars"]],
            inputs=list(d["inputs"]),
            outputs=list(d["outputs"]),
            updates=[tuple(u) for u in d.get("updates", [])],
            nodes=[Node.from_json(n) for n in d["nodes"]],
            name=d.get("name"),
        )

    def dumps(self, **kw):
        return json.dumps(self.to_json(), **kw)

    @staticmethod
    def loads(s):
        return Program.from_json(json.loads(s))

    def save(self, path):
        with open(path, "w") as f:
            f.write(self.dumps(indent=1))

    @staticmethod
    def load(path):
        with open(path) as f:
            return Program.loads(f.read())

    # ------------------------------------------------------------------
    def op_counts(self):
        out: Dict[str, int] = {}
        for n in self.nodes:
            out[n.op] = out.get(n.op, 0) + 1
        return out

    def summary(self):
        lines = [f"program {self.name or ''}: {len(self.nodes)} nodes"]
        for i, n in enumerate(self.nodes):
            lines.append(f"  {i:3d} {n.op:<14s} {n.inputs} -> {n.outputs}  {n.label or ''}")
        return "\n".join(lines)


# ---------------------------------------------------------------------------
# dtype helpers shared by code generators and the oracle
# ---------------------------------------------------------------------------
FLOAT_DTYPES = ("float16", "float32", "float64")
INT_DTYPES = ("int8", "int16", "int32", "int64")
UINT_DTYPES = ("uint8", "uint16", "uint32", "uint64")
DISCRETE_DTYPES = ("bool",) + INT_DTYPES + UINT_DTYPES
COMPLEX_DTYPES = ("complex64", "complex128")

# element type names in device code (see csrc/ab_types.cuh)
CTYPE = {
    "bool": "ab_bool",
    "int8": "ab_i8",
    "int16": "ab_i16",
    "int32": "ab_i32",
    "int64": "ab_i64",
    "uint8": "ab_u8",
    "uint16": "ab_u16",
    "uint32": "ab_u32",
    "uint64": "ab_u64",
    "float32": "float",
    "float64": "double",
}

# integer codes used across the C ABI (include/aesara_b200.h: ab_dtype)
DTYPE_CODE = {
    "bool": 0,
    "int8": 1,
    "int16": 2,
    "int32": 3,
    "int64": 4,
    "uint8": 5,
    "uint16": 6,
    "uint32": 7,
    "uint64": 8,
    "float16": 9,
    "float32": 10,
    "float64": 11,
    "bfloat16": 12,
}


def itemsize(dtype: str) -> int:
    return np.dtype(dtype).itemsize
