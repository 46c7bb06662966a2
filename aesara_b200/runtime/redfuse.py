"""Map-reduce fusion: a unary ``Elemwise`` node whose only consumer is a ``CAReduce`` runs
inside the reduction kernel (the mapped [B, H] array is never written).

The reference has the same rewrite for its C backend (``local_careduce_fusion``,
``aesara/tensor/rewriting/elemwise.py:943-1040``), but it is not part of the graphs the
default ``fast_run`` query produced for the BASELINE configs (App. A.3: ``Elemwise{Sqr}``
followed by ``Sum`` for the MSE loss), so it is applied here at executor level, like the
other regions (``rowfuse.py``, ``gemmfuse.py``): the lowered program is unchanged, the
Elemwise node is skipped and the CAReduce kernel is generated with the node's scalar
expression as its pre-map (``codegen/careduce.py``, ``AB_RED_PRE``).
"""

from __future__ import annotations

import os

from .device import DeviceArray


class ReducePreFusion:
    def __init__(self, program, e, r):
        self.program = program
        self.e, self.r = e, r
        self.members = [e, r]
        self.first, self.last = e, r
        self.broken = False
        self._kernel = None

    @staticmethod
    def detect(program, destroys, taken=()):
        if os.environ.get("AB_NO_RED_FUSE"):
            return []
        nodes = program.nodes
        consumers = {}
        for i, n in enumerate(nodes):
            for v in n.inputs:
                consumers.setdefault(v, []).append(i)
        producer = {v: i for i, n in enumerate(nodes) for v in n.outputs}
        found, used = [], set(taken)
        for r, n in enumerate(nodes):
            if n.op != "CAReduce" or r in used or not n.params.get("axis"):
                continue
            v = n.inputs[0]
            e = producer.get(v)
            if e is None or e in used or nodes[e].op != "Elemwise":
                continue
            en = nodes[e]
            expr = en.params["expr"]
            if len(en.inputs) != 1 or len(en.outputs) != 1 or consumers.get(v, []) != [r] or v in program.outputs:
                continue
            if program.vars[en.inputs[0]].ndim == 0 or expr["out_dtypes"][0] != n.params["in_dtype"]:
                continue
            if any(destroys[i] for i in range(e + 1, r)):
                continue
            found.append(ReducePreFusion(program, e, r))
            used.update((e, r))
        return found

    def kernel(self):
        if self._kernel is None:
            from . import kernels as K

            p = self.program.nodes[self.r].params
            expr = self.program.nodes[self.e].params["expr"]
            self._kernel = K.CAReduceKernel.get(p["scalar_op"], expr["inputs"][0], p["acc_dtype"],
                                                p["out_dtype"], pre_expr=expr)
        return self._kernel

    def compile_all(self):
        self.kernel().compile()
        return 1

    def run(self, ex, env):
        x = env[self.program.nodes[self.e].inputs[0]]
        p = self.program.nodes[self.r].params
        axis = tuple(p["axis"])
        if not isinstance(x, DeviceArray) or x.size == 0 or any(x.shape[a] == 0 for a in axis):
            return False  # host values / empty reductions keep their own code paths
        env[self.program.nodes[self.r].outputs[0]] = self.kernel().launch(x, axis)
        return True
