"""GEMM-epilogue fusion: ``Dot22``/``Gemm``/``Dot22Scalar`` followed by the one ``Elemwise``
node that consumes the product run as a single tcgen05 kernel whose epilogue applies the
Elemwise expression (``codegen/gemm_epilogue.py``, ``csrc/ab_gemm_tcgen05_kernel.cuh``).

In BASELINE config 3 that is ``tanh(X @ W1 + b1)``, ``(h @ W2 - Y) + b2`` and
``(dout @ W2.T) * (1 - h**2)``: three [B, H] round trips through HBM (and three re-reads
by the operand pack of the next product) that the reference's node-by-node execution pays
(``tensor/blas.py:872/1659`` then ``tensor/elemwise.py:835``).

Like ``rowfuse.RowFusion`` this is an executor-level region: the lowered program is not
changed; the GEMM node is deferred to the position of the Elemwise node and both run as one
launch.  Anything the fused kernel does not take (problem too small for the tensor-core
path, operands that do not broadcast over the [M, N] result, misaligned operands) runs node
by node as before.
"""

from __future__ import annotations

import os

import numpy as np

from ..codegen.gemm_epilogue import MAX_OPERANDS, gemm_epilogue_source
from . import lib as _lib
from .device import DeviceArray

GEMM_OPS = ("Dot22", "Gemm", "Dot22Scalar")


class EpilogueRequest:
    """What ``kernels.gemm`` needs to launch the fused variant (set on the executor by
    ``GemmEpilogueFusion.run`` for exactly one ``K.gemm`` call)."""

    def __init__(self, fusion, operands, want_shadow):
        self.fusion = fusion
        self.operands = operands          # DeviceArrays (2-D, broadcastable over [M, N])
        self.want_shadow = want_shadow
        self.applied = False
        self.shadow = None                # (torch buffer, pitch) when a bf16 plane was written


class GemmEpilogueFusion:
    def __init__(self, program, g, e, acc_input, operand_inputs, shadow_consumer):
        self.program = program
        self.g, self.e = g, e
        self.members = [g, e]
        self.first, self.last = g, e
        self.acc_input = acc_input
        self.operand_inputs = operand_inputs
        self.shadow_consumer = shadow_consumer  # the result feeds another GEMM as an operand
        self.broken = False
        self._src = None
        self._handle = None

    # ------------------------------------------------------------------ analysis
    @staticmethod
    def detect(program, destroys, taken=()):
        if os.environ.get("AB_NO_GEMM_FUSE"):
            return []
        nodes = program.nodes
        consumers = {}
        for i, n in enumerate(nodes):
            for v in n.inputs:
                consumers.setdefault(v, []).append(i)
        producer = {v: i for i, n in enumerate(nodes) for v in n.outputs}
        found, used = [], set(taken)
        for g, n in enumerate(nodes):
            if n.op not in GEMM_OPS or g in used:
                continue
            z = n.outputs[0]
            zv = program.vars[z]
            if zv.dtype != "float32" or zv.ndim != 2 or z in program.outputs:
                continue
            cons = consumers.get(z, [])
            if len(cons) != 1:
                continue
            e = cons[0]
            en = nodes[e]
            if en.op != "Elemwise" or e in used or len(en.outputs) != 1 or en.inputs.count(z) != 1:
                continue
            expr = en.params["expr"]
            ov = program.vars[en.outputs[0]]
            if ov.dtype != "float32" or ov.ndim != 2 or any(dt != "float32" for dt in expr["inputs"]):
                continue
            if len(en.inputs) - 1 > MAX_OPERANDS:
                continue
            if any(program.vars[v].kind != "tensor" or program.vars[v].ndim != 2 for v in en.inputs):
                continue
            # the GEMM is deferred to the Elemwise position: nothing in between may rewrite memory
            if any(destroys[i] for i in range(g + 1, e)):
                continue
            # Elemwise operands must exist before the (deferred) launch: always true, they
            # precede the Elemwise node; but they must not be produced from z (single consumer)
            acc_input = en.inputs.index(z)
            operand_inputs = [k for k in range(len(en.inputs)) if k != acc_input]
            out = en.outputs[0]
            shadow = False
            for c in consumers.get(out, []):
                cn = nodes[c]
                if cn.op in GEMM_OPS and out in _gemm_operands(cn):
                    shadow = True
                if cn.op == "DimShuffle" and list(cn.params.get("new_order", [])) == [1, 0]:
                    for c2 in consumers.get(cn.outputs[0], []):
                        if nodes[c2].op in GEMM_OPS and cn.outputs[0] in _gemm_operands(nodes[c2]):
                            shadow = True
            found.append(GemmEpilogueFusion(program, g, e, acc_input, operand_inputs, shadow))
            used.update((g, e))
        del producer
        return found

    # ------------------------------------------------------------------ execution
    def source(self):
        if self._src is None:
            expr = self.program.nodes[self.e].params["expr"]
            self._src = gemm_epilogue_source(expr, self.acc_input, self.operand_inputs)
        return self._src

    def compile_all(self):
        _lib.compile_cubin(self.source(), "gemm_ep")
        return 1

    def handle(self):
        if self._handle is None:
            self._handle = _lib.load_module(self.source(), "gemm_ep")
        return self._handle

    def run(self, ex, env):
        from .vm import _EXEC, _as_dev_inputs
        from ..ir import Node

        if self.broken:
            return False
        prog = self.program
        gn, en = prog.nodes[self.g], prog.nodes[self.e]
        gargs = [env[v] for v in gn.inputs]
        x, y = (gargs[0], gargs[1]) if gn.op != "Gemm" else (gargs[2], gargs[3])
        if not (isinstance(x, DeviceArray) and isinstance(y, DeviceArray)):
            return False
        M, N = x.shape[0], y.shape[1]
        op_vars = [en.inputs[k] for k in self.operand_inputs]
        ops = _as_dev_inputs(ex, self.e, Node("Elemwise", op_vars, []), [env[v] for v in op_vars])
        for a in ops:
            if a.dtype != np.float32 or a.ndim != 2 or a.shape[0] not in (1, M) or a.shape[1] not in (1, N):
                return False
        req = EpilogueRequest(self, ops, self.shadow_consumer and ex.precision == 2 and N % 8 == 0)
        ex._gemm_epilogue = req
        try:
            out = _EXEC[gn.op](ex, self.g, gn, gargs)
        except _lib.AbError:
            # the fused launch was refused (alignment, driver limits): never try again here
            self.broken = True
            ex._gemm_epilogue = None
            return False
        finally:
            pending, ex._gemm_epilogue = ex._gemm_epilogue, None
        if not req.applied:
            # the product took another path (SIMT / too small): finish with the plain Elemwise
            del pending
            env[gn.outputs[0]] = out
            eargs = [env[v] for v in en.inputs]
            env[en.outputs[0]] = _EXEC["Elemwise"](ex, self.e, en, eargs)
            env.pop(gn.outputs[0], None)
            return True
        env[en.outputs[0]] = out
        if req.shadow is not None:
            ex.pack_cache.adopt(out, *req.shadow)
        return True


def _gemm_operands(node):
    return (node.inputs[2], node.inputs[3]) if node.op == "Gemm" else (node.inputs[0], node.inputs[1])
