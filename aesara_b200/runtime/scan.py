"""``Scan`` on device: the general step loop.

Reference: ``aesara/scan/op.py:1673-2160`` (``Scan.perform``; Cython twin
``scan_perform.pyx:76-602``).  The recurrent outputs are circular buffers of
``store_steps`` rows living in device memory; each step gathers the tap rows
as *views*, runs the lowered inner program (itself a ``ProgramExecutor`` —
``Scan.make_thunk`` compiles its inner graph with the same linker,
``op.py:1431-1459``) and writes the results into the row at ``pos``.  No data
leaves the device between steps; the only host work per step is launching.
"""

from __future__ import annotations

import os

import numpy as np

from . import kernels as K
from .device import DeviceArray


WHILE_POLL = 8  # an `until` condition is read back every this many steps when steps may run ahead


class ScanRunner:
    def __init__(self, node, parent):
        from .vm import ProgramExecutor

        self.node = node
        self.parent = parent
        p = node.params
        self.info = info = p["info"]
        self.inner = ProgramExecutor(p["inner"], precision=parent.precision, host_outputs=False)
        self.destroy = {int(k) for k in p.get("destroy_map", {})}
        self.mm_in = info["mit_mot_in_slices"]
        self.mm_out = info["mit_mot_out_slices"]
        self.tap_array = self.mm_in + info["mit_sot_in_slices"] + info["sit_sot_in_slices"]
        self.n_mit_mot = len(self.mm_in)
        self.n_outs = len(self.tap_array)
        self.n_nit_sot = info["n_nit_sot"]
        self.n_shared = info["n_shared_outs"]
        self.n_seqs = info["n_seqs"]
        self.mintaps = [min(t) for t in self.tap_array] + [0] * self.n_nit_sot
        # fast path: "one Gemm with a loop-invariant matrix + Elemwise on its column slices"
        # (LSTM, tanh-RNN, gated units) -> one persistent kernel (runtime/scan_cell.py)
        self.cell = None
        if (self.n_seqs == 1 and not self.mm_in and 1 <= len(self.tap_array) <= 3
                and all(list(t) == [-1] for t in self.tap_array)
                and self.n_nit_sot == 0 and self.n_shared == 0 and info["n_non_seqs"] == 1
                and not info["as_while"] and parent.precision == 0
                and not os.environ.get("AB_SCAN_NO_FAST")):
            from .scan_cell import CellMatch

            self.cell = CellMatch(self.inner, len(self.tap_array))
        self.used_fast_path = False
        self.fast_path_kind = None   # "lstm" (ahead-of-time cell) or "jit" (generated cell)
        self.direct_writes = 0
        self.while_polls = 0

    def run(self, args):
        info = self.info
        ex = self.parent
        prog = ex.program
        node = self.node
        n_steps = int(np.asarray(args[0].to_numpy() if isinstance(args[0], DeviceArray) else args[0]).item())
        if n_steps < 0:
            raise IndexError(f"Scan was asked to run for negative number of step {n_steps}")
        n_seqs, n_outs, n_nit, n_sh = self.n_seqs, self.n_outs, self.n_nit_sot, self.n_shared

        def dev(v, k):
            if isinstance(v, DeviceArray):
                return v
            return ex.dev(v, dtype=prog.vars[node.inputs[k]].dtype)

        seqs = [dev(a, 1 + k) for k, a in enumerate(args[1 : 1 + n_seqs])]
        for idx, s in enumerate(seqs):
            if s.shape[0] < n_steps:
                raise ValueError(
                    f"Sequence {idx} has shape {s.shape} but the Scan's required number of steps is {n_steps}"
                )
        o0 = 1 + n_seqs
        states = [dev(a, o0 + k) for k, a in enumerate(args[o0 : o0 + n_outs])]
        shared_vals = list(args[o0 + n_outs : o0 + n_outs + n_sh])
        nit_len = [int(np.asarray(a.to_numpy() if isinstance(a, DeviceArray) else a).item())
                   for a in args[o0 + n_outs + n_sh : o0 + n_outs + n_sh + n_nit]]
        non_seqs = list(args[o0 + n_outs + n_sh + n_nit :])

        store_steps = [s.shape[0] for s in states] + nit_len
        bufs = []
        for idx, s in enumerate(states):
            if idx in self.destroy and isinstance(args[o0 + idx], DeviceArray):
                bufs.append(s)
            else:
                bufs.append(K.contiguous_copy(s))
        nit_bufs = [None] * n_nit
        out_vars = [prog.vars[v] for v in node.outputs]
        if n_steps == 0:
            return bufs + [
                DeviceArray.empty((0,) * out_vars[n_outs + j].ndim, out_vars[n_outs + j].dtype)
                for j in range(n_nit)
            ] + shared_vals

        pos = [(-self.mintaps[idx]) % store_steps[idx] for idx in range(n_outs + n_nit)]
        i, cond = 0, True
        self.used_fast_path = False
        self.fast_path_kind = None
        if self.cell is not None and len(non_seqs) == 1:
            from . import scan_cell

            U = dev(non_seqs[0], len(args) - 1)
            H = bufs[0].shape[2] if bufs[0].ndim == 3 else 0
            gates = U.shape[1] // H if (U.ndim == 2 and H and U.shape[1] % H == 0) else 0
            spec = None
            if gates and scan_cell.eligible(seqs[0], U, bufs, n_steps, gates):
                spec = self.cell.match(bufs[0].shape[1], H, gates)
            if spec is not None:
                if spec.is_lstm and not os.environ.get("AB_SCAN_JIT"):
                    scan_cell.run_lstm(n_steps, seqs[0], U, bufs[0], bufs[1], pos[0], pos[1])
                    self.fast_path_kind = "lstm"
                else:
                    scan_cell.run_cell(spec, n_steps, seqs[0], U, bufs, pos[:n_outs])
                    self.fast_path_kind = "jit"
                i = n_steps
                pos = [(p + n_steps) % s for p, s in zip(pos, store_steps)]
                self.used_fast_path = True
        # run ahead of an `until` condition only when no ring overwrites rows that matter
        speculate = bool(info["as_while"]) and n_sh == 0 and all(
            store_steps[idx] >= n_steps - self.mintaps[idx] for idx in range(n_outs)) and all(
            nl >= n_steps for nl in nit_len)
        flags, polled = None, i
        while i < n_steps and cond:
            inner_in = [s.index((i,)) for s in seqs]
            for idx, taps in enumerate(self.tap_array):
                for t in taps:
                    inner_in.append(bufs[idx].index(((pos[idx] + t) % store_steps[idx],)))
            inner_in += shared_vals
            inner_in += non_seqs
            # sit-sot / mit-sot results go straight into their ring rows when the row does not
            # overlap anything this step still reads (scan_perform.pyx copies, op.py:2005-2040)
            k0 = sum(len(self.mm_out[g]) for g in range(self.n_mit_mot))
            want = {}
            for j in range(self.n_mit_mot, n_outs):
                row = bufs[j].index((pos[j],))
                if row.is_c_contiguous() and not any(_overlaps(row, a) for a in inner_in):
                    want[k0 + j - self.n_mit_mot] = row
            inner_out = self.inner(*inner_in, out_storage=want)
            k = 0
            for g in range(self.n_mit_mot):
                for out_slice in self.mm_out[g]:
                    K.copy_into(bufs[g].index((out_slice + pos[g],)), _d(inner_out[k]))
                    k += 1
            for j in range(self.n_mit_mot, n_outs):
                if inner_out[k] is not want.get(k):
                    K.copy_into(bufs[j].index((pos[j],)), _d(inner_out[k]))
                else:
                    self.direct_writes += 1
                k += 1
            for j in range(n_nit):
                val = _d(inner_out[k])
                if i == 0:
                    nit_bufs[j] = DeviceArray.empty((store_steps[n_outs + j],) + val.shape,
                                                    out_vars[n_outs + j].dtype)
                K.copy_into(nit_bufs[j].index((pos[n_outs + j],)), val)
                k += 1
            for j in range(n_sh):
                shared_vals[j] = inner_out[k]
                k += 1
            pos = [(p + 1) % s for p, s in zip(pos, store_steps)]
            i += 1
            if info["as_while"]:
                # `until`: the reference reads the condition after every step (op.py:2041-2046),
                # a device->host synchronisation per step here.  When no output ring wraps (every
                # row of every step has its own slot) steps may run ahead: the conditions are
                # collected on the device and read back every WHILE_POLL steps; rows written past
                # the stopping step are cut off below exactly like the unused tail of the rings.
                c = inner_out[k]
                if not isinstance(c, DeviceArray):
                    cond = not bool(np.asarray(c).item())
                elif not speculate:
                    cond = not bool(c.item())
                else:
                    if flags is None:
                        flags = DeviceArray.empty((n_steps,), c.dtype)
                    K.copy_into(flags.index((slice(i - 1, i),)), c.reshape_view((1,)) if c.ndim == 0 else c)
                    if i - polled >= WHILE_POLL or i == n_steps:
                        got = flags.index((slice(polled, i),)).to_numpy().astype(bool)
                        self.while_polls += 1
                        if got.any():
                            stop = polled + int(np.argmax(got)) + 1
                            pos = [(p - (i - stop)) % s for p, s in zip(pos, store_steps)]
                            i = stop
                            cond = False
                        polled = i

        allb = bufs + nit_bufs
        for idx in range(self.n_mit_mot, n_outs + n_nit):
            st = store_steps[idx]
            b = allb[idx]
            if st < i - self.mintaps[idx] and pos[idx] < st and pos[idx] != 0:
                # un-rotate the circular buffer (op.py:2105-2134)
                pdx = pos[idx]
                tmp = K.contiguous_copy(b)
                K.copy_into(b.index((slice(0, st - pdx),)), tmp.index((slice(pdx, st),)))
                K.copy_into(b.index((slice(st - pdx, st),)), tmp.index((slice(0, pdx),)))
            elif st > i - self.mintaps[idx]:
                tail = b.index((slice(i - self.mintaps[idx], st),))
                if tail.size:
                    zero = DeviceArray.from_numpy(np.zeros((1,) * b.ndim, dtype=b.dtype))
                    K.copy_into(tail, zero)
                if i < n_steps:
                    allb[idx] = b.index((slice(0, st - (n_steps - i)),))
        return allb + shared_vals


def _overlaps(a, b):
    """Conservative byte-range overlap of two device arrays (host values never overlap)."""
    if not isinstance(a, DeviceArray) or not isinstance(b, DeviceArray) or a.owner is not b.owner:
        return False

    def span(x):
        lo = hi = 0
        for n, s in zip(x.shape, x.strides):
            if n == 0:
                return x.ptr, x.ptr
            if s >= 0:
                hi += (n - 1) * s
            else:
                lo += (n - 1) * s
        return x.ptr + lo * x.itemsize, x.ptr + (hi + 1) * x.itemsize

    a0, a1 = span(a)
    b0, b1 = span(b)
    return a0 < b1 and b0 < a1


def _d(v):
    if isinstance(v, DeviceArray):
        return v
    return DeviceArray.from_numpy(np.asarray(v))
