"""NVRTC-compile (no GPU needed) every kernel the committed fixtures use, so the
cubins in aesara_b200/_kcache/ travel with the tree."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from concurrent.futures import ThreadPoolExecutor

from aesara_b200.runtime.vm import ProgramExecutor
from aesara_b200.runtime import kernels as K
from tests._cases import case_names, load_case


def main():
    t = time.time()
    had = os.environ.get("AB_GEMM_FUSE_FP32")
    os.environ["AB_GEMM_FUSE_FP32"] = "1"  # also the (opt-in) regions of the fp32-faithful kernels: tests launch them
    try:
        _compile_all(t)
    finally:
        if had is None:
            del os.environ["AB_GEMM_FUSE_FP32"]
        else:
            os.environ["AB_GEMM_FUSE_FP32"] = had


def _compile_all(t):
    exs = [ProgramExecutor(load_case(n)[0]) for n in case_names()]
    # GEMM-epilogue regions are generated per precision policy (float-pair vs float32-tree sums)
    for n in case_names():
        prog = load_case(n)[0]
        if any(nd.op in ("Dot22", "Gemm", "Dot22Scalar") for nd in prog.nodes):
            exs += [ProgramExecutor(prog, precision=pr) for pr in (1, 2)]
    kerns = list(K.ElemwiseKernel._by_key.values()) + list(K.CAReduceKernel._by_key.values())
    for dt in ("float32", "float64", "int64", "int32", "int8", "bool"):
        kerns.append(K._identity_kernel(dt))
    fusions = [f for ex in exs for f in ex._fusions]  # row-region / GEMM-epilogue / map-reduce kernels
    with ThreadPoolExecutor(8) as pool:
        list(pool.map(lambda k: k.compile(), kerns))
        list(pool.map(lambda f: f.compile_all(), fusions))
    print(f"compiled {len(kerns)} modules + {len(fusions)} fused regions in {time.time() - t:.1f}s")


if __name__ == "__main__":
    main()
