#!/usr/bin/env python
"""Benchmark of the hot path (BASELINE.json metric): graph-evals/s of an
optimised Aesara graph executed by the B200 backend, with the roofline of its
dominant kernels and the CPU baseline timed beside it.

    python bench.py --gpus 1 --steps 10 --warmup 3                 # headline: MLP fwd+grad
    python bench.py --workload elemwise                            # fused Elemwise config
    python bench.py --impl reference                               # CPU arm (oracle port)
    torchrun --nproc-per-node N bench.py --gpus N ...              # weak scaling over B

A "step" is one evaluation of the compiled graph on one batch of synthetic
inputs.  `value` is measured with inputs resident in HBM; `e2e` goes through the
public host API (pinned host inputs -> H2D -> graph -> D2H of every output).
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
PRECISIONS = {"fp32": 0, "tf32": 1, "bf16": 2}


# ----------------------------------------------------------------------------- workloads
def workload_spec(name, args):
    if name == "mlp":
        B, H = args.batch or 65536, args.hidden or 4096
        return dict(
            name="mlp", program="cfg3_mlp", B=B, H=H,
            desc=f"cfg3: 2-layer tanh MLP fwd+grad (MSE), batch {B} x hidden {H}, f32 graph",
            gemm_flops=5 * 2.0 * B * H * H,            # 5 GEMMs x 2BH^2
            n_gemm=5,
            elemwise_bytes=14.0 * B * H * 4,           # 14*B*H*s, s=4 (f32 graph)
        )
    if name == "elemwise":
        n = args.n or (1 << 28)
        return dict(
            name="elemwise", program="cfg2_fused", n=n,
            desc=f"cfg2: fused Elemwise softplus(tanh(x)+y)*z on 3x{n} f32",
            elemwise_bytes=16.0 * n,                   # 4 arrays x N x 4 B
            gemm_flops=0.0, n_gemm=0,
        )
    if name == "lstm":
        T, B, H = args.steps_t or 128, args.batch or 8192, args.hidden or 1024
        return dict(
            name="lstm", program="cfg4_lstm", T=T, B=B, H=H, graph=True,
            desc=f"cfg4: Scan LSTM cell, T={T} steps, batch {B}, hidden {H}, f32",
            gemm_flops=2.0 * T * B * H * 4 * H, n_gemm=T,      # 8.80 TFLOP
            elemwise_bytes=(T * B * 4 * H + 4 * B * H + H * 4 * H) * 4.0,  # 17.33 GB ideal
        )
    if name == "logreg":
        N, D = args.n or (1 << 24), args.hidden or 512
        return dict(
            name="logreg", program="cfg5_logreg", N=N, D=D, graph=True,
            desc=f"cfg5: logistic-regression cost+grad, {N} rows x {D} f32 per GPU",
            gemm_flops=0.0, n_gemm=0,
            elemwise_bytes=(2.0 * N * D + 12.0 * N) * 4,      # SURVEY §8d graph-as-optimised
            # single-pass row-region fusion (runtime/rowfuse.py): X once, y read by `1 - y` and
            # by the fused kernel, `1 - y` written and read back
            fused_bytes=(1.0 * N * D + 4.0 * N) * 4,
        )
    if name == "readme":
        n = args.n or 1000
        return dict(
            name="readme", program="cfg1_readme", n=n, graph=True,
            desc=f"cfg1: README a/a + (M+a).dot(v), {n}x{n} f64",
            gemm_flops=0.0, n_gemm=0, elemwise_bytes=24.0 * n * n,
        )
    raise SystemExit(f"unknown workload {name}")


def make_inputs_numpy(spec, rng, scale_rows=None):
    if spec["name"] == "lstm":
        T, H = spec["T"], spec["H"]
        B = scale_rows or spec["B"]
        x = rng.standard_normal((T, B, 4 * H), dtype=np.float32)
        U = (rng.standard_normal((H, 4 * H), dtype=np.float32) / np.sqrt(H)).astype(np.float32)
        return [x, np.zeros((B, H), np.float32), np.zeros((B, H), np.float32), U]
    if spec["name"] == "logreg":
        N = scale_rows or spec["N"]
        D = spec["D"]
        X = rng.standard_normal((N, D), dtype=np.float32)
        y = (rng.random(N) < 0.5).astype(np.float32)
        w = (rng.standard_normal(D) * 0.01).astype(np.float32)
        return [X, y, w, np.float32(0.0)]
    if spec["name"] == "readme":
        n = spec["n"]
        return [np.float64(1.5), rng.standard_normal(n), rng.standard_normal((n, n))]
    if spec["name"] == "mlp":
        B = scale_rows or spec["B"]
        H = spec["H"]
        X = rng.standard_normal((B, H), dtype=np.float32)
        Y = rng.standard_normal((B, H), dtype=np.float32)
        W1 = (rng.standard_normal((H, H), dtype=np.float32) / np.sqrt(H)).astype(np.float32)
        W2 = (rng.standard_normal((H, H), dtype=np.float32) / np.sqrt(H)).astype(np.float32)
        return [X, Y, W1, np.zeros(H, np.float32), W2, np.zeros(H, np.float32)]
    n = scale_rows or spec["n"]
    return [rng.standard_normal(n, dtype=np.float32) for _ in range(3)]


def make_inputs_device(spec, seed):
    import torch

    from aesara_b200.runtime.device import DeviceArray

    g = torch.Generator(device="cuda").manual_seed(seed)
    if spec["name"] == "mlp":
        B, H = spec["B"], spec["H"]
        X = torch.randn(B, H, device="cuda", generator=g)
        Y = torch.randn(B, H, device="cuda", generator=g)
        W1 = torch.randn(H, H, device="cuda", generator=g) / H ** 0.5
        W2 = torch.randn(H, H, device="cuda", generator=g) / H ** 0.5
        b1 = torch.zeros(H, device="cuda")
        b2 = torch.zeros(H, device="cuda")
        ts = [X, Y, W1, b1, W2, b2]
    elif spec["name"] == "lstm":
        T, B, H = spec["T"], spec["B"], spec["H"]
        x = torch.randn(T, B, 4 * H, device="cuda", generator=g)
        U = torch.randn(H, 4 * H, device="cuda", generator=g) / H ** 0.5
        ts = [x, torch.zeros(B, H, device="cuda"), torch.zeros(B, H, device="cuda"), U]
    elif spec["name"] == "logreg":
        N, D = spec["N"], spec["D"]
        X = torch.randn(N, D, device="cuda", generator=g)
        y = (torch.rand(N, device="cuda", generator=g) < 0.5).float()
        w = torch.randn(D, device="cuda", generator=g) * 0.01
        ts = [X, y, w]
        return [DeviceArray.from_torch(t) for t in ts] + [np.float32(0.0)], ts
    elif spec["name"] == "readme":
        n = spec["n"]
        v = torch.randn(n, device="cuda", generator=g, dtype=torch.float64)
        M = torch.randn(n, n, device="cuda", generator=g, dtype=torch.float64)
        ts = [v, M]
        return [np.float64(1.5)] + [DeviceArray.from_torch(t) for t in ts], ts
    else:
        ts = [torch.randn(spec["n"], device="cuda", generator=g) for _ in range(3)]
    return [DeviceArray.from_torch(t) for t in ts], ts


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nme, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def ncu_traffic(summary, kernel_substr):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the kernel from
    a committed ``ncu --set full`` summary under profiles/ (None if it is not there)."""
    path = os.path.join(ROOT, "profiles", summary)
    if not os.path.exists(path):
        return None
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    total, inside, seen = 0.0, False, 0
    for line in open(path):
        if line.startswith("kernel:"):
            if inside and seen == 2:
                break
            inside, total, seen = kernel_substr in line, 0.0, 0
        elif inside and ("dram__bytes_read.sum " in line or "dram__bytes_write.sum " in line):
            parts = line.split()
            total += float(parts[1]) * unit.get(parts[2], 1.0)
            seen += 1
    return total if seen == 2 else None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm=d["hbm_gbs"], bf16=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    bf16_burst=d["bf16_tflops"], src="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, bf16=1400.0, bf16_burst=1590.0, src="fallback (B200_PROFILING.md)")


# ----------------------------------------------------------------------------- CPU arm
def cpu_baseline(spec, seconds_budget=20.0):
    """The oracle port (NumPy restatement of the reference's per-Op code) timed on
    the host cores on a bounded sample of the same workload; extrapolated per row."""
    from aesara_b200.ir import Program
    from oracle.program_np import run_program

    try:  # torchrun pins OMP_NUM_THREADS=1: give the BLAS under NumPy every host core back
        from threadpoolctl import threadpool_limits

        threadpool_limits(limits=os.cpu_count())
    except Exception:
        pass
    prog = Program.load(os.path.join(GOLDEN, spec["program"] + ".json"))
    rng = np.random.default_rng(0)
    if spec["name"] == "mlp":
        rows, full = min(spec["B"], 1024), spec["B"]
        sample = f"B={rows} rows of {full} (H={spec['H']}); time scaled by {full}/{rows}"
    elif spec["name"] == "lstm":
        rows, full = min(spec["B"], 128), spec["B"]
        sample = f"B={rows} batch rows of {full} (T={spec['T']}, H={spec['H']}); time scaled by {full}/{rows}"
    elif spec["name"] == "logreg":
        rows, full = min(spec["N"], 1 << 18), spec["N"]
        sample = f"{rows} of {full} rows (D={spec['D']}); time scaled by {full}/{rows}"
    elif spec["name"] == "readme":
        rows, full = None, 1
        sample = "full size"
    else:
        rows, full = min(spec["n"], 1 << 22), spec["n"]
        sample = f"{rows} of {full} elements; time scaled by {full}/{rows}"
    ins = make_inputs_numpy(spec, rng, scale_rows=rows)
    run_program(prog, ins)  # warm-up
    times = []
    t_end = time.perf_counter() + seconds_budget
    while len(times) < 3 or (time.perf_counter() < t_end and len(times) < 8):
        t0 = time.perf_counter()
        run_program(prog, ins)
        times.append(time.perf_counter() - t0)
    t = float(np.median(times)) * ((full / rows) if rows else 1.0)
    return {"value": 1.0 / t, "unit": "graph-evals/s", "cores": os.cpu_count(), "kind": "port",
            "sample": sample, "ms_per_eval_extrapolated": t * 1e3}


def run_reference_arm(args, spec, rank, world):
    if rank != 0:
        return
    cb = None
    t0 = time.perf_counter()
    vals = []
    for _ in range(max(1, min(args.steps, 3))):
        cb = cpu_baseline(spec, seconds_budget=5.0)
        vals.append(cb["value"])
        if time.perf_counter() - t0 > 120:
            break
    v = float(np.median(vals))
    cb["value"] = v
    line = {
        "impl": "reference", "metric": "graph-evals/s", "value": v, "unit": "graph-evals/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 / v, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": spec["desc"], "note": "oracle port (NumPy restatement of the reference "
                   "C-linker path) on host cores; the reference itself cannot travel to the GPU box"},
        "cpu_baseline": cb,
        "e2e": {"value": v, "unit": "graph-evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="mlp", choices=["mlp", "elemwise", "lstm", "logreg", "readme"])
    ap.add_argument("--steps-t", type=int, default=0, help="Scan length for --workload lstm")
    ap.add_argument("--graph", type=int, default=-1, help="1/0: replay the evaluation as a CUDA graph")
    ap.add_argument("--precision", default=None, choices=list(PRECISIONS),
                    help="GEMM compute policy; default: bf16 for mlp (BASELINE cfg3), fp32 (3xTF32) otherwise")
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--hidden", type=int, default=0)
    ap.add_argument("--n", type=int, default=0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    spec = workload_spec(args.workload, args)
    if args.precision is None:
        args.precision = "bf16" if args.workload == "mlp" else "fp32"
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, spec, rank, world)
        return

    import torch

    from aesara_b200.ir import Program
    from aesara_b200.runtime import lib
    from aesara_b200.runtime.device import DeviceArray
    from aesara_b200.runtime.vm import ProgramExecutor

    torch.cuda.set_device(local_rank)
    lib.check(lib.load().ab_init(local_rank))
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    prec = PRECISIONS[args.precision]
    prog = Program.load(os.path.join(GOLDEN, spec["program"] + ".json"))
    use_graph = bool(spec.get("graph")) if args.graph < 0 else bool(args.graph)
    ex = ProgramExecutor(prog, precision=prec, host_outputs=False, time_nodes=not use_graph)
    dev_in, keep = make_inputs_device(spec, seed=1234 + rank)
    run = ex
    if use_graph:
        from aesara_b200.runtime.graph import GraphReplay

        run = GraphReplay(ex)

    combiner = None
    if world > 1:
        from aesara_b200.shard import OutputCombiner

        combiner = OutputCombiner(world, mode="mean")

    def step():
        outs = run(*dev_in)
        if combiner is not None:
            outs = combiner(outs)
        return outs

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    launches0 = lib.load().ab_launch_count()
    clocks = ClockSampler(local_rank)
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    node_ms = {}
    barrier()
    e0.record()
    pending = []
    for _ in range(args.steps):
        step()
        if not use_graph:
            pending.append(ex.node_events)
    e1.record()
    barrier()
    clk = clocks.stop()
    ms_total = e0.elapsed_time(e1)
    launches = lib.load().ab_launch_count() - launches0
    if use_graph:
        # a replayed graph re-issues the kernels captured once: count them from one eager call
        l0 = lib.load().ab_launch_count()
        ex(*dev_in)
        torch.cuda.synchronize()
        launches = (lib.load().ab_launch_count() - l0) * args.steps if run.replays else launches
    for evs in pending:
        for i, a, b in evs:
            node_ms.setdefault(i, []).append(a.elapsed_time(b))
    t_ms = torch.tensor([ms_total], device="cuda")
    if dist is not None:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_step = t_ms.item() / args.steps
    value = world * 1e3 / ms_step  # every rank evaluates its shard once per step

    # per-kind device time (average per step) from the per-node CUDA events
    gemm_ms, hbm_ms, other_ms = 0.0, 0.0, 0.0
    fused_gemm = {f.last for f in ex._fusions if type(f).__name__ == "GemmEpilogueFusion"}
    for i, lst in node_ms.items():
        # a fused Gemm -> Elemwise region is timed at its last node: it is tensor-pipe work
        op = "Gemm" if i in fused_gemm else prog.nodes[i].op
        t = float(np.mean(lst))
        if op in ("Dot22", "Gemm", "Dot22Scalar", "Scan"):
            gemm_ms += t
        elif op in ("Elemwise", "CAReduce"):
            hbm_ms += t
        else:
            other_ms += t
    peaks = measured_peaks()
    if use_graph:
        # no per-node events inside a replayed graph: the roofline is taken over the whole
        # step (conservative: every kernel of the evaluation is charged to the bound resource)
        roofline_hbm = None
        if spec["n_gemm"]:
            ach = spec["gemm_flops"] / (ms_step * 1e-3) / 1e12
            peak = peaks["bf16"] if args.precision == "bf16" else peaks["bf16"] / 2.0
            roofline = {"bound": "tensor", "kernel": "whole evaluation (CUDA-graph replay); Gemm = gemm_tcgen05_kernel",
                        "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                        "peak_source": peaks["src"] + ("" if args.precision == "bf16" else "; tf32 = bf16/2 (nominal ratio)"),
                        "traffic": None, "ms_per_step": ms_step}
            ach_h = spec["elemwise_bytes"] / (ms_step * 1e-3) / 1e9
            roofline_hbm = {"bound": "hbm", "kernel": "whole evaluation", "achieved": ach_h,
                            "peak": peaks["hbm"], "unit": "GB/s", "frac": ach_h / peaks["hbm"],
                            "peak_source": peaks["src"], "traffic": None, "ms_per_step": ms_step}
        else:
            single_pass = bool(spec.get("fused_bytes")) and ex.fused_regions_run > 0
            nbytes = spec["fused_bytes"] if single_pass else spec["elemwise_bytes"]
            ach_h = nbytes / (ms_step * 1e-3) / 1e9
            roofline = {"bound": "hbm", "kernel": "whole evaluation (CUDA-graph replay)"
                        + ("; ab_rowfused single pass over X" if single_pass else ""),
                        "achieved": ach_h, "peak": peaks["hbm"], "unit": "GB/s",
                        "frac": ach_h / peaks["hbm"], "peak_source": peaks["src"], "traffic": None,
                        "ms_per_step": ms_step,
                        "algorithmic_bytes": nbytes,
                        "bytes_model": ("single pass: N*D*4 + 16 N (SURVEY 8d: 'report which is implemented')"
                                        if single_pass else "graph as optimised: 2*N*D*4 + 48 N (SURVEY 8d)")}
        gemm_ms = hbm_ms = other_ms = None
    elif spec["n_gemm"]:
        ach = spec["gemm_flops"] / (gemm_ms * 1e-3) / 1e12
        peak = peaks["bf16"] if args.precision == "bf16" else peaks["bf16"] / 2.0
        n_fused = sum(1 for f in ex._fusions if type(f).__name__ == "GemmEpilogueFusion" and not f.broken)
        roofline = {"bound": "tensor", "kernel": "gemm_tcgen05_kernel (+operand pack) per Gemm/Dot22 node"
                    + (f"; {n_fused} of them with the consuming Elemwise node fused into the epilogue" if n_fused else ""),
                    "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                    "peak_source": peaks["src"] + ("" if args.precision == "bf16" else "; tf32 = bf16/2 (nominal ratio)"),
                    "traffic": None, "ms_per_step": gemm_ms}
        ach_h = spec["elemwise_bytes"] / (hbm_ms * 1e-3) / 1e9 if hbm_ms else None
        roofline_hbm = {"bound": "hbm", "kernel": "fused Elemwise + CAReduce nodes"
                        + (f" (the algorithmic bytes of all of them over the time of those not absorbed into a GEMM "
                           f"epilogue: may exceed the peak, SURVEY 8d)" if n_fused else ""),
                        "achieved": ach_h, "peak": peaks["hbm"], "unit": "GB/s",
                        "frac": ach_h / peaks["hbm"] if ach_h else None,
                        "peak_source": peaks["src"], "traffic": None, "ms_per_step": hbm_ms}
    else:
        ach_h = spec["elemwise_bytes"] / (hbm_ms * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": "ab_ew_flat_vec (generated fused Elemwise)",
                    "achieved": ach_h, "peak": peaks["hbm"], "unit": "GB/s",
                    "frac": ach_h / peaks["hbm"], "peak_source": peaks["src"], "traffic": None,
                    "ms_per_step": hbm_ms}
        roofline_hbm = None

    if roofline.get("bound") == "tensor" and args.precision == "fp32":
        # fp32-faithful products issue three TF32 MMAs each: the pipe-level ceiling for the
        # algorithmic flops is a third of the TF32 peak
        roofline["mma_per_product"] = 3
        roofline["ceiling_3xtf32"] = roofline["peak"] / 3.0
        roofline["frac_of_3xtf32_ceiling"] = roofline["achieved"] / (roofline["peak"] / 3.0)

    # DRAM traffic of the dominant kernel, per launch, from the committed ncu --set full summary
    # of the same command (profiles/; cold-cache, serialised capture)
    tsrc = None
    if spec["name"] == "mlp" and args.precision == "bf16":
        fused = any(type(f).__name__ == "GemmEpilogueFusion" and not f.broken for f in ex._fusions)
        tsrc = (("r01_gemm_fused_epilogue.txt", "ab_gemm_ep_2cta_f16") if fused
                else ("r01_gemm_bf16_v4_8warp_epilogue.txt", "gemm_tcgen05_2cta_kernel"))
    elif spec["name"] == "logreg" and ex.fused_regions_run > 0:
        tsrc = ("r01_rowfused_logreg.txt", "ab_rowfused")
    elif spec["name"] == "elemwise":
        tsrc = ("r01_elemwise_cfg2_v2_unroll1.txt", "ab_ew_flat_vec")
    if tsrc is not None:
        roofline["traffic"] = ncu_traffic(*tsrc)
        roofline["traffic_source"] = f"profiles/{tsrc[0]} ({tsrc[1]}, dram__bytes_read.sum + dram__bytes_write.sum per launch)"

    # end to end through the host API: pinned host inputs, H2D + eval + D2H of every output
    e2e = None
    if not args.no_e2e and world == 1:
        host_in = []
        h2d = 0
        for t in keep:
            h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            h.copy_(t)
            host_in.append(h)
            h2d += h.numel() * h.element_size()
        template = [None if isinstance(a, DeviceArray) else a for a in dev_in]
        del keep[:]
        del dev_in[:]
        run = None
        torch.cuda.empty_cache()
        # the public call: host (page-locked) ndarrays in, host ndarrays out
        ex2 = ProgramExecutor(prog, precision=prec, host_outputs=True)
        it = iter(host_in)
        host_args = [slot if slot is not None else next(it).numpy() for slot in template]

        def e2e_step():
            res = ex2(*host_args)
            return sum(np.asarray(r).nbytes for r in res)

        d2h = e2e_step()
        torch.cuda.synchronize()
        n_e2e = max(2, min(args.steps, 5))
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n_e2e):
            e2e_step()
        b.record()
        torch.cuda.synchronize()
        e2e_ms = a.elapsed_time(b) / n_e2e
        e2e = {"value": 1e3 / e2e_ms, "unit": "graph-evals/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms, "steps": n_e2e}

    if rank == 0:
        cb = None if args.no_cpu else cpu_baseline(spec)
        if world == 1:
            par = "single"
        else:
            from aesara_b200.shard import exchange_plan

            how = exchange_plan(combiner.layout.total, world) if combiner.layout is not None else "allgather"
            par = (f"dp{world} (batch rows sharded; outputs combined by one NCCL "
                   + ("all-reduce of the pre-weighted packed outputs)" if how == "allreduce"
                      else "all-gather of the packed outputs + local weighted sum)"))
        line = {
            "metric": "graph-evals/s", "value": value, "unit": "graph-evals/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"bf16": "bf16", "tf32": "tf32", "fp32": "f32 (3xTF32)"}[args.precision]
            if spec["n_gemm"] else "f32",
            "data": "synthetic",
            "config": {"workload": spec["desc"], "parallelism": par,
                       "l2": ("working set fits L2: launch-latency bound, reported as evals/s only"
                              if spec["name"] == "readme" else "inputs >> 126 MB L2, no flush needed"),
                       "executor": "cuda-graph replay" if use_graph else "eager launches + per-node CUDA events",
                       "gemm_precision": args.precision if spec["n_gemm"] else None,
                       "per_gpu": {k: spec[k] for k in ("B", "H", "n", "N", "D", "T") if k in spec}},
            "roofline": roofline, "roofline_hbm": roofline_hbm,
            "device_ms": {"gemm": gemm_ms, "elemwise_careduce": hbm_ms, "other": other_ms},
            "cpu_baseline": cb, "e2e": e2e, "gpu_launches": int(launches), "clocks": clk,
        }
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
