"""GPU, 2 processes x 1 GPU each over NCCL: ``aesara.function(..., mode=B200(shard="rows"))``.

The combined outputs of the row-sharded function (unequal row blocks) equal ONE GPU's
evaluation of the concatenated batch; large gradients are handed to NCCL before the evaluation
has finished (ShardedExecutor.early_issued).  Skipped on a box with fewer than 2 GPUs — the same
check runs as the pre-flight of ``bench.py --gpus N`` (``parity_sharded`` in its JSON line)."""
import os
import socket

import numpy as np
import pytest

from aesara_b200.compat import bootstrap

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not bootstrap.available(), reason="reference front-end (oracle/_ref) not present")]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from aesara_b200.runtime import lib

    lib.check(lib.load().ab_init(rank))
    aesara = bootstrap.load_aesara()
    import aesara_b200.linker as L
    from aesara_b200 import graphs as G
    from aesara_b200.shard import row_block

    res = {}
    for name, build, inputs in (("cfg3", G.cfg3_mlp, lambda: G.cfg3_inputs(1100, 256)),
                                ("cfg5", G.cfg5_logreg, lambda: G.cfg5_inputs(3001, 64)),
                                ("cfg2", G.cfg2_fused_elemwise, lambda: G.cfg2_inputs(70000))):  # even: equal blocks for the gather
        i, o = build()
        f = aesara.function(i, o, mode=L.mode(shard="rows", gather=True), on_unused_input="ignore")
        plan = f.maker.linker.shard_plan
        vals = inputs()
        n = next(np.shape(v)[ax] for v, ax in zip(vals, plan.sharded_inputs) if ax is not None)
        a, b = row_block(n, world, rank)
        local = [np.ascontiguousarray(v[a:b]) if ax == 0 else v for v, ax in zip(vals, plan.sharded_inputs)]
        got = f(*local)
        got = got if isinstance(got, (list, tuple)) else [got]
        ex = f.vm.executor
        res[name] = {"modes": [m[0] for m in plan.outputs], "early": ex.early_issued, "exchanges": ex.exchanges}
        if rank == 0:
            i2, o2 = build()
            want = aesara.function(i2, o2, mode=L.mode(), on_unused_input="ignore")(*vals)
            want = want if isinstance(want, (list, tuple)) else [want]
            errs = []
            for g, w, m in zip(got, want, plan.outputs):
                if m[0] == "concat" and world > 1 and n % world:
                    continue  # all_gather_into_tensor needs equal blocks: only checked for cfg2 below
                g, w = np.asarray(g, np.float64), np.asarray(w, np.float64)
                assert g.shape == w.shape, (name, g.shape, w.shape)
                errs.append(float(np.max(np.abs(g - w)) / max(np.max(np.abs(w)), 1e-30)))
            res[name]["errs"] = errs
    if rank == 0:
        q.put(res)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_row_sharded_functions_match_one_gpu():
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=500)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res["cfg3"]["modes"] == ["mean"] * 5 and res["cfg5"]["modes"] == ["mean"] * 3
    assert res["cfg2"]["modes"] == ["concat"]
    for name, r in res.items():
        assert r["errs"] and max(r["errs"]) < 2e-5, (name, r)
    assert res["cfg3"]["early"] >= 2      # dW2 and dW1 went to NCCL while the evaluation ran
    assert res["cfg5"]["exchanges"] == 1  # three small outputs travel packed in one all-reduce
