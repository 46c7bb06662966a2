#!/bin/bash
# One validation trip on a B200 (run under gpurun): the sections are independent, each under
# its own timeout, risky kernels last.  usage: tools/gpu_trip.sh <section>...
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONPATH=$PWD
for sec in "$@"; do
case $sec in
  newtests)
    timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_function.py -m gpu -q -p no:cacheprovider -k "big or cuda_graph or launch_counts or matches_reference or epilogue_fusion" > gpurun_out/pytest_new.log 2>&1; echo "newtests rc=$?"; tail -5 gpurun_out/pytest_new.log | cut -c1-400;;
  alltests)
    timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_all.log 2>&1; echo "alltests rc=$?"; tail -8 gpurun_out/pytest_all.log | cut -c1-400;;
  c4test)
    timeout 300 python -m pytest tests/test_gpu_blas.py -m gpu -q -p no:cacheprovider -x -k "four_cta" > gpurun_out/pytest_c4.log 2>&1; echo "c4 rc=$?"; tail -12 gpurun_out/pytest_c4.log | cut -c1-300;;
  blas)
    timeout 600 python -m pytest tests/test_gpu_blas.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_blas.log 2>&1; echo "blas rc=$?"; tail -8 gpurun_out/pytest_blas.log | cut -c1-300;;
  bench)
    timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_default.err; cut -c1-2500 gpurun_out/bench_default.json;;
  bench_quick)
    timeout 400 python bench.py --gpus 1 --steps 10 --warmup 3 --no-also --no-cpu --no-e2e > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench_quick rc=$?"; tail -3 gpurun_out/bench_quick.err; python -c "import json;d=json.load(open('gpurun_out/bench_quick.json'));print(d['ms_per_step'], d['ms_per_step_eager'], [g['ms'] for g in d['gemm_nodes']], d['roofline']['frac'], d['clocks'], d['parity']['max_err'])";;
  bench_notplane)
    AB_EP_NO_TPLANE=1 timeout 400 python bench.py --gpus 1 --steps 10 --warmup 3 --no-also --no-cpu --no-e2e > gpurun_out/bench_notplane.json 2> gpurun_out/bench_notplane.err; echo "bench_notplane rc=$?"; tail -3 gpurun_out/bench_notplane.err; python -c "import json;d=json.load(open('gpurun_out/bench_notplane.json'));print(d['ms_per_step'], d['ms_per_step_eager'], [g['ms'] for g in d['gemm_nodes']])";;
  bench_tplane_r13)
    AB_EP_TPLANE_NO_FULLSUM=1 timeout 400 python bench.py --gpus 1 --steps 10 --warmup 3 --no-also --no-cpu --no-e2e > gpurun_out/bench_tplane_r13.json 2> gpurun_out/bench_tplane_r13.err; echo "bench_tplane_r13 rc=$?"; tail -3 gpurun_out/bench_tplane_r13.err; python -c "import json;d=json.load(open('gpurun_out/bench_tplane_r13.json'));print(d['ms_per_step'], d['ms_per_step_eager'], [g['ms'] for g in d['gemm_nodes']])";;
  bench_fp32)
    for v in ${FP32_VARIANTS:-default}; do
      case $v in default) envs="";; regions) envs="AB_GEMM_FUSE_FP32=1";; nostaging) envs="AB_GEMM_FUSE_FP32=1 AB_EP_NO_STAGING=1";; nofuse) envs="AB_NO_GEMM_FUSE=1";; esac
      env $envs timeout 400 python bench.py --gpus 1 --steps 4 --warmup 3 --precision fp32 --no-also --no-cpu --no-e2e --no-truth > gpurun_out/bench_fp32_$v.json 2> gpurun_out/bench_fp32_$v.err; echo "bench_fp32 $v rc=$?"; python -c "import json;d=json.load(open('gpurun_out/bench_fp32_$v.json'));print(d['ms_per_step'], d['ms_per_step_eager'], [g['ms'] for g in d['gemm_nodes']], d['device_ms'])"
    done;;
  bench_nostaging)
    AB_EP_NO_STAGING=1 timeout 400 python bench.py --gpus 1 --steps 10 --warmup 3 --no-also --no-cpu --no-e2e > gpurun_out/bench_nostaging.json 2> gpurun_out/bench_nostaging.err; echo "bench_nostaging rc=$?"; tail -3 gpurun_out/bench_nostaging.err; python -c "import json;d=json.load(open('gpurun_out/bench_nostaging.json'));print(d['ms_per_step'], d['ms_per_step_eager'], [g['ms'] for g in d['gemm_nodes']])";;
  bench_noc4)
    AB_GEMM_NO_CLUSTER4=1 timeout 400 python bench.py --gpus 1 --steps 10 --warmup 3 --no-also --no-cpu --no-e2e > gpurun_out/bench_noc4.json 2> gpurun_out/bench_noc4.err; echo "bench_noc4 rc=$?"; tail -3 gpurun_out/bench_noc4.err; cut -c1-1800 gpurun_out/bench_noc4.json;;
  bench_single)
    AB_GEMM_FUSE_SINGLE=1 timeout 400 python bench.py --gpus 1 --steps 10 --warmup 3 --no-also --no-cpu --no-e2e > gpurun_out/bench_single.json 2> gpurun_out/bench_single.err; echo "bench_single rc=$?"; tail -3 gpurun_out/bench_single.err; cut -c1-1800 gpurun_out/bench_single.json;;
  bench_single_noc4)
    AB_GEMM_NO_CLUSTER4=1 AB_GEMM_FUSE_SINGLE=1 timeout 400 python bench.py --gpus 1 --steps 10 --warmup 3 --no-also --no-cpu --no-e2e > gpurun_out/bench_single_noc4.json 2> gpurun_out/bench_single_noc4.err; echo "bench_single_noc4 rc=$?"; tail -3 gpurun_out/bench_single_noc4.err; cut -c1-400 gpurun_out/bench_single_noc4.json;;
  fullsize)
    timeout 900 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -s -p no:cacheprovider > gpurun_out/pytest_fullsize.log 2>&1; echo "fullsize rc=$?"; grep -a "cfg\|passed\|failed\|Error" gpurun_out/pytest_fullsize.log | cut -c1-300 | tail -20;;
  e2e)
    timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-also --no-cpu > gpurun_out/bench_e2e.json 2> gpurun_out/bench_e2e.err; echo "bench_e2e rc=$?"; tail -3 gpurun_out/bench_e2e.err; python -c "import json;d=json.load(open('gpurun_out/bench_e2e.json'));print(d['ms_per_step'], d['e2e'], d.get('parity'))";;
  e2e_probe)
    timeout 600 python tools/e2e_probe.py > gpurun_out/e2e_probe.log 2>&1; grep -a "host_chunks\|GB/s" gpurun_out/e2e_probe.log | tail -8;;
  gemm_probe)
    timeout 600 python tools/gemm_probe.py 20 > gpurun_out/gemm_probe.log 2>&1; tail -16 gpurun_out/gemm_probe.log;;
  gemm_probe_dw)
    PROBE_DW_LAYOUTS=1 timeout 600 python tools/gemm_probe.py 20 > gpurun_out/gemm_probe_dw.log 2>&1; tail -14 gpurun_out/gemm_probe_dw.log; cp gpurun_out/gemm_probe.json gpurun_out/gemm_probe_dw.json;;
  region2_ncu)
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ab_gemm_ep" -s 4 -c 1 -o gpurun_out/region2_prof -f python bench.py --steps 1 --warmup 1 --no-also --no-cpu --no-e2e --no-truth --graph 0 > gpurun_out/region2_ncu.log 2>&1; echo "region2_ncu rc=$?"; tail -2 gpurun_out/region2_ncu.log | cut -c1-200; ls -la gpurun_out/region2_prof.ncu-rep;;
  gemm_probe_half)
    PROBE_HALF_GRID=1 timeout 600 python tools/gemm_probe.py 10 > gpurun_out/gemm_probe_half.log 2>&1; tail -11 gpurun_out/gemm_probe_half.log; cp gpurun_out/gemm_probe.json gpurun_out/gemm_probe_half.json;;
  dw_ncu)
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_tcgen05_2cta" -c 6 -o gpurun_out/dw_layout_prof -f python tools/dw_layout_ncu.py > gpurun_out/dw_ncu.log 2>&1; echo "dw_ncu rc=$?"; tail -4 gpurun_out/dw_ncu.log | cut -c1-200; ls -la gpurun_out/dw_layout_prof.ncu-rep;;
  gemm_ncu)
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 6 -c 2 -o gpurun_out/gemm_prof -f python tools/gemm_probe.py 2 > gpurun_out/gemm_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/gemm_ncu.log | cut -c1-300; ls -la gpurun_out/gemm_prof.ncu-rep;;
  scantests)
    timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -p no:cacheprovider -k "lstm or scan or cfg4" > gpurun_out/pytest_scan.log 2>&1; echo "scantests rc=$?"; tail -6 gpurun_out/pytest_scan.log | cut -c1-300;;
  bench_lstm)
    timeout 600 python bench.py --workload lstm --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_lstm.json 2> gpurun_out/bench_lstm.err; echo "bench_lstm rc=$?"; tail -2 gpurun_out/bench_lstm.err; python -c "import json;d=json.load(open('gpurun_out/bench_lstm.json'));print(d['ms_per_step'], d['roofline'])";;
  lstm_ncu)
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:"lstm_scan" -s 1 -c 1 -o gpurun_out/lstm_prof -f python bench.py --workload lstm --steps 1 --warmup 1 --no-also --no-cpu --no-e2e --no-truth --graph 0 > gpurun_out/lstm_ncu.log 2>&1; echo "lstm_ncu rc=$?"; tail -2 gpurun_out/lstm_ncu.log | cut -c1-200; ls -la gpurun_out/lstm_prof.ncu-rep;;
  bench_ncu)
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ab_gemm_ep|gemm_tcgen05" -s 8 -c 5 -o gpurun_out/bench_prof -f python bench.py --steps 1 --warmup 1 --no-also --no-cpu --no-e2e --no-truth --graph 0 > gpurun_out/bench_ncu.log 2>&1; echo "bench_ncu rc=$?"; tail -2 gpurun_out/bench_ncu.log | cut -c1-200; ls -la gpurun_out/bench_prof.ncu-rep;;
  multitest)
    timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -p no:cacheprovider -x > gpurun_out/pytest_multi.log 2>&1; echo "multitest rc=$?"; tail -6 gpurun_out/pytest_multi.log | cut -c1-300;;
  scale_mlp)
    N=$(nvidia-smi -L | wc -l)
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_mlp_n$N.json 2> gpurun_out/bench_mlp_n$N.err; echo "scale_mlp N=$N rc=$?"; tail -3 gpurun_out/bench_mlp_n$N.err | cut -c1-300; python -c "import json;d=json.loads([l for l in open('gpurun_out/bench_mlp_n$N.json') if l.startswith('{')][-1]);print(d['n_gpus'], d['value'], d['ms_per_step'], d['config']['executor'], d.get('parity_sharded'), d.get('exchange'))";;
  scale_mlp_nograph)
    N=$(nvidia-smi -L | wc -l)
    AB_SHARD_GRAPH=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_mlp_nograph_n$N.json 2> gpurun_out/bench_mlp_nograph_n$N.err; echo "scale_mlp_nograph N=$N rc=$?"; python -c "import json;d=json.loads([l for l in open('gpurun_out/bench_mlp_nograph_n$N.json') if l.startswith('{')][-1]);print(d['n_gpus'], d['value'], d['ms_per_step'], d['config']['executor'])";;
  scale_logreg)
    N=$(nvidia-smi -L | wc -l)
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --workload logreg --gpus $N --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_logreg_n$N.json 2> gpurun_out/bench_logreg_n$N.err; echo "scale_logreg N=$N rc=$?"; tail -3 gpurun_out/bench_logreg_n$N.err | cut -c1-300; python -c "import json;d=json.loads([l for l in open('gpurun_out/bench_logreg_n$N.json') if l.startswith('{')][-1]);print(d['n_gpus'], d['value'], d['ms_per_step'], d['config']['executor'], d.get('parity_sharded'), d.get('exchange'))";;
  ew_probe)
    timeout 300 python tools/ew_probe.py 2>&1 | tail -8;;
  reference)
    timeout 500 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref rc=$?"; tail -2 gpurun_out/bench_reference.err; cut -c1-1200 gpurun_out/bench_reference.json;;
  smoke)
    timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3;;
  launches)
    timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-also --no-cpu --no-e2e --graph 0 > gpurun_out/launches.log 2>&1; echo "launches rc=$?"; tail -2 gpurun_out/launches.log | cut -c1-300;;
  *) echo "unknown section $sec";;
esac
done
