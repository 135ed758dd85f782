#!/bin/bash
# Multi-GPU hardware check on N GPUs of one box: the multi-GPU parity suite (only the cases of exactly N ranks when N > 2: an
# N-GPU call is charged N x its wall time), then bench.py.
#   gpurun --gpus N --timeout 1500 -- 'bash tools/gpu_multi_check.sh N'
set -u
n=${1:-2}
mkdir -p gpurun_out
export CNV_PEER_TIMEOUT_MS=${CNV_PEER_TIMEOUT_MS:-30000}
if [ "${SKIP_TESTS:-0}" != "1" ]; then
    f=tests/test_gpu_z_multi.py
    sel="$f"
    if [ "$n" -gt 2 ]; then   # only the cases of exactly N ranks (peer exchange, C driver)
        sel="$f::test_slab_poisson_bitwise[$n-peer] $f::test_c_driver_multi_gpu_matches_single_gpu[$n]"
    fi
    timeout 600 python -m pytest $sel -m gpu -x -q > gpurun_out/multi_tests_$n.log 2>&1
    echo "pytest exit $?" >> gpurun_out/multi_tests_$n.log
    tail -4 gpurun_out/multi_tests_$n.log
fi
run() { timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 --master-port "$1" "${@:2}"; }
run 29801 bench.py --gpus "$n" --steps ${STEPS:-10} --warmup 3 > gpurun_out/bench_peer_$n.json 2> gpurun_out/bench_peer_$n.err
echo "bench peer exit $?"; tail -c 400 gpurun_out/bench_peer_$n.err
