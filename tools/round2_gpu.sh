#!/bin/bash
# First GPU calls of the next round: verify on hardware what was written after round 1's GPU budget was spent
# (planner change 97f5c9d: taller chunks at the domain boundary; the lagged stop decision of the peer path,
# CNV_PEER_LAG=1), then measure it.  Everything writes into gpurun_out/.
#
# An N-GPU call is charged N x its wall time against the round's 180 GPU-minutes, so the multi-GPU modes are short:
#   1 GPU  (~25 min):  gpurun --timeout 2400 -- 'bash tools/round2_gpu.sh single'
#   2 GPUs (~8 min) :  gpurun --gpus 2 --timeout 900 -- 'bash tools/round2_gpu.sh lagtest 2'
#   2 GPUs (~4 min) :  gpurun --gpus 2 --timeout 600 -- 'bash tools/round2_gpu.sh scale 2'
#   8 GPUs (~4 min) :  gpurun --gpus 8 --timeout 600 -- 'bash tools/round2_gpu.sh scale 8'     (only after lagtest passed)
#   optional        :  gpurun --gpus N --timeout 600 -- 'bash tools/round2_gpu.sh extra N'
# Afterwards, here:  python tools/collect_round2.py gpurun_out > profiles/ab_r2.md   (one report of all A/B results)
set -u
mkdir -p gpurun_out
mode=${1:-single}
case "$mode" in
single)
    # parity first (the planner change is covered by the emulator on the CPU; this is the hardware confirmation)
    timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1
    echo "pytest exit $?" >> gpurun_out/r2_pytest_gpu.log
    tail -3 gpurun_out/r2_pytest_gpu.log
    # headline, with and without the taller chunks at the domain boundary (planner change 97f5c9d; CNV_POISSON_EDGE=0 = old plan)
    python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
    CNV_POISSON_EDGE=0 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/r2_bench_n1_edge0.json 2>&1
    # A/B: the leaner step body (libcnavier_b200_lean.so: about 82 instead of 111 instructions per thread and row-step on the
    # steady-state path, 2-long norm chain; bit-exact on the emulator) -- parity on hardware first, then the same bench
    CNV_LIB=lean timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "poisson or steps or golden" > gpurun_out/r2_pytest_lean.log 2>&1
    echo "pytest exit $?" >> gpurun_out/r2_pytest_lean.log
    CNV_LIB=lean python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/r2_bench_n1_lean.json 2> gpurun_out/r2_bench_n1_lean.err
    python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/r2_bench_n1_again.json 2>&1   # same box, same moment: the A of the A/B
    # stationary-tile kernel after the batching rework (loads of a batch of 4 rows issued together, select-free bodies, no
    # spills): vs the streaming kernel on the L2-resident shapes (planner's tile plans for T = 2, 4, 6, 8; with and without
    # programmatic dependent launch)
    for shape in "1024 1024" "2048 2048" "4096 4096" "528 4096" "256 256"; do
        python tools/probe_poisson.py $shape "8:0:0,t2:0:0:0,t4:0:0:0,t6:0:0:0,t8:0:0:0" 512 >> gpurun_out/r2_probe_tile.log 2>&1
        CNV_TILE_PDL=1 python tools/probe_poisson.py $shape "t2:0:0:0,t4:0:0:0,t6:0:0:0,t8:0:0:0" 512 >> gpurun_out/r2_probe_tile_pdl.log 2>&1
    done
    # launch list + one full capture of the pass kernel
    ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv \
        python bench.py --steps 2 --warmup 3 --sweeps 128 --no-cpu > gpurun_out/r2_launches.log 2>&1
    ncu --set full --clock-control none --import-source on -k regex:k_poisson_pass -s 20 -c 2 -o gpurun_out/r2_pass \
        python tools/prof_one.py 4096 8 > gpurun_out/r2_ncu_pass.log 2>&1
    ;;
lagtest)
    # correctness of the lagged decision on hardware (opt-in tests; each pytest case skips itself if it needs more GPUs than
    # the box has), then the plain multi-GPU suite.  Run this on 2 GPUs: an N-GPU call is charged N x its wall time.
    n=${2:-2}
    CNV_TEST_LAG=1 timeout 1200 python -m pytest tests/test_gpu_z_multi.py -m gpu -x -q -k "lagged" > gpurun_out/r2_lag_tests_$n.log 2>&1
    echo "pytest exit $?" >> gpurun_out/r2_lag_tests_$n.log
    tail -3 gpurun_out/r2_lag_tests_$n.log
    CNV_TEST_TILE_PEER=1 timeout 900 python -m pytest tests/test_gpu_z_multi.py -m gpu -x -q -k "tile_peer" > gpurun_out/r2_tile_peer_tests_$n.log 2>&1
    echo "pytest exit $?" >> gpurun_out/r2_tile_peer_tests_$n.log
    tail -3 gpurun_out/r2_tile_peer_tests_$n.log
    timeout 1200 python -m pytest tests/test_gpu_z_multi.py -m gpu -x -q > gpurun_out/r2_multi_tests_$n.log 2>&1
    echo "pytest exit $?" >> gpurun_out/r2_multi_tests_$n.log
    tail -3 gpurun_out/r2_multi_tests_$n.log
    ;;
scale)
    # plain peer path vs lagged decision, same box, back to back: weak and strong scaling (about 4 minutes of wall time)
    n=${2:-2}
    run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 --master-port "$1" "${@:2}"; }
    run 29801 bench.py --gpus "$n" --steps 10 --warmup 3 > gpurun_out/r2_scale_peer_$n.json 2> gpurun_out/r2_scale_peer_$n.err
    CNV_PEER_LAG=1 run 29802 bench.py --gpus "$n" --steps 10 --warmup 3 > gpurun_out/r2_scale_lag_$n.json 2> gpurun_out/r2_scale_lag_$n.err
    run 29803 bench.py --gpus "$n" --steps 10 --warmup 3 --scaling strong > gpurun_out/r2_strong_peer_$n.json 2>&1
    CNV_PEER_LAG=1 run 29804 bench.py --gpus "$n" --steps 10 --warmup 3 --scaling strong > gpurun_out/r2_strong_lag_$n.json 2>&1
    tail -n 2 gpurun_out/r2_scale_peer_$n.json gpurun_out/r2_scale_lag_$n.json
    ;;
extra)
    # optional: strong scaling with the stationary-tile kernel on the (L2-resident) slabs over the NCCL exchange (the peer
    # exchange lives in the streaming kernel only), and where the pass time goes (per-CTA stamps) for both peer variants
    n=${2:-2}
    run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 --master-port "$1" "${@:2}"; }
    CNV_DIST_BACKEND=nccl CNV_POISSON_TILE=1 run 29809 tests/dist/slab_gpu_check.py 1024 1024 4 > gpurun_out/r2_tile_slab_check_$n.log 2>&1   # parity first
    for T in 4 8; do
        CNV_DIST_BACKEND=nccl CNV_POISSON_TILE=1 run 29807 bench.py --gpus "$n" --steps 10 --warmup 3 --scaling strong --T $T \
            > gpurun_out/r2_strong_tile_T${T}_$n.json 2>&1
    done
    CNV_DIST_BACKEND=nccl run 29808 bench.py --gpus "$n" --steps 10 --warmup 3 --scaling strong > gpurun_out/r2_strong_nccl_$n.json 2>&1
    # ... and with the peer exchange inside the tile kernel (only after `lagtest` showed its tests green)
    for T in 4 8; do
        CNV_POISSON_TILE=1 CNV_TILE_PEER=1 run 29810 bench.py --gpus "$n" --steps 10 --warmup 3 --scaling strong --T $T \
            > gpurun_out/r2_strong_tilepeer_T${T}_$n.json 2>&1
    done
    run 29805 tools/peer_trace.py 4096 4096 24 > gpurun_out/r2_trace_peer_$n.log 2>&1
    CNV_PEER_LAG=1 run 29806 tools/peer_trace.py 4096 4096 24 > gpurun_out/r2_trace_lag_$n.log 2>&1
    ;;
*)
    echo "usage: $0 single | lagtest N | scale N | extra N" >&2
    exit 2
    ;;
esac
