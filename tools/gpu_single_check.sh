#!/bin/bash
# Single-GPU hardware check: GPU parity suite, on-chip kernel probes, bench.py, launch list + ncu captures.
#   gpurun --timeout 1800 -- 'bash tools/gpu_single_check.sh [tests|probe|bench|ncu ...]'   (default: all)
set -u
mkdir -p gpurun_out
what=${*:-tests probe bench ncu}
for w in $what; do
case "$w" in
tests)
    timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
    echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
    tail -5 gpurun_out/pytest_gpu.log
    ;;
octests)
    timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "onchip" > gpurun_out/pytest_onchip.log 2>&1
    echo "pytest exit $?" >> gpurun_out/pytest_onchip.log
    tail -5 gpurun_out/pytest_onchip.log
    ;;
probe)
    : > gpurun_out/probe_onchip.log
    for shape in "1024 1024" "512 512" "256 256" "128 128" "64 64" "768 1536"; do
        CNV_ONCHIP_PROF=1 timeout 300 python tools/probe_poisson.py $shape "8:0:0,o0:0:0,o4:12:12,o2:12:12,o4:16:9,o4:9:16,o2:0:0,o6:0:0,o8:0:0" 2048 >> gpurun_out/probe_onchip.log 2>&1
    done
    tail -60 gpurun_out/probe_onchip.log
    ;;
ncuoc)
    CNV_POISSON_ONCHIP=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_poisson_onchip -c 1 -f -o gpurun_out/ncu_onchip \
        python tools/prof_onchip.py 1024 64 > gpurun_out/ncu_onchip.log 2>&1
    ;;
bench)
    timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
    echo "bench exit $?"; tail -c 400 gpurun_out/bench_n1.err
    ;;
ncu)
    timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
        python bench.py --steps 2 --warmup 3 --sweeps 128 --no-cpu --series single > gpurun_out/launches.log 2>&1
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_poisson_pass -s 2 -c 1 -f -o gpurun_out/ncu_pass \
        python tools/prof_one.py 4096 8 > gpurun_out/ncu_pass.log 2>&1
    CNV_POISSON_ONCHIP=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_poisson_onchip -c 1 -f -o gpurun_out/ncu_onchip \
        python tools/prof_onchip.py 1024 64 > gpurun_out/ncu_onchip.log 2>&1
    ;;
esac
done
