python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r1_n1.json 2> gpurun_out/bench_r1_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 3 --sweeps 128 --no-cpu > gpurun_out/b_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_poisson_pass -s 6 -c 2 -f -o gpurun_out/prof_pass_r1 python tools/prof_one.py 4096 8 > gpurun_out/b_ncu2.log 2>&1
python tools/run_case.py default > gpurun_out/cases.log 2>&1
python tools/run_case.py high_re >> gpurun_out/cases.log 2>&1
python tools/run_case.py c3 100 >> gpurun_out/cases.log 2>&1
cat gpurun_out/cases.log; cut -c1-600 gpurun_out/bench_r1_n1.json
