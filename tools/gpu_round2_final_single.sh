set -u
bash tools/gpu_single_check.sh tests
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_stencil.csv python -c "
import fluid_dynamics1_b200 as fd
sim = fd.Simulation(dict(nx=4096, ny=4096, Re=1000.0, dt=5e-6, poisson_max_it=100000))
fd.lib().cnv_sim_stencil_phase(sim.h, 6, None); fd.lib().cnv_device_synchronize()
" > gpurun_out/launches_stencil.log 2>&1
: > gpurun_out/probe_onchip_final.log
for shape in "64 64" "128 128" "256 256" "512 512" "1024 1024" "768 1536"; do
    CNV_ONCHIP_PROF=1 timeout 300 python tools/probe_poisson.py $shape "8:0:0,o0:0:0" 2048 >> gpurun_out/probe_onchip_final.log 2>&1
done
bash tools/gpu_single_check.sh bench ncu
timeout 300 python tools/probe_small.py > gpurun_out/probe_small.log 2>&1
: > gpurun_out/cases.log; for c in "default" "high_re" "c3 10000" "c4 3"; do timeout 600 python tools/run_case.py $c >> gpurun_out/cases.log 2>&1; done
