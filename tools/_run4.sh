TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531"
export CNV_DIST_BACKEND=peer
timeout 200 $TR tests/dist/slab_gpu_check.py 160 96 4 > gpurun_out/n4_check_a.log 2>&1; echo "check_a rc=$?"
timeout 200 $TR tests/dist/slab_sim_gpu_check.py 128 4 > gpurun_out/n4_sim.log 2>&1; echo "sim rc=$?"
timeout 200 $TR tests/dist/slab_stress.py 1024 1024 8 4 > gpurun_out/n4_stress.log 2>&1; echo "stress rc=$?"
timeout 300 $TR bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/n4_weak_peer.log 2>&1; echo "weak peer rc=$?"
timeout 300 $TR bench.py --gpus 4 --steps 5 --warmup 3 --scaling strong > gpurun_out/n4_strong_peer.log 2>&1; echo "strong rc=$?"
for f in n4_check_a n4_sim n4_stress; do tail -n 1 gpurun_out/$f.log; done
grep -h '"metric"' gpurun_out/n4_weak_peer.log gpurun_out/n4_strong_peer.log | cut -c1-200
