TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521"
export CNV_DIST_BACKEND=peer
timeout 200 $TR tests/dist/slab_gpu_check.py 320 96 4 > gpurun_out/n8_check_a.log 2>&1; echo "check_a rc=$?"
timeout 200 $TR tests/dist/slab_gpu_check.py 1024 1024 8 > gpurun_out/n8_check_b.log 2>&1; echo "check_b rc=$?"
timeout 200 $TR tests/dist/slab_stress.py 800 96 4 6 > gpurun_out/n8_stress.log 2>&1; echo "stress rc=$?"
timeout 300 $TR bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/n8_weak_peer.log 2>&1; echo "weak peer rc=$?"
CNV_DIST_BACKEND=nccl timeout 300 $TR bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/n8_weak_nccl.log 2>&1; echo "weak nccl rc=$?"
timeout 300 $TR bench.py --gpus 8 --steps 3 --warmup 3 --n 16384 --scaling strong > gpurun_out/n8_c5_peer.log 2>&1; echo "c5 rc=$?"
timeout 300 $TR bench.py --gpus 8 --steps 5 --warmup 3 --scaling strong > gpurun_out/n8_strong_peer.log 2>&1; echo "strong rc=$?"
for f in n8_check_a n8_check_b n8_stress; do tail -n 2 gpurun_out/$f.log; done
grep -h '"metric"' gpurun_out/n8_weak_peer.log gpurun_out/n8_weak_nccl.log gpurun_out/n8_c5_peer.log gpurun_out/n8_strong_peer.log | cut -c1-330
