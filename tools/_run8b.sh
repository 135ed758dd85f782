TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541"
timeout 400 $TR bench.py --gpus 8 --steps 3 --warmup 3 --grid 16384 --scaling strong > gpurun_out/n8_c5_peer.log 2>&1; echo "c5 rc=$?"
grep -h '"metric"' gpurun_out/n8_c5_peer.log | cut -c1-1500
tail -3 gpurun_out/n8_c5_peer.log | cut -c1-300
