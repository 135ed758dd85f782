#!/bin/bash
# A/B of the peer path's CTA launch order (CNV_PEER_EDGE_FIRST) on N GPUs: headline workload only (bench.py --series
# headline), optionally the per-CTA trace (tools/peer_trace.py).
#   gpurun --gpus N --timeout 900 -- 'WITH_TRACE=1 bash tools/gpu_edge_ab.sh N "1 0"'
set -u
n=${1:-2}
variants=${2:-"1 0"}
mkdir -p gpurun_out
export CNV_PEER_TIMEOUT_MS=${CNV_PEER_TIMEOUT_MS:-30000}
port=29810
for edge in $variants; do
    port=$((port + 1))
    CNV_PEER_EDGE_FIRST=$edge timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 \
        --master-port $port bench.py --gpus "$n" --steps ${STEPS:-10} --warmup 3 --series headline > gpurun_out/edge_ab_${n}_edge${edge}.json 2> gpurun_out/edge_ab_${n}_edge${edge}.err
    echo "edge_first=$edge exit $?: $(python -c "
import json,sys
for l in open('gpurun_out/edge_ab_${n}_edge${edge}.json'):
    if l.startswith('{'):
        d=json.loads(l); print('value %.4e ms/step %.3f launch_us %.1f parity %s e2e %.4e' % (d['value'], d['ms_per_step'], d['roofline']['launch_us'], d['parity_check'], d['e2e']['value']))
" 2>&1)"
done
if [ "${WITH_TRACE:-0}" = "1" ]; then
    for edge in $variants; do
        port=$((port + 1))
        CNV_PEER_EDGE_FIRST=$edge timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 \
            --master-port $port tools/peer_trace.py 2048 16384 24 > gpurun_out/peer_trace_${n}_edge${edge}.log 2>&1
        echo "trace edge=$edge exit $?"
    done
fi
