#!/bin/bash
# A/B timing of the gradient-sync modes on 2 GPUs (ms per step, graph?, buckets?)
pick='import sys,json
for l in sys.stdin:
    l=l.strip()
    if l.startswith("{"):
        d=json.loads(l); print(round(d["ms_per_step"],4), d["config"].get("cuda_graph"), d["config"].get("grad_buckets"))'
run() { port=$1; shift; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 2 --steps 20 --warmup 5 --no_baseline --no_e2e "$@" 2>/dev/null | python -c "$pick"; }
echo "A buckets+pdl graph"; run 29601
echo "B buckets no-pdl graph"; LSTM_TS_BUCKET_PDL=0 run 29602
echo "C no buckets graph"; run 29603 --grad_buckets 0
echo "D buckets+pdl eager"; run 29604 --cuda_graph 0
echo "E no wavefront buckets+pdl graph"; LSTM_TS_WAVEFRONT=0 run 29605
echo "F no wavefront no buckets graph"; LSTM_TS_WAVEFRONT=0 run 29606 --grad_buckets 0
echo "G N=1 graph"; timeout 200 python bench.py --steps 20 --warmup 5 --no_baseline --no_e2e 2>/dev/null | python -c "$pick"
