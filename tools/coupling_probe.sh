#!/bin/bash
# Development aid: does a GPU run the 1/8-shard interact kernel slower when the other GPUs of the box are busy too?
N=${1:-8}
echo "== alone (GPU 0)"; CUDA_VISIBLE_DEVICES=0 PROBE_FRAC=8 PROBE_GROUP=0 PROBE_GPC=0 python tools/shard_probe.py 2>&1 | tail -1
echo "== $N at once"
for g in $(seq 0 $((N-1))); do CUDA_VISIBLE_DEVICES=$g PROBE_FRAC=8 PROBE_GROUP=0 PROBE_GPC=0 python tools/shard_probe.py > /tmp/coupling_$g.log 2>&1 & done
wait
for g in $(seq 0 $((N-1))); do tail -1 /tmp/coupling_$g.log; done
