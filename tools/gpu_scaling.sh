#!/bin/bash
# multi-GPU round: run with gpurun --gpus N.  torchrun bench at 1..N, in-process sharding check.
set -u
mkdir -p gpurun_out
tag="${1:-r01d}"; n="${2:-2}"
nvidia-smi --query-gpu=index,name --format=csv,noheader
python tools/multi_gpu_check.py > gpurun_out/${tag}_multi_inproc.jsonl 2> gpurun_out/${tag}_multi_inproc.err; echo "inproc exit $?"
cat gpurun_out/${tag}_multi_inproc.jsonl; tail -3 gpurun_out/${tag}_multi_inproc.err
for g in 1 2 4 8; do
  [ $g -le $n ] || continue
  if [ $g -eq 1 ]; then
    python bench.py --gpus 1 --no-cmp --no-tail > gpurun_out/${tag}_scale_$g.json 2> gpurun_out/${tag}_scale_$g.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $g --no-cmp --no-tail > gpurun_out/${tag}_scale_$g.json 2> gpurun_out/${tag}_scale_$g.err
  fi
  echo "bench --gpus $g exit $?"; tail -1 gpurun_out/${tag}_scale_$g.json | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); print(d['n_gpus'], 'GPUs: value', round(d['value'],1), d['unit'], 'e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],3))"
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 \
   bench.py --gpus $n --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_scale_ref_$n.json 2>&1; echo "ref arm at $n exit $?"
tail -c 600 gpurun_out/${tag}_scale_ref_$n.json
