#!/bin/bash
# One gpurun call: GPU parity tests, smoke, both bench arms, then the ncu launch list of the
# bench command and one full capture of each hot kernel.  Outputs under gpurun_out/.
set -u
mkdir -p gpurun_out
tag="${1:-r01}"
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_smi.csv 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest exit $?"
tail -3 gpurun_out/${tag}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke exit $?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err; echo "ref exit $?"
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench exit $?"
cat gpurun_out/${tag}_bench.json
python bench.py --steps 5 --warmup 3 > gpurun_out/${tag}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 5 --warmup 3 > gpurun_out/${tag}_ncu_launches.log 2>&1; echo "ncu launches exit $?"
python tools/ncu_target.py cfg2 0 0 3 > gpurun_out/${tag}_plain_sha.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sha512 -s 2 -c 1 -o gpurun_out/${tag}_sha_cfg2 -f \
    python tools/ncu_target.py cfg2 0 0 3 > gpurun_out/${tag}_ncu_sha.log 2>&1; echo "ncu sha exit $?"
python tools/ncu_target.py cmp 0 0 3 > gpurun_out/${tag}_plain_cmp.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:cmp_pairs -s 2 -c 1 -o gpurun_out/${tag}_cmp -f \
    python tools/ncu_target.py cmp 0 0 3 > gpurun_out/${tag}_ncu_cmp.log 2>&1; echo "ncu cmp exit $?"
ls -la gpurun_out | tail -20
