#!/bin/bash
# gpurun call 2 of round 1: tests, bench, e2e trace, config 3, then ncu captures of the bench command.
set -u
mkdir -p gpurun_out
tag="${1:-r01b}"
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest exit $?"
tail -3 gpurun_out/${tag}_pytest_gpu.log
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench exit $?"
cat gpurun_out/${tag}_bench.json
SNAPGPU_TRACE=1 python tools/e2e_probe.py > gpurun_out/${tag}_e2e_probe.log 2> gpurun_out/${tag}_e2e_trace.log; echo "e2e exit $?"
cat gpurun_out/${tag}_e2e_probe.log; grep -m3 "length binning of 100000" gpurun_out/${tag}_e2e_trace.log
timeout 600 python tools/cfg3_tail.py gpurun_out/${tag}_cfg3.json 1024 > gpurun_out/${tag}_cfg3.log 2>&1; echo "cfg3 exit $?"
cat gpurun_out/${tag}_cfg3.log | tail -3
python bench.py --steps 3 --no-cpu --no-e2e --no-tail > gpurun_out/${tag}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sha512_segments -s 4 -c 1 -o gpurun_out/${tag}_bench_sha -f \
    python bench.py --steps 3 --no-cpu --no-e2e --no-tail > gpurun_out/${tag}_ncu_bench_sha.log 2>&1; echo "ncu bench sha exit $?"
ncu --set full --clock-control none --import-source on -k regex:cmp_pairs -s 4 -c 1 -o gpurun_out/${tag}_bench_cmp -f \
    python bench.py --steps 3 --no-cpu --no-e2e --no-tail > gpurun_out/${tag}_ncu_bench_cmp.log 2>&1; echo "ncu bench cmp exit $?"
ls -la gpurun_out | grep ${tag}
