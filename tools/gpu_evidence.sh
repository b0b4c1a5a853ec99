#!/bin/bash
# One gpurun call that regenerates the evidence under profiles/ for a round (1 GPU):
# GPU tests, smoke, both bench arms, ncu launch list + full captures of the bench command,
# tree benchmark, config 3 tail latency.  usage: gpu_evidence.sh <tag> [skip-cfg3]
set -u
mkdir -p gpurun_out
tag="${1:-r01}"; skip="${2:-}"
o=gpurun_out/${tag}
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw --format=csv > ${o}_smi.csv 2>&1
python -m pytest tests -m gpu -x -q > ${o}_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -2 ${o}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > ${o}_smoke.log 2>&1; echo "smoke exit $?"; cat ${o}_smoke.log
python bench.py --impl reference --steps 3 --warmup 1 > ${o}_bench_reference_arm.json 2> ${o}_bench_ref.err; echo "ref exit $?"
python bench.py > ${o}_bench.json 2> ${o}_bench.err; echo "bench exit $?"
python - <<PY
import json
d = json.loads(open("${o}_bench.json").readlines()[-1])
print("value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "kernel", round(d["roofline"]["kernel_ms"], 3),
      "frac", round(d["roofline"]["frac"], 3), "e2e", round(d["e2e"]["value"], 1), "cmp frac", round(d["roofline_cmp"]["frac"], 3),
      "cpu", round(d["cpu_baseline"]["value"], 1), "clocks", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
PY
python bench.py --steps 5 --warmup 3 > ${o}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file ${o}_bench_launches.csv \
    python bench.py --steps 5 --warmup 3 > ${o}_ncu_launches.log 2>&1; echo "ncu launches exit $?"
python bench.py --steps 3 --no-cpu --no-e2e --no-tail > ${o}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sha512_segments -s 4 -c 1 -o ${o}_bench_sha -f \
    python bench.py --steps 3 --no-cpu --no-e2e --no-tail > ${o}_ncu_bench_sha.log 2>&1; echo "ncu bench sha exit $?"
ncu --set full --clock-control none --import-source on -k regex:cmp_pairs -s 4 -c 1 -o ${o}_bench_cmp -f \
    python bench.py --steps 3 --no-cpu --no-e2e --no-tail > ${o}_ncu_bench_cmp.log 2>&1; echo "ncu bench cmp exit $?"
SNAPGPU_TRACE=1 python tools/tree_bench.py ${o}_tree.jsonl > ${o}_tree.log 2> ${o}_tree_trace.log; echo "tree exit $?"; cat ${o}_tree.log
grep "writeHashes:" ${o}_tree_trace.log | tail -3; grep "batch " ${o}_tree_trace.log | tail -6
python tools/long_probe.py 16 > ${o}_long_probe.jsonl 2> ${o}_long_probe.err; echo "long probe exit $?"; grep '"files": 1,' ${o}_long_probe.jsonl
python tools/stress.py 60 2 > ${o}_stress.json 2> ${o}_stress.err; echo "stress exit $?"; tail -1 ${o}_stress.json
python tools/ncu_long_target.py 4 > ${o}_pair_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sha512_pair -c 1 -o ${o}_pair -f \
    python tools/ncu_long_target.py 4 > ${o}_ncu_pair.log 2>&1; echo "ncu pair exit $?"
if [ -z "$skip" ]; then
  timeout 900 python tools/cfg3_tail.py ${o}_cfg3.json 1024 > ${o}_cfg3.log 2>&1; echo "cfg3 exit $?"; tail -1 ${o}_cfg3.log
fi
