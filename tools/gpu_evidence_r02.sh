#!/bin/bash
# One gpurun call (1 GPU) that regenerates the round-2 evidence under gpurun_out/ (copied into
# profiles/ afterwards): GPU tests, smoke, both bench arms, the ncu launch list and full captures of
# the bench command, the tree-hasher timeline, host I/O scaling, lane-pair forms, stress.
# usage: tools/gpu_evidence_r02.sh [tag]
set -u
mkdir -p gpurun_out
tag="${1:-r02}"
o=gpurun_out/${tag}
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw --format=csv > ${o}_smi.csv 2>&1
nproc >> ${o}_smi.csv
python -m pytest tests -m gpu -q > ${o}_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -2 ${o}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > ${o}_smoke.log 2>&1; echo "smoke exit $?"; cat ${o}_smoke.log
python bench.py --impl reference --steps 3 --warmup 1 > ${o}_bench_reference_arm.json 2> ${o}_bench_ref.err; echo "ref exit $?"
python bench.py > ${o}_bench.json 2> ${o}_bench.err; echo "bench exit $?"
python bench.py --steps 5 --warmup 3 --no-tree --no-cfg5 > ${o}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file ${o}_bench_launches.csv \
    python bench.py --steps 5 --warmup 3 --no-tree --no-cfg5 > ${o}_ncu_launches.log 2>&1; echo "ncu launches exit $?"
python bench.py --steps 3 --no-cpu --no-e2e --no-tail --no-tree --no-cfg5 > ${o}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sha512_segments -s 4 -c 1 -o ${o}_bench_sha -f \
    python bench.py --steps 3 --no-cpu --no-e2e --no-tail --no-tree --no-cfg5 > ${o}_ncu_bench_sha.log 2>&1; echo "ncu bench sha exit $?"
ncu --set full --clock-control none --import-source on -k regex:cmp_pairs -s 4 -c 1 -o ${o}_bench_cmp -f \
    python bench.py --steps 3 --no-cpu --no-e2e --no-tail --no-tree --no-cfg5 > ${o}_ncu_bench_cmp.log 2>&1; echo "ncu bench cmp exit $?"
SNAPGPU_TRACE=1 python tools/tree_trace.py 4 2> ${o}_tree_trace.log; echo "tree trace exit $?"; grep "writeHashes:" ${o}_tree_trace.log | tail -2
tools/tree_probe.sh ${o}_tree_probe.txt; echo "tree probe exit $?"
python tools/pair_form_probe.py 4 > ${o}_pair_forms.jsonl 2> ${o}_pair_forms.err; echo "pair forms exit $?"; grep '"files": 1,' ${o}_pair_forms.jsonl | cut -c1-120
python tools/pair_form_probe.py 4 0 1,16,37,38,74,75,148,400 > ${o}_pair_regions.jsonl 2> ${o}_pair_regions.err
python tools/pair_form_probe.py 4 0 1,16,64 16 >> ${o}_pair_regions.jsonl 2>> ${o}_pair_regions.err; echo "pair regions exit $?"
python tools/long_bin_probe.py > ${o}_long_bin_probe.jsonl 2> ${o}_long_bin_probe.err; echo "long bin probe exit $?"
python tools/tree_tail_probe.py 30 > ${o}_tree_tail_probe.jsonl 2> ${o}_tree_tail_probe.err; echo "tree tail probe exit $?"; cat ${o}_tree_tail_probe.jsonl
SNAPGPU_TRACE=1 python tools/e2e_trace.py > ${o}_e2e_trace.txt 2> ${o}_e2e_trace.log; echo "e2e trace exit $?"; grep "shard done" ${o}_e2e_trace.log | tail -1
python tools/stress.py 60 2 > ${o}_stress.json 2> ${o}_stress.err; echo "stress exit $?"; tail -1 ${o}_stress.json
python tools/ncu_long_target.py 4 > ${o}_pair_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sha512_pair -c 1 -o ${o}_pair -f \
    python tools/ncu_long_target.py 4 > ${o}_ncu_pair.log 2>&1; echo "ncu pair exit $?"
python tools/shape_probe.py > ${o}_shape_probe.jsonl 2> ${o}_shape_probe.err; echo "shape probe exit $?"
python tools/tree_stress.py 60 1 > ${o}_tree_stress.json 2> ${o}_tree_stress.err; echo "tree stress exit $?"; cat ${o}_tree_stress.json
SNAPGPU_TRACE=1 python tools/tree_bench.py ${o}_tree_bench.jsonl cfg2 > ${o}_tree_bench.log 2> ${o}_tree_bench_trace.log; echo "tree bench exit $?"
python tools/tree_cfg3.py 20000 64 16 > ${o}_tree_cfg3_shape.json 2> ${o}_tree_cfg3_shape.err; echo "tree cfg3 exit $?"
python tools/h2d_piece_probe.py > ${o}_h2d_pieces.jsonl 2> ${o}_h2d_pieces.err; echo "h2d pieces exit $?"
python tools/cold_start_probe.py 3 > ${o}_cold_start.jsonl 2> ${o}_cold_start.err; echo "cold start exit $?"
[ -x tools/pin_probe ] || nvcc -O2 -gencode arch=compute_100a,code=sm_100a -o tools/pin_probe tools/pin_probe.cu 2>/dev/null
tools/pin_probe 256 2>&1 | grep -v "^  " > ${o}_pin_probe.txt; echo "pin probe exit $?"
