#!/bin/bash
# Host-side scaling of the tree packer on this box: raw open/read/close probes and the fake-GPU
# pipeline (tests/hostsim, no hashing) by thread count.  usage: tools/tree_probe.sh OUT
out=${1:-gpurun_out/tree_probe.txt}
python - <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np
from pathlib import Path
from snappy_b200 import synth
sys.argv = ["x"]
import bench
lengths = synth.lognormal_sizes(100_000)
offs, total = synth.layout(lengths)
data = np.random.default_rng(1).integers(0, 256, total, dtype=np.uint8)
import shutil
root = Path("/dev/shm/snapgpu_probe_tree")
shutil.rmtree(root, ignore_errors=True)
bench.materialise_tree(root, data, offs, lengths)
PY
head -c 1 /dev/zero > /dev/shm/snapgpu_probe_tar
{
echo "== nproc $(nproc); lscpu:"; lscpu | grep -E "Model name|Thread|Core|Socket|NUMA"
for t in 1 8 16 24 32; do for m in 2; do UNSHARE=1 tools/io_probe /dev/shm/snapgpu_probe_tree $t $m | tail -1; done; done
for t in 16; do tools/io_probe /dev/shm/snapgpu_probe_tree $t 2 | tail -1; NOREAD=1 UNSHARE=1 tools/io_probe /dev/shm/snapgpu_probe_tree $t 2 | tail -1; done
make -s -C tests/hostsim _build/hostsim_fast
for t in 8 16 24 32 48; do echo "-- hostsim pipeline, $t pool threads"; SNAPGPU_PACK_THREADS=$t SNAPGPU_TRACE=1 HOSTSIM_NOHASH=1 tests/hostsim/_build/hostsim_fast repeat 4 /dev/shm/snapgpu_probe_tree /dev/shm/snapgpu_probe_tar 2>&1 >/dev/null | grep writeHashes | tail -2; done
} > $out 2>&1
rm -rf /dev/shm/snapgpu_probe_tree /dev/shm/snapgpu_probe_tar
