import sys, time, shutil, os
sys.path.insert(0, '/root/repo')
from pathlib import Path
import numpy as np
from snappy_b200 import _native as N, build, synth
N.init([0])
src = open('/root/repo/tools/tree_bench.py').read()
ns = {}
exec("import shutil, os\nfrom pathlib import Path\nimport numpy as np\nfrom snappy_b200 import synth\n" + src[src.index('def make_tree'):src.index('for cfg in configs')], ns)
root = Path('/dev/shm/snapgpu_tree/cfg2')
tar = ns['make_tree'](root, synth.lognormal_sizes(100_000))
build.hashes_yaml(str(root), tar)
for rep in range(3):
    stage = Path('/dev/shm/snapgpu_tree/stage')
    if stage.exists():
        t0 = time.perf_counter(); shutil.rmtree(stage); print('rmtree', time.perf_counter() - t0, file=sys.stderr)
    N.lib().snapgpu_digest_cache_clear()
    t0 = time.perf_counter()
    build.copyToBuildDir(str(root), str(stage), no_link=True)
    print('copy ms', (time.perf_counter() - t0) * 1e3, file=sys.stderr)
shutil.rmtree('/dev/shm/snapgpu_tree')
