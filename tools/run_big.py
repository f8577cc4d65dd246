import sys, time
import numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import helpers, qldpc_b200
from qldpc_b200 import _lib
from qldpc_b200.simulation.engine import ShotEngine
B = int(sys.argv[1]); tag = sys.argv[2] if len(sys.argv) > 2 else "144"; p = float(sys.argv[3]) if len(sys.argv) > 3 else 0.005
s = helpers.code_setup(tag); M = helpers.matrices(tag, p)
eng = ShotEngine(s["cc"], s["Lx"], s["Lz"], M, max_batch=B)
import os
cfg = _lib.make_config(int(os.environ.get("QB_MAX_ITER", "20")), _lib.QB_ALPHA_DYNAMIC, precision=int(os.environ.get("QB_PRECISION", "0")), use_osd=os.environ.get("QB_NO_OSD") is None)
for it in range(int(sys.argv[4]) if len(sys.argv) > 4 else 3):
    t0 = time.time()
    counts, _ = eng.pipeline.run(1234, it * B, B, p, cfg)
    dt = time.time() - t0
    print(B, counts.tolist(), f"{dt*1e3:.1f} ms", eng.pipeline.stats(), "Z", eng.decZ.osd_stats(), "X", eng.decX.osd_stats(), flush=True)
