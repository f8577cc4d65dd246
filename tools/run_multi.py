"""Several batches in ONE qb_pipeline_run call (double-buffered workspaces) against the same shots in separate calls."""
import sys, time
import numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import helpers, qldpc_b200
from qldpc_b200 import _lib
from qldpc_b200.simulation.engine import ShotEngine
B = int(sys.argv[1]); K = int(sys.argv[2]); tag = sys.argv[3] if len(sys.argv) > 3 else "144"; p = float(sys.argv[4]) if len(sys.argv) > 4 else 0.005
s = helpers.code_setup(tag); M = helpers.matrices(tag, p)
eng = ShotEngine(s["cc"], s["Lx"], s["Lz"], M, max_batch=B)
cfg = _lib.make_config(20, _lib.QB_ALPHA_DYNAMIC)
eng.pipeline.run(1234, 0, B, p, cfg)
tot = np.zeros(8, dtype=np.int64)
t0 = time.time()
for it in range(K):
    c, _ = eng.pipeline.run(1234, it * B, B, p, cfg); tot += c
t_sep = time.time() - t0
t0 = time.time()
c2, _ = eng.pipeline.run(1234, 0, B * K, p, cfg)
t_one = time.time() - t0
c2b, f2 = eng.pipeline.run(1234, 0, B * K, p, cfg, want_flags=True)     # (pageable flag copies serialise the host: not timed)
assert np.array_equal(c2, c2b)
print("separate calls:", tot.tolist(), f"{t_sep*1e3/K:.2f} ms/batch")
print("one call      :", c2.tolist(), f"{t_one*1e3/K:.2f} ms/batch", eng.pipeline.stats())
assert np.array_equal(tot, c2)
c3, f3 = eng.pipeline.run(1234, B, B, p, cfg, want_flags=True)
assert np.array_equal(f3, f2[B:2 * B]), "flags of the second batch (second workspace) differ"
print("ok")
