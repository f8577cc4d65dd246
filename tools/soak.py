"""Soak run: many batches of every BASELINE code through both pipeline entry points (device sampler and host events),
alternating precisions and batch sizes; any CUDA fault or count mismatch between the two workspaces' batches aborts."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests")); sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import helpers, qldpc_b200
from qldpc_b200 import _lib
from qldpc_b200.simulation.engine import ShotEngine
t0 = time.time()
total = 0
for tag, p, it, batch, nb in (("144", 0.005, 20, 65536, 24), ("144", 0.006, 20, 30000, 9), ("72", 0.004, 20, 65536, 20), ("90", 0.006, 20, 50000, 8),
                              ("108", 0.005, 20, 65536, 8), ("288", 0.006, 100, 4096, 3), ("144", 0.001, 20, 65536, 6), ("72", 0.02, 20, 8192, 6)):
    s = helpers.code_setup(tag); M = helpers.matrices(tag, p)
    eng = ShotEngine(s["cc"], s["Lx"], s["Lz"], M, max_batch=batch)
    for prec in ((0, 1) if tag != "288" else (0,)):
        cfg = _lib.make_config(it, _lib.QB_ALPHA_DYNAMIC, precision=prec)
        c, f = eng.pipeline.run(99, 7 * batch, nb * batch + 123, p, cfg)
        c2, _ = eng.pipeline.run(99, 7 * batch, nb * batch + 123, p, cfg)
        assert np.array_equal(c, c2), (tag, p, prec, c, c2)
        total += 2 * int(c[3])
        print(tag, p, "precision", prec, "shots", int(c[3]), "LER %.4f" % (c[2] / c[3]), "Z", eng.decZ.osd_stats()["tier_b"], flush=True)
    eng.close()
print("soak ok:", total, "shots in %.0f s" % (time.time() - t0))
