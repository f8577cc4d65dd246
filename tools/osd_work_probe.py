"""Work counters of the free-row OSD kernel (library built with QB_EXTRA_NVCC_FLAGS=-DQB_OSD_STATS)."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests")); sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import helpers, qldpc_b200
from qldpc_b200 import _lib
from qldpc_b200.simulation.engine import ShotEngine
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
tag = sys.argv[2] if len(sys.argv) > 2 else "144"; p = float(sys.argv[3]) if len(sys.argv) > 3 else 0.005
s = helpers.code_setup(tag); M = helpers.matrices(tag, p)
eng = ShotEngine(s["cc"], s["Lx"], s["Lz"], M, max_batch=B)
cfg = _lib.make_config(20, _lib.QB_ALPHA_DYNAMIC)
counts, _ = eng.pipeline.run(1234, 0, B, p, cfg)
print(counts.tolist(), eng.pipeline.stats())
lib = _lib.load()
for nm, dec in (("Z", eng.decZ), ("X", eng.decX)):
    out = np.zeros(40, np.int32)
    rc = lib.qb_debug_osd_work(dec._h, out.ctypes.data_as(C.c_void_p))
    sides, piv, blocks, hitb, rows, cands, R = [int(x) for x in out[:7]]
    print(nm, "rc", rc, "sides", sides, "| per side: pivots %.1f candidates %.1f touched rows %.1f | per pivot: blocks scanned %.2f, blocks with a hit %.2f, rows updated %.2f"
          % (piv / max(1, sides), cands / max(1, sides), R / max(1, sides), blocks / max(1, piv), hitb / max(1, piv), rows / max(1, piv)))
    for nm2, o, cnt, cyc in (("tier A", 8, 16, 18), ("tier B", 12, 17, 19)):
        print("   ", nm2, "sides", int(out[cnt]), "mean k-cycles %.1f" % (out[cyc] * 0.256 / max(1, out[cnt])), "| slowest side: k-cycles %.0f pivots %d candidates %d rows %d" % (out[o] * 0.256, out[o + 1], out[o + 2], out[o + 3]))
    for nm2, o in (("select 1", 20), ("select 2", 26)):
        k = max(1, int(out[o]))
        print("   ", nm2, "sides", int(out[o]), "| k-cycles per side: residual+histogram %.1f, window scan %.1f, scatter %.1f, rank+write %.1f | candidates written %.0f"
              % tuple([out[o + i] * 0.256 / k for i in (1, 2, 3, 4)] + [out[o + 5] / k]))
    print("    tier A phases, k-cycles per side: setup %.1f, candidate loop %.1f, back substitution %.1f" % (out[7] * 0.256 / max(1, out[16]), out[32] * 0.256 / max(1, out[16]), (out[18] - out[7] - out[32]) * 0.256 / max(1, out[16])))
