"""Per-side cycle counts of osd0_kernel against pivots / candidates (needs a library built with
QB_EXTRA_NVCC_FLAGS=-DQB_OSD_PROFILE, loaded through QLDPC_B200_LIB)."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests")); sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import helpers
import qldpc_b200
from qldpc_b200 import _lib
from qldpc_b200.simulation.engine import ShotEngine

tag, p, shots = "144", 0.005, int(sys.argv[1]) if len(sys.argv) > 1 else 65536
s = helpers.code_setup(tag); M = helpers.matrices(tag, p)
eng = ShotEngine(s["cc"], s["Lx"], s["Lz"], M, max_batch=shots)
cfg = _lib.make_config(20, _lib.QB_ALPHA_DYNAMIC)
lib = _lib.load()
buf = np.zeros((1 << 18, 8), np.uint64)
eng.pipeline.run(1, 0, shots, p, cfg)
lib.qb_debug_osd_profile(buf.ctypes.data_as(C.c_void_p), 1 << 18)
eng.pipeline.run(1234, 0, shots, p, cfg)
n = lib.qb_debug_osd_profile(buf.ctypes.data_as(C.c_void_p), 1 << 18)
a = buf[:n].astype(np.float64)
t, c, cyc, gt = a[:, 0], a[:, 1], a[:, 2], a[:, 3]
print("sides", n, "stats", eng.pipeline.stats())
print("cycles per side: mean %.0f  p50 %.0f p90 %.0f p99 %.0f max %.0f" % (cyc.mean(), *np.percentile(cyc, [50, 90, 99, 100])))
ph = a[:, 4:8]
print("phase cycles per side (mean): setup+residual %.0f, histogram %.0f, windows (search, collect, rank) %.0f, write-back %.0f, elimination %.0f" % (
    *ph.mean(axis=0), (cyc - ph.sum(axis=1)).mean()))
print("phase share of all cycles: setup %.3f hist %.3f windows %.3f writeback %.3f elimination %.3f" % (
    *(ph.sum(axis=0) / cyc.sum()), 1 - ph.sum() / cyc.sum()))
edges = [0, 50, 100, 150, 200, 300, 400, 500, 700, 1100]
print("pivots bin: sides, share of sides, share of cycles, mean cycles, cycles per pivot")
for lo, hi in zip(edges[:-1], edges[1:]):
    k = (t >= lo) & (t < hi)
    if k.any():
        print(f"[{lo:4d},{hi:4d}) {k.sum():7d} {k.mean():6.3f} {cyc[k].sum() / cyc.sum():6.3f} {cyc[k].mean():10.0f} {cyc[k].sum() / max(1, t[k].sum()):8.0f}")
# the two launches (Z, X): records are appended in completion order, the second launch starts at the largest gap of end times
gap = int(np.argmax(np.diff(gt))) + 1
for name, sl in (("first launch", np.arange(0, gap)), ("second launch", np.arange(gap, n))):
    if len(sl) < 2: continue
    e = gt[sl]; dur = (e.max() - e.min()) * 1e-6
    late = e > e.max() - 0.05 * (e.max() - e.min())
    hist, _ = np.histogram(e, bins=20)
    print(name, "completions per 5 % of the launch:", hist.tolist())
    if len(sl) == 0 or not late.any(): continue
    print(f"{name}: {len(sl)} sides, end-time span {dur:.2f} ms; sides ending in the last 10 % of it: {late.sum()} (mean pivots {t[sl][late].mean():.0f}, max {t[sl][late].max():.0f})")
