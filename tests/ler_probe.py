"""Paired comparison with the real reference's per-shot flags (tests/golden/ler.npz) on identical faults, all configurations
(GPU box; prints one line per configuration)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import helpers
import qldpc_b200
from qldpc_b200 import _lib
from qldpc_b200.simulation.engine import ShotEngine
from test_gpu_osd_pipeline import _reference_stream_events, LER_CASES
g = np.load(os.path.join(helpers.GOLDEN, "ler.npz"))
for tag, p, max_iter, shots, _ in LER_CASES:
    key = f"{tag}_{int(round(p * 1e4))}"
    if key + "_z" not in g.files:
        continue
    rz = np.unpackbits(g[key + "_z"], bitorder="little")[:shots].astype(bool)
    rx = np.unpackbits(g[key + "_x"], bitorder="little")[:shots].astype(bool)
    s = helpers.code_setup(tag); M = helpers.matrices(tag, p)
    n = min(shots, 4000)
    eng = ShotEngine(s["cc"], s["Lx"], s["Lz"], M, max_batch=n)
    ev_ptr, ev = _reference_stream_events(s["cc"], n, p)
    counts, flags = eng.pipeline.run_events(ev_ptr, ev, _lib.make_config(max_iter, _lib.QB_ALPHA_DYNAMIC))
    eng.close()
    ez, ex = (flags & 1) != 0, (flags & 2) != 0
    print(f"{key}: same faults, {n} shots: reference errors z/x/total {rz[:n].sum()}/{rx[:n].sum()}/{(rz|rx)[:n].sum()}  "
          f"GPU {ez.sum()}/{ex.sum()}/{(ez|ex).sum()}  per-shot flag agreement z {np.mean(ez == rz[:n]):.4f} x {np.mean(ex == rx[:n]):.4f}", flush=True)
