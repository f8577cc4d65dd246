"""Per-iteration vs per-shot cost of the min-sum kernel: run the pipeline with 10/20/40 iterations (OSD off)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import qldpc_b200
from qldpc_b200 import _lib
from qldpc_b200.codes.bb_code import BB_CODES, BBCodeCircuit, make_bb_code
from qldpc_b200.noise.builder import fault_tables_for, matrices_from_tables
from qldpc_b200.noise.compiled import CompiledCircuit
from qldpc_b200.simulation.engine import ShotEngine
name, p, B = "[[144, 12, 12]]", 0.005, 65536
code = make_bb_code(name)
bb = {k: code[k] for k in ("ell", "m", "a_x_powers", "a_y_powers", "b_y_powers", "b_x_powers")}
d = BB_CODES[name]["distance"]
cc = CompiledCircuit.from_builder(BBCodeCircuit(code["Hx"], code["Hz"], num_cycles=d, **bb))
ft = fault_tables_for(cc, code["Lx"], code["Lz"])
M = matrices_from_tables(ft, p, d)
eng = ShotEngine(cc, code["Lx"], code["Lz"], M, max_batch=B)
res = {}
for it in (10, 20, 40):
    cfg = _lib.make_config(it, _lib.QB_ALPHA_DYNAMIC, use_osd=False)
    eng.pipeline.run(1, 0, B, p, cfg)
    eng.pipeline.run(1234, 0, B, p, cfg)
    st = eng.pipeline.stats()
    res[it] = st["ms_minsum"]
    print(it, "iterations: min-sum ms", round(st["ms_minsum"], 2), "edge messages", st["edge_messages"])
per_it = (res[40] - res[20]) / 20
print("per iteration ms (2 x 65536 sides): %.3f ; fixed per launch pair: %.2f ms = %.1f %% of the 20-iteration time" % (per_it, res[20] - 20 * per_it, 100 * (res[20] - 20 * per_it) / res[20]))
cyc = per_it * 1e-3 * 1.965e9 / (2 * B / 148)
print("cycles per iteration per shot-side per SM: %.0f" % cyc)
eng.close()
