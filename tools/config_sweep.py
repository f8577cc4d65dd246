"""Throughput / LER of every BASELINE.json configuration through the fused pipeline (one GPU)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import qldpc_b200
from qldpc_b200 import _lib
from qldpc_b200.codes.bb_code import BB_CODES, BBCodeCircuit, make_bb_code
from qldpc_b200.noise.builder import fault_tables_for, matrices_from_tables
from qldpc_b200.noise.compiled import CompiledCircuit
from qldpc_b200.simulation.engine import ShotEngine

CONFIGS = [("[[72, 12, 6]]", 0.004, 20, 1_000_000), ("[[144, 12, 12]]", 0.005, 20, 262_144), ("[[288, 12, 18]]", 0.006, 100, 16_384),
           ("[[90, 8, 10]]", 0.004, 20, 524_288), ("[[90, 8, 10]]", 0.005, 20, 524_288), ("[[90, 8, 10]]", 0.006, 20, 524_288),
           ("[[108, 8, 10]]", 0.004, 20, 524_288), ("[[108, 8, 10]]", 0.005, 20, 524_288), ("[[108, 8, 10]]", 0.006, 20, 524_288)]
only = [a for a in sys.argv[1:] if not a.startswith('-')]
if only:
    CONFIGS = [c for c in CONFIGS if any(o in c[0] for o in only)]
rows = []
cache = {}
for name, p, max_iter, shots in CONFIGS:
    if name not in cache:
        code = make_bb_code(name)
        bb = {k: code[k] for k in ("ell", "m", "a_x_powers", "a_y_powers", "b_y_powers", "b_x_powers")}
        d = BB_CODES[name]["distance"]
        t = time.time()
        cc = CompiledCircuit.from_builder(BBCodeCircuit(code["Hx"], code["Hz"], num_cycles=d, **bb))
        ft = fault_tables_for(cc, code["Lx"], code["Lz"])
        cache[name] = (code, cc, ft, d, time.time() - t)
    code, cc, ft, d, tb = cache[name]
    M = matrices_from_tables(ft, p, d)
    batch = min(65536, shots) if name != "[[288, 12, 18]]" else 8192
    eng = ShotEngine(cc, code["Lx"], code["Lz"], M, max_batch=batch)
    cfg = _lib.make_config(max_iter, _lib.QB_ALPHA_DYNAMIC, precision=int(os.environ.get("QB_PRECISION", "0")))
    eng.pipeline.run(1, 0, batch, p, cfg)
    counts, _ = eng.pipeline.run(1234, 0, shots, p, cfg)
    st = eng.pipeline.stats()
    row = dict(code=name, p=p, max_iter=max_iter, shots=int(counts[3]), shots_per_s=round(shots / (st["ms_total"] * 1e-3)),
               ler=round(counts[2] / counts[3], 5), z_ler=round(counts[0] / counts[3], 5), x_ler=round(counts[1] / counts[3], 5),
               nonconverged_side_frac=round((counts[4] + counts[5]) / (2 * counts[3]), 4),
               edge_messages_per_s=round(st["edge_messages"] / (st["ms_minsum"] * 1e-3) / 1e9, 1),
               ms=dict(minsum=round(st["ms_minsum"], 1), osd=round(st["ms_osd"], 1), sample=round(st["ms_sample"], 1)),
               table_build_s=round(tb, 2), osd_tier_exits=dict(z=eng.decZ.osd_stats(), x=eng.decX.osd_stats()),
               precision="half2" if os.environ.get("QB_PRECISION", "0") == "1" else "f32")
    print(json.dumps(row), flush=True)
    eng.close()
