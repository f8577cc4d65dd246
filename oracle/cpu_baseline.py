"""CPU baseline runner (TEST/BENCH INFRASTRUCTURE): the oracle port of the reference's per-shot loop
(_run_single_trial_fast, src/simulation/engine.py:68-122) on all host cores, structured like the
reference: a spawn pool, one task per shot, np.random.seed(base_seed + shot) per shot."""
import os
import sys
import time

_state = {}


def _init(root, tag_name, p, max_iter):
    sys.path.insert(0, root)
    import numpy as np
    import qldpc_b200  # noqa: F401  (host-side table builder only; no CUDA call)
    from qldpc_b200.codes.bb_code import BB_CODES, BBCodeCircuit, make_bb_code
    from qldpc_b200.noise.builder import fault_tables_for, matrices_from_tables
    from qldpc_b200.noise.compiled import CompiledCircuit
    from oracle import oracle as orc
    code = make_bb_code(tag_name)
    bb = {k: code[k] for k in ("ell", "m", "a_x_powers", "a_y_powers", "b_y_powers", "b_x_powers")}
    d = BB_CODES[tag_name]["distance"]
    cc = CompiledCircuit.from_builder(BBCodeCircuit(code["Hx"], code["Hz"], num_cycles=d, **bb))
    M = matrices_from_tables(fault_tables_for(cc, code["Lx"], code["Lz"]), p, d)
    m, k = M["first_logical_rowZ"], code["Lx"].shape[0]
    _state.update(np=np, orc=orc, cc=cc, p=p, Lx=code["Lx"], Lz=code["Lz"], max_iter=max_iter,
                  gz=orc.SideGraph(M["HdecZ"], M["HZ_full"][m:m + k], orc.llr_priors(M["channel_probsZ"])),
                  gx=orc.SideGraph(M["HdecX"], M["HX_full"][m:m + k], orc.llr_priors(M["channel_probsX"])))
    orc.lib()
    return True


def _shot(i):
    s = _state
    np, orc = s["np"], s["orc"]
    np.random.seed(1234 + i)
    sz, tz, sx, tx = orc.run_trial_fast(s["cc"], s["p"], s["Lx"], s["Lz"])
    ez, cz, iz = orc.decode_side(s["gz"], sz, tz, s["max_iter"])
    ex, cx, ix = orc.decode_side(s["gx"], sx, tx, s["max_iter"])
    return (ez, ex, iz * len(s["gz"].indices) + ix * len(s["gx"].indices))


def _ready(_):
    return os.getpid()


class CpuBaseline:
    def __init__(self, tag_name, p, max_iter, cores=None):
        import multiprocessing as mp
        self.cores = cores or len(os.sched_getaffinity(0))
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        self.pool = mp.get_context("spawn").Pool(self.cores, initializer=_init, initargs=(root, tag_name, p, max_iter))
        self.pool.map(_ready, range(self.cores * 2))      # wait until every worker built its tables
        self.next_shot = 0

    def run(self, n_shots):
        """Decode n_shots; returns (seconds, logical errors, edge messages)."""
        idx = range(self.next_shot, self.next_shot + n_shots)
        self.next_shot += n_shots
        t = time.perf_counter()
        res = self.pool.map(_shot, idx, chunksize=1)
        dt = time.perf_counter() - t
        return dt, sum(1 for r in res if r[0] or r[1]), sum(r[2] for r in res)

    def close(self):
        self.pool.terminate()
