"""``run_simulation`` on the GPU (reference ``src/simulation/engine.py:193-488``).

Same keyword arguments and result dict.  Differences that do not change semantics:
  * shots run in device batches through the fused pipeline (sample -> syndromes -> min-sum ->
    OSD-0 -> logical check) instead of a process pool; ``num_workers`` / ``use_jit`` are accepted
    and ignored;
  * the per-shot generator is Philox4x32-10 keyed by ``base_seed`` with the shot index as counter
    (reference: MT19937 reseeded per shot, engine.py:70), so results are reproducible and do not
    depend on the batch size or the number of GPUs; LERs agree statistically;
  * early stop (``target_logical_errors``) is replayed on the host over per-shot flags in shot
    order, which gives the reference's exact cut (engine.py:450-464);
  * with ``torch.distributed`` initialised every rank takes a contiguous slice of each round of
    shots and the counters are all-reduced (NCCL on GPUs, gloo in the CPU tests).
"""
import logging

import numpy as np

from .. import _lib
from ..codes.bb_code import BBCodeCircuit
from ..noise.builder import fault_tables_for, matrices_from_tables
from ..noise.compiled import CompiledCircuit

_logger = logging.getLogger(__name__)


def llr_priors(channel_probs):
    """engine.py:210-212."""
    with np.errstate(divide="ignore", invalid="ignore"):
        cp = np.asarray(channel_probs, dtype=np.float64)
        return np.clip(np.nan_to_num(np.log((1 - cp) / cp)), -50, 50)


def _dense_to_csr(H):
    mask = np.asarray(H) != 0
    indptr = np.concatenate([[0], np.cumsum(mask.sum(axis=1))]).astype(np.int32)
    return indptr, np.nonzero(mask)[1].astype(np.int32)


def shard_range(total, rank, world):
    """Contiguous slice [lo, hi) of ``total`` items owned by ``rank`` (sizes differ by at most 1)."""
    base, rem = divmod(int(total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def early_stop_cut(flags, target, errors_before=0):
    """Index (exclusive) at which the reference loop stops: first shot where the cumulative number
    of logical errors reaches ``target`` (engine.py:450-464); None if not reached in ``flags``."""
    cum = errors_before + np.cumsum(np.asarray(flags) != 0)
    hit = np.nonzero(cum >= target)[0]
    return None if hit.size == 0 else int(hit[0]) + 1


def _dist():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist, dist.get_rank(), dist.get_world_size()
    except Exception:
        pass
    return None, 0, 1


class _Progress:
    """Batch-level stand-in for the reference's per-trial rich progress bar (engine.py:436-460): the same line
    ``p=... | logical=<errors>/<target>``, advanced once per device round instead of once per shot."""

    def __init__(self, error_rate, total, target, enabled=True):
        self.bar = self.task = None
        if not enabled:
            return
        try:
            from rich.progress import BarColumn, Progress, TaskProgressColumn, TextColumn, TimeElapsedColumn
            self.bar = Progress(TextColumn("p={task.fields[p]:.4g} | logical={task.fields[errors]}/{task.fields[target]}", justify="left"),
                                BarColumn(), TaskProgressColumn(), TimeElapsedColumn())
            self.bar.start()
            self.task = self.bar.add_task("simulate", total=total, p=error_rate, errors=0, target=(target if target else "∞"))
        except Exception:          # rich missing or no usable console: log lines instead
            self.bar = None
            self.p, self.total, self.target = error_rate, total, target

    def update(self, trials, errors, estimate=False):
        if self.bar is not None:
            self.bar.update(self.task, completed=trials, errors=(f"~{errors}" if estimate else errors))
        elif hasattr(self, "p"):
            _logger.info("p=%.4g | logical=%s/%s | %d/%d shots", self.p, errors, self.target or "∞", trials, self.total)

    def close(self):
        if self.bar is not None:
            self.bar.stop()


class ShotEngine:
    """Device state of one (code, p): sampler, two decoders, pipeline."""

    def __init__(self, compiled, Lx, Lz, matrices, max_batch=65536, device=None):
        ft = fault_tables_for(compiled, Lx, Lz)
        self.ft = ft
        k = np.asarray(Lx).shape[0]
        mz, mx = int(matrices["first_logical_rowZ"]), int(matrices["first_logical_rowX"])
        self.llrs_z = llr_priors(matrices["channel_probsZ"])
        self.llrs_x = llr_priors(matrices["channel_probsX"])
        HZf, HXf = np.asarray(matrices["HZ_full"]), np.asarray(matrices["HX_full"])
        zp, zi = _dense_to_csr(matrices["HdecZ"])
        xp, xi = _dense_to_csr(matrices["HdecX"])
        self.nnz = (len(zi), len(xi))
        self.sampler = _lib.Sampler(ft, device)
        self.decZ = _lib.Decoder(zp, zi, np.asarray(matrices["HdecZ"]).shape[1], self.llrs_z, HZf[mz:mz + k], device)
        self.decX = _lib.Decoder(xp, xi, np.asarray(matrices["HdecX"]).shape[1], self.llrs_x, HXf[mx:mx + k], device)
        self.pipeline = _lib.Pipeline(self.sampler, self.decZ, self.decX, max_batch)
        self.max_batch = max_batch

    def close(self):
        self.pipeline.close(); self.decZ.close(); self.decX.close(); self.sampler.close()


def _estimation_trials(n_cols, error_rate, alpha_estimation_trials):
    """Dynamic trial count of engine.py:233-247: enough samples of true-1 bits for the histogram."""
    dynamic = max(500, min(50000, int(2000 / (n_cols * error_rate))))
    return alpha_estimation_trials if alpha_estimation_trials != 5000 else dynamic


def _alpha_setup(alpha_mode, use_dynamic_alpha, alvarado_alpha, matrices, llrs_z, llrs_x, error_rate, maxIter,
                 alpha_estimation_trials, alpha_estimation_bins, plot_dir):
    """alpha-mode handling of engine.py:216-344 -> (mode name, QB mode, alpha_z, alpha_x, extra result fields)."""
    extras = {}
    if alpha_mode is None:
        alpha_mode = "dynamical" if use_dynamic_alpha else "alvarado"
    fmt = lambda r: f"{r:.6g}".replace(".", "p")
    if alpha_mode == "dynamical":
        return alpha_mode, _lib.QB_ALPHA_DYNAMIC, 1.0, 1.0, extras
    if alpha_mode == "alvarado":
        if alvarado_alpha is None:
            from ..decoding.alpha import estimate_alpha_alvarado
            tz = _estimation_trials(np.asarray(matrices["HdecZ"]).shape[1], error_rate, alpha_estimation_trials)
            tx = _estimation_trials(np.asarray(matrices["HdecX"]).shape[1], error_rate, alpha_estimation_trials)
            _logger.info("Alpha estimation trials: Z=%d, X=%d (dynamic based on n*p)", tz, tx)
            az, r2z = estimate_alpha_alvarado(matrices["HdecZ"], error_rate, trials=tz, bins=alpha_estimation_bins, plot_dir=plot_dir,
                                              plot_prefix=f"alvarado_{fmt(error_rate)}_z", llrs=llrs_z)
            ax, r2x = estimate_alpha_alvarado(matrices["HdecX"], error_rate, trials=tx, bins=alpha_estimation_bins, plot_dir=plot_dir,
                                              plot_prefix=f"alvarado_{fmt(error_rate)}_x", llrs=llrs_x)
            extras.update(alpha_r2_z=r2z, alpha_r2_x=r2x)
        elif isinstance(alvarado_alpha, (list, tuple, np.ndarray)) and len(alvarado_alpha) == 2:
            az, ax = float(alvarado_alpha[0]), float(alvarado_alpha[1])
            extras.update(alpha_r2_z=None, alpha_r2_x=None)
        else:
            az = ax = float(alvarado_alpha)
            extras.update(alpha_r2_z=None, alpha_r2_x=None)
        _logger.info("Alvarado alpha for p=%.6g: alpha_z=%.6g, alpha_x=%.6g", error_rate, az, ax)
        if az <= 0 or ax <= 0:
            raise ValueError("alpha must be > 0 when alpha_mode='alvarado'")
        return alpha_mode, _lib.QB_ALPHA_FIXED, az, ax, extras
    if alpha_mode == "alvarado-autoregressive":
        if alvarado_alpha is not None:
            raise ValueError("alvarado_alpha must be None for alvarado-autoregressive")
        from ..decoding.alpha import estimate_alpha_alvarado_autoregressive
        tz = _estimation_trials(np.asarray(matrices["HdecZ"]).shape[1], error_rate, alpha_estimation_trials)
        tx = _estimation_trials(np.asarray(matrices["HdecX"]).shape[1], error_rate, alpha_estimation_trials)
        _logger.info("Autoregressive alpha estimation trials: Z=%d, X=%d (dynamic based on n*p)", tz, tx)
        az, r2z = estimate_alpha_alvarado_autoregressive(matrices["HdecZ"], error_rate, maxIter=maxIter, trials=tz, bins=alpha_estimation_bins,
                                                         plot_dir=plot_dir, plot_prefix=f"autoregressive_{fmt(error_rate)}_z", llrs=llrs_z)
        ax, r2x = estimate_alpha_alvarado_autoregressive(matrices["HdecX"], error_rate, maxIter=maxIter, trials=tx, bins=alpha_estimation_bins,
                                                         plot_dir=plot_dir, plot_prefix=f"autoregressive_{fmt(error_rate)}_x", llrs=llrs_x)
        extras.update(alpha_values_z=az, alpha_values_x=ax, alpha_r2_values_z=r2z, alpha_r2_values_x=r2x)
        return alpha_mode, _lib.QB_ALPHA_SEQUENCE, az, ax, extras
    raise ValueError(f"Unsupported alpha_mode: {alpha_mode}")


_TABLE_CACHE = {}


def _circuit_tables(Hx, Hz, Lx, Lz, num_cycles, bb_params):
    """Compiled circuit + fault signature tables of a code.  They depend on the code and the number of cycles only, not on
    the error rate, so a sweep over error rates (the reference's main.py:95-141 loop) builds them once per process
    (0.85 s for the gross code); the last four codes are kept."""
    import hashlib
    h = hashlib.sha1()
    for a in (Hx, Hz, Lx, Lz):
        a = np.ascontiguousarray(a)
        h.update(str(a.shape).encode()); h.update(str(a.dtype).encode()); h.update(a.tobytes())
    h.update(repr((int(num_cycles), sorted((k, np.asarray(v).tolist()) for k, v in bb_params.items()))).encode())
    key = h.hexdigest()
    hit = _TABLE_CACHE.pop(key, None)
    if hit is None:
        compiled = CompiledCircuit.from_builder(BBCodeCircuit(Hx, Hz, num_cycles=num_cycles, **bb_params))
        hit = (compiled, fault_tables_for(compiled, Lx, Lz))
    _TABLE_CACHE[key] = hit                                  # (re-inserted last: the dict is the LRU order)
    while len(_TABLE_CACHE) > 4:
        _TABLE_CACHE.pop(next(iter(_TABLE_CACHE)))
    return hit


def run_simulation(Hx, Hz, Lx, Lz, error_rate, num_trials=1000, num_cycles=12, maxIter=50, osd_order=0,
                   use_dynamic_alpha=True, alpha_mode=None, alvarado_alpha=None, alpha_estimation_trials=5000,
                   alpha_estimation_bins=50, precomputed_matrices=None, num_workers=None, base_seed=None,
                   use_jit=True, target_logical_errors=None, max_trials=None, scopt=False,
                   estimation_plot_dir=None, batch_size=None, progress=True, **bb_params):
    if base_seed is None:
        base_seed = np.random.randint(0, 2 ** 31)
    if alpha_mode not in (None, "dynamical", "alvarado", "alvarado-autoregressive"):
        raise ValueError(f"Unsupported alpha_mode: {alpha_mode}")
    compiled, ft = _circuit_tables(Hx, Hz, Lx, Lz, num_cycles, bb_params)
    matrices = precomputed_matrices or matrices_from_tables(ft, error_rate, num_cycles)
    if estimation_plot_dir is not None:
        import os
        os.makedirs(estimation_plot_dir, exist_ok=True)
    alpha_mode, qmode, alpha_z, alpha_x, extras = _alpha_setup(
        alpha_mode, use_dynamic_alpha, alvarado_alpha, matrices, llr_priors(matrices["channel_probsZ"]),
        llr_priors(matrices["channel_probsX"]), error_rate, maxIter, alpha_estimation_trials, alpha_estimation_bins,
        estimation_plot_dir)
    if scopt:
        # SCOPT beta pre-pass (engine.py:346-389; like the reference the estimate is reported, not yet used by the decoder)
        from ..decoding.scopt import estimate_scopt_beta
        fmt = lambda r: f"{r:.6g}".replace(".", "p")
        betas = {}
        for sd, key, ll in (("z", "HdecZ", llr_priors(matrices["channel_probsZ"])), ("x", "HdecX", llr_priors(matrices["channel_probsX"]))):
            n_sd = np.asarray(matrices[key]).shape[1]
            trials_sd = max(500, min(50000, int(2000 / (n_sd * error_rate))))
            alpha_arg = alpha_z if sd == "z" else alpha_x
            betas[sd] = estimate_scopt_beta(matrices[key], error_rate, trials=trials_sd, bins=alpha_estimation_bins, alpha=alpha_arg,
                                            alpha_mode=alpha_mode, maxIter=maxIter, plot_dir=estimation_plot_dir,
                                            plot_prefix=f"scopt_{fmt(error_rate)}_{sd}", llrs=ll)
        _logger.info("SCOPT beta (estimated) for p=%.6g: beta_z=%.6g, beta_x=%.6g", error_rate, betas["z"][0], betas["x"][0])
        extras.update(beta_z=betas["z"][0], beta_x=betas["x"][0], beta_r2_z=betas["z"][1], beta_r2_x=betas["x"][1])
    if max_trials is None:
        max_trials = num_trials if num_trials is not None else 1000000
    stop_on_errors = target_logical_errors is not None and target_logical_errors > 0

    dist, rank, world = _dist()
    if dist is not None and _dist_backend(dist) == "nccl":
        import torch
        torch.cuda.set_device(_lib.default_device())
    base_seed = _broadcast_seed(dist, base_seed)
    if batch_size is None:
        batch_size = int(min(65536, max(256, -(-max_trials // world))))
    eng = ShotEngine(compiled, Lx, Lz, matrices, max_batch=batch_size)
    cfg = _lib.make_config(maxIter, qmode, alpha_z, alpha_x, clip_llr=20.0, use_osd=True)
    report = _Progress(error_rate, max_trials, target_logical_errors if stop_on_errors else None, enabled=progress and rank == 0)
    try:
        z_errs = x_errs = tot_errs = trials_run = 0
        done = 0
        if stop_on_errors:
            # early stop: per round the per-shot flags of all ranks are gathered and replayed in shot order, which gives
            # the reference's exact cut (engine.py:450-464)
            while done < max_trials:
                round_total = min(world * batch_size, max_trials - done)
                lo, hi = shard_range(round_total, rank, world)
                counts, flags = eng.pipeline.run(base_seed, done + lo, hi - lo, error_rate, cfg, want_flags=True)
                all_flags = _gather_flags(dist, world, flags, round_total)
                cut = early_stop_cut(all_flags & 3, target_logical_errors, tot_errs)
                use = all_flags[:cut] if cut is not None else all_flags
                z_errs += int(np.sum(use & 1 != 0)); x_errs += int(np.sum(use & 2 != 0)); tot_errs += int(np.sum(use != 0))
                trials_run += len(use)
                report.update(trials_run, tot_errs)
                if cut is not None:
                    break
                done += round_total
        else:
            # fixed number of shots: no synchronisation between ranks inside the loop.  Every rank owns one contiguous
            # range of shot indices and walks it in chunks of a few batches (one library call each: the batches of a call
            # are pipelined on the device, the host only updates the progress line in between); the counters are reduced
            # once at the end.
            lo, hi = shard_range(max_trials, rank, world)
            local = np.zeros(8, dtype=np.int64)
            chunk = batch_size * max(1, min(32, (1 << 21) // batch_size))
            for start in range(lo, hi, chunk):
                n = min(chunk, hi - start)
                counts, _ = eng.pipeline.run(base_seed, start, n, error_rate, cfg)
                local += counts
                report.update(min(max_trials, int(local[3]) * world), int(local[2]) * world, estimate=world > 1)
            c = _reduce_counts(dist, local)
            z_errs, x_errs, tot_errs, trials_run = int(c[0]), int(c[1]), int(c[2]), int(c[3])
            report.update(trials_run, tot_errs)
    finally:
        report.close()
        eng.close()
    result = {
        "logical_error_rate": tot_errs / max(1, trials_run),
        "z_logical_error_rate": z_errs / max(1, trials_run),
        "x_logical_error_rate": x_errs / max(1, trials_run),
        "num_trials": trials_run,
        "logical_errors": tot_errs,
    }
    result.update(extras)          # alpha_values_* / alpha_r2_* like engine.py:474-481
    return result


def _dist_backend(dist):
    return dist.get_backend()


def _collective_device(dist):
    """Tensors for a collective live on this rank's own GPU under NCCL (LOCAL_RANK, like every handle of the
    package), on the host under gloo."""
    import torch
    if dist.get_backend() == "nccl":
        return torch.device("cuda", _lib.default_device())
    return torch.device("cpu")


def _broadcast_seed(dist, seed):
    """Every rank must key Philox with the same seed (rank 0's draw when the caller passed None)."""
    if dist is None:
        return seed
    import torch
    t = torch.tensor([int(seed)], dtype=torch.int64, device=_collective_device(dist))
    dist.broadcast(t, src=0)
    return int(t.item())


def _reduce_counts(dist, counts):
    if dist is None:
        return counts
    import torch
    dev = _collective_device(dist)
    t = torch.from_numpy(np.asarray(counts, dtype=np.int64).copy()).to(dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def _gather_flags(dist, world, flags, round_total):
    if dist is None:
        return flags
    import torch
    dev = _collective_device(dist)
    width = -(-round_total // world)
    buf = np.full(width, 255, dtype=np.uint8)       # 255 = padding marker
    buf[:len(flags)] = flags
    t = torch.from_numpy(buf).to(dev)
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    parts = []
    for r in range(world):
        lo, hi = shard_range(round_total, r, world)
        parts.append(out[r].cpu().numpy()[:hi - lo])
    return np.concatenate(parts)
