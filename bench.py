#!/usr/bin/env python
"""bench.py -- decoded shots/s of the Monte-Carlo decoding hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[2]): [[144,12,12]] gross code, circuit-level p = 0.005, min-sum
20 iterations (dynamical alpha) + OSD-0 on non-converged sides, both sides per shot.
A step = one pass of the whole hot path (Philox sampling -> syndromes -> min-sum Z,X -> OSD-0 ->
logical check) over --shots-per-step shots per GPU (one device batch of --batch shots by default).  One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CODE = "[[144, 12, 12]]"
P = 0.005
MAX_ITER = 20
METRIC = "decoded shots/s ([[144,12,12]], p=0.005)"
WORKLOAD = ("[[144,12,12]] gross code, circuit-level p=0.005, min-sum 20 it (dynamical alpha) + OSD-0 on "
            "non-converged sides, Z and X side per shot")
# identical in both arms (the driver compares the two config dicts)
CONFIG = {"workload": WORKLOAD, "code": CODE, "p": P, "max_iter": MAX_ITER, "alpha_mode": "dynamical", "osd_order": 0,
          "shots": "both sides of every shot decoded; LER counts z_err or x_err (engine.py:122)"}


def smi_sampler(stop, out, gpu_index):
    q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    try:
        pr = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(gpu_index)],
                              stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
    except Exception:
        return
    try:
        while not stop.is_set():
            line = pr.stdout.readline()
            if not line:
                break
            out.append(line.strip())
    finally:
        pr.kill()


def summarize_clocks(lines):
    sm, mx, reasons = [], [], set()
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    for ln in lines:
        f = [x.strip() for x in ln.split(",")]
        try:
            sm.append(float(f[0])); mx.append(float(f[1]))
        except Exception:
            continue
        for nm, v in zip(names, f[3:7]):
            if v.lower().startswith("active"):
                reasons.add(nm)
    if not sm:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
    sm.sort()
    return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def ncu_traffic(kernel):
    """Per-launch ncu counters of one kernel from profiles/r2b_traffic.json (written by tools/ncu_traffic.py from an
    `ncu --set full` capture of this file's own command line; the JSON records the command) -> (dict, shots per launch)."""
    try:
        with open(os.path.join(ROOT, "profiles", "r2b_traffic.json")) as f:
            d = json.load(f)
        for name, k in d["kernels"].items():
            if name.startswith(kernel):
                return k, int(d["shots_per_launch"])
    except Exception:
        pass
    return None, 0


def reference_cpu_timing():
    """The real (numba) reference timed on host cores.  /root/reference does not travel to the GPU box, so the record is
    the measurement made in the build container (profiles/r2_reference_cpu_timing.json says how)."""
    try:
        with open(os.path.join(ROOT, "profiles", "r2_reference_cpu_timing.json")) as f:
            d = json.load(f)
        return {"value": d["shots_per_s"], "unit": "shots/s", "cores": d["workers"], "kind": "reference",
                "sample": f"{d['shots']} shots in {d['seconds']:.0f} s; {d['what']}", "where": d["where"],
                "logical_error_rate": d["logical_error_rate"]}
    except Exception:
        return None


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def run_reference(args, rank, world):
    """--impl reference: the CPU port of the reference's per-shot loop on all host cores."""
    if rank != 0:
        return
    from oracle.cpu_baseline import CpuBaseline
    cores = len(os.sched_getaffinity(0))
    base = CpuBaseline(CODE, P, MAX_ITER, cores)
    per_step = max(cores, min(cores * 8, 512))
    for _ in range(args.warmup):
        base.run(max(cores, per_step // 4))
    tot_t = tot_n = tot_err = tot_em = 0
    for _ in range(args.steps):
        dt, errs, em = base.run(per_step)
        tot_t += dt; tot_n += per_step; tot_err += errs; tot_em += em
    base.close()
    value = tot_n / tot_t
    sample = f"{tot_n} shots ({per_step}/step) of the same workload, np.random.seed(1234+i) per shot, spawn pool of {cores} workers"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "shots/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": CONFIG, "shots_per_step": per_step,
        "cpu_baseline": {"value": value, "unit": "shots/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "shots/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "logical_error_rate": tot_err / tot_n, "edge_messages_per_s": tot_em / tot_t,
        "note": "C port (oracle/) of the reference's numba per-shot loop; the reference itself (pure Python + numba) measured "
                "~0.95 shots/s on 8 cores for this workload (BASELINE.md section 2)",
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--shots-per-step", type=int, default=65536)
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-run-simulation", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    # stdout carries exactly one JSON line: NCCL's own log lines (NCCL_DEBUG=VERSION/INFO on some boxes) go to stderr
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus > 1 and world == 1 and "RANK" not in os.environ:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(29400 + os.getpid() % 500), os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    import qldpc_b200  # noqa: F401
    from qldpc_b200 import _lib
    from qldpc_b200.codes.bb_code import BB_CODES, BBCodeCircuit, make_bb_code
    from qldpc_b200.noise.builder import fault_tables_for, matrices_from_tables
    from qldpc_b200.noise.compiled import CompiledCircuit
    from qldpc_b200.simulation.engine import ShotEngine

    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    code = make_bb_code(CODE)
    bb = {k: code[k] for k in ("ell", "m", "a_x_powers", "a_y_powers", "b_y_powers", "b_x_powers")}
    d = BB_CODES[CODE]["distance"]
    cc = CompiledCircuit.from_builder(BBCodeCircuit(code["Hx"], code["Hz"], num_cycles=d, **bb))
    ft = fault_tables_for(cc, code["Lx"], code["Lz"])
    M = matrices_from_tables(ft, P, d)
    eng = ShotEngine(cc, code["Lx"], code["Lz"], M, max_batch=args.batch, device=local_rank)
    cfg = _lib.make_config(MAX_ITER, _lib.QB_ALPHA_DYNAMIC)
    stream = torch.cuda.current_stream(dev)
    eng.pipeline.set_stream(stream.cuda_stream)
    B = args.shots_per_step
    seed = 1234

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # K steps = K device batches of B shots in ONE library call: the two batch workspaces of the pipeline overlap the
    # sampling / min-sum of batch i+1 with the OSD-0 tail of batch i, and nothing synchronises with the host in between
    def steps(first_step, k):
        first = (first_step * world + rank * k) * B          # disjoint shot ranges per call and rank
        counts, _ = eng.pipeline.run(seed, first, k * B, P, cfg)
        return counts, eng.pipeline.stats()

    if args.warmup:
        steps(0, args.warmup)
    clock_lines, stop = [], threading.Event()
    th = threading.Thread(target=smi_sampler, args=(stop, clock_lines, local_rank), daemon=True)
    th.start()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall = time.perf_counter()
    ev0.record(stream)
    tot, st = steps(args.warmup, args.steps)
    ev1.record(stream)
    barrier()
    t_wall = time.perf_counter() - t_wall
    stop.set()
    launches_timed, em_timed = int(st["kernel_launches"]), int(st["edge_messages"])
    # per-stage and per-kernel durations come from two extra single-batch calls after the timed region: inside the
    # pipelined call the stages of consecutive batches overlap, so their event-to-event times contain each other
    agg = dict(ms_sample=0.0, ms_minsum=0.0, ms_osd=0.0, ms_logical=0.0, kernel_launches=0, edge_messages=0, osd_sides=0)
    n_stage = 2
    for j in range(n_stage):
        _, st1 = steps(args.warmup + args.steps + j, 1)
        for k in agg:
            agg[k] += st1[k]
    agg["batches"] = n_stage * -(-B // args.batch)
    ms = ev0.elapsed_time(ev1)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    ctot = torch.from_numpy(tot.copy()).to(dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(ctot, op=dist.ReduceOp.SUM)          # the path's only collective: counter reduction
    ms_max = float(t.item())
    ctot = ctot.cpu().numpy()
    th.join(timeout=2)

    # ---- e2e: reference-facing C-ABI call with HOST buffers (pinned), H2D + D2H inside the timed region ----
    Be = min(args.batch, B)
    rng = np.random.default_rng(99 + rank)
    n_sets = 3
    sets = []
    kind = ft.loc_kind
    for _ in range(n_sets):
        # i.i.d. Bernoulli(P) over the Be x L (shot, location) grid by geometric skipping (exact, and cheap on the host)
        total = Be * ft.L
        gaps = rng.geometric(P, size=int(total * P * 1.05) + 1000)
        pos = np.cumsum(gaps) - 1
        while pos[-1] < total:
            more = np.cumsum(rng.geometric(P, size=len(gaps) // 10 + 1000)) + pos[-1]
            pos = np.concatenate([pos, more])
        pos = pos[pos < total]
        sh, loc = np.divmod(pos, ft.L)
        out = np.where(kind[loc] == 3, rng.integers(0, 15, len(loc)), np.where(kind[loc] == 2, rng.integers(0, 3, len(loc)), 0))
        evp = np.zeros(Be + 1, dtype=np.int64); np.add.at(evp, sh + 1, 1); evp = np.cumsum(evp).astype(np.int32)
        ev = (loc.astype(np.uint32) | (out.astype(np.uint32) << 24)).astype(np.uint32)
        pp = torch.empty(len(evp), dtype=torch.int32).pin_memory(); pp.numpy()[:] = evp
        pe = torch.empty(len(ev), dtype=torch.int32).pin_memory(); pe.numpy().view(np.uint32)[:] = ev
        sets.append((pp, pe))
    e2e_steps = max(3, args.steps)
    # the e2e call: ONE qb_pipeline_run_events_host over e2e_steps batches of host-sampled fault events (the three sets
    # tiled, offsets rebased), like the device-resident measurement one library call for all steps
    ptr_parts, ev_parts, base = [np.zeros(1, dtype=np.int64)], [], 0
    for i in range(e2e_steps):
        pp, pe = sets[i % n_sets]
        q = pp.numpy().astype(np.int64)
        ptr_parts.append(q[1:] + base); base += int(q[-1]); ev_parts.append(pe.numpy())
    big_ptr = torch.empty(e2e_steps * Be + 1, dtype=torch.int32).pin_memory(); big_ptr.numpy()[:] = np.concatenate(ptr_parts)
    big_ev = torch.empty(base, dtype=torch.int32).pin_memory(); big_ev.numpy()[:] = np.concatenate(ev_parts)
    pflags = torch.empty(e2e_steps * Be, dtype=torch.uint8).pin_memory()
    eng.pipeline.set_stream(None)
    # one untimed call of the same size first (the library sizes its device-side event buffer on the first call of a
    # size), then three timed calls, each bracketed by barrier + synchronize; the median is reported (a single wall-clock
    # sample of a ~0.5 s host call was seen to vary 2x with host noise), all three are listed in e2e.samples
    eng.pipeline.run_events(big_ptr.numpy(), big_ev.numpy().view(np.uint32), cfg, flags_out=pflags.numpy())
    h2d = big_ptr.numel() * 4 + big_ev.numel() * 4
    d2h = e2e_steps * Be + 64
    e2e_samples = []
    for _ in range(3):
        barrier()
        t0 = time.perf_counter()
        eng.pipeline.run_events(big_ptr.numpy(), big_ev.numpy().view(np.uint32), cfg, flags_out=pflags.numpy())
        barrier()
        te_t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(te_t, op=dist.ReduceOp.MAX)
        e2e_samples.append(world * Be * e2e_steps / float(te_t.item()))
    e2e_value = sorted(e2e_samples)[1]

    # ---- e2e through the public entry point a user of the reference calls: run_simulation(num_trials=2**20, ...) ----
    # wall clock of the whole call on this rank's GPU: host table build (fault signatures, priors), handle creation,
    # 2**20 shots (Philox), counters back.  N = 1 only (under torchrun the entry point shards shots over the ranks itself).
    e2e_rs = None
    if world == 1 and not args.no_run_simulation:
        from qldpc_b200.simulation.engine import run_simulation
        n_rs = 1 << 20
        t0 = time.perf_counter()
        res = run_simulation(code["Hx"], code["Hz"], code["Lx"], code["Lz"], P, num_trials=n_rs, num_cycles=d, maxIter=MAX_ITER,
                             osd_order=0, alpha_mode="dynamical", base_seed=seed, progress=False, **bb)
        dt_all = time.perf_counter() - t0
        t0 = time.perf_counter()
        res2 = run_simulation(code["Hx"], code["Hz"], code["Lx"], code["Lz"], P, num_trials=n_rs, num_cycles=d, maxIter=MAX_ITER,
                              osd_order=0, alpha_mode="dynamical", base_seed=seed, precomputed_matrices=M, progress=False, **bb)
        dt_pre = time.perf_counter() - t0
        e2e_rs = {"value": n_rs / dt_all, "unit": "shots/s", "shots": n_rs, "seconds": dt_all,
                  "with_precomputed_matrices": {"value": n_rs / dt_pre, "seconds": dt_pre,
                                                "note": "second call in the process: decoding matrices passed in, circuit tables from the in-process cache"},
                  "logical_error_rate": res["logical_error_rate"], "same_result_both_calls": res == res2,
                  "path": "qldpc_b200.simulation.engine.run_simulation(num_trials=2**20, maxIter=20, alpha_mode='dynamical', osd_order=0): "
                          "wall clock of the whole call incl. host-side table build and handle creation"}

    if rank == 0:
        clocks = summarize_clocks(clock_lines)
        hbm_peak, peak_src = measured_peaks()
        total_shots = world * B * args.steps
        value = total_shots / (ms_max * 1e-3)
        # dominant kernel: minsum_edge_kernel (two launches per batch: Z and X side)
        n_ms_launch = 2 * agg["batches"]
        ms_launch = agg["ms_minsum"] / max(1, n_ms_launch)
        shots_per_launch = min(args.batch, B)
        gz, gx = eng.decZ, eng.decX
        nonconv_frac = (ctot[4] + ctot[5]) / max(1, 2 * ctot[3])
        # algorithmic units of one launch: edge-messages = nnz x iterations executed, 8 B each on chip;
        # HBM: syndrome words in, hard bits + flags out, posteriors out only for non-converged sides
        em_per_s = agg["edge_messages"] / (agg["ms_minsum"] * 1e-3) if agg["ms_minsum"] > 0 else 0.0
        em_per_launch = agg["edge_messages"] / max(1, n_ms_launch)
        per_shot = 0.5 * ((gz.m + 31) // 32 * 4 + (gx.m + 31) // 32 * 4) + 0.5 * ((gz.n + 31) // 32 * 4 + (gx.n + 31) // 32 * 4) + 5 \
            + nonconv_frac * 0.5 * (gz.n + gx.n) * 4
        hbm_bytes_launch = per_shot * shots_per_launch
        hbm_achieved = hbm_bytes_launch / (ms_launch * 1e-3) / 1e9 if ms_launch > 0 else 0.0
        sm_mhz = clocks["sm_mhz"] or 1965.0
        smem_peak = 148 * 128 * sm_mhz * 1e6 / 1e9          # GB/s: 128 B/clk/SM
        ncu, ncu_shots = ncu_traffic("minsum_edge_kernel")
        scale = shots_per_launch / ncu_shots if ncu_shots else 0.0
        out = {
            "metric": METRIC, "value": value, "unit": "shots/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": CONFIG,
            "run": {"shots_per_step_per_gpu": B, "batch": args.batch, "steps_per_library_call": args.steps,
                    "sampler": "Philox4x32-10 on device, counter = global shot index",
                    "l2": "no flush needed: per-batch posterior/state buffers (>1 GB) exceed the 126 MB L2 and every step decodes new shots"},
            # the binding roofline of the dominant kernel is on chip: algorithmic 8 B per edge-message against
            # 148 SMs x 128 B/clk of shared-memory bandwidth at the SM clock sampled during the run
            "roofline": {"bound": "smem", "kernel": "minsum_edge_kernel<1024,1>", "achieved": em_per_s * 8 / 1e9, "peak": smem_peak,
                         "unit": "GB/s", "frac": em_per_s * 8 / 1e9 / smem_peak,
                         "traffic": (ncu["smem_bytes"] * scale) if ncu else None,
                         "traffic_unit": "bytes/launch: ncu l1tex__data_pipe_lsu_wavefronts_mem_shared.sum x 128 B of a 16384-shot launch "
                                         "(profiles/r2b_traffic.json), scaled to this launch size",
                         "algorithmic_bytes_per_launch": em_per_launch * 8, "edge_messages_per_launch": em_per_launch,
                         "edge_messages_per_s": em_per_s, "bytes_per_edge_message": 8, "ms_per_launch": ms_launch,
                         "shots_per_launch": shots_per_launch,
                         "peak_source": "148 SMs x 128 B/clk x median SM clock under load (MEASURED_PEAKS.json holds HBM and tensor peaks only)",
                         "issue_active_pct_ncu": ncu["issue_active_pct"] if ncu else None,
                         "note": "the kernel moves 16 B + 2 B of slot index per edge-message through shared memory and is bound by the alu pipe "
                                 "(check rows) and by shared-memory instruction issue (variables), DESIGN.md section 4"},
            "roofline_hbm": {"bound": "hbm", "kernel": "minsum_edge_kernel<1024,1>", "achieved": hbm_achieved, "peak": hbm_peak, "unit": "GB/s",
                             "frac": hbm_achieved / hbm_peak, "traffic": (ncu["dram_bytes"] * scale) if ncu else None,
                             "traffic_unit": "bytes/launch: ncu dram__bytes_read.sum + dram__bytes_write.sum (profiles/r2b_traffic.json), scaled",
                             "algorithmic_bytes_per_launch": hbm_bytes_launch, "peak_source": peak_src,
                             "note": "non-binding by design: messages never leave the SM; traffic = algorithmic bytes (posteriors of non-converged sides), no re-reads"},
            "e2e": {"value": e2e_value, "unit": "shots/s", "h2d_bytes_per_step": h2d // e2e_steps, "d2h_bytes_per_step": d2h // e2e_steps,
                    "shots_per_step": Be, "steps": e2e_steps, "samples": e2e_samples, "statistic": "median of 3 timed calls after one untimed call of the same size",
                    "path": "one qb_pipeline_run_events_host call (the C-ABI entry behind run_trial_fast + the decoders) over all steps: fault "
                            "events sampled on the host BEFORE the timed region (three sets, tiled, pinned) -> one H2D -> per batch K2 syndromes -> "
                            "min-sum -> OSD-0 -> logical check -> flags D2H (pinned); contains no sampling work; wall clock incl. the final host sync"},
            "e2e_run_simulation": e2e_rs,
            "gpu_launches": launches_timed,
            "clocks": clocks,
            "stage_ms_per_step": {**{k: agg[k] / n_stage for k in ("ms_sample", "ms_minsum", "ms_osd", "ms_logical")},
                                  "note": "two un-pipelined single-batch steps after the timed region (stages of consecutive batches overlap inside it)"},
            "wall_ms_per_step": 1e3 * t_wall / args.steps,
            "logical_error_rate": float(ctot[2] / max(1, ctot[3])), "z_ler": float(ctot[0] / max(1, ctot[3])), "x_ler": float(ctot[1] / max(1, ctot[3])),
            "nonconverged_side_fraction": float(nonconv_frac),
            "edge_messages_per_s_whole_job": world * em_timed / (ms_max * 1e-3),
            "osd_sides_per_s": (agg["osd_sides"] / (agg["ms_osd"] * 1e-3)) if agg["ms_osd"] > 0 else None,
            "cpu_baseline_reference": reference_cpu_timing(),
        }
        if world == 1 and not args.no_cpu_baseline:
            from oracle.cpu_baseline import CpuBaseline
            cores = len(os.sched_getaffinity(0))
            base = CpuBaseline(CODE, P, MAX_ITER, cores)
            base.run(cores)                                   # warm the workers
            n = cores * 4
            dt, errs, em = base.run(n)
            shots, secs, errs_t = n, dt, errs
            while secs < args.cpu_seconds:
                n2 = max(cores, int(cores * min(64, max(1, (args.cpu_seconds - secs) / max(dt / 4, 1e-3) / 4))))
                dt2, e2, em2 = base.run(n2)
                shots += n2; secs += dt2; errs_t += e2; em += em2
            base.close()
            out["cpu_baseline"] = {"value": shots / secs, "unit": "shots/s", "cores": cores, "kind": "port",
                                   "sample": f"{shots} shots of the same workload in {secs:.1f} s (oracle C port, spawn pool, np.random.seed(1234+i) per shot)",
                                   "logical_error_rate": errs_t / shots, "edge_messages_per_s": em / secs}
        print(json.dumps(out), flush=True)
    eng.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
