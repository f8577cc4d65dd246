"""The literal BASELINE.json configuration 3 through the public entry point: run_simulation on the [[144,12,12]] gross code,
circuit-level p = 0.005, min-sum 20 iterations (dynamical alpha) + OSD-0, N shots over the ranks of a torchrun launch.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29555 \
        tools/run_baseline_config.py 100000000

Prints one JSON line on rank 0: wall-clock seconds of the whole call (table build, handle creation, shots, reduction)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
import qldpc_b200
from qldpc_b200.codes.bb_code import BB_CODES, make_bb_code
from qldpc_b200.simulation.engine import run_simulation

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
name, p = "[[144, 12, 12]]", 0.005
rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
if world > 1:
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
code = make_bb_code(name)
bb = {k: code[k] for k in ("ell", "m", "a_x_powers", "a_y_powers", "b_y_powers", "b_x_powers")}
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
res = run_simulation(code["Hx"], code["Hz"], code["Lx"], code["Lz"], p, num_trials=n, num_cycles=BB_CODES[name]["distance"], maxIter=20,
                     osd_order=0, alpha_mode="dynamical", base_seed=1234, progress=False, **bb)
if world > 1:
    dist.barrier()
dt = time.perf_counter() - t0
if rank == 0:
    print(json.dumps({"config": "BASELINE configs[2]: [[144,12,12]] p=0.005, min-sum 20 it + OSD-0", "entry": "run_simulation", "shots": res["num_trials"],
                      "n_gpus": world, "seconds": dt, "shots_per_s": res["num_trials"] / dt, "logical_error_rate": res["logical_error_rate"],
                      "z_ler": res["z_logical_error_rate"], "x_ler": res["x_logical_error_rate"], "logical_errors": res["logical_errors"]}), flush=True)
if world > 1:
    dist.destroy_process_group()
