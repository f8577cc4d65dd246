"""Per-kernel traffic constants for bench.py from an `ncu --set full` report:

    python tools/ncu_traffic.py gpurun_out/r2_full.ncu-rep 16384 profiles/r2_traffic.json

For every kernel in the report (averaged over its captured launches): duration, DRAM bytes read + written
(dram__bytes_read.sum + dram__bytes_write.sum), shared-memory wavefronts and their bytes (128 B per wavefront),
warp instructions, issue-slot utilisation, achieved occupancy.  `shots_per_launch` is the batch size of the profiled
command; bench.py scales per shot.  The capture command is recorded in the JSON.
"""
import collections, csv, io, json, subprocess, sys

rep, shots, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "second": 1e3, "nsecond": 1e-6}


def val(r, name):
    x = float(r[col[name]].replace(",", ""))
    return x * UNIT.get(units[col[name]], 1.0)


acc = collections.defaultdict(lambda: collections.defaultdict(float))
cnt = collections.Counter()
for r in rows[2:]:
    name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").replace("qb::", "")
    cnt[name] += 1
    a = acc[name]
    a["ms_per_launch"] += val(r, "gpu__time_duration.sum")
    a["dram_bytes_read"] += val(r, "dram__bytes_read.sum")
    a["dram_bytes_write"] += val(r, "dram__bytes_write.sum")
    a["smem_wavefronts"] += val(r, "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum")
    a["warp_instructions"] += val(r, "smsp__inst_executed.sum")
    a["issue_active_pct"] += val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active")
    a["warps_active_pct"] += val(r, "sm__warps_active.avg.pct_of_peak_sustained_active")
    a["registers_per_thread"] += val(r, "launch__registers_per_thread")
kernels = {}
for name, a in acc.items():
    k = {key: v / cnt[name] for key, v in a.items()}
    k["launches_captured"] = cnt[name]
    k["dram_bytes"] = k["dram_bytes_read"] + k["dram_bytes_write"]
    k["smem_bytes"] = k["smem_wavefronts"] * 128.0
    kernels[name] = k
json.dump({"report": rep, "shots_per_launch": shots,
           "command": f"tools/ncu_capture.sh: ncu --set full --clock-control none --import-source on -c 18 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-run-simulation --shots-per-step {shots} --batch {shots}",
           "kernels": kernels}, open(out, "w"), indent=1, sort_keys=True)
for name, k in kernels.items():
    print(f"{name:40s} {k['ms_per_launch']:8.3f} ms  dram {k['dram_bytes']/1e6:9.1f} MB  smem {k['smem_bytes']/1e9:8.2f} GB  inst {k['warp_instructions']/1e6:8.1f} M  issue {k['issue_active_pct']:5.1f}%  warps {k['warps_active_pct']:5.1f}%")
