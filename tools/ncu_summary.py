"""Summarise an .ncu-rep: per-kernel headline metrics + SASS-level opcode mix / stall reasons / hottest source lines."""
import collections, csv, io, re, subprocess, sys

rep = sys.argv[1]
pat = sys.argv[2] if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "sm__cycles_elapsed.max",
        "launch__shared_mem_per_block_dynamic"]
for row in rows[2:]:
    print("-----")
    for w in want:
        if w in hdr:
            print(f"{w:80s} {row[hdr.index(w)][:80]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + (["--kernel-name", "regex:" + pat] if pat else []), capture_output=True, text=True).stdout
inst, cur = [], None
for r in csv.reader(io.StringIO(src)):
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}; inst.append(cur); continue
    if cur is not None:
        cur["rows"].append(r)
seen = set()
for k in inst:
    if k["name"] in seen or not k["rows"]:
        continue
    seen.add(k["name"])
    h = k["rows"][0]; data = k["rows"][1:]
    H = {x: i for i, x in enumerate(h)}
    def f(r, key):
        try: return float(r[H[key]])
        except Exception: return 0.0
    tot = sum(f(r, "Instructions Executed") for r in data) or 1
    print("\n=====", k["name"][:90], "SASS", len(data), "executed", int(tot))
    stalls = [x for x in h if x.startswith("stall_") and "Not Issued" not in x]
    ss = {s: sum(f(r, s) for r in data) for s in stalls}
    tots = sum(ss.values()) or 1
    print("stalls:", {s[6:]: f"{v / tots * 100:.1f}%" for s, v in sorted(ss.items(), key=lambda x: -x[1]) if v / tots > 0.01})
    op = collections.Counter()
    for r in data:
        m = re.match(r"\s*(@!?U?P\w+\s+)?([A-Z0-9_]+)", r[H["Source"]])
        if m: op[m.group(2)] += f(r, "Instructions Executed")
    print("opcodes:", {k2: f"{v / tot * 100:.1f}%" for k2, v in op.most_common(16)})
    print("smem wavefronts", int(sum(f(r, "L1 Wavefronts Shared") for r in data)), "ideal", int(sum(f(r, "L1 Wavefronts Shared Ideal") for r in data)))
