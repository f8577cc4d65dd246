"""Join an ncu SASS-level source page with nvdisasm line info: instructions executed and stall samples per CUDA source line.
usage: sass_lines.py <rep.ncu-rep> <kernel-regex> <mangled-substring> <cubin-sass-with-lineinfo> <source.cu>"""
import collections, csv, io, re, subprocess, sys
rep, kre, mangled, sassfile, srcfile = sys.argv[1:6]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
# first instance only
start = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
seg = rows[start[0] + 1:(start[1] if len(start) > 1 else len(rows))]
H = {x: i for i, x in enumerate(seg[0])}
stat = {}
for r in seg[1:]:
    try:
        addr = int(r[H["Address"]], 16) if r[H["Address"]].startswith("0x") else int(r[H["Address"]])
    except Exception:
        continue
    stat[addr] = (float(r[H["Instructions Executed"]] or 0), float(r[H["# Samples"]] or 0), r[H["Source"]])
base = min(stat)
# nvdisasm: find function section and its line annotations
lines = open(sassfile).read().split("\n")
cur_line, in_fn, per_line = None, False, collections.defaultdict(lambda: [0.0, 0.0, 0])
for ln in lines:
    if ln.startswith("//--------------------- .text."):
        in_fn = mangled in ln
        continue
    if not in_fn:
        continue
    m = re.search(r'//## File ".*?([^/"]+)", line (\d+)', ln)
    if m:
        cur_line = (m.group(1), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m and cur_line:
        off = int(m.group(1), 16)
        st = stat.get(base + off)
        if st:
            per_line[cur_line][0] += st[0]; per_line[cur_line][1] += st[1]; per_line[cur_line][2] += 1
tot_i = sum(v[0] for v in per_line.values()) or 1
tot_s = sum(v[1] for v in per_line.values()) or 1
srcl = open(srcfile).read().split("\n")
print(f"total executed {tot_i:.3g}, samples {tot_s:.0f}")
for (fn, l), v in sorted(per_line.items(), key=lambda x: -x[1][0])[:int(sys.argv[6]) if len(sys.argv) > 6 else 40]:
    text = srcl[l - 1].strip()[:95] if fn in srcfile and l - 1 < len(srcl) else fn
    print(f"{fn[:12]:12s} L{l:4d} inst {v[0] / tot_i * 100:5.1f}%  samples {v[1] / tot_s * 100:5.1f}%  sass {v[2]:4d} | {text}")
