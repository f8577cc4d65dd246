#!/bin/bash
# ncu --set full capture of one un-pipelined 16384-shot batch (first 18 launches = every kernel of the pipeline once per
# side) + the summaries bench.py and DESIGN.md read; the report itself travels back only when it is small enough
TAG=${1:-r2b}
python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-run-simulation --shots-per-step 16384 --batch 16384 > gpurun_out/${TAG}_plain16k.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -c 18 -f -o gpurun_out/${TAG}_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-run-simulation --shots-per-step 16384 --batch 16384 > gpurun_out/${TAG}_ncu_full.log 2>&1
python tools/ncu_traffic.py gpurun_out/${TAG}_full.ncu-rep 16384 gpurun_out/${TAG}_traffic.json
python tools/ncu_summary.py gpurun_out/${TAG}_full.ncu-rep > gpurun_out/${TAG}_ncu_full_summary.txt 2>&1
ls -la gpurun_out/
SZ=$(stat -c %s gpurun_out/${TAG}_full.ncu-rep)
if [ "$SZ" -gt 50000000 ]; then rm gpurun_out/${TAG}_full.ncu-rep; echo "report removed ($SZ bytes)"; fi
