#!/bin/bash
# final single-GPU evidence of a round: GPU tests, smoke, default bench, reference arm, launch list (ncu durations), config sweep
set -x
python -m pytest tests -m gpu -q > gpurun_out/final_pytest.log 2>&1; tail -3 gpurun_out/final_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; tail -2 gpurun_out/final_smoke.log
python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; tail -c 300 gpurun_out/final_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err; tail -c 300 gpurun_out/final_bench_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-run-simulation > gpurun_out/final_ncu_l.log 2>&1
python tools/config_sweep.py > gpurun_out/final_config_sweep.jsonl 2> gpurun_out/final_sweep.err; tail -c 300 gpurun_out/final_sweep.err
