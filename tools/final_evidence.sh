#!/bin/bash
# round-end evidence run on one B200: full GPU test suite, smoke, full bench line, VQA probe, per-kernel launch list of one step
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 300 -p no:cacheprovider 2>&1 | tail -3 > gpurun_out/final_tests.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1 > gpurun_out/final_smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 2>&1 | tail -1 > gpurun_out/final_bench.json
timeout 600 python tools/bench_vqa.py all > gpurun_out/final_bench_vqa.log 2>&1
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/final_step_launches.csv \
  python bench.py --steps 1 --warmup 1 --searches 8 --no-cpu-baseline --profile-range > gpurun_out/final_ncu_bench.log 2>&1
cat gpurun_out/final_tests.log gpurun_out/final_smoke.log
cut -c1-400 gpurun_out/final_bench.json
tail -5 gpurun_out/final_bench_vqa.log
wc -l gpurun_out/final_step_launches.csv
