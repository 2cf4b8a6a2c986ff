#!/bin/bash
# round-end evidence run on ONE B200: full GPU test suite, smoke, reference arm, full bench line, per-kernel launch list of one
# (half-size) step, DRAM bytes of the dominant GEMM launch with the shipped settings.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider 2>&1 | tail -3 > gpurun_out/final_tests.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1 > gpurun_out/final_smoke.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>gpurun_out/final_ref.err | tail -1 > gpurun_out/final_bench_reference.json
timeout 900 python bench.py --steps 5 --warmup 3 2>gpurun_out/final_bench.err | tail -1 > gpurun_out/final_bench.json
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
  --clock-control none -k regex:gemm_bf16_tcgen05 --csv --log-file gpurun_out/r02_gemm_dram_bytes.csv python tools/prof_gemm.py > gpurun_out/r02_prof_gemm.log 2>&1
timeout 700 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 3200 --csv --log-file gpurun_out/r02_step_launches.csv \
  python bench.py --steps 1 --warmup 1 --searches 4 --no-cpu-baseline --no-extra-legs --profile-range > gpurun_out/r02_ncu_bench.log 2>&1
python tools/summarize_launches.py gpurun_out/r02_step_launches.csv > gpurun_out/r02_step_launch_summary.csv
cat gpurun_out/final_tests.log gpurun_out/final_smoke.log
cut -c1-600 gpurun_out/final_bench_reference.json
cut -c1-300 gpurun_out/final_bench.json
head -12 gpurun_out/r02_step_launch_summary.csv; tail -1 gpurun_out/r02_step_launch_summary.csv
