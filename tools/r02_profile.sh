#!/bin/bash
# round-2 profiling pass on ONE B200 (never a multi-rank command): per-kernel launch list of one timed bench step, DRAM bytes of
# the dominant GEMM launch, ncu --set full of the shipped head_dim-64 attention kernel.  Outputs under gpurun_out/.
mkdir -p gpurun_out
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_step_launches.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extra-legs --profile-range > gpurun_out/r02_ncu_bench.log 2>&1
echo "launch list rc=$? lines=$(wc -l < gpurun_out/r02_step_launches.csv)"
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
  --clock-control none -k regex:gemm_bf16_tcgen05 --csv --log-file gpurun_out/r02_gemm_dram_bytes.csv python tools/prof_gemm.py > gpurun_out/r02_prof_gemm.log 2>&1
echo "gemm dram rc=$?"; tail -2 gpurun_out/r02_prof_gemm.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_tc2 -c 1 -s 2 -o gpurun_out/r02_attn_tc2_owl python tools/prof_attn.py > gpurun_out/r02_prof_attn.log 2>&1
echo "attn rc=$?"
ncu -i gpurun_out/r02_attn_tc2_owl.ncu-rep --page details --csv > gpurun_out/r02_attn_tc2_owl_ncu_details.csv 2>/dev/null
python tools/summarize_launches.py gpurun_out/r02_step_launches.csv > gpurun_out/r02_step_launch_summary.csv
head -40 gpurun_out/r02_step_launch_summary.csv
