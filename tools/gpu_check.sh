#!/bin/bash
# first-contact GPU run: kernel parity tests (non-GEMM and GEMM in separate processes so a hung kernel
# cannot take the other results down), then micro-benchmarks.  Everything is logged under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/smi.txt 2>&1
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "not gemm" --timeout 120 -p no:cacheprovider > gpurun_out/t_nongemm.log 2>&1
echo "nongemm exit $?" >> gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "gemm" --timeout 120 -p no:cacheprovider > gpurun_out/t_gemm.log 2>&1
echo "gemm exit $?" >> gpurun_out/summary.txt
tail -n 40 gpurun_out/t_nongemm.log
tail -n 60 gpurun_out/t_gemm.log
if [ "$1" == "bench" ]; then
  timeout 600 python tools/bench_kernels.py > gpurun_out/bench_kernels.log 2>&1
  tail -n 40 gpurun_out/bench_kernels.log
fi
cat gpurun_out/summary.txt
