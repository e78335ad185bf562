set -x
python tools/profile_step.py gemma-3-27b-q4_0 8 2 > gpurun_out/plain_step.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_launches_27b_step.csv python tools/profile_step.py gemma-3-27b-q4_0 8 2 > gpurun_out/ncu_step.log 2>&1
python tools/ncu_case.py Q4_0 5376 21504 0x0 8 > gpurun_out/plain_case.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemv_ring -s 4 -c 3 -o gpurun_out/r02_ring_gate27b python tools/ncu_case.py Q4_0 5376 21504 0x0 8 > gpurun_out/ncu_case.log 2>&1
FAST=1 CASE=2 python tools/prefill_gemm_bench.py 1024 > gpurun_out/plain_fast.log 2>&1 && \
FAST=1 CASE=2 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 2 -c 2 -o gpurun_out/r02_gemm_bf16_gate27b python tools/prefill_gemm_bench.py 1024 > gpurun_out/ncu_fast.log 2>&1
tail -2 gpurun_out/ncu_step.log gpurun_out/ncu_case.log gpurun_out/ncu_fast.log
ls -la gpurun_out/*.ncu-rep gpurun_out/r02_launches_27b_step.csv
