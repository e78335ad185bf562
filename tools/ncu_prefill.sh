python tools/prefill_mode_bench.py gemma-3-27b-q4_0 2048 2 > gpurun_out/plain_prefill.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches_prefill_fast_2l.csv python tools/prefill_mode_bench.py gemma-3-27b-q4_0 2048 2 > gpurun_out/ncu_prefill.log 2>&1
tail -n 2 gpurun_out/plain_prefill.log
