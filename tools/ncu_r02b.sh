set -x
python tools/ncu_case.py Q4_0 2560 10240 0x0 8 > gpurun_out/plain_case4b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemv_ -s 4 -c 3 -o gpurun_out/r02_gate4b python tools/ncu_case.py Q4_0 2560 10240 0x0 8 > gpurun_out/ncu_case4b.log 2>&1
python tools/ncu_case.py Q4_0 21504 5376 0x0 8 > gpurun_out/plain_case27d.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemv_ring -s 4 -c 3 -o gpurun_out/r02_ring_down27b python tools/ncu_case.py Q4_0 21504 5376 0x0 8 > gpurun_out/ncu_case27d.log 2>&1
ls -la gpurun_out/*.ncu-rep
