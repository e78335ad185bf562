for late in 0 1; do echo "LATE=$late"; LLMI_RING_LATE=$late timeout 600 python bench.py --steps 20 --warmup 5 --no-small --no-cpu 2>>gpurun_out/ab.err | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('ms/step', round(d['ms_per_step'],4), [ (k['kernel'][:16], round(k.get('ms_per_step',0),3)) for k in d['kernels']])
"; done
