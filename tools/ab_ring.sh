# A/B of the persistent ring mat-vec inside the whole decode step: LLMI_GEMV_RING=mode,cps,depth LLMI_RING_PF=items
run() {
  echo "== $3 RING=$1 PF=$2"
  LLMI_GEMV_RING=$1 LLMI_RING_PF=$2 timeout 600 python bench.py --steps 20 --warmup 5 --no-small --no-cpu --workload $3 2>>gpurun_out/ab.err | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('ms/step', round(d['ms_per_step'],4), [ (k['kernel'][:24], round(k.get('us_per_launch',0),2), round(k.get('GBps',0))) for k in d['kernels'][:1]], 'frac', round(d['roofline']['frac'],3), 'glue ms', round(d['kernels'][-1]['ms_per_step'],3))
"
}
run 1,0,0 0 gemma-3-27b-q4_0
run 2,3,2 0 gemma-3-27b-q4_0
run 2,3,2 32 gemma-3-27b-q4_0
run 2,3,2 96 gemma-3-27b-q4_0
run 2,2,2 96 gemma-3-27b-q4_0
run 2,4,2 96 gemma-3-27b-q4_0
run 2,3,2 64 gemma-3-1b-q4_0
