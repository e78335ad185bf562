# A/B of the persistent ring mat-vec inside the whole decode step: LLMI_GEMV_RING=mode,cps,depth,warps
run() {
  echo "== $2 RING=$1"
  LLMI_GEMV_RING=$1 timeout 600 python bench.py --steps 20 --warmup 5 --no-small --no-cpu --workload $2 2>>gpurun_out/ab.err | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('ms/step', round(d['ms_per_step'],4), [ (k['kernel'][:24], round(k.get('us_per_launch',0),2), round(k.get('GBps',0))) for k in d['kernels'][:1]], 'frac', round(d['roofline']['frac'],3), 'glue ms', round(d['kernels'][-1]['ms_per_step'],3))
"
}
for wl in ${WORKLOADS:-gemma-3-27b-q4_0 gemma-3-4b-q4_k_m gemma-3-12b-q8_0 gemma-3-1b-q4_0}; do
  for cfg in ${CONFIGS:-1,0,0,0 0,0,0,0 2,2,2,16 2,3,2,8}; do run $cfg $wl; done
done
