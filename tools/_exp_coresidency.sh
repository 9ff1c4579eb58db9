for cfg in "" "GPSO_SCR_STAGES=4" "GPSO_PRODUCT_PRIO=0" "GPSO_SCR_STAGES=4 GPSO_PRODUCT_PRIO=0" "GPSO_SCR_XCOV_PAD=100000" "GPSO_SCR_STAGES=6"; do
  echo "== $cfg"
  env $cfg timeout 300 python tools/screen_trace.py 2100000 gpurun_out/r02d_trace_$(echo $cfg | tr ' =' '__').json 2>&1 | grep digits3 | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l.split(' ',1)[1]); print('overlap ms', round(d['overlap']['ms_per_step'],2),'no_overlap', round(d['no_overlap']['ms_per_step'],2),'prod', round(d['overlap']['screen_product_ms_per_step'],2), {k:round(v,3) for k,v in d['summary'].items()})
"
done
