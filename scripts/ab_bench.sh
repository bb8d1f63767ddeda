#!/bin/bash
# A/B of shade-stage variants on the full bench workload: prints value / ms_per_step / kernel shares per variant.
run() { echo "== $1"; shift; env "$@" python bench.py --steps 4 --warmup 2 --no-cpu-baseline --no-raycast 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('value %.1fM  ms/step %.2f  e2e %.1fM  share %s  closest Mrays %.0f frac %.3f' % (d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, {k:round(v,3) for k,v in r['kernel_share_of_step'].items()}, d['mrays_per_s']['closest'], r['frac']))"; }
run "outline, sort from bounce 1" SELAS_SORT_FROM_BOUNCE=1
run "outline, sort from bounce 0" SELAS_SORT_FROM_BOUNCE=0
run "outline, no sort" SELAS_SORT_FROM_BOUNCE=9999
run "inline, sort from bounce 1" SELAS_CUDA_LIB=$PWD/selas_b200/lib_ab/libselas_cuda.so SELAS_SORT_FROM_BOUNCE=1
run "inline, no sort" SELAS_CUDA_LIB=$PWD/selas_b200/lib_ab/libselas_cuda.so SELAS_SORT_FROM_BOUNCE=9999
