#!/bin/bash
# A/B of trace-kernel tuning variants (scripts/build_variants.sh): C2 Mrays/s and the full-bench value per variant.
one() {
  name=$1; lib=$2
  r=$(SELAS_CUDA_LIB=$lib python scripts/bench_rows.py --rows raycast 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())['raycast_config2']
print(' '.join('%s=%d' % (k[:6], v['mrays_per_s']) for k,v in d.items() if isinstance(v,dict) and 'mrays_per_s' in v))")
  b=$(SELAS_CUDA_LIB=$lib python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-raycast 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value=%.1fM closest=%d shadow=%d' % (d['value']/1e6, d['mrays_per_s']['closest'], d['mrays_per_s']['occlusion']))")
  echo "$name: $r | $b"
}
one default $PWD/selas_b200/lib/libselas_cuda.so
for v in refill16 refill24 refill28 stack24 t64x16 ctas12 ctas6; do
  [ -f selas_b200/lib_ab/$v/libselas_cuda.so ] && one $v $PWD/selas_b200/lib_ab/$v/libselas_cuda.so
done
