set -x
EVPP_TC_CLUSTER=42 timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for cfg in "2 1.0" "42 1.0" "42 1.066" "42 1.13"; do
  set -- $cfg
  EVPP_TC_CLUSTER=$1 EVPP_TC_HYBRID_R=$2 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_h.err | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('C=$1 R=$2', d['value'], d['e2e']['value'], d['roofline']['achieved'], d['clocks']['sm_mhz'], d['gpu_launches'])
"
done
