#!/bin/bash
# profiles/ab_run.sh <name>...: parity subset + device-resident step time of every A/B build (on the GPU box)
for v in "$@"; do
  export VGAN_B200_LIB_DIR=$PWD/build_ab/$v
  t=$(python -m pytest tests/test_gpu_haplocart.py -m gpu -x -q -k "oracle or tie or edge or logterm" 2>&1 | tail -1)
  python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.log
  python - "$v" "$t" <<'PY'
import json, sys
v, t = sys.argv[1], sys.argv[2]
try:
    d = json.loads(open("gpurun_out/ab_%s.json" % v).read())
    print("AB %-10s kernel_ms %.4f frac %.3f step_ms %.4f e2e_ms %.3f | %s" % (v, d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["ms_per_step"], d["e2e"]["ms_per_step"], t))
except Exception as e:
    print("AB %-10s FAILED %s | %s" % (v, e, t))
PY
done
