# quick check: GPU tests, the default bench line, the launch list of the raw route (per-kernel times of one bench pass)
TAG=${1:-r02q}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/${TAG}_pytest.txt
timeout 900 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.log; tail -c 300 gpurun_out/bench_${TAG}.log
python - <<PY
import json
d = json.loads(open('gpurun_out/bench_${TAG}.json').read().strip().splitlines()[-1])
print('BENCH value %.3e ms %.3f kernel_ms %.4f frac %.3f | e2e %.2f ms (marshal %.1f, wait %.1f) | prepacked %.2f ms' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['e2e']['ms_per_step'], 1e3*d['e2e']['host_marshal_seconds_per_step'], 1e3*d['e2e']['host_wait_seconds_per_step'], d['e2e_prepacked']['ms_per_step']))
print('latency', json.dumps(d.get('latency'))[:2500])
for k in ('config4', 'config5'):
    c = d.get(k, {})
    print(k, {x: c.get(x) for x in ('value', 'ms_per_step', 'error')}, c.get('e2e', {}).get('ms_per_step'), c.get('roofline', {}).get('frac'))
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/${TAG}_ncu_list.log 2>&1
python profiles/launch_table.py gpurun_out/${TAG}_launches.csv 2>&1 | head -30
