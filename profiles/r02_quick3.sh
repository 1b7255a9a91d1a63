# GPU tests, A/B of the log-term kernel's block size, feeder thread sweep, default bench line
TAG=${1:-r02k}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -12 | tee gpurun_out/${TAG}_pytest.txt
bash profiles/ab_run.sh lt768 lt896 lt1024 2>&1 | grep "^AB" | tee gpurun_out/${TAG}_ab.txt
for T in 8 12 16; do
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --feeder-threads $T > gpurun_out/ft_$T.json 2>/dev/null
  python - $T <<'PY'
import json, sys
d = json.loads(open('gpurun_out/ft_%s.json' % sys.argv[1]).read().strip().splitlines()[-1])
print('FEEDER threads %s e2e %.2f ms marshal %.1f wait %.1f' % (sys.argv[1], d['e2e']['ms_per_step'], 1e3*d['e2e']['host_marshal_seconds_per_step'], 1e3*d['e2e']['host_wait_seconds_per_step']))
PY
done 2>&1 | tee gpurun_out/${TAG}_feeder_threads.txt
timeout 900 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.log; tail -c 300 gpurun_out/bench_${TAG}.log
python - <<PY
import json
d = json.loads(open('gpurun_out/bench_${TAG}.json').read().strip().splitlines()[-1])
print('BENCH value %.3e ms %.3f kernel_ms %.4f frac %.3f | e2e %.2f ms (marshal %.1f, wait %.1f) | prepacked %.2f ms' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['e2e']['ms_per_step'], 1e3*d['e2e']['host_marshal_seconds_per_step'], 1e3*d['e2e']['host_wait_seconds_per_step'], d['e2e_prepacked']['ms_per_step']))
print('latency', json.dumps(d.get('latency'))[:2500])
for k in ('config4', 'config5'):
    c = d.get(k, {})
    print(k, {x: c.get(x) for x in ('value', 'ms_per_step', 'error')}, json.dumps(c.get('e2e'))[:600], c.get('e2e_from_alignments', {}).get('ms_per_step'), c.get('roofline', {}).get('frac'))
PY
