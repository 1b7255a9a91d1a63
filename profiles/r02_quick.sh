timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 600 python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/bench_r02b.json 2> gpurun_out/bench_r02b.log; tail -c 300 gpurun_out/bench_r02b.log
python - <<PY
import json
d = json.loads(open('gpurun_out/bench_r02b.json').read().strip().splitlines()[-1])
print('BENCH value %.3e ms %.3f kernel_ms %.4f frac %.3f | e2e %.2f ms (marshal %.1f) | prepacked %.2f ms' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['e2e']['ms_per_step'], 1e3*d['e2e']['host_marshal_seconds_per_step'], d['e2e_prepacked']['ms_per_step']))
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hc_read_logterm -s 2 -c 1 -f -o gpurun_out/r02b_logterm python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r02b_ncu.log 2>&1
python profiles/ncu_summary.py gpurun_out/r02b_logterm.ncu-rep 998978 2>&1 | tee gpurun_out/r02b_logterm.txt
