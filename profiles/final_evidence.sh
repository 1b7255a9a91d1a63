python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/bench_r01_final.json 2> gpurun_out/b11.log
python -c "
import json; d=json.loads(open('gpurun_out/bench_r01_final.json').read()); print('BENCH', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['e2e']['value'], d['raw_path']['ms_per_step'], d['clocks'])"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_final_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/b12.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:hc_read_kernel -s 3 -c 1 -f -o gpurun_out/prof_r01_final python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/b13.log 2>&1
python profiles/step_breakdown.py 2>&1 | tail -1 > gpurun_out/step_breakdown_r01.json
python profiles/bench_siblings.py 2>&1 | tail -1 > gpurun_out/siblings_r01.json
tail -c 600 gpurun_out/step_breakdown_r01.json
