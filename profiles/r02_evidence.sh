#!/bin/bash
# Round-2 evidence run (one B200): GPU tests, the bench line, the reference arm, the launch list of one bench pass, one
# `ncu --set full` capture per kernel of every route, compute-sanitizer over the smoke run.  Everything lands in gpurun_out/;
# summaries are copied to profiles/ by hand.
# Usage (under gpurun): bash profiles/r02_evidence.sh <tag> [tests|notests] [ncu|noncu]
TAG=${1:-r02}
mkdir -p gpurun_out
if [ "${2:-tests}" = "tests" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/${TAG}_pytest.txt
  cat gpurun_out/${TAG}_pytest.txt
  timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python __graft_entry__.py smoke > gpurun_out/${TAG}_sanitizer_memcheck.txt 2>&1; echo "memcheck rc=$?" >> gpurun_out/${TAG}_sanitizer_memcheck.txt
  timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python __graft_entry__.py smoke > gpurun_out/${TAG}_sanitizer_racecheck.txt 2>&1; echo "racecheck rc=$?" >> gpurun_out/${TAG}_sanitizer_racecheck.txt
  tail -n 3 gpurun_out/${TAG}_sanitizer_memcheck.txt; tail -n 3 gpurun_out/${TAG}_sanitizer_racecheck.txt
fi
timeout 1200 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.log
tail -c 400 gpurun_out/bench_${TAG}.log
python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/bench_${TAG}.json').read().strip().splitlines()[-1])
    print('BENCH value %.3e ms %.3f kernel_ms %.4f frac %.3f | e2e %.3e (%.2f ms, marshal %.1f ms, wait %.1f ms) | prepacked %.2f ms' % (
        d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['ms_per_step'],
        1e3 * d['e2e']['host_marshal_seconds_per_step'], 1e3 * d['e2e']['host_wait_seconds_per_step'], d['e2e_prepacked']['ms_per_step']))
    print('latency', json.dumps(d.get('latency'))[:1500])
    for k in ('config4', 'config5'):
        c = d.get(k, {})
        print(k, {x: c.get(x) for x in ('value', 'ms_per_step', 'error')}, c.get('e2e', {}).get('ms_per_step'), c.get('roofline', {}).get('frac'), c.get('roofline', {}).get('kernel_ms'))
    print('cpu', d.get('cpu_baseline', {}).get('kind'), d.get('cpu_baseline', {}).get('value'), d.get('cpu_baseline', {}).get('cores'), 'clocks', d['clocks'], 'launches', d['gpu_launches'])
except Exception as e:
    print('bench parse failed', e)
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${TAG}_reference.json 2>> gpurun_out/bench_${TAG}.log
cat gpurun_out/bench_${TAG}_reference.json | cut -c1-300
if [ "${3:-ncu}" = "ncu" ]; then
  # launch list of one bench pass (cold-cache, serialised: shares of the step, not absolute times)
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${TAG}_launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --sibling-reads 200000 > gpurun_out/${TAG}_ncu_list.log 2>&1
  # one full capture per kernel (a launch after the warm-up of its route)
  for K in hc_read_logterm_kernel hc_collapse_kernel hc_path_kernel hc_path_samples_kernel hc_posterior_kernel hc_recon_kernel hc_pack_kernel hc_expand_kernel \
           euka_read_kernel soibean_edit_kernel soibean_path_kernel soibean_mcmc_kernel gam_count_kernel gam_fill_kernel dedup_insert_kernel; do
    SKIP=2; case $K in soibean_mcmc_kernel|dedup_insert_kernel) SKIP=1;; esac   # (kernels the bench launches only twice)
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:${K} -s $SKIP -c 1 -f -o gpurun_out/${TAG}_${K} \
        python bench.py --steps 1 --warmup 3 --no-cpu-baseline --sibling-reads 200000 > gpurun_out/${TAG}_ncu_${K}.log 2>&1
    python profiles/ncu_summary.py gpurun_out/${TAG}_${K}.ncu-rep > gpurun_out/${TAG}_${K}.txt 2>&1
    echo "== ${K}"; head -3 gpurun_out/${TAG}_${K}.txt
  done
  # the general read kernel (what runs when a batch cannot take the log-term kernel), pinned with VGL_HC_TABLE_MODE=0
  VGL_HC_TABLE_MODE=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:hc_read_kernel -s 2 -c 1 -f -o gpurun_out/${TAG}_hc_read_kernel_general \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/${TAG}_ncu_general.log 2>&1
  python profiles/ncu_summary.py gpurun_out/${TAG}_hc_read_kernel_general.ncu-rep 998978 > gpurun_out/${TAG}_hc_read_kernel_general.txt 2>&1
  head -3 gpurun_out/${TAG}_hc_read_kernel_general.txt
  python profiles/ncu_summary.py gpurun_out/${TAG}_hc_read_logterm_kernel.ncu-rep 998978 > gpurun_out/${TAG}_hc_read_logterm_kernel.txt 2>&1
  rm -f gpurun_out/${TAG}_hc_path_samples_kernel.ncu-rep gpurun_out/${TAG}_gam_*.ncu-rep gpurun_out/${TAG}_dedup_*.ncu-rep gpurun_out/${TAG}_soibean_*.ncu-rep \
        gpurun_out/${TAG}_hc_expand_kernel.ncu-rep gpurun_out/${TAG}_hc_pack_kernel.ncu-rep gpurun_out/${TAG}_hc_posterior_kernel.ncu-rep gpurun_out/${TAG}_hc_collapse_kernel.ncu-rep   # (size limit of the return channel; summaries kept)
fi
