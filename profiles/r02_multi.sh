# multi-GPU check on N GPUs of one box: the two-rank tests, then the bench line under torchrun
N=${1:-2}; TAG=${2:-r02}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -4 | tee gpurun_out/${TAG}_multi_pytest_${N}gpu.txt
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 30 --warmup 5 \
    > gpurun_out/bench_${TAG}_${N}gpu.json 2> gpurun_out/bench_${TAG}_${N}gpu.log
tail -c 400 gpurun_out/bench_${TAG}_${N}gpu.log
python - <<PY
import json
d = json.loads(open('gpurun_out/bench_${TAG}_${N}gpu.json').read().strip().splitlines()[-1])
print('BENCH N=%d value %.3e ms %.3f (sync %.3f) kernel_ms %.4f | e2e %.3e (%.2f ms; marshal %.1f wait %.1f threads %s) | parity_ok %s' % (
    d['n_gpus'], d['value'], d['ms_per_step'], d.get('ms_per_step_one_result_at_a_time', 0.0), d['roofline']['kernel_ms'], d['e2e']['value'], d['e2e']['ms_per_step'],
    1e3*d['e2e']['host_marshal_seconds_per_step'], 1e3*d['e2e']['host_wait_seconds_per_step'], d['e2e'].get('host_marshal_threads'), d.get('parity_ok')))
print('config3', json.dumps(d.get('config3'))[:1200])
PY
