# why is K1 slower at N = 2?  Same box: one rank alone, two ranks with the NCCL exchange, two ranks with the fused peer exchange
mkdir -p gpurun_out
show() { python - "$1" "$2" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print('%-14s N=%d value %.3e step %.4f ms kernel %.4f ms e2e %.2f ms parity %s' % (sys.argv[2], d['n_gpus'], d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['ms_per_step'], d.get('parity_ok')))
PY
}
python bench.py --steps 30 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/probe_single.json 2>/dev/null; show gpurun_out/probe_single.json single
for X in nccl peer; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 30 --warmup 5 --no-extras --exchange $X \
      > gpurun_out/probe_$X.json 2>/dev/null; show gpurun_out/probe_$X.json $X
done
nvidia-smi --query-gpu=index,clocks.sm,clocks.mem,power.draw,power.limit --format=csv
nproc; lscpu | grep -E "Model name|Socket|NUMA node\(s\)|^CPU\(s\)"
