# last captures of the round: the read kernel, the collapse and path kernels (ncu --set full), then tests and the bench line
TAG=${1:-r02z}
mkdir -p gpurun_out
for K in hc_read_logterm_kernel hc_collapse_kernel hc_path_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:${K} -s 2 -c 1 -f -o gpurun_out/${TAG}_${K} \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/${TAG}_ncu_${K}.log 2>&1
  python profiles/ncu_summary.py gpurun_out/${TAG}_${K}.ncu-rep 998978 > gpurun_out/${TAG}_${K}.txt 2>&1
  echo "== ${K}"; head -3 gpurun_out/${TAG}_${K}.txt
done
rm -f gpurun_out/${TAG}_hc_collapse_kernel.ncu-rep gpurun_out/${TAG}_hc_path_kernel.ncu-rep
bash profiles/r02_quick2.sh ${TAG}
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${TAG}_reference.json 2>> gpurun_out/bench_${TAG}.log
cut -c1-200 gpurun_out/bench_${TAG}_reference.json
