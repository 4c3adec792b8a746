#!/bin/bash
# One-GPU evidence pass for profiles/: tests, bench lines of every workload, ncu launch list and --set full captures of cfg5.
# usage (through gpurun): bash tools/measure_1gpu.sh <tag>     -> gpurun_out/<tag>_*
tag=${1:-rX}
out=gpurun_out
mkdir -p $out
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 > $out/${tag}_pytest_gpu.log
timeout 400 python bench.py > $out/${tag}_bench_cfg5.json 2> $out/${tag}_bench_cfg5.err
for w in cfg1 cfg2 cfg2r cfg3 cfg4; do
  timeout 300 python bench.py --workload $w > $out/${tag}_bench_$w.json 2> /dev/null
done
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches_cfg5.csv \
  python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --sustained-seconds 0 > $out/${tag}_ncu_launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k 'regex:gibbs_chain_kernel|gibbs_gemm_kernel_2cta' -s 6 -c 2 -f -o $out/${tag}_ncu_gemm \
  python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --sustained-seconds 0 > $out/${tag}_ncu_gemm.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k 'regex:update_w_kernel' -s 3 -c 1 -f -o $out/${tag}_ncu_update \
  python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --sustained-seconds 0 > $out/${tag}_ncu_update.log 2>&1
cat $out/${tag}_pytest_gpu.log
cut -c1-260 $out/${tag}_bench_cfg5.json
