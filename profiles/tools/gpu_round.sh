#!/bin/bash
# one GPU round: tests, bench, ncu launch list [, full captures].  usage: gpu_round.sh <tag> [full] [notests]
tag=${1:-x}
if [ "$3" != "notests" ]; then
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/pytest_gpu_$tag.txt; tail -3 gpurun_out/pytest_gpu_$tag.txt
fi
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
python -c "
import json;d=json.loads(open('gpurun_out/bench_$tag.json').read().strip().split(chr(10))[-1]);print(d['value'],d['ms_per_step'],d['log_likelihood']);print({k:round(v['avg_us'],1) for k,v in d['roofline']['kernels'].items()}); print(d['small_filter_c3']); print(json.dumps(d.get('configs'), indent=1)[:3000])"
python profiles/tools/weights_bench.py > gpurun_out/weights_$tag.txt 2>&1; cat gpurun_out/weights_$tag.txt
if [ "$2" = "full" ]; then
MCB_BENCH_CONFIGS=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 330 --csv --log-file gpurun_out/ncu_launches_$tag.csv python bench.py --steps 1 --warmup 1 > gpurun_out/ncu_l.log 2>&1
for k in k_run_compare k_weights_scan k_ancestors; do
MCB_BENCH_CONFIGS=0 ncu --set full --clock-control none --import-source on -k regex:$k -s 50 -c 1 -o gpurun_out/${k}_$tag -f python bench.py --steps 1 --warmup 1 > gpurun_out/ncu_$k.log 2>&1
done
fi
