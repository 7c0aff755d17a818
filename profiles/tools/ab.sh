#!/bin/bash
# usage: profiles/tools/ab.sh v1 v2 ... ; short bench per variant library (build/var/lib_<v>.so): per-kernel us, ms/step
for v in "$@"; do
  MCB_BENCH_CONFIGS=0 MCSTATE_B200_LIB=$PWD/build/var/lib_$v.so python bench.py --steps 5 --warmup 3 2>/dev/null > /tmp/b_$v.json
  python - <<P
import json
d=json.loads(open("/tmp/b_$v.json").read().strip().split(chr(10))[-1])
k=d["roofline"]["kernels"]
print("%-10s"%"$v", "ms/step %.3f"%d["ms_per_step"], " ".join("%s %.1f"%(n[2:6],k[n]["avg_us"]) for n in k), "c3 %.3f/%.3f"%(d["small_filter_c3"]["single_launch"],d["small_filter_c3"]["two_launch"]), "ll", d["log_likelihood"])
P
done
