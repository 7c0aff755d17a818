#!/bin/bash
# usage: build/ab.sh v1 v2 ... ; runs bench (short) per variant, prints K1/K2 us and ms/step
for v in "$@"; do
  for pdl in 1; do export MCB_PDL=$pdl; echo -n "pdl=$pdl "
  MCSTATE_B200_LIB=$PWD/build/var/lib_$v.so python bench.py --steps 5 --warmup 3 2>/dev/null > /tmp/b_$v.json
  python - <<P
import json
d=json.load(open("/tmp/b_$v.json"))
k=d["roofline"]["kernels"]
print("$v", "ms/step %.3f"%d["ms_per_step"], "K1 %.1f"%k["k_run_compare"]["avg_us"], "K2 %.1f"%k["k_weights_scan"]["avg_us"], "ll", d["log_likelihood"])
P
done; done
