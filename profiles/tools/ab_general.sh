#!/bin/bash
# (variants built by hand into build/var/ with -DMCB_K1_MINBLOCKS_GENERAL=N [-DMCB_BTRS_INLINE=__noinline__])
# A/B of the launch bound of K1 for models without a lean path (SEIR user model, volatility)
for v in g12 g8 g6 g4 g8n g6n; do
  timeout 60 python profiles/tools/user_model_bench.py build/var/libseir_$v.so 2>&1 | tail -1
done
for v in g12 g8 g6; do
  echo -n "$v "; MCSTATE_B200_LIB=$PWD/build/var/lib_$v.so timeout 60 python profiles/tools/vol_bench.py 2>&1 | tail -1
done
