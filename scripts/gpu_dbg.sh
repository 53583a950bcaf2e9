#!/bin/bash
mkdir -p gpurun_out
for w in c op_nograd op_detached prob_badmode op; do
  echo "== $w"; LD_PRELOAD=$PWD/scripts/dbg/libsegv_bt.so timeout 120 python scripts/dbg/err_path.py $w 2>&1 | tail -40; echo "exit ${PIPESTATUS[0]}"
done > gpurun_out/r02_dbg_err_path.log 2>&1
cat gpurun_out/r02_dbg_err_path.log
