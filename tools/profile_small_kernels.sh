#!/bin/bash
# ncu --set full captures of the kernels that are not the headline quadrature (run under gpurun on one B200).
# Each command first exits 0 without ncu (B200_PROFILING.md); reports land in gpurun_out/, summaries are made
# here afterwards with tools/ncu_summary.py full <rep> profiles/<name>.txt
set -u
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
python tools/dev_beta_time.py 256 1024 > gpurun_out/beta_plain.log 2>&1 &&
  $NCU -k regex:hm_beta2h -s 2 -c 1 -f -o gpurun_out/r2_beta2h python tools/dev_beta_time.py 256 1024 > gpurun_out/ncu_beta.log 2>&1
python tools/dev_c5_time.py 256 256 > gpurun_out/c5_plain.log 2>&1 &&
  $NCU -k regex:'hm_window_nfw|hm_window_iso' -s 2 -c 2 -f -o gpurun_out/r2_windows python tools/dev_c5_time.py 256 256 > gpurun_out/ncu_win.log 2>&1
python tools/dev_gram_time.py > gpurun_out/gram_plain.log 2>&1 &&
  $NCU -k regex:'hm_gram_epilogue|hm_gram_weights' -s 2 -c 2 -f -o gpurun_out/r2_gram_small python tools/dev_gram_time.py > gpurun_out/ncu_gram_small.log 2>&1
cat gpurun_out/beta_plain.log gpurun_out/c5_plain.log gpurun_out/gram_plain.log
for f in gpurun_out/ncu_beta.log gpurun_out/ncu_win.log gpurun_out/ncu_gram_small.log; do tail -n 2 $f; done
