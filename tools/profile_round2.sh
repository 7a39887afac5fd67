#!/bin/bash
# Round-2 evidence run (one B200, under gpurun): the default bench line, its launch list, and ncu --set full captures
# of the kernels the configs rest on.  Every command exits 0 without ncu before it runs under ncu (B200_PROFILING.md).
# Reports land in gpurun_out/; tools/ncu_summary.py turns them into the text summaries under profiles/.
set -u
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
python bench.py --steps 2 --warmup 3 --sub 4 --no-cpu --no-configs > gpurun_out/r2_bench_short.json 2> /dev/null &&
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'hm_weights|hm_quad|hm_epilogue' -c 200 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 2 --warmup 3 --sub 4 --no-cpu --no-configs > /dev/null 2>&1
python tools/dev_c3_time.py 256 5 > gpurun_out/c3_plain.log 2>&1 &&
  $NCU -k regex:hm_quad_rows -s 3 -c 1 -f -o gpurun_out/r2_c3_final python tools/dev_c3_time.py 256 5 > gpurun_out/ncu_c3.log 2>&1
python tools/dev_c3_time.py 256 5 > gpurun_out/c3_plain.log 2>&1 &&
  $NCU -k regex:hm_weights_sweep -s 3 -c 1 -f -o gpurun_out/r2_weights_final python tools/dev_c3_time.py 256 5 > gpurun_out/ncu_w.log 2>&1
python tools/dev_gram_time.py > gpurun_out/gram_plain.log 2>&1 &&
  $NCU -k regex:'hm_gram_kernel|hm_gram_epilogue' -s 2 -c 2 -f -o gpurun_out/r2_gram_final python tools/dev_gram_time.py > gpurun_out/ncu_gram.log 2>&1
python tools/dev_sigma_time.py 64 > gpurun_out/sig_plain.log 2>&1 &&
  $NCU -k regex:hm_sigma_tophat -s 1 -c 1 -f -o gpurun_out/r2_sigma_final python tools/dev_sigma_time.py 64 > gpurun_out/ncu_sig.log 2>&1
cat gpurun_out/c3_plain.log gpurun_out/gram_plain.log gpurun_out/sig_plain.log
for f in gpurun_out/ncu_c3.log gpurun_out/ncu_gram.log gpurun_out/ncu_sig.log; do tail -n 1 $f; done
