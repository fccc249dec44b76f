#!/bin/bash
# Round-end evidence (run under gpurun, ONE GPU): ncu launch list of the bench command + one `--set full` capture of
# four consecutive encoder GEMM launches (qkv+RoPE, proj+residual, fc1+GELU, fc2+residual) of a timed cfg-2 forward.
set -x
ncu --metrics gpu__time_duration.sum --clock-control none -c 1300 --csv --log-file gpurun_out/launches_final.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_final.csv 399 > gpurun_out/launches_final_summary.txt
head -30 gpurun_out/launches_final_summary.txt
# warm-up forward = 244 GEMM-family launches; skip them + patch embed + 2 blocks, then capture one encoder block's four GEMMs
ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 253 -c 4 -o gpurun_out/gemm_full \
  python tools/one_forward.py 1 > gpurun_out/gemm_full.log 2>&1
ncu -i gpurun_out/gemm_full.ncu-rep --page raw --csv > gpurun_out/gemm_full_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/gemm_full_raw.csv | tee gpurun_out/gemm_full_summary.txt
ncu --set full --clock-control none --import-source on -k regex:attention_fwd -s 26 -c 1 -o gpurun_out/attn_full \
  python tools/one_forward.py 1 > gpurun_out/attn_full.log 2>&1
ncu -i gpurun_out/attn_full.ncu-rep --page raw --csv > gpurun_out/attn_full_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/attn_full_raw.csv | tee gpurun_out/attn_full_summary.txt
