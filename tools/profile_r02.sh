#!/bin/bash
# Round-2 evidence (run under gpurun, ONE GPU).  (1) ncu launch list of the bench command; (2) `--set full` captures inside the
# NVTX range of one timed cfg-2 forward (tools/one_forward.py): one encoder block's four GEMMs, the last encoder attention +
# the first decoder self / cross attention (n = 769, kv_batch_shift), refinenet convs, head.0, head.2 (fused post-process),
# LayerNorm and the bilinear upsample.  Summaries: tools/ncu_summary.py -> gpurun_out/r02_ncu_*_summary.txt
set -x
OUT=gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 1300 --csv --log-file $OUT/r02_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/r02_bench_under_ncu.log 2>&1
python tools/summarize_launches.py $OUT/r02_launches.csv 399 > $OUT/r02_launches_summary.txt
head -30 $OUT/r02_launches_summary.txt
cap() {  # name, kernel regex, skip, count
  ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "timed_forward/" --kernel-name-base demangled \
    -k "regex:$2" -s $3 -c $4 -o $OUT/r02_$1 python tools/one_forward.py 1 > $OUT/r02_$1.log 2>&1
  ncu -i $OUT/r02_$1.ncu-rep --page raw --csv > $OUT/r02_$1_raw.csv 2>/dev/null
  python tools/ncu_summary.py $OUT/r02_$1_raw.csv > $OUT/r02_ncu_$1_summary.txt
  cat $OUT/r02_ncu_$1_summary.txt | cut -c1-400
}
cap gemm_full 'gemm_tc_kernel<\(int\)256, \(int\)0' 5 4
cap attn_full 'attention_fwd_kernel' 23 3
cap conv_full 'gemm_tc_kernel<\(int\)256, \(int\)1, \(int\)0' 10 4
cap head_full 'gemm_tc_kernel<\(int\)128, \(int\)1' 0 2
cap ln_ups_full 'layernorm_kernel|upsample2x_kernel' 60 4
rm -f $OUT/r02_*_raw.csv
