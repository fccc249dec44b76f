#!/bin/bash
# Round-2 closing evidence (run under gpurun, ONE GPU) for the kernels that changed after tools/profile_r02.sh ran:
# (1) ncu launch list of the bench command; (2) `--set full` captures inside the NVTX range of one timed cfg-2 forward
# (tools/one_forward.py): the last encoder attention (query-tile-pair kernel, n = 768), the first decoder self / cross
# attention (one-query-tile kernel, n = 769), the full-resolution refinenet convolutions (halo-staged; 8 and 16 epilogue
# warps), head.0 and head.2 (halo-staged, N = 128).  Summaries: tools/ncu_summary.py -> gpurun_out/r02b_ncu_*_summary.txt
set -x
OUT=gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 1300 --csv --log-file $OUT/r02b_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/r02b_bench_under_ncu.log 2>&1
python tools/summarize_launches.py $OUT/r02b_launches.csv 399 > $OUT/r02b_launches_summary.txt
head -30 $OUT/r02b_launches_summary.txt
cap() {  # name, kernel regex, skip, count
  ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "timed_forward/" --kernel-name-base demangled \
    -k "regex:$2" -s $3 -c $4 -o $OUT/r02b_$1 python tools/one_forward.py 1 > $OUT/r02b_$1.log 2>&1
  ncu -i $OUT/r02b_$1.ncu-rep --page raw --csv > $OUT/r02b_$1_raw.csv 2>/dev/null
  python tools/ncu_summary.py $OUT/r02b_$1_raw.csv > $OUT/r02b_ncu_$1_summary.txt
  cat $OUT/r02b_ncu_$1_summary.txt | cut -c1-400
}
cap attn_pair_full 'attention_fwd_kernel' 23 1
cap attn_1q_full 'attention_1q_kernel' 0 2
cap conv_full 'gemm_tc_kernel<\(int\)256, \(int\)2, \(int\)0, \(int\)2, \(int\)8' 5 2
cap conv_skip_full 'gemm_tc_kernel<\(int\)256, \(int\)2, \(int\)0, \(int\)2, \(int\)16' 9 2
cap head_full 'gemm_tc_kernel<\(int\)128, \(int\)2, \(int\)[05], \(int\)2' 0 2
rm -f $OUT/r02b_*_raw.csv $OUT/r02b_conv_full.ncu-rep $OUT/r02b_conv_skip_full.ncu-rep $OUT/r02b_attn_pair_full.ncu-rep
