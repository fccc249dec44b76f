#!/bin/bash
# per-kernel GPU durations of one steady-state keyframe (eager launches so that ncu sees every kernel)
STA_CUDA_GRAPHS=0 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/slam_launches.csv \
  python tools/slam_stream.py --size ${1:-224x224} --edges 2 --keyframes 4 > gpurun_out/slam_profile.log 2>&1
python tools/summarize_launches.py gpurun_out/slam_launches.csv ${2:-420} 2>/dev/null | head -40
