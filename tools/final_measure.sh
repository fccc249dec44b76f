set -x
O=gpurun_out
timeout 700 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee $O/final_pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3 | tee $O/final_smoke.log
python bench.py --steps 20 --warmup 5 > $O/final_bench_n1.json 2>$O/final_bench_n1.err; tail -1 $O/final_bench_n1.json | cut -c1-600
python bench.py --impl reference --steps 3 --warmup 1 > $O/final_bench_reference.json 2>$O/final_bench_reference.err; tail -1 $O/final_bench_reference.json | cut -c1-400
rm -f $O/sweep_R1.jsonl; timeout 400 python tools/sweep.py 1 8 16 32 64 128 256 2>&1 | tail -6
timeout 200 python tools/slam_stream.py > $O/final_slam_stream_224.log 2>&1; tail -4 $O/final_slam_stream_224.log | cut -c1-300
timeout 200 python tools/slam_stream.py --size 512x384 > $O/final_slam_stream_512.log 2>&1; tail -4 $O/final_slam_stream_512.log | cut -c1-300
timeout 300 python tools/parity_report.py --precision bf16 --no-emu 2>&1 | tail -12
timeout 300 python tools/parity_report.py --precision x3 --no-emu 2>&1 | tail -12
