"""Summarise an `ncu --set full` report exported with `ncu -i X.ncu-rep --page raw --csv`: one line per captured launch
with the metrics DESIGN.md / bench.py quote (duration, DRAM bytes, tensor-pipe activity, registers)."""
import csv
import sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if not l.startswith("==")))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "lts__t_sector_hit_rate.pct",
        "sm__cycles_elapsed.max", "smsp__issue_active.avg.pct", "launch__grid_size", "launch__block_size"]
idx = {h: i for i, h in enumerate(hdr)}
kcol = idx.get("Kernel Name")
for r in data:
    out = {"kernel": r[kcol][:70]}
    for w in want:
        cands = [h for h in hdr if h.startswith(w)]
        if cands:
            i = idx[cands[0]]
            out[w] = "%s %s" % (r[i], units[i])
    print(out)
