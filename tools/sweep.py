"""cfg-5: synthetic 512x384 pair-batch throughput sweep, total batch B in {8..256} split contiguously over R ranks
(BASELINE.json configs[4]).  Run under `gpurun --gpus R`:

    python tools/sweep.py R [B ...]  >  one JSON line per (R, B) in gpurun_out/sweep_R<R>.jsonl

Each point is `bench.py --gpus R --total-pairs B` (strong scaling: B/R pairs per rank, down to 1 pair per rank), launched
under torchrun for R > 1, without the CPU baseline leg.
"""
import json
import os
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")


def main():
    R = int(sys.argv[1])
    Bs = [int(a) for a in sys.argv[2:]] or [8, 16, 32, 64, 128, 256]
    out = os.path.join(ROOT, "gpurun_out", "sweep_R%d.jsonl" % R)
    os.makedirs(os.path.dirname(out), exist_ok=True)
    with open(out, "a") as f:
        for B in Bs:
            if B < R:
                continue
            steps = 10 if B <= 64 else 5
            base = [os.path.join(ROOT, "bench.py"), "--gpus", str(R), "--total-pairs", str(B), "--steps", str(steps), "--warmup", "3",
                    "--no-cpu-baseline"]
            if R == 1:
                cmd = [sys.executable] + base
            else:
                cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(R), "--master-addr",
                       "127.0.0.1", "--master-port", "29531"] + base
            r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
            line = next((ln for ln in r.stdout.splitlines() if ln.startswith("{")), None)
            if line is None:
                print("R=%d B=%d FAILED: %s" % (R, B, (r.stderr or r.stdout)[-500:]), flush=True)
                continue
            d = json.loads(line)
            print("R=%d B=%3d: %.1f pairs/s (%.2f ms/step), e2e %.1f, pairs/rank %d" %
                  (R, B, d["value"], d["ms_per_step"], d["e2e"]["value"], d["config"]["pairs_per_gpu"]), flush=True)
            f.write(line + "\n")
            f.flush()


if __name__ == "__main__":
    main()
