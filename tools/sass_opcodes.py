"""Blackwell-native evidence without a GPU: SASS opcode census of libsta_b200.so per kernel (cuobjdump -sass).

    python tools/sass_opcodes.py > profiles/r02_sass_opcodes.txt

UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG / UTMAREDG = TMA tensor loads / stores / reductions,
HMMA would be the legacy mma.sync path (must be 0), MUFU.EX2 = the softmax / GELU exponentials, FMNMX3 = 3-input max.
"""
import collections
import os
import re
import subprocess
import sys

LIB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "vista_slam_b200", "csrc", "libsta_b200.so")
OPS = ["UTCHMMA.2CTA", "UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTCBAR", "SYNCS", "HMMA", "MUFU.EX2", "FMNMX3",
       "STL", "LDL"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    cur = None
    for ln in out.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = cur.replace("(anonymous namespace)::", "").replace("sta::", "")
            cur = re.sub(r"\(.*", "", cur)
            per.setdefault(cur, collections.Counter())
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if not m:
            continue
        op = m.group(1)
        per[cur]["_instr"] += 1
        for o in OPS:
            if op == o or op.startswith(o + "."):
                if o == "UTCHMMA" and ".2CTA" in op:
                    continue
                per[cur][o] += 1
    tot = collections.Counter()
    print("SASS opcode census of vista_slam_b200/csrc/libsta_b200.so (sm_100a), per kernel; columns:", " ".join(OPS))
    for k, c in per.items():
        if not any(c[o] for o in OPS[:8]) and c["_instr"] < 200:
            continue
        print("%-78s instr %6d | %s" % (k[:78], c["_instr"], " ".join("%s=%d" % (o, c[o]) for o in OPS if c[o])))
        tot.update(c)
    print("TOTAL over %d kernels: %s" % (len(per), " ".join("%s=%d" % (o, tot[o]) for o in OPS)))
    assert tot["HMMA"] == 0, "legacy HMMA found"


if __name__ == "__main__":
    main()
