"""Opcode evidence per kernel of libb200vo.so (cuobjdump -sass): the SASS mnemonics that prove the Blackwell paths
(B200_PROFILING.md): UTC*MMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG = TMA tensor load, UTCBAR = tcgen05.commit,
SYNCS = mbarrier, LDGSTS = cp.async, IDP = dp4a/dp2a, REDUX = warp reduce, FMNMX3 = three-input min.
    python profiles/sass_summary.py > profiles/sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "monocular_visual_odometry_va4mr_b200", "csrc", "libb200vo.so")
KEYS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "LDGSTS", "IDP", "REDUX",
        "FMNMX3", "SHF", "DFMA", "MUFU", "HMMA", "BAR", "LDS", "STS", "ATOMS", "ATOMG", "RED"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
    fn, hist = None, collections.OrderedDict()
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
            hist[fn] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and fn:
            hist[fn][m.group(1)] += 1
            hist[fn]["_total"] += 1
    head = subprocess.run(["cuobjdump", "-lelf", SO], capture_output=True, text=True).stdout.strip().splitlines()
    print(f"# {os.path.relpath(SO, ROOT)}: SASS opcode counts per kernel (static instruction counts; cuobjdump -sass), sm_100a only: {head}")
    print(f"# {'kernel':58s} {'instr':>6s}  " + " ".join(f"{k}" for k in KEYS))
    for fn, h in hist.items():
        cells = " ".join(f"{k}={h[k]}" for k in KEYS if h[k])
        print(f"{fn[:58]:58s} {h['_total']:6d}  {cells}")


if __name__ == "__main__":
    sys.exit(main())
