"""Opcode histogram of the built library's SASS, per kernel: the mnemonics that show which hardware paths a kernel uses
(UTMALDG = TMA tensor loads, UBLKCP = bulk copies, SYNCS = mbarrier, LDGSTS = cp.async, F*2 = packed f32 pipe ...).
Usage: python tools/sass_summary.py > profiles/sass_summary.txt"""
import collections
import re
import subprocess
import sys
from pathlib import Path

LIB = Path(__file__).resolve().parent.parent / "phantomsdr_b200" / "lib" / "libphantomsdr_b200.so"
KEY = ["UTMALDG", "UBLKCP", "UTMASTG", "SYNCS", "LDGSTS", "LDGDEPBAR", "FMUL2", "FADD2", "FFMA2", "FFMA", "FMUL", "FADD", "MUFU", "LDS", "STS",
       "LDG", "STG", "SHFL", "BAR", "ATOM", "RED", "LDC", "IMAD", "PRMT", "LOP3", "HMMA", "UTCMMA"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    name, hist = None, collections.OrderedDict()
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(.*", "", name)
            hist[name] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and name:
            hist[name][m.group(1)] += 1
    print(f"# {LIB.name}: static SASS opcode counts per kernel (cuobjdump -sass, sm_100a)")
    for name, c in hist.items():
        total = sum(c.values())
        keys = " ".join(f"{k}={c[k]}" for k in KEY if c[k])
        print(f"{name}\n    {total} instructions: {keys}")


if __name__ == "__main__":
    sys.exit(main())
