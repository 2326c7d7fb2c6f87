#!/usr/bin/env python
"""sass_pipes.py <object-or-.so> <kernel-substring> — static SASS instruction count of one kernel by issue pipe (ALU / FMA-lite / FMA-heavy /
LSU / other) and the two pipe-time estimates used in DESIGN.md §4: ALU = 2 issue cycles per warp instruction, FMA = 2 (IMAD, IMAD.X, IMAD.MOV)
or 4.9 (IMAD.WIDE, IMAD.HI) - xfg_pipe_probe's measured rates.  A static count: loops are counted once."""
import re
import subprocess
import sys

ALU = ("IADD3", "SEL", "VIADD", "ISETP", "LOP3", "SHF", "LEA", "PRMT", "IABS", "FSEL", "PLOP3", "VIMNMX", "IMNMX", "POPC", "FLO", "BREV", "P2R", "R2P")
HEAVY = ("IMAD.WIDE", "IMAD.HI")


def main():
    obj, pat = sys.argv[1], sys.argv[2]
    names = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    fn = None; counts = {}
    for line in names.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = m.group(1); continue
        if fn is None or pat not in fn:
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_.]+)", line)
        if m:
            counts.setdefault(fn, {}).setdefault(m.group(1), 0); counts[fn][m.group(1)] += 1
    for fn, c in counts.items():
        alu = sum(v for k, v in c.items() if k.startswith(ALU))
        heavy = sum(v for k, v in c.items() if k.startswith(HEAVY))
        lite = sum(v for k, v in c.items() if k.startswith("IMAD") and not k.startswith(HEAVY))
        lsu = sum(v for k, v in c.items() if k.startswith(("LDG", "STG", "LDS", "STS", "LDC", "LDL", "STL", "RED", "ATOM")))
        tot = sum(c.values())
        print(f"{fn[:70]:70s} total {tot:5d}  ALU {alu:5d}  FMA-lite {lite:5d}  FMA-heavy {heavy:4d}  LSU {lsu:4d}  other {tot - alu - lite - heavy - lsu:4d}  |  ALU cycles {2 * alu:6d}  FMA cycles {2 * lite + 4.9 * heavy:8.0f}")


if __name__ == "__main__":
    main()
