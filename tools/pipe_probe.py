"""Instruction-mix probe on the GPU box: what the FMA pipe (IMAD) can take off the ALU pipe.  python tools/pipe_probe.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import xfg_stark_b200 as xs
names = {0: "xor,shf,xor,shf (ALU only)", 1: "4 IMAD", 2: "xor,IMAD,shf,IMAD", 3: "xor,shf,xor,IMAD", 4: "4 IMAD.WIDE", 5: "xor,shf,IMAD.WIDE,xor", 6: "4 IADD3"}
with xs.Context(device=0, max_n_log2=16) as c:
    for m in range(7):
        print(f"mode {m}: {c.pipe_probe(m):9.1f} G instr/s   {names[m]}")
