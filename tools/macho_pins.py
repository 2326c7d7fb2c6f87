#!/usr/bin/env python
"""macho_pins.py — command-line front of oracle/a64emu for the reference's shipped binary (/root/reference/test-dist/xfg-stark-cli).

  python tools/macho_pins.py symbols [REGEX]        function symbols (address, size, demangled name) from LC_SYMTAB
  python tools/macho_pins.py imports                libSystem imports behind __stubs and how the interpreter binds each one
  python tools/macho_pins.py words ADDR|REGEX [N]   raw A64 instruction words of a function (for reading immediates such as movz/movk constants)
  python tools/macho_pins.py pins                   protocol constants obtained by EXECUTING the binary's functions (the values
                                                    tests/test_reference_binary_pins.py asserts against include/xfg/spec.h)
  python tools/macho_pins.py prove [--ext 1|2]      run the reference's prove_burn_mint (64 rows) and print the proof's sha256 / size / tail

The binary is opened read-only.  See oracle/a64emu/refbin.py (loader), a64emu.cpp (interpreter), make_reference_vectors.py (golden proofs)."""
import hashlib
import os
import struct
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle", "a64emu"))
import refbin  # noqa: E402

P = 0xFFFFFFFF00000001
R = (1 << 64) % P


def movwide_constants(words):
    """values materialised by movz/movk runs (how rustc builds 64-bit constants such as the Montgomery image of the 2^32-th root of unity)"""
    regs, out = {}, []
    for w in words:
        if (w & 0x1F800000) == 0x12800000:
            opc, hw, imm, rd = (w >> 29) & 3, (w >> 21) & 3, (w >> 5) & 0xFFFF, w & 31
            if opc == 2:
                regs[rd] = imm << (16 * hw)
            elif opc == 3 and rd in regs:
                regs[rd] = (regs[rd] & ~(0xFFFF << (16 * hw))) | (imm << (16 * hw))
            if rd in regs and regs[rd] > 0xFFFFFFFF:
                out.append(regs[rd])
    return sorted(set(out))


def main(argv):
    cmd = argv[1] if len(argv) > 1 else "pins"
    m = refbin.MachO()
    if cmd == "symbols":
        for a, d in m.find(argv[2] if len(argv) > 2 else "."):
            print(f"{a:#x} {m.func_size(a):6d} {d}")
    elif cmd == "imports":
        for a, n in sorted(m.stubs.items()):
            how = "native" if n in refbin.NATIVES else "returns 0" if n in refbin.RET0_IMPORTS else "python" if hasattr(refbin.RefBinary, "imp" + n) else "unimplemented (raises)"
            print(f"{a:#x} {n:32s} {how}")
    elif cmd == "words":
        a = int(argv[2], 16) if argv[2].startswith("0x") else m.find(argv[2], 0)
        n = int(argv[3]) if len(argv) > 3 else m.func_size(a) // 4
        ws = struct.unpack("<%dI" % n, m.read(a, 4 * n))
        for i, w in enumerate(ws):
            print(f"{a + 4 * i:#x}: {w:08x}")
        print("movz/movk constants:", [hex(v) for v in movwide_constants(ws)])
    elif cmd == "pins":
        rb = refbin.RefBinary()
        unmont = lambda v: (v * pow(R, -1, P)) % P
        print("binary sha256           ", hashlib.sha256(m.data).hexdigest())
        print("2^32-th root of unity    ", unmont(rb.call(r"StarkField::get_root_of_unity::", (32,))))
        a = m.find(r"StarkField::get_root_of_unity::", 0)
        ws = struct.unpack("<%dI" % (m.func_size(a) // 4), m.read(a, m.func_size(a)))
        print("  movz/movk constants in get_root_of_unity:", [f"{v:#x} (= Montgomery image of {unmont(v)})" for v in movwide_constants(ws)])
        print("w_8, w_64                ", unmont(rb.call(r"StarkField::get_root_of_unity::", (3,))), unmont(rb.call(r"StarkField::get_root_of_unity::", (6,))))
        print("ProofOptions(42,8,4,None,8,31) struct bytes:", rb.call(r"winter_air::options::ProofOptions::new::", (42, 8, 4, 1, 8, 31)).to_bytes(8, "little")[:6].hex())
        po = rb.put(bytes([2, 42, 8, 4, 8, 31]) + b"\0" * 10); vec = rb.malloc(32)
        rb.call(r"ProofOptions as winter_math::field::traits::ToElements<E>>::to_elements::", (po,), x8=vec)
        cap, ptr, ln = rb.u64s(vec, 3)
        print("ProofOptions::to_elements (quadratic):", [hex(unmont(v)) for v in rb.u64s(ptr, ln)])
        for deg in (1, 2, 3, 4):
            d = rb.malloc(64); rb.call(r"TransitionConstraintDegree::new::", (deg,), x8=d)
            print(f"declared degree {deg}: min_blowup_factor {rb.call(r'TransitionConstraintDegree::min_blowup_factor::', (d,))}, evaluation degree at n=64 "
                  f"{rb.call(r'TransitionConstraintDegree::get_evaluation_degree::', (d, 64))}")
    elif cmd == "prove":
        import make_reference_vectors as mk
        ext = int(argv[argv.index("--ext") + 1]) if "--ext" in argv else 1
        ref = mk.Reference()
        proof, ic = ref.prove64(bytes(range(1, 33)), bytes(range(20)), bytes([1, 2, 3, 4] * 8), (42, 8, 4, ext, 8, 31))
        print(f"{len(proof)} bytes, sha256 {hashlib.sha256(proof).hexdigest()}, {ic} guest instructions, context {proof[:21].hex()}, tail {proof[-9:].hex()}")
    else:
        print(__doc__)


if __name__ == "__main__":
    main(sys.argv)
