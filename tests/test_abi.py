"""CPU tests (-m "not gpu") of the drop-in boundary: libxfgstark.so loads without a GPU, exports every symbol declared in
include/xfg_stark.h, refuses to create a context without a device (no CPU fallback), and its host-side mirror of
XfgBurnMintProver's input handling (src/burn_mint_prover.rs) agrees with the oracle."""
import os
import re
import subprocess

import numpy as np
import pytest

import orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "xfg_stark.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(xfg_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import xfg_stark_b200 as xs
    lib = xs.load_library()
    syms = header_symbols()
    assert len(syms) >= 18
    for name in syms:
        assert hasattr(lib, name), name
    assert set(xs.EXPORTED_SYMBOLS) <= set(syms)
    out = subprocess.run(["nm", "-D", "--defined-only", xs.library_path()], capture_output=True, text=True).stdout
    for name in syms:
        assert re.search(rf"\bT {name}\b", out), name


def test_rust_sys_crate_declares_every_symbol():
    """rust/xfg-stark-gpu-sys (source only: no cargo in the image) must bind every entry point of the header"""
    rs = open(os.path.join(ROOT, "rust", "xfg-stark-gpu-sys", "src", "lib.rs")).read()
    missing = [n for n in header_symbols() if not re.search(rf"\bpub fn {n}\(", rs)]
    assert missing == []


def test_header_is_plain_c_and_the_library_links_from_c(tmp_path):
    """include/xfg_stark.h must be consumable by a C compiler (cgo / FFI generators read it as C) and the library must link from C"""
    import xfg_stark_b200 as xs
    exe = str(tmp_path / "abi_example")
    libdir = os.path.dirname(xs.library_path())
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "abi_example.c"),
                           "-o", exe, "-L", libdir, "-lxfgstark", "-Wl,-rpath," + libdir])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "ABI_EXAMPLE_OK" in out.stdout, out.stdout + out.stderr


def test_library_is_built_for_sm_100a():
    import xfg_stark_b200 as xs
    out = subprocess.run(["cuobjdump", "-lelf", xs.library_path()], capture_output=True, text=True).stdout
    assert "sm_100a" in out and "sm_90" not in out


def test_no_cpu_fallback():
    import torch
    import xfg_stark_b200 as xs
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(xs.XfgError) as e:
        xs.Context()
    assert e.value.code == 7
    assert xs.load_library().xfg_strerror(5) == b"UnsatisfiedTransitionConstraintError"


def test_product_never_touches_the_oracle():
    """the package and its C sources must not include, import or load anything under oracle/"""
    pkg = os.path.join(ROOT, "xfg-stark_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".inc", "Makefile")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "oracle/" not in txt and "libxfg_oracle" not in txt and "import orc" not in txt, os.path.join(dp, f)
    out = subprocess.run(["ldd", os.path.join(pkg, "libxfgstark.so")], capture_output=True, text=True).stdout
    assert "oracle" not in out


def test_product_has_no_host_execution_of_the_general_pipeline():
    """tests/host_emul runs the general-options kernel bodies on the host for CPU-only parity tests; the product must never do that: the package
    does not load, import or build the harness (the sources only name it in comments), the host backend (`HostBK`) exists only in the harness,
    and the library exports no emulation entry point"""
    pkg = os.path.join(ROOT, "xfg-stark_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".inc", "Makefile")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "go_emul" not in txt and "goemul" not in txt and "HostBK" not in txt, os.path.join(dp, f)
                for line in txt.splitlines():
                    if "host_emul" in line:
                        assert line.lstrip().startswith(("//", "*", "#")), (f, line)      # named in comments only
    syms = subprocess.run(["nm", "-D", "--defined-only", os.path.join(pkg, "libxfgstark.so")], capture_output=True, text=True).stdout
    assert "emul" not in syms
    out = subprocess.run(["ldd", os.path.join(pkg, "libxfgstark.so")], capture_output=True, text=True).stdout
    assert "go_emul" not in out


def test_host_mirror_matches_oracle():
    import xfg_stark_b200 as xs
    for idx in range(6):
        s = orc.synthetic_inputs(idx)
        assert s == xs.synthetic_inputs(idx)
        air = xs.pack_inputs(s["burn"], s["mint"], s["tx_prefix_hash"], s["recipient"], s["secret"], s["network_id"], s["target_chain_id"], s["version"])
        pi, ac, _ = orc.pack_inputs(s["burn"], s["mint"], s["tx_prefix_hash"], s["recipient"], s["secret"], s["network_id"], s["target_chain_id"], s["version"])
        assert list(air.pub_inputs) == [int(v) for v in pi]
        assert [air.txn_hash, air.recipient_hash, air.nullifier, air.commitment] == [int(v) for v in ac]
        for lg in (3, 6, 11):
            assert (xs.build_trace(air, lg) == orc.build_trace(pi, ac, 1 << lg)).all()
    s = orc.synthetic_inputs(0)
    for bad in (dict(burn=1000, mint=1000), dict(mint=1), dict(tx_prefix_hash=bytes(32)), dict(recipient=b"x" * 21), dict(secret=b"abc"), dict(secret=b"abcdef")):
        k = dict(s); k.update(bad)
        with pytest.raises(xs.XfgError) as e:
            xs.pack_inputs(k["burn"], k["mint"], k["tx_prefix_hash"], k["recipient"], k["secret"], 4, 42161, 1)
        assert e.value.code == 8


def test_reference_interface_mirror():
    import xfg_stark_b200 as xs
    o = xs.ProofOptions()
    assert o.as_tuple() == (42, 8, 4, 1, 8, 31)                  # src/burn_mint_prover.rs:28-35
    assert xs.XfgBurnMintProver.xfg_to_atomic_units(0.8) == 8_000_000 and xs.XfgBurnMintProver.atomic_units_to_xfg(8_000_000) == 0.8
    assert xs.STAGE_NAMES[0] == "extend_execution_trace" and len(xs.STAGE_NAMES) == 9
