"""xfg-stark_b200 — B200 (sm_100a) proving backend for the XFG burn-mint STARK.

The product is ``libxfgstark.so`` (CUDA kernels + C ABI, ``include/xfg_stark.h``).  This package is the thin Python
host-side mirror of the reference's operator surface for that path, used by the tests and ``bench.py``:

* ``ProofOptions``           — winter_air::ProofOptions as built at src/burn_mint_prover.rs:28-35
* ``XfgBurnMintProver``      — src/burn_mint_prover.rs:18-237 (``new``/``with_options``/``prove_burn_mint``/...)
* ``XfgBurnMintVerifier`` / ``BatchBurnMintVerifier`` — src/burn_mint_verifier.rs:18-408 over the CUDA batch verifier
* ``AirBuilder`` / ``Context.prove_air`` — generic AIR front-end: a user-defined ``impl Air`` (src/winterfell_air.rs:87-127) as data
* ``Context``                — one device + workspaces; stage-level entry points for kernel parity tests

There is no CPU fallback: importing works anywhere (so the CPU test-suite can check the exported symbols), but creating
a ``Context`` without a CUDA device raises ``XfgError``.  The directory name carries a hyphen, so import it through the
``xfg_stark_b200`` shim at the repo root.
"""
from ._binding import (WideTrace, Context, ProofOptions, StageTimes, XfgBurnMintProver, XfgBurnMintVerifier, BatchBurnMintVerifier, XfgError, AirConsts, STAGE_NAMES, load_library,
                       library_path, EXPORTED_SYMBOLS, FieldExtension, pack_inputs, build_trace, air_compile_check, P)
from .synthetic import synthetic_inputs
from .air import AirBuilder
from . import air
from . import multi

__all__ = ["WideTrace", "Context", "ProofOptions", "StageTimes", "XfgBurnMintProver", "XfgBurnMintVerifier", "BatchBurnMintVerifier", "XfgError", "AirConsts", "STAGE_NAMES",
           "load_library", "library_path", "EXPORTED_SYMBOLS", "FieldExtension", "pack_inputs", "build_trace", "air_compile_check", "synthetic_inputs", "multi", "AirBuilder", "air", "P"]
