"""ctypes binding of the CPU oracle (oracle/_build/libxfg_oracle.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this module; the
product package (xfg-stark_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(_ROOT, "oracle", "_build", "libxfg_oracle.so")
P = 0xFFFFFFFF00000001
STAGES = ["extend_execution_trace", "compute_execution_trace_commitment", "evaluate_constraints",
          "commit_to_constraint_evaluations", "build_deep_composition_poly", "evaluate_deep_composition_poly",
          "compute_fri_layers", "determine_query_positions", "build_proof_object"]


def build():
    """Compile the oracle (gcc only, no GPU) unless the .so is already newer than its sources."""
    src_dir = os.path.join(_ROOT, "oracle")
    srcs = [os.path.join(src_dir, f) for f in os.listdir(src_dir) if f.endswith((".hpp", ".cpp"))]
    srcs.append(os.path.join(_ROOT, "include", "xfg", "spec.h"))
    if os.path.exists(_SO) and all(os.path.getmtime(_SO) >= os.path.getmtime(s) for s in srcs):
        return _SO
    subprocess.check_call(["make", "-C", src_dir], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        u64, u32, sz, vp, cp = C.c_uint64, C.c_uint32, C.c_size_t, C.c_void_p, C.c_char_p
        for name in ("orc_fadd", "orc_fsub", "orc_fmul", "orc_fmul_slow", "orc_fpow"):
            getattr(L, name).restype = u64
            getattr(L, name).argtypes = [u64, u64]
        L.orc_finv.restype = u64; L.orc_finv.argtypes = [u64]
        L.orc_root_of_unity.restype = u64; L.orc_root_of_unity.argtypes = [C.c_uint]
        L.orc_f2_mul.argtypes = [vp, vp, vp]; L.orc_f2_inv.argtypes = [vp, vp]
        L.orc_ntt.argtypes = [vp, sz, C.c_int, C.c_int]
        L.orc_lde.argtypes = [vp, sz, sz, u64, vp]
        L.orc_interpolate_offset.argtypes = [vp, sz, u64]
        L.orc_blake3.argtypes = [vp, sz, vp]; L.orc_keccak256.argtypes = [vp, sz, vp]
        L.orc_hash_rows.argtypes = [vp, sz, sz, vp]
        L.orc_merkle.argtypes = [vp, sz, vp, vp]
        L.orc_merkle_prove_batch.restype = C.c_long; L.orc_merkle_prove_batch.argtypes = [vp, sz, vp, sz, vp, sz]
        L.orc_pack_inputs.argtypes = [u64, u64, vp, vp, sz, vp, sz, u32, u32, u32, vp, vp, vp, cp, sz]
        L.orc_build_trace.argtypes = [vp, vp, sz, vp]
        L.orc_validate_options.argtypes = [vp, cp, sz]
        L.orc_prove.argtypes = [vp, C.c_uint, vp, vp, vp, vp, sz, vp, vp, C.c_int, cp, sz]
        L.orc_verify.argtypes = [vp, sz, vp, vp, vp, cp, sz]
        L.orc_prove_air.argtypes = [vp, C.c_uint, vp, vp, vp, vp, vp, vp, vp, vp, sz, vp, vp, C.c_int, cp, sz]
        L.orc_verify_air.argtypes = [vp, sz, vp, vp, vp, vp, vp, vp, vp, cp, sz]
        L.orc_debug_get.restype = C.c_long; L.orc_debug_get.argtypes = [cp, vp, sz]
        L.orc_set_threads.argtypes = [C.c_int]
        L.orc_fri_fold.argtypes = [vp, sz, C.c_int, sz, vp, vp]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def fmul(a, b): return lib().orc_fmul(a, b)
def finv(a): return lib().orc_finv(a)
def fpow(a, e): return lib().orc_fpow(a, e)
def root_of_unity(k): return lib().orc_root_of_unity(k)
def max_threads(): return lib().orc_max_threads()
def set_threads(t): lib().orc_set_threads(t)


def blake3(data: bytes) -> bytes:
    out = C.create_string_buffer(32); lib().orc_blake3(data, len(data), out); return out.raw


def keccak256(data: bytes) -> bytes:
    out = C.create_string_buffer(32); lib().orc_keccak256(data, len(data), out); return out.raw


def ntt(data: np.ndarray, deg=1, mode=0) -> np.ndarray:
    """mode 0 forward, 1 interpolate_poly, 2 naive DFT.  data: n (deg=1) or n x 2 (deg=2) canonical u64."""
    a = np.ascontiguousarray(data, dtype=np.uint64).copy()
    lib().orc_ntt(_p(a), a.size // deg, deg, mode)
    return a


def lde(coeffs: np.ndarray, blowup=8, offset=7) -> np.ndarray:
    c = np.ascontiguousarray(coeffs, dtype=np.uint64)
    out = np.empty(c.size * blowup, dtype=np.uint64)
    lib().orc_lde(_p(c), c.size, blowup, offset, _p(out))
    return out


def interpolate_offset(evals: np.ndarray, offset=7) -> np.ndarray:
    a = np.ascontiguousarray(evals, dtype=np.uint64).copy()
    lib().orc_interpolate_offset(_p(a), a.size, offset)
    return a


def hash_rows(colmajor: np.ndarray) -> np.ndarray:
    """colmajor: (cols, rows) u64 -> (rows, 32) u8 leaf digests."""
    m = np.ascontiguousarray(colmajor, dtype=np.uint64)
    cols, rows = m.shape
    out = np.empty((rows, 32), dtype=np.uint8)
    lib().orc_hash_rows(_p(m), rows, cols, _p(out))
    return out


def merkle(leaves: np.ndarray):
    """leaves: (n, 32) u8 -> (root bytes, nodes (n, 32) u8)."""
    lv = np.ascontiguousarray(leaves, dtype=np.uint8)
    n = lv.shape[0]
    root = C.create_string_buffer(32)
    nodes = np.zeros((n, 32), dtype=np.uint8)
    lib().orc_merkle(_p(lv), n, root, _p(nodes))
    return root.raw, nodes


def merkle_prove_batch(leaves: np.ndarray, indexes) -> bytes:
    lv = np.ascontiguousarray(leaves, dtype=np.uint8)
    idx = np.asarray(indexes, dtype=np.uint64)
    out = C.create_string_buffer(1 << 20)
    k = lib().orc_merkle_prove_batch(_p(lv), lv.shape[0], _p(idx), idx.size, out, len(out))
    if k < 0:
        raise ValueError("prove_batch failed")
    return out.raw[:k]


def fri_fold(evals: np.ndarray, alpha) -> np.ndarray:
    """evals: (Nl, deg) u64 element-major; folding factor 8, offset 7 -> (Nl/8, deg)."""
    e = np.ascontiguousarray(evals, dtype=np.uint64); nl, deg = e.shape
    a = np.ascontiguousarray(alpha, dtype=np.uint64)
    out = np.empty((nl // 8, deg), dtype=np.uint64)
    lib().orc_fri_fold(_p(e), nl, deg, 8, _p(a), _p(out))
    return out


DEFAULT_OPTIONS = (42, 8, 4, 1, 8, 31)   # num_queries, blowup, grinding, ext(1=None,2=Quadratic), folding, rem_max_deg


def _opts(o):
    return np.asarray(o, dtype=np.uint32)


def pack_inputs(burn, mint, tx_prefix_hash: bytes, recipient: bytes, secret: bytes, network_id, target_chain_id, version):
    """-> (pub_inputs[12] u64, consts[4] u64, secret_elem).  Raises ValueError with the reference's message."""
    pi = np.zeros(12, dtype=np.uint64); ac = np.zeros(4, dtype=np.uint64); se = C.c_uint64(0)
    err = C.create_string_buffer(256)
    rc = lib().orc_pack_inputs(burn, mint, tx_prefix_hash, recipient, len(recipient), secret, len(secret), network_id,
                               target_chain_id, version, _p(pi), _p(ac), C.byref(se), err, len(err))
    if rc:
        raise ValueError(err.value.decode())
    return pi, ac, se.value


def build_trace(pi, ac, n) -> np.ndarray:
    out = np.empty((7, n), dtype=np.uint64)
    lib().orc_build_trace(_p(pi), _p(ac), n, _p(out))
    return out


def prove(trace: np.ndarray, pi, ac, options=DEFAULT_OPTIONS, keep_debug=False, want_times=False):
    """trace: (7, n) u64 column-major.  Returns proof bytes (and a dict of stage ms when want_times)."""
    t = np.ascontiguousarray(trace, dtype=np.uint64)
    n = t.shape[1]
    n_log2 = n.bit_length() - 1
    cap = 1 << 22
    out = C.create_string_buffer(cap); ln = C.c_size_t(0)
    ms = np.zeros(len(STAGES), dtype=np.float64); err = C.create_string_buffer(256)
    o = _opts(options)
    rc = lib().orc_prove(_p(t), n_log2, _p(pi), _p(ac), _p(o), out, cap, C.byref(ln), _p(ms), int(keep_debug), err, len(err))
    if rc:
        raise RuntimeError(err.value.decode())
    proof = out.raw[:ln.value]
    return (proof, dict(zip(STAGES, ms.tolist()))) if want_times else proof


def verify(proof: bytes, pi, ac, options=DEFAULT_OPTIONS) -> str:
    """Returns '' when accepted, else the rejection reason."""
    err = C.create_string_buffer(256); o = _opts(options)
    rc = lib().orc_verify(proof, len(proof), _p(pi), _p(ac), _p(o), err, len(err))
    return "" if rc == 0 else (err.value.decode() or "rejected")


def _air_arrays(flat):
    """flat: dict from AirBuilder.flatten() (plain arrays: the AIR as data, no product code involved)."""
    return (np.ascontiguousarray(flat["desc"], dtype=np.uint32), np.ascontiguousarray(flat["pub"], dtype=np.uint64),
            np.ascontiguousarray(flat["consts"], dtype=np.uint64), np.ascontiguousarray(flat["code"], dtype=np.uint32),
            np.ascontiguousarray(flat["outs"], dtype=np.uint32), np.ascontiguousarray(flat["asr"], dtype=np.uint64))


def prove_air(flat, trace: np.ndarray, options=DEFAULT_OPTIONS, keep_debug=False, want_times=False):
    """Generic degree-<=2 AIR.  trace: (width, n) u64 column-major."""
    desc, pub, consts, code, outs, asr = _air_arrays(flat)
    t = np.ascontiguousarray(trace, dtype=np.uint64)
    n_log2 = t.shape[1].bit_length() - 1
    cap = 1 << 22
    out = C.create_string_buffer(cap); ln = C.c_size_t(0)
    ms = np.zeros(len(STAGES), dtype=np.float64); err = C.create_string_buffer(256)
    o = _opts(options)
    rc = lib().orc_prove_air(_p(t), n_log2, _p(desc), _p(pub), _p(consts), _p(code), _p(outs), _p(asr), _p(o), out, cap, C.byref(ln), _p(ms),
                             int(keep_debug), err, len(err))
    if rc:
        raise RuntimeError(err.value.decode())
    proof = out.raw[:ln.value]
    return (proof, dict(zip(STAGES, ms.tolist()))) if want_times else proof


def verify_air(proof: bytes, flat, options=DEFAULT_OPTIONS) -> str:
    desc, pub, consts, code, outs, asr = _air_arrays(flat)
    err = C.create_string_buffer(256); o = _opts(options)
    rc = lib().orc_verify_air(proof, len(proof), _p(desc), _p(pub), _p(consts), _p(code), _p(outs), _p(asr), _p(o), err, len(err))
    return "" if rc == 0 else (err.value.decode() or "rejected")


def debug_get(name: str, cap=1 << 26) -> np.ndarray:
    out = np.empty(cap, dtype=np.uint64)
    k = lib().orc_debug_get(name.encode(), _p(out), cap)
    if k < 0:
        raise KeyError(name)
    return out[:k].copy()


# ---- synthetic inputs of SURVEY.md §8(d): SplitMix64(seed "XFGSTARK" + proof_index) ----
SEED = 0x584647535441524B


def splitmix64(state):
    mask = (1 << 64) - 1
    while True:
        state = (state + 0x9E3779B97F4A7C15) & mask
        z = state
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & mask
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & mask
        yield z ^ (z >> 31)


def synthetic_inputs(index=0):
    g = splitmix64(SEED + index)
    raw = b"".join(next(g).to_bytes(8, "little") for _ in range(11))
    txp, rcpt, secret = raw[:32], raw[32:52], raw[56:88]
    return dict(burn=8_000_000, mint=8_000_000, tx_prefix_hash=txp, recipient=rcpt, secret=secret,
                network_id=4, target_chain_id=42161, version=1)


def synthetic_case(n, index=0):
    """-> (trace (7,n), pub_inputs, consts) of the normalised BurnMintAir for proof `index`."""
    s = synthetic_inputs(index)
    pi, ac, _ = pack_inputs(s["burn"], s["mint"], s["tx_prefix_hash"], s["recipient"], s["secret"], s["network_id"],
                            s["target_chain_id"], s["version"])
    return build_trace(pi, ac, n), pi, ac
