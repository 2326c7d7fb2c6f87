"""ctypes binding of tests/host_emul/libgo_emul.so: the product's general-options pipeline (launch sequence + per-thread kernel bodies of
xfg-stark_b200/csrc/general_*.cuh) executed on the host.  TEST INFRASTRUCTURE ONLY - the product library never loads it."""
import ctypes as C
import os
import subprocess

import numpy as np

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "host_emul")
_SO = os.environ.get("GO_EMUL_LIB") or os.path.join(_DIR, "libgo_emul.so")
_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_lib = None


def build():
    if os.environ.get("GO_EMUL_LIB"):
        return _SO
    csrc = os.path.join(_ROOT, "xfg-stark_b200", "csrc")
    srcs = [os.path.join(_DIR, "go_emul.cpp")] + [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cuh", ".hpp"))]
    srcs += [os.path.join(_ROOT, "include", "xfg_stark.h"), os.path.join(_ROOT, "include", "xfg", "spec.h")]
    if os.path.exists(_SO) and all(os.path.getmtime(_SO) >= os.path.getmtime(s) for s in srcs):
        return _SO
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-Wno-unknown-pragmas", "-I", cuda_inc, "-o", _SO, os.path.join(_DIR, "go_emul.cpp")])
    return _SO


def build_asan():
    """the same harness under AddressSanitizer, with poisoned guard zones between the workspace regions; None when g++ has no libasan"""
    so = os.path.join(_DIR, "libgo_emul_asan.so")
    asan = subprocess.run(["g++", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    if not os.path.isabs(asan) or not os.path.exists(asan):
        return None, None
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    src = os.path.join(_DIR, "go_emul.cpp")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(_SO) if os.path.exists(_SO) else 0):
        subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-shared", "-fsanitize=address", "-fno-omit-frame-pointer", "-Wno-unknown-pragmas",
                               "-I", cuda_inc, "-o", so, src])
    return so, asan


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        vp, sz, u32, cp = C.c_void_p, C.c_size_t, C.c_uint32, C.c_char_p
        L.go_emul_prove_air.argtypes = [vp, vp, vp, vp, vp, vp, vp, u32, vp, vp, sz, vp, cp, sz]
        L.go_emul_prove_burn_mint.argtypes = [vp, vp, vp, u32, vp, C.c_int, vp, sz, vp, cp, sz]
        L.go_emul_verify_air.argtypes = [vp, vp, vp, vp, vp, vp, cp, sz, vp]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class EmulError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{code}: {msg}")
        self.code = code


def prove_air(flat, trace, options):
    desc = np.ascontiguousarray(flat["desc"], dtype=np.uint32); pub = np.ascontiguousarray(flat["pub"], dtype=np.uint64)
    consts = np.ascontiguousarray(flat["consts"], dtype=np.uint64); code = np.ascontiguousarray(flat["code"], dtype=np.uint32)
    outs = np.ascontiguousarray(flat["outs"], dtype=np.uint32); asr = np.ascontiguousarray(flat["asr"], dtype=np.uint64)
    t = np.ascontiguousarray(trace, dtype=np.uint64); o = np.asarray(options, dtype=np.uint32)
    cap = 1 << 22; out = C.create_string_buffer(cap); ln = C.c_size_t(0); err = C.create_string_buffer(256)
    rc = lib().go_emul_prove_air(_p(desc), _p(pub), _p(consts), _p(code), _p(outs), _p(asr), _p(t), t.shape[1].bit_length() - 1, _p(o), out, cap, C.byref(ln), err, len(err))
    if rc:
        raise EmulError(rc, err.value.decode())
    return out.raw[:ln.value]


def prove_burn_mint(trace, pi, ac, options, montgomery=False):
    t = np.ascontiguousarray(trace, dtype=np.uint64); o = np.asarray(options, dtype=np.uint32)
    pi = np.ascontiguousarray(pi, dtype=np.uint64); ac = np.ascontiguousarray(ac, dtype=np.uint64)
    cap = 1 << 22; out = C.create_string_buffer(cap); ln = C.c_size_t(0); err = C.create_string_buffer(256)
    rc = lib().go_emul_prove_burn_mint(_p(pi), _p(ac), _p(t), t.shape[1].bit_length() - 1, _p(o), int(montgomery), out, cap, C.byref(ln), err, len(err))
    if rc:
        raise EmulError(rc, err.value.decode())
    return out.raw[:ln.value]


VERDICTS = ["", "ProofDeserializationError", "UnacceptableProofOptions", "InconsistentOodConstraintEvaluations", "QuerySeedProofOfWorkVerificationFailed",
            "NumberOfQueriesMismatch", "TraceQueryDoesNotMatchCommitment", "ConstraintQueryDoesNotMatchCommitment", "LayerCommitmentMismatch", "InvalidLayerFolding",
            "RemainderCommitmentMismatch", "RemainderDegreeMismatch", "InvalidRemainderFolding", "DegreeTruncation", "FailedToDrawFieldElement"]      # XFG_VERIFY_* in order


def verify_air(flat, proof, options):
    """-> '' when the product's general verifier body accepts, else the name of the failing check (winterfell's VerifierError variants)"""
    desc = np.ascontiguousarray(flat["desc"], dtype=np.uint32); pub = np.ascontiguousarray(flat["pub"], dtype=np.uint64)
    consts = np.ascontiguousarray(flat["consts"], dtype=np.uint64); code = np.ascontiguousarray(flat["code"], dtype=np.uint32)
    outs = np.ascontiguousarray(flat["outs"], dtype=np.uint32); asr = np.ascontiguousarray(flat["asr"], dtype=np.uint64)
    o = np.asarray(options, dtype=np.uint32)
    rc = lib().go_emul_verify_air(_p(desc), _p(pub), _p(consts), _p(code), _p(outs), _p(asr), bytes(proof), len(proof), _p(o))
    return VERDICTS[rc]
