"""ctypes binding of libxfgstark.so (include/xfg_stark.h) and the Python mirror of the reference's prover interface."""
import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.environ.get("XFG_LIB") or os.path.join(_HERE, "libxfgstark.so")   # XFG_LIB: debug override (A/B builds)

NUM_STAGES = 9
STAGE_NAMES = ["extend_execution_trace", "compute_execution_trace_commitment", "evaluate_constraints",
               "commit_to_constraint_evaluations", "build_deep_composition_poly", "evaluate_deep_composition_poly",
               "compute_fri_layers", "determine_query_positions", "build_proof_object"]   # winter-prover span names
EXPORTED_SYMBOLS = ["xfg_create", "xfg_destroy", "xfg_strerror", "xfg_last_error", "xfg_prove_burn_mint",
                    "xfg_prove_burn_mint_device", "xfg_prove_burn_mint_batch", "xfg_burn_mint_pack_inputs",
                    "xfg_burn_mint_build_trace", "xfg_prove_burn_mint_from_inputs", "xfg_ntt", "xfg_lde_commit",
                    "xfg_merkle_root", "xfg_eval_constraints", "xfg_fri_fold_layer", "xfg_hash_rows", "xfg_set_profiling", "xfg_set_graphs", "xfg_int_pipe_peak", "xfg_get_profile", "xfg_field_selftest", "xfg_wide_create", "xfg_wide_destroy",
                    "xfg_wide_recv_ptr", "xfg_wide_ipc_handle", "xfg_wide_open_peers", "xfg_wide_set_peer_ptrs", "xfg_wide_extend",
                    "xfg_wide_commit", "xfg_wide_read_recv", "xfg_verify_burn_mint_batch", "xfg_verify_air_batch", "xfg_verify_strerror", "xfg_create_ex", "xfg_prove_air", "xfg_prove_air_device", "xfg_prove_air_batch", "xfg_air_compile_check", "xfg_pipe_probe",
                    "xfg_prove_burn_mint_cols", "xfg_host_register", "xfg_host_unregister",
                    "xfg_debug_guard_fill", "xfg_debug_guard_check", "xfg_debug_poke_guard"]


P = 0xFFFFFFFF00000001          # the Winterfell base field modulus 2^64 - 2^32 + 1


class FieldExtension:           # winterfell::FieldExtension discriminants
    NONE = 1
    QUADRATIC = 2
    CUBIC = 3


class XfgError(RuntimeError):
    """Mirrors XfgStarkError::CryptoError("Prover error: ...") (src/burn_mint_prover.rs:124-126)."""

    def __init__(self, code, message):
        super().__init__(f"[{code}] {message}")
        self.code = code
        self.message = message


class _Options(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("num_queries", "blowup_factor", "grinding_factor", "field_extension",
                                          "fri_folding_factor", "fri_remainder_max_degree")]


class AirConsts(C.Structure):
    _fields_ = [("pub_inputs", C.c_uint64 * 12), ("txn_hash", C.c_uint64), ("recipient_hash", C.c_uint64),
                ("nullifier", C.c_uint64), ("commitment", C.c_uint64)]


class StageTimes(C.Structure):
    _fields_ = [("stage_ms", C.c_float * NUM_STAGES), ("h2d_ms", C.c_float), ("device_ms", C.c_float),
                ("total_ms", C.c_float), ("kernel_launches", C.c_uint32), ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64)]

    def as_dict(self):
        d = {n: float(self.stage_ms[i]) for i, n in enumerate(STAGE_NAMES)}
        d.update(h2d_ms=float(self.h2d_ms), device_ms=float(self.device_ms), total_ms=float(self.total_ms),
                 kernel_launches=int(self.kernel_launches), h2d_bytes=int(self.h2d_bytes), d2h_bytes=int(self.d2h_bytes))
        return d


class _AirInstr(C.Structure):
    _fields_ = [("op", C.c_uint32), ("a", C.c_uint32), ("b", C.c_uint32)]


class _Assertion(C.Structure):
    _fields_ = [("column", C.c_uint32), ("step", C.c_uint32), ("value", C.c_uint64)]


class _AirDesc(C.Structure):      # xfg_air_desc
    _fields_ = [(n, C.c_uint32) for n in ("width", "num_pub_inputs", "num_constants", "num_instr", "num_constraints", "num_assertions")] + \
               [("pub_inputs", C.c_void_p), ("constants", C.c_void_p), ("code", C.c_void_p), ("constraint_values", C.c_void_p), ("assertions", C.c_void_p)]


class VerifyTimes(C.Structure):
    _fields_ = [("host_parse_ms", C.c_float), ("h2d_ms", C.c_float), ("kernel_ms", C.c_float), ("total_ms", C.c_float),
                ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


@dataclass(frozen=True)
class ProofOptions:
    """winter_air::ProofOptions::new(num_queries, blowup_factor, grinding_factor, field_extension, fri_folding_factor,
    fri_remainder_max_degree); defaults are the reference's (src/burn_mint_prover.rs:28-35)."""
    num_queries: int = 42
    blowup_factor: int = 8
    grinding_factor: int = 4
    field_extension: int = FieldExtension.NONE
    fri_folding_factor: int = 8
    fri_remainder_max_degree: int = 31

    def _c(self):
        return _Options(self.num_queries, self.blowup_factor, self.grinding_factor, self.field_extension,
                        self.fri_folding_factor, self.fri_remainder_max_degree)

    def as_tuple(self):
        return (self.num_queries, self.blowup_factor, self.grinding_factor, self.field_extension,
                self.fri_folding_factor, self.fri_remainder_max_degree)


def library_path():
    return _SO


_lib = None


def load_library():
    """Loads libxfgstark.so (no GPU needed to load it).  Fails loudly if the extension was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_SO):
        raise XfgError(-1, f"{_SO} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(make -C xfg-stark_b200/csrc); there is no CPU fallback")
    L = C.CDLL(_SO)
    vp, u32, u64, sz, i = C.c_void_p, C.c_uint32, C.c_uint64, C.c_size_t, C.c_int
    L.xfg_create.argtypes = [i, u32, u32, C.POINTER(vp)]
    L.xfg_create_ex.argtypes = [i, u32, u32, u32, C.POINTER(vp)]
    L.xfg_destroy.argtypes = [vp]; L.xfg_destroy.restype = None
    L.xfg_strerror.argtypes = [i]; L.xfg_strerror.restype = C.c_char_p
    L.xfg_last_error.argtypes = [vp]; L.xfg_last_error.restype = C.c_char_p
    prove_tail = [vp, sz, C.POINTER(sz), C.POINTER(StageTimes)]
    L.xfg_prove_burn_mint.argtypes = [vp, vp, u32, C.POINTER(AirConsts), C.POINTER(_Options)] + prove_tail
    L.xfg_prove_burn_mint_device.argtypes = [vp, vp, u32, C.POINTER(AirConsts), C.POINTER(_Options)] + prove_tail
    L.xfg_prove_burn_mint_cols.argtypes = [vp, vp, u32, u32, C.POINTER(AirConsts), C.POINTER(_Options)] + prove_tail
    L.xfg_host_register.argtypes = [vp, vp, sz]; L.xfg_host_unregister.argtypes = [vp, vp]
    L.xfg_prove_air.argtypes = [vp, C.POINTER(_AirDesc), vp, u32, C.POINTER(_Options)] + prove_tail
    L.xfg_prove_air_device.argtypes = [vp, C.POINTER(_AirDesc), vp, u32, C.POINTER(_Options)] + prove_tail
    L.xfg_air_compile_check.argtypes = [C.POINTER(_AirDesc), u32, C.POINTER(u32), C.POINTER(u32), C.POINTER(u32), vp, vp, vp]
    L.xfg_prove_air_batch.argtypes = [vp, u32, vp, vp, u32, C.POINTER(_Options), vp, sz, vp, C.POINTER(C.c_float)]
    L.xfg_prove_burn_mint_batch.argtypes = [vp, u32, vp, u32, C.POINTER(AirConsts), C.POINTER(_Options), vp, sz, vp, C.POINTER(C.c_float)]
    inputs = [u64, u64, vp, vp, sz, vp, sz, u32, u32, u32]
    L.xfg_burn_mint_pack_inputs.argtypes = [vp] + inputs + [C.POINTER(AirConsts)]
    L.xfg_burn_mint_build_trace.argtypes = [C.POINTER(AirConsts), u32, vp]
    L.xfg_prove_burn_mint_from_inputs.argtypes = [vp] + inputs + [u32, C.POINTER(_Options)] + prove_tail
    L.xfg_verify_burn_mint_batch.argtypes = [vp, u32, vp, vp, C.POINTER(AirConsts), C.POINTER(_Options), vp, C.POINTER(VerifyTimes)]
    L.xfg_verify_air_batch.argtypes = [vp, u32, vp, vp, vp, C.POINTER(_Options), vp, C.POINTER(VerifyTimes)]
    L.xfg_verify_strerror.argtypes = [i]; L.xfg_verify_strerror.restype = C.c_char_p
    L.xfg_ntt.argtypes = [vp, vp, u32, u32, i]
    L.xfg_lde_commit.argtypes = [vp, vp, u32, u32, vp, vp]
    L.xfg_merkle_root.argtypes = [vp, vp, sz, vp, vp]
    L.xfg_eval_constraints.argtypes = [vp, vp, u32, C.POINTER(AirConsts), u32, vp, vp]
    L.xfg_fri_fold_layer.argtypes = [vp, vp, u32, u32, vp, vp]
    L.xfg_hash_rows.argtypes = [vp, vp, sz, u32, vp]
    L.xfg_set_profiling.argtypes = [vp, i]
    L.xfg_set_graphs.argtypes = [vp, i]
    L.xfg_int_pipe_peak.argtypes = [vp, C.POINTER(C.c_double)]
    L.xfg_pipe_probe.argtypes = [vp, i, C.POINTER(C.c_double)]
    L.xfg_wide_create.argtypes = [vp, u32, u32, u32, u32, C.POINTER(vp)]
    L.xfg_wide_destroy.argtypes = [vp]; L.xfg_wide_destroy.restype = None
    L.xfg_wide_recv_ptr.argtypes = [vp]; L.xfg_wide_recv_ptr.restype = vp
    L.xfg_wide_ipc_handle.argtypes = [vp, vp]
    L.xfg_wide_open_peers.argtypes = [vp, vp]
    L.xfg_wide_set_peer_ptrs.argtypes = [vp, vp]
    L.xfg_wide_extend.argtypes = [vp, vp, C.POINTER(C.c_float)]
    L.xfg_wide_commit.argtypes = [vp, vp, C.POINTER(C.c_float)]
    L.xfg_wide_read_recv.argtypes = [vp, vp]
    L.xfg_field_selftest.argtypes = [vp, u32, vp, vp, sz, vp]
    L.xfg_get_profile.argtypes = [vp, u32, C.POINTER(u32), vp, vp, vp]
    _lib = L
    return L


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


MAX_PROOF_BYTES = 1 << 20


def pack_inputs(burn_amount, mint_amount, tx_prefix_hash, recipient_address, secret, network_id, target_chain_id, commitment_version, _ctx=None):
    """validate_inputs + public-input packing + Keccak scalars (src/burn_mint_prover.rs:62-107, src/burn_mint_air.rs:124-202).  Host only."""
    L = load_library(); air = AirConsts()
    h = _ctx._h if _ctx is not None else None
    rc = L.xfg_burn_mint_pack_inputs(h, burn_amount, mint_amount, bytes(tx_prefix_hash), bytes(recipient_address), len(recipient_address),
                                     bytes(secret), len(secret), network_id, target_chain_id, commitment_version, C.byref(air))
    if rc:
        msg = (L.xfg_last_error(h) or b"").decode() if h else ""
        raise XfgError(rc, msg or L.xfg_strerror(rc).decode())
    return air


def build_trace(air, n_log2):
    """XfgBurnMintAir::build_trace (src/burn_mint_air.rs:442-476) for 2**n_log2 rows -> (7, n) uint64.  Host only."""
    t = np.empty((7, 1 << n_log2), dtype=np.uint64)
    rc = load_library().xfg_burn_mint_build_trace(C.byref(air), n_log2, _ptr(t))
    if rc:
        raise XfgError(rc, "bad arguments")
    return t


def air_compile_check(air, n_log2, cur=None, nxt=None):
    """Host-only (no GPU): validate + compile an AirBuilder as xfg_prove_air would.  -> dict(num_instr, num_slots, num_groups[, results]); results =
    the compiled program's constraint values on the frame (cur, nxt) when given.  Raises XfgError with the library's code otherwise."""
    L = load_library()
    desc, keep, w = Context._air_desc(air)
    ni, ns, ng = C.c_uint32(0), C.c_uint32(0), C.c_uint32(0)
    out = None; pc = pn = po = None
    if cur is not None:
        c = np.ascontiguousarray(cur, dtype=np.uint64); n = np.ascontiguousarray(nxt, dtype=np.uint64); out = np.zeros(desc.num_constraints, dtype=np.uint64)
        pc, pn, po = _ptr(c), _ptr(n), _ptr(out)
    rc = L.xfg_air_compile_check(C.byref(desc), n_log2, C.byref(ni), C.byref(ns), C.byref(ng), pc, pn, po)
    if rc:
        raise XfgError(rc, L.xfg_strerror(rc).decode())
    r = dict(num_instr=ni.value, num_slots=ns.value, num_groups=ng.value)
    if out is not None:
        r["results"] = out
    return r


class Context:
    """xfg_ctx: one CUDA device, `num_slots` proof workspaces sized for traces of up to 2**max_n_log2 rows."""

    def __init__(self, device=0, max_n_log2=16, num_slots=1, max_width=7):
        self._lib = load_library()
        self._h = C.c_void_p()
        rc = self._lib.xfg_create_ex(device, max_n_log2, num_slots, max_width, C.byref(self._h))
        if rc:
            raise XfgError(rc, "xfg_create failed: " + self._lib.xfg_strerror(rc).decode() + " (a CUDA device is required; no CPU fallback)")
        self.device, self.max_n_log2, self.num_slots = device, max_n_log2, num_slots
        self._out = C.create_string_buffer(MAX_PROOF_BYTES)      # reused output buffer (a fresh 1 MiB zero-filled buffer per call costs ~40 us)

    def close(self):
        if self._h:
            self._lib.xfg_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc):
        if rc:
            raise XfgError(rc, (self._lib.xfg_last_error(self._h) or b"").decode() or self._lib.xfg_strerror(rc).decode())

    # ---- host-side mirror of XfgBurnMintProver's input half ----
    def pack_inputs(self, burn_amount, mint_amount, tx_prefix_hash, recipient_address, secret, network_id, target_chain_id, commitment_version):
        return pack_inputs(burn_amount, mint_amount, tx_prefix_hash, recipient_address, secret, network_id, target_chain_id, commitment_version, _ctx=self)

    def build_trace(self, air, n_log2):
        return build_trace(air, n_log2)

    @staticmethod
    def _trace_array(trace, width, n_log2=None):
        """(width, n) uint64, C-contiguous, n a power of two (and equal to 2^n_log2 when given); anything else would make the C side
        read past the end of the buffer, so it is rejected here with the prover's own error code."""
        t = np.ascontiguousarray(trace, dtype=np.uint64)
        if t.ndim != 2 or t.shape[0] != width or t.shape[1] < 1 or (t.shape[1] & (t.shape[1] - 1)):
            raise XfgError(1, f"trace must be a ({width}, 2^k) uint64 array, got shape {tuple(t.shape)}")
        if n_log2 is not None and t.shape[1] != (1 << n_log2):
            raise XfgError(1, "all traces of a batch must have the same length")
        return t

    # ---- whole proof ----
    def prove(self, trace, air, options=ProofOptions(), want_times=False):
        """trace: (7, n) uint64 column-major host array (numpy; a pinned torch tensor's numpy view is uploaded without staging)."""
        t = self._trace_array(trace, 7)
        return self._prove_ptr(self._lib.xfg_prove_burn_mint, _ptr(t), t.shape[1].bit_length() - 1, air, options, want_times)

    def prove_cols(self, cols, air, options=ProofOptions(), form=0, want_times=False):
        """cols: seven 1-D uint64 arrays (one per trace column, as `TraceTable::get_column` hands them out); form: 0 = canonical integers,
        1 = Montgomery form (winter-math BaseElement's in-memory representation, read without conversion)."""
        if len(cols) != 7:
            raise XfgError(1, "seven columns are required")
        arrs = [np.ascontiguousarray(c, dtype=np.uint64) for c in cols]
        n = arrs[0].shape[0]
        if any(a.ndim != 1 or a.shape[0] != n for a in arrs) or n < 1 or (n & (n - 1)):
            raise XfgError(1, "columns must be 1-D uint64 arrays of one power-of-two length")
        ptrs = (C.c_void_p * 7)(*[a.ctypes.data for a in arrs])
        out = self._out; ln = C.c_size_t(0); st = StageTimes(); o = options._c()
        self._check(self._lib.xfg_prove_burn_mint_cols(self._h, ptrs, form, n.bit_length() - 1, C.byref(air), C.byref(o), out, len(out), C.byref(ln),
                                                       C.byref(st) if want_times else None))
        proof = C.string_at(out, ln.value)
        return (proof, st.as_dict()) if want_times else proof

    def guard_fill(self):
        """paints slot 0's workspace; see guard_check"""
        self._lib.xfg_debug_guard_fill.argtypes = [C.c_void_p]
        self._check(self._lib.xfg_debug_guard_fill(self._h))

    def guard_check(self, n_log2, options=ProofOptions(), width=7):
        """-> (damaged guard words, index of the first damaged zone) after proofs of that shape since guard_fill()"""
        v = C.c_uint64(0); r = C.c_int32(-1)
        self._lib.xfg_debug_guard_check.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint64), C.POINTER(C.c_int32)]
        self._check(self._lib.xfg_debug_guard_check(self._h, n_log2, options.field_extension, width, options.fri_remainder_max_degree, C.byref(v), C.byref(r)))
        return int(v.value), int(r.value)

    def host_register(self, array):
        """page-locks a caller-owned numpy array so that traces in it upload without staging (xfg_host_register)"""
        self._check(self._lib.xfg_host_register(self._h, C.c_void_p(array.ctypes.data), array.nbytes))

    def host_unregister(self, array):
        self._check(self._lib.xfg_host_unregister(self._h, C.c_void_p(array.ctypes.data)))

    def prove_device(self, device_ptr, n_log2, air, options=ProofOptions(), want_times=False):
        """device_ptr: integer device address of a (7, n) uint64 column-major buffer on this context's device."""
        return self._prove_ptr(self._lib.xfg_prove_burn_mint_device, C.c_void_p(device_ptr), n_log2, air, options, want_times)

    def _prove_ptr(self, fn, ptr, n_log2, air, options, want_times):
        out = self._out; ln = C.c_size_t(0); st = StageTimes(); o = options._c()
        self._check(fn(self._h, ptr, n_log2, C.byref(air), C.byref(o), out, len(out), C.byref(ln), C.byref(st) if want_times else None))
        proof = C.string_at(out, ln.value)
        return (proof, st.as_dict()) if want_times else proof

    # ---- generic AIR front-end ----
    @staticmethod
    def _air_desc(air):
        """air: an AirBuilder or the dict of its flatten().  -> (xfg_air_desc, arrays kept alive, width)"""
        f = air.flatten() if hasattr(air, "flatten") else air
        d = [int(v) for v in f["desc"]]
        pub = np.ascontiguousarray(f["pub"], dtype=np.uint64); consts = np.ascontiguousarray(f["consts"], dtype=np.uint64)
        code = np.ascontiguousarray(f["code"], dtype=np.uint32); outs = np.ascontiguousarray(f["outs"], dtype=np.uint32)
        a = np.asarray(f["asr"], dtype=np.uint64).reshape(-1, 3)
        asr = (_Assertion * max(1, len(a)))(*[_Assertion(int(r[0]), int(r[1]), int(r[2])) for r in a])
        desc = _AirDesc(d[0], d[1], d[2], d[3], d[4], d[5], pub.ctypes.data, consts.ctypes.data, code.ctypes.data, outs.ctypes.data, C.addressof(asr))
        return desc, (pub, consts, code, outs, asr), d[0]

    def prove_air(self, air, trace, options=ProofOptions(), want_times=False):
        """air: AirBuilder (xfg_air_desc); trace: (width, n) uint64 column-major host array.  Replaces `air.prove(trace)` for a
        user-defined AIR (e.g. src/winterfell_air.rs:169)."""
        desc, keep, w = self._air_desc(air)
        t = self._trace_array(trace, w)
        out = self._out; ln = C.c_size_t(0); st = StageTimes(); o = options._c()
        self._check(self._lib.xfg_prove_air(self._h, C.byref(desc), _ptr(t), t.shape[1].bit_length() - 1, C.byref(o), out, len(out), C.byref(ln),
                                            C.byref(st) if want_times else None))
        proof = C.string_at(out, ln.value)
        return (proof, st.as_dict()) if want_times else proof

    def prove_air_device(self, air, device_ptr, n_log2, options=ProofOptions(), want_times=False):
        desc, keep, w = self._air_desc(air)
        out = self._out; ln = C.c_size_t(0); st = StageTimes(); o = options._c()
        self._check(self._lib.xfg_prove_air_device(self._h, C.byref(desc), C.c_void_p(device_ptr), n_log2, C.byref(o), out, len(out), C.byref(ln),
                                                   C.byref(st) if want_times else None))
        proof = C.string_at(out, ln.value)
        return (proof, st.as_dict()) if want_times else proof

    def prove_air_batch(self, airs, traces, options=ProofOptions(), out_stride=1 << 17):
        """airs: list of AirBuilder (one per proof); traces: list of (width, n) uint64 arrays of equal n.  -> (list of proof bytes, device wall ms)"""
        cnt = len(traces)
        if cnt == 0:
            return [], 0.0
        if len(airs) != cnt:
            raise XfgError(1, "one AIR description per trace is required")
        descs, keep, widths = [], [], []
        for a in airs:
            d, k, w = self._air_desc(a); descs.append(d); keep.append(k); widths.append(w)
        n_log2 = np.asarray(traces[0]).shape[-1].bit_length() - 1
        ts = [self._trace_array(t, w, n_log2) for t, w in zip(traces, widths)]
        darr = (_AirDesc * cnt)(*descs); ptrs = (C.c_void_p * cnt)(*[t.ctypes.data for t in ts])
        out = np.empty(cnt * out_stride, dtype=np.uint8); lens = np.zeros(cnt, dtype=np.uint64); ms = C.c_float(0); o = options._c()
        self._check(self._lib.xfg_prove_air_batch(self._h, cnt, darr, ptrs, n_log2, C.byref(o), _ptr(out), out_stride, _ptr(lens), C.byref(ms)))
        return [out[i * out_stride:i * out_stride + int(lens[i])].tobytes() for i in range(cnt)], float(ms.value)

    def prove_batch(self, traces, airs, options=ProofOptions(), out_stride=1 << 17):
        """traces: list of (7, n) uint64 arrays; airs: list of AirConsts.  Returns (list of proof bytes, device wall ms)."""
        cnt = len(traces)
        if cnt == 0:
            return [], 0.0
        if len(airs) != cnt:
            raise XfgError(1, "one AirConsts per trace is required")
        n_log2 = np.asarray(traces[0]).shape[-1].bit_length() - 1
        ts = [self._trace_array(t, 7, n_log2) for t in traces]
        ptrs = (C.c_void_p * cnt)(*[t.ctypes.data for t in ts])
        air_arr = (AirConsts * cnt)(*airs)
        out = np.empty(cnt * out_stride, dtype=np.uint8); lens = np.zeros(cnt, dtype=np.uint64); ms = C.c_float(0); o = options._c()
        rc = self._lib.xfg_prove_burn_mint_batch(self._h, cnt, ptrs, n_log2, air_arr, C.byref(o), _ptr(out), out_stride, _ptr(lens), C.byref(ms))
        proofs = [out[i * out_stride:i * out_stride + int(lens[i])].tobytes() for i in range(cnt)]
        if rc:      # first error of the batch; proofs with a non-zero length are complete and valid (XfgError.partial)
            try:
                self._check(rc)
            except XfgError as e:
                e.partial = proofs
                raise
        return proofs, float(ms.value)

    def verify_batch(self, proofs, airs, options=ProofOptions(), want_times=False):
        """proofs: list of proof bytes; airs: list of AirConsts (public inputs + AIR constants of each proof).
        Returns a list of rejection reasons ('' = accepted), named after winterfell's VerifierError variants."""
        cnt = len(proofs)
        if cnt == 0:
            return ([], VerifyTimes().as_dict()) if want_times else []
        bufs = [np.frombuffer(bytes(p), dtype=np.uint8) if len(p) else np.zeros(1, dtype=np.uint8) for p in proofs]
        ptrs = (C.c_void_p * cnt)(*[b.ctypes.data for b in bufs])
        lens = (C.c_size_t * cnt)(*[len(p) for p in proofs])
        air_arr = (AirConsts * cnt)(*airs)
        res = np.zeros(cnt, dtype=np.int32); vt = VerifyTimes(); o = options._c()
        self._check(self._lib.xfg_verify_burn_mint_batch(self._h, cnt, ptrs, lens, air_arr, C.byref(o), _ptr(res), C.byref(vt)))
        out = [self._lib.xfg_verify_strerror(int(r)).decode() for r in res]
        return (out, vt.as_dict()) if want_times else out

    def verify_air_batch(self, proofs, airs, options=ProofOptions(), want_times=False):
        """proofs of AIRs given as data (AirBuilder, one per proof): any ProofOptions, transition degree <= 9.  -> list of rejection reasons ('' = accepted)"""
        cnt = len(proofs)
        if cnt == 0:
            return ([], VerifyTimes().as_dict()) if want_times else []
        if len(airs) != cnt:
            raise XfgError(1, "one AIR description per proof is required")
        descs, keep = [], []
        for a in airs:
            d, k, _ = self._air_desc(a); descs.append(d); keep.append(k)
        bufs = [np.frombuffer(bytes(p), dtype=np.uint8) if len(p) else np.zeros(1, dtype=np.uint8) for p in proofs]
        ptrs = (C.c_void_p * cnt)(*[b.ctypes.data for b in bufs])
        lens = (C.c_size_t * cnt)(*[len(p) for p in proofs])
        darr = (_AirDesc * cnt)(*descs)
        res = np.zeros(cnt, dtype=np.int32); vt = VerifyTimes(); o = options._c()
        self._check(self._lib.xfg_verify_air_batch(self._h, cnt, ptrs, lens, darr, C.byref(o), _ptr(res), C.byref(vt)))
        out = [self._lib.xfg_verify_strerror(int(r)).decode() for r in res]
        return (out, vt.as_dict()) if want_times else out

    def prove_from_inputs(self, burn_amount, mint_amount, tx_prefix_hash, recipient_address, secret, network_id, target_chain_id,
                          commitment_version, n_log2=6, options=ProofOptions(), want_times=False):
        out = self._out; ln = C.c_size_t(0); st = StageTimes(); o = options._c()
        self._check(self._lib.xfg_prove_burn_mint_from_inputs(self._h, burn_amount, mint_amount, bytes(tx_prefix_hash), bytes(recipient_address),
                                                              len(recipient_address), bytes(secret), len(secret), network_id, target_chain_id,
                                                              commitment_version, n_log2, C.byref(o), out, len(out), C.byref(ln),
                                                              C.byref(st) if want_times else None))
        proof = C.string_at(out, ln.value)
        return (proof, st.as_dict()) if want_times else proof

    def int_pipe_peak(self):
        """measured 32-bit integer ALU-pipe peak of this device in Gop/s (IADD3 / LOP3 / SHF mix)"""
        g = C.c_double(0); self._check(self._lib.xfg_int_pipe_peak(self._h, C.byref(g))); return float(g.value)

    def pipe_probe(self, mode):
        """thread-level instructions per second (1e9) of instruction mix `mode` (xfg_pipe_probe)"""
        g = C.c_double(0); self._check(self._lib.xfg_pipe_probe(self._h, mode, C.byref(g))); return float(g.value)

    def set_graphs(self, on=True):
        self._check(self._lib.xfg_set_graphs(self._h, int(on)))

    def set_profiling(self, on=True):
        self._check(self._lib.xfg_set_profiling(self._h, int(on)))

    def get_profile(self):
        """-> list of (kernel family, device ms, launches) of the last proof run with want_times=True while profiling."""
        cap = 64; cnt = C.c_uint32(0); names = (C.c_char_p * cap)(); ms = (C.c_float * cap)(); ln = (C.c_uint32 * cap)()
        self._check(self._lib.xfg_get_profile(self._h, cap, C.byref(cnt), names, ms, ln))
        return [(names[i].decode(), float(ms[i]), int(ln[i])) for i in range(min(cnt.value, cap))]

    # ---- stage entry points ----
    def ntt(self, data, inverse=False):
        """data: (batch, n) uint64; forward evaluation or fft::interpolate_poly per row."""
        a = np.ascontiguousarray(data, dtype=np.uint64).copy()
        a2 = a.reshape(-1, a.shape[-1])
        self._check(self._lib.xfg_ntt(self._h, _ptr(a2), a2.shape[1].bit_length() - 1, a2.shape[0], int(inverse)))
        return a

    def lde_commit(self, cols, want_lde=True):
        c = np.ascontiguousarray(cols, dtype=np.uint64); ncols, n = c.shape
        lde = np.empty((ncols, 8 * n), dtype=np.uint64) if want_lde else None
        root = C.create_string_buffer(32)
        self._check(self._lib.xfg_lde_commit(self._h, _ptr(c), n.bit_length() - 1, ncols, _ptr(lde) if want_lde else None, root))
        return lde, root.raw

    def merkle_root(self, leaves, want_nodes=False):
        lv = np.ascontiguousarray(leaves, dtype=np.uint8); cnt = lv.shape[0]
        nodes = np.empty((cnt, 32), dtype=np.uint8) if want_nodes else None
        root = C.create_string_buffer(32)
        self._check(self._lib.xfg_merkle_root(self._h, _ptr(lv), cnt, root, _ptr(nodes) if want_nodes else None))
        return (root.raw, nodes) if want_nodes else root.raw

    def eval_constraints(self, lde, air, ext, coeffs):
        l = np.ascontiguousarray(lde, dtype=np.uint64); n = l.shape[1] // 8
        cf = np.ascontiguousarray(coeffs, dtype=np.uint64)
        out = np.empty((2 * n, ext), dtype=np.uint64)
        self._check(self._lib.xfg_eval_constraints(self._h, _ptr(l), n.bit_length() - 1, C.byref(air), ext, _ptr(cf), _ptr(out)))
        return out

    def fri_fold_layer(self, evals, alpha):
        e = np.ascontiguousarray(evals, dtype=np.uint64); nl, ext = e.shape
        a = np.ascontiguousarray(alpha, dtype=np.uint64)
        out = np.empty((nl // 8, ext), dtype=np.uint64)
        self._check(self._lib.xfg_fri_fold_layer(self._h, _ptr(e), nl.bit_length() - 1, ext, _ptr(a), _ptr(out)))
        return out

    def field_selftest(self, op, a, b):
        x = np.ascontiguousarray(a, dtype=np.uint64); y = np.ascontiguousarray(b, dtype=np.uint64)
        out = np.empty_like(x)
        self._check(self._lib.xfg_field_selftest(self._h, op, _ptr(x), _ptr(y), x.size, _ptr(out)))
        return out

    def hash_rows(self, rows):
        r = np.ascontiguousarray(rows, dtype=np.uint64); cnt, limbs = r.shape
        out = np.empty((cnt, 32), dtype=np.uint8)
        self._check(self._lib.xfg_hash_rows(self._h, _ptr(r), cnt, limbs, _ptr(out)))
        return out


class WideTrace:
    """xfg_wide: this rank's share of one wide trace (BASELINE config 5): W/G columns in, rows [r*N/G, (r+1)*N/G) of all W
    columns out (the all-to-all is fused into the last NTT pass as peer stores), then the local Merkle subtree."""

    def __init__(self, ctx, n_log2, total_cols, num_ranks, rank):
        self.ctx, self.n_log2, self.total_cols, self.num_ranks, self.rank = ctx, n_log2, total_cols, num_ranks, rank
        self._h = C.c_void_p()
        ctx._check(ctx._lib.xfg_wide_create(ctx._h, n_log2, total_cols, num_ranks, rank, C.byref(self._h)))

    def close(self):
        if self._h:
            self.ctx._lib.xfg_wide_destroy(self._h); self._h = C.c_void_p()

    def recv_ptr(self):
        return self.ctx._lib.xfg_wide_recv_ptr(self._h)

    def ipc_handle(self):
        buf = C.create_string_buffer(64); self.ctx._check(self.ctx._lib.xfg_wide_ipc_handle(self._h, buf)); return buf.raw

    def open_peers(self, handles):
        """handles: list of num_ranks 64-byte IPC handles (entry r from rank r), one process per GPU."""
        self.ctx._check(self.ctx._lib.xfg_wide_open_peers(self._h, b"".join(handles)))

    def set_peer_ptrs(self, ptrs):
        arr = (C.c_void_p * len(ptrs))(*ptrs); self.ctx._check(self.ctx._lib.xfg_wide_set_peer_ptrs(self._h, arr))

    def extend(self, device_ptr):
        ms = C.c_float(0); self.ctx._check(self.ctx._lib.xfg_wide_extend(self._h, C.c_void_p(device_ptr), C.byref(ms))); return float(ms.value)

    def commit(self):
        root = C.create_string_buffer(32); ms = C.c_float(0)
        self.ctx._check(self.ctx._lib.xfg_wide_commit(self._h, root, C.byref(ms))); return root.raw, float(ms.value)

    def read_recv(self):
        n_local = (1 << self.n_log2) // self.num_ranks
        out = np.empty((self.total_cols, 8, n_local), dtype=np.uint64)
        self.ctx._check(self.ctx._lib.xfg_wide_read_recv(self._h, _ptr(out))); return out


class XfgBurnMintProver:
    """Python mirror of XfgBurnMintProver (src/burn_mint_prover.rs:18-237) over the CUDA backend.

    `new(security_parameter)` / `with_options(security_parameter, options)` / `prove_burn_mint(8 args)` keep the reference's
    names and argument meaning; `trace_log2` selects the (normalised) trace length, 6 = the reference's 64 rows.
    """

    def __init__(self, security_parameter=128, proof_options=None, trace_log2=6, context=None, device=0):
        self.security_parameter_ = security_parameter
        self.proof_options_ = proof_options or ProofOptions()
        self.trace_log2 = trace_log2
        self.ctx = context or Context(device=device, max_n_log2=max(trace_log2, 6))

    @classmethod
    def new(cls, security_parameter=128, **kw):
        return cls(security_parameter, None, **kw)

    @classmethod
    def with_options(cls, security_parameter, proof_options, **kw):
        return cls(security_parameter, proof_options, **kw)

    def prove_burn_mint(self, burn_amount, mint_amount, tx_prefix_hash, recipient_address, secret, network_id, target_chain_id, commitment_version):
        return self.ctx.prove_from_inputs(burn_amount, mint_amount, tx_prefix_hash, recipient_address, secret, network_id, target_chain_id,
                                          commitment_version, self.trace_log2, self.proof_options_)

    @staticmethod
    def get_proof_size(proof):
        return len(proof)

    def security_parameter(self):
        return self.security_parameter_

    def proof_options(self):
        return self.proof_options_

    @staticmethod
    def xfg_to_atomic_units(xfg_amount):   # src/burn_mint_prover.rs:184-186
        return int(xfg_amount * 10_000_000.0)

    @staticmethod
    def atomic_units_to_xfg(atomic_units):  # :190-192
        return atomic_units / 10_000_000.0


class XfgBurnMintVerifier:
    """Python mirror of XfgBurnMintVerifier (src/burn_mint_verifier.rs:18-358) over the CUDA batch verifier.

    `verify_with_winterfell(proof, air)` raises XfgError with the rejection reason, as the reference maps any
    winterfell::VerifierError to XfgStarkError::CryptoError("STARK verification failed: ...") (src/burn_mint_verifier.rs:278-282);
    `verify_with_public_inputs` returns a bool.  `air` is the AirConsts of the statement (public inputs and derived constants).
    """

    def __init__(self, security_parameter=128, proof_options=None, context=None, device=0):
        self.security_parameter_ = security_parameter
        self.proof_options_ = proof_options or ProofOptions()
        self.ctx = context or Context(device=device, max_n_log2=6)

    @classmethod
    def new(cls, security_parameter=128, **kw):
        return cls(security_parameter, None, **kw)

    def security_parameter(self):
        return self.security_parameter_

    def verify_with_winterfell(self, proof, air):
        reason = self.ctx.verify_batch([proof], [air], self.proof_options_)[0]
        if reason:
            raise XfgError(0, "STARK verification failed: " + reason)

    def verify_with_public_inputs(self, proof, air):
        return self.ctx.verify_batch([proof], [air], self.proof_options_)[0] == ""


class BatchBurnMintVerifier:
    """Python mirror of BatchBurnMintVerifier (src/burn_mint_verifier.rs:371-408): one kernel launch verifies the whole batch."""

    def __init__(self, security_parameter=128, proof_options=None, context=None, device=0):
        self.verifier = XfgBurnMintVerifier(security_parameter, proof_options, context, device)

    @classmethod
    def new(cls, security_parameter=128, **kw):
        return cls(security_parameter, None, **kw)

    def verify_batch(self, proofs_and_airs):
        proofs = [p for p, _ in proofs_and_airs]; airs = [a for _, a in proofs_and_airs]
        return [r == "" for r in self.verifier.ctx.verify_batch(proofs, airs, self.verifier.proof_options_)]

    def verify_all(self, proofs_and_airs):
        return all(self.verify_batch(proofs_and_airs))
