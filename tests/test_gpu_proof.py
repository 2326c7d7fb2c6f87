"""GPU parity tests, whole proof: proof bytes from the CUDA backend (through the C ABI) must equal the CPU oracle's bytes
on the same trace, public inputs and options, and the oracle verifier (independent recomputation) must accept them."""
import numpy as np
import pytest

import orc

pytestmark = pytest.mark.gpu


def gpu_case(xs, index, n_log2):
    s = orc.synthetic_inputs(index)
    air = xs.pack_inputs(s["burn"], s["mint"], s["tx_prefix_hash"], s["recipient"], s["secret"], s["network_id"], s["target_chain_id"], s["version"])
    return air, xs.build_trace(air, n_log2)


@pytest.mark.parametrize("ext", [1, 2])
@pytest.mark.parametrize("n_log2", [3, 4, 5, 6, 7, 9, 10, 12, 13, 16])
def test_proof_bytes_equal_oracle(ctx, n_log2, ext):
    import xfg_stark_b200 as xs
    opts = xs.ProofOptions(field_extension=ext)
    air, trace = gpu_case(xs, n_log2, n_log2)
    proof = ctx.prove(trace, air, opts)
    tr, pi, ac = orc.synthetic_case(1 << n_log2, n_log2)
    assert (tr == trace).all()
    expect = orc.prove(tr, pi, ac, opts.as_tuple())
    assert len(proof) == len(expect)
    assert proof == expect
    assert orc.verify(proof, pi, ac, opts.as_tuple()) == ""


@pytest.mark.parametrize("n_log2,ext", [(17, 1), (18, 2), (19, 1), (20, 2)])
def test_large_proofs_equal_oracle(n_log2, ext):
    """BASELINE configs 2/3 sizes (up to the 2^20-row, quadratic-extension headline): bytes equal the oracle's, verifier accepts."""
    import xfg_stark_b200 as xs
    from test_gpu_stages import big_ctx
    c = big_ctx()
    opts = xs.ProofOptions(field_extension=ext)
    air, trace = gpu_case(xs, n_log2, n_log2)
    proof = c.prove(trace, air, opts)
    tr, pi, ac = orc.synthetic_case(1 << n_log2, n_log2)
    orc.set_threads(orc.max_threads())
    expect = orc.prove(tr, pi, ac, opts.as_tuple())
    orc.set_threads(1)
    assert proof == expect
    assert orc.verify(proof, pi, ac, opts.as_tuple()) == ""


@pytest.mark.parametrize("opts_t", [(42, 8, 0, 1, 8, 31), (1, 8, 4, 2, 8, 7), (255, 8, 10, 1, 8, 255), (27, 8, 16, 2, 8, 63), (100, 8, 20, 1, 8, 15)])
def test_option_sweep(ctx, opts_t):
    import xfg_stark_b200 as xs
    opts = xs.ProofOptions(*opts_t)
    air, trace = gpu_case(xs, 11, 11)
    proof = ctx.prove(trace, air, opts)
    tr, pi, ac = orc.synthetic_case(1 << 11, 11)
    assert proof == orc.prove(tr, pi, ac, opts_t)
    assert orc.verify(proof, pi, ac, opts_t) == ""


def test_reference_entry_point_64_rows(ctx):
    """XfgBurnMintProver::prove_burn_mint (src/burn_mint_prover.rs:62-129) with the reference's 64-row trace (BASELINE config 1)."""
    import xfg_stark_b200 as xs
    s = orc.synthetic_inputs(0)
    prover = xs.XfgBurnMintProver.new(128, context=ctx)
    proof = prover.prove_burn_mint(s["burn"], s["mint"], s["tx_prefix_hash"], s["recipient"], s["secret"], s["network_id"], s["target_chain_id"], s["version"])
    tr, pi, ac = orc.synthetic_case(64, 0)
    assert proof == orc.prove(tr, pi, ac)
    assert proof[:21].hex() == "07000006000008010000" "00ffffffff2a080401081f"      # Context bytes, SURVEY.md A.12
    assert prover.get_proof_size(proof) == len(proof)
    with pytest.raises(xs.XfgError) as e:
        prover.prove_burn_mint(1000, 1000, s["tx_prefix_hash"], s["recipient"], s["secret"], 4, 42161, 1)
    assert e.value.code == 8 and "Burn amount must be exactly" in e.value.message


@pytest.mark.parametrize("n_log2,ext", [(3, 1), (6, 2), (11, 1), (17, 2)])
def test_from_inputs_builds_the_trace_on_the_device(n_log2, ext):
    """the 8-argument entry (src/burn_mint_prover.rs:62-72) uploads no trace: build_trace (src/burn_mint_air.rs:442-476) is a device kernel;
    bytes equal the oracle's proof of the host-built trace, and the only host->device bytes are the coin seed elements"""
    import xfg_stark_b200 as xs
    s = orc.synthetic_inputs(n_log2)
    opts = xs.ProofOptions(field_extension=ext)
    with xs.Context(device=0, max_n_log2=n_log2) as c:
        for _ in range(2):          # second call replays the cached CUDA graph
            proof, times = c.prove_from_inputs(s["burn"], s["mint"], s["tx_prefix_hash"], s["recipient"], s["secret"], s["network_id"], s["target_chain_id"],
                                               s["version"], n_log2=n_log2, options=opts, want_times=True)
            tr, pi, ac = orc.synthetic_case(1 << n_log2, n_log2)
            assert proof == orc.prove(tr, pi, ac, opts.as_tuple())
            assert times["h2d_bytes"] == 128 * 8 + 16 + 64        # the init block of the proof state (seed elements, flags): no trace bytes
            proof_b = c.prove_from_inputs(s["burn"], s["mint"], s["tx_prefix_hash"], s["recipient"], s["secret"], s["network_id"], s["target_chain_id"],
                                          s["version"], n_log2=n_log2, options=opts)
            assert proof_b == proof


def to_montgomery(col):
    """x -> x * 2^64 mod p: the in-memory representation of winter-math 0.8 f64::BaseElement"""
    return np.array([(int(v) << 64) % orc.P for v in col], dtype=np.uint64)


@pytest.mark.parametrize("n_log2,ext", [(5, 1), (10, 2), (17, 1)])
def test_montgomery_form_columns(n_log2, ext):
    """xfg_prove_burn_mint_cols: seven separate column buffers in Montgomery form (TraceTable::get_column memory, src/burn_mint_air.rs:475),
    pageable and registered, give the bytes of the canonical contiguous trace"""
    import xfg_stark_b200 as xs
    opts = xs.ProofOptions(field_extension=ext)
    air, trace = gpu_case(xs, 3, n_log2)
    consts = {c: (int(trace[c, 0]) << 64) % orc.P for c in range(7)}
    cols = []
    for c in range(7):
        if c == 4:
            lut = np.array([(v << 64) % orc.P for v in range(4)], dtype=np.uint64); cols.append(lut[trace[4].astype(np.int64)].copy())
        else:
            cols.append(np.full(1 << n_log2, consts[c], dtype=np.uint64))
    assert (cols[0][:3] == to_montgomery(trace[0][:3])).all() and (cols[4][-3:] == to_montgomery(trace[4][-3:])).all()
    with xs.Context(device=0, max_n_log2=n_log2) as c:
        expect = c.prove(trace, air, opts)
        assert c.prove_cols(cols, air, opts, form=1) == expect                          # pageable columns, staged by the library
        assert c.prove_cols([trace[k].copy() for k in range(7)], air, opts, form=0) == expect
        for a in cols:
            c.host_register(a)
        try:
            p1, t = c.prove_cols(cols, air, opts, form=1, want_times=True)              # DMA straight from the caller's columns
            assert p1 == expect and c.prove_cols(cols, air, opts, form=1) == expect
            assert t["h2d_bytes"] == 7 * 8 * (1 << n_log2) + 128 * 8 + 16 + 64
        finally:
            for a in cols:
                c.host_unregister(a)
        bad = cols[2].copy(); bad[7] = orc.P + 3                                         # non-canonical Montgomery word
        with pytest.raises(xs.XfgError) as e:
            c.prove_cols(cols[:2] + [bad] + cols[3:], air, opts, form=1)
        assert e.value.code == 1
        with pytest.raises(xs.XfgError):
            c.prove_cols(cols, air, opts, form=7)
    tr, pi, ac = orc.synthetic_case(1 << n_log2, 3)
    assert expect == orc.prove(tr, pi, ac, opts.as_tuple())


def test_device_resident_trace_and_times(ctx):
    import torch
    import xfg_stark_b200 as xs
    air, trace = gpu_case(xs, 1, 14)
    d = torch.from_numpy(trace.view(np.int64)).cuda()
    proof, times = ctx.prove_device(d.data_ptr(), 14, air, want_times=True)
    assert proof == ctx.prove(trace, air)
    assert times["kernel_launches"] > 20 and times["device_ms"] > 0


def test_batch_equals_single(ctx):
    import xfg_stark_b200 as xs
    cases = [gpu_case(xs, i, 10) for i in range(7)]
    proofs, ms = ctx.prove_batch([t for _, t in cases], [a for a, _ in cases])
    for i, (air, trace) in enumerate(cases):
        assert proofs[i] == ctx.prove(trace, air)
        tr, pi, ac = orc.synthetic_case(1 << 10, i)
        assert orc.verify(proofs[i], pi, ac) == ""
    assert ms > 0


def test_invalid_trace_is_rejected(ctx):
    """mirrors ProverError::UnsatisfiedTransitionConstraintError: a state jump of 2 breaks d(d-1) = 0 (src/burn_mint_air.rs:240-246)."""
    import xfg_stark_b200 as xs
    air, trace = gpu_case(xs, 2, 8)
    bad = trace.copy(); bad[4, 100] = 3
    with pytest.raises(xs.XfgError) as e:
        ctx.prove(bad, air)
    assert e.value.code == 5
    bad = trace.copy(); bad[0, 5] = orc.P          # non-canonical element
    with pytest.raises(xs.XfgError) as e:
        ctx.prove(bad, air)
    assert e.value.code == 1
    assert ctx.prove(trace, air) == orc.prove(*orc.synthetic_case(256, 2))   # the context stays usable


@pytest.mark.parametrize("n_log2", [3, 6, 11, 12, 13, 15, 16, 17])
def test_non_canonical_trace_element_is_rejected_at_every_size(n_log2):
    """the canonicity check is fused into the first load of the interpolation: every NTT kernel shape (single pass, radix-2 four-step,
    register-radix four-step, split upload) must report an element >= p, wherever it sits, and accept the largest canonical value"""
    import xfg_stark_b200 as xs
    from test_gpu_stages import big_ctx
    c = big_ctx()
    air, trace = gpu_case(xs, 1, n_log2)
    n = 1 << n_log2
    for col, row, val in ((0, 0, orc.P), (6, n - 1, (1 << 64) - 1), (3, n // 2 + 1, orc.P + 5)):
        bad = trace.copy(); bad[col, row] = val
        with pytest.raises(xs.XfgError) as e:
            c.prove(bad, air)
        assert e.value.code == 1 and "non-canonical" in e.value.message
    assert c.prove(trace, air) == orc.prove(*orc.synthetic_case(n, 1))


def test_output_buffer_too_small_and_empty_batch(ctx):
    import ctypes as C
    import xfg_stark_b200 as xs
    air, trace = gpu_case(xs, 0, 6)
    lib = xs.load_library()
    out = C.create_string_buffer(100); ln = C.c_size_t(0); o = xs.ProofOptions()._c()
    rc = lib.xfg_prove_burn_mint(ctx._h, trace.ctypes.data_as(C.c_void_p), 6, C.byref(air), C.byref(o), out, len(out), C.byref(ln), None)
    assert rc == 6 and ln.value == len(ctx.prove(trace, air))          # XFG_ERR_BUFFER_TOO_SMALL, *out_len = required size
    rc = lib.xfg_prove_burn_mint(ctx._h, None, 6, C.byref(air), C.byref(o), out, len(out), C.byref(ln), None)
    assert rc == 1                                                      # XFG_ERR_BAD_ARGS
    proofs, _ = ctx.prove_batch([], [])
    assert proofs == []


def test_stage_times_and_profile(ctx):
    import xfg_stark_b200 as xs
    air, trace = gpu_case(xs, 3, 12)
    ctx.set_profiling(True)
    proof, times = ctx.prove(trace, air, want_times=True)
    prof = ctx.get_profile()
    ctx.set_profiling(False)
    assert set(xs.STAGE_NAMES) <= set(times) and all(times[k] >= 0 for k in xs.STAGE_NAMES)
    names = [p[0] for p in prof]
    for fam in ("ntt.lde_trace", "commit_rows.trace", "constraints", "deep", "fri.fold", "transcript", "gather"):
        assert fam in names
    assert sum(p[2] for p in prof) == times["kernel_launches"]
    assert times["h2d_bytes"] >= 7 * 4096 * 8 and times["d2h_bytes"] > 0


def test_option_errors(ctx):
    import xfg_stark_b200 as xs
    air, trace = gpu_case(xs, 0, 6)
    for o, code in [((42, 8, 4, 4, 8, 31), 2), ((0, 8, 4, 1, 8, 31), 2), ((256, 8, 4, 1, 8, 31), 2), ((42, 6, 4, 1, 8, 31), 2), ((42, 256, 4, 1, 8, 31), 2),
                    ((42, 8, 33, 1, 8, 31), 2), ((42, 8, 4, 1, 8, 30), 2), ((42, 8, 4, 1, 32, 31), 2), ((30, 4, 0, 2, 16, 0), 2)]:
        with pytest.raises(xs.XfgError) as e:
            ctx.prove(trace, air, xs.ProofOptions(*o))
        assert e.value.code == code, o
    tr, pi, ac = orc.synthetic_case(64, 0)
    for o in [(42, 8, 4, 3, 8, 31), (42, 4, 4, 1, 8, 31), (42, 8, 4, 1, 4, 31)]:      # refused in round 1, served by the general-options pipeline now
        assert ctx.prove(trace, air, xs.ProofOptions(*o)) == orc.prove(tr, pi, ac, o)
    big = np.zeros((7, 1 << 17), dtype=np.uint64)
    with pytest.raises(xs.XfgError) as e:
        ctx.prove(big, air)
    assert e.value.code == 9


def test_random_cases_equal_oracle(ctx):
    """seeded sweep over trace lengths, extensions, query counts, grinding factors, remainder degrees and input sets"""
    import random
    import xfg_stark_b200 as xs
    rng = random.Random(20261018)
    for case in range(24):
        n_log2 = rng.choice([3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14])
        opts_t = (rng.randrange(1, min(255, (8 << n_log2) - 1) + 1), 8, rng.randrange(0, 13), rng.choice([1, 2]), 8, rng.choice([7, 15, 31, 63, 127]))
        index = rng.randrange(1 << 20)
        air, trace = gpu_case(xs, index, n_log2)
        tr, pi, ac = orc.synthetic_case(1 << n_log2, index)
        try:
            expect = orc.prove(tr, pi, ac, opts_t)
        except RuntimeError:
            continue
        try:
            proof = ctx.prove(trace, air, xs.ProofOptions(*opts_t))
        except xs.XfgError as e:
            assert e.code == 3, (case, n_log2, opts_t, e)       # only "unsupported options" may differ from the oracle's coverage
            continue
        assert proof == expect, (case, n_log2, opts_t)
        assert orc.verify(proof, pi, ac, opts_t) == ""


def test_two_contexts_two_devices_one_process():
    """distinct contexts are independent (include/xfg_stark.h): one process driving two GPUs gets the same bytes from both"""
    import torch
    import xfg_stark_b200 as xs
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    air, trace = gpu_case(xs, 5, 16)
    with xs.Context(device=0, max_n_log2=16) as c0, xs.Context(device=1, max_n_log2=16) as c1:
        p0 = c0.prove(trace, air); p1 = c1.prove(trace, air); p0b = c0.prove(trace, air)
    assert p0 == p1 == p0b
    assert p0 == orc.prove(*orc.synthetic_case(1 << 16, 5))


def test_batch_error_leaves_context_usable(ctx):
    """one bad proof in a batch (non-canonical public input) is reported, nothing dangles, and the context keeps working"""
    import xfg_stark_b200 as xs
    cases = [gpu_case(xs, i, 9) for i in range(5)]
    airs = [a for a, _ in cases]; traces = [t for _, t in cases]
    bad = xs.AirConsts.from_buffer_copy(bytes(airs[3])); bad.pub_inputs[9] = orc.P + 5
    with pytest.raises(xs.XfgError) as e:
        ctx.prove_batch(traces, airs[:3] + [bad] + airs[4:])
    assert e.value.code == 1
    bad_trace = traces[2].copy(); bad_trace[4, 17] = 2            # breaks the state-transition constraint of proof 2 only
    with pytest.raises(xs.XfgError) as e:
        ctx.prove_batch(traces[:2] + [bad_trace] + traces[3:], airs)
    assert e.value.code == 5
    # the failed proof has length 0, every other proof of the batch is complete and valid (out_lens contract of the header)
    part = e.value.partial
    assert len(part) == 5 and part[2] == b""
    for i in (0, 1, 3, 4):
        assert part[i] == orc.prove(*orc.synthetic_case(1 << 9, i))
    with pytest.raises(xs.XfgError):                                # malformed inputs never reach the C side
        ctx.prove(traces[0][:4], airs[0])
    with pytest.raises(xs.XfgError):
        ctx.prove_batch(traces[:2] + [traces[2][:, :256]], airs[:3])
    with pytest.raises(xs.XfgError):
        ctx.prove_batch(traces[:3], airs[:2])
    proofs, _ = ctx.prove_batch(traces, airs)
    for i, (air, trace) in enumerate(cases):
        assert proofs[i] == orc.prove(*orc.synthetic_case(1 << 9, i))


def test_maximum_size_proof_2p22_equals_oracle():
    """a 2^22-row trace (N = 2^25 LDE points, 6 FRI layers, 2^11-point NTT tiles): beyond the headline size, still byte-equal"""
    import xfg_stark_b200 as xs
    n_log2 = 22
    opts = xs.ProofOptions()
    air, trace = gpu_case(xs, 7, n_log2)
    with xs.Context(device=0, max_n_log2=n_log2, num_slots=1) as c:
        proof = c.prove(trace, air, opts)
    tr, pi, ac = orc.synthetic_case(1 << n_log2, 7)
    orc.set_threads(orc.max_threads())
    expect = orc.prove(tr, pi, ac, opts.as_tuple())
    orc.set_threads(1)
    assert proof == expect
    assert orc.verify(proof, pi, ac, opts.as_tuple()) == ""
