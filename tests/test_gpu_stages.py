"""GPU parity tests, stage level: each CUDA kernel family is called through the C ABI (include/xfg_stark.h) on seeded
inputs and compared bit-for-bit with the CPU oracle (integer work: the bar is exact equality)."""
import numpy as np
import pytest

import orc

pytestmark = pytest.mark.gpu
P = orc.P


_BIG = []


def big_ctx():
    """2^20-row context for the sizes that use the 512..1024-point register-radix NTT tiles (created on first use)."""
    import xfg_stark_b200 as xs
    if not _BIG:
        _BIG.append(xs.Context(device=0, max_n_log2=20, num_slots=1))
    return _BIG[0]


def rand_elems(rng, shape):
    return (rng.integers(0, 1 << 63, size=shape, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=shape, dtype=np.uint64)) % np.uint64(P)


EDGE = [0, 1, 2, 0xFFFFFFFF, 1 << 32, (1 << 32) + 1, 1 << 63, P - 1, P - 2, P - (1 << 32), 0xFFFFFFFF00000000, 0xFFFFFFFEFFFFFFFF,
        P, P + 1, (1 << 64) - 1, (1 << 64) - 2, (1 << 64) - (1 << 32), 0x8000000000000001, 0x00000001FFFFFFFF, 0xFFFFFFFF]


def _operands(rng, canonical_b):
    """all pairs of edge values (incl. weak aliases >= p for `a`) + random values"""
    ea = np.array(EDGE, dtype=np.uint64)
    eb = np.array([v for v in EDGE if v < P] if canonical_b else EDGE, dtype=np.uint64)
    a = np.concatenate([np.repeat(ea, eb.size), rng.integers(0, 1 << 64, size=4096, dtype=np.uint64, endpoint=False)])
    b = np.concatenate([np.tile(eb, ea.size), rand_elems(rng, (4096,)) if canonical_b else rng.integers(0, 1 << 64, size=4096, dtype=np.uint64, endpoint=False)])
    return a, b


@pytest.mark.parametrize("op", [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10] + [100 + s for s in (0, 1, 12, 24, 31, 32, 33, 36, 48, 60, 63, 64, 65, 72, 84, 95)])
def test_field_arithmetic_exact(ctx, op):
    """device field arithmetic (incl. the weak forms used inside NTT butterflies) against Python big integers, on every pair
    of edge values: 0, 1, p-1, the non-canonical aliases p..2^64-1, values around 2^32 and 2^63."""
    rng = np.random.default_rng(op)
    canonical_b = op in (2, 3, 4, 5)
    a, b = _operands(rng, canonical_b)
    if op in (4, 5, 6):
        a = a % np.uint64(P)
    got = ctx.field_selftest(op, a, b)
    ai = [int(x) for x in a]; bi = [int(x) for x in b]
    if op in (0, 1): exp = [(x * y) % P for x, y in zip(ai, bi)]
    elif op in (2, 4): exp = [(x + y) % P for x, y in zip(ai, bi)]
    elif op in (3, 5): exp = [(x - y) % P for x, y in zip(ai, bi)]
    elif op in (6, 10): exp = [pow(x, P - 2, P) for x in ai]           # 10: inversion over weak products, any u64 operand
    elif op == 7: exp = [(x + ((y & 0xFFFFFFFF) << 32)) % P for x, y in zip(ai, bi)]
    elif op == 8: exp = [(x - ((y & 0xFFFFFFFF) << 32)) % P for x, y in zip(ai, bi)]
    elif op == 9: exp = [(37 * x * y + y * y) % P for x, y in zip(ai, bi)]
    else: exp = [(x << (op - 100)) % P for x in ai]
    assert [int(g) for g in got] == exp


@pytest.mark.parametrize("limbs", [1, 2, 7, 8, 16])
def test_hash_rows_matches_blake3(ctx, limbs):
    import blake3
    rng = np.random.default_rng(limbs)
    rows = rand_elems(rng, (257, limbs))
    rows[0] = 0; rows[1] = P - 1                      # edge values
    got = ctx.hash_rows(rows)
    for i in range(rows.shape[0]):
        assert got[i].tobytes() == blake3.blake3(rows[i].astype("<u8").tobytes()).digest()


@pytest.mark.parametrize("count", [2, 4, 16, 512, 2048, 4096, 1 << 15, 1 << 17])
def test_merkle_tree_matches_oracle(ctx, count):
    rng = np.random.default_rng(count)
    leaves = rng.integers(0, 256, size=(count, 32), dtype=np.uint8)
    root, nodes = ctx.merkle_root(leaves, want_nodes=True)
    eroot, enodes = orc.merkle(leaves)
    assert root == eroot
    assert (nodes[1:] == enodes[1:]).all()


@pytest.mark.parametrize("n_log2", [3, 4, 6, 9, 11, 12, 13, 15, 16, 17, 18, 19, 20])
@pytest.mark.parametrize("inverse", [False, True])
def test_ntt_matches_oracle(ctx, n_log2, inverse):
    rng = np.random.default_rng(100 * n_log2 + inverse)
    c = ctx if n_log2 <= 16 else big_ctx()
    data = rand_elems(rng, (3, 1 << n_log2))
    data[0, :2] = [P - 1, 0]
    data[2, :] = P - 1          # all-(p-1) column: drives sums towards the weak range
    got = c.ntt(data, inverse=inverse)
    for b in range(3):
        assert (got[b] == orc.ntt(data[b], 1, 1 if inverse else 0)).all()


@pytest.mark.parametrize("n_log2", [21, 22, 23, 24])
def test_ntt_largest_tiles_match_oracle(n_log2):
    """2^21 .. 2^24 points: the 2^11- and 2^12-point register-radix tiles (1024-thread CTAs) used by the wide trace of config 5"""
    import xfg_stark_b200 as xs
    rng = np.random.default_rng(n_log2)
    data = rand_elems(rng, (1, 1 << n_log2))
    with xs.Context(device=0, max_n_log2=n_log2, num_slots=1) as c:
        fwd = c.ntt(data)
        assert (fwd[0] == orc.ntt(data[0], 1, 0)).all()
        assert (c.ntt(fwd, inverse=True) == data).all()


def test_ntt_round_trip_full_size(ctx):
    """size-independent property at the context's largest size: interpolate(evaluate(p)) == p, and linearity."""
    rng = np.random.default_rng(7)
    a = rand_elems(rng, (2, 1 << 16))
    fa = ctx.ntt(a)
    assert (ctx.ntt(fa, inverse=True) == a).all()
    s = ((a[0].astype(object) + a[1].astype(object)) % P).astype(np.uint64)
    fs = ctx.ntt(s[None, :])[0]
    assert (fs == ((fa[0].astype(object) + fa[1].astype(object)) % P).astype(np.uint64)).all()


@pytest.mark.parametrize("n_log2,cols", [(3, 7), (6, 7), (10, 7), (12, 7), (14, 7), (9, 1), (13, 2)])
def test_lde_commit_matches_oracle(ctx, n_log2, cols):
    rng = np.random.default_rng(n_log2 * 10 + cols)
    n = 1 << n_log2
    tr = rand_elems(rng, (cols, n))
    lde, root = ctx.lde_commit(tr)
    exp = np.stack([orc.lde(orc.ntt(tr[c], 1, 1)) for c in range(cols)])
    assert (lde == exp).all()
    eroot, _ = orc.merkle(orc.hash_rows(exp))
    assert root == eroot
    # the LDE restricted to every 8th point of the un-shifted domain would be the trace; on the coset, check degree instead:
    # interpolating the LDE over the coset gives back the n coefficients followed by zeros
    c0 = orc.interpolate_offset(lde[0])
    assert (c0[:n] == orc.ntt(tr[0], 1, 1)).all() and not c0[n:].any()


def test_lde_three_pass_plan_2p24_matches_oracle():
    """2^24 points run as three passes of 2^8-point tiles (16 columns, 128-byte runs) instead of two passes of 2^12-point tiles: the
    interpolation and all eight coset transforms of one column equal the oracle's, element by element (config 5's transform size)"""
    import xfg_stark_b200 as xs
    n_log2 = 24; n = 1 << n_log2
    rng = np.random.default_rng(24)
    col = rand_elems(rng, (1, n))
    orc.set_threads(orc.max_threads())
    coef = orc.ntt(col[0], 1, 1)
    with xs.Context(device=0, max_n_log2=n_log2, num_slots=1) as c:
        lde, _ = c.lde_commit(col)
    exp = orc.lde(coef)
    orc.set_threads(1)
    assert lde.shape == (1, 8 * n)
    assert (lde[0] == exp).all()


@pytest.mark.parametrize("ext", [1, 2])
@pytest.mark.parametrize("n_log2", [3, 6, 11])
def test_eval_constraints_matches_oracle(ctx, ext, n_log2):
    import xfg_stark_b200 as xs
    n = 1 << n_log2
    s = orc.synthetic_inputs(n_log2)
    air = xs.pack_inputs(s["burn"], s["mint"], s["tx_prefix_hash"], s["recipient"], s["secret"], s["network_id"], s["target_chain_id"], s["version"])
    tr, pi, ac = orc.synthetic_case(n, n_log2)
    opts = (42, 8, 4, ext, 8, 31)
    orc.prove(tr, pi, ac, opts, keep_debug=True)
    coeffs = np.concatenate([orc.debug_get("tcoef"), orc.debug_get("bcoef")])
    exp = orc.debug_get("ce_evals").reshape(2 * n, ext)
    lde = np.stack([orc.lde(orc.ntt(tr[c], 1, 1)) for c in range(7)])
    got = ctx.eval_constraints(lde, air, ext, coeffs)
    assert (got == exp).all()


@pytest.mark.parametrize("ext", [1, 2])
@pytest.mark.parametrize("nl_log2", [6, 9, 14, 17])
def test_fri_fold_matches_oracle(ctx, ext, nl_log2):
    rng = np.random.default_rng(nl_log2 + 50 * ext)
    ev = rand_elems(rng, (1 << nl_log2, ext))
    alpha = rand_elems(rng, (ext,))
    assert (ctx.fri_fold_layer(ev, alpha) == orc.fri_fold(ev, alpha)).all()
