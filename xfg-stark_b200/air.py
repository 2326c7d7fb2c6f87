"""Generic AIR front-end: describe an AIR as data and prove it with the burn-mint pipeline.

``AirBuilder`` records what the body of a Winterfell ``Air::evaluate_transition`` computes (e.g. the 4-column ``XfgBurnAir``
sketch, src/winterfell_air.rs:87-127) as a straight-line program over ``frame.current()`` / ``frame.next()``, plus the
``Assertion::single`` list of ``get_assertions`` (src/winterfell_air.rs:117-124) and the public-input elements.  ``flatten()``
gives the arrays of ``xfg_air_desc`` (include/xfg_stark.h).  Constraints may have degree <= 9 (degree <= 2: ce_blowup 2, one composition
column); the library rejects anything else.

The module also holds the example AIRs used by the tests and ``bench.py --workload air``; their traces are generated on the
host with exact Goldilocks arithmetic.
"""
import numpy as np

P = 0xFFFFFFFF00000001
OP_ADD, OP_SUB, OP_MUL = 0, 1, 2


class Expr:
    """A value of the straight-line program (an index into the value numbering of xfg_air_desc)."""
    __slots__ = ("air", "vid")

    def __init__(self, air, vid):
        self.air, self.vid = air, vid

    def _bin(self, op, other, swap=False):
        o = other if isinstance(other, Expr) else self.air.const(other)
        a, b = (o, self) if swap else (self, o)
        return self.air._emit(op, a.vid, b.vid)

    def __add__(self, o): return self._bin(OP_ADD, o)
    def __radd__(self, o): return self._bin(OP_ADD, o, True)
    def __sub__(self, o): return self._bin(OP_SUB, o)
    def __rsub__(self, o): return self._bin(OP_SUB, o, True)
    def __mul__(self, o): return self._bin(OP_MUL, o)
    def __rmul__(self, o): return self._bin(OP_MUL, o, True)


class AirBuilder:
    def __init__(self, width, pub_inputs=()):
        self.width = int(width)
        self.pub_inputs = [int(v) % P for v in pub_inputs]
        self._consts, self._const_ids = [], {}
        self._code, self._cse = [], {}
        self._outs, self._asr = [], []

    # ---- frame access (Air::evaluate_transition: frame.current()[i], frame.next()[i]) ----
    def cur(self, i):
        assert 0 <= i < self.width
        return Expr(self, ("cur", i))

    def nxt(self, i):
        assert 0 <= i < self.width
        return Expr(self, ("nxt", i))

    def const(self, v):
        v = int(v) % P
        if v not in self._const_ids:
            self._const_ids[v] = len(self._consts); self._consts.append(v)
        return Expr(self, ("const", self._const_ids[v]))

    def _emit(self, op, a, b):
        key = (op, a, b)
        if key not in self._cse:
            self._cse[key] = len(self._code); self._code.append(key)
        return Expr(self, ("instr", self._cse[key]))

    def constraint(self, expr):
        """result[j] = expr for the next j (must vanish on every step but the last)."""
        self._outs.append(expr.vid)

    def assert_single(self, column, step, value):
        """Assertion::single(column, step, value)."""
        self._asr.append((int(column), int(step), int(value) % P))

    def evaluate(self, expr, cur, nxt=None):
        """Value of `expr` on a frame given as sequences of Python ints (exact arithmetic mod p); used to generate traces."""
        memo = {}

        def val(v):
            kind, i = v
            if kind == "cur":
                return int(cur[i]) % P
            if kind == "nxt":
                return int(nxt[i]) % P
            if kind == "const":
                return self._consts[i]
            if v not in memo:
                op, a, b = self._code[i]
                x, y = val(a), val(b)
                memo[v] = (x + y) % P if op == OP_ADD else (x - y) % P if op == OP_SUB else x * y % P
            return memo[v]
        return val(expr.vid)

    # ---- xfg_air_desc arrays ----
    def flatten(self):
        w, C = self.width, len(self._consts)

        def vid(v):
            kind, i = v
            return {"cur": i, "nxt": w + i, "const": 2 * w + i, "instr": 2 * w + C + i}[kind]
        code = np.array([[op, vid(a), vid(b)] for op, a, b in self._code], dtype=np.uint32).reshape(-1, 3)
        return dict(desc=np.array([w, len(self.pub_inputs), C, len(self._code), len(self._outs), len(self._asr)], dtype=np.uint32),
                    pub=np.array(self.pub_inputs, dtype=np.uint64), consts=np.array(self._consts, dtype=np.uint64), code=code,
                    outs=np.array([vid(v) for v in self._outs], dtype=np.uint32),
                    asr=np.array(self._asr, dtype=np.uint64).reshape(-1, 3))


# ------------------------------------------------------------------------------------------------------------------
# example AIRs
# ------------------------------------------------------------------------------------------------------------------
def xfg_burn_air(commitment, nullifier, amount, network_id, n):
    """The 4-register XfgBurnAir sketch (src/winterfell_air.rs:87-127): every row is (commitment, nullifier, amount, network_id),
    constraint i is `current[i] - expected_i` (:104-113), assertions pin step 0 (:117-124).  -> (AirBuilder, trace (4, n))."""
    vals = [int(v) % P for v in (commitment, nullifier, amount, network_id)]
    air = AirBuilder(4)
    for i, v in enumerate(vals):
        air.constraint(air.cur(i) - v)
        air.assert_single(i, 0, v)
    return air, np.repeat(np.array(vals, dtype=np.uint64)[:, None], n, axis=1)


def fibonacci_air(n, a0=1, b0=1):
    """Two-register Fibonacci (winterfell's fib2 example shape): a' = a + b, b' = b + a'.  Degree 1, assertions at steps 0 and n-1."""
    air = AirBuilder(2)
    air.constraint(air.nxt(0) - (air.cur(0) + air.cur(1)))
    air.constraint(air.nxt(1) - (air.cur(1) + air.nxt(0)))
    t = np.zeros((2, n), dtype=np.uint64); a, b = a0 % P, b0 % P
    for i in range(n):
        t[0, i], t[1, i] = a, b
        a = (a + b) % P; b = (b + a) % P
    air.assert_single(0, 0, int(t[0, 0])); air.assert_single(1, 0, int(t[1, 0])); air.assert_single(1, n - 1, int(t[1, n - 1]))
    air.pub_inputs = [int(t[1, n - 1])]
    return air, t


def gl_mul_np(a, b):
    """element-wise a * b mod p on uint64 arrays (exact; 32-bit limb products, 2^64 = 2^32 - 1, 2^96 = -1 mod p)"""
    a = np.asarray(a, dtype=np.uint64); b = np.asarray(b, dtype=np.uint64)
    M = np.uint64(0xFFFFFFFF); S = np.uint64(32); PP = np.uint64(P); EPS = np.uint64(0xFFFFFFFF)
    a0, a1, b0, b1 = a & M, a >> S, b & M, b >> S
    ll, lh, hl, hh = a0 * b0, a0 * b1, a1 * b0, a1 * b1
    mid = (ll >> S) + (lh & M) + (hl & M)                     # < 3 * 2^32
    lo = (ll & M) | ((mid & M) << S)
    hi = hh + (lh >> S) + (hl >> S) + (mid >> S)              # < 2^64
    hh_, hl_ = hi >> S, hi & M
    t0 = lo - hh_
    t0 = np.where(lo < hh_, t0 - EPS, t0)                     # borrow: + p = - EPS (mod 2^64)
    t1 = hl_ * EPS
    r = t0 + t1
    r = np.where(r < t1, r + EPS, r)
    return np.where(r >= PP, r - PP, r)


def gl_add_np(a, b):
    a = np.asarray(a, dtype=np.uint64); b = np.asarray(b, dtype=np.uint64)
    s = a + b
    s = np.where(s < a, s + np.uint64(0xFFFFFFFF), s)
    return np.where(s >= np.uint64(P), s - np.uint64(P), s)


def power_map_air(width, n, degree, seed=1):
    """`width` registers with x_j' = x_j^(degree-1) * x_{(j+1) mod width} + c_j: transition constraints of actual degree `degree` (3 .. 9), so the
    composition polynomial really fills degree - 1 columns (and, from degree 4, a constraint-evaluation domain of 4 n or 8 n points).
    Assertions: every register at step 0, register 0 at step n - 1."""
    rng = np.random.default_rng(seed)
    c = rng.integers(0, P, size=width, dtype=np.uint64)
    air = AirBuilder(width)
    for j in range(width):
        pw = air.cur((j + 1) % width)
        for _ in range(degree - 1):
            pw = pw * air.cur(j)
        air.constraint(air.nxt(j) - (pw + int(c[j])))
    t = np.zeros((width, n), dtype=np.uint64)
    x = rng.integers(0, P, size=width, dtype=np.uint64)
    with np.errstate(over="ignore"):
        for i in range(n):
            t[:, i] = x
            y = np.roll(x, -1)
            for _ in range(degree - 1):
                y = gl_mul_np(y, x)
            x = gl_add_np(y, c)
    for j in range(width):
        air.assert_single(j, 0, int(t[j, 0]))
    air.assert_single(0, n - 1, int(t[0, n - 1]))
    air.pub_inputs = [int(v) for v in t[:min(width, 8), 0]] + [int(t[0, n - 1])]
    return air, t


def wide_quadratic_air(width, n, seed=1, extra_steps=()):
    """The "wide synthetic AIR" of BASELINE config 5 with actual constraints: `width` registers, register j evolves as
    x_j' = x_j * x_{(j+1) mod width} + c_j (degree 2, every register reads its neighbour).  Assertions: every register at step 0,
    register 0 at step n-1 and at each of `extra_steps` (each distinct step is one more boundary divisor)."""
    rng = np.random.default_rng(seed)
    c = rng.integers(0, P, size=width, dtype=np.uint64)
    air = AirBuilder(width)
    for j in range(width):
        air.constraint(air.nxt(j) - (air.cur(j) * air.cur((j + 1) % width) + int(c[j])))
    t = np.zeros((width, n), dtype=np.uint64)
    x = rng.integers(0, P, size=width, dtype=np.uint64)
    with np.errstate(over="ignore"):
        for i in range(n):
            t[:, i] = x
            x = gl_add_np(gl_mul_np(x, np.roll(x, -1)), c)
    for j in range(width):
        air.assert_single(j, 0, int(t[j, 0]))
    for s in (n - 1,) + tuple(extra_steps):
        air.assert_single(0, s, int(t[0, s]))
    air.pub_inputs = [int(v) for v in t[:min(width, 8), 0]] + [int(t[0, n - 1])]
    return air, t


def burn_mint_air(pub_inputs, txn, rcpt, nullifier, commitment, n, last_step=None, pad_degree=None):
    """The normalised XfgBurnMintAir (src/burn_mint_air.rs:356-377 constraints, :383-394 assertions, :54-71 public inputs) written
    with the builder: the generic pipeline must emit the same proof bytes as the hand-written burn-mint kernels."""
    pi = [int(v) for v in pub_inputs]
    air = AirBuilder(7, pi)
    c = [air.cur(i) for i in range(7)]
    air.constraint((c[0] - 8_000_000) * (c[0] - 8_000_000_000))
    if pad_degree is None:
        air.constraint(c[1] - c[0])
    else:
        # the same values written as an expression of degree `pad_degree` (c0^d - c0^d = 0): what the reference computes when constraint 1 DECLARES that
        # degree (TransitionConstraintDegree::new(d)) - more composition columns (d - 1) and, from d = 4, a larger constraint-evaluation domain
        pw = c[0]
        for _ in range(pad_degree - 1):
            pw = pw * c[0]
        air.constraint((c[1] - c[0]) + pw - pw)
    air.constraint(c[2] - int(txn))
    air.constraint(c[3] - int(rcpt))
    d = air.nxt(4) - c[4]
    air.constraint(d * (d - 1))
    air.constraint(c[5] - int(nullifier))
    air.constraint(c[6] - int(commitment))
    for col, v in enumerate([pi[0], pi[1], pi[2], pi[3], 0, int(nullifier), int(commitment)]):
        air.assert_single(col, 0, v)
    # last_step: the reference source pins the final state at step 63 whatever the trace length (src/burn_mint_air.rs:393); the normalised AIR uses n - 1
    air.assert_single(4, n - 1 if last_step is None else last_step, 3)
    return air
