"""
pyspec.py — big-integer executable SPEC of the hot path.  TEST INFRASTRUCTURE ONLY.

Pure-Python (arbitrary precision ints + hashlib) statement of the *mathematical
definitions* behind the reference's hot path, written independently of both the
C oracle (oracle/gl_oracle.c) and the CUDA product so it can pin them on small
cases.  Everything here works on CANONICAL integers in [0, p); `to_mont` /
`from_mont` convert to the memory representation the reference uses
(x * 2^64 mod p, gpu/src/metal/felt_u64.h.metal:118,127).

PARITY UNPINNED: no reference golden vectors exist for this path.
"""
import hashlib

P = 2**64 - 2**32 + 1
R = 2**64
ONE_MONT = R % P                      # 4294967295          felt_u64.h.metal:118
R2_MONT = (R * R) % P                 # 18446744065119617025 felt_u64.h.metal:127
GENERATOR = 7                         # Fp::GENERATOR, coset offset (src/air.rs:42-44)
TWO_ADICITY = 32
TWO_ADIC_ROOT = pow(GENERATOR, (P - 1) >> TWO_ADICITY, P)   # 1753635133440165772
NONRESIDUE = 2                        # Fq3 = Fp[X]/(X^3 - 2)  gpu/src/fields.rs:80-83
R_INV = pow(R, -1, P)


def to_mont(x):
    return (x % P) * R % P


def from_mont(w):
    return w * R_INV % P


def root_of_unity(log_n):
    """ark-ff FftField::get_root_of_unity(2^log_n)."""
    assert 0 <= log_n <= TWO_ADICITY
    return pow(TWO_ADIC_ROOT, 1 << (TWO_ADICITY - log_n), P)


def bit_reverse_index(n, i):
    """gpu/src/utils.rs:4-11."""
    bits = n.bit_length() - 1
    r = 0
    for b in range(bits):
        r |= ((i >> b) & 1) << (bits - 1 - b)
    return r


def bit_reverse(v):
    n = len(v)
    return [v[bit_reverse_index(n, i)] for i in range(n)]


# ------------------------------------------------------------------ Fq3 ----
def fq3_add(a, b):
    return tuple((x + y) % P for x, y in zip(a, b))


def fq3_sub(a, b):
    return tuple((x - y) % P for x, y in zip(a, b))


def fq3_mul(a, b):
    prod = [0] * 5
    for i in range(3):
        for j in range(3):
            prod[i + j] += a[i] * b[j]
    return ((prod[0] + NONRESIDUE * prod[3]) % P, (prod[1] + NONRESIDUE * prod[4]) % P, prod[2] % P)


def fq3_pow(a, e):
    r = (1, 0, 0)
    while e:
        if e & 1:
            r = fq3_mul(r, a)
        a = fq3_mul(a, a)
        e >>= 1
    return r


def fq3_inv(a):
    # a^(p^3 - 2): Fermat in the field of order p^3
    return fq3_pow(a, P**3 - 2)


def lift(x, lanes):
    """element (int for lanes=1, 3-tuple for lanes=3) -> Fq3 tuple"""
    return (x, 0, 0) if lanes == 1 else tuple(x)


# ------------------------------------------------------------------ NTT ----
def dft_naive(coeffs, log_n, offset=1):
    """evals[i] = sum_j c_j (offset * g^i)^j  — the definition of
    Radix2EvaluationDomain::fft on the coset offset*<g>."""
    n = 1 << log_n
    g = root_of_unity(log_n)
    out = []
    for i in range(n):
        x = offset * pow(g, i, P) % P
        acc = 0
        for c in reversed(coeffs):
            acc = (acc * x + c) % P
        out.append(acc)
    return out


def ntt(coeffs, log_n, offset=1):
    """O(n log n) recursive evaluation, same result as dft_naive (zero-pads)."""
    n = 1 << log_n
    a = [c * pow(offset, j, P) % P for j, c in enumerate(coeffs)] + [0] * (n - len(coeffs))

    def rec(v, w):
        m = len(v)
        if m == 1:
            return v
        e = rec(v[0::2], w * w % P)
        o = rec(v[1::2], w * w % P)
        out = [0] * m
        t = 1
        for k in range(m // 2):
            x = t * o[k] % P
            out[k] = (e[k] + x) % P
            out[k + m // 2] = (e[k] - x) % P
            t = t * w % P
        return out

    return rec(a, root_of_unity(log_n))


def intt(evals, log_n, offset=1):
    """inverse of ntt: coefficients c with ntt(c) == evals."""
    n = 1 << log_n
    g_inv = pow(root_of_unity(log_n), -1, P)
    n_inv = pow(n, -1, P)
    off_inv = pow(offset, -1, P)

    def rec(v, w):
        m = len(v)
        if m == 1:
            return v
        e = rec(v[0::2], w * w % P)
        o = rec(v[1::2], w * w % P)
        out = [0] * m
        t = 1
        for k in range(m // 2):
            x = t * o[k] % P
            out[k] = (e[k] + x) % P
            out[k + m // 2] = (e[k] - x) % P
            t = t * w % P
        return out

    raw = rec(list(evals), g_inv)
    return [raw[j] * n_inv % P * pow(off_inv, j, P) % P for j in range(n)]


def lde(coeffs, log_n, log_blowup, offset=GENERATOR, bitrev=True):
    """Matrix::into_(bit_reversed_)evaluations for one column (src/matrix.rs:165-234)."""
    ev = ntt(coeffs, log_n + log_blowup, offset)
    return bit_reverse(ev) if bitrev else ev


# -------------------------------------------------------------- hashing ----
def serialize_fp(x_canon):
    """ark-serialize of Fp: into_bigint() as 8 little-endian bytes."""
    return int(x_canon).to_bytes(8, "little")


def hash_row(row, lanes):
    """Sha256HashFn::hash_elements (src/hash.rs:92-99); row: canonical elements."""
    buf = b""
    for el in row:
        if lanes == 1:
            buf += serialize_fp(el)
        else:
            buf += b"".join(serialize_fp(c) for c in el)
    return hashlib.sha256(buf).digest()


def merkle_nodes(leaves):
    """build_merkle_nodes (src/merkle.rs:485-508): heap layout, nodes[0] = default."""
    n = len(leaves)
    nodes = [bytes(32)] * n
    for i in range(n // 2):
        nodes[n // 2 + i] = hashlib.sha256(leaves[2 * i] + leaves[2 * i + 1]).digest()
    for k in range(n // 2 - 1, 0, -1):
        nodes[k] = hashlib.sha256(nodes[2 * k] + nodes[2 * k + 1]).digest()
    return nodes


# ------------------------------------------------------------------ FRI ----
def fri_fold_definition(evals_bitrev, log_n, log_ff, alpha, lanes, offset=1):
    """Next FRI codeword by its per-coset definition, mirroring the verifier
    (src/fri.rs:393-412): row k of the layer (ff consecutive bit-reversed-order
    evaluations) is interpolated over its coset and evaluated at alpha, times ff.
    Elements are ints (lanes=1) or 3-tuples (lanes=3); alpha likewise."""
    n, ff = 1 << log_n, 1 << log_ff
    m = n // ff
    g = root_of_unity(log_n)
    al = lift(alpha, lanes)
    out = []
    for k in range(m):
        chunk = [lift(e, lanes) for e in evals_bitrev[k * ff:(k + 1) * ff]]
        # natural index of chunk entry t is bitrev_n(k*ff + t) = bitrev_m(k) + m * bitrev_ff(t)
        chunk = bit_reverse(chunk)
        coset_off = offset * pow(g, bit_reverse_index(m, k) if m > 1 else 0, P) % P
        # interpolate over coset_off * <g^m> (size ff), per lane (twiddles in Fp)
        cols = list(zip(*chunk))
        coeffs = list(zip(*[intt(list(c), log_ff, coset_off) for c in cols]))
        acc = (0, 0, 0)
        for c in reversed(coeffs):
            acc = fq3_add(fq3_mul(acc, al), tuple(c))
        acc = tuple(x * ff % P for x in acc)
        out.append(acc[0] if lanes == 1 else acc)
    return out


def fri_apply_drp(evals_bitrev, log_n, log_ff, alpha, lanes, offset=1):
    """apply_drp exactly as written (src/fri.rs:526-567)."""
    n, ff = 1 << log_n, 1 << log_ff
    ev = bit_reverse([lift(e, lanes) for e in evals_bitrev])
    cols = list(zip(*ev))
    coeffs = list(zip(*[intt(list(c), log_n, offset) for c in cols]))
    al = lift(alpha, lanes)
    apow = [(1, 0, 0)]
    for _ in range(ff - 1):
        apow.append(fq3_mul(apow[-1], al))
    drp = []
    for k in range(n // ff):
        s = (0, 0, 0)
        for i in range(ff):
            c = tuple(x * ff % P for x in coeffs[k * ff + i])
            s = fq3_add(s, fq3_mul(c, apow[i]))
        drp.append(s)
    cols = list(zip(*drp))
    ev2 = list(zip(*[ntt(list(c), log_n - log_ff, pow(offset, ff, P)) for c in cols]))
    ev2 = bit_reverse([tuple(e) for e in ev2])
    return [e[0] if lanes == 1 else e for e in ev2]


# ----------------------------------------------------------------- DEEP ----
def horner(coeffs, x, lanes):
    acc = (0, 0, 0)
    for c in reversed(coeffs):
        acc = fq3_add(fq3_mul(acc, x), lift(c, lanes))
    return acc


def divide_out_points(coeffs, zs, cs):
    """sum_k c_k (P(X) - P(z_k)) / (X - z_k), coefficient form (src/utils.rs:163-175)."""
    n = len(coeffs)
    out = [(0, 0, 0)] * n
    for z, c in zip(zs, cs):
        rem = (0, 0, 0)
        q = [None] * n
        for i in range(n - 1, -1, -1):
            q[i] = rem
            rem = fq3_add(fq3_mul(rem, z), coeffs[i])
        out = [fq3_add(o, fq3_mul(c, qi)) for o, qi in zip(out, q)]
    return out
