"""stark_oracle.py — CPU restatement of the reference's whole prove -> verify protocol.  TEST INFRASTRUCTURE ONLY
(imported by tests/ and bench.py's CPU legs; the product package never imports it).

    verify(...)      default_verify          src/verifier.rs:27-183
                     ood_constraint_evaluation / deep_composition_evaluations   src/verifier.rs:207-297
                     FriVerifier::new / verify_generic / verify_remainder        src/fri.rs:293-524
                     MerkleTreeImpl::verify / verify_rows                        src/merkle.rs:209-281,364-386
    cpu_prove(...)   default_prove           src/prover.rs:25-174 — in the REFERENCE's formulation: per-column iNTT/LDE,
                     whole-column constraint evaluation, coefficient-form DEEP quotients (synthetic division, column sum,
                     degree adjustment), apply_drp through two transforms — using the C oracle for the bulk arithmetic.
    Coin             PublicCoinImpl          src/random.rs:91-196
    parse_proof      the ark-serialize layout of Proof (src/proof.rs:43-66 and the structs it contains)

The AIR *description* (constraint expressions, degree bookkeeping) is shared with the product through
ministark_b200.air — as the reference's prover and verifier share `Air` — but every field operation, hash and
transcript step below is computed independently of the product code: big-int Python for the verifier, the C
oracle for the CPU prover.

PARITY UNPINNED: no reference proof bytes exist to compare with; conventions that live in
un-vendored crates (Fp::rand, rand::gen_range, ark-serialize) are restated from upstream knowledge.
"""
import hashlib
from collections import deque

import numpy as np

from . import oracle as orc
from . import pyspec as S

P = S.P


# ------------------------------------------------------------------ Fq helpers (elements are 3-tuples of canonical ints)
def q(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (int(v) % P, 0, 0)


def q_pow(a, e):
    return S.fq3_pow(a, e)


def q_div(a, b):
    return S.fq3_mul(a, S.fq3_inv(b))


def q_scale(a, s):
    return tuple(c * s % P for c in a)


def ser(v, lanes):
    v = q(v)
    return b"".join(int(c).to_bytes(8, "little") for c in v[:lanes])


def sha(*chunks):
    h = hashlib.sha256()
    for c in chunks:
        h.update(c)
    return h.digest()


# ------------------------------------------------------------------ public coin
class Coin:
    def __init__(self, seed, lanes):
        self.seed, self.counter, self.buf, self.lanes = seed, 0, [], lanes

    def _set(self, seed):
        self.seed, self.counter, self.buf = seed, 0, []

    def reseed_digest(self, d):
        self._set(sha(self.seed, d))

    def reseed_elements(self, vals):
        for v in vals:
            self._set(sha(self.seed, sha(ser(v, self.lanes))))

    def reseed_int(self, v):
        self._set(sha(self.seed, int(v).to_bytes(8, "big")))

    def check_pow(self, bits, nonce):
        d = sha(self.seed, int(nonce).to_bytes(8, "big"))
        z = 0
        for b in d:
            if b == 0:
                z += 8
                continue
            z += 8 - b.bit_length()
            break
        return z >= bits

    def next_u64(self):
        out = 0
        for _ in range(8):
            if not self.buf:
                self.counter += 1
                self.buf = list(sha(self.seed, self.counter.to_bytes(8, "big")))
            out = (out << 8) | self.buf.pop()
        return out

    def draw_fp(self):
        while True:
            w = self.next_u64()
            if w < P:
                return S.from_mont(w)              # the raw word is the Montgomery representation

    def draw(self):
        return tuple(self.draw_fp() for _ in range(3)) if self.lanes == 3 else (self.draw_fp(), 0, 0)

    def draw_queries(self, n, domain_size):
        zone = ((domain_size << (64 - domain_size.bit_length())) - 1) % 2**64
        out = set()
        for _ in range(n):
            while True:
                m = self.next_u64() * domain_size
                if m % 2**64 <= zone:
                    out.add(m >> 64)
                    break
        return sorted(out)


# ------------------------------------------------------------------ wire format
class _Reader:
    def __init__(self, b):
        self.b, self.i = b, 0

    def take(self, k):
        if self.i + k > len(self.b):
            raise ValueError("truncated proof")
        v = self.b[self.i:self.i + k]
        self.i += k
        return v

    def u64(self):
        return int.from_bytes(self.take(8), "little")

    def u32(self):
        return int.from_bytes(self.take(4), "little")

    def digest(self):
        if self.u64() != 32:
            raise ValueError("bad digest length")
        return self.take(32)

    def elem(self, lanes):
        v = [int.from_bytes(self.take(8), "little") for _ in range(lanes)]
        if any(c >= P for c in v):
            raise ValueError("non-canonical field element")
        return tuple(v + [0] * (3 - lanes))

    def vec(self, f):
        return [f() for _ in range(self.u64())]

    def option(self, f):
        tag = self.take(1)[0]
        if tag not in (0, 1):
            raise ValueError("bad option tag")
        return f() if tag else None

    def view(self):
        return dict(nodes=self.vec(self.digest), initial_leaves=self.vec(self.digest), sibling_leaves=self.vec(self.digest),
                    height=self.u32())


def parse_proof(b, fq_lanes):
    r = _Reader(bytes(b))
    opts = tuple(r.take(5))
    p = dict(options=opts, trace_len=r.u64(), base_root=r.digest(), ext_root=r.option(r.digest), comp_root=r.digest())
    fq = lambda: r.elem(fq_lanes)
    fp = lambda: r.elem(1)
    layer = lambda: dict(rows=r.vec(fq), view=r.view(), root=r.digest())
    p["fri_layers"] = r.vec(layer)
    p["remainder"] = r.vec(fq)
    p["pow_nonce"] = r.u64()
    p["base_values"], p["ext_values"], p["comp_values"] = r.vec(fp), r.vec(fq), r.vec(fq)
    p["base_view"], p["ext_view"], p["comp_view"] = r.view(), r.option(r.view), r.view()
    p["trace_oods"], p["comp_oods"] = r.vec(fq), r.vec(fq)
    if r.i != len(r.b):
        raise ValueError("trailing bytes after proof")
    return p


# ------------------------------------------------------------------ Merkle verification (src/merkle.rs:209-281)
class VerificationError(Exception):
    pass


def merkle_verify(root, view, indices):
    height = view["height"]
    n = 1 << height
    if any(i >= n for i in indices):
        raise VerificationError("leaf index out of bounds")
    indices = sorted(set(indices))
    siblings, nodes = deque(view["sibling_leaves"]), deque(view["nodes"])
    leaf_q = deque(zip(indices, view["initial_leaves"]))
    node_q = deque()
    while leaf_q:
        index, leaf = leaf_q.popleft()
        node_index = (n + index) >> 1
        if leaf_q and leaf_q[0][0] == index ^ 1:
            node_q.append((node_index, sha(leaf, leaf_q.popleft()[1])))
            continue
        if not siblings:
            raise VerificationError("proof is invalid")
        sib = siblings.popleft()
        node_q.append((node_index, sha(leaf, sib) if index % 2 == 0 else sha(sib, leaf)))
    if siblings:
        raise VerificationError("proof is invalid")
    while node_q:
        index, h = node_q.popleft()
        if index.bit_length() - 1 == 0:
            if node_q or h != root:
                raise VerificationError("proof is invalid")
            return
        if node_q and node_q[0][0] == index ^ 1:
            node_q.append((index >> 1, sha(h, node_q.popleft()[1])))
            continue
        if not nodes:
            raise VerificationError("proof is invalid")
        sib = nodes.popleft()
        node_q.append((index >> 1, sha(h, sib) if index % 2 == 0 else sha(sib, h)))


def verify_rows(root, row_ids, rows, lanes, view):
    """MatrixMerkleTree::verify_rows (src/merkle.rs:364-386)"""
    inst = sorted(dict(zip(row_ids, rows)).items())
    leaves = [sha(b"".join(ser(v, lanes) for v in row)) for _, row in inst]
    if view["initial_leaves"] != leaves:
        raise VerificationError("proof is invalid")
    merkle_verify(root, view, [i for i, _ in inst])


# ------------------------------------------------------------------ expression evaluation at a point (graph_eval)
def eval_expr_at(expr, x, trace_map, challenges, hints, coeffs):
    memo = {}
    order, seen, stack = [], set(), [(expr, False)]
    while stack:
        node, done = stack.pop()
        if done:
            order.append(node)
            continue
        if id(node) in seen:
            continue
        seen.add(id(node))
        stack.append((node, True))
        stack.extend((a, False) for a in node[1:] if isinstance(a, tuple) and a and isinstance(a[0], str) and id(a) not in seen)
    for nd in order:
        k = nd[0]
        if k == "x":
            v = x
        elif k == "const":
            v = tuple(nd[1])
        elif k == "chal":
            v = q(challenges[nd[1]])
        elif k == "hint":
            v = q(hints[nd[1]])
        elif k == "ccoef":
            v = q(coeffs[nd[1]])
        elif k == "trace":
            v = trace_map[(nd[1], nd[2])]
        elif k == "neg":
            v = S.fq3_sub((0, 0, 0), memo[id(nd[1])])
        elif k == "add":
            v = S.fq3_add(memo[id(nd[1])], memo[id(nd[2])])
        elif k == "mul":
            v = S.fq3_mul(memo[id(nd[1])], memo[id(nd[2])])
        elif k == "div":
            v = q_div(memo[id(nd[1])], memo[id(nd[2])])
        elif k == "pow":
            v = q_pow(memo[id(nd[1])], nd[2])
        else:
            raise ValueError(k)
        memo[id(nd)] = v
    return memo[id(expr)]


def horner_q(coeffs, x):
    acc = (0, 0, 0)
    for c in reversed(coeffs):
        acc = S.fq3_add(S.fq3_mul(acc, x), c)
    return acc


def security_level_bits(options, trace_len, fq_lanes):
    """Proof::security_level_bits (src/proof.rs:126-146); field_bits = extension degree * 64"""
    nq, blowup, grind = options[0], options[1], options[2]
    field_security = fq_lanes * 64 - (trace_len * blowup).bit_length() + 1
    fri_query_security = (blowup.bit_length() - 1) * nq + grind
    return min(field_security, fri_query_security, 128, 128)


# ------------------------------------------------------------------ default_verify
def verify(stark, proof_bytes, required_security_bits, make_air):
    """stark: the claim (AirConfig, public inputs); make_air(trace_len, options5) -> AIR description.
    Raises VerificationError, returns the VerifierChannelArtifacts as a dict."""
    cfg = stark.AirConfig
    lanes = 1 if cfg.FQ_IS_FP else 3
    nbase, next_ = cfg.NUM_BASE_COLUMNS, cfg.NUM_EXTENSION_COLUMNS
    pr = parse_proof(proof_bytes, lanes)
    options, n = pr["options"], pr["trace_len"]
    nq, beta, grind, ff, max_rem = options
    if security_level_bits(options, n, lanes) < required_security_bits:
        raise VerificationError("proof params do not satisfy security requirements")
    air = make_air(n, options)
    seed = stark.public_inputs_bytes(air.public_inputs) + n.to_bytes(8, "little") + bytes(options)
    coin = Coin(sha(seed), lanes)
    coin.reseed_digest(pr["base_root"])
    challenges = [coin.draw() for _ in range(air.num_challenges())]
    unq = (lambda v: v[0]) if lanes == 1 else (lambda v: v)
    hints = air.gen_hints([unq(c) for c in challenges])
    if pr["ext_root"] is not None:
        coin.reseed_digest(pr["ext_root"])
    comp_coeffs = [coin.draw() for _ in range(air.num_composition_constraint_coeffs())]
    coin.reseed_digest(pr["comp_root"])
    z = coin.draw()
    coin.reseed_elements(pr["trace_oods"] + pr["comp_oods"])
    trace_args = air.trace_arguments()
    if len(trace_args) != len(pr["trace_oods"]) or len(pr["comp_oods"]) != air.ce_blowup_factor:
        raise VerificationError("wrong number of out-of-domain evaluations")
    ood_map = dict(zip(trace_args, pr["trace_oods"]))
    calculated = eval_expr_at(air.composition_constraint.to_tuple(), z, ood_map, challenges, hints, comp_coeffs)
    provided = horner_q(pr["comp_oods"], z)
    if calculated != provided:
        raise VerificationError("constraint evaluations at the out-of-domain point are inconsistent")
    ex_alphas = [coin.draw() for _ in range(len(trace_args))]
    co_alphas = [coin.draw() for _ in range(air.ce_blowup_factor)]
    d_alpha, d_beta = coin.draw(), coin.draw()

    # FriVerifier::new (src/fri.rs:310-352): max_poly_degree = trace_len - 1
    N = (1 << (n - 1).bit_length()) * beta if n > 1 else beta
    fri_alphas, cw = [], N
    for i, layer in enumerate(pr["fri_layers"]):
        coin.reseed_digest(layer["root"])
        fri_alphas.append(coin.draw())
        if i != len(pr["fri_layers"]) - 1 and cw % ff:
            raise VerificationError("codeword truncation")
        cw //= ff
    coin.reseed_elements(pr["remainder"])
    if grind:
        if not coin.check_pow(grind, pr["pow_nonce"]):
            raise VerificationError("insufficient proof of work on fri commitments")
        coin.reseed_int(pr["pow_nonce"])
    lde_size = n * beta
    positions = coin.draw_queries(nq, lde_size)

    def chunks(vals, k):
        return [vals[i:i + k] for i in range(0, len(vals), k)] if k else []

    base_rows = chunks(pr["base_values"], nbase)
    ext_rows = chunks(pr["ext_values"], next_) if next_ else []
    comp_rows = chunks(pr["comp_values"], air.ce_blowup_factor)
    if len(base_rows) != len(positions) or len(comp_rows) != len(positions) or (next_ and len(ext_rows) != len(positions)):
        raise VerificationError("wrong number of queried rows")
    try:
        verify_rows(pr["base_root"], positions, base_rows, 1, pr["base_view"])
    except VerificationError:
        raise VerificationError("query does not resolve to the base trace commitment")
    if pr["ext_root"] is not None:
        try:
            verify_rows(pr["ext_root"], positions, ext_rows, lanes, pr["ext_view"])
        except VerificationError:
            raise VerificationError("query does not resolve to the extension trace commitment")
    try:
        verify_rows(pr["comp_root"], positions, comp_rows, lanes, pr["comp_view"])
    except VerificationError:
        raise VerificationError("query does not resolve to the composition trace commitment")

    # deep_composition_evaluations (src/verifier.rs:231-297)
    log_n = n.bit_length() - 1
    g = S.root_of_unity(log_n)
    g_inv = pow(g, -1, P)
    g_lde = S.root_of_unity(lde_size.bit_length() - 1)
    z_n = q_pow(z, air.ce_blowup_factor)
    deep_evals = []
    for i, pos in enumerate(positions):
        x = S.GENERATOR * pow(g_lde, S.bit_reverse_index(lde_size, pos), P) % P
        xq = (x, 0, 0)
        ev = (0, 0, 0)
        for j, ((col, off), ood) in enumerate(sorted(ood_map.items())):
            if col < nbase:
                tv = base_rows[i][col]
            elif col < nbase + next_:
                tv = ext_rows[i][col - nbase]
            else:
                raise VerificationError(f"column {col} does not exist")
            shift = pow(g if off >= 0 else g_inv, abs(off), P)
            term = q_div(S.fq3_mul(ex_alphas[j], S.fq3_sub(tv, ood)), S.fq3_sub(xq, q_scale(z, shift)))
            ev = S.fq3_add(ev, term)
        for j, value in enumerate(comp_rows[i]):
            term = q_div(S.fq3_mul(co_alphas[j], S.fq3_sub(value, pr["comp_oods"][j])), S.fq3_sub(xq, z_n))
            ev = S.fq3_add(ev, term)
        ev = S.fq3_mul(ev, S.fq3_add(d_alpha, q_scale(d_beta, x)))
        deep_evals.append(ev)

    # FriVerifier::verify_generic (src/fri.rs:354-439); the folding domain has offset 1 (domain_generator powers only)
    evaluations, pos, domain_size, gen = deep_evals, positions, N, g_lde
    num_layers = 0
    d = N
    while d > max_rem * beta:
        d //= ff
        num_layers += 1
    if num_layers != len(pr["fri_layers"]):
        raise VerificationError("wrong number of FRI layers")
    w_ff_inv = pow(S.root_of_unity(ff.bit_length() - 1), -1, P)
    for li in range(num_layers):
        layer, alpha = pr["fri_layers"][li], fri_alphas[li]
        folded = sorted(set(p_ // ff for p_ in pos))
        rows = chunks(layer["rows"], ff)
        if len(rows) != len(folded):
            raise VerificationError(f"queries do not resolve to their commitment in layer {li}")
        try:
            verify_rows(layer["root"], folded, rows, lanes, layer["view"])
        except VerificationError:
            raise VerificationError(f"queries do not resolve to their commitment in layer {li}")
        query_values = [rows[folded.index(p_ // ff)][p_ % ff] for p_ in pos]
        if evaluations != query_values:
            raise VerificationError(f"degree respecting projection is invalid for layer {li}")
        nxt = []
        for row, fp_ in zip(rows, folded):
            offset = pow(gen, S.bit_reverse_index(domain_size // ff, fp_), P)
            # bit-reverse the chunk, inverse DFT over the coset offset*<w_ff>, times ff, evaluate at alpha
            nat = [row[S.bit_reverse_index(ff, t)] for t in range(ff)]
            off_inv = pow(offset, -1, P)
            coeffs = []
            for j in range(ff):
                acc = (0, 0, 0)
                for t in range(ff):
                    acc = S.fq3_add(acc, q_scale(nat[t], pow(w_ff_inv, j * t, P)))
                # (1/ff) * offset^-j from the interpolation cancels against the reference's * ff
                coeffs.append(q_scale(acc, pow(off_inv, j, P)))
            nxt.append(horner_q(coeffs, alpha))
        evaluations, pos = nxt, folded
        gen = pow(gen, ff, P)
        domain_size //= ff
    # verify_remainder (src/fri.rs:479-512)
    rem = pr["remainder"]
    deg = len(rem) - 1
    while deg > 0 and rem[deg] == (0, 0, 0):
        deg -= 1
    if deg > domain_size // beta - 1:
        raise VerificationError("remainder degree mismatch")
    for p_, want in zip(pos, evaluations):
        xr = pow(gen, S.bit_reverse_index(domain_size, p_), P)
        if horner_q(rem, (xr, 0, 0)) != want:
            raise VerificationError("remainder is invalid")
    return dict(air_challenges=challenges, air_hints=hints, fri_alphas=fri_alphas, query_positions=positions)


# ------------------------------------------------------------------ default_prove on the CPU (reference formulation)
def _mont_vec(vals, lanes):
    out = np.empty(len(vals) * lanes, dtype=np.uint64)
    for i, v in enumerate(vals):
        v = q(v)
        for l in range(lanes):
            out[i * lanes + l] = S.to_mont(v[l])
    return out


def _canon(words, lanes):
    flat = [S.from_mont(int(w)) for w in np.asarray(words).ravel()]
    return [tuple(flat[i:i + lanes] + [0] * (3 - lanes)) for i in range(0, len(flat), lanes)]


def _merkle_prove(leaves, nodes, indices):
    """MerkleTreeImpl::prove (src/merkle.rs:149-207)"""
    n = leaves.shape[0]
    idx = sorted(set(indices))
    init, sib, path = [], [], []
    node_q, leaf_q = deque(), deque(idx)
    while leaf_q:
        i = leaf_q.popleft()
        init.append(leaves[i].tobytes())
        node_q.append((n + i) >> 1)
        if leaf_q and leaf_q[0] == i ^ 1:
            init.append(leaves[leaf_q.popleft()].tobytes())
            continue
        sib.append(leaves[i ^ 1].tobytes())
    while node_q:
        i = node_q.popleft()
        if i > 2:
            node_q.append(i >> 1)
        if node_q and node_q[0] == i ^ 1:
            node_q.popleft()
            continue
        path.append(nodes[i ^ 1].tobytes())
    return dict(nodes=path, initial_leaves=init, sibling_leaves=sib, height=n.bit_length() - 1)


def _ser_view(v):
    d = lambda x: (32).to_bytes(8, "little") + x
    vec = lambda xs: len(xs).to_bytes(8, "little") + b"".join(d(x) for x in xs)
    return vec(v["nodes"]) + vec(v["initial_leaves"]) + vec(v["sibling_leaves"]) + v["height"].to_bytes(4, "little")


def _ser_vec(vals, lanes):
    return len(vals).to_bytes(8, "little") + b"".join(ser(v, lanes) for v in vals)


def cpu_prove(stark, options5, base_trace, make_air, ext_builder=None, timings=None):
    """base_trace: (ncols, n) Montgomery words.  Returns the proof bytes."""
    import time
    t0 = time.perf_counter()

    def lap(name):
        nonlocal t0
        t = time.perf_counter()
        if timings is not None:
            timings[name] = t - t0
        t0 = t

    cfg = stark.AirConfig
    lanes = 1 if cfg.FQ_IS_FP else 3
    nbase, next_ = cfg.NUM_BASE_COLUMNS, cfg.NUM_EXTENSION_COLUMNS
    nq, beta, grind, ff, max_rem = options5
    n = base_trace.shape[1]
    log_n, log_b = n.bit_length() - 1, beta.bit_length() - 1
    N, log_N = n * beta, log_n + log_b
    air = make_air(n, options5)
    gen = orc.generator()
    seed = stark.public_inputs_bytes(air.public_inputs) + n.to_bytes(8, "little") + bytes(options5)
    coin = Coin(sha(seed), lanes)
    unq = (lambda v: v[0]) if lanes == 1 else (lambda v: v)
    lap("init_air")

    def commit(evals_or_polys, cl, is_evals):
        polys = orc.ntt(evals_or_polys, cl, log_n, inverse=True) if is_evals else evals_or_polys
        lde = orc.lde(polys, cl, log_n, log_b, gen, bitrev=True)
        leaves = orc.hash_rows(lde, cl)
        nodes = orc.merkle_nodes(leaves)
        return polys, lde, leaves, nodes

    base_polys, base_lde, base_leaves, base_nodes = commit(np.ascontiguousarray(base_trace), 1, True)
    base_root = base_nodes[1].tobytes()
    coin.reseed_digest(base_root)
    lap("base_trace_commitment")
    challenges = [coin.draw() for _ in range(air.num_challenges())]
    hints = air.gen_hints([unq(c) for c in challenges])
    ext_root = None
    ext_polys = ext_lde = None
    if next_:
        ext = ext_builder([unq(c) for c in challenges])
        ext_polys, ext_lde, ext_leaves, ext_nodes = commit(np.ascontiguousarray(ext), lanes, True)
        ext_root = ext_nodes[1].tobytes()
        coin.reseed_digest(ext_root)
    lap("extension_trace_commitment")

    # constraint evaluation over the ce domain: bit_reverse_ce_trace, eval, (bit reverse back: our copies are untouched)
    from . import eval_oracle
    ce_blowup = air.ce_blowup_factor
    log_ce = log_n + ce_blowup.bit_length() - 1
    M = n * ce_blowup
    comp_coeffs = [coin.draw() for _ in range(air.num_composition_constraint_coeffs())]
    expr = air.substitute_composition_coeffs([unq(c) for c in comp_coeffs]).to_tuple()
    base_ce = np.stack([orc.bit_reverse(np.ascontiguousarray(base_lde[c][:M]), 1, log_ce) for c in range(nbase)])
    ext_ce = np.stack([orc.bit_reverse(np.ascontiguousarray(ext_lde[c][:M * lanes]), lanes, log_ce) for c in range(next_)]) if next_ else None
    comp_evals = eval_oracle.evaluate(expr, log_ce, gen, base_ce, ext_ce, fq_lanes=lanes, challenges=[unq(c) for c in challenges],
                                      hints=hints, lde_step=ce_blowup)
    lap("constraint_eval")
    comp_poly = orc.ntt(comp_evals.reshape(1, -1), lanes, log_ce, gen, inverse=True)[0]
    comp_polys = np.ascontiguousarray(comp_poly.reshape(n, ce_blowup, lanes).transpose(1, 0, 2).reshape(ce_blowup, n * lanes))
    _, comp_lde, comp_leaves, comp_nodes = commit(comp_polys, lanes, False)
    comp_root = comp_nodes[1].tobytes()
    coin.reseed_digest(comp_root)
    lap("composition_trace_commitment")

    # DEEP (src/composer.rs)
    z = coin.draw()
    g = S.root_of_unity(log_n)
    g_inv = pow(g, -1, P)
    trace_args = air.trace_arguments()

    def zpt(off):
        return q_scale(z, pow(g if off >= 0 else g_inv, abs(off), P))

    def col_poly(col):
        return (base_polys[col], 1) if col < nbase else (ext_polys[col - nbase], lanes)

    # one Horner evaluation per trace argument / composition column, spread over the threads as the reference's
    # cfg_into_iter does (composer.rs:60-83)
    z_n = q_pow(z, ce_blowup)
    oods = _canon(orc.horner_jobs([col_poly(col) + (_mont_vec([zpt(off)], 3),) for col, off in trace_args] +
                                  [(comp_polys[j], lanes, _mont_vec([z_n], 3)) for j in range(ce_blowup)]).reshape(-1), 3)
    trace_oods, comp_oods = oods[:len(trace_args)], oods[len(trace_args):]
    if lanes == 1 and any(v[1] or v[2] for v in trace_oods + comp_oods):
        raise AssertionError("ood value left the base field")
    coin.reseed_elements(trace_oods + comp_oods)
    ex_alphas = [coin.draw() for _ in range(len(trace_args))]
    co_alphas = [coin.draw() for _ in range(ce_blowup)]
    d_alpha, d_beta = coin.draw(), coin.draw()

    def lift_col(col, cl):
        if cl == 3:
            return col
        out = np.zeros(3 * n, dtype=np.uint64)
        out[0::3] = col
        return out

    # every column lifted to Fq3 and divided by its own points, columns across the threads (composer.rs:108-158)
    quotients = np.empty((ce_blowup + nbase + next_, 3 * n), dtype=np.uint64)
    zs, cs = [], []
    for j in range(ce_blowup):
        quotients[j] = lift_col(comp_polys[j], lanes)
        zs.append(_mont_vec([z_n], 3))
        cs.append(_mont_vec([co_alphas[j]], 3))
    for col in range(nbase + next_):
        sel = [(zpt(off), a) for (c, off), a in zip(trace_args, ex_alphas) if c == col]
        cp, cl = col_poly(col)
        quotients[ce_blowup + col] = lift_col(cp, cl)
        zs.append(_mont_vec([s[0] for s in sel], 3))
        cs.append(_mont_vec([s[1] for s in sel], 3))
    orc.divide_out_points_columns(quotients, zs, cs)
    deep_poly = orc.degree_adjust(orc.sum_columns(quotients, 3), _mont_vec([d_alpha], 3), _mont_vec([d_beta], 3))
    if lanes == 1:
        assert not deep_poly[1::3].any() and not deep_poly[2::3].any()
        deep_poly = np.ascontiguousarray(deep_poly[0::3])
    deep_lde = orc.lde(deep_poly.reshape(1, -1), lanes, log_n, log_b, gen, bitrev=True)[0]
    lap("deep_composition")

    # FRI (src/fri.rs:179-249)
    layers, cur, ln = [], deep_lde, log_N
    num_layers, d = 0, N
    while d > max_rem * beta:
        d //= ff
        num_layers += 1
    log_ff = ff.bit_length() - 1
    for _ in range(num_layers):
        nrows = 1 << (ln - log_ff)
        mat = np.ascontiguousarray(cur.reshape(nrows, ff, lanes).transpose(1, 0, 2).reshape(ff, nrows * lanes))   # Matrix::from_arrays
        leaves = orc.hash_rows(mat, lanes)
        nodes = orc.merkle_nodes(leaves)
        root = nodes[1].tobytes()
        coin.reseed_digest(root)
        layers.append((cur, leaves, nodes, root, nrows))
        alpha = coin.draw()
        cur = orc.fri_apply_drp(cur, lanes, ln, log_ff, _mont_vec([alpha], 3))
        ln -= log_ff
    rem = orc.ntt(orc.bit_reverse(cur, lanes, ln).reshape(1, -1), lanes, ln, inverse=True)[0]
    rem_coeffs = _canon(rem, lanes)
    keep = (1 << ln) // beta
    assert all(c == (0, 0, 0) for c in rem_coeffs[keep:]), "remainder is not low degree"
    remainder = rem_coeffs[:keep]
    coin.reseed_elements(remainder)
    lap("fri")
    nonce = 0
    if grind:
        nonce = orc.pow_grind(coin.seed, grind)
        coin.reseed_int(nonce)
    lap("proof_of_work")
    positions = coin.draw_queries(nq, N)

    out = bytes(options5) + n.to_bytes(8, "little") + (32).to_bytes(8, "little") + base_root
    out += b"\x00" if ext_root is None else b"\x01" + (32).to_bytes(8, "little") + ext_root
    out += (32).to_bytes(8, "little") + comp_root
    out += len(layers).to_bytes(8, "little")
    folded = positions
    for evals, leaves, nodes, root, nrows in layers:
        folded = sorted(set(p_ // ff for p_ in folded))
        rows = evals.reshape(nrows, ff * lanes)[folded]
        out += _ser_vec(_canon(rows, lanes), lanes) + _ser_view(_merkle_prove(leaves, nodes, folded)) + (32).to_bytes(8, "little") + root
    out += _ser_vec(remainder, lanes) + nonce.to_bytes(8, "little")
    rows_of = lambda lde, cl: _canon(np.stack([lde[:, p_ * cl:(p_ + 1) * cl].reshape(-1) for p_ in positions]), cl)
    out += _ser_vec(rows_of(base_lde, 1), 1)
    out += _ser_vec(rows_of(ext_lde, lanes), lanes) if next_ else (0).to_bytes(8, "little")
    out += _ser_vec(rows_of(comp_lde, lanes), lanes)
    out += _ser_view(_merkle_prove(base_leaves, base_nodes, positions))
    out += (b"\x01" + _ser_view(_merkle_prove(ext_leaves, ext_nodes, positions))) if next_ else b"\x00"
    out += _ser_view(_merkle_prove(comp_leaves, comp_nodes, positions))
    out += _ser_vec(trace_oods, lanes) + _ser_vec(comp_oods, lanes)
    lap("queries")
    return out
