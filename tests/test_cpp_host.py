"""CPU-only: the C++ host layer (include/ministark_host.hpp — the compiled-language mirror of the reference's Rust host
code above the C ABI) against the Python mirror: public coin, proof wire format, AIR bookkeeping of examples/fib, and the
expression compiler, whose emitted programs are EXECUTED by the big-integer interpreter of tests/test_expr_compiler.py
and compared with a direct evaluation of the same expression."""
import hashlib
import json
import os
import random
import subprocess

import numpy as np
import pytest

from ministark_b200 import channel as CH
from ministark_b200 import expr as E
from ministark_b200 import proof as PR
from ministark_b200.air import Air, ProofOptions
from ministark_b200.examples import fib
from tests_helpers_expr import direct, random_columns, run_program

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = E.P


@pytest.fixture(scope="module")
def host_test(tmp_path_factory):
    exe = tmp_path_factory.mktemp("cpp") / "host_test"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "host_test.cpp"), "-o", str(exe)])

    def run(*args, stdin=None):
        return subprocess.run([str(exe)] + [str(a) for a in args], input=stdin, capture_output=True, text=True, check=True).stdout
    return run


class Prog:
    def __init__(self, d):
        self.code = np.array(d["code"], dtype=np.uint32).reshape(-1, 4)
        self.consts = np.array(d["consts"], dtype=np.uint64).reshape(-1, 3)
        self.bindings = [(s, {2: "chal", 3: "hint", 4: "ccoef"}[k], i) for s, k, i in d["bindings"]]
        self.nregs = d["nregs"]

    def bind(self, challenges=(), hints=(), ccoefs=()):
        src = {"chal": challenges, "hint": hints, "ccoef": ccoefs}
        out = Prog.__new__(Prog)
        out.code, out.nregs, out.bindings = self.code, self.nregs, self.bindings
        out.consts = self.consts.copy()
        for slot, kind, idx in self.bindings:
            out.consts[slot] = [c * 2**64 % P for c in E._q(src[kind][idx])]
        return out


@pytest.mark.parametrize("lanes", [1, 3])
def test_public_coin(host_test, lanes):
    seed = hashlib.sha256(b"cpp coin").digest()
    got = json.loads(host_test("coin", seed.hex(), lanes))
    c = CH.PublicCoin(seed, ext=lanes == 3)
    draws = []
    for step in range(24):
        v = c.draw()
        draws.append(list(E._q(v)))
        if step % 5 == 1:
            c.reseed_with_digest(hashlib.sha256(bytes([step])).digest())
        if step % 7 == 2:
            c.reseed_with_field_elements([v, v])
        if step % 9 == 3:
            c.reseed_with_int(step * 1234567)
    assert got["draws"] == draws
    assert got["queries"] == c.draw_queries(32, 1 << 23)
    assert got["queries96"] == c.draw_queries(5, 96)
    assert got["pow"] == [int(c.verify_proof_of_work(3, n)) for n in range(1, 64)]
    assert got["seed"] == c.seed.hex()


@pytest.mark.parametrize("lanes,with_ext", [(3, 1), (1, 0)])
def test_proof_wire_format(host_test, lanes, with_ext):
    class Lcg:
        def __init__(self, s):
            self.s = s

        def next(self):
            self.s = (self.s * 6364136223846793005 + 1442695040888963407) % 2**64
            return self.s >> 11

        def fp(self):
            return self.next() % P

        def fq(self, l):
            return (self.fp(), self.fp(), self.fp()) if l == 3 else self.fp()

        def digest(self):
            return bytes(self.next() & 0xFF for _ in range(32))

    r = Lcg(12345)

    def view(k):
        nodes = [r.digest() for _ in range(k)]
        init = [r.digest() for _ in range(3)]
        sib = [r.digest() for _ in range(2)]
        return PR.MerkleView(nodes, init, sib, 17)

    base = r.digest()
    ext = r.digest() if with_ext else None
    comp = r.digest()
    layers = []
    for l in range(2):
        rows = [r.fq(lanes) for _ in range(16 >> l)]
        mv = view(5 - 5 * l)
        layers.append(PR.LayerProof(rows, mv, r.digest()))
    rem = [r.fq(lanes) for _ in range(2)]
    nonce = r.next()
    bvals = [r.fq(1) for _ in range(6)]
    evals = [r.fq(lanes) for _ in range(4)] if with_ext else []
    cvals = [r.fq(lanes) for _ in range(4)]
    bview = view(7)
    eview = view(6) if with_ext else None
    cview = view(4)
    toods = [r.fq(lanes) for _ in range(5)]
    coods = [r.fq(lanes) for _ in range(2)]
    p = PR.Proof(ProofOptions(19, 16, 20, 16, 16), 2048, base, ext, comp, PR.FriProof(layers, rem), nonce,
                 PR.Queries(bvals, evals, cvals, bview, eview, cview), toods, coods)
    assert host_test("proof", 12345, lanes, with_ext).strip() == p.to_bytes().hex()


@pytest.mark.parametrize("log_n", [4, 6, 10])
def test_fib_air_and_composition_program(host_test, log_n):
    got = json.loads(host_test("fib", log_n))
    n = 1 << log_n
    air = Air(fib.FibAirConfig, n, 5, fib.OPTIONS)
    assert got["ce_blowup"] == air.ce_blowup_factor and got["num_challenges"] == air.num_challenges()
    assert got["num_coeffs"] == air.num_composition_constraint_coeffs()
    assert [tuple(t) for t in got["trace_arguments"]] == air.trace_arguments()
    from ministark_b200.air import degree
    assert [tuple(d) for d in got["degrees"]] == [degree(c, n - 1) for c in air.constraints]
    prog = Prog(got["program"])
    assert prog.nregs <= E.MAX_REGS
    rng = random.Random(log_n)
    cols, is_q = random_columns(rng, 8, 0, n)
    hints = [rng.randrange(P)]
    cc = [rng.randrange(P) for _ in range(34)]
    bound = prog.bind(hints=hints, ccoefs=cc)
    for row in (0, 1, n - 1):
        x = rng.randrange(2, P)
        assert run_program(bound, x, cols, is_q, row, n) == direct(air.composition_constraint, x, cols, row, n, (), hints, cc)


def _postfix(e, out, seen_tokens):
    """serialise an Expr as postfix tokens (shared sub-expressions are re-emitted; the C++ side hash-conses them again)"""
    k, a = e.kind, e.args
    if k == "x":
        out.append("x")
    elif k == "const":
        out.append(("q %d %d %d" % a[0]) if a[1] else ("c %d" % a[0][0]))
    elif k in ("chal", "hint", "ccoef"):
        out.append({"chal": "ch", "hint": "h", "ccoef": "cc"}[k] + " %d" % a[0])
    elif k == "trace":
        out.append("t %d %d" % a)
    elif k == "neg":
        _postfix(a[0], out, seen_tokens)
        out.append("neg")
    elif k == "pow":
        _postfix(a[0], out, seen_tokens)
        out.append("pow %d" % a[1])
    else:
        _postfix(a[0], out, seen_tokens)
        _postfix(a[1], out, seen_tokens)
        out.append(k)


@pytest.mark.parametrize("seed", range(5))
def test_compiler_on_random_expressions(host_test, seed):
    rng = random.Random(100 + seed)
    nbase, next_, m = 3, 2, 8
    cols, is_q = random_columns(rng, nbase, next_, m)
    leaves = [E.X()] + [E.Trace(c, o) for c in range(nbase + next_) for o in (0, 1)] + [E.Constant(rng.randrange(P)) for _ in range(3)] \
        + [E.Constant(tuple(rng.randrange(P) for _ in range(3))), E.Challenge(0), E.Hint(0), E.Hint(1)]
    pool = list(leaves)
    for _ in range(14):          # kept small: the postfix form expands shared sub-expressions
        a, b = rng.choice(pool), rng.choice(pool)
        pool.append(rng.choice([lambda: a + b, lambda: a - b, lambda: a * b, lambda: a / (b + E.Constant(1)), lambda: -a,
                                lambda: a ** rng.randrange(0, 5)])())
    expr = pool[-1] + pool[-2] * pool[-3]
    toks = []
    _postfix(expr, toks, None)
    prog = Prog(json.loads(host_test("expr", nbase, 1, 3, stdin=" ".join(toks))))
    ch = [tuple(rng.randrange(P) for _ in range(3))]
    hi = [tuple(rng.randrange(P) for _ in range(3)) for _ in range(2)]
    bound = prog.bind(challenges=ch, hints=hi)
    for row in range(m):
        x = rng.randrange(1, P)
        assert run_program(bound, x, cols, is_q, row, m) == direct(expr, x, cols, row, m, ch, hi)


@pytest.mark.parametrize("which", ["hello", "burner", "echo"])
def test_brainfuck_vm_and_air(host_test, which):
    """include/ministark_examples.hpp: the C++ brainfuck VM produces the same base trace as the Python one, the AIR has
    the same degrees / blow-up / trace arguments / hints, and its compiled composition evaluates like the Python AIR's"""
    from ministark_b200.air import degree
    from ministark_b200.examples import brainfuck as bf
    if which == "hello":
        src, inp, args = bf.HELLO_WORLD, b"", ("hello",)
    elif which == "burner":
        src, inp, args = bf.cycle_burner(4, 4, 4), b"", ("burner", 4, 4, 4)
    else:
        src, inp, args = ",>,<.>.+[-].", b"hi", ("src", ",>,<.>.+[-].", "hi")
    got = json.loads(host_test("bf", *args))
    trace, out = bf.simulate(src, inp)
    n = len(trace)
    assert got["n"] == n and bytes.fromhex(got["output"]) == out
    assert got["base_sha256"] == hashlib.sha256(np.ascontiguousarray(trace.base_columns()).tobytes()).hexdigest()
    claim = bf.BrainfuckClaim(src, inp, out)
    air = Air(bf.BrainfuckAirConfig, n, claim, bf.OPTIONS)
    assert (got["ce_blowup"], got["nconstraints"], got["num_challenges"], got["num_coeffs"]) == \
        (air.ce_blowup_factor, len(air.constraints), air.num_challenges(), air.num_composition_constraint_coeffs())
    assert [tuple(t) for t in got["trace_arguments"]] == air.trace_arguments()
    assert [tuple(d) for d in got["degrees"]] == [degree(c, n - 1) for c in air.constraints]
    ch = [(i + 1, i + 2, i + 3) for i in range(11)]
    hints = air.gen_hints(ch)
    assert [tuple(h) for h in got["hints"]] == [tuple(h) for h in hints]
    prog = Prog(got["program"])
    assert prog.nregs <= E.MAX_REGS
    rng = random.Random(5)
    m = n * air.ce_blowup_factor
    pick = [0, 1, 17, m - 1, m // 2]
    nb, ne = 17, 9
    lazy = {}                   # 26 columns x 32768 rows: sampled lazily, only the cells the constraints touch

    class Col:
        def __init__(self, c, is_q):
            self.c, self.q = c, is_q

        def __getitem__(self, r):
            key = (self.c, r)
            if key not in lazy:
                lazy[key] = tuple(rng.randrange(P) for _ in range(3)) if self.q else rng.randrange(P)
            return lazy[key]

    cols = [Col(c, False) for c in range(nb)] + [Col(nb + c, True) for c in range(ne)]
    is_q = [False] * nb + [True] * ne
    cc = [tuple(rng.randrange(P) for _ in range(3)) for _ in range(air.num_composition_constraint_coeffs())]
    bound = prog.bind(challenges=ch, hints=hints, ccoefs=cc)
    for row in pick:
        x = rng.randrange(2, P)
        want = direct(air.composition_constraint, x, cols, row, m, ch, hints, cc, lde_step=air.ce_blowup_factor)
        assert run_program(bound, x, cols, is_q, row, m) == want


def test_cpp_verifier_accepts_and_rejects_like_the_python_restatement(host_test, orc):
    """include/ministark_verifier.hpp (default_verify as a C++ library): accepts the CPU prover's
    proofs for examples/fib and examples/brainfuck with the same query positions as the Python restatement, refuses a wrong
    claim, too little security, and every sampled single-bit corruption (which the Python verifier refuses too)"""
    from ministark_b200.examples import brainfuck as bf
    from oracle import stark_oracle as SO
    rng = random.Random(77)

    def cpp(kind, proof, *args):
        return host_test("verify", kind, *args, stdin=proof.hex()).strip()

    # ---- fib
    opts = (32, 4, 8, 8, 64)
    trace, last = fib.gen_trace(8 << 7)
    claim = fib.FibClaim(last)
    mk = lambda n, o: Air(claim.AirConfig, n, claim.get_public_inputs(), ProofOptions(*o))
    proof = SO.cpu_prove(claim, opts, trace.base_columns(), mk)
    art = SO.verify(claim, proof, 30, mk)
    assert cpp("fib", proof, last, 30) == "ok " + " ".join(str(p) for p in art["query_positions"])
    assert cpp("fib", proof, (last + 1) % P, 30).startswith("error: constraint evaluations at the out-of-domain point")
    assert cpp("fib", proof, last, 100).startswith("error: proof params do not satisfy security")
    for _ in range(30):
        b = bytearray(proof)
        b[rng.randrange(5, len(b))] ^= 1 << rng.randrange(8)
        with pytest.raises((SO.VerificationError, ValueError)):
            SO.verify(claim, bytes(b), 30, mk)
        assert cpp("fib", bytes(b), last, 30).startswith("error:")
    # ---- brainfuck hello world (Fq3, extension columns, ce_blowup 16, 2 FRI layers of folding factor 16)
    btrace, out = bf.simulate(bf.HELLO_WORLD)
    bclaim = bf.BrainfuckClaim(bf.HELLO_WORLD, b"", out)
    bmk = lambda n, o: Air(bclaim.AirConfig, n, bclaim, ProofOptions(*o))
    bproof = SO.cpu_prove(bclaim, (19, 16, 20, 16, 16), btrace.base_columns(), bmk, ext_builder=btrace.build_extension_columns)
    bart = SO.verify(bclaim, bproof, 96, bmk)
    assert cpp("bf", bproof, bf.HELLO_WORLD, out.hex(), 96) == "ok " + " ".join(str(p) for p in bart["query_positions"])
    assert cpp("bf", bproof, bf.HELLO_WORLD, b"Hello World?".hex(), 96).startswith("error:")
    for _ in range(12):
        b = bytearray(bproof)
        b[rng.randrange(5, len(b))] ^= 1 << rng.randrange(8)
        assert cpp("bf", bytes(b), bf.HELLO_WORLD, out.hex(), 96).startswith("error:")
