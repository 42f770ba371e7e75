"""CPU-only: the restated protocol layer (oracle/stark_oracle.py) is self-consistent — the reference-formulation CPU
prover's bytes are accepted by the restated `default_verify`, any tampering is rejected — and the product's host-side
Fiat–Shamir code (ministark_b200/channel.py, air.py) agrees with the oracle's independent implementation.
Mirrors the reference's own end-to-end check (examples/fib/main.rs:227-243: prove, then verify) and
src/merkle.rs:528-581 (prove/verify round trips)."""
import hashlib
import random

import numpy as np
import pytest

from ministark_b200 import channel as CH
from ministark_b200.air import Air, ProofOptions, blowup_factor, degree
from ministark_b200 import expr as E
from ministark_b200.examples import fib, perm
from oracle import stark_oracle as SO


def air_factory(stark):
    return lambda n, o: Air(stark.AirConfig, n, stark.get_public_inputs(), ProofOptions(*o))


@pytest.mark.parametrize("log_rows,opts", [(7, (32, 4, 8, 8, 64)), (9, (20, 8, 5, 4, 16)), (6, (10, 2, 0, 2, 8)), (8, (16, 16, 3, 16, 4))])
def test_fib_cpu_prove_verify(orc, log_rows, opts):
    trace, last = fib.gen_trace(8 << log_rows)
    claim = fib.FibClaim(last)
    proof = SO.cpu_prove(claim, opts, trace.base_columns(), air_factory(claim))
    art = SO.verify(claim, proof, 10, air_factory(claim))
    assert len(art["query_positions"]) <= opts[0] and art["query_positions"] == sorted(set(art["query_positions"]))
    # a wrong public input fails at the OOD consistency check (verifier.rs:84-86)
    bad = fib.FibClaim((last + 1) % fib.P)
    with pytest.raises(SO.VerificationError, match="out-of-domain"):
        SO.verify(bad, proof, 10, air_factory(bad))
    # insufficient security is refused up front (verifier.rs:34-36)
    with pytest.raises(SO.VerificationError, match="security"):
        SO.verify(claim, proof, 129, air_factory(claim))
    # every single-bit corruption of the proof is rejected (sampled)
    rng = random.Random(log_rows)
    for _ in range(40):
        b = bytearray(proof)
        i = rng.randrange(5, len(b))
        b[i] ^= 1 << rng.randrange(8)
        with pytest.raises((SO.VerificationError, ValueError)):
            SO.verify(claim, bytes(b), 10, air_factory(claim))


def test_fib_trace_satisfies_recurrence():
    trace, last = fib.gen_trace(1 << 12)
    cols = [[int(w) * pow(2**64, -1, fib.P) % fib.P for w in c] for c in trace.base_columns()]
    flat = [cols[k][r] for r in range(len(cols[0])) for k in range(8)]
    assert flat[0] == 1 and flat[1] == 2 and flat[-1] == last
    assert all(flat[k] == flat[k - 2] * flat[k - 1] % fib.P for k in range(2, len(flat)))


def test_invalid_trace_is_rejected_by_the_verifier(orc):
    # the prover never checks the AIR (src/debug.rs is debug-only) and every polynomial it commits to is low degree by
    # construction, so a bad trace still yields a proof — which the verifier's OOD consistency check refuses
    trace, last = fib.gen_trace(8 << 6)
    cols = trace.base_columns().copy()
    cols[3, 17] ^= np.uint64(1)
    claim = fib.FibClaim(last)
    proof = SO.cpu_prove(claim, (10, 4, 0, 8, 4), cols, air_factory(claim))
    with pytest.raises(SO.VerificationError, match="out-of-domain"):
        SO.verify(claim, proof, 10, air_factory(claim))


@pytest.mark.parametrize("log_n,opts", [(6, (12, 8, 4, 4, 8)), (8, (20, 16, 6, 16, 4))])
def test_perm_fq3_cpu_prove_verify(orc, log_n, opts):
    claim = perm.PermClaim()
    tr = perm.gen_trace(1 << log_n)
    air = air_factory(claim)(1 << log_n, opts)
    assert air.ce_blowup_factor == 4 and air.num_challenges() == 1
    proof = SO.cpu_prove(claim, opts, tr.base_columns(), air_factory(claim), ext_builder=tr.build_extension_columns)
    SO.verify(claim, proof, 10, air_factory(claim))
    rng = random.Random(log_n)
    for _ in range(25):
        b = bytearray(proof)
        b[rng.randrange(5, len(b))] ^= 1 << rng.randrange(8)
        with pytest.raises((SO.VerificationError, ValueError)):
            SO.verify(claim, bytes(b), 10, air_factory(claim))


def test_degree_rules_match_the_reference_examples():
    # examples/fib: every constraint and the composition have blowup 1 (SURVEY.md §8d config 3 note)
    n = 1 << 10
    cs = fib.FibAirConfig.constraints(n)
    assert [blowup_factor(c, n) for c in cs] == [1] * 17
    assert degree(cs[0], n - 1) == (n - 1, 1) and degree(cs[9], n - 1) == (2 * (n - 1) + 1, n)
    air = Air(fib.FibAirConfig, n, 0, fib.OPTIONS)
    assert air.ce_blowup_factor == 1 and air.num_composition_constraint_coeffs() == 34
    assert air.trace_arguments() == [(c, o) for c in range(8) for o in (0, 1)]
    # tests/constraint.rs:199-217: x*(x-1) has degree 2(n-1)
    c = E.Trace(0, 0) * (E.Trace(0, 0) - E.Constant(1))
    assert degree(c, n - 1) == (2 * (n - 1), 0)


def test_public_coin_two_implementations_agree():
    seed = hashlib.sha256(b"coin").digest()
    for ext in (False, True):
        a, b = CH.PublicCoin(seed, ext=ext), SO.Coin(seed, 3 if ext else 1)
        for step in range(40):
            va, vb = a.draw(), b.draw()
            assert SO.q(va) == vb
            if step % 7 == 3:
                d = hashlib.sha256(bytes([step])).digest()
                a.reseed_with_digest(d); b.reseed_digest(d)
            if step % 11 == 5:
                a.reseed_with_field_elements([va, va]); b.reseed_elements([vb, vb])
            if step % 13 == 6:
                a.reseed_with_int(step * 1234567); b.reseed_int(step * 1234567)
        assert a.draw_queries(32, 1 << 23) == b.draw_queries(32, 1 << 23)
        assert a.draw_queries(5, 96) == b.draw_queries(5, 96)          # non power of two range: zone rejection path
        for nonce in range(1, 200):
            assert a.verify_proof_of_work(4, nonce) == b.check_pow(4, nonce)
    # byte order of the coin: bytes are popped from the end of SHA-256(seed || counter_be), assembled big-endian
    c = CH.PublicCoin(seed)
    d = hashlib.sha256(seed + (1).to_bytes(8, "big")).digest()
    assert c.next_u64() == int.from_bytes(d[::-1][:8], "big")


def test_merkle_view_roundtrip_like_reference(orc):
    # src/merkle.rs:528-581: prove then verify, single leaf / all leaves / large tree
    rng = np.random.default_rng(3)
    for n, ids in ((8, [3]), (4, [0, 1, 2, 3]), (1 << 10, [378]), (64, [5, 4, 63, 17, 16, 5]), (2, [1])):
        leaves = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
        nodes = orc.merkle_nodes(leaves)
        view = SO._merkle_prove(leaves, nodes, ids)
        SO.merkle_verify(nodes[1].tobytes(), view, ids)
        if n > 2:
            key = "nodes" if view["nodes"] else "initial_leaves"     # all leaves opened: there are no path nodes
            view[key][0] = bytes(32)
            with pytest.raises(SO.VerificationError):
                SO.merkle_verify(nodes[1].tobytes(), view, ids)


def test_brainfuck_hello_world_cpu_prove_verify(orc):
    """BASELINE config 2: examples/brainfuck hello_world.bf with the reference's ProofOptions(19, 16, 20, 16, 16).
    SURVEY.md §8d: 108 program words, n = 2048, 17 Fp + 9 Fq3 columns, ce_blowup 16."""
    from ministark_b200.examples import brainfuck as bf
    trace, out = bf.simulate(bf.HELLO_WORLD)
    assert out == b"Hello World" and len(trace) == 2048 and len(bf.compile_program(bf.HELLO_WORLD)) == 108
    assert trace.base_columns().shape == (17, 2048)
    claim = bf.BrainfuckClaim(bf.HELLO_WORLD, b"", out)
    mk = lambda n, o: Air(claim.AirConfig, n, claim, ProofOptions(*o))
    air = mk(2048, (19, 16, 20, 16, 16))
    assert air.ce_blowup_factor == 16 and air.num_challenges() == 11 and len(air.constraints) == 48
    proof = SO.cpu_prove(claim, (19, 16, 20, 16, 16), trace.base_columns(), mk, ext_builder=trace.build_extension_columns)
    SO.verify(claim, proof, bf.SECURITY_LEVEL, mk)
    pr = SO.parse_proof(proof, 3)
    assert len(pr["fri_layers"]) == 2 and len(pr["remainder"]) == 8            # 32768 -> 2048 -> 128 evals -> 8 coefficients
    bad = bf.BrainfuckClaim(bf.HELLO_WORLD, b"", b"Hello World?")
    with pytest.raises(SO.VerificationError):
        SO.verify(bad, proof, bf.SECURITY_LEVEL, lambda n, o: Air(bad.AirConfig, n, bad, ProofOptions(*o)))


def test_brainfuck_vm_with_input():
    from ministark_b200.examples import brainfuck as bf
    trace, out = bf.simulate(",>,<.>.", b"hi")           # echo two bytes
    assert out == b"hi"
    rows = trace.rows
    assert sum(1 for r in rows if r[bf.CURR_INSTR] == bf.READ) == 2
    assert [r[bf.IN_VALUE] for r in rows[:3]] == [ord("h"), ord("i"), 0]
    # ChaCha (test_rng restatement): the quarter-round core reproduces the RFC 7539 section 2.3.2 block (20 rounds)
    key = [int.from_bytes(bytes(range(4 * i, 4 * i + 4)), "little") for i in range(8)]
    st = [0x61707865, 0x3320646E, 0x79622D32, 0x6B206574] + key + [1, 0x09000000, 0x4A000000, 0]
    w = bf._chacha_core(st, 20)
    assert w[:4] == [0xE4E7F110, 0x15593BD1, 0x1FDD0F50, 0xC47120A3] and w[15] == 0x4E3C50A2
    a, b = bf.test_rng_fq3(2)
    assert a != b and all(0 <= c < bf.P for c in a + b)


def test_proof_wire_format_two_implementations_agree():
    """ministark_b200/proof.py (writer, product) and oracle/stark_oracle.parse_proof (reader, checker) implement the
    ark-serialize layout of Proof independently: a synthetic proof object round-trips field by field"""
    from ministark_b200 import proof as PR
    rng = random.Random(11)
    dg = lambda: bytes(rng.randrange(256) for _ in range(32))
    fq = lambda: tuple(rng.randrange(fib.P) for _ in range(3))
    fp = lambda: rng.randrange(fib.P)
    view = lambda k: PR.MerkleView([dg() for _ in range(k)], [dg() for _ in range(3)], [dg() for _ in range(2)], 17)
    for with_ext in (True, False):
        layers = [PR.LayerProof([fq() for _ in range(16)], view(5), dg()), PR.LayerProof([fq() for _ in range(8)], view(0), dg())]
        q = PR.Queries([fp() for _ in range(6)], [fq() for _ in range(4)] if with_ext else [], [fq() for _ in range(4)], view(7),
                       view(6) if with_ext else None, view(4))
        p = PR.Proof(ProofOptions(19, 16, 20, 16, 16), 1 << 11, dg(), dg() if with_ext else None, dg(), PR.FriProof(layers, [fq(), fq()]),
                     0xDEADBEEF12345, q, [fq() for _ in range(5)], [fq() for _ in range(2)])
        b = p.to_bytes()
        d = SO.parse_proof(b, 3)
        assert d["options"] == (19, 16, 20, 16, 16) and d["trace_len"] == 2048 and d["pow_nonce"] == 0xDEADBEEF12345
        assert (d["base_root"], d["ext_root"], d["comp_root"]) == (p.base_trace_commitment, p.extension_trace_commitment,
                                                                   p.composition_trace_commitment)
        assert [l["rows"] for l in d["fri_layers"]] == [l.flattenend_rows for l in layers]
        assert [l["root"] for l in d["fri_layers"]] == [l.commitment for l in layers]
        assert d["fri_layers"][0]["view"] == dict(nodes=layers[0].merkle_proof.nodes, initial_leaves=layers[0].merkle_proof.initial_leaves,
                                                  sibling_leaves=layers[0].merkle_proof.sibling_leaves, height=17)
        assert d["remainder"] == p.fri_proof.remainder_coeffs
        assert d["base_values"] == [(v, 0, 0) for v in q.base_trace_values] and d["ext_values"] == q.extension_trace_values
        assert d["comp_values"] == q.composition_trace_values and (d["ext_view"] is None) == (not with_ext)
        assert d["trace_oods"] == p.execution_trace_ood_evals and d["comp_oods"] == p.composition_trace_ood_evals
        with pytest.raises(ValueError):
            SO.parse_proof(b + b"\\x00", 3)
        with pytest.raises(ValueError):
            SO.parse_proof(b[:-1], 3)
