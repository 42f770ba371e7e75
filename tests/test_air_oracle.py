"""CPU-only: the independent restatement of the AIR bookkeeping (oracle/air_oracle.py: degree rules, blow-up factors,
composition constraint, trace arguments — src/constraints.rs:340-455, src/air.rs:50-82,142-246) against the product's
ministark_b200/air.py, and an examples/fib proof verified with an AIR description that shares NOTHING with the prover's."""
import pytest

from ministark_b200.air import Air, ProofOptions
from ministark_b200.examples import brainfuck as bf
from ministark_b200.examples import fib, perm
from oracle import air_oracle as AO


def _cases():
    yield "fib", fib.FibAirConfig, 64, 5, (32, 4, 8, 8, 64)
    yield "fib-large", fib.FibAirConfig, 1 << 12, 5, (32, 4, 8, 8, 64)
    yield "perm", perm.PermAirConfig, 64, [], (8, 8, 0, 4, 4)
    yield "brainfuck", bf.BrainfuckAirConfig, 2048, bf.BrainfuckClaim("+.", b"", b"\x01"), (19, 16, 20, 16, 16)


@pytest.mark.parametrize("name,cfg,n,pub,opts", list(_cases()), ids=[c[0] for c in _cases()])
def test_air_bookkeeping_matches_the_independent_restatement(name, cfg, n, pub, opts):
    air = Air(cfg, n, pub, ProofOptions(*opts))
    cons = [c.to_tuple() for c in air.constraints]
    oa = AO.OracleAir(cons, n, pub, opts, cfg.gen_hints, cfg.NUM_BASE_COLUMNS)
    assert oa.ce_blowup_factor == air.ce_blowup_factor
    assert oa.trace_arguments() == air.trace_arguments()
    assert oa.num_challenges() == air.num_challenges()
    assert oa.num_composition_constraint_coeffs() == air.num_composition_constraint_coeffs()
    assert AO.structurally_equal(oa.composition_tuple, air.composition_constraint.to_tuple())
    for c in air.constraints:
        from ministark_b200.air import blowup_factor, degree
        assert AO.degree(c.to_tuple(), n - 1) == degree(c, n - 1)
        assert AO.blowup_factor(c.to_tuple(), n) == blowup_factor(c, n)


def test_fib_constraints_restated_from_the_example():
    for n in (16, 128, 1 << 10):
        mine = [c.to_tuple() for c in fib.FibAirConfig.constraints(n)]
        ref = AO.fib_constraints(n)
        assert len(mine) == len(ref) == 17
        assert all(AO.structurally_equal(a, b) for a, b in zip(mine, ref))


def test_fib_proof_verifies_against_the_independent_air(orc):
    """prove with the CPU restatement of default_prove using the PRODUCT's AIR description, verify with the oracle's own
    AIR description (constraints from examples/fib restated in oracle/air_oracle.py, bookkeeping rules restated there too)"""
    from oracle import stark_oracle as SO
    opts = (16, 4, 6, 8, 16)
    trace, last = fib.gen_trace(8 << 6)
    claim = fib.FibClaim(last)
    mk_product = lambda n, o: Air(claim.AirConfig, n, claim.get_public_inputs(), ProofOptions(*o))
    proof = SO.cpu_prove(claim, opts, trace.base_columns(), mk_product)
    mk_oracle = lambda n, o: AO.OracleAir(AO.fib_constraints(n), n, last, tuple(o), AO.fib_hints, 8)
    SO.verify(claim, proof, 10, mk_oracle)
    assert SO.cpu_prove(claim, opts, trace.base_columns(), mk_oracle) == proof      # and the prover restatement agrees too
    bad = fib.FibClaim((last + 1) % fib.P)
    with pytest.raises(SO.VerificationError):
        SO.verify(bad, proof, 10, lambda n, o: AO.OracleAir(AO.fib_constraints(n), n, (last + 1) % fib.P, tuple(o), AO.fib_hints, 8))
