//! The AIR of examples/fib (examples/fib/main.rs:54-172) stated against the reference's public items, and one proof.
//! 8 columns hold 8 consecutive terms of v_k = v_(k-2) * v_(k-1) (v_0 = 1, v_1 = 2) per row.
use ark_ff::One;
use ark_poly::{EvaluationDomain, Radix2EvaluationDomain};
use ark_serialize::CanonicalSerialize;
use ministark::air::AirConfig;
use ministark::challenges::Challenges;
use ministark::constraints::{AlgebraicItem, Constraint, ExecutionTraceColumn};
use ministark::expression::Expr;
use ministark::hash::{HashFn, Sha256HashFn};
use ministark::hints::Hints;
use ministark::merkle::MatrixMerkleTreeImpl;
use ministark::random::{PublicCoin, PublicCoinImpl};
use ministark::stark::Stark;
use ministark::utils::{FieldVariant, SerdeOutput};
use ministark::{Air, Matrix, ProofOptions, Trace};
use ministark_gpu::fields::p18446744069414584321::ark::Fp;
use num_traits::Pow;
use sha2::{Digest as _, Sha256};

pub struct FibTrace(Matrix<Fp>);

impl Trace for FibTrace {
    type Fp = Fp;
    type Fq = Fp;
    fn len(&self) -> usize {
        self.0.num_rows()
    }
    fn base_columns(&self) -> &Matrix<Self::Fp> {
        &self.0
    }
}

pub struct FibAir;

impl AirConfig for FibAir {
    const NUM_BASE_COLUMNS: usize = 8;
    type Fp = Fp;
    type Fq = Fp;
    type PublicInputs = Fp;

    fn gen_hints(_n: usize, claimed: &Fp, _: &Challenges<Fp>) -> Hints<Fp> {
        Hints::new(vec![(0, *claimed)])
    }

    fn constraints(trace_len: usize) -> Vec<Constraint<FieldVariant<Fp, Fp>>> {
        use AlgebraicItem::*;
        let xs = Radix2EvaluationDomain::<Fp>::new(trace_len).unwrap();
        let first = Constant(FieldVariant::Fp(xs.element(0)));
        let last = Constant(FieldVariant::Fp(xs.element(trace_len - 1)));
        let one = Constant(FieldVariant::Fp(Fp::one()));
        // the first row holds v_0 .. v_7 (written out like the example: a sum of two items is an Expr, products take refs)
        let v0 = one;
        let v1 = v0 + v0;
        let v2 = &v1 * v0;
        let v3 = &v1 * &v2;
        let v4 = &v2 * &v3;
        let v5 = &v3 * &v4;
        let v6 = &v4 * &v5;
        let v7 = &v5 * &v6;
        let mut cs = vec![
            (0.curr() - v0) / (X - first),
            (1.curr() - v1) / (X - first),
            (2.curr() - v2) / (X - first),
            (3.curr() - v3) / (X - first),
            (4.curr() - v4) / (X - first),
            (5.curr() - v5) / (X - first),
            (6.curr() - v6) / (X - first),
            (7.curr() - v7) / (X - first),
        ];
        cs.push((7.curr() - Hint(0)) / (X - last));
        // sequence value j of a row pair: the current row holds 0..8, the next row 8..16; v_(c+8) = v_(c+6) * v_(c+7)
        let cell = |j: usize| -> Expr<AlgebraicItem<FieldVariant<Fp, Fp>>> { if j < 8 { j.curr() } else { (j - 8).next() } };
        for c in 0..8usize {
            let step = cell(c + 8) - cell(c + 6) * cell(c + 7);
            // every row but the last: (x - t_(n-1)) / (x^n - 1)
            cs.push(step * ((X - last) / (X.pow(trace_len) - one)));
        }
        cs.into_iter().map(Constraint::new).collect()
    }
}

pub struct FibClaim(Fp);

impl Stark for FibClaim {
    type Fp = Fp;
    type Fq = Fp;
    type AirConfig = FibAir;
    type Digest = SerdeOutput<Sha256>;
    type PublicCoin = PublicCoinImpl<Fp, Sha256HashFn>;
    type MerkleTree = MatrixMerkleTreeImpl<Sha256HashFn>;
    type Witness = FibTrace;
    type Trace = FibTrace;

    fn get_public_inputs(&self) -> Fp {
        self.0
    }
    fn generate_trace(&self, witness: FibTrace) -> FibTrace {
        witness
    }
    fn gen_public_coin(&self, air: &Air<FibAir>) -> Self::PublicCoin {
        let mut seed = Vec::new();
        air.public_inputs().serialize_compressed(&mut seed).unwrap();
        air.trace_len().serialize_compressed(&mut seed).unwrap();
        air.options().serialize_compressed(&mut seed).unwrap();
        PublicCoinImpl::new(Sha256HashFn::hash_chunks([&*seed]))
    }
}

fn gen_trace(num_rows: usize) -> (FibTrace, Fp) {
    let mut rows: Vec<Vec<Fp>> = Vec::with_capacity(num_rows);
    let mut v = vec![Fp::one(), Fp::one() + Fp::one()];
    for i in 2..8 {
        let t = v[i - 2] * v[i - 1];
        v.push(t);
    }
    for _ in 0..num_rows {
        rows.push(v.clone());
        let mut w = vec![v[6] * v[7]];
        w.push(v[7] * w[0]);
        for i in 2..8 {
            let t = w[i - 2] * w[i - 1];
            w.push(t);
        }
        v = w;
    }
    let last = rows[num_rows - 1][7];
    (FibTrace(Matrix::from_rows(rows)), last)
}

/// proof bytes for 2^7 rows with the example's own options; tests/golden/golden_r01.json holds what the restatement
/// produces for the same claim ("fib_2p7_rows_proof_sha256")
pub fn proof_vector() -> String {
    let options = ProofOptions::new(32, 4, 8, 8, 64);
    let (trace, last) = gen_trace(1 << 7);
    let claim = FibClaim(last);
    let proof = pollster::block_on(claim.prove(options, trace)).expect("prover failed");
    let mut bytes = Vec::new();
    proof.serialize_compressed(&mut bytes).unwrap();
    let digest: String = Sha256::digest(&bytes).iter().map(|b| format!("{b:02x}")).collect();
    claim.verify(proof, 30).expect("verification failed");
    use ark_ff::PrimeField;
    format!(
        "\"fib_proof\": {{\"log_rows\": 7, \"options\": [32,4,8,8,64], \"claim_canonical\": {}, \"proof_len\": {}, \"proof_sha256\": \"{}\", \"pow_nonce_rule\": \"serial find: smallest nonce >= 1\"}}",
        last.into_bigint().0[0],
        bytes.len(),
        digest
    )
}
