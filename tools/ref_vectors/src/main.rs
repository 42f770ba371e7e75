//! Reference vectors for ministark_b200's parity tests — see ../Cargo.toml.
//!
//! Each section pins one convention that SURVEY.md §8(c) lists as "known only from upstream memory" or as
//! unpinned by the reference's own tests; tests/test_ref_vectors.py names the oracle function it checks.
//!
//!   consts      Fp::ONE / GENERATOR / TWO_ADIC_ROOT_OF_UNITY as in-memory (Montgomery) and canonical words
//!   ntt         Radix2EvaluationDomain::{fft, ifft} subgroup + coset (src/matrix.rs:134,185 call sites)
//!   lde         Matrix::interpolate + bit_reversed_evaluate (src/matrix.rs:157-163,225-234)
//!   hash        Sha256HashFn::hash_elements (src/hash.rs:92-99), Fp and Fq3 rows
//!   merkle      MatrixMerkleTreeImpl::from_matrix().root() and prove_rows() bytes (src/merkle.rs:149-207,359-361)
//!   coin        PublicCoinImpl draws, queries, proof of work (src/random.rs:91-196)
//!   serialize   ark-serialize layouts of Fp, Fq3, Vec, Option, digest (src/utils.rs:552-582)
//!   test_rng    ark_std::test_rng() Fq3 draws (examples/brainfuck/trace.rs:82-84)
//!   fib_proof   examples/fib AIR, Stark::prove bytes (examples/fib/main.rs:54-243), feature "proof"
#![feature(allocator_api)]

use ark_ff::{FftField, Field, One, PrimeField, UniformRand, Zero};
use ark_poly::{EvaluationDomain, Radix2EvaluationDomain};
use ark_serialize::CanonicalSerialize;
use ministark::hash::{Digest as _, ElementHashFn, HashFn, Sha256HashFn};
use ministark::merkle::{MatrixMerkleTree, MatrixMerkleTreeImpl, MerkleTree};
use ministark::random::{PublicCoin, PublicCoinImpl};
use ministark::Matrix;
use ministark_gpu::fields::p18446744069414584321::ark::{Fp, Fq3};
use sha2::{Digest as _, Sha256};

#[cfg(feature = "proof")]
mod fib_air;

const P: u64 = 0xFFFF_FFFF_0000_0001;

/// SURVEY.md §8(d) generator (oracle/gl_oracle.c orc_splitmix_fill): splitmix64, reject draws >= p; canonical value
fn splitmix_column(n: usize, seed: u64) -> Vec<Fp> {
    let mut s = seed;
    let mut out = Vec::with_capacity(n);
    while out.len() < n {
        s = s.wrapping_add(0x9E37_79B9_7F4A_7C15);
        let mut z = s;
        z = (z ^ (z >> 30)).wrapping_mul(0xBF58_476D_1CE4_E5B9);
        z = (z ^ (z >> 27)).wrapping_mul(0x94D0_49BB_1331_11EB);
        z ^= z >> 31;
        if z < P {
            out.push(Fp::from(z));
        }
    }
    out
}

fn canon(x: &Fp) -> u64 {
    x.into_bigint().0[0]
}

fn hex(bytes: &[u8]) -> String {
    bytes.iter().map(|b| format!("{b:02x}")).collect()
}

fn list(v: &[Fp]) -> String {
    format!("[{}]", v.iter().map(|x| canon(x).to_string()).collect::<Vec<_>>().join(","))
}

fn fq3_list(v: &[Fq3]) -> String {
    format!(
        "[{}]",
        v.iter().map(|x| format!("[{},{},{}]", canon(&x.c0), canon(&x.c1), canon(&x.c2))).collect::<Vec<_>>().join(",")
    )
}

/// SHA-256 of the canonical little-endian words (what tests/test_ref_vectors.py hashes on its side)
fn digest_of(cols: &[&[Fp]]) -> String {
    let mut h = Sha256::new();
    for col in cols {
        for x in col.iter() {
            h.update(canon(x).to_le_bytes());
        }
    }
    hex(&h.finalize())
}

fn ser<T: CanonicalSerialize>(v: &T) -> String {
    let mut b = Vec::new();
    v.serialize_compressed(&mut b).unwrap();
    hex(&b)
}

fn main() {
    let mut out: Vec<String> = Vec::new();

    // ---- consts: in-memory representation (ark-ff-optimized keeps the Montgomery word in .0)
    out.push(format!(
        "\"consts\": {{\"one_canonical\": {}, \"generator_canonical\": {}, \"two_adic_root_canonical\": {}, \"two_adicity\": {}}}",
        canon(&Fp::one()),
        canon(&Fp::GENERATOR),
        canon(&Fp::TWO_ADIC_ROOT_OF_UNITY),
        Fp::TWO_ADICITY
    ));

    // ---- ntt
    let mut ntt = Vec::new();
    for (log_n, full) in [(4usize, true), (10, false), (16, false)] {
        let n = 1usize << log_n;
        let coeffs = splitmix_column(n, 0x9E37_79B9_7F4A_7C15 ^ (77 + log_n as u64));
        for coset in [false, true] {
            let dom = if coset {
                Radix2EvaluationDomain::<Fp>::new_coset(n, Fp::GENERATOR).unwrap()
            } else {
                Radix2EvaluationDomain::<Fp>::new(n).unwrap()
            };
            let ev = dom.fft(&coeffs);
            let back = dom.ifft(&ev);
            assert_eq!(back, coeffs);
            let inv_of_input = dom.ifft(&coeffs);
            ntt.push(format!(
                "{{\"log_n\": {log_n}, \"coset\": {coset}, \"seed\": {}, \"fft_sha256\": \"{}\", \"ifft_sha256\": \"{}\"{}}}",
                77 + log_n,
                digest_of(&[&ev]),
                digest_of(&[&inv_of_input]),
                if full { format!(", \"input\": {}, \"fft\": {}", list(&coeffs), list(&ev)) } else { String::new() }
            ));
        }
    }
    out.push(format!("\"ntt\": [{}]", ntt.join(",")));

    // ---- lde + hash + merkle: 3 columns x 2^6 rows, blow-up 4
    let (log_n, log_b, ncols) = (6usize, 2usize, 3usize);
    let n = 1usize << log_n;
    let cols: Vec<Vec<Fp>> = (0..ncols).map(|c| splitmix_column(n, 0x9E37_79B9_7F4A_7C15 ^ (500 + c as u64))).collect();
    let rows: Vec<Vec<Fp>> = (0..n).map(|r| cols.iter().map(|c| c[r]).collect()).collect();
    let trace = Matrix::from_rows(rows);
    let trace_dom = Radix2EvaluationDomain::<Fp>::new(n).unwrap();
    let lde_dom = Radix2EvaluationDomain::<Fp>::new_coset(n << log_b, Fp::GENERATOR).unwrap();
    let polys = trace.interpolate(trace_dom);
    let lde = polys.bit_reversed_evaluate(lde_dom);
    let lde_cols: Vec<&[Fp]> = lde.0.iter().map(|c| &c[..]).collect();
    let poly_cols: Vec<&[Fp]> = polys.0.iter().map(|c| &c[..]).collect();
    out.push(format!(
        "\"lde\": {{\"log_n\": {log_n}, \"log_blowup\": {log_b}, \"ncols\": {ncols}, \"seed\": 500, \"polys_sha256\": \"{}\", \"lde_bitrev_sha256\": \"{}\", \"lde_col0_first8\": {}}}",
        digest_of(&poly_cols),
        digest_of(&lde_cols),
        list(&lde.0[0][..8])
    ));
    let row5: Vec<Fp> = lde.0.iter().map(|c| c[5]).collect();
    let q = [Fq3::new(row5[0], row5[1], row5[2]), Fq3::new(Fp::one(), Fp::zero(), Fp::GENERATOR)];
    out.push(format!(
        "\"hash\": {{\"fp_row\": {}, \"fp_row_digest\": \"{}\", \"fq3_row\": {}, \"fq3_row_digest\": \"{}\"}}",
        list(&row5),
        hex(&<Sha256HashFn as ElementHashFn<Fp>>::hash_elements(row5.iter().copied()).as_bytes()),
        fq3_list(&q),
        hex(&<Sha256HashFn as ElementHashFn<Fq3>>::hash_elements(q.iter().copied()).as_bytes())
    ));
    let tree = <MatrixMerkleTreeImpl<Sha256HashFn> as MatrixMerkleTree<Fp>>::from_matrix(&lde);
    let ids = [1usize, 5, 6, 77, 200];
    let view = <MatrixMerkleTreeImpl<Sha256HashFn> as MatrixMerkleTree<Fp>>::prove_rows(&tree, &ids).unwrap();
    out.push(format!(
        "\"merkle\": {{\"root\": \"{}\", \"row_ids\": [1,5,6,77,200], \"view_bytes\": \"{}\"}}",
        hex(&tree.root().as_bytes()),
        ser(&view)
    ));

    // ---- coin
    let seed = Sha256HashFn::hash_chunks([&b"ministark_b200 reference vectors"[..]]);
    let mut coin = PublicCoinImpl::<Fp, Sha256HashFn>::new(seed.clone());
    let fp_draws: Vec<Fp> = (0..4).map(|_| coin.draw()).collect();
    coin.reseed_with_field_elements(&fp_draws[..2]);
    coin.reseed_with_int(12345);
    coin.reseed_with_digest(&seed);
    let after: Vec<Fp> = (0..2).map(|_| coin.draw()).collect();
    let queries = coin.draw_queries(16, 1 << 20);
    let mut coin3 = PublicCoinImpl::<Fq3, Sha256HashFn>::new(seed.clone());
    let q3: Vec<Fq3> = (0..3).map(|_| coin3.draw()).collect();
    let grind = PublicCoinImpl::<Fp, Sha256HashFn>::new(seed.clone());
    let nonce = grind.grind_proof_of_work(12).unwrap();
    out.push(format!(
        "\"coin\": {{\"seed\": \"{}\", \"fp_draws\": {}, \"fp_draws_after_reseeds\": {}, \"queries_16_of_2p20\": [{}], \"fq3_draws\": {}, \"pow_bits\": 12, \"pow_nonce\": {}}}",
        hex(&seed.as_bytes()),
        list(&fp_draws),
        list(&after),
        queries.iter().map(|q| q.to_string()).collect::<Vec<_>>().join(","),
        fq3_list(&q3),
        nonce
    ));

    // ---- serialize
    let some: Option<Fp> = Some(Fp::GENERATOR);
    let none: Option<Fp> = None;
    out.push(format!(
        "\"serialize\": {{\"fp_generator\": \"{}\", \"fq3\": \"{}\", \"vec_fp\": \"{}\", \"option_some\": \"{}\", \"option_none\": \"{}\", \"digest\": \"{}\", \"usize_5\": \"{}\"}}",
        ser(&Fp::GENERATOR),
        ser(&q[1]),
        ser(&fp_draws),
        ser(&some),
        ser(&none),
        ser(&seed),
        ser(&5usize)
    ));

    // ---- ark_std::test_rng(): the fixed-seed StdRng examples/brainfuck uses for its extension columns
    let mut rng = ark_std::test_rng();
    let t: Vec<Fq3> = (0..2).map(|_| Fq3::rand(&mut rng)).collect();
    out.push(format!("\"test_rng\": {{\"fq3_draws\": {}}}", fq3_list(&t)));

    #[cfg(feature = "proof")]
    out.push(fib_air::proof_vector());

    println!("{{\n{}\n}}", out.join(",\n"));
}
