// examples/dump_vectors.rs for the plonky2 crate: prints, from the REAL crates (p3-goldilocks, p3-baby-bear,
// p3-poseidon2 as the reference links them, Cargo.toml:17-24), the constants and vectors that SURVEY.md 8(c) lists as
// unpinned by any literal in the reference tree, as JSON in the schema of tests/golden/.  Run it wherever cargo exists
//     cargo run --release --example dump_vectors > external_vectors.json
// and drop the file into tests/golden/: tests/test_oracle.py::test_external_vectors_if_present then pins the oracle
// (and through it the CUDA path) on bytes produced by the reference itself.  Written against the reference's API; NOT
// compiled in this repository's image (no Rust toolchain).
use p3_baby_bear::BabyBear;
use p3_field::{AbstractField, Field, PrimeField64, TwoAdicField};
use p3_goldilocks::Goldilocks;
use p3_symmetric::Permutation;
use plonky2::field::polynomial::PolynomialValues;
use plonky2::fri::oracle::PolynomialBatch;
use plonky2::hash::poseidon2_babybear::poseidon2;
use plonky2::hash::poseidon_goldilocks::Poseidon;
use plonky2::plonk::config::{Poseidon2BabyBearConfig, PoseidonGoldilocksConfig};
use plonky2::util::proving_process_info::ProvingProcessInfo;
use plonky2_field::types::HasExtension;
use plonky2_field::{BABYBEAR_EXTENSION_FIELD_DEGREE, BABYBEAR_NUM_HASH_OUT_ELTS, GOLDILOCKS_EXTENSION_FIELD_DEGREE, GOLDILOCKS_NUM_HASH_OUT_ELTS};

/// SplitMix64 of seed ^ (col << 32 | row), reduced mod p: the synthetic input of SURVEY.md 8(d) / tests/util.py
fn synthetic(seed: u64, col: u64, row: u64) -> u64 {
    let mut z = (seed ^ (col << 32 | row)).wrapping_add(0x9E3779B97F4A7C15);
    z = (z ^ (z >> 30)).wrapping_mul(0xBF58476D1CE4E5B9);
    z = (z ^ (z >> 27)).wrapping_mul(0x94D049BB133111EB);
    z ^ (z >> 31)
}

fn list(v: impl IntoIterator<Item = u64>) -> String {
    format!("[{}]", v.into_iter().map(|x| x.to_string()).collect::<Vec<_>>().join(", "))
}

fn commit_cap<F, C, const D: usize, const N: usize>(log_n: usize, cols: usize) -> String
where
    F: plonky2::hash::hash_types::RichField + TwoAdicField + HasExtension<D>,
    C: plonky2::plonk::config::GenericConfig<D, N, F = F, FE = F::Extension>,
{
    let values: Vec<PolynomialValues<F>> = (0..cols as u64)
        .map(|c| PolynomialValues::new((0..1u64 << log_n).map(|r| F::from_canonical_u64(synthetic(42, c, r) % F::ORDER_U64)).collect()))
        .collect();
    let mut timing = ProvingProcessInfo::new("dump", log::Level::Debug);
    let batch = PolynomialBatch::<F, C, D, N>::from_values(values, 3, false, 4, &mut timing, None);
    let cap: Vec<String> = batch.merkle_tree.cap.0.iter().map(|h| list(h.elements.iter().map(|e| e.as_canonical_u64()))).collect();
    format!("{{\"log_n\": {log_n}, \"cols\": {cols}, \"rate_bits\": 3, \"cap_height\": 4, \"seed\": 42, \"cap\": [{}]}}", cap.join(", "))
}

fn main() {
    let gl_roots = list((0..=32).map(|k| Goldilocks::two_adic_generator(k).as_canonical_u64()));
    let bb_roots = list((0..=27).map(|k| BabyBear::two_adic_generator(k).as_canonical_u64()));
    // Poseidon2 width 16 on 0..15 (the input of gates/poseidon2_babybear.rs:960-1004) and on one synthetic state
    let mut s0: [BabyBear; 16] = core::array::from_fn(|i| BabyBear::from_canonical_u32(i as u32));
    poseidon2.permute_mut(&mut s0);
    let mut s1: [BabyBear; 16] = core::array::from_fn(|i| BabyBear::from_canonical_u64(synthetic(7, 0, i as u64) % BabyBear::ORDER_U64));
    let s1_in = list(s1.iter().map(|x| x.as_canonical_u64()));
    poseidon2.permute_mut(&mut s1);
    // Poseidon width 12 on 0..11 (also a literal KAT of poseidon_goldilocks.rs:1169-1190; here as a cross-check of the dump itself)
    let g0: [Goldilocks; 12] = core::array::from_fn(|i| Goldilocks::from_canonical_u64(i as u64));
    let g0_out = <Goldilocks as Poseidon>::poseidon(g0);
    println!("{{");
    println!("  \"source\": \"integration/rust/dump_vectors.rs run against the reference crates\",");
    println!("  \"goldilocks\": {{\"generator\": {}, \"two_adic_generators\": {}, \"extension_degree\": {}, \"extension_w\": {},",
        Goldilocks::generator().as_canonical_u64(), gl_roots, GOLDILOCKS_EXTENSION_FIELD_DEGREE,
        <Goldilocks as HasExtension<GOLDILOCKS_EXTENSION_FIELD_DEGREE>>::Extension::w().as_canonical_u64());
    println!("    \"poseidon_0_to_11\": {},", list(g0_out.iter().map(|x| x.as_canonical_u64())));
    println!("    \"from_values\": {}}},", commit_cap::<Goldilocks, PoseidonGoldilocksConfig, GOLDILOCKS_EXTENSION_FIELD_DEGREE, GOLDILOCKS_NUM_HASH_OUT_ELTS>(8, 9));
    println!("  \"babybear\": {{\"generator\": {}, \"two_adic_generators\": {}, \"extension_degree\": {}, \"extension_w\": {},",
        BabyBear::generator().as_canonical_u64(), bb_roots, BABYBEAR_EXTENSION_FIELD_DEGREE,
        <BabyBear as HasExtension<BABYBEAR_EXTENSION_FIELD_DEGREE>>::Extension::w().as_canonical_u64());
    println!("    \"poseidon2_0_to_15\": {},", list(s0.iter().map(|x| x.as_canonical_u64())));
    println!("    \"poseidon2_synthetic\": {{\"input\": {}, \"output\": {}}},", s1_in, list(s1.iter().map(|x| x.as_canonical_u64())));
    println!("    \"from_values\": {}}}", commit_cap::<BabyBear, Poseidon2BabyBearConfig, BABYBEAR_EXTENSION_FIELD_DEGREE, BABYBEAR_NUM_HASH_OUT_ELTS>(8, 9));
    println!("}}");
}
