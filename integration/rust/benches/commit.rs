// benches/commit.rs for the plonky2 crate -- the commit-path benchmark SURVEY.md 8(d) asks for, so that the genuine
// rayon number can be taken wherever cargo exists (this repository's image has no Rust toolchain: bench.py times the
// C++ restatement under oracle/ instead and says so in its `cpu_baseline.kind`).
//
//   RUSTFLAGS=-Ctarget-cpu=native cargo bench --bench commit                 # reference CPU path (rayon)
//   RUSTFLAGS=-Ctarget-cpu=native cargo bench --bench commit --features cuda # the same calls through libp2g_cuda.so
//
// Shapes are BASELINE.json's: Goldilocks 2^20 x 135 (config 2, reduce with P2G_BENCH_LOG_N for a quick run) and
// BabyBear 2^22 x 256 (config 3), rate_bits 3, cap_height 4, blinding off; inputs are the SplitMix64 columns of
// SURVEY.md 8(d) (value[col][row] = splitmix64(42 ^ (col << 32 | row)) mod p), the ones bench.py and the parity tests use.
// Add to plonky2/Cargo.toml:   [[bench]] name = "commit"  harness = false
use criterion::{criterion_group, criterion_main, BatchSize, Criterion, Throughput};
use plonky2::fri::oracle::PolynomialBatch;
use plonky2::plonk::config::{Poseidon2BabyBearConfig, PoseidonGoldilocksConfig};
use plonky2::util::proving_process_info::ProvingProcessInfo;
use plonky2_field::polynomial::PolynomialValues;
use plonky2_field::types::Field;

fn splitmix64(x: u64) -> u64 {
    let x = x.wrapping_add(0x9E37_79B9_7F4A_7C15);
    let mut z = x;
    z = (z ^ (z >> 30)).wrapping_mul(0xBF58_476D_1CE4_E5B9);
    z = (z ^ (z >> 27)).wrapping_mul(0x94D0_49BB_1331_11EB);
    z ^ (z >> 31)
}

fn synthetic_columns<F: Field>(cols: usize, log_n: usize, order: u64) -> Vec<PolynomialValues<F>> {
    (0..cols as u64)
        .map(|c| PolynomialValues::new((0..1u64 << log_n).map(|r| F::from_canonical_u64(splitmix64(42 ^ (c << 32 | r)) % order)).collect()))
        .collect()
}

fn bench_commit(c: &mut Criterion) {
    let log_gl: usize = std::env::var("P2G_BENCH_LOG_N").ok().and_then(|s| s.parse().ok()).unwrap_or(20);
    let mut g = c.benchmark_group("polynomial_batch_from_values");
    g.sample_size(10);
    {
        type F = p3_goldilocks::Goldilocks;
        const D: usize = 2;
        let values = synthetic_columns::<F>(135, log_gl, 0xFFFF_FFFF_0000_0001);
        g.throughput(Throughput::Elements((135u64) << log_gl)); // rows x cols, the unit of BASELINE.json's metric
        g.bench_function(format!("goldilocks_poseidon_2^{log_gl}x135_r3_h4"), |b| {
            b.iter_batched(
                || values.clone(),
                |v| PolynomialBatch::<F, PoseidonGoldilocksConfig, D, 4>::from_values(v, 3, false, 4, &mut ProvingProcessInfo::default(), None),
                BatchSize::LargeInput,
            )
        });
    }
    {
        type F = p3_baby_bear::BabyBear;
        const D: usize = 4;
        let log_bb = log_gl + 2;
        let values = synthetic_columns::<F>(256, log_bb, 2013265921);
        g.throughput(Throughput::Elements((256u64) << log_bb));
        g.bench_function(format!("babybear_poseidon2_2^{log_bb}x256_r3_h4"), |b| {
            b.iter_batched(
                || values.clone(),
                |v| PolynomialBatch::<F, Poseidon2BabyBearConfig, D, 8>::from_values(v, 3, false, 4, &mut ProvingProcessInfo::default(), None),
                BatchSize::LargeInput,
            )
        });
    }
    g.finish();
}

criterion_group!(benches, bench_commit);
criterion_main!(benches);
