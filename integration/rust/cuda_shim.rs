// Bodies for the `cuda` feature of the plonky2 crate. Written against the ABI; not compiled in this repository's image.
// plonky2/src/fri/oracle.rs
#[cfg(feature = "cuda")]
pub fn from_values(values: Vec<PolynomialValues<F>>, rate_bits: usize, blinding: bool, cap_height: usize,
                   timing: &mut ProvingProcessInfo, _fft_root_table: Option<&FftRootTable<F>>) -> Self {
    cuda::commit::<F, C, D, NUM_HASH_OUT_ELTS>(
        values.into_iter().map(|v| v.values).collect(), P2G_INPUT_VALUES, rate_bits, blinding, cap_height, timing)
}
#[cfg(feature = "cuda")]
pub fn from_coeffs(polynomials: Vec<PolynomialCoeffs<F>>, /* same */) -> Self { /* P2G_INPUT_COEFFS */ }

// plonky2/src/fri/cuda.rs (new)
pub(crate) fn commit<F: RichField + HasExtension<D>, C, const D: usize, const N: usize>(
    cols: Vec<Vec<F>>, kind: i32, rate_bits: usize, blinding: bool, cap_height: usize,
    timing: &mut ProvingProcessInfo) -> PolynomialBatch<F, C, D, N> {
    // field / hasher dispatch exactly like hash_types.rs:269-275 does for bytes:
    let (field, hash) = match F::ORDER_U64 {
        Goldilocks::ORDER_U64 => (P2G_FIELD_GOLDILOCKS, P2G_HASH_POSEIDON_GL12),
        BabyBear::ORDER_U64   => (P2G_FIELD_BABYBEAR,   P2G_HASH_POSEIDON2_BB16),
        _ => unreachable!(),
    };
    assert!(TypeId::of::<C::Hasher>() != TypeId::of::<KeccakHash<25>>(), "cuda feature: Poseidon hashers only");
    let degree = cols[0].len();
    for c in &cols { assert_eq!(c.len(), degree, "Polynomial degrees inconsistent"); }   // oracle.rs:139
    let n_total = degree << rate_bits;
    let salts: Vec<Vec<F>> = if blinding { (0..SALT_SIZE).map(|_| F::rand_vec(n_total)).collect() } else { vec![] };
    let width = cols.len() + salts.len();
    // p3's BabyBear is a u32 in Montgomery form, p3's Goldilocks a u64: the Vec<F> buffers are passed as they are.
    let repr = if field == P2G_FIELD_BABYBEAR { P2G_REPR_MONTY_R32 } else { P2G_REPR_CANONICAL };
    let mut coeffs: Vec<Vec<F>> = if kind == P2G_INPUT_VALUES { cols.iter().map(|_| Vec::with_capacity(degree)).collect() } else { vec![] };
    let mut leaves_flat: Vec<F> = Vec::with_capacity(n_total * width);
    let num_digests = 2 * (n_total - (1 << cap_height));
    let mut digests: Vec<HashOut<F, N>> = Vec::with_capacity(num_digests);
    let mut cap: Vec<HashOut<F, N>> = Vec::with_capacity(1 << cap_height);
    let mut desc = p2g_commit_desc { field, hash, repr, input_kind: kind, log_n: log2_strict(degree) as u32,
        num_cols: cols.len() as u32, rate_bits: rate_bits as u32, cap_height: cap_height as u32,
        num_salt_cols: salts.len() as u32, flags: 0, /* pointer tables built from the Vecs above */ .. };
    let rc = unsafe { p2g_commit(context(), &mut desc) };
    if rc != 0 { panic!("{}", last_error()); }            // the reference panics on misuse; so do we
    unsafe { /* set_len on coeffs / leaves_flat / digests / cap */ }
    for (name, s) in ["IFFT", "FFT + blinding", "transpose LDEs", "build Merkle tree"].iter().zip(desc.stage_seconds) {
        timing.push(name, log::Level::Debug); timing.pop_with_duration(s);   // keeps the four timed! scopes
    }
    let leaves: Vec<Vec<F>> = leaves_flat.chunks_exact(width).map(|r| r.to_vec()).collect();  // MerkleTree.leaves is Vec<Vec<F>>
    PolynomialBatch { polynomials: /* coeffs or the input */, merkle_tree: MerkleTree { leaves, digests, cap: MerkleCap(cap) },
                      degree_log: log2_strict(degree), rate_bits, blinding }
}

// plonky2/src/hash/merkle_tree.rs
#[cfg(feature = "cuda")]
pub fn new(leaves: Vec<Vec<F>>, cap_height: usize) -> Self {
    let log2_leaves_len = log2_strict(leaves.len());
    assert!(cap_height <= log2_leaves_len, "cap_height={cap_height} should be at most log2(leaves.len())={log2_leaves_len}");
    // flatten rows (all rows of one tree have the same length on the commit path), call p2g_merkle, wrap the outputs
}
