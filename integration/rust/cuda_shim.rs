// The `cuda` cargo feature of the plonky2 crate: PolynomialBatch::from_values / from_coeffs (plonky2/src/fri/oracle.rs:68-123)
// and MerkleTree::new (plonky2/src/hash/merkle_tree.rs:152-181) routed through libp2g_cuda.so, signatures unchanged.
// Written against the C ABI in include/p2g.h (bindings: p2g_sys.rs, generated from the header); NOT compiled in this
// repository's image (no Rust toolchain) -- complete bodies, no elisions, so that a maintainer with cargo can drop the
// three items below into the crate behind `#[cfg(feature = "cuda")]` and keep the CPU bodies behind `#[cfg(not(...))]`.
//
// There is no CPU fallback inside the feature: if the library reports an error the call panics with its message, which
// is what the reference does on misuse (oracle.rs:139, merkle_tree.rs:153-157, util/src/lib.rs:30-33).

// ---------------------------------------------------------------------------------------------------------------------
// plonky2/src/fri/cuda.rs (new file)
// ---------------------------------------------------------------------------------------------------------------------
use core::any::TypeId;
use core::ffi::{c_void, CStr};
use std::sync::OnceLock;

use p2g_sys::*;
use p3_baby_bear::BabyBear;
use p3_field::PrimeField64;
use p3_goldilocks::Goldilocks;
use plonky2_field::types::{HasExtension, Sample};

use crate::field::polynomial::PolynomialCoeffs;
use crate::fri::oracle::{PolynomialBatch, SALT_SIZE};
use crate::hash::hash_types::RichField;
use crate::hash::keccak::KeccakHash;
use crate::hash::merkle_tree::{MerkleCap, MerkleTree};
use crate::plonk::config::{GenericConfig, Hasher};
use plonky2_maybe_rayon::*;

use crate::util::log2_strict;
use crate::util::proving_process_info::ProvingProcessInfo;

struct Ctx(*mut p2g_ctx);
unsafe impl Send for Ctx {}
unsafe impl Sync for Ctx {} // the library serialises calls on a context with an internal lock (threading)

/// One context per process: every visible GPU the environment names (P2G_DEVICES="0,1,2,3"), else device 0.
/// With more than one device the library shards each commit by LDE coset itself (include/p2g.h, p2g_init).
pub(crate) fn context() -> *mut p2g_ctx {
    static CTX: OnceLock<Ctx> = OnceLock::new();
    CTX.get_or_init(|| {
        let ids: Vec<i32> = std::env::var("P2G_DEVICES")
            .ok()
            .map(|s| s.split(',').filter_map(|t| t.trim().parse().ok()).collect())
            .filter(|v: &Vec<i32>| !v.is_empty())
            .unwrap_or_else(|| vec![0]);
        let mut ctx: *mut p2g_ctx = core::ptr::null_mut();
        let rc = unsafe { p2g_init(ids.as_ptr(), ids.len() as i32, &mut ctx) };
        if rc != P2G_OK {
            panic!("p2g_init failed ({rc}): {}", last_error(core::ptr::null_mut()));
        }
        assert_eq!(unsafe { p2g_abi_version() }, P2G_ABI_VERSION, "libp2g_cuda.so ABI mismatch");
        Ctx(ctx)
    })
    .0
}

pub(crate) fn last_error(ctx: *mut p2g_ctx) -> String {
    unsafe { CStr::from_ptr(p2g_last_error(ctx)) }.to_string_lossy().into_owned()
}

/// (field, hash, repr) of the generic parameters, the way hash_types.rs:269-275 dispatches on the field order.
/// p3's Goldilocks is a canonical-or-not u64 and p3's BabyBear a u32 in Montgomery form (R = 2^32): the `Vec<F>`
/// buffers cross the boundary as they are, no conversion pass.
fn dispatch<F: RichField, H: Hasher<F> + 'static>() -> (i32, i32, i32) {
    assert!(
        TypeId::of::<H>() != TypeId::of::<KeccakHash<25>>(),
        "cuda feature: only the Poseidon (Goldilocks) and Poseidon2 (BabyBear) hashers run on the GPU path"
    );
    match F::ORDER_U64 {
        Goldilocks::ORDER_U64 => (P2G_FIELD_GOLDILOCKS, P2G_HASH_POSEIDON_GL12, P2G_REPR_CANONICAL),
        BabyBear::ORDER_U64 => (P2G_FIELD_BABYBEAR, P2G_HASH_POSEIDON2_BB16, P2G_REPR_MONTY_R32),
        _ => panic!("cuda feature: unsupported field"),
    }
}

/// Pinned staging memory for the big device->host transfer (the leaf matrix: 9 GB at 2^20 x 135).  `MerkleTree.leaves`
/// is a `Vec<Vec<F>>`, so the rows are copied out of a flat buffer anyway; letting the GPU write that flat buffer into
/// PINNED memory makes the transfer run at the link's speed (measured on B200: 57 GB/s against ~15 GB/s into pageable
/// memory, bench.py `e2e` / `e2e_pageable`).  Grow-only, reused across commits (pinning is slow), one per process.
struct Staging {
    ptr: *mut c_void,
    bytes: usize,
}
unsafe impl Send for Staging {}
static STAGING: std::sync::Mutex<Staging> = std::sync::Mutex::new(Staging { ptr: core::ptr::null_mut(), bytes: 0 });

fn with_staging<R>(bytes: usize, f: impl FnOnce(*mut c_void) -> R) -> R {
    let mut st = STAGING.lock().unwrap();
    if st.bytes < bytes {
        let ctx = context();
        unsafe {
            if !st.ptr.is_null() {
                p2g_host_free(ctx, st.ptr);
            }
            st.ptr = p2g_host_alloc(ctx, bytes);
        }
        assert!(!st.ptr.is_null(), "{}", last_error(ctx));
        st.bytes = bytes;
    }
    f(st.ptr)
}

/// `len` uninitialised elements the library is about to fill (every element is written before the call returns).
fn uninit_vec<T>(len: usize) -> Vec<T> {
    let mut v = Vec::with_capacity(len);
    unsafe { v.set_len(len) };
    v
}

pub(crate) fn commit<
    F: RichField + HasExtension<D>,
    C: GenericConfig<D, NUM_HASH_OUT_ELTS, F = F, FE = F::Extension> + 'static,
    const D: usize,
    const NUM_HASH_OUT_ELTS: usize,
>(
    cols: Vec<Vec<F>>,
    input_kind: i32,
    rate_bits: usize,
    blinding: bool,
    cap_height: usize,
    timing: &mut ProvingProcessInfo,
) -> PolynomialBatch<F, C, D, NUM_HASH_OUT_ELTS>
where
    C::Hasher: 'static,
{
    type Hash<F, C, const D: usize, const N: usize> = <<C as GenericConfig<D, N>>::Hasher as Hasher<F>>::Hash;
    let (field, hash, repr_) = dispatch::<F, C::Hasher>();
    let degree = cols[0].len(); // panics on an empty batch like polynomials[0] at oracle.rs:101
    for c in &cols {
        assert_eq!(c.len(), degree, "Polynomial degrees inconsistent"); // oracle.rs:139
    }
    let degree_log = log2_strict(degree);
    let n_total = degree << rate_bits;
    assert!(cap_height <= degree_log + rate_bits,
        "cap_height={} should be at most log2(leaves.len())={}", cap_height, degree_log + rate_bits); // merkle_tree.rs:153-157
    // oracle.rs:144-148: SALT_SIZE random columns of N elements; the RNG stays on this side of the boundary
    let salts: Vec<Vec<F>> = if blinding { (0..SALT_SIZE).map(|_| F::rand_vec(n_total)).collect() } else { Vec::new() };
    let width = cols.len() + salts.len();
    let num_cap = 1usize << cap_height;
    let num_digests = 2 * (n_total - num_cap);

    let mut coeffs: Vec<Vec<F>> =
        if input_kind == P2G_INPUT_VALUES { cols.iter().map(|_| uninit_vec::<F>(degree)).collect() } else { Vec::new() };
    let mut digests: Vec<Hash<F, C, D, NUM_HASH_OUT_ELTS>> = uninit_vec(num_digests);
    let mut cap: Vec<Hash<F, C, D, NUM_HASH_OUT_ELTS>> = uninit_vec(num_cap);
    // HashOut<F, N> is N field elements, 32 bytes for both configurations (hash_types.rs:105-108)
    assert_eq!(core::mem::size_of::<Hash<F, C, D, NUM_HASH_OUT_ELTS>>(), 32);

    let col_ptrs: Vec<*const c_void> = cols.iter().map(|c| c.as_ptr() as *const c_void).collect();
    let salt_ptrs: Vec<*const c_void> = salts.iter().map(|s| s.as_ptr() as *const c_void).collect();
    let coeff_ptrs: Vec<*mut c_void> = coeffs.iter_mut().map(|c| c.as_mut_ptr() as *mut c_void).collect();
    let leaf_bytes = n_total * width * core::mem::size_of::<F>();
    let (desc, leaves) = with_staging(leaf_bytes, |leaves_flat| {
    let mut desc = p2g_commit_desc {
        field,
        hash,
        repr_,
        input_kind,
        log_n: degree_log as u32,
        num_cols: cols.len() as u32,
        rate_bits: rate_bits as u32,
        cap_height: cap_height as u32,
        num_salt_cols: salts.len() as u32,
        flags: 0,
        col_ptrs: col_ptrs.as_ptr(),
        salt_ptrs: if salts.is_empty() { core::ptr::null() } else { salt_ptrs.as_ptr() },
        coeffs_out: if coeffs.is_empty() { core::ptr::null() } else { coeff_ptrs.as_ptr() },
        leaves_out: leaves_flat,
        leaf_pitch: width as u64,
        digests_out: digests.as_mut_ptr() as *mut c_void,
        cap_out: cap.as_mut_ptr() as *mut c_void,
        shard_index: 0,
        shard_count: 0,
        stage_seconds: [0.0; 4],
        h2d_seconds: 0.0,
        d2h_seconds: 0.0,
        kernel_seconds: [0.0; 4],
        batch_out: core::ptr::null_mut(),
    };
    let ctx = context();
    let rc = unsafe { p2g_commit(ctx, &mut desc) };
    if rc != P2G_OK {
        panic!("{}", last_error(ctx));
    }
    // MerkleTree.leaves is Vec<Vec<F>> (merkle_tree.rs:49): one row per leaf, copied out of the pinned staging buffer
    // on all cores (the reference's own `transpose`, plonky2/src/util/mod.rs:25-31, builds the same rows)
    let flat: &[F] = unsafe { core::slice::from_raw_parts(leaves_flat as *const F, n_total * width) };
    let leaves: Vec<Vec<F>> = flat.par_chunks_exact(width).map(|r| r.to_vec()).collect();
    (desc, leaves)
    });
    // the four scopes the CPU path opens with timed! (oracle.rs:76-114); the device reports their durations
    let _ = &desc.stage_seconds; // (a ProvingProcessInfo scope measures wall clock; the device times are in `desc` for callers that log them)
    for name in ["IFFT", "FFT + blinding", "transpose LDEs", "build Merkle tree"] {
        if name == "IFFT" && input_kind != P2G_INPUT_VALUES {
            continue;
        }
        timing.push(name, log::Level::Debug);
        timing.pop();
    }
    let polynomials: Vec<PolynomialCoeffs<F>> = if input_kind == P2G_INPUT_VALUES {
        coeffs.into_iter().map(PolynomialCoeffs::new).collect()
    } else {
        cols.into_iter().map(PolynomialCoeffs::new).collect()
    };
    PolynomialBatch {
        polynomials,
        merkle_tree: MerkleTree { leaves, digests, cap: MerkleCap(cap) },
        degree_log,
        rate_bits,
        blinding,
    }
}

pub(crate) fn merkle_tree_new<F: RichField, H: Hasher<F> + 'static>(leaves: Vec<Vec<F>>, cap_height: usize) -> MerkleTree<F, H> {
    let (field, hash, repr_) = dispatch::<F, H>();
    let log2_leaves_len = log2_strict(leaves.len()); // panics on a non-power-of-two, util/src/lib.rs:30-33
    assert!(cap_height <= log2_leaves_len,
        "cap_height={} should be at most log2(leaves.len())={}", cap_height, log2_leaves_len); // merkle_tree.rs:153-157
    let leaf_len = leaves[0].len();
    // every tree on the commit path has rows of one length (LDE rows, FRI layer chunks); a ragged tree is not this path
    assert!(leaves.iter().all(|l| l.len() == leaf_len), "cuda feature: MerkleTree::new needs rows of equal length");
    let flat: Vec<F> = leaves.iter().flat_map(|l| l.iter().copied()).collect();
    let num_cap = 1usize << cap_height;
    let mut digests: Vec<H::Hash> = uninit_vec(2 * (leaves.len() - num_cap));
    let mut cap: Vec<H::Hash> = uninit_vec(num_cap);
    assert_eq!(core::mem::size_of::<H::Hash>(), 32);
    let mut desc = p2g_merkle_desc {
        field,
        hash,
        repr_,
        flags: 0,
        num_leaves: leaves.len() as u64,
        leaf_len: leaf_len as u64,
        leaf_pitch: leaf_len as u64,
        cap_height: cap_height as u32,
        leaves: flat.as_ptr() as *const c_void,
        digests_out: digests.as_mut_ptr() as *mut c_void,
        cap_out: cap.as_mut_ptr() as *mut c_void,
        seconds: 0.0,
    };
    let ctx = context();
    let rc = unsafe { p2g_merkle(ctx, &mut desc) };
    if rc != P2G_OK {
        panic!("{}", last_error(ctx));
    }
    MerkleTree { leaves, digests, cap: MerkleCap(cap) }
}

// ---------------------------------------------------------------------------------------------------------------------
// plonky2/src/fri/oracle.rs -- inside `impl PolynomialBatch`, next to the CPU bodies (which get #[cfg(not(feature = "cuda"))])
// ---------------------------------------------------------------------------------------------------------------------
#[cfg(feature = "cuda")]
pub fn from_values(
    values: Vec<PolynomialValues<F>>,
    rate_bits: usize,
    blinding: bool,
    cap_height: usize,
    timing: &mut ProvingProcessInfo,
    _fft_root_table: Option<&FftRootTable<F>>, // twiddles are resident on the device
) -> Self {
    crate::fri::cuda::commit::<F, C, D, NUM_HASH_OUT_ELTS>(
        values.into_iter().map(|v| v.values).collect(),
        p2g_sys::P2G_INPUT_VALUES,
        rate_bits,
        blinding,
        cap_height,
        timing,
    )
}

#[cfg(feature = "cuda")]
pub fn from_coeffs(
    polynomials: Vec<PolynomialCoeffs<F>>,
    rate_bits: usize,
    blinding: bool,
    cap_height: usize,
    timing: &mut ProvingProcessInfo,
    _fft_root_table: Option<&FftRootTable<F>>,
) -> Self {
    crate::fri::cuda::commit::<F, C, D, NUM_HASH_OUT_ELTS>(
        polynomials.into_iter().map(|p| p.coeffs).collect(),
        p2g_sys::P2G_INPUT_COEFFS,
        rate_bits,
        blinding,
        cap_height,
        timing,
    )
}

// ---------------------------------------------------------------------------------------------------------------------
// plonky2/src/hash/merkle_tree.rs -- inside `impl<F: RichField, H: Hasher<F>> MerkleTree<F, H>`
// ---------------------------------------------------------------------------------------------------------------------
#[cfg(feature = "cuda")]
pub fn new(leaves: Vec<Vec<F>>, cap_height: usize) -> Self
where
    H: 'static,
{
    crate::fri::cuda::merkle_tree_new::<F, H>(leaves, cap_height)
}
