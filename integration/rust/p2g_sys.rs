// Bindings to libp2g_cuda.so (include/p2g.h). Written against the ABI; not compiled in this repository's image.
// p2g-sys/src/lib.rs  -- mirrors include/p2g.h one to one
#[repr(C)] pub struct p2g_ctx { _private: [u8; 0] }
#[repr(C)] pub struct p2g_batch { _private: [u8; 0] }

pub const P2G_FLAG_DEVICE_INPUT: u32 = 1;
pub const P2G_FLAG_DEVICE_OUTPUT: u32 = 2;
pub const P2G_FLAG_NO_LEAVES_OUT: u32 = 4;
pub const P2G_FLAG_KEEP_ON_DEVICE: u32 = 8;
pub const P2G_FLAG_LEAVES_ONLY: u32 = 16;   // commit split over column ranges: stop after the leaves
pub const P2G_FLAG_LEAVES_READY: u32 = 32;  // ... and build the tree over a finished leaf matrix

#[repr(C)]
pub struct p2g_commit_desc {
    pub field: i32, pub hash: i32, pub repr: i32, pub input_kind: i32,
    pub log_n: u32, pub num_cols: u32, pub rate_bits: u32, pub cap_height: u32,
    pub num_salt_cols: u32, pub flags: u32,
    pub col_ptrs: *const *const core::ffi::c_void,
    pub salt_ptrs: *const *const core::ffi::c_void,
    pub coeffs_out: *const *mut core::ffi::c_void,
    pub leaves_out: *mut core::ffi::c_void, pub leaf_pitch: u64,
    pub digests_out: *mut core::ffi::c_void, pub cap_out: *mut core::ffi::c_void,
    pub shard_index: u32, pub shard_count: u32,
    pub stage_seconds: [f64; 4], pub h2d_seconds: f64, pub d2h_seconds: f64,
    pub kernel_seconds: [f64; 4],
    pub batch_out: *mut *mut p2g_batch,   // with P2G_FLAG_KEEP_ON_DEVICE
}
#[repr(C)]
pub struct p2g_merkle_desc {
    pub field: i32, pub hash: i32, pub repr: i32, pub flags: u32,
    pub num_leaves: u64, pub leaf_len: u64, pub leaf_pitch: u64, pub cap_height: u32,
    pub leaves: *const core::ffi::c_void, pub digests_out: *mut core::ffi::c_void,
    pub cap_out: *mut core::ffi::c_void, pub seconds: f64,
}
extern "C" {
    pub fn p2g_init(device_ids: *const i32, n_devices: i32, out: *mut *mut p2g_ctx) -> i32;
    pub fn p2g_destroy(ctx: *mut p2g_ctx);
    pub fn p2g_last_error(ctx: *mut p2g_ctx) -> *const core::ffi::c_char;
    pub fn p2g_commit(ctx: *mut p2g_ctx, desc: *mut p2g_commit_desc) -> i32;
    pub fn p2g_merkle(ctx: *mut p2g_ctx, desc: *mut p2g_merkle_desc) -> i32;
    pub fn p2g_host_alloc(ctx: *mut p2g_ctx, bytes: usize) -> *mut core::ffi::c_void;
    pub fn p2g_host_free(ctx: *mut p2g_ctx, ptr: *mut core::ffi::c_void);
    // device-resident batches (get_lde_values / MerkleTree::get / prove served from HBM)
    pub fn p2g_batch_get_rows(ctx: *mut p2g_ctx, b: *mut p2g_batch, rows: *const u64, count: u64, out: *mut core::ffi::c_void) -> i32;
    pub fn p2g_batch_prove(ctx: *mut p2g_ctx, b: *mut p2g_batch, leaves: *const u64, count: u64, siblings_out: *mut core::ffi::c_void) -> i32;
    pub fn p2g_batch_download(ctx: *mut p2g_ctx, b: *mut p2g_batch, what: i32, out: *mut core::ffi::c_void) -> i32;
    pub fn p2g_batch_free(ctx: *mut p2g_ctx, b: *mut p2g_batch);
}
// build.rs: println!("cargo:rustc-link-search=native={}", env::var("P2G_LIB_DIR")?);
//           println!("cargo:rustc-link-lib=dylib=p2g_cuda");
