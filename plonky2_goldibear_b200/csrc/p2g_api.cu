// C ABI + host orchestration of the commit path (see include/p2g.h for the contract).
//
// PolynomialBatch::from_values / from_coeffs (plonky2/src/fri/oracle.rs:68-123) become:
//   H2D columns -> [iNTT passes] -> coefficients -> per LDE coset: [scale + NTT passes -> leaf rows] ->
//   leaf sponge -> tree levels -> D2H.   No CPU arithmetic on the data path; the host only builds small
//   twiddle tables (cached per context) and launches kernels.
#include <cuda_runtime.h>

#include <algorithm>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <tuple>
#include <vector>

#include "../../include/p2g.h"
#include "p2g_internal.h"

using namespace p2g;

namespace {

thread_local std::string g_init_error;

struct Buffer {
    void *ptr = nullptr;
    size_t bytes = 0;
};

struct FieldInfo {
    int id;
    size_t elem;          // bytes per element
    unsigned two_adicity;
    uint64_t generator;      // canonical
    uint64_t two_adic_root;  // canonical, order 2^two_adicity
    unsigned digest_elems;
};
// Goldilocks: generator 7, two_adic_generator(32) = 1753635133440165772; BabyBear: generator 31,
// two_adic_generator(27) = 0x1a427a41 -- values of the external Plonky3 crates the reference links
// (call sites field/src/fft.rs:16, fri/oracle.rs:141); see SURVEY.md 8(c) for how each is (un)pinned.
const FieldInfo kGl = {FIELD_GL, 8, 32, 7ull, 1753635133440165772ull, 4};
const FieldInfo kBb = {FIELD_BB, 4, 27, 31ull, 0x1a427a41ull, 8};

uint64_t f_mul(const FieldInfo &f, uint64_t a, uint64_t b) {
    return f.id == FIELD_GL ? GlHost::mul(a, b) : BbHost::mul((u32)a, (u32)b);
}
uint64_t f_pow(const FieldInfo &f, uint64_t a, uint64_t e) {
    return f.id == FIELD_GL ? GlHost::pow(a, e) : BbHost::pow((u32)a, e);
}
uint64_t f_inv(const FieldInfo &f, uint64_t a) { return f.id == FIELD_GL ? GlHost::inv(a) : BbHost::inv((u32)a); }
uint64_t f_internal(const FieldInfo &f, uint64_t a) {
    return f.id == FIELD_GL ? GlHost::to_internal(a) : BbHost::to_internal((u32)a);
}
uint64_t two_adic_generator(const FieldInfo &f, unsigned bits) {
    uint64_t g = f.two_adic_root;
    for (unsigned i = bits; i < f.two_adicity; i++) g = f_mul(f, g, g);
    return g;
}
unsigned bitrev(unsigned x, unsigned bits) {
    unsigned r = 0;
    for (unsigned i = 0; i < bits; i++) r |= ((x >> i) & 1u) << (bits - 1 - i);
    return r;
}

}  // namespace

// Device-resident result of one commit (P2G_FLAG_KEEP_ON_DEVICE): owns its buffers.
struct p2g_batch {
    size_t leaves_bytes = 0, digests_bytes = 0, cap_bytes = 0, coeffs_bytes = 0;
    int field = 0;
    size_t elem = 0;
    uint64_t rows = 0, width = 0, num_cols = 0, n = 0;
    uint32_t log_leaves = 0, cap_height = 0;
    uint64_t num_digests = 0, num_cap = 0;
    void *leaves = nullptr, *digests = nullptr, *cap = nullptr, *coeffs = nullptr;
    // result of a commit on a multi-device context: one device-resident batch per device (coset shard d = leaf rows
    // [d * rows, (d + 1) * rows), the matching digest slice and cap entries); `rows` etc. above then describe ONE shard
    std::vector<p2g_batch *> shards;
};

// State of one FRI commit phase between layers (fri/prover.rs:83-133): the coefficient vector of the current
// codeword, one extension coordinate per column, and its evaluations in bit-reversed order.
struct p2g_fri {
    int field = 0, hash = 0, repr = 0;
    unsigned degree = 0;    // D, base elements per extension element
    size_t elem = 0;
    unsigned log_n = 0;     // current codeword length
    uint64_t shift = 0;     // coset shift of the current evaluations (canonical)
    void *coeffs = nullptr; // [D][n]
    void *values = nullptr; // [n][D], bit-reversed order; null once a layer tree owns it
    size_t coeffs_bytes = 0, values_bytes = 0;
    int pending_arity_bits = -1;  // set by commit_layer, consumed by fold
};

struct p2g_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;  // D2H of finished outputs overlaps the hashing kernels
    cudaEvent_t ev[8];
    cudaStream_t h2d_stream = nullptr;   // H2D of the next column group overlaps the transforms of the current one
    cudaEvent_t ready[64];               // "rows are final" markers for the copy stream
    enum { MAX_GROUPS = 32 };
    cudaEvent_t gready[MAX_GROUPS + 1], gdone[MAX_GROUPS];  // column group arrived / its coefficients are final
    cudaEvent_t tev[3 * MAX_GROUPS];     // stage timing per column group
    std::mutex mu;
    std::string error;
    uint64_t launches = 0;
    // workspace (grow-only)
    Buffer in, coeffs, tmp, leaves, digests, cap, salts, misc;
    // cached device tables
    std::map<std::pair<int, uint64_t>, void *> pow_tables;   // (field, base) -> [3][1024]
    std::map<std::pair<int, uint64_t>, void *> root_tables;  // (field, (inverse<<8)|lp) -> [2^(lp-1)]
    // per-pass factor tables built on the device: (field, kind, log_min, log_np, w base, s base) -> table
    std::map<std::tuple<int, int, unsigned, unsigned, uint64_t, uint64_t>, void *> pass_tables;
    size_t pass_table_bytes = 0;
    // buffers of freed device-resident batches, reused by the next kept commit (cudaMalloc / cudaFree of multi-GB
    // blocks costs ~100 ms and synchronises the device)
    std::vector<Buffer> pool;
    size_t pool_bytes = 0;
    // pinned staging ring for large copies from / to PAGEABLE host memory (two slots, see staged_copy)
    void *stage_ring = nullptr;
    cudaEvent_t stage_ev[2] = {nullptr, nullptr};
    // multi-device context (p2g_init with n_devices > 1): this object only coordinates; every device has its own
    // single-device context, a gather buffer that receives all coefficient columns over NVLink, and an event that says
    // "this device's share of the inverse NTT has been stored everywhere"
    std::vector<p2g_ctx *> subs;
    std::vector<Buffer> gather;
    std::vector<cudaEvent_t> intt_done;

    int fail(int code, const char *fmt, ...) {
        char buf[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof buf, fmt, ap);
        va_end(ap);
        error = buf;
        return code;
    }
};

namespace {

#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess)                                                                           \
            return ctx->fail(e_ == cudaErrorMemoryAllocation ? P2G_ERR_OOM : P2G_ERR_CUDA, "%s failed: %s", \
                             #call, cudaGetErrorString(e_));                                             \
    } while (0)

int ensure(p2g_ctx *ctx, Buffer &b, size_t bytes) {
    if (b.bytes >= bytes) return 0;
    if (b.ptr) {
        cudaStreamSynchronize(ctx->stream);
        cudaFree(b.ptr);
        b.ptr = nullptr;
        b.bytes = 0;
    }
    if (bytes == 0) return 0;
    cudaError_t e = cudaMalloc(&b.ptr, bytes);
    if (e != cudaSuccess && !ctx->pool.empty()) {
        // buffers of freed device-resident batches are only parked: give them back and try once more
        cudaGetLastError();
        cudaStreamSynchronize(ctx->stream);
        for (auto &pb : ctx->pool) cudaFree(pb.ptr);
        ctx->pool.clear();
        ctx->pool_bytes = 0;
        e = cudaMalloc(&b.ptr, bytes);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        b.ptr = nullptr;
        return ctx->fail(P2G_ERR_OOM, "cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
    }
    b.bytes = bytes;
    return 0;
}

void store_elem(const FieldInfo &f, void *base, size_t idx, uint64_t internal_value) {
    if (f.id == FIELD_GL)
        ((uint64_t *)base)[idx] = internal_value;
    else
        ((uint32_t *)base)[idx] = (uint32_t)internal_value;
}

// [3][1024] table of base^(k * 1024^level), internal representation
int get_pow_table(p2g_ctx *ctx, const FieldInfo &f, uint64_t base, const void **out) {
    auto key = std::make_pair(f.id, base);
    auto it = ctx->pow_tables.find(key);
    if (it != ctx->pow_tables.end()) {
        *out = it->second;
        return 0;
    }
    std::vector<unsigned char> host(POW_LEVELS * POW_DIGIT * f.elem);
    uint64_t step = base;
    for (int lvl = 0; lvl < POW_LEVELS; lvl++) {
        uint64_t cur = 1;
        for (int k = 0; k < POW_DIGIT; k++) {
            store_elem(f, host.data(), (size_t)lvl * POW_DIGIT + k, f_internal(f, cur));
            cur = f_mul(f, cur, step);
        }
        step = cur;  // base^(1024^(lvl+1))
    }
    void *d = nullptr;
    CK(cudaMalloc(&d, host.size()));
    CK(cudaMemcpyAsync(d, host.data(), host.size(), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->pow_tables[key] = d;
    *out = d;
    return 0;
}

// In-tile twiddles of a 2^lp-point tile, w the primitive 2^lp-th root (or its inverse).  BabyBear: w^k, k < 2^(lp-1).
// Goldilocks: [2^lp][4], entry k = w^k * 2^(24 i), i < 4 -- the four W_i a limb join multiplies by (gl_limbs.cuh)
int get_root_table(p2g_ctx *ctx, const FieldInfo &f, unsigned lp, bool inverse, const void **out) {
    auto key = std::make_pair(f.id, ((uint64_t)inverse << 8) | lp);
    auto it = ctx->root_tables.find(key);
    if (it != ctx->root_tables.end()) {
        *out = it->second;
        return 0;
    }
    const bool gl = f.id == FIELD_GL;
    // Goldilocks: [np][4] join entries; BabyBear: [np/2] pairs (w, floor(w 2^32 / p)), w canonical (NttBb::mul_root)
    size_t cnt = gl ? ((size_t)4 << lp) : 2 * (lp ? ((size_t)1 << (lp - 1)) : 1);
    std::vector<unsigned char> host(cnt * f.elem);
    uint64_t w = two_adic_generator(f, lp);
    if (inverse) w = f_inv(f, w);
    uint64_t cur = 1;
    if (gl) {
        const uint64_t x24 = (uint64_t)1 << 24;
        for (size_t k = 0; k < ((size_t)1 << lp); k++) {
            uint64_t v = cur;
            for (int i = 0; i < 4; i++) {
                store_elem(f, host.data(), 4 * k + i, v);
                v = f_mul(f, v, x24);
            }
            cur = f_mul(f, cur, w);
        }
    } else {
        for (size_t k = 0; k < cnt / 2; k++) {
            store_elem(f, host.data(), 2 * k, cur);
            store_elem(f, host.data(), 2 * k + 1, (cur << 32) / P2G_BB_P);
            cur = f_mul(f, cur, w);
        }
    }
    void *d = nullptr;
    CK(cudaMalloc(&d, host.size()));
    CK(cudaMemcpyAsync(d, host.data(), host.size(), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->root_tables[key] = d;
    *out = d;
    return 0;
}

void *pool_alloc(p2g_ctx *ctx, size_t bytes) {
    bytes = std::max<size_t>(bytes, 32);
    int best = -1;
    for (size_t i = 0; i < ctx->pool.size(); i++)
        if (ctx->pool[i].bytes >= bytes && ctx->pool[i].bytes <= bytes + bytes / 2 + 4096 &&
            (best < 0 || ctx->pool[i].bytes < ctx->pool[best].bytes))
            best = (int)i;
    if (best >= 0) {
        void *p = ctx->pool[best].ptr;
        ctx->pool_bytes -= ctx->pool[best].bytes;
        ctx->pool.erase(ctx->pool.begin() + best);
        return p;
    }
    void *p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess) {
        cudaGetLastError();
        // make room and retry once
        for (auto &b : ctx->pool) cudaFree(b.ptr);
        ctx->pool.clear();
        ctx->pool_bytes = 0;
        if (cudaMalloc(&p, bytes) != cudaSuccess) {
            cudaGetLastError();
            return nullptr;
        }
    }
    return p;
}
void pool_free(p2g_ctx *ctx, void *p, size_t bytes) {
    if (!p) return;
    bytes = std::max<size_t>(bytes, 32);
    if (!ctx || ctx->pool_bytes + bytes > ((size_t)48 << 30) || ctx->pool.size() >= 96) {
        cudaFree(p);
        return;
    }
    ctx->pool.push_back(Buffer{p, bytes});
    ctx->pool_bytes += bytes;
}

void drop_pass_tables(p2g_ctx *ctx) {
    cudaStreamSynchronize(ctx->stream);
    for (auto &kv : ctx->pass_tables) cudaFree(kv.second);
    ctx->pass_tables.clear();
    ctx->pass_table_bytes = 0;
}

// kind 0: tw_table[m_in] = w^(f*j0) (* s^j0 when sbase != 0); kind 1: scale_table[np] = s^(i << log_mout)
int get_pass_table(p2g_ctx *ctx, const FieldInfo &f, int kind, unsigned log_min, unsigned log_np, uint64_t wbase,
                   uint64_t sbase, const void **out) {
    auto key = std::make_tuple(f.id, kind, log_min, log_np, wbase, sbase);
    auto it = ctx->pass_tables.find(key);
    if (it != ctx->pass_tables.end()) {
        *out = it->second;
        return 0;
    }
    size_t bytes = ((size_t)1 << (kind == 0 ? log_min : log_np)) * f.elem;
    const void *wpow = nullptr, *spow = nullptr;
    int rc;
    if (kind == 0 && (rc = get_pow_table(ctx, f, wbase, &wpow))) return rc;
    if (sbase && (rc = get_pow_table(ctx, f, sbase, &spow))) return rc;
    void *d = nullptr;
    CK(cudaMalloc(&d, bytes));
    if (kind == 0)
        launch_build_tw_table(f.id, d, log_min, log_np, wpow, spow, ctx->stream, &ctx->launches);
    else
        launch_build_scale_table(f.id, d, log_np, log_min - log_np, spow, ctx->stream, &ctx->launches);
    CK(cudaGetLastError());
    ctx->pass_tables[key] = d;
    ctx->pass_table_bytes += bytes;
    *out = d;
    return 0;
}

// split log_n into at most three tile sizes, each <= MAX_TILE_LOG, as evenly as possible
void plan_passes(unsigned log_n, unsigned lps[3], unsigned *num) {
    // test hook: P2G_TEST_MAX_TILE_LOG=k (1 <= k <= MAX_TILE_LOG) forces smaller tiles, so that the two- and three-pass
    // plans can be exercised at sizes the CPU oracle checks in seconds
    unsigned max_tile = MAX_TILE_LOG;
    if (const char *e = getenv("P2G_TEST_MAX_TILE_LOG")) {
        int k = atoi(e);
        if (k >= 1 && k <= (int)MAX_TILE_LOG) max_tile = (unsigned)k;
    }
    unsigned P = log_n <= max_tile ? 1 : (log_n <= 2 * max_tile ? 2 : 3);
    unsigned rem = log_n;
    for (unsigned i = 0; i < P; i++) {
        unsigned take = (rem + (P - i) - 1) / (P - i);
        lps[i] = take;
        rem -= take;
    }
    for (unsigned i = P; i < 3; i++) lps[i] = 0;
    *num = P;
}

static bool multi_pass(unsigned log_n) {  // the strided passes of a forward transform need the scratch buffer
    unsigned lps[3], P;
    plan_passes(log_n, lps, &P);
    return P > 1;
}

// How the coefficients of a column-sharded inverse transform reach the other devices' gather buffers:
//   0 "fused": the last pass stores every 64-byte run into the peers' buffers as well (NttPass::peer_dst) -- the transfer
//      hides behind the butterflies while there are enough of them (2 devices: 67 columns each);
//   1 "push": the last pass stores locally and a copy kernel sends the finished block (whole columns, contiguous) with
//      16 bytes per lane, 512 contiguous bytes per warp instruction -- with 4 or 8 devices a device transforms 34 / 17
//      columns and sends 3 / 7 times that, the kernel is store-bound, and 64-byte runs to seven targets reach ~500 GB/s
//      of the link's 900;
//   2 "ce": the same block through the copy engines, one copy per peer.
// P2G_EXCHANGE=fused|push|ce overrides the choice.
static int exchange_mode(unsigned num_peers, bool sendable) {  // sendable: one contiguous, 16-byte aligned block
    if (!num_peers) return 0;
    int mode = num_peers >= 3 ? 1 : 0;
    if (const char *e = getenv("P2G_EXCHANGE")) {
        if (!strcmp(e, "fused")) mode = 0;
        else if (!strcmp(e, "push")) mode = 1;
        else if (!strcmp(e, "ce")) mode = 2;
    }
    return sendable ? mode : 0;
}

// Runs size-2^log_n transforms over `num_cols` columns.
//   inverse: values -> coefficients (natural order, scaled by 1/n), result in `dst` (column-major, user repr); nb = 1
//   forward: nb coset LDE blocks in one launch per pass: coefficients (user repr, `src`) scaled by powers of
//            shifts[b], rows written to the leaf matrix `dst` starting at row_base + b * row_stride.
// `src` is overwritten when inverse (the strided passes run in place); forward uses `tmp` for the strided passes
// (nb * num_cols * n elements).
// pass_mode (forward only; the split LDE of p2g_lde_columns): 0 = every pass; 1 = the strided passes only, of the
// num_cols columns at `src`, into column slots [tmp_col_base, ...) of a scratch that is tmp_cols_total columns wide per
// coset; 2 = the last pass only, over num_cols = tmp_cols_total columns whose strided passes were run before (`src` is
// read by single-pass sizes only, which have no strided pass).
int run_transform_batch(p2g_ctx *ctx, const FieldInfo &f, int repr, bool inverse, unsigned log_n, unsigned num_cols,
                        void *src, void *tmp, void *dst, uint64_t dst_stride, const uint64_t *shifts, unsigned nb,
                        uint64_t row_base, uint64_t row_stride, void *const *peer_dst = nullptr, unsigned num_peers = 0,
                        int pass_mode = 0, uint64_t tmp_cols_total = 0, uint64_t tmp_col_base = 0) {
    unsigned lps[3], P;
    plan_passes(log_n, lps, &P);
    const uint64_t n = (uint64_t)1 << log_n;
    int xmode = 0;
    if (nb == 0 || nb > 8 || (inverse && nb != 1)) return ctx->fail(P2G_ERR_INVALID, "bad transform batch %u", nb);
    unsigned log_min = log_n;
    const void *cur_src = src;
    uint64_t cur_src_batch_stride = 0;
    void *stage_buf = inverse ? src : tmp;
    const uint64_t stage_cols = tmp_cols_total ? tmp_cols_total : num_cols;  // columns per coset in the scratch
    if (pass_mode && !inverse) stage_buf = (char *)tmp + (size_t)tmp_col_base * n * f.elem;
    for (unsigned i = 0; i < P; i++) {
        NttPass p;
        memset(&p, 0, sizeof p);
        p.num_cols = num_cols;
        p.log_n = log_n;
        p.log_min = log_min;
        p.log_np = lps[i];
        p.inverse = inverse;
        p.gl_eps = 0xffffffffu;
        p.num_batch = nb;
        int rc = get_root_table(ctx, f, lps[i], inverse, &p.roots);
        if (rc) return rc;
        p.src = cur_src;
        p.src_col_stride = n;
        p.src_batch_stride = cur_src_batch_stride;
        p.async_ok = ((uintptr_t)cur_src % 16 == 0) && ((n * f.elem) % 16 == 0 || (num_cols == 1 && nb == 1));
        p.in_user_repr = (i == 0);
        if (i == 0 && !inverse) {
            // coset factor s^j, j = i*m_out + j0: the row part s^(m_out*i) on load, s^j0 merged into the pass table
            for (unsigned b = 0; b < nb; b++)
                if ((rc = get_pass_table(ctx, f, 1, log_min, lps[i], 0, shifts[b], &p.scale_tables[b]))) return rc;
        }
        if (i + 1 < P) {
            // strided pass with inter-pass twiddles w_{m_in}^(f*j0)
            uint64_t w = two_adic_generator(f, log_min);
            if (inverse) w = f_inv(f, w);
            for (unsigned b = 0; b < nb; b++)
                if ((rc = get_pass_table(ctx, f, 0, log_min, lps[i], w, (i == 0 && !inverse) ? shifts[b] : 0, &p.tw_tables[b]))) return rc;
            p.dst = stage_buf;
            p.dst_col_stride = n;
            p.dst_batch_stride = stage_cols * n;
            if (pass_mode != 2) launch_ntt_pass(f.id, repr, NTT_STRIDED, p, ctx->stream, &ctx->launches);
            cur_src = stage_buf;
            cur_src_batch_stride = stage_cols * n;
            log_min -= lps[i];
        } else if (pass_mode == 1) {
            break;  // the last pass follows in a later call, over all columns
        } else {
            p.lp1 = P >= 2 ? lps[0] : 0;
            p.lp2 = P >= 3 ? lps[1] : 0;
            p.dst = dst;
            p.dst_col_stride = dst_stride;
            p.out_user_repr = 1;
            if (inverse) {
                bool sendable = dst_stride == n && ((size_t)num_cols * n * f.elem) % 16 == 0 && (uintptr_t)dst % 16 == 0;
                for (unsigned k = 0; k < num_peers && k < 8; k++) sendable = sendable && (uintptr_t)peer_dst[k] % 16 == 0;
                xmode = exchange_mode(num_peers, sendable);
                if (xmode == 0) {
                    for (unsigned k = 0; k < num_peers && k < 8; k++) p.peer_dst[k] = peer_dst[k];
                    p.num_peers = num_peers;
                }
                const uint64_t n_inv = f_pow(f, f_inv(f, 2), log_n);  // n^-1 = (1/2)^lg n, fft.rs:78
                if (f.id == FIELD_GL) {
                    // n^-1 is a power of two here: it rides in the last round's join (gl_limbs.cuh) for free
                    for (int k = 0; k < 4; k++) p.join_w[k] = f_mul(f, n_inv, f_pow(f, 2, 24 * k));
                } else {
                    p.use_final_scale = 1;
                    p.final_scale = f_internal(f, n_inv);
                }
                launch_ntt_pass(f.id, repr, NTT_FINAL_NATURAL, p, ctx->stream, &ctx->launches);
            } else {
                p.row_base = row_base;
                p.row_batch_stride = row_stride;
                launch_ntt_pass(f.id, repr, NTT_FINAL_LEAVES, p, ctx->stream, &ctx->launches);
            }
        }
    }
    if (inverse && num_peers && xmode) {
        // the coefficient block of this device (num_cols whole columns, contiguous) goes to the peers behind the last pass
        const size_t bytes = (size_t)num_cols * n * f.elem;
        if (xmode == 1) {
            launch_peer_push(dst, peer_dst, num_peers, bytes, ctx->stream, &ctx->launches);
        } else {
            for (unsigned k = 0; k < num_peers; k++) CK(cudaMemcpyAsync(peer_dst[k], dst, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
        }
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return ctx->fail(P2G_ERR_CUDA, "NTT launch failed: %s", cudaGetErrorString(e));
    return 0;
}

int run_transform(p2g_ctx *ctx, const FieldInfo &f, int repr, bool inverse, unsigned log_n, unsigned num_cols,
                  void *src, void *tmp, void *dst, uint64_t dst_stride, uint64_t shift, uint64_t row_base,
                  void *const *peer_dst = nullptr, unsigned num_peers = 0) {
    return run_transform_batch(ctx, f, repr, inverse, log_n, num_cols, src, tmp, dst, dst_stride, &shift, 1, row_base, 0,
                               peer_dst, num_peers);
}

// how many cosets one LDE launch carries: as many as the scratch buffer of the strided passes allows (it holds the
// first pass' output of every coset in flight).  P2G_LDE_BATCH overrides (experiments).
unsigned lde_batch(unsigned log_n, unsigned cosets, size_t bytes_per_coset) {
    unsigned nb = 1;
    const size_t cap = (size_t)10 << 30;
    while (nb * 2 <= cosets && nb * 2 <= 8 && (!multi_pass(log_n) || (size_t)(nb * 2) * bytes_per_coset <= cap)) nb *= 2;
    if (const char *e = getenv("P2G_LDE_BATCH")) {
        int k = atoi(e);
        if (k >= 1 && k <= 8) nb = std::min<unsigned>((unsigned)k, cosets);
    }
    return nb ? nb : 1;
}

int check_field_hash(p2g_ctx *ctx, int field, int hash, const FieldInfo **out) {
    if (field == P2G_FIELD_GOLDILOCKS && hash == P2G_HASH_POSEIDON_GL12) {
        *out = &kGl;
        return 0;
    }
    if (field == P2G_FIELD_BABYBEAR && hash == P2G_HASH_POSEIDON2_BB16) {
        *out = &kBb;
        return 0;
    }
    return ctx->fail(P2G_ERR_UNSUPPORTED, "unsupported field/hash combination (%d, %d): only Goldilocks+Poseidon(w12) "
                                         "and BabyBear+Poseidon2(w16) run on the GPU path", field, hash);
}

// cudaMemcpyAsync from/to pageable memory blocks the calling thread, which would stall the kernel launches behind it:
// only pinned destinations are copied early on the copy stream.
bool is_pinned(const void *p) {
    if (!p) return false;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost;
}

// every buffer of a pointer table (a table mixing pinned and pageable buffers takes the plain path)
bool all_pinned(const void *const *ptrs, unsigned count) {
    if (!ptrs || !count) return false;
    for (unsigned j = 0; j < count; j++)
        if (!is_pinned(ptrs[j])) return false;
    return true;
}

unsigned ilog2(uint64_t x) {
    unsigned l = 0;
    while (((uint64_t)1 << l) < x) l++;
    return l;
}

// ---- large copies between device memory and PAGEABLE host memory -------------------------------------------------------
// cudaMemcpy on pageable memory makes the driver stage every piece through its own small pinned buffer on one thread:
// ~15 GB/s device->host and ~11 GB/s host->device on this box, against 57 / 56 GB/s for pinned memory -- 756 ms instead of
// 199 ms for the outputs of a 2^20 x 135 commit (bench.py e2e_pageable, round-2 first measurement).  Callers that cannot
// pin their buffers (a Vec<F> from the Rust allocator, a numpy array) get the same pipeline inside the library: the
// copy engine moves 128 MB slots between HBM and a two-slot pinned ring at link speed while several host threads copy
// the previous slot to / from the caller's memory.
enum { STAGE_SLOT_LOG = 27 };  // 128 MB
static const size_t kStageSlot = (size_t)1 << STAGE_SLOT_LOG;
static const size_t kStageMin = (size_t)32 << 20;  // smaller copies are not worth the pipeline

struct HostPiece {
    char *ptr;
    size_t bytes;
};
struct CopyTask {
    char *dst;
    const char *src;
    size_t bytes;
};

void run_copy_tasks(std::vector<CopyTask> &tasks) {
    // split into units of at most 4 MB so that the threads balance, then deal them out round robin
    std::vector<CopyTask> units;
    const size_t unit = (size_t)4 << 20;
    for (const CopyTask &t : tasks)
        for (size_t off = 0; off < t.bytes; off += unit) units.push_back(CopyTask{t.dst + off, t.src + off, std::min(unit, t.bytes - off)});
    unsigned hw = std::thread::hardware_concurrency();
    unsigned T = std::max(1u, std::min(8u, hw ? hw / 2 : 2u));
    T = (unsigned)std::min<size_t>(T, units.size());
    auto work = [&](unsigned t) {
        for (size_t k = t; k < units.size(); k += T) memcpy(units[k].dst, units[k].src, units[k].bytes);
    };
    std::vector<std::thread> th;
    for (unsigned t = 1; t < T; t++) th.emplace_back(work, t);
    if (T) work(0);
    for (auto &x : th) x.join();
}

int stage_ring_ready(p2g_ctx *ctx) {
    if (ctx->stage_ring) return 0;
    cudaError_t e = cudaHostAlloc(&ctx->stage_ring, 2 * kStageSlot, cudaHostAllocDefault);
    if (e != cudaSuccess) {
        cudaGetLastError();
        ctx->stage_ring = nullptr;
        return ctx->fail(P2G_ERR_OOM, "cudaHostAlloc of the staging ring failed: %s", cudaGetErrorString(e));
    }
    for (auto &ev : ctx->stage_ev) cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    return 0;
}

// The host side is the concatenation of `pieces` (a caller's columns need not be adjacent), the device side one
// contiguous block.  Synchronous; `stream` must already be ordered behind whatever produces / consumes the device side.
int staged_copy(p2g_ctx *ctx, const std::vector<HostPiece> &pieces, void *dev, bool to_host, cudaStream_t stream) {
    int rc = stage_ring_ready(ctx);
    if (rc) return rc;
    char *ring = (char *)ctx->stage_ring;
    size_t bytes = 0;
    for (const HostPiece &hp : pieces) bytes += hp.bytes;
    if (!bytes) return 0;
    const size_t nchunks = (bytes + kStageSlot - 1) / kStageSlot;
    auto len = [&](size_t i) { return std::min(kStageSlot, bytes - i * kStageSlot); };
    // host <-> ring for chunk i: walk the pieces that overlap [i * slot, i * slot + len)
    size_t piece = 0, piece_start = 0;  // cursor: pieces[piece] starts at byte piece_start of the concatenation
    auto host_tasks = [&](size_t i, std::vector<CopyTask> &tasks) {
        const size_t lo = i * kStageSlot, hi = lo + len(i);
        while (piece < pieces.size() && piece_start + pieces[piece].bytes <= lo) piece_start += pieces[piece++].bytes;
        size_t p = piece, ps = piece_start;
        char *slot = ring + (i & 1) * kStageSlot;
        while (p < pieces.size() && ps < hi) {
            const size_t a = std::max(lo, ps), b = std::min(hi, ps + pieces[p].bytes);
            if (b > a) {
                char *h = pieces[p].ptr + (a - ps), *r = slot + (a - lo);
                tasks.push_back(to_host ? CopyTask{h, r, b - a} : CopyTask{r, h, b - a});
            }
            ps += pieces[p++].bytes;
        }
    };
    std::vector<CopyTask> tasks;
    if (to_host) {
        CK(cudaMemcpyAsync(ring, dev, len(0), cudaMemcpyDeviceToHost, stream));
        CK(cudaEventRecord(ctx->stage_ev[0], stream));
        for (size_t i = 0; i < nchunks; i++) {
            if (i + 1 < nchunks) {  // the other slot was drained in the previous iteration
                CK(cudaMemcpyAsync(ring + ((i + 1) & 1) * kStageSlot, (char *)dev + (i + 1) * kStageSlot, len(i + 1), cudaMemcpyDeviceToHost, stream));
                CK(cudaEventRecord(ctx->stage_ev[(i + 1) & 1], stream));
            }
            CK(cudaEventSynchronize(ctx->stage_ev[i & 1]));
            tasks.clear();
            host_tasks(i, tasks);
            run_copy_tasks(tasks);
        }
    } else {
        for (size_t i = 0; i < nchunks; i++) {
            if (i >= 2) CK(cudaEventSynchronize(ctx->stage_ev[i & 1]));  // the upload that last used this slot has left it
            tasks.clear();
            host_tasks(i, tasks);
            run_copy_tasks(tasks);
            CK(cudaMemcpyAsync((char *)dev + i * kStageSlot, ring + (i & 1) * kStageSlot, len(i), cudaMemcpyHostToDevice, stream));
            CK(cudaEventRecord(ctx->stage_ev[i & 1], stream));
        }
        CK(cudaStreamSynchronize(stream));  // the ring is free again when this returns
    }
    return 0;
}

// device columns [first, last) -> the caller's per-column host buffers; buffers adjacent in host memory share one copy
int copy_columns_home(p2g_ctx *ctx, void *const *dst, unsigned first, unsigned last, const void *src, size_t col_bytes,
                      cudaStream_t stream) {
    for (unsigned j = first; j < last;) {
        if (!dst[j]) {
            j++;
            continue;
        }
        unsigned k = j + 1;
        while (k < last && dst[k] && (char *)dst[k] == (char *)dst[k - 1] + col_bytes) k++;
        CK(cudaMemcpyAsync(dst[j], (const char *)src + (size_t)j * col_bytes, (size_t)(k - j) * col_bytes, cudaMemcpyDeviceToHost, stream));
        j = k;
    }
    return 0;
}

double ev_seconds(cudaEvent_t a, cudaEvent_t b) {
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    return ms * 1e-3;
}

int32_t init_multi(const int32_t *device_ids, int32_t n_devices, p2g_ctx **out_ctx);
int32_t commit_multi(p2g_ctx *ctx, p2g_commit_desc *d);

// a call every thread of a multi-device commit makes once: returns when all of them have arrived
struct ThreadBarrier {
    std::mutex mu;
    std::condition_variable cv;
    unsigned waiting = 0, generation = 0, parties;
    explicit ThreadBarrier(unsigned n) : parties(n) {}
    void arrive_and_wait() {
        std::unique_lock<std::mutex> lock(mu);
        unsigned gen = generation;
        if (++waiting == parties) {
            waiting = 0;
            generation++;
            cv.notify_all();
        } else {
            cv.wait(lock, [&] { return gen != generation; });
        }
    }
};

}  // namespace

// ---- multi-device contexts (p2g_init(device_ids, n) holds everything the coset partition needs) ----------
// One process, one host thread per device for the duration of a call.  Partition: device d runs the inverse
// NTT of its column slice and stores the coefficients into EVERY device's gather buffer from the transform's epilogue
// (peer stores over NVLink, run_transform's peer list); an event per device + a host-side rendezvous is the cross-device
// barrier (each stream then waits on every other device's event: no collective library); then device d extends and
// hashes its own cosets [d * 2^r / G, ...) and writes leaf rows [d * N / G, ...), the matching digest slice and cap
// entries straight into the caller's buffers over its own PCIe link.  Nothing else crosses devices.
namespace {

int32_t init_multi(const int32_t *device_ids, int32_t n_devices, p2g_ctx **out_ctx) {
    p2g_ctx *ctx = new p2g_ctx();
    ctx->device = device_ids[0];
    for (int d = 0; d < n_devices; d++) {
        p2g_ctx *sub = nullptr;
        int32_t rc = p2g_init(&device_ids[d], 1, &sub);
        if (rc != P2G_OK) {
            for (p2g_ctx *s2 : ctx->subs) p2g_destroy(s2);
            delete ctx;
            return rc;  // g_init_error was set by the failing single-device init
        }
        ctx->subs.push_back(sub);
    }
    // every device stores into every other device's gather buffer (the same device may be listed more than once: its
    // shards then share it, which is how the partition is exercised on a one-GPU box)
    for (int a = 0; a < n_devices; a++)
        for (int b = 0; b < n_devices; b++) {
            if (device_ids[a] == device_ids[b]) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, device_ids[a], device_ids[b]);
            if (!can) {
                g_init_error = "p2g_init: devices " + std::to_string(device_ids[a]) + " and " + std::to_string(device_ids[b]) +
                               " cannot access each other's memory (no NVLink / PCIe peer path)";
                for (p2g_ctx *s2 : ctx->subs) p2g_destroy(s2);
                delete ctx;
                return P2G_ERR_UNSUPPORTED;
            }
            cudaSetDevice(device_ids[a]);
            cudaError_t e = cudaDeviceEnablePeerAccess(device_ids[b], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
                g_init_error = std::string("p2g_init: cudaDeviceEnablePeerAccess failed: ") + cudaGetErrorString(e);
                for (p2g_ctx *s2 : ctx->subs) p2g_destroy(s2);
                delete ctx;
                return P2G_ERR_CUDA;
            }
            cudaGetLastError();
        }
    ctx->gather.resize(n_devices);
    ctx->intt_done.resize(n_devices);
    for (int d = 0; d < n_devices; d++) {
        cudaSetDevice(device_ids[d]);
        cudaEventCreateWithFlags(&ctx->intt_done[d], cudaEventDisableTiming);
    }
    *out_ctx = ctx;
    return P2G_OK;
}

int32_t commit_multi(p2g_ctx *ctx, p2g_commit_desc *d) {
    std::lock_guard<std::mutex> lock(ctx->mu);
    const unsigned G = (unsigned)ctx->subs.size(), lg = ilog2(G);
    const FieldInfo *fp;
    int rc = check_field_hash(ctx, d->field, d->hash, &fp);
    if (rc) return rc;
    const FieldInfo &f = *fp;
    const int repr = f.id == FIELD_BB ? d->repr : 0;
    const unsigned allowed = P2G_FLAG_NO_LEAVES_OUT | P2G_FLAG_KEEP_ON_DEVICE | P2G_FLAG_DEVICE_INPUT;
    if (d->flags & ~allowed)
        return ctx->fail(P2G_ERR_UNSUPPORTED, "multi-device context: only NO_LEAVES_OUT, KEEP_ON_DEVICE and DEVICE_INPUT "
                                              "(one column-slice pointer per device) are supported");
    if (d->shard_count) return ctx->fail(P2G_ERR_INVALID, "multi-device context: the library shards the job itself (shard_count must be 0)");
    if (d->num_cols == 0) return ctx->fail(P2G_ERR_INVALID, "empty polynomial batch (the reference panics on polynomials[0])");
    if (!d->col_ptrs) return ctx->fail(P2G_ERR_INVALID, "col_ptrs is NULL");
    if (d->input_kind != P2G_INPUT_VALUES && d->input_kind != P2G_INPUT_COEFFS) return ctx->fail(P2G_ERR_INVALID, "bad input_kind");
    const unsigned L = d->log_n, r = d->rate_bits, LN = L + r, h = d->cap_height;
    if (lg > r) return ctx->fail(P2G_ERR_UNSUPPORTED, "%u devices exceed the 2^rate_bits = %u cosets", G, 1u << r);
    if (h > LN) return ctx->fail(P2G_ERR_INVALID, "cap_height=%u should be at most log2(leaves.len())=%u", h, LN);
    if (h < lg)
        return ctx->fail(P2G_ERR_UNSUPPORTED, "multi-device context: cap_height %u < log2(devices) %u leaves cap entries that span "
                                              "devices; commit per shard and join with p2g_merkle_finish_cap", h, lg);
    if (LN > f.two_adicity || LN > 30 || L > 3 * MAX_TILE_LOG) return ctx->fail(P2G_ERR_INVALID, "LDE size 2^%u unsupported", LN);
    const bool keep = d->flags & P2G_FLAG_KEEP_ON_DEVICE, dev_in = d->flags & P2G_FLAG_DEVICE_INPUT;
    if (keep && !d->batch_out) return ctx->fail(P2G_ERR_INVALID, "P2G_FLAG_KEEP_ON_DEVICE needs batch_out");
    const bool values = d->input_kind == P2G_INPUT_VALUES;
    if (dev_in && !values) return ctx->fail(P2G_ERR_UNSUPPORTED, "multi-device context: DEVICE_INPUT takes values (column slices per device)");
    const uint64_t n = (uint64_t)1 << L, N = (uint64_t)1 << LN;
    const unsigned c = d->num_cols, width = c + d->num_salt_cols;
    const size_t col_bytes = n * f.elem;
    const unsigned cpr = (c + G - 1) / G;  // columns whose inverse NTT a device owns
    const uint64_t rows_local = N >> lg, num_cap_local = ((uint64_t)1 << h) >> lg;
    const uint64_t digests_local = 2 * (rows_local - num_cap_local);
    const uint64_t out_pitch = d->leaf_pitch ? d->leaf_pitch : width;
    if (out_pitch < width) return ctx->fail(P2G_ERR_INVALID, "leaf_pitch smaller than the leaf length");
    if (!dev_in)
        for (unsigned j = 0; j < c; j++)
            if (!d->col_ptrs[j]) return ctx->fail(P2G_ERR_INVALID, "col_ptrs[%u] is NULL", j);

    // gather buffers: all coefficient columns on every device (column slot j of device e's buffer = column j)
    if (values)
        for (unsigned e = 0; e < G; e++) {
            p2g_ctx *sub = ctx->subs[e];
            cudaSetDevice(sub->device);
            std::lock_guard<std::mutex> sl(sub->mu);
            int rc2 = ensure(sub, ctx->gather[e], (size_t)G * cpr * col_bytes);
            if (rc2) return ctx->fail(rc2, "device %d: %s", sub->device, sub->error.c_str());
        }

    std::vector<p2g_commit_desc> descs(G, *d);
    std::vector<int> rcs(G, 0);
    std::vector<std::string> errs(G);
    std::vector<p2g_batch *> shard_batches(G, nullptr);
    std::vector<const void *> gather_base(G, nullptr);
    ThreadBarrier barrier(G);
    auto worker = [&](unsigned e) {
        p2g_ctx *sub = ctx->subs[e];
        int &wrc = rcs[e];
        auto fail_here = [&](int code, const std::string &msg) {
            wrc = code;
            errs[e] = "device " + std::to_string(sub->device) + ": " + msg;
        };
        if (cudaSetDevice(sub->device) != cudaSuccess) fail_here(P2G_ERR_CUDA, "cudaSetDevice failed");
        const unsigned col0 = std::min(e * cpr, c), ncol = std::min(cpr, c - col0);
        if (values && !wrc) {
            std::lock_guard<std::mutex> sl(sub->mu);
            cudaStream_t st = sub->stream;
            cudaError_t ce = cudaSuccess;
            int rc2 = ncol ? ensure(sub, sub->in, (size_t)ncol * col_bytes) : 0;
            if (rc2) fail_here(rc2, sub->error);
            // this device's column slice: up from the host (adjacent columns in one copy), or across from the caller's device buffer
            if (!wrc && ncol) {
                if (dev_in) {
                    ce = cudaMemcpyAsync(sub->in.ptr, d->col_ptrs[e], (size_t)ncol * col_bytes, cudaMemcpyDeviceToDevice, st);
                } else {
                    for (unsigned j = col0; j < col0 + ncol && ce == cudaSuccess;) {
                        unsigned k = j + 1;
                        while (k < col0 + ncol && (const char *)d->col_ptrs[k] == (const char *)d->col_ptrs[k - 1] + col_bytes) k++;
                        ce = cudaMemcpyAsync((char *)sub->in.ptr + (size_t)(j - col0) * col_bytes, d->col_ptrs[j], (size_t)(k - j) * col_bytes,
                                             cudaMemcpyHostToDevice, st);
                        j = k;
                    }
                }
                if (ce != cudaSuccess) fail_here(P2G_ERR_CUDA, std::string("input copy failed: ") + cudaGetErrorString(ce));
            }
            // inverse NTT of the slice, coefficients stored into every device's gather buffer from the last pass' epilogue
            if (!wrc && ncol) {
                void *peers[8];
                unsigned np_ = 0;
                for (unsigned o = 0; o < G; o++)
                    if (o != e) peers[np_++] = (char *)ctx->gather[o].ptr + (size_t)col0 * col_bytes;
                rc2 = run_transform(sub, f, repr, true, L, ncol, sub->in.ptr, nullptr, (char *)ctx->gather[e].ptr + (size_t)col0 * col_bytes, n, 0,
                                    0, peers, np_);
                if (rc2) fail_here(rc2, sub->error);
            }
            if (cudaEventRecord(ctx->intt_done[e], st) != cudaSuccess && !wrc) fail_here(P2G_ERR_CUDA, "cudaEventRecord failed");
            // `polynomials` of this slice start their way home on the copy stream while the LDE runs
            if (!wrc && ncol && d->coeffs_out && !keep) {
                cudaStreamWaitEvent(sub->copy_stream, ctx->intt_done[e], 0);
                rc2 = copy_columns_home(sub, d->coeffs_out, col0, col0 + ncol, ctx->gather[e].ptr, col_bytes, sub->copy_stream);
                if (rc2) fail_here(rc2, sub->error);
            }
        }
        barrier.arrive_and_wait();  // every device's event is recorded (or its failure noted)
        bool any_failed = false;
        for (unsigned o = 0; o < G; o++) any_failed |= rcs[o] != 0;
        if (any_failed) {
            cudaStreamSynchronize(sub->stream);
            cudaStreamSynchronize(sub->copy_stream);
            return;
        }
        if (values)
            for (unsigned o = 0; o < G; o++)
                if (o != e) cudaStreamWaitEvent(sub->stream, ctx->intt_done[o], 0);  // the cross-device barrier, on the streams
        // this device's cosets: LDE + tree of shard e, outputs straight into the caller's slices
        p2g_commit_desc &sd = descs[e];
        sd.shard_index = e;
        sd.shard_count = G;
        sd.flags = d->flags & (P2G_FLAG_NO_LEAVES_OUT | P2G_FLAG_KEEP_ON_DEVICE);
        if (values) {
            sd.input_kind = P2G_INPUT_COEFFS;
            sd.flags |= P2G_FLAG_DEVICE_INPUT;
            gather_base[e] = ctx->gather[e].ptr;
            sd.col_ptrs = &gather_base[e];
            sd.coeffs_out = nullptr;
        }
        if (sd.leaves_out) sd.leaves_out = (char *)d->leaves_out + (size_t)e * rows_local * out_pitch * f.elem;
        if (sd.digests_out) sd.digests_out = (char *)d->digests_out + (size_t)e * digests_local * 32;
        if (sd.cap_out) sd.cap_out = (char *)d->cap_out + (size_t)e * num_cap_local * 32;
        sd.leaf_pitch = out_pitch;
        if (keep) sd.batch_out = &shard_batches[e];
        int rc3 = p2g_commit(sub, &sd);
        if (rc3) fail_here(rc3, sub->error);
        cudaStreamSynchronize(sub->copy_stream);
    };
    std::vector<std::thread> threads;
    for (unsigned e = 1; e < G; e++) threads.emplace_back(worker, e);
    worker(0);
    for (auto &t : threads) t.join();
    cudaSetDevice(ctx->subs[0]->device);
    for (unsigned e = 0; e < G; e++)
        if (rcs[e]) {
            for (unsigned o = 0; o < G; o++)
                if (shard_batches[o]) p2g_batch_free(ctx->subs[o], shard_batches[o]);
            ctx->error = errs[e];
            return rcs[e];
        }
    // timings: the slowest device
    for (int k = 0; k < 4; k++) d->stage_seconds[k] = d->kernel_seconds[k] = 0.0;
    d->h2d_seconds = d->d2h_seconds = 0.0;
    for (unsigned e = 0; e < G; e++) {
        for (int k = 0; k < 4; k++) {
            d->stage_seconds[k] = std::max(d->stage_seconds[k], descs[e].stage_seconds[k]);
            d->kernel_seconds[k] = std::max(d->kernel_seconds[k], descs[e].kernel_seconds[k]);
        }
        d->h2d_seconds = std::max(d->h2d_seconds, descs[e].h2d_seconds);
        d->d2h_seconds = std::max(d->d2h_seconds, descs[e].d2h_seconds);
    }
    if (keep) {
        p2g_batch *b = new p2g_batch(*shard_batches[0]);
        b->leaves = b->digests = b->cap = b->coeffs = nullptr;  // the shards own the memory
        b->leaves_bytes = b->digests_bytes = b->cap_bytes = b->coeffs_bytes = 0;
        b->shards = shard_batches;
        *d->batch_out = b;  // (every shard has already written its cap entries to cap_out)
    }
    return P2G_OK;
}

}  // namespace

extern "C" {

int32_t p2g_abi_version(void) { return P2G_ABI_VERSION; }

int32_t p2g_init(const int32_t *device_ids, int32_t n_devices, p2g_ctx **out_ctx) {
    if (!out_ctx) return P2G_ERR_INVALID;
    *out_ctx = nullptr;
    if (!device_ids || n_devices < 1 || n_devices > 8 || (n_devices & (n_devices - 1))) {
        g_init_error = "p2g_init: device list of 1, 2, 4 or 8 entries expected";
        return P2G_ERR_INVALID;
    }
    if (n_devices > 1) return init_multi(device_ids, n_devices, out_ctx);
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        g_init_error = std::string("p2g_init: no usable CUDA device (") + cudaGetErrorString(e) +
                       "); this library has no CPU fallback";
        return P2G_ERR_NO_DEVICE;
    }
    if (device_ids[0] < 0 || device_ids[0] >= count) {
        g_init_error = "p2g_init: device id out of range";
        return P2G_ERR_INVALID;
    }
    p2g_ctx *ctx = new p2g_ctx();
    ctx->device = device_ids[0];
    if ((e = cudaSetDevice(ctx->device)) != cudaSuccess || (e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) {
        g_init_error = std::string("p2g_init: ") + cudaGetErrorString(e);
        delete ctx;
        return P2G_ERR_NO_DEVICE;
    }
    cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking);
    cudaStreamCreateWithFlags(&ctx->h2d_stream, cudaStreamNonBlocking);
    for (auto &ev : ctx->gready) cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    for (auto &ev : ctx->gdone) cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    for (auto &ev : ctx->tev) cudaEventCreate(&ev);
    for (auto &ev : ctx->ev) cudaEventCreate(&ev);
    for (auto &ev : ctx->ready) cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    static PoseidonGlTables gl_tables;
    static Poseidon2BbTables bb_tables;
    static std::once_flag once;
    std::call_once(once, [] {
        build_poseidon_gl_tables(&gl_tables);
        build_poseidon2_bb_tables(&bb_tables);
    });
    if ((e = upload_hash_tables(&gl_tables, &bb_tables)) != cudaSuccess) {
        g_init_error = std::string("p2g_init: constant upload failed: ") + cudaGetErrorString(e);
        delete ctx;
        return P2G_ERR_CUDA;
    }
    *out_ctx = ctx;
    return P2G_OK;
}

void p2g_destroy(p2g_ctx *ctx) {
    if (!ctx) return;
    if (!ctx->subs.empty()) {
        for (size_t d = 0; d < ctx->subs.size(); d++) {
            cudaSetDevice(ctx->subs[d]->device);
            cudaStreamSynchronize(ctx->subs[d]->stream);
            if (ctx->gather[d].ptr) cudaFree(ctx->gather[d].ptr);
            cudaEventDestroy(ctx->intt_done[d]);
        }
        for (p2g_ctx *sub : ctx->subs) p2g_destroy(sub);
        delete ctx;
        return;
    }
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    p2g_release_workspace(ctx);
    for (auto &kv : ctx->pow_tables) cudaFree(kv.second);
    for (auto &kv : ctx->root_tables) cudaFree(kv.second);
    drop_pass_tables(ctx);
    for (auto &ev : ctx->ev) cudaEventDestroy(ev);
    for (auto &ev : ctx->ready) cudaEventDestroy(ev);
    for (auto &ev : ctx->gready) cudaEventDestroy(ev);
    for (auto &ev : ctx->gdone) cudaEventDestroy(ev);
    for (auto &ev : ctx->tev) cudaEventDestroy(ev);
    if (ctx->stage_ring) {
        cudaFreeHost(ctx->stage_ring);
        for (auto &ev : ctx->stage_ev) cudaEventDestroy(ev);
    }
    cudaStreamDestroy(ctx->h2d_stream);
    cudaStreamDestroy(ctx->copy_stream);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char *p2g_last_error(p2g_ctx *ctx) {
    if (!ctx) return g_init_error.c_str();
    // entry points a multi-device context hands to its first device leave their message there
    if (!ctx->subs.empty() && ctx->error.empty()) return ctx->subs[0]->error.c_str();
    return ctx->error.c_str();
}

int32_t p2g_num_devices(p2g_ctx *ctx) { return !ctx ? 0 : (ctx->subs.empty() ? 1 : (int32_t)ctx->subs.size()); }
p2g_ctx *p2g_sub_context(p2g_ctx *ctx, int32_t index) {
    if (!ctx || index < 0) return nullptr;
    if (ctx->subs.empty()) return index == 0 ? ctx : nullptr;
    return (size_t)index < ctx->subs.size() ? ctx->subs[index] : nullptr;
}

void *p2g_host_alloc(p2g_ctx *ctx, size_t bytes) {
    if (!ctx) return nullptr;
    cudaSetDevice(ctx->device);
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess) {  // pinned for every device of the process
        cudaGetLastError();
        ctx->fail(P2G_ERR_OOM, "cudaHostAlloc(%zu) failed", bytes);
        return nullptr;
    }
    return p;
}
void p2g_host_free(p2g_ctx *ctx, void *ptr) {
    (void)ctx;
    if (ptr) cudaFreeHost(ptr);
}

uint64_t p2g_workspace_bytes(p2g_ctx *ctx) {
    if (!ctx) return 0;
    if (!ctx->subs.empty()) {
        uint64_t total = 0;
        for (size_t e = 0; e < ctx->subs.size(); e++) total += p2g_workspace_bytes(ctx->subs[e]) + ctx->gather[e].bytes;
        return total;
    }
    return ctx->in.bytes + ctx->coeffs.bytes + ctx->tmp.bytes + ctx->leaves.bytes + ctx->digests.bytes + ctx->cap.bytes +
           ctx->salts.bytes + ctx->misc.bytes + ctx->pass_table_bytes + ctx->pool_bytes;
}
void p2g_release_workspace(p2g_ctx *ctx) {
    if (!ctx) return;
    if (!ctx->subs.empty()) {
        for (size_t e = 0; e < ctx->subs.size(); e++) {
            p2g_release_workspace(ctx->subs[e]);
            if (ctx->gather[e].ptr) cudaFree(ctx->gather[e].ptr);  // (release_workspace left that device current)
            ctx->gather[e] = Buffer();
        }
        return;
    }
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    drop_pass_tables(ctx);
    for (auto &pb : ctx->pool) cudaFree(pb.ptr);
    ctx->pool.clear();
    ctx->pool_bytes = 0;
    for (Buffer *b : {&ctx->in, &ctx->coeffs, &ctx->tmp, &ctx->leaves, &ctx->digests, &ctx->cap, &ctx->salts, &ctx->misc}) {
        if (b->ptr) cudaFree(b->ptr);
        b->ptr = nullptr;
        b->bytes = 0;
    }
}
void *p2g_stream(p2g_ctx *ctx) {
    if (ctx && !ctx->subs.empty()) ctx = ctx->subs[0];
    return ctx ? (void *)ctx->stream : nullptr;
}
uint64_t p2g_kernel_launches(p2g_ctx *ctx) {
    if (!ctx) return 0;
    uint64_t total = ctx->launches;
    for (p2g_ctx *sub : ctx->subs) total += sub->launches;
    return total;
}

int32_t p2g_commit(p2g_ctx *ctx, p2g_commit_desc *d) {
    if (!ctx) return P2G_ERR_NO_DEVICE;
    if (!d) return P2G_ERR_INVALID;
    if (!ctx->subs.empty()) return commit_multi(ctx, d);
    std::lock_guard<std::mutex> lock(ctx->mu);
    CK(cudaSetDevice(ctx->device));
    if (ctx->pass_table_bytes > ((size_t)2 << 30)) drop_pass_tables(ctx);  // bounded cache; tables are rebuilt on demand
    const FieldInfo *fp;
    int rc = check_field_hash(ctx, d->field, d->hash, &fp);
    if (rc) return rc;
    const FieldInfo &f = *fp;
    const int repr = f.id == FIELD_BB ? d->repr : 0;
    if (repr != P2G_REPR_CANONICAL && repr != P2G_REPR_MONTY_R32) return ctx->fail(P2G_ERR_INVALID, "bad repr %d", repr);
    if (d->num_cols == 0) return ctx->fail(P2G_ERR_INVALID, "empty polynomial batch (the reference panics on polynomials[0])");
    if (d->input_kind != P2G_INPUT_VALUES && d->input_kind != P2G_INPUT_COEFFS) return ctx->fail(P2G_ERR_INVALID, "bad input_kind");
    const unsigned L = d->log_n, r = d->rate_bits, LN = L + r;
    if (LN > f.two_adicity) return ctx->fail(P2G_ERR_INVALID, "log_n + rate_bits = %u exceeds the field's two-adicity %u", LN, f.two_adicity);
    if (LN > 30 || L > 3 * MAX_TILE_LOG) return ctx->fail(P2G_ERR_UNSUPPORTED, "LDE size 2^%u too large for this build", LN);
    if (d->cap_height > LN) return ctx->fail(P2G_ERR_INVALID, "cap_height=%u should be at most log2(leaves.len())=%u", d->cap_height, LN);
    if (d->num_salt_cols && !d->salt_ptrs) return ctx->fail(P2G_ERR_INVALID, "salt columns requested without salt_ptrs");
    const bool leaves_only = d->flags & P2G_FLAG_LEAVES_ONLY, leaves_ready = d->flags & P2G_FLAG_LEAVES_READY;
    if (leaves_only && leaves_ready) return ctx->fail(P2G_ERR_INVALID, "LEAVES_ONLY and LEAVES_READY are exclusive");
    if ((leaves_only || leaves_ready) && (!(d->flags & P2G_FLAG_DEVICE_OUTPUT) || (d->flags & P2G_FLAG_KEEP_ON_DEVICE) || !d->leaves_out))
        return ctx->fail(P2G_ERR_INVALID, "LEAVES_ONLY / LEAVES_READY need P2G_FLAG_DEVICE_OUTPUT and a device leaves_out");
    if (!d->col_ptrs && !leaves_ready) return ctx->fail(P2G_ERR_INVALID, "col_ptrs is NULL");
    unsigned G = d->shard_count ? d->shard_count : 1, gi = d->shard_count ? d->shard_index : 0;
    if (G & (G - 1)) return ctx->fail(P2G_ERR_INVALID, "shard_count must be a power of two");
    if (gi >= G) return ctx->fail(P2G_ERR_INVALID, "shard_index out of range");
    const unsigned lg = ilog2(G);
    if (lg > r) return ctx->fail(P2G_ERR_UNSUPPORTED, "shard_count %u exceeds the 2^rate_bits = %u cosets", G, 1u << r);

    const uint64_t n = (uint64_t)1 << L, N = (uint64_t)1 << LN;
    const unsigned c = d->num_cols, salt = d->num_salt_cols, width = c + salt;
    const unsigned cosets_local = (1u << r) >> lg, coset0 = gi * cosets_local;
    const uint64_t rows_local = N >> lg, row0 = (uint64_t)gi * rows_local;
    const unsigned tree_log = LN - lg, cap_local = d->cap_height >= lg ? d->cap_height - lg : 0;
    const uint64_t num_cap_local = (uint64_t)1 << cap_local;
    const uint64_t num_digests_local = 2 * (rows_local - num_cap_local);
    const bool keep = d->flags & P2G_FLAG_KEEP_ON_DEVICE;
    // device-resident leaves are written (and hashed) in place at the caller's pitch
    const uint64_t pitch_dev = ((d->flags & P2G_FLAG_DEVICE_OUTPUT) && !keep && d->leaves_out && d->leaf_pitch &&
                                !(d->flags & P2G_FLAG_NO_LEAVES_OUT)) ? d->leaf_pitch : width;
    if (keep && !d->batch_out) return ctx->fail(P2G_ERR_INVALID, "P2G_FLAG_KEEP_ON_DEVICE needs batch_out");
    if (keep && (d->flags & P2G_FLAG_DEVICE_OUTPUT)) return ctx->fail(P2G_ERR_INVALID, "KEEP_ON_DEVICE and DEVICE_OUTPUT are exclusive");
    const bool dev_in = d->flags & P2G_FLAG_DEVICE_INPUT, dev_out = (d->flags & P2G_FLAG_DEVICE_OUTPUT) || keep;
    const bool want_leaves = keep || (!(d->flags & P2G_FLAG_NO_LEAVES_OUT) && d->leaves_out);
    const uint64_t out_pitch = d->leaf_pitch ? d->leaf_pitch : width;
    if (out_pitch < width) return ctx->fail(P2G_ERR_INVALID, "leaf_pitch smaller than the leaf length");
    if ((leaves_only || leaves_ready) && !want_leaves) return ctx->fail(P2G_ERR_INVALID, "LEAVES_ONLY / LEAVES_READY with NO_LEAVES_OUT");

    const size_t col_bytes = n * f.elem;
    cudaStream_t st = ctx->stream;
    // an error return must not leave copies in flight that still reference the caller's host buffers
    struct StreamGuard {
        p2g_ctx *c;
        bool ok;
        ~StreamGuard() {
            if (ok) return;
            cudaStreamSynchronize(c->stream);
            cudaStreamSynchronize(c->copy_stream);
            cudaStreamSynchronize(c->h2d_stream);
        }
    } stream_guard{ctx, false};
    // ---- workspace
    if (!leaves_ready) {
        if (!dev_in || d->input_kind == P2G_INPUT_VALUES) {
            if ((rc = ensure(ctx, ctx->in, (size_t)c * col_bytes))) return rc;
        }
        if (d->input_kind == P2G_INPUT_VALUES)
            if ((rc = ensure(ctx, ctx->coeffs, (size_t)c * col_bytes))) return rc;
    }
    // a kept batch owns fresh device buffers instead of borrowing the context's workspace
    p2g_batch *batch = nullptr;
    struct BatchGuard {
        p2g_ctx *ctx;
        p2g_batch *&b;
        ~BatchGuard() {
            if (!b) return;
            cudaStreamSynchronize(ctx->stream);
            pool_free(ctx, b->leaves, b->leaves_bytes);
            pool_free(ctx, b->digests, b->digests_bytes);
            pool_free(ctx, b->cap, b->cap_bytes);
            pool_free(ctx, b->coeffs, b->coeffs_bytes);
            delete b;
        }
    } batch_guard{ctx, batch};
    if (keep) {
        batch = new p2g_batch();
        batch->field = f.id;
        batch->elem = f.elem;
        batch->rows = rows_local;
        batch->width = width;
        batch->num_cols = c;
        batch->n = n;
        batch->log_leaves = tree_log;
        batch->cap_height = cap_local;
        batch->num_digests = num_digests_local;
        batch->num_cap = num_cap_local;
        batch->leaves_bytes = (size_t)rows_local * pitch_dev * f.elem;
        batch->digests_bytes = num_digests_local * 32;
        batch->cap_bytes = num_cap_local * 32;
        batch->coeffs_bytes = (size_t)c * col_bytes;
        batch->leaves = pool_alloc(ctx, batch->leaves_bytes);
        batch->digests = pool_alloc(ctx, batch->digests_bytes);
        batch->cap = pool_alloc(ctx, batch->cap_bytes);
        batch->coeffs = pool_alloc(ctx, batch->coeffs_bytes);
        if (!batch->leaves || !batch->digests || !batch->cap || !batch->coeffs)
            return ctx->fail(P2G_ERR_OOM, "device allocation for the kept batch failed");
    }
    void *d_leaves = nullptr;
    if (keep) {
        d_leaves = batch->leaves;
    } else if (dev_out && want_leaves && out_pitch == pitch_dev) {
        d_leaves = d->leaves_out;
    } else {
        if ((rc = ensure(ctx, ctx->leaves, (size_t)rows_local * pitch_dev * f.elem))) return rc;
        d_leaves = ctx->leaves.ptr;
    }
    void *d_digests = nullptr, *d_cap = nullptr;
    if (leaves_only) {
        // no tree in this call
    } else if (keep) {
        d_digests = batch->digests;
    } else if (dev_out && d->digests_out) {
        d_digests = d->digests_out;
    } else {
        if ((rc = ensure(ctx, ctx->digests, std::max<size_t>(32, num_digests_local * 32)))) return rc;
        d_digests = ctx->digests.ptr;
    }
    if (leaves_only) {
    } else if (keep) {
        d_cap = batch->cap;
    } else if (dev_out && d->cap_out) {
        d_cap = d->cap_out;
    } else {
        if ((rc = ensure(ctx, ctx->cap, num_cap_local * 32))) return rc;
        d_cap = ctx->cap.ptr;
    }

    if (leaves_ready) {
        // the caller filled this shard's rows (several LEAVES_ONLY calls over column ranges): "build Merkle tree" only
        CK(cudaEventRecord(ctx->ev[3], st));
        TreeShape tree = {tree_log, cap_local, d_digests, d_cap};
        launch_leaf_hash(f.id, repr, d_leaves, pitch_dev, width, 0, rows_local, tree, st, &ctx->launches);
        CK(cudaEventRecord(ctx->ev[6], st));
        launch_tree_levels(f.id, repr, tree, st, &ctx->launches);
        CK(cudaGetLastError());
        CK(cudaEventRecord(ctx->ev[4], st));
        CK(cudaStreamSynchronize(st));
        for (int k = 0; k < 4; k++) d->stage_seconds[k] = d->kernel_seconds[k] = 0.0;
        d->h2d_seconds = d->d2h_seconds = 0.0;
        d->stage_seconds[3] = ev_seconds(ctx->ev[3], ctx->ev[4]);
        d->kernel_seconds[2] = ev_seconds(ctx->ev[3], ctx->ev[6]);
        d->kernel_seconds[3] = ev_seconds(ctx->ev[6], ctx->ev[4]);
        stream_guard.ok = true;
        return P2G_OK;
    }

    // ---- H2D, "IFFT", "FFT + blinding" + "transpose LDEs"
    // Columns are independent until the leaves are hashed, so with pinned host input they are processed in groups:
    // group g+1 crosses PCIe (h2d stream) while group g runs its inverse NTT and its 2^rate_bits coset NTTs, and the
    // coefficients of a finished group start their way home (d2h stream).  Device input / pageable memory = one group.
    CK(cudaEventRecord(ctx->ev[0], st));
    cudaStream_t cs = ctx->copy_stream, hs = ctx->h2d_stream;
    const bool overlap_d2h = !dev_out;
    const bool values = d->input_kind == P2G_INPUT_VALUES;
    const unsigned tile_cols = f.id == FIELD_GL ? 8 : 16;
    bool pipelined = !dev_in && c > 2 * tile_cols && all_pinned(d->col_ptrs, c);
    unsigned group_cols = c;
    if (pipelined) {
        group_cols = 2 * tile_cols;
        while ((c + group_cols - 1) / group_cols > p2g_ctx::MAX_GROUPS) group_cols += 2 * tile_cols;
    }
    const unsigned num_groups = (c + group_cols - 1) / group_cols;
    const bool coeffs_early = overlap_d2h && values && d->coeffs_out && all_pinned(d->coeffs_out, c);
    const bool coeffs_sent = coeffs_early;
    // from_coeffs into a kept batch: the coefficients land directly in the batch's own buffer
    void *h2d_base = (keep && !values) ? batch->coeffs : ctx->in.ptr;
    void *d_input = dev_in ? const_cast<void *>(d->col_ptrs[0]) : h2d_base;
    if (dev_in && !values && keep) {
        CK(cudaMemcpyAsync(batch->coeffs, d_input, (size_t)c * col_bytes, cudaMemcpyDeviceToDevice, st));
        d_input = batch->coeffs;
    }
    if (dev_in && values) {
        // the inverse transform runs in place on its input: work on a copy, the caller keeps its buffer
        CK(cudaMemcpyAsync(ctx->in.ptr, d_input, (size_t)c * col_bytes, cudaMemcpyDeviceToDevice, st));
        d_input = ctx->in.ptr;
    }
    if (!dev_in)
        for (unsigned j = 0; j < c; j++)
            if (!d->col_ptrs[j]) return ctx->fail(P2G_ERR_INVALID, "col_ptrs[%u] is NULL", j);
    if (salt) {
        if ((rc = ensure(ctx, ctx->salts, (size_t)salt * N * f.elem))) return rc;
        for (unsigned j = 0; j < salt; j++)
            // host or device pointers alike (the Rust shim draws its salts on the host even when the columns are
            // already device resident): the runtime resolves the direction from the pointer
            CK(cudaMemcpyAsync((char *)ctx->salts.ptr + (size_t)j * N * f.elem, d->salt_ptrs[j], N * f.elem, cudaMemcpyDefault, st));
    }
    void *d_coeffs = d_input;
    if (values) d_coeffs = keep ? batch->coeffs : ((dev_out && d->coeffs_out && d->coeffs_out[0]) ? d->coeffs_out[0] : ctx->coeffs.ptr);
    const uint64_t g = f.generator;
    const uint64_t wN = two_adic_generator(f, LN);
    if (pipelined) {
        // the h2d stream must not overwrite the input buffer while earlier work of this context still reads it
        CK(cudaEventRecord(ctx->gready[p2g_ctx::MAX_GROUPS], st));
        CK(cudaStreamWaitEvent(hs, ctx->gready[p2g_ctx::MAX_GROUPS], 0));
    }
    // cosets transformed inside the per-group loop (see below); 5/8 of them balance a group's transforms against its upload
    const unsigned early_cosets = pipelined ? std::max(1u, cosets_local * 5 / 8) : cosets_local;
    // several cosets share one launch per pass (the coefficients are then read from DRAM once per batch); the scratch
    // buffer of the strided passes holds the first pass' output of every coset in flight
    const unsigned nb_group = lde_batch(L, early_cosets, (size_t)std::min(group_cols, c) * col_bytes);
    const unsigned nb_all = lde_batch(L, cosets_local - early_cosets ? cosets_local - early_cosets : 1, (size_t)c * col_bytes);
    if (multi_pass(L))
        if ((rc = ensure(ctx, ctx->tmp, std::max((size_t)nb_group * std::min(group_cols, c), (size_t)nb_all * c) * col_bytes))) return rc;
    auto lde_cosets = [&](unsigned b_begin, unsigned b_end, unsigned nb, unsigned ncol, void *coeffs, void *leaves) -> int {
        for (unsigned b = b_begin; b < b_end; b += nb) {
            const unsigned cur = std::min(nb, b_end - b);
            uint64_t shifts[8];
            for (unsigned k = 0; k < cur; k++) shifts[k] = f_mul(f, g, f_pow(f, wN, bitrev(coset0 + b + k, r)));
            int rc2 = run_transform_batch(ctx, f, repr, false, L, ncol, coeffs, ctx->tmp.ptr, leaves, pitch_dev, shifts, cur,
                                          (uint64_t)b * n, n);
            if (rc2) return rc2;
        }
        return 0;
    };
    for (unsigned gi2 = 0; gi2 < num_groups; gi2++) {
        const unsigned col0 = gi2 * group_cols, ncol = std::min(group_cols, c - col0);
        char *g_in = (char *)d_input + (size_t)col0 * col_bytes;
        char *g_co = (char *)d_coeffs + (size_t)col0 * col_bytes;
        if (!dev_in && !pipelined && (size_t)ncol * col_bytes >= kStageMin && !is_pinned(d->col_ptrs[col0])) {
            // pageable input: through the pinned ring on several host threads (staged_copy), not the driver's own staging
            std::vector<HostPiece> pieces;
            for (unsigned j = col0; j < col0 + ncol; j++) pieces.push_back(HostPiece{(char *)const_cast<void *>(d->col_ptrs[j]), col_bytes});
            if ((rc = staged_copy(ctx, pieces, (char *)h2d_base + (size_t)col0 * col_bytes, false, st))) return rc;
        } else if (!dev_in) {
            cudaStream_t up = pipelined ? hs : st;
            for (unsigned j = col0; j < col0 + ncol;) {
                // columns that are adjacent in host memory (a 2-D array behind the pointer table) go up in one copy
                unsigned k = j + 1;
                while (k < col0 + ncol && (const char *)d->col_ptrs[k] == (const char *)d->col_ptrs[k - 1] + col_bytes) k++;
                CK(cudaMemcpyAsync((char *)h2d_base + (size_t)j * col_bytes, d->col_ptrs[j], (size_t)(k - j) * col_bytes, cudaMemcpyHostToDevice, up));
                j = k;
            }
            if (pipelined) {
                CK(cudaEventRecord(ctx->gready[gi2], hs));
                CK(cudaStreamWaitEvent(st, ctx->gready[gi2], 0));
            }
        }
        if (gi2 == 0) CK(cudaEventRecord(ctx->ev[1], st));
        CK(cudaEventRecord(ctx->tev[3 * gi2], st));
        if (values)
            if ((rc = run_transform(ctx, f, repr, true, L, ncol, g_in, nullptr, g_co, n, 0, 0))) return rc;
        CK(cudaEventRecord(ctx->tev[3 * gi2 + 1], st));
        if (coeffs_early) {
            // `polynomials` of this group are final: send them home while the LDE and the hashing run
            CK(cudaEventRecord(ctx->gdone[gi2], st));
            CK(cudaStreamWaitEvent(cs, ctx->gdone[gi2], 0));
            if ((rc = copy_columns_home(ctx, d->coeffs_out, col0, col0 + ncol, d_coeffs, col_bytes, cs))) return rc;
        }
        // one coset of n rows at a time, written straight into leaf order.  Only the first `early_cosets` of them
        // here: that much transform work hides under the upload of the next column group; the rest follows below,
        // coset by coset, so that finished cosets can start their way home while later ones are still computed.
        if ((rc = lde_cosets(0, early_cosets, nb_group, ncol, g_co, (char *)d_leaves + (size_t)col0 * f.elem))) return rc;
        CK(cudaEventRecord(ctx->tev[3 * gi2 + 2], st));
    }
    const bool leaves_early = overlap_d2h && want_leaves && !salt && cosets_local <= 62 && is_pinned(d->leaves_out);
    auto send_coset = [&](unsigned b) -> int {
        // every row of coset b is final (no salt columns to add): start its D2H on the copy stream
        CK(cudaEventRecord(ctx->ready[b], st));
        CK(cudaStreamWaitEvent(cs, ctx->ready[b], 0));
        char *src = (char *)d_leaves + (size_t)b * n * pitch_dev * f.elem;
        char *dsth = (char *)d->leaves_out + (size_t)b * n * out_pitch * f.elem;
        if (out_pitch == pitch_dev)
            CK(cudaMemcpyAsync(dsth, src, (size_t)n * pitch_dev * f.elem, cudaMemcpyDeviceToHost, cs));
        else
            CK(cudaMemcpy2DAsync(dsth, out_pitch * f.elem, src, pitch_dev * f.elem, width * f.elem, n, cudaMemcpyDeviceToHost, cs));
        return 0;
    };
    if (leaves_early)
        for (unsigned b = 0; b < early_cosets; b++)
            if ((rc = send_coset(b))) return rc;
    // every column is on the device by now: one launch per pass over all of them.  When finished cosets travel home
    // during the computation they are transformed one at a time, so that each can leave as soon as it is complete.
    for (unsigned b = early_cosets; b < cosets_local;) {
        const unsigned cur = leaves_early ? 1 : std::min(nb_all, cosets_local - b);
        if ((rc = lde_cosets(b, b + cur, cur, c, d_coeffs, d_leaves))) return rc;
        if (leaves_early)
            for (unsigned k = 0; k < cur; k++)
                if ((rc = send_coset(b + k))) return rc;
        b += cur;
    }
    const bool leaves_sent = leaves_early;
    if (salt)
        launch_salt_scatter(f.id, ctx->salts.ptr, salt, LN, row0, rows_local, d_leaves, pitch_dev, c, st, &ctx->launches);
    CK(cudaEventRecord(ctx->ev[3], st));
    if (leaves_only) {
        // the rows' column range [0, width) at leaves_out is final; the tree is built by a later LEAVES_READY call
        CK(cudaStreamSynchronize(st));
        CK(cudaStreamSynchronize(cs));
        CK(cudaStreamSynchronize(hs));
        d->h2d_seconds = ev_seconds(ctx->ev[0], ctx->ev[1]);
        d->d2h_seconds = 0.0;
        d->stage_seconds[0] = d->stage_seconds[1] = 0.0;
        for (unsigned gi2 = 0; gi2 < num_groups; gi2++) {
            d->stage_seconds[0] += ev_seconds(ctx->tev[3 * gi2], ctx->tev[3 * gi2 + 1]);
            d->stage_seconds[1] += ev_seconds(ctx->tev[3 * gi2 + 1], ctx->tev[3 * gi2 + 2]);
        }
        d->stage_seconds[1] += ev_seconds(ctx->tev[3 * (num_groups - 1) + 2], ctx->ev[3]);
        d->stage_seconds[2] = d->stage_seconds[3] = 0.0;
        d->kernel_seconds[0] = d->stage_seconds[0];
        d->kernel_seconds[1] = d->stage_seconds[1];
        d->kernel_seconds[2] = d->kernel_seconds[3] = 0.0;
        stream_guard.ok = true;
        return P2G_OK;
    }

    // ---- "build Merkle tree"
    // (Running the leaf sponge of finished cosets on a second, low-priority stream under the transforms of the next
    // ones was measured: 136.7 against 134.7 ms per 2^20 x 135 commit -- the two kernels do not share an SM's issue
    // slots productively.)
    TreeShape tree = {tree_log, cap_local, d_digests, d_cap};
    launch_leaf_hash(f.id, repr, d_leaves, pitch_dev, width, 0, rows_local, tree, st, &ctx->launches);
    CK(cudaEventRecord(ctx->ev[6], st));
    launch_tree_levels(f.id, repr, tree, st, &ctx->launches);
    CK(cudaGetLastError());
    CK(cudaEventRecord(ctx->ev[4], st));

    // ---- D2H
    // Pageable destinations: every launch is enqueued by now, so this thread is free to pull the big outputs home
    // through the pinned ring (staged_copy, on the copy stream) as soon as each is final -- the coefficients after the
    // inverse NTT, the leaves after the LDE (while the tree is still being hashed), the digests after the tree.
    bool coeffs_staged = false, leaves_staged = false, digests_staged = false;
    if (!dev_out) {
        const size_t leaf_bytes_local = (size_t)rows_local * pitch_dev * f.elem;
        if (values && d->coeffs_out && !coeffs_sent && (size_t)c * col_bytes >= kStageMin && !is_pinned(d->coeffs_out[0])) {
            bool all = true;
            for (unsigned j = 0; j < c; j++) all &= d->coeffs_out[j] != nullptr;
            if (all) {
                std::vector<HostPiece> pieces;
                for (unsigned j = 0; j < c; j++) pieces.push_back(HostPiece{(char *)d->coeffs_out[j], col_bytes});
                CK(cudaEventSynchronize(ctx->tev[3 * (num_groups - 1) + 1]));  // the last group's inverse NTT is done
                if ((rc = staged_copy(ctx, pieces, d_coeffs, true, cs))) return rc;
                coeffs_staged = true;
            }
        }
        if (want_leaves && !leaves_sent && out_pitch == pitch_dev && leaf_bytes_local >= kStageMin && !is_pinned(d->leaves_out)) {
            CK(cudaEventSynchronize(ctx->ev[3]));  // LDE (and salts) done; the sponge kernels read the same rows meanwhile
            if ((rc = staged_copy(ctx, {HostPiece{(char *)d->leaves_out, leaf_bytes_local}}, d_leaves, true, cs))) return rc;
            leaves_staged = true;
        }
        if (d->digests_out && num_digests_local * 32 >= kStageMin && !is_pinned(d->digests_out)) {
            CK(cudaEventSynchronize(ctx->ev[4]));
            if ((rc = staged_copy(ctx, {HostPiece{(char *)d->digests_out, (size_t)num_digests_local * 32}}, d_digests, true, cs))) return rc;
            digests_staged = true;
        }
    }
    if (!dev_out) {
        if (d->input_kind == P2G_INPUT_VALUES && d->coeffs_out && !coeffs_sent && !coeffs_staged)
            if ((rc = copy_columns_home(ctx, d->coeffs_out, 0, c, d_coeffs, col_bytes, st))) return rc;
        if (want_leaves && !leaves_sent && !leaves_staged) {
            if (out_pitch == pitch_dev)
                CK(cudaMemcpyAsync(d->leaves_out, d_leaves, (size_t)rows_local * pitch_dev * f.elem, cudaMemcpyDeviceToHost, st));
            else
                CK(cudaMemcpy2DAsync(d->leaves_out, out_pitch * f.elem, d_leaves, pitch_dev * f.elem, width * f.elem, rows_local,
                                     cudaMemcpyDeviceToHost, st));
        }
        if (d->digests_out && num_digests_local && !digests_staged)
            CK(cudaMemcpyAsync(d->digests_out, d_digests, num_digests_local * 32, cudaMemcpyDeviceToHost, st));
        if (d->cap_out) CK(cudaMemcpyAsync(d->cap_out, d_cap, num_cap_local * 32, cudaMemcpyDeviceToHost, st));
    } else if (keep) {
        if (d->cap_out) CK(cudaMemcpyAsync(d->cap_out, d_cap, num_cap_local * 32, cudaMemcpyDeviceToHost, st));
    } else if (want_leaves && d_leaves != d->leaves_out) {
        CK(cudaMemcpy2DAsync(d->leaves_out, out_pitch * f.elem, d_leaves, pitch_dev * f.elem, width * f.elem, rows_local,
                             cudaMemcpyDeviceToDevice, st));
    }
    CK(cudaEventRecord(ctx->ev[5], st));
    CK(cudaStreamSynchronize(st));
    CK(cudaStreamSynchronize(cs));
    CK(cudaStreamSynchronize(hs));
    d->h2d_seconds = ev_seconds(ctx->ev[0], ctx->ev[1]);  // exposed part: until the first group is on the device
    d->stage_seconds[0] = d->stage_seconds[1] = 0.0;
    for (unsigned gi2 = 0; gi2 < num_groups; gi2++) {
        d->stage_seconds[0] += ev_seconds(ctx->tev[3 * gi2], ctx->tev[3 * gi2 + 1]);
        d->stage_seconds[1] += ev_seconds(ctx->tev[3 * gi2 + 1], ctx->tev[3 * gi2 + 2]);
    }
    d->stage_seconds[1] += ev_seconds(ctx->tev[3 * (num_groups - 1) + 2], ctx->ev[3]);  // salt scatter
    d->stage_seconds[2] = 0.0;
    d->stage_seconds[3] = ev_seconds(ctx->ev[3], ctx->ev[4]);
    d->d2h_seconds = ev_seconds(ctx->ev[4], ctx->ev[5]);
    d->kernel_seconds[0] = d->stage_seconds[0];
    d->kernel_seconds[1] = d->stage_seconds[1];
    d->kernel_seconds[2] = ev_seconds(ctx->ev[3], ctx->ev[6]);
    d->kernel_seconds[3] = ev_seconds(ctx->ev[6], ctx->ev[4]);
    if (keep) {
        *d->batch_out = batch;
        batch = nullptr;  // ownership passes to the caller
    }
    stream_guard.ok = true;
    return P2G_OK;
}

int32_t p2g_batch_shape(p2g_batch *b, uint64_t *rows, uint64_t *width, uint32_t *num_layers, uint64_t *num_cap) {
    if (!b) return P2G_ERR_INVALID;
    const uint64_t g = b->shards.empty() ? 1 : b->shards.size();  // a multi-device batch: g shards of b->rows rows each
    if (rows) *rows = b->rows * g;
    if (width) *width = b->width;
    if (num_layers) *num_layers = b->log_leaves - b->cap_height;
    if (num_cap) *num_cap = b->num_cap * g;
    return P2G_OK;
}

// multi-device batch: leaf index -> (shard, local index); the shards' trees are the cap subtrees of the whole tree
// (cap_height >= log2(devices)), so a Merkle path never leaves its shard
static int32_t batch_route(p2g_ctx *ctx, p2g_batch *b, const uint64_t *idx, uint64_t count, size_t out_stride, void *out, bool prove) {
    if (ctx->subs.size() != b->shards.size()) return ctx->fail(P2G_ERR_INVALID, "batch does not belong to this context");
    const uint64_t g = b->shards.size();
    std::vector<std::vector<uint64_t>> local(g), where(g);
    for (uint64_t i = 0; i < count; i++) {
        if (idx[i] >= b->rows * g) return ctx->fail(P2G_ERR_INVALID, "index %llu out of range", (unsigned long long)idx[i]);
        const uint64_t s_ = idx[i] / b->rows;
        local[s_].push_back(idx[i] - s_ * b->rows);
        where[s_].push_back(i);
    }
    std::vector<unsigned char> tmp;
    for (uint64_t s_ = 0; s_ < g; s_++) {
        if (local[s_].empty()) continue;
        tmp.resize(local[s_].size() * out_stride);
        int32_t rc = prove ? p2g_batch_prove(ctx->subs[s_], b->shards[s_], local[s_].data(), local[s_].size(), tmp.data())
                           : p2g_batch_get_rows(ctx->subs[s_], b->shards[s_], local[s_].data(), local[s_].size(), tmp.data());
        if (rc) return ctx->fail(rc, "device %d: %s", ctx->subs[s_]->device, ctx->subs[s_]->error.c_str());
        for (size_t k = 0; k < where[s_].size(); k++)
            memcpy((char *)out + where[s_][k] * out_stride, tmp.data() + k * out_stride, out_stride);
    }
    return P2G_OK;
}

void p2g_batch_free(p2g_ctx *ctx, p2g_batch *b) {
    if (!b) return;
    if (!b->shards.empty()) {
        for (size_t s_ = 0; s_ < b->shards.size(); s_++)
            p2g_batch_free(ctx && s_ < ctx->subs.size() ? ctx->subs[s_] : nullptr, b->shards[s_]);
        delete b;
        return;
    }
    if (ctx) {
        std::lock_guard<std::mutex> lock(ctx->mu);
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->stream);
        pool_free(ctx, b->leaves, b->leaves_bytes);
        pool_free(ctx, b->digests, b->digests_bytes);
        pool_free(ctx, b->cap, b->cap_bytes);
        pool_free(ctx, b->coeffs, b->coeffs_bytes);
    } else {
        for (void *q : {b->leaves, b->digests, b->cap, b->coeffs})
            if (q) cudaFree(q);
    }
    delete b;
}

int32_t p2g_batch_get_rows(p2g_ctx *ctx, p2g_batch *b, const uint64_t *row_indices, uint64_t count, void *rows_out) {
    if (!ctx) return P2G_ERR_NO_DEVICE;
    if (!b || (count && (!row_indices || !rows_out))) return P2G_ERR_INVALID;
    std::lock_guard<std::mutex> lock(ctx->mu);
    if (!b->shards.empty()) return batch_route(ctx, b, row_indices, count, (size_t)b->width * b->elem, rows_out, false);
    CK(cudaSetDevice(ctx->device));
    if (!count) return P2G_OK;
    for (uint64_t i = 0; i < count; i++)
        if (row_indices[i] >= b->rows) return ctx->fail(P2G_ERR_INVALID, "row index %llu out of range", (unsigned long long)row_indices[i]);
    size_t idx_bytes = count * 8, out_bytes = count * b->width * b->elem;
    size_t off = (idx_bytes + 255) & ~(size_t)255;
    int rc = ensure(ctx, ctx->misc, off + std::max<size_t>(out_bytes, 16));
    if (rc) return rc;
    cudaStream_t st = ctx->stream;
    CK(cudaMemcpyAsync(ctx->misc.ptr, row_indices, idx_bytes, cudaMemcpyHostToDevice, st));
    launch_gather_rows(b->field, b->leaves, b->width, (uint32_t)b->width, (const uint64_t *)ctx->misc.ptr, count,
                       (char *)ctx->misc.ptr + off, st, &ctx->launches);
    CK(cudaGetLastError());
    if (out_bytes) CK(cudaMemcpyAsync(rows_out, (char *)ctx->misc.ptr + off, out_bytes, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return P2G_OK;
}

int32_t p2g_batch_prove(p2g_ctx *ctx, p2g_batch *b, const uint64_t *leaf_indices, uint64_t count, void *siblings_out) {
    if (!ctx) return P2G_ERR_NO_DEVICE;
    if (!b || (count && !leaf_indices)) return P2G_ERR_INVALID;
    std::lock_guard<std::mutex> lock(ctx->mu);
    if (!b->shards.empty()) {
        if (!count || b->log_leaves == b->cap_height) return P2G_OK;
        if (!siblings_out) return P2G_ERR_INVALID;
        return batch_route(ctx, b, leaf_indices, count, (size_t)(b->log_leaves - b->cap_height) * 32, siblings_out, true);
    }
    CK(cudaSetDevice(ctx->device));
    const uint32_t layers = b->log_leaves - b->cap_height;
    if (!count || !layers) return P2G_OK;
    if (!siblings_out) return P2G_ERR_INVALID;
    for (uint64_t i = 0; i < count; i++)
        if (leaf_indices[i] >= b->rows) return ctx->fail(P2G_ERR_INVALID, "leaf index %llu out of range", (unsigned long long)leaf_indices[i]);
    size_t idx_bytes = count * 8, out_bytes = count * layers * 32;
    size_t off = (idx_bytes + 255) & ~(size_t)255;
    int rc = ensure(ctx, ctx->misc, off + out_bytes);
    if (rc) return rc;
    cudaStream_t st = ctx->stream;
    CK(cudaMemcpyAsync(ctx->misc.ptr, leaf_indices, idx_bytes, cudaMemcpyHostToDevice, st));
    launch_gather_proofs(b->digests, b->log_leaves, b->cap_height, (const uint64_t *)ctx->misc.ptr, count,
                         (char *)ctx->misc.ptr + off, st, &ctx->launches);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(siblings_out, (char *)ctx->misc.ptr + off, out_bytes, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return P2G_OK;
}

int32_t p2g_batch_download(p2g_ctx *ctx, p2g_batch *b, int32_t what, void *out) {
    if (!ctx) return P2G_ERR_NO_DEVICE;
    if (!b || !out) return P2G_ERR_INVALID;
    std::lock_guard<std::mutex> lock(ctx->mu);
    if (!b->shards.empty()) {
        // leaves / digests / cap: the shards' arrays one after the other ARE the whole arrays; coefficients: any shard has all
        if (ctx->subs.size() != b->shards.size()) return ctx->fail(P2G_ERR_INVALID, "batch does not belong to this context");
        const size_t per = what == 0 ? (size_t)b->rows * b->width * b->elem : (what == 1 ? (size_t)b->num_digests * 32 : (size_t)b->num_cap * 32);
        for (size_t s_ = 0; s_ < b->shards.size(); s_++) {
            if (what == 3 && s_) break;
            int32_t rc = p2g_batch_download(ctx->subs[s_], b->shards[s_], what, (char *)out + (what == 3 ? 0 : s_ * per));
            if (rc) return ctx->fail(rc, "device %d: %s", ctx->subs[s_]->device, ctx->subs[s_]->error.c_str());
        }
        return P2G_OK;
    }
    CK(cudaSetDevice(ctx->device));
    const void *src;
    size_t bytes;
    switch (what) {
        case 0: src = b->leaves; bytes = (size_t)b->rows * b->width * b->elem; break;
        case 1: src = b->digests; bytes = (size_t)b->num_digests * 32; break;
        case 2: src = b->cap; bytes = (size_t)b->num_cap * 32; break;
        case 3: src = b->coeffs; bytes = (size_t)b->num_cols * b->n * b->elem; break;
        default: return ctx->fail(P2G_ERR_INVALID, "unknown array %d", what);
    }
    if (bytes) CK(cudaMemcpyAsync(out, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return P2G_OK;
}

static int32_t merkle_locked(p2g_ctx *ctx, p2g_merkle_desc *d) {
    CK(cudaSetDevice(ctx->device));
    const FieldInfo *fp;
    int rc = check_field_hash(ctx, d->field, d->hash, &fp);
    if (rc) return rc;
    const FieldInfo &f = *fp;
    const int repr = f.id == FIELD_BB ? d->repr : 0;
    if (d->num_leaves == 0 || (d->num_leaves & (d->num_leaves - 1)))
        return ctx->fail(P2G_ERR_INVALID, "num_leaves must be a power of two (log2_strict)");
    unsigned lg = ilog2(d->num_leaves);
    if (d->cap_height > lg) return ctx->fail(P2G_ERR_INVALID, "cap_height=%u should be at most log2(leaves.len())=%u", d->cap_height, lg);
    if (lg > 31) return ctx->fail(P2G_ERR_UNSUPPORTED, "too many leaves");
    uint64_t pitch = d->leaf_pitch ? d->leaf_pitch : d->leaf_len;
    if (pitch < d->leaf_len) return ctx->fail(P2G_ERR_INVALID, "leaf_pitch smaller than leaf_len");
    if (d->leaf_len > 0xffffffffull) return ctx->fail(P2G_ERR_UNSUPPORTED, "leaf too long");
    if (!d->leaves && d->leaf_len) return ctx->fail(P2G_ERR_INVALID, "leaves is NULL");
    const bool dev_in = d->flags & P2G_FLAG_DEVICE_INPUT, dev_out = d->flags & P2G_FLAG_DEVICE_OUTPUT;
    uint64_t num_cap = (uint64_t)1 << d->cap_height, num_digests = 2 * (d->num_leaves - num_cap);
    cudaStream_t st = ctx->stream;
    const void *d_leaves = d->leaves;
    size_t leaf_bytes = (size_t)d->num_leaves * pitch * f.elem;
    if (!dev_in) {
        if ((rc = ensure(ctx, ctx->leaves, std::max<size_t>(leaf_bytes, 32)))) return rc;
        if (leaf_bytes) CK(cudaMemcpyAsync(ctx->leaves.ptr, d->leaves, leaf_bytes, cudaMemcpyHostToDevice, st));
        d_leaves = ctx->leaves.ptr;
    }
    void *d_dig = d->digests_out, *d_cap = d->cap_out;
    if (!dev_out) {
        if ((rc = ensure(ctx, ctx->digests, std::max<size_t>(32, num_digests * 32)))) return rc;
        if ((rc = ensure(ctx, ctx->cap, num_cap * 32))) return rc;
        d_dig = ctx->digests.ptr;
        d_cap = ctx->cap.ptr;
    }
    CK(cudaEventRecord(ctx->ev[0], st));
    TreeShape tree = {lg, d->cap_height, d_dig, d_cap};
    launch_leaf_hash(f.id, repr, d_leaves, pitch, (uint32_t)d->leaf_len, 0, d->num_leaves, tree, st, &ctx->launches);
    launch_tree_levels(f.id, repr, tree, st, &ctx->launches);
    CK(cudaGetLastError());
    CK(cudaEventRecord(ctx->ev[1], st));
    if (!dev_out) {
        if (d->digests_out && num_digests) CK(cudaMemcpyAsync(d->digests_out, d_dig, num_digests * 32, cudaMemcpyDeviceToHost, st));
        if (d->cap_out) CK(cudaMemcpyAsync(d->cap_out, d_cap, num_cap * 32, cudaMemcpyDeviceToHost, st));
    }
    CK(cudaStreamSynchronize(st));
    d->seconds = ev_seconds(ctx->ev[0], ctx->ev[1]);
    return P2G_OK;
}

int32_t p2g_merkle(p2g_ctx *ctx, p2g_merkle_desc *d) {
    if (ctx && !ctx->subs.empty()) {
        int32_t rc_ = p2g_merkle(ctx->subs[0], d);
        if (rc_) ctx->error = ctx->subs[0]->error;
        return rc_;
    }
    if (!ctx) return P2G_ERR_NO_DEVICE;
    if (!d) return P2G_ERR_INVALID;
    std::lock_guard<std::mutex> lock(ctx->mu);
    return merkle_locked(ctx, d);
}

int32_t p2g_merkle_finish_cap(p2g_ctx *ctx, int32_t field, int32_t hash, int32_t repr, const void *roots,
                              uint32_t shard_count, uint32_t cap_height, void *top_digests_out, void *cap_out) {
    if (!ctx) return P2G_ERR_NO_DEVICE;
    if (!ctx->subs.empty()) ctx = ctx->subs[0];  // multi-device context: its first device serves this call
    std::lock_guard<std::mutex> lock(ctx->mu);
    CK(cudaSetDevice(ctx->device));
    const FieldInfo *fp;
    int rc = check_field_hash(ctx, field, hash, &fp);
    if (rc) return rc;
    if (!shard_count || (shard_count & (shard_count - 1))) return ctx->fail(P2G_ERR_INVALID, "shard_count must be a power of two");
    unsigned lg = ilog2(shard_count);
    if (cap_height > lg) return ctx->fail(P2G_ERR_INVALID, "cap_height exceeds log2(shard_count)");
    // The gathered roots are the leaves (already digests: hash_or_noop copies 32-byte leaves) of a tiny tree whose
    // inner digests, in the reference layout, are exactly the slots that join the shards' local digest blocks.
    p2g_merkle_desc m;
    memset(&m, 0, sizeof m);
    m.field = field;
    m.hash = hash;
    m.repr = fp->id == FIELD_BB ? repr : 0;
    m.num_leaves = shard_count;
    m.leaf_len = fp->digest_elems;
    m.cap_height = cap_height;
    m.leaves = roots;
    m.digests_out = top_digests_out;
    m.cap_out = cap_out;
    return merkle_locked(ctx, &m);
}

// ---- FRI commit phase ---------------------------------------------------------------------------------
static uint64_t ext_w(const FieldInfo &f) { return f.id == FIELD_GL ? 7 : 11; }  // X^D = W of the binomial extension
static unsigned ext_degree(const FieldInfo &f) { return f.id == FIELD_GL ? 2 : 4; }

int32_t p2g_fri_begin(p2g_ctx *ctx, int32_t field, int32_t hash, int32_t repr_in, uint32_t degree, uint32_t log_n,
                      const void *coeffs, const void *values, uint32_t flags, p2g_fri **out) {
    if (!ctx) return P2G_ERR_NO_DEVICE;
    if (!ctx->subs.empty()) ctx = ctx->subs[0];  // multi-device context: its first device serves this call
    std::lock_guard<std::mutex> lock(ctx->mu);
    CK(cudaSetDevice(ctx->device));
    if (!out) return ctx->fail(P2G_ERR_INVALID, "out is NULL");
    *out = nullptr;
    const FieldInfo *fp;
    int rc = check_field_hash(ctx, field, hash, &fp);
    if (rc) return rc;
    const FieldInfo &f = *fp;
    const int repr = f.id == FIELD_BB ? repr_in : 0;
    if (repr != P2G_REPR_CANONICAL && repr != P2G_REPR_MONTY_R32) return ctx->fail(P2G_ERR_INVALID, "bad repr %d", repr);
    if (degree != ext_degree(f))
        return ctx->fail(P2G_ERR_UNSUPPORTED, "extension degree %u: the GPU path has D = 2 for Goldilocks and D = 4 for BabyBear", degree);
    if (log_n > f.two_adicity || log_n > 30) return ctx->fail(P2G_ERR_INVALID, "codeword length 2^%u exceeds the field's two-adicity", log_n);
    if (!coeffs) return ctx->fail(P2G_ERR_INVALID, "coeffs is NULL");
    const uint64_t n = (uint64_t)1 << log_n;
    const size_t bytes = (size_t)n * degree * f.elem;
    const bool dev_in = flags & P2G_FLAG_DEVICE_INPUT;
    p2g_fri *s = new p2g_fri();
    s->field = field;
    s->hash = hash;
    s->repr = repr;
    s->degree = degree;
    s->elem = f.elem;
    s->log_n = log_n;
    s->shift = f.generator;  // prover.rs:100
    s->coeffs_bytes = s->values_bytes = std::max<size_t>(bytes, 32);
    s->coeffs = pool_alloc(ctx, s->coeffs_bytes);
    s->values = pool_alloc(ctx, s->values_bytes);
    struct Guard {
        p2g_ctx *ctx;
        p2g_fri *s;
        ~Guard() {
            if (!s) return;
            cudaStreamSynchronize(ctx->stream);
            pool_free(ctx, s->coeffs, s->coeffs_bytes);
            pool_free(ctx, s->values, s->values_bytes);
            delete s;
        }
    } guard{ctx, s};
    if (!s->coeffs || !s->values) return ctx->fail(P2G_ERR_OOM, "device allocation for the FRI state failed");
    cudaStream_t st = ctx->stream;
    const void *src = coeffs;
    if (!dev_in) {
        if ((rc = ensure(ctx, ctx->in, bytes))) return rc;
        CK(cudaMemcpyAsync(ctx->in.ptr, coeffs, bytes, cudaMemcpyHostToDevice, st));
        src = ctx->in.ptr;
    }
    launch_fri_split(f.id, src, n, s->coeffs, st, &ctx->launches);
    if (values) {
        src = values;
        if (!dev_in) {
            CK(cudaMemcpyAsync(ctx->in.ptr, values, bytes, cudaMemcpyHostToDevice, st));
            src = ctx->in.ptr;
        }
        launch_fri_bitrev16(src, log_n, s->values, st, &ctx->launches);  // reverse_index_bits_in_place, prover.rs:103
    } else if (log_n == 0) {
        CK(cudaMemcpyAsync(s->values, src, bytes, cudaMemcpyDeviceToDevice, st));
    } else {
        // lde_polynomial_values = lde_polynomial_coeffs.coset_fft(F::generator()) (fri/oracle.rs:226-231), computed here,
        // straight into bit-reversed order
        if (multi_pass(log_n) && (rc = ensure(ctx, ctx->tmp, bytes))) return rc;
        if ((rc = run_transform(ctx, f, repr, false, log_n, degree, s->coeffs, ctx->tmp.ptr, s->values, degree, f.generator, 0))) return rc;
    }
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(st));
    guard.s = nullptr;
    *out = s;
    return P2G_OK;
}

int32_t p2g_fri_commit_layer(p2g_ctx *ctx, p2g_fri *s, uint32_t arity_bits, uint32_t cap_height, void *cap_out,
                             p2g_batch **tree_out) {
    if (!ctx) return P2G_ERR_NO_DEVICE;
    if (!ctx->subs.empty()) ctx = ctx->subs[0];  // multi-device context: its first device serves this call
    std::lock_guard<std::mutex> lock(ctx->mu);
    CK(cudaSetDevice(ctx->device));
    if (!s || !tree_out) return ctx->fail(P2G_ERR_INVALID, "state / tree_out is NULL");
    if (!s->values || s->pending_arity_bits >= 0) return ctx->fail(P2G_ERR_INVALID, "layer already committed: call p2g_fri_fold first");
    if (arity_bits > s->log_n) return ctx->fail(P2G_ERR_INVALID, "arity 2^%u exceeds the codeword length 2^%u", arity_bits, s->log_n);
    const FieldInfo *fp;
    int rc = check_field_hash(ctx, s->field, s->hash, &fp);
    if (rc) return rc;
    const FieldInfo &f = *fp;
    const unsigned tree_log = s->log_n - arity_bits;
    if (cap_height > tree_log)
        return ctx->fail(P2G_ERR_INVALID, "cap_height=%u should be at most log2(leaves.len())=%u", cap_height, tree_log);
    p2g_batch *b = new p2g_batch();
    b->field = f.id;
    b->elem = f.elem;
    b->rows = (uint64_t)1 << tree_log;
    b->width = (uint64_t)s->degree << arity_bits;  // flatten(chunk): arity extension elements, prover.rs:104-108
    b->num_cols = 0;
    b->n = 0;
    b->log_leaves = tree_log;
    b->cap_height = cap_height;
    b->num_cap = (uint64_t)1 << cap_height;
    b->num_digests = 2 * (b->rows - b->num_cap);
    b->digests_bytes = std::max<size_t>(32, b->num_digests * 32);
    b->cap_bytes = b->num_cap * 32;
    b->digests = pool_alloc(ctx, b->digests_bytes);
    b->cap = pool_alloc(ctx, b->cap_bytes);
    if (!b->digests || !b->cap) {
        pool_free(ctx, b->digests, b->digests_bytes);
        pool_free(ctx, b->cap, b->cap_bytes);
        delete b;
        return ctx->fail(P2G_ERR_OOM, "device allocation for the layer tree failed");
    }
    // the bit-reversed evaluations ARE the leaf matrix; the tree takes them over
    b->leaves = s->values;
    b->leaves_bytes = s->values_bytes;
    s->values = nullptr;
    s->values_bytes = 0;
    cudaStream_t st = ctx->stream;
    TreeShape tree = {tree_log, cap_height, b->digests, b->cap};
    launch_leaf_hash(f.id, s->repr, b->leaves, b->width, (uint32_t)b->width, 0, b->rows, tree, st, &ctx->launches);
    launch_tree_levels(f.id, s->repr, tree, st, &ctx->launches);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess && cap_out) e = cudaMemcpyAsync(cap_out, b->cap, b->cap_bytes, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    s->pending_arity_bits = (int)arity_bits;
    *tree_out = b;  // handed over even on a CUDA error so the caller can free it
    if (e != cudaSuccess) return ctx->fail(P2G_ERR_CUDA, "layer tree failed: %s", cudaGetErrorString(e));
    return P2G_OK;
}

int32_t p2g_fri_fold(p2g_ctx *ctx, p2g_fri *s, const void *beta) {
    if (!ctx) return P2G_ERR_NO_DEVICE;
    if (!ctx->subs.empty()) ctx = ctx->subs[0];  // multi-device context: its first device serves this call
    std::lock_guard<std::mutex> lock(ctx->mu);
    CK(cudaSetDevice(ctx->device));
    if (!s || !beta) return ctx->fail(P2G_ERR_INVALID, "state / beta is NULL");
    if (s->pending_arity_bits < 0) return ctx->fail(P2G_ERR_INVALID, "no committed layer to fold: call p2g_fri_commit_layer first");
    const FieldInfo *fp;
    int rc = check_field_hash(ctx, s->field, s->hash, &fp);
    if (rc) return rc;
    const FieldInfo &f = *fp;
    const unsigned a = (unsigned)s->pending_arity_bits, D = s->degree;
    const unsigned log_out = s->log_n - a;
    const uint64_t n_in = (uint64_t)1 << s->log_n, n_out = (uint64_t)1 << log_out;
    // beta arrives in the caller's representation; the kernel wants the internal one
    uint64_t beta_int[4] = {0, 0, 0, 0};
    for (unsigned k = 0; k < D; k++) {
        uint64_t v = f.id == FIELD_GL ? ((const uint64_t *)beta)[k] : ((const uint32_t *)beta)[k];
        if (f.id == FIELD_GL)
            beta_int[k] = v >= P2G_GL_P ? v - P2G_GL_P : v;
        else if (s->repr == P2G_REPR_MONTY_R32)
            beta_int[k] = v % P2G_BB_P;
        else
            beta_int[k] = f_internal(f, v % P2G_BB_P);
    }
    const size_t out_bytes = std::max<size_t>((size_t)n_out * D * f.elem, 32);
    void *nc = pool_alloc(ctx, out_bytes), *nv = pool_alloc(ctx, out_bytes);
    if (!nc || !nv) {
        pool_free(ctx, nc, out_bytes);
        pool_free(ctx, nv, out_bytes);
        return ctx->fail(P2G_ERR_OOM, "device allocation for the folded layer failed");
    }
    cudaStream_t st = ctx->stream;
    launch_fri_fold(f.id, s->repr, s->coeffs, n_in, a, beta_int, f_internal(f, ext_w(f)), nc, st, &ctx->launches);
    const uint64_t shift = f_pow(f, s->shift, (uint64_t)1 << a);  // prover.rs:124
    rc = 0;
    if (log_out == 0) {
        // a single coefficient is its own evaluation
        for (unsigned k = 0; k < D && !rc; k++)
            if (cudaMemcpyAsync((char *)nv + k * f.elem, (char *)nc + k * f.elem, f.elem, cudaMemcpyDeviceToDevice, st) != cudaSuccess)
                rc = ctx->fail(P2G_ERR_CUDA, "copy failed");
    } else {
        if (multi_pass(log_out)) rc = ensure(ctx, ctx->tmp, (size_t)n_out * D * f.elem);
        // coset_fft(shift): D base-field transforms, written as rows of D in bit-reversed order = the next leaf matrix
        if (!rc) rc = run_transform(ctx, f, s->repr, false, log_out, D, nc, ctx->tmp.ptr, nv, D, shift, 0);
    }
    cudaError_t e = cudaStreamSynchronize(st);
    if (!rc && e != cudaSuccess) rc = ctx->fail(P2G_ERR_CUDA, "fold failed: %s", cudaGetErrorString(e));
    if (rc) {
        pool_free(ctx, nc, out_bytes);
        pool_free(ctx, nv, out_bytes);
        return rc;
    }
    pool_free(ctx, s->coeffs, s->coeffs_bytes);
    s->coeffs = nc;
    s->values = nv;
    s->coeffs_bytes = s->values_bytes = out_bytes;
    s->log_n = log_out;
    s->shift = shift;
    s->pending_arity_bits = -1;
    return P2G_OK;
}

int32_t p2g_fri_finish(p2g_ctx *ctx, p2g_fri *s, uint32_t rate_bits, void *final_coeffs_out, uint64_t *count_out) {
    if (!ctx) return P2G_ERR_NO_DEVICE;
    if (!ctx->subs.empty()) ctx = ctx->subs[0];  // multi-device context: its first device serves this call
    std::lock_guard<std::mutex> lock(ctx->mu);
    CK(cudaSetDevice(ctx->device));
    if (!s) return ctx->fail(P2G_ERR_INVALID, "state is NULL");
    if (s->pending_arity_bits >= 0) return ctx->fail(P2G_ERR_INVALID, "a committed layer has not been folded");
    const uint64_t n = (uint64_t)1 << s->log_n;
    const uint64_t count = rate_bits >= 64 ? 0 : n >> rate_bits;  // coeffs.truncate(len >> rate_bits), prover.rs:129-131
    if (count_out) *count_out = count;
    if (final_coeffs_out && count) {
        std::vector<unsigned char> cols((size_t)count * s->degree * s->elem);
        for (unsigned k = 0; k < s->degree; k++)
            CK(cudaMemcpyAsync(cols.data() + (size_t)k * count * s->elem, (char *)s->coeffs + (size_t)k * n * s->elem, count * s->elem,
                               cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        for (uint64_t i = 0; i < count; i++)
            for (unsigned k = 0; k < s->degree; k++)
                memcpy((char *)final_coeffs_out + ((size_t)i * s->degree + k) * s->elem, cols.data() + ((size_t)k * count + i) * s->elem, s->elem);
    }
    return P2G_OK;
}

int32_t p2g_fri_shape(p2g_fri *s, uint32_t *log_n, uint64_t *shift) {
    if (!s) return P2G_ERR_INVALID;
    if (log_n) *log_n = s->log_n;
    if (shift) *shift = s->shift;
    return P2G_OK;
}

void p2g_fri_free(p2g_ctx *ctx, p2g_fri *s) {
    if (!ctx || !s) return;
    if (!ctx->subs.empty()) ctx = ctx->subs[0];
    std::lock_guard<std::mutex> lock(ctx->mu);
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    pool_free(ctx, s->coeffs, s->coeffs_bytes);
    pool_free(ctx, s->values, s->values_bytes);
    delete s;
}

// fri_proof_of_work (fri/prover.rs:136-188): the search only; the caller observes the witness afterwards (:181-186).
int32_t p2g_fri_pow(p2g_ctx *ctx, int32_t field, int32_t hash, const void *sponge_state, const void *input_buffer,
                    uint32_t input_len, uint32_t min_leading_zeros, uint64_t *witness_out) {
    if (!ctx) return P2G_ERR_NO_DEVICE;
    if (!ctx->subs.empty()) ctx = ctx->subs[0];  // multi-device context: its first device serves this call
    std::lock_guard<std::mutex> lock(ctx->mu);
    CK(cudaSetDevice(ctx->device));
    const FieldInfo *fp;
    int rc = check_field_hash(ctx, field, hash, &fp);
    if (rc) return rc;
    const FieldInfo &f = *fp;
    if (!sponge_state || !witness_out || (input_len && !input_buffer)) return ctx->fail(P2G_ERR_INVALID, "NULL argument");
    if (input_len >= 8) return ctx->fail(P2G_ERR_INVALID, "input_buffer holds %u elements: a Challenger never keeps a full rate", input_len);
    const unsigned width = f.id == FIELD_GL ? 12 : 16;
    // duplex_intermediate_state (:166-168): the sponge state with the buffered inputs written over its first lanes
    unsigned char st_host[16 * 8];
    memcpy(st_host, sponge_state, width * f.elem);
    memcpy(st_host, input_buffer, input_len * f.elem);
    if ((rc = ensure(ctx, ctx->misc, 256))) return rc;
    unsigned char *dev = (unsigned char *)ctx->misc.ptr;
    unsigned long long *best = (unsigned long long *)(dev + 128);
    cudaStream_t st = ctx->stream;
    CK(cudaMemcpyAsync(dev, st_host, width * f.elem, cudaMemcpyHostToDevice, st));
    const uint64_t last = (f.id == FIELD_GL ? P2G_GL_P : (uint64_t)P2G_BB_P) - 1;  // candidates run over 0..=F::neg_one() (:170)
    // batches in increasing candidate order, so the smallest hit of the first batch with a hit is the smallest overall;
    // they start at 2^16 candidates (a 16-bit grind ends in the first or second one) and double up to 2^22
    uint64_t batch = (uint64_t)1 << 16;
    const unsigned long long none = ~0ull;
    for (uint64_t start = 0; start <= last; start += batch, batch = std::min<uint64_t>(batch * 2, (uint64_t)1 << 22)) {
        CK(cudaMemcpyAsync(best, &none, 8, cudaMemcpyHostToDevice, st));
        uint64_t count = std::min<uint64_t>(batch, last - start + 1);
        launch_pow_search(f.id, dev, input_len, min_leading_zeros, start, count, last, best, st, &ctx->launches);
        unsigned long long got = none;
        CK(cudaMemcpyAsync(&got, best, 8, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        if (got != none) {
            *witness_out = got;
            return P2G_OK;
        }
        if (start + batch < start) break;  // wrapped past 2^64
    }
    return ctx->fail(P2G_ERR_INVALID, "Proof of work failed. This is highly unlikely!");  // the reference's expect(), :179
}

int32_t p2g_permute(p2g_ctx *ctx, int32_t hash, int32_t repr, void *states, uint64_t count) {
    if (ctx && !ctx->subs.empty()) {
        int32_t rc_ = p2g_permute(ctx->subs[0], hash, repr, states, count);
        if (rc_) ctx->error = ctx->subs[0]->error;
        return rc_;
    }
    if (!ctx) return P2G_ERR_NO_DEVICE;
    std::lock_guard<std::mutex> lock(ctx->mu);
    CK(cudaSetDevice(ctx->device));
    const FieldInfo &f = hash == P2G_HASH_POSEIDON_GL12 ? kGl : kBb;
    if (hash != P2G_HASH_POSEIDON_GL12 && hash != P2G_HASH_POSEIDON2_BB16) return ctx->fail(P2G_ERR_UNSUPPORTED, "unknown hash %d", hash);
    size_t bytes = (size_t)count * (f.id == FIELD_GL ? 12 * 8 : 16 * 4);
    if (!count) return P2G_OK;
    int rc = ensure(ctx, ctx->misc, bytes);
    if (rc) return rc;
    CK(cudaMemcpyAsync(ctx->misc.ptr, states, bytes, cudaMemcpyHostToDevice, ctx->stream));
    launch_permute(f.id, f.id == FIELD_BB ? repr : 0, ctx->misc.ptr, count, ctx->stream, &ctx->launches);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(states, ctx->misc.ptr, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return P2G_OK;
}

int32_t p2g_transform(p2g_ctx *ctx, int32_t field, int32_t repr, int32_t kind, uint32_t log_n, uint32_t num_cols,
                      uint32_t rate_bits, uint32_t flags, const void *in, void *out) {
    if (ctx && !ctx->subs.empty()) {
        int32_t rc_ = p2g_transform(ctx->subs[0], field, repr, kind, log_n, num_cols, rate_bits, flags, in, out);
        if (rc_) ctx->error = ctx->subs[0]->error;
        return rc_;
    }
    if (!ctx) return P2G_ERR_NO_DEVICE;
    std::lock_guard<std::mutex> lock(ctx->mu);
    CK(cudaSetDevice(ctx->device));
    if (ctx->pass_table_bytes > ((size_t)2 << 30)) drop_pass_tables(ctx);
    if (field != P2G_FIELD_GOLDILOCKS && field != P2G_FIELD_BABYBEAR) return ctx->fail(P2G_ERR_UNSUPPORTED, "unknown field");
    const FieldInfo &f = field == P2G_FIELD_GOLDILOCKS ? kGl : kBb;
    if (f.id == FIELD_GL) repr = 0;
    if (!num_cols || !in || !out) return ctx->fail(P2G_ERR_INVALID, "empty transform");
    if (kind == 0 || kind == 2) rate_bits = 0;
    if (log_n + rate_bits > f.two_adicity || log_n + rate_bits > 30 || log_n > 3 * MAX_TILE_LOG) return ctx->fail(P2G_ERR_INVALID, "transform size unsupported");
    const bool dev_in = flags & P2G_FLAG_DEVICE_INPUT, dev_out = flags & P2G_FLAG_DEVICE_OUTPUT;
    const uint64_t n = (uint64_t)1 << log_n, N = n << rate_bits;
    size_t in_bytes = (size_t)num_cols * n * f.elem;
    cudaStream_t st = ctx->stream;
    int rc;
    if ((rc = ensure(ctx, ctx->in, in_bytes))) return rc;
    CK(cudaMemcpyAsync(ctx->in.ptr, in, in_bytes, dev_in ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
    if (kind == 0 || kind == 2) {
        void *dst = out;
        if (!dev_out) {
            if ((rc = ensure(ctx, ctx->coeffs, in_bytes))) return rc;
            dst = ctx->coeffs.ptr;
        }
        if ((rc = run_transform(ctx, f, repr, true, log_n, num_cols, ctx->in.ptr, nullptr, dst, n, 0, 0))) return rc;
        if (kind == 2) {
            // coset_ifft (polynomial/mod.rs:62-72): coefficient j of the plain inverse transform times shift^-j
            const void *spow;
            if ((rc = get_pow_table(ctx, f, f_inv(f, f.generator), &spow))) return rc;
            launch_scale_powers(f.id, repr, dst, log_n, num_cols, n, spow, st, &ctx->launches);
            CK(cudaGetLastError());
        }
        if (!dev_out) CK(cudaMemcpyAsync(out, dst, in_bytes, cudaMemcpyDeviceToHost, st));
    } else if (kind == 1) {
        size_t out_bytes = (size_t)num_cols * N * f.elem;
        if ((rc = ensure(ctx, ctx->leaves, out_bytes))) return rc;
        if ((rc = ensure(ctx, ctx->misc, out_bytes))) return rc;
        const unsigned cosets = 1u << rate_bits, nb = lde_batch(log_n, cosets, in_bytes);
        if (multi_pass(log_n))
            if ((rc = ensure(ctx, ctx->tmp, (size_t)nb * in_bytes))) return rc;
        const uint64_t wN = two_adic_generator(f, log_n + rate_bits);
        for (unsigned b = 0; b < cosets; b += nb) {
            uint64_t shifts[8];
            const unsigned cur = std::min(nb, cosets - b);
            for (unsigned k = 0; k < cur; k++) shifts[k] = f_mul(f, f.generator, f_pow(f, wN, bitrev(b + k, rate_bits)));
            if ((rc = run_transform_batch(ctx, f, repr, false, log_n, num_cols, ctx->in.ptr, ctx->tmp.ptr, ctx->leaves.ptr, num_cols,
                                          shifts, cur, (uint64_t)b * n, n)))
                return rc;
        }
        void *dst = dev_out ? out : ctx->misc.ptr;
        launch_leaves_to_columns(f.id, ctx->leaves.ptr, num_cols, num_cols, log_n + rate_bits, dst, st, &ctx->launches);
        CK(cudaGetLastError());
        if (!dev_out) CK(cudaMemcpyAsync(out, dst, out_bytes, cudaMemcpyDeviceToHost, st));
    } else {
        return ctx->fail(P2G_ERR_INVALID, "unknown transform kind %d", kind);
    }
    CK(cudaStreamSynchronize(st));
    return P2G_OK;
}

// ---- multi-GPU plumbing: device buffers that other processes can map, and the fused inverse-NTT + all-gather --------
void *p2g_dev_alloc(p2g_ctx *ctx, size_t bytes) {
    if (!ctx) return nullptr;
    if (!ctx->subs.empty()) ctx = ctx->subs[0];
    cudaSetDevice(ctx->device);
    void *p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess) {
        cudaGetLastError();
        ctx->fail(P2G_ERR_OOM, "cudaMalloc(%zu) failed", bytes);
        return nullptr;
    }
    return p;
}
void p2g_dev_free(p2g_ctx *ctx, void *ptr) {
    if (ctx && !ctx->subs.empty()) ctx = ctx->subs[0];
    if (ctx) {
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->stream);
    }
    if (ptr) cudaFree(ptr);
}
int32_t p2g_ipc_export(p2g_ctx *ctx, void *dev_ptr, uint8_t handle_out[64]) {
    if (!ctx) return P2G_ERR_NO_DEVICE;
    if (!ctx->subs.empty()) ctx = ctx->subs[0];  // multi-device context: its first device serves this call
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    CK(cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, dev_ptr));
    memcpy(handle_out, &h, 64);
    return P2G_OK;
}
int32_t p2g_ipc_import(p2g_ctx *ctx, const uint8_t handle[64], void **dev_ptr_out) {
    if (!ctx) return P2G_ERR_NO_DEVICE;
    if (!ctx->subs.empty()) ctx = ctx->subs[0];  // multi-device context: its first device serves this call
    CK(cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    CK(cudaIpcOpenMemHandle(dev_ptr_out, h, cudaIpcMemLazyEnablePeerAccess));
    return P2G_OK;
}
int32_t p2g_ipc_close(p2g_ctx *ctx, void *dev_ptr) {
    if (!ctx) return P2G_ERR_NO_DEVICE;
    if (!ctx->subs.empty()) ctx = ctx->subs[0];  // multi-device context: its first device serves this call
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaIpcCloseMemHandle(dev_ptr));
    return P2G_OK;
}

int32_t p2g_memcpy_async(p2g_ctx *ctx, void *dst, const void *src, size_t bytes, int32_t kind) {
    if (!ctx) return P2G_ERR_NO_DEVICE;
    if (!ctx->subs.empty()) ctx = ctx->subs[0];  // multi-device context: its first device serves this call
    CK(cudaSetDevice(ctx->device));
    cudaMemcpyKind k = kind == 0 ? cudaMemcpyHostToDevice : (kind == 1 ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice);
    CK(cudaMemcpyAsync(dst, src, bytes, k, ctx->stream));
    return P2G_OK;
}

int32_t p2g_intt_allgather(p2g_ctx *ctx, int32_t field, int32_t repr, uint32_t log_n, uint32_t num_cols, uint32_t col_offset,
                           const void *values_dev, void *const *gather_bufs, uint32_t num_bufs) {
    if (!ctx) return P2G_ERR_NO_DEVICE;
    if (!ctx->subs.empty()) ctx = ctx->subs[0];  // multi-device context: its first device serves this call
    std::lock_guard<std::mutex> lock(ctx->mu);
    CK(cudaSetDevice(ctx->device));
    if (field != P2G_FIELD_GOLDILOCKS && field != P2G_FIELD_BABYBEAR) return ctx->fail(P2G_ERR_UNSUPPORTED, "unknown field");
    const FieldInfo &f = field == P2G_FIELD_GOLDILOCKS ? kGl : kBb;
    if (f.id == FIELD_GL) repr = 0;
    if (!num_cols || !values_dev || !gather_bufs || num_bufs == 0 || num_bufs > 9) return ctx->fail(P2G_ERR_INVALID, "bad arguments");
    if (log_n > f.two_adicity || log_n > 3 * MAX_TILE_LOG) return ctx->fail(P2G_ERR_INVALID, "transform size unsupported");
    const uint64_t n = (uint64_t)1 << log_n;
    const size_t bytes = (size_t)num_cols * n * f.elem;
    int rc;
    if ((rc = ensure(ctx, ctx->in, bytes))) return rc;
    // the strided passes run in place: work on a copy so that the caller's values survive
    CK(cudaMemcpyAsync(ctx->in.ptr, values_dev, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    void *peers[8];
    for (unsigned k = 1; k < num_bufs; k++) peers[k - 1] = (char *)gather_bufs[k] + (size_t)col_offset * n * f.elem;
    void *own = (char *)gather_bufs[0] + (size_t)col_offset * n * f.elem;
    if ((rc = run_transform(ctx, f, repr, true, log_n, num_cols, ctx->in.ptr, nullptr, own, n, 0, 0, peers, num_bufs - 1))) return rc;
    // asynchronous on the context's stream: the caller orders the cross-GPU barrier and the commit behind it
    return P2G_OK;
}

// The LDE of a shard split by PASS (see p2g.h): the strided passes column block by column block, as the blocks arrive,
// and the last pass -- which writes the leaf rows -- once over all columns.
int32_t p2g_lde_columns(p2g_ctx *ctx, int32_t field, int32_t repr, uint32_t log_n, uint32_t rate_bits, uint32_t shard_index,
                        uint32_t shard_count, uint32_t total_cols, uint32_t col_base, uint32_t num_cols, const void *coeffs_dev,
                        void *leaves_dev, uint64_t leaf_pitch, int32_t phase) {
    if (!ctx) return P2G_ERR_NO_DEVICE;
    if (!ctx->subs.empty()) ctx = ctx->subs[0];
    std::lock_guard<std::mutex> lock(ctx->mu);
    CK(cudaSetDevice(ctx->device));
    if (field != P2G_FIELD_GOLDILOCKS && field != P2G_FIELD_BABYBEAR) return ctx->fail(P2G_ERR_UNSUPPORTED, "unknown field");
    const FieldInfo &f = field == P2G_FIELD_GOLDILOCKS ? kGl : kBb;
    if (f.id == FIELD_GL) repr = 0;
    const unsigned G = shard_count ? shard_count : 1, gi = shard_count ? shard_index : 0;
    if ((G & (G - 1)) || gi >= G) return ctx->fail(P2G_ERR_INVALID, "bad shard");
    const unsigned lg = ilog2(G), L = log_n, r = rate_bits, LN = L + r;
    if (lg > r) return ctx->fail(P2G_ERR_UNSUPPORTED, "shard_count %u exceeds the 2^rate_bits = %u cosets", G, 1u << r);
    if (LN > f.two_adicity || LN > 30 || L > 3 * MAX_TILE_LOG) return ctx->fail(P2G_ERR_INVALID, "LDE size 2^%u unsupported", LN);
    if (phase != P2G_LDE_STRIDED && phase != P2G_LDE_LAST) return ctx->fail(P2G_ERR_INVALID, "bad phase");
    if (!total_cols || !num_cols || (uint64_t)col_base + num_cols > total_cols) return ctx->fail(P2G_ERR_INVALID, "bad column range");
    if (phase == P2G_LDE_LAST && (col_base || num_cols != total_cols || !leaves_dev || leaf_pitch < total_cols))
        return ctx->fail(P2G_ERR_INVALID, "the last pass covers all columns and needs the leaf rows");
    const uint64_t n = (uint64_t)1 << L;
    const unsigned cosets_local = (1u << r) >> lg, coset0 = gi * cosets_local;
    const bool multi = multi_pass(L);
    if ((multi ? phase == P2G_LDE_STRIDED : phase == P2G_LDE_LAST) && !coeffs_dev) return ctx->fail(P2G_ERR_INVALID, "coeffs_dev is NULL");
    if (!multi && phase == P2G_LDE_STRIDED) return P2G_OK;  // a single-pass transform has no strided pass
    int rc;
    // every coset of the shard in one launch per pass; the scratch holds the strided passes' output of all of them
    if (multi && (rc = ensure(ctx, ctx->tmp, (size_t)cosets_local * total_cols * n * f.elem))) return rc;
    uint64_t shifts[8];
    const uint64_t wN = two_adic_generator(f, LN);
    for (unsigned k = 0; k < cosets_local; k++) shifts[k] = f_mul(f, f.generator, f_pow(f, wN, bitrev(coset0 + k, r)));
    return run_transform_batch(ctx, f, repr, false, L, num_cols, const_cast<void *>(coeffs_dev), ctx->tmp.ptr, leaves_dev, leaf_pitch,
                               shifts, cosets_local, 0, n, nullptr, 0, phase == P2G_LDE_STRIDED ? 1 : 2, total_cols, col_base);
}

}  // extern "C"
