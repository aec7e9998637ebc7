// Host-side construction of the Poseidon / Poseidon2 constant tables uploaded to __constant__ memory.
#include "p2g_constants.h"
#include "p2g_internal.h"

namespace p2g {

void build_poseidon_gl_tables(PoseidonGlTables *T) {
    for (int r = 0; r < 31; r++)
        for (int i = 0; i < 12; i++) {
            u64 c = r < 30 ? P2G_GL_POSEIDON_RC[r * 12 + i] % GlHost::P : 0;
            if (r == 0) T->rc0[i] = c;
            T->rcl[r][0][i] = (unsigned)(c & 0x3fffff);
            T->rcl[r][1][i] = (unsigned)((c >> 22) & 0x3fffff);
            T->rcl[r][2][i] = (unsigned)(c >> 44);
        }
    // limb-domain partial rounds: push the constants of lanes 1..11 forward through the MDS layers.
    // true state x = x' + delta (delta[0] = 0);  M S(x) + c = M S(x') + (M delta + c) =: M S(x') + full;  the kernel adds
    // full[0] to lane 0 and the bare bias to lanes 1..11, so delta' = full - bias on lanes 1..11.
    u64 delta[12] = {0};
    for (int k = 0; k < P2G_PGL_LIMB_ROUNDS; k++) {
        u64 full[12];
        for (int r = 0; r < 12; r++) {
            u64 acc = P2G_GL_POSEIDON_RC[(5 + k) * 12 + r] % GlHost::P;
            for (int i = 0; i < 12; i++) acc = GlHost::add(acc, GlHost::mul(P2G_GL_MDS_CIRC[i], delta[(i + r) % 12]));
            full[r] = acc;  // the diagonal entry multiplies delta[0] = 0
        }
        u64 d = GlHost::sub(full[0], P2G_PGL_LIMB_BIAS);
        T->rclb[k][0] = (unsigned)(d & 0x3fffff) + P2G_PGL_LIMB_BIAS;
        T->rclb[k][1] = (unsigned)((d >> 22) & 0x3fffff);
        T->rclb[k][2] = (unsigned)(d >> 44);
        T->rclb[k][3] = 0;
        delta[0] = 0;
        for (int i = 1; i < 12; i++) delta[i] = GlHost::sub(full[i], P2G_PGL_LIMB_BIAS);
    }
    T->rc_limb_exit[0] = 0;
    for (int i = 1; i < 12; i++) T->rc_limb_exit[i] = GlHost::sub(delta[i], 1024);  // the join adds 2^10 to limb 0
}

void build_poseidon2_bb_tables(Poseidon2BbTables *T) {
    for (int r = 0; r < 8; r++)
        for (int i = 0; i < 16; i++) T->ext_rc[r][i] = BbHost::to_internal(P2G_BB_P2_EXTERNAL_RC[r][i] % P2G_BB_P) - P2G_BB_P;
    for (int r = 0; r < 13; r++) T->int_rc[r] = BbHost::to_internal(P2G_BB_P2_INTERNAL_RC[r] % P2G_BB_P) - P2G_BB_P;
}

}  // namespace p2g
