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
}

void build_poseidon2_bb_tables(Poseidon2BbTables *T) {
    for (int r = 0; r < 8; r++)
        for (int i = 0; i < 16; i++) T->ext_rc[r][i] = BbHost::to_internal(P2G_BB_P2_EXTERNAL_RC[r][i] % P2G_BB_P);
    for (int r = 0; r < 13; r++) T->int_rc[r] = BbHost::to_internal(P2G_BB_P2_INTERNAL_RC[r] % P2G_BB_P);
}

}  // namespace p2g
