// NEXT ROUND — CPU model of the planned warp-resident RANK / MTFT inverse (tests/test_sbrt_inv_warp_host.py): the list lives in 32 lanes
// x 8 consecutive ranks, the new rank of the decoded symbol is the number of entries above it whose q is larger than the new q (a ballot /
// popc over a comparison, valid because the list is always sorted by q), and the move-up is a shift of the ranks in between.
// Reference: v2/transform/SBRT.go:177-226 (Inverse), modes :93-111.
#include <cstdint>

extern "C" int64_t sbrt_inverse_warp_model(int mode, const uint8_t* src, int64_t n, uint8_t* dst) {
    const int m2 = mode == 2 ? -1 : 0, s = mode == 2 ? 1 : 0;  // MTFT: qc = i; RANK: qc = (i + p[c]) >> 1
    struct E {
        int32_t q, p;
        uint8_t sym;
    };
    E lane[32][8];  // lane l holds ranks 8 l .. 8 l + 7
    for (int r = 0; r < 256; r++) lane[r >> 3][r & 7] = E{0, 0, uint8_t(r)};
    for (int64_t i = 0; i < n; i++) {
        const int r = src[i];
        const E cur = lane[r >> 3][r & 7];  // one shuffle from lane r / 8
        dst[i] = cur.sym;
        const int32_t qc = int32_t(((i & -1) + (int64_t(cur.p) & m2)) >> s);
        // every lane counts its entries above rank r with q > qc; the sum is the new rank (one popc-add reduction)
        int nr = 0;
        for (int l = 0; l < 32; l++)
            for (int k = 0; k < 8; k++) {
                const int rank = 8 * l + k;
                if (rank < r && lane[l][k].q > qc) nr++;
            }
        // ranks nr .. r-1 move down by one: inside a lane a register shift, across lanes one shuffle of the lane's last entry
        for (int rank = r; rank > nr; rank--) lane[rank >> 3][rank & 7] = lane[(rank - 1) >> 3][(rank - 1) & 7];
        lane[nr >> 3][nr & 7] = E{qc, int32_t(i), cur.sym};
    }
    return n;
}
