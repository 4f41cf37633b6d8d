// CPU-tier exhaustive check of conan_slam_b200/csrc/shard_map.h (compiled by tests/test_shard_map.py):
// for every (nt, world, rank) the linear tile index t must enumerate exactly the owned tiles of the upper
// triangle in row-major order, and shard_lrow / shard_owns must agree with the block-cyclic row layout.
#include <cstdio>
#include <cstdlib>

#include "../conan_slam_b200/csrc/shard_map.h"

using namespace cslam;

static int check(int nt, int world) {
    for (int rank = 0; rank < world; rank++) {
        const Shard sh{rank, world};
        long long t = 0;
        int l = 0;
        for (int tr = rank; tr < nt; tr += world, l++) {
            if (shard_first_tile(l, nt, sh) != t) return 1;
            for (int tc = tr; tc < nt; tc++, t++) {
                int a, b;
                shard_tile(t, nt, sh, a, b);
                if (a != tr || b != tc) {
                    printf("nt=%d world=%d rank=%d t=%lld: got (%d,%d) want (%d,%d)\n", nt, world, rank, t, a, b, tr, tc);
                    return 1;
                }
            }
            // rows of the tile row are consecutive in local storage, and owned by this rank only
            for (int r = 0; r < kShardRows; r += 37) {
                const int i = tr * kShardRows + r;
                if (!shard_owns(sh, i) || shard_lrow(sh, i) != (size_t)l * kShardRows + (size_t)r) return 2;
                for (int o = 0; o < world; o++)
                    if (o != rank && shard_owns(Shard{o, world}, i)) return 3;
            }
        }
        if (t != shard_tile_count(nt, sh)) return 4;
    }
    return 0;
}

int main() {
    const int worlds[] = {1, 2, 3, 4, 8};
    for (int w : worlds) {
        for (int nt = 1; nt <= 330; nt++)
            if (int rc = check(nt, w)) return rc;
        const int big[] = {626, 938, 1251, 1876, 3907};  // 40k / 60k-state maps at 64/128-wide tiles, 500k rows
        for (int nt : big)
            if (int rc = check(nt, w)) return rc;
    }
    printf("shard map ok\n");
    return 0;
}
