// shard_map.h — index arithmetic of the row-sharded, tiled upper triangle.  Plain C++ (no CUDA types) so
// that the CPU tier can check it exhaustively (tests/test_shard_map.py compiles this header with g++):
// every kernel that walks the covariance maps its blockIdx through shard_tile().
#pragma once
#include <math.h>
#include <stddef.h>

#if defined(__CUDACC__)
#define CSLAM_HD __host__ __device__ __forceinline__
#else
#define CSLAM_HD inline
#endif

namespace cslam {

// Row sharding of the covariance over GPUs: block-cyclic 128-row tiles (tile tr belongs to rank
// tr % world; local tile index tr / world).  world == 1 is the identity, so every kernel is
// written shard-aware and the single-GPU path pays nothing.
constexpr int kShardRows = 128;
struct Shard {
    int rank;
    int world;
};
CSLAM_HD bool shard_owns(Shard s, int i) { return ((i >> 7) % s.world) == s.rank; }
CSLAM_HD size_t shard_lrow(Shard s, int i) {
    return (size_t)((i >> 7) / s.world) * kShardRows + (size_t)(i & 127);
}

// Linear tile index -> (tr, tc) over the tiles of the upper triangle that THIS RANK owns, row-major:
// owned tile rows are tr = l*world + rank (l = 0,1,...), row tr holds the nt - tr tiles tc >= tr, so
//   first(l) = l*(nt - rank) - world*l*(l-1)/2          (world = 1: the plain triangular index)
CSLAM_HD long long shard_first_tile(long long l, int nt, Shard sh) {
    return l * (long long)(nt - sh.rank) - (long long)sh.world * l * (l - 1) / 2;
}
// The quadratic first(l) <= t is inverted in double precision and the two integer loops make the
// result exact whatever the rounding of the first guess.  (A single-precision first guess that also
// handed the local row index to the caller saved ~60 instructions per CTA but pushed the 64-register
// multi-update pass into spills: 2.25 -> 2.51 ms; measured and reverted.)
CSLAM_HD void shard_tile(long long t, int nt, Shard sh, int& tr, int& tc) {
    const double w = (double)sh.world;
    const double b = (double)(nt - sh.rank) + 0.5 * w;
    long long l = (long long)floor((b - sqrt(b * b - 2.0 * w * (double)t)) / w);
    if (l < 0) l = 0;
    while (shard_first_tile(l, nt, sh) > t) l--;
    while (shard_first_tile(l + 1, nt, sh) <= t) l++;
    tr = (int)l * sh.world + sh.rank;
    tc = tr + (int)(t - shard_first_tile(l, nt, sh));
}
// number of owned tiles (host side)
inline long long shard_tile_count(int nt, Shard sh) {
    long long cnt = 0;
    for (int tr = sh.rank; tr < nt; tr += sh.world) cnt += nt - tr;
    return cnt;
}

}  // namespace cslam
