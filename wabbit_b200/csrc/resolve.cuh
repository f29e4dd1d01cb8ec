// Geometric resolution of lattice points on a graded leaf grid (device side).
//
// On a leaf grid every point of the level-L lattice (global integer coordinates, spacing 2^-L L_dom / Bs) lies in
// exactly one block.  The ghost-node synchronisation of the reference (sync_ghosts_generic,
// LIB/MPI/synchronize_ghosts_generic.f90:181-343) fills a ghost point of a level-L block with
//   - the coincident interior value if the owner is on level L         (stage 1, copy,        lvl_diff  0),
//   - the coincident interior value if the owner is on level L+1       (stage 2, decimation,  lvl_diff +1, ignore_Filter),
//   - an interpolation of the level-(L-1) lattice if the owner is coarser (stage 3, prediction, lvl_diff -1), where the
//     coarse lattice values are themselves of the first two kinds (gradedness: the sender's ghosts that the interpolation
//     stencil touches were filled in stages 1 and 2).
// `SrcTable` resolves the first two kinds for all points of a small box through a block lookup keyed by (level, ix, iy, iz).
#pragma once

#include <stdint.h>

struct BlockLookup {
    const unsigned long long *keys;   // open addressing, ~0 = empty
    const int *vals;
    unsigned mask;                    // capacity - 1 (power of two)
};

__host__ __device__ inline unsigned long long blk_key(int level, int ix, int iy, int iz)
{
    return ((unsigned long long)level << 57) | ((unsigned long long)(unsigned)iz << 38) | ((unsigned long long)(unsigned)iy << 19) |
           (unsigned long long)(unsigned)ix;
}

__host__ __device__ inline unsigned blk_hash(unsigned long long k)
{
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdULL;
    k ^= k >> 33;
    k *= 0xc4ceb9fe1a85ec53ULL;
    k ^= k >> 33;
    return (unsigned)k;
}

__device__ __forceinline__ int blk_lookup(const BlockLookup &L, int level, int ix, int iy, int iz)
{
    const unsigned long long k = blk_key(level, ix, iy, iz);
    unsigned h = blk_hash(k) & L.mask;
    for (;;) {
        const unsigned long long kk = L.keys[h];
        if (kk == k) return L.vals[h];
        if (kk == ~0ull) return -1;
        h = (h + 1) & L.mask;
    }
}

__device__ __forceinline__ int floor_div(int a, int b) { return a >= 0 ? a / b : -((-a + b - 1) / b); }

// Sources of a box of level-L lattice points that spans at most 3 blocks per axis.
struct SrcTable {
    int blk[27];        // level-L block of segment (sx,sy,sz), or -1
    int child[27][8];   // level-(L+1) blocks covering the same region (bit0 x, bit1 y, bit2 z), or -1
    int b0[3];          // unwrapped block coordinate of segment 0 per axis
};

// all threads of the CTA call this; lo/hi are unwrapped global lattice coordinates (level L)
__device__ inline void src_table_build(SrcTable &T, const BlockLookup &L, int level, const int lo[3], const int hi[3], int Bs, int dim,
                                       const int periodic[3], int tid, int nt)
{
    int b0[3], ns[3];
    for (int a = 0; a < 3; ++a) {
        b0[a] = a < dim ? floor_div(lo[a], Bs) : 0;
        ns[a] = a < dim ? floor_div(hi[a], Bs) - b0[a] + 1 : 1;
    }
    if (tid < 3) T.b0[tid] = b0[tid];
    const int nb = 1 << level;
    for (int t = tid; t < 27 * 9; t += nt) {
        const int e = t / 9, k = t % 9;   // k = 0: the level-L block, k = 1..8: child k-1
        const int s[3] = {e % 3, (e / 3) % 3, e / 9};
        int bc[3];
        bool ok = true;
        for (int a = 0; a < 3; ++a) {
            bc[a] = b0[a] + s[a];
            if (s[a] >= ns[a]) ok = false;
            if (a < dim && (bc[a] < 0 || bc[a] >= nb)) {
                if (periodic[a]) bc[a] = ((bc[a] % nb) + nb) % nb;
                else ok = false;
            }
            if (a >= dim) bc[a] = 0;
        }
        int v = -1;
        if (ok) {
            if (k == 0) v = blk_lookup(L, level, bc[0], bc[1], bc[2]);
            else {
                const int c = k - 1;
                if (!(dim == 2 && (c & 4)))
                    v = blk_lookup(L, level + 1, 2 * bc[0] + (c & 1), 2 * bc[1] + ((c >> 1) & 1), dim == 3 ? 2 * bc[2] + ((c >> 2) & 1) : 0);
            }
        }
        if (v < -1) v = -1;   // known by position only (wgpu_set_grid with hvy id <= 0): no data here
        if (k == 0) T.blk[e] = v;
        else T.child[e][k - 1] = v;
    }
}

// element offset (block index, offset inside the Bs^3 component) of lattice point P; blk = -1 if no leaf of level L or L+1 owns it
// roff >= 0: the owner is one level finer and roff is the point's offset inside a decimated (Bs/2)^dim component (else -1)
__device__ __forceinline__ void src_resolve(const SrcTable &T, const int P[3], int Bs, int dim, int &blk, int &off, int &roff)
{
    roff = -1;
    int seg[3], loc[3];
    for (int a = 0; a < 3; ++a) {
        if (a < dim) {
            const int bc = floor_div(P[a], Bs);
            seg[a] = bc - T.b0[a];
            loc[a] = P[a] - bc * Bs;
        } else {
            seg[a] = 0;
            loc[a] = 0;
        }
    }
    const int e = seg[0] + 3 * seg[1] + 9 * seg[2];
    blk = T.blk[e];
    if (blk < 0) {
        const int half = Bs / 2;
        const int c = (loc[0] >= half ? 1 : 0) | (loc[1] >= half ? 2 : 0) | ((dim == 3 && loc[2] >= half) ? 4 : 0);
        blk = T.child[e][c];
        for (int a = 0; a < dim; ++a) loc[a] = 2 * loc[a] - (loc[a] >= half ? Bs : 0);
        roff = ((loc[2] >> 1) * half + (loc[1] >> 1)) * half + (loc[0] >> 1);
    }
    off = (loc[2] * Bs + loc[1]) * Bs + loc[0];
}

__device__ __forceinline__ void src_resolve(const SrcTable &T, const int P[3], int Bs, int dim, int &blk, int &off)
{
    int roff;
    src_resolve(T, P, Bs, dim, blk, off, roff);
}
