// libwabbit_host.so -- host-side forest metadata (see include/wabbit_host.h).
// Plain C++17; produces hvy_active / level / hvy_neighbor(168) tables with WABBIT's conventions.
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <array>
#include <numeric>
#include <unordered_map>
#include <utility>
#include <vector>

#include "wabbit_host.h"

namespace {

struct Blk {
    int level;
    int ix[3];
    int64_t tc;
    uint64_t key;
    int rank, hvy;   // owner and 1-based hvy id
};

inline uint64_t pos_hash(int level, const int ix[3])
{
    return ((uint64_t)level << 58) ^ ((uint64_t)(uint32_t)ix[2] << 38) ^ ((uint64_t)(uint32_t)ix[1] << 19) ^ (uint64_t)(uint32_t)ix[0];
}

// WABBIT's Hilbert curve (treecode_to_hilbertcode_2D / _3D, LIB/MESH/treecode_to_hilbertcode_2D.f90, ..._3D.f90:11; the pattern automaton of
// M. Bader, "Space-Filling Curves", 2012, p. 115): walking down the treecode, the digit of level k is mapped to its position in the current
// basic pattern (4 patterns in 2-D, 12 in 3-D; the first is pattern 1) and the pattern of level k+1 follows from the pattern and the digit of
// level k.  The position replaces the digit in place: the Hilbert code of a block on level J has its digit k at bit (Jmax - k) * dim, zeros
// behind, exactly like the numerical treecode, which is what balanceLoad_tree sorts by (balanceLoad_tree.f90:225-250).  The two tables per
// dimension are the reference's (rows: pattern 1.., columns: treecode digit); tests/test_host.py pins them against the per-rank block
// lists of the reference's multi-rank fixture files.
const unsigned char HIL2_NEXT[4][4] = {{2, 4, 1, 1}, {1, 2, 3, 2}, {3, 3, 2, 4}, {4, 1, 4, 3}};
const unsigned char HIL2_POS[4][4] = {{0, 3, 1, 2}, {0, 1, 3, 2}, {2, 1, 3, 0}, {2, 3, 1, 0}};
const unsigned char HIL3_NEXT[12][8] = {{3, 5, 8, 5, 4, 11, 8, 11},  {12, 3, 12, 7, 6, 4, 6, 7}, {5, 12, 1, 2, 9, 9, 1, 2},   {10, 10, 1, 2, 11, 6, 1, 2},
                                        {1, 6, 7, 6, 3, 3, 10, 10},  {4, 4, 9, 9, 5, 2, 5, 8},   {2, 5, 10, 5, 2, 11, 9, 11}, {12, 1, 12, 10, 6, 1, 6, 9},
                                        {7, 8, 3, 3, 7, 8, 11, 6},   {7, 8, 5, 12, 7, 8, 4, 4},  {4, 4, 9, 9, 1, 12, 7, 12},  {11, 2, 11, 8, 3, 3, 10, 10}};
const unsigned char HIL3_POS[12][8] = {{0, 1, 3, 2, 7, 6, 4, 5}, {6, 7, 5, 4, 1, 0, 2, 3}, {0, 7, 1, 6, 3, 4, 2, 5}, {4, 3, 5, 2, 7, 0, 6, 1},
                                       {0, 3, 7, 4, 1, 2, 6, 5}, {2, 1, 5, 6, 3, 0, 4, 7}, {4, 5, 7, 6, 3, 2, 0, 1}, {2, 3, 1, 0, 5, 4, 6, 7},
                                       {2, 5, 3, 4, 1, 6, 0, 7}, {6, 1, 7, 0, 5, 2, 4, 3}, {6, 5, 1, 2, 7, 4, 0, 3}, {4, 7, 3, 0, 5, 6, 2, 1}};

uint64_t hilbert_code(int dim, int level, int Jmax, int64_t tc)
{
    const unsigned dmask = (1u << dim) - 1u;
    uint64_t code = 0;
    int pattern = 1, prev = 0;
    for (int k = 1; k <= level; ++k) {
        const int sh = (Jmax - k) * dim;
        const int digit = (int)((uint64_t)tc >> sh) & (int)dmask;
        if (k > 1) pattern = dim == 3 ? HIL3_NEXT[pattern - 1][prev] : HIL2_NEXT[pattern - 1][prev];
        const uint64_t pos = dim == 3 ? HIL3_POS[pattern - 1][digit] : HIL2_POS[pattern - 1][digit];
        code |= pos << sh;
        prev = digit;
    }
    return code;
}

}  // namespace

struct whost_forest {
    int dim, Jmax, sfc, n_ranks, N;
    int periodic[3];
    std::vector<Blk> blocks;                       // SFC order
    std::vector<std::vector<int>> rank_blocks;     // indices into blocks per rank
    // open-addressing lookup (level, ix) -> index into blocks; read-only after build(), so the neighbour search can run in parallel
    std::vector<uint64_t> hkeys;
    std::vector<int> hvals;
    uint64_t hmask = 0;
    std::vector<std::vector<int32_t>> nbr;         // per rank: ld*168, built on first use (ensure_neighbors): the device derives its own
    bool nbr_ready = false;                        // topology from the block positions (wgpu_set_grid), so most forests never need it
    bool uniform = true;

    static uint64_t mix(uint64_t k)
    {
        k ^= k >> 33;
        k *= 0xff51afd7ed558ccdULL;
        k ^= k >> 33;
        return k;
    }
    void lookup_build()
    {
        uint64_t cap = 64;
        while (cap < blocks.size() * 2 + 2) cap <<= 1;
        hkeys.assign(cap, ~0ull);
        hvals.assign(cap, -1);
        hmask = cap - 1;
        for (int i = 0; i < (int)blocks.size(); ++i) {
            const uint64_t key = pos_hash(blocks[i].level, blocks[i].ix);
            uint64_t h = mix(key) & hmask;
            while (hkeys[h] != ~0ull && hkeys[h] != key) h = (h + 1) & hmask;
            hkeys[h] = key;
            hvals[h] = i;
        }
    }
    int find(int level, const int ix[3]) const
    {
        const uint64_t key = pos_hash(level, ix);
        uint64_t h = mix(key) & hmask;
        while (hkeys[h] != ~0ull) {
            if (hkeys[h] == key) return hvals[h];
            h = (h + 1) & hmask;
        }
        return -1;
    }
};

extern "C" {

int64_t whost_encode(int32_t dim, int32_t level, int32_t Jmax, const int32_t ixyz[3])
{
    // digit bit0 <- y, bit1 <- x, bit2 <- z ; level bit i of the coordinate goes to digit (i + Jmax - level)
    int64_t tc = 0;
    const int p[3] = {ixyz[1], ixyz[0], dim == 3 ? ixyz[2] : 0};
    for (int d = 0; d < dim; ++d)
        for (int i = 0; i < level; ++i) tc |= (int64_t)((p[d] >> i) & 1) << ((i + Jmax - level) * dim + d);
    return tc;
}

int32_t whost_decode(int32_t dim, int32_t level, int32_t Jmax, int64_t tc, int32_t ixyz[3])
{
    int p[3] = {0, 0, 0};
    for (int d = 0; d < dim; ++d)
        for (int i = 0; i < level; ++i) p[d] |= (int)((tc >> ((i + Jmax - level) * dim + d)) & 1) << i;
    ixyz[0] = p[1];
    ixyz[1] = p[0];
    ixyz[2] = p[2];
    return 0;
}

uint64_t whost_sfc_key(int32_t dim, int32_t sfc, int32_t level, int32_t Jmax, const int32_t ixyz[3])
{
    const int64_t tc = whost_encode(dim, level, Jmax, ixyz);
    if (sfc == WHOST_SFC_HILBERT) return hilbert_code(dim, level, Jmax, tc);
    return (uint64_t)tc;   // sfc_z: the position on the Z curve is the treecode itself (balanceLoad_tree.f90:204-222)
}

static void build(whost_forest *f)
{
    const int dim = f->dim;
    for (auto &b : f->blocks) {
        b.tc = whost_encode(dim, b.level, f->Jmax, b.ix);
        b.key = whost_sfc_key(dim, f->sfc, b.level, f->Jmax, b.ix);
    }
    std::sort(f->blocks.begin(), f->blocks.end(), [](const Blk &a, const Blk &b) { return a.key < b.key; });
    const int nb = (int)f->blocks.size();
    // contiguous chunks: the first (nb mod P) ranks hold one block more (balanceLoad_tree.f90:600-640)
    f->rank_blocks.assign(f->n_ranks, {});
    int pos = 0;
    for (int r = 0; r < f->n_ranks; ++r) {
        const int cnt = nb / f->n_ranks + (r < nb % f->n_ranks ? 1 : 0);
        for (int k = 0; k < cnt; ++k, ++pos) {
            f->blocks[pos].rank = r;
            f->blocks[pos].hvy = k + 1;
            f->rank_blocks[r].push_back(pos);
        }
    }
    f->lookup_build();
    // every leaf on one level <=> every neighbour relation is same-level (the grid is complete)
    f->uniform = true;
    for (int i = 1; i < nb; ++i)
        if (f->blocks[i].level != f->blocks[0].level) f->uniform = false;
    f->nbr_ready = false;
}

static void ensure_neighbors(const whost_forest *cf)
{
    whost_forest *f = const_cast<whost_forest *>(cf);
    if (f->nbr_ready) return;
    const int dim = f->dim, nb = (int)f->blocks.size();
    // neighbour search, one direction at a time (find_neighbor, LIB/MESH/find_neighbors.f90:18-180)
    // hvy_neighbor(ld, 168) per rank with ld = number of active blocks of the rank (hvy ids are 1..ld): the table costs
    // 672 B per block instead of 672 B per allocated slot
    f->nbr.resize(f->n_ranks);
    for (int r = 0; r < f->n_ranks; ++r) f->nbr[r].assign(f->rank_blocks[r].size() * 168, -1);
    f->uniform = true;
    const int vary_tc[3] = {2, 1, 4};   // digit bit of x, y, z
    int any_jump = 0;
    // direction-major: for one direction all blocks write into (at most 9) rows of the column-major table at consecutive positions
    // (blocks are in space-filling-curve order = hvy order), instead of every block scattering its ~30 entries over 168 rows
    for (int dz = (dim == 3 ? -1 : 0); dz <= (dim == 3 ? 1 : 0); ++dz)
        for (int dy = -1; dy <= 1; ++dy)
            for (int dx = -1; dx <= 1; ++dx) {
                if (!dx && !dy && !dz) continue;
                const int d[3] = {dx, dy, dz};
                const int nzero = (dx == 0) + (dy == 0) + (dz == 0);
                int n_free = 1 << nzero;
                if (dim == 2) n_free /= 2;
                // digits of the (virtual) children touching this side
                int append[4] = {0, 0, 0, 0};
                int apply_free = 1;
                for (int a = 0; a < dim; ++a) {
                    if (d[a] == 0) {
                        for (int k = 0; k < 4; ++k) append[k] += vary_tc[a] * ((k / apply_free) % 2);
                        apply_free += 1;
                    } else if (d[a] == 1) {
                        for (int k = 0; k < 4; ++k) append[k] += vary_tc[a];
                    }
                }
                // same-level slot code
                int code;
                if (nzero == 2) {
                    code = 1;
                    for (int a = 0; a < 3; ++a) {
                        if (d[a] != 0) code += 8 * a;
                        if (d[a] == 1) code += 4;
                    }
                } else if (nzero == 1) {
                    code = 25;
                    int af = 1;
                    for (int a = 0; a < 3; ++a) {
                        if (d[a] == 0) code += 8 * (2 - a);
                        else {
                            if (d[a] == 1) code += af * 2;
                            af++;
                        }
                    }
                } else {
                    code = 49;
                    for (int a = 0; a < 3; ++a)
                        if (d[a] == 1) code += 1 << a;
                }
#pragma omp parallel for schedule(static) reduction(| : any_jump)
                for (int i = 0; i < nb; ++i) {
                    const Blk &b = f->blocks[i];
                    int32_t *row = f->nbr[b.rank].data();
                    const size_t ld = f->rank_blocks[b.rank].size();
                    auto set = [&](int cd, int j) { row[(size_t)(cd - 1) * ld + (b.hvy - 1)] = f->blocks[j].rank * f->N + f->blocks[j].hvy; };
                    const int tc_last = b.level > 0 ? (int)((b.tc >> ((f->Jmax - b.level) * dim)) & ((1 << dim) - 1)) : 0;
                    int code_coarser = -1;
                    for (int k = 0; k < n_free; ++k)
                        if (tc_last == append[k]) code_coarser = code + k;
                    // same level
                    const int nblk = 1 << b.level;
                    int p[3] = {0, 0, 0};
                    bool outside = false;
                    for (int a = 0; a < dim; ++a) {
                        p[a] = b.ix[a] + d[a];
                        if (p[a] < 0 || p[a] >= nblk) {
                            if (f->periodic[a]) p[a] = (p[a] + nblk) & (nblk - 1);
                            else outside = true;
                        }
                    }
                    if (outside) continue;
                    int j = f->find(b.level, p);
                    if (j >= 0) {
                        set(code, j);
                        continue;
                    }
                    // finer neighbours: neighbour of each virtual child across this side
                    bool found_finer = false;
                    if (b.level < f->Jmax) {
                        for (int k = 0; k < n_free; ++k) {
                            int q[3] = {0, 0, 0};
                            const int nblk2 = nblk * 2;
                            for (int a = 0; a < dim; ++a) {
                                const int bit = (append[k] & vary_tc[a]) ? 1 : 0;
                                q[a] = 2 * b.ix[a] + bit + d[a];
                                q[a] = (q[a] + nblk2) & (nblk2 - 1);
                            }
                            j = f->find(b.level + 1, q);
                            if (j < 0) break;
                            set(code + k + 112, j);
                            found_finer = true;
                            any_jump |= 1;
                        }
                    }
                    if (found_finer) continue;
                    // coarser neighbour
                    if (code_coarser != -1 && b.level > 0) {
                        int q[3] = {p[0] >> 1, p[1] >> 1, p[2] >> 1};
                        j = f->find(b.level - 1, q);
                        if (j >= 0) {
                            set(code_coarser + 56, j);
                            any_jump |= 1;
                        }
                    }
                }
            }
    (void)any_jump;
    f->nbr_ready = true;
}

int32_t whost_create_from_blocks(int32_t dim, int32_t Jmax, int32_t sfc, int32_t n_ranks, int32_t N, const int32_t periodic[3], int32_t n,
                                 const int32_t *level, const int32_t *ixyz, whost_forest **out)
{
    if (!out || (dim != 2 && dim != 3) || n_ranks < 1 || N < 1 || n < 0 || Jmax < 0 || Jmax > 20) return 1;
    if ((int64_t)n > (int64_t)N * n_ranks) return 2;
    if ((n + n_ranks - 1) / n_ranks > N) return 2;
    whost_forest *f = new whost_forest();
    f->dim = dim;
    f->Jmax = Jmax;
    f->sfc = sfc;
    f->n_ranks = n_ranks;
    f->N = N;
    for (int a = 0; a < 3; ++a) f->periodic[a] = periodic ? periodic[a] : 1;
    f->blocks.resize(n);
    for (int i = 0; i < n; ++i) {
        Blk &b = f->blocks[i];
        b.level = level[i];
        if (b.level < 0 || b.level > Jmax) {
            delete f;
            return 3;
        }
        for (int a = 0; a < 3; ++a) b.ix[a] = a < dim ? ixyz[3 * i + a] : 0;
    }
    build(f);
    *out = f;
    return 0;
}

int32_t whost_create_uniform(int32_t dim, int32_t J, int32_t Jmax, int32_t sfc, int32_t n_ranks, int32_t N, const int32_t periodic[3],
                             whost_forest **out)
{
    if (J < 0 || J > Jmax) return 3;
    const int nb1 = 1 << J;
    const int64_t n = dim == 3 ? (int64_t)nb1 * nb1 * nb1 : (int64_t)nb1 * nb1;
    std::vector<int32_t> level((size_t)n, J), ixyz((size_t)n * 3, 0);
    int64_t i = 0;
    for (int z = 0; z < (dim == 3 ? nb1 : 1); ++z)
        for (int y = 0; y < nb1; ++y)
            for (int x = 0; x < nb1; ++x, ++i) {
                ixyz[3 * i] = x;
                ixyz[3 * i + 1] = y;
                ixyz[3 * i + 2] = z;
            }
    return whost_create_from_blocks(dim, Jmax, sfc, n_ranks, N, periodic, (int32_t)n, level.data(), ixyz.data(), out);
}

int32_t whost_destroy(whost_forest *f)
{
    delete f;
    return 0;
}

int32_t whost_n_blocks(const whost_forest *f) { return f ? (int32_t)f->blocks.size() : 0; }
int32_t whost_n_active(const whost_forest *f, int32_t rank) { return (f && rank >= 0 && rank < f->n_ranks) ? (int32_t)f->rank_blocks[rank].size() : 0; }

int32_t whost_get_active(const whost_forest *f, int32_t rank, int32_t *hvy_active, int32_t *level, int32_t *ixyz, int64_t *treecode)
{
    if (!f || rank < 0 || rank >= f->n_ranks) return 1;
    int k = 0;
    for (int i : f->rank_blocks[rank]) {
        const Blk &b = f->blocks[i];
        if (hvy_active) hvy_active[k] = b.hvy;
        if (level) level[k] = b.level;
        if (ixyz)
            for (int a = 0; a < 3; ++a) ixyz[3 * k + a] = b.ix[a];
        if (treecode) treecode[k] = b.tc;
        ++k;
    }
    return 0;
}

int32_t whost_get_neighbors(const whost_forest *f, int32_t rank, int32_t *hvy_neighbor)
{
    if (!f || rank < 0 || rank >= f->n_ranks || !hvy_neighbor) return 1;
    ensure_neighbors(f);
    memcpy(hvy_neighbor, f->nbr[rank].data(), sizeof(int32_t) * f->nbr[rank].size());
    return 0;
}

const int32_t *whost_neighbors_ptr(const whost_forest *f, int32_t rank)
{
    if (!f || rank < 0 || rank >= f->n_ranks) return nullptr;
    ensure_neighbors(f);
    return f->nbr[rank].data();
}

int32_t whost_is_uniform(const whost_forest *f) { return f && f->uniform ? 1 : 0; }

// ---------------------------------------------------------------------------------------------------------------------
// Halo plan of one rank, from the block positions (no 168-slot table): which blocks of other ranks appear in a neighbour relation of the
// rank's blocks (same level, finer, coarser: find_neighbor's three cases, LIB/MESH/find_neighbors.f90:60-180) and which of its own blocks
// the peers mirror.  Relations are symmetric (A sees B on its level <=> B sees A; A sees the finer B <=> B sees the coarser A), so both
// lists follow from the rank's own blocks.  The reference derives the same information per synchronisation from hvy_neighbor in
// prepare_ghost_synch_metadata (LIB/MPI/synchronize_ghosts_generic.f90:352-694).
//   halo_*      blocks of other ranks, ascending lgt id (= owner-major, hvy ascending: the order they arrive in and of the halo slots);
//               halo_fine[k] = 1 if the block is a FINER neighbour of one of the rank's blocks (its filtered copy travels too)
//   send_*      own blocks (hvy ids) the peers mirror, peer-major, hvy ascending; fine_send_*: own blocks that are finer neighbours of a
//               peer's block
// Output arrays hold whost_n_blocks entries at most (send lists: n_ranks * n_active(rank)); counts arrays hold n_ranks entries.
// ---------------------------------------------------------------------------------------------------------------------
int32_t whost_halo_plan(const whost_forest *f, int32_t rank, int32_t *n_halo, int32_t *halo_lgt, int32_t *halo_level, int64_t *halo_tc,
                        int32_t *halo_fine, int32_t *recv_counts, int32_t *n_send, int32_t *send_hvy, int32_t *send_counts, int32_t *n_fine_send,
                        int32_t *fine_send_hvy, int32_t *fine_send_counts)
{
    if (!f || rank < 0 || rank >= f->n_ranks || !n_halo || !n_send || !n_fine_send || !recv_counts || !send_counts || !fine_send_counts) return 1;
    const int dim = f->dim, W = f->n_ranks, nchild = 1 << dim;
    const std::vector<int> &mine = f->rank_blocks[rank];
    const int nm = (int)mine.size(), ntot = (int)f->blocks.size();
    // per foreign block: bit0 related, bit1 finer neighbour of one of mine; per (own block, peer): bit0 related, bit1 the own block is the finer one
    // (every own block writes only its own row of `own`; marks of foreign blocks are collected per thread and merged: the loop over the
    // own blocks -- ~60 hash lookups each -- runs on the rank's share of the host cores)
    std::vector<unsigned char> foreign(ntot, 0), own((size_t)nm * W, 0);
#pragma omp parallel
    {
    std::vector<std::pair<int, unsigned char>> marks;
#pragma omp for schedule(static) nowait
    for (int m = 0; m < nm; ++m) {
        const Blk &b = f->blocks[mine[m]];
        const int nblk = 1 << b.level;
        auto mark = [&](int j, bool j_is_finer, bool j_is_coarser) {
            const Blk &o = f->blocks[j];
            if (o.rank == rank) return;
            unsigned char v = 1 | (j_is_finer ? 2 : 0);
            marks.emplace_back(j, v);
            own[(size_t)m * W + o.rank] |= 1 | (j_is_coarser ? 2 : 0);
        };
        for (int dz = (dim == 3 ? -1 : 0); dz <= (dim == 3 ? 1 : 0); ++dz)
            for (int dy = -1; dy <= 1; ++dy)
                for (int dx = -1; dx <= 1; ++dx) {
                    if (!dx && !dy && !dz) continue;
                    const int d[3] = {dx, dy, dz};
                    int p[3] = {0, 0, 0};
                    bool outside = false;
                    for (int a = 0; a < dim; ++a) {
                        p[a] = b.ix[a] + d[a];
                        if (p[a] < 0 || p[a] >= nblk) {
                            if (f->periodic[a]) p[a] = (p[a] + nblk) & (nblk - 1);
                            else outside = true;
                        }
                    }
                    if (outside) continue;
                    int j = f->find(b.level, p);
                    if (j >= 0) {
                        mark(j, false, false);
                        continue;
                    }
                    bool finer = false;
                    if (b.level < f->Jmax) {
                        const int nblk2 = nblk * 2;
                        for (int c = 0; c < nchild; ++c) {   // virtual daughters that touch this side
                            bool touches = true;
                            int q[3] = {0, 0, 0};
                            for (int a = 0; a < dim; ++a) {
                                const int bit = (c >> a) & 1;
                                if ((d[a] > 0 && !bit) || (d[a] < 0 && bit)) touches = false;
                                q[a] = (2 * b.ix[a] + bit + d[a] + nblk2) & (nblk2 - 1);
                            }
                            if (!touches) continue;
                            j = f->find(b.level + 1, q);
                            if (j >= 0) {
                                mark(j, true, false);
                                finer = true;
                            }
                        }
                    }
                    if (finer || b.level == 0) continue;
                    const int q[3] = {p[0] >> 1, p[1] >> 1, p[2] >> 1};
                    j = f->find(b.level - 1, q);
                    if (j >= 0) mark(j, false, true);
                }
    }
#pragma omp critical
    for (const auto &mk : marks) foreign[mk.first] |= mk.second;
    }
    for (int r = 0; r < W; ++r) recv_counts[r] = send_counts[r] = fine_send_counts[r] = 0;
    int nh = 0;
    for (int r = 0; r < W; ++r) {   // blocks are stored rank by rank in hvy order: ascending lgt id
        if (r == rank) continue;
        for (int j : f->rank_blocks[r]) {
            if (!foreign[j]) continue;
            const Blk &o = f->blocks[j];
            if (halo_lgt) halo_lgt[nh] = o.rank * f->N + o.hvy;
            if (halo_level) halo_level[nh] = o.level;
            if (halo_tc) halo_tc[nh] = o.tc;
            if (halo_fine) halo_fine[nh] = (foreign[j] & 2) ? 1 : 0;
            ++nh;
            ++recv_counts[r];
        }
    }
    *n_halo = nh;
    int ns = 0, nf = 0;
    for (int r = 0; r < W; ++r) {
        if (r == rank) continue;
        for (int m = 0; m < nm; ++m) {
            const unsigned char v = own[(size_t)m * W + r];
            if (v & 1) {
                if (send_hvy) send_hvy[ns] = f->blocks[mine[m]].hvy;
                ++ns;
                ++send_counts[r];
            }
            if (v & 2) {
                if (fine_send_hvy) fine_send_hvy[nf] = f->blocks[mine[m]].hvy;
                ++nf;
                ++fine_send_counts[r];
            }
        }
    }
    *n_send = ns;
    *n_fine_send = nf;
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// Grid adaptation, light data only (single rank).  Stand-ins for refinement_execute_tree's id bookkeeping +
// balanceLoad_tree (LIB/MESH/refinementExecute.f90, balanceLoad_tree.f90) and for respectJmaxJmin_tree / completeness /
// ensureGradedness_tree (LIB/MESH/ensureGradedness_tree.f90:13) so that the drivers of this repository run at 10^5 blocks
// without a Python loop.  In a WABBIT build this logic stays in host Fortran.
// ---------------------------------------------------------------------------------------------------------------------
static whost_forest *clone_with_blocks(const whost_forest *f, int32_t N, std::vector<Blk> &&blocks, int n_ranks = 1)
{
    whost_forest *g = new whost_forest();
    g->dim = f->dim;
    g->Jmax = f->Jmax;
    g->sfc = f->sfc;
    g->n_ranks = n_ranks;
    g->N = N;
    for (int a = 0; a < 3; ++a) g->periodic[a] = f->periodic[a];
    g->blocks = std::move(blocks);
    build(g);
    return g;
}

static int32_t refine_impl(const whost_forest *f, const int32_t *flags, int32_t max_blocks, whost_forest **out, int32_t *n_mothers, int32_t *mothers,
                           int32_t *daughters, int32_t *n_keep, int32_t *keep_src, int32_t *keep_dst, bool global);
static int32_t coarsen_impl(const whost_forest *f, int32_t *status, int32_t Jmin, int32_t max_blocks, whost_forest **out, int32_t *n_mothers,
                            int32_t *mothers, int32_t *daughters, int32_t *n_keep, int32_t *keep_src, int32_t *keep_dst, bool global);

int32_t whost_refine(const whost_forest *f, const int32_t *flags, int32_t max_blocks, whost_forest **out, int32_t *n_mothers, int32_t *mothers,
                     int32_t *daughters, int32_t *n_keep, int32_t *keep_src, int32_t *keep_dst)
{
    return refine_impl(f, flags, max_blocks, out, n_mothers, mothers, daughters, n_keep, keep_src, keep_dst, false);
}

int32_t whost_refine_global(const whost_forest *f, const int32_t *flags, int32_t max_blocks_per_rank, whost_forest **out, int32_t *n_mothers,
                            int32_t *mothers, int32_t *daughters, int32_t *n_keep, int32_t *keep_src, int32_t *keep_dst)
{
    return refine_impl(f, flags, max_blocks_per_rank, out, n_mothers, mothers, daughters, n_keep, keep_src, keep_dst, true);
}

// global = true: any number of ranks; flags and all ids are 1-based positions in the global space-filling-curve order (rank-major order of
// the active lists) of the old resp. new grid, and the new grid is partitioned over the same number of ranks
static int32_t refine_impl(const whost_forest *f, const int32_t *flags, int32_t max_blocks, whost_forest **out, int32_t *n_mothers, int32_t *mothers,
                           int32_t *daughters, int32_t *n_keep, int32_t *keep_src, int32_t *keep_dst, bool global)
{
    if (!f || !out || (!global && f->n_ranks != 1) || !n_mothers || !n_keep) return 1;
    const int dim = f->dim, nd = 1 << dim, n = (int)f->blocks.size();
    std::vector<Blk> nb;
    nb.reserve((size_t)n * nd);
    std::vector<char> ref(n, 0);
    for (int k = 0; k < n; ++k) {   // blocks are in SFC order == hvy order on a single rank
        const Blk &b = f->blocks[k];
        ref[k] = (!flags || flags[k] > 0) && b.level < f->Jmax;   // respectJmaxJmin_tree
        if (!ref[k]) {
            nb.push_back(b);
            continue;
        }
        for (int d = 0; d < nd; ++d) {
            Blk c = b;
            c.level = b.level + 1;
            c.ix[0] = 2 * b.ix[0] + ((d >> 1) & 1);   // digit: bit0 -> y, bit1 -> x, bit2 -> z
            c.ix[1] = 2 * b.ix[1] + (d & 1);
            c.ix[2] = dim == 3 ? 2 * b.ix[2] + ((d >> 2) & 1) : 0;
            nb.push_back(c);
        }
    }
    if ((int64_t)nb.size() > (int64_t)max_blocks * (global ? f->n_ranks : 1)) return 2;   // error_OOM of refine_tree
    whost_forest *g = clone_with_blocks(f, max_blocks, std::move(nb), global ? f->n_ranks : 1);
    for (int r = 0; r < g->n_ranks; ++r)
        if ((int)g->rank_blocks[r].size() > max_blocks) {
            delete g;
            return 2;
        }
    int nm = 0, nk = 0;
    for (int k = 0; k < n; ++k) {
        const Blk &b = f->blocks[k];
        if (!ref[k]) {
            if (keep_src) keep_src[nk] = global ? k + 1 : b.hvy;
            if (keep_dst) keep_dst[nk] = global ? g->find(b.level, b.ix) + 1 : g->blocks[g->find(b.level, b.ix)].hvy;
            ++nk;
            continue;
        }
        if (mothers) mothers[nm] = global ? k + 1 : b.hvy;
        for (int d = 0; d < nd; ++d) {
            const int ix[3] = {2 * b.ix[0] + ((d >> 1) & 1), 2 * b.ix[1] + (d & 1), dim == 3 ? 2 * b.ix[2] + ((d >> 2) & 1) : 0};
            if (daughters) daughters[(size_t)nm * nd + d] = global ? g->find(b.level + 1, ix) + 1 : g->blocks[g->find(b.level + 1, ix)].hvy;
        }
        ++nm;
    }
    *n_mothers = nm;
    *n_keep = nk;
    *out = g;
    return 0;
}

int32_t whost_coarsen(const whost_forest *f, int32_t *status, int32_t Jmin, int32_t max_blocks, whost_forest **out, int32_t *n_mothers,
                      int32_t *mothers, int32_t *daughters, int32_t *n_keep, int32_t *keep_src, int32_t *keep_dst)
{
    return coarsen_impl(f, status, Jmin, max_blocks, out, n_mothers, mothers, daughters, n_keep, keep_src, keep_dst, false);
}

int32_t whost_coarsen_global(const whost_forest *f, int32_t *status, int32_t Jmin, int32_t max_blocks_per_rank, whost_forest **out,
                             int32_t *n_mothers, int32_t *mothers, int32_t *daughters, int32_t *n_keep, int32_t *keep_src, int32_t *keep_dst)
{
    return coarsen_impl(f, status, Jmin, max_blocks_per_rank, out, n_mothers, mothers, daughters, n_keep, keep_src, keep_dst, true);
}

static int32_t coarsen_impl(const whost_forest *f, int32_t *status, int32_t Jmin, int32_t max_blocks, whost_forest **out, int32_t *n_mothers,
                            int32_t *mothers, int32_t *daughters, int32_t *n_keep, int32_t *keep_src, int32_t *keep_dst, bool global)
{
    if (!f || !out || !status || (!global && f->n_ranks != 1) || !n_mothers || !n_keep) return 1;
    const int dim = f->dim, nd = 1 << dim, n = (int)f->blocks.size();
    ensure_neighbors(f);
    // finer neighbours of block k (global position): row of its owner's table, lgt ids -> global positions
    std::vector<int> roff(f->n_ranks + 1, 0), owner(n), local(n);
    for (int r = 0; r < f->n_ranks; ++r) {
        roff[r + 1] = roff[r] + (int)f->rank_blocks[r].size();
        for (size_t i = 0; i < f->rank_blocks[r].size(); ++i) {
            owner[f->rank_blocks[r][i]] = r;
            local[f->rank_blocks[r][i]] = (int)i;
        }
    }
    auto finer = [&](int k, int slot) -> int {   // global position (0-based) of the neighbour in `slot`, or -1
        const int r = owner[k], ld = (int)f->rank_blocks[r].size();
        const int lgt = f->nbr[r][(size_t)slot * ld + local[k]];
        if (lgt < 1) return -1;
        const int rr = (lgt - 1) / f->N, hv = (lgt - 1) % f->N;
        return roff[rr] + hv;
    };
    for (int k = 0; k < n; ++k) status[k] = (status[k] == -1 && f->blocks[k].level > Jmin) ? -1 : 0;
    auto mkey = [&](const Blk &b) {
        const int m[3] = {b.ix[0] >> 1, b.ix[1] >> 1, b.ix[2] >> 1};
        return pos_hash(b.level - 1, m);
    };
    bool changed = true;
    std::unordered_map<uint64_t, std::vector<int>> groups;
    while (changed) {
        changed = false;
        groups.clear();
        for (int k = 0; k < n; ++k)
            if (status[k] == -1) groups[mkey(f->blocks[k])].push_back(k);
        for (auto &kv : groups) {
            bool ok = (int)kv.second.size() == nd;                     // completeness: all sisters are leaves that want to coarsen
            for (size_t i = 0; ok && i < kv.second.size(); ++i) {
                const int k = kv.second[i];
                for (int slot = 112; slot < 168 && ok; ++slot) {       // gradedness: a finer neighbour must coarsen as well
                    const int j = finer(k, slot);
                    if (j >= 0 && status[j] != -1) ok = false;
                }
            }
            if (!ok) {
                for (int k : kv.second) status[k] = 0;
                changed = true;
            }
        }
    }
    std::vector<Blk> nb;
    nb.reserve(n);
    std::vector<std::pair<uint64_t, Blk>> moth;
    for (int k = 0; k < n; ++k) {
        const Blk &b = f->blocks[k];
        if (status[k] != -1) {
            nb.push_back(b);
            continue;
        }
        const int d = (b.ix[1] & 1) | ((b.ix[0] & 1) << 1) | ((dim == 3 ? b.ix[2] & 1 : 0) << 2);
        if (d == 0) {   // one mother per group
            Blk m = b;
            m.level = b.level - 1;
            for (int a = 0; a < 3; ++a) m.ix[a] = b.ix[a] >> 1;
            nb.push_back(m);
        }
    }
    if ((int64_t)nb.size() > (int64_t)max_blocks * (global ? f->n_ranks : 1)) return 2;
    whost_forest *g = clone_with_blocks(f, max_blocks, std::move(nb), global ? f->n_ranks : 1);
    int nm = 0, nk = 0;
    for (int k = 0; k < n; ++k) {
        const Blk &b = f->blocks[k];
        if (status[k] != -1) {
            if (keep_src) keep_src[nk] = global ? k + 1 : b.hvy;
            if (keep_dst) keep_dst[nk] = global ? g->find(b.level, b.ix) + 1 : g->blocks[g->find(b.level, b.ix)].hvy;
            ++nk;
            continue;
        }
        const int d = (b.ix[1] & 1) | ((b.ix[0] & 1) << 1) | ((dim == 3 ? b.ix[2] & 1 : 0) << 2);
        if (d != 0) continue;
        const int mix[3] = {b.ix[0] >> 1, b.ix[1] >> 1, b.ix[2] >> 1};
        if (mothers) mothers[nm] = global ? g->find(b.level - 1, mix) + 1 : g->blocks[g->find(b.level - 1, mix)].hvy;
        for (int dd = 0; dd < nd; ++dd) {
            const int ix[3] = {2 * mix[0] + ((dd >> 1) & 1), 2 * mix[1] + (dd & 1), dim == 3 ? 2 * mix[2] + ((dd >> 2) & 1) : 0};
            if (daughters) daughters[(size_t)nm * nd + dd] = global ? f->find(b.level, ix) + 1 : f->blocks[f->find(b.level, ix)].hvy;
        }
        ++nm;
    }
    *n_mothers = nm;
    *n_keep = nk;
    *out = g;
    return 0;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------------------
// Full tree (leaves + all ancestors) of adapt_tree's full wavelet transformation: neighbourhood tables and the grid decision.
// Light data only; stand-ins for updateMetadata_tree on the full tree (LIB/MESH/updateMetadata_tree.f90) and for
// respectJmaxJmin_tree + ensureGradedness_tree(check_daughters) (LIB/MESH/ensureGradedness_tree.f90, ensure_completeness_block.f90).
// Blocks are given as (level, ix, iy, iz) in any order; all outputs are indices into that list or -1.
// ---------------------------------------------------------------------------------------------------------------------
extern "C" {

// nb[n][3^dim - 1]: same-level neighbour per direction (dz, dy, dx ascending, (0,0,0) skipped; periodic), par[n]: mother,
// child[n][2^dim]: daughters, column = x offset + 2 * y offset + 4 * z offset
// (level, position) -> index lookup for a list of tree blocks: one dense array per level when the levels are shallow enough (8^Jmax entries:
// 1 MB at level 6, direct addressing, cache resident), else open addressing
struct TreeIndex {
    int dim = 3, lmax = 0;
    bool dense = false;
    std::vector<size_t> base;          // dense: offset of level l
    std::vector<int> cell;
    std::vector<uint64_t> hkeys;
    std::vector<int> hvals;
    uint64_t hmask = 0;

    bool build(int dim_, int n, const int32_t *level, const int32_t *pos)
    {
        dim = dim_;
        lmax = 0;
        for (int i = 0; i < n; ++i) lmax = std::max(lmax, (int)level[i]);
        size_t total = 0;
        base.assign(lmax + 2, 0);
        for (int l = 0; l <= lmax; ++l) {
            base[l] = total;
            total += (size_t)1 << (l * dim);
        }
        dense = total <= ((size_t)1 << 25);
        if (dense) {
            cell.assign(total, -1);
            for (int i = 0; i < n; ++i) {
                int &c = cell[slot(level[i], pos + 3 * i)];
                if (c >= 0) return false;   // duplicate position
                c = i;
            }
            return true;
        }
        size_t cap = 64;
        while (cap < (size_t)n * 2 + 2) cap <<= 1;
        hmask = cap - 1;
        hkeys.assign(cap, ~0ull);
        hvals.assign(cap, -1);
        for (int i = 0; i < n; ++i) {
            const uint64_t key = pos_hash(level[i], pos + 3 * i);
            uint64_t h = whost_forest::mix(key) & hmask;
            while (hkeys[h] != ~0ull && hkeys[h] != key) h = (h + 1) & hmask;
            hkeys[h] = key;
            hvals[h] = i;
        }
        return true;
    }
    size_t slot(int l, const int p[3]) const
    {
        return base[l] + ((((size_t)(dim == 3 ? p[2] : 0) << l) | (size_t)p[1]) << l | (size_t)p[0]);
    }
    int find(int l, const int p[3]) const
    {
        if (l < 0 || l > lmax) return -1;
        if (dense) return cell[slot(l, p)];
        const uint64_t key = pos_hash(l, p);
        uint64_t h = whost_forest::mix(key) & hmask;
        while (hkeys[h] != ~0ull) {
            if (hkeys[h] == key) return hvals[h];
            h = (h + 1) & hmask;
        }
        return -1;
    }
};

int32_t whost_ft_tables(int32_t dim, int32_t n, const int32_t *level, const int32_t *pos, int32_t *nb, int32_t *par, int32_t *child)
{
    if (n < 0 || (n > 0 && (!level || !pos || !nb || !par || !child))) return 1;
    TreeIndex T;
    if (!T.build(dim, n, level, pos)) return 3;
    const int nd = 1 << dim, ndir = (dim == 3 ? 27 : 9) - 1;
    for (int i = 0; i < n; ++i) {   // (serial: 35 direct-address lookups per block, ~10 ms at 5e4 blocks; threads cost more than they gain here)
        const int l = level[i], nb_l = 1 << l;
        const int *p = pos + 3 * i;
        int q = 0;
        for (int dz = (dim == 3 ? -1 : 0); dz <= (dim == 3 ? 1 : 0); ++dz)
            for (int dy = -1; dy <= 1; ++dy)
                for (int dx = -1; dx <= 1; ++dx) {
                    if (!dx && !dy && !dz) continue;
                    const int m_l = nb_l - 1;   // nb_l is a power of two: periodic wrap by masking
                    const int np[3] = {(p[0] + dx) & m_l, (p[1] + dy) & m_l, dim == 3 ? ((p[2] + dz) & m_l) : 0};
                    nb[(size_t)i * ndir + q++] = T.find(l, np);
                }
        const int pp[3] = {p[0] >> 1, p[1] >> 1, p[2] >> 1};
        par[i] = T.find(l - 1, pp);
        for (int c = 0; c < nd; ++c) {
            const int cp[3] = {2 * p[0] + (c & 1), 2 * p[1] + ((c >> 1) & 1), dim == 3 ? 2 * p[2] + ((c >> 2) & 1) : 0};
            child[(size_t)i * nd + c] = T.find(l + 1, cp);
        }
    }
    return 0;
}

// the (significant block, direction) pairs addSecurityZone_CE_tree examines (LIB/MESH/securityZone_tree.f90:140-298): block i is significant
// (sig[i] != 0), its same-level neighbour in direction q exists and is insignificant (status -1).  Pairs come out direction-major (all blocks
// of direction 0 first), blocks ascending: blk[k], dir[k] (index q into the direction list of whost_ft_tables).  Returns the count.
int32_t whost_ft_security_pairs(int32_t dim, int32_t n, const int32_t *nb, const int32_t *status, const uint8_t *sig, int32_t cap, int32_t *blk,
                                int32_t *dir)
{
    if (n < 0 || (n > 0 && (!nb || !status || !sig))) return -1;
    const int ndir = (dim == 3 ? 27 : 9) - 1;
    int64_t k = 0;
    for (int q = 0; q < ndir; ++q)
        for (int i = 0; i < n; ++i) {
            if (!sig[i]) continue;
            const int j = nb[(size_t)i * ndir + q];
            if (j < 0 || status[j] != -1) continue;
            if (k < cap) {
                blk[k] = i;
                dir[k] = q;
            }
            ++k;
        }
    return k > cap ? -2 : (int32_t)k;
}

// init_full_tree (LIB/MESH/adapt_tree.f90:268-330, the light-data half): all ancestors of the leaves down to Jmin are added; the tree comes back
// sorted by position code (level << 57 | z << 38 | y << 19 | x), leaf_of[i] = index of the block in the input list or -1 for a mother.
int32_t whost_ft_build(int32_t dim, int32_t Jmin, int32_t n_leaf, const int32_t *level, const int32_t *pos, int32_t cap, int32_t *n_out,
                       int32_t *level_out, int32_t *pos_out, int32_t *leaf_of)
{
    if (n_leaf < 0 || !n_out || (n_leaf > 0 && (!level || !pos)) || !level_out || !pos_out || !leaf_of) return 1;
    (void)dim;
    auto code_of = [](int l, const int *p) {
        return ((uint64_t)l << 57) | ((uint64_t)(uint32_t)p[2] << 38) | ((uint64_t)(uint32_t)p[1] << 19) | (uint64_t)(uint32_t)p[0];
    };
    std::vector<std::pair<uint64_t, int>> all;
    all.reserve((size_t)n_leaf + n_leaf / 6 + 16);
    for (int i = 0; i < n_leaf; ++i) all.emplace_back(code_of(level[i], pos + 3 * i), i);
    // ancestors, one level at a time: parents of the newest layer that are not in the tree yet
    std::vector<uint64_t> layer(all.size());
    for (size_t i = 0; i < all.size(); ++i) layer[i] = all[i].first;
    std::vector<uint64_t> seen(layer);
    std::sort(seen.begin(), seen.end());
    while (!layer.empty()) {
        std::vector<uint64_t> par;
        par.reserve(layer.size() / 4 + 8);
        for (uint64_t c : layer) {
            const int l = (int)(c >> 57);
            if (l <= Jmin) continue;
            const int p[3] = {(int)(c & 0x7FFFF) >> 1, (int)((c >> 19) & 0x7FFFF) >> 1, (int)((c >> 38) & 0x7FFFF) >> 1};
            par.push_back(code_of(l - 1, p));
        }
        std::sort(par.begin(), par.end());
        par.erase(std::unique(par.begin(), par.end()), par.end());
        std::vector<uint64_t> fresh;
        fresh.reserve(par.size());
        for (uint64_t c : par)
            if (!std::binary_search(seen.begin(), seen.end(), c)) fresh.push_back(c);
        for (uint64_t c : fresh) all.emplace_back(c, -1);
        std::vector<uint64_t> merged(seen.size() + fresh.size());
        std::merge(seen.begin(), seen.end(), fresh.begin(), fresh.end(), merged.begin());
        seen.swap(merged);
        layer.swap(fresh);
    }
    if ((int64_t)all.size() > cap) return 2;
    std::sort(all.begin(), all.end());
    for (size_t i = 0; i < all.size(); ++i) {
        const uint64_t c = all[i].first;
        level_out[i] = (int)(c >> 57);
        pos_out[3 * i] = (int)(c & 0x7FFFF);
        pos_out[3 * i + 1] = (int)((c >> 19) & 0x7FFFF);
        pos_out[3 * i + 2] = (int)((c >> 38) & 0x7FFFF);
        leaf_of[i] = all[i].second;
    }
    *n_out = (int32_t)all.size();
    return 0;
}

// numerical treecodes of many blocks at once (encoding_b, LIB/TREE/module_treelib.f90:837-871)
int32_t whost_encode_many(int32_t dim, int32_t Jmax, int32_t n, const int32_t *level, const int32_t *pos, int64_t *tc)
{
    if (n < 0 || (n > 0 && (!level || !pos || !tc))) return 1;
    for (int i = 0; i < n; ++i) tc[i] = whost_encode(dim, level[i], Jmax, pos + 3 * i);
    return 0;
}

// hvy_neighbor rows of every block of a full tree (wabbit_b200/fulltree.py: _upload_rows): same-level relations to whatever block sits
// there (row dir_code[q] - 1); for the blocks flagged in `leaf`, directions without a same-level block become a coarser relation
// (row dir_code[q] - 1 + 56) to the block one level up that covers the neighbour position.  rows: [168][ld] int32, column = slot - 1;
// entries that are not set keep their value (the caller presets -1).  Direction-major, so the writes stream.
int32_t whost_ft_rows(int32_t dim, int32_t n, const int32_t *level, const int32_t *pos, const int32_t *nb, const int32_t *slots,
                      const int32_t *leaf, const int32_t *dir_code, int64_t ld, int32_t *rows)
{
    if (n < 0 || ld < 0 || (n > 0 && (!level || !pos || !nb || !slots || !leaf || !dir_code || !rows))) return 1;
    size_t cap = 64;
    while (cap < (size_t)n * 2 + 2) cap <<= 1;
    const uint64_t hmask = cap - 1;
    std::vector<uint64_t> hkeys(cap, ~0ull);
    std::vector<int> hvals(cap, -1);
    for (int i = 0; i < n; ++i) {
        if (slots[i] < 1 || slots[i] > ld) return 2;
        const uint64_t key = pos_hash(level[i], pos + 3 * i);
        uint64_t h = whost_forest::mix(key) & hmask;
        while (hkeys[h] != ~0ull && hkeys[h] != key) h = (h + 1) & hmask;
        hkeys[h] = key;
        hvals[h] = i;
    }
    auto find = [&](int l, const int p[3]) -> int {
        if (l < 0) return -1;
        const uint64_t key = pos_hash(l, p);
        uint64_t h = whost_forest::mix(key) & hmask;
        while (hkeys[h] != ~0ull) {
            if (hkeys[h] == key) return hvals[h];
            h = (h + 1) & hmask;
        }
        return -1;
    };
    const int ndir = (dim == 3 ? 27 : 9) - 1;
    int q = 0;
    for (int dz = (dim == 3 ? -1 : 0); dz <= (dim == 3 ? 1 : 0); ++dz)
        for (int dy = -1; dy <= 1; ++dy)
            for (int dx = -1; dx <= 1; ++dx) {
                if (!dx && !dy && !dz) continue;
                const int d[3] = {dx, dy, dz};
                int32_t *same = rows + (size_t)(dir_code[q] - 1) * ld, *coarse = rows + (size_t)(dir_code[q] - 1 + 56) * ld;
#pragma omp parallel for schedule(static)
                for (int i = 0; i < n; ++i) {
                    const int j = nb[(size_t)i * ndir + q];
                    if (j >= 0) {
                        same[slots[i] - 1] = slots[j];
                    } else if (leaf[i]) {
                        const int l = level[i], nb_l = 1 << l;
                        const int *p = pos + 3 * i;
                        const int m_l = nb_l - 1;
                        const int cp[3] = {((p[0] + d[0]) & m_l) >> 1, ((p[1] + d[1]) & m_l) >> 1, dim == 3 ? (((p[2] + d[2]) & m_l) >> 1) : 0};
                        const int c = find(l - 1, cp);
                        if (c >= 0) coarse[slots[i] - 1] = slots[c];
                    }
                }
                ++q;
            }
    return 0;
}

// status[n] in: -1 = insignificant; out: -1 only for the blocks that are deleted, 9 (REF_UNSIGNIFICANT_STAY) for demoted ones.
// A block keeps -1 only if it sits above Jmin, all its 2^dim sisters carry -1, none of its daughters stays and none of its finer
// neighbours (the daughters of its same-level neighbours that touch it) stays.  Statuses only move from -1 to 9: unique fixed point.
int32_t whost_ft_decide(int32_t dim, int32_t n, const int32_t *level, const int32_t *nb, const int32_t *par, const int32_t *child, int32_t Jmin,
                        int32_t *status)
{
    if (n < 0 || (n > 0 && (!level || !nb || !par || !child || !status))) return 1;
    const int nd = 1 << dim, ndir = (dim == 3 ? 27 : 9) - 1;
    // daughters of the neighbour in direction q that touch the block: offset 1 on axes with d < 0, 0 with d > 0, both with d = 0
    std::vector<std::vector<int>> cols(ndir);
    {
        int q = 0;
        for (int dz = (dim == 3 ? -1 : 0); dz <= (dim == 3 ? 1 : 0); ++dz)
            for (int dy = -1; dy <= 1; ++dy)
                for (int dx = -1; dx <= 1; ++dx) {
                    if (!dx && !dy && !dz) continue;
                    const int d[3] = {dx, dy, dz};
                    for (int c = 0; c < nd; ++c) {
                        bool ok = true;
                        for (int a = 0; a < dim; ++a) {
                            const int off = (c >> a) & 1;
                            if ((d[a] < 0 && off != 1) || (d[a] > 0 && off != 0)) ok = false;
                        }
                        if (ok) cols[q].push_back(c);
                    }
                    ++q;
                }
    }
    for (int i = 0; i < n; ++i)
        if (status[i] == -1 && level[i] <= Jmin) status[i] = 9;
    // Every rule looks at the block's own level (sisters) or one level finer (daughters, finer neighbours), and statuses only move from -1
    // to 9: one sweep from the finest level to the coarsest -- first the rules that read the finer level, then completeness -- reaches the
    // fixed point the reference iterates to.
    int lmax = 0;
    for (int i = 0; i < n; ++i) lmax = std::max(lmax, (int)level[i]);
    std::vector<std::vector<int>> by_level(lmax + 1);
    for (int i = 0; i < n; ++i) by_level[level[i]].push_back(i);
    for (int l = lmax; l >= 0; --l) {
        const std::vector<int> &B = by_level[l];
        for (int i : B) {
            if (status[i] != -1) continue;
            bool stay = par[i] < 0;
            for (int c = 0; c < nd && !stay; ++c) {
                const int s = child[(size_t)i * nd + c];
                if (s >= 0 && status[s] != -1) stay = true;                                      // check_daughters
            }
            for (int q = 0; q < ndir && !stay; ++q) {
                const int j = nb[(size_t)i * ndir + q];
                if (j < 0) continue;
                for (int c : cols[q]) {
                    const int f = child[(size_t)j * nd + c];
                    if (f >= 0 && status[f] != -1) { stay = true; break; }                       // gradedness
                }
            }
            if (stay) status[i] = 9;
        }
        for (int i : B) {                                                                         // completeness: decided per sister group
            if (status[i] != -1) continue;
            bool stay = false;
            for (int c = 0; c < nd && !stay; ++c) {
                const int s = child[(size_t)par[i] * nd + c];
                if (s < 0 || (status[s] != -1 && status[s] != -3)) stay = true;
            }
            // mark the verdict without disturbing the sisters' view of this round: -3 = "-1, but the group stays"
            if (stay) status[i] = -3;
        }
        for (int i : B)
            if (status[i] == -3) status[i] = 9;
    }
    return 0;
}

}  // extern "C"
