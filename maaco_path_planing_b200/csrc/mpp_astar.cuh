// mpp_astar.cuh -- warp-cooperative exact emulation of the reference's two A* connectors.
//
//   variant 0: AStarSolver.solve  (astar.py:33-101)  -- closed set seeded from nodes_to_avoid minus
//              {start,target}, relax from the popped g, in-place decrease-key.
//   variant 1: MPA._a_star        (MPA.py:106-151)   -- no closed set, avoid-set filters neighbours only,
//              relax from g_score[current], keys of nodes already in the open set are left stale,
//              nodes not in the open set (incl. already expanded ones) are re-pushed.
//
// Both are fully described by "extract-min over the total order (f, g, r, c)" + their relax rules
// (validated 0/1500 mismatches against the unmodified reference); heap internals do
// not matter.  One warp runs one search: a 32-ary min-heap in HBM (one coalesced 32-child load per level,
// warp arg-min by REDUX), lanes 0..7 relax the 8 neighbours in parallel.  Decrease-key is done by lazy
// deletion (variant 0: the fresher entry always sorts first, the stale one is dropped when popped because
// its cell is closed), which keeps the pop sequence identical to the reference's.
#pragma once
#include "mpp_common.cuh"

#define MPP_INF_BITS 0x7ff0000000000000ll
#define MPP_ASTAR_SQRT2 1.4142135623730951

// neighbour order helper.py:30-36 == MPA.py:71-77: (0,1),(0,-1),(1,0),(-1,0),(1,1),(1,-1),(-1,1),(-1,-1)
// (delta+1) packed two bits per move
#define MPP_NB_R 0x0A25u  // r+1: 1,1,2,0,2,2,0,0
#define MPP_NB_C 0x2252u  // c+1: 2,0,1,1,2,0,2,0

struct AStarGrid {
    const uint32_t *occ;  // padded occupancy bits (shared or global)
    int pitch, R, C;
    int allow_diag, restrict_corner;
};

// per-warp scratch slot (HBM).  meta word: [31:8] stamp, bit4 closed, bit3 in_open, [2:0] parent move.
struct AStarSlot {
    double *g;
    uint32_t *meta;
    double *hf, *hg;
    int32_t *hc;
    uint32_t *hdr;  // hdr[0] = stamp counter
    int heap_cap;
};

__host__ __device__ __forceinline__ size_t astar_slot_bytes(int rc, int heap_cap) {
    size_t b = 256;                                  // header
    b += ((size_t)rc * 8 + 255) & ~(size_t)255;      // g
    b += ((size_t)rc * 4 + 255) & ~(size_t)255;      // meta
    size_t hs = (size_t)heap_cap + 64;               // storage index = node + 31, padded
    b += 2 * ((hs * 8 + 255) & ~(size_t)255);        // hf, hg
    b += (hs * 4 + 255) & ~(size_t)255;              // hc
    return b;
}
__device__ __forceinline__ AStarSlot astar_slot_at(char *base, int rc, int heap_cap) {
    AStarSlot s;
    s.hdr = (uint32_t *)base; base += 256;
    s.g = (double *)base; base += ((size_t)rc * 8 + 255) & ~(size_t)255;
    s.meta = (uint32_t *)base; base += ((size_t)rc * 4 + 255) & ~(size_t)255;
    size_t hs = (size_t)heap_cap + 64;
    s.hf = (double *)base; base += (hs * 8 + 255) & ~(size_t)255;
    s.hg = (double *)base; base += (hs * 8 + 255) & ~(size_t)255;
    s.hc = (int32_t *)base;
    s.heap_cap = heap_cap;
    return s;
}

__device__ __forceinline__ bool occ_bit(const AStarGrid &G, int r, int c) {  // true = blocked (incl. out of bounds)
    const int pb = c + 1;
    return (G.occ[(r + 1) * G.pitch + (pb >> 5)] >> (pb & 31)) & 1u;
}
__device__ __forceinline__ double hdist_dev(int r0, int c0, int r1, int c1) {  // helper.py:8-12
    const long long dr = r0 - r1, dc = c0 - c1;
    return sqrt((double)(dr * dr + dc * dc));                                 // IEEE correctly rounded
}

// lexicographic (f, g, cell) "a < b"
__device__ __forceinline__ bool key_less(double fa, double ga, int ca, double fb, double gb, int cb) {
    if (fa != fb) return fa < fb;
    if (ga != gb) return ga < gb;
    return ca < cb;
}

// warp arg-min over lanes of (f, g, cell); lanes with f = +inf are empty.  All lanes get the winner lane.
__device__ __forceinline__ int warp_argmin_key(double f, double g, int cell) {
    // non-negative doubles: unsigned bit order == numeric order (+inf sorts last)
    const unsigned long long kf = (unsigned long long)__double_as_longlong(f);
    uint32_t hi = (uint32_t)(kf >> 32), lo = (uint32_t)kf;
    uint32_t mh = __reduce_min_sync(0xffffffffu, hi);
    uint32_t ml = __reduce_min_sync(0xffffffffu, hi == mh ? lo : 0xffffffffu);
    uint32_t mask = __ballot_sync(0xffffffffu, hi == mh && lo == ml);
    if (__popc(mask) > 1) {
        const unsigned long long kg = (unsigned long long)__double_as_longlong(g);
        const bool in = (mask >> (threadIdx.x & 31)) & 1u;
        hi = in ? (uint32_t)(kg >> 32) : 0xffffffffu;
        lo = (uint32_t)kg;
        mh = __reduce_min_sync(0xffffffffu, hi);
        ml = __reduce_min_sync(0xffffffffu, (in && hi == mh) ? lo : 0xffffffffu);
        mask = __ballot_sync(0xffffffffu, in && hi == mh && lo == ml);
        if (__popc(mask) > 1) {
            const bool in2 = (mask >> (threadIdx.x & 31)) & 1u;
            const uint32_t mc = __reduce_min_sync(0xffffffffu, in2 ? (uint32_t)cell : 0xffffffffu);
            mask = __ballot_sync(0xffffffffu, in2 && (uint32_t)cell == mc);
        }
    }
    return __ffs(mask) - 1;
}

#define HIDX(k) ((k) + 31)  // node k -> storage index; children of k = nodes 32k+1..32k+32 (storage aligned to 32)

// Push (f,g,cell) -- warp-uniform arguments.  Returns false on overflow.
__device__ __forceinline__ bool heap_push(AStarSlot &S, int &n, double f, double g, int cell) {
    if (n >= S.heap_cap) return false;
    int k = n++;
    const int lane = threadIdx.x & 31;
    while (k > 0) {
        const int p = (k - 1) >> 5;
        const double pf = S.hf[HIDX(p)], pg = S.hg[HIDX(p)];
        const int pc = S.hc[HIDX(p)];
        if (!key_less(f, g, cell, pf, pg, pc)) break;
        if (lane == 0) { S.hf[HIDX(k)] = pf; S.hg[HIDX(k)] = pg; S.hc[HIDX(k)] = pc; }
        k = p;
    }
    if (lane == 0) { S.hf[HIDX(k)] = f; S.hg[HIDX(k)] = g; S.hc[HIDX(k)] = cell; }
    __syncwarp();
    return true;
}

// Pop the minimum -- returns it in (f,g,cell) (warp-uniform).  n > 0 required.
__device__ __forceinline__ void heap_pop(AStarSlot &S, int &n, double &f, double &g, int &cell) {
    const int lane = threadIdx.x & 31;
    f = S.hf[HIDX(0)]; g = S.hg[HIDX(0)]; cell = S.hc[HIDX(0)];
    --n;
    if (n == 0) return;
    const double xf = S.hf[HIDX(n)], xg = S.hg[HIDX(n)];
    const int xc = S.hc[HIDX(n)];
    __syncwarp();
    int k = 0;
    for (;;) {
        const int c0 = 32 * k + 1;
        if (c0 >= n) break;
        const int ci = c0 + lane;
        double cf = __longlong_as_double(MPP_INF_BITS), cg = 0.0;
        int cc = 0x7fffffff;
        if (ci < n) { cf = S.hf[HIDX(ci)]; cg = S.hg[HIDX(ci)]; cc = S.hc[HIDX(ci)]; }
        const int w = warp_argmin_key(cf, cg, cc);
        const double mf = __shfl_sync(0xffffffffu, cf, w), mg = __shfl_sync(0xffffffffu, cg, w);
        const int mc = __shfl_sync(0xffffffffu, cc, w);
        if (!key_less(mf, mg, mc, xf, xg, xc)) break;
        if (lane == 0) { S.hf[HIDX(k)] = mf; S.hg[HIDX(k)] = mg; S.hc[HIDX(k)] = mc; }
        k = c0 + w;
    }
    if (lane == 0) { S.hf[HIDX(k)] = xf; S.hg[HIDX(k)] = xg; S.hc[HIDX(k)] = xc; }
    __syncwarp();
}

// One search by one warp.  Returns number of path cells written to out[0..) in forward order
// (0 = no path / invalid endpoints, -1 = heap overflow).  avoid: bitmap over cells or nullptr.
// *g_out = g of the popped target entry (+inf when no path).
static __device__ int astar_search(const AStarGrid &G, AStarSlot &S, int variant, int src, int dst,
                            const uint32_t *avoid, int32_t *out, int out_cap, double *g_out,
                            unsigned long long *n_exp, unsigned long long *n_rel) {
    const int lane = threadIdx.x & 31;
    const int C = G.C;
    const int sr = src / C, sc = src % C, tr = dst / C, tc = dst % C;
    const double INF = __longlong_as_double(MPP_INF_BITS);
    if (g_out) *g_out = INF;
    if (variant == 1 && src == dst) { if (lane == 0 && out_cap > 0) out[0] = src; if (g_out) *g_out = 0.0; __syncwarp(); return 1; }  // MPA.py:107-108
    if (occ_bit(G, sr, sc) || occ_bit(G, tr, tc)) return 0;        // astar.py:37-39 / MPA.py:109-111
    if (variant == 0 && src == dst) { if (lane == 0 && out_cap > 0) out[0] = src; if (g_out) *g_out = 0.0; __syncwarp(); return 1; }  // astar.py:41-42
    // new stamp for this search (meta words of older searches become invalid without clearing)
    uint32_t stamp = S.hdr[0] + 1;
    if (stamp >= (1u << 24)) {  // wrap: clear the meta array once every 16M searches
        for (int i = lane; i < G.R * C; i += 32) S.meta[i] = 0;
        stamp = 1;
    }
    __syncwarp();
    if (lane == 0) S.hdr[0] = stamp;
    const uint32_t stamp_hi = stamp << 8;
    int n = 0;
    if (lane == 0) { S.g[src] = 0.0; S.meta[src] = stamp_hi | 8u; }
    __syncwarp();
    heap_push(S, n, hdist_dev(sr, sc, tr, tc), 0.0, src);          // astar.py:45 / MPA.py:113
    const long long max_steps = (long long)G.R * C * (variant == 0 ? 3 : 2);   // astar.py:58 / MPA.py:118
    long long steps = 0;
    unsigned long long exps = 0, rels = 0;
    // per-lane neighbour deltas
    const bool nb_lane = lane < (G.allow_diag ? 8 : 4);
    const int dr = (int)((MPP_NB_R >> (2 * (lane & 7))) & 3u) - 1, dc = (int)((MPP_NB_C >> (2 * (lane & 7))) & 3u) - 1;
    const double step_cost = (lane & 7) >= 4 ? MPP_ASTAR_SQRT2 : 1.0;  // distance_euclidean: sqrt(1)=1, sqrt(2)
    int found = 0;
    double cur_g = 0.0;
    while (n > 0 && steps < max_steps) {
        double cf;
        int cur;
        heap_pop(S, n, cf, cur_g, cur);
        uint32_t mcur = S.meta[cur];
        if (variant == 0) {
            if (cur == dst) { found = 1; ++steps; break; }                     // astar.py:64
            if (mcur & 16u) continue;                                         // stale (lazily deleted) entry
            ++steps;
            if (lane == 0) S.meta[cur] = mcur | 16u;                           // closed_set.add astar.py:74
        } else {
            ++steps;
            if (cur == dst) { found = 1; break; }                              // MPA.py:123
            if (lane == 0) S.meta[cur] = mcur & ~8u;                           // left the open set
        }
        ++exps;
        const int cr = cur / C, cc = cur % C;
        const double gbase = (variant == 0) ? cur_g : S.g[cur];                // astar.py:85 vs MPA.py:135
        __syncwarp();
        // ---- relax the neighbours, one per lane ----
        bool push = false;
        int j = 0;
        double tg = 0.0;
        if (nb_lane) {
            const int nr = cr + dr, nc = cc + dc;
            bool blocked = occ_bit(G, nr, nc);
            if (!blocked && (lane >= 4) && G.restrict_corner)                  // helper.py:45-49 / MPA.py:86-96
                blocked = occ_bit(G, cr + dr, cc) || occ_bit(G, cr, cc + dc);
            if (!blocked) {
                j = nr * C + nc;
                const bool av = avoid ? ((avoid[j >> 5] >> (j & 31)) & 1u) : false;
                uint32_t mj = S.meta[j];
                const bool seen = (mj & 0xffffff00u) == stamp_hi;
                if (!seen) mj = 0;
                bool excluded;
                if (variant == 0) excluded = (av && j != src && j != dst) || (mj & 16u);   // closed set astar.py:51-56,80
                else excluded = av;                                                       // MPA.py:132
                if (!excluded) {
                    tg = gbase + step_cost;
                    const double gj = seen ? S.g[j] : INF;
                    if (tg < gj) {                                             // astar.py:87 / MPA.py:137
                        S.g[j] = tg;
                        // parent = the move taken cur -> j; variant 1 keeps a stale key if already in open
                        push = (variant == 0) ? true : !(mj & 8u);
                        S.meta[j] = stamp_hi | (mj & 16u) | 8u | (uint32_t)(lane & 7);
                        ++rels;
                    }
                }
            }
        }
        uint32_t pm = __ballot_sync(0xffffffffu, push);
        while (pm) {
            const int l = __ffs(pm) - 1;
            pm &= pm - 1;
            const int pj = __shfl_sync(0xffffffffu, j, l);
            const double ptg = __shfl_sync(0xffffffffu, tg, l);
            const double pf = ptg + hdist_dev(pj / C, pj % C, tr, tc);         // astar.py:90 / MPA.py:140
            if (!heap_push(S, n, pf, ptg, pj)) return -1;
        }
    }
    rels = __reduce_add_sync(0xffffffffu, (uint32_t)rels);
    if (lane == 0) {
        if (n_exp) atomicAdd(n_exp, exps);
        if (n_rel) atomicAdd(n_rel, rels);
    }
    if (!found) return 0;
    if (g_out) *g_out = cur_g;
    // ---- rebuild: follow parent moves target -> start (astar.py:65-69 / MPA.py:124-130) ----
    int len = 0;
    if (lane == 0) {
        int t = dst;
        while (true) {
            if (len < out_cap) out[len] = t;
            ++len;
            if (t == src) break;
            const uint32_t mv = S.meta[t] & 7u;
            const int pdr = (int)((MPP_NB_R >> (2 * mv)) & 3u) - 1, pdc = (int)((MPP_NB_C >> (2 * mv)) & 3u) - 1;
            t -= pdr * C + pdc;
        }
    }
    len = __shfl_sync(0xffffffffu, len, 0);
    __syncwarp();
    const int m = len < out_cap ? len : out_cap;  // (if truncated the caller sees len > cap)
    for (int i = lane; i < m / 2; i += 32) {       // reverse in place
        const int32_t a = out[i], b = out[m - 1 - i];
        out[i] = b; out[m - 1 - i] = a;
    }
    __syncwarp();
    return len;
}
