// mpp_maaco.cu -- MAACO colony pass on B200: per-pass move ranking, tour construction (one thread per
// ant by default; warp / 16- / 8-lane cooperative forms kept), order-dependent best tracking, and the
// atomics-free ordered pheromone update.
// Reference semantics: MAACO.py:58-91 (tables), :100-181 (filter), :197-262 (selection),
// :278-302 (tour), :304-332 (pheromone), :343-358 (best tracking).
#include <cmath>
#include <cstdlib>
#include <thread>
#include <vector>

#include "mpp_common.cuh"

// ---------------------------------------------------------------------------------------------
// host: tables (libm exp/pow == what CPython / NumPy call, so the tables are bit-identical)
// ---------------------------------------------------------------------------------------------
static inline double hdist(int r0, int c0, int r1, int c1) {
    long long dr = r0 - r1, dc = c0 - c1;
    return std::sqrt((double)(dr * dr + dc * dc));
}

extern "C" double mpp_maaco_q0(int K, int k, double q0_initial) {  // MAACO.py:212-226
    double k0 = 0.7 * (double)K, q0;
    if ((double)k < k0) {
        if (std::fabs((double)K - k0) < 1e-6) q0 = q0_initial;
        else q0 = ((double)(K - k) / (double)K) * q0_initial;
    } else {
        double q0_at_k0 = (((double)K - k0) / (double)K) * q0_initial;
        q0 = q0_at_k0 +
             (((double)k - k0) / ((double)K - k0 + 1e-9)) * (q0_initial * (1 - ((double)K - k0) / (double)K) / 2.0);
    }
    q0 = q0 > 0.01 ? q0 : 0.01;
    return q0 < 0.99 ? q0 : 0.99;
}

extern "C" int mpp_maaco_tables(const mpp_map *map, const mpp_maaco_params *p, double *tau0_dev, double *E01_dev,
                                double *dist_t_dev, void *stream) {
    MPP_REQUIRE(map && p && tau0_dev && E01_dev, "mpp_maaco_tables: null argument");
    MPP_REQUIRE(map->start >= 0 && map->target >= 0, "mpp_maaco_tables: map has no start/target");
    const int R = map->rows, C = map->cols;
    const size_t n = (size_t)R * C;
    const int sr = map->start / C, sc = map->start % C, tr = map->target / C, tc = map->target % C;
    double dsT = hdist(sr, sc, tr, tc);
    if (dsT < 1e-9) dsT = 1e-9;  // MAACO.py:43-45
    std::vector<double> buf(4 * n);
    double *tau0 = buf.data(), *E01 = tau0 + n, *dt = E01 + 2 * n;
    const uint8_t *grid = map->grid_host;
    auto work = [&](int r_lo, int r_hi) {
        for (int r = r_lo; r < r_hi; ++r)
            for (int c = 0; c < C; ++c) {
                const size_t i = (size_t)r * C + c;
                const double diT = hdist(r, c, tr, tc);
                const double dsi = hdist(sr, sc, r, c);
                dt[i] = diT;
                if (grid[i] == 1) {
                    tau0[i] = 1e-9;
                } else {
                    const double den = dsi + diT;
                    double factor;
                    if (den < 1e-9) factor = (dsi < 1e-6 || diT < 1e-6) ? 1.0 : 0.1;
                    else factor = dsT / den;
                    const double v = factor * p->C0_initial_pheromone;
                    tau0[i] = v < 1e-9 ? 1e-9 : v;
                }
                double h;
                if (dsT < 1e-9) h = p->wh_min;
                else h = p->wh_max - (p->wh_max - p->wh_min) * std::exp(-p->k_h_adaptive * diT / dsT);
                const double g = 1.0 - h;
                const double base = g * dsi + h * diT;
                double d0 = base + p->a_turn_coef * 0.0, d1 = base + p->a_turn_coef * 1.0;
                d0 = d0 > 1e-9 ? d0 : 1e-9;
                d1 = d1 > 1e-9 ? d1 : 1e-9;
                E01[2 * i] = std::pow(1.0 / d0, p->beta);        // turn flag 0
                E01[2 * i + 1] = std::pow(1.0 / d1, p->beta);    // turn flag 1
            }
    };
    unsigned nt = std::thread::hardware_concurrency();
    if (nt > 16) nt = 16;
    if (nt < 2 || n < 65536) {
        work(0, R);
    } else {
        std::vector<std::thread> th;
        int per = (R + (int)nt - 1) / (int)nt;
        for (unsigned t = 0; t < nt; ++t) {
            int lo = (int)t * per, hi = lo + per > R ? R : lo + per;
            if (lo < hi) th.emplace_back(work, lo, hi);
        }
        for (auto &t : th) t.join();
    }
    cudaStream_t s = (cudaStream_t)stream;
    MPP_CUDA(cudaSetDevice(map->device));
    MPP_CUDA(cudaMemcpyAsync(tau0_dev, tau0, n * 8, cudaMemcpyHostToDevice, s));
    MPP_CUDA(cudaMemcpyAsync(E01_dev, E01, 2 * n * 8, cudaMemcpyHostToDevice, s));
    if (dist_t_dev) MPP_CUDA(cudaMemcpyAsync(dist_t_dev, dt, n * 8, cudaMemcpyHostToDevice, s));
    MPP_CUDA(cudaStreamSynchronize(s));  // buf is freed on return
    return MPP_OK;
}

// ---------------------------------------------------------------------------------------------
// K2: tour construction
// ---------------------------------------------------------------------------------------------
// Buffer layout (uint32 words): [margin | 9*R*C strategy-1 words | margin | pad to even] [9*R*C full entries (2 words)].
// The margins let the tour kernel read the word of a neighbour cell without bounds checks.
struct RankLayout { size_t margin, fast_words, total_words; };
static RankLayout rank_layout(int R, int C) {
    RankLayout L;
    L.margin = (size_t)(C + 2) * 9;
    L.fast_words = ((size_t)9 * R * C + 2 * L.margin + 1) & ~(size_t)1;
    L.total_words = L.fast_words + (size_t)18 * R * C;
    return L;
}
static uint32_t host_orient_mask(int dR, int dC) {                // MAACO.py:146-157
    uint32_t k = 0xffu;
    if (dC > 0) k &= ~0x29u;
    if (dC < 0) k &= ~0x94u;
    if (dR > 0) k &= ~0x07u;
    if (dR < 0) k &= ~0xE0u;
    return k;
}

struct TourS1 {              // strategy-1 constants (P1 = orientation start -> target, MAACO.py:146-165)
    uint32_t P1;
    int fast_ok;            // P1 has exactly three moves
    int sm[3];              // those moves
    int dpr[3], dc[3];      // window row pointer delta (bytes) and column delta of each
    int so[3];              // ranking-word offset of (neighbour cell, context move+1)
};

struct TourArgs {
    TourS1 s1;
    const uint8_t *svalid;  // per-cell static move mask (MAACO move order)
    const uint32_t *rank;   // ranking buffer (mpp_maaco_rank) or null; layout: rank_layout()
    const uint32_t *rank_fast;   // word (cell*9 + ctx) of the strategy-1 table (margins on both sides)
    const uint2 *rank_slow;      // entry (cell*9 + ctx) of the full ranking
    int R, C, start, target;
    const double *tau, *E01;
    uint32_t it;
    double q0, alpha;
    int n_ants, ant_offset;
    uint32_t k0, k1;
    uint32_t *visitT;
    int32_t *cells;
    int max_cells;
    mpp_ant_result *result;
    unsigned long long *steps;
};

#ifndef MPP_TOUR_THREADS
#define MPP_TOUR_THREADS 256
#endif
#ifndef MPP_TOUR_MIN_BLOCKS
#define MPP_TOUR_MIN_BLOCKS 4
#endif
#define MPP_SQRT2 1.4142135623730951  // sqrt(2.0) correctly rounded == math.sqrt(2)


// Full roulette branch MAACO.py:255-262 (rare; kept out of line so it does not cost registers
// in the step loop).  Returns the rank (among candidates, in move order) of the selected move.
__device__ __noinline__ int roulette_rank(double attr, uint32_t cand, uint32_t gmask, int gshift, double S, double u1) {
    double at[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) at[i] = __shfl_sync(gmask, attr, gshift + i);
    const int n = __popc(cand);
    double pr[8], ps = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        pr[i] = at[i] / S;                                            // :255
        if ((cand >> i) & 1u) ps += pr[i];
    }
    if (fabs(ps - 1.0) > 1e-6) {                                      // :257-258
        const double ps0 = ps;
        ps = 0.0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            pr[i] = pr[i] / ps0;
            if ((cand >> i) & 1u) ps += pr[i];
        }
    }
    int k;
    if (!(fabs(ps - 1.0) <= 1.4901161193847656e-08)) {                // np.random.choice ValueError -> :262
        k = (int)(u1 * (double)n);
    } else {
        // RandomState.choice: cdf = cumsum(p); cdf /= cdf[-1]; searchsorted(u, side='right')
        double acc = 0.0, cdf[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if ((cand >> i) & 1u) acc += pr[i];
            cdf[i] = acc;
        }
        k = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (((cand >> i) & 1u) && (cdf[i] / acc <= u1)) ++k;
    }
    return k < n ? k : n - 1;
}

__device__ __noinline__ double pow_slow(double x, double y) { return pow(x, y); }

// Philox block `blk` of stream (seed, TOUR, it, ant) -> the two uniforms of ant step `blk` (out of line:
// runs once per LPA steps, keeps the ten rounds out of the step loop's register budget)
__device__ __noinline__ void tour_uniforms(uint32_t blk, uint32_t ant, uint32_t it, uint32_t k0, uint32_t k1,
                                           double &u0, double &u1) {
    const mpp_u4 rb = mpp_philox(blk, ant, it, MPP_CLS_MAACO_TOUR, k0, k1);
    u0 = mpp_u53(rb.x, rb.y);
    u1 = mpp_u53(rb.z, rb.w);
}

template <int LPA>
__device__ __forceinline__ uint32_t group_ballot(uint32_t gmask, int gshift, bool pred) {
    uint32_t b = __ballot_sync(gmask, pred);
    return (LPA == 32) ? (b & 0xffu) : ((b >> gshift) & 0xffu);  // only lanes m < 8 can vote true
}

// Per step each group (LPA lanes; lane m < 8 owns move m) does ONE round of loads -- the cell's static
// move mask (bounds + obstacles + crossing prohibition, precomputed per map), and per move the visited
// word, tau and E -- then one ballot, a REDUX max and the literal selection rules.  Philox blocks are
// generated LPA steps at a time (lane m computes the block of step base+m): two double shuffles per step.
template <int LPA>
__global__ void __launch_bounds__(MPP_TOUR_THREADS, (LPA == 32) ? MPP_TOUR_MIN_BLOCKS : 2) mpp_maaco_tour_kernel(const TourArgs A) {
    const int lane = threadIdx.x & 31;
    const int m = lane % LPA;                      // move index handled by this lane (m < 8 active)
    const int gshift = (LPA == 32) ? 0 : (lane / LPA) * LPA;
    const uint32_t gmask = (LPA == 32) ? 0xffffffffu : (((1u << LPA) - 1u) << gshift);
    const int a = (blockIdx.x * MPP_TOUR_THREADS + threadIdx.x) / LPA;
    if (a >= A.n_ants) return;
    // move order MAACO.py:98: (-1,-1),(-1,0),(-1,1),(0,-1),(0,1),(1,-1),(1,0),(1,1); (delta+1) packed 2 bits/move
    const int C = A.C, RC = A.R * A.C;
    const int mm = m & 7;
    const int delta = ((int)((0xA940u >> (2 * mm)) & 3u) - 1) * C + ((int)((0x9224u >> (2 * mm)) & 3u) - 1);
    const int target = A.target;
    const int tr = target / C, tc = target % C;
    int cur = A.start;
    // orientation masks MAACO.py:146-157
    auto orient_mask = [](int dR, int dC) -> uint32_t {
        uint32_t k = 0xffu;
        if (dC > 0) k &= ~0x29u;  // moves with dc<0: m0,m3,m5
        if (dC < 0) k &= ~0x94u;  // dc>0: m2,m4,m7
        if (dR > 0) k &= ~0x07u;  // dr<0: m0,m1,m2
        if (dR < 0) k &= ~0xE0u;  // dr>0: m5,m6,m7
        return k;
    };
    const uint32_t P1 = orient_mask(tr - cur / C, tc - cur % C);
    const uint32_t ant_global = (uint32_t)(A.ant_offset + a);
    const size_t n_ants = (size_t)A.n_ants;
    uint32_t *const visit_a = A.visitT + a;                      // word w of this ant at visit_a[w * n_ants]
    int32_t *const cells_a = A.cells + (size_t)a * A.max_cells;
    const double *const __restrict__ tau = A.tau;
    const double *const __restrict__ E01 = A.E01;                // interleaved: E[2*cell + turn]
    const uint8_t *const __restrict__ svalid = A.svalid;
    const uint32_t *const __restrict__ rank = A.rank;
    int n_path = 1, prev_m = -1, turns = 0;                       // steps taken == n_path - 1
    double len = 0.0;
    const int max_path = 2 * RC + 1;                              // step cap 2*R*C (MAACO.py:283); R*C < 2^30
    bool failed = false;
    double u0_l = 0.0, u1_l = 0.0;
    if (m == 0) {
        visit_a[(size_t)(cur >> 5) * n_ants] = 1u << (cur & 31);
        cells_a[0] = cur;
    }
    __syncwarp(gmask);
    int32_t *cell_out = cells_a + 1;                              // next path slot
    while (cur != target && n_path < max_path) {
        // ---- one round of loads: the cell's ranking word (or static mask) and this lane's visited word ----
        // ranking word (mpp_maaco_rank): [31:24] static move mask, [23:0] rank position of each move by
        // attractiveness (3 bits/move) for this (cell, previous move); 0xFFFFFF = "not small, use the full rule"
        const uint32_t ctx = (n_path >= 2) ? (uint32_t)(prev_m + 1) : 0u;
        const uint32_t rw = rank ? A.rank_slow[(size_t)cur * 9 + ctx].y : (((uint32_t)svalid[cur] << 24) | 0xFFFFFFu);
        const uint32_t sv = rw >> 24;                                 // bounds / obstacle / corner-cut (:93-120)
        int j = cur + delta;
        j = j < 0 ? 0 : (j >= RC ? RC - 1 : j);                       // clamp: lanes outside the mask are ignored
        const uint32_t tw = visit_a[(size_t)(j >> 5) * n_ants];
        // uniforms: draws 2s (q-test) and 2s+1 (selection) of stream (seed, TOUR, it, ant); block s
        const uint32_t step = (uint32_t)(n_path - 1), sub = step & (uint32_t)(LPA - 1);
        if (sub == 0) tour_uniforms(step + (uint32_t)m, ant_global, A.it, A.k0, A.k1, u0_l, u1_l);
        const double u0 = __shfl_sync(gmask, u0_l, gshift + (int)sub);
        const double u1 = __shfl_sync(gmask, u1_l, gshift + (int)sub);
        const bool free_lane = (m < 8) && ((sv >> m) & 1u) && !((tw >> (j & 31)) & 1u);   // + tabu :93-95
        const uint32_t valid = group_ballot<LPA>(gmask, gshift, free_lane);
        uint32_t cand = valid & P1;                                   // strategy 1 :165
        if (!cand) cand = valid & orient_mask(tr - cur / C, tc - cur % C);   // strategy 2 :169
        if (!cand) cand = valid;                                      // strategy 3 :172-180
        if (!cand) { failed = true; break; }                          // :287-288
        const bool in_c = (cand >> m) & 1u;                           // cand has 8 bits -> false for m >= 8
        uint32_t pool;   // the set the final uniform index is taken from
        int k = -1;      // >= 0: rank already decided by the roulette
        if ((rw & 0xFFFFFFu) != 0xFFFFFFu) {
            // every attractiveness around this cell is < 1e-10 (mpp_maaco_rank), hence:
            //  greedy  :241-250 -> |attr_i - max| < 1e-9 for all i: pool = first arg-max + every later candidate;
            //  roulette:251-254 -> sum < 1e-9: uniform over all candidates.
            // Only the ORDER of the attractiveness values matters, and that was ranked once per cell.
            if (u0 <= A.q0) {
                const uint32_t key = in_c ? ((((rw >> (3 * m)) & 7u) << 3) | (uint32_t)m) : 0xFFu;
                const int r = (int)(__reduce_min_sync(gmask, key) & 7u);   // best-ranked candidate == first arg-max
                pool = cand & ~((1u << r) - 1u);
            } else {
                pool = cand;
            }
        } else {
        const bool turn = (n_path >= 2) && (m != prev_m);             // MAACO.py:184-195
        const double tv = tau[j];
        const double ev = E01[2 * (size_t)j + (turn ? 1 : 0)];
        const double ta = (A.alpha == 1.0) ? tv : pow_slow(tv, A.alpha);   // tau**alpha (x**1.0 == x exactly)
        const double attr = ta * ev;                                  // :238
        // group max of attr over the candidates (attr >= 0: IEEE order == unsigned bit order)
        const unsigned long long key = in_c ? (unsigned long long)__double_as_longlong(attr) : 0ull;
        const uint32_t hi = (uint32_t)(key >> 32);
        const uint32_t mhi = __reduce_max_sync(gmask, hi);
        const uint32_t lo = (in_c && hi == mhi) ? (uint32_t)key : 0u;
        const uint32_t mlo = __reduce_max_sync(gmask, lo);
        const double mx = __longlong_as_double((long long)(((unsigned long long)mhi << 32) | mlo));
        if (u0 <= A.q0) {
            // greedy :241-250.  Sequential rule == {first arg-max r} U {i>r : |attr_i - max| < 1e-9}
            const uint32_t eq = group_ballot<LPA>(gmask, gshift, in_c && attr == mx);
            const int r = __ffs(eq) - 1;
            if (mx < 1e-9) {
                // every candidate lies in [0, max] with max < 1e-9, so |attr_i - max| < 1e-9 holds for all of
                // them: the pool is the first arg-max and every later candidate
                pool = cand & ~((1u << r) - 1u);
            } else {
                pool = group_ballot<LPA>(gmask, gshift, in_c && (m == r || (m > r && fabs(attr - mx) < 1e-9)));
            }
        } else {
            pool = cand;
            // :251-262.  sum() over np.float64 items == plain left-to-right.  n <= 8 terms <= mx, so
            // 8*mx < 0.9e-9 already implies S < 1e-9
            if (!(mx * 8.0 < 0.9e-9)) {
                double S = 0.0;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const double v = __shfl_sync(gmask, attr, gshift + i);
                    if ((cand >> i) & 1u) S += v;
                }
                if (!(S < 1e-9)) k = roulette_rank(attr, cand, gmask, gshift, S, u1);  // rare: near T only
            }
        }
        }
        if (k < 0) {                                                  // random.choice(pool) -> pool[floor(u*n)]
            const int n = __popc(pool);
            k = (int)(u1 * (double)n);
            k = k < n ? k : n - 1;
        }
        const uint32_t sel = group_ballot<LPA>(gmask, gshift, ((pool >> m) & 1u) && __popc(pool & ((1u << m) - 1u)) == k);
        const int pick = __ffs(sel) - 1;
        // ---- advance :293-297 ----
        len += ((0xA5u >> pick) & 1u) ? MPP_SQRT2 : 1.0;              // diagonal moves m0,m2,m5,m7
        if (n_path >= 2 && pick != prev_m) ++turns;                   // :264-276 counted on the fly
        prev_m = pick;
        if (m == pick) {
            visit_a[(size_t)(j >> 5) * n_ants] = tw | (1u << (j & 31));
            if (n_path < A.max_cells) *cell_out = j;
        }
        cur = __shfl_sync(gmask, j, gshift + pick);
        ++n_path;
        ++cell_out;
        __syncwarp(gmask);
    }
    if (m == 0) {
        const bool ok = !failed && cur == target;
        mpp_ant_result res;
        res.length = ok ? len : __longlong_as_double(0x7ff0000000000000ll);
        res.n_cells = ok ? n_path : 0;
        res.turns = ok ? turns : -1;
        A.result[a] = res;
        if (A.steps) atomicAdd(A.steps, (unsigned long long)(n_path - 1));
    }
}

// ---------------------------------------------------------------------------------------------
// K2, thread-per-ant form (default).  One ant step is a serial chain (position -> tabu bits -> candidate
// set -> selection -> position) and a colony pass lasts as long as its longest tour times the latency of that
// chain, so the chain stays inside ONE thread (no shuffles / votes / warp reductions on it) and touches only
// shared memory; a lone warp issues about one instruction per 5 cycles here, so what counts is the number of
// instructions on the chain and every exposed memory latency (profiles/r01_tour1_summary.md):
//   * tabu bits (:93-95) come from a 64x64-cell WINDOW of the ant's visited set in shared memory (rows of two
//     words; the ant stays at least U cells inside it).  A global store to a line evicts it from L1 (measured:
//     tools/ubench/l1_store.cu, 123 -> 432 cycles per dependent load), so re-reading a bitmap the ant itself
//     keeps writing costs an L2 round trip per step wherever it lives in global memory.  When the ant reaches
//     the window's edge the window slides by one 32-cell tile: the warp writes the leaving tile to the
//     word-major visitT (what the pheromone update streams) and reloads the entering one from there; the four
//     tiles still in the window are written at the end of the tour;
//   * strategy 1 (:165, almost every step): the strategy-1 word of (cell, previous move) from mpp_maaco_rank
//     -- static move mask + what greedy selection keeps of every subset of P1's three moves -- plus the
//     precomputed greedy flag / floor(u1*n) of the step index one shared-memory table row that yields the move
//     and all its deltas.  The words of the three possible next cells are prefetched one step ahead;
//   * every other step (no strategy-1 candidate, P1 without three moves, a cell whose attractiveness is not
//     tiny): all eight moves, strategies 2/3, the full ranking entry or the literal rules (tour_select_slow);
//   * the warp's lanes generate Philox blocks together: lane L makes the block of ant L%apw, step s+L/apw.
// `apw` lanes of each warp own an ant: fewer ants per warp = more warps to spread over the SMs.
// ---------------------------------------------------------------------------------------------
#define MPP_TOUR1_THREADS 128

// literal selection rules MAACO.py:228-262 for one ant; cand = candidate move mask (move order :98).
__device__ __noinline__ int tour_select_slow(uint32_t cand, int cr, int cc, int C, bool have_prev, int prev_m,
                                             const double *__restrict__ tau, const double *__restrict__ E01,
                                             double alpha, double q0, double u0, double u1) {
    int mv[8], n = 0;
    double attr[8];
#pragma unroll
    for (int m = 0; m < 8; ++m) {
        if (!((cand >> m) & 1u)) continue;
        const int dr = (int)((0xA940u >> (2 * m)) & 3u) - 1, dc = (int)((0x9224u >> (2 * m)) & 3u) - 1;
        const int j = (cr + dr) * C + (cc + dc);
        const bool turn = have_prev && (m != prev_m);                 // :184-195
        const double tv = tau[j];
        const double ta = (alpha == 1.0) ? tv : pow_slow(tv, alpha);
        attr[n] = ta * E01[2 * (size_t)j + (turn ? 1 : 0)];           // :238
        mv[n] = m;
        ++n;
    }
    int k;
    if (u0 <= q0) {                                                   // greedy :241-250
        double mx = -1.0;
        int best[8], nb = 0;
        for (int i = 0; i < n; ++i) {
            if (attr[i] > mx) { mx = attr[i]; nb = 0; best[nb++] = i; }
            else if (fabs(attr[i] - mx) < 1e-9) best[nb++] = i;
        }
        int q = (int)(u1 * (double)nb);
        q = q < nb ? q : nb - 1;
        k = best[q];
    } else {
        double S = 0.0;                                               // :252 plain left-to-right
        for (int i = 0; i < n; ++i) S += attr[i];
        bool uniform = S < 1e-9;                                      // :253-254
        double pr[8];
        double ps = 0.0;
        if (!uniform) {
            for (int i = 0; i < n; ++i) { pr[i] = attr[i] / S; ps += pr[i]; }   // :255
            if (fabs(ps - 1.0) > 1e-6) {                              // :257-258
                const double ps0 = ps;
                ps = 0.0;
                for (int i = 0; i < n; ++i) { pr[i] = pr[i] / ps0; ps += pr[i]; }
            }
            if (!(fabs(ps - 1.0) <= 1.4901161193847656e-08)) uniform = true;    // choice() ValueError -> :262
        }
        if (uniform) {
            k = (int)(u1 * (double)n);
        } else {
            // RandomState.choice: cdf = cumsum(p); cdf /= cdf[-1]; searchsorted(u, side='right')
            double cdf[8], acc = 0.0;
            for (int i = 0; i < n; ++i) { acc += pr[i]; cdf[i] = acc; }
            k = 0;
            for (int i = 0; i < n; ++i)
                if (cdf[i] / acc <= u1) ++k;
        }
        k = k < n ? k : n - 1;
    }
    return mv[k];
}

// visited bits of cells (r, 32*tcx .. 32*tcx+31) from the word-major bitmap (flat cell index, 32 per word)
__device__ __forceinline__ uint32_t tour_tile_word(const uint32_t *visit_a, size_t n_ants, int n_words, int r, int tcx,
                                                   int R, int C, int TC) {
    if (r < 0 || r >= R || tcx < 0 || tcx >= TC) return 0u;
    const int f = r * C + (tcx << 5), w = f >> 5, sh = f & 31;
    const uint32_t lo = __ldcg(visit_a + (size_t)w * n_ants);
    if (sh == 0) return lo;
    const uint32_t hi = (w + 1 < n_words) ? __ldcg(visit_a + (size_t)(w + 1) * n_ants) : 0u;
    return __funnelshift_r(lo, hi, sh);
}

// the reverse: merge a window word into the word-major bitmap.  Rows are always handled by lane (r & 31), so a later
// tour_tile_word of the same row is ordered after this by program order.  With C % 32 == 0 the window word IS the
// bitmap word (and holds everything the bitmap held: it was loaded from there), so a plain store does.
__device__ __forceinline__ void tour_tile_store(uint32_t *visit_a, size_t n_ants, int n_words, int r, int tcx, int R, int C,
                                                int TC, uint32_t v) {
    if (v == 0u || r < 0 || r >= R || tcx < 0 || tcx >= TC) return;
    const int f = r * C + (tcx << 5), w = f >> 5, sh = f & 31;
    if ((C & 31) == 0) { visit_a[(size_t)w * n_ants] = v; return; }
    atomicOr(visit_a + (size_t)w * n_ants, v << sh);
    if (sh != 0 && w + 1 < n_words && (v >> (32 - sh)) != 0u) atomicOr(visit_a + (size_t)(w + 1) * n_ants, v >> (32 - sh));
}

// a load the compiler may not sink to its use (it would turn "select among three loaded entries" into "one load
// from the selected address" and put the L2 latency back on the ant's serial chain)
__device__ __forceinline__ uint32_t ldg32_pinned(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// shared-memory loads by 32-bit shared address (kept in registers; the generic form re-derives the window base)
__device__ __forceinline__ uint2 lds_u2(uint32_t addr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint4 lds_u4(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ double lds_f64(uint32_t addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}

// shared-memory tables of the thread-per-ant kernel (per block)
struct Tour1Move {          // one per move (order MAACO.py:98)
    int dcur;               // flat cell delta  dr*C + dc
    int dprow;              // byte delta of the window row pointer  dr*8
    int dc;                 // column delta
    int pad;
};
#define T1_KTH_OFF 0                        // uint8  kth[256*8]   : index of the k-th set bit
#define T1_SPREAD_OFF 2048                  // uint2  spread[256]  : byte m = 0xFF if bit m set
#define T1_MOVE_OFF 4096                    // Tour1Move move[8]
#define T1_DLEN_OFF (4096 + 128)            // double dlen[8]      : 1.0 or sqrt(2) (:293)
#define T1_FAST_OFF (4096 + 256)            // uint4 fast[4*8]     : (k, subset) -> {dcur, dprow, dc, which | move << 8}
#define T1_RNG_OFF (4096 + 256 + 512)       // per warp: double2 u[32] (512 B) + uint32 pack[32] (128 B)
#define T1_WIN_OFF (T1_RNG_OFF + (MPP_TOUR1_THREADS / 32) * 640)

__global__ void __launch_bounds__(MPP_TOUR1_THREADS) mpp_maaco_tour1_kernel(const TourArgs A, const int apw) {
    extern __shared__ __align__(16) uint8_t t1_smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int R = A.R, C = A.C, TC = (C + 31) >> 5;
    {   // ---- tables ----
        uint8_t *const kth = t1_smem + T1_KTH_OFF;
        uint2 *const spread = (uint2 *)(t1_smem + T1_SPREAD_OFF);
        for (int p = threadIdx.x; p < 256; p += MPP_TOUR1_THREADS) {
            int idx = 0;
            uint32_t lo = 0, hi = 0;
            for (int b = 0; b < 8; ++b)
                if ((p >> b) & 1) {
                    kth[p * 8 + idx++] = (uint8_t)b;
                    if (b < 4) lo |= 0xFFu << (8 * b); else hi |= 0xFFu << (8 * (b - 4));
                }
            for (; idx < 8; ++idx) kth[p * 8 + idx] = 0;
            spread[p] = make_uint2(lo, hi);
        }
        if (threadIdx.x < 8) {
            const int m = threadIdx.x;
            const int dr = (int)((0xA940u >> (2 * m)) & 3u) - 1, dc = (int)((0x9224u >> (2 * m)) & 3u) - 1;
            Tour1Move mv;
            mv.dcur = dr * C + dc; mv.dprow = dr * 8; mv.dc = dc; mv.pad = 0;
            ((Tour1Move *)(t1_smem + T1_MOVE_OFF))[m] = mv;
            ((double *)(t1_smem + T1_DLEN_OFF))[m] = (dr != 0 && dc != 0) ? MPP_SQRT2 : 1.0;
        }
        if (threadIdx.x < 32) {                                    // (k, subset of P1's three moves) -> k-th member
            const int k = threadIdx.x >> 3, p3 = threadIdx.x & 7;
            const uint32_t P1t = A.s1.P1;
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if (__popc(P1t) == 3) {
                const int s[3] = {__ffs(P1t) - 1, __ffs(P1t & (P1t - 1)) - 1, 31 - __clz(P1t)};
                int seen = 0, which = 2;
                for (int i = 0; i < 3; ++i)
                    if ((p3 >> i) & 1) { if (seen == k) { which = i; break; } ++seen; }
                const int mm = s[which];
                const int dr = (int)((0xA940u >> (2 * mm)) & 3u) - 1, dc = (int)((0x9224u >> (2 * mm)) & 3u) - 1;
                v = make_uint4((uint32_t)(dr * C + dc), (uint32_t)(dr * 8), (uint32_t)dc, (uint32_t)which | ((uint32_t)mm << 8));
            }
            ((uint4 *)(t1_smem + T1_FAST_OFF))[threadIdx.x] = v;
        }
    }
    double2 *const rngu = (double2 *)(t1_smem + T1_RNG_OFF + wib * 640);
    uint32_t *const rngp = (uint32_t *)(t1_smem + T1_RNG_OFF + wib * 640 + 512);
    uint2 *const win_w = (uint2 *)(t1_smem + T1_WIN_OFF) + (size_t)wib * apw * 64;   // the warp's windows: 64 rows x uint2 each
    for (int i = lane; i < apw * 64; i += 32) win_w[i] = make_uint2(0u, 0u);          // visitT is zero on entry, so is the window
    __syncthreads();
    const int warp = (blockIdx.x * MPP_TOUR1_THREADS + threadIdx.x) >> 5;
    const int a0 = warp * apw;                                     // first ant of this warp
    if (a0 >= A.n_ants) return;                                    // whole warp
    const int a = a0 + lane;
    bool active = lane < apw && a < A.n_ants;
    const int n_words = (R * C + 31) >> 5;
    const int target = A.target;
    int cur = A.start;
    auto orient_mask = [](int dR, int dC) -> uint32_t {           // MAACO.py:146-157
        uint32_t k = 0xffu;
        if (dC > 0) k &= ~0x29u;
        if (dC < 0) k &= ~0x94u;
        if (dR > 0) k &= ~0x07u;
        if (dR < 0) k &= ~0xE0u;
        return k;
    };
    const int tr = target / C, tc = target % C;
    const uint32_t P1 = A.s1.P1;                                  // the same for every ant: start and target are the map's
    // the ranking entry of the next cell is fetched before the move is chosen: one load per strategy-1 move (the
    // first three of P1; P1 has 3 moves unless start and target share a row or column)
    // strategy 1 usually leaves three moves (a quadrant); then each step works on those three only
    const bool fast_ok = A.s1.fast_ok;
    const int sm0 = A.s1.sm[0], sm1 = A.s1.sm[1], sm2 = A.s1.sm[2];
    const int dpr0 = A.s1.dpr[0], dpr1 = A.s1.dpr[1], dpr2 = A.s1.dpr[2];
    const int dc0 = A.s1.dc[0], dc1 = A.s1.dc[1], dc2 = A.s1.dc[2];
    const int so0 = A.s1.so[0], so1 = A.s1.so[1], so2 = A.s1.so[2];
    // (kept in registers: re-reading kernel parameters inside the step costs a constant-bank round trip each time)
#define T1_KEEP(x) x = __shfl_sync(0xffffffffu, x, 0)   /* a value ptxas cannot re-derive from the parameter bank */
    int k_sm0 = sm0, k_sm1 = sm1, k_sm2 = sm2, k_dpr0 = dpr0, k_dpr1 = dpr1, k_dpr2 = dpr2;
    int k_dc0 = dc0, k_dc1 = dc1, k_dc2 = dc2, k_so0 = so0, k_so1 = so1, k_so2 = so2;
    T1_KEEP(k_sm0); T1_KEEP(k_sm1); T1_KEEP(k_sm2); T1_KEEP(k_dpr0); T1_KEEP(k_dpr1); T1_KEEP(k_dpr2);
    T1_KEEP(k_dc0); T1_KEEP(k_dc1); T1_KEEP(k_dc2); T1_KEEP(k_so0); T1_KEEP(k_so1); T1_KEEP(k_so2);
    const size_t n_ants = (size_t)A.n_ants;
    int32_t *const cells_a = A.cells + (size_t)(active ? a : a0) * A.max_cells;
    const uint32_t *__restrict__ rank_fast = A.rank_fast;
    {
        unsigned long long rf = (unsigned long long)rank_fast;
        rf = __shfl_sync(0xffffffffu, rf, 0);
        rank_fast = (const uint32_t *)rf;
    }
    int k_target = A.target, k_max_cells = A.max_cells;
    T1_KEEP(k_target); T1_KEEP(k_max_cells);
    int n_path = 1, prev_m = -1, turns = 0;
    double len = 0.0;
    const int max_path = 2 * R * C + 1;
    bool failed = false;
    // window = tile rows {wr, wr+1} x tile cols {wc, wc+1} (64 x 64 cells); row lrow of it is the uint2 at prow,
    // bit lcol of that 64-bit row is the cell; the ant stays in [1, 62] x [1, 62]
    int wr, wc, lrow, lcol;
    {
        const int cr = cur / C, cc = cur % C;
        wr = (cr >> 5) - (((cr & 31) < 16) ? 1 : 0);
        wc = (cc >> 5) - (((cc & 31) < 16) ? 1 : 0);
        lrow = cr - (wr << 5);
        lcol = cc - (wc << 5);
    }
    uint32_t prow_s = (uint32_t)__cvta_generic_to_shared(win_w + (size_t)(lane < apw ? lane : 0) * 64 + lrow);   // the ant's window row
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(t1_smem);
    uint32_t rngp_s = (uint32_t)__cvta_generic_to_shared(rngp) + 4u * (uint32_t)lane;
    uint32_t sbase_k = sbase;                                    // (opaque copy: otherwise re-derived from %cluster_ctaid every use)
    sbase_k = __shfl_sync(0xffffffffu, sbase_k, 0);
    uint32_t fw = 0u;                                             // strategy-1 word of (cell, previous move)
    uint32_t pf0 = 0u, pf1 = 0u, pf2 = 0u;                        // ... of the three strategy-1 neighbours, prefetched
    if (active) {
        asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(prow_s + 4u * ((uint32_t)lcol >> 5)), "r"(1u << (lcol & 31)) : "memory");
        cells_a[0] = cur;
        fw = rank_fast[(size_t)cur * 9];                          // context 0: no previous move
        // (read it here: a first use inside the loop would make every step wait on this load's scoreboard, which the
        // loop's own prefetches share.)  Field 0 is 0 or 7, so the test is never true.
        if (cur == target || (fw & 7u) == 3u) active = false;
        if (fast_ok) {
            const uint32_t *const rk = rank_fast + (size_t)cur * 9;
            pf0 = ldg32_pinned(rk + so0); pf1 = ldg32_pinned(rk + so1); pf2 = ldg32_pinned(rk + so2);
        }
    }
    __syncwarp();
    const uint32_t gmask = (uint32_t)(32 / apw) - 1u;              // a cooperative Philox pass covers 32/apw steps
    // U steps per pass of the warp-level bookkeeping below (liveness vote, Philox refill, window slides): an ant may
    // then be U cells from where the checks saw it, so the window keeps a margin of U cells instead of one
    const int U = (apw <= 16) ? 2 : 1;                             // a refill covers 32 / apw steps: must be a multiple of U
    const unsigned edge_lo = (unsigned)U, edge_span = 63u - 2u * (unsigned)U;
    for (uint32_t step = 0;; step += (uint32_t)U) {
        const uint32_t need0 = __ballot_sync(0xffffffffu, active && ((unsigned)lrow - edge_lo > edge_span || (unsigned)lcol - edge_lo > edge_span));
        if (!__any_sync(0xffffffffu, active)) break;
        if ((step & gmask) == 0 || need0) {
            if ((step & gmask) == 0) {
                // uniforms: draws 2s (q-test) and 2s+1 (selection) of stream (seed, TOUR, it, ant) = Philox block s;
                // lane L makes the block of ant L % apw for step + L / apw.  Besides u0/u1 (for the full rule) the
                // pass leaves what the ranking path needs: the greedy flag u0 <= q0 (:240) and floor(u1 * n) for
                // every pool size n = 1..8 (random.choice on n items).
                const int la = lane & (apw - 1);
                const mpp_u4 rb = mpp_philox(step + (uint32_t)(lane / apw), (uint32_t)(A.ant_offset + a0 + la), A.it,
                                             MPP_CLS_MAACO_TOUR, A.k0, A.k1);
                const double u0 = mpp_u53(rb.x, rb.y), u1 = mpp_u53(rb.z, rb.w);
                uint32_t pack = (u0 <= A.q0) ? (1u << 24) : 0u;
#pragma unroll
                for (int n = 2; n <= 8; ++n) {
                    int k = (int)(u1 * (double)n);
                    k = k < n ? k : n - 1;
                    pack |= (uint32_t)k << (3 * (n - 1));
                }
                __syncwarp();
                rngu[lane] = make_double2(u0, u1);
                rngp[lane] = pack;
                __syncwarp();
            }
            // ---- slide the windows whose ant reached their edge (at most every 31 steps per ant) ----
            uint32_t need = need0;
            while (need) {
                const int src = __ffs(need) - 1;
                const int s_lrow = __shfl_sync(0xffffffffu, lrow, src), s_lcol = __shfl_sync(0xffffffffu, lcol, src);
                const int s_wr = __shfl_sync(0xffffffffu, wr, src), s_wc = __shfl_sync(0xffffffffu, wc, src);
                uint32_t *const v_s = A.visitT + (a0 + src);
                uint2 *const w_s = win_w + src * 64;
                int d_wr = 0, d_wc = 0;
                if ((unsigned)s_lrow - edge_lo > edge_span) {      // vertical: rows move by 32, one tile row leaves, one enters
                    const bool up = s_lrow < U;
                    d_wr = up ? -1 : 1;
                    const int r_out = ((up ? s_wr + 1 : s_wr) << 5) + lane, r_in = ((up ? s_wr - 1 : s_wr + 2) << 5) + lane;
                    const uint2 keep = w_s[up ? lane : lane + 32], out = w_s[up ? lane + 32 : lane];
                    tour_tile_store(v_s, n_ants, n_words, r_out, s_wc, R, C, TC, out.x);
                    tour_tile_store(v_s, n_ants, n_words, r_out, s_wc + 1, R, C, TC, out.y);
                    uint2 nw;
                    nw.x = tour_tile_word(v_s, n_ants, n_words, r_in, s_wc, R, C, TC);
                    nw.y = tour_tile_word(v_s, n_ants, n_words, r_in, s_wc + 1, R, C, TC);
                    w_s[up ? lane + 32 : lane] = keep;
                    w_s[up ? lane : lane + 32] = nw;
                } else {                                           // horizontal: the two words of each row shift
                    const bool left = s_lcol < U;
                    d_wc = left ? -1 : 1;
                    const int t_out = left ? s_wc + 1 : s_wc, t_in = left ? s_wc - 1 : s_wc + 2;
                    const int r = (s_wr << 5) + lane;
                    const uint2 o0 = w_s[lane], o1 = w_s[lane + 32];
                    tour_tile_store(v_s, n_ants, n_words, r, t_out, R, C, TC, left ? o0.y : o0.x);
                    tour_tile_store(v_s, n_ants, n_words, r + 32, t_out, R, C, TC, left ? o1.y : o1.x);
                    const uint32_t x0 = tour_tile_word(v_s, n_ants, n_words, r, t_in, R, C, TC);
                    const uint32_t x1 = tour_tile_word(v_s, n_ants, n_words, r + 32, t_in, R, C, TC);
                    w_s[lane] = left ? make_uint2(x0, o0.x) : make_uint2(o0.y, x0);
                    w_s[lane + 32] = left ? make_uint2(x1, o1.x) : make_uint2(o1.y, x1);
                }
                if (lane == src) {
                    wr += d_wr; wc += d_wc;
                    lrow -= d_wr << 5; lcol -= d_wc << 5;
                    prow_s -= (uint32_t)(d_wr * 256);
                }
                __syncwarp();
                // a diagonal step can leave through a corner: this ant may still need the other direction
                const bool again = (lane == src) && ((unsigned)lrow - edge_lo > edge_span || (unsigned)lcol - edge_lo > edge_span);
                need = (need & (need - 1)) | __ballot_sync(0xffffffffu, again);
            }
        }
        for (uint32_t su = step; su < step + (uint32_t)U; ++su)
        if (active) {
            int m = -1, dcur = 0, dpr = 0, dcc = 0;                   // the move taken this step and its deltas
            uint32_t pi = 0u;                                         // which prefetched word applies next step
            const uint32_t pack = lds_u32(rngp_s + 4u * ((su & gmask) * (uint32_t)apw));
            if (fast_ok) {
                // ---- strategy 1 (:165) on the three moves of P1 only ----
                const uint2 ra = lds_u2(prow_s + k_dpr0), rb = lds_u2(prow_s + k_dpr1), rc = lds_u2(prow_s + k_dpr2);
                const int c0 = lcol + k_dc0, c1 = lcol + k_dc1, c2 = lcol + k_dc2;           // 0..63
                const uint32_t v0 = ((c0 & 32) ? ra.y : ra.x) >> (c0 & 31);            // tabu bit (:93-95) in bit 0
                const uint32_t v1 = ((c1 & 32) ? rb.y : rb.x) >> (c1 & 31);
                const uint32_t v2 = ((c2 & 32) ? rc.y : rc.x) >> (c2 & 31);
                const uint32_t sv = fw >> 24;                                          // static mask (:93-120)
                const uint32_t c3 = (((sv >> k_sm0) & ~v0) & 1u) | ((((sv >> k_sm1) & ~v1) & 1u) << 1) |
                                    ((((sv >> k_sm2) & ~v2) & 1u) << 2);                 // candidate subset of {k_sm0, k_sm1, k_sm2}
                if (c3 != 0u && (fw & 7u) == 0u) {
                    // greedy (:241-250): field c3 of the word = first arg-max + every later candidate, precomputed;
                    // roulette (:251-254): all candidates.  Then random.choice: the floor(u1*n)-th member.
                    const uint32_t p3 = (pack & (1u << 24)) ? ((fw >> (3u * c3)) & 7u) : c3;
                    const uint32_t sh = (0x63303000u >> (4u * p3)) & 15u;              // 3 * (popc(p3) - 1)
                    const uint32_t k = (pack >> sh) & 7u;
                    // k-th member of subset p3 and everything that follows from the move: one table row
                    const uint4 fv = lds_u4(sbase_k + T1_FAST_OFF + 16u * ((k << 3) | p3));
                    dcur = (int)fv.x; dpr = (int)fv.y; dcc = (int)fv.z;
                    pi = fv.w & 3u;
                    m = (int)(fv.w >> 8);
                }
            }
            if (m < 0) {
                // ---- general step: all eight moves, strategies 1-3, full ranking entry or the literal rules ----
                const uint2 q0r = lds_u2(prow_s - 8u), q1r = lds_u2(prow_s), q2r = lds_u2(prow_s + 8u);
                const int rot = lcol - 1;                             // 0..61
                const bool sw = rot & 32;
                const uint32_t t3 = __funnelshift_r(sw ? q0r.y : q0r.x, q0r.y, rot) & 7u;
                const uint32_t m3 = __funnelshift_r(sw ? q1r.y : q1r.x, q1r.y, rot) & 5u;
                const uint32_t b3 = __funnelshift_r(sw ? q2r.y : q2r.x, q2r.y, rot) & 7u;
                const uint32_t vis = t3 | ((m3 & 1u) << 3) | ((m3 & 4u) << 2) | (b3 << 5);
                const uint32_t valid = (fw >> 24) & ~vis;             // static mask (:93-120) minus tabu (:93-95)
                uint32_t cand = valid & P1;                           // strategy 1 :165
                if (cand == 0u) {
                    cand = valid & orient_mask(tr - cur / C, tc - cur % C);   // strategy 2 :169
                    if (!cand) cand = valid;                          // strategy 3 :172-180
                }
                if (cand == 0u) {                                     // :287-288
                    failed = true;
                    active = false;
                } else {
                    if ((cand & (cand - 1u)) == 0u) {
                        m = __ffs(cand) - 1;                          // one candidate: every rule picks it
                    } else if ((fw & 7u) != 0u) {                     // some attractiveness >= 1e-10: literal rules
                        const double2 uu = rngu[(su & gmask) * apw + lane];
                        m = tour_select_slow(cand, cur / C, cur % C, C, n_path >= 2, prev_m, A.tau, A.E01, A.alpha, A.q0, uu.x, uu.y);
                    } else {
                        uint32_t pool = cand;                         // roulette over tiny values: uniform (:253-254)
                        if (pack & (1u << 24)) {
                            // greedy: first arg-max = best-ranked candidate (full entry: permute the candidate flags
                            // into rank order, take the first) and every later candidate
                            const uint2 rws = A.rank_slow[(size_t)cur * 9 + (n_path >= 2 ? prev_m + 1 : 0)];
                            const uint2 sp = lds_u2(sbase_k + T1_SPREAD_OFF + 8u * cand);
                            const uint32_t f0 = __byte_perm(sp.x, sp.y, rws.x & 0xFFFFu), f1 = __byte_perm(sp.x, sp.y, rws.x >> 16);
                            const uint32_t ff = f0 ? f0 : f1;
                            const int pos4 = ((__ffs(ff) - 1) >> 1) + (f0 ? 0 : 16);
                            const uint32_t best = (rws.x >> pos4) & 7u;
                            pool = cand & ~((1u << best) - 1u);
                        }
                        const int n = __popc(pool);
                        const int k = (pack >> (3 * n - 3)) & 7u;
                        m = (int)lds_u8(sbase_k + T1_KTH_OFF + pool * 8u + (uint32_t)k);
                    }
                    const uint4 mvv = lds_u4(sbase_k + T1_MOVE_OFF + 16u * (uint32_t)m);
                    dcur = (int)mvv.x; dpr = (int)mvv.y; dcc = (int)mvv.z;
                    // (loaded into a prefetch register, never straight into fw: see the note at the first load)
                    pf0 = ldg32_pinned(rank_fast + (size_t)(cur + dcur) * 9 + (m + 1));
                    pi = 0u;
                }
            }
            if (active) {
                // ---- advance :293-297 ----
                cur += dcur;
                // bitwise select on `pi`, opaque to the compiler: as a ternary the default operand is copied early,
                // and that copy waits for the prefetches a whole selection too soon
                {
                    const uint32_t m1 = 0u - (pi & 1u), m2 = 0u - (pi >> 1);
                    uint32_t t;
                    asm("lop3.b32 %0, %1, %2, %3, 0xD8;" : "=r"(t) : "r"(pf0), "r"(pf1), "r"(m1));
                    asm("lop3.b32 %0, %1, %2, %3, 0xD8;" : "=r"(fw) : "r"(t), "r"(pf2), "r"(m2));
                }
                if (fast_ok) {
                    // the word of each strategy-1 neighbour of the NEW cell, for the end of the next step (the table has
                    // margins: every index is readable); a whole step of work hides the L2 latency
                    const uint32_t *const rk = rank_fast + (size_t)cur * 9;
                    pf0 = ldg32_pinned(rk + k_so0); pf1 = ldg32_pinned(rk + k_so1); pf2 = ldg32_pinned(rk + k_so2);
                }
                prow_s += (uint32_t)dpr;
                lrow += dpr >> 3;
                lcol += dcc;
                asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(prow_s + 4u * ((uint32_t)lcol >> 5)), "r"(1u << (lcol & 31)) : "memory");
                if (n_path < k_max_cells) cells_a[n_path] = cur;
                len += lds_f64(sbase_k + T1_DLEN_OFF + 8u * (uint32_t)m);
                if (n_path >= 2 && m != prev_m) ++turns;              // :264-276 counted on the fly
                prev_m = m;
                ++n_path;
                if (cur == k_target || n_path >= max_path) active = false;
            }
        }
    }
    // ---- the windows still hold the marks of their four tiles: merge them into visitT ----
    for (int src = 0; src < apw && a0 + src < A.n_ants; ++src) {
        const int s_wr = __shfl_sync(0xffffffffu, wr, src), s_wc = __shfl_sync(0xffffffffu, wc, src);
        uint32_t *const v_s = A.visitT + (a0 + src);
        const uint2 o0 = win_w[src * 64 + lane], o1 = win_w[src * 64 + lane + 32];
        const int r = (s_wr << 5) + lane;
        tour_tile_store(v_s, n_ants, n_words, r, s_wc, R, C, TC, o0.x);
        tour_tile_store(v_s, n_ants, n_words, r, s_wc + 1, R, C, TC, o0.y);
        tour_tile_store(v_s, n_ants, n_words, r + 32, s_wc, R, C, TC, o1.x);
        tour_tile_store(v_s, n_ants, n_words, r + 32, s_wc + 1, R, C, TC, o1.y);
    }
    if (lane < apw && a < A.n_ants) {
        const bool ok = !failed && cur == target;
        mpp_ant_result res;
        res.length = ok ? len : __longlong_as_double(0x7ff0000000000000ll);
        res.n_cells = ok ? n_path : 0;
        res.turns = ok ? turns : -1;
        A.result[a] = res;
        if (A.steps) atomicAdd(A.steps, (unsigned long long)(n_path - 1));
    }
}

extern "C" int mpp_maaco_tours(const mpp_map *map, const double *tau_dev, const double *E01_dev,
                               const uint32_t *rank_dev, int iteration, double q0, double alpha, int n_ants, int ant_offset, uint64_t seed,
                               uint32_t *visitT_dev, int32_t *cells_dev, int max_cells, mpp_ant_result *result_dev,
                               unsigned long long *steps_dev, int lanes_per_ant, void *stream) {
    MPP_REQUIRE(map && tau_dev && E01_dev && visitT_dev && cells_dev && result_dev, "mpp_maaco_tours: null argument");
    MPP_REQUIRE(map->start >= 0 && map->target >= 0, "mpp_maaco_tours: map has no start/target");
    MPP_REQUIRE(n_ants > 0 && max_cells > 0, "mpp_maaco_tours: n_ants=%d max_cells=%d", n_ants, max_cells);
    int apw_hint = 0;                                             // lanes_per_ant = -k: thread per ant, k ants per warp
    if (lanes_per_ant < 0) { apw_hint = -lanes_per_ant; lanes_per_ant = 1; }
    if (lanes_per_ant == 0) {
        const char *e = getenv("MPP_TOUR_LPA");
        lanes_per_ant = e ? atoi(e) : 1;   // measured on B200: thread-per-ant beats the cooperative forms from 256 to 16k ants
    }
    MPP_REQUIRE(lanes_per_ant == 1 || lanes_per_ant == 8 || lanes_per_ant == 16 || lanes_per_ant == 32,
                "mpp_maaco_tours: lanes_per_ant must be 1, 8, 16 or 32");
    MPP_CUDA(cudaSetDevice(map->device));
    TourArgs A;
    A.svalid = map->svalid_dev;
    A.rank = rank_dev;
    {
        TourS1 &S = A.s1;
        S.P1 = host_orient_mask(map->target / map->cols - map->start / map->cols, map->target % map->cols - map->start % map->cols);
        S.fast_ok = __builtin_popcount(S.P1) == 3;
        uint32_t rest = S.P1;
        for (int i = 0; i < 3; ++i) {
            const int m = S.fast_ok ? __builtin_ctz(rest) : __builtin_ctz(S.P1);
            rest &= rest - 1;
            const int dr = (int)((0xA940u >> (2 * m)) & 3u) - 1, dc = (int)((0x9224u >> (2 * m)) & 3u) - 1;
            S.sm[i] = m; S.dpr[i] = dr * 8; S.dc[i] = dc; S.so[i] = (dr * map->cols + dc) * 9 + m + 1;
        }
    }
    {
        const RankLayout L = rank_layout(map->rows, map->cols);
        A.rank_fast = rank_dev ? rank_dev + L.margin : nullptr;
        A.rank_slow = rank_dev ? (const uint2 *)(rank_dev + L.fast_words) : nullptr;
    }
    A.R = map->rows; A.C = map->cols; A.start = map->start; A.target = map->target;
    A.tau = tau_dev; A.E01 = E01_dev;
    A.it = (uint32_t)iteration; A.q0 = q0; A.alpha = alpha;
    A.n_ants = n_ants; A.ant_offset = ant_offset;
    A.k0 = (uint32_t)seed; A.k1 = (uint32_t)(seed >> 32);
    A.visitT = visitT_dev; A.cells = cells_dev; A.max_cells = max_cells;
    A.result = result_dev; A.steps = steps_dev;
    if (lanes_per_ant == 1 && !rank_dev) lanes_per_ant = 32;      // the thread-per-ant kernel needs the ranking table
    if (lanes_per_ant == 1) {
        // ants per warp: enough warps for every SM sub-partition first, full warps only for big colonies
        int apw = 32;
        const char *e = getenv("MPP_TOUR_APW");
        if (e) apw = atoi(e);
        else if (apw_hint) apw = apw_hint;
        else while (apw > 2 && (n_ants + apw - 1) / apw < 12 * map->sm_count) apw >>= 1;   // measured: 2 ants/warp at 4096 ants
        MPP_REQUIRE(apw == 1 || apw == 2 || apw == 4 || apw == 8 || apw == 16 || apw == 32, "ants per warp (MPP_TOUR_APW / -lanes_per_ant) must be a power of two <= 32");
        const int warps = (n_ants + apw - 1) / apw, wpb = MPP_TOUR1_THREADS / 32;
        const size_t smem = T1_WIN_OFF + (size_t)wpb * apw * 512;   // tables + per-warp RNG + 512-byte windows
        {   // raise the kernel's dynamic shared-memory limit only when it grows (the call is slow; devices tracked apart)
            static size_t smem_set[64] = {0};
            const int dv = map->device & 63;
            if (smem > smem_set[dv]) {
                MPP_CUDA(cudaFuncSetAttribute(mpp_maaco_tour1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                smem_set[dv] = smem;
            }
        }
        mpp_maaco_tour1_kernel<<<(warps + wpb - 1) / wpb, MPP_TOUR1_THREADS, smem, (cudaStream_t)stream>>>(A, apw);
        MPP_CUDA(cudaGetLastError());
        return MPP_OK;
    }
    const int ants_per_block = MPP_TOUR_THREADS / lanes_per_ant;
    const int blocks = (n_ants + ants_per_block - 1) / ants_per_block;
    cudaStream_t s = (cudaStream_t)stream;
    if (lanes_per_ant == 32) mpp_maaco_tour_kernel<32><<<blocks, MPP_TOUR_THREADS, 0, s>>>(A);
    else if (lanes_per_ant == 16) mpp_maaco_tour_kernel<16><<<blocks, MPP_TOUR_THREADS, 0, s>>>(A);
    else mpp_maaco_tour_kernel<8><<<blocks, MPP_TOUR_THREADS, 0, s>>>(A);
    MPP_CUDA(cudaGetLastError());
    return MPP_OK;
}

// ---------------------------------------------------------------------------------------------
// K2a: per (cell, turn context) ranking of the 8 moves by attractiveness tau**alpha * eta'**beta (MAACO.py:238).
// With beta = 7 the values are ~1e-8..1e-20, so the selection rules (:241-262) only depend on their ORDER
// (|attr_i - max| < 1e-9 and sum < 1e-9 always hold); the tour kernels then need one or two words per step
// instead of 16 fp64 loads.  Context 0 = no previous move (turn flag 0 for every candidate, :185-186),
// context p+1 = previous move p (turn flag = (m != p)).  Entry = two words.  Word 1: [31:24] static move mask,
// [23:0] rank position of each move (0 = largest attractiveness, ties -> lower move index, exactly the order the
// sequential scan sees), or 0xFFFFFF when some attractiveness is >= 1e-10 (the tour kernels then apply the full
// rule).  Word 0: the inverse permutation, nibble p = move at rank position p (a PRMT selector).
// ---------------------------------------------------------------------------------------------

extern "C" long long mpp_maaco_rank_words(const mpp_map *map) {
    if (!map) return 0;
    return (long long)rank_layout(map->rows, map->cols).total_words;
}

// Strategy-1 word (one per (cell, context)): [31:24] static move mask; field c (3 bits at 3c, c = 1..7 = a subset of
// P1's three moves in move order) = what greedy selection (:241-250) keeps of candidate set c, as a subset again;
// field 0 = 0 when the ranking applies (all attractiveness < 1e-10), 7 when it does not; fields 1..7 are only
// filled when P1 has exactly three moves.
__global__ void __launch_bounds__(128) mpp_maaco_rank_kernel(const uint8_t *__restrict__ svalid,
                                                             const double *__restrict__ tau,
                                                             const double *__restrict__ E01, double alpha, int R, int C,
                                                             uint32_t P1, uint32_t *__restrict__ rank_fast,
                                                             uint2 *__restrict__ rank_slow) {
    // one thread per cell: the eight neighbours' tau / eta' are read once and serve all nine contexts
    const int cell = blockIdx.x * 128 + threadIdx.x;
    if (cell >= R * C) return;
    const uint32_t sv = svalid[cell];
    double a0[8], a1[8];                                          // attractiveness without / with the turn factor (:238)
    double mx = 0.0;
#pragma unroll
    for (int m = 0; m < 8; ++m) {
        a0[m] = a1[m] = -1.0;
        if ((sv >> m) & 1u) {
            const int j = cell + ((int)((0xA940u >> (2 * m)) & 3u) - 1) * C + ((int)((0x9224u >> (2 * m)) & 3u) - 1);
            const double tv = tau[j];
            const double ta = (alpha == 1.0) ? tv : pow_slow(tv, alpha);
            const double2 e = *(const double2 *)(E01 + 2 * (size_t)j);
            a0[m] = ta * e.x;
            a1[m] = ta * e.y;
            mx = fmax(mx, fmax(a0[m], a1[m]));
        }
    }
    const bool p1_three = __popc(P1) == 3;
    const int s0 = __ffs(P1) - 1, s1 = __ffs(P1 & (P1 - 1)) - 1, s2 = 31 - __clz(P1);
    for (int ctx = 0; ctx < 9; ++ctx) {
        double a[8];
#pragma unroll
        for (int m = 0; m < 8; ++m) a[m] = (ctx > 0 && m != ctx - 1) ? a1[m] : a0[m];   // turn flag :184-195
        // "small" is decided per context on the values that context uses (as the tour's full rule would see them)
        double cmx = 0.0;
#pragma unroll
        for (int m = 0; m < 8; ++m) cmx = fmax(cmx, a[m]);
        uint32_t word = 0xFFFFFFu, perm = 0u, fast = 7u;
        if (cmx < 1e-10) {
            word = 0u;
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                int pos = 0;
#pragma unroll
                for (int q = 0; q < 8; ++q) pos += (a[q] > a[m]) || (a[q] == a[m] && q < m);
                word |= (uint32_t)pos << (3 * m);
                perm |= (uint32_t)m << (4 * pos);
            }
            fast = 0u;                                            // field 0: the ranking applies
            if (p1_three) {
                const uint32_t p0 = (word >> (3 * s0)) & 7u, p1 = (word >> (3 * s1)) & 7u, p2 = (word >> (3 * s2)) & 7u;
#pragma unroll
                for (uint32_t c = 1; c < 8; ++c) {
                    // best-ranked member of subset c (rank positions are distinct)
                    uint32_t bp = 8u, bi = 0u;
                    if ((c & 1u) && p0 < bp) { bp = p0; bi = 0u; }
                    if ((c & 2u) && p1 < bp) { bp = p1; bi = 1u; }
                    if ((c & 4u) && p2 < bp) { bp = p2; bi = 2u; }
                    fast |= (c & ~((1u << bi) - 1u)) << (3u * c);     // the best one and every later member
                }
            }
        }
        const size_t t = (size_t)cell * 9 + ctx;
        rank_fast[t] = fast | (sv << 24);
        rank_slow[t] = make_uint2(perm, word | (sv << 24));
    }
    (void)mx;
}

extern "C" int mpp_maaco_rank(const mpp_map *map, const double *tau_dev, const double *E01_dev, double alpha,
                              uint32_t *rank_dev, void *stream) {
    MPP_REQUIRE(map && tau_dev && E01_dev && rank_dev, "mpp_maaco_rank: null argument");
    MPP_REQUIRE(map->start >= 0 && map->target >= 0, "mpp_maaco_rank: map has no start/target");
    MPP_CUDA(cudaSetDevice(map->device));
    const int total = map->rows * map->cols;
    const RankLayout L = rank_layout(map->rows, map->cols);
    const uint32_t P1 = host_orient_mask(map->target / map->cols - map->start / map->cols,
                                         map->target % map->cols - map->start % map->cols);
    mpp_maaco_rank_kernel<<<(total + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
        map->svalid_dev, tau_dev, E01_dev, alpha, map->rows, map->cols, P1, rank_dev + L.margin,
        (uint2 *)(rank_dev + L.fast_words));
    MPP_CUDA(cudaGetLastError());
    return MPP_OK;
}

// ---------------------------------------------------------------------------------------------
// K4: iteration-best / overall-best (MAACO.py:343-358) + deposits (:307-308)
// ---------------------------------------------------------------------------------------------
#define MPP_BEST_THREADS 1024
__device__ __forceinline__ bool lt_len_idx(double la, int ia, double lb, int ib) {
    return la < lb || (la == lb && ia < ib);
}

__global__ void __launch_bounds__(MPP_BEST_THREADS)
mpp_maaco_best_kernel(const mpp_ant_result *__restrict__ res, const int32_t *__restrict__ cells, int max_cells,
                      int cells_ant_offset, int cells_n_ants, int n, double Q, int iteration, mpp_maaco_state *state,
                      int32_t *best_cells, double *deposit, double *log) {
    __shared__ double s_len[32];
    __shared__ int s_idx[32];
    __shared__ int s_t[32];
    __shared__ double b_len;
    __shared__ int b_r, b_ant, b_turns, b_copy;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const double INF = __longlong_as_double(0x7ff0000000000000ll);
    // phase 1: global min length and its first index r (the last strict record of the scan)
    double ml = INF;
    int mi = 0x7fffffff;
    for (int i = tid; i < n; i += MPP_BEST_THREADS) {
        const mpp_ant_result ri = res[i];
        const double l = ri.length;
        if (lt_len_idx(l, i, ml, mi)) { ml = l; mi = i; }
        // deposit amount MAACO.py:307-308 (0.0 == "does not deposit"; x + 0.0 is exact anyway)
        deposit[i] = (l != INF && ri.n_cells > 0 && l > 1e-6) ? Q / l : 0.0;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        const double ol = __shfl_xor_sync(0xffffffffu, ml, o);
        const int oi = __shfl_xor_sync(0xffffffffu, mi, o);
        if (lt_len_idx(ol, oi, ml, mi)) { ml = ol; mi = oi; }
    }
    if (lane == 0) { s_len[wid] = ml; s_idx[wid] = mi; }
    __syncthreads();
    if (wid == 0) {
        ml = s_len[lane]; mi = s_idx[lane];
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const double ol = __shfl_xor_sync(0xffffffffu, ml, o);
            const int oi = __shfl_xor_sync(0xffffffffu, mi, o);
            if (lt_len_idx(ol, oi, ml, mi)) { ml = ol; mi = oi; }
        }
        if (lane == 0) { b_len = ml; b_r = mi; }
    }
    __syncthreads();
    const double L = b_len;
    const int r = b_r;
    // phase 2: among {r} U {i>r : |len_i - L| < 1e-9} the first index with the fewest turns
    int bt = 0x7fffffff, bi = 0x7fffffff;
    if (L != INF) {
        for (int i = tid; i < n; i += MPP_BEST_THREADS) {
            if (i < r) continue;
            const mpp_ant_result ri = res[i];
            const double l = ri.length;
            if (i == r || fabs(l - L) < 1e-9) {
                const int t = ri.turns;
                if (t >= 0 && (t < bt || (t == bt && i < bi))) { bt = t; bi = i; }
            }
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        const int ot = __shfl_xor_sync(0xffffffffu, bt, o), oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ot < bt || (ot == bt && oi < bi)) { bt = ot; bi = oi; }
    }
    if (lane == 0) { s_t[wid] = bt; s_idx[wid] = bi; }
    __syncthreads();
    if (wid == 0) {
        bt = s_t[lane]; bi = s_idx[lane];
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const int ot = __shfl_xor_sync(0xffffffffu, bt, o), oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ot < bt || (ot == bt && oi < bi)) { bt = ot; bi = oi; }
        }
        if (lane == 0) {
            const bool any = (L != INF);
            b_ant = any ? bi : -1;
            b_turns = any ? bt : -1;
            // overall best MAACO.py:351-358
            mpp_maaco_state st = *state;
            int copy = 0;
            if (any) {
                if (L < st.best_len) {
                    st.best_len = L; st.best_turns = bt; copy = 1;
                } else if (fabs(L - st.best_len) < 1e-9) {
                    if (bt < st.best_turns) { st.best_turns = bt; copy = 1; }
                }
            }
            if (copy) { st.best_n_cells = res[bi].n_cells; st.best_iter = iteration; st.best_ant = bi; }
            st.iter_best_len = L; st.iter_best_turns = b_turns; st.iter_best_ant = b_ant;
            *state = st;
            b_copy = copy;
            if (log) {
                double *lg = log + 4 * (size_t)(iteration - 1);
                lg[0] = L; lg[1] = (double)b_turns; lg[2] = st.best_len; lg[3] = (double)st.best_turns;
            }
        }
    }
    __syncthreads();
    // the path is copied only by the rank that owns the ant (sharded colony: cells holds ants
    // [cells_ant_offset, cells_ant_offset + cells_n_ants)); the owner broadcasts it after the solve
    if (b_copy && b_ant >= cells_ant_offset && b_ant < cells_ant_offset + cells_n_ants) {
        const int a = b_ant - cells_ant_offset;
        int nc = res[b_ant].n_cells;
        if (nc > max_cells) nc = max_cells;
        const int32_t *src = cells + (size_t)a * max_cells;
        for (int k = tid; k < nc; k += MPP_BEST_THREADS) best_cells[k] = src[k];
    }
}

extern "C" int mpp_maaco_best(const mpp_ant_result *result_dev, const int32_t *cells_dev, int max_cells,
                              int cells_ant_offset, int cells_n_ants, int n_ants, double Q, int iteration,
                              mpp_maaco_state *state_dev, int32_t *best_cells_dev, double *deposit_dev,
                              double *log_dev, void *stream) {
    MPP_REQUIRE(result_dev && cells_dev && state_dev && best_cells_dev && deposit_dev, "mpp_maaco_best: null argument");
    MPP_REQUIRE(n_ants > 0 && iteration >= 1, "mpp_maaco_best: n_ants=%d iteration=%d", n_ants, iteration);
    mpp_maaco_best_kernel<<<1, MPP_BEST_THREADS, 0, (cudaStream_t)stream>>>(
        result_dev, cells_dev, max_cells, cells_ant_offset, cells_n_ants, n_ants, Q, iteration, state_dev,
        best_cells_dev, deposit_dev, log_dev);
    MPP_CUDA(cudaGetLastError());
    return MPP_OK;
}

// ---------------------------------------------------------------------------------------------
// K3: pheromone evaporate + ordered deposit + MMAS clip (MAACO.py:304-332)
// One warp owns 32 consecutive cells (one bitmap word); it streams that word of every ant
// (word-major layout => 128-B coalesced loads of 32 ants) and adds deposits in ant order.
// ---------------------------------------------------------------------------------------------
#define MPP_PHER_THREADS 256
struct __align__(16) PherEntry { double d; uint32_t w; uint32_t pad; };

// MMAS clip + obstacle reset (MAACO.py:312-332) for one cell
__device__ __forceinline__ double pher_finalize(double t, int cell, const uint32_t *occ, int pitch, int R, int C,
                                                double rho, const mpp_maaco_state *state) {
    double b = state->best_len;                                               // :312-316
    if (b == __longlong_as_double(0x7ff0000000000000ll)) b = (double)(R + C);
    if (b < 1e-6) b = 1e-6;
    const double tmax = (1.0 / (1.0 - rho)) * (1.0 / b);                      // :317
    int mx = C > R ? C : R;
    if (mx < 1) mx = 1;
    const double tmin = tmax / (2.0 * (double)mx);                            // :323
    const int r = cell / C, c = cell % C, pb = c + 1;
    const bool obst = (occ[(r + 1) * pitch + (pb >> 5)] >> (pb & 31)) & 1u;
    if (obst) return 1e-9;                                                    // :332
    t = t > tmin ? t : tmin;                                                  // :327-331 np.clip
    return t < tmax ? t : tmax;
}

// One CTA per bitmap word.  A block issues the loads of 4096 ants
// (16 words per lane) at once, so a word costs ONE round trip.  The hits -- typically a few dozen per word -- are
// compacted in ant order into a small list (warp-count prefix over shared memory) that warp 0 folds; words with more
// hits than the list holds (around the start cell) fall back to one warp's 512 ants at a time.
#define MPP_PHER_SR 4096
#define MPP_PHER_SRU (MPP_PHER_SR / 8 / 32)   // loads per lane per super-round (8 warps)
#define MPP_PHER_CAP 512                     // list entries per buffer (>= ants per warp per super-round)

__device__ __forceinline__ double pher_fold_list(const PherEntry *lst, int n, int lane, double t) {
    int i = 0;
    for (; i + 4 <= n; i += 4) {
        // select the operand, not the sum: the loop-carried chain is a bare DADD (t + 0.0 == t exactly)
        const PherEntry x0 = lst[i], x1 = lst[i + 1], x2 = lst[i + 2], x3 = lst[i + 3];
        const double a0 = ((x0.w >> lane) & 1u) ? x0.d : 0.0, a1 = ((x1.w >> lane) & 1u) ? x1.d : 0.0;
        const double a2 = ((x2.w >> lane) & 1u) ? x2.d : 0.0, a3 = ((x3.w >> lane) & 1u) ? x3.d : 0.0;
        t += a0;                                                             // :311
        t += a1;
        t += a2;
        t += a3;
    }
    for (; i < n; ++i) {
        const PherEntry x = lst[i];
        t += ((x.w >> lane) & 1u) ? x.d : 0.0;
    }
    return t;
}

__global__ void __launch_bounds__(MPP_PHER_THREADS, 3)
mpp_maaco_pheromone_sr_kernel(const uint32_t *__restrict__ occ, int pitch, int R, int C, double *__restrict__ tau,
                              uint32_t *__restrict__ visitT, const double *__restrict__ deposit, int n_seg,
                              int seg_ants, int word0, int n_words, double rho,
                              const mpp_maaco_state *__restrict__ state, int clear_visit) {
    __shared__ PherEntry s_list[2][MPP_PHER_CAP];
    __shared__ int s_cnt[2][8];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int wl = blockIdx.x;                               // one bitmap word (32 cells) per block
    const int cell = (word0 + wl) * 32 + lane;
    const bool live = cell < R * C;
    double t = 0.0;
    if (wid == 0 && live) t = tau[cell] * (1.0 - rho);       // :305
    const int sr_per_seg = (seg_ants + MPP_PHER_SR - 1) / MPP_PHER_SR;
    const int n_sr = n_seg * sr_per_seg;
    const uint32_t lt = (1u << lane) - 1u;
    for (int sr = 0; sr < n_sr; ++sr) {
        const int buf = sr & 1;
        const int seg = sr / sr_per_seg, a_base = (sr % sr_per_seg) * MPP_PHER_SR + wid * (MPP_PHER_SR / 8);
        uint32_t *const row = visitT + ((size_t)seg * n_words + wl) * seg_ants;
        const double *const dep = deposit + (size_t)seg * seg_ants;
        uint32_t wd[MPP_PHER_SRU];
#pragma unroll
        for (int u = 0; u < MPP_PHER_SRU; ++u) {
            const int a = a_base + u * 32 + lane;
            wd[u] = (a < seg_ants) ? row[a] : 0u;
        }
        int cnt = 0;
#pragma unroll
        for (int u = 0; u < MPP_PHER_SRU; ++u) cnt += __popc(__ballot_sync(0xffffffffu, wd[u] != 0u));
        if (lane == 0) s_cnt[buf][wid] = cnt;
        double dv[MPP_PHER_SRU];                             // deposits of the hits: all loads in flight before the barrier
#pragma unroll
        for (int u = 0; u < MPP_PHER_SRU; ++u) {
            dv[u] = 0.0;
            if (wd[u] != 0u) {
                const int a = a_base + u * 32 + lane;
                dv[u] = dep[a];
                if (clear_visit) row[a] = 0u;
            }
        }
        __syncthreads();
        int off = 0, total = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int c = s_cnt[buf][k];
            if (k < wid) off += c;
            total += c;
        }
        if (total <= MPP_PHER_CAP) {
            int o = off;
#pragma unroll
            for (int u = 0; u < MPP_PHER_SRU; ++u) {
                const uint32_t nz = __ballot_sync(0xffffffffu, wd[u] != 0u);
                if (wd[u] != 0u) {
                    PherEntry e; e.d = dv[u]; e.w = wd[u]; e.pad = 0u;
                    s_list[buf][o + __popc(nz & lt)] = e;
                }
                o += __popc(nz);
            }
            __syncthreads();
            if (wid == 0) t = pher_fold_list(s_list[buf], total, lane, t);   // ants in index order :306
        } else {
            // crowded word: one warp's ants (<= 512 hits) at a time through the same list
            for (int c = 0; c < 8; ++c) {
                if (wid == c) {
                    int o = 0;
#pragma unroll
                    for (int u = 0; u < MPP_PHER_SRU; ++u) {
                        const uint32_t nz = __ballot_sync(0xffffffffu, wd[u] != 0u);
                        if (wd[u] != 0u) {
                            PherEntry e; e.d = dv[u]; e.w = wd[u]; e.pad = 0u;
                            s_list[buf][o + __popc(nz & lt)] = e;
                        }
                        o += __popc(nz);
                    }
                }
                __syncthreads();
                if (wid == 0) t = pher_fold_list(s_list[buf], s_cnt[buf][c], lane, t);
                __syncthreads();
            }
        }
    }
    if (wid == 0 && live) tau[cell] = pher_finalize(t, cell, occ, pitch, R, C, rho, state);
}

extern "C" int mpp_maaco_pheromone(const mpp_map *map, double *tau_dev, uint32_t *visitT_dev,
                                   const double *deposit_dev, int n_seg, int seg_ants, int word0, int n_words,
                                   double rho, const mpp_maaco_state *state_dev, int clear_visit, void *stream) {
    MPP_REQUIRE(map && tau_dev && visitT_dev && deposit_dev && state_dev, "mpp_maaco_pheromone: null argument");
    MPP_REQUIRE(n_seg > 0 && seg_ants > 0 && word0 >= 0 && n_words > 0, "mpp_maaco_pheromone: bad shape");
    MPP_CUDA(cudaSetDevice(map->device));
    mpp_maaco_pheromone_sr_kernel<<<n_words, MPP_PHER_THREADS, 0, (cudaStream_t)stream>>>(
        map->occ_dev, map->pitch_words, map->rows, map->cols, tau_dev, visitT_dev, deposit_dev, n_seg, seg_ants,
        word0, n_words, rho, state_dev, clear_visit);
    MPP_CUDA(cudaGetLastError());
    return MPP_OK;
}

// ---------------------------------------------------------------------------------------------
// Sharded-colony exchange: tours travel between GPUs as 1-byte move codes (MAACO move order), ~10x
// smaller than dense visited bitmaps and 4x smaller than int32 cell lists; every rank replays the codes of
// all ants into the visited words of ITS slice of the map.
// ---------------------------------------------------------------------------------------------
// offsets[i] = byte offset of ant i's codes inside its segment's packed buffer; totals[seg] = bytes of
// segment seg.  One block per segment (exclusive scan over the segment's ants of n_cells-1, 0 for failed ants).
__global__ void __launch_bounds__(1024) mpp_maaco_move_offsets_kernel(const mpp_ant_result *__restrict__ res,
                                                                      int seg_ants, int32_t *__restrict__ offsets,
                                                                      int32_t *__restrict__ totals) {
    __shared__ int s_warp[32];
    __shared__ int s_carry;
    const int seg = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < seg_ants; base += 1024) {
        const int a = base + tid;
        int len = 0;
        if (a < seg_ants) { const int n = res[(size_t)seg * seg_ants + a].n_cells; len = n > 0 ? n - 1 : 0; }
        int x = len;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        if (lane == 31) s_warp[wid] = x;
        __syncthreads();
        if (wid == 0) {
            int w = s_warp[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += y; }
            s_warp[lane] = w;
        }
        __syncthreads();
        const int excl = s_carry + (wid ? s_warp[wid - 1] : 0) + x - len;
        if (a < seg_ants) offsets[(size_t)seg * seg_ants + a] = excl;
        __syncthreads();
        if (tid == 1023) s_carry += s_warp[31];
        __syncthreads();
    }
    if (tid == 0) totals[seg] = s_carry;
}

// one warp per local ant: cells -> move codes at packed[offsets[a] ...]
__global__ void __launch_bounds__(256) mpp_maaco_pack_moves_kernel(const int32_t *__restrict__ cells, int max_cells,
                                                                   const mpp_ant_result *__restrict__ res_local,
                                                                   const int32_t *__restrict__ offsets_local, int n_local,
                                                                   int C, uint8_t *__restrict__ packed, int cap,
                                                                   int32_t *__restrict__ status) {
    const int a = (blockIdx.x * 256 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (a >= n_local) return;
    const int n = res_local[a].n_cells;
    if (n <= 1) return;
    if (n > max_cells || offsets_local[a] + (n - 1) > cap) { if (lane == 0) atomicMax(status, 2); return; }
    const int32_t *p = cells + (size_t)a * max_cells;
    uint8_t *out = packed + offsets_local[a];
    for (int i = lane; i + 1 < n; i += 32) {
        const int d = p[i + 1] - p[i];                     // -C-1,-C,-C+1,-1,+1,C-1,C,C+1 -> move 0..7
        const int dr = (d + C + 1) / C - 1 + ((d + C + 1) < 0 ? -1 : 0);  // rows: d in [-C-1,-C+1] -> -1, [-1,1] -> 0, [C-1,C+1] -> 1
        const int dc = d - dr * C;
        const int i9 = (dr + 1) * 3 + (dc + 1);
        out[i] = (uint8_t)(i9 - (i9 > 4));
    }
}

// one warp per global ant: replay its moves 32 at a time (warp prefix sum of the cell deltas) and set the
// visited bits that fall into words [word0, word0 + n_words); visit_seg is [n_seg][n_words][seg_ants] and
// must be zero on entry.  Bits are set with fire-and-forget RED.OR (a column belongs to one ant).
__global__ void __launch_bounds__(256) mpp_maaco_rebuild_visits_kernel(const uint8_t *__restrict__ packed_all, int cap,
                                                                       const int32_t *__restrict__ offsets,
                                                                       const mpp_ant_result *__restrict__ res, int n_seg,
                                                                       int seg_ants, int start, int C, int word0,
                                                                       int n_words, uint32_t *__restrict__ visit_seg) {
    const int g = (blockIdx.x * 256 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (g >= n_seg * seg_ants) return;
    const int n = res[g].n_cells;
    if (n <= 0) return;                                      // failed ants deposit nothing (MAACO.py:307)
    const int seg = g / seg_ants, a = g - seg * seg_ants;
    const uint8_t *codes = packed_all + (size_t)seg * cap + offsets[g];
    uint32_t *col = visit_seg + (size_t)seg * n_words * seg_ants + a;     // word w of this ant at col[w * seg_ants]
    int base_cell = start;                                   // cell before the first move of the current chunk
    if (lane == 0) {
        const int w = (start >> 5) - word0;
        if (w >= 0 && w < n_words) atomicOr(&col[(size_t)w * seg_ants], 1u << (start & 31));
    }
    for (int i0 = 0; i0 < n - 1; i0 += 32) {
        const int i = i0 + lane;
        int d = 0;
        if (i < n - 1) {
            const int m = codes[i];
            d = ((int)((0xA940u >> (2 * m)) & 3u) - 1) * C + ((int)((0x9224u >> (2 * m)) & 3u) - 1);
        }
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, d, o); if (lane >= o) d += y; }
        const int cell = base_cell + d;                      // cell reached after move i
        if (i < n - 1) {
            const int w = (cell >> 5) - word0;
            if (w >= 0 && w < n_words) atomicOr(&col[(size_t)w * seg_ants], 1u << (cell & 31));
        }
        base_cell = __shfl_sync(0xffffffffu, cell, 31);
    }
}

extern "C" int mpp_maaco_move_offsets(const mpp_ant_result *result_dev, int n_seg, int seg_ants, int32_t *offsets_dev,
                                      int32_t *totals_dev, void *stream) {
    MPP_REQUIRE(result_dev && offsets_dev && totals_dev && n_seg > 0 && seg_ants > 0, "mpp_maaco_move_offsets: bad argument");
    mpp_maaco_move_offsets_kernel<<<n_seg, 1024, 0, (cudaStream_t)stream>>>(result_dev, seg_ants, offsets_dev, totals_dev);
    MPP_CUDA(cudaGetLastError());
    return MPP_OK;
}

extern "C" int mpp_maaco_pack_moves(const mpp_map *map, const int32_t *cells_dev, int max_cells,
                                    const mpp_ant_result *result_local_dev, const int32_t *offsets_local_dev, int n_local,
                                    uint8_t *packed_dev, int capacity, int32_t *status_dev, void *stream) {
    MPP_REQUIRE(map && cells_dev && result_local_dev && offsets_local_dev && packed_dev && status_dev && n_local > 0,
                "mpp_maaco_pack_moves: bad argument");
    MPP_CUDA(cudaSetDevice(map->device));
    mpp_maaco_pack_moves_kernel<<<(n_local + 7) / 8, 256, 0, (cudaStream_t)stream>>>(
        cells_dev, max_cells, result_local_dev, offsets_local_dev, n_local, map->cols, packed_dev, capacity, status_dev);
    MPP_CUDA(cudaGetLastError());
    return MPP_OK;
}

extern "C" int mpp_maaco_rebuild_visits(const mpp_map *map, const uint8_t *packed_all_dev, int capacity,
                                        const int32_t *offsets_dev, const mpp_ant_result *result_dev, int n_seg,
                                        int seg_ants, int word0, int n_words, uint32_t *visit_seg_dev, void *stream) {
    MPP_REQUIRE(map && packed_all_dev && offsets_dev && result_dev && visit_seg_dev && n_seg > 0 && seg_ants > 0,
                "mpp_maaco_rebuild_visits: bad argument");
    MPP_CUDA(cudaSetDevice(map->device));
    const int total = n_seg * seg_ants;
    mpp_maaco_rebuild_visits_kernel<<<(total + 7) / 8, 256, 0, (cudaStream_t)stream>>>(
        packed_all_dev, capacity, offsets_dev, result_dev, n_seg, seg_ants, map->start, map->cols, word0, n_words,
        visit_seg_dev);
    MPP_CUDA(cudaGetLastError());
    return MPP_OK;
}
