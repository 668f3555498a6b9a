// mpp_astar.cu -- batched A* connectors, waypoint-chain fitness (PSO/GA) and path statistics.
// Reference semantics: astar.py:33-101, MPA.py:106-151, helper.py:58-113, MPA.py:176-229,
// pso.py:56-94, ga_solver.py:58-93.
#include <cmath>
#include <vector>

#include "mpp_astar.cuh"
#include "mpp_stats.cuh"

#define MPP_AS_THREADS (MPP_GL == 32 ? 256 : 128)
#define MPP_AS_WARPS (MPP_AS_THREADS / 32)
#define MPP_AS_GROUPS (MPP_AS_THREADS / MPP_GL)     // searches (lane groups) per CTA

// ---------------------------------------------------------------------------------------------
// K1b: safety-class table (helper.py:67-80 as a per-cell function of the map)
//   class[cell] = min squared distance to an obstacle within the (2*radius+1)^2 stencil, radius = floor(msd); 0 = none.
//   lut[d2]     = (msd - sqrt(d2))**2 if sqrt(d2) < msd else 0   -- evaluated on the host with libm pow.
// Only obstacles nearer than msd contribute (helper.py:76), and those lie inside the stencil.
// Out-of-bounds is NOT an obstacle here (obstacle_nodes = argwhere(grid == 1), helper.py:125).
// ---------------------------------------------------------------------------------------------
__global__ void mpp_safety_kernel(const uint32_t *__restrict__ occ, int pitch, int R, int C, int radius,
                                  uint16_t *__restrict__ cls) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R * C) return;
    const int r = i / C, c = i % C;
    int best = 1 << 30;
    for (int dr = -radius; dr <= radius; ++dr) {
        const int rr = r + dr;
        if (rr < 0 || rr >= R) continue;
        for (int dc = -radius; dc <= radius; ++dc) {
            const int c2 = c + dc;
            if (c2 < 0 || c2 >= C) continue;
            const int pb = c2 + 1;
            if ((occ[(rr + 1) * pitch + (pb >> 5)] >> (pb & 31)) & 1u) {
                const int d2 = dr * dr + dc * dc;
                best = d2 < best ? d2 : best;
            }
        }
    }
    cls[i] = (best == (1 << 30)) ? 0 : (uint16_t)best;          // best <= 2*radius^2 < 65536 (radius <= 180)
}

extern "C" int mpp_map_safety_table(mpp_map *map, double msd, void *stream) {
    MPP_REQUIRE(map, "mpp_map_safety_table: null map");
    MPP_REQUIRE(msd >= 0.0 && msd <= 180.0, "mpp_map_safety_table: min_safe_distance %g unsupported (0 <= msd <= 180)", msd);
    if (map->safety_msd == msd && map->safety_d2_dev) return MPP_OK;
    MPP_CUDA(cudaSetDevice(map->device));
    const int n = map->rows * map->cols;
    const int radius = (int)msd;                               // an obstacle nearer than msd has |dr|, |dc| <= floor(msd)
    const int lut_n = 2 * radius * radius + 2;
    cudaStream_t s = (cudaStream_t)stream;
    if (!map->safety_d2_dev) MPP_CUDA(cudaMalloc(&map->safety_d2_dev, (size_t)n * sizeof(uint16_t)));
    if (map->safety_lut_n < lut_n) {
        MPP_CUDA(cudaStreamSynchronize(s));                     // a kernel of an earlier call may still read the old table
        if (map->safety_lut_dev) MPP_CUDA(cudaFree(map->safety_lut_dev));
        map->safety_lut_dev = nullptr;
        MPP_CUDA(cudaMalloc(&map->safety_lut_dev, (size_t)lut_n * sizeof(double)));
        map->safety_lut_n = lut_n;
    }
    std::vector<double> lut((size_t)lut_n, 0.0);
    for (int d2 = 1; d2 < lut_n; ++d2) {
        const double d = std::sqrt((double)d2);
        lut[d2] = (d < msd) ? std::pow(msd - d, 2.0) : 0.0;     // helper.py:77-78 (float ** 2 -> libm pow)
    }
    MPP_CUDA(cudaMemcpyAsync(map->safety_lut_dev, lut.data(), (size_t)lut_n * sizeof(double), cudaMemcpyHostToDevice, s));
    MPP_CUDA(cudaStreamSynchronize(s));
    mpp_safety_kernel<<<(n + 255) / 256, 256, 0, s>>>(map->occ_dev, map->pitch_words, map->rows, map->cols, radius,
                                                      map->safety_d2_dev);
    MPP_CUDA(cudaGetLastError());
    map->safety_msd = msd;
    return MPP_OK;
}

__global__ void __launch_bounds__(MPP_AS_THREADS)
mpp_path_stats_kernel(StatsCtx X, const int32_t *__restrict__ cells, int max_cells, const int32_t *__restrict__ n_cells,
                      int n_paths, double *stats) {
    const LaneGroup L = lane_group();
    const int w = (blockIdx.x * MPP_AS_THREADS + threadIdx.x) / MPP_GL;
    if (w >= n_paths) return;
    int n = n_cells[w];
    if (n > max_cells) n = max_cells;
    path_stats_warp(L, X, cells + (size_t)w * max_cells, n, stats + (size_t)w * 5);
}

static int make_stats_ctx(mpp_map *map, const mpp_policy *pol, void *stream, StatsCtx *X) {
    X->occ = map->occ_dev; X->pitch = map->pitch_words; X->R = map->rows; X->C = map->cols;
    X->pol = *pol;
    X->cls = nullptr; X->lut = nullptr;
    if (pol->mode == 0 && map->n_obstacles > 0) {
        int rc = mpp_map_safety_table(map, pol->min_safe_distance, stream);
        if (rc) return rc;
        X->cls = map->safety_d2_dev; X->lut = map->safety_lut_dev;
    }
    return MPP_OK;
}

extern "C" int mpp_path_stats(mpp_map *map, const int32_t *cells_dev, int max_cells, const int32_t *n_cells_dev,
                              int n_paths, const mpp_policy *policy, double *stats_dev, void *stream) {
    MPP_REQUIRE(map && cells_dev && n_cells_dev && policy && stats_dev, "mpp_path_stats: null argument");
    MPP_REQUIRE(n_paths > 0 && max_cells > 0, "mpp_path_stats: bad sizes");
    MPP_CUDA(cudaSetDevice(map->device));
    StatsCtx X;
    int rc = make_stats_ctx(map, policy, stream, &X);
    if (rc) return rc;
    const int blocks = (n_paths + MPP_AS_GROUPS - 1) / MPP_AS_GROUPS;
    mpp_path_stats_kernel<<<blocks, MPP_AS_THREADS, 0, (cudaStream_t)stream>>>(X, cells_dev, max_cells, n_cells_dev,
                                                                             n_paths, stats_dev);
    MPP_CUDA(cudaGetLastError());
    return MPP_OK;
}

// ---------------------------------------------------------------------------------------------
// scratch layout: [0,256) work counter + error flag; then n_slots search slots
// ---------------------------------------------------------------------------------------------
extern "C" size_t mpp_astar_scratch_bytes(const mpp_map *map, int n_slots, int heap_cap) {
    if (!map || n_slots <= 0 || heap_cap <= 0) return 0;
    return 256 + (size_t)n_slots * astar_slot_bytes(map->rows * map->cols, heap_cap);
}

extern "C" size_t mpp_astar_slot_bytes(int rows, int cols, int heap_cap) {
    if (rows <= 0 || cols <= 0 || heap_cap <= 0) return 0;
    return astar_slot_bytes(rows * cols, heap_cap);
}

// search slots (lane groups) that are resident at once: three CTAs of MPP_AS_GROUPS groups per SM
extern "C" int mpp_astar_max_slots(const mpp_map *map) { return map ? map->sm_count * 3 * MPP_AS_GROUPS : 0; }

struct BatchArgs {
    AStarGrid G;
    int occ_words;
    int variant;
    const int32_t *src, *dst;
    const uint32_t *avoid;  // n x words or null
    int n, words;
    int32_t *cells;
    int max_cells;
    int32_t *n_cells;
    double *g;
    char *scratch;
    int n_slots, heap_cap;
    unsigned long long *counters;
};

template <bool OCC_SMEM>
__global__ void __launch_bounds__(MPP_AS_THREADS, 3) mpp_astar_batch_kernel(BatchArgs A) {
    extern __shared__ __align__(16) uint32_t s_occ[];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ __align__(16) uint8_t s_cnt[MPP_AS_GROUPS][MPP_PQ_NB];
    AStarGrid G = A.G;
    if (OCC_SMEM) {
        mpp_stage_bulk(s_occ, A.G.occ, (uint32_t)A.occ_words * 4u, &s_bar);
        G.occ = s_occ;
    }
    const LaneGroup L = lane_group();
    const int lane = L.gl;
    const int slot = (blockIdx.x * MPP_AS_THREADS + threadIdx.x) / MPP_GL;
    if (slot >= A.n_slots) return;
    const int rc = G.R * G.C;
    AStarSlot S = astar_slot_at(A.scratch + 256 + (size_t)slot * astar_slot_bytes(rc, A.heap_cap), rc, A.heap_cap,
                                s_cnt[threadIdx.x / MPP_GL]);
    unsigned int *next = (unsigned int *)A.scratch;
    for (;;) {
        int i = 0;
        if (lane == 0) i = (int)atomicAdd(next, 1u);
        i = grp_shfl(L, i, 0);
        if (i >= A.n) break;
        double g;
        const int len = astar_search(L, G, S, A.variant, A.src[i], A.dst[i], A.avoid ? A.avoid + (size_t)i * A.words : nullptr,
                                     A.cells + (size_t)i * A.max_cells, A.max_cells, &g, A.counters);
        if (lane == 0) { A.n_cells[i] = len; if (A.g) A.g[i] = g; }
    }
}

static int check_scratch(const mpp_map *map, size_t scratch_bytes, int n_slots, int heap_cap) {
    // the searches pack a node as (row << 16 | col) in a signed int (same (r, c) order as the reference's tuples)
    MPP_REQUIRE(map->rows < 32768 && map->cols < 65536, "A*: map %dx%d exceeds the packed-node limit (rows < 32768, cols < 65536)",
                map->rows, map->cols);
    MPP_REQUIRE(n_slots > 0 && heap_cap >= 64, "A*: n_slots=%d heap_cap=%d", n_slots, heap_cap);
    MPP_REQUIRE(scratch_bytes >= mpp_astar_scratch_bytes(map, n_slots, heap_cap),
                "A*: scratch too small (%zu < %zu)", scratch_bytes, mpp_astar_scratch_bytes(map, n_slots, heap_cap));
    return MPP_OK;
}

extern "C" int mpp_astar_batch(mpp_map *map, int variant, const int32_t *src_dev, const int32_t *dst_dev,
                               const uint32_t *avoid_bits_dev, int n, int allow_diagonal, int restrict_corner,
                               int32_t *cells_dev, int max_cells, int32_t *n_cells_dev, double *g_dev,
                               void *scratch_dev, size_t scratch_bytes, int n_slots, int heap_cap,
                               unsigned long long *counters_dev, void *stream) {
    MPP_REQUIRE(map && src_dev && dst_dev && cells_dev && n_cells_dev && scratch_dev, "mpp_astar_batch: null argument");
    MPP_REQUIRE(variant >= 0 && variant <= 2, "mpp_astar_batch: variant must be 0 (astar.py), 1 (MPA.py) or 2 (dijkstra.py)");
    MPP_REQUIRE(n > 0 && max_cells > 0, "mpp_astar_batch: bad sizes");
    int rc = check_scratch(map, scratch_bytes, n_slots, heap_cap);
    if (rc) return rc;
    MPP_CUDA(cudaSetDevice(map->device));
    cudaStream_t s = (cudaStream_t)stream;
    MPP_CUDA(cudaMemsetAsync(scratch_dev, 0, 256, s));
    BatchArgs A;
    A.G.occ = map->occ_dev; A.G.pitch = map->pitch_words; A.G.R = map->rows; A.G.C = map->cols;
    A.G.allow_diag = allow_diagonal; A.G.restrict_corner = restrict_corner;
    A.occ_words = map->occ_words; A.variant = variant; A.src = src_dev; A.dst = dst_dev; A.avoid = avoid_bits_dev;
    A.n = n; A.words = (map->rows * map->cols + 31) / 32; A.cells = cells_dev; A.max_cells = max_cells;
    A.n_cells = n_cells_dev; A.g = g_dev; A.scratch = (char *)scratch_dev; A.n_slots = n_slots; A.heap_cap = heap_cap;
    A.counters = counters_dev;
    const size_t smem = (size_t)map->occ_words * 4;
    const int blocks = (n_slots + MPP_AS_GROUPS - 1) / MPP_AS_GROUPS;
    if (smem <= 32 * 1024) {
        mpp_astar_batch_kernel<true><<<blocks, MPP_AS_THREADS, smem, s>>>(A);
    } else if (smem <= 40 * 1024) {
        MPP_CUDA(cudaFuncSetAttribute(mpp_astar_batch_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        mpp_astar_batch_kernel<true><<<blocks, MPP_AS_THREADS, smem, s>>>(A);
    } else {
        mpp_astar_batch_kernel<false><<<blocks, MPP_AS_THREADS, 0, s>>>(A);
    }
    MPP_CUDA(cudaGetLastError());
    return MPP_OK;
}

// ---------------------------------------------------------------------------------------------
// K6 + K7: waypoint chain -> path -> fitness (pso.py:56-94 / ga_solver.py:58-93 + helper.py:98-113)
// One warp per individual: W+1 sequential connector searches with the growing avoid set
// (nodes_in_path_so_far - {current_start, waypoint}), then the statistics of the joined path.
// n_cells: 0 = invalid individual ([]), -1 = heap overflow, > max_cells = path truncated.
// ---------------------------------------------------------------------------------------------
struct ChainArgs {
    AStarGrid G;
    int occ_words;
    StatsCtx X;
    int start, target;
    const int32_t *wps;  // N x W cells
    int N, W, words;
    int32_t *cells;
    int max_cells;
    int32_t *n_cells;
    double *stats;
    uint32_t *visited;  // N x words
    char *scratch;
    int n_slots, heap_cap;
    unsigned long long *counters;
    const int32_t *order;  // optional processing order of the individuals (null: index order)
};

template <bool OCC_SMEM>
__global__ void __launch_bounds__(MPP_AS_THREADS, 3) mpp_waypoint_fitness_kernel(ChainArgs A) {
    extern __shared__ __align__(16) uint32_t s_occ[];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ __align__(16) uint8_t s_cnt[MPP_AS_GROUPS][MPP_PQ_NB];
    AStarGrid G = A.G;
    StatsCtx X = A.X;
    if (OCC_SMEM) {
        mpp_stage_bulk(s_occ, A.G.occ, (uint32_t)A.occ_words * 4u, &s_bar);
        G.occ = s_occ;
        X.occ = s_occ;
    }
    const LaneGroup L = lane_group();
    const int lane = L.gl;
    const int slot = (blockIdx.x * MPP_AS_THREADS + threadIdx.x) / MPP_GL;
    if (slot >= A.n_slots) return;
    const int rc = G.R * G.C;
    AStarSlot S = astar_slot_at(A.scratch + 256 + (size_t)slot * astar_slot_bytes(rc, A.heap_cap), rc, A.heap_cap,
                                s_cnt[threadIdx.x / MPP_GL]);
    unsigned int *next = (unsigned int *)A.scratch;
    for (;;) {
        int i = 0;
        if (lane == 0) i = (int)atomicAdd(next, 1u);
        i = grp_shfl(L, i, 0);
        if (i >= A.N) break;
        if (A.order) i = A.order[i];                                   // longest chains first (see mpp_waypoint_fitness)
        uint32_t *vis = A.visited + (size_t)i * A.words;
        int32_t *path = A.cells + (size_t)i * A.max_cells;
        for (int w = lane; w < A.words; w += MPP_GL) vis[w] = 0u;
        grp_sync(L);
        int n = 1, cur = A.start, status = 1;
        if (lane == 0) { path[0] = A.start; vis[A.start >> 5] |= 1u << (A.start & 31); }   // pso.py:63-65
        grp_sync(L);
        for (int k = 0; k <= A.W; ++k) {
            const int goal = (k < A.W) ? A.wps[(size_t)i * A.W + k] : A.target;
            // the segment is written over the tail cell of the path (segment[0] == current_start)
            const int cap = A.max_cells - (n - 1);
            const int sl = astar_search(L, G, S, 0, cur, goal, vis, path + (n - 1), cap, nullptr, A.counters);
            if (sl < 0) { status = -1; break; }
            if (sl == 0 || (sl == 1 && cur != goal)) { status = 0; break; }          // pso.py:77,87 -> []
            if (sl > cap) { status = 2; n += sl - 1; break; }                        // truncated
            for (int t = 1 + lane; t < sl; t += MPP_GL) {                            // nodes_in_path_so_far.update
                const int c = path[n - 1 + t];
                atomicOr(&vis[c >> 5], 1u << (c & 31));
            }
            grp_sync(L);
            n += sl - 1;
            cur = goal;
        }
        // (consecutive duplicates cannot occur: a segment never repeats its first cell; pso.py:91-93 is a no-op)
        int n_out = status == 1 ? n : (status == 2 ? n : status);
        if (lane == 0) A.n_cells[i] = n_out;
        path_stats_warp(L, X, path, status == 1 ? n : 0, A.stats + (size_t)i * 5);
    }
}

extern "C" int mpp_waypoint_fitness(mpp_map *map, const int32_t *waypoints_dev, int n_individuals, int n_waypoints,
                                    const mpp_policy *policy, int32_t *cells_dev, int max_cells, int32_t *n_cells_dev,
                                    double *stats_dev, uint32_t *visited_dev, void *scratch_dev, size_t scratch_bytes,
                                    int n_slots, int heap_cap, unsigned long long *counters_dev, const int32_t *order_dev,
                                    void *stream) {
    MPP_REQUIRE(map && policy && cells_dev && n_cells_dev && stats_dev && visited_dev && scratch_dev,
                "mpp_waypoint_fitness: null argument");
    MPP_REQUIRE(n_individuals > 0 && n_waypoints >= 0 && max_cells > 1, "mpp_waypoint_fitness: bad sizes");
    MPP_REQUIRE(n_waypoints == 0 || waypoints_dev, "mpp_waypoint_fitness: null waypoints");
    MPP_REQUIRE(map->start >= 0 && map->target >= 0, "mpp_waypoint_fitness: map has no start/target");
    int rc = check_scratch(map, scratch_bytes, n_slots, heap_cap);
    if (rc) return rc;
    MPP_CUDA(cudaSetDevice(map->device));
    cudaStream_t s = (cudaStream_t)stream;
    ChainArgs A;
    rc = make_stats_ctx(map, policy, stream, &A.X);
    if (rc) return rc;
    MPP_CUDA(cudaMemsetAsync(scratch_dev, 0, 256, s));
    A.G.occ = map->occ_dev; A.G.pitch = map->pitch_words; A.G.R = map->rows; A.G.C = map->cols;
    A.G.allow_diag = policy->allow_diagonal; A.G.restrict_corner = policy->restrict_policy;
    A.occ_words = map->occ_words; A.start = map->start; A.target = map->target;
    A.wps = waypoints_dev; A.N = n_individuals; A.W = n_waypoints; A.words = (map->rows * map->cols + 31) / 32;
    A.cells = cells_dev; A.max_cells = max_cells; A.n_cells = n_cells_dev; A.stats = stats_dev; A.visited = visited_dev;
    A.scratch = (char *)scratch_dev; A.n_slots = n_slots; A.heap_cap = heap_cap; A.counters = counters_dev;
    A.order = order_dev;
    const size_t smem = (size_t)map->occ_words * 4;
    const int blocks = (n_slots + MPP_AS_GROUPS - 1) / MPP_AS_GROUPS;
    if (smem <= 32 * 1024) {
        mpp_waypoint_fitness_kernel<true><<<blocks, MPP_AS_THREADS, smem, s>>>(A);
    } else if (smem <= 40 * 1024) {
        MPP_CUDA(cudaFuncSetAttribute(mpp_waypoint_fitness_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        mpp_waypoint_fitness_kernel<true><<<blocks, MPP_AS_THREADS, smem, s>>>(A);
    } else {
        mpp_waypoint_fitness_kernel<false><<<blocks, MPP_AS_THREADS, 0, s>>>(A);
    }
    MPP_CUDA(cudaGetLastError());
    return MPP_OK;
}
