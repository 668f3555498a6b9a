/*
 * mpp.h -- C ABI of libmpp_b200.so: the B200 (sm_100a) implementation of the
 * population-evaluation hot path of dvnam1605/MAACO-path-planing.
 *
 * The reference has no FFI of its own (pure Python); this header is the boundary a
 * maintainer binds with ctypes underneath the reference's Python classes (see
 * INTEGRATION.md for the stub).  Each entry point names the reference code it
 * replaces (file:line relative to the reference tree).
 *
 * Conventions
 *   - every function returns 0 on success, a negative MPP_E* code otherwise; nothing
 *     throws across the ABI; mpp_last_error() returns a thread-local message.
 *   - all pointers named *_dev are device pointers on the map's device; everything is
 *     enqueued on `stream` (a cudaStream_t passed as void*) and is asynchronous unless
 *     stated.  The caller owns every buffer; the library owns only map handles.
 *   - cells are int32 `r*cols + c`; paths are int32 arrays; all reals are IEEE binary64.
 *   - there is NO CPU fallback: with no sm_100 device the calls fail with MPP_ENODEVICE.
 *   - RNG: counter-based Philox4x32-10 streams keyed (seed, class, iteration, individual)
 *     with an in-stream cursor (DESIGN.md "RNG contract").
 */
#ifndef MPP_H_
#define MPP_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPP_ABI_VERSION 1

#define MPP_OK 0
#define MPP_EINVAL (-1)
#define MPP_ENODEVICE (-2)
#define MPP_ECUDA (-3)
#define MPP_ENOMEM (-4)
#define MPP_EUNSUPPORTED (-5)

/* stream classes of the RNG contract */
#define MPP_CLS_MAACO_TOUR 1
#define MPP_CLS_PSO_INIT 2
#define MPP_CLS_PSO_PAD 3
#define MPP_CLS_PSO_UPDATE 4
#define MPP_CLS_GA_INIT 5
#define MPP_CLS_GA_PAD 6
#define MPP_CLS_GA_SELECT 7
#define MPP_CLS_GA_BREED 8
#define MPP_CLS_MPA_PHASE 9
#define MPP_CLS_MPA_FADS 10

typedef struct mpp_map mpp_map; /* opaque: bit-packed occupancy grid + per-map tables on one device */

int mpp_abi_version(void);
const char *mpp_last_error(void);
/* number of sm_100 devices visible (0 => every compute entry point fails) */
int mpp_device_count(void);

/* ---- grid map: env.py grid convention (0 free, 1 obstacle, 2 start, 3 target) ------------
 * replaces: np.array(grid) + argwhere(grid==2/3) in every constructor
 * (MAACO.py:15-41, helper.py:121-125, astar.py:17-22, pso.py:17-22, ga_solver.py:17-22, MPA.py:20-43).
 * grid_host: rows*cols bytes, row-major.  Start/target = first row-major 2 / 3.  Synchronous. */
int mpp_map_create(const uint8_t *grid_host, int rows, int cols, int device, mpp_map **out);
void mpp_map_destroy(mpp_map *map);
int mpp_map_rows(const mpp_map *map);
int mpp_map_cols(const mpp_map *map);
int mpp_map_start(const mpp_map *map);  /* cell id or -1 */
int mpp_map_target(const mpp_map *map); /* cell id or -1 */
int mpp_map_device(const mpp_map *map);
/* device pointer to the border-padded bit-packed occupancy grid ((rows+2) x pitch_words uint32,
 * bit (r+1, c+1); the border is marked occupied) and its row pitch in 32-bit words */
const uint32_t *mpp_map_occ_bits(const mpp_map *map, int *pitch_words);

/* ---- MAACO (MAACO.py) -------------------------------------------------------------------- */
typedef struct {
    double alpha, beta, rho, Q, a_turn_coef, wh_max, wh_min, k_h_adaptive, q0_initial, C0_initial_pheromone;
    int num_iterations;
} mpp_maaco_params;

/* replaces MAACO._initialize_pheromones_maaco (MAACO.py:58-84), _precompute_dist_to_target
 * (:86-91) and the per-candidate heuristic eta'**beta (:197-210, :238) folded into two per-cell
 * tables E[cell][c], c = turn flag (interleaved: E01_dev[2*cell + c], 2*rows*cols doubles).  exp/pow
 * are evaluated on the host with libm (the same functions CPython/NumPy call) so the tables are
 * bit-identical to the reference; tau0_dev / dist_t_dev are rows*cols doubles (dist_t_dev may be
 * NULL).  Synchronous. */
int mpp_maaco_tables(const mpp_map *map, const mpp_maaco_params *p, double *tau0_dev, double *E01_dev,
                     double *dist_t_dev, void *stream);

/* replaces MAACO._calculate_adaptive_q0 (MAACO.py:212-226); pure host function */
double mpp_maaco_q0(int num_iterations, int iteration, double q0_initial);

/* ranks, per cell and previous-move context, the 8 moves by attractiveness tau**alpha * eta'**beta
 * (MAACO.py:235-239).  Where every value is < 1e-10 the selection rules (:241-262) depend only on that order
 * and the tour kernels use the ranking; elsewhere the entry says "evaluate".  Must be re-run after every
 * pheromone update.  rank_dev: mpp_maaco_rank_words(map) uint32 (opaque: a per-(cell, context) word for the
 * moves of the start->target quadrant, padded for unchecked neighbour reads, and a full two-word entry). */
int mpp_maaco_rank(const mpp_map *map, const double *tau_dev, const double *E01_dev, double alpha, uint32_t *rank_dev,
                   void *stream);
long long mpp_maaco_rank_words(const mpp_map *map);

/* per-ant outcome of one tour (MAACO.py:300-302): failed ants have n_cells=0, length=+inf, turns=-1 */
typedef struct {
    double length;
    int32_t n_cells;
    int32_t turns;
} mpp_ant_result;

/* replaces the ant loop MAACO.py:340-342 -> _construct_ant_solution_maaco (:278-302) with the
 * orientation filter + crossing-prohibition (:100-181) and pseudo-random-proportional selection
 * (:228-262).  Ant i of this call is global ant `ant_offset + i` (its RNG stream id).
 *   tau / E01      rows*cols / 2*rows*cols doubles (mpp_maaco_tables)
 *   rank_dev       optional mpp_maaco_rank_words(map) words from mpp_maaco_rank for the CURRENT tau (NULL = evaluate the
 *                  attractiveness of every candidate at every step)
 *   visitT_dev     word-major visited bitmaps: word w of ant i at [w*n_ants + i], ceil(rows*cols/32)
 *                  words per ant; MUST be zero on entry; holds each ant's visited set on return
 *   cells_dev      n_ants x max_cells path cells (cells beyond max_cells are dropped; n_cells still
 *                  counts them)
 *   result_dev     n_ants x mpp_ant_result (:288,:292,:300-302)
 *   steps_dev      optional counter, += number of ant steps taken
 *   lanes_per_ant  0 = library default; 1 = one thread per ant (needs rank_dev), ants per warp chosen from n_ants;
 *                  -k (k = 1,2,4,..,32) = one thread per ant, k ants per warp (for callers that overlap several
 *                  colonies and know the total load); 8, 16 or 32 = that many lanes cooperate on one ant              */
int mpp_maaco_tours(const mpp_map *map, const double *tau_dev, const double *E01_dev, const uint32_t *rank_dev,
                    int iteration, double q0, double alpha, int n_ants, int ant_offset, uint64_t seed,
                    uint32_t *visitT_dev, int32_t *cells_dev, int max_cells, mpp_ant_result *result_dev,
                    unsigned long long *steps_dev, int lanes_per_ant, void *stream);

/* colony state kept on the device so an entire solve can be enqueued without host syncs */
typedef struct {
    double best_len;       /* best_path_length_overall (+inf = none) */
    int32_t best_turns;    /* -1 = inf */
    int32_t best_n_cells;
    int32_t best_iter;     /* iteration that set best_path (0 = none) */
    int32_t best_ant;
    double iter_best_len;  /* last iteration */
    int32_t iter_best_turns;
    int32_t iter_best_ant; /* -1 = all ants failed */
} mpp_maaco_state;

/* replaces the order-dependent best tracking MAACO.py:343-358 for one iteration over all
 * n_ants results (in global ant order) and prepares the per-ant deposit Q/length (:307-308; 0 for
 * ants that do not deposit).  Updates *state_dev and appends to the per-iteration log
 * log_dev[4*(iteration-1) + {0: iter best len, 1: iter best turns, 2: overall len, 3: overall turns}]
 * (log_dev may be NULL).  cells_dev holds the paths of ants [cells_ant_offset, cells_ant_offset +
 * cells_n_ants) (a sharded colony keeps only its own); the new overall-best path is copied into
 * best_cells_dev (capacity max_cells) when the best ant lies in that range. */
int mpp_maaco_best(const mpp_ant_result *result_dev, const int32_t *cells_dev, int max_cells, int cells_ant_offset,
                   int cells_n_ants, int n_ants, double Q, int iteration, mpp_maaco_state *state_dev,
                   int32_t *best_cells_dev, double *deposit_dev, double *log_dev, void *stream);

/* replaces MAACO._update_pheromone_trails_maaco (MAACO.py:304-332) for the cells of bitmap words
 * [word0, word0 + n_words): evaporate, deposit in global ant order (bit-exact, atomics-free: one warp
 * per 32 cells streams the visited words of every ant), MMAS clip with tau_max from state->best_len,
 * obstacles <- 1e-9.  visitT_dev is [n_seg][n_words][seg_ants] (word-major per segment; global ant =
 * seg*seg_ants + a; one segment per source rank in a sharded colony; n_seg=1, word0=0,
 * n_words=ceil(rows*cols/32) on a single GPU); deposit_dev is indexed by global ant.  Clears
 * visitT_dev behind itself when clear_visit != 0. */
int mpp_maaco_pheromone(const mpp_map *map, double *tau_dev, uint32_t *visitT_dev, const double *deposit_dev,
                        int n_seg, int seg_ants, int word0, int n_words, double rho,
                        const mpp_maaco_state *state_dev, int clear_visit, void *stream);

/* ---- sharded-colony exchange (no reference counterpart: the reference is single-process; these implement the
 * per-iteration exchange BASELINE's north star asks for, bit-exactly) ----------------------------------------
 * Tours travel between GPUs as 1-byte move codes (MAACO move order, MAACO.py:98).
 * mpp_maaco_move_offsets: from the all-gathered results ([n_seg*seg_ants], segment = source rank) compute each
 *   ant's byte offset inside its segment's packed buffer and the per-segment totals (n_cells-1 bytes per
 *   successful ant, 0 for failed ants).
 * mpp_maaco_pack_moves: encode this rank's paths (cells_dev, n_local x max_cells) at offsets_local_dev.
 *   status_dev[0] = 2 if a path was truncated at max_cells or does not fit `capacity`.
 * mpp_maaco_rebuild_visits: replay the codes of every ant (packed_all_dev = n_seg buffers of `capacity` bytes)
 *   and set the visited bits falling into words [word0, word0+n_words) of visit_seg_dev
 *   ([n_seg][n_words][seg_ants], zero on entry) -- the input layout of mpp_maaco_pheromone. */
int mpp_maaco_move_offsets(const mpp_ant_result *result_dev, int n_seg, int seg_ants, int32_t *offsets_dev,
                           int32_t *totals_dev, void *stream);
int mpp_maaco_pack_moves(const mpp_map *map, const int32_t *cells_dev, int max_cells,
                         const mpp_ant_result *result_local_dev, const int32_t *offsets_local_dev, int n_local,
                         uint8_t *packed_dev, int capacity, int32_t *status_dev, void *stream);
int mpp_maaco_rebuild_visits(const mpp_map *map, const uint8_t *packed_all_dev, int capacity,
                             const int32_t *offsets_dev, const mpp_ant_result *result_dev, int n_seg, int seg_ants,
                             int word0, int n_words, uint32_t *visit_seg_dev, void *stream);

/* ---- A* connectors, waypoint-chain fitness, path statistics (astar.py, MPA.py, helper.py, pso.py, ga_solver.py) ---- */
typedef struct {
    double turn_penalty_factor, safety_penalty_factor, min_safe_distance, diagonal_obstacle_penalty_value;
    int restrict_policy;  /* restrict_diagonal_near_obstacle_policy: connector bans corner cutting; stats charge it */
    int allow_diagonal;   /* allow_diagonal_moves */
    int mode;             /* 0 = helper.calculate_path_stats, 1 = MPA._calculate_path_stats (safety hard-wired 0.0) */
} mpp_policy;

/* replaces calculate_path_safety_penalty's O(len x n_obstacles) scan (helper.py:67-80) by a per-cell class
 * table (u16: min squared distance to an obstacle within floor(msd)) + a LUT of (msd - d)**2 per class, evaluated
 * with libm on the host.  0 <= msd <= 180 (u16 classes).  Cached in the map for one msd at a time; called
 * implicitly by the stats entry points. */
int mpp_map_safety_table(mpp_map *map, double min_safe_distance, void *stream);

/* replaces helper.calculate_path_stats (helper.py:98-113; count_turns :58-65, safety :67-80, diagonal
 * penalty :82-96) and MPA._calculate_path_stats (MPA.py:215-229) for n_paths paths, one warp each.
 * length uses CPython 3.12's Neumaier-compensated sum().  stats_dev: n_paths x 5 doubles =
 * (length, turns, safety_penalty, diag_penalty, fitness); empty path -> (inf, 0, 0, 0, inf). */
int mpp_path_stats(mpp_map *map, const int32_t *cells_dev, int max_cells, const int32_t *n_cells_dev, int n_paths,
                   const mpp_policy *policy, double *stats_dev, void *stream);

/* scratch sizing for the search entry points: n_slots concurrent searches (one warp each; mpp_astar_max_slots
 * = the number that are resident at once), each with a heap of heap_cap entries.  The scratch buffer must be
 * zero-filled once before its first use (search stamps live in it). */
size_t mpp_astar_scratch_bytes(const mpp_map *map, int n_slots, int heap_cap);
int mpp_astar_max_slots(const mpp_map *map);

/* replaces n independent calls of AStarSolver.solve (variant 0, astar.py:33-101, with get_valid_neighbors
 * helper.py:18-53), MPA._a_star (variant 1, MPA.py:106-151) or DijkstraSolver.solve (variant 2,
 * dijkstra.py:32-97 = variant 0 with a zero heuristic).  avoid_bits_dev: optional n x ceil(rows*cols/32)
 * bitmaps (nodes_to_avoid).  cells_dev: n x max_cells; n_cells_dev[i] = path length (0 = no path / invalid
 * endpoint, -1 = heap_cap overflow, > max_cells = truncated); g_dev[i] = g of the popped target (nullable).
 * counters_dev: optional [4] = (node expansions, successful relaxations, ring-bucket pushes, overflow-heap
 * pushes), accumulated. */
int mpp_astar_batch(mpp_map *map, int variant, const int32_t *src_dev, const int32_t *dst_dev,
                    const uint32_t *avoid_bits_dev, int n, int allow_diagonal, int restrict_corner,
                    int32_t *cells_dev, int max_cells, int32_t *n_cells_dev, double *g_dev, void *scratch_dev,
                    size_t scratch_bytes, int n_slots, int heap_cap, unsigned long long *counters_dev, void *stream);

/* replaces PSOSolver._reconstruct_path_from_position (pso.py:56-94, after rounding/clamping) /
 * GASolver._reconstruct_path_from_chromosome (ga_solver.py:58-93) followed by
 * BasePathfinder._calculate_stats_for_path (helper.py:138-147) for a whole population: one warp per
 * individual runs its W+1 connector searches with the growing avoid set and the statistics of the joined
 * path.  waypoints_dev: N x W cells.  visited_dev: N x ceil(rows*cols/32) words of scratch.
 * n_cells_dev[i]: 0 = invalid individual ([]), -1 = heap overflow, > max_cells = truncated. */
int mpp_waypoint_fitness(mpp_map *map, const int32_t *waypoints_dev, int n_individuals, int n_waypoints,
                         const mpp_policy *policy, int32_t *cells_dev, int max_cells, int32_t *n_cells_dev,
                         double *stats_dev, uint32_t *visited_dev, void *scratch_dev, size_t scratch_bytes,
                         int n_slots, int heap_cap, unsigned long long *counters_dev, void *stream);

/* ---- PSO / GA population updates (pso.py, ga_solver.py) ------------------------------------------- */

/* replaces the velocity/position update of the particle loop (pso.py:183-206) for particles
 * [particle_offset, particle_offset + n_particles) (device arrays start at that particle; a suffix is
 * re-run when an earlier particle improved gbest -- the reference's loop is asynchronous, pso.py:222-229).
 * pos/vel/pbest: n x W x 2 doubles (row, col); gbest: W x 2.  Exactly 4W uniforms per particle, stream
 * (seed, PSO_UPDATE, iteration, particle).  Also writes the integer waypoints pso.py:61,69-70
 * (round-half-even, clamped) as cells into waypoint_cells_dev (n x W). */
int mpp_pso_update(const mpp_map *map, double *pos_dev, double *vel_dev, const double *pbest_pos_dev,
                   const double *gbest_pos_dev, int n_particles, int particle_offset, int n_waypoints, double w,
                   double c1, double c2, double max_vel, uint64_t seed, int iteration, int32_t *waypoint_cells_dev,
                   void *stream);
/* pso.py:61,69-70 alone (used for the initial particles) */
int mpp_pso_round(const mpp_map *map, const double *pos_dev, int n_particles, int n_waypoints,
                  int32_t *waypoint_cells_dev, void *stream);

/* replaces GASolver._selection (ga_solver.py:136-142): n tournaments of min(tournament_size, n) individuals
 * drawn with CPython's random.sample algorithm (randbelow(n) = floor(u*n)); winner = first minimum in sample
 * order.  fitness_dev is in population (sorted) order; parents_dev receives n population indices.
 * Stream (seed, GA_SELECT, generation, tournament). */
int mpp_ga_select(const double *fitness_dev, int n, int tournament_size, uint64_t seed, int generation,
                  int32_t *parents_dev, void *stream);

/* replaces the breeding loop ga_solver.py:186-194 (_crossover :144-152, _mutate :154-160,
 * _generate_random_waypoint :48-53): pair p = (parents[2p % n], parents[(2p+1) % n]) -> children 2p, 2p+1.
 * chrom_dev: n x W cells in population order; children_dev: n x W.  Stream (seed, GA_BREED, generation, pair). */
int mpp_ga_breed(const mpp_map *map, const int32_t *chrom_dev, const int32_t *parents_dev, int n, int n_waypoints,
                 double crossover_rate, double mutation_rate, uint64_t seed, int generation, int32_t *children_dev,
                 void *stream);

/* ---- MPA (MPA.py) ------------------------------------------------------------------------------- */

/* replaces one iteration of MPA.solve_path_planning for the whole population (MPA.py:339-410): the
 * phase loop (:340-377: start index + gate draw, Levy :250-264 / Brownian :266-282 target,
 * _reconstruct_path_segment :284-318 with the private A* :106-151 and stats :215-229), the
 * marine-memory step (:380-384) and the FADs step (:387-410).  The old population (cells/n_cells/stats,
 * N x max_cells) must be sorted by fitness (row 0 = elite, :333-334); the next population is written
 * unsorted to out_*.  phase = 1/2/3 (:339,:348,:365), CF as computed at :336, levy_sigma as at :251-253.
 * tmp_cells_dev: n_slots x max_cells, avoid_dev: n_slots x ceil(rows*cols/32) words of scratch.
 * Only predators [pred_begin, pred_end) are processed (a rank's shard of the population; their rows of out_* are
 * written).  status_dev[0] (zero it first): 1 = heap overflow, 2 = a path exceeded max_cells (repeat with larger
 * buffers).  Streams (seed, MPA_PHASE, iteration, i) and (seed, MPA_FADS, iteration, i). */
int mpp_mpa_iteration(mpp_map *map, const mpp_policy *policy, int n_predators, int pred_begin, int pred_end,
                      int iteration, int phase,
                      double P_const, double CF, double FADs_rate, double levy_sigma, double levy_beta, uint64_t seed,
                      const int32_t *cells_dev, const int32_t *n_cells_dev, const double *stats_dev, int max_cells,
                      int32_t *out_cells_dev, int32_t *out_n_dev, double *out_stats_dev, int32_t *tmp_cells_dev,
                      uint32_t *avoid_dev, void *scratch_dev, size_t scratch_bytes, int n_slots, int heap_cap,
                      int32_t *status_dev, unsigned long long *counters_dev, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* MPP_H_ */
