/*
 * libs4mpi -- C-ABI of the B200-native Streets4MPI traffic-assignment hot path.
 *
 * The reference (jfietkau/Streets4MPI) has no FFI layer: its hot path is the
 * Python method surface of Simulation / StreetNetwork plus one MPI collective.
 * Every entry point below names the reference lines it replaces (paths relative
 * to the reference checkout).  The host side (the streets4mpi_b200 package) keeps the
 * reference's class and method names and forwards to these through ctypes; the
 * stub a reference maintainer would add is shown in INTEGRATION.md.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / CUDA types in signatures (a CUDA
 *     stream crosses as void*).
 *   - every function returns S4_OK (0) or a negative S4_ERR_*; the message of
 *     the last failure on the calling thread is s4_last_error().  No exceptions,
 *     no exit().
 *   - node ids are dense int32 0..V-1, assigned by the host in ascending order
 *     of the original ids (so id-based tie-breaks commute with the mapping);
 *     street ids are the reference's street_index 0..S-1 (streetnetwork.py:40,57).
 *   - host arrays are borrowed for the duration of the call only.
 *   - one host thread drives one s4_ctx; work is stream-ordered on the ctx
 *     stream and asynchronous unless a function hands results back to the host.
 *   - there is NO CPU fallback: without a CUDA device every call fails with
 *     S4_ERR_CUDA.
 */
#ifndef S4MPI_H
#define S4MPI_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define S4_ABI_VERSION 1

#define S4_OK 0
#define S4_ERR_ARG (-1)    /* bad argument (null pointer, id out of range, ...) */
#define S4_ERR_CUDA (-2)   /* CUDA runtime error, or no device */
#define S4_ERR_OOM (-3)    /* host or device allocation failed */
#define S4_ERR_STATE (-4)  /* call not valid in the current state */

typedef struct s4_ctx s4_ctx;     /* one GPU + one stream                                   */
typedef struct s4_graph s4_graph; /* replicated street graph: symmetric CSR in HBM          */
typedef struct s4_sim s4_sim;     /* one (virtual) MPI rank: trips, jam tolerance, loads    */

/* settings.py:36-44 -- the knobs the hot path reads */
typedef struct s4_physics {
    double car_length;            /* settings["car_length"]            (simulation.py:132) */
    double min_breaking_distance; /* settings["min_breaking_distance"] (simulation.py:132) */
    double braking_deceleration;  /* settings["braking_deceleration"]  (simulation.py:134) */
    uint32_t trip_volume;         /* settings["trip_volume"]           (simulation.py:86)  */
    uint32_t reserved;
} s4_physics;

/* which endpoint of a trip roots its shortest-path tree */
#define S4_ROOT_ORIGIN 0 /* reference grouping: one tree per distinct origin (simulation.py:71) */
#define S4_ROOT_GOAL 1   /* tree per distinct goal; identical loads where paths are unique       */
#define S4_ROOT_AUTO 2   /* whichever endpoint set is smaller (the graph is undirected)          */

/* device-side work and time accounting since the last s4_sim_reset_counters */
typedef struct s4_counters {
    uint64_t days;              /* s4_sim_step calls                                              */
    uint64_t sssp_instances;    /* shortest-path trees computed                                   */
    uint64_t arcs_algorithmic;  /* sum over trees of arcs out of reachable nodes (each arc once)  */
    uint64_t nodes_algorithmic; /* sum over trees of reachable nodes                              */
    uint64_t arcs_relaxed;      /* arcs actually scanned, re-relaxations included                 */
    uint64_t nodes_popped;      /* frontier entries expanded                                      */
    uint64_t stale_pops;        /* frontier entries skipped as superseded                         */
    uint64_t queue_pushes;      /* near + far insertions                                          */
    uint64_t bucket_advances;   /* far-pile splits                                                */
    uint64_t relax_rounds;      /* CTA-wide relaxation rounds                                     */
    uint64_t fallback_sweeps;   /* Bellman-Ford sweeps after a queue overflow (normally 0)        */
    uint64_t trips_routed;      /* trips handed to the walk (reachable or not)                    */
    uint64_t trips_unreachable; /* goal not in the tree (simulation.py:80)                        */
    uint64_t trips_stuck;       /* no predecessor found (zero-weight plateau); normally 0         */
    uint64_t hops;              /* street traversals == sum of load increments / trip_volume      */
    uint64_t kernel_launches;   /* kernels this library launched                                  */
    uint64_t sssp_launches;     /* launches of the SSSP kernel                                    */
    double ms_weights;          /* K1 (CUDA events, summed)                                       */
    double ms_sssp;             /* K2                                                             */
    double ms_walk;             /* K3                                                             */
    double ms_adapt;            /* K5                                                             */
    double ms_step_total;       /* whole s4_sim_step, first kernel to last                        */
} s4_counters;

int s4_abi_version(void);
const char *s4_last_error(void);

/* ---- context ------------------------------------------------------------ */
/* Replaces the per-process setup of streets4mpi.py:44-46.  `stream` is a
 * cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream) or NULL to let
 * the library create its own non-blocking stream. */
int s4_ctx_create(int device, void *stream, s4_ctx **out);
int s4_ctx_destroy(s4_ctx *ctx);
int s4_ctx_sync(s4_ctx *ctx);
/* Tunables; 0 means "chosen by the library" where noted.
 *   "delta_factor"       near-bucket width in typical (median) street weights (default 1.5)
 *   "adaptive_delta"     1 (default): the width follows the wave (grows on thin waves, shrinks on floods)
 *   "block_threads"      0 (default: from the graph's wave width) | 32 | 64 | 128 | 256 | 512 | 1024
 *   "ctas_per_sm"        0 (default: 1024 threads per SM)
 *   "near_smem_entries"  0 (default: 6 per thread) - shared-memory part of each near queue
 *   "queue_entries"      0 (default: from V, within 15 % of the free device memory) - per-CTA spill / far-pile capacity
 *   "pool_bytes"         cap on the bytes of per-tree distance arrays (default 0: 55 % of the free device memory)
 *   "batch_roots"        0 (default: from pool_bytes) - trees in flight
 *   "root_mode"          S4_ROOT_* used by the next s4_sim_create
 *   "walk_tile"          0 (default: 8) | 4 | 8 | 16 | 32 lanes per walking trip
 *   "preread"            2 (default): the filtering read decides the push, the minimum goes out as a
 *                        fire-and-forget reduction; 1: filtering read, then returning atomicMin
 *   "early_advance"      2 (default): the next bucket is admitted as soon as a relaxation round (not one of the first
 *                        two of its bucket) holds fewer entries than 2/8 of a CTA - the thinning tail of a bucket
 *                        overlaps the head of the next; 0: only when the near queue is empty
 *   "walk_ell"           1 (default): the walk reads fixed-stride rows {head, street_index, weight} (one 64-byte
 *                        read per hop, no row_ptr lookup; nodes with more than four arcs use the CSR); 0: CSR only.
 *                        Read at s4_sim_create.
 *   "lanes"              2 (default): a day's batches alternate between two concurrent launch lanes;
 *                        1: strictly serial kernels (per-kernel timing)
 *   "group_lanes"        8 (default) | 4: lanes that expand one node for a whole group of trees on one shared
 *                        frontier (source-batched SSSP); 0: one tree per CTA (the round-1 kernel)
 *   "group_words"        0 (default: 4 when twice as many slots as SMs fit in the pool, else 2) | 1 | 2 | 4: distance
 *                        words per lane; trees per group = group_lanes * group_words - 1 (7 | 15 | 31)
 *   "lag_factor"         1 (default): a node is held back by this multiple of the root-to-root distance of its group
 *   "min_batches"        4 (default): a day is cut into at least this many (SSSP launch -> walk) batches
 *   "balance_waves"      1 (default): fewer trees per group when that fills the last wave of a launch
 *   "cluster"            0 (default: from the number of slots that fit in the pool) | 1 | 2 | 4 | 8: CTAs of a
 *                        thread-block cluster that share one group (graphs with 10^7 .. 10^8 nodes: fewer slots than SMs)
 *   "walk_codes"         2 (default): when the day's walks are long against a pass over every (node, tree) -- hops of the
 *                        previous day (first day: trips x half the graph's hop depth) >= 0.1 * V * trees -- a streaming pass writes the predecessor of every
 *                        (node, tree) as a 2-bit arc-slot code into the rows' meta words and the walks take one lane and
 *                        one dependent access per hop; 1: always; 0: never (predecessors derived on the fly)
 *   "l2_persist"         1: access-policy window that keeps the arc rows in L2 (measured: no gain; default 0)
 *   "group_log"          1: keep per-group cycles / rounds / expansions of the last day (s4_graph_group_log)
 *   "delta_factor" counts MEDIAN street weights (the mean follows the 1 km/h streets road construction leaves behind).
 * No option changes a result. */
int s4_ctx_set_option(s4_ctx *ctx, const char *key, double value);
int s4_ctx_get_option(s4_ctx *ctx, const char *key, double *value);
/* name, SM count, bytes of HBM; name_out may be NULL */
int s4_ctx_device_info(s4_ctx *ctx, char *name_out, int name_cap, int *sm_count, uint64_t *hbm_bytes);

/* ---- rank-level exchange (streets4mpi.py:44-46, :97-98) --------------------------------------------------
 * One s4_ctx = one MPI rank of the reference.  The collective is NCCL's all-reduce over NVLink, bound at run time
 * to the libnccl the process already holds (torch's under Python; else $S4MPI_NCCL_LIB or libnccl.so.2), so a
 * non-Python host gets the exchange from this library alone:
 *     rank 0:      s4_comm_unique_id(id)   -> ship the 128 bytes to the other ranks (any transport)
 *     every rank:  s4_comm_init(ctx, nranks, rank, id)              (mpi4py: COMM_WORLD, Get_rank, Get_size)
 *     every day:   s4_sim_step(sim); s4_sim_exchange(sim);          (communicator.Allreduce(..., op=MPI.SUM))
 * nranks == 1 needs no id and no NCCL (the reference started without mpiexec, README.md:73-80). */
int s4_comm_unique_id(void *id_out_128_bytes);
int s4_comm_init(s4_ctx *ctx, int nranks, int rank, const void *unique_id_128_bytes);
int s4_comm_free(s4_ctx *ctx);
int s4_comm_size(s4_ctx *ctx, int *nranks, int *rank);
/* communicator.Allreduce(local, total, op=MPI.SUM) for a driver that keeps `traffic_load` in a host array
 * (streets4mpi.py:97-98 literally): in place on inout[n], through a pinned staging buffer. */
int s4_comm_allreduce_host(s4_ctx *ctx, uint32_t *inout, int64_t n);

/* ---- graph (streetnetwork.py:35-57, 113-125) ------------------------------ */
/* Street i joins u[i] and v[i] (either orientation, the graph is undirected:
 * streetnetwork.py:24,37), has length[i] metres and speed limit max_speed[i]
 * km/h (> 0).  Builds the symmetric CSR (2 arcs per street, arcs of a node in
 * street_index order) and uploads it.  node_rank (may be NULL) is a permutation
 * of 0..V-1 giving every node its position in the device's internal numbering
 * (the host passes the order along a space-filling curve over the node
 * coordinates: neighbours then share cache sectors of the per-tree distance
 * arrays).  It changes no result: all ids crossing this ABI and the lowest-id
 * tie-break stay in the caller's numbering. */
int s4_graph_upload(s4_ctx *ctx, int32_t V, int64_t S, const int32_t *u, const int32_t *v,
                    const double *length, const int32_t *max_speed, const int32_t *node_rank, s4_graph **out);
int s4_graph_free(s4_graph *g);
int s4_graph_dims(s4_graph *g, int32_t *V, int64_t *S, int64_t *A);
/* Diagnostics (option "group_log" = 1): 4 words per group of trees of the last s4_sim_step on this graph --
 * {SM cycles >> 10, relaxation rounds, frontier entries popped, bucket advances}.  out may be NULL to size the buffer. */
int s4_graph_group_log(s4_graph *g, uint32_t *out, int64_t cap_words, int64_t *n_words);
/* StreetNetwork.calculate_shortest_paths (streetnetwork.py:108-109) with
 * explicit per-street weights (the network's current driving times,
 * streetnetwork.py:60-61).  dist_out[V] (+inf = unreachable), pred_out[V]
 * (-1 for the origin and for unreachable nodes); either may be NULL.
 * Predecessor rule: lowest neighbour id among those reproducing dist exactly. */
int s4_graph_sssp(s4_graph *g, const double *w_street, int32_t origin, double *dist_out, int32_t *pred_out);

/* ---- simulation = one (virtual) rank (simulation.py:37-45, streets4mpi.py:67-82) */
/* T trips (origin[i] -> goal[i]); jam_tolerance as drawn at streets4mpi.py:78.
 * Trips are grouped by tree root on the device (tripgenerator.py:37-42 did that
 * with a dict).  The simulation owns a private copy of the speed limits, like
 * every reference rank owns its StreetNetwork. */
int s4_sim_create(s4_graph *g, int64_t T, const int32_t *origin, const int32_t *goal,
                  double jam_tolerance, const s4_physics *phys, s4_sim **out);
/* The same with the trips drawn on the device (tripgenerator.py:33-35: one uniform origin and one uniform goal per
 * resident): trip t takes Philox4x32-10 at counter (t, 0) under key `seed`, words 0,1 index origins[], words 2,3
 * goals[] (index = floor(r64 * n / 2^64)); only the two populations cross PCIe.  streets4mpi.py:67-75. */
int s4_sim_create_sampled(s4_graph *g, int64_t T, int64_t n_origins, const int32_t *origins, int64_t n_goals,
                          const int32_t *goals, uint64_t seed, double jam_tolerance, const s4_physics *phys, s4_sim **out);
/* the simulation's T trips, grouped by tree root (origin_out[T], goal_out[T]) */
int s4_sim_get_trips(s4_sim *sim, int32_t *origin_out, int32_t *goal_out);
int s4_sim_free(s4_sim *sim);
/* number of distinct tree roots and the root mode actually used */
int s4_sim_roots(s4_sim *sim, int64_t *n_roots, int *root_mode);

/* Simulation.step (simulation.py:48-88): weights from the resident traffic_load
 * (:53-65), reset it (:68), one tree per root (:71-75), walk every trip and add
 * trip_volume per street (:78-88).  Asynchronous. */
int s4_sim_step(s4_sim *sim);
/* The per-street driving times (simulation.py:53-65).  refresh != 0: run K1 alone
 * on the resident load without resetting it; refresh == 0: the weights the last
 * step (or refresh) routed on.  w_out[S] may be NULL. */
int s4_sim_weights(s4_sim *sim, int refresh, double *w_out);

/* `simulation.traffic_load` (uint32[S], array('I')): read (D2H) / assign (H2D,
 * streets4mpi.py:99) / device address for an in-place collective (:97-98). */
int s4_sim_get_load(s4_sim *sim, uint32_t *out);
int s4_sim_set_load(s4_sim *sim, const uint32_t *in);
void *s4_sim_load_device_ptr(s4_sim *sim);
/* dst.load += src.load  /  dst.load = src.load -- several virtual ranks on one GPU */
int s4_sim_load_add(s4_sim *dst, s4_sim *src);
int s4_sim_load_copy(s4_sim *dst, s4_sim *src);
/* streets4mpi.py:100 + utils.py:26-34: cumulative = load + (cumulative or 0) */
int s4_sim_commit_total(s4_sim *sim);
/* streets4mpi.py:97-100 in one stream-ordered call: traffic_load = Allreduce(traffic_load, SUM) over the ranks of
 * s4_comm_init (ncclAllReduce, ncclUint32, in place on the device, no host synchronisation), then
 * cumulative = traffic_load + (cumulative or 0).  With one rank it is s4_sim_commit_total. */
int s4_sim_exchange(s4_sim *sim);
/* 64-bit FNV-1a of traffic_load and of cumulative_traffic_load (0 when None); either pointer may be NULL.
 * uint32 sums are exact, so the hashes of a day do not depend on how its trips were partitioned over GPUs. */
int s4_sim_load_hash(s4_sim *sim, uint64_t *load_hash, uint64_t *cumulative_hash);
/* `simulation.cumulative_traffic_load`: *is_none = 1 when it is None */
int s4_sim_get_cumulative(s4_sim *sim, uint32_t *out, int *is_none);
int s4_sim_set_cumulative(s4_sim *sim, const uint32_t *in_or_null);

/* Simulation.road_construction (simulation.py:91-109, streetnetwork.py:79-82):
 * rank streets by (cumulative load, street_index), apply the literal sweep,
 * clamp to [1,140], cumulative = None.  adaptable[S] (may be NULL = all) marks
 * the streets whose change is visible to the weight loop (insertion orientation
 * u <= v, see DESIGN.md "orientation aliasing"); every street's
 * insertion-orientation value changes regardless. */
int s4_sim_road_construction(s4_sim *sim, const uint8_t *adaptable);
/* speed limits as the weight loop sees them / as stored on the insertion orientation */
int s4_sim_get_max_speed(s4_sim *sim, int32_t *visible_out, int32_t *insertion_out);
int s4_sim_set_max_speed(s4_sim *sim, const int32_t *visible, const int32_t *insertion_or_null);

/* Diagnostics: the day's root slots as the SSSP kernel consumes them -- consecutive groups of *group_size slots, one
 * group = the trees that share a frontier, -1 = empty slot; node ids in the caller's numbering.  out may be NULL to
 * query *n_slots.  Valid after the first s4_sim_step (the group size follows the launch geometry). */
int s4_sim_get_root_slots(s4_sim *sim, int32_t *out, int64_t cap, int64_t *n_slots, int *group_size);
int s4_sim_get_counters(s4_sim *sim, s4_counters *out);
/* Trips whose walk found no predecessor since the last reset (simulation.py:83-85 would raise KeyError there): only
 * possible on a zero-weight plateau of more than 32 nodes.  Cheap (8 bytes D2H behind the stream); the host layer
 * warns when the number grows. */
int s4_sim_trips_stuck(s4_sim *sim, uint64_t *out);
int s4_sim_reset_counters(s4_sim *sim);

/* calculate_driving_speed (simulation.py:129-138) for n streets, on the device */
int s4_driving_speed(s4_ctx *ctx, int64_t n, const double *length, const int32_t *max_speed,
                     const uint32_t *trips, const s4_physics *phys, double *out);

#ifdef __cplusplus
}
#endif
#endif /* S4MPI_H */
