// Device code of libs4mpi: the per-day traffic-assignment path of Streets4MPI,
// written for sm_100a (B200).  Nothing here is a dense contraction, so there is
// no tcgen05 / TMEM: the path is HBM- and latency-bound integer / fp64 work and
// the design rules are coalesced 16-byte arc records, L2-coherent (ld.cg)
// distance words, shared-memory frontier staging, warp ballot compaction and a
// persistent grid sized to the SM count.
//
//   K1  weights_street_kernel / weights_arc_kernel   simulation.py:53-65, :129-138
//   K2  sssp_nearfar_kernel                          simulation.py:71-75 -> streetnetwork.py:108-109
//   K3  walk_kernel                                  simulation.py:78-88
//   K4b merge kernels                                streets4mpi.py:100, utils.py:26-34
//   K5  adapt_kernel                                 simulation.py:91-109, streetnetwork.py:79-82
#pragma once

#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace cg = cooperative_groups;

namespace s4 {

typedef unsigned long long u64;
typedef unsigned int u32;

// +inf as a bit pattern.  Distances are non-negative doubles, whose IEEE-754 bit
// patterns order exactly like unsigned integers: all distance compares and the
// atomic min run on the u64 view.
static constexpr u64 kInfBits = 0x7FF0000000000000ull;

// device-side counters (indices into a u64 array)
enum Ctr {
    C_ARCS_RELAXED = 0,
    C_NODES_POPPED,
    C_STALE_POPS,
    C_PUSHES,
    C_ADVANCES,
    C_ROUNDS,
    C_FALLBACK_SWEEPS,
    C_TRIPS,
    C_UNREACHABLE,
    C_STUCK,
    C_HOPS,
    C_COUNT
};

struct Phys {
    double car_length, min_brk, decel;
};

// One CSR arc: head node, street_index, weight of the day.  16 bytes so that a
// node's arcs are fetched with LDG.128 and consecutive lanes coalesce.
struct __align__(16) Arc {
    u32 col;
    u32 street;
    double w;
};

__device__ __forceinline__ u64 ld_dist(const u64 *p) { return __ldcg(p); }   // L2-coherent with the atomics
__device__ __forceinline__ u32 hi32(u64 x) { return (u32)(x >> 32); }

__device__ __forceinline__ Arc ld_arc(const Arc *p)
{
    uint4 r = __ldg(reinterpret_cast<const uint4 *>(p));
    Arc a;
    a.col = r.x;
    a.street = r.y;
    a.w = __hiloint2double((int)r.w, (int)r.z);
    return a;
}

// ---------------------------------------------------------------------------
// K1  simulation.py:129-138.  Every operation is an explicitly rounded
// intrinsic: nvcc must not contract a*b+c (the file is also built -fmad=false).
// ---------------------------------------------------------------------------
__device__ __forceinline__ double driving_speed(double length, double max_speed, double trips, Phys ph)
{
    double per_car = __ddiv_rn(length, trips > 1.0 ? trips : 1.0);            // :131  length / max(n, 1)
    double room = __dadd_rn(per_car, -ph.car_length);                          // :132  space - car_length
    if (!(room > ph.min_brk)) room = ph.min_brk;                               //       max(.., min_breaking_distance)
    double pot = __dsqrt_rn(__dmul_rn(__dmul_rn(ph.decel, room), 2.0));        // :134  sqrt(decel * room * 2)
    double kmh = __dmul_rn(pot, 3.6);                                          // :136
    return max_speed <= kmh ? max_speed : kmh;                                 //       min(max_speed, kmh)
}

__device__ __forceinline__ double driving_time(double length, double max_speed, double load, double jam, Phys ph)
{
    double ideal = driving_speed(length, max_speed, 0.0, ph);                  // :57
    double actual = driving_speed(length, max_speed, load, ph);                // :59
    double perceived = __dadd_rn(actual, __dmul_rn(__dadd_rn(ideal, -actual), jam));   // :61
    return __ddiv_rn(length, perceived);                                       // :63
}

// per street: weight of the day from yesterday's total load; optionally reset
// the load counter (simulation.py:68) and accumulate sum(w) for the bucket width.
__global__ void weights_street_kernel(int64_t S, const double *__restrict__ length,
                                      const int32_t *__restrict__ max_speed, u32 *__restrict__ load,
                                      double jam, Phys ph, int reset_load, double *__restrict__ w_street,
                                      double *__restrict__ wsum)
{
    double local = 0.0;
    for (int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; s < S; s += (int64_t)gridDim.x * blockDim.x) {
        double w = driving_time(length[s], (double)max_speed[s], (double)load[s], jam, ph);
        w_street[s] = w;
        if (reset_load) load[s] = 0u;
        local += w;
    }
    // block reduction of the (heuristic-only) weight sum
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    __shared__ double part[32];
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = threadIdx.x < (blockDim.x + 31) / 32 ? part[threadIdx.x] : 0.0;
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0) atomicAdd(wsum, v);
    }
}

// per arc: build the 16-byte record {head, street, w[street]}
__global__ void weights_arc_kernel(int64_t A, const uint2 *__restrict__ arc_cs, const double *__restrict__ w_street,
                                   Arc *__restrict__ arcs)
{
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < A; e += (int64_t)gridDim.x * blockDim.x) {
        uint2 cs = arc_cs[e];
        Arc a;
        a.col = cs.x;
        a.street = cs.y;
        a.w = w_street[cs.y];
        arcs[e] = a;
    }
}

__global__ void driving_speed_kernel(int64_t n, const double *length, const int32_t *max_speed, const u32 *trips,
                                     Phys ph, double *out)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = driving_speed(length[i], (double)max_speed[i], (double)trips[i], ph);
}

// ---------------------------------------------------------------------------
// K2  batched multi-source SSSP, near-far (delta-stepping with one near bucket
// and a far pile), one CTA per tree, persistent grid pulling roots from a queue.
//
// Per tree the CTA owns one distance slot dist[V] (u64 view of fp64) in the
// pool.  A frontier entry is (node, high 32 bits of the distance it was pushed
// with): popping it re-reads dist[node] and drops the entry if the word has
// since improved past those bits (superseded), so a node is expanded once per
// improvement, not once per push.  Relaxation is TILE lanes per frontier node:
// the lanes of a tile read consecutive 16-byte arc records (one coalesced
// request per node), then dist[head] (ld.cg), then atomicMin on the u64 view.
// Improved heads go to the next near queue (shared memory, spilling to global)
// when below the bucket threshold, else to the far pile (global); both pushes
// are compacted with one warp ballot and one shared-memory atomic per warp.
//
// The fixed point reached is schedule-independent and equals Dijkstra's:
// dist[v] = min over neighbours of fl(dist[u] + w) with left-to-right rounding.
// ---------------------------------------------------------------------------
struct SsspParams {
    const u32 *row_ptr;     // V+1
    const Arc *arcs;        // A
    u64 *dist_pool;         // n_roots slots of V words (slot i <-> roots[i])
    const int32_t *roots;   // roots of this batch
    int n_roots;
    u32 V;
    u32 *work_counter;      // zeroed before launch
    const double *wsum;     // sum of street weights (device scalar)
    double delta_scale;     // delta = delta_scale * wsum   (= delta_factor / S)
    uint2 *ws_near;         // per CTA: 2 * near_gcap
    uint2 *ws_far;          // per CTA: 2 * far_cap
    u32 near_scap;          // entries per near queue in shared memory
    u32 near_gcap;          // spill entries per near queue in global memory
    u32 far_cap;            // entries per far pile
    u64 *counters;
};

template <int BLOCK, int TILE>
__global__ void __launch_bounds__(BLOCK) sssp_nearfar_kernel(SsspParams p)
{
    static_assert(BLOCK % 32 == 0 && 32 % TILE == 0, "tile must divide the warp");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint2 *sq0 = reinterpret_cast<uint2 *>(smem_raw);
    uint2 *sq1 = sq0 + p.near_scap;
    __shared__ u32 s_cnt[3];       // near queue fill counters, rotating (one barrier per round)
    __shared__ u32 s_far_cnt[2];
    __shared__ u32 s_far_min[2];   // min high word among the entries of each far pile
    __shared__ int s_root_idx;
    __shared__ u32 s_overflow;
    __shared__ u32 s_changed;

    const u32 tid = threadIdx.x;
    const u32 lane = tid & 31u;
    const u32 lt_mask = (1u << lane) - 1u;
    const u32 sub = tid % TILE;
    const u32 grp = tid / TILE;
    constexpr u32 NG = BLOCK / TILE;

    uint2 *gq0 = p.ws_near + (size_t)blockIdx.x * 2u * p.near_gcap;
    uint2 *gq1 = gq0 + p.near_gcap;
    uint2 *far0 = p.ws_far + (size_t)blockIdx.x * 2u * p.far_cap;
    uint2 *far1 = far0 + p.far_cap;
    const u32 scap = p.near_scap, gcap = p.near_gcap, fcap = p.far_cap;
    const u32 V = p.V;

    double delta = p.delta_scale * (*p.wsum);
    if (!(delta > 0.0)) delta = 0.0;

    u64 c_relax = 0, c_pop = 0, c_stale = 0, c_push = 0, c_adv = 0, c_round = 0, c_sweep = 0;

    for (;;) {
        __syncthreads();                       // previous tree fully retired before s_* are reused
        if (tid == 0) s_root_idx = (int)atomicAdd(p.work_counter, 1u);
        __syncthreads();
        const int ri = s_root_idx;
        if (ri >= p.n_roots) break;
        u64 *dist = p.dist_pool + (size_t)ri * V;
        const u32 root = (u32)p.roots[ri];

        // ---- fill the slot with +inf (16-byte stores; V*8 bytes, slot is 16-byte aligned when V is even)
        {
            const size_t base_addr = (size_t)ri * V;
            u32 head = (u32)(base_addr & 1u);              // leading word if the slot starts on an odd word
            if (head && tid == 0) dist[0] = kInfBits;
            ulonglong2 *d2 = reinterpret_cast<ulonglong2 *>(dist + head);
            const u32 n2 = (V - head) >> 1;
            const ulonglong2 inf2 = make_ulonglong2(kInfBits, kInfBits);
            for (u32 i = tid; i < n2; i += BLOCK) d2[i] = inf2;
            if (((V - head) & 1u) && tid == 0) dist[V - 1] = kInfBits;
        }
        if (tid == 0) {
            s_cnt[0] = 1u; s_cnt[1] = 0u; s_cnt[2] = 0u;
            s_far_cnt[0] = 0u; s_far_cnt[1] = 0u;
            s_far_min[0] = 0xFFFFFFFFu; s_far_min[1] = 0xFFFFFFFFu;
            s_overflow = 0u;
            sq0[0] = make_uint2(root, 0u);
        }
        __syncthreads();
        if (tid == 0) dist[root] = 0ull;                   // after the fill
        __syncthreads();

        u32 c = 0;          // index of the current near counter
        u32 b = 0;          // current near buffer
        u32 f = 0;          // current far pile (push target)
        u64 thr_bits = (u64)__double_as_longlong(delta);   // near bucket = [.., thr)

        for (;;) {
            const u32 n = s_cnt[c];
            if (n > 0) {
                // ------------------------- relaxation round -------------------------
                const u32 cn = c == 2 ? 0 : c + 1;
                if (tid == 0) s_cnt[cn == 2 ? 0 : cn + 1] = 0u;        // counter of the round after next
                uint2 *qc = b ? sq1 : sq0, *gqc = b ? gq1 : gq0;
                uint2 *qn = b ? sq0 : sq1, *gqn = b ? gq0 : gq1;
                uint2 *farp = f ? far1 : far0;
                const u32 n_s = n < scap + gcap ? n : scap + gcap;      // entries that were actually stored
                ++c_round;
                for (u32 base = 0; base < n_s; base += NG) {
                    const u32 i = base + grp;
                    bool valid = i < n_s;
                    u32 rb = 0, deg = 0;
                    u64 du = 0;
                    if (valid) {
                        const uint2 ent = i < scap ? qc[i] : gqc[i - scap];
                        du = ld_dist(dist + ent.x);
                        if (hi32(du) < ent.y) {
                            valid = false;                              // superseded by a later push
                            if (sub == 0) ++c_stale;
                        } else {
                            rb = __ldg(p.row_ptr + ent.x);
                            deg = __ldg(p.row_ptr + ent.x + 1) - rb;
                            if (sub == 0) ++c_pop;
                        }
                    }
                    for (u32 k = sub; __any_sync(0xffffffffu, valid && k < deg); k += TILE) {
                        const bool act = valid && k < deg;
                        bool imp = false;
                        u32 v = 0;
                        u64 ndb = 0;
                        if (act) {
                            const Arc a = ld_arc(p.arcs + rb + k);
                            v = a.col;
                            ndb = (u64)__double_as_longlong(__dadd_rn(__longlong_as_double((long long)du), a.w));
                            u64 old = ld_dist(dist + v);
                            if (ndb < old) {
                                old = atomicMin(dist + v, ndb);
                                imp = ndb < old;
                            }
                            ++c_relax;
                        }
                        const bool to_near = imp && ndb < thr_bits;
                        const bool to_far = imp && !to_near;
                        const u32 mn = __ballot_sync(0xffffffffu, to_near);
                        if (mn) {
                            const int leader = __ffs(mn) - 1;
                            u32 basei = 0;
                            if ((int)lane == leader) basei = atomicAdd(&s_cnt[cn], (u32)__popc(mn));
                            basei = __shfl_sync(0xffffffffu, basei, leader);
                            if (to_near) {
                                const u32 idx = basei + (u32)__popc(mn & lt_mask);
                                const uint2 e = make_uint2(v, hi32(ndb));
                                if (idx < scap) qn[idx] = e;
                                else if (idx - scap < gcap) gqn[idx - scap] = e;
                                else s_overflow = 1u;
                                ++c_push;
                            }
                        }
                        const u32 mf = __ballot_sync(0xffffffffu, to_far);
                        if (mf) {
                            const int leader = __ffs(mf) - 1;
                            const u32 wmin = __reduce_min_sync(0xffffffffu, to_far ? hi32(ndb) : 0xFFFFFFFFu);
                            u32 basei = 0;
                            if ((int)lane == leader) {
                                basei = atomicAdd(&s_far_cnt[f], (u32)__popc(mf));
                                atomicMin(&s_far_min[f], wmin);
                            }
                            basei = __shfl_sync(0xffffffffu, basei, leader);
                            if (to_far) {
                                const u32 idx = basei + (u32)__popc(mf & lt_mask);
                                if (idx < fcap) farp[idx] = make_uint2(v, hi32(ndb));
                                else s_overflow = 1u;
                                ++c_push;
                            }
                        }
                    }
                }
                __syncthreads();
                c = cn;
                b ^= 1u;
                continue;
            }
            // ------------------------- near queue empty: advance the bucket -------------------------
            const u32 nf_raw = s_far_cnt[f];
            if (nf_raw == 0u) break;                                    // tree complete
            const u32 nf = nf_raw < fcap ? nf_raw : fcap;
            ++c_adv;
            // new threshold: smallest (lower bound of a) distance in the pile + delta
            const double lo = __longlong_as_double((long long)((u64)s_far_min[f] << 32));
            const double thr = __dadd_rn(lo, delta);
            thr_bits = (u64)__double_as_longlong(thr);
            const u32 thr_hi = hi32(thr_bits);
            uint2 *src = f ? far1 : far0;
            uint2 *dst = f ? far0 : far1;
            uint2 *qc = b ? sq1 : sq0, *gqc = b ? gq1 : gq0;            // the round after this reads (b, c)
            const u32 fo = f ^ 1u;
            for (u32 base = 0; base < nf; base += BLOCK) {
                const u32 i = base + tid;
                const bool have = i < nf;
                uint2 e = make_uint2(0u, 0u);
                if (have) e = src[i];
                const bool to_near = have && e.y <= thr_hi;
                const bool keep = have && !to_near;
                const u32 mn = __ballot_sync(0xffffffffu, to_near);
                if (mn) {
                    const int leader = __ffs(mn) - 1;
                    u32 basei = 0;
                    if ((int)lane == leader) basei = atomicAdd(&s_cnt[c], (u32)__popc(mn));
                    basei = __shfl_sync(0xffffffffu, basei, leader);
                    if (to_near) {
                        const u32 idx = basei + (u32)__popc(mn & lt_mask);
                        if (idx < scap) qc[idx] = e;
                        else if (idx - scap < gcap) gqc[idx - scap] = e;
                        else s_overflow = 1u;
                    }
                }
                const u32 mk = __ballot_sync(0xffffffffu, keep);
                if (mk) {
                    const int leader = __ffs(mk) - 1;
                    const u32 wmin = __reduce_min_sync(0xffffffffu, keep ? e.y : 0xFFFFFFFFu);
                    u32 basei = 0;
                    if ((int)lane == leader) {
                        basei = atomicAdd(&s_far_cnt[fo], (u32)__popc(mk));
                        atomicMin(&s_far_min[fo], wmin);
                    }
                    basei = __shfl_sync(0xffffffffu, basei, leader);
                    if (keep) dst[basei + (u32)__popc(mk & lt_mask)] = e;   // never exceeds nf <= fcap
                }
            }
            __syncthreads();
            if (tid == 0) { s_far_cnt[f] = 0u; s_far_min[f] = 0xFFFFFFFFu; }
            f = fo;
            // no barrier needed: pile f^1 is not a push target until the next advance,
            // and a relaxation round (with its barrier) always intervenes.
        }

        // ---- safety net: a queue overflowed and entries were dropped.  Sweep all
        // nodes Bellman-Ford style until nothing changes (same fixed point).
        if (s_overflow) {
            for (;;) {
                __syncthreads();
                if (tid == 0) s_changed = 0u;
                __syncthreads();
                ++c_sweep;
                bool any = false;
                for (u32 u = tid; u < V; u += BLOCK) {
                    const u64 du = ld_dist(dist + u);
                    if (du >= kInfBits) continue;
                    const u32 rb = __ldg(p.row_ptr + u), re = __ldg(p.row_ptr + u + 1);
                    for (u32 e = rb; e < re; ++e) {
                        const Arc a = ld_arc(p.arcs + e);
                        const u64 ndb = (u64)__double_as_longlong(__dadd_rn(__longlong_as_double((long long)du), a.w));
                        if (ndb < ld_dist(dist + a.col)) {
                            atomicMin(dist + a.col, ndb);
                            any = true;
                        }
                    }
                }
                if (any) s_changed = 1u;
                __syncthreads();
                if (!s_changed) break;
            }
        }
    }

    // ---- flush per-thread counters: one atomic per warp per counter
    u64 vals[7] = {c_relax, c_pop, c_stale, c_push, tid == 0 ? c_adv : 0, tid == 0 ? c_round : 0, tid == 0 ? c_sweep : 0};
    const int ids[7] = {C_ARCS_RELAXED, C_NODES_POPPED, C_STALE_POPS, C_PUSHES, C_ADVANCES, C_ROUNDS, C_FALLBACK_SWEEPS};
#pragma unroll
    for (int j = 0; j < 7; ++j) {
        u64 v = vals[j];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0 && v) atomicAdd(p.counters + ids[j], v);
    }
}

// ---------------------------------------------------------------------------
// Canonical predecessor of `cur` from converged distances, evaluated by the
// lanes of one tile (lane k looks at arcs rb+k, rb+k+TILE, ...):
//   min (u, street) over arcs (cur,u), u != cur, with fl(dist[u]+w) == dist[cur]
//   and (dist[u] < dist[cur] or u < cur); if empty, over all equal-cost arcs.
// Returns the packed key (u << 32 | street), ~0 if there is none; *du_out is
// dist[u] of the winner (valid on every lane of the tile).
// ---------------------------------------------------------------------------
template <int TILE>
__device__ __forceinline__ u64 canonical_pred(const cg::thread_block_tile<TILE> &tile, const u32 *__restrict__ row_ptr,
                                              const Arc *__restrict__ arcs, const u64 *dist, u32 cur, u64 dv,
                                              u64 *du_out)
{
    const u32 rb = __ldg(row_ptr + cur), re = __ldg(row_ptr + cur + 1);
    u64 strict = ~0ull, loose = ~0ull, du_s = 0, du_l = 0;
    for (u32 e = rb + tile.thread_rank(); tile.any(e < re); e += TILE) {
        if (e < re) {
            const Arc a = ld_arc(arcs + e);
            if (a.col != cur) {
                const u64 du = ld_dist(dist + a.col);
                const u64 ndb = (u64)__double_as_longlong(__dadd_rn(__longlong_as_double((long long)du), a.w));
                if (du < kInfBits && ndb == dv) {
                    const u64 key = ((u64)a.col << 32) | a.street;
                    if (key < loose) { loose = key; du_l = du; }
                    if ((du < dv || a.col < cur) && key < strict) { strict = key; du_s = du; }
                }
            }
        }
    }
    // tile-wide minimum of both keys, carrying the winner's distance along
    for (int o = TILE / 2; o > 0; o >>= 1) {
        const u64 ks = tile.shfl_xor(strict, o), ds = tile.shfl_xor(du_s, o);
        if (ks < strict) { strict = ks; du_s = ds; }
        const u64 kl = tile.shfl_xor(loose, o), dl = tile.shfl_xor(du_l, o);
        if (kl < loose) { loose = kl; du_l = dl; }
    }
    if (strict != ~0ull) { *du_out = du_s; return strict; }
    *du_out = du_l;
    return loose;
}

// ---------------------------------------------------------------------------
// K3  simulation.py:78-88: one tile of lanes per trip walks target -> root up
// the canonical predecessor tree (derived on the fly from the distance slot,
// no pred[] array is ever materialised) and scatter-adds trip_volume into the
// per-street uint32 counters.
// ---------------------------------------------------------------------------
struct WalkParams {
    const u32 *row_ptr;
    const Arc *arcs;
    const u64 *dist_pool;
    const int32_t *roots;       // distinct roots (all of them)
    const int32_t *trip_target; // trips sorted by root
    const u32 *trip_root;       // index of the trip's root among the distinct roots
    int64_t t0, t1;             // trips of this batch
    u32 root0;                  // first root of this batch (slot = trip_root - root0)
    u32 V;
    u32 volume;
    u32 *load;
    u64 *counters;
};

template <int TILE>
__global__ void __launch_bounds__(256) walk_kernel(WalkParams p)
{
    cg::thread_block block = cg::this_thread_block();
    cg::thread_block_tile<TILE> tile = cg::tiled_partition<TILE>(block);
    const int64_t tiles_per_grid = (int64_t)gridDim.x * (blockDim.x / TILE);
    const int64_t tile_id = (int64_t)blockIdx.x * (blockDim.x / TILE) + threadIdx.x / TILE;
    u64 c_hops = 0, c_unreach = 0, c_stuck = 0, c_trips = 0;
    for (int64_t t = p.t0 + tile_id; t < p.t1; t += tiles_per_grid) {
        const u32 ri = p.trip_root[t];
        const u32 root = (u32)p.roots[ri];
        const u64 *dist = p.dist_pool + (size_t)(ri - p.root0) * p.V;
        u32 cur = (u32)p.trip_target[t];
        u64 dv = ld_dist(dist + cur);
        ++c_trips;
        if (dv >= kInfBits) { ++c_unreach; continue; }                     // simulation.py:80
        u32 guard = 0;
        while (cur != root) {                                              // :83
            u64 du;
            const u64 key = canonical_pred<TILE>(tile, p.row_ptr, p.arcs, dist, cur, dv, &du);
            if (key == ~0ull || ++guard > p.V) { ++c_stuck; break; }
            if (tile.thread_rank() == 0) atomicAdd(p.load + (u32)key, p.volume);   // :86-88
            ++c_hops;
            cur = (u32)(key >> 32);                                        // :85
            dv = du;
        }
    }
    if (tile.thread_rank() == 0) {
        if (c_trips) atomicAdd(p.counters + C_TRIPS, c_trips);
        if (c_hops) atomicAdd(p.counters + C_HOPS, c_hops);
        if (c_unreach) atomicAdd(p.counters + C_UNREACHABLE, c_unreach);
        if (c_stuck) atomicAdd(p.counters + C_STUCK, c_stuck);
    }
}

// predecessor map of one tree (StreetNetwork.calculate_shortest_paths, streetnetwork.py:108-109)
template <int TILE>
__global__ void __launch_bounds__(256) pred_kernel(const u32 *row_ptr, const Arc *arcs, const u64 *dist, u32 V,
                                                   u32 root, int32_t *pred)
{
    cg::thread_block block = cg::this_thread_block();
    cg::thread_block_tile<TILE> tile = cg::tiled_partition<TILE>(block);
    const u32 tiles_per_grid = gridDim.x * (blockDim.x / TILE);
    for (u32 v = blockIdx.x * (blockDim.x / TILE) + threadIdx.x / TILE; v < V; v += tiles_per_grid) {
        const u64 dv = ld_dist(dist + v);
        int32_t pr = -1;
        if (v != root && dv < kInfBits) {
            u64 du;
            const u64 key = canonical_pred<TILE>(tile, row_ptr, arcs, dist, v, dv, &du);
            if (key != ~0ull) pr = (int32_t)(key >> 32);
        }
        if (tile.thread_rank() == 0) pred[v] = pr;
    }
}

// ---------------------------------------------------------------------------
// small elementwise kernels
// ---------------------------------------------------------------------------
__global__ void add_u32_kernel(int64_t n, u32 *__restrict__ dst, const u32 *__restrict__ src)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        dst[i] += src[i];
}

__global__ void iota_kernel(int64_t n, u32 *out)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = (u32)i;
}

// trip grouping: after sorting trips by root, mark the first trip of every root
__global__ void head_flag_kernel(int64_t T, const int32_t *__restrict__ sorted_root, u32 *__restrict__ flag)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < T; i += (int64_t)gridDim.x * blockDim.x)
        flag[i] = (i == 0 || sorted_root[i] != sorted_root[i - 1]) ? 1u : 0u;
}

// rank[i] = inclusive_scan(flag)[i]; root index = rank-1; heads record root node and first-trip offset
__global__ void group_scatter_kernel(int64_t T, const int32_t *__restrict__ sorted_root, const u32 *__restrict__ flag,
                                     const u32 *__restrict__ rank, u32 *__restrict__ trip_root,
                                     int32_t *__restrict__ roots, int64_t *__restrict__ root_off)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < T; i += (int64_t)gridDim.x * blockDim.x) {
        const u32 r = rank[i] - 1u;
        trip_root[i] = r;
        if (flag[i]) { roots[r] = sorted_root[i]; root_off[r] = i; }
    }
}

// ---------------------------------------------------------------------------
// K5  simulation.py:99-108 + streetnetwork.py:79-82 on the ranked streets.
// order[r] = street with rank r in (cumulative load, street_index) order.  The
// literal sweep decreases ranks [0, n_dec) and increases ranks [S-n_inc, S),
// iteration i touching rank i and then rank S-1-i.
// ---------------------------------------------------------------------------
__device__ __forceinline__ int32_t clamp_speed(int32_t x) { return x < 1 ? 1 : (x > 140 ? 140 : x); }

__global__ void adapt_kernel(int64_t S, const u32 *__restrict__ order, int64_t n_dec, int64_t n_inc,
                             const uint8_t *__restrict__ adaptable, int32_t *__restrict__ ms_visible,
                             int32_t *__restrict__ ms_insertion)
{
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < S; r += (int64_t)gridDim.x * blockDim.x) {
        const bool dec = r < n_dec;
        const bool inc = r >= S - n_inc;
        if (!dec && !inc) continue;
        const u32 s = order[r];
        int32_t x = ms_insertion[s];
        if (dec && inc) {
            // both sweeps reach this rank: dec in iteration r, inc in iteration S-1-r (dec first on a tie)
            if (r <= S - 1 - r) { x = clamp_speed(x - 20); x = clamp_speed(x + 20); }
            else { x = clamp_speed(x + 20); x = clamp_speed(x - 20); }
        } else if (dec) x = clamp_speed(x - 20);
        else x = clamp_speed(x + 20);
        ms_insertion[s] = x;
        if (!adaptable || adaptable[s]) ms_visible[s] = x;
    }
}

}  // namespace s4
