// Device code of libs4mpi: the per-day traffic-assignment path of Streets4MPI,
// written for sm_100a (B200).  Nothing here is a dense contraction, so there is
// no tcgen05 / TMEM: the path is HBM- and latency-bound integer / fp64 work and
// the design rules are coalesced 16-byte arc records, L2-coherent (ld.cg)
// distance words, shared-memory frontier staging, warp ballot compaction and a
// persistent grid sized to the SM count.
//
//   K1  weights_street_kernel / weight_scale_kernel / weights_arc_kernel / weights_ellw_kernel
//                                                    simulation.py:53-65, :129-138
//   K2  sssp_group_kernel (s4_sssp_group.cuh: 15 / 31 trees per CTA or cluster on one shared frontier);
//       sssp_nearfar_kernel (here: one tree per CTA, round 1, still selectable)
//                                                    simulation.py:71-75 -> streetnetwork.py:108-109
//   K3  walk_kernel / walk_slow_kernel (here); pred_code_kernel / walk_code_kernel (s4_walk_codes.cuh)
//                                                    simulation.py:78-88
//   K4b merge kernels (the all-reduce itself is NCCL, s4_capi.cu)      streets4mpi.py:100, utils.py:26-34
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

// per street: weight of the day from yesterday's total load; optionally reset the load counter (simulation.py:68) and
// build the histogram the bucket width is derived from.
// Bucket width (a heuristic: no result depends on it).  delta is a multiple of the TYPICAL street weight.  The mean is
// a poor "typical" once road construction has pushed unused streets down to 1 km/h (weights 30-50x the bulk: the mean
// follows them, buckets hold most of the graph, frontiers overflow); the median does not move.  The weights are
// binned by the top 14 bits of their fp64 pattern (exponent + 2 mantissa bits: bin edges 19 % apart, 8192 bins for
// positive finite values), privatised per block in shared memory; weight_scale_kernel then reads off the median bin.
static constexpr int kWeightBins = 8192;

__global__ void __launch_bounds__(256) weights_street_kernel(int64_t S, const double *__restrict__ length,
                                                             const int32_t *__restrict__ max_speed, u32 *__restrict__ load,
                                                             double jam, Phys ph, int reset_load, double *__restrict__ w_street,
                                                             u32 *__restrict__ hist)
{
    __shared__ u32 h[kWeightBins];
    for (int i = threadIdx.x; i < kWeightBins; i += blockDim.x) h[i] = 0u;
    __syncthreads();
    for (int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; s < S; s += (int64_t)gridDim.x * blockDim.x) {
        double w = driving_time(length[s], (double)max_speed[s], (double)load[s], jam, ph);
        w_street[s] = w;
        if (reset_load) load[s] = 0u;
        const u32 hi = hi32((u64)__double_as_longlong(w));
        if (w > 0.0 && hi < 0x7FF00000u) atomicAdd(&h[hi >> 18], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kWeightBins; i += blockDim.x)
        if (h[i]) atomicAdd(hist + i, h[i]);
}

// one block: the median bin of the histogram -> scale[0] = (typical weight) * S, the quantity the SSSP kernels
// multiply by delta_factor / S; clears the histogram for the next day
__global__ void __launch_bounds__(1024) weight_scale_kernel(u32 *__restrict__ hist, int64_t S, double *__restrict__ scale)
{
    constexpr int PER = kWeightBins / 1024;
    __shared__ unsigned long long part[1024];
    u32 mine[PER];
    unsigned long long sum = 0;
    for (int k = 0; k < PER; ++k) { mine[k] = hist[threadIdx.x * PER + k]; hist[threadIdx.x * PER + k] = 0u; sum += mine[k]; }
    part[threadIdx.x] = sum;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {                       // inclusive scan
        const unsigned long long v = threadIdx.x >= (u32)o ? part[threadIdx.x - o] : 0ull;
        __syncthreads();
        part[threadIdx.x] += v;
        __syncthreads();
    }
    const unsigned long long total = part[1023], half = (total + 1) / 2;
    if (total == 0) { if (threadIdx.x == 0) scale[0] = 0.0; return; }
    unsigned long long before = part[threadIdx.x] - sum;
    if (before < half && part[threadIdx.x] >= half) {          // exactly one thread
        int k = 0;
        while (before + mine[k] < half) before += mine[k++];
        const u32 bin = threadIdx.x * PER + k;
        // centre of the bin: lower edge * 2^(1/8)
        scale[0] = __hiloint2double((int)(bin << 18), 0) * 1.0905077326652577 * (double)S;
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

// per ELL slot: {head | overflow flag, slot of the reverse arc in the head's row, w[street]}.  The template
// packs (reverse slot << 29 | street); padding slots carry ~0 and get w = +inf.
__global__ void weights_ell_kernel(int64_t n_slots, const uint2 *__restrict__ ell_cs, const double *__restrict__ w_street,
                                   Arc *__restrict__ ell)
{
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n_slots; e += (int64_t)gridDim.x * blockDim.x) {
        uint2 cs = ell_cs[e];
        Arc a;
        a.col = cs.x;
        a.street = (cs.x & 0x0FFFFFFFu) | ((cs.y == 0xFFFFFFFFu ? 7u : cs.y >> 29) << 28);   // entry word of the head
        a.w = cs.y == 0xFFFFFFFFu ? __longlong_as_double((long long)kInfBits) : w_street[cs.y & 0x1FFFFFFFu];
        ell[e] = a;
    }
}

// fixed-stride rows for the walk: {head | overflow flag, street_index, w[street]} (padding: {self, -, +inf}); one
// 64-byte read replaces the row_ptr -> arcs chain of the CSR
__global__ void weights_ellw_kernel(int64_t n_slots, const uint2 *__restrict__ ell_cs, const double *__restrict__ w_street,
                                    Arc *__restrict__ ellw)
{
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n_slots; e += (int64_t)gridDim.x * blockDim.x) {
        const uint2 cs = ell_cs[e];
        Arc a;
        a.col = cs.x;
        a.street = cs.y & 0x1FFFFFFFu;
        a.w = cs.y == 0xFFFFFFFFu ? __longlong_as_double((long long)0x7FF0000000000000ll) : w_street[cs.y & 0x1FFFFFFFu];
        ellw[e] = a;
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
// Layout.  Road graphs have degree <= 4 almost everywhere, so next to the CSR the
// arcs are kept in a fixed-stride table ell[V][4] of 16-byte records (64 bytes,
// two sectors, per node): no row_ptr lookup sits on the critical path.  Unused
// slots hold {col = the node itself, w = +inf}, so they relax to +inf and never
// win -- no degree branch.  Bit 31 of slot 0's col flags a node with more than 4
// arcs; the extra arcs are read from the CSR row (rare).
//
// Per tree the CTA owns one distance slot dist[V] (u64 view of fp64) in the
// pool.  A frontier entry is (node, exact distance it was pushed with): popping
// needs nothing from memory before the arc fetch, and the entry is dropped when
// dist[node] has improved since (superseded), so a node is expanded once per
// surviving improvement.  One thread expands one node: the 64-byte row as two
// 256-bit loads and the 32-byte sector of dist[] holding the node (stale check,
// neighbours in the same sector) in flight together, then one 8-byte filtering
// read (ld.cg) per remaining head, then RED.MIN on the u64 view (fire and forget:
// the push is decided on the filtering read).  Improved heads go to the next near
// queue (shared memory, spilling to global) when below the bucket threshold, else
// to the far pile (global); pushes are compacted with a warp scan and one
// shared-memory atomic per warp and queue.  The next bucket is admitted as soon
// as a round runs thin (overlapped buckets).
//
// What binds it (ncu): the L1TEX wavefront pipe -- every lane
// of a scattered access is a wavefront of its own, ~10 per node (2 row + 1 sector
// per popped entry, live or superseded; ~1.7 filtering reads and ~1.4 RED per
// live entry) -- together with the dependent latency of a pass.
//
// The fixed point reached is schedule-independent and equals Dijkstra's:
// dist[v] = min over neighbours of fl(dist[u] + w) with left-to-right rounding.
// ---------------------------------------------------------------------------
static constexpr int kEllWidth = 4;
static constexpr u32 kOvfBit = 0x80000000u;
static constexpr u32 kNodeMask = 0x0FFFFFFFu;   // frontier entry word: node | (skip slot << 28)
static constexpr u32 kNoSkip = 7u << 28;

struct SsspParams {
    const Arc *ell;         // V * kEllWidth
    const u32 *row_ptr;     // V+1  (overflow arcs, fallback sweeps)
    const Arc *arcs;        // A
    u64 *dist_pool;         // n_roots slots of slot_stride words (slot i <-> roots[i]); 32-byte aligned slots
    size_t slot_stride;     // V rounded up to a multiple of 4
    const int32_t *roots;   // roots of this batch
    int n_roots;
    u32 V;
    u32 *work_counter;      // zeroed before launch
    const double *wsum;     // sum of street weights (device scalar)
    double delta_scale;     // delta = delta_scale * wsum   (= delta_factor / S)
    u32 *ws_v;              // per CTA: [near0 | near1] gcap entries each, then [far0 | far1] fcap each
    u64 *ws_d;
    u32 near_scap;          // entries per near queue in shared memory
    u32 near_gcap;          // spill entries per near queue in global memory
    u32 far_cap;            // entries per far pile
    u32 preread;            // 1: filtering read, then returning atomicMin; 2: filtering read decides the push, RED.MIN
    u32 adaptive;           // 1: the bucket width follows the width of the wave
    u32 early_advance;      // admit the next bucket as soon as a round (not the first two of its bucket) has fewer entries
    u64 *counters;
};

// 256-bit loads (LDG.E.256 on sm_100a): one L1 wavefront instead of two per 32 bytes
__device__ __forceinline__ void ld_row_half(const Arc *p, Arc &x, Arc &y)
{
    u32 r0, r1, r2, r3, r4, r5, r6, r7;
    asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3), "=r"(r4), "=r"(r5), "=r"(r6), "=r"(r7) : "l"(p));
    x.col = r0; x.street = r1; x.w = __hiloint2double((int)r3, (int)r2);
    y.col = r4; y.street = r5; y.w = __hiloint2double((int)r7, (int)r6);
}
__device__ __forceinline__ void ld_dist_sector(const u64 *p, u64 (&s)[4])
{
    asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(s[0]), "=l"(s[1]), "=l"(s[2]), "=l"(s[3]) : "l"(p));
}
__device__ __forceinline__ u64 pick4(const u64 (&s)[4], u32 k)
{
    const u64 lo = (k & 1u) ? s[1] : s[0], hi = (k & 1u) ? s[3] : s[2];
    return (k & 2u) ? hi : lo;
}

// Expansion of one frontier entry (u, du) by one thread.  Memory operations, in flight together:
// the 64-byte ELL row as two 256-bit loads and the 32-byte sector of dist[] that holds u (stale check
// plus every neighbour that shares the sector -- with a locality-preserving numbering about half of
// them); then one 8-byte read per remaining neighbour; then atomicMin for the arcs that can win.
// kind[j]: 0 nothing, 1 improved and inside the bucket, 2 improved and beyond it.  Returns false for a
// superseded entry.  `ovf` reports a node with more than kEllWidth arcs.
__device__ __forceinline__ bool expand_entry(const Arc *__restrict__ ell, u64 *dist, u32 ent, u64 du, u32 thr_hi, u32 preread,
                                             u32 (&kind)[kEllWidth], u32 (&pv)[kEllWidth], u64 (&pd)[kEllWidth], bool &ovf)
{
    // entry word: node in bits 0..27, bits 28..30 = slot of the arc back to the node that pushed it (7 = none):
    // relaxing that arc can never win (the pusher's distance is smaller), so it is neither read nor relaxed
    const u32 u = ent & kNodeMask, skip = ent >> 28;
    Arc a[kEllWidth];
    u64 sec[4];
    const Arc *row = ell + (size_t)u * kEllWidth;
    ld_row_half(row, a[0], a[1]);
    ld_row_half(row + 2, a[2], a[3]);
    ld_dist_sector(dist + (u & ~3u), sec);
    if (pick4(sec, u & 3u) < du) return false;              // superseded by a later push
    ovf = (a[0].col & kOvfBit) != 0u;
    u64 old[kEllWidth];
#pragma unroll
    for (int j = 0; j < kEllWidth; ++j) {
        pv[j] = a[j].col & ~kOvfBit;
        pd[j] = (u64)__double_as_longlong(__dadd_rn(__longlong_as_double((long long)du), a[j].w));
        if ((u32)j == skip) pd[j] = kInfBits;
        old[j] = kInfBits;
        if (pd[j] < kInfBits) {                              // padding slots have w == +inf
            if ((pv[j] >> 2) == (u >> 2)) old[j] = pick4(sec, pv[j] & 3u);
            else old[j] = ld_dist(dist + pv[j]);
        }
    }
    if (preread == 2u) {
        // fire-and-forget: the push is decided on the filtering read alone and the minimum is sent as a
        // reduction (RED, no return value to wait for).  If another thread got a smaller value in first,
        // the entry pushed here is simply superseded and dropped when it is popped (exact compare).
#pragma unroll
        for (int j = 0; j < kEllWidth; ++j) {
            if (pd[j] < old[j]) {
                atomicMin(dist + pv[j], pd[j]);
                kind[j] = hi32(pd[j]) <= thr_hi ? 1u : 2u;
                pv[j] = a[j].street;
            }
        }
        return true;
    }
#pragma unroll
    for (int j = 0; j < kEllWidth; ++j) {
        if (pd[j] < old[j]) {                                // padding slots have pd == +inf
            const u64 o2 = atomicMin(dist + pv[j], pd[j]);
            if (pd[j] < o2) { kind[j] = hi32(pd[j]) <= thr_hi ? 1u : 2u; pv[j] = a[j].street; }
        }
    }
    return true;
}

// Where a round's pushes go (CTA-uniform).  The shared-memory part of the next near queue is addressed directly;
// its global spill and the far pile are 32-bit offsets into the CTA's workspace (SoA: entry words / distances),
// so a store costs one select and one scaled add instead of a re-derived 64-bit pointer per home.
struct PushTarget {
    u32 *sv; u64 *sd;       // next near queue, shared-memory part (scap entries)
    u32 *wv; u64 *wd;       // this CTA's workspace
    u32 near_off;           // workspace offset of that queue's spill, minus scap (entry idx >= scap lives at near_off + idx)
    u32 far_off;            // workspace offset of the far pile
    u32 scap, near_lim;     // near_lim = scap + spill capacity
    u32 fcap;
};

// Warp-cooperative append of up to N candidates per lane: kind[j] 1 = near queue, 2 = far pile.
// One scan over packed (near, far) counts, one shared-memory atomic per queue per warp.
template <int N>
__device__ __forceinline__ void warp_push(const u32 (&kind)[N], const u32 (&pv)[N], const u64 (&pd)[N], u32 lane,
                                          const PushTarget &t, u32 *near_cnt, u32 *far_cnt, u32 *far_min, u32 *overflow,
                                          u64 &c_push)
{
    u32 packed = 0, mymin = 0xFFFFFFFFu;
#pragma unroll
    for (int j = 0; j < N; ++j) {
        packed += (kind[j] & 1u) | ((kind[j] & 2u) << 15);     // near count low half, far count high half
        if (kind[j] == 2u) mymin = min(mymin, hi32(pd[j]));
    }
    u32 incl = packed;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const u32 t_ = __shfl_up_sync(0xffffffffu, incl, o);
        if ((int)lane >= o) incl += t_;
    }
    const u32 total = __shfl_sync(0xffffffffu, incl, 31);
    if (total == 0u) return;                                   // warp-uniform
    u32 base_n = 0, base_f = 0;
    if (total >> 16) {
        const u32 wmin = __reduce_min_sync(0xffffffffu, mymin);
        if (lane == 31) { base_f = atomicAdd(far_cnt, total >> 16); atomicMin(far_min, wmin); }
    }
    if ((total & 0xFFFFu) && lane == 31) base_n = atomicAdd(near_cnt, total & 0xFFFFu);
    base_n = __shfl_sync(0xffffffffu, base_n, 31);
    base_f = __shfl_sync(0xffffffffu, base_f, 31);
    const u32 excl = incl - packed;
    u32 on = base_n + (excl & 0xFFFFu), of = base_f + (excl >> 16);
#pragma unroll
    for (int j = 0; j < N; ++j) {
        if (kind[j]) {
            const bool is_near = kind[j] == 1u;
            const u32 idx = is_near ? on : of;
            on += is_near ? 1u : 0u;
            of += is_near ? 0u : 1u;
            if (is_near && idx < t.scap) { t.sv[idx] = pv[j]; t.sd[idx] = pd[j]; }
            else if (idx < (is_near ? t.near_lim : t.fcap)) {
                const u32 off = (is_near ? t.near_off : t.far_off) + idx;
                __stcs(t.wv + off, pv[j]);
                __stcs(t.wd + off, pd[j]);
            } else *overflow = 1u;
        }
    }
    c_push += (packed & 0xFFFFu) + (packed >> 16);
}

template <int BLOCK>
__global__ void __launch_bounds__(BLOCK, (BLOCK <= 32 ? 32 : BLOCK <= 512 ? 1024 / BLOCK : 1)) sssp_nearfar_kernel(SsspParams p)
{
    static_assert(BLOCK % 32 == 0, "whole warps");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const u32 scap = p.near_scap, gcap = p.near_gcap, fcap = p.far_cap;
    u64 *sd0 = reinterpret_cast<u64 *>(smem_raw);
    u64 *sd1 = sd0 + scap;
    u32 *sv0 = reinterpret_cast<u32 *>(sd1 + scap);
    u32 *sv1 = sv0 + scap;
    __shared__ u32 s_cnt[3];       // near queue fill counters, rotating (one barrier per round)
    __shared__ u32 s_far_cnt[2];
    __shared__ u32 s_far_min[2];   // min high word among the entries of each far pile
    __shared__ int s_root_idx;
    __shared__ u32 s_overflow;
    __shared__ u32 s_changed;
    __shared__ u32 s_stat[3];           // bucket advances, relaxation rounds, fallback sweeps (thread 0 counts)
    __shared__ u32 s_bucket_rounds;     // relaxation rounds since the last bucket advance
    __shared__ u32 s_bucket_entries;    // entries popped in them

    const u32 tid = threadIdx.x;
    const u32 lane = tid & 31u;
    const u32 V = p.V;
    const size_t ws_stride = (size_t)2 * gcap + (size_t)2 * fcap;
    u32 *wv = p.ws_v + (size_t)blockIdx.x * ws_stride;
    u64 *wd = p.ws_d + (size_t)blockIdx.x * ws_stride;
    // queue b / far pile f by arithmetic (no dynamically indexed local arrays)
    // pushes of a round go to near queue nb (shared part + spill) and far pile nf_ (32-bit workspace offsets)
    auto target = [&](u32 nb, u32 nf_) {
        return PushTarget{sv0 + nb * scap, sd0 + nb * scap, wv, wd, nb * gcap - scap, 2u * gcap + nf_ * fcap, scap, scap + gcap, fcap};
    };

    double delta = p.delta_scale * (*p.wsum);
    if (!(delta > 0.0)) delta = 0.0;

    u64 c_relax = 0, c_pop = 0, c_stale = 0, c_push = 0;
    if (tid == 0) { s_stat[0] = 0u; s_stat[1] = 0u; s_stat[2] = 0u; }

    for (;;) {
        __syncthreads();                       // previous tree fully retired before s_* are reused
        if (tid == 0) s_root_idx = (int)atomicAdd(p.work_counter, 1u);
        __syncthreads();
        const int ri = s_root_idx;
        if (ri >= p.n_roots) break;
        u64 *dist = p.dist_pool + (size_t)ri * p.slot_stride;
        const u32 root = (u32)p.roots[ri];

        // ---- fill the slot with +inf (16-byte stores)
        {
            ulonglong2 *d2 = reinterpret_cast<ulonglong2 *>(dist);
            const u32 n2 = (u32)(p.slot_stride >> 1);
            const ulonglong2 inf2 = make_ulonglong2(kInfBits, kInfBits);
            for (u32 i = tid; i < n2; i += BLOCK) __stcs(d2 + i, inf2);      // streaming: not needed again before the wave arrives
        }
        if (tid == 0) {
            s_cnt[0] = 1u; s_cnt[1] = 0u; s_cnt[2] = 0u;
            s_far_cnt[0] = 0u; s_far_cnt[1] = 0u;
            s_far_min[0] = 0xFFFFFFFFu; s_far_min[1] = 0xFFFFFFFFu;
            s_overflow = 0u;
            s_bucket_rounds = 0u; s_bucket_entries = 0u;
            sv0[0] = root | kNoSkip; sd0[0] = 0ull;
        }
        __syncthreads();
        if (tid == 0) dist[root] = 0ull;                   // after the fill
        __syncthreads();

        double dmul = 1.0;  // adaptive multiplier of the bucket width
        u32 c = 0;          // index of the current near counter
        u32 b = 0;          // current near buffer
        u32 f = 0;          // current far pile (push target)
        u32 thr_hi;                                            // near bucket = entries with hi32(d) <= thr_hi
        {
            double delta = p.delta_scale * __ldg(p.wsum);
            thr_hi = hi32((u64)__double_as_longlong(delta > 0.0 ? delta : 0.0));
        }

        u32 rib = 0;                // rounds in the current bucket
        bool force_round = false;   // an early advance found the far pile empty: run the round instead
        for (;;) {
            const u32 n = s_cnt[c];
            // Overlapped buckets: the last rounds of a bucket are a thinning cascade that leaves most of the CTA idle.
            // Once a round is that small, the next bucket is admitted on top of it (the split appends to the queue the
            // round would have read).  The threshold still moves one bucket width per advance, so there are no more
            // split passes than before, only fewer near-empty rounds.
            // (n, rib and force_round are CTA-uniform: s_cnt[c] is not written during the round that follows.)
            const bool early = n > 0u && n < p.early_advance && rib >= 2u && !force_round;
            if (n > 0 && !early) {
                ++rib;
                force_round = false;
                // ------------------------- relaxation round -------------------------
                const u32 cn = c == 2 ? 0 : c + 1;
                if (tid == 0) {
                    s_cnt[cn == 2 ? 0 : cn + 1] = 0u;                  // counter of the round after next
                    s_bucket_rounds += 1u;
                    s_bucket_entries += n;
                    s_stat[1] += 1u;
                }
                const u32 *const csv = sv0 + b * scap;                      // the queue this round reads
                const u64 *const csd = sd0 + b * scap;
                const u32 cur_off = b * gcap - scap;                        // its spill: entry i >= scap at cur_off + i
                const PushTarget tg = target(b ^ 1u, f);
                const u32 n_s = n < scap + gcap ? n : scap + gcap;      // entries that were actually stored
                for (u32 base = 0; base < n_s; base += BLOCK) {
                    const u32 i = base + tid;
                    bool valid = i < n_s;
                    u32 u = 0;
                    u64 du = 0;
                    u32 kind[kEllWidth], pv[kEllWidth];
                    u64 pd[kEllWidth];
#pragma unroll
                    for (int j = 0; j < kEllWidth; ++j) { kind[j] = 0u; pv[j] = 0u; pd[j] = 0ull; }
                    bool ovf = false;
                    if (valid) {
                        if (i < scap) { u = csv[i]; du = csd[i]; }
                        else { u = wv[cur_off + i]; du = wd[cur_off + i]; }
                        valid = expand_entry(p.ell, dist, u, du, thr_hi, p.preread, kind, pv, pd, ovf);
                        if (valid) { ++c_pop; c_relax += kEllWidth; } else ++c_stale;
                    }
                    warp_push<kEllWidth>(kind, pv, pd, lane, tg, &s_cnt[cn], &s_far_cnt[f], &s_far_min[f], &s_overflow, c_push);
                    // nodes with more than kEllWidth arcs: the rest of the CSR row, one arc per step
                    if (__any_sync(0xffffffffu, ovf)) {
                        u32 e = 0, re = 0;
                        if (ovf) { e = __ldg(p.row_ptr + (u & kNodeMask)) + kEllWidth; re = __ldg(p.row_ptr + (u & kNodeMask) + 1); }
                        while (__any_sync(0xffffffffu, e < re)) {
                            u32 k1[1] = {0u}, v1[1] = {0u};
                            u64 d1[1] = {0ull};
                            if (e < re) {
                                const Arc x = ld_arc(p.arcs + e);
                                v1[0] = x.col | kNoSkip;
                                d1[0] = (u64)__double_as_longlong(__dadd_rn(__longlong_as_double((long long)du), x.w));
                                if (d1[0] < ld_dist(dist + x.col)) {
                                    const u64 o2 = atomicMin(dist + x.col, d1[0]);
                                    if (d1[0] < o2) k1[0] = hi32(d1[0]) <= thr_hi ? 1u : 2u;
                                }
                                ++c_relax;
                                ++e;
                            }
                            warp_push<1>(k1, v1, d1, lane, tg, &s_cnt[cn], &s_far_cnt[f], &s_far_min[f], &s_overflow, c_push);
                        }
                    }
                }
                __syncthreads();
                c = cn;
                b ^= 1u;
                continue;
            }
            // ------------------------- near queue empty (or thin): advance the bucket -------------------------
            // every thread must have read s_cnt[c] before the split starts refilling that queue; past this barrier no
            // push is in flight, so the far pile's counters are stable
            __syncthreads();
            const u32 nf_raw = s_far_cnt[f];
            if (nf_raw == 0u) {
                if (n == 0u) break;                                     // tree complete
                force_round = true;                                     // nothing to admit yet
                continue;
            }
            const u32 nf = nf_raw < fcap ? nf_raw : fcap;
            const u32 far_min_hi = s_far_min[f];
            rib = 0;
            if (tid == 0) s_stat[0] += 1u;
            // new threshold: smallest (lower bound of a) distance in the pile + delta
            const double lo = __longlong_as_double((long long)((u64)far_min_hi << 32));
            // bucket width follows the wave: widen while the rounds of a bucket leave most of the CTA idle,
            // narrow again when a bucket floods it (priority order pays off once there is enough parallelism)
            if (p.adaptive) {
                const u32 br = s_bucket_rounds, be = s_bucket_entries;
                if (br > 0u) {
                    if (be < br * (BLOCK / 2u) && dmul < 16.0) dmul *= 1.5;
                    else if (be > br * (BLOCK * 4u) && dmul > 1.0) dmul *= (1.0 / 1.5);
                }
            }
            {
                double delta = p.delta_scale * __ldg(p.wsum);
                if (!(delta > 0.0)) delta = 0.0;
                thr_hi = hi32((u64)__double_as_longlong(__dadd_rn(lo, delta * dmul)));
            }
            const u32 fo = f ^ 1u;
            const u32 src_off = 2u * gcap + f * fcap;
            const PushTarget tg = target(b, fo);
            for (u32 base = 0; base < nf; base += BLOCK) {
                const u32 i = base + tid;
                u32 k1[1] = {0u}, v1[1] = {0u};
                u64 d1[1] = {0ull};
                if (i < nf) {
                    v1[0] = __ldcs(wv + src_off + i);
                    d1[0] = __ldcs(wd + src_off + i);
                    k1[0] = hi32(d1[0]) <= thr_hi ? 1u : 2u;
                }
                // the round after this reads (b, c); kept entries go to the other pile (never more than nf)
                warp_push<1>(k1, v1, d1, lane, tg, &s_cnt[c], &s_far_cnt[fo], &s_far_min[fo], &s_overflow, c_push);
            }
            __syncthreads();
            if (tid == 0) { s_far_cnt[f] = 0u; s_far_min[f] = 0xFFFFFFFFu; s_bucket_rounds = 0u; s_bucket_entries = 0u; }
            f = fo;
            // no barrier needed: pile f^1 is not a push target until the next advance,
            // and a relaxation round (with its barrier) always intervenes.
        }

        // ---- safety net: a queue overflowed and entries were dropped.  Sweep all
        // nodes Bellman-Ford style until nothing changes (same fixed point).
        if (s_overflow) {
            for (;;) {
                __syncthreads();
                if (tid == 0) { s_changed = 0u; s_stat[2] += 1u; }
                __syncthreads();
                bool any = false;
                for (u32 u = tid; u < V; u += BLOCK) {
                    const u64 du = ld_dist(dist + u);
                    if (du >= kInfBits) continue;
                    const u32 rb = __ldg(p.row_ptr + u), re = __ldg(p.row_ptr + u + 1);
                    for (u32 e = rb; e < re; ++e) {
                        const Arc x = ld_arc(p.arcs + e);
                        const u64 ndb = (u64)__double_as_longlong(__dadd_rn(__longlong_as_double((long long)du), x.w));
                        if (ndb < ld_dist(dist + x.col)) {
                            atomicMin(dist + x.col, ndb);
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
    __syncthreads();
    u64 vals[7] = {c_relax, c_pop, c_stale, c_push, tid == 0 ? s_stat[0] : 0u, tid == 0 ? s_stat[1] : 0u, tid == 0 ? s_stat[2] : 0u};
    const int ids[7] = {C_ARCS_RELAXED, C_NODES_POPPED, C_STALE_POPS, C_PUSHES, C_ADVANCES, C_ROUNDS, C_FALLBACK_SWEEPS};
#pragma unroll
    for (int j = 0; j < 7; ++j) {
        u64 v = vals[j];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0 && v) atomicAdd(p.counters + ids[j], v);
    }
}

// ---------------------------------------------------------------------------
// Canonical predecessor of `cur` from converged distances (the rule of oracle/port.py canonical_tree and
// oracle/c/s4_oracle.c canon_pred_root), evaluated by the lanes of one tile (lane k looks at arcs rb+k, rb+k+TILE, ...).
// Candidates: arcs (cur,u), u != cur, dist[u] finite, fl(dist[u]+w) == dist[cur].
//   - some candidate strictly closer to the root (dist[u] < dist[cur]): min (id(u), street) among those;
//   - else `cur` sits on a plateau (zero / absorbed weights): bounded breadth-first search over the plateau to the
//     nearest node that has a strictly closer candidate (or the root), first hop of that path (plateau_first_hop).
// id() is the caller's node id: the device numbers nodes along a space-filling curve (locality), `orig` maps that
// numbering back so that the tie-break is the one the host sees (nullptr = identical numbering).
// Returns the packed key (id(u) << 32 | street), ~0 if there is none; *col_out is the device number of the winner
// and *du_out its distance (valid on every lane).
// ---------------------------------------------------------------------------
static constexpr int kPlateauMax = 32;     // nodes a plateau search may touch (oracle: PLATEAU_MAX)

// smallest candidate of x above `after` in (id(u), street) order (has_after = false: the smallest of all);
// closer_only: dist[u] < dist[x].  Sequential over the CSR row: the slow path behind zero-weight streets.
__device__ __noinline__ u64 next_candidate(const u32 *__restrict__ row_ptr, const Arc *__restrict__ arcs, const u32 *__restrict__ orig,
                                           const u64 *dist, u32 stride, u32 x, bool has_after, u64 after, bool closer_only, u32 *col_out)
{
    const u64 dx = ld_dist(dist + (size_t)x * stride);
    u64 best = ~0ull;
    u32 bcol = 0;
    for (u32 e = __ldg(row_ptr + x), re = __ldg(row_ptr + x + 1); e < re; ++e) {
        const Arc a = ld_arc(arcs + e);
        if (a.col == x) continue;
        const u64 du = ld_dist(dist + (size_t)a.col * stride);
        if (du >= kInfBits) continue;
        if ((u64)__double_as_longlong(__dadd_rn(__longlong_as_double((long long)du), a.w)) != dx) continue;
        if (closer_only && !(du < dx)) continue;
        const u64 key = ((u64)(orig ? __ldg(orig + a.col) : a.col) << 32) | a.street;
        if (has_after && key <= after) continue;
        if (key < best) { best = key; bcol = a.col; }
    }
    *col_out = bcol;
    return best;
}

// every lane of the tile runs the same search on the same data (uniform, the loads broadcast).  The result comes
// back by value: an out-pointer into the caller's frame would force the walk's loop variables into local memory.
struct PlateauHop {
    u64 key;    // (id(u) << 32 | street) of the first hop, ~0: none (plateau larger than the bound)
    u32 col;    // its device number
};
__device__ __noinline__ PlateauHop plateau_first_hop(const u32 *__restrict__ row_ptr, const Arc *__restrict__ arcs, const u32 *__restrict__ orig,
                                                     const u64 *dist, u32 stride, u32 v, u32 root)
{
    u32 seen[kPlateauMax], first_col[kPlateauMax];
    u64 first_key[kPlateauMax];
    int n_seen = 1;
    seen[0] = v; first_col[0] = 0u; first_key[0] = ~0ull;
    for (int head = 0; head < n_seen; ++head) {
        const u32 x = seen[head];
        bool has = false;
        u64 after = 0ull;
        for (;;) {
            u32 cu;
            const u64 key = next_candidate(row_ptr, arcs, orig, dist, stride, x, has, after, false, &cu);
            if (key == ~0ull) break;
            has = true;
            after = key;
            bool dup = false;
            for (int k = 0; k < n_seen; ++k) dup = dup || seen[k] == cu;
            if (dup) continue;
            if (n_seen >= kPlateauMax) return PlateauHop{~0ull, 0u};
            seen[n_seen] = cu;
            first_key[n_seen] = head == 0 ? key : first_key[head];
            first_col[n_seen] = head == 0 ? cu : first_col[head];
            u32 dummy;
            if (cu == root || next_candidate(row_ptr, arcs, orig, dist, stride, cu, false, 0ull, true, &dummy) != ~0ull)
                return PlateauHop{first_key[n_seen], first_col[n_seen]};
            ++n_seen;
        }
    }
    return PlateauHop{~0ull, 0u};
}

// DEFER_PLATEAU: do not search, return kPlateauKey (the hot walk kernel hands such trips to walk_slow_kernel: a call
// to the search would raise the register count of the whole kernel to the callee's).
static constexpr u64 kPlateauKey = ~1ull;

template <int TILE, bool WANT_ID, bool DEFER_PLATEAU = false>
__device__ __forceinline__ u64 canonical_pred(const cg::thread_block_tile<TILE> &tile, const u32 *__restrict__ row_ptr,
                                              const Arc *__restrict__ arcs, const u32 *__restrict__ orig, const u64 *dist,
                                              u32 stride, u32 cur, u64 dv, u32 root, u32 *col_out, u64 *du_out)
{
    const u32 rb = __ldg(row_ptr + cur), re = __ldg(row_ptr + cur + 1);
    // pass 1: which arcs reproduce dist[cur] exactly, and from a strictly smaller distance?  Almost always exactly
    // one (unique shortest path): then the ids do not matter and no id lookup is needed.
    u64 du1 = 0;
    u32 street1 = 0, col1 = 0, n_cand = 0, n_close = 0;
    for (u32 e = rb + tile.thread_rank(); tile.any(e < re); e += TILE) {
        if (e < re) {
            const Arc a = ld_arc(arcs + e);
            if (a.col != cur) {
                const u64 du = ld_dist(dist + (size_t)a.col * stride);
                const u64 ndb = (u64)__double_as_longlong(__dadd_rn(__longlong_as_double((long long)du), a.w));
                if (du < kInfBits && ndb == dv) {
                    ++n_cand;
                    if (du < dv) { ++n_close; street1 = a.street; du1 = du; col1 = a.col; }
                }
            }
        }
    }
    u32 total = n_cand | (n_close << 16);
    for (int o = TILE / 2; o > 0; o >>= 1) total += tile.shfl_xor(total, o);
    const u32 closer = total >> 16;
    if (closer == 1u) {
        // broadcast the single closer candidate from the lane that holds it
        const u32 src = (u32)__ffs((int)tile.ballot(n_close != 0u)) - 1u;
        const u32 street = tile.shfl(street1, src);
        *col_out = tile.shfl(col1, src);
        *du_out = tile.shfl(du1, src);
        const u32 id_u = !WANT_ID ? 0u : orig ? __ldg(orig + *col_out) : *col_out;     // the walk only needs the street
        return ((u64)id_u << 32) | street;
    }
    if ((total & 0xFFFFu) == 0u) { *du_out = 0; *col_out = 0; return ~0ull; }
    if (closer == 0u) {
        // plateau: every candidate is at dist[cur]
        if constexpr (DEFER_PLATEAU) { *du_out = dv; *col_out = cur; return kPlateauKey; }
        const PlateauHop h = plateau_first_hop(row_ptr, arcs, orig, dist, stride, cur, root);
        *du_out = dv;
        *col_out = h.col;
        return h.key;
    }
    // several closer candidates: smallest (id, street)
    u64 best = ~0ull, du_b = 0;
    u32 col_b = 0;
    for (u32 e = rb + tile.thread_rank(); tile.any(e < re); e += TILE) {
        if (e < re) {
            const Arc a = ld_arc(arcs + e);
            if (a.col != cur) {
                const u64 du = ld_dist(dist + (size_t)a.col * stride);
                const u64 ndb = (u64)__double_as_longlong(__dadd_rn(__longlong_as_double((long long)du), a.w));
                if (du < dv && ndb == dv) {
                    const u64 key = ((u64)(orig ? __ldg(orig + a.col) : a.col) << 32) | a.street;
                    if (key < best) { best = key; du_b = du; col_b = a.col; }
                }
            }
        }
    }
    for (int o = TILE / 2; o > 0; o >>= 1) {
        const u64 kb = tile.shfl_xor(best, o), db = tile.shfl_xor(du_b, o);
        const u32 cb = tile.shfl_xor(col_b, o);
        if (kb < best) { best = kb; du_b = db; col_b = cb; }
    }
    *du_out = du_b;
    *col_out = col_b;
    return best;
}

// The same rule on the fixed-stride rows (ellw: first kEllWidth arcs of every node, lane k of the tile looks at slot
// k): one dependent memory stage less per hop than the CSR (no row_ptr lookup).  Nodes with more arcs fall back to
// canonical_pred.  Padding slots point back at `cur` and are skipped like self loops.
template <int TILE, bool WANT_ID, bool DEFER_PLATEAU = false>
__device__ __forceinline__ u64 canonical_pred_ell(const cg::thread_block_tile<TILE> &tile, const Arc *__restrict__ ellw,
                                                  const u32 *__restrict__ row_ptr, const Arc *__restrict__ arcs,
                                                  const u32 *__restrict__ orig, const u64 *dist, u32 stride, u32 cur, u64 dv,
                                                  u32 root, u32 *col_out, u64 *du_out, u32 skip = ~0u)
{
    // skip: a neighbour known to be strictly farther from the root than `cur` (the node a walk just came from through a
    // closer-candidate hop): it cannot reproduce dist[cur], its distance is not read (one DRAM sector less per hop)
    static_assert(TILE >= kEllWidth, "one lane per slot");
    const u32 lane = tile.thread_rank();
    Arc a;
    a.col = cur; a.street = 0u; a.w = 0.0;
    if (lane < (u32)kEllWidth) a = ld_arc(ellw + (size_t)cur * kEllWidth + lane);
    if (tile.any(lane == 0u && (a.col & kOvfBit) != 0u))
        return canonical_pred<TILE, WANT_ID, DEFER_PLATEAU>(tile, row_ptr, arcs, orig, dist, stride, cur, dv, root, col_out, du_out);
    const u32 col = a.col;
    u64 du = 0;
    bool match = false;
    if (lane < (u32)kEllWidth && col != cur && col != skip) {
        du = ld_dist(dist + (size_t)col * stride);
        const u64 ndb = (u64)__double_as_longlong(__dadd_rn(__longlong_as_double((long long)du), a.w));
        match = du < kInfBits && ndb == dv;
    }
    const bool close = match && du < dv;
    const u32 mc = tile.ballot(close);
    if (mc != 0u && (mc & (mc - 1u)) == 0u) {
        // exactly one arc reproduces dist[cur] from a smaller distance (unique shortest path): no id lookup needed
        const u32 src = (u32)__ffs((int)mc) - 1u;
        const u32 street = tile.shfl(a.street, src);
        *col_out = tile.shfl(col, src);
        *du_out = tile.shfl(du, src);
        const u32 id_u = !WANT_ID ? 0u : orig ? __ldg(orig + *col_out) : *col_out;
        return ((u64)id_u << 32) | street;
    }
    if (mc == 0u) {
        if (!tile.any(match)) { *du_out = 0; *col_out = 0; return ~0ull; }
        if constexpr (DEFER_PLATEAU) { *du_out = dv; *col_out = cur; return kPlateauKey; }
        const PlateauHop h = plateau_first_hop(row_ptr, arcs, orig, dist, stride, cur, root);   // plateau
        *du_out = dv;
        *col_out = h.col;
        return h.key;
    }
    // several closer candidates: smallest (id, street)
    u64 best = ~0ull, du_b = 0;
    u32 col_b = 0;
    if (close) { best = ((u64)(orig ? __ldg(orig + col) : col) << 32) | a.street; du_b = du; col_b = col; }
    for (int o = TILE / 2; o > 0; o >>= 1) {
        const u64 kb = tile.shfl_xor(best, o), db = tile.shfl_xor(du_b, o);
        const u32 cb = tile.shfl_xor(col_b, o);
        if (kb < best) { best = kb; du_b = db; col_b = cb; }
    }
    *du_out = du_b;
    *col_out = col_b;
    return best;
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
    const Arc *ellw;            // fixed-stride rows with street indices (nullptr: CSR only)
    const uint2 *ell_cs;        // static fixed-stride rows {head | overflow flag, street | reverse slot << 29} (walk_code_kernel)
    const u32 *orig;            // device number -> caller's node id (nullptr = identity)
    const u64 *dist_pool;       // first slot of this batch
    size_t slot_stride;         // words per slot (a slot holds the rows of `group` trees)
    u32 group;                  // trees per slot: tree i of the batch is word i % group of the rows of slot i / group
    u32 node_stride;            // words per row (1 = one tree per slot, plain dist[V])
    const int32_t *roots;       // distinct roots (all of them), device numbers
    const int32_t *trip_target; // trips sorted by root, device numbers
    const u32 *trip_root;       // index of the trip's root among the distinct roots
    int64_t t0, t1;             // trips of this batch
    u32 root0;                  // first root of this batch (slot = trip_root - root0)
    u32 V;
    u32 volume;
    u32 *load;
    u64 *counters;
    uint2 *deferred;            // {trip - t_base, node}: trips stopped on a zero-weight plateau (walk_slow_kernel)
    u32 *deferred_count;        // zeroed before the walk of a batch
    int64_t t_base;             // first trip of this batch (deferred entries are relative to it)
};

// one trip: walk from `cur` (distance dv) to the root on the converged distances of its tree (dist + node * stride), add
// `volume` per street.  DEFER: on a zero-weight plateau stop and return the node (the caller queues the trip for
// walk_slow_kernel); otherwise search the plateau in place.  Returns ~0u when the trip is finished.
template <int TILE, bool DEFER>
__device__ __forceinline__ u32 walk_trip(const cg::thread_block_tile<TILE> &tile, const Arc *__restrict__ ellw,
                                         const u32 *__restrict__ row_ptr, const Arc *__restrict__ arcs, const u32 *__restrict__ orig,
                                         const u64 *dist, u32 stride, u32 root, u32 cur, u64 dv, u32 V, u32 volume, u32 *load,
                                         u32 &c_hops, u32 &c_stuck)
{
    u32 guard = 0;
    u32 prev = ~0u;                  // DEFER: every hop so far went to a strictly closer node, so `prev` is strictly farther
    while (cur != root) {                                                  // simulation.py:83
        u64 du;
        u32 nxt;
        const u64 key = ellw ? canonical_pred_ell<TILE, false, DEFER>(tile, ellw, row_ptr, arcs, orig, dist, stride, cur, dv, root, &nxt, &du, prev)
                             : canonical_pred<TILE, false, DEFER>(tile, row_ptr, arcs, orig, dist, stride, cur, dv, root, &nxt, &du);
        if (DEFER && key == kPlateauKey) return cur;
        if (key == ~0ull || ++guard > V) { ++c_stuck; break; }
        if (tile.thread_rank() == 0) atomicAdd(load + (u32)key, volume);   // :86-88
        ++c_hops;
        if (DEFER) prev = cur;
        cur = nxt;                                                         // :85
        dv = du;
    }
    return ~0u;
}

template <int TILE>
__global__ void __launch_bounds__(256) walk_kernel(WalkParams p)
{
    cg::thread_block block = cg::this_thread_block();
    cg::thread_block_tile<TILE> tile = cg::tiled_partition<TILE>(block);
    const int64_t tiles_per_grid = (int64_t)gridDim.x * (blockDim.x / TILE);
    const int64_t tile_id = (int64_t)blockIdx.x * (blockDim.x / TILE) + threadIdx.x / TILE;
    u32 c_hops = 0, c_unreach = 0, c_stuck = 0, c_trips = 0;
    for (int64_t t = p.t0 + tile_id; t < p.t1; t += tiles_per_grid) {
        const u32 ri = p.trip_root[t];
        const u32 ti = ri - p.root0;
        const u64 *dist = p.dist_pool + (size_t)(ti / p.group) * p.slot_stride + (ti % p.group);
        const u32 target = (u32)p.trip_target[t];
        ++c_trips;
        const u64 dv = ld_dist(dist + (size_t)target * p.node_stride);
        if (dv >= kInfBits) { ++c_unreach; continue; }                     // simulation.py:80
        const u32 at = walk_trip<TILE, true>(tile, p.ellw, p.row_ptr, p.arcs, p.orig, dist, p.node_stride, (u32)p.roots[ri], target, dv,
                                             p.V, p.volume, p.load, c_hops, c_stuck);
        if (at != ~0u && tile.thread_rank() == 0) {
            // a zero-weight plateau: the rest of this trip is walked by walk_slow_kernel
            const u32 slot = atomicAdd(p.deferred_count, 1u);
            p.deferred[slot] = make_uint2((u32)(t - p.t_base), at);
        }
        if (c_hops > 0x40000000u) { if (tile.thread_rank() == 0) atomicAdd(p.counters + C_HOPS, (u64)c_hops); c_hops = 0; }
    }
    if (tile.thread_rank() == 0) {
        if (c_trips) atomicAdd(p.counters + C_TRIPS, (u64)c_trips);
        if (c_hops) atomicAdd(p.counters + C_HOPS, (u64)c_hops);
        if (c_unreach) atomicAdd(p.counters + C_UNREACHABLE, (u64)c_unreach);
        if (c_stuck) atomicAdd(p.counters + C_STUCK, (u64)c_stuck);
    }
}

// The trips walk_kernel stopped on a zero-weight plateau, continued with the full rule (bounded plateau search).  Rare
// (zero-length streets); launched behind every walk_kernel on the same stream, a no-op when the list is empty.
template <int TILE>
__global__ void __launch_bounds__(256) walk_slow_kernel(WalkParams p)
{
    cg::thread_block block = cg::this_thread_block();
    cg::thread_block_tile<TILE> tile = cg::tiled_partition<TILE>(block);
    const u32 tiles_per_grid = gridDim.x * (blockDim.x / TILE);
    const u32 n = *p.deferred_count;
    u32 c_hops = 0, c_stuck = 0;
    for (u32 i = blockIdx.x * (blockDim.x / TILE) + threadIdx.x / TILE; i < n; i += tiles_per_grid) {
        const uint2 d = p.deferred[i];
        const int64_t t = p.t_base + (int64_t)d.x;
        const u32 ri = p.trip_root[t];
        const u32 ti = ri - p.root0;
        const u64 *dist = p.dist_pool + (size_t)(ti / p.group) * p.slot_stride + (ti % p.group);
        const u64 dv = ld_dist(dist + (size_t)d.y * p.node_stride);
        walk_trip<TILE, false>(tile, p.ellw, p.row_ptr, p.arcs, p.orig, dist, p.node_stride, (u32)p.roots[ri], d.y, dv, p.V, p.volume,
                               p.load, c_hops, c_stuck);
    }
    if (tile.thread_rank() == 0) {
        if (c_hops) atomicAdd(p.counters + C_HOPS, (u64)c_hops);
        if (c_stuck) atomicAdd(p.counters + C_STUCK, (u64)c_stuck);
    }
}

// predecessor map and distances of one tree in the caller's numbering
// (StreetNetwork.calculate_shortest_paths, streetnetwork.py:108-109)
template <int TILE>
__global__ void __launch_bounds__(256) pred_kernel(const u32 *row_ptr, const Arc *arcs, const u32 *orig, const u64 *dist,
                                                   u32 stride, u32 V, u32 root, int32_t *pred_out, double *dist_out)
{
    cg::thread_block block = cg::this_thread_block();
    cg::thread_block_tile<TILE> tile = cg::tiled_partition<TILE>(block);
    const u32 tiles_per_grid = gridDim.x * (blockDim.x / TILE);
    for (u32 v = blockIdx.x * (blockDim.x / TILE) + threadIdx.x / TILE; v < V; v += tiles_per_grid) {
        const u64 dv = ld_dist(dist + (size_t)v * stride);
        int32_t pr = -1;
        if (v != root && dv < kInfBits) {
            u64 du;
            u32 nxt;
            const u64 key = canonical_pred<TILE, true>(tile, row_ptr, arcs, orig, dist, stride, v, dv, root, &nxt, &du);
            if (key != ~0ull) pr = (int32_t)(key >> 32);
        }
        if (tile.thread_rank() == 0) {
            const u32 id_v = orig ? orig[v] : v;
            pred_out[id_v] = pr;
            dist_out[id_v] = __longlong_as_double((long long)dv);
        }
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

// ---------------------------------------------------------------------------
// Trip sampling on the device (tripgenerator.py:33-35 draws one origin and one goal per resident): trip t takes the
// four words of Philox4x32-10 at counter (t, 0) under the key `seed` -- words 0,1 pick the origin, words 2,3 the
// goal, index = floor(r64 * n / 2^64).  Counter based, so the list does not depend on the launch shape;
// tripgenerator.philox_trip_indices is the same function in numpy.
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const u32 hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const u32 hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}

__global__ void sample_trips_kernel(int64_t T, u64 seed, const int32_t *__restrict__ pop_root, u64 n_root, int root_words_hi,
                                    const int32_t *__restrict__ pop_target, u64 n_target, u32 *__restrict__ root_out,
                                    u32 *__restrict__ target_out)
{
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < T; t += (int64_t)gridDim.x * blockDim.x) {
        const uint4 r = philox4x32_10(make_uint4((u32)t, (u32)((u64)t >> 32), 0u, 0u), make_uint2((u32)seed, (u32)(seed >> 32)));
        const u64 r_origin = ((u64)r.x << 32) | r.y, r_goal = ((u64)r.z << 32) | r.w;
        // root_words_hi: the roots are the goals (words 2,3), the targets the origins (words 0,1)
        const u64 rr = root_words_hi ? r_goal : r_origin, rt = root_words_hi ? r_origin : r_goal;
        root_out[t] = (u32)pop_root[__umul64hi(rr, n_root)];
        target_out[t] = (u32)pop_target[__umul64hi(rt, n_target)];
    }
}

// trip grouping: after sorting trips by root, mark the first trip of every root
__global__ void head_flag_kernel(int64_t T, const int32_t *__restrict__ sorted_root, u32 *__restrict__ flag)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < T; i += (int64_t)gridDim.x * blockDim.x)
        flag[i] = (i == 0 || sorted_root[i] != sorted_root[i - 1]) ? 1u : 0u;
}

// rank[i] = inclusive_scan(flag)[i]; root index = rank-1; heads record root node and first-trip offset
// (sorted_root holds positions in the grouping order; ginv maps a position back to the node)
__global__ void group_scatter_kernel(int64_t T, const int32_t *__restrict__ sorted_root, const u32 *__restrict__ flag,
                                     const u32 *__restrict__ rank, const u32 *__restrict__ ginv, u32 *__restrict__ trip_root,
                                     int32_t *__restrict__ roots, int64_t *__restrict__ root_off)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < T; i += (int64_t)gridDim.x * blockDim.x) {
        const u32 r = rank[i] - 1u;
        trip_root[i] = r;
        if (flag[i]) { roots[r] = (int32_t)ginv[sorted_root[i]]; root_off[r] = i; }
    }
}

// trips from the canonical order (sorted by root) into the order of the root slots of the current grouping
__global__ void regroup_trips_kernel(int64_t T, const u32 *__restrict__ trip_root0, const int32_t *__restrict__ trip_target0,
                                     const int64_t *__restrict__ root_off0, const u32 *__restrict__ slot_of,
                                     const int64_t *__restrict__ new_start, u32 *__restrict__ trip_root, int32_t *__restrict__ trip_target)
{
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < T; t += (int64_t)gridDim.x * blockDim.x) {
        const u32 r = trip_root0[t];
        const int64_t dst = new_start[r] + (t - root_off0[r]);
        trip_root[dst] = slot_of[r];
        trip_target[dst] = trip_target0[t];
    }
}

// x[i] = table[x[i]]
__global__ void gather_u32_kernel(int64_t n, u32 *__restrict__ x, const u32 *__restrict__ table)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        x[i] = table[x[i]];
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
