// K2, source-batched: K shortest-path trees per CTA on one shared frontier.
//   simulation.py:71-75 -> streetnetwork.py:108-109 (one Dijkstra per distinct origin);
//   north_star item (2): "batched multi-source SSSP".
//
// Why.  All trees of a day run on the same weights, and the roots of a day, sorted along the space-filling
// curve the nodes are numbered by, are neighbours.  Far from the roots their shortest-path trees coincide
// and their waves pass a node within a few bucket widths of each other.  So K neighbouring roots share ONE
// frontier: an entry is a node, and expanding it relaxes its arcs for all K trees at once.  One 64-byte arc
// row, one queue pop, one push serve K trees; per (node, tree) the kernel issues ~K times fewer
// instructions and L1 wavefronts than the one-tree-per-CTA kernel it replaces (sssp_nearfar_kernel).
//
// Layout.  A group owns a slot of V rows of RW = K+1 words (u64): row v = { dist_0[v] .. dist_{K-1}[v], meta[v] }.
// RW*8 = 64 or 128 bytes, so the LP lanes that expand a node read a node's row as ONE coalesced access
// (lane t holds words [t*TL, t*TL+TL)); the low half of meta[v] is the id of the round v was last queued for
// (de-duplicates the pushes of the K trees and of a node's several upstream neighbours).
//
// Algorithm.  Near-far (delta-stepping with one near bucket and a far pile) on the key min_k dist_k[v] (+ lag):
// label-correcting, every improvement of any (v, k) queues v (unless v is already queued for the coming round),
// an expansion relaxes all K trees with the values it finds.  The fixed point is schedule-independent and equal
// to Dijkstra's for every tree: dist_k[v] = min_u fl(dist_k[u] + w(u,v)), sums rounded left to right.
// Distances are non-negative doubles compared and atomically minimised on their u64 bit patterns.
//
// Lag.  With lag_factor > 0 the key of v is min_k dist_k[v] + L, L = lag_factor * (largest finite distance between
// two roots of the group so far): v is expanded once the slowest tree's wave has had time to arrive too.
// No result depends on it.
//
// Clusters (CL > 1).  A slot is V * RW * 8 bytes: on graphs with 10^7 .. 10^8 nodes only a few dozen fit in HBM, fewer
// than there are SMs.  Then the CL CTAs of a thread-block cluster share ONE group: its frontier queues and far piles
// live in the cluster's global workspace, the queue counters in a small control block in global memory, every round
// ends with a cluster barrier instead of __syncthreads, and the entries of a round / the far pile / the slot fill
// are dealt out to the CTAs in strides.  CL = 1 is the plain one-CTA-per-group kernel (shared-memory counters and
// queue heads; the code paths below are selected at compile time).
#pragma once

#include "s4_kernels.cuh"

namespace s4 {

struct GroupParams {
    const Arc *ellw;        // V * kEllWidth rows {head | overflow flag, street, w}; padding slots {self, -, +inf}
    const u32 *row_ptr;     // V+1 (overflow arcs, fallback sweeps)
    const Arc *arcs;        // A
    u64 *pool;              // first group slot of this launch
    size_t slot_stride;     // words per group slot = V * RW
    const int32_t *roots;   // root slots of this batch; group g owns slots [g*K, (g+1)*K), -1 = empty (always at the end of a group)
    int n_roots;            // slots, a multiple of K unless the caller hands in a plain list of roots
    int n_groups;
    u32 V;
    u32 *work_counter;      // zeroed before launch
    const double *wsum;     // sum of street weights (device scalar)
    double delta_scale;     // delta = delta_scale * wsum
    double lag_factor;
    u32 *ws_v;              // per CTA: [near0 | near1] gcap node entries, then [far0 | far1] fcap node entries
    u32 *ws_k;              // per CTA, same offsets: keys (high words) of the far entries
    u32 near_scap;          // entries per near queue in shared memory
    u32 near_gcap;          // spill entries per near queue in global memory
    u32 far_cap;            // entries per far pile
    u32 adaptive;           // 1: the bucket width follows the width of the wave
    u32 early_advance;      // admit the next bucket when a round (not the first two of its bucket) has fewer entries
    u64 *counters;
    u32 *group_log;         // diagnostics (may be null): per group {SM cycles >> 10, rounds, entries popped, bucket advances}
    u32 *cluster_ctl;       // CL > 1: 16 control words per cluster (queue counters, overflow flag, current group)
};

// control words of a cluster (CL > 1)
enum GroupCtl { GC_CNT = 0, GC_FAR_CNT = 3, GC_FAR_MIN = 5, GC_OVERFLOW = 7, GC_GROUP = 8, GC_CHANGED = 9, GC_WORDS = 16 };

template <int TL>
__device__ __forceinline__ void ld_words(const u64 *p, u64 (&x)[TL])
{
    if constexpr (TL == 1) {
        x[0] = __ldcg(p);
    } else if constexpr (TL == 2) {
        asm volatile("ld.global.cg.v2.u64 {%0,%1}, [%2];" : "=l"(x[0]), "=l"(x[1]) : "l"(p));
    } else {
        static_assert(TL == 4, "one, two or four words per lane");
        asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(x[0]), "=l"(x[1]), "=l"(x[2]), "=l"(x[3]) : "l"(p));   // LDG.E.256
    }
}

__device__ __forceinline__ u32 sel4(u32 j, u32 a, u32 b, u32 c, u32 d)
{
    const u32 lo = (j & 1u) ? b : a, hi = (j & 1u) ? d : c;
    return (j & 2u) ? hi : lo;
}

struct GroupPushTarget {
    u32 *sv;                // next near queue, shared-memory part (scap entries)
    u32 *wv, *wk;           // this CTA's workspace
    u32 near_off;           // workspace offset of that queue's spill, minus scap
    u32 far_off;            // workspace offset of the far pile
    u32 scap, near_lim, fcap;
};

// Warp-cooperative append: kind 1 = node pv to the near queue, 2 = (pv, key pk) to the far pile.  One ballot per
// queue, one shared-memory atomic per warp and queue.
__device__ __forceinline__ void warp_push_g(u32 kind, u32 pv, u32 pk, u32 lane, const GroupPushTarget &t, u32 *near_cnt,
                                            u32 *far_cnt, u32 *far_min, u32 *overflow, u32 &c_push)
{
    const u32 m_near = __ballot_sync(0xffffffffu, kind == 1u), m_far = __ballot_sync(0xffffffffu, kind == 2u);
    if ((m_near | m_far) == 0u) return;                        // warp-uniform
    const u32 lt = (1u << lane) - 1u;
    if (m_near) {
        const u32 leader = (u32)__ffs((int)m_near) - 1u;
        u32 base = 0;
        if (lane == leader) base = atomicAdd(near_cnt, (u32)__popc(m_near));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (kind == 1u) {
            const u32 idx = base + (u32)__popc(m_near & lt);
            if (idx < t.scap) t.sv[idx] = pv;
            else if (idx < t.near_lim) __stcs(t.wv + (t.near_off + idx), pv);      // 32-bit sum: near_off is biased by -scap
            else *overflow = 1u;
        }
    }
    if (m_far) {
        const u32 leader = (u32)__ffs((int)m_far) - 1u;
        const u32 wmin = __reduce_min_sync(0xffffffffu, kind == 2u ? pk : 0xFFFFFFFFu);
        u32 base = 0;
        if (lane == leader) { base = atomicAdd(far_cnt, (u32)__popc(m_far)); atomicMin(far_min, wmin); }
        base = __shfl_sync(0xffffffffu, base, leader);
        if (kind == 2u) {
            const u32 idx = base + (u32)__popc(m_far & lt);
            if (idx < t.fcap) { __stcs(t.wv + (t.far_off + idx), pv); __stcs(t.wk + (t.far_off + idx), pk); }
            else *overflow = 1u;
        }
    }
    c_push += (kind != 0u);
}

// atomic minimum without a return value, predicated instead of branched around (RED.E.MIN.64)
__device__ __forceinline__ void red_min_if(u64 *addr, u64 val, bool pred)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p red.global.min.u64 [%0], %1;\n\t}" ::"l"(addr), "l"(val), "r"((u32)pred) : "memory");
}

// One lane's share of relaxing one arc (u -> v, weight w) for the trees it holds: my[] are u's distances (words this
// lane owns), dv[] v's; ok[] marks the trees whose distance at u lies inside the current bucket (only those are
// propagated: a larger value cannot be final yet, and the improvement that set it queued its own entry).  Sends the
// improvements as reductions; returns the lane's minimum over its (updated) words of v in mn, whether any of its
// trees improved, and whether an improved value lies beyond the bucket.
template <int TL>
__device__ __forceinline__ void relax_words(u64 *vrow, const u64 (&my)[TL], const bool (&ok)[TL], double w, const u64 (&dv)[TL],
                                            bool meta_lane, u32 thr_hi, u32 &mn, bool &im, bool &beyond)
{
    mn = 0x7FFFFFFFu;
    im = false;
    beyond = false;
#pragma unroll
    for (int q = 0; q < TL; ++q) {
        const bool is_meta = meta_lane && q == TL - 1;
        const u64 cand = ok[q] ? (u64)__double_as_longlong(__dadd_rn(__longlong_as_double((long long)my[q]), w)) : kInfBits;
        const bool better = cand < dv[q];                      // ok[] is false on the meta word
        red_min_if(vrow + q, cand, better);
        const u32 h = hi32(better ? cand : dv[q]);
        if (!is_meta) mn = min(mn, h);
        im = im || better;
        beyond = beyond || (better && hi32(cand) > thr_hi);
    }
}

// minimum over the LP lanes of a sub-warp of four per-lane values (one per arc slot), "transposed": on return lane t
// holds the full minimum for arc slot j = (LP == 8 ? t >> 1 : t) in the return value.  LP/2 + 1 shuffles instead of
// 4 * log2(LP).
template <int LP>
__device__ __forceinline__ u32 subwarp_min4(u32 x0, u32 x1, u32 x2, u32 x3, u32 t)
{
    static_assert(LP == 8 || LP == 4, "sub-warps of 4 or 8 lanes");
    constexpr u32 H = LP / 2;                                   // first exchange distance: 4 (LP 8) or 2 (LP 4)
    const bool up = (t & H) != 0u;
    // lanes in the upper half keep slots {2,3} and send {0,1}; the lower half the other way round
    const u32 r0 = __shfl_xor_sync(0xffffffffu, up ? x0 : x2, H);
    const u32 r1 = __shfl_xor_sync(0xffffffffu, up ? x1 : x3, H);
    const u32 y0 = min(up ? x2 : x0, r0), y1 = min(up ? x3 : x1, r1);
    const bool up2 = (t & (H >> 1)) != 0u;
    const u32 r = __shfl_xor_sync(0xffffffffu, up2 ? y0 : y1, H >> 1);
    u32 z = min(up2 ? y1 : y0, r);
    if constexpr (LP == 8) z = min(z, __shfl_xor_sync(0xffffffffu, z, 1));
    return z;
}

// CTAs per SM: 1024 threads per SM at 64 registers; with four words per lane a 512-thread CTA takes the SM alone (128
// registers: the distances of two neighbour rows alone are 16)
template <int BLOCK, int LP, int TL, int CL = 1>
__global__ void __launch_bounds__(BLOCK, (BLOCK <= 32 ? 32 : BLOCK <= 512 ? (TL >= 4 && BLOCK == 512 ? 1 : 1024 / BLOCK) : 1)) sssp_group_kernel(GroupParams p)
{
    static_assert(BLOCK % 32 == 0, "whole warps");
    constexpr int RW = LP * TL;           // words per row
    constexpr int K = RW - 1;             // trees per group; the last word of a row is meta
    constexpr u32 PW = BLOCK / LP;        // nodes expanded per pass
    constexpr u32 LPMASK = (1u << LP) - 1u;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const u32 scap = p.near_scap, gcap = p.near_gcap, fcap = p.far_cap;
    u32 *sv0 = reinterpret_cast<u32 *>(smem_raw);
    __shared__ u32 s_cnt[3];       // near queue fill counters, rotating (one barrier per round)
    __shared__ u32 s_far_cnt[2];
    __shared__ u32 s_far_min[2];   // min key among the entries of each far pile
    __shared__ int s_group;
    __shared__ u32 s_overflow;
    __shared__ u32 s_changed;
    __shared__ u32 s_stat[3];           // bucket advances, relaxation rounds, fallback sweeps (thread 0 counts)
    __shared__ u32 s_bucket_rounds;     // relaxation rounds since the last bucket advance
    __shared__ u32 s_bucket_entries;    // entries popped in them
    __shared__ u32 s_lag_hi;            // high word of the largest finite root-to-root distance seen so far
    __shared__ u32 s_roots[K];
    __shared__ u32 s_glog[3];           // diagnostics of the group: rounds, entries popped, bucket advances
    __shared__ long long s_t0;
    __shared__ double s_lag_scale;      // halved whenever the far pile runs half full (a long lag parks many nodes there)
    __shared__ double s_dmul, s_lag;    // adaptive multiplier of the bucket width; lag (loop-invariant state kept out of registers)

    const u32 tid = threadIdx.x;
    const u32 lane = tid & 31u;
    const u32 t = tid % LP;                       // lane within the sub-warp = which words of a row
    const u32 sub = tid / LP;                     // sub-warp within the CTA = which entry of a pass
    const u32 sub_shift = (lane / LP) * LP;       // position of the sub-warp's bits in a warp ballot
    const bool meta_lane = t == (u32)LP - 1u;
    const u32 V = p.V;
    const size_t ws_stride = (size_t)2 * gcap + (size_t)2 * fcap;
    // cluster: rank of this CTA, the cluster's workspace and control words; the accessors pick shared or global state
    const u32 crank = CL > 1 ? blockIdx.x % (u32)CL : 0u;
    const u32 unit = CL > 1 ? blockIdx.x / (u32)CL : blockIdx.x;
    u32 *const ctl = CL > 1 ? p.cluster_ctl + (size_t)unit * GC_WORDS : nullptr;
    const bool lead = tid == 0 && crank == 0u;                    // the one thread that owns the shared control state
    // the queue counters of a cluster live in the shared memory of its first CTA and are reached through distributed
    // shared memory (an atomic there returns in ~215 cycles, a global one in ~700: it sits on every pass's push)
    u32 *cl_cnt = s_cnt, *cl_far_cnt = s_far_cnt, *cl_far_min = s_far_min, *cl_overflow = &s_overflow;
    if constexpr (CL > 1) {
        cg::cluster_group cluster = cg::this_cluster();
        cl_cnt = cluster.map_shared_rank(s_cnt, 0);
        cl_far_cnt = cluster.map_shared_rank(s_far_cnt, 0);
        cl_far_min = cluster.map_shared_rank(s_far_min, 0);
        cl_overflow = cluster.map_shared_rank(&s_overflow, 0);
    }
    auto CNT = [&](u32 i) -> u32 * { if constexpr (CL > 1) return cl_cnt + i; else return &s_cnt[i]; };
    auto FAR_CNT = [&](u32 i) -> u32 * { if constexpr (CL > 1) return cl_far_cnt + i; else return &s_far_cnt[i]; };
    auto FAR_MIN = [&](u32 i) -> u32 * { if constexpr (CL > 1) return cl_far_min + i; else return &s_far_min[i]; };
    auto OVERFLOW = [&]() -> u32 * { if constexpr (CL > 1) return cl_overflow; else return &s_overflow; };
    auto rd = [&](const u32 *q) -> u32 { if constexpr (CL > 1) return *reinterpret_cast<const volatile u32 *>(q); else return *q; };
    auto sync_all = [&]() { if constexpr (CL > 1) cg::this_cluster().sync(); else __syncthreads(); };
    u32 *wv = p.ws_v + (size_t)unit * ws_stride;
    u32 *wk = p.ws_k + (size_t)unit * ws_stride;
    auto target = [&](u32 nb, u32 nf_) {
        return GroupPushTarget{sv0 + nb * scap, wv, wk, nb * gcap - scap, 2u * gcap + nf_ * fcap, scap, scap + gcap, fcap};
    };

    u32 c_pop = 0, c_dup = 0, c_push = 0;                  // per-thread counts of one launch
    u64 c_relax = 0, c_pop_all = 0;
    if (tid == 0) { s_stat[0] = 0u; s_stat[1] = 0u; s_stat[2] = 0u; }

    for (;;) {
        sync_all();                            // previous group fully retired before the control state is reused
        if constexpr (CL > 1) { if (lead) ctl[GC_GROUP] = atomicAdd(p.work_counter, 1u); }
        else if (tid == 0) s_group = (int)atomicAdd(p.work_counter, 1u);
        sync_all();
        const int gi = CL > 1 ? (int)__ldcg(ctl + GC_GROUP) : s_group;
        if (gi >= p.n_groups) break;
        u64 *rows = p.pool + (size_t)gi * p.slot_stride;
        int k_live = 0;                                         // trees of this group: the slots holding a root
        {
            const int k_max = min(K, p.n_roots - gi * K);
            for (int k = 0; k < k_max; ++k) k_live += p.roots[gi * K + k] >= 0 ? 1 : 0;
        }

        // ---- fill the slot: +inf distances, meta = 0 (16-byte streaming stores)
        {
            ulonglong2 *d2 = reinterpret_cast<ulonglong2 *>(rows);
            const size_t n2 = p.slot_stride >> 1;
            const ulonglong2 inf2 = make_ulonglong2(kInfBits, kInfBits), tail2 = make_ulonglong2(kInfBits, 0ull);
            for (size_t i = tid + (size_t)crank * BLOCK; i < n2; i += (size_t)BLOCK * CL)
                __stcs(d2 + i, (i % (RW / 2)) == (size_t)(RW / 2 - 1) ? tail2 : inf2);
        }
        if (lead) {
            *CNT(0) = (u32)k_live; *CNT(1) = 0u; *CNT(2) = 0u;
            *FAR_CNT(0) = 0u; *FAR_CNT(1) = 0u;
            *FAR_MIN(0) = 0xFFFFFFFFu; *FAR_MIN(1) = 0xFFFFFFFFu;
            *OVERFLOW() = 0u;
        }
        if (tid == 0) {                                         // replicated per CTA: every CTA of a cluster computes the same
            s_bucket_rounds = 0u; s_bucket_entries = 0u;
            s_lag_hi = 0u;
            s_glog[0] = 0u; s_glog[1] = 0u; s_glog[2] = 0u;
            s_t0 = clock64();
            s_dmul = 1.0; s_lag = 0.0; s_lag_scale = 1.0;
        }
        sync_all();
        if (tid < (u32)k_live) {                                // after the fill: the roots, queued for round 1
            const u32 r = (u32)p.roots[gi * K + (int)tid];
            s_roots[tid] = r;
            if (crank == 0u) {
                rows[(size_t)r * RW + tid] = 0ull;
                reinterpret_cast<u32 *>(rows + (size_t)r * RW + K)[0] = 1u;
                if constexpr (CL > 1) wv[tid] = r; else sv0[tid] = r;
            }
        }
        sync_all();

        u32 c = 0;          // index of the current near counter
        u32 b = 0;          // current near buffer
        u32 f = 0;          // current far pile (push target)
        u32 rid = 1;        // id of the queue the coming round reads (pushes of a round are stamped rid + 1)
        u32 thr_hi;         // near bucket = keys with high word <= thr_hi
        {
            const double delta = p.delta_scale * __ldg(p.wsum);
            thr_hi = hi32((u64)__double_as_longlong(delta > 0.0 ? delta : 0.0));
        }

        u32 rib = 0;                // rounds in the current bucket
        bool force_round = false;   // an early advance found the far pile empty: run the round instead
        for (;;) {
            const u32 n = rd(CNT(c));
            const bool early = n > 0u && n < p.early_advance && rib >= 2u && !force_round;
            if (n > 0 && !early) {
                ++rib;
                force_round = false;
                // ------------------------- relaxation round -------------------------
                const u32 cn = c == 2 ? 0 : c + 1;
                if (lead) *CNT(cn == 2 ? 0 : cn + 1) = 0u;             // counter of the round after next
                if (tid == 0) {
                    s_bucket_rounds += 1u;
                    s_bucket_entries += n;
                    s_stat[1] += 1u;
                    s_glog[0] += 1u;
                    s_glog[1] += n;
                }
                const u32 *const csv = sv0 + b * scap;                      // the queue this round reads
                const u32 cur_off = b * gcap - scap;                        // its spill: entry i >= scap at cur_off + i
                const GroupPushTarget tg = target(b ^ 1u, f);
                const u32 n_s = n < scap + gcap ? n : scap + gcap;          // entries that were actually stored
                const u32 rid_next = rid + 1u;
                const double thr_up = __hiloint2double((int)(thr_hi + 1u), 0);   // upper end of the bucket
                const double lag = s_lag;
                u32 u_next = 0u;                                            // cluster: the entry of the next pass, fetched a pass ahead
                if constexpr (CL > 1) {
                    const u32 i0 = crank * PW + sub;
                    if (i0 < n_s) u_next = __ldcg(wv + cur_off + i0);
                }
                for (u32 base = 0; base < n_s; base += PW * (u32)CL) {
                    const u32 i = base + crank * PW + sub;
                    const bool valid = i < n_s;                             // uniform in the sub-warp
                    u32 u = 0;
                    Arc a[kEllWidth];
                    u64 my[TL];
#pragma unroll
                    for (int q = 0; q < TL; ++q) my[q] = kInfBits;
#pragma unroll
                    for (int j = 0; j < kEllWidth; ++j) { a[j].col = 0u; a[j].street = 0u; a[j].w = 0.0; }
                    if (valid) {
                        if constexpr (CL > 1) {
                            // written by any CTA of the cluster (global workspace); a global read would be one more
                            // dependent stage in front of the row loads, so it was issued during the previous pass
                            u = u_next;
                            const u32 i_next = i + PW * (u32)CL;
                            if (i_next < n_s) u_next = __ldcg(wv + cur_off + i_next);
                        } else u = i < scap ? csv[i] : wv[cur_off + i];
                        const Arc *row = p.ellw + (size_t)u * kEllWidth;
                        ld_row_half(row, a[0], a[1]);
                        ld_row_half(row + 2, a[2], a[3]);
                        ld_words<TL>(rows + (size_t)u * RW + t * TL, my);
                    }
                    const bool ovf = valid && (a[0].col & kOvfBit) != 0u;
                    u32 vj[kEllWidth];
#pragma unroll
                    for (int j = 0; j < kEllWidth; ++j) vj[j] = a[j].col & ~kOvfBit;
                    bool ok[TL];
#pragma unroll
                    for (int q = 0; q < TL; ++q) ok[q] = !(meta_lane && q == TL - 1) && hi32(my[q]) <= thr_hi;
                    u32 xmin[kEllWidth];
                    u32 bal_imp[kEllWidth], bal_new[kEllWidth], bal_bey[kEllWidth];
                    // the neighbour rows of STAGE arcs are in flight together: all four with up to two words per lane; with
                    // four words (rows of 256 bytes, 31 trees) two at a time, or the distances alone would take 32 registers
                    constexpr int STAGE = TL >= 4 ? 2 : kEllWidth;
#pragma unroll
                    for (int s0 = 0; s0 < kEllWidth; s0 += STAGE) {
                        u64 dv[STAGE][TL];
#pragma unroll
                        for (int jj = 0; jj < STAGE; ++jj) {
#pragma unroll
                            for (int q = 0; q < TL; ++q) dv[jj][q] = 0ull;       // nothing improves on 0
                            // padding slots and self loops point back at u: nothing to relax, no read
                            if (valid && vj[s0 + jj] != u) ld_words<TL>(rows + (size_t)vj[s0 + jj] * RW + t * TL, dv[jj]);
                        }
#pragma unroll
                        for (int jj = 0; jj < STAGE; ++jj) {
                            const int j = s0 + jj;
                            bool im, beyond;
                            relax_words<TL>(rows + (size_t)vj[j] * RW + t * TL, my, ok, a[j].w, dv[jj], meta_lane, thr_hi, xmin[j], im, beyond);
                            bal_imp[j] = __ballot_sync(0xffffffffu, im);
                            bal_bey[j] = __ballot_sync(0xffffffffu, beyond);
                            // the meta lane holds v's stamp: not yet queued for the coming round?
                            bal_new[j] = __ballot_sync(0xffffffffu, meta_lane && (u32)dv[jj][TL - 1] != rid_next);
                        }
                    }
                    // sub-warp minimum of the (updated) distances of every head: lane t gets slot j(t)
                    const u32 zmin = subwarp_min4<LP>(xmin[0], xmin[1], xmin[2], xmin[3], t);
                    const bool pusher = LP == 8 ? (t & 1u) == 0u : true;
                    const u32 j = LP == 8 ? t >> 1 : t;
                    const bool any_imp = ((sel4(j, bal_imp[0], bal_imp[1], bal_imp[2], bal_imp[3]) >> sub_shift) & LPMASK) != 0u;
                    const bool any_bey = ((sel4(j, bal_bey[0], bal_bey[1], bal_bey[2], bal_bey[3]) >> sub_shift) & LPMASK) != 0u;
                    const bool fresh = ((sel4(j, bal_new[0], bal_new[1], bal_new[2], bal_new[3]) >> sub_shift) & LPMASK) != 0u;
                    const u32 pv = sel4(j, vj[0], vj[1], vj[2], vj[3]);
                    // key of the entry: the group's earliest distance at the head (+ lag); an improved value beyond the
                    // bucket must not be expanded before the bucket reaches it: it is at most bucket end + w
                    u32 key = zmin;
                    if (lag > 0.0) key = hi32((u64)__double_as_longlong(__dadd_rn(__hiloint2double((int)zmin, 0), lag)));
                    if (any_bey) {
                        const u32 b0 = hi32((u64)__double_as_longlong(__dadd_rn(thr_up, a[0].w))), b1 = hi32((u64)__double_as_longlong(__dadd_rn(thr_up, a[1].w)));
                        const u32 b2 = hi32((u64)__double_as_longlong(__dadd_rn(thr_up, a[2].w))), b3 = hi32((u64)__double_as_longlong(__dadd_rn(thr_up, a[3].w)));
                        key = max(max(key, thr_hi + 1u), sel4(j, b0, b1, b2, b3));
                    }
                    u32 kind = 0u;
                    if (pusher && any_imp) {
                        if (key <= thr_hi) {
                            if (fresh) {
                                kind = 1u;
                                reinterpret_cast<u32 *>(rows + (size_t)pv * RW + K)[0] = rid_next;   // stamp: queued for the coming round
                            } else ++c_dup;
                        } else kind = 2u;
                    }
                    warp_push_g(kind, pv, key, lane, tg, CNT(cn), FAR_CNT(f), FAR_MIN(f), OVERFLOW(), c_push);
                    if (valid && t == 0u) ++c_pop;
                    // nodes with more than kEllWidth arcs: the rest of the CSR row, one arc per step
                    if (__any_sync(0xffffffffu, ovf)) {
                        u32 e = 0, re = 0;
                        if (ovf) { e = __ldg(p.row_ptr + u) + kEllWidth; re = __ldg(p.row_ptr + u + 1); }
                        while (__any_sync(0xffffffffu, e < re)) {
                            const bool act = e < re;
                            u32 v1 = u;
                            double w1 = 0.0;
                            if (act) { const Arc x = ld_arc(p.arcs + e); v1 = x.col; w1 = x.w; }
                            u64 d1[TL];
#pragma unroll
                            for (int q = 0; q < TL; ++q) d1[q] = 0ull;
                            if (act && v1 != u) ld_words<TL>(rows + (size_t)v1 * RW + t * TL, d1);
                            u32 mn;
                            bool im, beyond;
                            relax_words<TL>(rows + (size_t)v1 * RW + t * TL, my, ok, w1, d1, meta_lane, thr_hi, mn, im, beyond);
                            const u32 bi = __ballot_sync(0xffffffffu, im);
                            const u32 bb = __ballot_sync(0xffffffffu, beyond);
                            const u32 bn = __ballot_sync(0xffffffffu, meta_lane && (u32)d1[TL - 1] != rid_next);
#pragma unroll
                            for (int o = LP / 2; o > 0; o >>= 1) mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
                            u32 key1 = mn;
                            if (lag > 0.0) key1 = hi32((u64)__double_as_longlong(__dadd_rn(__hiloint2double((int)mn, 0), lag)));
                            if (((bb >> sub_shift) & LPMASK) != 0u)
                                key1 = max(max(key1, thr_hi + 1u), hi32((u64)__double_as_longlong(__dadd_rn(thr_up, w1))));
                            u32 k1 = 0u;
                            if (t == 0u && ((bi >> sub_shift) & LPMASK) != 0u) {
                                if (key1 <= thr_hi) {
                                    if (((bn >> sub_shift) & LPMASK) != 0u) {
                                        k1 = 1u;
                                        reinterpret_cast<u32 *>(rows + (size_t)v1 * RW + K)[0] = rid_next;
                                    } else ++c_dup;
                                } else k1 = 2u;
                            }
                            warp_push_g(k1, v1, key1, lane, tg, CNT(cn), FAR_CNT(f), FAR_MIN(f), OVERFLOW(), c_push);
                            if (act) ++e;
                        }
                    }
                }
                sync_all();
                c = cn;
                b ^= 1u;
                rid = rid_next;
                continue;
            }
            // ------------------------- near queue empty (or thin): advance the bucket -------------------------
            // every thread must have read s_cnt[c] before the split starts refilling that queue; past this barrier no
            // push is in flight, so the far pile's counters are stable
            sync_all();
            const u32 nf_raw = rd(FAR_CNT(f));
            if (nf_raw == 0u) {
                if (n == 0u) break;                                     // group complete
                force_round = true;                                     // nothing to admit yet
                continue;
            }
            const u32 nf = nf_raw < fcap ? nf_raw : fcap;
            const u32 far_min_hi = rd(FAR_MIN(f));
            rib = 0;
            if (tid == 0) { s_stat[0] += 1u; s_glog[2] += 1u; }
            // lag: the largest finite distance between two roots of the group (values only decrease towards the
            // final ones; a heuristic either way)
            if (p.lag_factor > 0.0) {
                for (u32 idx = tid; idx < (u32)(k_live * k_live); idx += BLOCK) {
                    const u32 jr = idx / (u32)k_live, kt = idx % (u32)k_live;
                    const u32 h = hi32(__ldcg(rows + (size_t)s_roots[jr] * RW + kt));
                    if (h < 0x7FF00000u) atomicMax(&s_lag_hi, h);
                }
                __syncthreads();
                if (tid == 0) {
                    if (nf_raw > fcap / 2u) s_lag_scale *= 0.5;
                    s_lag = p.lag_factor * s_lag_scale * __hiloint2double((int)s_lag_hi, 0);
                }
            }
            const double lo = __hiloint2double((int)far_min_hi, 0);
            {
                double dmul = s_dmul;                                   // every thread computes the same value ...
                if (p.adaptive) {
                    const u32 br = s_bucket_rounds, be = s_bucket_entries;
                    if (br > 0u) {
                        // a cluster widens its buckets only half as eagerly as its width would suggest: wider buckets
                        // mean more re-expansions (cfg3, two CTAs per group: 1.73x -> the work of one CTA's buckets)
                        constexpr u32 AW = PW * (CL > 1 ? (u32)CL / 2u : 1u);
                        if (be < br * (AW / 2u) && dmul < 16.0) dmul *= 1.5;
                        else if (be > br * (AW * 4u) && dmul > 1.0) dmul *= (1.0 / 1.5);
                    }
                }
                double delta = p.delta_scale * __ldg(p.wsum);
                if (!(delta > 0.0)) delta = 0.0;
                thr_hi = hi32((u64)__double_as_longlong(__dadd_rn(lo, delta * dmul)));
                __syncthreads();                                        // ... from the old one, then thread 0 stores the new one
                if (tid == 0) s_dmul = dmul;
            }
            const u32 fo = f ^ 1u;
            const u32 src_off = 2u * gcap + f * fcap;
            const GroupPushTarget tg = target(b, fo);
            // four entries per thread and step, their loads (and the stamps of the admitted ones) in flight together:
            // with heavy-tailed weights (streets adapted down to 1 km/h) the pile holds many distant entries that are
            // re-filed at every advance, and a step is one memory round trip
            constexpr u32 SPLIT = 4;
            for (u32 base = 0; base < nf; base += BLOCK * (u32)CL * SPLIT) {
                u32 kx[SPLIT], vx[SPLIT], keyx[SPLIT];
#pragma unroll
                for (u32 s4 = 0; s4 < SPLIT; ++s4) {
                    const u32 i = base + (s4 * (u32)CL + crank) * BLOCK + tid;
                    kx[s4] = 0u; vx[s4] = 0u; keyx[s4] = 0u;
                    if (i < nf) {
                        if constexpr (CL > 1) { vx[s4] = __ldcg(wv + src_off + i); keyx[s4] = __ldcg(wk + src_off + i); }
                        else { vx[s4] = __ldcs(wv + src_off + i); keyx[s4] = __ldcs(wk + src_off + i); }
                        kx[s4] = 2u;
                    }
                }
                u32 oldx[SPLIT];
#pragma unroll
                for (u32 s4 = 0; s4 < SPLIT; ++s4) {
                    oldx[s4] = rid;
                    // admitted: unless the node is already queued for the coming round
                    if (kx[s4] != 0u && keyx[s4] <= thr_hi) {
                        oldx[s4] = atomicExch(reinterpret_cast<u32 *>(rows + (size_t)vx[s4] * RW + K), rid);
                        kx[s4] = 1u;
                    }
                }
#pragma unroll
                for (u32 s4 = 0; s4 < SPLIT; ++s4) {
                    if (kx[s4] == 1u && oldx[s4] == rid) { kx[s4] = 0u; ++c_dup; }
                    // the round after this reads (b, c); kept entries go to the other pile (never more than nf)
                    warp_push_g(kx[s4], vx[s4], keyx[s4], lane, tg, CNT(c), FAR_CNT(fo), FAR_MIN(fo), OVERFLOW(), c_push);
                }
            }
            sync_all();
            if (lead) { *FAR_CNT(f) = 0u; *FAR_MIN(f) = 0xFFFFFFFFu; }
            if (tid == 0) { s_bucket_rounds = 0u; s_bucket_entries = 0u; }
            f = fo;
        }

        c_relax += (u64)c_pop * (u64)(kEllWidth * k_live);
        c_pop_all += c_pop;
        c_pop = 0u;     // arcs x trees of the expansions (nodes with more arcs not counted)
        if (p.group_log && lead) {
            u32 *gl = p.group_log + (size_t)gi * 4;
            gl[0] = (u32)((clock64() - s_t0) >> 10); gl[1] = s_glog[0]; gl[2] = s_glog[1]; gl[3] = s_glog[2];
        }
        // ---- safety net: a queue overflowed and entries were dropped.  Sweep all nodes Bellman-Ford style
        // until nothing changes (same fixed point).
        if (rd(OVERFLOW())) {
            for (;;) {
                sync_all();
                if constexpr (CL > 1) { if (lead) ctl[GC_CHANGED] = 0u; }
                if (tid == 0) { s_changed = 0u; s_stat[2] += 1u; }
                sync_all();
                bool any = false;
                for (u32 u = tid + crank * BLOCK; u < V; u += BLOCK * (u32)CL) {
                    const u32 rb = __ldg(p.row_ptr + u), re = __ldg(p.row_ptr + u + 1);
                    for (int k = 0; k < k_live; ++k) {
                        const u64 du = __ldcg(rows + (size_t)u * RW + k);
                        if (du >= kInfBits) continue;
                        for (u32 e = rb; e < re; ++e) {
                            const Arc x = ld_arc(p.arcs + e);
                            const u64 ndb = (u64)__double_as_longlong(__dadd_rn(__longlong_as_double((long long)du), x.w));
                            if (ndb < __ldcg(rows + (size_t)x.col * RW + k)) {
                                atomicMin(rows + (size_t)x.col * RW + k, ndb);
                                any = true;
                            }
                        }
                    }
                }
                if constexpr (CL > 1) { if (any) ctl[GC_CHANGED] = 1u; }
                else if (any) s_changed = 1u;
                sync_all();
                if (!(CL > 1 ? __ldcg(ctl + GC_CHANGED) : s_changed)) break;
            }
        }
    }

    // ---- flush per-thread counters: one atomic per warp per counter
    __syncthreads();
    u64 vals[7] = {c_relax, c_pop_all, c_dup, c_push, lead ? s_stat[0] : 0u, lead ? s_stat[1] : 0u, lead ? s_stat[2] : 0u};
    const int ids[7] = {C_ARCS_RELAXED, C_NODES_POPPED, C_STALE_POPS, C_PUSHES, C_ADVANCES, C_ROUNDS, C_FALLBACK_SWEEPS};
#pragma unroll
    for (int jx = 0; jx < 7; ++jx) {
        u64 v = vals[jx];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0 && v) atomicAdd(p.counters + ids[jx], v);
    }
}

}  // namespace s4
