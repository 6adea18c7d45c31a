// K3 for trip-heavy days: predecessor codes + one-lane walks.
//   simulation.py:78-88 (walk every trip goal -> origin, load[street] += trip_volume).
//
// walk_kernel derives the predecessor of every node it passes from the converged distances: per hop the arc row, then
// the distance words of up to three neighbours -- 5.4 random 32-byte DRAM sectors, two dependent memory stages, eight
// lanes.  With hundreds of trips per tree (BASELINE north_star: 10 M trips/day, 477 per tree) the same nodes are
// resolved again and again, and the walk costs as much as the shortest-path trees.  Then it pays to resolve every
// (node, tree) ONCE, in a streaming pass over the group's slot, and let the walks follow the result:
//
// pred_code_kernel   for every node of a group slot: the arc slot (2 bits) of the canonical predecessor in each tree of
//                    the group, packed into the row's `meta` word (the stamps in it are dead once the group has
//                    converged): tree k at bits [2k, 2k+2), bit 63 = "slow node" -- some tree has several closer
//                    candidates (tie: the rule needs ids), sits on a zero-weight plateau, or the node has more arcs
//                    than the fixed-stride row holds.  Same lane layout and coalesced row accesses as the SSSP expansion.
// walk_code_kernel   ONE lane per trip and one dependent stage per hop: meta word and arc row are fetched together, the
//                    code selects the arc.  At a slow node the lane applies the rule itself over the CSR row (ties, wide
//                    nodes); only zero-weight plateaus are handed to walk_slow_kernel (the bounded search, from that node
//                    on), so the result is the canonical one bit for bit.
#pragma once

#include "s4_sssp_group.cuh"

namespace s4 {

struct CodeParams {
    const Arc *ellw;
    u64 *pool;              // first group slot of the batch
    size_t slot_stride;     // words per slot
    int n_groups;
    u32 V;
};

static constexpr u64 kSlowBit = 1ull << 63;

template <int TL>
__global__ void __launch_bounds__(256, 4) pred_code_kernel(CodeParams p)
{
    constexpr int LP = 8, RW = LP * TL;
    const u32 t = threadIdx.x % LP;
    const bool meta_lane = t == (u32)LP - 1u;
    const u64 total = (u64)p.n_groups * (u64)p.V;
    const u64 per_grid = (u64)gridDim.x * (blockDim.x / LP);
    const u64 warp_first = ((u64)blockIdx.x * blockDim.x + (threadIdx.x & ~31u)) / LP;     // first node of this warp
    for (u64 base = warp_first; base < total; base += per_grid) {                          // warp-uniform trip count
        const u64 idx = base + (threadIdx.x & 31u) / LP;
        const bool valid = idx < total;
        const u32 gi = valid ? (u32)(idx / p.V) : 0u, v = valid ? (u32)(idx % p.V) : 0u;
        u64 *rows = p.pool + (size_t)gi * p.slot_stride;
        Arc a[kEllWidth];
        u64 my[TL];
#pragma unroll
        for (int j = 0; j < kEllWidth; ++j) { a[j].col = v; a[j].street = 0u; a[j].w = 0.0; }
#pragma unroll
        for (int q = 0; q < TL; ++q) my[q] = kInfBits;
        if (valid) {
            const Arc *row = p.ellw + (size_t)v * kEllWidth;
            ld_row_half(row, a[0], a[1]);
            ld_row_half(row + 2, a[2], a[3]);
            ld_words<TL>(rows + (size_t)v * RW + t * TL, my);
        }
        u64 dv[kEllWidth][TL];
#pragma unroll
        for (int j = 0; j < kEllWidth; ++j) {
#pragma unroll
            for (int q = 0; q < TL; ++q) dv[j][q] = kInfBits;
            const u32 vj = a[j].col & ~kOvfBit;
            if (valid && vj != v) ld_words<TL>(rows + (size_t)vj * RW + t * TL, dv[j]);    // padding slots point back at v
        }
        u64 word = valid && (a[0].col & kOvfBit) != 0u ? kSlowBit : 0ull;
#pragma unroll
        for (int q = 0; q < TL; ++q) {
            if (meta_lane && q == TL - 1) continue;                                        // the meta word itself
            u32 n_close = 0, n_equal = 0, code = 0;
#pragma unroll
            for (int j = 0; j < kEllWidth; ++j) {
                const u64 du = dv[j][q];
                const u64 cand = (u64)__double_as_longlong(__dadd_rn(__longlong_as_double((long long)du), a[j].w));
                if (du < kInfBits && cand == my[q]) {
                    if (du < my[q]) { ++n_close; code = (u32)j; }
                    else ++n_equal;
                }
            }
            if (n_close > 1u || (n_close == 0u && n_equal > 0u)) word |= kSlowBit;
            word |= (u64)code << (2u * (t * (u32)TL + (u32)q));
        }
#pragma unroll
        for (int o = 1; o < LP; o <<= 1) word |= __shfl_xor_sync(0xffffffffu, word, o);
        if (valid && meta_lane) rows[(size_t)v * RW + (RW - 1)] = word;
    }
}

// one lane per trip; lanes fetch their next trip as soon as the current one ends (no lane waits for its warp)
__global__ void __launch_bounds__(256) walk_code_kernel(WalkParams p)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t t_next = p.t0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    u32 c_hops = 0, c_unreach = 0, c_stuck = 0, c_trips = 0;
    bool active = false;
    const u64 *rows = nullptr;
    u32 cur = 0, root = 0, shift = 0, guard = 0, t_rel = 0;
    const u32 RW = p.node_stride;
    for (;;) {
        if (!active) {
            if (t_next >= p.t1) break;
            const int64_t t = t_next;
            t_next += stride;
            const u32 ri = p.trip_root[t];
            const u32 ti = ri - p.root0;
            rows = p.dist_pool + (size_t)(ti / p.group) * p.slot_stride;
            shift = 2u * (ti % p.group);
            root = (u32)p.roots[ri];
            cur = (u32)p.trip_target[t];
            t_rel = (u32)(t - p.t_base);
            guard = 0;
            ++c_trips;
            if (ld_dist(rows + (size_t)cur * RW + (shift >> 1)) >= kInfBits) { ++c_unreach; continue; }   // simulation.py:80
            if (cur == root) continue;
            active = true;
        }
        const u64 meta = __ldcg(rows + (size_t)cur * RW + (RW - 1u));
        // {head | overflow flag, street | reverse slot << 29} x 4: the static 32-byte row of the graph (one LDG.E.256,
        // 33 MB on cfg3: stays in L2), not the 64-byte row with the day's weights
        u32 h0, s0, h1, s1, h2, s2, h3, s3;
        asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(h0), "=r"(s0), "=r"(h1), "=r"(s1), "=r"(h2), "=r"(s2), "=r"(h3), "=r"(s3) : "l"(p.ell_cs + (size_t)cur * kEllWidth));
        u32 nxt, street;
        if (meta & kSlowBit) {
            // a tie in some tree of the group or a node with more arcs than the row holds: this lane applies the rule
            // itself over the node's CSR row -- smallest (id, street) among the strictly closer candidates
            const u64 *dist = rows + (shift >> 1);
            const u64 dv = ld_dist(dist + (size_t)cur * RW);
            u64 best = ~0ull;
            u32 best_col = 0u, n_equal = 0u;
            for (u32 e = __ldg(p.row_ptr + cur), re = __ldg(p.row_ptr + cur + 1); e < re; ++e) {
                const Arc a = ld_arc(p.arcs + e);
                if (a.col == cur) continue;
                const u64 du = ld_dist(dist + (size_t)a.col * RW);
                if (du >= kInfBits || (u64)__double_as_longlong(__dadd_rn(__longlong_as_double((long long)du), a.w)) != dv) continue;
                if (du < dv) {
                    const u64 key = ((u64)(p.orig ? __ldg(p.orig + a.col) : a.col) << 32) | a.street;
                    if (key < best) { best = key; best_col = a.col; }
                } else ++n_equal;
            }
            if (best == ~0ull) {
                if (n_equal) {
                    // a zero-weight plateau: the bounded search of walk_slow_kernel takes over from here
                    const u32 slot = atomicAdd(p.deferred_count, 1u);
                    p.deferred[slot] = make_uint2(t_rel, cur);
                } else ++c_stuck;
                active = false;
                continue;
            }
            nxt = best_col;
            street = (u32)best;
        } else {
            const u32 j = (u32)(meta >> shift) & 3u;
            nxt = sel4(j, h0, h1, h2, h3) & ~kOvfBit;
            street = sel4(j, s0, s1, s2, s3) & 0x1FFFFFFFu;
        }
        if (nxt == cur || ++guard > p.V) { ++c_stuck; active = false; continue; }      // never with a converged slot
        atomicAdd(p.load + street, p.volume);                                          // :86-88
        ++c_hops;
        cur = nxt;                                                                     // :85
        if (cur == root) active = false;
        if (c_hops > 0x40000000u) { atomicAdd(p.counters + C_HOPS, (u64)c_hops); c_hops = 0; }
    }
    if (c_trips) atomicAdd(p.counters + C_TRIPS, (u64)c_trips);
    if (c_hops) atomicAdd(p.counters + C_HOPS, (u64)c_hops);
    if (c_unreach) atomicAdd(p.counters + C_UNREACHABLE, (u64)c_unreach);
    if (c_stuck) atomicAdd(p.counters + C_STUCK, (u64)c_stuck);
}

}  // namespace s4
