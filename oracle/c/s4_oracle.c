/*
 * CPU oracle ("C twin") for the Streets4MPI per-day traffic-assignment path.
 *
 * TEST INFRASTRUCTURE ONLY -- linked by tests/, __graft_entry__.smoke() and the
 * cpu_baseline leg of bench.py (through ctypes), never by the product library.
 * Same arithmetic as oracle/port.py (which restates the reference line by
 * line); this file exists because pure Python cannot route 10^5..10^6-node
 * graphs in test time.  Build: oracle/Makefile  (-O2 -ffp-contract=off so that
 * a*b+c is never fused -- simulation.py:61 must round twice).
 *
 * Reference lines restated (paths relative to /root/reference):
 *   o_driving_speed / o_weights   simulation.py:129-138, :53-65
 *   o_sssp                        streetnetwork.py:108-109 -> python-graph Dijkstra
 *                                 (strict '<'; parity unpinned on ties, see oracle/__init__.py)
 *   o_day                         simulation.py:68-88
 *   o_merge                       utils.py:26-34, streets4mpi.py:97-100
 *   o_road_construction           simulation.py:91-109, streetnetwork.py:79-82
 *
 * Graph layout: symmetric CSR over dense node ids 0..V-1.  arc e of node u is
 * (col[e], arc_street[e]); weights are per street (w[street]).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

typedef struct { double d; int32_t v; } heap_item;

/* simulation.py:129-138 */
double o_driving_speed(double length, double max_speed, double trips,
                       double car_length, double min_brk, double decel)
{
    double per_car = length / (trips > 1.0 ? trips : 1.0);
    double room = per_car - car_length;
    if (!(room > min_brk)) room = min_brk;           /* max(room, min_brk) */
    double pot = sqrt(decel * room * 2.0);
    double kmh = pot * 3.6;
    return max_speed <= kmh ? max_speed : kmh;       /* min(max_speed, kmh) */
}

/* simulation.py:53-65 for every street */
void o_weights(int64_t S, const double *length, const int32_t *max_speed,
               const uint32_t *load, double jam, double car_length,
               double min_brk, double decel, double *w_out)
{
    for (int64_t s = 0; s < S; ++s) {
        double ms = (double)max_speed[s];
        double ideal = o_driving_speed(length[s], ms, 0.0, car_length, min_brk, decel);
        double actual = o_driving_speed(length[s], ms, (double)load[s], car_length, min_brk, decel);
        double diff = ideal - actual;
        double scaled = diff * jam;
        double perceived = actual + scaled;
        w_out[s] = length[s] / perceived;
    }
}

/* ---- binary heap on (d, v), lazy deletion -------------------------------- */
static inline int item_less(heap_item a, heap_item b)
{
    return a.d < b.d || (a.d == b.d && a.v < b.v);
}
static void heap_push(heap_item *h, int64_t *n, heap_item x)
{
    int64_t i = (*n)++;
    while (i > 0) {
        int64_t p = (i - 1) >> 1;
        if (!item_less(x, h[p])) break;
        h[i] = h[p];
        i = p;
    }
    h[i] = x;
}
static heap_item heap_pop(heap_item *h, int64_t *n)
{
    heap_item top = h[0];
    heap_item x = h[--(*n)];
    int64_t i = 0, m = *n;
    for (;;) {
        int64_t c = 2 * i + 1;
        if (c >= m) break;
        if (c + 1 < m && item_less(h[c + 1], h[c])) ++c;
        if (!item_less(h[c], x)) break;
        h[i] = h[c];
        i = c;
    }
    if (m > 0) h[i] = x;
    return top;
}

/* Dijkstra distances (strict '<' relaxation).  dist = +inf for unreachable. */
static void dijkstra(int32_t V, const int64_t *row_ptr, const int32_t *col,
                     const int32_t *arc_street, const double *w, int32_t src,
                     double *dist, heap_item *heap, uint8_t *done)
{
    for (int32_t i = 0; i < V; ++i) dist[i] = INFINITY;
    memset(done, 0, (size_t)V);
    int64_t n = 0;
    dist[src] = 0.0;
    heap_push(heap, &n, (heap_item){0.0, src});
    while (n > 0) {
        heap_item it = heap_pop(heap, &n);
        int32_t u = it.v;
        if (done[u]) continue;
        done[u] = 1;
        for (int64_t e = row_ptr[u]; e < row_ptr[u + 1]; ++e) {
            int32_t v = col[e];
            if (done[v]) continue;
            double alt = it.d + w[arc_street[e]];
            if (alt < dist[v]) {
                dist[v] = alt;
                heap_push(heap, &n, (heap_item){alt, v});
            }
        }
    }
}

/*
 * Canonical predecessor of v given converged distances (same rule as oracle/port.py canonical_tree).
 * Candidates: arcs (v,u), u != v, dist[u] finite, fl(dist[u]+w) == dist[v].
 *   - some candidate strictly closer (dist[u] < dist[v]): min (u, street) among those;
 *   - else v is on a plateau (zero / absorbed weights): breadth-first search over the plateau, candidates in
 *     ascending (u, street) order, at most PLATEAU_MAX nodes, to the nearest node with a strictly closer
 *     candidate (or the root); the first hop of that path.  Levels decrease along the chain: no cycles.
 * Returns -1 if none.  *n_cand = number of candidates of v.
 */
#define PLATEAU_MAX 32

/* smallest candidate of x that is > (after_u, after_s) in (u, street) order; closer_only: dist[u] < dist[x] */
static int32_t next_cand(int32_t x, int32_t after_u, int32_t after_s, int closer_only, const int64_t *row_ptr,
                         const int32_t *col, const int32_t *arc_street, const double *w, const double *dist,
                         int32_t *street_out)
{
    double dx = dist[x];
    int32_t best_u = -1, best_s = -1;
    for (int64_t e = row_ptr[x]; e < row_ptr[x + 1]; ++e) {
        int32_t u = col[e], s = arc_street[e];
        if (u == x) continue;
        double du = dist[u];
        if (isinf(du) || !(du + w[s] == dx)) continue;
        if (closer_only && !(du < dx)) continue;
        if (u < after_u || (u == after_u && s <= after_s)) continue;
        if (best_u < 0 || u < best_u || (u == best_u && s < best_s)) { best_u = u; best_s = s; }
    }
    *street_out = best_s;
    return best_u;
}

static int32_t canon_pred_root(int32_t v, int32_t root, const int64_t *row_ptr, const int32_t *col,
                               const int32_t *arc_street, const double *w,
                               const double *dist, int32_t *street_out, int *n_cand)
{
    int32_t s = -1;
    if (n_cand) {
        int nc = 0;
        int32_t cu = -1, cs = -1;
        for (;;) { cu = next_cand(v, cu, cs, 0, row_ptr, col, arc_street, w, dist, &cs); if (cu < 0) break; ++nc; }
        *n_cand = nc;
    }
    int32_t u = next_cand(v, -1, -1, 1, row_ptr, col, arc_street, w, dist, &s);
    if (u >= 0) { *street_out = s; return u; }
    /* plateau search */
    int32_t seen[PLATEAU_MAX], first_u[PLATEAU_MAX], first_s[PLATEAU_MAX];
    int n_seen = 1, head = 0;
    seen[0] = v; first_u[0] = -1; first_s[0] = -1;
    while (head < n_seen) {
        int32_t x = seen[head];
        int32_t cu = -1, cs = -1;
        for (;;) {
            cu = next_cand(x, cu, cs, 0, row_ptr, col, arc_street, w, dist, &cs);
            if (cu < 0) break;
            int dup = 0;
            for (int k = 0; k < n_seen; ++k) if (seen[k] == cu) { dup = 1; break; }
            if (dup) continue;
            if (n_seen >= PLATEAU_MAX) { *street_out = -1; return -1; }
            seen[n_seen] = cu;
            first_u[n_seen] = head == 0 ? cu : first_u[head];
            first_s[n_seen] = head == 0 ? cs : first_s[head];
            int32_t dummy;
            if (cu == root || next_cand(cu, -1, -1, 1, row_ptr, col, arc_street, w, dist, &dummy) >= 0) {
                *street_out = first_s[n_seen];
                return first_u[n_seen];
            }
            ++n_seen;
        }
        ++head;
    }
    *street_out = -1;
    return -1;
}

/* streetnetwork.py:108-109: one tree.  pred[src] = -1, pred[unreachable] = -1,
 * tied[v] = 1 when v had more than one candidate predecessor (tie audit). */
int o_sssp(int32_t V, const int64_t *row_ptr, const int32_t *col,
           const int32_t *arc_street, const double *w, int32_t src,
           double *dist_out, int32_t *pred_out, uint8_t *tied_out)
{
    int64_t A = row_ptr[V];
    heap_item *heap = (heap_item *)malloc(sizeof(heap_item) * (size_t)(A + 1));
    uint8_t *done = (uint8_t *)malloc((size_t)V);
    if (!heap || !done) { free(heap); free(done); return -1; }
    dijkstra(V, row_ptr, col, arc_street, w, src, dist_out, heap, done);
    if (pred_out) {
        for (int32_t v = 0; v < V; ++v) {
            int32_t s; int nc = 0;
            if (v == src || isinf(dist_out[v])) { pred_out[v] = -1; if (tied_out) tied_out[v] = 0; continue; }
            pred_out[v] = canon_pred_root(v, src, row_ptr, col, arc_street, w, dist_out, &s, &nc);
            if (tied_out) tied_out[v] = nc > 1;
        }
    }
    free(heap); free(done);
    return 0;
}

/*
 * simulation.py:68-88: one day for one (virtual) rank.  Trips arrive grouped:
 * roots[i] owns targets[off[i] .. off[i+1]).  load_out (uint32[S]) is zeroed
 * here (:68) and receives trip_volume per hop (:86-88).  counters (may be
 * NULL): [0] hops walked, [1] unreachable trips, [2] trips through a tied node,
 * [3] stuck walks (no predecessor found; never on positive weights).
 * nthreads worker threads (pthreads) pull roots from a shared counter; integer
 * adds commute, so the result does not depend on the schedule.
 */
typedef struct {
    int32_t V; const int64_t *row_ptr; const int32_t *col; const int32_t *arc_street;
    const double *w; int64_t n_roots; const int32_t *roots; const int64_t *off;
    const int32_t *targets; uint32_t trip_volume; uint32_t *load; int64_t *next;
    int64_t hops, unreachable, tied_trips, stuck; int err;
} day_job;

static void *day_worker(void *arg)
{
    day_job *j = (day_job *)arg;
    int32_t V = j->V;
    int64_t A = j->row_ptr[V];
    heap_item *heap = (heap_item *)malloc(sizeof(heap_item) * (size_t)(A + 1));
    uint8_t *done = (uint8_t *)malloc((size_t)V);
    double *dist = (double *)malloc(sizeof(double) * (size_t)V);
    if (!heap || !done || !dist) { j->err = -1; free(heap); free(done); free(dist); return NULL; }
    for (;;) {
        int64_t i = __atomic_fetch_add(j->next, 1, __ATOMIC_RELAXED);
        if (i >= j->n_roots) break;
        int32_t root = j->roots[i];
        dijkstra(V, j->row_ptr, j->col, j->arc_street, j->w, root, dist, heap, done);
        for (int64_t t = j->off[i]; t < j->off[i + 1]; ++t) {
            int32_t here = j->targets[t];
            if (isinf(dist[here])) { ++j->unreachable; continue; }           /* :80 */
            int tied = 0;
            int64_t guard = 0;
            while (here != root) {                                           /* :83 */
                int32_t s; int nc = 0;
                int32_t nxt = canon_pred_root(here, root, j->row_ptr, j->col, j->arc_street, j->w, dist, &s, &nc);
                if (nxt < 0 || ++guard > (int64_t)V) { ++j->stuck; break; }
                if (nc > 1) tied = 1;
                __atomic_fetch_add(&j->load[s], j->trip_volume, __ATOMIC_RELAXED);   /* :86-88 */
                ++j->hops;
                here = nxt;                                                  /* :85 */
            }
            j->tied_trips += tied;
        }
    }
    free(heap); free(done); free(dist);
    return NULL;
}

int o_day(int32_t V, int64_t S, const int64_t *row_ptr, const int32_t *col,
          const int32_t *arc_street, const double *w, int64_t n_roots,
          const int32_t *roots, const int64_t *off, const int32_t *targets,
          uint32_t trip_volume, uint32_t *load_out, int64_t *counters, int nthreads)
{
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 1024) nthreads = 1024;
    memset(load_out, 0, sizeof(uint32_t) * (size_t)S);
    int64_t next = 0;
    day_job *jobs = (day_job *)calloc((size_t)nthreads, sizeof(day_job));
    pthread_t *th = (pthread_t *)calloc((size_t)nthreads, sizeof(pthread_t));
    if (!jobs || !th) { free(jobs); free(th); return -1; }
    for (int k = 0; k < nthreads; ++k) {
        jobs[k] = (day_job){V, row_ptr, col, arc_street, w, n_roots, roots, off, targets,
                            trip_volume, load_out, &next, 0, 0, 0, 0, 0};
    }
    int started = 0, err = 0;
    for (int k = 1; k < nthreads; ++k) {
        if (pthread_create(&th[k], NULL, day_worker, &jobs[k]) != 0) break;
        started = k;
    }
    day_worker(&jobs[0]);
    for (int k = 1; k <= started; ++k) pthread_join(th[k], NULL);
    int64_t c[4] = {0, 0, 0, 0};
    for (int k = 0; k <= started; ++k) {
        c[0] += jobs[k].hops; c[1] += jobs[k].unreachable; c[2] += jobs[k].tied_trips; c[3] += jobs[k].stuck;
        if (jobs[k].err) err = -1;
    }
    if (counters) memcpy(counters, c, sizeof c);
    free(jobs); free(th);
    return err;
}

/* utils.py:26-34 as used at streets4mpi.py:100: acc += add (uint32 wrap) */
void o_merge(int64_t S, uint32_t *acc, const uint32_t *add)
{
    for (int64_t i = 0; i < S; ++i) acc[i] += add[i];
}

typedef struct { uint32_t load; int32_t idx; } rank_item;
static int rank_cmp(const void *a, const void *b)
{
    const rank_item *x = (const rank_item *)a, *y = (const rank_item *)b;
    if (x->load != y->load) return x->load < y->load ? -1 : 1;
    return (x->idx > y->idx) - (x->idx < y->idx);       /* stable by street_index */
}

/*
 * simulation.py:91-109 executed literally (the always-true `if not` of :101/:105
 * included) on streets ranked by (cumulative load, street_index).  adaptable
 * (may be NULL = all ones) masks streets whose speed change is invisible to the
 * weight loop (orientation aliasing).  Returns #decreased in
 * counts[0], #increased in counts[1].
 */
int o_road_construction(int64_t S, const uint32_t *cumulative, int32_t *max_speed,
                        const uint8_t *adaptable, int64_t *counts)
{
    rank_item *r = (rank_item *)malloc(sizeof(rank_item) * (size_t)(S > 0 ? S : 1));
    if (!r) return -1;
    for (int64_t i = 0; i < S; ++i) { r[i].load = cumulative[i]; r[i].idx = (int32_t)i; }
    qsort(r, (size_t)S, sizeof(rank_item), rank_cmp);
    double dec_limit = 0.15 * (double)S, inc_limit = 0.95 * (double)S;
    int64_t ndec = 0, ninc = 0;
    for (int64_t i = 0; i < S; ++i) {
        if ((double)i <= dec_limit) {
            int32_t s = r[i].idx;
            if (!adaptable || adaptable[s]) {
                int32_t x = max_speed[s] - 20; if (x > 140) x = 140; if (x < 1) x = 1;
                max_speed[s] = x;
            }
            ++ndec; dec_limit += 1.0;
        }
        int64_t j = S - i - 1;
        if ((double)j >= inc_limit) {
            int32_t s = r[j].idx;
            if (!adaptable || adaptable[s]) {
                int32_t x = max_speed[s] + 20; if (x > 140) x = 140; if (x < 1) x = 1;
                max_speed[s] = x;
            }
            ++ninc; inc_limit -= 1.0;
        }
        if (dec_limit >= inc_limit) break;
    }
    if (counts) { counts[0] = ndec; counts[1] = ninc; }
    free(r);
    return 0;
}
