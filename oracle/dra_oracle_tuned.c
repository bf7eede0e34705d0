/*
 * dra_oracle_tuned.c — a CPU port of spec/ALLOCATION.md written for SPEED.  TEST INFRASTRUCTURE ONLY, like
 * dra_oracle.c (which stays the parity checker: written for obviousness).  This one exists so that the CPU baseline
 * reported beside the GPU numbers is not a strawman:
 *   - no allocation inside the timed call (scratch grows once and is kept),
 *   - the stable bucketing by node is parallel (per-thread histograms over contiguous chunks, prefix, scatter),
 *   - nodes are taken from an atomic counter by a persistent pool whose threads also did the bucketing,
 *   - plain MIG / GPU / SHARED claims take a lean step: shift/AND fit map + count-trailing-zeros, and the
 *     monotone "this shape already failed on this node" memo (a batch only takes capacity, spec §2),
 *   - nodes with co-location groups or selectors fall back to the plain oracle's node_process (same file, included).
 * Same spec, same bytes out: tests/test_oracle_tuned.py compares it with the plain oracle on every workload class.
 * Build: -O3 -march=native (oracle/Makefile).   PARITY UNPINNED for the search itself, as everywhere.
 */
#include "dra_oracle.c"

#include <stdatomic.h>

typedef struct tuned_job {
    dra_gpu_rec* gpus; uint32_t n_gpu; const uint32_t* node_off; uint32_t n_node;
    const dra_profile_tbl* tbl; const dra_claim_rec* claims; uint32_t n_claim;
    const uint32_t* out_off; dra_out_rec* out; uint32_t n_out;
    int n_thr;
} tuned_job;

static struct {
    pthread_t* th; int n;                        /* workers; the caller is thread 0 */
    _Atomic uint64_t gen; _Atomic int stop; uint64_t start_gen;
    _Atomic uint32_t bar_cnt, bar_gen, next_node, done;
    tuned_job job;
    /* scratch, grown on demand, kept */
    uint32_t* idx; size_t idx_cap;               /* claim indices grouped by node */
    uint32_t* cnt; size_t cnt_cap;               /* [n_thr][n_node+1] */
    uint32_t* off; size_t off_cap;               /* [n_node+2] */
} tp;

static void tbarrier(int n)
{
    uint32_t g = atomic_load_explicit(&tp.bar_gen, memory_order_acquire);
    if (atomic_fetch_add_explicit(&tp.bar_cnt, 1u, memory_order_acq_rel) == (uint32_t)n - 1) {
        atomic_store_explicit(&tp.bar_cnt, 0u, memory_order_relaxed);
        atomic_fetch_add_explicit(&tp.bar_gen, 1u, memory_order_release);
    } else {
        uint32_t spins = 0;
        while (atomic_load_explicit(&tp.bar_gen, memory_order_acquire) == g) { if (++spins < 4000u) __builtin_ia32_pause(); else sched_yield(); }
    }
}

static inline uint32_t fit16(uint32_t free16, uint32_t size)
{
    uint32_t t = free16, r = size - 1u, s = 1u;
    while (r) { uint32_t m = s < r ? s : r; t &= t >> m; r -= m; s <<= 1; }
    return t;
}

/* one node, lean: returns 0 if the node needs the plain path (groups / selectors) */
static int node_lean(const tuned_job* j, uint32_t n, const uint32_t* idx, uint32_t cnt)
{
    const int have_off = j->out_off != NULL;
    dra_gpu_rec* G = j->gpus + j->node_off[n];
    const uint32_t g0 = j->node_off[n], ng = j->node_off[n + 1] - g0;
    for (uint32_t k = 0; k < cnt; k++) {
        const dra_claim_rec* c = &j->claims[idx[k]];
        if (c->kind > DRA_KIND_SHARED) continue;
        if ((c->kind == DRA_KIND_MIG && (c->group != 0 || c->mem_limit_mib != 0)) ||
            (c->kind == DRA_KIND_GPU && c->mem_limit_mib != 0) || (c->kind == DRA_KIND_SHARED && c->group != 0)) return 0;
    }
    uint32_t dead_nocap = 0, dead_bad = 0, gpu_min = 0xFFFFu; uint64_t sh_min = 1ull << 32;
    for (uint32_t k = 0; k < cnt; k++) {
        const uint32_t ci = idx[k];
        const dra_claim_rec* c = &j->claims[ci];
        dra_out_rec* o = &j->out[have_off ? j->out_off[ci] : ci];
        if (claim_bad_shape(c, j->n_node, have_off)) { fail_all(o, 1, c, DRA_ST_INVALID); continue; }
        if (c->kind == DRA_KIND_MIG) {
            const uint32_t pb = 1u << c->profile;
            if ((dead_nocap | dead_bad) & pb) { put(o, DRA_GPU_NONE, 0, 0, c->profile, (dead_bad & pb) ? DRA_ST_BAD_PROFILE : DRA_ST_NO_CAPACITY); continue; }
            int any = 0, done = 0;
            for (uint32_t g = 0; g < ng; g++) {
                dra_gpu_rec* r = &G[g];
                const dra_prof_ent e = j->tbl[r->model].ent[c->profile];
                if ((r->flags & (DRA_GPU_MIG_ENABLED | DRA_GPU_UNAVAILABLE)) != DRA_GPU_MIG_ENABLED || !e.start_mask) continue;
                any = 1;
                if (r->flags & DRA_GPU_FULL_ALLOCATED) continue;
                const uint32_t cand = fit16(~(uint32_t)r->busy & 0xFFFFu, e.size) & e.start_mask;
                if (!cand) continue;
                const uint32_t s = (uint32_t)__builtin_ctz(cand);
                r->busy |= (uint16_t)(((1u << e.size) - 1u) << s);
                put(o, g0 + g, (uint8_t)s, e.size, c->profile, DRA_ST_OK);
                done = 1; break;
            }
            if (!done) {
                if (any) dead_nocap |= pb; else dead_bad |= pb;
                put(o, DRA_GPU_NONE, 0, 0, c->profile, any ? DRA_ST_NO_CAPACITY : DRA_ST_BAD_PROFILE);
            }
        } else if (c->kind == DRA_KIND_GPU) {
            if (c->count >= gpu_min) { fail_all(o, c->count, c, DRA_ST_NO_CAPACITY); continue; }
            uint32_t elig = 0;
            for (uint32_t g = 0; g < ng; g++)
                elig += !(G[g].flags & (DRA_GPU_MIG_ENABLED | DRA_GPU_FULL_ALLOCATED | DRA_GPU_UNAVAILABLE)) && G[g].share_cnt == 0;
            if (elig < c->count) { gpu_min = c->count; fail_all(o, c->count, c, DRA_ST_NO_CAPACITY); continue; }
            uint32_t t = 0;
            for (uint32_t g = 0; g < ng && t < c->count; g++)
                if (!(G[g].flags & (DRA_GPU_MIG_ENABLED | DRA_GPU_FULL_ALLOCATED | DRA_GPU_UNAVAILABLE)) && G[g].share_cnt == 0) {
                    G[g].flags |= DRA_GPU_FULL_ALLOCATED;
                    put(&o[t++], g0 + g, 0, 0, DRA_PROFILE_GPU, DRA_ST_OK);
                }
        } else {
            if ((uint64_t)c->mem_limit_mib >= sh_min) { put(o, DRA_GPU_NONE, 0, 0, DRA_PROFILE_SHARED, DRA_ST_MEM_LIMIT); continue; }
            int done = 0;
            for (uint32_t g = 0; g < ng; g++) {
                dra_gpu_rec* r = &G[g];
                if ((r->flags & (DRA_GPU_MIG_ENABLED | DRA_GPU_FULL_ALLOCATED | DRA_GPU_UNAVAILABLE)) || r->share_cnt == 0xFFFFu ||
                    r->mem_free_mib < c->mem_limit_mib) continue;
                r->mem_free_mib -= c->mem_limit_mib; r->share_cnt++;
                put(o, g0 + g, 0, 0, DRA_PROFILE_SHARED, DRA_ST_OK);
                done = 1; break;
            }
            if (!done) { sh_min = c->mem_limit_mib; put(o, DRA_GPU_NONE, 0, 0, DRA_PROFILE_SHARED, DRA_ST_MEM_LIMIT); }
        }
    }
    return 1;
}

static void tuned_work(int tid)
{
    const tuned_job* j = &tp.job;
    const int T = j->n_thr;
    const uint32_t nb = j->n_node + 1;                       /* bucket n_node = claims naming no node */
    const uint32_t lo = (uint32_t)((uint64_t)j->n_claim * (uint32_t)tid / (uint32_t)T);
    const uint32_t hi = (uint32_t)((uint64_t)j->n_claim * (uint32_t)(tid + 1) / (uint32_t)T);
    uint32_t* mycnt = tp.cnt + (size_t)tid * nb;
    /* 1. per-thread histogram of a contiguous chunk (stability: chunks are in input order) */
    memset(mycnt, 0, nb * sizeof(uint32_t));
    for (uint32_t i = lo; i < hi; i++) { uint32_t n = j->claims[i].node; mycnt[n < j->n_node ? n : j->n_node]++; }
    tbarrier(T);
    /* 2. offsets: thread t owns a slice of the nodes; node totals, then (3) a serial prefix over nodes by thread 0 */
    const uint32_t n0 = (uint32_t)((uint64_t)nb * (uint32_t)tid / (uint32_t)T), n1 = (uint32_t)((uint64_t)nb * (uint32_t)(tid + 1) / (uint32_t)T);
    for (uint32_t n = n0; n < n1; n++) {
        uint32_t run = 0;
        for (int t = 0; t < T; t++) { uint32_t v = tp.cnt[(size_t)t * nb + n]; tp.cnt[(size_t)t * nb + n] = run; run += v; }
        tp.off[n + 1] = run;
    }
    tbarrier(T);
    if (tid == 0) { tp.off[0] = 0; for (uint32_t n = 0; n < nb; n++) tp.off[n + 1] += tp.off[n]; }
    tbarrier(T);
    /* 4. scatter the chunk's claim indices */
    for (uint32_t i = lo; i < hi; i++) {
        uint32_t n = j->claims[i].node; n = n < j->n_node ? n : j->n_node;
        tp.idx[tp.off[n] + mycnt[n]++] = i;
    }
    tbarrier(T);
    /* 5. nodes from an atomic counter */
    const int have_off = j->out_off != NULL;
    for (;;) {
        uint32_t n = atomic_fetch_add_explicit(&tp.next_node, 1u, memory_order_relaxed);
        if (n > j->n_node) break;
        const uint32_t* idx = tp.idx + tp.off[n]; const uint32_t cnt = tp.off[n + 1] - tp.off[n];
        if (n == j->n_node) {                                 /* claims naming no node: INVALID (spec §3) */
            for (uint32_t k = 0; k < cnt; k++) { const uint32_t ci = idx[k]; put(&j->out[have_off ? j->out_off[ci] : ci], DRA_GPU_NONE, 0, 0, out_profile(&j->claims[ci]), DRA_ST_INVALID); }
            continue;
        }
        if (!cnt) continue;
        if (g_nsel == 0 && node_lean(j, n, idx, cnt)) continue;
        node_job nj; memset(&nj, 0, sizeof nj);
        nj.gpus = j->gpus + j->node_off[n]; nj.g0 = j->node_off[n]; nj.ng = j->node_off[n + 1] - j->node_off[n];
        nj.tbl = j->tbl; nj.claims = j->claims; nj.idx = idx; nj.cnt = cnt; nj.n_node = j->n_node;
        nj.out_off = j->out_off; nj.out = j->out; nj.commit = 1;
        node_process(&nj);
    }
}

static void* tuned_worker(void* arg)
{
    const int tid = (int)(intptr_t)arg;
    uint64_t seen = tp.start_gen;                       /* batches before this worker existed are not its business */
    for (;;) {
        uint32_t spins = 0;
        while (atomic_load_explicit(&tp.gen, memory_order_acquire) == seen && !atomic_load(&tp.stop)) {
            if (++spins < 4000u) __builtin_ia32_pause(); else { sched_yield(); }
        }
        if (atomic_load(&tp.stop)) return NULL;
        seen = atomic_load_explicit(&tp.gen, memory_order_acquire);
        if (tid < tp.job.n_thr) tuned_work(tid);
        atomic_fetch_add_explicit(&tp.done, 1u, memory_order_release);
    }
}

static void tuned_shutdown(void)
{
    if (!tp.n) return;
    atomic_store(&tp.stop, 1);
    for (int t = 0; t < tp.n; t++) pthread_join(tp.th[t], NULL);
    free(tp.th); tp.th = NULL; tp.n = 0; atomic_store(&tp.stop, 0);
}

int dra_oracle_tuned_threads(int n_threads)        /* (re)creates the pool: n_threads - 1 workers spinning between batches */
{
    if (n_threads < 1) n_threads = 1;
    if (tp.n == n_threads - 1) return 0;
    tuned_shutdown();
    if (n_threads == 1) return 0;
    tp.th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)(n_threads - 1));
    if (!tp.th) return -1;
    static int at_exit_set;
    if (!at_exit_set) { atexit(tuned_shutdown); at_exit_set = 1; }
    tp.start_gen = atomic_load(&tp.gen);
    for (int t = 0; t < n_threads - 1; t++)
        if (pthread_create(&tp.th[t], NULL, tuned_worker, (void*)(intptr_t)(t + 1))) { tp.n = t; tuned_shutdown(); return -1; }
    tp.n = n_threads - 1;
    return 0;
}

/* Same contract as dra_oracle_allocate_mt.  Validation of the arguments (one pass over the claims) is part of the call. */
int dra_oracle_tuned_allocate(dra_gpu_rec* gpus, uint32_t n_gpu, const uint32_t* node_off, uint32_t n_node,
                              const dra_profile_tbl* tbl, const dra_claim_rec* claims, uint32_t n_claim,
                              const uint32_t* out_off, dra_out_rec* out, uint32_t n_out, int n_threads)
{
    if (n_threads < 1) n_threads = 1;
    if ((uint32_t)n_threads > n_node + 1) n_threads = (int)n_node + 1;
    if (dra_oracle_tuned_threads(n_threads)) return -1;
    if (check_inventory(gpus, n_gpu, node_off, n_node, tbl)) return -1;
    const int have_off = out_off != NULL;
    if (!have_off) { if (n_out < n_claim) return -1; }
    else for (uint32_t i = 0; i < n_claim; i++) {
        uint32_t sl = claim_slots(&claims[i], n_node, 1);
        if (out_off[i] > n_out || sl > n_out - out_off[i]) return -1;
    }
    const size_t nb = (size_t)n_node + 2;
    if (tp.idx_cap < (size_t)n_claim + 1) { free(tp.idx); tp.idx_cap = (size_t)n_claim * 2 + 64; tp.idx = (uint32_t*)malloc(tp.idx_cap * 4); }
    if (tp.cnt_cap < nb * (size_t)n_threads) { free(tp.cnt); tp.cnt_cap = nb * (size_t)n_threads * 2; tp.cnt = (uint32_t*)malloc(tp.cnt_cap * 4); }
    if (tp.off_cap < nb + 1) { free(tp.off); tp.off_cap = nb * 2 + 8; tp.off = (uint32_t*)malloc(tp.off_cap * 4); }
    if (!tp.idx || !tp.cnt || !tp.off) return -1;
    tuned_job* j = &tp.job;
    j->gpus = gpus; j->n_gpu = n_gpu; j->node_off = node_off; j->n_node = n_node; j->tbl = tbl; j->claims = claims;
    j->n_claim = n_claim; j->out_off = out_off; j->out = out; j->n_out = n_out; j->n_thr = n_threads;
    atomic_store_explicit(&tp.next_node, 0u, memory_order_relaxed);
    atomic_store_explicit(&tp.done, 0u, memory_order_relaxed);
    atomic_store_explicit(&tp.bar_cnt, 0u, memory_order_relaxed);
    atomic_fetch_add_explicit(&tp.gen, 1u, memory_order_release);
    tuned_work(0);
    { uint32_t spins = 0; while (atomic_load_explicit(&tp.done, memory_order_acquire) != (uint32_t)tp.n) { if (++spins < 4000u) __builtin_ia32_pause(); else sched_yield(); } }
    return 0;
}
