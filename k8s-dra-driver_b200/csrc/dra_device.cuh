// dra_device.cuh — sm_100a kernels of the allocation hot path (spec/ALLOCATION.md).
//
// Pipeline of one Allocate batch (all integer / bitwise, HBM-bound, no tensor cores):
//   k_bucket_hist    stable counting sort, pass 1: per-tile node histogram + stable intra-tile rank
//                    (__match_any_sync gives every lane its rank among equal-node lanes)
//   k_bucket_scan    pass 2: per-node exclusive scan over tiles, node offsets (CSR claim_off)
//   k_bucket_scatter pass 3: claims -> node-sorted array, first OutRec slot embedded in the record
//   k_pack           one warp per node, one lane per GPU: the node's GpuRecs and its claim span are
//                    staged into shared memory with 1-D TMA bulk copies (cp.async.bulk + mbarrier), each
//                    claim is tested on all GPUs at once, __ballot_sync + __ffs pick lowest GPU / lowest
//                    start (first-fit), the winner lane updates its register-resident occupancy mask.
//   k_unsuitable     one warp per (pod, candidate node) pair, same step function on a snapshot
//   k_dealloc        inverse updates with atomics
//
// What this replaces in the reference: the per-node search behind Allocate()/UnsuitableNodes() that
// north_star names (absent from the snapshot), over the device model of
// cmd/nvidia-dra-plugin/deviceinfo.go:30-64,199-204 and the placement enumeration of nvlib.go:244-295.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/dra_alloc.h"

namespace dra {

constexpr uint32_t FULLMASK = 0xFFFFFFFFu;
constexpr uint32_t SEG = 32;          // claims per TMA segment
constexpr uint32_t RING = 4;          // ring slots per warp
constexpr uint32_t ERR_OUT_RANGE = 1u;    // index into the error-flag words
constexpr uint32_t ERR_NOT_SORTED = 2u;
constexpr uint32_t ERR_WORDS = 4u;

// Error flags: `dev` (device memory) gates later kernels of the same batch, `host` (mapped pinned
// memory) is what the host reads after the stream drains.  Idempotent plain stores, errors are rare.
struct Err {
    uint32_t* dev; volatile uint32_t* host;
    __device__ __forceinline__ void set(uint32_t which) const { dev[which] = 1u; host[which] = 1u; __threadfence_system(); }
};

// ---- record views ---------------------------------------------------------------------------------
// GpuRec   as uint4: x = busy | flags<<16 | model<<24 ; y = mem_free_mib ; z = node ; w = share|rsvd<<16
// ClaimRec as uint4: x = kind | profile<<8 | count<<16 ; y = node (sorted copy: first out slot) ;
//                    z = mem_limit_mib ; w = group
// OutRec   as uint2: x = gpu ; y = start | size<<8 | profile<<16 | status<<24

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t addr = smem_u32(bar), ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    } while (!ok);
}
// 1-D TMA bulk copy global -> shared, completion counted on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_1d(void* sdst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(sdst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ uint32_t lanemask_lt() {
    uint32_t m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m;
}

// bit i of the result: slices i .. i+size-1 of `free16` are all free (log-step AND of shifted copies)
__device__ __forceinline__ uint32_t fit_map(uint32_t free16, uint32_t size) {
    uint32_t t = free16, r = size - 1u, s = 1u;
    while (r) { uint32_t m = s < r ? s : r; t &= t >> m; r -= m; s <<= 1; }
    return t;
}

__device__ __forceinline__ uint32_t meta(uint32_t start, uint32_t size, uint32_t prof, uint32_t st) {
    return start | (size << 8) | (prof << 16) | (st << 24);
}

// ---- one lane = one GPU of the node ---------------------------------------------------------------
struct Lane {
    uint32_t busy, flags, model, mem, share;
    bool valid;
    __device__ __forceinline__ void load(uint4 r, bool v) {
        valid = v;
        busy = v ? (r.x & 0xFFFFu) : 0xFFFFu;
        flags = v ? ((r.x >> 16) & 0xFFu) : DRA_GPU_UNAVAILABLE;
        model = v ? ((r.x >> 24) & 0xFu) : 0u;
        mem = v ? r.y : 0u;
        share = v ? (r.w & 0xFFFFu) : 0u;
    }
    __device__ __forceinline__ uint4 store(uint4 r) const {
        r.x = (r.x & 0xFF000000u) | (flags << 16) | busy;
        r.y = mem;
        r.w = (r.w & 0xFFFF0000u) | share;
        return r;
    }
};

// Per-warp monotone failure memo (spec §2: a batch only ever takes capacity, so a request shape that
// failed once on this node fails for the rest of the batch).  Lets failing claims skip the ballot.
struct Dead {
    uint32_t nocap = 0, bad = 0;     // per MIG profile
    uint32_t gpu_min = 0xFFFFu;      // smallest GPU count that failed
    uint64_t sh_min = 1ull << 32;    // smallest SHARED limit that failed
};

// Sink for Allocate: writes OutRecs.
struct OutSink {
    uint2* out; uint32_t n_out; Err err; uint32_t lane;
    __device__ __forceinline__ bool range(uint32_t dst, uint32_t slots) const {
        if (dst > n_out || slots > n_out - dst) { if (lane == 0) err.set(ERR_OUT_RANGE); return false; }
        return true;
    }
    __device__ __forceinline__ void put(uint32_t idx, uint32_t gpu, uint32_t m) const {
        out[idx] = make_uint2(gpu, m);
    }
    __device__ __forceinline__ bool in_range(uint32_t idx) const { return idx < n_out; }
    __device__ __forceinline__ void fail(uint32_t dst, uint32_t slots, uint32_t prof, uint32_t st) {
        if (lane < slots) put(dst + lane, DRA_GPU_NONE, meta(0, 0, prof, st));
    }
    __device__ __forceinline__ void mark_failed() {}
    __device__ __forceinline__ bool stop() const { return false; }
};

// Sink for UnsuitableNodes: only remembers whether anything failed.
struct FlagSink {
    bool failed = false;
    __device__ __forceinline__ bool range(uint32_t, uint32_t) const { return true; }
    __device__ __forceinline__ void put(uint32_t, uint32_t, uint32_t) const {}
    __device__ __forceinline__ bool in_range(uint32_t) const { return true; }
    __device__ __forceinline__ void fail(uint32_t, uint32_t, uint32_t, uint32_t) { failed = true; }
    __device__ __forceinline__ void mark_failed() { failed = true; }
    __device__ __forceinline__ bool stop() const { return failed; }
};

__device__ __forceinline__ bool claim_invalid(uint32_t kind, uint32_t prof, uint32_t count, bool have_off) {
    if (kind > DRA_KIND_SHARED) return true;
    if (kind == DRA_KIND_GPU) return count == 0 || count > DRA_MAX_COUNT || (!have_off && count != 1);
    if (kind == DRA_KIND_MIG) return prof >= DRA_MAX_PROFILES;
    return false;
}

// One claim (or one co-location run starting at k) on the node held by this warp.  Returns the number of
// claims consumed.  `get(m)` returns claim m of the node's span (uniform address, all lanes).
template <class Get, class Sink>
__device__ __forceinline__ uint32_t node_step(Lane& L, Dead& D, uint32_t lane, uint32_t g0,
                                              const uint32_t* __restrict__ tbl_s, Get get, uint32_t k,
                                              uint32_t cnt, Sink& sink, bool have_off) {
    const uint4 c = get(k);
    const uint32_t kind = c.x & 0xFFu, prof = (c.x >> 8) & 0xFFu, count = c.x >> 16;
    const uint32_t dst = c.y, mem = c.z, group = c.w;
    constexpr uint32_t BLOCKED = DRA_GPU_MIG_ENABLED | DRA_GPU_FULL_ALLOCATED | DRA_GPU_UNAVAILABLE;

    if (claim_invalid(kind, prof, count, have_off)) {               // spec §3
        const uint32_t op = kind == DRA_KIND_GPU ? DRA_PROFILE_GPU
                          : kind == DRA_KIND_SHARED ? DRA_PROFILE_SHARED : prof;
        if (sink.range(dst, 1)) sink.fail(dst, 1, op, DRA_ST_INVALID);
        return 1;
    }

    if (kind == DRA_KIND_MIG && group == 0) {                       // spec §5
        if (!sink.range(dst, 1)) return 1;
        const uint32_t pbit = 1u << prof;
        if ((D.bad | D.nocap) & pbit) {
            sink.fail(dst, 1, prof, (D.bad & pbit) ? DRA_ST_BAD_PROFILE : DRA_ST_NO_CAPACITY);
            return 1;
        }
        const uint32_t e = tbl_s[L.model * DRA_MAX_PROFILES + prof];
        const uint32_t smask = e >> 16, size = e & 0xFFu;
        const bool offers = L.valid && (L.flags & (DRA_GPU_MIG_ENABLED | DRA_GPU_UNAVAILABLE)) == DRA_GPU_MIG_ENABLED
                            && smask != 0;
        const bool elig = offers && !(L.flags & DRA_GPU_FULL_ALLOCATED);
        const uint32_t cand = elig ? (fit_map(~L.busy & 0xFFFFu, size) & smask) : 0u;
        const uint32_t b = __ballot_sync(FULLMASK, cand != 0);
        if (b) {
            if (lane == (uint32_t)__ffs(b) - 1u) {                  // lowest GPU
                const uint32_t st = (uint32_t)__ffs(cand) - 1u;     // lowest start
                L.busy |= ((1u << size) - 1u) << st;
                sink.put(dst, g0 + lane, meta(st, size, prof, DRA_ST_OK));
            }
        } else {
            const bool any = __ballot_sync(FULLMASK, offers) != 0;
            if (any) D.nocap |= pbit; else D.bad |= pbit;
            sink.fail(dst, 1, prof, any ? DRA_ST_NO_CAPACITY : DRA_ST_BAD_PROFILE);
        }
        return 1;
    }

    if (kind == DRA_KIND_MIG) {                                     // spec §6: co-location run
        uint32_t e_ = k + 1;
        const uint32_t lim = (cnt - k) < DRA_MAX_GROUP ? cnt : k + DRA_MAX_GROUP;
        while (e_ < lim) {
            const uint4 cm = get(e_);
            const uint32_t km = cm.x & 0xFFu, pm = (cm.x >> 8) & 0xFFu;
            if (km != DRA_KIND_MIG || cm.w != group || pm >= DRA_MAX_PROFILES) break;
            ++e_;
        }
        // every lane tries the whole run on a private copy of its own GPU
        uint32_t tb = L.busy;
        bool ok = L.valid && (L.flags & BLOCKED) == DRA_GPU_MIG_ENABLED;
        for (uint32_t m = k; m < e_; ++m) {
            const uint32_t pm = (get(m).x >> 8) & 0xFFu;
            const uint32_t en = tbl_s[L.model * DRA_MAX_PROFILES + pm];
            const uint32_t sm = en >> 16, sz = en & 0xFFu;
            const uint32_t cd = sm ? (fit_map(~tb & 0xFFFFu, sz) & sm) : 0u;
            ok = ok && cd != 0;
            if (cd) tb |= ((1u << sz) - 1u) << ((uint32_t)__ffs(cd) - 1u);
        }
        const uint32_t b = __ballot_sync(FULLMASK, ok);
        if (b) {
            if (lane == (uint32_t)__ffs(b) - 1u) {                  // winner replays and emits
                uint32_t nb = L.busy;
                for (uint32_t m = k; m < e_; ++m) {
                    const uint4 cm = get(m);
                    const uint32_t pm = (cm.x >> 8) & 0xFFu;
                    const uint32_t en = tbl_s[L.model * DRA_MAX_PROFILES + pm];
                    const uint32_t sm = en >> 16, sz = en & 0xFFu;
                    const uint32_t st = (uint32_t)__ffs(fit_map(~nb & 0xFFFFu, sz) & sm) - 1u;
                    nb |= ((1u << sz) - 1u) << st;
                    if (sink.in_range(cm.y)) sink.put(cm.y, g0 + lane, meta(st, sz, pm, DRA_ST_OK));
                }
                L.busy = nb;
            }
        } else {
            if (lane < e_ - k) {
                const uint4 cm = get(k + lane);
                if (sink.in_range(cm.y))
                    sink.put(cm.y, DRA_GPU_NONE, meta(0, 0, (cm.x >> 8) & 0xFFu, DRA_ST_GROUP));
            }
            sink.mark_failed();
        }
        // members whose slot is out of range
        for (uint32_t m = k; m < e_; ++m) (void)sink.range(get(m).y, 1);
        return e_ - k;
    }

    if (kind == DRA_KIND_GPU) {                                     // spec §4
        if (!sink.range(dst, count)) return 1;
        if (count >= D.gpu_min) { sink.fail(dst, count, DRA_PROFILE_GPU, DRA_ST_NO_CAPACITY); return 1; }
        const bool elig = L.valid && !(L.flags & BLOCKED) && L.share == 0;
        const uint32_t b = __ballot_sync(FULLMASK, elig);
        if ((uint32_t)__popc(b) >= count) {
            const uint32_t r = (uint32_t)__popc(b & lanemask_lt());
            if (elig && r < count) {
                L.flags |= DRA_GPU_FULL_ALLOCATED;
                sink.put(dst + r, g0 + lane, meta(0, 0, DRA_PROFILE_GPU, DRA_ST_OK));
            }
        } else {
            D.gpu_min = count;
            sink.fail(dst, count, DRA_PROFILE_GPU, DRA_ST_NO_CAPACITY);
        }
        return 1;
    }

    // SHARED, spec §7
    if (!sink.range(dst, 1)) return 1;
    if ((uint64_t)mem >= D.sh_min) { sink.fail(dst, 1, DRA_PROFILE_SHARED, DRA_ST_MEM_LIMIT); return 1; }
    {
        const bool elig = L.valid && !(L.flags & BLOCKED) && L.share < 0xFFFFu && L.mem >= mem;
        const uint32_t b = __ballot_sync(FULLMASK, elig);
        if (b) {
            if (lane == (uint32_t)__ffs(b) - 1u) {
                L.mem -= mem; L.share += 1;
                sink.put(dst, g0 + lane, meta(0, 0, DRA_PROFILE_SHARED, DRA_ST_OK));
            }
        } else {
            D.sh_min = mem;
            sink.fail(dst, 1, DRA_PROFILE_SHARED, DRA_ST_MEM_LIMIT);
        }
    }
    return 1;
}

// ====================================================================================================
// bucketing: stable counting sort of claims by node
// ====================================================================================================

// grid = n_tiles, block = 32.  Tile t owns claims [t*T, (t+1)*T).  Dynamic smem: (n_node+1) u16 counters.
// hist[t][n] = number of tile-t claims on node n (bucket n_node = claims naming no node);
// rank[i]    = number of earlier claims of the same tile with the same node (stable rank).
__global__ void __launch_bounds__(32)
k_bucket_hist(const uint4* __restrict__ claims, uint32_t n_claim, uint32_t n_node, uint32_t T,
              uint32_t* __restrict__ hist, uint16_t* __restrict__ rank) {
    extern __shared__ uint16_t cnt[];
    const uint32_t lane = threadIdx.x, nb = n_node + 1;
    for (uint32_t n = lane; n < nb; n += 32) cnt[n] = 0;
    __syncwarp();
    const uint32_t base = blockIdx.x * T;
    const uint32_t end = min(n_claim, base + T);
    for (uint32_t i0 = base; i0 < end; i0 += 32) {
        const uint32_t i = i0 + lane;
        const bool act = i < end;
        uint32_t key = 0xFFFFFFFFu;
        if (act) { key = __ldg(&claims[i]).y; key = key < n_node ? key : n_node; }
        const uint32_t m = __match_any_sync(FULLMASK, key);
        const uint32_t r = (uint32_t)__popc(m & lanemask_lt());
        uint32_t old = 0;
        if (act) { old = cnt[key]; rank[i] = (uint16_t)(old + r); }
        __syncwarp();
        if (act && r == 0) cnt[key] = (uint16_t)(old + (uint32_t)__popc(m));
        __syncwarp();
    }
    uint32_t* h = hist + (size_t)blockIdx.x * nb;
    for (uint32_t n = lane; n < nb; n += 32) h[n] = cnt[n];
}

// one CTA of 1024 threads.  In place: hist[t][n] <- sum_{t' < t} hist[t'][n];
// claim_off[n] <- sum_{n' < n} total[n'] for n in [0, n_node+1]  (claim_off[n_node+1] = n_claim).
__global__ void __launch_bounds__(1024)
k_bucket_scan(uint32_t* __restrict__ hist, uint32_t n_tiles, uint32_t n_node,
              uint32_t* __restrict__ claim_off) {
    __shared__ uint32_t wsum[32];
    __shared__ uint32_t carry_s, total_s;
    const uint32_t nb = n_node + 1, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (uint32_t n0 = 0; n0 < nb; n0 += 1024) {
        const uint32_t n = n0 + tid;
        uint32_t run = 0;
        if (n < nb) {
            uint32_t t = 0;
            for (; t + 4 <= n_tiles; t += 4) {
                uint32_t* p = hist + (size_t)t * nb + n;
                const uint32_t v0 = p[0], v1 = p[nb], v2 = p[2 * (size_t)nb], v3 = p[3 * (size_t)nb];
                p[0] = run; run += v0; p[nb] = run; run += v1;
                p[2 * (size_t)nb] = run; run += v2; p[3 * (size_t)nb] = run; run += v3;
            }
            for (; t < n_tiles; ++t) { uint32_t* p = hist + (size_t)t * nb + n; const uint32_t v = *p; *p = run; run += v; }
        }
        // block-wide exclusive scan of `run`
        uint32_t x = run;
        #pragma unroll
        for (int d = 1; d < 32; d <<= 1) { uint32_t y = __shfl_up_sync(FULLMASK, x, d); if (lane >= (uint32_t)d) x += y; }
        if (lane == 31) wsum[wid] = x;
        __syncthreads();
        if (wid == 0) {
            const uint32_t w = wsum[lane];
            uint32_t ws = w;
            #pragma unroll
            for (int d = 1; d < 32; d <<= 1) { uint32_t y = __shfl_up_sync(FULLMASK, ws, d); if (lane >= (uint32_t)d) ws += y; }
            wsum[lane] = ws - w;                       // exclusive warp offsets
            if (lane == 31) total_s = ws;
        }
        __syncthreads();
        const uint32_t excl = carry_s + wsum[wid] + (x - run);
        if (n < nb) claim_off[n] = excl;
        __syncthreads();
        if (tid == 0) carry_s += total_s;
        __syncthreads();
    }
    if (tid == 0) claim_off[nb] = carry_s;
}

// thread per claim.  sorted[dest] = claim with .y replaced by its first OutRec slot.
// Claims that name no node (bucket n_node) get their INVALID record here and are not sorted in.
__global__ void __launch_bounds__(256)
k_bucket_scatter(const uint4* __restrict__ claims, uint32_t n_claim, uint32_t n_node, uint32_t T,
                 const uint32_t* __restrict__ hist, const uint16_t* __restrict__ rank,
                 const uint32_t* __restrict__ claim_off, const uint32_t* __restrict__ out_off,
                 uint4* __restrict__ sorted, uint2* __restrict__ out, uint32_t n_out, Err err) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_claim) return;
    uint4 c = __ldg(&claims[i]);
    const uint32_t dst = out_off ? __ldg(&out_off[i]) : i;
    const uint32_t nb = n_node + 1;
    if (c.y >= n_node) {
        const uint32_t kind = c.x & 0xFFu;
        const uint32_t op = kind == DRA_KIND_GPU ? DRA_PROFILE_GPU
                          : kind == DRA_KIND_SHARED ? DRA_PROFILE_SHARED : ((c.x >> 8) & 0xFFu);
        if (dst < n_out) out[dst] = make_uint2(DRA_GPU_NONE, meta(0, 0, op, DRA_ST_INVALID));
        else err.set(ERR_OUT_RANGE);
        return;
    }
    const uint32_t t = i / T;
    const uint32_t dest = __ldg(&claim_off[c.y]) + __ldg(&hist[(size_t)t * nb + c.y]) + rank[i];
    c.y = dst;
    sorted[dest] = c;
}

// DRA_F_NODE_SORTED input: verify order, build claim_off by boundary detection, copy with slot embedded.
__global__ void __launch_bounds__(256)
k_sorted_prep(const uint4* __restrict__ claims, uint32_t n_claim, uint32_t n_node,
              const uint32_t* __restrict__ out_off, uint32_t* __restrict__ claim_off,
              uint4* __restrict__ sorted, uint2* __restrict__ out, uint32_t n_out, Err err) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0 && n_claim == 0) { for (uint32_t n = 0; n <= n_node + 1; ++n) claim_off[n] = 0; }
    if (i >= n_claim) return;
    uint4 c = __ldg(&claims[i]);
    const uint32_t key = c.y < n_node ? c.y : n_node;
    uint32_t prev = 0; bool first = i == 0;
    if (!first) { prev = __ldg(&claims[i - 1]).y; prev = prev < n_node ? prev : n_node; }
    if (!first && prev > key) err.set(ERR_NOT_SORTED);
    // claim_off[n] = i for every bucket n in (prev, key]  (first claim: [0, key])
    if (first) { for (uint32_t n = 0; n <= key; ++n) claim_off[n] = 0; }
    else if (prev < key) { for (uint32_t n = prev + 1; n <= key; ++n) claim_off[n] = i; }
    if (i == n_claim - 1) { for (uint32_t n = key + 1; n <= n_node + 1; ++n) claim_off[n] = n_claim; }
    const uint32_t dst = out_off ? __ldg(&out_off[i]) : i;
    if (key == n_node) {
        const uint32_t kind = c.x & 0xFFu;
        const uint32_t op = kind == DRA_KIND_GPU ? DRA_PROFILE_GPU
                          : kind == DRA_KIND_SHARED ? DRA_PROFILE_SHARED : ((c.x >> 8) & 0xFFu);
        if (dst < n_out) out[dst] = make_uint2(DRA_GPU_NONE, meta(0, 0, op, DRA_ST_INVALID));
        else err.set(ERR_OUT_RANGE);
    }
    c.y = dst;
    sorted[i] = c;
}

// ====================================================================================================
// pack: first-fit per node
// ====================================================================================================

struct PackArgs {
    const uint4* sorted;          // node-sorted claims, .y = first out slot
    const uint32_t* claim_off;    // [n_node+2]
    const uint4* inv_src;         // inventory read from here ...
    uint4* inv_dst;               // ... and written back here (may alias inv_src)
    const uint32_t* node_off;     // [n_node+1]
    const uint32_t* tbl;          // [16*16] ProfEnt as u32
    uint2* out;
    uint32_t n_out, n_node, have_off;
    Err err;
};

struct __align__(16) WarpSmem {
    uint4 inv[DRA_MAX_GPUS_PER_NODE];          // 512 B  node's GpuRecs
    uint4 ring[RING][SEG];                     // 2 KiB  claim segments
    uint64_t bar[RING + 1];                    // ring slot barriers + inventory barrier
    uint64_t pad;
};

// claim m of the node's span from the shared-memory ring (all lanes read one address: broadcast)
struct RingGet {
    const WarpSmem* ws; uint32_t segbase;
    __device__ __forceinline__ uint4 operator()(uint32_t m) const {
        return ws->ring[(segbase + (m >> 5)) & (RING - 1)][m & 31];
    }
};

template <int WPC>
__global__ void __launch_bounds__(WPC * 32)
k_pack(const PackArgs a) {
    __shared__ uint32_t tbl_s[DRA_MAX_MODELS * DRA_MAX_PROFILES];
    __shared__ WarpSmem wsm[WPC];
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    WarpSmem& ws = wsm[wid];

    for (uint32_t i = threadIdx.x; i < DRA_MAX_MODELS * DRA_MAX_PROFILES; i += WPC * 32) tbl_s[i] = __ldg(&a.tbl[i]);
    if (lane == 0) {
        #pragma unroll
        for (int q = 0; q <= (int)RING; ++q) mbar_init(&ws.bar[q], 1);
        mbar_fence_init();
    }
    __syncthreads();

    if (a.err.dev[ERR_NOT_SORTED]) return;     // input contract violated upstream: leave the inventory alone

    uint32_t segbase = 0;      // running segment counter: ring slot = (segbase+q) % RING
    uint32_t inv_phase = 0;
    const uint32_t nwarps = gridDim.x * WPC;
    for (uint32_t node = blockIdx.x * WPC + wid; node < a.n_node; node += nwarps) {
        const uint32_t g0 = __ldg(&a.node_off[node]);
        const uint32_t ng = __ldg(&a.node_off[node + 1]) - g0;
        const uint32_t c0 = __ldg(&a.claim_off[node]);
        const uint32_t cnt = __ldg(&a.claim_off[node + 1]) - c0;
        if (cnt == 0) {
            if (a.inv_src != a.inv_dst && lane < ng) a.inv_dst[g0 + lane] = __ldg(&a.inv_src[g0 + lane]);
            continue;
        }
        const uint32_t nseg = (cnt + SEG - 1) / SEG;
        uint32_t issued = 0, waited = 0;

        auto issue = [&](uint32_t q) {
            if (lane == 0) {
                const uint32_t n = min(SEG, cnt - q * SEG) * 16u;
                uint64_t* bar = &ws.bar[(segbase + q) & (RING - 1)];
                mbar_arrive_expect_tx(bar, n);
                tma_load_1d(&ws.ring[(segbase + q) & (RING - 1)][0], a.sorted + c0 + q * SEG, n, bar);
            }
        };
        // stage the node: GpuRecs + first claim segments, all in flight at once
        if (lane == 0 && ng) {
            mbar_arrive_expect_tx(&ws.bar[RING], ng * 16u);
            tma_load_1d(&ws.inv[0], a.inv_src + g0, ng * 16u, &ws.bar[RING]);
        }
        while (issued < nseg && issued < 3) issue(issued++);

        uint4 rec = make_uint4(0, 0, 0, 0);
        if (ng) { mbar_wait(&ws.bar[RING], inv_phase); inv_phase ^= 1; if (lane < ng) rec = ws.inv[lane]; }
        Lane L; L.load(rec, lane < ng);
        Dead D;
        OutSink sink{a.out, a.n_out, a.err, lane};
        RingGet get{&ws, segbase};

        uint32_t k = 0;
        while (k < cnt) {
            const uint32_t seg = k >> 5;
            const uint32_t need = min(nseg - 1, seg + 1);       // current + look-ahead for runs
            while (waited <= need) {
                mbar_wait(&ws.bar[(segbase + waited) & (RING - 1)], ((segbase + waited) / RING) & 1);
                ++waited;
            }
            // slots of seg and seg+1 are live; seg+2 may be in flight; seg-1's slot is free
            __syncwarp();
            while (issued < nseg && issued <= seg + 2) issue(issued++);
            k += node_step(L, D, lane, g0, tbl_s, get, k, cnt, sink, a.have_off != 0);
        }
        if (lane < ng) a.inv_dst[g0 + lane] = L.store(rec);
        segbase += nseg;
        __syncwarp();
    }
}

// ====================================================================================================
// UnsuitableNodes: warp per (pod, candidate) pair on a snapshot
// ====================================================================================================

struct UnsArgs {
    const uint4* claims; const uint32_t* pod_off; uint32_t n_pod;
    const uint32_t* cand_nodes; const uint32_t* cand_off; const uint32_t* pair_pod;   // pair -> pod
    uint32_t n_pair;
    const uint4* inv; const uint32_t* node_off; uint32_t n_node; const uint32_t* tbl;
    uint32_t* bits;     // n_pair bits, zeroed by the caller, set with atomicOr
};

struct GlobalGet {
    const uint4* base;
    __device__ __forceinline__ uint4 operator()(uint32_t m) const { return __ldg(&base[m]); }
};

template <int WPC>
__global__ void __launch_bounds__(WPC * 32)
k_unsuitable(const UnsArgs a) {
    __shared__ uint32_t tbl_s[DRA_MAX_MODELS * DRA_MAX_PROFILES];
    for (uint32_t i = threadIdx.x; i < DRA_MAX_MODELS * DRA_MAX_PROFILES; i += WPC * 32) tbl_s[i] = __ldg(&a.tbl[i]);
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t nwarps = gridDim.x * WPC;
    for (uint32_t pair = blockIdx.x * WPC + wid; pair < a.n_pair; pair += nwarps) {
        const uint32_t pod = __ldg(&a.pair_pod[pair]);
        const uint32_t node = __ldg(&a.cand_nodes[pair]);
        if (node >= a.n_node) continue;                               // unknown node: unsuitable
        const uint32_t c0 = __ldg(&a.pod_off[pod]);
        const uint32_t cnt = __ldg(&a.pod_off[pod + 1]) - c0;
        const uint32_t g0 = __ldg(&a.node_off[node]);
        const uint32_t ng = __ldg(&a.node_off[node + 1]) - g0;
        uint4 rec = make_uint4(0, 0, 0, 0);
        if (lane < ng) rec = __ldg(&a.inv[g0 + lane]);
        Lane L; L.load(rec, lane < ng);
        Dead D; FlagSink sink; GlobalGet get{a.claims + c0};
        uint32_t k = 0;
        while (k < cnt && !sink.stop()) k += node_step(L, D, lane, g0, tbl_s, get, k, cnt, sink, true);
        if (!sink.failed && lane == 0) atomicOr(&a.bits[pair >> 5], 1u << (pair & 31));
    }
}

// ====================================================================================================
// Deallocate: thread per claim, atomics on the GpuRec words (updates commute, spec §9)
// ====================================================================================================
__global__ void __launch_bounds__(256)
k_dealloc(const uint4* __restrict__ claims, uint32_t n_claim, const uint32_t* __restrict__ out_off,
          const uint2* __restrict__ out, uint32_t n_out, uint32_t* __restrict__ inv, uint32_t n_gpu,
          Err err) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_claim) return;
    const uint4 c = __ldg(&claims[i]);
    const uint32_t kind = c.x & 0xFFu, count = c.x >> 16;
    const bool have_off = out_off != nullptr;
    const uint32_t base = have_off ? __ldg(&out_off[i]) : i;
    const uint32_t slots = (kind == DRA_KIND_GPU && count >= 1 && count <= DRA_MAX_COUNT &&
                            (have_off || count == 1)) ? count : 1u;
    if (base > n_out || slots > n_out - base) { err.set(ERR_OUT_RANGE); return; }
    for (uint32_t s = 0; s < slots; ++s) {
        const uint2 o = __ldg(&out[base + s]);
        if ((o.y >> 24) != DRA_ST_OK || o.x >= n_gpu) continue;
        uint32_t* w = inv + (size_t)o.x * 4;
        if (kind == DRA_KIND_GPU) atomicAnd(&w[0], ~((uint32_t)DRA_GPU_FULL_ALLOCATED << 16));
        else if (kind == DRA_KIND_MIG) {
            const uint32_t st = o.y & 0xFFu, sz = (o.y >> 8) & 0xFFu;
            atomicAnd(&w[0], ~(((((1u << sz) - 1u) << st)) & 0xFFFFu));
        } else if (kind == DRA_KIND_SHARED) {
            atomicAdd(&w[1], c.z);
            uint32_t old = w[3], assumed;
            do { assumed = old; old = atomicCAS(&w[3], assumed, (assumed & 0xFFFF0000u) | ((assumed - 1u) & 0xFFFFu)); }
            while (old != assumed);
        }
    }
}

}  // namespace dra
