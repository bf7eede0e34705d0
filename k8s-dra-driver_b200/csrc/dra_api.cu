// dra_api.cu — C ABI of libdra_alloc.so (include/dra_alloc.h): context, staging, kernel chain, NCCL.
//
// No CPU fallback lives here: every entry point that computes anything launches the sm_100a kernels of
// dra_device.cuh, and dra_ctx_create fails when there is no usable device.
#include "dra_device.cuh"

#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

using namespace dra;

namespace {

thread_local std::string g_create_err;

// registry of dra_host_alloc buffers: passing one of these skips the staging copy
std::mutex g_pin_mu;
std::map<uintptr_t, size_t> g_pinned;

bool is_pinned(const void* p, size_t bytes) {
    if (!p) return false;
    std::lock_guard<std::mutex> lk(g_pin_mu);
    auto it = g_pinned.upper_bound((uintptr_t)p);
    if (it == g_pinned.begin()) return false;
    --it;
    return (uintptr_t)p + bytes <= it->first + it->second;
}

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool load(std::string& err) {
        if (lib) return true;
        // libnccl.so.2 already mapped by the host runtime (torch) is reused by soname
        lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) { err = std::string("dlopen libnccl: ") + dlerror(); return false; }
        GetUniqueId = (decltype(GetUniqueId))dlsym(lib, "ncclGetUniqueId");
        CommInitRank = (decltype(CommInitRank))dlsym(lib, "ncclCommInitRank");
        AllGather = (decltype(AllGather))dlsym(lib, "ncclAllGather");
        CommDestroy = (decltype(CommDestroy))dlsym(lib, "ncclCommDestroy");
        GetErrorString = (decltype(GetErrorString))dlsym(lib, "ncclGetErrorString");
        if (!GetUniqueId || !CommInitRank || !AllGather || !CommDestroy) { err = "libnccl: missing symbols"; return false; }
        return true;
    }
};
NcclApi g_nccl;
std::mutex g_nccl_mu;

}  // namespace

struct dra_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    uint32_t cfg_flags = 0;

    // inventory
    uint32_t n_gpu = 0, n_node = 0, max_width = 0;     // max_width: most GPUs on one node
    uint4* d_inv_live = nullptr;
    uint4* d_inv_pristine = nullptr;
    uint32_t* d_node_off = nullptr;
    uint32_t* d_tbl = nullptr;
    dra_profile_tbl h_tbl[DRA_MAX_MODELS];
    bool tbl_dirty = true;
    uint4* d_attrs = nullptr;  uint32_t n_attr = 0;      // spec §10
    uint4* d_sels = nullptr;   uint32_t n_sel = 0;

    // batch buffers (device)
    size_t cap_claims = 0, cap_out = 0, cap_hist = 0, cap_nodes = 0, cap_pairs = 0, cap_pods = 0, cap_bits_only = 0;
    uint4* d_claims = nullptr;
    uint4* d_sorted = nullptr;
    uint32_t* d_out_off = nullptr;
    uint2* d_out = nullptr;
    uint16_t* d_rank = nullptr;
    uint32_t* d_hist = nullptr;
    uint32_t* d_claim_off = nullptr;
    uint32_t *d_pod_off = nullptr, *d_cand_off = nullptr, *d_cand_nodes = nullptr, *d_pair_pod = nullptr, *d_bits = nullptr;
    uint4* d_podrec = nullptr; size_t cap_podrec = 0;     // pod mode (spec §12)

    // error flags
    uint32_t* d_err = nullptr;            // device memory
    volatile uint32_t* h_err = nullptr;   // mapped pinned host memory
    uint32_t* h_err_dev = nullptr;        // device alias of h_err

    // pinned staging
    uint8_t* h_in = nullptr;  size_t h_in_cap = 0;
    uint8_t* h_out = nullptr; size_t h_out_cap = 0;

    uint64_t launches = 0;
    bool profiling = false;
    cudaEvent_t ev[8] = {};
    bool ev_ok = false;
    float timings[5] = {0, 0, 0, 0, 0};
    uint32_t ev_mask = 0;
    // dynamic shared-memory limits are raised ONCE per device to the opt-in maximum (raise_smem_limits): the attribute is
    // per function, not per context — a lazily "grown" per-context value would LOWER it under another context's feet
    int hist8_smem_set = 1 << 30, hist_smem_set = 1 << 30, small_smem_set = 1 << 30, fused_smem_set = 1 << 30, fused_smem_set_stage = 1 << 30, fused_smem_set_cl = 1 << 30;
    uint64_t fused_max_work = 3000000ull;   // n_node * n_claim up to which the single-launch kernel is used
    uint32_t fused_max_one_wave = 14000u, fused_max_multi = 7500u;   // claims up to which it beats the sort path: measured on one box
                                                                     // (profiles/path_crossover_r01f.txt); dra_calibrate re-measures them here
    // direct host I/O of the single-launch kernel (DirectIO in dra_device.cuh)
    DirectIO dio_pending{};                 // set by dra_allocate_batch for the next launch_allocate, then cleared
    uint32_t* d_gbar = nullptr;             // grid barrier words
    int n_sm = 0, coop_ok = 0;
    int dio_cap_smem = -1, dio_cap_cta = 0; // co-resident CTA capacity of k_fused at dio_cap_smem bytes of shared memory
    unsigned long long* d_scan_status = nullptr; uint32_t cap_scan_status = 0;   // k_bucket_scan_rows: one word per CTA
    int pack4_occ = -1;              // resident CTAs of k_pack<4> per SM (occupancy query, once)
    int tail_cap_smem = -1, tail_cap_stage = -1, tail_cap_cta = 0;   // the same for the gather tail
    const void* dio_seen[3] = {nullptr, nullptr, nullptr};   // host pointers already checked to be device-visible as-is

    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;

    // resident mode (k_serve): the single-launch kernel stays up and takes batches by doorbell
    bool serve_on = false;
    cudaStream_t serve_stream = nullptr;
    ServeCmd* h_cmd = nullptr; ServeCmd* h_cmd_dev = nullptr;          // mapped host memory (one cache line)
    volatile uint32_t* h_stat = nullptr; uint32_t* h_stat_dev = nullptr;
    uint32_t* d_go = nullptr;
    uint32_t serve_seq = 0, serve_cap = 0; size_t serve_smem = 0; uint64_t serve_epoch = 0;
    uint64_t serve_batches = 0;

    // CUDA-graph replay of the host-buffer Allocate call
    struct GraphKeyT { const void* c; const void* o; void* d; uint32_t n_claim, n_out, flags; uint64_t epoch;
                       bool operator==(const GraphKeyT& k) const { return c == k.c && o == k.o && d == k.d && n_claim == k.n_claim && n_out == k.n_out && flags == k.flags && epoch == k.epoch && c != nullptr; } };
    GraphKeyT graph_key{};
    cudaGraphExec_t graph_exec = nullptr;
    uint32_t graph_launches = 0;
    uint64_t state_epoch = 1;     // bumped by every call that reallocates or re-points device state

    // peer-memory all-gather (packets over NVLink, dra_device.cuh PktGather)
    bool peer_ready = false;
    uint32_t peer_form = 0;                     // 1: per-rank slices [world][n_per]   2: global slots [tab_len]
    uint32_t peer_n_per = 0, peer_tab_len = 0, peer_cap = 0, peer_epoch = 0;
    size_t peer_bytes = 0, off_table[2] = {0, 0}, off_stage[2] = {0, 0}, off_hdr[2] = {0, 0}, off_flags = 0;
    uint32_t rendezvous_seq = 0;
    uint8_t* peer_local = nullptr;              // this rank's buffer (cudaMalloc, exported by IPC handle)
    uint8_t* peer_base[PEER_MAX] = {};          // every rank's buffer as mapped here
    bool peer_ipc[PEER_MAX] = {};               // mapped with cudaIpcOpenMemHandle (else: same process)
    uint32_t* d_cursor = nullptr;               // [2] packet reservation cursors (local)
    long long peer_spin = 4000000000ll;         // cycles a receiver waits for a packet before ERR_PEER_TIMEOUT
    // sharded global batch
    bool shard_on = false; uint32_t shard_lo = 0, shard_hi = 0, shard_stray = 0, shard_epoch = 0;
    unsigned long long* d_sc_times = nullptr;   // instrumentation (DRA_TIMELINE)
    uint32_t shard_last_n = 0;                  // batch size of the last sharded call (the plan hint is per batch size)
    bool shard_map_on = false; uint32_t shard_bounds[PEER_MAX + 1] = {}; uint32_t shard_stray_rank = 0;
    uint32_t* d_rank_slots = nullptr;           // [2][PEER_MAX] OutRec slots per rank, counted by the compaction
    uint4* d_cclaims = nullptr; uint32_t* d_coff = nullptr; size_t cap_cclaims = 0;
    unsigned long long* d_sc_status = nullptr; size_t cap_sc_status = 0;
    uint32_t* d_sc_counts = nullptr;            // device: [0] claims kept, [1] slots
    volatile uint32_t* h_sc_counts = nullptr;   // mapped host copy of the last call's counts
    uint32_t* h_sc_counts_dev = nullptr;
    uint2* d_gtable = nullptr; size_t cap_gtable = 0;   // result table when no peers are set up (world 1)
    uint32_t* d_ticket = nullptr;
    unsigned long long* d_timeline = nullptr; size_t tl_cap = 0; uint32_t tl_n = 0;
    const dra_out_rec* gather_table = nullptr;  // where the last gather's complete table lives (device)
    uint32_t gather_n_per = 0, gather_len = 0;   // gather_len != 0: a global table of that many records

    std::string err;
};

using GraphKey = dra_ctx::GraphKeyT;

namespace {

int fail(dra_ctx* c, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    if (c) c->err = buf; else g_create_err = buf;
    return code;
}

#define CU(call)                                                                                       \
    do { cudaError_t e_ = (call);                                                                     \
         if (e_ != cudaSuccess) return fail(ctx, DRA_E_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); } while (0)

template <class T>
int grow(dra_ctx* ctx, T*& p, size_t& cap, size_t need, size_t slack = 16) {
    if (need <= cap && p) return DRA_OK;
    size_t ncap = std::max(need, cap + cap / 2) + slack;
    if (p) CU(cudaFree(p));
    p = nullptr;
    CU(cudaMalloc((void**)&p, ncap * sizeof(T)));
    cap = ncap;
    ctx->state_epoch++;
    return DRA_OK;
}

template <class T>
int grow_nc(dra_ctx* ctx, T*& p, size_t have_cap, size_t new_cap) {   // companion buffer, cap tracked elsewhere
    if (p && new_cap <= have_cap) return DRA_OK;
    if (p) CU(cudaFree(p));
    p = nullptr;
    CU(cudaMalloc((void**)&p, new_cap * sizeof(T)));
    ctx->state_epoch++;
    return DRA_OK;
}

int grow_pinned(dra_ctx* ctx, uint8_t*& p, size_t& cap, size_t need) {
    if (need <= cap && p) return DRA_OK;
    size_t ncap = std::max(need, cap + cap / 2) + 256;
    if (p) CU(cudaFreeHost(p));
    p = nullptr;
    CU(cudaHostAlloc((void**)&p, ncap, cudaHostAllocDefault));
    cap = ncap;
    return DRA_OK;
}

struct Tiling { uint32_t T, n_tiles; bool cta_wide; };
Tiling tiling(uint32_t n_claim, uint32_t n_node, int n_sm) {
    // CTA-wide tiles (k_bucket_hist8) whenever 8 x (n_node+1) u16 counters fit in shared memory: 2048 claims, or — once
    // such tiles would outnumber the SMs — the multiple of 2048 that fills the SMs once (<= 32768: ranks are u16)
    if ((size_t)8 * ((size_t)n_node + 2) * 2 <= 200 * 1024) {
        static const bool fixed = getenv("DRA_HIST_TILE_2048") != nullptr;           // experiment switch: the round-1 tiling
        uint32_t T = H8_TILE;
        static const char* force = getenv("DRA_HIST_TILE");                          // experiment switch: a given multiple of 2048
        if (force && n_claim > (uint32_t)n_sm * H8_TILE) T = std::min<uint32_t>(32768u, std::max<uint32_t>(H8_TILE, (uint32_t)atoi(force) / H8_TILE * H8_TILE));
        else if (!fixed && n_claim > (uint32_t)n_sm * H8_TILE)
            T = std::min<uint32_t>(32768u, ((n_claim + (uint32_t)n_sm - 1) / (uint32_t)n_sm + H8_TILE - 1) / H8_TILE * H8_TILE);
        uint32_t n_tiles = n_claim ? (n_claim + T - 1) / T : 1;
        return {T, n_tiles, true};
    }
    uint32_t T = (n_claim + 63) / 64;
    T = (T + 31) & ~31u;
    T = std::max(256u, std::min(T, 65504u));
    uint32_t n_tiles = n_claim ? (n_claim + T - 1) / T : 1;
    return {T, n_tiles, false};
}

int ensure_batch(dra_ctx* ctx, uint32_t n_claim, uint32_t n_out, bool own_io) {
    size_t need_c = std::max<size_t>(n_claim, 1), need_o = std::max<size_t>(n_out, 1);
    if (need_c > ctx->cap_claims || !ctx->d_sorted) {
        size_t ncap = std::max(need_c, ctx->cap_claims + ctx->cap_claims / 2) + 64;
        int rc;
        if ((rc = grow_nc(ctx, ctx->d_claims, 0, ncap))) return rc;
        if ((rc = grow_nc(ctx, ctx->d_sorted, 0, ncap))) return rc;
        if ((rc = grow_nc(ctx, ctx->d_out_off, 0, ncap))) return rc;
        if ((rc = grow_nc(ctx, ctx->d_rank, 0, ncap))) return rc;
        ctx->cap_claims = ncap;
    }
    if (own_io) { int rc = grow(ctx, ctx->d_out, ctx->cap_out, need_o, 64); if (rc) return rc; }
    Tiling t = tiling(n_claim, ctx->n_node, ctx->n_sm);
    size_t need_h = (size_t)t.n_tiles * (ctx->n_node + 1);
    int rc = grow(ctx, ctx->d_hist, ctx->cap_hist, need_h, 64);
    if (rc) return rc;
    // k_bucket_scan_rows: one status word per CTA of 32 nodes, zeroed once (here, never inside a stream capture)
    const uint32_t n_cta = (ctx->n_node + 1 + 31) / 32;
    if (ctx->cap_scan_status < n_cta) {
        CU(cudaStreamSynchronize(ctx->stream));
        if (ctx->d_scan_status) CU(cudaFree(ctx->d_scan_status));
        ctx->d_scan_status = nullptr; ctx->cap_scan_status = 0;
        const uint32_t ncap = n_cta + n_cta / 2 + 64;
        CU(cudaMalloc((void**)&ctx->d_scan_status, (size_t)ncap * 8));
        CU(cudaMemset(ctx->d_scan_status, 0, (size_t)ncap * 8));
        ctx->cap_scan_status = ncap;
        ctx->state_epoch++;
    }
    return DRA_OK;
}

SelCtx sel_of(dra_ctx* ctx) { return SelCtx{ctx->n_attr == ctx->n_gpu ? ctx->d_attrs : nullptr, ctx->d_sels, ctx->n_sel}; }

Err err_of(dra_ctx* ctx) { return Err{ctx->d_err, (volatile uint32_t*)ctx->h_err_dev}; }

int upload_table(dra_ctx* ctx) {
    if (!ctx->tbl_dirty) return DRA_OK;
    CU(cudaMemcpyAsync(ctx->d_tbl, ctx->h_tbl, sizeof ctx->h_tbl, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));   // h_tbl is pageable; keep it simple (rare call)
    ctx->tbl_dirty = false;
    return DRA_OK;
}

// Kernel launch with (optionally) the programmatic-dependent-launch attribute: the kernel's CTAs may be scheduled
// while its predecessor in the stream drains; it blocks in griddepcontrol.wait until that grid has completed.
template <typename... KA, typename... A>
static cudaError_t launch_k(void (*k)(KA...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl, A&&... args) {
    cudaLaunchConfig_t lc; memset(&lc, 0, sizeof lc);
    lc.gridDim = grid; lc.blockDim = block; lc.dynamicSmemBytes = smem; lc.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = at; lc.numAttrs = pdl ? 1u : 0u;
    return cudaLaunchKernelEx(&lc, k, static_cast<KA>(args)...);
}

struct Prof {          // event i brackets stage i: [hist, scan, scatter, pack/fused, all-gather]
    dra_ctx* c; int i = 0;
    explicit Prof(dra_ctx* ctx) : c(ctx) { c->ev_mask = 0; rec(0); }
    void rec(int k) { if (c->profiling) { cudaEventRecord(c->ev[k], c->stream); c->ev_mask |= 1u << k; } i = k; }
    void mark() { rec(i + 1); }
    void skip_to(int k) { rec(k); }      // stages before k did not run in this call
};

// The kernel chain of one Allocate batch on device-resident inputs.  Enqueues only.
constexpr int FUSED_NW = 8;
// Which kernel chain an Allocate batch takes.  Small batches: ONE launch — every node's CTA filters the claim
// stream for itself (n_node * n_claim key tests spread over n_node SMs, data from L2) and packs; no sort, no copy.
struct FusedPlan { bool fused, stage; size_t smem; };
FusedPlan fused_plan(const dra_ctx* ctx, uint32_t n_claim, uint32_t flags, uint32_t n_node) {
    static const bool no_stage = getenv("DRA_NO_STAGE") != nullptr;          // experiment switch
    FusedPlan p;
    p.stage = !no_stage && n_claim <= FUSED_NW * FU_PIECE * FU_MAXPIECE && fused_smem_bytes(n_claim, FUSED_NW, true) <= 225 * 1024;
    p.smem = fused_smem_bytes(n_claim, FUSED_NW, p.stage);
    // Measured crossover against the sort path (profiles/path_crossover_r01f.txt): the single launch costs about
    // 4 us + 1.2 us per 1000 claims while all CTAs fit one wave (10 us + 1.4 us beyond), the sort path about
    // 20 us + 0.2 us per 1000 claims.
    const uint32_t fused_max_claims = (int)(n_node + 1) <= ctx->n_sm ? ctx->fused_max_one_wave : ctx->fused_max_multi;   // one wave of CTAs or not
    p.fused = !(flags & DRA_F_NODE_SORTED) && !(ctx->cfg_flags & DRA_CFG_NO_FUSED) && n_claim <= fused_max_claims &&
              (uint64_t)n_node * n_claim <= ctx->fused_max_work && p.smem <= 225 * 1024 && n_node <= 16384;
    return p;
}

// Stable counting sort of 16-byte records by their node key (.y) into ctx->d_sorted + ctx->d_claim_off — the
// front half of the sort path.  Records naming no node get an INVALID OutRec (d_out != nullptr) or are dropped
// (d_out == nullptr: pod records).  Enqueues only.
int launch_sort(dra_ctx* ctx, const uint4* d_claims, uint32_t n_claim, const uint32_t* d_out_off, uint2* d_out,
                uint32_t n_out, uint32_t flags, Prof& prof, const void* inv_src, bool pdl,
                uint32_t n_node, const uint32_t* d_node_off, const uint32_t* n_dev = nullptr) {
    Err err = err_of(ctx);
    // sort path: the kernels after the first are programmatic dependents of their predecessor (launch latency and
    // prologue overlap the predecessor's tail); not while per-kernel events are being recorded
    if (flags & DRA_F_NODE_SORTED) {
        if (n_dev) return fail(ctx, DRA_E_INVAL, "DRA_F_NODE_SORTED is not supported by the sharded call");
        uint32_t blocks = std::max(1u, (n_claim + 255) / 256);
        k_sorted_prep<<<blocks, 256, 0, ctx->stream>>>(d_claims, n_claim, n_node, d_out_off, ctx->d_claim_off,
                                                       ctx->d_sorted, d_out, n_out, err);
        ctx->launches += 1;
        prof.mark(); prof.skip_to(3);
    } else {
        const size_t nbp = ((size_t)n_node + 2) & ~(size_t)1;
        const size_t small_smem = nbp * 4 + 32 * nbp * 2 + (size_t)n_claim * 2 + 16;
        if (n_claim <= 8192 && small_smem <= 200 * 1024) {
            // one launch: the whole stable counting sort in a single CTA
            if (small_smem > 48 * 1024 && ctx->small_smem_set < (int)small_smem) {
                CU(cudaFuncSetAttribute(k_bucket_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)small_smem));
                ctx->small_smem_set = (int)small_smem;
            }
            Prefetch pf;
            pf.p[0] = inv_src;
            pf.bytes[0] = std::min<uint32_t>(ctx->n_gpu * 16u, 1u << 20) & ~15u;
            pf.p[1] = (const void*)((uintptr_t)d_node_off & ~(uintptr_t)15);      // (a shard's view starts mid-array: bulk prefetch wants 16-byte alignment)
            pf.bytes[1] = std::min<uint32_t>((n_node + 1) * 4u, 1u << 20) & ~15u;
            pf.p[2] = ctx->d_tbl; pf.bytes[2] = 1024;
            k_bucket_small<<<1, 1024, small_smem, ctx->stream>>>(d_claims, n_claim, n_node, d_out_off, ctx->d_claim_off,
                                                                 ctx->d_sorted, d_out, n_out, err, pf, n_dev);
            ctx->launches += 1;
            prof.mark(); prof.skip_to(3);
        } else {
            Tiling t = tiling(n_claim, n_node, ctx->n_sm);
            if (t.cta_wide) {
                const size_t smem = (size_t)8 * (((size_t)n_node + 2) & ~(size_t)1) * 2;
                if (smem > 48 * 1024 && ctx->hist8_smem_set < (int)smem) {
                    CU(cudaFuncSetAttribute(k_bucket_hist8<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                    CU(cudaFuncSetAttribute(k_bucket_hist8<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                    ctx->hist8_smem_set = (int)smem;
                }
                if (!ctx->d_ticket) { CU(cudaMalloc((void**)&ctx->d_ticket, 64)); CU(cudaMemsetAsync(ctx->d_ticket, 0, 64, ctx->stream)); }
                if (t.T == H8_TILE) k_bucket_hist8<false><<<t.n_tiles, 256, smem, ctx->stream>>>(d_claims, n_claim, n_node, ctx->d_hist, ctx->d_rank, n_dev, t.T);
                else k_bucket_hist8<true><<<t.n_tiles, 256, smem, ctx->stream>>>(d_claims, n_claim, n_node, ctx->d_hist, ctx->d_rank, n_dev, t.T);
                prof.mark();
                // the scan: a warp per node with every load in flight at once (latency), or — big matrices — coalesced rows
                static const bool no_rows = getenv("DRA_SCAN_BY_NODE") != nullptr;
                if (!no_rows && (size_t)t.n_tiles * (n_node + 1) > (size_t)512 * 1024) {
                    const uint32_t n_cta = (n_node + 1 + 31) / 32;
                    if (ctx->cap_scan_status < n_cta) return fail(ctx, DRA_E_STATE, "scan status words not allocated (ensure_batch)");
                    CU(launch_k(k_bucket_scan_rows, dim3(n_cta), dim3(256), 0, ctx->stream, pdl, ctx->d_hist, t.n_tiles, n_node, ctx->d_claim_off, ctx->d_ticket + 10, ctx->d_scan_status, err));
                }
                else
                    CU(launch_k(k_bucket_scan8, dim3((n_node + 1 + 7) / 8), dim3(256), 0, ctx->stream, pdl, ctx->d_hist, t.n_tiles, n_node, ctx->d_claim_off, ctx->d_ticket + 8));
                prof.mark();
            } else {
                if (n_dev) return fail(ctx, DRA_E_INVAL, "sharded call: node range too wide for the CTA-wide histogram");
                size_t smem = ((size_t)n_node + 1) * sizeof(uint16_t);
                if (smem > 200 * 1024) return fail(ctx, DRA_E_INVAL, "n_node=%u exceeds the bucketing limit", n_node);
                if (smem > 48 * 1024 && ctx->hist_smem_set < (int)smem) {
                    CU(cudaFuncSetAttribute(k_bucket_hist, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                    ctx->hist_smem_set = (int)smem;
                }
                k_bucket_hist<<<t.n_tiles, 32, smem, ctx->stream>>>(d_claims, n_claim, n_node, t.T, ctx->d_hist, ctx->d_rank);
                prof.mark();
                CU(launch_k(k_bucket_scan, dim3(1), dim3(1024), 0, ctx->stream, pdl, ctx->d_hist, t.n_tiles, n_node, ctx->d_claim_off));
                prof.mark();
            }
            uint32_t blocks = std::max(1u, (n_claim + 255) / 256);
            CU(launch_k(k_bucket_scatter, dim3(blocks), dim3(256), 0, ctx->stream, pdl, d_claims, n_claim, n_node, t.T, ctx->d_hist, ctx->d_rank,
                        ctx->d_claim_off, d_out_off, ctx->d_sorted, d_out, n_out, err, n_dev));
            prof.mark();
            ctx->launches += 3;
        }
    }

    return DRA_OK;
}

// The part of the inventory a launch works on (sharded call: this rank's node range; claims carry LOCAL node
// indices, GPU indices stay global because node_off keeps its global values) and where the claim count comes from.
struct AllocView {
    uint32_t node_lo = 0, n_node = 0;
    const uint32_t* n_dev = nullptr;      // device-side claim count (the claim list was compacted on the device)
    int have_off = -1;                    // -1: from d_out_off; else the CALLER's out_off convention (spec §1/§3)
    bool no_pdl_first = false;
    const ShardArgs* sa = nullptr;        // the compaction that produces the claim list: launched by launch_allocate — as its
    uint32_t sc_rows = 1; bool sc_flat = true;   // own kernel, or run inside k_fused when the shard takes the single-launch kernel
};

int launch_compaction(dra_ctx* ctx, const ShardArgs& sa, uint32_t sc_rows, bool sc_flat) {
    if (!sc_flat) k_shard_compact<<<sa.n_tiles, 256, 0, ctx->stream>>>(sa);
    else if (sc_rows == 1) k_shard_compact_flat<1><<<sa.n_tiles, 256, 0, ctx->stream>>>(sa);
    else k_shard_compact_flat<8><<<sa.n_tiles, 256, 0, ctx->stream>>>(sa);
    ctx->launches += 1;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(ctx, DRA_E_CUDA, "shard compaction launch: %s", cudaGetErrorString(e));
    return DRA_OK;
}

int launch_allocate(dra_ctx* ctx, const uint4* d_claims, uint32_t n_claim, const uint32_t* d_out_off,
                    uint2* d_out, uint32_t n_out, uint32_t flags, const PktGather* tail = nullptr, bool* tail_done = nullptr,
                    const AllocView* view = nullptr) {
    if (!ctx->d_inv_live) return fail(ctx, DRA_E_STATE, "dra_set_inventory has not been called");
    int rc = upload_table(ctx);
    if (rc) return rc;
    const uint32_t n_node = view ? view->n_node : ctx->n_node;
    const uint32_t* d_node_off = ctx->d_node_off + (view ? view->node_lo : 0u);
    const uint32_t* n_dev = view ? view->n_dev : nullptr;
    Err err = err_of(ctx);
    Prof prof(ctx);

    PackArgs a;
    memset(&a, 0, sizeof a);
    a.inv_src = (flags & DRA_F_FRESH_INVENTORY) ? ctx->d_inv_pristine : ctx->d_inv_live;
    a.inv_dst = ctx->d_inv_live;
    a.node_off = d_node_off;
    a.tbl = ctx->d_tbl;
    a.out = d_out; a.n_out = n_out; a.n_node = n_node;
    a.have_off = (view && view->have_off >= 0) ? (uint32_t)view->have_off : (d_out_off != nullptr);
    a.n_dev = n_dev;
    a.err = err;
    a.sel = sel_of(ctx);

    const FusedPlan plan = fused_plan(ctx, n_claim, flags, n_node);
    const bool fused = plan.fused, stage = plan.stage;
    const size_t fused_smem = plan.smem;
    const DirectIO dio = ctx->dio_pending;
    ctx->dio_pending = DirectIO{};
    if (dio.h_claims && !(fused && stage)) return fail(ctx, DRA_E_STATE, "direct host I/O was planned for a batch that does not take the staged single-launch kernel");
    if (fused) {
        // the gather tail has every CTA wait for its peers' packets: all CTAs of the grid must be resident at once.
        // Shared memory left over goes to the send queue (records leave for the peers while the pack still runs).
        bool tail_ok = tail != nullptr;
        const bool want_in = view && view->sa;             // a compaction comes with this launch
        size_t launch_smem = fused_smem;
        uint32_t q_cap = 0;
        if (tail_ok || want_in) {
            static const bool no_q = getenv("DRA_NO_SENDQ") != nullptr;
            const size_t room = fused_smem + 32 <= 225 * 1024 ? (225 * 1024 - fused_smem - 32) / 16 : 0;
            q_cap = no_q ? 0u : (uint32_t)std::min<size_t>(room, 2048);
            launch_smem = fused_smem + 32 + (size_t)q_cap * 16;
            if (ctx->tail_cap_smem != (int)launch_smem || ctx->tail_cap_stage != (int)stage) {
                int nb = 0;
                if (stage) CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_fused<FUSED_NW, true, 1>, FUSED_NW * 32, launch_smem));
                else CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_fused<FUSED_NW, false, 1>, FUSED_NW * 32, launch_smem));
                ctx->tail_cap_smem = (int)launch_smem; ctx->tail_cap_stage = (int)stage; ctx->tail_cap_cta = nb * ctx->n_sm;
            }
            const bool resident = (int)(n_node + 1) <= ctx->tail_cap_cta;
            tail_ok = tail_ok && resident;
            // the compaction inside the kernel: one 1024-claim tile per CTA, then a grid barrier (needs every CTA resident).
            // An experiment, off by default (DRA_SHARD_IN_KERNEL=1 turns it on): it saves the second launch but pays a
            // cooperative launch and a grid barrier, and loses the overlap of this kernel's prologue with the compaction —
            // measured at N = 2: 30.97 us per step against 29.02 us with the compaction as its own kernel + PDL
            // (profiles/r02_inkernel_compaction.txt)
            const bool want_inside = getenv("DRA_SHARD_IN_KERNEL") != nullptr;
            if (want_in && resident && stage && want_inside && ctx->coop_ok && !ctx->profiling && view->sc_flat &&
                (view->sa->n_claim + 256u * SC_IN_ROWS - 1) / (256u * SC_IN_ROWS) <= n_node + 1) {
                a.sh = *view->sa; a.sh_on = 1;
                a.sh.n_tiles = std::max(1u, (view->sa->n_claim + 256u * SC_IN_ROWS - 1) / (256u * SC_IN_ROWS));
                if (!ctx->d_gbar) { CU(cudaMalloc((void**)&ctx->d_gbar, 64)); CU(cudaMemsetAsync(ctx->d_gbar, 0, 64, ctx->stream)); }
                a.dio.gbar = ctx->d_gbar;
            }
        }
        if (want_in && !a.sh_on && (rc = launch_compaction(ctx, *view->sa, view->sc_rows, view->sc_flat))) return rc;
        if (tail_ok) { a.peer = *tail; a.q_cap = q_cap; if (tail_done) *tail_done = true; }
        else if (!a.sh_on) launch_smem = fused_smem;
        // clusters of 8 CTAs + TMA multicast: measured slower (profiles/cluster_multicast_r01e.txt: 28.1 vs 18.8 us per batch),
        // kept as an opt-in experiment (DRA_CLUSTER=1)
        constexpr int CLS = 8;
        static const bool want_cluster = getenv("DRA_CLUSTER") != nullptr;
        const bool cluster = stage && want_cluster && n_node + 1 >= (uint32_t)CLS && !tail_ok;
        a.claims = d_claims; a.out_off = d_out_off; a.n_claim = n_claim;
        if (getenv("DRA_TIMELINE")) {           // instrumentation only: per-CTA clock stamps of the last fused launch
            if (ctx->tl_cap < (size_t)(2 * n_node + 24) * 8) {
                if (ctx->d_timeline) CU(cudaFree(ctx->d_timeline));
                ctx->tl_cap = (size_t)(2 * n_node + 24) * 8 + 64;
                CU(cudaMalloc((void**)&ctx->d_timeline, ctx->tl_cap * 8));
                CU(cudaMemsetAsync(ctx->d_timeline, 0, ctx->tl_cap * 8, ctx->stream));     // (only here: a memset between the kernels of a step would sit on the measured path)
            }
            a.timeline = ctx->d_timeline; ctx->tl_n = (2 * n_node + 6) * 8;
        }
        prof.skip_to(3);
        // grid = one CTA per node + one CTA for the claims that name no node (+ padding to whole clusters)
        if (cluster) {
            cudaLaunchConfig_t lc; memset(&lc, 0, sizeof lc);
            lc.gridDim = dim3(((n_node + 1 + CLS - 1) / CLS) * CLS); lc.blockDim = dim3(FUSED_NW * 32);
            lc.dynamicSmemBytes = fused_smem; lc.stream = ctx->stream;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CLS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            lc.attrs = at; lc.numAttrs = 1;
            CU(cudaLaunchKernelEx(&lc, k_fused<FUSED_NW, true, CLS>, a));
        }
        else if (stage && dio.h_claims) {
            // direct host I/O: two grid barriers inside, so every CTA must be resident — cooperative launch
            a.dio = dio;
            cudaLaunchConfig_t lc; memset(&lc, 0, sizeof lc);
            lc.gridDim = dim3(n_node + 1); lc.blockDim = dim3(FUSED_NW * 32); lc.dynamicSmemBytes = fused_smem; lc.stream = ctx->stream;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeCooperative; at[0].val.cooperative = 1;
            lc.attrs = at; lc.numAttrs = 1;
            CU(cudaLaunchKernelEx(&lc, k_fused<FUSED_NW, true, 1>, a));
        }
        else {
            static const bool no_pdl_fused = getenv("DRA_NO_PDL") != nullptr;
            const bool dep = n_dev != nullptr && !ctx->profiling && !no_pdl_fused && !a.sh_on;      // programmatic dependent of the compaction
            if (a.sh_on) {                                  // compaction inside: a grid barrier, so a cooperative launch
                cudaLaunchConfig_t lc; memset(&lc, 0, sizeof lc);
                lc.gridDim = dim3(n_node + 1); lc.blockDim = dim3(FUSED_NW * 32); lc.dynamicSmemBytes = launch_smem; lc.stream = ctx->stream;
                cudaLaunchAttribute at[1];
                at[0].id = cudaLaunchAttributeCooperative; at[0].val.cooperative = 1;
                lc.attrs = at; lc.numAttrs = 1;
                if (stage) CU(cudaLaunchKernelEx(&lc, k_fused<FUSED_NW, true, 1>, a));
                else CU(cudaLaunchKernelEx(&lc, k_fused<FUSED_NW, false, 1>, a));
            }
            else if (stage) CU(launch_k(k_fused<FUSED_NW, true, 1>, dim3(n_node + 1), dim3(FUSED_NW * 32), launch_smem, ctx->stream, dep, a));
            else CU(launch_k(k_fused<FUSED_NW, false, 1>, dim3(n_node + 1), dim3(FUSED_NW * 32), launch_smem, ctx->stream, dep, a));
        }
        ctx->launches += 1;
        prof.mark();
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return fail(ctx, DRA_E_CUDA, "kernel launch: %s", cudaGetErrorString(e));
        return DRA_OK;
    }

    if (view && view->sa && (rc = launch_compaction(ctx, *view->sa, view->sc_rows, view->sc_flat))) return rc;
    static const bool no_pdl = getenv("DRA_NO_PDL") != nullptr;
    const bool pdl = !no_pdl && !ctx->profiling;
    if ((rc = launch_sort(ctx, d_claims, n_claim, d_out_off, d_out, n_out, flags, prof, a.inv_src, pdl, n_node, d_node_off, n_dev))) return rc;

    a.sorted = ctx->d_sorted;
    a.claim_off = ctx->d_claim_off;
    if (n_node) {
        if (n_node <= (uint32_t)ctx->n_sm * 16u) CU(launch_k(k_pack<1>, dim3(n_node), dim3(32), pack_smem_bytes(1), ctx->stream, pdl, a));
        else {
            // more nodes than warps can be resident: CTAs of 4 warps, exactly as many as fit at once, each warp walking nodes
            if (ctx->pack4_occ < 0) {
                int nb = 0;
                CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_pack<4>, 128, pack_smem_bytes(4)));
                ctx->pack4_occ = std::max(1, nb);
            }
            CU(launch_k(k_pack<4>, dim3(std::min((n_node + 3) / 4, (uint32_t)(ctx->n_sm * ctx->pack4_occ))), dim3(128), pack_smem_bytes(4), ctx->stream, pdl, a));
        }
        ctx->launches += 1;
    }
    prof.mark();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(ctx, DRA_E_CUDA, "kernel launch: %s", cudaGetErrorString(e));
    return DRA_OK;
}

int collect_timings(dra_ctx* ctx, int n_marks) {
    if (!ctx->profiling) return DRA_OK;
    for (int i = 0; i < 5; ++i) ctx->timings[i] = 0.f;
    for (int i = 0; i < n_marks && i < 5; ++i) {
        float ms = 0.f;
        if (!((ctx->ev_mask >> i) & 1u) || !((ctx->ev_mask >> (i + 1)) & 1u)) continue;
        if (cudaEventElapsedTime(&ms, ctx->ev[i], ctx->ev[i + 1]) == cudaSuccess) ctx->timings[i] = ms * 1000.f;
    }
    (void)cudaGetLastError();   // an event that was never recorded is not an error of the batch
    return DRA_OK;
}

int check_err(dra_ctx* ctx) {
    uint32_t oor = ctx->h_err[ERR_OUT_RANGE], ns = ctx->h_err[ERR_NOT_SORTED], pt = ctx->h_err[ERR_PEER_TIMEOUT], sp = ctx->h_err[ERR_SHARD_PLAN];
    const uint32_t dw = ctx->h_err[ERR_DEVICE_WAIT];
    if (!oor && !ns && !pt && !sp && !dw) return DRA_OK;
    for (uint32_t i = 0; i < ERR_WORDS; ++i) ctx->h_err[i] = 0;
    cudaMemsetAsync(ctx->d_err, 0, ERR_WORDS * sizeof(uint32_t), ctx->stream);
    cudaStreamSynchronize(ctx->stream);
    if (sp) return fail(ctx, DRA_E_STATE, "sharded call: %u claims fell into this shard, more than the launch was laid out for; nothing was changed, call again", ctx->h_sc_counts ? ctx->h_sc_counts[0] : 0u);
    if (dw) return fail(ctx, DRA_E_CUDA, "a device-side wait inside one kernel (grid barrier / look-back over CTAs) timed out: CTAs were not scheduled together; state is undefined, reload the inventory");
    if (pt) return fail(ctx, DRA_E_NCCL, "all-gather: a rank did not deliver its records in time (or aborted its batch)");
    if (ns) return fail(ctx, DRA_E_INVAL, "DRA_F_NODE_SORTED given but claims are not sorted by node; inventory unchanged");
    return fail(ctx, DRA_E_INVAL, "out_off/n_out: a claim's slots fall outside out[]; inventory state is undefined, reset it");
}

}  // namespace

namespace {
template <class K>
int raise_one(dra_ctx* ctx, K k, int optin) {
    cudaFuncAttributes fa;
    CU(cudaFuncGetAttributes(&fa, k));
    const int dyn = optin - (int)fa.sharedSizeBytes;          // the opt-in limit covers static + dynamic
    if (dyn > 48 * 1024) CU(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn));
    return DRA_OK;
}
int raise_smem_limits(dra_ctx* ctx, int optin) {
    int rc;
    if ((rc = raise_one(ctx, k_fused<FUSED_NW, true, 1>, optin))) return rc;
    if ((rc = raise_one(ctx, k_fused<FUSED_NW, false, 1>, optin))) return rc;
    if ((rc = raise_one(ctx, k_fused<FUSED_NW, true, 8>, optin))) return rc;
    if ((rc = raise_one(ctx, k_bucket_small, optin))) return rc;
    if ((rc = raise_one(ctx, k_bucket_hist8<false>, optin))) return rc;
    if ((rc = raise_one(ctx, k_bucket_hist8<true>, optin))) return rc;
    if ((rc = raise_one(ctx, k_bucket_hist, optin))) return rc;
    if ((rc = raise_one(ctx, k_serve<FUSED_NW>, optin))) return rc;
    // The compaction runs right before the single-launch kernel, which wants the SM's L1/shared split at "all shared".  An SM
    // cannot hold CTAs of kernels with different splits at once: with the default split the dependent kernel's CTAs could
    // not be placed early (programmatic dependent launch) and the SMs were re-configured in between (~6 us from the
    // compaction's last CTA to k_fused's first, profiles/tail_timeline_r02_n2.txt).  Same split for both.
    CU(cudaFuncSetAttribute(k_shard_compact_flat<1>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CU(cudaFuncSetAttribute(k_shard_compact_flat<8>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CU(cudaFuncSetAttribute(k_shard_compact, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    return DRA_OK;
}

// ---- resident mode -----------------------------------------------------------------------------------------------
// Stops the resident kernel (if any): EXIT command, then wait for the kernel to leave.  Every entry point that launches
// other work or touches device state calls this first — the resident kernel owns the SMs' shared memory while it runs.
int serve_stop(dra_ctx* ctx) {
    if (!ctx->serve_on) return DRA_OK;
    CU(cudaSetDevice(ctx->device));
    if (ctx->h_stat[1] != 2u) {
        ServeCmd* c = ctx->h_cmd;
        c->n_claim = 0; c->n_out = 0; c->flags = SERVE_EXIT; c->claims = 0; c->out_off = 0; c->out = 0;
        __atomic_store_n(&c->seq, ctx->serve_seq + 1u, __ATOMIC_RELEASE);
    }
    CU(cudaStreamSynchronize(ctx->serve_stream));
    ctx->serve_on = false;
    ctx->serve_seq += 1;                        // the EXIT consumed a sequence number (or none was taken: harmless gap)
    return DRA_OK;
}

// largest claim count whose staged layout fits the single-launch kernel's shared memory
uint32_t serve_capacity() {
    uint32_t lo = 256, hi = FUSED_NW * FU_PIECE * FU_MAXPIECE;
    while (lo < hi) { const uint32_t mid = (lo + hi + 1) / 2; if (fused_smem_bytes(mid, FUSED_NW, true) <= 225 * 1024) lo = mid; else hi = mid - 1; }
    return lo;
}

int serve_start(dra_ctx* ctx) {
    CU(cudaSetDevice(ctx->device));
    if (!ctx->serve_stream) CU(cudaStreamCreateWithFlags(&ctx->serve_stream, cudaStreamNonBlocking));
    if (!ctx->h_cmd) {
        void* p = nullptr;
        CU(cudaHostAlloc(&p, 256, cudaHostAllocMapped));
        memset(p, 0, 256);
        ctx->h_cmd = (ServeCmd*)p; ctx->h_stat = (volatile uint32_t*)((uint8_t*)p + 128);
        void* dp = nullptr;
        CU(cudaHostGetDevicePointer(&dp, p, 0));
        ctx->h_cmd_dev = (ServeCmd*)dp; ctx->h_stat_dev = (uint32_t*)((uint8_t*)dp + 128);
        CU(cudaMalloc((void**)&ctx->d_go, 256)); CU(cudaMemset(ctx->d_go, 0, 256));
    }
    if (!ctx->d_gbar) { CU(cudaMalloc((void**)&ctx->d_gbar, 64)); CU(cudaMemset(ctx->d_gbar, 0, 64)); }
    int rc = upload_table(ctx);
    if (rc) return rc;
    CU(cudaStreamSynchronize(ctx->stream));                  // earlier work of this context is done before the kernel takes over
    ctx->serve_cap = serve_capacity();
    ctx->serve_smem = fused_smem_bytes(ctx->serve_cap, FUSED_NW, true);
    if ((rc = ensure_batch(ctx, ctx->serve_cap, ctx->serve_cap * 2, true))) return rc;
    ServeArgs sa; memset(&sa, 0, sizeof sa);
    PackArgs& a = sa.base;
    a.inv_dst = ctx->d_inv_live; a.inv_src = ctx->d_inv_live; a.node_off = ctx->d_node_off; a.tbl = ctx->d_tbl;
    a.out = ctx->d_out; a.n_node = ctx->n_node; a.err = err_of(ctx); a.sel = sel_of(ctx);
    a.claims = ctx->d_claims;
    a.dio.d_claims = ctx->d_claims; a.dio.d_out_off = ctx->d_out_off; a.dio.gbar = ctx->d_gbar;
    sa.inv_pristine = ctx->d_inv_pristine;
    sa.h_cmd = ctx->h_cmd_dev; sa.h_stat = ctx->h_stat_dev; sa.d_go = ctx->d_go;
    sa.first_seq = ctx->serve_seq + 1; sa.cap_claims = ctx->serve_cap;
    static const long long idle_ms = getenv("DRA_SERVE_IDLE_MS") ? atoll(getenv("DRA_SERVE_IDLE_MS")) : 20;
    sa.idle_cycles = std::max(1ll, idle_ms) * 1900000ll;
    ctx->h_stat[1] = 0; ctx->h_stat[2] = 0;
    CU(cudaMemsetAsync(ctx->d_go, 0, 256, ctx->serve_stream));
    cudaLaunchConfig_t lc; memset(&lc, 0, sizeof lc);
    lc.gridDim = dim3(ctx->n_node + 1); lc.blockDim = dim3(FUSED_NW * 32); lc.dynamicSmemBytes = ctx->serve_smem; lc.stream = ctx->serve_stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeCooperative; at[0].val.cooperative = 1;
    lc.attrs = at; lc.numAttrs = 1;
    CU(cudaLaunchKernelEx(&lc, k_serve<FUSED_NW>, sa));
    ctx->launches += 1;
    ctx->serve_on = true; ctx->serve_epoch = ctx->state_epoch;
    return DRA_OK;
}

// one batch through the resident kernel; *took = false when the kernel had gone idle before it saw the doorbell
int serve_call(dra_ctx* ctx, const void* src_c, uint32_t n_claim, const void* src_o, void* dst_o, uint32_t n_out, uint32_t flags) {
    for (int attempt = 0; attempt < 3; ++attempt) {
        if (!ctx->serve_on || ctx->h_stat[1] == 2u || ctx->serve_epoch != ctx->state_epoch) {
            int rc = serve_stop(ctx);
            if (rc) return rc;
            if ((rc = serve_start(ctx))) return rc;
        }
        const uint32_t seq = ctx->serve_seq + 1u;
        ServeCmd* c = ctx->h_cmd;
        c->n_claim = n_claim; c->n_out = n_out; c->flags = flags & DRA_F_FRESH_INVENTORY;
        c->claims = (unsigned long long)(uintptr_t)src_c; c->out_off = (unsigned long long)(uintptr_t)src_o; c->out = (unsigned long long)(uintptr_t)dst_o;
        __atomic_store_n(&c->seq, seq, __ATOMIC_RELEASE);                  // the doorbell
        uint64_t spins = 0;
        bool done = false, gone = false;
        while (true) {
            if (__atomic_load_n((const uint32_t*)&ctx->h_stat[0], __ATOMIC_ACQUIRE) == seq) { done = true; break; }
            if ((++spins & 0xFFu) == 0) {
                if (ctx->h_stat[1] == 2u) { gone = __atomic_load_n((const uint32_t*)&ctx->h_stat[0], __ATOMIC_ACQUIRE) != seq; if (gone) break; done = true; break; }
                if (spins > (1ull << 31)) return fail(ctx, DRA_E_CUDA, "resident kernel does not answer");
            }
        }
        if (done) { ctx->serve_seq = seq; ctx->serve_batches++; return DRA_OK; }
        // the kernel left (idle time-out) before it saw this doorbell: start it again and ring once more
        CU(cudaStreamSynchronize(ctx->serve_stream));
        ctx->serve_on = false;
    }
    return fail(ctx, DRA_E_CUDA, "resident kernel keeps leaving before it takes the batch");
}

#define QUIESCE() do { int q_ = serve_stop(ctx); if (q_) return q_; } while (0)
}  // namespace

extern "C" {

int dra_abi_version(void) { return (int)DRA_ABI_VERSION; }

const char* dra_last_error(const dra_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

int dra_ctx_create(const dra_cfg* cfg, dra_ctx** out) {
    dra_ctx* ctx = nullptr;   // for CU(): errors go to the thread-local create error
    if (!cfg || !out) return fail(nullptr, DRA_E_INVAL, "null argument");
    if (cfg->abi_version != DRA_ABI_VERSION) return fail(nullptr, DRA_E_INVAL, "abi_version %u != %u", cfg->abi_version, DRA_ABI_VERSION);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, DRA_E_CUDA, "no CUDA device (%s); this library has no CPU fallback", cudaGetErrorString(e));
    if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, DRA_E_INVAL, "device %d out of range", cfg->device);
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major != 10) return fail(nullptr, DRA_E_CUDA, "device %d is sm_%d%d; this build is sm_100a only", cfg->device, prop.major, prop.minor);
    CU(cudaSetDevice(cfg->device));
    dra_ctx* c = new dra_ctx();
    c->device = cfg->device;
    c->cfg_flags = cfg->flags;
    c->n_sm = prop.multiProcessorCount;
    c->coop_ok = prop.cooperativeLaunch && prop.unifiedAddressing && prop.canUseHostPointerForRegisteredMem;
    ctx = c;
    auto bail = [&](int rc) { g_create_err = c->err; dra_ctx_destroy(c); return rc; };
    { int rc = raise_smem_limits(c, (int)prop.sharedMemPerBlockOptin); if (rc) return bail(rc); }
    if (cfg->stream) c->stream = (cudaStream_t)cfg->stream;
    else {
        if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) return bail(fail(c, DRA_E_CUDA, "cudaStreamCreate"));
        c->own_stream = true;
    }
    memset(c->h_tbl, 0, sizeof c->h_tbl);
    if (cudaMalloc((void**)&c->d_tbl, sizeof c->h_tbl) != cudaSuccess) return bail(fail(c, DRA_E_NOMEM, "cudaMalloc table"));
    if (cudaMalloc((void**)&c->d_err, ERR_WORDS * sizeof(uint32_t)) != cudaSuccess) return bail(fail(c, DRA_E_NOMEM, "cudaMalloc err"));
    cudaMemset(c->d_err, 0, ERR_WORDS * sizeof(uint32_t));
    void* he = nullptr;
    if (cudaHostAlloc(&he, ERR_WORDS * sizeof(uint32_t), cudaHostAllocMapped) != cudaSuccess) return bail(fail(c, DRA_E_NOMEM, "cudaHostAlloc err"));
    c->h_err = (volatile uint32_t*)he;
    for (uint32_t i = 0; i < ERR_WORDS; ++i) c->h_err[i] = 0;
    if (cudaHostGetDevicePointer((void**)&c->h_err_dev, he, 0) != cudaSuccess) return bail(fail(c, DRA_E_CUDA, "cudaHostGetDevicePointer"));
    if (cudaMalloc((void**)&c->d_ticket, 64) != cudaSuccess) return bail(fail(c, DRA_E_NOMEM, "cudaMalloc ticket"));
    cudaMemset(c->d_ticket, 0, 64);
    for (int i = 0; i < 8; ++i) if (cudaEventCreate(&c->ev[i]) != cudaSuccess) return bail(fail(c, DRA_E_CUDA, "cudaEventCreate"));
    c->ev_ok = true;
    if (cfg->max_claims) { int rc = ensure_batch(c, cfg->max_claims, cfg->max_claims, true); if (rc) return bail(rc); }
    *out = c;
    return DRA_OK;
}

void dra_ctx_destroy(dra_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    (void)serve_stop(c);
    if (c->serve_stream) cudaStreamDestroy(c->serve_stream);
    if (c->h_cmd) cudaFreeHost(c->h_cmd);
    if (c->d_go) cudaFree(c->d_go);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->graph_exec) cudaGraphExecDestroy(c->graph_exec);
    if (c->comm) { std::lock_guard<std::mutex> lk(g_nccl_mu); if (g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm); }
    for (int r = 0; r < (int)PEER_MAX; ++r) if (c->peer_base[r] && c->peer_ipc[r]) cudaIpcCloseMemHandle(c->peer_base[r]);
    if (c->peer_local) cudaFree(c->peer_local);
    if (c->d_ticket) cudaFree(c->d_ticket);
    if (c->d_scan_status) cudaFree(c->d_scan_status);
    if (c->d_gbar) cudaFree(c->d_gbar);
    void* dev[] = {c->d_inv_live, c->d_inv_pristine, c->d_node_off, c->d_tbl, c->d_claims, c->d_sorted, c->d_out_off,
                   c->d_out, c->d_rank, c->d_hist, c->d_claim_off, c->d_pod_off, c->d_cand_off, c->d_cand_nodes,
                   c->d_pair_pod, c->d_bits, c->d_err, c->d_attrs, c->d_sels, c->d_podrec, c->d_cursor, c->d_cclaims, c->d_coff,
                   c->d_sc_status, c->d_sc_counts, c->d_gtable, c->d_rank_slots, c->d_sc_times};
    for (void* p : dev) if (p) cudaFree(p);
    if (c->h_err) cudaFreeHost((void*)c->h_err);
    if (c->h_sc_counts) cudaFreeHost((void*)c->h_sc_counts);
    if (c->h_in) cudaFreeHost(c->h_in);
    if (c->h_out) cudaFreeHost(c->h_out);
    if (c->ev_ok) for (int i = 0; i < 8; ++i) cudaEventDestroy(c->ev[i]);
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

int dra_set_placement_table(dra_ctx* ctx, uint32_t model, const dra_profile_tbl* tbl) {
    if (!ctx || !tbl) return DRA_E_INVAL;
    QUIESCE();
    if (model >= DRA_MAX_MODELS) return fail(ctx, DRA_E_INVAL, "model %u >= %u", model, DRA_MAX_MODELS);
    for (uint32_t p = 0; p < DRA_MAX_PROFILES; ++p) {
        const dra_prof_ent& e = tbl->ent[p];
        if (!e.start_mask) continue;
        if (e.size < 1 || e.size > 16) return fail(ctx, DRA_E_INVAL, "profile %u: size %u", p, e.size);
        int hi = 15; while (!((e.start_mask >> hi) & 1)) --hi;
        if (hi + e.size > 16) return fail(ctx, DRA_E_INVAL, "profile %u: start %d + size %u > 16", p, hi, e.size);
    }
    ctx->h_tbl[model] = *tbl;
    ctx->tbl_dirty = true;
    ctx->state_epoch++;
    return DRA_OK;
}

int dra_set_inventory(dra_ctx* ctx, const dra_gpu_rec* gpus, uint32_t n_gpu, const uint32_t* node_off, uint32_t n_node) {
    if (!ctx || !node_off || (n_gpu && !gpus)) return DRA_E_INVAL;
    QUIESCE();
    CU(cudaSetDevice(ctx->device));
    if (node_off[0] != 0 || node_off[n_node] != n_gpu) return fail(ctx, DRA_E_INVAL, "node_off must span [0, n_gpu]");
    for (uint32_t n = 0; n < n_node; ++n) {
        if (node_off[n + 1] < node_off[n]) return fail(ctx, DRA_E_INVAL, "node_off not monotone at node %u", n);
        if (node_off[n + 1] - node_off[n] > DRA_MAX_GPUS_PER_NODE) return fail(ctx, DRA_E_INVAL, "node %u has more than %u GPUs", n, DRA_MAX_GPUS_PER_NODE);
        for (uint32_t g = node_off[n]; g < node_off[n + 1]; ++g) {
            if (gpus[g].model >= DRA_MAX_MODELS) return fail(ctx, DRA_E_INVAL, "gpu %u: model %u", g, gpus[g].model);
            if (gpus[g].node != n) return fail(ctx, DRA_E_INVAL, "gpu %u: node field %u, expected %u", g, gpus[g].node, n);
        }
    }
    CU(cudaStreamSynchronize(ctx->stream));
    if (n_gpu + 8 > ctx->cap_nodes || !ctx->d_inv_live) {   // cap_nodes doubles as "inventory capacity"
        if (ctx->d_inv_live) CU(cudaFree(ctx->d_inv_live));
        if (ctx->d_inv_pristine) CU(cudaFree(ctx->d_inv_pristine));
        ctx->d_inv_live = ctx->d_inv_pristine = nullptr;
        size_t cap = (size_t)n_gpu + 64;
        CU(cudaMalloc((void**)&ctx->d_inv_live, cap * 16));
        CU(cudaMalloc((void**)&ctx->d_inv_pristine, cap * 16));
        ctx->cap_nodes = cap;
    }
    if (ctx->d_node_off) CU(cudaFree(ctx->d_node_off));
    if (ctx->d_claim_off) CU(cudaFree(ctx->d_claim_off));
    ctx->d_node_off = ctx->d_claim_off = nullptr;
    CU(cudaMalloc((void**)&ctx->d_node_off, ((size_t)n_node + 8) * 4));
    CU(cudaMalloc((void**)&ctx->d_claim_off, ((size_t)n_node + 8) * 4));
    if (n_gpu) CU(cudaMemcpy(ctx->d_inv_pristine, gpus, (size_t)n_gpu * 16, cudaMemcpyHostToDevice));
    if (n_gpu) CU(cudaMemcpy(ctx->d_inv_live, gpus, (size_t)n_gpu * 16, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(ctx->d_node_off, node_off, ((size_t)n_node + 1) * 4, cudaMemcpyHostToDevice));
    ctx->n_gpu = n_gpu; ctx->n_node = n_node;
    ctx->state_epoch++;
    ctx->max_width = 0;
    for (uint32_t n = 0; n < n_node; ++n) ctx->max_width = std::max(ctx->max_width, node_off[n + 1] - node_off[n]);
    ctx->cap_hist = 0;  // histogram geometry depends on n_node
    if (ctx->d_hist) { CU(cudaFree(ctx->d_hist)); ctx->d_hist = nullptr; }
    return DRA_OK;
}

int dra_set_gpu_attrs(dra_ctx* ctx, const dra_gpu_attr* attrs, uint32_t n_gpu) {
    if (!ctx || (n_gpu && !attrs)) return DRA_E_INVAL;
    QUIESCE();
    static_assert(sizeof(dra_gpu_attr) == 16 && sizeof(dra_selector) == 64, "selector record layout");
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    if (ctx->d_attrs) { CU(cudaFree(ctx->d_attrs)); ctx->d_attrs = nullptr; }
    ctx->n_attr = 0;
    ctx->state_epoch++;                      // a captured graph holds the freed pointer: never replay it
    if (!n_gpu) return DRA_OK;
    CU(cudaMalloc((void**)&ctx->d_attrs, (size_t)n_gpu * 16 + 64));
    CU(cudaMemcpy(ctx->d_attrs, attrs, (size_t)n_gpu * 16, cudaMemcpyHostToDevice));
    ctx->n_attr = n_gpu;
    ctx->state_epoch++;
    return DRA_OK;
}

int dra_set_selectors(dra_ctx* ctx, const dra_selector* sels, uint32_t n_sel) {
    if (!ctx || (n_sel && !sels)) return DRA_E_INVAL;
    QUIESCE();
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    if (ctx->d_sels) { CU(cudaFree(ctx->d_sels)); ctx->d_sels = nullptr; }
    ctx->n_sel = 0;
    ctx->state_epoch++;
    if (!n_sel) return DRA_OK;
    CU(cudaMalloc((void**)&ctx->d_sels, (size_t)n_sel * 64 + 64));
    CU(cudaMemcpy(ctx->d_sels, sels, (size_t)n_sel * 64, cudaMemcpyHostToDevice));
    ctx->n_sel = n_sel;
    ctx->state_epoch++;
    return DRA_OK;
}

int dra_get_inventory(dra_ctx* ctx, dra_gpu_rec* gpus, uint32_t n_gpu) {
    if (!ctx || !gpus) return DRA_E_INVAL;
    QUIESCE();
    if (n_gpu != ctx->n_gpu) return fail(ctx, DRA_E_INVAL, "n_gpu %u != inventory size %u", n_gpu, ctx->n_gpu);
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    if (n_gpu) CU(cudaMemcpy(gpus, ctx->d_inv_live, (size_t)n_gpu * 16, cudaMemcpyDeviceToHost));
    return DRA_OK;
}

int dra_reset_inventory(dra_ctx* ctx) {
    if (!ctx) return DRA_E_INVAL;
    QUIESCE();
    if (!ctx->d_inv_live) return fail(ctx, DRA_E_STATE, "no inventory");
    CU(cudaSetDevice(ctx->device));
    if (ctx->n_gpu) CU(cudaMemcpyAsync(ctx->d_inv_live, ctx->d_inv_pristine, (size_t)ctx->n_gpu * 16, cudaMemcpyDeviceToDevice, ctx->stream));
    return DRA_OK;
}

int dra_ctx_sync(dra_ctx* ctx) {
    if (!ctx) return DRA_E_INVAL;
    QUIESCE();
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    collect_timings(ctx, 5);
    return check_err(ctx);
}

int dra_allocate_batch_device(dra_ctx* ctx, const dra_claim_rec* d_claims, uint32_t n_claim, const uint32_t* d_out_off,
                              dra_out_rec* d_out, uint32_t n_out, uint32_t flags) {
    if (!ctx || (n_claim && (!d_claims || !d_out))) return DRA_E_INVAL;
    QUIESCE();
    CU(cudaSetDevice(ctx->device));
    int rc = ensure_batch(ctx, n_claim, n_out, false);
    if (rc) return rc;
    return launch_allocate(ctx, (const uint4*)d_claims, n_claim, d_out_off, (uint2*)d_out, n_out, flags);
}

int dra_allocate_batch(dra_ctx* ctx, const dra_claim_rec* claims, uint32_t n_claim, const uint32_t* out_off,
                       dra_out_rec* out, uint32_t n_out, uint32_t flags) {
    if (!ctx || (n_claim && (!claims || !out))) return DRA_E_INVAL;
    if (!out_off && n_out < n_claim) return fail(ctx, DRA_E_INVAL, "n_out %u < n_claim %u without out_off", n_out, n_claim);
    CU(cudaSetDevice(ctx->device));
    int rc = ensure_batch(ctx, n_claim, n_out, true);
    if (rc) return rc;
    const size_t cb = (size_t)n_claim * 16, ob = out_off ? (size_t)n_claim * 4 : 0, rb = (size_t)n_out * 8;
    // host -> device (staged through pinned memory unless the caller's buffers are dra_host_alloc'ed)
    const void* src_c = claims; const void* src_o = out_off;
    if (cb && !is_pinned(claims, cb)) {
        if ((rc = grow_pinned(ctx, ctx->h_in, ctx->h_in_cap, cb + ob))) return rc;
        memcpy(ctx->h_in, claims, cb); src_c = ctx->h_in;
        if (ob) { memcpy(ctx->h_in + cb, out_off, ob); src_o = ctx->h_in + cb; }
    } else if (ob && !is_pinned(out_off, ob)) {
        if ((rc = grow_pinned(ctx, ctx->h_in, ctx->h_in_cap, ob))) return rc;
        memcpy(ctx->h_in, out_off, ob); src_o = ctx->h_in;
    }
    const bool direct = rb && is_pinned(out, rb);
    if (rb && !direct && (rc = grow_pinned(ctx, ctx->h_out, ctx->h_out_cap, rb))) return rc;
    void* dst_o = direct ? (void*)out : (void*)ctx->h_out;

    // Resident mode (DRA_CFG_RESIDENT): the kernel is already up — write the command, ring the doorbell, spin on the
    // completion word.  Same eligibility as direct host I/O below; anything else stops the resident kernel first.
    {
        bool ok = (ctx->cfg_flags & DRA_CFG_RESIDENT) && ctx->coop_ok && n_claim && rb && !ctx->profiling && !getenv("DRA_TIMELINE") &&
                  !(flags & ~DRA_F_FRESH_INVENTORY) && (int)(ctx->n_node + 1) <= ctx->n_sm && ctx->d_inv_live &&
                  ((uintptr_t)src_c & 15) == 0 && ((uintptr_t)dst_o & 15) == 0 && (!src_o || ((uintptr_t)src_o & 3) == 0);
        if (ok) {
            if (!ctx->serve_cap) ctx->serve_cap = serve_capacity();
            ok = n_claim <= ctx->serve_cap && n_out <= 2 * ctx->serve_cap;
        }
        if (ok) {
            const void* hp[3] = {src_c, src_o, dst_o};
            for (int k = 0; k < 3 && ok; ++k) {
                if (!hp[k] || ctx->dio_seen[k] == hp[k]) continue;
                void* dp = nullptr;
                if (cudaHostGetDevicePointer(&dp, const_cast<void*>(hp[k]), 0) != cudaSuccess || dp != hp[k]) { (void)cudaGetLastError(); ok = false; }
                else ctx->dio_seen[k] = hp[k];
            }
        }
        if (ok) {
            rc = serve_call(ctx, src_c, n_claim, src_o, dst_o, n_out, flags);
            if (rc) return rc;
            if ((rc = check_err(ctx))) return rc;
            if (rb && !direct) memcpy(out, ctx->h_out, rb);
            return DRA_OK;
        }
        QUIESCE();
    }
    // Direct host I/O: when the batch takes the single-launch kernel with the claim array staged in shared memory
    // and all its CTAs can be resident at once, the kernel itself reads the claims from the (pinned, device-mapped)
    // host buffer and writes the OutRecs back there — no copy-engine transfers, no graph, one cooperative launch.
    {
        static const bool no_direct = getenv("DRA_NO_DIRECT") != nullptr || getenv("DRA_CLUSTER") != nullptr;   // (the cluster experiment has no ingest)
        const FusedPlan plan = fused_plan(ctx, n_claim, flags, ctx->n_node);
        bool ok = !no_direct && !(ctx->cfg_flags & DRA_CFG_NO_DIRECT) && ctx->coop_ok && plan.fused && plan.stage && n_claim && rb &&
                  !ctx->profiling && ((uintptr_t)src_c & 15) == 0 && ((uintptr_t)dst_o & 15) == 0 &&
                  (!src_o || ((uintptr_t)src_o & 3) == 0);
        if (ok && ctx->dio_cap_smem != (int)plan.smem) {
            // co-resident capacity at this shared-memory size (the attribute must be raised before the query)
            if (plan.smem > 48 * 1024 && ctx->fused_smem_set_stage < (int)plan.smem) {
                CU(cudaFuncSetAttribute(k_fused<FUSED_NW, true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem));
                ctx->fused_smem_set_stage = (int)plan.smem;
            }
            int nb = 0;
            CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_fused<FUSED_NW, true, 1>, FUSED_NW * 32, plan.smem));
            ctx->dio_cap_smem = (int)plan.smem; ctx->dio_cap_cta = nb * ctx->n_sm;
        }
        ok = ok && (int)(ctx->n_node + 1) <= ctx->dio_cap_cta;
        if (ok) {
            // the host pointers must be usable by the device as they are (UVA identity mapping); checked once each
            const void* hp[3] = {src_c, src_o, dst_o};
            for (int k = 0; k < 3 && ok; ++k) {
                if (!hp[k] || ctx->dio_seen[k] == hp[k]) continue;
                void* dp = nullptr;
                if (cudaHostGetDevicePointer(&dp, const_cast<void*>(hp[k]), 0) != cudaSuccess || dp != hp[k]) { (void)cudaGetLastError(); ok = false; }
                else ctx->dio_seen[k] = hp[k];
            }
        }
        if (ok && !ctx->d_gbar) { CU(cudaMalloc((void**)&ctx->d_gbar, 64)); CU(cudaMemsetAsync(ctx->d_gbar, 0, 64, ctx->stream)); }
        if (ok) {
            DirectIO d;
            d.h_claims = (const uint4*)src_c; d.h_out_off = (const uint32_t*)src_o; d.h_out = (uint2*)dst_o;
            d.d_claims = ctx->d_claims; d.d_out_off = ctx->d_out_off; d.gbar = ctx->d_gbar;
            ctx->dio_pending = d;
            rc = launch_allocate(ctx, ctx->d_claims, n_claim, out_off ? ctx->d_out_off : nullptr, ctx->d_out, n_out, flags);
            ctx->dio_pending = DirectIO{};
            if (rc) return rc;
            CU(cudaStreamSynchronize(ctx->stream));
            if ((rc = check_err(ctx))) return rc;
            if (rb && !direct) memcpy(out, ctx->h_out, rb);
            return DRA_OK;
        }
    }

    // DRA_CFG_USE_GRAPH: H2D -> kernel chain -> D2H replayed as ONE cudaGraphLaunch when the call has the same
    // shape and buffers as the previous one (the first call of a shape runs eagerly: it also sets the kernels'
    // shared-memory attributes, which cannot happen inside a capture).
    GraphKey key{src_c, src_o, dst_o, n_claim, n_out, flags, ctx->state_epoch};
    const bool want_graph = (ctx->cfg_flags & DRA_CFG_USE_GRAPH) && !ctx->profiling && !getenv("DRA_TIMELINE");
    if (want_graph && ctx->graph_exec && key == ctx->graph_key) {
        CU(cudaGraphLaunch(ctx->graph_exec, ctx->stream));
        ctx->launches += ctx->graph_launches;
    } else {
        const bool capture = want_graph && key == ctx->graph_key;      // second call of this shape: capture it
        if (want_graph && !capture) { ctx->graph_key = key; if (ctx->graph_exec) { cudaGraphExecDestroy(ctx->graph_exec); ctx->graph_exec = nullptr; } }
        const uint64_t l0 = ctx->launches;
        if (capture) CU(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
        cudaError_t ce = cudaSuccess;
        if (cb) ce = cudaMemcpyAsync(ctx->d_claims, src_c, cb, cudaMemcpyHostToDevice, ctx->stream);
        if (ce == cudaSuccess && ob) ce = cudaMemcpyAsync(ctx->d_out_off, src_o, ob, cudaMemcpyHostToDevice, ctx->stream);
        if (ce == cudaSuccess) rc = launch_allocate(ctx, ctx->d_claims, n_claim, out_off ? ctx->d_out_off : nullptr, ctx->d_out, n_out, flags);
        if (ce == cudaSuccess && rc == DRA_OK && rb) ce = cudaMemcpyAsync(dst_o, ctx->d_out, rb, cudaMemcpyDeviceToHost, ctx->stream);
        if (capture) {
            cudaGraph_t g = nullptr;
            cudaError_t ee = cudaStreamEndCapture(ctx->stream, &g);
            if (ce != cudaSuccess || rc != DRA_OK || ee != cudaSuccess || !g) {
                if (g) cudaGraphDestroy(g);
                (void)cudaGetLastError();
                ctx->graph_key = GraphKey{};
                return rc != DRA_OK ? rc : fail(ctx, DRA_E_CUDA, "graph capture failed: %s", cudaGetErrorString(ce != cudaSuccess ? ce : ee));
            }
            ctx->graph_launches = (uint32_t)(ctx->launches - l0);
            cudaError_t ie = cudaGraphInstantiate(&ctx->graph_exec, g, 0);
            cudaGraphDestroy(g);
            if (ie != cudaSuccess) { ctx->graph_exec = nullptr; return fail(ctx, DRA_E_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(ie)); }
            CU(cudaGraphLaunch(ctx->graph_exec, ctx->stream));
        } else {
            if (ce != cudaSuccess) return fail(ctx, DRA_E_CUDA, "enqueue: %s", cudaGetErrorString(ce));
            if (rc) return rc;
        }
    }
    CU(cudaStreamSynchronize(ctx->stream));
    collect_timings(ctx, 4);
    if ((rc = check_err(ctx))) return rc;
    if (rb && !direct) memcpy(out, ctx->h_out, rb);
    return DRA_OK;
}

int dra_unsuitable_batch(dra_ctx* ctx, const dra_claim_rec* claims, uint32_t n_claim, const uint32_t* pod_off, uint32_t n_pod,
                         const uint32_t* cand_nodes, const uint32_t* cand_off, uint8_t* suitable_bits, uint32_t flags) {
    if (!ctx || !pod_off || (n_claim && !claims)) return DRA_E_INVAL;
    QUIESCE();
    if (flags & ~DRA_F_EXHAUSTIVE) return fail(ctx, DRA_E_INVAL, "dra_unsuitable_batch: unknown flags 0x%x", flags);
    if (!ctx->d_inv_live) return fail(ctx, DRA_E_STATE, "dra_set_inventory has not been called");
    if (pod_off[0] != 0 || pod_off[n_pod] != n_claim) return fail(ctx, DRA_E_INVAL, "pod_off must span [0, n_claim]");
    const bool dense = cand_nodes == nullptr && cand_off == nullptr;      // every pod against every node
    if (!dense && !cand_off) return DRA_E_INVAL;
    if (!dense && cand_off[0] != 0) return fail(ctx, DRA_E_INVAL, "cand_off[0] must be 0");
    const uint64_t n_pair64 = dense ? (uint64_t)n_pod * ctx->n_node : cand_off[n_pod];
    if (n_pair64 > 0xFFFFFFF0ull) return fail(ctx, DRA_E_INVAL, "too many (pod, node) pairs");
    const uint32_t n_pair = (uint32_t)n_pair64;
    if (n_pair && (!suitable_bits || (!dense && !cand_nodes))) return DRA_E_INVAL;
    CU(cudaSetDevice(ctx->device));
    int rc = ensure_batch(ctx, n_claim, 1, true);
    if (rc) return rc;
    if ((rc = upload_table(ctx))) return rc;
    if (n_pod + 8 > ctx->cap_pods) {
        size_t cap = (size_t)n_pod + n_pod / 2 + 64;
        if ((rc = grow_nc(ctx, ctx->d_pod_off, 0, cap))) return rc;
        if ((rc = grow_nc(ctx, ctx->d_cand_off, 0, cap))) return rc;
        ctx->cap_pods = cap;
    }
    const size_t words = ((size_t)n_pair + 31) / 32;
    if (!dense && (size_t)n_pair + 64 > ctx->cap_pairs) {
        size_t cap = (size_t)n_pair + n_pair / 2 + 128;
        if ((rc = grow_nc(ctx, ctx->d_cand_nodes, 0, cap))) return rc;
        if ((rc = grow_nc(ctx, ctx->d_pair_pod, 0, cap))) return rc;
        ctx->cap_pairs = cap;
    }
    if (words + 8 > ctx->cap_bits_only) {
        size_t cap = words + words / 2 + 64;
        if ((rc = grow_nc(ctx, ctx->d_bits, 0, cap))) return rc;
        ctx->cap_bits_only = cap;
    }
    // staging: claims | pod_off | (cand_off | cand_nodes | pair->pod expansion)
    const size_t stage = (size_t)n_claim * 16 + ((size_t)n_pod + 1) * 8 + (dense ? 0 : (size_t)n_pair * 8) + 64;
    if ((rc = grow_pinned(ctx, ctx->h_in, ctx->h_in_cap, stage))) return rc;
    uint8_t* p = ctx->h_in;
    if (n_claim) memcpy(p, claims, (size_t)n_claim * 16);
    uint8_t* h_claims = p; p += (size_t)n_claim * 16;
    memcpy(p, pod_off, ((size_t)n_pod + 1) * 4); uint8_t* h_pod = p; p += ((size_t)n_pod + 1) * 4;
    for (uint32_t q = 0; q < n_pod; ++q)
        if (pod_off[q + 1] < pod_off[q]) return fail(ctx, DRA_E_INVAL, "pod_off not monotone at pod %u", q);
    uint8_t *h_cn = nullptr; uint32_t* h_pp = nullptr;
    if (!dense) {
        p += ((size_t)n_pod + 1) * 4;                    // (cand_off is only needed on the host)
        if (n_pair) memcpy(p, cand_nodes, (size_t)n_pair * 4);
        h_cn = p; p += (size_t)n_pair * 4;
        h_pp = (uint32_t*)p;
        for (uint32_t q = 0; q < n_pod; ++q) {
            if (cand_off[q + 1] < cand_off[q] || cand_off[q + 1] > n_pair) return fail(ctx, DRA_E_INVAL, "cand_off not monotone at pod %u", q);
            for (uint32_t k = cand_off[q]; k < cand_off[q + 1]; ++k) h_pp[k] = q;
        }
    }
    if (n_claim) CU(cudaMemcpyAsync(ctx->d_claims, h_claims, (size_t)n_claim * 16, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->d_pod_off, h_pod, ((size_t)n_pod + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
    if (n_pair) {
        if (!dense) {
            CU(cudaMemcpyAsync(ctx->d_cand_nodes, h_cn, (size_t)n_pair * 4, cudaMemcpyHostToDevice, ctx->stream));
            CU(cudaMemcpyAsync(ctx->d_pair_pod, h_pp, (size_t)n_pair * 4, cudaMemcpyHostToDevice, ctx->stream));
        }
        CU(cudaMemsetAsync(ctx->d_bits, 0, words * 4, ctx->stream));
        UnsArgs a; memset(&a, 0, sizeof a);
        a.claims = ctx->d_claims; a.pod_off = ctx->d_pod_off; a.n_pod = n_pod;
        a.cand_nodes = ctx->d_cand_nodes; a.cand_off = ctx->d_cand_off; a.pair_pod = ctx->d_pair_pod; a.n_pair = n_pair;
        a.inv = ctx->d_inv_live; a.node_off = ctx->d_node_off; a.n_node = ctx->n_node; a.tbl = ctx->d_tbl; a.bits = ctx->d_bits;
        a.sel = sel_of(ctx);
        a.dense = dense ? 1u : 0u;
        a.exhaustive = (flags & DRA_F_EXHAUSTIVE) ? 1u : 0u;
        Prof prof(ctx);
        // lanes per pair from the widest node: 8 lanes = 4 pairs per warp
        const uint32_t W = ctx->max_width <= 8 ? 8u : (ctx->max_width <= 16 ? 16u : 32u);
        const uint32_t per_cta = 8 * (32 / W);
        const uint32_t grid = std::max(1u, std::min((n_pair + per_cta - 1) / per_cta, (uint32_t)ctx->n_sm * 8u));
        if (a.exhaustive) {
            a.work = ctx->d_ticket + 12;
            CU(cudaMemsetAsync(a.work, 0, 4, ctx->stream));
            if (W == 8) k_unsuitable<8, 8, true><<<grid, 256, 0, ctx->stream>>>(a);
            else if (W == 16) k_unsuitable<8, 16, true><<<grid, 256, 0, ctx->stream>>>(a);
            else k_unsuitable<8, 32, true><<<grid, 256, 0, ctx->stream>>>(a);
        } else {
            if (W == 8) k_unsuitable<8, 8, false><<<grid, 256, 0, ctx->stream>>>(a);
            else if (W == 16) k_unsuitable<8, 16, false><<<grid, 256, 0, ctx->stream>>>(a);
            else k_unsuitable<8, 32, false><<<grid, 256, 0, ctx->stream>>>(a);
        }
        prof.mark();
        ctx->launches += 1;
        if ((rc = grow_pinned(ctx, ctx->h_out, ctx->h_out_cap, words * 4))) return rc;
        CU(cudaMemcpyAsync(ctx->h_out, ctx->d_bits, words * 4, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CU(cudaStreamSynchronize(ctx->stream));
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(ctx, DRA_E_CUDA, "k_unsuitable: %s", cudaGetErrorString(e));
    collect_timings(ctx, 1);
    if (n_pair) memcpy(suitable_bits, ctx->h_out, ((size_t)n_pair + 7) / 8);
    return DRA_OK;
}

// Allocate with pod boundaries (spec §12): pod records -> stable counting sort of the PODS by node (the claim path's
// bucketing kernels, the record has a ClaimRec's layout) -> one warp per node walks its pods in input order.
int dra_allocate_pods_batch(dra_ctx* ctx, const dra_claim_rec* claims, uint32_t n_claim, const uint32_t* pod_off, uint32_t n_pod,
                            const uint32_t* out_off, dra_out_rec* out, uint32_t n_out, uint32_t flags) {
    if (!ctx || !pod_off || (n_claim && (!claims || !out))) return DRA_E_INVAL;
    QUIESCE();
    if (flags & ~(DRA_F_EXHAUSTIVE | DRA_F_FRESH_INVENTORY)) return fail(ctx, DRA_E_INVAL, "dra_allocate_pods_batch: unknown flags 0x%x", flags);
    if (!ctx->d_inv_live) return fail(ctx, DRA_E_STATE, "dra_set_inventory has not been called");
    if (pod_off[0] != 0 || pod_off[n_pod] != n_claim) return fail(ctx, DRA_E_INVAL, "pod_off must span [0, n_claim]");
    if (!out_off && n_out < n_claim) return fail(ctx, DRA_E_INVAL, "n_out %u < n_claim %u without out_off", n_out, n_claim);
    for (uint32_t q = 0; q < n_pod; ++q)
        if (pod_off[q + 1] < pod_off[q]) return fail(ctx, DRA_E_INVAL, "pod_off not monotone at pod %u", q);
    if (!n_pod) return DRA_OK;
    CU(cudaSetDevice(ctx->device));
    int rc = ensure_batch(ctx, std::max(n_claim, n_pod), n_out, true);
    if (rc) return rc;
    if ((rc = upload_table(ctx))) return rc;
    if (n_pod + 8 > ctx->cap_pods) {
        size_t cap = (size_t)n_pod + n_pod / 2 + 64;
        if ((rc = grow_nc(ctx, ctx->d_pod_off, 0, cap))) return rc;
        if ((rc = grow_nc(ctx, ctx->d_cand_off, 0, cap))) return rc;
        ctx->cap_pods = cap;
    }
    if (n_pod + 8 > ctx->cap_podrec) {
        size_t cap = (size_t)n_pod + n_pod / 2 + 64;
        if ((rc = grow_nc(ctx, ctx->d_podrec, 0, cap))) return rc;
        ctx->cap_podrec = cap;
    }
    const size_t cb = (size_t)n_claim * 16, pb = ((size_t)n_pod + 1) * 4, ob = out_off ? (size_t)n_claim * 4 : 0, rb = (size_t)n_out * 8;
    if ((rc = grow_pinned(ctx, ctx->h_in, ctx->h_in_cap, cb + pb + ob))) return rc;
    if ((rc = grow_pinned(ctx, ctx->h_out, ctx->h_out_cap, rb + 16))) return rc;
    if (cb) memcpy(ctx->h_in, claims, cb);
    memcpy(ctx->h_in + cb, pod_off, pb);
    if (ob) memcpy(ctx->h_in + cb + pb, out_off, ob);
    if (cb) CU(cudaMemcpyAsync(ctx->d_claims, ctx->h_in, cb, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->d_pod_off, ctx->h_in + cb, pb, cudaMemcpyHostToDevice, ctx->stream));
    if (ob) CU(cudaMemcpyAsync(ctx->d_out_off, ctx->h_in + cb + pb, ob, cudaMemcpyHostToDevice, ctx->stream));
    Err err = err_of(ctx);
    Prof prof(ctx);
    uint2* d_out = ctx->d_out;
    const uint32_t* d_oo = out_off ? ctx->d_out_off : nullptr;
    k_pod_records<<<(n_pod + 255) / 256, 256, 0, ctx->stream>>>(ctx->d_claims, ctx->d_pod_off, n_pod, ctx->n_node, d_oo, d_out, n_out, ctx->d_podrec, err);
    ctx->launches += 1;
    static const bool no_pdl = getenv("DRA_NO_PDL") != nullptr;
    const bool pdl = !no_pdl && !ctx->profiling;
    const uint4* inv_src = (flags & DRA_F_FRESH_INVENTORY) ? ctx->d_inv_pristine : ctx->d_inv_live;
    if ((rc = launch_sort(ctx, ctx->d_podrec, n_pod, nullptr, nullptr, 0, 0, prof, inv_src, false, ctx->n_node, ctx->d_node_off))) return rc;
    PodArgs a; memset(&a, 0, sizeof a);
    a.claims = ctx->d_claims; a.out_off = d_oo; a.out = d_out;
    a.sorted = ctx->d_sorted; a.claim_off = ctx->d_claim_off;
    a.inv_src = inv_src; a.inv_dst = ctx->d_inv_live; a.node_off = ctx->d_node_off; a.tbl = ctx->d_tbl;
    a.n_node = ctx->n_node; a.exhaustive = (flags & DRA_F_EXHAUSTIVE) ? 1u : 0u; a.sel = sel_of(ctx); a.err = err;
    if (ctx->n_node) {
        const uint32_t grid = std::max(1u, std::min((ctx->n_node + 3) / 4, (uint32_t)ctx->n_sm * 8u));
        CU(launch_k(k_pods<4>, dim3(grid), dim3(128), 0, ctx->stream, pdl, a));
        ctx->launches += 1;
    }
    if (rb) CU(cudaMemcpyAsync(ctx->h_out, d_out, rb, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(ctx, DRA_E_CUDA, "pod kernels: %s", cudaGetErrorString(e));
    if ((rc = check_err(ctx))) return rc;
    if (rb) memcpy(out, ctx->h_out, rb);
    return DRA_OK;
}

int dra_deallocate_batch(dra_ctx* ctx, const dra_claim_rec* claims, uint32_t n_claim, const uint32_t* out_off,
                         const dra_out_rec* out, uint32_t n_out) {
    if (!ctx || (n_claim && (!claims || !out))) return DRA_E_INVAL;
    QUIESCE();
    if (!ctx->d_inv_live) return fail(ctx, DRA_E_STATE, "dra_set_inventory has not been called");
    if (!n_claim) return DRA_OK;
    CU(cudaSetDevice(ctx->device));
    int rc = ensure_batch(ctx, n_claim, n_out, true);
    if (rc) return rc;
    const size_t cb = (size_t)n_claim * 16, ob = out_off ? (size_t)n_claim * 4 : 0, rb = (size_t)n_out * 8;
    if ((rc = grow_pinned(ctx, ctx->h_in, ctx->h_in_cap, cb + ob + rb))) return rc;
    memcpy(ctx->h_in, claims, cb);
    if (ob) memcpy(ctx->h_in + cb, out_off, ob);
    memcpy(ctx->h_in + cb + ob, out, rb);
    CU(cudaMemcpyAsync(ctx->d_claims, ctx->h_in, cb, cudaMemcpyHostToDevice, ctx->stream));
    if (ob) CU(cudaMemcpyAsync(ctx->d_out_off, ctx->h_in + cb, ob, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->d_out, ctx->h_in + cb + ob, rb, cudaMemcpyHostToDevice, ctx->stream));
    k_dealloc<<<(n_claim + 255) / 256, 256, 0, ctx->stream>>>(ctx->d_claims, n_claim, out_off ? ctx->d_out_off : nullptr,
                                                             ctx->d_out, n_out, (uint32_t*)ctx->d_inv_live, ctx->n_gpu, ctx->n_node, err_of(ctx));
    ctx->launches += 1;
    CU(cudaStreamSynchronize(ctx->stream));
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(ctx, DRA_E_CUDA, "k_dealloc: %s", cudaGetErrorString(e));
    return check_err(ctx);
}

// ---- multi-GPU --------------------------------------------------------------------------------------

int dra_comm_unique_id(void* id128) {
    if (!id128) return DRA_E_INVAL;
    std::lock_guard<std::mutex> lk(g_nccl_mu);
    std::string err;
    if (!g_nccl.load(err)) { g_create_err = err; return DRA_E_NCCL; }
    ncclUniqueId id;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclResult_t r = g_nccl.GetUniqueId(&id);
    if (r != ncclSuccess) { g_create_err = "ncclGetUniqueId failed"; return DRA_E_NCCL; }
    memcpy(id128, &id, 128);
    return DRA_OK;
}

int dra_comm_init(dra_ctx* ctx, const void* id128, int rank, int world) {
    if (!ctx || !id128 || world < 1 || rank < 0 || rank >= world) return DRA_E_INVAL;
    CU(cudaSetDevice(ctx->device));
    {
        std::lock_guard<std::mutex> lk(g_nccl_mu);
        std::string err;
        if (!g_nccl.load(err)) return fail(ctx, DRA_E_NCCL, "%s", err.c_str());
    }
    ncclUniqueId id; memcpy(&id, id128, 128);
    ncclResult_t r = g_nccl.CommInitRank(&ctx->comm, world, id, rank);
    if (r != ncclSuccess) return fail(ctx, DRA_E_NCCL, "ncclCommInitRank: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
    ctx->rank = rank; ctx->world = world;
    return DRA_OK;
}

namespace {

// (Re)allocates this rank's gather buffer:  [table x2][staging x2: world slices of cap packets][headers x2]
int peer_setup(dra_ctx* ctx, uint32_t form, uint32_t n_per, uint32_t tab_len, uint32_t cap) {
    if (ctx->world > (int)PEER_MAX) return fail(ctx, DRA_E_INVAL, "world %d > %u", ctx->world, PEER_MAX);
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->peer_ready = false;
    for (int r = 0; r < (int)PEER_MAX; ++r) {
        if (ctx->peer_base[r] && ctx->peer_ipc[r]) cudaIpcCloseMemHandle(ctx->peer_base[r]);
        ctx->peer_base[r] = nullptr; ctx->peer_ipc[r] = false;
    }
    if (ctx->peer_local) { CU(cudaFree(ctx->peer_local)); ctx->peer_local = nullptr; }
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = (off + bytes + 255) & ~(size_t)255; return o; };
    for (int p = 0; p < 2; ++p) ctx->off_table[p] = take((size_t)tab_len * 8 + 16);
    for (int p = 0; p < 2; ++p) ctx->off_stage[p] = take((size_t)ctx->world * cap * 16);
    for (int p = 0; p < 2; ++p) ctx->off_hdr[p] = take((size_t)ctx->world * 16);
    ctx->off_flags = take((size_t)PEER_MAX * 4); ctx->rendezvous_seq = 0;
    ctx->peer_bytes = off;
    CU(cudaMalloc((void**)&ctx->peer_local, ctx->peer_bytes));
    CU(cudaMemset(ctx->peer_local, 0, ctx->peer_bytes));
    if (!ctx->d_cursor) CU(cudaMalloc((void**)&ctx->d_cursor, 64));
    CU(cudaMemset(ctx->d_cursor, 0, 64));
    ctx->peer_form = form; ctx->peer_n_per = n_per; ctx->peer_tab_len = tab_len; ctx->peer_cap = cap; ctx->peer_epoch = 0;
    if (const char* e = getenv("DRA_PEER_TIMEOUT_MS")) ctx->peer_spin = std::max(1ll, atoll(e)) * 1900000ll;
    ctx->gather_table = nullptr;
    ctx->state_epoch++;
    return DRA_OK;
}

PktGather make_gather(dra_ctx* ctx, uint32_t epoch, uint32_t slot_base, uint32_t n_per, uint32_t count, const uint32_t* count_dev) {
    PktGather g; memset(&g, 0, sizeof g);
    const uint32_t par = epoch & 1u;
    for (int r = 0; r < ctx->world; ++r) {
        g.stage[r] = (uint4*)(ctx->peer_base[r] + ctx->off_stage[par]) + (size_t)ctx->rank * ctx->peer_cap;
        g.hdr[r] = (uint4*)(ctx->peer_base[r] + ctx->off_hdr[par]) + ctx->rank;
    }
    g.my_stage = (const uint4*)(ctx->peer_local + ctx->off_stage[par]);
    g.my_hdr = (const uint4*)(ctx->peer_local + ctx->off_hdr[par]);
    g.table = (uint2*)(ctx->peer_local + ctx->off_table[par]);
    g.cursor = ctx->d_cursor;
    g.world = ctx->world; g.rank = ctx->rank; g.cap = ctx->peer_cap; g.epoch = epoch; g.parity = par;
    g.slot_base = slot_base; g.n_per = n_per; g.count = count; g.count_dev = count_dev; g.spin_limit = ctx->peer_spin;
    return g;
}

int launch_gather(dra_ctx* ctx, const PktGather& g, const uint4* d_claims, uint32_t n_claim, const uint32_t* n_dev,
                  const uint32_t* d_out_off, uint32_t n_out, uint32_t have_off, uint32_t n_node) {
    const uint32_t blocks = std::max(1u, std::min((uint32_t)ctx->n_sm * 2u, (std::max(n_claim, g.world * g.cap / 4u) + 255u) / 256u));
    k_pkt_gather<<<blocks, 256, 0, ctx->stream>>>(g, d_claims, n_claim, n_dev, d_out_off, n_out, have_off, n_node, err_of(ctx));
    ctx->launches += 1;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(ctx, DRA_E_CUDA, "k_pkt_gather: %s", cudaGetErrorString(e));
    return DRA_OK;
}

}  // namespace

int dra_peer_export(dra_ctx* ctx, uint32_t n_per_rank, void* handle64) {
    if (!ctx) return DRA_E_INVAL;
    if (!n_per_rank) { ctx->peer_ready = false; return DRA_OK; }          // switch the peer path off (NCCL again)
    if (!handle64) return DRA_E_INVAL;
    int rc = peer_setup(ctx, 1, n_per_rank, (uint32_t)ctx->world * n_per_rank, n_per_rank);
    if (rc) return rc;
    cudaIpcMemHandle_t h;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    CU(cudaIpcGetMemHandle(&h, ctx->peer_local));
    memcpy(handle64, &h, 64);
    return DRA_OK;
}

int dra_shard_export(dra_ctx* ctx, uint32_t n_out_max, uint32_t cap_per_rank, void* handle64) {
    if (!ctx || !n_out_max) return DRA_E_INVAL;
    if (n_out_max > (1u << 26)) return fail(ctx, DRA_E_INVAL, "n_out_max %u exceeds 2^26", n_out_max);
    int rc = peer_setup(ctx, 2, 0, n_out_max, cap_per_rank ? std::min(cap_per_rank, n_out_max) : n_out_max);
    if (rc) return rc;
    if (handle64) {
        cudaIpcMemHandle_t h;
        CU(cudaIpcGetMemHandle(&h, ctx->peer_local));
        memcpy(handle64, &h, 64);
    }
    return DRA_OK;
}

int dra_peer_import(dra_ctx* ctx, const void* handles) {
    if (!ctx || !handles) return DRA_E_INVAL;
    if (!ctx->peer_local) return fail(ctx, DRA_E_STATE, "dra_peer_export / dra_shard_export has not been called");
    CU(cudaSetDevice(ctx->device));
    for (int r = 0; r < ctx->world; ++r) {
        if (r == ctx->rank) { ctx->peer_base[r] = ctx->peer_local; ctx->peer_ipc[r] = false; continue; }
        cudaIpcMemHandle_t h; memcpy(&h, (const uint8_t*)handles + (size_t)r * 64, 64);
        void* p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) { (void)cudaGetLastError(); return fail(ctx, DRA_E_CUDA, "cudaIpcOpenMemHandle(rank %d): %s", r, cudaGetErrorString(e)); }
        ctx->peer_base[r] = (uint8_t*)p; ctx->peer_ipc[r] = true;
    }
    ctx->peer_ready = true;
    return DRA_OK;
}

int dra_peer_import_local(dra_ctx* ctx, dra_ctx* const* ctxs) {
    if (!ctx || !ctxs) return DRA_E_INVAL;
    if (!ctx->peer_local) return fail(ctx, DRA_E_STATE, "dra_peer_export / dra_shard_export has not been called");
    CU(cudaSetDevice(ctx->device));
    for (int r = 0; r < ctx->world; ++r) {
        dra_ctx* o = ctxs[r];
        if (!o || !o->peer_local || o->world != ctx->world || o->rank != r || o->peer_form != ctx->peer_form ||
            o->peer_cap != ctx->peer_cap || o->peer_tab_len != ctx->peer_tab_len)
            return fail(ctx, DRA_E_INVAL, "dra_peer_import_local: context %d is not set up like this one", r);
        if (o->device != ctx->device) {
            int ok = 0; CU(cudaDeviceCanAccessPeer(&ok, ctx->device, o->device));
            if (!ok) return fail(ctx, DRA_E_CUDA, "no peer access from device %d to %d", ctx->device, o->device);
            cudaError_t e = cudaDeviceEnablePeerAccess(o->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(ctx, DRA_E_CUDA, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e));
            (void)cudaGetLastError();
        }
        ctx->peer_base[r] = o->peer_local; ctx->peer_ipc[r] = false;
    }
    ctx->peer_ready = true;
    return DRA_OK;
}

int dra_peer_rendezvous_device(dra_ctx* ctx) {
    if (!ctx) return DRA_E_INVAL;
    QUIESCE();
    if (ctx->world <= 1) return DRA_OK;
    if (!ctx->peer_ready) return fail(ctx, DRA_E_STATE, "dra_peer_import has not been called");
    CU(cudaSetDevice(ctx->device));
    RendezvousArgs a; memset(&a, 0, sizeof a);
    for (int r = 0; r < ctx->world; ++r) a.peer_flags[r] = (uint32_t*)(ctx->peer_base[r] + ctx->off_flags);
    a.my_flags = (const uint32_t*)(ctx->peer_local + ctx->off_flags);
    a.world = (uint32_t)ctx->world; a.rank = (uint32_t)ctx->rank; a.seq = ++ctx->rendezvous_seq; a.spin_limit = ctx->peer_spin; a.err = err_of(ctx);
    k_rendezvous<<<1, 32, 0, ctx->stream>>>(a);              // (not counted in dra_launch_count: it is no part of a batch)
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(ctx, DRA_E_CUDA, "k_rendezvous: %s", cudaGetErrorString(e));
    return DRA_OK;
}

int dra_comm_init_local(dra_ctx* ctx, int rank, int world) {
    if (!ctx || world < 1 || world > (int)PEER_MAX || rank < 0 || rank >= world) return DRA_E_INVAL;
    ctx->rank = rank; ctx->world = world;
    return DRA_OK;
}

int dra_allocate_batch_gather_device(dra_ctx* ctx, const dra_claim_rec* d_claims, uint32_t n_claim, const uint32_t* d_out_off,
                                     dra_out_rec* d_out_all, uint32_t n_out, uint32_t n_per_rank, uint32_t flags) {
    if (!ctx || (n_claim && !d_claims)) return DRA_E_INVAL;
    QUIESCE();
    if (!ctx->comm && !ctx->peer_ready) return fail(ctx, DRA_E_STATE, "dra_comm_init has not been called");
    if (!d_out_all && !ctx->peer_ready) return fail(ctx, DRA_E_INVAL, "d_out_all may be NULL only with the peer all-gather set up");
    if (n_out > n_per_rank) return fail(ctx, DRA_E_INVAL, "n_out %u > n_per_rank %u", n_out, n_per_rank);
    CU(cudaSetDevice(ctx->device));
    int rc = ensure_batch(ctx, n_claim, n_out, false);
    if (rc) return rc;

    if (ctx->peer_ready && ctx->peer_form == 1 && n_per_rank == ctx->peer_n_per) {
        // ---- packet all-gather over NVLink: rides in the tail of the fused kernel, or runs as its own kernel ----
        ctx->peer_epoch += 1;
        const PktGather g = make_gather(ctx, ctx->peer_epoch, (uint32_t)ctx->rank * n_per_rank, n_per_rank, n_out, nullptr);
        uint2* mine = g.table + g.slot_base;
        bool tail_done = false;
        rc = launch_allocate(ctx, (const uint4*)d_claims, n_claim, d_out_off, mine, n_out, flags, &g, &tail_done);
        if (rc) return rc;
        if (!tail_done && (rc = launch_gather(ctx, g, (const uint4*)d_claims, n_claim, nullptr, d_out_off, n_out, d_out_off != nullptr, ctx->n_node))) return rc;
        ctx->gather_table = (const dra_out_rec*)g.table; ctx->gather_n_per = n_per_rank; ctx->gather_len = 0;
        if (d_out_all) CU(cudaMemcpyAsync(d_out_all, g.table, (size_t)ctx->world * n_per_rank * 8, cudaMemcpyDeviceToDevice, ctx->stream));
        if (ctx->profiling && !tail_done) { cudaEventRecord(ctx->ev[5], ctx->stream); ctx->ev_mask |= 1u << 5; }
        return DRA_OK;
    }

    if (!ctx->comm) return fail(ctx, DRA_E_STATE, "dra_comm_init has not been called (NCCL path)");
    if (!d_out_all) return fail(ctx, DRA_E_INVAL, "d_out_all is required on the NCCL path");
    ctx->gather_table = d_out_all; ctx->gather_n_per = n_per_rank; ctx->gather_len = 0;
    dra_out_rec* mine = d_out_all + (size_t)ctx->rank * n_per_rank;
    if (n_per_rank > n_out) CU(cudaMemsetAsync(mine + n_out, 0, (size_t)(n_per_rank - n_out) * 8, ctx->stream));
    rc = launch_allocate(ctx, (const uint4*)d_claims, n_claim, d_out_off, (uint2*)mine, n_out, flags);
    if (rc) return rc;
    // the one collective of the path: every rank's OutRec slice to every rank (in place), same stream
    ncclResult_t r = g_nccl.AllGather(mine, d_out_all, (size_t)n_per_rank * 8, ncclUint8, ctx->comm, ctx->stream);
    if (r != ncclSuccess) return fail(ctx, DRA_E_NCCL, "ncclAllGather: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
    if (ctx->profiling) { cudaEventRecord(ctx->ev[5], ctx->stream); ctx->ev_mask |= 1u << 5; }
    return DRA_OK;
}

// ---- sharded global batch -------------------------------------------------------------------------------

int dra_set_shard(dra_ctx* ctx, uint32_t node_lo, uint32_t node_hi, int take_stray) {
    if (!ctx) return DRA_E_INVAL;
    if (node_lo > node_hi || node_hi > ctx->n_node) return fail(ctx, DRA_E_INVAL, "shard [%u, %u) outside the inventory's %u nodes", node_lo, node_hi, ctx->n_node);
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    if (ctx->h_sc_counts) ctx->h_sc_counts[3] = 0u;           // a new partition: the old call's count is no plan hint any more
    ctx->shard_on = true; ctx->shard_lo = node_lo; ctx->shard_hi = node_hi; ctx->shard_stray = take_stray ? 1u : 0u;
    ctx->shard_map_on = false;
    ctx->state_epoch++;
    return DRA_OK;
}

int dra_set_shard_map(dra_ctx* ctx, const uint32_t* bounds, int stray_rank) {
    if (!ctx || !bounds) return DRA_E_INVAL;
    if (stray_rank < 0 || stray_rank >= ctx->world) return fail(ctx, DRA_E_INVAL, "stray_rank %d outside the world of %d", stray_rank, ctx->world);
    if (bounds[0] != 0 || bounds[ctx->world] != ctx->n_node) return fail(ctx, DRA_E_INVAL, "shard map must span [0, %u]", ctx->n_node);
    for (int r = 0; r < ctx->world; ++r) if (bounds[r + 1] < bounds[r]) return fail(ctx, DRA_E_INVAL, "shard map not monotone at rank %d", r);
    int rc = dra_set_shard(ctx, bounds[ctx->rank], bounds[ctx->rank + 1], stray_rank == ctx->rank);
    if (rc) return rc;
    for (int r = 0; r <= ctx->world; ++r) ctx->shard_bounds[r] = bounds[r];
    ctx->shard_stray_rank = (uint32_t)stray_rank; ctx->shard_map_on = true;
    return DRA_OK;
}

int dra_allocate_batch_global_device(dra_ctx* ctx, const dra_claim_rec* d_claims, uint32_t n_claim, const uint32_t* d_out_off,
                                     uint32_t n_out, uint32_t flags) {
    if (!ctx || (n_claim && !d_claims)) return DRA_E_INVAL;
    QUIESCE();
    if (!ctx->shard_on) return fail(ctx, DRA_E_STATE, "dra_set_shard has not been called");
    if (flags & ~DRA_F_FRESH_INVENTORY) return fail(ctx, DRA_E_INVAL, "dra_allocate_batch_global_device: unsupported flags 0x%x", flags);
    if (n_claim > (1u << 26) || n_out > (1u << 26)) return fail(ctx, DRA_E_INVAL, "batch too large for the sharded call (2^26)");
    if (!d_out_off && n_out < n_claim) return fail(ctx, DRA_E_INVAL, "n_out %u < n_claim %u without out_off", n_out, n_claim);
    const bool gather = ctx->world > 1;
    if (gather && !(ctx->peer_ready && ctx->peer_form == 2)) return fail(ctx, DRA_E_STATE, "dra_shard_export + dra_peer_import have not been called");
    if (gather && n_out > ctx->peer_tab_len) return fail(ctx, DRA_E_INVAL, "n_out %u exceeds the exported table (%u)", n_out, ctx->peer_tab_len);
    CU(cudaSetDevice(ctx->device));
    const uint32_t n_local_node = ctx->shard_hi - ctx->shard_lo;
    // Plan.  The single-launch kernel lays its shared memory out for a CAPACITY of claims; the real number is only known
    // on the device (the compaction's count).  First call of a shape: capacity = the whole batch (always safe).  Later
    // calls: the last completed call's count plus slack — as much slack as still lets the claim array be STAGED in shared
    // memory (the fast form).  If more claims than the capacity fall into the shard the kernel touches nothing, tells its
    // peers, and the call fails with DRA_E_STATE ("call again": the next plan uses the new count).
    uint32_t cap = n_claim;
    FusedPlan plan = fused_plan(ctx, cap, flags, n_local_node);
    if (ctx->h_sc_counts && ctx->h_sc_counts[3] != 0u && ctx->shard_last_n == n_claim) {
        const uint32_t expect = ctx->h_sc_counts[0];
        const uint32_t hi = (uint32_t)std::min<uint64_t>(n_claim, (uint64_t)expect + expect / 4 + 256);
        const uint32_t lo = (uint32_t)std::min<uint64_t>(n_claim, (uint64_t)expect + expect / 32 + 64);
        FusedPlan ph = fused_plan(ctx, hi, flags, n_local_node);
        cap = hi; plan = ph;
        if (ph.fused && !ph.stage) {                            // shrink the slack until the array fits shared memory
            for (uint32_t c = hi; c >= lo; c -= std::max(1u, (hi - lo) / 16u)) {
                const FusedPlan pc = fused_plan(ctx, c, flags, n_local_node);
                if (pc.fused && pc.stage) { cap = c; plan = pc; break; }
                if (c == lo || c - lo < std::max(1u, (hi - lo) / 16u)) {
                    const FusedPlan pl = fused_plan(ctx, lo, flags, n_local_node);
                    if (pl.fused && pl.stage) { cap = lo; plan = pl; }
                    break;
                }
            }
        }
    }
    if (!plan.fused) cap = n_claim;                           // sort path: no layout depends on the count
    ctx->shard_last_n = n_claim;
    int rc = ensure_batch(ctx, std::max(cap, 1u), n_out, false);
    if (rc) return rc;
    if ((size_t)n_claim + 64 > ctx->cap_cclaims) {
        const size_t ncap = (size_t)n_claim + n_claim / 4 + 64;
        if ((rc = grow_nc(ctx, ctx->d_cclaims, 0, ncap))) return rc;
        if ((rc = grow_nc(ctx, ctx->d_coff, 0, ncap))) return rc;
        ctx->cap_cclaims = ncap;
    }
    // tiles: 256 claims while that keeps them within 1024 (one round of predecessor reads), else 2048; beyond 2M claims
    // the look-back kernel
    const uint32_t sc_rows = n_claim <= 256u * 1024u ? 1u : 8u;
    const bool sc_flat = n_claim <= 2048u * 1024u;
    const uint32_t n_tiles = std::max(1u, (n_claim + 256u * sc_rows - 1) / (256u * sc_rows));
    if ((size_t)n_tiles + 8 > ctx->cap_sc_status) {
        const size_t ncap = (size_t)n_tiles * 2 + 64;
        if ((rc = grow_nc(ctx, ctx->d_sc_status, 0, ncap))) return rc;
        CU(cudaMemsetAsync(ctx->d_sc_status, 0, ncap * 8, ctx->stream));
        ctx->cap_sc_status = ncap;
    }
    if (!ctx->d_sc_counts) {
        CU(cudaMalloc((void**)&ctx->d_sc_counts, 64)); CU(cudaMemsetAsync(ctx->d_sc_counts, 0, 64, ctx->stream));
        void* hp = nullptr;
        CU(cudaHostAlloc(&hp, 64, cudaHostAllocMapped));
        memset(hp, 0, 64);
        ctx->h_sc_counts = (volatile uint32_t*)hp;
        CU(cudaHostGetDevicePointer((void**)&ctx->h_sc_counts_dev, hp, 0));
    }
    if (!ctx->d_ticket) { CU(cudaMalloc((void**)&ctx->d_ticket, 64)); CU(cudaMemsetAsync(ctx->d_ticket, 0, 64, ctx->stream)); }
    if (!ctx->d_rank_slots) { CU(cudaMalloc((void**)&ctx->d_rank_slots, 2 * PEER_MAX * 4)); CU(cudaMemsetAsync(ctx->d_rank_slots, 0, 2 * PEER_MAX * 4, ctx->stream)); }
    const bool counted = gather && ctx->shard_map_on;          // every rank counts every rank's slots: the gather needs no header
    // result table
    uint2* table = nullptr; PktGather g; memset(&g, 0, sizeof g);
    if (gather) {
        ctx->peer_epoch += 1;
        g = make_gather(ctx, ctx->peer_epoch, 0, 0, 0, ctx->d_sc_counts + 1);
        table = g.table;
    } else {
        if ((size_t)n_out + 16 > ctx->cap_gtable) {
            const size_t ncap = (size_t)n_out + n_out / 4 + 64;
            if ((rc = grow_nc(ctx, ctx->d_gtable, 0, ncap))) return rc;
            ctx->cap_gtable = ncap;
        }
        table = ctx->d_gtable;
    }
    // 1. this rank's claims, compacted in input order, node indices local to the shard
    ctx->shard_epoch += 1;
    if (counted) g.expect = ctx->d_rank_slots + (ctx->shard_epoch & 1u) * PEER_MAX;
    ShardArgs sa; memset(&sa, 0, sizeof sa);
    if (counted) {
        sa.world = (uint32_t)ctx->world; sa.stray_rank = ctx->shard_stray_rank; sa.rank_slots = ctx->d_rank_slots;
        for (int r = 0; r < ctx->world; ++r) sa.rank_hi[r] = ctx->shard_bounds[r + 1];
    }
    sa.claims = (const uint4*)d_claims; sa.n_claim = n_claim; sa.out_off = d_out_off;
    sa.node_lo = ctx->shard_lo; sa.node_hi = ctx->shard_hi; sa.n_node_global = ctx->n_node; sa.take_stray = ctx->shard_stray;
    sa.have_off = d_out_off != nullptr;
    sa.cclaims = ctx->d_cclaims; sa.coff = ctx->d_coff; sa.cap = (uint32_t)std::min<size_t>(ctx->cap_cclaims, 0xFFFFFFFFu);
    sa.status = ctx->d_sc_status; sa.ticket = ctx->d_ticket + 4; sa.n_tiles = n_tiles; sa.epoch = ctx->shard_epoch;
    sa.counts = ctx->d_sc_counts; sa.h_counts = ctx->h_sc_counts_dev; sa.err = err_of(ctx);
    if (getenv("DRA_TIMELINE")) {                              // instrumentation: two words, read back with dra_debug_shard_times
        if (!ctx->d_sc_times) CU(cudaMalloc((void**)&ctx->d_sc_times, 16));
        const unsigned long long init[2] = {~0ull, 0ull};
        CU(cudaMemcpyAsync(ctx->d_sc_times, init, 16, cudaMemcpyHostToDevice, ctx->stream));
        sa.timeline = ctx->d_sc_times;
    }
    // (launched by launch_allocate: as its own kernel, or inside k_fused)
    // 2. the usual chain on the shard's view of the inventory, results at the claims' GLOBAL slots
    AllocView view; view.node_lo = ctx->shard_lo; view.n_node = n_local_node; view.n_dev = ctx->d_sc_counts; view.have_off = d_out_off != nullptr;
    view.sa = &sa; view.sc_rows = sc_rows; view.sc_flat = sc_flat;
    bool tail_done = false;
    rc = launch_allocate(ctx, ctx->d_cclaims, cap, ctx->d_coff, table, n_out, flags, gather ? &g : nullptr, &tail_done, &view);
    if (rc) return rc;
    if (gather && !tail_done && (rc = launch_gather(ctx, g, ctx->d_cclaims, cap, ctx->d_sc_counts, ctx->d_coff, n_out, d_out_off != nullptr, n_local_node))) return rc;
    ctx->gather_table = (const dra_out_rec*)table; ctx->gather_n_per = n_out; ctx->gather_len = n_out;
    return DRA_OK;
}

int dra_gather_table(dra_ctx* ctx, const dra_out_rec** d_table, uint32_t* n_per_rank) {
    if (!ctx || !d_table) return DRA_E_INVAL;
    if (!ctx->gather_table) return fail(ctx, DRA_E_STATE, "no gather has run");
    *d_table = ctx->gather_table;
    if (n_per_rank) *n_per_rank = ctx->gather_n_per;
    return DRA_OK;
}

int dra_gather_read(dra_ctx* ctx, dra_out_rec* out_all, uint32_t n_rec) {
    if (!ctx || !out_all) return DRA_E_INVAL;
    QUIESCE();
    if (!ctx->gather_table) return fail(ctx, DRA_E_STATE, "no gather has run");
    if (n_rec > (ctx->gather_len ? ctx->gather_len : (uint32_t)ctx->world * ctx->gather_n_per)) return fail(ctx, DRA_E_INVAL, "n_rec %u exceeds the table", n_rec);
    CU(cudaSetDevice(ctx->device));
    const size_t rb = (size_t)n_rec * 8;
    const bool direct = is_pinned(out_all, rb);
    int rc;
    if (!direct && (rc = grow_pinned(ctx, ctx->h_out, ctx->h_out_cap, rb))) return rc;
    CU(cudaMemcpyAsync(direct ? (void*)out_all : (void*)ctx->h_out, ctx->gather_table, rb, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if ((rc = check_err(ctx))) return rc;
    if (!direct) memcpy(out_all, ctx->h_out, rb);
    return DRA_OK;
}

// ---- adjacent integer searches, batched ------------------------------------------------
int dra_mps_limits_batch(dra_ctx* ctx, const int64_t* bytes, uint32_t n, int64_t* mib, uint8_t* valid) {
    if (!ctx || (n && (!bytes || !mib || !valid))) return DRA_E_INVAL;
    QUIESCE();
    if (!n) return DRA_OK;
    CU(cudaSetDevice(ctx->device));
    void *d_b = nullptr, *d_m = nullptr, *d_v = nullptr;
    CU(cudaMalloc(&d_b, (size_t)n * 8)); CU(cudaMalloc(&d_m, (size_t)n * 8)); CU(cudaMalloc(&d_v, n));
    cudaError_t e = cudaMemcpyAsync(d_b, bytes, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream);
    k_mps_limits<<<(n + 255) / 256, 256, 0, ctx->stream>>>((const long long*)d_b, n, (long long*)d_m, (uint8_t*)d_v);
    ctx->launches += 1;
    if (e == cudaSuccess) e = cudaMemcpyAsync(mib, d_m, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(valid, d_v, n, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_b); cudaFree(d_m); cudaFree(d_v);
    if (e != cudaSuccess) return fail(ctx, DRA_E_CUDA, "dra_mps_limits_batch: %s", cudaGetErrorString(e));
    return DRA_OK;
}

int dra_imex_offsets_batch(dra_ctx* ctx, const int32_t* used, const uint32_t* dom_off, uint32_t n_dom, int32_t step, int32_t limit, int32_t* out) {
    if (!ctx || !dom_off || (n_dom && !out) || step <= 0 || limit < 0) return DRA_E_INVAL;
    QUIESCE();
    if (!n_dom) return DRA_OK;
    const uint32_t n_used = dom_off[n_dom];
    if (n_used && !used) return DRA_E_INVAL;
    CU(cudaSetDevice(ctx->device));
    void *d_u = nullptr, *d_o = nullptr, *d_r = nullptr;
    CU(cudaMalloc(&d_u, (size_t)n_used * 4 + 16)); CU(cudaMalloc(&d_o, ((size_t)n_dom + 1) * 4)); CU(cudaMalloc(&d_r, (size_t)n_dom * 4));
    cudaError_t e = cudaSuccess;
    if (n_used) e = cudaMemcpyAsync(d_u, used, (size_t)n_used * 4, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_o, dom_off, ((size_t)n_dom + 1) * 4, cudaMemcpyHostToDevice, ctx->stream);
    k_imex_offsets<<<(n_dom + 7) / 8, 256, 0, ctx->stream>>>((const int*)d_u, (const uint32_t*)d_o, n_dom, step, limit, (int*)d_r);
    ctx->launches += 1;
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_r, (size_t)n_dom * 4, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_u); cudaFree(d_o); cudaFree(d_r);
    if (e != cudaSuccess) return fail(ctx, DRA_E_CUDA, "dra_imex_offsets_batch: %s", cudaGetErrorString(e));
    return DRA_OK;
}

// One-time calibration of the single-launch / sort-path crossover on THIS device and THIS inventory: synthetic MIG claims
// (uniform over the nodes) at a few batch sizes through both chains, CUDA events, best of 5; the live inventory is
// saved and restored.  Sets the claim count up to which dra_allocate_batch takes the single-launch kernel.
int dra_calibrate(dra_ctx* ctx, uint32_t* crossover_claims) {
    if (!ctx) return DRA_E_INVAL;
    QUIESCE();
    if (!ctx->d_inv_live || !ctx->n_node) return fail(ctx, DRA_E_STATE, "dra_set_inventory has not been called");
    CU(cudaSetDevice(ctx->device));
    const uint32_t sizes[] = {2000, 4000, 6000, 8000, 10000, 12000, 14000, 16000, 20000, 24000};
    const uint32_t n_max = 24000;
    int rc = ensure_batch(ctx, n_max, n_max, true);
    if (rc) return rc;
    std::vector<dra_claim_rec> h(n_max);
    uint64_t x = 0x9E3779B97F4A7C15ull;
    for (uint32_t i = 0; i < n_max; ++i) {
        x ^= x << 13; x ^= x >> 7; x ^= x << 17;
        h[i].kind = DRA_KIND_MIG; h[i].profile = (uint8_t)((x >> 40) % 3); h[i].count = 1;
        h[i].node = (uint32_t)(x % ctx->n_node); h[i].mem_limit_mib = 0; h[i].group = 0;
    }
    uint4* d_save = nullptr;
    CU(cudaMalloc((void**)&d_save, (size_t)ctx->n_gpu * 16 + 16));
    CU(cudaMemcpy(d_save, ctx->d_inv_live, (size_t)ctx->n_gpu * 16, cudaMemcpyDeviceToDevice));
    CU(cudaMemcpy(ctx->d_claims, h.data(), (size_t)n_max * 16, cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1; CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
    const uint32_t saved_flags = ctx->cfg_flags;
    const uint32_t saved_a = ctx->fused_max_one_wave, saved_b = ctx->fused_max_multi;
    const bool prof = ctx->profiling; ctx->profiling = false;
    uint32_t best = 0; bool lost = false;
    for (uint32_t n : sizes) {
        float t[2] = {1e30f, 1e30f};
        for (int path = 0; path < 2 && rc == DRA_OK; ++path) {
            ctx->cfg_flags = (saved_flags & ~DRA_CFG_NO_FUSED) | (path ? DRA_CFG_NO_FUSED : 0u);
            ctx->fused_max_one_wave = ctx->fused_max_multi = path ? 0u : 0xFFFFFFFFu;
            if (path == 0 && !fused_plan(ctx, n, DRA_F_FRESH_INVENTORY, ctx->n_node).fused) { t[0] = 1e30f; continue; }   // not launchable at this size
            for (int rep = 0; rep < 6 && rc == DRA_OK; ++rep) {
                cudaEventRecord(e0, ctx->stream);
                rc = launch_allocate(ctx, ctx->d_claims, n, nullptr, ctx->d_out, n, DRA_F_FRESH_INVENTORY);
                cudaEventRecord(e1, ctx->stream);
                if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = fail(ctx, DRA_E_CUDA, "calibration run failed");
                float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
                if (rep) t[path] = std::min(t[path], ms);
            }
        }
        if (rc) break;
        if (t[0] <= t[1] && !lost) best = n; else lost = true;          // the first size at which the sort path wins ends the range
    }
    ctx->cfg_flags = saved_flags; ctx->profiling = prof;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaMemcpy(ctx->d_inv_live, d_save, (size_t)ctx->n_gpu * 16, cudaMemcpyDeviceToDevice);
    cudaFree(d_save);
    (void)check_err(ctx);
    if (rc) { ctx->fused_max_one_wave = saved_a; ctx->fused_max_multi = saved_b; return rc; }
    if ((int)(ctx->n_node + 1) <= ctx->n_sm) { ctx->fused_max_one_wave = best; ctx->fused_max_multi = saved_b; }
    else { ctx->fused_max_multi = best; ctx->fused_max_one_wave = saved_a; }
    ctx->state_epoch++;
    if (crossover_claims) *crossover_claims = best;
    return DRA_OK;
}

// ---- resident mode ---------------------------------------------------------------------------------------------------
int dra_serve_start(dra_ctx* ctx) {
    if (!ctx) return DRA_E_INVAL;
    if (!ctx->d_inv_live) return fail(ctx, DRA_E_STATE, "dra_set_inventory has not been called");
    if (!ctx->coop_ok || (int)(ctx->n_node + 1) > ctx->n_sm) return fail(ctx, DRA_E_STATE, "resident mode needs every node's CTA resident at once: %u nodes, %d SMs", ctx->n_node, ctx->n_sm);
    int rc = serve_stop(ctx);
    if (rc) return rc;
    return serve_start(ctx);
}
int dra_serve_stop(dra_ctx* ctx) { if (!ctx) return DRA_E_INVAL; return serve_stop(ctx); }
uint64_t dra_serve_batches(const dra_ctx* ctx) { return ctx ? ctx->serve_batches : 0; }

// ---- host memory + instrumentation -----------------------------------------------------------------

void* dra_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 16, cudaHostAllocDefault) != cudaSuccess) return nullptr;
    std::lock_guard<std::mutex> lk(g_pin_mu);
    g_pinned[(uintptr_t)p] = bytes ? bytes : 16;
    return p;
}

void dra_host_free(void* p) {
    if (!p) return;
    { std::lock_guard<std::mutex> lk(g_pin_mu); g_pinned.erase((uintptr_t)p); }
    cudaFreeHost(p);
}

uint64_t dra_launch_count(const dra_ctx* ctx) { return ctx ? ctx->launches : 0; }

namespace { __global__ void k_noop(int x) { extern __shared__ uint8_t sm[]; if (x == 12345) sm[0] = 1; } }

// Instrumentation: enqueue an empty kernel with the given shape (calibrates the launch floor of the box).
int dra_debug_noop(dra_ctx* ctx, uint32_t grid, uint32_t block, uint32_t smem) {
    if (!ctx) return DRA_E_INVAL;
    QUIESCE();
    CU(cudaSetDevice(ctx->device));
    if (smem > 48 * 1024) CU(cudaFuncSetAttribute(k_noop, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_noop<<<grid, block, smem, ctx->stream>>>(0);
    return DRA_OK;
}

int dra_debug_serve_times(dra_ctx* ctx, unsigned long long* four) {
    if (!ctx || !four || !ctx->d_go) return 0;
    cudaSetDevice(ctx->device);
    if (serve_stop(ctx)) return 0;
    cudaMemcpy(four, ctx->d_go + 32, 32, cudaMemcpyDeviceToHost);
    return 4;
}

int dra_debug_shard_times(dra_ctx* ctx, unsigned long long* two) {
    if (!ctx || !two || !ctx->d_sc_times) return 0;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    cudaMemcpy(two, ctx->d_sc_times, 16, cudaMemcpyDeviceToHost);
    return 2;
}

int dra_debug_timeline(dra_ctx* ctx, unsigned long long* host, uint32_t n) {
    if (!ctx || !host || !ctx->d_timeline) return 0;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    uint32_t m = std::min(n, ctx->tl_n);
    cudaMemcpy(host, ctx->d_timeline, (size_t)m * 8, cudaMemcpyDeviceToHost);
    return (int)m;
}

int dra_set_profiling(dra_ctx* ctx, int enabled) {
    if (!ctx) return DRA_E_INVAL;
    ctx->profiling = enabled != 0;
    return DRA_OK;
}

int dra_get_timings(dra_ctx* ctx, float* us, int n) {
    if (!ctx || !us) return 0;
    int m = std::min(n, 5);
    for (int i = 0; i < m; ++i) us[i] = ctx->timings[i];
    return m;
}

}  // extern "C"
