// peer_bench.cu — microbenchmark of the multi-GPU tail: how should a rank publish its OutRecs to its peers?
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o profiles/peer_bench profiles/peer_bench.cu
//   profiles/peer_bench [n_gpu] [iters]          (ONE process, peer access between all devices)
//
// Every device runs the same kernel (126 CTAs x 256 threads, like k_fused on cfg2).  The kernels first meet at a
// cross-GPU start barrier (so inter-process launch skew is NOT part of the number), stamp t0, "produce" 10,000
// OutRecs at permuted slots of their slice (unsorted input: a node's records are scattered over the slice), then
// publish the slice to every peer and wait until every peer's slice has arrived.  Reported: t_end - t0 of the last
// CTA, max over devices, median / p90 over iterations.
//
//   scatter   what r01 shipped: each CTA stores ITS records (8 B each, scattered) into every peer, one fence.sys per
//             CTA, ticket; the last CTA publishes the epoch flags and waits
//   scatter+c same stores, but completion by per-CTA remote counters (red.release.sys) instead of ticket -> flag
//   coalesced grid barrier, then every CTA pushes 1/126 of the slice as 16-byte stores; fence per CTA, ticket, flags
//   lastcta   ticket first; the last CTA pushes the whole slice (16-byte stores), one fence, flags
//   lasttma   ticket first; the last CTA stages the slice in shared memory and issues ONE bulk store per peer
//   flagonly  no data at all: start barrier -> ticket -> flags -> wait   (the protocol floor)
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

constexpr int MAXG = 8, GRID = 126, BLOCK = 256, NREC = 10000;
constexpr long long SPIN_LIMIT = 600000000ll;     // ~0.3 s: fail, never hang

struct Args {
    uint2* table[MAXG];       // every device's table: [world][NREC] records
    uint32_t* flags[MAXG];    // every device's flag block: [0..15] start, [16..31] done, [32] go, [33] counter
    uint32_t* ticket;         // local: [0] ticket, [1] gbar count, [2] gbar gen
    uint2* local;             // local scratch slice [NREC]
    const uint32_t* perm;     // permutation of 0..NREC-1 (local)
    unsigned long long* stamps;   // local: [iter][4]
    uint32_t* err;
    int world, rank, iter, mode;
};

__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ void st_rel_sys(uint32_t* p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_acq_sys(const uint32_t* p) { uint32_t v; asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ uint32_t ld_acq_gpu(const uint32_t* p) { uint32_t v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void red_rel_sys(uint32_t* p, uint32_t v) { asm volatile("red.release.sys.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

__device__ bool spin_ge(const uint32_t* p, uint32_t want, bool sys, uint32_t* err) {
    const long long t0 = clock64();
    while (true) {
        const uint32_t v = sys ? ld_acq_sys(p) : ld_acq_gpu(p);
        if ((int32_t)(v - want) >= 0) return true;
        if (clock64() - t0 > SPIN_LIMIT) { *err = 1; return false; }
    }
}

__device__ void grid_barrier(uint32_t* gb, uint32_t n, uint32_t* err) {
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t gen = ld_acq_gpu(gb + 1);
        __threadfence();
        if (atomicAdd(gb, 1u) == n - 1) { gb[0] = 0; __threadfence(); asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(gb + 1), "r"(gen + 1) : "memory"); }
        else spin_ge(gb + 1, gen + 1, false, err);
    }
    __syncthreads();
}

__global__ void __launch_bounds__(BLOCK, 1) k_proto(const Args a) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint32_t last_s;
    const uint32_t tid = threadIdx.x, cta = blockIdx.x, ep = (uint32_t)a.iter;
    uint32_t* myflags = a.flags[a.rank];
    // ---- cross-GPU start barrier (takes the launch skew out of the measurement) ----
    if (cta == 0) {
        if (tid < (uint32_t)a.world) { st_rel_sys(a.flags[tid] + a.rank, ep); spin_ge(myflags + tid, ep, true, a.err); }
        __syncthreads();
        if (tid == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(myflags + 32), "r"(ep) : "memory");
    } else if (tid == 0) spin_ge(myflags + 32, ep, false, a.err);
    __syncthreads();
    unsigned long long t0 = 0, t1 = 0, t2 = 0;
    if (tid == 0) t0 = gtime();
    // ---- produce: CTA c owns records c*80 .. (scattered slots of the slice, as with unsorted claims) ----
    const uint32_t per = (NREC + GRID - 1) / GRID;
    uint2* mine_in[MAXG];
    #pragma unroll
    for (int p = 0; p < MAXG; ++p) mine_in[p] = p < a.world ? a.table[p] + (size_t)a.rank * NREC : nullptr;
    const int mode = a.mode;
    if (mode != 5) {
        const uint32_t k = cta * per + tid;
        if (tid < per && k < NREC) {
            const uint32_t slot = a.perm[k];
            const uint2 rec = make_uint2(a.rank * 1000000u + k, ep);
            mine_in[a.rank][slot] = rec;
            if (mode == 0 || mode == 1)
                for (int p = 0; p < a.world; ++p) if (p != a.rank) mine_in[p][slot] = rec;
        }
    }
    if (mode == 2) {                                      // coalesced: everybody pushes a share after a grid barrier
        grid_barrier(a.ticket + 1, GRID, a.err);
        const uint32_t n16 = NREC / 2, share = (n16 + GRID - 1) / GRID;
        const uint4* src = reinterpret_cast<const uint4*>(mine_in[a.rank]);
        for (int p = 0; p < a.world; ++p) {
            if (p == a.rank) continue;
            uint4* dst = reinterpret_cast<uint4*>(mine_in[p]);
            for (uint32_t i = cta * share + tid; i < min(n16, (cta + 1) * share); i += BLOCK) dst[i] = __ldcg(src + i);
        }
    }
    __syncthreads();
    if (tid == 0) {
        if (mode == 0 || mode == 1 || mode == 2) __threadfence_system(); else __threadfence();
        t1 = gtime();
        if (mode == 1) {                                  // completion by remote counters
            for (int p = 0; p < a.world; ++p) red_rel_sys(a.flags[p] + 33, 1u);
        }
        last_s = atomicAdd(a.ticket, 1u) == GRID - 1 ? 1u : 0u;
    }
    __syncthreads();
    if (!last_s) { if (tid == 0) { a.stamps[(size_t)a.iter * 4 * GRID + cta * 4 + 0] = t0; a.stamps[(size_t)a.iter * 4 * GRID + cta * 4 + 1] = t1; } return; }
    if (tid == 0) { *a.ticket = 0; }
    __threadfence();
    __syncthreads();
    if (mode == 3) {                                      // the last CTA pushes the whole slice
        const uint4* src = reinterpret_cast<const uint4*>(mine_in[a.rank]);
        for (int p = 0; p < a.world; ++p) {
            if (p == a.rank) continue;
            uint4* dst = reinterpret_cast<uint4*>(mine_in[p]);
            for (uint32_t i = tid; i < NREC / 2; i += BLOCK) dst[i] = __ldcg(src + i);
        }
        __threadfence_system();
        __syncthreads();
    } else if (mode == 4) {                               // slice -> shared memory, one bulk store per peer
        const uint4* src = reinterpret_cast<const uint4*>(mine_in[a.rank]);
        uint4* s4 = reinterpret_cast<uint4*>(smem);
        for (uint32_t i = tid; i < NREC / 2; i += BLOCK) s4[i] = __ldcg(src + i);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid < (uint32_t)a.world && (int)tid != a.rank) {
            const uint32_t sa = (uint32_t)__cvta_generic_to_shared(smem);
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(mine_in[tid]), "r"(sa), "r"((uint32_t)NREC * 8u) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
            asm volatile("fence.proxy.async;" ::: "memory");
            __threadfence_system();
        }
        __syncthreads();
    }
    if (mode == 1) {
        if (tid == 0) spin_ge(myflags + 33, ep * (uint32_t)a.world * GRID, true, a.err);
    } else if (tid < (uint32_t)a.world) {
        __threadfence_system();
        st_rel_sys(a.flags[tid] + 16 + a.rank, ep);
        spin_ge(myflags + 16 + tid, ep, true, a.err);
    }
    __syncthreads();
    if (tid == 0) {
        t2 = gtime();
        unsigned long long* s = a.stamps + (size_t)a.iter * 4 * GRID + cta * 4;
        s[0] = t0; s[1] = t1; s[2] = t2; s[3] = 1;
    }
}

// one-way flag latency: dev0 writes, dev1 echoes; round trip / 2
__global__ void k_ping(uint32_t* mine, uint32_t* peer, int iters, int leader, unsigned long long* out, uint32_t* err) {
    unsigned long long t0 = gtime();
    for (int i = 1; i <= iters; ++i) {
        if (leader) { st_rel_sys(peer, (uint32_t)i); if (!spin_ge(mine, (uint32_t)i, true, err)) break; }
        else { if (!spin_ge(mine, (uint32_t)i, true, err)) break; st_rel_sys(peer, (uint32_t)i); }
    }
    if (leader) *out = gtime() - t0;
}

int main(int argc, char** argv) {
    int ndev = 0; CK(cudaGetDeviceCount(&ndev));
    int world = argc > 1 ? atoi(argv[1]) : ndev; world = std::min(world, std::min(ndev, MAXG));
    const int iters = argc > 2 ? atoi(argv[2]) : 300;
    printf("peer_bench: %d device(s), %d iterations, %d records (%d B slice), grid %d x %d\n", world, iters, NREC, NREC * 8, GRID, BLOCK);
    if (world < 2) { printf("needs >= 2 devices\n"); return 0; }
    std::vector<cudaStream_t> st(world);
    uint2* table[MAXG]; uint32_t* flags[MAXG]; uint32_t* ticket[MAXG]; uint2* local[MAXG]; uint32_t* perm[MAXG];
    unsigned long long* stamps[MAXG]; uint32_t* err[MAXG];
    std::vector<uint32_t> hperm(NREC);
    for (int i = 0; i < NREC; ++i) hperm[i] = i;
    uint64_t x = 88172645463325252ull;
    for (int i = NREC - 1; i > 0; --i) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; std::swap(hperm[i], hperm[x % (i + 1)]); }
    for (int d = 0; d < world; ++d) {
        CK(cudaSetDevice(d));
        for (int p = 0; p < world; ++p) if (p != d) { int ok = 0; CK(cudaDeviceCanAccessPeer(&ok, d, p)); if (!ok) { printf("no P2P %d->%d\n", d, p); return 0; } cudaError_t e = cudaDeviceEnablePeerAccess(p, 0); if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CK(e); (void)cudaGetLastError(); }
        CK(cudaStreamCreateWithFlags(&st[d], cudaStreamNonBlocking));
        CK(cudaMalloc(&table[d], (size_t)world * NREC * 8)); CK(cudaMemset(table[d], 0, (size_t)world * NREC * 8));
        CK(cudaMalloc(&flags[d], 256)); CK(cudaMemset(flags[d], 0, 256));
        CK(cudaMalloc(&ticket[d], 64)); CK(cudaMemset(ticket[d], 0, 64));
        CK(cudaMalloc(&local[d], NREC * 8));
        CK(cudaMalloc(&perm[d], NREC * 4)); CK(cudaMemcpy(perm[d], hperm.data(), NREC * 4, cudaMemcpyHostToDevice));
        CK(cudaMalloc(&stamps[d], (size_t)(iters + 2) * 4 * GRID * 8));
        CK(cudaMalloc(&err[d], 4)); CK(cudaMemset(err[d], 0, 4));
        CK(cudaFuncSetAttribute(k_proto, cudaFuncAttributeMaxDynamicSharedMemorySize, NREC * 8 + 128));
    }
    for (int d = 0; d < world; ++d) { CK(cudaSetDevice(d)); CK(cudaDeviceSynchronize()); }

    // ---- flag ping-pong between device 0 and device 1 ----
    {
        unsigned long long* out; CK(cudaSetDevice(0)); CK(cudaMalloc(&out, 8));
        const int n = 2000;
        CK(cudaSetDevice(1)); k_ping<<<1, 1, 0, st[1]>>>(flags[1] + 40, flags[0] + 40, n, 0, nullptr, err[1]);
        CK(cudaSetDevice(0)); k_ping<<<1, 1, 0, st[0]>>>(flags[0] + 40, flags[1] + 40, n, 1, out, err[0]);
        CK(cudaStreamSynchronize(st[0])); CK(cudaSetDevice(1)); CK(cudaStreamSynchronize(st[1]));
        unsigned long long ns = 0; CK(cudaMemcpy(&ns, out, 8, cudaMemcpyDeviceToHost));
        printf("flag ping-pong 0<->1: round trip %.2f us, one way %.2f us\n", ns / 1e3 / n, ns / 2e3 / n);
    }

    const char* names[] = {"scatter", "scatter+c", "coalesced", "lastcta", "lasttma", "flagonly"};
    uint32_t epoch = 0;
    for (int mode : {5, 0, 1, 2, 3, 4}) {
        // counters are cumulative per mode: reset everything
        for (int d = 0; d < world; ++d) { CK(cudaSetDevice(d)); CK(cudaDeviceSynchronize()); CK(cudaMemset(flags[d], 0, 256)); CK(cudaMemset(ticket[d], 0, 64)); CK(cudaMemset(table[d], 0, (size_t)world * NREC * 8)); CK(cudaMemset(stamps[d], 0, (size_t)(iters + 2) * 4 * GRID * 8)); }
        for (int d = 0; d < world; ++d) { CK(cudaSetDevice(d)); CK(cudaDeviceSynchronize()); }
        epoch = 0;
        std::vector<double> dur, t_fence;
        for (int it = 0; it < iters; ++it) {
            ++epoch;
            for (int d = 0; d < world; ++d) {
                Args a{};
                for (int p = 0; p < world; ++p) { a.table[p] = table[p]; a.flags[p] = flags[p]; }
                a.ticket = ticket[d]; a.local = local[d]; a.perm = perm[d]; a.stamps = stamps[d]; a.err = err[d];
                a.world = world; a.rank = d; a.iter = (int)epoch; a.mode = mode;
                CK(cudaSetDevice(d));
                k_proto<<<GRID, BLOCK, mode == 4 ? NREC * 8 + 128 : 0, st[d]>>>(a);
                CK(cudaGetLastError());
            }
            double worst = 0, worst_f = 0;
            for (int d = 0; d < world; ++d) {
                CK(cudaSetDevice(d)); CK(cudaStreamSynchronize(st[d]));
                std::vector<unsigned long long> h(4 * GRID);
                CK(cudaMemcpy(h.data(), stamps[d] + (size_t)epoch * 4 * GRID, 4 * GRID * 8, cudaMemcpyDeviceToHost));
                unsigned long long tmin = ~0ull, tend = 0, tf = 0;
                for (int c = 0; c < GRID; ++c) { tmin = std::min(tmin, h[c * 4]); tf = std::max(tf, h[c * 4 + 1]); if (h[c * 4 + 3]) tend = h[c * 4 + 2]; }
                worst = std::max(worst, (double)(tend - tmin) / 1e3); worst_f = std::max(worst_f, (double)(tf - tmin) / 1e3);
                uint32_t e = 0; CK(cudaMemcpy(&e, err[d], 4, cudaMemcpyDeviceToHost));
                if (e) { printf("mode %s: device %d timed out at iteration %d\n", names[mode], d, it); return 2; }
            }
            if (it >= 20) { dur.push_back(worst); t_fence.push_back(worst_f); }
        }
        // verify the last iteration's table on device 0 (modes that move data)
        if (mode != 5) {
            CK(cudaSetDevice(0));
            std::vector<uint2> h((size_t)world * NREC);
            CK(cudaMemcpy(h.data(), table[0], h.size() * 8, cudaMemcpyDeviceToHost));
            size_t bad = 0;
            for (int r = 0; r < world; ++r) for (int k = 0; k < NREC; ++k) { const uint2 v = h[(size_t)r * NREC + hperm[k]]; if (v.x != r * 1000000u + k || v.y != epoch) ++bad; }
            if (bad) printf("mode %s: %zu WRONG records\n", names[mode], bad);
        }
        std::sort(dur.begin(), dur.end()); std::sort(t_fence.begin(), t_fence.end());
        printf("%-10s world %d: tail (start barrier -> all slices here)  median %6.2f us  p90 %6.2f  min %6.2f   | to last CTA's fence: median %5.2f us\n",
               names[mode], world, dur[dur.size() / 2], dur[dur.size() * 9 / 10], dur[0], t_fence[t_fence.size() / 2]);
    }
    return 0;
}
