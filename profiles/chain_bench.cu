// chain_bench.cu — microbenchmark of the first-fit serial step (profiles/, not product code).
//
// One CTA, one warp, lane = GPU (8 valid lanes), NREC prepared records in shared memory, clock64 around the
// loop.  Each variant removes or replaces one ingredient of the production step (k_fused / segment_run fast
// loop) so that the cost of that ingredient can be read off as a difference.
//
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o chain_bench chain_bench.cu && ./chain_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define FULL 0xFFFFFFFFu
constexpr int NREC = 64;

__device__ __forceinline__ uint32_t lop3_or(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d; asm("lop3.b32 %0, %1, %2, %3, 0xFE;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d;
}
__device__ __forceinline__ uint32_t lop3_nor(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d; asm("lop3.b32 %0, %1, %2, %3, 0x01;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d;
}
__device__ __forceinline__ void sts32_if(bool p, uint32_t addr, uint32_t v) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %0, 0;\n\t@q st.shared.u32 [%1], %2;\n\t}" :: "r"((uint32_t)p), "r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v; asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr)); return v;
}

struct Rec { uint4 r0, r1; };     // r0 = shift schedule, r1 = {smask | prof<<16, dst, size<<8, sizebits}

template <int V>
__global__ void __launch_bounds__(32, 1) k_chain(long long* cycles, uint32_t* sink, int reps) {
    __shared__ __align__(16) Rec recs[NREC];
    __shared__ uint32_t res[NREC];
    const uint32_t lane = threadIdx.x, lanebit = 1u << lane, lemask = lanebit | (lanebit - 1u);
    // cfg2 mix: 40 % 1g (size 1), 30 % 2g (2), 20 % 3g (4), 10 % 7g (8); A100-40GB start masks
    for (int i = lane; i < NREC; i += 32) {
        uint32_t h = (i * 2654435761u) >> 24, size, sm, prof;
        if (h < 102) { size = 1; sm = 0x7F; prof = 0; } else if (h < 179) { size = 2; sm = 0x15; prof = 1; }
        else if (h < 230) { size = 4; sm = 0x11; prof = 2; } else { size = 8; sm = 1; prof = 4; }
        uint32_t c = 1, s[4];
        for (int k = 0; k < 4; ++k) { s[k] = min(c, size - c) & 15u; c += s[k]; }
        recs[i].r0 = make_uint4(s[0], s[1], s[2], s[3]);
        recs[i].r1 = make_uint4(sm | (prof << 16), i, size << 8, (1u << size) - 1u);
    }
    __syncwarp();
    const uint32_t rec_addr = (uint32_t)__cvta_generic_to_shared(recs), res_addr = (uint32_t)__cvta_generic_to_shared(res);
    const uint32_t gate = lane < 8 ? 0xFFFFu : 0u;
    const bool offer_any = true;
    uint32_t busy = 0, bad = 0, nocap = 0, acc = 0;
    long long best = 1ll << 60;
    for (int rep = 0; rep < reps; ++rep) {
        busy = 0; bad = 0; nocap = 0;
        __syncwarp();
        const long long t0 = clock64();
        uint4 a0 = lds128(rec_addr), a1 = lds128(rec_addr + 16), b0, b1;
        auto stepf = [&](const uint4 r0, const uint4 r1, const uint32_t q) {
            const uint32_t pj = (r1.x >> 16) & 0xFFu, smask = r1.x & 0xFFFFu, dbit = 1u << pj;
            if (V == 0) {                      // production r01e: 4 log rounds, branch on failure
                const uint32_t deadm = bad | nocap;
                uint32_t t = ~busy & gate;
                t &= t >> r0.x; t &= t >> r0.y; t &= t >> r0.z; t &= t >> r0.w;
                uint32_t cand = t & smask;
                cand = ((deadm >> pj) & 1u) ? 0u : cand;
                const uint32_t b = __ballot_sync(FULL, cand != 0);
                const uint32_t low = cand & (0u - cand);
                const bool win = (b & (0u - b)) == lanebit;
                busy |= win ? r1.w * low : 0u;
                sts32_if(win, res_addr + (q << 2), lane | (low << 8));
                if (b == 0) {
                    uint32_t stt;
                    if ((deadm >> pj) & 1u) stt = ((bad >> pj) & 1u) ? 2 : 1;
                    else {
                        const bool any = __ballot_sync(FULL, gate != 0 && smask != 0) != 0;
                        if (any) nocap |= dbit; else bad |= dbit;
                        stt = any ? 1 : 2;
                    }
                    sts32_if(lane == 0, res_addr + (q << 2), 0xFFu | (stt << 24));
                }
            } else {
                // V1: tree fit, branch-free (r01f candidate).  V2: V1 without the fail bookkeeping.
                // V3: V2 without the result store.  V4: V3 with the ballot replaced by a shuffle-free local test
                // (chain without any warp-wide op).  V5: ballot only (cand = busy-independent).  V6: V1 with 4 log rounds.
                const uint32_t sz1 = ((r1.z >> 8) & 0xFFu) - 1u;
                const uint32_t sm = (V == 1 || V == 6) ? (((bad | nocap) & dbit) ? 0u : (smask & gate)) : (smask & gate);
                const uint32_t u = busy, nsm = ~sm;
                uint32_t cand;
                if (V == 6) {
                    uint32_t t = ~busy;
                    t &= t >> r0.x; t &= t >> r0.y; t &= t >> r0.z; t &= t >> r0.w;
                    cand = t & sm;
                } else if (V == 5) {
                    cand = sm & ~(q & 1u);
                } else {
                    const uint32_t n0 = lop3_or(u, u >> min(1u, sz1), u >> min(2u, sz1));
                    const uint32_t n1 = lop3_or(u >> min(3u, sz1), u >> min(4u, sz1), u >> min(5u, sz1));
                    const uint32_t n2 = lop3_or(u >> min(6u, sz1), u >> min(7u, sz1), nsm);
                    cand = lop3_nor(n0, n1, n2);
                }
                uint32_t b;
                if (V == 4) b = cand ? lanebit : 0u;            // no warp-wide op: every lane "wins" for itself
                else b = __ballot_sync(FULL, cand != 0);
                const uint32_t low = cand & (0u - cand);
                const bool win = (b & lemask) == lanebit;
                busy |= win ? r1.w * low : 0u;
                if (V != 3 && V != 4 && V != 5) sts32_if(win, res_addr + (q << 2), lane | (low << 8));
                if (V == 1 || V == 6) {
                    const bool fail = b == 0, offered = offer_any && smask != 0;
                    nocap |= (fail && offered) ? dbit : 0u;
                    bad |= (fail && !offered) ? dbit : 0u;
                    sts32_if(fail && lane == 0, res_addr + (q << 2), 0xFFu | ((offered ? 1u : 2u) << 24));
                }
            }
        };
        for (uint32_t q = 0; q < NREC; q += 2) {
            const uint32_t nb = rec_addr + ((q + 1) << 5);
            if (q + 1 < NREC) { b0 = lds128(nb); b1 = lds128(nb + 16); }
            stepf(a0, a1, q);
            if (q + 1 >= NREC) break;
            if (q + 2 < NREC) { a0 = lds128(nb + 32); a1 = lds128(nb + 48); }
            stepf(b0, b1, q + 1);
        }
        __syncwarp();
        const long long t1 = clock64();
        best = min(best, t1 - t0);
        acc += busy + bad + nocap;
    }
    if (lane == 0) { cycles[0] = best; sink[0] = acc + res[5]; }
}

// VOTE / SHFL / LDS latency: dependent chains of N ops
__global__ void __launch_bounds__(32, 1) k_lat(long long* out) {
    const uint32_t lane = threadIdx.x;
    __shared__ uint32_t sm[64];
    sm[lane] = lane; sm[lane + 32] = lane;
    __syncwarp();
    uint32_t x = lane + 1;
    long long t0 = clock64();
    #pragma unroll 1
    for (int i = 0; i < 256; ++i) { const uint32_t b = __ballot_sync(FULL, x != 0); x = (b >> lane) | 1u; }
    long long t1 = clock64();
    out[0] = (t1 - t0) / 256;               // ISETP + VOTE + SHF + LOP
    uint32_t y = x;
    t0 = clock64();
    #pragma unroll 1
    for (int i = 0; i < 256; ++i) { y = (y >> 1) | 1u; y = (y << 1) ^ lane; }    // 4 dependent ALU ops
    t1 = clock64();
    out[1] = (t1 - t0) / 256;
    uint32_t z = y & 31u;
    t0 = clock64();
    #pragma unroll 1
    for (int i = 0; i < 256; ++i) z = __shfl_sync(FULL, z, z & 31u);
    t1 = clock64();
    out[2] = (t1 - t0) / 256;
    uint32_t w = z & 31u;
    t0 = clock64();
    #pragma unroll 1
    for (int i = 0; i < 256; ++i) w = ((volatile uint32_t*)sm)[w & 63u];
    t1 = clock64();
    out[3] = (t1 - t0) / 256;
    // uniform taken branch cost: loop with an opaque always-true condition guarding a tiny block
    uint32_t v = w; uint32_t one = (out != nullptr);
    t0 = clock64();
    #pragma unroll 1
    for (int i = 0; i < 256; ++i) { if (one) { asm volatile("" ::: "memory"); v = v * 3u + 1u; } else { v = sm[v & 63u] + 7u; sm[v & 63u] = v; } }
    t1 = clock64();
    out[4] = (t1 - t0) / 256;
    if (lane == 0) out[5] = x + y + z + w + v;
}


// ---- second batch: lean steps.  State is `free` (= ~busy & gate), the start mask is static (no dead-memo feedback:
// a dead shape simply finds no candidate), the fail word is pre-initialised outside the loop, the winner test is one
// LOP3 with a predicate output, the update is a select between two precomputed values.
//   W: 0 = 4 log rounds, 1 = 3 log rounds (sizes <= 8), 2 = two-level tree (2+2 shifts), 3 = one-level tree (7 shifts)
template <int W>
__global__ void __launch_bounds__(32, 1) k_lean(long long* cycles, uint32_t* sink, int reps) {
    __shared__ __align__(16) Rec recs[NREC];
    __shared__ uint32_t res[NREC];
    const uint32_t lane = threadIdx.x, lanebit = 1u << lane, lemask = lanebit | (lanebit - 1u);
    for (int i = lane; i < NREC; i += 32) {
        uint32_t h = (i * 2654435761u) >> 24, size, sm, prof;
        if (h < 102) { size = 1; sm = 0x7F; prof = 0; } else if (h < 179) { size = 2; sm = 0x15; prof = 1; }
        else if (h < 230) { size = 4; sm = 0x11; prof = 2; } else { size = 8; sm = 1; prof = 4; }
        uint32_t c = 1, s[4];
        for (int k = 0; k < 4; ++k) { s[k] = min(c, size - c) & 15u; c += s[k]; }
        if (W == 2) { const uint32_t z = size - 1; s[0] = min(1u, z); s[1] = min(2u, z); s[3] = z - s[1]; s[2] = min(3u, s[3]); }
        recs[i].r0 = make_uint4(s[0], s[1], s[2], s[3]);
        recs[i].r1 = make_uint4(sm | (prof << 16), i, size << 8, (1u << size) - 1u);
    }
    __syncwarp();
    const uint32_t rec_addr = (uint32_t)__cvta_generic_to_shared(recs), res_addr = (uint32_t)__cvta_generic_to_shared(res);
    const uint32_t gate = lane < 8 ? 0xFFFFu : 0u;
    uint32_t acc = 0, fre = 0;
    long long best = 1ll << 60;
    for (int rep = 0; rep < reps; ++rep) {
        fre = gate;
        for (int i = lane; i < NREC; i += 32) res[i] = 0x010000FFu;      // fail word, overwritten by a winner
        __syncwarp();
        const long long t0 = clock64();
        uint4 a0 = lds128(rec_addr), a1 = lds128(rec_addr + 16), b0, b1;
        auto stepf = [&](const uint4 r0, const uint4 r1, const uint32_t q) {
            const uint32_t sm = r1.x & gate;                 // static: 16 bits
            uint32_t cand;
            if (W == 0) { uint32_t t = fre; t &= t >> r0.x; t &= t >> r0.y; t &= t >> r0.z; t &= t >> r0.w; cand = t & sm; }
            else if (W == 1) { uint32_t t = fre; t &= t >> r0.x; t &= t >> r0.y; t &= t >> r0.z; cand = t & sm; }
            else if (W == 2) {
                uint32_t n; asm("lop3.b32 %0, %1, %2, %3, 0x80;" : "=r"(n) : "r"(fre), "r"(fre >> r0.x), "r"(fre >> r0.y));
                uint32_t m; asm("lop3.b32 %0, %1, %2, %3, 0x80;" : "=r"(m) : "r"(n), "r"(n >> r0.z), "r"(n >> r0.w));
                cand = m & sm;
            } else {
                const uint32_t z = ((r1.z >> 8) & 0xFFu) - 1u, u = fre;
                uint32_t n0, n1, n2;
                asm("lop3.b32 %0, %1, %2, %3, 0x80;" : "=r"(n0) : "r"(u), "r"(u >> min(1u, z)), "r"(u >> min(2u, z)));
                asm("lop3.b32 %0, %1, %2, %3, 0x80;" : "=r"(n1) : "r"(u >> min(3u, z)), "r"(u >> min(4u, z)), "r"(u >> min(5u, z)));
                asm("lop3.b32 %0, %1, %2, %3, 0x80;" : "=r"(n2) : "r"(u >> min(6u, z)), "r"(u >> min(7u, z)), "r"(sm));
                asm("lop3.b32 %0, %1, %2, %3, 0x80;" : "=r"(cand) : "r"(n0), "r"(n1), "r"(n2));
            }
            const uint32_t b = __ballot_sync(FULL, cand != 0);
            const uint32_t low = cand & (0u - cand);
            const uint32_t freW = fre & ~(r1.w * low);       // off the chain (overlaps the vote)
            const bool lose = ((b & lemask) ^ lanebit) != 0;
            fre = lose ? fre : freW;
            sts32_if(!lose, res_addr + (q << 2), lane + (cand << 8));    // start = ffs(cand) in the epilogue
        };
        for (uint32_t q = 0; q < NREC; q += 2) {
            const uint32_t nb = rec_addr + ((q + 1) << 5);
            if (q + 1 < NREC) { b0 = lds128(nb); b1 = lds128(nb + 16); }
            stepf(a0, a1, q);
            if (q + 1 >= NREC) break;
            if (q + 2 < NREC) { a0 = lds128(nb + 32); a1 = lds128(nb + 48); }
            stepf(b0, b1, q + 1);
        }
        __syncwarp();
        const long long t1 = clock64();
        best = min(best, t1 - t0);
        acc += fre + res[(rep * 7) & 63];
    }
    // outcome digest, to check that all variants agree
    uint32_t dig = 0;
    for (int i = lane; i < NREC; i += 32) { const uint32_t r = res[i]; const uint32_t w = r & 0xFF; dig += (w == 0xFF ? 0xABCDu : (w * 16u + (__ffs(r >> 8) - 1))) * (i + 1); }
    for (int o = 16; o; o >>= 1) dig += __shfl_xor_sync(FULL, dig, o);
    if (lane == 0) { cycles[0] = best; sink[0] = dig; sink[1] = acc; }
}

template <int W> static void run_lean(const char* what, long long* d_c, uint32_t* d_s) {
    k_lean<W><<<1, 32>>>(d_c, d_s, 20);
    long long c; uint32_t dg; cudaMemcpy(&c, d_c, 8, cudaMemcpyDeviceToHost); cudaMemcpy(&dg, d_s, 4, cudaMemcpyDeviceToHost);
    printf("L%d %-72s %6lld cycles / %d records = %6.1f cycles/record   digest %08x\n", W, what, c, NREC, (double)c / NREC, dg);
}

template <int V> static void run(const char* what, long long* d_c, uint32_t* d_s) {
    k_chain<V><<<1, 32>>>(d_c, d_s, 20);
    long long c; cudaMemcpy(&c, d_c, 8, cudaMemcpyDeviceToHost);
    printf("V%d %-72s %6lld cycles / %d records = %6.1f cycles/record\n", V, what, c, NREC, (double)c / NREC);
}

int main() {
    long long* d_c; uint32_t* d_s;
    cudaMalloc(&d_c, 64); cudaMalloc(&d_s, 64);
    run<0>("production r01e: 4 log rounds + fail branch", d_c, d_s);
    run<6>("4 log rounds, branch-free fail bookkeeping", d_c, d_s);
    run<1>("tree fit (7 shifts + LOP3 tree), branch-free fail bookkeeping", d_c, d_s);
    run<2>("tree fit, no fail bookkeeping", d_c, d_s);
    run<3>("tree fit, no fail bookkeeping, no result store", d_c, d_s);
    run<4>("tree fit, no warp-wide op at all (lane-local chain)", d_c, d_s);
    run<5>("ballot + winner + busy update only (no fit)", d_c, d_s);
    run_lean<0>("lean: 4 log rounds", d_c, d_s);
    run_lean<1>("lean: 3 log rounds (sizes <= 8)", d_c, d_s);
    run_lean<2>("lean: two-level tree, 2+2 shifts", d_c, d_s);
    run_lean<3>("lean: one-level tree, 7 shifts", d_c, d_s);
    k_lat<<<1, 32>>>(d_c);
    long long l[6]; cudaMemcpy(l, d_c, 48, cudaMemcpyDeviceToHost);
    printf("latency: ISETP+VOTE+SHF+LOP chain %lld | 4 dependent ALU %lld | SHFL %lld | LDS %lld | loop iteration with a guarded block %lld\n",
           l[0], l[1], l[2], l[3], l[4]);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
