// atomic_microbench.cu -- what can B200's L2 do with random counter updates?
// Measures, for table sizes from L2-resident to HBM-resident, the throughput of the candidate
// update primitives for khmer-style saturating byte counters ("fraction of a
// measured L2-atomic peak from a microbenchmark over a table of the same size").
//   red_add    : atomicAdd without return (RED.ADD) on the containing u32        (NOT exact: carries)
//   atom_add   : atomicAdd with return (ATOM.ADD)
//   cas_blind  : one atomicCAS with a guessed expected value (transaction cost of a CAS)
//   ld_cas     : ld.cg + CAS loop = the exact saturating byte increment shipped in kv_sat_inc
//   ld_only    : the random 32-bit loads alone
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o atomic_microbench atomic_microbench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t mix(uint64_t x)
{
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}

template <int MODE>
__global__ void __launch_bounds__(256) bench(unsigned *table, uint64_t nbytes, uint64_t n_ops, unsigned *sink)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    unsigned acc = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_ops; i += stride) {
        uint64_t h = mix(i + 0x1234567);
        // 4 independent targets per item, like the 4 tables of a sketch
        unsigned *w[4]; unsigned sh[4], old[4];
#pragma unroll
        for (int t = 0; t < 4; t++) {
            uint64_t byte = mix(h + t) % nbytes;
            w[t] = table + (byte >> 2);
            sh[t] = (unsigned)(byte & 3) * 8;
        }
        if (MODE == 0) {
#pragma unroll
            for (int t = 0; t < 4; t++) atomicAdd(w[t], 1u << sh[t]);
        } else if (MODE == 1) {
#pragma unroll
            for (int t = 0; t < 4; t++) acc += atomicAdd(w[t], 1u << sh[t]);
        } else if (MODE == 2) {
#pragma unroll
            for (int t = 0; t < 4; t++) acc += atomicCAS(w[t], 0u, 1u << sh[t]);
        } else if (MODE == 3) {
#pragma unroll
            for (int t = 0; t < 4; t++) old[t] = __ldcg(w[t]);
#pragma unroll
            for (int t = 0; t < 4; t++) {
                unsigned o = old[t];
                while (((o >> sh[t]) & 255u) != 255u) {
                    unsigned a = o;
                    o = atomicCAS(w[t], a, a + (1u << sh[t]));
                    if (o == a) break;
                }
            }
        } else if (MODE == 4) {
#pragma unroll
            for (int t = 0; t < 4; t++) acc += __ldcg(w[t]);
        } else if (MODE == 5) {   // ld + CAS, first attempt of all four issued back to back, rare retry loop after
#pragma unroll
            for (int t = 0; t < 4; t++) old[t] = __ldcg(w[t]);
            unsigned res[4];
#pragma unroll
            for (int t = 0; t < 4; t++)
                res[t] = ((old[t] >> sh[t]) & 255u) != 255u ? atomicCAS(w[t], old[t], old[t] + (1u << sh[t])) : old[t];
#pragma unroll
            for (int t = 0; t < 4; t++) {
                if (res[t] == old[t]) continue;
                unsigned o = res[t];
                while (((o >> sh[t]) & 255u) != 255u) {
                    unsigned a = o;
                    o = atomicCAS(w[t], a, a + (1u << sh[t]));
                    if (o == a) break;
                }
            }
        } else if (MODE == 6) {   // ld + ATOM.ADD with return (speculative add, overflow check on the old value)
#pragma unroll
            for (int t = 0; t < 4; t++) old[t] = __ldcg(w[t]);
#pragma unroll
            for (int t = 0; t < 4; t++)
                if (((old[t] >> sh[t]) & 255u) != 255u) acc += (atomicAdd(w[t], 1u << sh[t]) >> sh[t]) & 255u;
        } else if (MODE == 7) {   // ld + RED.ADD
#pragma unroll
            for (int t = 0; t < 4; t++) old[t] = __ldcg(w[t]);
#pragma unroll
            for (int t = 0; t < 4; t++)
                if (((old[t] >> sh[t]) & 255u) != 255u) atomicAdd(w[t], 1u << sh[t]);
        } else if (MODE == 8) {   // ld.ca + CAS loop
#pragma unroll
            for (int t = 0; t < 4; t++) old[t] = *(volatile unsigned *)w[t];
#pragma unroll
            for (int t = 0; t < 4; t++) {
                unsigned o = old[t];
                while (((o >> sh[t]) & 255u) != 255u) {
                    unsigned a = o;
                    o = atomicCAS(w[t], a, a + (1u << sh[t]));
                    if (o == a) break;
                }
            }
        } else if (MODE >= 10 && MODE <= 13) {   // how the expected value is fetched: cv / relaxed.gpu / atom.or 0 / acquire
#pragma unroll
            for (int t = 0; t < 4; t++) {
                if (MODE == 10) old[t] = __ldcv(w[t]);
                else if (MODE == 11) asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(old[t]) : "l"(w[t]) : "memory");
                else if (MODE == 12) old[t] = atomicOr(w[t], 0u);
                else asm volatile("ld.global.lu.u32 %0, [%1];" : "=r"(old[t]) : "l"(w[t]) : "memory");
            }
#pragma unroll
            for (int t = 0; t < 4; t++) {
                unsigned o = old[t];
                while (((o >> sh[t]) & 255u) != 255u) {
                    unsigned a = o;
                    o = atomicCAS(w[t], a, a + (1u << sh[t]));
                    if (o == a) break;
                }
            }
        } else if (MODE == 14) {   // side bitmap (1 bit per counter, separate array) + speculative ATOM.ADD
            const unsigned *bitmap = table + (nbytes >> 2);   // caller allocates nbytes + nbytes/8
#pragma unroll
            for (int t = 0; t < 4; t++) {
                uint64_t byte = (uint64_t)((const char *)w[t] - (const char *)table) + (sh[t] >> 3);
                old[t] = __ldg(bitmap + (byte >> 5)) >> (byte & 31);
            }
#pragma unroll
            for (int t = 0; t < 4; t++)
                if (!(old[t] & 1u)) acc += (atomicAdd(w[t], 1u << sh[t]) >> sh[t]) & 255u;
        } else if (MODE == 9) {   // 16-bit CAS on the containing half-word (halves the false sharing)
#pragma unroll
            for (int t = 0; t < 4; t++) old[t] = __ldcg(w[t]);
#pragma unroll
            for (int t = 0; t < 4; t++) {
                unsigned short *hw = (unsigned short *)w[t] + (sh[t] >> 4);
                unsigned hs = sh[t] & 8;
                unsigned short o = (unsigned short)(old[t] >> (sh[t] & 16));
                while (((o >> hs) & 255u) != 255u) {
                    unsigned short a = o;
                    o = atomicCAS(hw, a, (unsigned short)(a + (1u << hs)));
                    if (o == a) break;
                }
            }
        }
    }
    if (acc == 0xdeadbeef) *sink = acc;
}

int main()
{
    const char *names[15] = {"red_add", "atom_add", "cas_blind", "ld_cas", "ld_only", "ld_cas_batched", "ld_atom_add", "ld_red_add", "ldca_cas", "ld_cas16", "ldcv_cas", "ldrelaxed_cas", "atomor0_cas", "ldlu_cas", "bitmap_atom_add"};
    const uint64_t sizes[] = {64ull << 20, 1ull << 30};
    const uint64_t n_items = 21000000;   // one C2 sample
    unsigned *sink;
    cudaMalloc(&sink, 4);
    int sm = 148;
    cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
    printf("mode,table_MB,ms,G_updates_per_s,G_items_per_s\n");
    for (uint64_t nbytes : sizes) {
        unsigned *table;
        if (cudaMalloc(&table, nbytes + nbytes / 8 + 256) != cudaSuccess) { printf("alloc %llu failed\n", (unsigned long long)nbytes); continue; }
        for (int mode = 0; mode < 15; mode++) {
            float best = 1e30f;
            for (int rep = 0; rep < 4; rep++) {
                cudaMemset(table, 0, nbytes + nbytes / 8 + 256);
                cudaEvent_t e0, e1;
                cudaEventCreate(&e0); cudaEventCreate(&e1);
                cudaEventRecord(e0);
                switch (mode) {
                case 0: bench<0><<<sm * 8, 256>>>(table, nbytes, n_items, sink); break;
                case 1: bench<1><<<sm * 8, 256>>>(table, nbytes, n_items, sink); break;
                case 2: bench<2><<<sm * 8, 256>>>(table, nbytes, n_items, sink); break;
                case 3: bench<3><<<sm * 8, 256>>>(table, nbytes, n_items, sink); break;
                case 4: bench<4><<<sm * 8, 256>>>(table, nbytes, n_items, sink); break;
                case 5: bench<5><<<sm * 8, 256>>>(table, nbytes, n_items, sink); break;
                case 6: bench<6><<<sm * 8, 256>>>(table, nbytes, n_items, sink); break;
                case 7: bench<7><<<sm * 8, 256>>>(table, nbytes, n_items, sink); break;
                case 8: bench<8><<<sm * 8, 256>>>(table, nbytes, n_items, sink); break;
                case 9: bench<9><<<sm * 8, 256>>>(table, nbytes, n_items, sink); break;
                case 10: bench<10><<<sm * 8, 256>>>(table, nbytes, n_items, sink); break;
                case 11: bench<11><<<sm * 8, 256>>>(table, nbytes, n_items, sink); break;
                case 12: bench<12><<<sm * 8, 256>>>(table, nbytes, n_items, sink); break;
                case 13: bench<13><<<sm * 8, 256>>>(table, nbytes, n_items, sink); break;
                default: bench<14><<<sm * 8, 256>>>(table, nbytes, n_items, sink); break;
                }
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                if (rep && ms < best) best = ms;
            }
            printf("%s,%llu,%.4f,%.2f,%.2f\n", names[mode], (unsigned long long)(nbytes >> 20), best,
                   4.0 * n_items / best / 1e6, n_items / best / 1e6);
        }
        cudaFree(table);
    }
    return 0;
}
