// dsmem_order_probe.cu — microbenchmark for the open question of DESIGN.md §1a-iv / §7-4 (round 3): can a reader in CTA 0 of a
// cluster see the "stamp" of a record before all of the record's words, when the words are plain stores from 32 lanes of another
// CTA into CTA 0's shared memory (DSMEM), followed by [a fence in every lane], __syncwarp() and a st.release.cluster of the stamp
// by lane 0 — the publication protocol of stage A's builders (parse_v7.cuh: build_record7)?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dsmem_order_probe tools/dsmem_order_probe.cu && ./dsmem_order_probe [rounds] [fence]
//
// One cluster of 2 CTAs.  CTA 1, warp 0 is the writer: for sequence number s it stores word[l] = s * 64 + l from every lane l (and a
// second row word[32 + l]), optionally fences in every lane, __syncwarp(), lane 0 releases stamp = s.  CTA 0, warp 0 is the reader:
// it polls the stamp with volatile loads (no acquire, like the decide team), and as soon as it sees s it reads the 64 words; any
// word that is not s * 64 + l is a word that arrived after its stamp.  The reader then acknowledges s (a flag the writer polls
// through DSMEM) so that the writer never runs more than one record ahead.  Prints how many of the `rounds` records were seen torn.
// Development tool; not built by __graft_entry__.build(), not linked into the library.
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

namespace cg = cooperative_groups;

struct Shared {
    unsigned stamp;      // written last by the writer (release)
    unsigned ack;        // written by the reader: the last sequence number it has checked
    unsigned words[64];  // the record
};

__global__ void __cluster_dims__(2, 1, 1) probe(unsigned rounds, int fence_mode, unsigned *torn_records, unsigned *torn_words) {
    __shared__ Shared sh;
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned rank = cluster.block_rank();
    const int lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        sh.stamp = 0;
        sh.ack = 0;
    }
    for (int i = threadIdx.x; i < 64; i += blockDim.x) sh.words[i] = 0;
    __syncthreads();
    cluster.sync();
    Shared *remote = cluster.map_shared_rank(&sh, 0); // CTA 0's copy, seen from either CTA
    if (threadIdx.x < 32) {
        if (rank == 1) { // ---------------- writer
            for (unsigned s = 1; s <= rounds; s++) {
                while (*reinterpret_cast<volatile unsigned *>(&remote->ack) != s - 1) {
                }
                remote->words[lane] = s * 64u + (unsigned)lane;
                remote->words[32 + lane] = s * 64u + 32u + (unsigned)lane;
                if (fence_mode == 1) asm volatile("fence.acq_rel.cluster;" ::: "memory");
                if (fence_mode == 2) __threadfence();
                __syncwarp();
                if (lane == 0) asm volatile("st.release.cluster.u32 [%0], %1;" ::"l"(&remote->stamp), "r"(s) : "memory");
                __syncwarp();
            }
        } else { // ---------------- reader
            unsigned bad_records = 0, bad_words = 0;
            for (unsigned s = 1; s <= rounds; s++) {
                while (*reinterpret_cast<volatile unsigned *>(&sh.stamp) != s) {
                }
                const unsigned w0 = *reinterpret_cast<volatile unsigned *>(&sh.words[lane]);
                const unsigned w1 = *reinterpret_cast<volatile unsigned *>(&sh.words[32 + lane]);
                const unsigned miss = __ballot_sync(0xffffffffu, w0 != s * 64u + (unsigned)lane) | __ballot_sync(0xffffffffu, w1 != s * 64u + 32u + (unsigned)lane);
                if (miss) {
                    bad_records++;
                    bad_words += __popc(miss);
                }
                __syncwarp();
                if (lane == 0) *reinterpret_cast<volatile unsigned *>(&sh.ack) = s;
                __syncwarp();
            }
            if (lane == 0) {
                *torn_records = bad_records;
                *torn_words = bad_words;
            }
        }
    }
    cluster.sync();
}

int main(int argc, char **argv) {
    const unsigned rounds = argc > 1 ? (unsigned)strtoul(argv[1], nullptr, 10) : 2000000u;
    const int fence_mode = argc > 2 ? atoi(argv[2]) : 0; // 0: __syncwarp + release only; 1: fence.acq_rel.cluster in every lane; 2: __threadfence in every lane
    unsigned *d = nullptr, h[2] = {0, 0};
    cudaMalloc(&d, 2 * sizeof(unsigned));
    cudaMemset(d, 0, 2 * sizeof(unsigned));
    probe<<<2, 64>>>(rounds, fence_mode, d, d + 1);
    const cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
    printf("{\"rounds\": %u, \"fence_mode\": %d, \"records_seen_before_all_their_words\": %u, \"words_missing\": %u, \"cuda\": \"%s\"}\n", rounds, fence_mode, h[0], h[1],
           cudaGetErrorString(e));
    return e == cudaSuccess ? 0 : 1;
}
