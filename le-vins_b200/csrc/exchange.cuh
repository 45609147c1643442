// exchange.cuh -- the path's only exchange step: every rank ends up with the per-feature results of all ranks.
//
// Two ways, both behind the C ABI:
//   * levins_gather        one ncclAllGather of the rank's contiguous result block (17 B / feature) on the context's stream;
//   * peer mode            no separate collective at all: k_plane stores each feature's result straight into slot `rank` of
//                          EVERY rank's gathered block (peer memory mapped through CUDA IPC, stores travel over NVLink /
//                          NVSwitch), k_peer_signal publishes "step s of rank r is complete" to every rank, and k_peer_wait
//                          holds the consumer's stream until all ranks have published.  One process per GPU; the kernels of
//                          different ranks run on different devices, so the wait cannot starve the signal.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace lv {

constexpr int MAX_PEERS = 16;

// where k_plane also puts its results in peer mode
struct PeerOut {
    int n_ranks; // 0: peer mode off
    int rank;
    unsigned long long slot_bytes;  // bytes of one rank's slot: 17 * nq_cap
    unsigned long long nq_cap;      // queries per slot (xyz_depth first, then type, like the local result block)
    unsigned long long parity_off;  // byte offset of the half (step parity) written by this launch
    unsigned char* base[MAX_PEERS]; // gathered block of rank r as mapped in this process: [2][n_ranks][slot_bytes]
};

struct PeerFlags {
    int n_ranks, rank;
    unsigned long long step;
    unsigned long long* flags[MAX_PEERS]; // rank r's flag array [n_ranks] as mapped here; we write flags[r][rank]
};

// after k_plane on the same stream: everything this rank stored for `step` is visible system-wide before the flag is
__global__ void k_peer_signal(const __grid_constant__ PeerFlags f) {
    const int r = threadIdx.x;
    if (r >= f.n_ranks) return;
    __threadfence_system();
    unsigned long long* p = f.flags[r] + f.rank;
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(f.step) : "memory");
}

// holds the stream until every rank has published `step` into OUR flag array; gives up (and says so) after ~8 s so that a
// missing peer cannot hang the GPU
__global__ void k_peer_wait(const unsigned long long* my_flags, int n_ranks, unsigned long long step, int* timed_out) {
    const int r = threadIdx.x;
    if (r >= n_ranks) return;
    const long long t0 = clock64();
    unsigned long long v;
    do {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(my_flags + r) : "memory");
        if (v >= step) return;
        __nanosleep(200);
    } while (clock64() - t0 < 16000000000ll);
    *timed_out = 1;
}

} // namespace lv
