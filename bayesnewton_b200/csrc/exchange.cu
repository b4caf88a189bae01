// Carry exchange between the time shards of a multi-GPU run over NVLink peer memory.
// The two-level scan needs, between its phases, one O(d^2) carry from every rank on every rank (33 / 21 doubles at
// d = 3).  Instead of an NCCL all-gather, ONE small kernel per exchange does the whole collective on the compute
// stream: every rank stores its carry straight into a slot of every peer's inbox (peer-mapped symmetric memory: the
// stores travel over NVLink / NVSwitch), publishes a sequence number per peer with a system-scope release, spins on
// its own flags until all peers' numbers have arrived, and copies the inbox to the output -- no host round trip, no
// communicator kernel, deterministic layout.  Slots are double-buffered by sequence parity: a rank can only be two
// exchanges ahead of a peer after that peer has launched (hence finished reading) the exchange in between.
#include "common.cuh"

namespace bn {

// per-rank symmetric buffer:  data[2][world][kMaxCarry] doubles, then flags[2][world] (unsigned long long)
constexpr int kMaxCarry = 256;

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// grid = 1 CTA of world warps: warp p talks to peer p
__global__ void carry_exchange_kernel(const unsigned long long* peer_bufs, int world, int rank, const double* carry, int len,
                                      unsigned long long seq, double* out) {
    const int p = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int slot = (int)(seq & 1ULL);
    if (p < world) {
        double* data = reinterpret_cast<double*>(peer_bufs[p]);
        unsigned long long* flags = reinterpret_cast<unsigned long long*>(data + 2 * world * kMaxCarry);
        double* dst = data + ((size_t)slot * world + rank) * kMaxCarry;
        for (int i = lane; i < len; i += 32) dst[i] = carry[i];
        __threadfence_system();
        __syncwarp();
        if (lane == 0) st_release_sys(flags + (size_t)slot * world + rank, seq);
    }
    __syncthreads();
    // wait for every peer's carry of this sequence number in the local inbox
    double* mine = reinterpret_cast<double*>(peer_bufs[rank]);
    const unsigned long long* myflags = reinterpret_cast<const unsigned long long*>(mine + 2 * world * kMaxCarry);
    if (p < world) {
        if (lane == 0)
            while (ld_acquire_sys(myflags + (size_t)slot * world + p) < seq) { }
        __syncwarp();
        const double* src = mine + ((size_t)slot * world + p) * kMaxCarry;
        for (int i = lane; i < len; i += 32) out[(size_t)p * len + i] = __ldcv(src + i);
    }
}

}  // namespace bn

using namespace bn;

extern "C" size_t bn_carry_exchange_bytes(int world) {
    return world < 1 ? 0 : (size_t)2 * world * kMaxCarry * sizeof(double) + (size_t)2 * world * sizeof(unsigned long long);
}

extern "C" int bn_carry_exchange(const uint64_t* peer_buffers_dev, int world, int rank, const double* carry, int len,
                                 uint64_t seq, double* out, void* stream) {
    BN_REQUIRE(peer_buffers_dev && carry && out, "null array");
    BN_REQUIRE(world >= 1 && world <= 32 && rank >= 0 && rank < world, "bad world / rank (%d, %d)", world, rank);
    BN_REQUIRE(len >= 1 && len <= kMaxCarry, "carry length %d outside [1, %d]", len, kMaxCarry);
    BN_REQUIRE(seq >= 1, "sequence numbers start at 1 (the flags start at 0)");
    BN_LAUNCH("carry_exchange", (cudaStream_t)stream,
              carry_exchange_kernel<<<1, 32 * world, 0, (cudaStream_t)stream>>>(
                  (const unsigned long long*)peer_buffers_dev, world, rank, carry, len, (unsigned long long)seq, out));
    BN_CUDA(cudaGetLastError());
    return 0;
}
