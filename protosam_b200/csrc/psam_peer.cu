// One-sided exchanges over NVLink peer memory: the prototype "broadcast" and the record "gather" of the sharded volume
// path (the reference is single-GPU and has neither) without a collective library on the path.
//
// Every rank owns one REGION per channel, a symmetric allocation mapped into every peer (the host side gets the peer
// addresses from torch.distributed._symmetric_memory, include/psam_b200.h says how):
//     [ 4 KB of signal words | payload ("mailbox") ]
// A sender copies its data straight into the RECEIVER's mailbox with ordinary 16-byte stores (they travel over
// NVLink / NVSwitch), fences, and then raises a signal word in the receiver's region; the receiver's kernel spins on its
// own, local signal word, copies the mailbox into private memory and acknowledges by raising a word in the sender's
// region, which is what the sender's NEXT transfer waits for before it overwrites the mailbox.  All counters are epochs
// kept in device memory, so the kernels can be captured into CUDA graphs and replayed.
//
// Why not NCCL for these: the messages are small (3.8 MB of live prototype rows, 82 KB of compact records per rank) and
// the GPU is full of persistent compute CTAs; an NCCL kernel per volume and direction costs SM slots while it waits for
// its peers and carries the table's padding (8.0 MB), while these kernels move only live rows, need no channel / proxy
// machinery and their waits are one thread per CTA.  The NCCL path stays (engine.broadcast_prototypes / gather_packed):
// it is what runs over gloo in the CPU tests and when symmetric memory is unavailable.
#include "psam_common.cuh"

namespace psam {

constexpr int PEER_THREADS = 256;
constexpr int PEER_MAX_WORLD = 64;
constexpr unsigned long long PEER_TIMEOUT_NS = 120ull * 1000 * 1000 * 1000;    // a lost peer traps instead of hanging the GPU

// signal words (uint32) of a region: READY[r] raised by writer rank r, ACK[r] raised by reader rank r
__device__ __forceinline__ uint32_t* sig_ready(void* region, int r) { return reinterpret_cast<uint32_t*>(region) + r; }
__device__ __forceinline__ uint32_t* sig_ack(void* region, int r) { return reinterpret_cast<uint32_t*>(region) + 256 + r; }
__device__ __forceinline__ uint8_t* payload(void* region) { return reinterpret_cast<uint8_t*>(region) + PSAM_PEER_SIGNAL_BYTES; }

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p)
{
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// epochs compare modulo 2^32
__device__ __forceinline__ void wait_at_least(const uint32_t* p, uint32_t e)
{
    const unsigned long long t0 = trace_now();
    while ((int32_t)(ld_acquire_sys(p) - e) < 0) {
        __nanosleep(200);
        if (trace_now() - t0 > PEER_TIMEOUT_NS) __trap();
    }
}

struct PeerCtr {            // per channel, local device memory
    uint32_t epoch;         // exchanges completed on this rank
    uint32_t done;          // CTAs of the running kernel that finished their copy
};

// 16-byte grid-stride copy.  Loads bypass L1 (ld.global.cg): a mailbox is written by another GPU, and an L1 line left by
// the previous epoch's read of the same address must not be served.
__device__ __forceinline__ void copy16(uint8_t* dst, const uint8_t* src, size_t nbytes, size_t tid, size_t nthreads)
{
    const uint4* s = reinterpret_cast<const uint4*>(src);
    uint4* d = reinterpret_cast<uint4*>(dst);
    for (size_t i = tid; i < nbytes / 16; i += nthreads) d[i] = __ldcg(s + i);
}

// the last CTA to finish runs `fn` (after every CTA's stores were fenced) and closes the epoch
template <typename F>
__device__ __forceinline__ void finish(PeerCtr* ctr, uint32_t e, F fn)
{
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        if (atomicAdd(&ctr->done, 1u) == gridDim.x - 1) {
            __threadfence_system();
            fn();
            ctr->done = 0;
            ctr->epoch = e;
            __threadfence();
        }
    }
}

struct TableGeom {
    int nsets, cap_rows, C;
    size_t ints_offset, ints_bytes;       // counts | eff_modes | status behind the rows
};

// live rows of every set + the integer arrays, src -> dst (same offsets on both sides)
__device__ __forceinline__ void copy_table(uint8_t* dst, const uint8_t* src, const TableGeom& g, const int32_t* counts)
{
    const size_t row_bytes = (size_t)g.C * 4;
    for (int s = 0; s < g.nsets; ++s) {
        const int n = min(max(counts[s], 0), g.cap_rows);
        const size_t off = (size_t)s * g.cap_rows * row_bytes;
        copy16(dst + off, src + off, (size_t)n * row_bytes, (size_t)blockIdx.x * blockDim.x + threadIdx.x,
               (size_t)gridDim.x * blockDim.x);
    }
    copy16(dst + g.ints_offset, src + g.ints_offset, g.ints_bytes, (size_t)blockIdx.x * blockDim.x + threadIdx.x,
           (size_t)gridDim.x * blockDim.x);
}

// The push / receive of a table share ONE epoch counter per channel (every rank takes part in an exchange exactly once,
// on either side), put and collect have one each (the collecting rank runs both).
// src rank: wait until every peer acknowledged the previous table, write the live rows into every peer's mailbox, raise
// READY[rank] there
__global__ void __launch_bounds__(PEER_THREADS) k_peer_push_table(const uint8_t* __restrict__ table, TableGeom g,
                                                                   void* const* __restrict__ regions, int world, int rank,
                                                                   PeerCtr* ctr)
{
    const uint32_t e = ctr->epoch + 1;
    if (threadIdx.x < world && threadIdx.x != rank) wait_at_least(sig_ack(regions[rank], threadIdx.x), e - 1);
    __syncthreads();
    const int32_t* counts = reinterpret_cast<const int32_t*>(table + g.ints_offset);
    for (int p = 0; p < world; ++p)
        if (p != rank) copy_table(payload(regions[p]), table, g, counts);
    // the source also raises its own ACK word everywhere: it has no mailbox to drain in this exchange, and a DIFFERENT
    // source of the next exchange waits for the acknowledgement of every other rank, this one included
    finish(ctr, e, [&] {
        for (int p = 0; p < world; ++p)
            if (p != rank) {
                st_release_sys(sig_ack(regions[p], rank), e);
                st_release_sys(sig_ready(regions[p], rank), e);
            }
    });
}

// other ranks: wait for READY[src] in the own region, copy the mailbox into the private table, raise ACK[rank] in EVERY
// region (the next table may come from a different source rank, which then needs everybody's acknowledgement)
__global__ void __launch_bounds__(PEER_THREADS) k_peer_recv_table(uint8_t* __restrict__ table, TableGeom g,
                                                                   void* const* __restrict__ regions, int world, int rank,
                                                                   int src, PeerCtr* ctr)
{
    const uint32_t e = ctr->epoch + 1;
    if (threadIdx.x == 0) wait_at_least(sig_ready(regions[rank], src), e);
    __syncthreads();
    const uint8_t* box = payload(regions[rank]);
    // the counts arrive with the table: read them from the mailbox (volatile: written by the peer, not by a kernel of
    // this device)
    __shared__ int32_t s_counts[1024];
    const volatile int32_t* vc = reinterpret_cast<const volatile int32_t*>(box + g.ints_offset);
    for (int i = threadIdx.x; i < g.nsets; i += blockDim.x) s_counts[i] = vc[i];
    __syncthreads();
    copy_table(table, box, g, s_counts);
    finish(ctr, e, [&] {
        for (int p = 0; p < world; ++p)
            if (p != rank) st_release_sys(sig_ack(regions[p], rank), e);
    });
}

// every rank: wait until dst acknowledged the previous records, write `nbytes` into slot `rank` of dst's mailbox, raise
// READY[rank] there
__global__ void __launch_bounds__(PEER_THREADS) k_peer_put(const uint8_t* __restrict__ src, size_t nbytes, size_t slot_bytes,
                                                            void* const* __restrict__ regions, int rank, int dst, PeerCtr* ctr)
{
    const uint32_t e = ctr->epoch + 1;
    if (threadIdx.x == 0) wait_at_least(sig_ack(regions[rank], 0), e - 1);     // one ACK word: whoever collected last
    __syncthreads();
    copy16(payload(regions[dst]) + (size_t)rank * slot_bytes, src, nbytes, (size_t)blockIdx.x * blockDim.x + threadIdx.x,
           (size_t)gridDim.x * blockDim.x);
    finish(ctr, e, [&] { st_release_sys(sig_ready(regions[dst], rank), e); });
}

// dst: wait for READY[r] of every rank, copy the mailbox (world slots) into private memory, raise the ACK word everywhere.
// The exchange number is the one this rank's own put (earlier on the same stream) just closed: the collecting rank may
// change between exchanges, so it cannot count by itself.
__global__ void __launch_bounds__(PEER_THREADS) k_peer_collect(uint8_t* __restrict__ out, size_t slot_bytes,
                                                                void* const* __restrict__ regions, int world, int rank,
                                                                const PeerCtr* put_ctr, PeerCtr* ctr)
{
    const uint32_t e = put_ctr->epoch;
    if (threadIdx.x < world) wait_at_least(sig_ready(regions[rank], threadIdx.x), e);
    __syncthreads();
    copy16(out, payload(regions[rank]), slot_bytes * world, (size_t)blockIdx.x * blockDim.x + threadIdx.x,
           (size_t)gridDim.x * blockDim.x);
    finish(ctr, e, [&] {
        for (int p = 0; p < world; ++p) st_release_sys(sig_ack(regions[p], 0), e);
    });
}

static int grid_for(size_t nbytes)
{
    const size_t per_cta = (size_t)PEER_THREADS * 16 * 8;       // ~8 stores per thread
    return (int)std::max<size_t>(1, std::min<size_t>(32, (nbytes + per_cta - 1) / per_cta));
}

}  // namespace psam

using namespace psam;

PSAM_TRACE_TU();

#define PEER_COMMON_CHECKS(fn)                                                                                             \
    PSAM_CHECK_ARG(regions && ctr, fn ": null pointer");                                                                   \
    PSAM_CHECK_ARG(world >= 2 && world <= PEER_MAX_WORLD && rank >= 0 && rank < world, fn ": bad world / rank (%d, %d)", world, rank)

extern "C" size_t psam_peer_region_bytes(size_t payload_bytes) { return PSAM_PEER_SIGNAL_BYTES + align_up(payload_bytes, 256); }

static int table_geom(const char* fn, int nsets, int cap_rows, int C, size_t ints_offset, size_t ints_bytes, TableGeom* g)
{
    if (!(nsets >= 1 && nsets <= 1024 && cap_rows >= 1 && C >= 4 && C % 4 == 0 && ints_offset % 16 == 0 && ints_bytes % 16 == 0 &&
          ints_offset >= (size_t)nsets * cap_rows * C * 4 && ints_bytes >= (size_t)nsets * 4)) {
        set_error("%s: bad table geometry (nsets=%d cap_rows=%d C=%d ints at %zu + %zu)", fn, nsets, cap_rows, C, ints_offset, ints_bytes);
        return PSAM_ERR_ARG;
    }
    *g = TableGeom{nsets, cap_rows, C, ints_offset, ints_bytes};
    return PSAM_OK;
}

extern "C" int psam_peer_push_table(const void* table, int nsets, int cap_rows, int C, size_t ints_offset, size_t ints_bytes,
                                    void* const* regions, int world, int rank, void* ctr, psam_stream_t stream_)
{
    PSAM_TRACE("psam_peer_push_table");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    PEER_COMMON_CHECKS("psam_peer_push_table");
    PSAM_CHECK_ARG(table && (reinterpret_cast<uintptr_t>(table) & 15) == 0, "psam_peer_push_table: table must be 16-byte aligned");
    TableGeom g;
    if (int rc = table_geom("psam_peer_push_table", nsets, cap_rows, C, ints_offset, ints_bytes, &g)) return rc;
    PSAM_PROF_BEGIN(stream);
    k_peer_push_table<<<grid_for(ints_offset / 2), PEER_THREADS, 0, stream>>>(static_cast<const uint8_t*>(table), g, regions, world, rank,
                                                                             static_cast<PeerCtr*>(ctr));
    PSAM_CHECK_LAUNCH("k_peer_push_table");
    return PSAM_OK;
}

extern "C" int psam_peer_recv_table(void* table, int nsets, int cap_rows, int C, size_t ints_offset, size_t ints_bytes,
                                    void* const* regions, int world, int rank, int src, void* ctr, psam_stream_t stream_)
{
    PSAM_TRACE("psam_peer_recv_table");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    PEER_COMMON_CHECKS("psam_peer_recv_table");
    PSAM_CHECK_ARG(src >= 0 && src < world && src != rank, "psam_peer_recv_table: src %d", src);
    PSAM_CHECK_ARG(table && (reinterpret_cast<uintptr_t>(table) & 15) == 0, "psam_peer_recv_table: table must be 16-byte aligned");
    TableGeom g;
    if (int rc = table_geom("psam_peer_recv_table", nsets, cap_rows, C, ints_offset, ints_bytes, &g)) return rc;
    PSAM_PROF_BEGIN(stream);
    k_peer_recv_table<<<grid_for(ints_offset / 2), PEER_THREADS, 0, stream>>>(static_cast<uint8_t*>(table), g, regions, world, rank, src,
                                                                             static_cast<PeerCtr*>(ctr));
    PSAM_CHECK_LAUNCH("k_peer_recv_table");
    return PSAM_OK;
}

extern "C" int psam_peer_put(const void* src, size_t nbytes, size_t slot_bytes, void* const* regions, int world, int rank, int dst,
                             void* ctr, psam_stream_t stream_)
{
    PSAM_TRACE("psam_peer_put");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    PEER_COMMON_CHECKS("psam_peer_put");
    PSAM_CHECK_ARG(src && dst >= 0 && dst < world && nbytes % 16 == 0 && slot_bytes % 16 == 0 && nbytes <= slot_bytes &&
                       (reinterpret_cast<uintptr_t>(src) & 15) == 0,
                   "psam_peer_put: bad arguments (16-byte multiples, nbytes <= slot_bytes)");
    PSAM_PROF_BEGIN(stream);
    k_peer_put<<<grid_for(nbytes), PEER_THREADS, 0, stream>>>(static_cast<const uint8_t*>(src), nbytes, slot_bytes, regions, rank, dst,
                                                             static_cast<PeerCtr*>(ctr));
    PSAM_CHECK_LAUNCH("k_peer_put");
    return PSAM_OK;
}

extern "C" int psam_peer_collect(void* out, size_t slot_bytes, void* const* regions, int world, int rank, const void* put_ctr,
                                 void* ctr, psam_stream_t stream_)
{
    PSAM_TRACE("psam_peer_collect");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    PEER_COMMON_CHECKS("psam_peer_collect");
    PSAM_CHECK_ARG(out && put_ctr && slot_bytes % 16 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0, "psam_peer_collect: bad arguments");
    PSAM_PROF_BEGIN(stream);
    k_peer_collect<<<grid_for(slot_bytes * world), PEER_THREADS, 0, stream>>>(static_cast<uint8_t*>(out), slot_bytes, regions, world, rank,
                                                                             static_cast<const PeerCtr*>(put_ctr), static_cast<PeerCtr*>(ctr));
    PSAM_CHECK_LAUNCH("k_peer_collect");
    return PSAM_OK;
}
