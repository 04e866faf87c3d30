// daliti_b200/csrc/dlt_peer.cuh
//
// Sharded map (BASELINE config C4): the per-iteration exchange of the partial normal equations and the
// per-scan exchange of the map_incremental decisions, done by the kernels themselves over NVLink peer memory instead of
// a collective library call between them.
//
// Every rank owns one MAILBOX in its own HBM, mapped into every peer process (CUDA IPC; peer access inside one process).
// A sender stores its message straight into slot [parity][sender] of every receiver's mailbox, fences system-wide and
// then stores the message's sequence number into the slot's flag; a receiver spins on the flags in its OWN memory (the
// point of coherence for peer stores is the owner's L2, so volatile loads see them) and then adds the slots up in RANK
// ORDER -- every rank computes the same bits, so every rank takes the same control decisions of
// laserMapping.cpp:1040, 1069-1101.  Two parities suffice: a rank can only be one exchange ahead of its peers, because
// finishing exchange s needs every peer's message s, which a peer posts after it has finished reading exchange s - 1.
//
// No reference counterpart (the reference is a single-process CPU node); replaces the ncclAllReduce of SURVEY.md 8e on
// this path (dlt_set_shard_reduce / the dlt_iekf_update callback stay as the library-collective alternative).
#pragma once
#include "../../include/daliti_b200.h"
#include "dlt_common.cuh"
#include "dlt_map_kernels.cuh"
#if defined(DLT_EMU)
#include <chrono>
#endif

namespace dlt {

constexpr int kPeerMax = DLT_MAX_PEERS;
constexpr int kPeerEqDoubles = 160;  // >= kNormalEqDoubles (158), whole 128-byte lines
constexpr unsigned long long kPeerTimeoutNs = 5000000000ull;  // a peer that never posts: give up, flag the handle (no hang)

struct PeerBox {  // one rank's mailbox; dec[2][dec_cap] bytes follow
    unsigned long long eq_flag[2][kPeerMax];   // sequence number of the message in eq[parity][sender]
    unsigned long long dec_flag[2][kPeerMax];  // sequence number of the sender's decisions in dec[parity]
    double eq[2][kPeerMax][kPeerEqDoubles];
};
static_assert(sizeof(PeerBox) % 128 == 0, "mailbox header in whole lines");

struct PeerComm {  // device-resident, one per attached handle
    int world, rank;
    int dec_cap;
    int status;                  // sticky: 1 = timed out waiting for a peer
    unsigned long long eq_seq;   // exchanges of normal equations completed so far (device-owned: no-op launches do not exchange)
    unsigned long long dec_seq;  // exchanges of map_incremental decisions completed so far (device-owned: gated launches do not exchange)
    unsigned dec_ticket, pull_ticket;
    PeerBox *box[kPeerMax];      // every rank's mailbox as mapped in this process; box[rank] is this rank's own
};

DLT_D unsigned long long peer_now_ns() {
#if defined(DLT_EMU)
    return (unsigned long long)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
#else
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
#endif
}
DLT_D unsigned char *peer_dec(PeerBox *b, int parity, int dec_cap) {
    return reinterpret_cast<unsigned char *>(b + 1) + (size_t)parity * (size_t)dec_cap;
}
// spin until *flag >= seq; false on timeout
DLT_D bool peer_wait(const unsigned long long *flag, unsigned long long seq) {
    const volatile unsigned long long *f = flag;
    if (*f >= seq) return true;
    const unsigned long long t0 = peer_now_ns();
    for (;;) {
        for (int k = 0; k < 64; k++)
            if (*f >= seq) return true;
        if (peer_now_ns() - t0 > kPeerTimeoutNs) return false;
    }
}

// Sum of R[0 .. n) over the ranks, in place, by one whole block (blockDim.x >= kPeerMax, n <= kPeerEqDoubles).  R holds this
// rank's partial sums, written by this block.  false (block-uniform) = a peer never posted: R is not a sum, pc->status is set.
DLT_D bool peer_allreduce_block(PeerComm *pc, double *R, int n) {
    __shared__ unsigned long long s_seq;
    __shared__ int s_timeout;
    const int tid = threadIdx.x, W = pc->world, me = pc->rank;
    if (tid == 0) {
        s_seq = pc->eq_seq + 1ull;
        s_timeout = 0;
    }
    __syncthreads();  // (also: the block's own stores to R are visible to all its threads)
    const unsigned long long seq = s_seq;
    const int par = (int)(seq & 1ull);
    // ---- push my partial sums into slot [par][me] of every peer's mailbox
    for (int k = tid; k < W * n; k += blockDim.x) {
        const int p = k / n, j = k - p * n;
        if (p != me) pc->box[p]->eq[par][me][j] = R[j];
    }
    __threadfence_system();
    __syncthreads();
    if (tid < W && tid != me) *(volatile unsigned long long *)&pc->box[tid]->eq_flag[par][me] = seq;  // my message is complete at peer `tid`
    __syncthreads();  // every flag is posted before anybody waits
    if (tid < W && tid != me && !peer_wait(&pc->box[me]->eq_flag[par][tid], seq)) s_timeout = 1;  // peer `tid`'s message is complete here
    __threadfence_system();
    __syncthreads();
    // ---- fixed rank order: bit-identical on every rank
    if (tid < n) {
        double acc = 0.0;
        for (int r = 0; r < W; r++) acc += (r == me) ? R[tid] : *(const volatile double *)&pc->box[me]->eq[par][r][tid];
        R[tid] = acc;
    }
    if (tid == 0) {
        pc->eq_seq = seq;
        if (s_timeout) pc->status = 1;
    }
    __syncthreads();
    return s_timeout == 0;
}

// map_incremental on a sharded map: the owner of a query point decides (laserMapping.cpp:588-625 needs the point's
// neighbours, which only the owner holds) and stores the decision code (0 drop, 1 PointToAdd, 2 PointNoNeedDownsample)
// into dec[parity][i] of EVERY rank's mailbox, its own included; the last block to finish publishes the sequence number.
// gate_go / gate_n (optional): behind the device-resident loop the exchange only runs when the loop armed the insert
// (*gate_go == 1, the same on every rank), with feats_down_size read from the device.
__global__ void k_incr_push(PeerComm *pc, const unsigned char *__restrict__ ds_flag, const unsigned char *__restrict__ add_flag,
                            const unsigned char *__restrict__ flags, int n, unsigned char foreign_bit, const int *gate_go, const int *gate_n) {
    DLT_PDL_WAIT();
    __shared__ int s_last;
    if (gate_go && *gate_go != 1) return;  // block-uniform
    if (gate_n) n = *gate_n;
    const unsigned long long seq = pc->dec_seq + 1ull;  // (only the LAST block of k_incr_pull advances dec_seq, in stream order behind this kernel)
    const int W = pc->world, me = pc->rank, par = (int)(seq & 1ull);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && !(flags[i] & foreign_bit)) {
        const unsigned char code = ds_flag[i] ? 1 : (add_flag[i] ? 2 : 0);
        for (int p = 0; p < W; p++) peer_dec(pc->box[p], par, pc->dec_cap)[i] = code;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(&pc->dec_ticket, 1u) == gridDim.x - 1u) ? 1 : 0;
    __syncthreads();
    if (!s_last) return;
    __threadfence_system();
    if ((int)threadIdx.x < W) *(volatile unsigned long long *)&pc->box[threadIdx.x]->dec_flag[par][me] = seq;
    if (threadIdx.x == 0) pc->dec_ticket = 0u;
}
// ... and every rank, once all owners have published, reads all n decisions from its own mailbox.
// fi.on: the first two insert phases ride along (k_map_claim with the shard filter + k_ds_bid for the point just decided: neither
// needs another thread's result), which saves two launches per scan.
__global__ void k_incr_pull(PeerComm *pc, int n, unsigned char *__restrict__ ds_flag, unsigned char *__restrict__ add_flag,
                            int *__restrict__ class_counts, const int *gate_go, const int *gate_n, MapView m, const float4 *__restrict__ pts,
                            FuseInsert fi) {
    DLT_PDL_WAIT();
    __shared__ int s_timeout;
    if (gate_go && *gate_go != 1) return;  // block-uniform
    if (gate_n) n = *gate_n;
    const unsigned long long seq = pc->dec_seq + 1ull;  // read before this block's ticket: the last ticket holder writes it
    const int W = pc->world, me = pc->rank, par = (int)(seq & 1ull);
    if (threadIdx.x == 0) s_timeout = 0;
    __syncthreads();
    if ((int)threadIdx.x < W && !peer_wait(&pc->box[me]->dec_flag[par][threadIdx.x], seq)) s_timeout = 1;
    __threadfence_system();
    __syncthreads();
    if (s_timeout && threadIdx.x == 0) pc->status = 1;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    int code = 0;
    if (i < n) {
        code = *(const volatile unsigned char *)&peer_dec(pc->box[me], par, pc->dec_cap)[i];
        ds_flag[i] = code == 1 ? 1 : 0;
        add_flag[i] = code == 2 ? 1 : 0;
        if (fi.on) {
            int cs = -1, vs = -1;
            const float4 p = pts[i];
            if (code != 0) {
                int cx, cy, cz;
                cell_of_point(m, p.x, p.y, p.z, cx, cy, cz);
                if (shard_keeps_cell(m, cx, cy, cz, kShardHalo)) cs = map_claim(m, pack_key(cx, cy, cz));
            }
            if (code == 1) vs = ds_bid_point(m, fi.sc, p, i);
            fi.cell_slot[i] = cs;
            fi.vslot[i] = vs;
        }
    }
    unsigned bd = __ballot_sync(0xffffffffu, code == 1), ba = __ballot_sync(0xffffffffu, code == 2);
    if ((threadIdx.x & 31) == 0) {
        if (bd) atomicAdd(&class_counts[0], __popc(bd));
        if (ba) atomicAdd(&class_counts[1], __popc(ba));
    }
    __syncthreads();
    if (threadIdx.x == 0 && atomicAdd(&pc->pull_ticket, 1u) == gridDim.x - 1u) {  // every block has read dec_seq by now
        pc->pull_ticket = 0u;
        pc->dec_seq = seq;
    }
}

}  // namespace dlt
