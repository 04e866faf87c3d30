// daliti_b200/csrc/dlt_loop_kernels.cuh
//
// The IEKF iteration loop of eskf_lio (eskf_lio/src/laserMapping.cpp:820-1102) as ONE kernel per iteration, or as one
// persistent cooperative kernel for the whole loop.
//
// The classic device path spends three launches per iteration (k_knn8, k_knn, k_residual) and pays a kernel boundary
// between the match pass and the residual pass although point i's residual only needs point i's neighbours.  Here the
// downsampled scan is cut into chunks of C points (C a multiple of 32, at most kLoopChunkMax, chosen so that one chunk per
// block covers the scan) and a block does everything for its chunk back to back:
//     match pass    8 lanes / query over the 3^3 block (knn8_group); the queries that cannot prove exact go to a
//                   block-local list and are searched warp-per-query (knn_warp_query) right away  (:835-854)
//     residual pass plane fit / residual / gates / Jacobian row per point (residual_point), warp-shuffle + shared
//                   memory reduction to one partial per block                                       (:863-979, :1015)
// and the last block to finish sums the partials in fixed order, runs the 24-state solve / control step
// (iekf_step_block: :899-918, :1012-1101, and the zeta blend :1105-1131 at loop exit) for everybody.
//
//   k_iekf_iter   one launch = one iteration; later launches of a scan return at once when the loop has ended
//   k_iekf_loop   cooperative launch, grid = what is co-resident: all iterations in one launch with one grid-wide
//                 barrier per iteration; the 6 x 6 eigen-decomposition of the degeneracy output runs on an otherwise
//                 idle block WHILE the last block solves, so it costs no time of its own and no launch
//
// Results equal the classic kernels' up to the fp64 summation order of the normal equations (same per-point code).
#pragma once
#include "dlt_measure_kernels.cuh"

namespace dlt {

constexpr int kLoopBlock = 256;
constexpr int kLoopChunkMax = 256;
#ifndef DLT_LOOP_MINBLOCKS
#define DLT_LOOP_MINBLOCKS 3
#endif
static_assert(kLoopBlock == kResidBlock, "residual_final_reduce and iekf_step_block run on the loop kernels' block size");

struct GridBar {       // grid-wide barrier state of one handle (device memory, zeroed once)
    unsigned count;    // arrivals of the current barrier
    unsigned gen;      // completed barriers
    unsigned r_seq;    // normal equations published (iteration count, scan-local + base): the eigen block waits on it
    unsigned pad;
};

struct LoopK {  // everything a loop kernel needs, by value
    MapView m;
    const float4 *down;
    KnnOut knn;
    MeasureBufs mb;
    IekfDev *dev;
    const int *n_ptr;    // feats_down_size (ScanScalars::n_down)
    const int *vox_ptr;  // VoxelGrid status of the scan
    float max_sq_dist, plane_thr;
    PeerComm *peer;
    GridBar *bar;
    double *eig_out;     // eigvals[6], eigvecs[36] of the last iteration's H^T H [0:6, 0:6]
    int max_iter;
};

struct LoopSmem {
    Pose P;
    // the two halves of a match pass run one after the other (a block barrier in between): their scratch shares the space
    union {
        struct {  // knn_warp_query
            float4 cand[kKnnWarps][kCandSlots];
            int cid[kKnnWarps][kCandSlots];
            int wl[kKnnWarps][kWlMax];
            Cand bestw[kKnnWarps][kK];
            int nb[kKnnWarps];
        };
        int wl8[kLoopBlock / 32][kKnn8WlInts];  // knn8_group: bucket index of every visit
    };
    int unres[kLoopChunkMax];
    int n_unres;
    int last;
    double part[kLoopBlock / 32][NormalEq<true>::NR];
};

DLT_D int loop_chunk_size(int n, int n_blocks) {  // smallest multiple of 32 with one chunk per block, capped
    int c = ((n + n_blocks - 1) / n_blocks + 31) & ~31;
    if (c < 32) c = 32;
    return c > kLoopChunkMax ? kLoopChunkMax : c;
}

// instrumentation: the solver's thread 0 leaves globaltimer stamps (ns) of the iteration's stages in dev->clocks[it][9..14]
#if defined(DLT_EMU)
#define DLT_GSTAMP(dev, it, slot)
DLT_D void spin_pause() {}
#else
#define DLT_GSTAMP(dev, it, slot)                                              \
    do {                                                                       \
        if (threadIdx.x == 0) (dev)->clocks[it][slot] = (long long)peer_now_ns(); \
    } while (0)
DLT_D void spin_pause() { __nanosleep(40); }
#endif

// grid-wide barrier over co-resident blocks (cooperative launch): sense by generation count
DLT_D void grid_barrier(GridBar *b, unsigned n_blocks) {
    __syncthreads();
    if (threadIdx.x == 0) {
        volatile unsigned *gen = &b->gen;
        const unsigned g = *gen;
        __threadfence();
        if (atomicAdd(&b->count, 1u) == n_blocks - 1u) {
            b->count = 0u;
            __threadfence();
            *gen = g + 1u;
        } else {
            while (*gen == g) spin_pause();
        }
        __threadfence();
    }
    __syncthreads();
}

// One iteration's per-chunk work of this block, reduced to one partial per block.  Ends with the block's arrival on the
// iteration's ticket; returns true in the block that arrived last (whole block).
template <bool EXT>
DLT_D bool loop_iter_chunks(const LoopK &a, LoopSmem &sm, int n, int do_match) {
    using NE = NormalEq<EXT>;
    constexpr int D = NE::D, NR = NE::NR;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = (int)gridDim.x;
    const int C = loop_chunk_size(n, G);
    const int n_chunks = (n + C - 1) / C;
    const unsigned n_parts = (unsigned)(n_chunks < G ? n_chunks : G);  // blocks that own at least one chunk
    if (lane == 0)
        for (int k = 0; k < NR; k++) sm.part[warp][k] = 0.0;
    for (int c = blockIdx.x; c < n_chunks; c += G) {  // block-uniform
        const int c0 = c * C, c1 = min(n, c0 + C);
        if (do_match) {
            if (tid == 0) sm.n_unres = 0;
            __syncthreads();
            for (int q0 = c0 + warp * 4; q0 < c1; q0 += (kLoopBlock / 32) * 4)  // warp-uniform
                knn8_group(a.m, a.down, c1, 1, sm.P, a.max_sq_dist, a.knn, sm.unres, &sm.n_unres, q0, lane, sm.wl8[warp], kLoopChunkMax);
            __syncthreads();
            const int nu = min(sm.n_unres, kLoopChunkMax);
            if (warp < kKnnWarps)
                for (int u = warp; u < nu; u += kKnnWarps) {  // warp-uniform
                    knn_warp_query(a.m, a.down, sm.unres[u], 1, sm.P, a.max_sq_dist, a.knn, sm.cand[warp], sm.cid[warp], sm.wl[warp], sm.bestw[warp],
                                   &sm.nb[warp], lane);
                    __syncwarp();
                }
            __syncthreads();  // this chunk's neighbour sets and flags are complete (and visible to the block)
        }
        for (int i0 = c0 + warp * 32; i0 < c1; i0 += kLoopBlock) {  // warp-uniform
            const int i = i0 + lane;
            double row[D];
            double meas = 0.0, absr = 0.0;
#pragma unroll
            for (int d = 0; d < D; d++) row[d] = 0.0;
            bool effective = false;
            if (i < c1) effective = residual_point<EXT>(a.mb, i, do_match, sm.P, a.plane_thr, row, meas, absr);
            residual_warp_reduce<EXT, true>(row, meas, absr, effective, sm.part[warp], lane);
        }
    }
    __syncthreads();
    if ((unsigned)blockIdx.x < n_parts && tid < NR) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < kLoopBlock / 32; w++) v += sm.part[w][tid];
        a.mb.partials[(size_t)blockIdx.x * NR + tid] = v;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const unsigned t = atomicAdd(a.mb.ticket, 1u);
        sm.last = (t == (unsigned)G - 1u) ? 1 : 0;
    }
    __syncthreads();
    return sm.last != 0;
}

// What the last block of an iteration does: normal equations -> (sum over the ranks of a sharded map) -> solve / control.
// Returns false when a peer never posted (the loop is ended, the host reports it).
template <bool EXT>
DLT_D bool loop_iter_finish_reduce(const LoopK &a, int n, const LoopArgs &la) {
    const int G = (int)gridDim.x;
    const int C = loop_chunk_size(n, G);
    const int n_chunks = (n + C - 1) / C;
    const unsigned n_parts = (unsigned)(n_chunks < G ? n_chunks : G);
    residual_final_reduce<EXT>(a.mb, n_parts, n, la);
    if (a.peer && !peer_allreduce_block(a.peer, a.mb.result, kNormalEqDoubles)) {  // block-uniform; starts and ends with a barrier
        if (threadIdx.x == 0) a.dev->b.done = 1;
        return false;
    }
    __syncthreads();  // R complete and visible to the block
    return true;
}

DLT_D void loop_load_pose(const LoopK &a, LoopSmem &sm) {
    if (threadIdx.x < 24) reinterpret_cast<double *>(&sm.P)[threadIdx.x] = a.dev->b.state[threadIdx.x];
    __syncthreads();
}

// ------------------------------------------------------------------ one launch = one iteration
template <bool EXT>
__global__ void __launch_bounds__(kLoopBlock, DLT_LOOP_MINBLOCKS) k_iekf_iter(LoopK a) {
    DLT_PDL_WAIT();
    __shared__ LoopSmem sm;
    const dlt_iekf_block &c = a.dev->b;
    if (c.done) return;  // block-uniform: the loop ended in an earlier launch of this scan
    const int do_match = (c.iter == 0 || c.rematch_en) ? 1 : 0;
    const int n = *a.n_ptr;
    loop_load_pose(a, sm);
    if (!loop_iter_chunks<EXT>(a, sm, n, do_match)) return;
    __threadfence();
    LoopArgs la = {a.dev, a.n_ptr, a.vox_ptr, 1, 0, 0ull, 0ull, a.peer};
    if (!loop_iter_finish_reduce<EXT>(a, n, la)) return;
    iekf_step_block(a.dev, a.mb.result, a.n_ptr, a.vox_ptr, NormalEq<EXT>::D);
    if (threadIdx.x == 0 && !a.dev->b.done && a.dev->b.rematch_en) {  // the next launch matches again: re-arm its counters
        *a.knn.far_count = 0;
        a.knn.nn_count[0] = 0;
        a.knn.nn_count[1] = 0;
    }
}

#if !defined(DLT_EMU)
// ------------------------------------------------------------------ the whole loop in one cooperative launch
// Inside one launch L1 is never invalidated, so everything one block writes and ANOTHER block reads goes through L2
// (__ldcg / volatile); to keep that list short the solve always runs on the same block (the last one of the grid, which
// owns a chunk only for very large scans and otherwise just waits for the others' partials), so the IEKF block is read and
// written by one SM only, and the eigen-decomposition runs next to it on the second to last block.
template <bool EXT>
__global__ void __launch_bounds__(kLoopBlock, DLT_LOOP_MINBLOCKS) k_iekf_loop(LoopK a) {
    DLT_PDL_WAIT();
    __shared__ LoopSmem sm;
    const unsigned G = gridDim.x;
    const int n = *a.n_ptr;
    const bool solver = blockIdx.x == G - 1u;
    const bool eig_block = blockIdx.x == (G > 1u ? G - 2u : 0u);
    LoopArgs la = {a.dev, a.n_ptr, a.vox_ptr, 1, 0, 0ull, 0ull, a.peer};
    for (int it = 0; it < a.max_iter; it++) {
        // control state of this iteration: written by the solver before the barrier everybody just left
        if (*(volatile const int *)&a.dev->b.done) break;  // grid-uniform
        const int do_match = (it == 0 || *(volatile const int *)&a.dev->b.rematch_en) ? 1 : 0;
        if (threadIdx.x < 24) reinterpret_cast<double *>(&sm.P)[threadIdx.x] = __ldcg(&a.dev->b.state[threadIdx.x]);
        __syncthreads();
        if (solver) DLT_GSTAMP(a.dev, it, 9);
        loop_iter_chunks<EXT>(a, sm, n, do_match);
        if (solver) {
            DLT_GSTAMP(a.dev, it, 10);
            if (threadIdx.x == 0) {
                while (*(volatile unsigned *)a.mb.ticket != G) spin_pause();  // every block's partial is in place
                __threadfence();
            }
            __syncthreads();
            DLT_GSTAMP(a.dev, it, 11);
            const bool ok = loop_iter_finish_reduce<EXT>(a, n, la);
            if (threadIdx.x == 0) {
                __threadfence();
                *(volatile unsigned *)&a.bar->r_seq = (unsigned)it + 1u;  // the normal equations are in a.mb.result
            }
            DLT_GSTAMP(a.dev, it, 12);
            if (ok) {
                iekf_step_block(a.dev, a.mb.result, a.n_ptr, a.vox_ptr, NormalEq<EXT>::D);
                if (threadIdx.x == 0 && !a.dev->b.done && a.dev->b.rematch_en) {  // the next iteration matches again: re-arm its counters
                    *a.knn.far_count = 0;
                    a.knn.nn_count[0] = 0;
                    a.knn.nn_count[1] = 0;
                }
            }
            DLT_GSTAMP(a.dev, it, 13);
        }
        if (eig_block) {  // eigen-decomposition of this iteration's pose block while the solver solves (behind it when they coincide)
            if (threadIdx.x == 0) {
                while (*(volatile unsigned *)&a.bar->r_seq != (unsigned)it + 1u) spin_pause();
                __threadfence();
            }
            __syncthreads();
            if (threadIdx.x < 32) eigen6_warp(a.mb.result, a.eig_out, threadIdx.x);
        }
        grid_barrier(a.bar, G);
        if (solver) DLT_GSTAMP(a.dev, it, 14);
    }
    if (solver && threadIdx.x == 0) a.bar->r_seq = 0u;  // (nobody reads it past the last barrier)
}
#endif

}  // namespace dlt
