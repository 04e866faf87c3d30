// daliti_b200/csrc/dlt_measure_kernels.cuh
//
// The IEKF measurement model of eskf_lio (eskf_lio/src/laserMapping.cpp:820-979) on the
// device, for one iteration at a given pose:
//   k_knn      body->world transform (:835-840) + exact k=5 nearest neighbours in the
//              voxel-hash map (replaces KD_TREE::Nearest_Search, ikd_Tree.cpp:425-461,
//              1061-1244) + the match gate (:852-854).  One warp per query point; the
//              buckets of the 3x3x3 / 5x5x5 / 7x7x7 cell rings are staged through shared
//              memory and the 5 best are selected with warp shuffles.
//   k_residual plane fit (esti_plane, common_lib.h:267-299; cached between matches, F7),
//              point-to-plane residual and gates (:866-880, :889), the 1x12 Jacobian row
//              (:948-979) and its contribution to H^T H, H^T r, the effective count and
//              the residual sum, reduced with warp shuffles + shared memory per block and
//              deterministically across blocks; the last block also runs the 6x6 Jacobi
//              eigen-decomposition of the pose block (the new degeneracy output).
#pragma once
#include "../../include/daliti_b200.h"
#include "dlt_common.cuh"
#include "dlt_map_kernels.cuh"
#include "dlt_peer.cuh"

namespace dlt {

// Device-resident iteration loop (dlt_iekf_update): the kernels of one enqueued iteration take the
// pose, the do_match decision and the scan size from device memory instead of from their
// parameters, and return at once when the loop has already ended.  ctl == nullptr: classic path.
struct IekfDev {
    dlt_iekf_block b;
    double K1c[24 * 12];  // K_1[:, :12] of the last Kalman update (laserMapping.cpp:1019)
    double HtH12[144];    // its H^T H (for G = K H at :1084)
    long long clocks[DLT_IEKF_MAX_ITER][16];  // instrumentation: SM clock at the stages of k_iekf_step (dlt_get_iekf_clocks)
};
#if defined(DLT_EMU)
#define DLT_STAMP(i)
#else
#define DLT_STAMP(i)                                          \
    do {                                                      \
        if (threadIdx.x == 0) dev->clocks[it][i] = clock64(); \
    } while (0)
#endif
struct LoopArgs {
    const IekfDev *ctl;
    const int *n_ptr;    // feats_down_size on the device (ScanScalars::n_down)
    const int *vox_ptr;  // VoxelGrid status of the scan (ScanScalars::vox_status)
    int fuse_step;       // k_residual: its last block also runs the iteration's solve / control step (iekf_step_block)
    // CUDA-graph form of the loop: a WHILE node around {IF(match) {k_knn8, k_knn}, k_residual}; the fused step drives both
    // conditions from the device (cudaGraphSetConditional), so iterations that do not run are never launched.
    int use_cond;
    unsigned long long cond_while, cond_match;
    // sharded map with attached peers (dlt_peer_attach): the last block of k_residual sums the partial normal equations over
    // the ranks through the peer mailboxes (dlt_peer.cuh) before the solve step -- no collective launch in between
    PeerComm *peer;
};
// Block-wide: resolve (n, do_match, pose) for this launch; false = nothing to do.  The pose ends
// up in shared memory either way so that both paths run the same code.
DLT_D bool loop_resolve(const LoopArgs &la, const Pose &P_param, Pose *sP, int &n, int *do_match) {
    bool go = true;
    if (la.ctl) {
        const dlt_iekf_block &c = la.ctl->b;
        const int dm = (c.iter == 0 || c.rematch_en) ? 1 : 0;
        if (c.done) go = false;
        if (do_match) {
            if (*do_match < 0) {  // a match-pass kernel: only runs when this iteration (re)matches
                if (!dm) go = false;
            } else {
                *do_match = dm;
            }
        }
        if (!go) return false;  // block-uniform: no barrier needed on the way out (this is the no-op launch's whole cost)
        n = *la.n_ptr;
        if (threadIdx.x < 24) reinterpret_cast<double *>(sP)[threadIdx.x] = c.state[threadIdx.x];
    } else {
        if (la.n_ptr) n = *la.n_ptr;  // host loop right behind dlt_scan_downsample_async: feats_down_size is still on the device
        if (threadIdx.x < 24) reinterpret_cast<double *>(sP)[threadIdx.x] = reinterpret_cast<const double *>(&P_param)[threadIdx.x];
    }
    __syncthreads();
    return go;
}

// per-point flag bits
constexpr unsigned char kFlagMatched = 1;     // 5 neighbours found and d2[4] <= max_sq_dist
constexpr unsigned char kFlagUnresolved = 2;  // ring-3 search could not prove exactness (far / crowded)
constexpr unsigned char kFlagForeign = 4;     // query owned by another shard
constexpr unsigned char kFlagNeedNN = 8;      // unresolved AND nothing at all within the 7^3 block: map_incremental needs the true nearest point

#ifndef DLT_KNN_MINBLOCKS
#define DLT_KNN_MINBLOCKS 8
#endif
#ifndef DLT_KNN_BATCH
#define DLT_KNN_BATCH 1  // bucket load steps (4 buckets each) in flight per warp
#endif
constexpr int kKnnWarps = 4;
constexpr int kCandMax = 128;
constexpr int kWlMax = 256;  // buckets of one ring batch (dense, raw-built maps: ten and more buckets per cell)

DLT_D Cand cand_inf() {
    Cand c;
    c.d2 = INFINITY;
    c.x = c.y = c.z = 0.f;
    c.id = 0x7FFFFFFF;
    return c;
}
DLT_D Cand warp_min_cand(Cand c) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        Cand t;
        t.d2 = __shfl_xor_sync(0xffffffffu, c.d2, o);
        t.x = __shfl_xor_sync(0xffffffffu, c.x, o);
        t.y = __shfl_xor_sync(0xffffffffu, c.y, o);
        t.z = __shfl_xor_sync(0xffffffffu, c.z, o);
        t.id = __shfl_xor_sync(0xffffffffu, c.id, o);
        if (cand_less(t, c)) c = t;
    }
    return c;
}

// keep the kK smallest (stated order) of a stream of candidates, sorted ascending, in registers
DLT_D void topk_insert(Cand (&b)[kK], const Cand &c) {
    if (!cand_less(c, b[kK - 1])) return;
    b[kK - 1] = c;
#pragma unroll
    for (int t = kK - 1; t > 0; t--) {
        if (cand_less(b[t], b[t - 1])) {
            Cand tmp = b[t];
            b[t] = b[t - 1];
            b[t - 1] = tmp;
        }
    }
}

struct KnnOut {
    float4 *qw;            // [n] world-frame query (x y z intensity) of this match pass
    float4 *nbr;           // [n][5] neighbour x y z d2, ascending
    int *nbr_id;           // [n][5] bucket*8+slot
    int *nbr_cnt;          // [n]
    unsigned char *flags;  // [n]
    float *d6lb;           // [n] lower bound of the squared distance from the query to every map point that is NOT one of its five
                           //     neighbours (0 = none known): lets a later match pass of the same scan prove the set unchanged
    int *far_list;         // indices of unresolved queries
    int *far_count;
    // The subset of them that map_incremental cannot classify from what the rings saw (see k_nn1): their single
    // nearest map point is searched for by brute force.  nn_count[0] = size, nn_count[1] = queries whose bucket
    // chains overflowed the work list (pathological; they need the full exact fallback).
    int *nn_list;
    int *nn_pos;           // [n] position of a query in nn_list
    unsigned long long *nn_key;  // [.] (d2 bits << 32 | id) of the nearest point, all-ones = none yet
    int *nn_count;
};

// distance from q to the slab of cell c along one axis (conservative: shrunk by a rounding slack)
DLT_D float axis_gap(float q, int c, float cell_edge, float slack) {
    float lo = (float)c * cell_edge, hi = (float)(c + 1) * cell_edge;
    float g = fmaxf(fmaxf(lo - q, q - hi), 0.f) - slack;
    return g > 0.f ? g : 0.f;
}

// map_incremental kernels enqueued behind the device-resident loop: run only when the loop armed them
DLT_D bool insert_gate(const LoopArgs &la, int &n) {
    if (!la.ctl) return true;
    if (la.ctl->b.insert_status != 1) return false;
    n = *la.n_ptr;
    return true;
}

// ---- selection of the 5 best among the staged candidates ------------------------------------
// The running best lives in shared memory (s_best, ascending) and is appended to the staging list
// before every selection, so one list holds everything that can still win.
//
// Exact version (stated order d2, x, y, z, id): five rounds of "smallest key greater than the
// previous pick", each a warp-wide lexicographic argmin.  Only runs when the fast path below sees
// an exact d2 tie among the winners.  Out of line: it is rare.
__device__ __noinline__ void knn_select_exact(Cand *best, int *nbest_out, const float4 *cand, const int *cid, int n, int lane) {
    Cand prev;
    prev.d2 = -1.f;
    prev.x = prev.y = prev.z = 0.f;
    prev.id = -1;
    int nb = 0;
    for (int t = 0; t < kK; t++) {
        Cand loc = cand_inf();
        for (int c = lane; c < n; c += 32) {
            float4 e = cand[c];
            Cand k;
            k.d2 = e.w;
            k.x = e.x;
            k.y = e.y;
            k.z = e.z;
            k.id = cid[c];
            if (cand_less(prev, k) && cand_less(k, loc)) loc = k;
        }
        loc = warp_min_cand(loc);
        __syncwarp();
        if (lane == 0) best[t] = loc;
        if (loc.d2 < INFINITY) nb++;
        prev = loc;
    }
    if (lane == 0) *nbest_out = nb;
    __syncwarp();
}

DLT_D unsigned warp_min_u32(unsigned k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) k = min(k, __shfl_xor_sync(0xffffffffu, k, o));
    return k;
}

constexpr int kCandSlots = kCandMax + 8;        // staging list + room for the running best
constexpr int kKeysPerLane = (kCandSlots + 31) / 32;

// Fast path.  d2 >= 0, so its bit pattern orders like the float: every lane keeps the d2 bits of
// its (at most kKeysPerLane) list entries in registers and five rounds of warp-min extract the
// five smallest DISTINCT values v0 < ... < v4.  If exactly five entries are <= v4 each value has
// one owner and the result is the stated order; otherwise two entries share a d2 (a tie inside
// the winners or at the k-th boundary) and the exact version decides.
DLT_D void knn_select(Cand *s_best, int *s_nbest, int &nbest, float &d5, float4 *cand, int *cid, int ncand, int lane) {
    // append the running best to the list
    if (lane < nbest) {
        Cand b = s_best[lane];
        cand[ncand + lane] = make_float4(b.x, b.y, b.z, b.d2);
        cid[ncand + lane] = b.id;
    }
    const int n = ncand + nbest;
    __syncwarp();
    unsigned key[kKeysPerLane];
#pragma unroll
    for (int j = 0; j < kKeysPerLane; j++) {
        const int c = lane + 32 * j;
        key[j] = (c < n) ? __float_as_uint(cand[c].w) : 0xFFFFFFFFu;
    }
    unsigned v[kK];
    unsigned lo = 0u;
#pragma unroll
    for (int t = 0; t < kK; t++) {
        unsigned loc = 0xFFFFFFFFu;
#pragma unroll
        for (int j = 0; j < kKeysPerLane; j++)
            if (key[j] >= lo) loc = min(loc, key[j]);
        loc = warp_min_u32(loc);
        v[t] = loc;
        lo = (loc == 0xFFFFFFFFu) ? loc : loc + 1u;
    }
    int nb = 0;
#pragma unroll
    for (int t = 0; t < kK; t++) nb += (v[t] != 0xFFFFFFFFu) ? 1 : 0;
    // how many entries are <= the largest selected value?  (all finite entries when fewer than 5 exist)
    const unsigned vmax = (nb == kK) ? v[kK - 1] : 0xFFFFFFFEu;
    int cnt = 0;
#pragma unroll
    for (int j = 0; j < kKeysPerLane; j++) cnt += __popc(__ballot_sync(0xffffffffu, key[j] <= vmax));
    if (cnt != nb) {  // warp-uniform
        knn_select_exact(s_best, s_nbest, cand, cid, n, lane);
        nbest = *s_nbest;
        d5 = (nbest == kK) ? s_best[kK - 1].d2 : INFINITY;
        __syncwarp();
        return;
    }
#pragma unroll
    for (int j = 0; j < kKeysPerLane; j++) {
        // is this entry one of the winners?  (each selected value has exactly one owner)
        int t = -1;
#pragma unroll
        for (int u = 0; u < kK; u++)
            if (key[j] == v[u]) t = u;
        if (t >= 0 && key[j] != 0xFFFFFFFFu) {
            const int c = lane + 32 * j;
            float4 e = cand[c];
            Cand r;
            r.d2 = e.w;
            r.x = e.x;
            r.y = e.y;
            r.z = e.z;
            r.id = cid[c];
            s_best[t] = r;
        }
    }
    nbest = nb;
    d5 = (nb == kK) ? __uint_as_float(v[kK - 1]) : INFINITY;
    __syncwarp();
}

// One query, one warp: rings 3^3 / 5^3 / 7^3 until the 5 best are proven exact.
DLT_D void knn_warp_query(const MapView &m, const float4 *__restrict__ q_pts, int qi, int body_frame, const Pose &P, float max_sq_dist,
                          const KnnOut &out, float4 *cand, int *cid, int *wl, Cand *best, int *s_nb_slot, int lane) {
    const unsigned FULL = 0xffffffffu;
    const unsigned lt_mask = (1u << lane) - 1u;
    float4 qb = q_pts[qi];
    float qx = qb.x, qy = qb.y, qz = qb.z;
    if (body_frame) body_to_world(P, qb.x, qb.y, qb.z, qx, qy, qz);
    const float cell_edge = m.ds * (float)(1 << m.cell_shift);
    int cx, cy, cz;
    cell_of_point(m, qx, qy, qz, cx, cy, cz);
    if (lane == 0) out.qw[qi] = make_float4(qx, qy, qz, qb.w);

    if (m.shard_count > 1 && tile_owner(cx, cy, cz, m.tile_shift, m.shard_count) != m.shard_rank) {
        if (lane == 0) {
            out.nbr_cnt[qi] = 0;
            out.flags[qi] = kFlagForeign;
        }
        return;
    }

    int nbest = 0, ncand = 0;
    float d5 = INFINITY;
    bool overflow = false, resolved = false;
    const float slack = 4e-7f * (fabsf(qx) + fabsf(qy) + fabsf(qz) + 16.f * cell_edge);
    const int grp = lane >> 3, sub = lane & 7;

    for (int R = 1; R <= 3 && !resolved; R++) {
        const int side = 2 * R + 1, total = side * side * side;
        const int mdiv = (R == 1) ? 21846 : (R == 2) ? 13108 : 9363;  // (i * mdiv) >> 16 == i / side for i < 407
        for (int base = 0; base < total; base += 32) {
            // ---- probe up to 32 cells of this ring
            int idx = base + lane;
            if (R == 1) idx = (idx == 0) ? 13 : (idx == 13) ? 0 : idx;  // the query's own cell first: d5 tightens early
            int b = -1;
            if (idx < total) {
                const int q1 = (idx * mdiv) >> 16, q2 = (q1 * mdiv) >> 16;
                const int dz = idx - q1 * side - R, dy = q1 - q2 * side - R, dx = q2 - R;
                int cheb = max(max(abs(dx), abs(dy)), abs(dz));
                if (cheb == R || (R == 1 && cheb == 0)) {
                    bool prune = false;
                    if (nbest == kK) {
                        float gx = axis_gap(qx, cx + dx, cell_edge, slack), gy = axis_gap(qy, cy + dy, cell_edge, slack),
                              gz = axis_gap(qz, cz + dz, cell_edge, slack);
                        float bd = gx * gx + gy * gy + gz * gz;
                        prune = bd * 0.99999f > d5;
                    }
                    if (!prune) b = map_find(m, pack_key(cx + dx, cy + dy, cz + dz));
                }
            }
            unsigned found = __ballot_sync(FULL, b >= 0);
            int nwl = __popc(found);
            if (b >= 0) wl[__popc(found & lt_mask)] = b;
            __syncwarp();
            // ---- stage the buckets: 8 lanes x 16 B = one 128-byte line each, 4 lines per load step and up to
            //      4 load steps (16 buckets) in flight before the first one is consumed
            for (int j0 = 0; j0 < nwl;) {
                const int batch = min(4 * DLT_KNN_BATCH, nwl - j0);  // chain buckets pushed below are picked up by a later batch
                float4 vb[DLT_KNN_BATCH];
                int bi[DLT_KNN_BATCH];
#pragma unroll
                for (int u = 0; u < DLT_KNN_BATCH; u++) {
                    const int k = 4 * u + grp;
                    bi[u] = (k < batch) ? wl[j0 + k] : -1;
                    vb[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (bi[u] >= 0) vb[u] = reinterpret_cast<const float4 *>(&m.buckets[bi[u]])[sub];
                }
#pragma unroll
                for (int u = 0; u < DLT_KNN_BATCH; u++) {
                    if (4 * u >= batch) break;  // warp-uniform
                    const float4 v = vb[u];
                    const int bidx = bi[u];
                    const bool act = bidx >= 0;
                    int hdr_next = __shfl_sync(FULL, __float_as_int(v.z), lane & ~7);
                    unsigned hdr_mask = __shfl_sync(FULL, __float_as_uint(v.w), lane & ~7);
                    bool pushn = act && sub == 0 && hdr_next >= 0;
                    unsigned pm = __ballot_sync(FULL, pushn);
                    if (pushn) {
                        int pos = nwl + __popc(pm & lt_mask);
                        if (pos < kWlMax) wl[pos] = hdr_next;
                    }
                    bool has = act && sub >= 1 && ((hdr_mask >> (sub - 1)) & 1u);
                    float d2 = has ? calc_dist(qx, qy, qz, v.x, v.y, v.z) : INFINITY;
                    bool keep = has && (nbest < kK || d2 <= d5);
                    unsigned km = __ballot_sync(FULL, keep);
                    if (keep) {
                        int pos = ncand + __popc(km & lt_mask);
                        cand[pos] = make_float4(v.x, v.y, v.z, d2);
                        cid[pos] = bidx * 8 + sub;
                    }
                    ncand += __popc(km);
                    nwl += __popc(pm);
                    if (nwl > kWlMax) {
                        overflow = true;  // pathological chain length; exact result comes from k_far_*
                        nwl = kWlMax;
                    }
                    __syncwarp();
                    if (ncand > kCandMax - 32) {  // staging area nearly full: fold it into the running best
                        knn_select(best, s_nb_slot, nbest, d5, cand, cid, ncand, lane);
                        ncand = 0;
                    }
                }
                j0 += batch;
            }
        }
        if (ncand > 0) {
            knn_select(best, s_nb_slot, nbest, d5, cand, cid, ncand, lane);
            ncand = 0;
        }
        // ---- exactness: every unseen point lies outside the (2R+1)^3 block of cells
        if (nbest == kK && !overflow) {
            float cov = INFINITY;
            cov = fminf(cov, qx - (float)(cx - R) * cell_edge);
            cov = fminf(cov, (float)(cx + R + 1) * cell_edge - qx);
            cov = fminf(cov, qy - (float)(cy - R) * cell_edge);
            cov = fminf(cov, (float)(cy + R + 1) * cell_edge - qy);
            cov = fminf(cov, qz - (float)(cz - R) * cell_edge);
            cov = fminf(cov, (float)(cz + R + 1) * cell_edge - qz);
            cov -= slack;
            if (cov > 0.f && d5 < cov * cov * 0.99999f) resolved = true;
        }
    }

    // The match gate (laserMapping.cpp:852-854) is decided here even for unresolved queries:
    // the 7^3 block covers >= 3 cell edges > sqrt(max_sq_dist) (dlt_create picks cell_shift so),
    // hence "5 found with d2[4] <= max_sq_dist" implies resolved.  Unresolved queries (nothing
    // close, or a pathological chain) get their exact neighbours lazily from k_far_* before
    // map_incremental consumes them.
    unsigned char fl = 0;
    if (resolved && d5 <= max_sq_dist) fl |= kFlagMatched;
    if (!resolved) fl |= kFlagUnresolved;
    __syncwarp();
    if (lane < kK) {
        const bool ok = lane < nbest;
        Cand r = best[ok ? lane : 0];
        out.nbr[(size_t)qi * kK + lane] = ok ? make_float4(r.x, r.y, r.z, r.d2) : make_float4(0.f, 0.f, 0.f, -1.f);
        out.nbr_id[(size_t)qi * kK + lane] = ok ? r.id : -1;
    }
    if (lane == 0) {
        out.nbr_cnt[qi] = nbest;
        if (!resolved) {
            int pos = atomicAdd(out.far_count, 1);
            out.far_list[pos] = qi;
            // What map_incremental (laserMapping.cpp:593-617) reads of Nearest_Points: the nearest point, and which of the
            // five lie within half a voxel diagonal of the voxel centre.  Every point closer than the distance `cov3` to the
            // faces of the 7^3 block was seen, so once ONE candidate is closer than that the nearest point is exact and every
            // unseen point is too far to matter.  Otherwise the true nearest point has to be searched for.
            float cov3 = INFINITY;
            cov3 = fminf(cov3, qx - (float)(cx - 3) * cell_edge);
            cov3 = fminf(cov3, (float)(cx + 4) * cell_edge - qx);
            cov3 = fminf(cov3, qy - (float)(cy - 3) * cell_edge);
            cov3 = fminf(cov3, (float)(cy + 4) * cell_edge - qy);
            cov3 = fminf(cov3, qz - (float)(cz - 3) * cell_edge);
            cov3 = fminf(cov3, (float)(cz + 4) * cell_edge - qz);
            cov3 -= slack;
            const bool have_near = !overflow && nbest > 0 && cov3 > 0.f && best[0].d2 < cov3 * cov3 * 0.99999f;
            if (overflow) atomicAdd(out.nn_count + 1, 1);
            if (!have_near) {
                fl |= kFlagNeedNN;
                int np = atomicAdd(out.nn_count, 1);
                out.nn_list[np] = qi;
                out.nn_pos[qi] = np;
                out.nn_key[np] = 0xFFFFFFFFFFFFFFFFull;
            }
        }
        out.flags[qi] = fl;
    }
}

// ------------------------------------------------------------------ first pass: 8 lanes per query over the 3x3x3 block
// Four queries per warp.  The 8 lanes of a group probe the 27 cells in 4 rounds (the table slot of the NEXT round is
// requested before the buckets of the current one are consumed) and fetch two 128-byte buckets per step (lane 0 the
// header, lanes 1-7 one point each; both loads in flight before the first is consumed).
//
// Selection works on 32-bit keys.  d2 >= 0, so its bit pattern orders like the float; the low 9 bits of the pattern are
// replaced by a tag that names the candidate inside the group: (visit << 3) | lane-in-group, `visit` = which of the
// group's (at most 128) bucket fetches brought it in.  Keys are unique, a key identifies its point (the bucket index of
// every visit is kept in shared memory) and sorting keys sorts by d2 truncated to 13 mantissa bits.  Every lane keeps the
// five smallest keys it saw with a branch-free min / max insertion plus the smallest key it ever dropped; the 8 sorted
// lists merge with five 8-lane mins.  Whenever the SIX best keys have pairwise different truncated distances the five
// winners are exactly the five smallest d2 in exact order (anything else has a truncated distance >= the sixth's, hence
// a larger d2 than the fifth); their coordinates are re-read from the (L1-hot) buckets and d2 is recomputed -- the same
// expression on the same inputs, so the same bits.  Otherwise (two of the six agree to 13 bits: a few queries in a
// thousand; this includes every exact d2 tie, which the stated order (d2, x, y, z, id) has to break), or when the 3^3
// block cannot prove the result exact, or after more than 128 bucket fetches, the query goes to `unres_list` for the
// warp-per-query kernel above, which searches from scratch with exact comparisons.
#ifndef DLT_KNN8_BLOCK
#define DLT_KNN8_BLOCK 256
#endif
constexpr int kKnn8Block = DLT_KNN8_BLOCK;
#ifndef DLT_KNN8_MINBLOCKS
#define DLT_KNN8_MINBLOCKS 6
#endif
constexpr unsigned kKey32Inf = 0xFFFFFFFFu;
constexpr int kKnn8TagBits = 10;                      // 7 bits visit + 3 bits lane-in-group
constexpr int kKnn8Visits = 128;                      // bucket fetches per group that a tag can name (896 map points in the 3^3 block)
constexpr unsigned kKey32Finite = 0x7F800000u >> kKnn8TagBits;  // truncated pattern of +inf
constexpr int kKnn8WlInts = 4 * kKnn8Visits;          // shared-memory ints per warp (bucket index of every visit)

DLT_D unsigned group8_min_u32(unsigned k) {
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) k = min(k, __shfl_xor_sync(0xffffffffu, k, o));
    return k;
}
// keep the kK smallest keys, ascending (branch-free; k = kKey32Inf is a no-op); `dropped` <- smallest key that fell off the list
DLT_D void topk_insert_k32(unsigned (&b)[kK], unsigned k, unsigned &dropped) {
    dropped = min(dropped, max(k, b[kK - 1]));
    b[kK - 1] = min(b[kK - 1], k);
#pragma unroll
    for (int t = kK - 1; t > 0; t--) {
        const unsigned lo = min(b[t], b[t - 1]), hi = max(b[t], b[t - 1]);
        b[t - 1] = lo;
        b[t] = hi;
    }
}
// one staged bucket line: lane `sub` of the group holds 16 bytes of bucket bb (sub 0 = header); tag = (visit << 3) | sub
DLT_D void knn8_consume(const float4 v, int &bb, int lane, int sub, float qx, float qy, float qz, unsigned (&best)[kK], unsigned &dropped,
                        unsigned tag, bool tag_ok) {
    const int hdr_next = __shfl_sync(0xffffffffu, __float_as_int(v.z), lane & ~7);
    const unsigned hdr_mask = __shfl_sync(0xffffffffu, __float_as_uint(v.w), lane & ~7);
    const float d2 = calc_dist(qx, qy, qz, v.x, v.y, v.z);
    const bool has = bb >= 0 && tag_ok && (((hdr_mask << 1) >> sub) & 1u);  // (bit 0 of the shifted mask = the header lane: never set)
    const unsigned key = has ? ((__float_as_uint(d2) & ~((1u << kKnn8TagBits) - 1u)) | tag) : kKey32Inf;
    topk_insert_k32(best, key, dropped);
    bb = (bb >= 0) ? hdr_next : -1;
}
// first table slot of a cell's probe sequence, requested early; knn8_find_finish completes the lookup
struct Probe {
    unsigned long long key;
    unsigned h;
    Slot s;
    bool on;
};
DLT_D Probe knn8_find_begin(const MapView &m, bool on, int cx, int cy, int cz) {
    Probe p;
    p.on = on;
    p.key = pack_key(cx, cy, cz);
    p.h = hash_key(p.key) & m.table_mask;
    p.s.key = kEmptyKey;
    p.s.bucket = -1;
    p.s.pad = 0;
    if (on) p.s = load_slot(&m.table[p.h]);
    return p;
}
DLT_D int knn8_find_finish(const MapView &m, const Probe &p) {
    if (!p.on) return -1;
    Slot s = p.s;
    unsigned h = p.h;
    for (unsigned probe = 0; probe <= m.table_mask; probe++) {
        if (s.key == p.key) return s.bucket;
        if (s.key == kEmptyKey) return -1;
        h = (h + 1) & m.table_mask;
        s = load_slot(&m.table[h]);
    }
    return -1;
}

// the four queries q0 .. q0+3 of one warp; wl = this warp's kKnn8WlInts ints of shared memory
DLT_D void knn8_group(const MapView &m, const float4 *__restrict__ q_pts, int n, int body_frame, const Pose &P, float max_sq_dist, const KnnOut &out,
                      int *unres_list, int *unres_count, int q0, int lane, int *wl, int unres_cap = 0x7FFFFFFF, const int *qlist = nullptr) {
    const unsigned FULL = 0xffffffffu;
    const int grp = lane >> 3, sub = lane & 7;
    const bool live = q0 + grp < n;  // (with a list: n = its length)
    const int qi = qlist ? (live ? qlist[q0 + grp] : 0) : q0 + grp;
    float qx = 0.f, qy = 0.f, qz = 0.f;
    int cx = 0, cy = 0, cz = 0;
    bool work = false;
    const float cell_edge = m.ds * (float)(1 << m.cell_shift);
    if (live) {
        float4 qb = q_pts[qi];
        qx = qb.x;
        qy = qb.y;
        qz = qb.z;
        if (body_frame) body_to_world(P, qb.x, qb.y, qb.z, qx, qy, qz);
        cell_of_point(m, qx, qy, qz, cx, cy, cz);
        if (sub == 0) out.qw[qi] = make_float4(qx, qy, qz, qb.w);
        work = true;
        if (m.shard_count > 1 && tile_owner(cx, cy, cz, m.tile_shift, m.shard_count) != m.shard_rank) {
            work = false;
            if (sub == 0) {
                out.nbr_cnt[qi] = 0;
                out.flags[qi] = kFlagForeign;
            }
        }
    }
    unsigned best[kK];
#pragma unroll
    for (int t = 0; t < kK; t++) best[t] = kKey32Inf;
    unsigned dropped = kKey32Inf;
    int *gwl = wl + grp * kKnn8Visits;
    int visit = 0;        // group-uniform: buckets this group fetched so far
    bool ovf = false;     // this group needed a fetch beyond what a tag can name

    // cell ci = r * 8 + sub of the 3^3 block, in the order (x fastest)
    Probe pr = knn8_find_begin(m, work, cx + (sub % 3) - 1, cy + ((sub / 3) % 3) - 1, cz + (sub / 9) - 1);
    for (int r = 0; r < 4; r++) {
        const int cn = (r + 1) * 8 + sub;  // next round's cell: its slot is in flight while this round's buckets are consumed
        Probe nx = knn8_find_begin(m, work && r < 3 && cn < 27, cx + (cn % 3) - 1, cy + ((cn / 3) % 3) - 1, cz + (cn / 9) - 1);
        const int b = knn8_find_finish(m, pr);
        pr = nx;
        unsigned gb = (__ballot_sync(FULL, b >= 0) >> (grp * 8)) & 0xFFu;  // this group's found cells
        while (__any_sync(FULL, gb != 0u)) {                                // warp-uniform
            const int s0 = gb ? (__ffs((int)gb) - 1) : -1;
            gb &= gb - 1u;  // (0 stays 0)
            const int s1 = gb ? (__ffs((int)gb) - 1) : -1;
            gb &= gb - 1u;
            int bb0 = __shfl_sync(FULL, b, (grp << 3) + (s0 < 0 ? 0 : s0));
            int bb1 = __shfl_sync(FULL, b, (grp << 3) + (s1 < 0 ? 0 : s1));
            if (s0 < 0) bb0 = -1;
            if (s1 < 0) bb1 = -1;
            while (__any_sync(FULL, bb0 >= 0 || bb1 >= 0)) {  // the two cells' bucket chains
                float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
                if (bb0 >= 0) v0 = reinterpret_cast<const float4 *>(&m.buckets[bb0])[sub];  // 8 lanes x 16 B = one line
                if (bb1 >= 0) v1 = reinterpret_cast<const float4 *>(&m.buckets[bb1])[sub];
                const int t0 = visit, t1 = visit + (bb0 >= 0 ? 1 : 0);
                visit = t1 + (bb1 >= 0 ? 1 : 0);
                if (visit > kKnn8Visits) ovf = true;
                if (sub == 0) {
                    if (bb0 >= 0 && t0 < kKnn8Visits) gwl[t0] = bb0;
                    if (bb1 >= 0 && t1 < kKnn8Visits) gwl[t1] = bb1;
                }
                knn8_consume(v0, bb0, lane, sub, qx, qy, qz, best, dropped, ((unsigned)t0 << 3) | (unsigned)sub, t0 < kKnn8Visits);
                knn8_consume(v1, bb1, lane, sub, qx, qy, qz, best, dropped, ((unsigned)t1 << 3) | (unsigned)sub, t1 < kKnn8Visits);
            }
        }
    }
    // ---- merge the group's 8 sorted lists: five rounds of 8-lane min, the owner (named by the tag) pops its head
    int nb = 0;
    unsigned prevt = 0xFFFFFFFFu;  // truncated distance of the previous winner
    bool tie = false;
    unsigned mine = kKey32Inf;
#pragma unroll
    for (int t = 0; t < kK; t++) {
        const unsigned w = group8_min_u32(best[0]);
        const unsigned wt = w >> kKnn8TagBits;
        const bool valid = wt < kKey32Finite;
        if (valid && (int)(w & 7u) == sub) {
#pragma unroll
            for (int u = 0; u < kK - 1; u++) best[u] = best[u + 1];
            best[kK - 1] = kKey32Inf;
        }
        nb += valid ? 1 : 0;
        if (valid && wt == prevt) tie = true;
        prevt = wt;
        if (sub == t) mine = w;
    }
    // the group's 6th-smallest key: a remaining list head or something a lane dropped earlier
    const unsigned sixth = group8_min_u32(min(best[0], dropped));
    if (nb == kK && (sixth >> kKnn8TagBits) == prevt) tie = true;
    __syncwarp();  // the visit list (written by lane 0 of the group) is read by lanes 0-4
    // exact distances of the winners, from their coordinates
    const bool cand_ok = work && nb == kK && !tie && !ovf;
    float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
    float myd2 = INFINITY;
    int id = -1;
    if (cand_ok && sub < kK) {
        const int osub = (int)(mine & 7u);
        id = gwl[(mine >> 3) & (unsigned)(kKnn8Visits - 1)] * 8 + osub;
        p = m.buckets[id >> 3].pts[osub - 1];
        myd2 = calc_dist(qx, qy, qz, p.x, p.y, p.z);
    }
    const float d5 = __shfl_sync(FULL, myd2, (grp << 3) + (kK - 1));
    __syncwarp();  // (the next group of this warp reuses the visit list)
    if (!work) return;
    const float slack = 4e-7f * (fabsf(qx) + fabsf(qy) + fabsf(qz) + 16.f * cell_edge);
    float cov = INFINITY;
    cov = fminf(cov, qx - (float)(cx - 1) * cell_edge);
    cov = fminf(cov, (float)(cx + 2) * cell_edge - qx);
    cov = fminf(cov, qy - (float)(cy - 1) * cell_edge);
    cov = fminf(cov, (float)(cy + 2) * cell_edge - qy);
    cov = fminf(cov, qz - (float)(cz - 1) * cell_edge);
    cov = fminf(cov, (float)(cz + 2) * cell_edge - qz);
    cov -= slack;
    const bool resolved = cand_ok && cov > 0.f && d5 < cov * cov * 0.99999f;
#if defined(DLT_EMU) && defined(DLT_KNN8_STATS)
    if (sub == 0) {
        extern long long g_knn8_stats[8];
        g_knn8_stats[0]++;
        if (nb < kK) g_knn8_stats[1]++;
        else if (tie) g_knn8_stats[2]++;
        else if (ovf) g_knn8_stats[3]++;
        else if (!resolved) g_knn8_stats[4]++;
    }
#endif
    if (!resolved) {
        if (sub == 0) {
            out.d6lb[qi] = 0.f;
            const int pos = atomicAdd(unres_count, 1);
            if (pos < unres_cap) unres_list[pos] = qi;  // (the fused loop kernels size their block-local list to the chunk: never full)
        }
        return;
    }
    if (sub < kK) {
        out.nbr[(size_t)qi * kK + sub] = make_float4(p.x, p.y, p.z, myd2);
        out.nbr_id[(size_t)qi * kK + sub] = id;
    }
    if (sub == 0) {
        out.nbr_cnt[qi] = kK;
        out.flags[qi] = (d5 <= max_sq_dist) ? kFlagMatched : (unsigned char)0;
        // every other point of the block is at least as far as the sixth key says (its distance bits were truncated downwards),
        // every point outside the block farther than the block's faces
        out.d6lb[qi] = fminf(__uint_as_float(sixth & ~((1u << kKnn8TagBits) - 1u)), cov * cov * 0.99999f);
    }
}

// ---- a later match pass of the same scan (rematch, laserMapping.cpp:847 with rematch_en): most neighbour sets can be PROVEN
// unchanged.  The query moved by delta since the pass that found its five neighbours (largest distance d5) and every other map
// point was then at least sqrt(d6lb) away: if d5 + delta < sqrt(d6lb) - delta the same five are still strictly nearer than
// everything else, so only their distances are recomputed (the same fp32 expression a search evaluates) and re-ordered in the
// stated order.  Anything else is searched again.  Returns true when the set was reused.
DLT_D bool knn_try_reuse(const MapView &m, const float4 *__restrict__ down, int i, const Pose &P, float max_sq_dist, const KnnOut &out) {
    const float4 qb = down[i];
    float qx, qy, qz;
    body_to_world(P, qb.x, qb.y, qb.z, qx, qy, qz);
    const float4 old = out.qw[i];
    out.qw[i] = make_float4(qx, qy, qz, qb.w);
    if (m.shard_count > 1) {  // sharded map: the owner of the cell the query is in NOW evaluates it
        int cx, cy, cz;
        cell_of_point(m, qx, qy, qz, cx, cy, cz);
        if (tile_owner(cx, cy, cz, m.tile_shift, m.shard_count) != m.shard_rank) {
            out.nbr_cnt[i] = 0;
            out.flags[i] = kFlagForeign;
            out.d6lb[i] = 0.f;
            return true;  // nothing to search here
        }
        // (a query this rank did not own in the earlier pass carries kFlagForeign and no bound: searched below)
    }
    const float d6 = out.d6lb[i];
    if (!(d6 > 0.f) || out.nbr_cnt[i] != kK || (out.flags[i] & ~kFlagMatched)) return false;
    Cand c[kK];
#pragma unroll
    for (int j = 0; j < kK; j++) {
        const float4 e = out.nbr[(size_t)i * kK + j];
        c[j].x = e.x;
        c[j].y = e.y;
        c[j].z = e.z;
        c[j].d2 = e.w;
        c[j].id = out.nbr_id[(size_t)i * kK + j];
    }
    const float delta = sqrtf(calc_dist(qx, qy, qz, old.x, old.y, old.z)) * 1.00001f + 1e-12f;
    const float reach = sqrtf(c[kK - 1].d2) * 1.00001f + 2.f * delta;  // no neighbour is farther than this from the new position
    const float wall = sqrtf(d6) * 0.99999f;                            // no other point was nearer than this to the old one
    if (!(delta < 0.25f) || !(reach < wall)) return false;
#pragma unroll
    for (int j = 0; j < kK; j++) c[j].d2 = calc_dist(qx, qy, qz, c[j].x, c[j].y, c[j].z);
#pragma unroll
    for (int a = 1; a < kK; a++)  // insertion sort in the stated order (d2, x, y, z, id)
#pragma unroll
        for (int b = a; b > 0; b--)
            if (cand_less(c[b], c[b - 1])) {
                const Cand t = c[b];
                c[b] = c[b - 1];
                c[b - 1] = t;
            }
#pragma unroll
    for (int j = 0; j < kK; j++) {
        out.nbr[(size_t)i * kK + j] = make_float4(c[j].x, c[j].y, c[j].z, c[j].d2);
        out.nbr_id[(size_t)i * kK + j] = c[j].id;
    }
    out.flags[i] = (c[kK - 1].d2 <= max_sq_dist) ? kFlagMatched : (unsigned char)0;
    const float w2 = wall - delta;  // the bound seen from the new position (for a third pass)
    out.d6lb[i] = w2 > 0.f ? w2 * w2 * 0.99999f : 0.f;
    return true;
}

// Warp-stride over groups of four queries, so the grid may be sized from an estimate of n (the
// device-resident loop launches before the host knows feats_down_size) and capped at what is resident in one wave.
// Also zeroes far_count for the warp-per-query pass that follows in stream order.
__global__ void __launch_bounds__(kKnn8Block, DLT_KNN8_MINBLOCKS)
    k_knn8(MapView m, const float4 *__restrict__ q_pts, int n, int body_frame, Pose P, float max_sq_dist, KnnOut out, int *__restrict__ unres_list,
           int *__restrict__ unres_count, LoopArgs la) {
    DLT_PDL_WAIT();
    __shared__ Pose sP;
    __shared__ int s_wl[kKnn8Block / 32][kKnn8WlInts];
    int is_match = -1;
    if (!loop_resolve(la, P, &sP, n, &is_match)) return;  // block-uniform
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        *out.far_count = 0;
        out.nn_count[0] = 0;
        out.nn_count[1] = 0;
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int stride = gridDim.x * (kKnn8Block / 32) * 4;
    for (int q0 = (blockIdx.x * (kKnn8Block / 32) + warp) * 4; q0 < n; q0 += stride)  // warp-uniform
        knn8_group(m, q_pts, n, body_frame, sP, max_sq_dist, out, unres_list, unres_count, q0, lane, s_wl[warp]);
}

// First kernel of a LATER match pass of the same scan (knn_try_reuse): one query per thread; what cannot be proven unchanged is
// handed to the warp-per-query kernel (bulk mode), which searches it again from scratch -- typically one query in ten.
// Also zeroes far_count for that kernel, as k_knn8 does.
constexpr int kReuseBlock = 128;
__global__ void __launch_bounds__(kReuseBlock)
    k_knn_reuse(MapView m, const float4 *__restrict__ q_pts, int n, Pose P, float max_sq_dist, KnnOut out, int *__restrict__ unres_list,
                int *__restrict__ unres_count, LoopArgs la, int *__restrict__ reuse_stats) {
    DLT_PDL_WAIT();
    __shared__ Pose sP;
    int is_match = -1;
    if (!loop_resolve(la, P, &sP, n, &is_match)) return;  // block-uniform
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        *out.far_count = 0;
        out.nn_count[0] = 0;
        out.nn_count[1] = 0;
    }
    int seen = 0, redo = 0;
    for (int i = blockIdx.x * kReuseBlock + threadIdx.x; i < n; i += gridDim.x * kReuseBlock) {
        seen++;
        if (!knn_try_reuse(m, q_pts, i, sP, max_sq_dist, out)) {
            redo++;
            out.d6lb[i] = 0.f;
            unres_list[atomicAdd(unres_count, 1)] = i;
        }
    }
    if (reuse_stats) {  // instrumentation: queries seen / searched again by the reuse passes
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            seen += __shfl_xor_sync(0xffffffffu, seen, o);
            redo += __shfl_xor_sync(0xffffffffu, redo, o);
        }
        if ((threadIdx.x & 31) == 0) {
            atomicAdd(reuse_stats, seen);
            atomicAdd(reuse_stats + 1, redo);
        }
    }
}

// Warp-per-query kernel.  list == nullptr: queries 0..n-1, one per warp.  Otherwise the queries list[0..*count)
// (the ones k_knn_ring1 could not resolve) with a warp-stride loop.
__global__ void __launch_bounds__(kKnnWarps * 32, DLT_KNN_MINBLOCKS)
    k_knn(MapView m, const float4 *__restrict__ q_pts, int n, int body_frame, Pose P, float max_sq_dist, KnnOut out, const int *__restrict__ list,
          const int *__restrict__ count, LoopArgs la, int bulk) {
    DLT_PDL_WAIT();
    __shared__ float4 s_cand[kKnnWarps][kCandSlots];
    __shared__ int s_cid[kKnnWarps][kCandSlots];
    __shared__ int s_wl[kKnnWarps][kWlMax];
    __shared__ Cand s_bestw[kKnnWarps][kK];
    __shared__ int s_nb[kKnnWarps];
    __shared__ Pose sP;
    int is_match = -1;
    if (!loop_resolve(la, P, &sP, n, &is_match)) return;  // block-uniform
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total_warps = gridDim.x * kKnnWarps;
    const int limit = list ? *count : n;
    if (bulk && list && limit > total_warps) {  // block-uniform.  The list holds ordinary queries (a rematch pass could not prove
                         // their old neighbour sets) and is longer than one round of a warp per query: 8 lanes per query first, a
                         // warp per query only for what that leaves -- efficient however long the list is.  (A list that fits one
                         // round is finished sooner by the plain loop below: one search latency instead of two.)
        __shared__ int s_left[kKnnWarps][4];
        __shared__ int s_nleft[kKnnWarps];
        __shared__ int s_wl8[kKnnWarps][kKnn8WlInts];
        for (int g = (blockIdx.x * kKnnWarps + warp) * 4; g < limit; g += total_warps * 4) {  // warp-uniform
            if (lane == 0) s_nleft[warp] = 0;
            __syncwarp();
            knn8_group(m, q_pts, limit, body_frame, sP, max_sq_dist, out, s_left[warp], &s_nleft[warp], g, lane, s_wl8[warp], 4, list);
            __syncwarp();
            const int nl = min(s_nleft[warp], 4);
            for (int u = 0; u < nl; u++) {
                knn_warp_query(m, q_pts, s_left[warp][u], body_frame, sP, max_sq_dist, out, s_cand[warp], s_cid[warp], s_wl[warp], s_bestw[warp], &s_nb[warp], lane);
                __syncwarp();
            }
        }
        return;
    }
    for (int w = blockIdx.x * kKnnWarps + warp; w < limit; w += total_warps) {  // warp-uniform; no block-level barrier
        const int qi = list ? list[w] : w;
        knn_warp_query(m, q_pts, qi, body_frame, sP, max_sq_dist, out, s_cand[warp], s_cid[warp], s_wl[warp], s_bestw[warp], &s_nb[warp], lane);
        __syncwarp();
    }
}

// ------------------------------------------------------------------ exact fallback for unresolved queries
// Streams the whole bucket pool once; lane = query, so a warp serves 32 queries and keeps
// their running 5 best in registers.  grid = (query groups, map slices); partial results
// are merged by k_far_merge.  Only needed for points with no 5 neighbours within ~3 cells
// (or > kCandMax candidates): their exact neighbours feed map_incremental
// (laserMapping.cpp:593-617), never the residuals.
constexpr int kFarWarps = 4;
constexpr int kFarTile = 128;  // buckets per shared-memory tile

__global__ void __launch_bounds__(kFarWarps * 32)
    k_far_scan(MapView m, int n_buckets, const float4 *__restrict__ qw, const int *__restrict__ far_list, int far_off, int nfar,
               int n_slices, Cand *__restrict__ partial /* [far][slice][5] */) {
    DLT_PDL_WAIT();
    __shared__ float4 tile[kFarTile * 8];
    const int groups = (nfar + kFarWarps * 32 - 1) / (kFarWarps * 32);
    const int slice = blockIdx.y;
    const int per = (n_buckets + n_slices - 1) / n_slices;
    const int b0 = slice * per, b1 = min(n_buckets, b0 + per);
    for (int g = blockIdx.x; g < groups; g += gridDim.x) {  // block-uniform loop
        const int f = g * kFarWarps * 32 + threadIdx.x;
        const bool live = f < nfar;
        float qx = 0.f, qy = 0.f, qz = 0.f;
        if (live) {
            float4 q = qw[far_list[far_off + f]];
            qx = q.x;
            qy = q.y;
            qz = q.z;
        }
        Cand best[kK];
#pragma unroll
        for (int t = 0; t < kK; t++) best[t] = cand_inf();
        for (int tb = b0; tb < b1; tb += kFarTile) {
            const int nb = min(kFarTile, b1 - tb);
            __syncthreads();
            for (int k = threadIdx.x; k < nb * 8; k += kFarWarps * 32) tile[k] = reinterpret_cast<const float4 *>(&m.buckets[tb])[k];
            __syncthreads();
            if (live) {
                for (int k = 0; k < nb; k++) {
                    unsigned msk = __float_as_uint(tile[k * 8].w) & 0x7Fu;
                    while (msk) {
                        int s = __ffs((int)msk) - 1;
                        msk &= msk - 1u;
                        float4 e = tile[k * 8 + 1 + s];
                        Cand c;
                        c.d2 = calc_dist(qx, qy, qz, e.x, e.y, e.z);
                        c.x = e.x;
                        c.y = e.y;
                        c.z = e.z;
                        c.id = (tb + k) * 8 + 1 + s;
                        topk_insert(best, c);
                    }
                }
            }
        }
        if (live) {
#pragma unroll
            for (int t = 0; t < kK; t++) partial[((size_t)f * n_slices + slice) * kK + t] = best[t];
        }
    }
}

__global__ void k_far_merge(const int *__restrict__ far_list, int far_off, int nfar, int n_slices, const Cand *__restrict__ partial,
                            float max_sq_dist, KnnOut out) {
    DLT_PDL_WAIT();
    for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < nfar; f += gridDim.x * blockDim.x) {
        Cand best[kK];
#pragma unroll
        for (int t = 0; t < kK; t++) best[t] = cand_inf();
        for (int s = 0; s < n_slices; s++)
            for (int t = 0; t < kK; t++) {
                Cand c = partial[((size_t)f * n_slices + s) * kK + t];
                if (c.d2 < INFINITY) topk_insert(best, c);
            }
        const int qi = far_list[far_off + f];
        int nb = 0;
#pragma unroll
        for (int t = 0; t < kK; t++) {
            bool ok = best[t].d2 < INFINITY;
            nb += ok ? 1 : 0;
            out.nbr[(size_t)qi * kK + t] = ok ? make_float4(best[t].x, best[t].y, best[t].z, best[t].d2) : make_float4(0.f, 0.f, 0.f, -1.f);
            out.nbr_id[(size_t)qi * kK + t] = ok ? best[t].id : -1;
        }
        out.nbr_cnt[qi] = nb;
        unsigned char fl = 0;
        if (nb == kK && best[kK - 1].d2 <= max_sq_dist) fl |= kFlagMatched;
        out.flags[qi] = fl;  // now exact
    }
}

// ------------------------------------------------------------------ nearest map point of the kFlagNeedNN queries
// Exact, k = 1, over the whole bucket pool -- but not all of it is looked at.  The pool is cut into runs of kPoolRun
// consecutive buckets whose cell boxes are kept up to date where buckets are allocated (pool_box_touch; allocation order is
// first-touch order, so a run is spatially compact).
//   k_nn1_seed  one live sample point per run against every query: the smallest of those distances is a TRUE upper bound
//               of each query's nearest distance (combined with a 64-bit atomicMin on (d2 bits << 32 | id), like everything here)
//   k_nn1       lane = query, the pool in gridDim.y slices of whole runs streamed through shared memory; a run is only staged
//               when some query of the block could still find something nearer (or equally near, for the id tie-break) inside
//               its box than what it already holds
// On a map that covers the scene the far queries' nearest points lie in a handful of runs; the pass that took 0.8 ms of brute
// force on the 2.2 M-point C2 map reads a few percent of the pool.  Same keys as the brute force by construction.
constexpr int kNn1Block = 128;
constexpr int kNn1Tile = kPoolRun;  // buckets per shared-memory tile = one pool run

// A point at squared distance d2 from the query lies in a cell at most this many cells away (Chebyshev distance between the cell
// indices, which come from floor(x / ds) >> shift): a point of a cell k cells away is at least (k - 1) cell edges away.
DLT_D int nn1_reach(float d2, float cell_edge) {
    const float r = sqrtf(d2) * 1.00001f / cell_edge;
    return r < 1.0e9f ? (int)r + 2 : 0x3FFFFFFF;
}
DLT_D bool nn1_gate(const int *nn_count, int *gate) {
    if (gate && *gate != 1) return false;
    if (gate && nn_count[1] > 0) {  // a bucket chain overflowed the ring search: the host runs the full exact fallback instead
        if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *gate = 2;
        return false;
    }
    return nn_count[0] > 0;
}

constexpr int kSeedBlock = 256;
constexpr int kSeedBatch = 8;  // queries per block barrier
__global__ void __launch_bounds__(kSeedBlock)
    k_nn1_seed(MapView m, const int *__restrict__ n_buckets_ptr, const float4 *__restrict__ qw, const int *__restrict__ nn_list,
               const int *__restrict__ nn_count, unsigned long long *__restrict__ nn_key, int *gate) {
    DLT_PDL_WAIT();
    __shared__ unsigned long long s_min[kSeedBatch][kSeedBlock / 32];
    if (!nn1_gate(nn_count, gate)) return;  // block-uniform
    const int nn = nn_count[0];
    const int n_buckets = min(*n_buckets_ptr, m.bucket_cap);
    const int n_runs = (n_buckets + kPoolRun - 1) / kPoolRun;
    // the work is queries x sampled runs: with many queries every `step`-th run is sampled (any subset gives a valid bound)
    const int step = 1 + nn / 512;
    const int n_samp = (n_runs + step - 1) / step;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int s0 = blockIdx.x * kSeedBlock; s0 < n_samp; s0 += gridDim.x * kSeedBlock) {  // block-uniform
        // this thread's sample: the first live point of the first bucket of its run that has one (of the first four)
        const int r = (s0 + threadIdx.x) * step;
        float4 sp = make_float4(0.f, 0.f, 0.f, 0.f);
        int sid = -1;
        if (r < n_runs) {
            for (int k = 0; k < 4 && sid < 0; k++) {
                const int b = r * kPoolRun + k;
                if (b >= n_buckets) break;
                const unsigned msk = m.buckets[b].mask & 0x7Fu;
                if (msk) {
                    const int sl = __ffs((int)msk) - 1;
                    sp = m.buckets[b].pts[sl];
                    sid = b * 8 + 1 + sl;
                }
            }
        }
        for (int f0 = 0; f0 < nn; f0 += kSeedBatch) {  // block-uniform: every query against the block's samples, one atomic per query and block
            unsigned long long key[kSeedBatch];
#pragma unroll
            for (int u = 0; u < kSeedBatch; u++) {
                key[u] = 0xFFFFFFFFFFFFFFFFull;
                if (f0 + u < nn && sid >= 0) {
                    const float4 q = qw[nn_list[f0 + u]];
                    key[u] = ((unsigned long long)__float_as_uint(calc_dist(q.x, q.y, q.z, sp.x, sp.y, sp.z)) << 32) | (unsigned)sid;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const unsigned long long t = __shfl_xor_sync(0xffffffffu, key[u], o);
                    key[u] = t < key[u] ? t : key[u];
                }
            }
            __syncthreads();  // (s_min of the previous batch has been read)
            if (lane == 0) {
#pragma unroll
                for (int u = 0; u < kSeedBatch; u++) s_min[u][warp] = key[u];
            }
            __syncthreads();
            if (threadIdx.x < kSeedBatch && f0 + (int)threadIdx.x < nn) {
                unsigned long long best = s_min[threadIdx.x][0];
#pragma unroll
                for (int w = 1; w < kSeedBlock / 32; w++) best = s_min[threadIdx.x][w] < best ? s_min[threadIdx.x][w] : best;
                if (best != 0xFFFFFFFFFFFFFFFFull) atomicMin(&nn_key[f0 + threadIdx.x], best);
            }
        }
    }
}

__global__ void __launch_bounds__(kNn1Block)
    k_nn1(MapView m, const int *__restrict__ n_buckets_ptr, const float4 *__restrict__ qw, const int *__restrict__ nn_list,
          const int *__restrict__ nn_count, unsigned long long *__restrict__ nn_key, int *gate) {
    DLT_PDL_WAIT();
    __shared__ float4 tile[kNn1Tile * 8];
    __shared__ int s_need;
    if (!nn1_gate(nn_count, gate)) return;  // block-uniform
    const int nn = nn_count[0];
    const int n_buckets = min(*n_buckets_ptr, m.bucket_cap);
    const int n_runs = (n_buckets + kPoolRun - 1) / kPoolRun;
    const int n_slices = gridDim.y;
    const int per = (n_runs + n_slices - 1) / n_slices;  // whole runs per slice
    const int r_lo = blockIdx.y * per, r_hi = min(n_runs, r_lo + per);
    const int groups = (nn + kNn1Block - 1) / kNn1Block;
    const float cell_edge = m.ds * (float)(1 << m.cell_shift);
    for (int g = blockIdx.x; g < groups; g += gridDim.x) {  // block-uniform
        const int f = g * kNn1Block + threadIdx.x;
        const bool live = f < nn;
        float qx = 0.f, qy = 0.f, qz = 0.f, slack = 0.f;
        int qcx = 0, qcy = 0, qcz = 0, reach = 0x3FFFFFFF;  // cells farther than `reach` (Chebyshev) cannot hold anything as near as `best`
        unsigned long long best = 0xFFFFFFFFFFFFFFFFull;
        if (live) {
            const float4 q = qw[nn_list[f]];
            qx = q.x;
            qy = q.y;
            qz = q.z;
            slack = 4e-7f * (fabsf(qx) + fabsf(qy) + fabsf(qz) + 16.f * cell_edge);
            best = nn_key[f];  // the seed pass's upper bound (other blocks may lower it meanwhile: any value read is valid)
            cell_of_point(m, qx, qy, qz, qcx, qcy, qcz);
            if (best != 0xFFFFFFFFFFFFFFFFull) reach = nn1_reach(__uint_as_float((unsigned)(best >> 32)), cell_edge);
        }
        const unsigned long long best0 = best;
        for (int r = r_lo; r < r_hi; r++) {  // block-uniform
            // can this run hold anything at most as far as what the lane has?  (box of the cells its buckets belong to)
            bool need = false;
            if (live) {
                float bd = 0.f;
                bool empty = false;
#pragma unroll
                for (int a = 0; a < 3; a++) {
                    const int cmin = m.pool_box_min[3 * r + a], cmax = m.pool_box_max[3 * r + a];
                    if (cmin > cmax) empty = true;
                    const float q = (a == 0) ? qx : (a == 1) ? qy : qz;
                    const float lo = (float)cmin * cell_edge, hi = (float)(cmax + 1) * cell_edge;
                    float gap = fmaxf(fmaxf(lo - q, q - hi), 0.f) - slack;
                    gap = gap > 0.f ? gap : 0.f;
                    bd = bd + gap * gap;
                }
                const float bestd = __uint_as_float((unsigned)(best >> 32));  // (+inf pattern or NaN pattern of all-ones: the test below then passes)
                need = !empty && !(bd * 0.99999f > bestd);
            }
            __syncthreads();  // (the tile of the previous run has been consumed; s_need is free)
            if (threadIdx.x == 0) s_need = 0;
            __syncthreads();
            if (need) s_need = 1;
            __syncthreads();
            if (!s_need) continue;  // block-uniform
            const int tb = r * kPoolRun, nb = min(kPoolRun, n_buckets - tb);
            for (int k = threadIdx.x; k < nb * 8; k += kNn1Block) tile[k] = reinterpret_cast<const float4 *>(&m.buckets[tb])[k];
            __syncthreads();
            if (need) {
                for (int k = 0; k < nb; k++) {
                    const float4 hd = tile[k * 8];
                    unsigned msk = __float_as_uint(hd.w) & 0x7Fu;
                    // the bucket's own cell: every point of it is at least (Chebyshev cell distance - 1) cell edges away
                    int cx, cy, cz;
                    unpack_key(((unsigned long long)__float_as_uint(hd.y) << 32) | (unsigned long long)__float_as_uint(hd.x), cx, cy, cz);
                    const int dcell = max(max(abs(cx - qcx), abs(cy - qcy)), abs(cz - qcz));
                    if (dcell > reach) continue;
                    while (msk) {
                        const int sl = __ffs((int)msk) - 1;
                        msk &= msk - 1u;
                        const float4 e = tile[k * 8 + 1 + sl];
                        const float d2 = calc_dist(qx, qy, qz, e.x, e.y, e.z);
                        const unsigned long long key = ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned)((tb + k) * 8 + 1 + sl);
                        if (key < best) {
                            best = key;
                            reach = nn1_reach(d2, cell_edge);
                        }
                    }
                }
            }
        }
        if (live && best < best0) atomicMin(&nn_key[f], best);
    }
}

constexpr int kNormalEqDoubles = 158;                 // HtH[144], Htr[12], effective count, residual sum
constexpr int kFetchDoubles = kNormalEqDoubles + 2;   // + far_count of the last match pass + feats_down_size: one device->host copy per iteration
constexpr int kEigOffset = 160;                       // eigvals[6], eigvecs[36] (k_eigen6)
constexpr int kResultDoubles = kEigOffset + 42;

// ------------------------------------------------------------------ the iteration loop on the device
// One launch per enqueued iteration, after that iteration's k_residual: laserMapping.cpp:899-918
// (degradation window), :1012-1053 (Kalman update in the reduced form of SURVEY.md 8b), :1054-1063
// (stop branch), :1069-1101 (rematch / convergence control) and :1084-1085 (covariance update).
// One block; the 24 x 24 solve is a Gauss-Jordan elimination in shared
// memory, the SO(3) pieces run on one thread.  Mirrors dlt_host::LaserMapping::process_scan.
namespace iekf {
DLT_D void mat3_mul(const double *a, const double *b, double *r) {
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) r[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
}
DLT_D void mat3T_mul(const double *a, const double *b, double *r) {  // a^T b
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) r[3 * i + j] = a[i] * b[j] + a[3 + i] * b[3 + j] + a[6 + i] * b[6 + j];
}
// Exp(v1, v2, v3), so3_math.h:54-72
DLT_D void exp3(double v1, double v2, double v3, double *R) {
    const double n = sqrt(v1 * v1 + v2 * v2 + v3 * v3);
    if (n > 0.00001) {
        so3_rodrigues(v1 / n, v2 / n, v3 / n, n, R);
    } else {
#pragma unroll
        for (int i = 0; i < 9; i++) R[i] = (i % 4 == 0) ? 1.0 : 0.0;
    }
}
// Log(R), so3_math.h:75-81
DLT_D void log3(const double *R, double *o) {
    const double tr = R[0] + R[4] + R[8];
    const double theta = (tr > 3.0 - 1e-6) ? 0.0 : acos(0.5 * (tr - 1));
    const double K[3] = {R[7] - R[5], R[2] - R[6], R[3] - R[1]};
    const double f = (fabs(theta) < 0.001) ? 0.5 : (0.5 * theta / sin(theta));
    o[0] = K[0] * f;
    o[1] = K[1] * f;
    o[2] = K[2] * f;
}
// a (-) b over the 24-dim error state, common_lib.h:173-187
DLT_D void boxminus(const double *a, const double *b, double *o) {
    double M[9];
    mat3T_mul(b, a, M);
    log3(M, o);
    mat3T_mul(b + 12, a + 12, M);
    log3(M, o + 6);
#pragma unroll
    for (int k = 0; k < 3; k++) {
        o[3 + k] = a[9 + k] - b[9 + k];
        o[9 + k] = a[21 + k] - b[21 + k];
    }
#pragma unroll
    for (int k = 0; k < 12; k++) o[12 + k] = a[24 + k] - b[24 + k];
}
// s (+)= d, common_lib.h:147-158
DLT_D void boxplus_inplace(double *s, const double *d) {
    double E[9], M[9];
    exp3(d[0], d[1], d[2], E);
    mat3_mul(s, E, M);
#pragma unroll
    for (int k = 0; k < 9; k++) s[k] = M[k];
    exp3(d[6], d[7], d[8], E);
    mat3_mul(s + 12, E, M);
#pragma unroll
    for (int k = 0; k < 9; k++) s[12 + k] = M[k];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        s[9 + k] += d[3 + k];
        s[21 + k] += d[9 + k];
    }
#pragma unroll
    for (int k = 0; k < 12; k++) s[24 + k] += d[12 + k];
}
// o = a * s, common_lib.h:190-205 (R_L_I is not scaled)
DLT_D void scaled(const double *a, double s, double *o) {
    double so3[3];
    log3(a, so3);
    exp3(so3[0] * s, so3[1] * s, so3[2] * s, o);
#pragma unroll
    for (int k = 0; k < 9; k++) o[12 + k] = a[12 + k];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        o[9 + k] = a[9 + k] * s;
        o[21 + k] = a[21 + k] * s;
    }
#pragma unroll
    for (int k = 24; k < 36; k++) o[k] = a[k] * s;
}
// a <- a + b (StatesGroup + StatesGroup), common_lib.h:131-144: biases and gravity of the left operand
DLT_D void compose_inplace(double *a, const double *b) {
    double M[9];
    mat3_mul(a, b, M);
#pragma unroll
    for (int k = 0; k < 9; k++) a[k] = M[k];
    mat3_mul(a + 12, b + 12, M);
#pragma unroll
    for (int k = 0; k < 9; k++) a[12 + k] = M[k];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        a[9 + k] += b[9 + k];
        a[21 + k] += b[21 + k];
        a[24 + k] += b[24 + k];
    }
}
}  // namespace iekf

constexpr int kIekfBlock = 256;
// state_propagat: the shared copy, unless the stop branch reused that buffer for the thermal delta
DLT_D const double *sprop_or(const double *shared_copy, const double *global_copy, int stop) { return stop ? global_copy : shared_copy; }

// Kalman gain in covariance (Woodbury) form.  The reference inverts twice, K_1 = (H^T H + (P/R)^-1)^-1
// (laserMapping.cpp:1017-1018); with P' = P/R, U = [I_12; 0] and H^T H = U H U^T the same matrix is
//     K_1 = P' - P' U (I + H P'_11)^-1 H U^T P',     so     K_1[:, :12] = P'_1 - P'_1 Y,   (I + Q) Y = Q,   Q = H P'_11
// (P'_1 = first 12 columns, P'_11 = top-left 12 x 12).  Without extrinsic estimation only the top-left D = 6 block of
// H is non-zero, rows D..11 of Q and Y vanish and the system is D x D with 12 right-hand sides: 6 elimination steps
// instead of two 24 x 24 inversions.  Equal to the reference form up to fp64 round-off (tests: 1e-7 on the solution).
// The step as a block-wide device function (any block size >= 128 that is a multiple of 32): called by the last block
// of k_residual (fused: no launch, no extra round trip for the normal equations) or by k_iekf_step when a reduction
// over ranks has to run in between.
DLT_D void iekf_step_block(IekfDev *dev, const double *__restrict__ result, const int *__restrict__ n_down_ptr,
                           const int *__restrict__ vox_status_ptr, int D /* 6 or 12 */) {
    dlt_iekf_block &c = dev->b;
    constexpr int N = 24;
    __shared__ double PN[N * N];                      // P' = cov / LASER_POINT_COV
    __shared__ double W[12 * 24];                     // [I + Q[:, :D] | Q], later G = K H (24 x 12)
    __shared__ double T[N * 12];                      // K_1[:, :12]
    __shared__ double C12[12 * N];                    // the first 12 rows of the covariance
    __shared__ double R[kNormalEqDoubles + 2];        // the normal equations of this iteration
    __shared__ double st[36], sprop[36], sother[36];  // state, state_propagat | thermal delta, last_nodegared (non-covariance parts)
    __shared__ double vec[N], rhs[12], sol[N], rcpv[12];
    __shared__ int s_i[16];
    __shared__ int s_q[10];
    __shared__ int s_fail, s_piv;
    enum { I_QLEN, I_THR, I_CONV, I_RNUM, I_MAXIT, I_INITED, I_GAIN, I_NDOWN, I_VOX };
    const int tid = threadIdx.x, NT = blockDim.x;
    const int it = c.iter;
    dlt_iekf_iter &rec = c.iters[it];
    const int did_match = (it == 0 || c.rematch_en) ? 1 : 0;
    const int AW = D + 12;
    constexpr int kCovPer = 5;  // covariance elements per thread at the smallest block size (128)

    DLT_STAMP(0);
    // ---- one round trip to global memory for everything the step needs
    for (int k = tid; k < kNormalEqDoubles + 1; k += NT) R[k] = result[k];
    double cov_own[kCovPer];
#pragma unroll
    for (int q = 0; q < kCovPer; q++) {
        const int k = tid + q * NT;
        cov_own[q] = (k < N * N) ? c.state[36 + k] : 0.0;
    }
    const double lpc = c.laser_point_cov;
    if (tid < 36) {
        st[tid] = c.state[tid];
        sprop[tid] = c.state_propagat[tid];
    } else if (tid >= 64 && tid < 74) {
        s_q[tid - 64] = c.effct_queue[tid - 64];
    } else if (tid == 96) {
        s_i[I_QLEN] = c.queue_len;
        s_i[I_THR] = c.threshold;
        s_i[I_CONV] = c.converged;
    } else if (tid == 97) {
        s_i[I_RNUM] = c.rematch_num;
        s_i[I_MAXIT] = c.max_iteration;
        s_i[I_INITED] = c.flg_EKF_inited;
    } else if (tid == 98) {
        s_i[I_GAIN] = c.have_gain;
        s_i[I_NDOWN] = *n_down_ptr;
        s_i[I_VOX] = *vox_status_ptr;
    }
    if (tid == 0) {
        s_fail = 0;
        if (it == 0) c.insert_status = 0;  // map_incremental stays disarmed until the loop ends cleanly
    }
#pragma unroll
    for (int q = 0; q < kCovPer; q++) {
        const int k = tid + q * NT;
        if (k < N * N) {
            PN[k] = cov_own[q] / lpc;
            if (k < 12 * N) C12[k] = cov_own[q];
        }
    }
    __syncthreads();
    DLT_STAMP(1);

    // effct_feat_numQueue, :899-918 -- every thread derives the same decision from shared memory
    const int effct = (int)(R[156] + 0.5);
    int ql = s_i[I_QLEN];
    const int shift = (ql >= 10) ? 1 : 0;
    ql -= shift;
    int stop = (effct <= s_i[I_THR]) ? 1 : 0;
    for (int k = 0; k < ql; k++)
        if (s_q[k + shift] <= s_i[I_THR]) stop = 1;
    if (tid < 10) c.effct_queue[tid] = (tid < ql) ? s_q[tid + shift] : (tid == ql ? effct : 0);
    for (int k = tid; k < 144; k += NT) rec.HtH[k] = R[k];
    if (tid < 12) rec.Htr[tid] = R[144 + tid];
    if (tid < 24) rec.pose_in[tid] = st[tid];
    if (tid == 0) {
        rec.iter = it;
        rec.effct_feat_num = effct;
        rec.total_residual = R[157];
        rec.did_match = did_match;
        rec.reserved = 0;
        rec.ekf_stop = stop;
        c.n_down = s_i[I_NDOWN];
        c.reserved1 = s_i[I_VOX];  // VoxelGrid status of this scan (2 = bitmap capacity exceeded)
        if (did_match) c.n_unresolved = (int)(R[158] + 0.5);
        c.queue_len = ql + 1;
        c.ekf_stop = stop;
    }
    int conv = s_i[I_CONV];
    int have_gain = s_i[I_GAIN];
    int inited = s_i[I_INITED];

    if (!stop) {  // ---- Kalman update, :1012-1053
#pragma unroll
        for (int q = 0; q < kCovPer; q++) {
            const int k = tid + q * NT;
            if (k < N * N) c.last_nodegared[36 + k] = cov_own[q];  // :1050 (the covariance does not change inside the loop)
        }
        for (int k = tid; k < D * 12; k += NT) {  // Q = H P'_11 (rows 0..D-1)
            const int i = k / 12, j = k % 12;
            double qv = 0;
            for (int a = 0; a < D; a++) qv += R[i * 12 + a] * PN[a * N + j];
            W[i * AW + D + j] = qv;
            if (j < D) W[i * AW + j] = qv + ((i == j) ? 1.0 : 0.0);
        }
        if (tid == NT - 32) iekf::boxminus(sprop, st, vec);  // :1028 (lane 0 of the last warp)
        DLT_STAMP(2);
        __syncthreads();
        DLT_STAMP(3);
        // (I + Q[:, :D]) Y = Q by Gauss-Jordan with partial pivoting, the row exchange folded into the update
        for (int col = 0; col < D; col++) {
            if (tid < 32) {
                const int r = col + tid;
                double v = (r < D) ? fabs(W[r * AW + col]) : -1.0;
                int idx = r;
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) {  // D <= 12 rows: 16 lanes suffice
                    const double ov = __shfl_xor_sync(0xffffffffu, v, o);
                    const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
                    if (ov > v || (ov == v && oi < idx)) {
                        v = ov;
                        idx = oi;
                    }
                }
                if (tid == 0) {
                    s_piv = idx;
                    if (!(v > 1e-300) || !(v < 1e300)) s_fail = 1;
                }
            }
            __syncthreads();
            if (s_fail) break;  // block-uniform
            const int p = s_piv;
            const int ncols = AW - 1 - col;
            const double inv = 1.0 / W[p * AW + col];
            double upd[3];
            int dst[3];
#pragma unroll
            for (int q = 0; q < 3; q++) {  // D * ncols <= 12 * 23 elements
                const int k = tid + q * NT;
                dst[q] = -1;
                upd[q] = 0.0;
                if (k < D * ncols) {
                    const int r = k / ncols, j = col + 1 + k % ncols;
                    const int src = (r == col) ? p : (r == p) ? col : r;  // row it comes from (rows p / col exchanged)
                    double v = W[src * AW + j];
                    if (r != col) v = v - (W[src * AW + col] * inv) * W[p * AW + j];
                    upd[q] = v;
                    dst[q] = r * AW + j;
                }
            }
            double dcol = 0.0;  // the pivot moves to the diagonal; the rest of the column is never read again
            if (tid == 0) dcol = W[p * AW + col];
            __syncthreads();
#pragma unroll
            for (int q = 0; q < 3; q++)
                if (dst[q] >= 0) W[dst[q]] = upd[q];
            if (tid == 0) W[col * AW + col] = dcol;
            __syncthreads();
        }
        if (s_fail) {
            if (tid == 0) {
                c.status = 1;
                c.done = 1;
                c.n_iters = it;
            }
            return;
        }
        DLT_STAMP(4);
        if (tid < D) rcpv[tid] = 1.0 / W[tid * AW + tid];
        __syncthreads();
        for (int k = tid; k < N * 12; k += NT) {  // K_1[:, :12] = P'_1 - P'_1[:, :D] Y
            const int i = k / 12, j = k % 12;
            double x = PN[i * N + j];
            for (int a = 0; a < D; a++) x -= PN[i * N + a] * (W[a * AW + D + j] * rcpv[a]);
            T[k] = x;
            dev->K1c[k] = x;
        }
        for (int k = tid; k < 144; k += NT) dev->HtH12[k] = R[k];
        if (tid >= 32 && tid < 44) {  // solution = K_1[:, :12] (H^T r - H^T H vec_12) + vec, :1032
            const int a = tid - 32;
            double sacc = 0;
            for (int b = 0; b < 12; b++) sacc += R[a * 12 + b] * vec[b];
            rhs[a] = R[144 + a] - sacc;
        }
        __syncthreads();
        if (tid < N) {
            double sacc = 0;
            for (int a = 0; a < 12; a++) sacc += T[tid * 12 + a] * rhs[a];
            sol[tid] = sacc + vec[tid];
            rec.solution[tid] = sol[tid];
        }
        __syncthreads();
        DLT_STAMP(5);
        // state (+)= solution, :1033 -- the two rotations on two warps, the vector parts on a third
        if (tid == 0 || tid == 32) {
            const int o = (tid == 0) ? 0 : 12, d = (tid == 0) ? 0 : 6;
            double E[9], M[9];
            iekf::exp3(sol[d], sol[d + 1], sol[d + 2], E);
            iekf::mat3_mul(st + o, E, M);
#pragma unroll
            for (int k = 0; k < 9; k++) st[o + k] = M[k];
        } else if (tid >= 64 && tid < 67) {
            st[9 + tid - 64] += sol[3 + tid - 64];
            st[21 + tid - 64] += sol[9 + tid - 64];
        } else if (tid >= 96 && tid < 108) {
            st[24 + tid - 96] += sol[12 + tid - 96];
        }
        const double rn = sqrt(sol[0] * sol[0] + sol[1] * sol[1] + sol[2] * sol[2]);
        const double tn = sqrt(sol[3] * sol[3] + sol[4] * sol[4] + sol[5] * sol[5]);
        conv = ((rn * 57.3 < 0.01) && (tn * 100 < 0.015)) ? 1 : 0;  // :1040
        have_gain = 1;
        __syncthreads();
        DLT_STAMP(6);
        if (tid < 36) {
            c.state[tid] = st[tid];
            c.last_nodegared[tid] = st[tid];  // :1050
        }
    } else {  // ---- stop branch, :1054-1063: state = last_nodegared_state + odomToStateGruop(delta)
        if (tid < N) rec.solution[tid] = 0.0;
        if (tid < 36) sother[tid] = c.last_nodegared[tid];
        if (tid >= 64 && tid < 100) sprop[tid - 64] = c.thermal_delta[tid - 64];  // state_propagat is not needed on this branch
#pragma unroll
        for (int q = 0; q < kCovPer; q++) {
            const int k = tid + q * NT;
            if (k < N * N) {
                cov_own[q] = c.last_nodegared[36 + k];
                c.state[36 + k] = cov_own[q];
            }
        }
        __syncthreads();  // (block-uniform branch)
        if (tid == 0) {
            const double *L = sother, *Dl = sprop;
            iekf::mat3_mul(L, Dl, st);
            iekf::mat3_mul(L + 12, Dl + 12, st + 12);
            for (int k = 0; k < 3; k++) {
                st[9 + k] = L[9 + k] + Dl[9 + k];
                st[21 + k] = L[21 + k] + Dl[21 + k];
                st[24 + k] = L[24 + k] + Dl[24 + k];
            }
            for (int k = 27; k < 36; k++) st[k] = L[k];  // bias_g, bias_a, gravity of the left operand
        }
        inited = 0;
        __syncthreads();
        if (tid < 36) c.state[tid] = st[tid];
    }
    if (tid < 36) rec.state_out[tid] = st[tid];
    // ---- :1069-1101, every thread derives the same control decisions
    int rematch_en = 0, rematch_num = s_i[I_RNUM];
    if (conv || (rematch_num == 0 && it == s_i[I_MAXIT] - 2)) {
        rematch_en = 1;
        rematch_num++;
    }
    int fin = 0, upd = 0;
    if (rematch_num >= 2 || it == s_i[I_MAXIT] - 1) {
        fin = 1;
        upd = (inited && have_gain) ? 1 : 0;
    } else if (stop) {
        fin = 1;
    }
    if (tid == 0) {
        rec.converged = conv;
        c.converged = conv;
        c.have_gain = have_gain;
        c.flg_EKF_inited = inited;
        c.rematch_en = rematch_en;
        c.rematch_num = rematch_num;
        c.iter = it + 1;
        c.n_iters = it + 1;
    }
    if (fin && upd) {  // G[:, :12] = K_1[:, :12] H^T H;  cov = (I - G) cov, :1084-1085   (block-uniform)
        if (stop) {  // the gain of an earlier iteration: back from global memory (unreachable today: a stop clears inited)
            for (int k = tid; k < N * 12; k += NT) T[k] = dev->K1c[k];
            for (int k = tid; k < 144; k += NT) R[k] = dev->HtH12[k];
        }
        __syncthreads();
        double *G = W;  // 24 x 12
        for (int k = tid; k < N * 12; k += NT) {
            const int i = k / 12, j = k % 12;
            double sacc = 0;
            for (int a = 0; a < 12; a++) sacc += T[i * 12 + a] * R[a * 12 + j];
            G[k] = sacc;
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < kCovPer; q++) {
            const int k = tid + q * NT;
            if (k < N * N) {
                const int i = k / N, j = k % N;
                double sacc = cov_own[q];
                for (int a = 0; a < 12; a++) sacc -= G[i * 12 + a] * C12[a * N + j];
                c.state[36 + k] = sacc;
            }
        }
    }
    DLT_STAMP(7);
    if (fin && c.finish) {  // ---- zeta blend, :1105-1131, then arm map_incremental (block-uniform)
        __syncthreads();
        __shared__ double lastS[36], v1[N], v2[N], bl[36];
        if (tid < 36) lastS[tid] = c.last_state[tid];
        if (tid >= 64 && tid < 100 && c.blend_mode != 1) sother[tid - 64] = c.l2l_state[tid - 64];
        __syncthreads();
        const double zeta_t = c.zeta_t;
        if (c.blend_mode == 1) {
            const double alpha_l = effct / (c.beta * 65536.0);  // Nla, :106
            const double zeta_l = 2.0 / (1.0 + exp(-alpha_l)) - 1;
            double zn = zeta_l / (zeta_l + zeta_t);
            if (c.lidar_cnt_lt_100) zn = 1;
            if (tid == 0) {
                iekf::boxminus(sprop_or(sprop, c.state_propagat, stop), lastS, v1);
                for (int k = 0; k < N; k++) v1[k] *= (1 - zn);
                c.zeta_l = zeta_l;
            } else if (tid == 32) {
                iekf::boxminus(st, lastS, v2);
                for (int k = 0; k < N; k++) v2[k] *= zn;
            }
            __syncthreads();
            if (tid == 0) {  // state = (last_state + v1) + v2, :1119
                for (int k = 0; k < 36; k++) bl[k] = lastS[k];
                iekf::boxplus_inplace(bl, v1);
                iekf::boxplus_inplace(bl, v2);
            }
        } else {
            if (tid == 0) {  // state = (last_state + v1) + l2l * zeta_t_norm, :1122-1127
                const double ztn = zeta_t / (c.zeta_l + zeta_t);
                iekf::boxminus(st, lastS, v1);
                for (int k = 0; k < N; k++) v1[k] *= (1 - ztn);
                for (int k = 0; k < 36; k++) bl[k] = lastS[k];
                iekf::boxplus_inplace(bl, v1);
                double o[36];
                iekf::scaled(sother, ztn, o);
                iekf::compose_inplace(bl, o);
            }
        }
        __syncthreads();
        if (tid < 36) c.blend_state[tid] = bl[tid];
        // :1165: no map update while the EKF is stopped.  Unresolved queries need the exact-neighbour fallback first:
        // when its kernels were not queued (no such queries in the recent scans) the host runs map_incremental itself.
        if (tid == 0) c.insert_status = stop ? 0 : ((c.n_unresolved > 0 && !c.far_enqueued) ? 2 : 1);
    }
    DLT_STAMP(8);
    if (tid == 0 && fin) c.done = 1;
}


__global__ void __launch_bounds__(kIekfBlock) k_iekf_step(IekfDev *dev, const double *__restrict__ result, const int *__restrict__ n_down_ptr,
                                                           const int *__restrict__ vox_status_ptr, int D /* 6 or 12 */) {
    DLT_PDL_WAIT();
    if (dev->b.done) return;  // block-uniform
    iekf_step_block(dev, result, n_down_ptr, vox_status_ptr, D);
}

// ------------------------------------------------------------------ residual / Jacobian / reduction
template <bool EXT>
struct NormalEq {
    static constexpr int D = EXT ? 12 : 6;
    static constexpr int TRI = D * (D + 1) / 2;
    static constexpr int NR = TRI + D + 2;  // upper triangle of H^T H, H^T r, effective count, sum |r|
};
constexpr int kResidBlock = 256;
static_assert(kResidBlock == 256, "the final reduce combines exactly 8 warp slices");

struct MeasureBufs {
    const float4 *down;       // [n] body-frame downsampled scan (x y z intensity)
    const float4 *nbr;        // [n][5]
    const unsigned char *flags;
    float4 *plane;            // [n] cached plane (a b c d)
    float4 *coeff;            // [n] (a b c pd2) of the current iteration (coeffSel_tmpt)
    unsigned char *sel;       // [n] point_selected_surf, persistent across iterations
    unsigned char *eff;       // [n] effective this iteration
    double *partials;         // [blocks][NR]
    unsigned *ticket;
    const int *far_count;     // unresolved queries of the last match pass
    int *unres_count;         // k_knn8 -> k_knn hand-over counter, re-armed here for the next match pass
    double *result;           // kResultDoubles
    // opt-in (DLT_ZEROCOPY=1): the last block also stores the kFetchDoubles result block straight into pinned HOST memory and
    // then the launch's sequence number into a host flag the caller spins on -- no device->host copy, no stream synchronisation
    double *zc_result = nullptr;
    unsigned long long *zc_flag = nullptr;
    unsigned long long zc_seq = 0;
};
DLT_D void zc_publish(const MeasureBufs &mb, const double *R) {  // whole block; R complete and visible to the block
    for (int k = threadIdx.x; k < kFetchDoubles; k += blockDim.x) mb.zc_result[k] = R[k];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) *(volatile unsigned long long *)mb.zc_flag = mb.zc_seq;
}

// One point of the residual pass: plane (fitted on a match iteration, cached otherwise), residual + gates, Jacobian row.
// Returns `effective`; row / meas / absr are only meaningful then.  Writes plane / coeff / sel / eff of the point.
template <bool EXT>
DLT_D bool residual_point(const MeasureBufs &mb, int i, int do_match, const Pose &P, float plane_thr, double (&row)[NormalEq<EXT>::D], double &meas,
                          double &absr) {
    constexpr int D = NormalEq<EXT>::D;
    bool effective = false;
    float4 pb = mb.down[i];
    float wx, wy, wz;
    body_to_world(P, pb.x, pb.y, pb.z, wx, wy, wz);
    bool sel;
    float4 pl;
    if (do_match) {  // :847-863
        sel = (mb.flags[i] & kFlagMatched) != 0;
        pl = make_float4(0.f, 0.f, 0.f, 0.f);
        if (sel) {
            float px[5], py[5], pz[5];
#pragma unroll
            for (int j = 0; j < 5; j++) {
                float4 e = mb.nbr[(size_t)i * kK + j];
                px[j] = e.x;
                py[j] = e.y;
                pz[j] = e.z;
            }
            float pabcd[4];
            sel = esti_plane(pabcd, px, py, pz, plane_thr);
            pl = make_float4(pabcd[0], pabcd[1], pabcd[2], pabcd[3]);
        }
        mb.plane[i] = pl;
    } else {
        sel = mb.sel[i] != 0;
        pl = mb.plane[i];
    }
    bool sel_out = false;
    if (sel) {
        float pd2 = pl.x * wx + pl.y * wy;  // :866
        pd2 = pd2 + pl.z * wz;
        pd2 = pd2 + pl.w;
        double bn = sqrt(((double)pb.x * (double)pb.x + (double)pb.y * (double)pb.y) + (double)pb.z * (double)pb.z);
        float s = (float)(1 - 0.9 * (double)fabsf(pd2) / sqrt(bn));  // :868
        if ((double)s > 0.9) {                                         // :870
            sel_out = true;
            mb.coeff[i] = make_float4(pl.x, pl.y, pl.z, pd2);
            absr = (double)fabsf(pd2);  // res_last, :879
            if (absr <= 2.0) {           // :889
                effective = true;
                // Jacobian row, :948-977
                double tx, ty, tz, Cx, Cy, Cz;
                mat3_vec(P.R_L_I, (double)pb.x, (double)pb.y, (double)pb.z, tx, ty, tz);
                tx += P.T_L_I[0];
                ty += P.T_L_I[1];
                tz += P.T_L_I[2];
                mat3T_vec(P.rot_end, (double)pl.x, (double)pl.y, (double)pl.z, Cx, Cy, Cz);
                row[0] = ty * Cz - tz * Cy;  // [point_this]x * C
                row[1] = tz * Cx - tx * Cz;
                row[2] = tx * Cy - ty * Cx;
                row[3] = (double)pl.x;
                row[4] = (double)pl.y;
                row[5] = (double)pl.z;
                if (EXT) {
                    double ux, uy, uz;  // R_L_I^T * C
                    mat3T_vec(P.R_L_I, Cx, Cy, Cz, ux, uy, uz);
                    double bx = (double)pb.x, by = (double)pb.y, bz = (double)pb.z;
                    row[D - 6] = by * uz - bz * uy;  // [point_this_be]x * R_L_I^T * C
                    row[D - 5] = bz * ux - bx * uz;
                    row[D - 4] = bx * uy - by * ux;
                    row[D - 3] = Cx;
                    row[D - 2] = Cy;
                    row[D - 1] = Cz;
                }
                meas = -(double)pd2;  // :977
            }
        }
    }
    mb.sel[i] = sel_out ? 1 : 0;
    mb.eff[i] = effective ? 1 : 0;
    return effective;
}

// warp shuffle reduction of the NR accumulators of one warp's 32 points into dst[0..NR) (lane 0 stores or adds)
template <bool EXT, bool ACCUM>
DLT_D void residual_warp_reduce(const double (&row)[NormalEq<EXT>::D], double meas, double absr, bool effective, double *dst, int lane) {
    constexpr int D = NormalEq<EXT>::D;
    int k = 0;
#pragma unroll
    for (int a = 0; a < D; a++) {
#pragma unroll
        for (int b = a; b < D; b++) {
            double v = effective ? row[a] * row[b] : 0.0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) dst[k] = ACCUM ? dst[k] + v : v;
            k++;
        }
    }
#pragma unroll
    for (int a = 0; a < D; a++) {
        double v = effective ? row[a] * meas : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) dst[k] = ACCUM ? dst[k] + v : v;
        k++;
    }
    double c = effective ? 1.0 : 0.0, r = effective ? absr : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        c += __shfl_xor_sync(0xffffffffu, c, o);
        r += __shfl_xor_sync(0xffffffffu, r, o);
    }
    if (lane == 0) {
        dst[k] = ACCUM ? dst[k] + c : c;
        dst[k + 1] = ACCUM ? dst[k + 1] + r : r;
    }
}

// Last block of a residual pass (256 threads): fixed-order sum over the per-block partials (deterministic): 8 interleaved
// slices per value (one per warp), each summed in block order with 16 loads in flight, then combined in slice order;
// unpacked into the 12 x 12 result block R together with the scalars that ride along (far_count, feats_down_size).
template <bool EXT>
DLT_D void residual_final_reduce(const MeasureBufs &mb, unsigned n_parts, int n, const LoopArgs &la) {
    using NE = NormalEq<EXT>;
    constexpr int D = NE::D, NR = NE::NR;
    constexpr unsigned S = kResidBlock / 32;
    __shared__ double s_sum[S][32];
    __shared__ double s_fin[NR];
    for (int base = 0; base < NR; base += 32) {  // block-uniform
        const int v = base + (threadIdx.x & 31), sl = threadIdx.x >> 5;
        double acc = 0.0;
        if (v < NR) {
            const double *vp = mb.partials;
            for (unsigned b = sl; b < n_parts; b += S * 16) {  // (adding 0.0 for the padding is exact)
                double t[16];
#pragma unroll
                for (unsigned u = 0; u < 16; u++) t[u] = (b + S * u < n_parts) ? __ldcg(vp + (size_t)(b + S * u) * NR + v) : 0.0;
#pragma unroll
                for (unsigned u = 0; u < 16; u++) acc += t[u];
            }
        }
        s_sum[sl][threadIdx.x & 31] = acc;
        __syncthreads();
        if (threadIdx.x < 32 && v < NR) {
            double f = s_sum[0][threadIdx.x];
#pragma unroll
            for (unsigned w = 1; w < S; w++) f += s_sum[w][threadIdx.x];
            s_fin[v] = f;
        }
        __syncthreads();
    }
    // unpack the upper triangle into the 12x12 block of the result
    double *R = mb.result;
    for (int k = threadIdx.x; k < 144 + 12; k += kResidBlock) R[k] = 0.0;
    __syncthreads();
    if (threadIdx.x < NR) {
        const int k = threadIdx.x;
        if (k < NE::TRI) {  // k-th element of the row-major upper triangle -> (a, b)
            int a = 0, rem = k;
            while (rem >= D - a) {
                rem -= D - a;
                a++;
            }
            const int b = a + rem;
            const double v = s_fin[k];
            R[a * 12 + b] = v;
            R[b * 12 + a] = v;
        } else if (k < NE::TRI + D) {
            R[144 + (k - NE::TRI)] = s_fin[k];
        } else {
            R[156 + (k - NE::TRI - D)] = s_fin[k];
        }
    }
    if (threadIdx.x == 0) {
        R[158] = (double)(*(volatile const int *)mb.far_count);
        R[159] = (la.vox_ptr && *la.vox_ptr == 2) ? -1.0 : (double)n;  // feats_down_size (-1: VoxelGrid capacity exceeded)
        *mb.ticket = 0u;
        *mb.unres_count = 0;
    }
}

template <bool EXT>
__global__ void __launch_bounds__(kResidBlock)
    k_residual(MeasureBufs mb, int n, int do_match, Pose P_param, float plane_thr, LoopArgs la) {
    DLT_PDL_WAIT();
    using NE = NormalEq<EXT>;
    constexpr int D = NE::D, NR = NE::NR;
    __shared__ double s_part[kResidBlock / 32][NR];
    __shared__ int s_last;
    __shared__ Pose sP;
    if (!loop_resolve(la, P_param, &sP, n, &do_match)) return;  // block-uniform
    // the grid may be larger than needed (sized before the host knows n): only the first n_blocks blocks work
    const unsigned n_blocks = (unsigned)((n + kResidBlock - 1) / kResidBlock);
    if (n_blocks == 0u) {  // empty scan: the normal equations are zero
        if (blockIdx.x == 0) {
            for (int k = threadIdx.x; k < kNormalEqDoubles; k += kResidBlock) mb.result[k] = 0.0;
            if (threadIdx.x == 0) {
                mb.result[158] = (double)(*mb.far_count);
                mb.result[159] = (la.vox_ptr && *la.vox_ptr == 2) ? -1.0 : 0.0;
                *mb.unres_count = 0;
            }
            if (mb.zc_result) {  // block-uniform
                __syncthreads();
                zc_publish(mb, mb.result);
            }
        }
        return;
    }
    if (blockIdx.x >= n_blocks) return;
    const Pose &P = sP;
    const int i = blockIdx.x * kResidBlock + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    double row[D];
    double meas = 0.0, absr = 0.0;
    bool effective = false;
#pragma unroll
    for (int d = 0; d < D; d++) row[d] = 0.0;
    if (i < n) effective = residual_point<EXT>(mb, i, do_match, P, plane_thr, row, meas, absr);

    // ---- warp shuffle reduction of the NR accumulators, then shared memory across warps
    residual_warp_reduce<EXT, false>(row, meas, absr, effective, s_part[warp], lane);
    __syncthreads();
    if (threadIdx.x < NR) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < kResidBlock / 32; w++) v += s_part[w][threadIdx.x];
        mb.partials[(size_t)blockIdx.x * NR + threadIdx.x] = v;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned t = atomicAdd(mb.ticket, 1u);
        s_last = (t == n_blocks - 1u) ? 1 : 0;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    residual_final_reduce<EXT>(mb, n_blocks, n, la);
    double *R = mb.result;
    if (la.peer && !peer_allreduce_block(la.peer, R, kNormalEqDoubles)) {  // block-uniform; starts and ends with a barrier
        if (la.ctl && threadIdx.x == 0) const_cast<IekfDev *>(la.ctl)->b.done = 1;  // a peer never posted: end the loop, the host reports it
        return;
    }
    if (mb.zc_result) {  // block-uniform
        __syncthreads();  // thread 0's R[158], R[159] and (with peers) the summed block are visible to all threads
        zc_publish(mb, R);
    }
    if (la.ctl && la.fuse_step) {  // block-uniform: the solve / control step of this iteration, no launch in between
        __syncthreads();           // the block's own global writes to R are visible to it after the barrier
        iekf_step_block(const_cast<IekfDev *>(la.ctl), R, la.n_ptr, la.vox_ptr, D);
#if !defined(DLT_EMU)
        if (la.use_cond && threadIdx.x == 0) {  // (the step's own writes of done / rematch_en by this thread)
            cudaGraphSetConditional((cudaGraphConditionalHandle)la.cond_while, la.ctl->b.done ? 0u : 1u);
            cudaGraphSetConditional((cudaGraphConditionalHandle)la.cond_match, la.ctl->b.rematch_en ? 1u : 0u);
        }
#endif
    }
}

// ------------------------------------------------------------------ degeneracy: eigen-decomposition of H^T H [0:6, 0:6]
// New output (the reference has no eigen check).  Parallel cyclic Jacobi on one warp:
// the 15 index pairs of a sweep are visited in 5 rounds of 3 disjoint pairs (round-robin
// tournament); lanes 0-2 compute the three rotations of a round, then 18 lanes apply them to
// the columns / rows of A and the columns of V.  Ascending eigenvalues, eigenvectors in columns.
DLT_D void eigen6_warp(const double *__restrict__ result, double *__restrict__ eig_out /* eigvals[6], eigvecs[36] */, int lane) {
    __shared__ double A[36], V[36], cs[3][2];
    __shared__ int pq[3][2];
    for (int i = lane; i < 36; i += 32) {
        A[i] = __ldcg(result + (i / 6) * 12 + (i % 6));  // (inside the persistent loop kernel another block wrote it)
        V[i] = (i % 7 == 0) ? 1.0 : 0.0;
    }
    __syncwarp();
    for (int sweep = 0; sweep < 30; sweep++) {
        double off = 0.0, diag = 0.0;
        for (int i = 0; i < 36; i++) {
            double v = A[i] * A[i];
            if (i % 7 == 0)
                diag += v;
            else
                off += v;
        }
        if (off <= 1e-26 * diag || off == 0.0) break;  // warp-uniform (same data in every lane)
        for (int round = 0; round < 5; round++) {
            // round-robin tournament on players {0..5}: player 5 fixed, others rotate
            if (lane < 3) {
                int a = (lane == 0) ? 5 : (round + lane) % 5;
                int b = (lane == 0) ? round % 5 : (round + 5 - lane) % 5;
                int p = a < b ? a : b, q = a < b ? b : a;
                pq[lane][0] = p;
                pq[lane][1] = q;
                double apq = A[p * 6 + q];
                double c = 1.0, sn = 0.0;
                if (apq != 0.0) {
                    double theta = (A[q * 6 + q] - A[p * 6 + p]) / (2.0 * apq);
                    double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                    c = 1.0 / sqrt(t * t + 1.0);
                    sn = t * c;
                }
                cs[lane][0] = c;
                cs[lane][1] = sn;
            }
            __syncwarp();
            // columns: A <- A J (and V <- V J); lane = pair * 6 + k
            if (lane < 18) {
                const int pr = lane / 6, k = lane % 6;
                const int p = pq[pr][0], q = pq[pr][1];
                const double c = cs[pr][0], sn = cs[pr][1];
                double akp = A[k * 6 + p], akq = A[k * 6 + q];
                A[k * 6 + p] = c * akp - sn * akq;
                A[k * 6 + q] = sn * akp + c * akq;
                double vkp = V[k * 6 + p], vkq = V[k * 6 + q];
                V[k * 6 + p] = c * vkp - sn * vkq;
                V[k * 6 + q] = sn * vkp + c * vkq;
            }
            __syncwarp();
            // rows: A <- J^T A
            if (lane < 18) {
                const int pr = lane / 6, k = lane % 6;
                const int p = pq[pr][0], q = pq[pr][1];
                const double c = cs[pr][0], sn = cs[pr][1];
                double apk = A[p * 6 + k], aqk = A[q * 6 + k];
                A[p * 6 + k] = c * apk - sn * aqk;
                A[q * 6 + k] = sn * apk + c * aqk;
            }
            __syncwarp();
        }
    }
    if (lane == 0) {
        double ev[6];
        int ord[6];
        for (int i = 0; i < 6; i++) {
            ev[i] = A[i * 6 + i];
            ord[i] = i;
        }
        for (int i = 0; i < 5; i++) {
            int mn = i;
            for (int j = i + 1; j < 6; j++)
                if (ev[ord[j]] < ev[ord[mn]]) mn = j;
            int t = ord[i];
            ord[i] = ord[mn];
            ord[mn] = t;
        }
        for (int i = 0; i < 6; i++) {
            eig_out[i] = ev[ord[i]];
            for (int k = 0; k < 6; k++) eig_out[6 + k * 6 + i] = V[k * 6 + ord[i]];
        }
    }
}
__global__ void __launch_bounds__(32) k_eigen6(const double *__restrict__ result, double *__restrict__ eig_out, double *zc_out, unsigned long long *zc_flag,
                                               unsigned long long zc_seq) {
    DLT_PDL_WAIT();
    eigen6_warp(result, eig_out, threadIdx.x);
    if (zc_out) {  // publish into pinned host memory, then raise the flag dlt_degeneracy spins on
        __syncwarp();
        for (int k = threadIdx.x; k < 42; k += 32) zc_out[k] = eig_out[k];
        __threadfence_system();
        __syncwarp();
        if (threadIdx.x == 0) *(volatile unsigned long long *)zc_flag = zc_seq;
    }
}

// ------------------------------------------------------------------ map_incremental classification
// laserMapping.cpp:582-630: decide per downsampled point whether it is added raw
// (PointNoNeedDownsample), through downsample-on-insert (PointToAdd) or dropped.
__global__ void k_incr_classify(const float4 *__restrict__ down, int n, Pose P_param, const float4 *__restrict__ nbr, const int *__restrict__ nbr_cnt,
                                double fs, int ekf_inited, float4 *__restrict__ pw, unsigned char *__restrict__ ds_flag,
                                unsigned char *__restrict__ add_flag, int *__restrict__ class_counts /* [0] downsample adds, [1] raw adds */,
                                LoopArgs la, MapView m, const int *__restrict__ live_ptr, const unsigned char *__restrict__ flags,
                                const int *__restrict__ nn_pos, const unsigned long long *__restrict__ nn_key, FuseInsert fi) {
    DLT_PDL_WAIT();
    __shared__ Pose sP;
    if (!insert_gate(la, n)) return;  // block-uniform
    if (la.ctl) {  // pose after the zeta blend, flg_EKF_inited after the loop
        if (threadIdx.x < 24) reinterpret_cast<double *>(&sP)[threadIdx.x] = la.ctl->b.blend_state[threadIdx.x];
        ekf_inited = la.ctl->b.flg_EKF_inited;
    } else {
        if (threadIdx.x < 24) reinterpret_cast<double *>(&sP)[threadIdx.x] = reinterpret_cast<const double *>(&P_param)[threadIdx.x];
    }
    __syncthreads();
    const Pose &P = sP;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned char ds = 0, add = 0;
    if (i < n) {
    float4 pb = down[i];
    float wx, wy, wz;
    body_to_world(P, pb.x, pb.y, pb.z, wx, wy, wz);  // :591
    pw[i] = make_float4(wx, wy, wz, pb.w);
    // Nearest_Points[i].size(): what the reference's search returned = min(5, live map points).  The rings may have seen
    // fewer (unresolved query): the unseen ones are farther than the 7^3 block's faces, too far to decide anything below.
    int seen = nbr_cnt[i];
    const int live_pts = *live_ptr;
    int cnt = live_pts < kK ? (live_pts < seen ? seen : live_pts) : kK;
    float4 n0 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (seen > 0) n0 = nbr[(size_t)i * kK];
    const bool foreign = flags && (flags[i] & kFlagForeign);  // another shard owns this query and decides for it (k_incr_pack / unpack)
    if (!flags || foreign) {  // no match pass behind this scan: Nearest_Points empty
        cnt = 0;
        seen = 0;
    } else if (flags[i] & kFlagNeedNN) {  // nothing within the block: the true nearest point from k_nn1
        const unsigned long long key = nn_key[nn_pos[i]];
        seen = 0;
        if (key == 0xFFFFFFFFFFFFFFFFull) {
            cnt = 0;
        } else {
            const int id = (int)(unsigned)(key & 0xFFFFFFFFull);
            n0 = m.buckets[id >> 3].pts[(id & 7) - 1];
        }
    }
    ds = 1;  // default: PointToAdd (:621-624)
    if (cnt > 0 && ekf_inited) {    // :593
        float mx = (float)(floor((double)wx / fs) * fs + 0.5 * fs);  // :599-601
        float my = (float)(floor((double)wy / fs) * fs + 0.5 * fs);
        float mz = (float)(floor((double)wz / fs) * fs + 0.5 * fs);
        float dist = calc_dist(wx, wy, wz, mx, my, mz);  // :602
        if ((double)fabsf(n0.x - mx) > 0.5 * fs && (double)fabsf(n0.y - my) > 0.5 * fs && (double)fabsf(n0.z - mz) > 0.5 * fs) {  // :603
            ds = 0;
            add = 1;  // PointNoNeedDownsample
        } else {
            bool need_add = true;
            if (cnt >= kK) {  // :610
                for (int j = 0; j < seen; j++) {
                    float4 e = nbr[(size_t)i * kK + j];
                    if (calc_dist(e.x, e.y, e.z, mx, my, mz) < dist) {  // :612
                        need_add = false;
                        break;
                    }
                }
            }
            ds = need_add ? 1 : 0;
        }
    }
    if (foreign) ds = 0;
    ds_flag[i] = ds;
    add_flag[i] = add;
    if (fi.on) {  // k_map_claim + k_ds_bid for this point
        int cs = -1, vs = -1;
        if (ds || add) {
            int cx, cy, cz;
            cell_of_point(m, wx, wy, wz, cx, cy, cz);
            cs = map_claim(m, pack_key(cx, cy, cz));
        }
        if (ds) vs = ds_bid_point(m, fi.sc, make_float4(wx, wy, wz, pb.w), i);
        fi.cell_slot[i] = cs;
        fi.vslot[i] = vs;
    }
    }
    unsigned bd = __ballot_sync(0xffffffffu, ds != 0), ba = __ballot_sync(0xffffffffu, add != 0);
    if ((threadIdx.x & 31) == 0) {
        if (bd) atomicAdd(&class_counts[0], __popc(bd));
        if (ba) atomicAdd(&class_counts[1], __popc(ba));
    }
}

// Sharded map: every rank classifies the queries it owns; the decisions (0 drop, 1 PointToAdd, 2 PointNoNeedDownsample) are
// exchanged as doubles so that the same sum all-reduce that carries the normal equations can carry them (one owner per
// query, so the sum IS the decision); afterwards every rank knows every decision and inserts what falls into its tiles + halo.
__global__ void k_incr_pack(const unsigned char *__restrict__ ds_flag, const unsigned char *__restrict__ add_flag, int n, double *__restrict__ buf) {
    DLT_PDL_WAIT();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) buf[i] = ds_flag[i] ? 1.0 : (add_flag[i] ? 2.0 : 0.0);
}
__global__ void k_incr_unpack(const double *__restrict__ buf, int n, unsigned char *__restrict__ ds_flag, unsigned char *__restrict__ add_flag,
                              int *__restrict__ class_counts) {
    DLT_PDL_WAIT();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int code = 0;
    if (i < n) {
        code = (int)(buf[i] + 0.5);
        ds_flag[i] = code == 1 ? 1 : 0;
        add_flag[i] = code == 2 ? 1 : 0;
    }
    unsigned bd = __ballot_sync(0xffffffffu, code == 1), ba = __ballot_sync(0xffffffffu, code == 2);
    if ((threadIdx.x & 31) == 0) {
        if (bd) atomicAdd(&class_counts[0], __popc(bd));
        if (ba) atomicAdd(&class_counts[1], __popc(ba));
    }
}

// pointBodyToWorld over the downsampled scan (laserMapping.cpp:786-789), every point flagged for a raw add
__global__ void k_scan_to_world(const float4 *__restrict__ down, int n, Pose P, float4 *__restrict__ pw, unsigned char *__restrict__ ds_flag,
                                unsigned char *__restrict__ add_flag) {
    DLT_PDL_WAIT();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 pb = down[i];
    float wx, wy, wz;
    body_to_world(P, pb.x, pb.y, pb.z, wx, wy, wz);
    pw[i] = make_float4(wx, wy, wz, pb.w);
    ds_flag[i] = 0;
    add_flag[i] = 1;
}

}  // namespace dlt
