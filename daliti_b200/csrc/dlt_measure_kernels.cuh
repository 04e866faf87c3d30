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
#include "dlt_common.cuh"
#include "dlt_map_kernels.cuh"

namespace dlt {

// per-point flag bits
constexpr unsigned char kFlagMatched = 1;     // 5 neighbours found and d2[4] <= max_sq_dist
constexpr unsigned char kFlagUnresolved = 2;  // ring-3 search could not prove exactness (far / crowded)
constexpr unsigned char kFlagForeign = 4;     // query owned by another shard

#ifndef DLT_KNN_MINBLOCKS
#define DLT_KNN_MINBLOCKS 8
#endif
#ifndef DLT_KNN_BATCH
#define DLT_KNN_BATCH 1  // bucket load steps (4 buckets each) in flight per warp
#endif
constexpr int kKnnWarps = 4;
constexpr int kCandMax = 128;
constexpr int kWlMax = 96;

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
    int *far_list;         // indices of unresolved queries
    int *far_count;
};

// distance from q to the slab of cell c along one axis (conservative: shrunk by a rounding slack)
DLT_D float axis_gap(float q, int c, float cell_edge, float slack) {
    float lo = (float)c * cell_edge, hi = (float)(c + 1) * cell_edge;
    float g = fmaxf(fmaxf(lo - q, q - hi), 0.f) - slack;
    return g > 0.f ? g : 0.f;
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
        out.flags[qi] = fl;
        if (!resolved) {
            int pos = atomicAdd(out.far_count, 1);
            out.far_list[pos] = qi;
        }
    }
}

// Warp-per-query kernel.  list == nullptr: queries 0..n-1, one per warp.  Otherwise the queries list[0..*count)
// (the ones k_knn_ring1 could not resolve) with a warp-stride loop.
__global__ void __launch_bounds__(kKnnWarps * 32, DLT_KNN_MINBLOCKS)
    k_knn(MapView m, const float4 *__restrict__ q_pts, int n, int body_frame, Pose P, float max_sq_dist, KnnOut out, const int *__restrict__ list,
          const int *__restrict__ count) {
    __shared__ float4 s_cand[kKnnWarps][kCandSlots];
    __shared__ int s_cid[kKnnWarps][kCandSlots];
    __shared__ int s_wl[kKnnWarps][kWlMax];
    __shared__ Cand s_bestw[kKnnWarps][kK];
    __shared__ int s_nb[kKnnWarps];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total_warps = gridDim.x * kKnnWarps;
    const int limit = list ? *count : n;
    for (int w = blockIdx.x * kKnnWarps + warp; w < limit; w += total_warps) {  // warp-uniform; no block-level barrier
        const int qi = list ? list[w] : w;
        knn_warp_query(m, q_pts, qi, body_frame, P, max_sq_dist, out, s_cand[warp], s_cid[warp], s_wl[warp], s_bestw[warp], &s_nb[warp], lane);
        __syncwarp();
    }
}

// ------------------------------------------------------------------ first pass: 8 lanes per query over the 3x3x3 block
// Four queries per warp.  The 8 lanes of a group probe the 27 cells in 4 rounds, fetch one 128-byte bucket
// per step (lane 0 the header, lanes 1-7 one point each) and every lane keeps the 5 best of the points it
// saw in registers (stated order d2, x, y, z, id); the 8 sorted lists are merged with five 8-lane argmins.
// A query is finished here when the 3^3 block proves its 5 best exact (the common case on a mapped
// surface); the rest go to `unres_list` for the warp-per-query kernel above, which widens the search.
constexpr int kKnn8Block = 128;

DLT_D Cand group8_min_cand(Cand c) {
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
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

__global__ void __launch_bounds__(kKnn8Block)
    k_knn8(MapView m, const float4 *__restrict__ q_pts, int n, int body_frame, Pose P, float max_sq_dist, KnnOut out, int *__restrict__ unres_list,
           int *__restrict__ unres_count) {
    const unsigned FULL = 0xffffffffu;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int grp = lane >> 3, sub = lane & 7;
    const int q0 = (blockIdx.x * (kKnn8Block / 32) + warp) * 4;
    if (q0 >= n) return;  // warp-uniform
    const int qi = q0 + grp;
    const bool live = qi < n;
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
    Cand best[kK];
#pragma unroll
    for (int t = 0; t < kK; t++) best[t] = cand_inf();

    for (int r = 0; r < 4; r++) {
        const int ci = r * 8 + sub;
        int b = -1;
        if (work && ci < 27) b = map_find(m, pack_key(cx + (ci % 3) - 1, cy + ((ci / 3) % 3) - 1, cz + (ci / 9) - 1));
        unsigned gb = (__ballot_sync(FULL, b >= 0) >> (grp * 8)) & 0xFFu;  // this group's found cells
        while (__any_sync(FULL, gb != 0u)) {                                // warp-uniform
            const int src = gb ? (__ffs((int)gb) - 1) : 0;
            int bb = __shfl_sync(FULL, b, (grp << 3) + src);
            if (!gb) bb = -1;
            gb &= gb - 1u;
            while (__any_sync(FULL, bb >= 0)) {  // the cell's bucket chain
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (bb >= 0) v = reinterpret_cast<const float4 *>(&m.buckets[bb])[sub];  // 8 lanes x 16 B = one line
                const int hdr_next = __shfl_sync(FULL, __float_as_int(v.z), lane & ~7);
                const unsigned hdr_mask = __shfl_sync(FULL, __float_as_uint(v.w), lane & ~7);
                if (bb >= 0 && sub >= 1 && ((hdr_mask >> (sub - 1)) & 1u)) {
                    Cand c;
                    c.d2 = calc_dist(qx, qy, qz, v.x, v.y, v.z);
                    c.x = v.x;
                    c.y = v.y;
                    c.z = v.z;
                    c.id = bb * 8 + sub;
                    topk_insert(best, c);
                }
                bb = (bb >= 0) ? hdr_next : -1;
            }
        }
    }
    // ---- merge the group's 8 sorted lists: five rounds of 8-lane argmin, the owner pops its head
    int nb = 0;
    float d5 = INFINITY;
    Cand mine_out = cand_inf();
#pragma unroll
    for (int t = 0; t < kK; t++) {
        const Cand w = group8_min_cand(best[0]);
        const bool valid = w.d2 < INFINITY;
        if (valid && best[0].id == w.id) {
#pragma unroll
            for (int u = 0; u < kK - 1; u++) best[u] = best[u + 1];
            best[kK - 1] = cand_inf();
        }
        nb += valid ? 1 : 0;
        if (t == kK - 1) d5 = w.d2;
        if (sub == t) {  // lane t of the group keeps result t until the exactness test below
            mine_out = w;
            if (!valid) mine_out.id = -1;
        }
    }
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
    const bool resolved = (nb == kK) && cov > 0.f && d5 < cov * cov * 0.99999f;
    if (!resolved) {
        if (sub == 0) unres_list[atomicAdd(unres_count, 1)] = qi;
        return;
    }
    if (sub < kK) {
        out.nbr[(size_t)qi * kK + sub] = make_float4(mine_out.x, mine_out.y, mine_out.z, mine_out.d2);
        out.nbr_id[(size_t)qi * kK + sub] = mine_out.id;
    }
    if (sub == 0) {
        out.nbr_cnt[qi] = kK;
        out.flags[qi] = (d5 <= max_sq_dist) ? kFlagMatched : (unsigned char)0;
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

// ------------------------------------------------------------------ residual / Jacobian / reduction
template <bool EXT>
struct NormalEq {
    static constexpr int D = EXT ? 12 : 6;
    static constexpr int TRI = D * (D + 1) / 2;
    static constexpr int NR = TRI + D + 2;  // upper triangle of H^T H, H^T r, effective count, sum |r|
};
constexpr int kResidBlock = 128;
constexpr int kNormalEqDoubles = 158;                 // HtH[144], Htr[12], effective count, residual sum
constexpr int kFetchDoubles = kNormalEqDoubles + 1;   // + far_count of the last match pass: one device->host copy per iteration
constexpr int kEigOffset = 160;                       // eigvals[6], eigvecs[36] (k_eigen6)
constexpr int kResultDoubles = kEigOffset + 42;
static_assert(kResidBlock == 128, "the final reduce combines exactly 4 warp slices");

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
    double *result;           // kResultDoubles
};

template <bool EXT>
__global__ void __launch_bounds__(kResidBlock)
    k_residual(MeasureBufs mb, int n, int do_match, Pose P, float plane_thr) {
    using NE = NormalEq<EXT>;
    constexpr int D = NE::D, NR = NE::NR;
    __shared__ double s_part[kResidBlock / 32][NR];
    __shared__ int s_last;
    const int i = blockIdx.x * kResidBlock + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    double row[D];
    double meas = 0.0, absr = 0.0;
    bool effective = false;
#pragma unroll
    for (int d = 0; d < D; d++) row[d] = 0.0;

    if (i < n) {
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
    }

    // ---- warp shuffle reduction of the NR accumulators, then shared memory across warps
    {
        int k = 0;
#pragma unroll
        for (int a = 0; a < D; a++) {
#pragma unroll
            for (int b = a; b < D; b++) {
                double v = effective ? row[a] * row[b] : 0.0;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lane == 0) s_part[warp][k] = v;
                k++;
            }
        }
#pragma unroll
        for (int a = 0; a < D; a++) {
            double v = effective ? row[a] * meas : 0.0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) s_part[warp][k] = v;
            k++;
        }
        double c = effective ? 1.0 : 0.0, r = effective ? absr : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            c += __shfl_xor_sync(0xffffffffu, c, o);
            r += __shfl_xor_sync(0xffffffffu, r, o);
        }
        if (lane == 0) {
            s_part[warp][k] = c;
            s_part[warp][k + 1] = r;
        }
    }
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
        s_last = (t == gridDim.x - 1) ? 1 : 0;
    }
    __syncthreads();
    if (!s_last) return;
    // ---- last block: fixed-order sum over blocks (deterministic): 4 interleaved slices per value,
    //      each summed in block order, then combined ((s0+s1)+s2)+s3
    __threadfence();
    __shared__ double s_sum[4][32];
    __shared__ double s_fin[NR];
    for (int base = 0; base < NR; base += 32) {  // block-uniform
        const int v = base + (threadIdx.x & 31), sl = threadIdx.x >> 5;
        double acc = 0.0;
        if (v < NR) {
            // 8 loads in flight per thread; the adds stay in block order (adding 0.0 for the padding is exact)
            const double *vp = mb.partials;
            constexpr unsigned S = kResidBlock / 32;
            for (unsigned b = sl; b < gridDim.x; b += S * 8) {
                double t[8];
#pragma unroll
                for (unsigned u = 0; u < 8; u++) t[u] = (b + S * u < gridDim.x) ? __ldcg(vp + (size_t)(b + S * u) * NR + v) : 0.0;
#pragma unroll
                for (unsigned u = 0; u < 8; u++) acc += t[u];
            }
        }
        s_sum[sl][threadIdx.x & 31] = acc;
        __syncthreads();
        if (threadIdx.x < 32 && v < NR)
            s_fin[v] = ((s_sum[0][threadIdx.x] + s_sum[1][threadIdx.x]) + s_sum[2][threadIdx.x]) + s_sum[3][threadIdx.x];
        __syncthreads();
    }
    // unpack the upper triangle into the 12x12 block of the result
    double *R = mb.result;
    for (int k = threadIdx.x; k < 144 + 12; k += kResidBlock) R[k] = 0.0;
    __syncthreads();
    if (threadIdx.x == 0) {
        int k = 0;
        for (int a = 0; a < D; a++)
            for (int b = a; b < D; b++) {
                double v = s_fin[k++];
                R[a * 12 + b] = v;
                R[b * 12 + a] = v;
            }
        for (int a = 0; a < D; a++) R[144 + a] = s_fin[k++];
        R[156] = s_fin[k];
        R[157] = s_fin[k + 1];
        R[158] = (double)(*mb.far_count);
        *mb.ticket = 0u;
    }
}

// ------------------------------------------------------------------ degeneracy: eigen-decomposition of H^T H [0:6, 0:6]
// New output (the reference has no eigen check).  Parallel cyclic Jacobi on one warp:
// the 15 index pairs of a sweep are visited in 5 rounds of 3 disjoint pairs (round-robin
// tournament); lanes 0-2 compute the three rotations of a round, then 18 lanes apply them to
// the columns / rows of A and the columns of V.  Ascending eigenvalues, eigenvectors in columns.
__global__ void __launch_bounds__(32) k_eigen6(double *__restrict__ result) {
    __shared__ double A[36], V[36], cs[3][2];
    __shared__ int pq[3][2];
    const int lane = threadIdx.x;
    for (int i = lane; i < 36; i += 32) {
        A[i] = result[(i / 6) * 12 + (i % 6)];
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
            result[kEigOffset + i] = ev[ord[i]];
            for (int k = 0; k < 6; k++) result[kEigOffset + 6 + k * 6 + i] = V[k * 6 + ord[i]];
        }
    }
}

// ------------------------------------------------------------------ map_incremental classification
// laserMapping.cpp:582-630: decide per downsampled point whether it is added raw
// (PointNoNeedDownsample), through downsample-on-insert (PointToAdd) or dropped.
__global__ void k_incr_classify(const float4 *__restrict__ down, int n, Pose P, const float4 *__restrict__ nbr, const int *__restrict__ nbr_cnt,
                                double fs, int ekf_inited, float4 *__restrict__ pw, unsigned char *__restrict__ ds_flag,
                                unsigned char *__restrict__ add_flag, int *__restrict__ class_counts /* [0] downsample adds, [1] raw adds */) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned char ds = 0, add = 0;
    if (i < n) {
    float4 pb = down[i];
    float wx, wy, wz;
    body_to_world(P, pb.x, pb.y, pb.z, wx, wy, wz);  // :591
    pw[i] = make_float4(wx, wy, wz, pb.w);
    int cnt = nbr_cnt[i];
    ds = 1;  // default: PointToAdd (:621-624)
    if (cnt > 0 && ekf_inited) {    // :593
        float mx = (float)(floor((double)wx / fs) * fs + 0.5 * fs);  // :599-601
        float my = (float)(floor((double)wy / fs) * fs + 0.5 * fs);
        float mz = (float)(floor((double)wz / fs) * fs + 0.5 * fs);
        float dist = calc_dist(wx, wy, wz, mx, my, mz);  // :602
        float4 n0 = nbr[(size_t)i * kK];
        if ((double)fabsf(n0.x - mx) > 0.5 * fs && (double)fabsf(n0.y - my) > 0.5 * fs && (double)fabsf(n0.z - mz) > 0.5 * fs) {  // :603
            ds = 0;
            add = 1;  // PointNoNeedDownsample
        } else {
            bool need_add = true;
            if (cnt >= kK) {  // :610
                for (int j = 0; j < kK; j++) {
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
    ds_flag[i] = ds;
    add_flag[i] = add;
    }
    unsigned bd = __ballot_sync(0xffffffffu, ds != 0), ba = __ballot_sync(0xffffffffu, add != 0);
    if ((threadIdx.x & 31) == 0) {
        if (bd) atomicAdd(&class_counts[0], __popc(bd));
        if (ba) atomicAdd(&class_counts[1], __popc(ba));
    }
}

// pointBodyToWorld over the downsampled scan (laserMapping.cpp:786-789), every point flagged for a raw add
__global__ void k_scan_to_world(const float4 *__restrict__ down, int n, Pose P, float4 *__restrict__ pw, unsigned char *__restrict__ ds_flag,
                                unsigned char *__restrict__ add_flag) {
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
