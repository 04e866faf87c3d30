// daliti_b200/csrc/dlt_map_kernels.cuh
//
// Device-resident voxel-hash map that mirrors the CONTENTS of the reference ikd-Tree
// (eskf_lio/include/ikd-Tree/ikd_Tree.cpp) without its structure:
//   Build (no de-duplication)                          ikd_Tree.cpp:408-423
//   Add_Points(downsample=true): sequential per point   ikd_Tree.cpp:477-522
//   Add_Points(downsample=false): raw append            ikd_Tree.cpp:549-554
//   Delete_Point_Boxes: half-open [min,max) boxes        ikd_Tree.cpp:631-658, :796
//   flatten (export of live points)                     ikd_Tree.cpp:1626-1658
//
// Layout: open-addressing table of 16-byte slots keyed by the packed search-cell
// coordinate -> chain of 128-byte buckets (header + 7 float4 points, live bitmask).
// Mutation is phase-separated (claim cells | resolve voxels | append), one kernel per
// phase, so no kernel both scans and rewrites a chain and no thread ever spins.
#pragma once
#include "dlt_common.cuh"

namespace dlt {

DLT_D Slot load_slot(const Slot *p) {
#if defined(DLT_EMU)
    return *p;
#else
    Slot s;
    const ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(p);
    s.key = v.x;
    s.bucket = (int)(unsigned)(v.y & 0xFFFFFFFFull);
    s.pad = 0;
    return s;
#endif
}

// head bucket of a cell, -1 if the cell does not exist
DLT_D int map_find(const MapView &m, unsigned long long key) {
    unsigned h = hash_key(key) & m.table_mask;
    for (unsigned probe = 0; probe <= m.table_mask; probe++) {
        Slot s = load_slot(&m.table[h]);
        if (s.key == key) return s.bucket;
        if (s.key == kEmptyKey) return -1;
        h = (h + 1) & m.table_mask;
    }
    return -1;
}

DLT_D void bucket_init(Bucket *B, unsigned long long key) {
    B->key = key;
    B->next = -1;
    B->mask = 0u;
}
// a bucket of cell `key` was allocated at pool index b: widen the box of its pool run
DLT_D void pool_box_touch(const MapView &m, int b, unsigned long long key) {
    int c[3];
    unpack_key(key, c[0], c[1], c[2]);
    const int r = b / kPoolRun;
#pragma unroll
    for (int a = 0; a < 3; a++) {
        atomicMin(&m.pool_box_min[3 * r + a], c[a]);
        atomicMax(&m.pool_box_max[3 * r + a], c[a]);
    }
}

// Find-or-create the table slot of a cell; returns the slot index (-1: table full).
// The head bucket index written by the creator is only read by LATER kernels.
// created (optional): the caller allocates the head bucket itself (k_map_claim reserves one contiguous range per block, which
// keeps the pool runs of a bulk insert spatially compact); *created <- 1 when this call made the cell.
DLT_D int map_claim(const MapView &m, unsigned long long key, int *created = nullptr) {
    unsigned h = hash_key(key) & m.table_mask;
    for (unsigned probe = 0; probe <= m.table_mask; probe++) {
        unsigned long long old = atomicCAS(&m.table[h].key, kEmptyKey, key);
        if (old == kEmptyKey) {
            if (created) {
                *created = 1;
                return (int)h;
            }
            int b = atomicAdd(m.n_buckets, 1);
            if (b >= m.bucket_cap) {
                atomicExch(m.error, 1);
                b = -1;
            } else {
                bucket_init(&m.buckets[b], key);
                pool_box_touch(m, b, key);
            }
            m.table[h].bucket = b;
            return (int)h;
        }
        if (old == key) return (int)h;
        h = (h + 1) & m.table_mask;
    }
    atomicExch(m.error, 2);
    return -1;
}

// Append one point to the chain starting at bucket b (append phase only).
DLT_D bool map_append(const MapView &m, int b, float4 p) {
    while (b >= 0) {
        Bucket *B = &m.buckets[b];
        unsigned msk = *(volatile unsigned *)&B->mask;
        while ((msk & 0x7Fu) != 0x7Fu) {
            int s = __ffs((int)(~msk & 0x7Fu)) - 1;
            unsigned old = atomicOr(&B->mask, 1u << s);
            if (!(old & (1u << s))) {
                B->pts[s] = p;
                atomicAdd(m.n_live, 1);
                return true;
            }
            msk = old | (1u << s);
        }
        int nb = *(volatile int *)&B->next;
        if (nb < 0) {
            int fresh = atomicAdd(m.n_buckets, 1);
            if (fresh >= m.bucket_cap) {
                atomicExch(m.error, 1);
                return false;
            }
            bucket_init(&m.buckets[fresh], B->key);
            pool_box_touch(m, fresh, B->key);
            __threadfence();
            int prev = atomicCAS(&B->next, -1, fresh);
            nb = (prev == -1) ? fresh : prev;  // a lost race leaves `fresh` empty and unlinked
        }
        b = nb;
    }
    return false;
}

DLT_D void cell_of_point(const MapView &m, float x, float y, float z, int &cx, int &cy, int &cz) {
    cx = cell_of_voxel(voxel_index(x, m.ds), m.cell_shift);
    cy = cell_of_voxel(voxel_index(y, m.ds), m.cell_shift);
    cz = cell_of_voxel(voxel_index(z, m.ds), m.cell_shift);
}

// Under spatial sharding a rank keeps every point whose cell lies within `halo` cells of
// a tile it owns (halo < tile edge, so the 8 corners of the halo cube suffice).
DLT_D bool shard_keeps_cell(const MapView &m, int cx, int cy, int cz, int halo) {
    if (m.shard_count <= 1) return true;
#pragma unroll
    for (int c = 0; c < 8; c++) {
        int ox = (c & 1) ? halo : -halo, oy = (c & 2) ? halo : -halo, oz = (c & 4) ? halo : -halo;
        if (tile_owner(cx + ox, cy + oy, cz + oz, m.tile_shift, m.shard_count) == m.shard_rank) return true;
    }
    return false;
}
constexpr int kShardHalo = 4;  // cells: ring-3 search + 1

// Kernels enqueued behind the device-resident IEKF loop take their point count and a go/no-go flag from device
// memory (both null on the classic path): go == 1 runs, anything else returns at once.
struct InsertGate {
    const int *go;
    const int *n_ptr;
};
DLT_D bool gate_open(const InsertGate &g, int &n) {
    if (!g.go) return true;
    if (*g.go != 1) return false;
    n = *g.n_ptr;
    return true;
}

// ------------------------------------------------------------------ phase 1: claim cells
// Points with neither flag set (dropped by map_incremental) take no part.
// cell_slot[i] <- table slot of the point's cell (-1: not taking part / not kept by this shard).
__global__ void k_map_claim(MapView m, const float4 *__restrict__ pts, int n, const unsigned char *__restrict__ ds_flag,
                            const unsigned char *__restrict__ add_flag, int *__restrict__ cell_slot, int apply_shard_filter, InsertGate gate) {
    DLT_PDL_WAIT();
    __shared__ int s_warp[32];
    __shared__ int s_base;
    if (!gate_open(gate, n)) return;  // block-uniform
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    int slot = -1, created = 0;
    unsigned long long key = 0ull;
    if (i < n && (ds_flag[i] || add_flag[i])) {
        float4 p = pts[i];
        int cx, cy, cz;
        cell_of_point(m, p.x, p.y, p.z, cx, cy, cz);
        if (!(apply_shard_filter && !shard_keeps_cell(m, cx, cy, cz, kShardHalo))) {
            key = pack_key(cx, cy, cz);
            slot = map_claim(m, key, &created);
        }
    }
    if (i < n) cell_slot[i] = slot;
    // the cells this block created get consecutive head buckets: one reservation per block
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = (blockDim.x + 31) >> 5;
    const unsigned bal = __ballot_sync(0xffffffffu, created != 0);
    const int rank_in_warp = __popc(bal & ((1u << lane) - 1u));
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    if (threadIdx.x == 0) {
        int total = 0;
        for (int w = 0; w < n_warps; w++) {
            const int c = s_warp[w];
            s_warp[w] = total;
            total += c;
        }
        s_base = total > 0 ? atomicAdd(m.n_buckets, total) : 0;
    }
    __syncthreads();
    if (created) {
        int b = s_base + s_warp[warp] + rank_in_warp;
        if (b >= m.bucket_cap) {
            atomicExch(m.error, 1);
            b = -1;
        } else {
            bucket_init(&m.buckets[b], key);
            pool_box_touch(m, b, key);
        }
        m.table[slot].bucket = b;
    }
}

// ------------------------------------------------------------------ phase 3: append
__global__ void k_map_append(MapView m, const float4 *__restrict__ pts, int n, const unsigned char *__restrict__ add_flag,
                             const int *__restrict__ cell_slot, InsertGate gate) {
    DLT_PDL_WAIT();
    if (!gate_open(gate, n)) return;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (add_flag && !add_flag[i]) return;
    int s = cell_slot[i];
    if (s < 0) return;
    int b = m.table[s].bucket;
    if (b < 0) return;
    map_append(m, b, pts[i]);
}

// ------------------------------------------------------------------ downsample-on-insert
// Sequential reference semantics (ikd_Tree.cpp:487-522), per incoming point p in order:
//   S = live points in p's voxel box [min,max); winner = argmin dist-to-centre over
//   {p} U S, p winning ties; if |S| > 1 or the winner is p: delete S, add the winner.
// After the first incoming point a voxel therefore holds exactly one point, and each
// later incoming point replaces it iff it is at least as close to the centre.  The
// parallel equivalent: per touched voxel take the incoming point with the smallest
// (dist, then LARGEST index) and resolve it once against S.
struct DsScratch {
    unsigned long long *vkeys;  // voxel key table (kEmptyKey = free)
    unsigned long long *vwin;   // (dist bits << 32) | (0xFFFFFFFF - index), all-ones = none
    unsigned mask;
};

DLT_D void voxel_box(float x, float ds, float &mn, float &mx, float &mid) {
    mn = floorf(x / ds) * ds;                               // ikd_Tree.cpp:491
    mx = mn + ds;                                           // :492
    mid = (float)((double)mn + (double)(mx - mn) / 2.0);   // :497
}

// phase 1b: per ds point, claim a scratch entry for its voxel and bid for it; returns the scratch slot
DLT_D int ds_bid_point(const MapView &m, const DsScratch &sc, float4 p, int i);

__global__ void k_ds_bid(MapView m, DsScratch sc, const float4 *__restrict__ pts, int n, const unsigned char *__restrict__ ds_flag,
                         int *__restrict__ vslot, InsertGate gate) {
    DLT_PDL_WAIT();
    if (!gate_open(gate, n)) return;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    vslot[i] = ds_flag[i] ? ds_bid_point(m, sc, pts[i], i) : -1;
}

DLT_D int ds_bid_point(const MapView &m, const DsScratch &sc, float4 p, int i) {
    float mnx, mxx, mdx, mny, mxy, mdy, mnz, mxz, mdz;
    voxel_box(p.x, m.ds, mnx, mxx, mdx);
    voxel_box(p.y, m.ds, mny, mxy, mdy);
    voxel_box(p.z, m.ds, mnz, mxz, mdz);
    float d = calc_dist(p.x, p.y, p.z, mdx, mdy, mdz);
    unsigned long long vkey = pack_key(voxel_index(p.x, m.ds), voxel_index(p.y, m.ds), voxel_index(p.z, m.ds));
    unsigned h = hash_key(vkey ^ 0x9E3779B97F4A7C15ull) & sc.mask;
    for (unsigned probe = 0; probe <= sc.mask; probe++) {
        unsigned long long old = atomicCAS(&sc.vkeys[h], kEmptyKey, vkey);
        if (old == kEmptyKey || old == vkey) break;
        h = (h + 1) & sc.mask;
    }
    unsigned long long bid = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)i);
    atomicMin(&sc.vwin[h], bid);
    return (int)h;
}

// map_incremental on an unsharded map: the classification kernel claims the cell and bids for the voxel of the point it
// has just classified (no other thread's result is needed for either), which saves two launches per scan
struct FuseInsert {
    int on;
    DsScratch sc;
    int *cell_slot;
    int *vslot;
};

// phase 2: the winning incoming point of each voxel resolves the voxel against the map.
// Clears the live bits of the points it removes; sets add_flag[i] when it must be added.
__global__ void k_ds_resolve(MapView m, DsScratch sc, const float4 *__restrict__ pts, int n, const unsigned char *__restrict__ ds_flag,
                             const int *__restrict__ vslot, const int *__restrict__ cell_slot, unsigned char *__restrict__ add_flag, InsertGate gate) {
    DLT_PDL_WAIT();
    if (!gate_open(gate, n)) return;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (!ds_flag[i]) return;  // raw points keep the add_flag the caller set
    add_flag[i] = 0;
    int vs = vslot[i];
    if (vs < 0 || cell_slot[i] < 0) return;
    unsigned long long win = sc.vwin[vs];
    if ((unsigned)(win & 0xFFFFFFFFull) != 0xFFFFFFFFu - (unsigned)i) return;
    float4 p = pts[i];
    float mnx, mxx, mdx, mny, mxy, mdy, mnz, mxz, mdz;
    voxel_box(p.x, m.ds, mnx, mxx, mdx);
    voxel_box(p.y, m.ds, mny, mxy, mdy);
    voxel_box(p.z, m.ds, mnz, mxz, mdz);
    float dist_in = calc_dist(p.x, p.y, p.z, mdx, mdy, mdz);
    int head = m.table[cell_slot[i]].bucket;
    // pass 1: closest existing point of the voxel (ties: smallest x, y, z)
    int n_in = 0;
    Cand best;
    best.d2 = INFINITY;
    best.x = best.y = best.z = 0.f;
    best.id = -1;
    for (int b = head; b >= 0; b = m.buckets[b].next) {
        const Bucket *B = &m.buckets[b];
        unsigned msk = B->mask;
        for (int s = 0; s < kBucketSlots; s++) {
            if (!((msk >> s) & 1u)) continue;
            float4 e = B->pts[s];
            if (mnx <= e.x && mxx > e.x && mny <= e.y && mxy > e.y && mnz <= e.z && mxz > e.z) {  // ikd_Tree.cpp:1263
                n_in++;
                Cand c;
                c.d2 = calc_dist(e.x, e.y, e.z, mdx, mdy, mdz);
                c.x = e.x;
                c.y = e.y;
                c.z = e.z;
                c.id = b * 8 + s + 1;
                if (best.id < 0 || cand_less(c, best)) best = c;
            }
        }
    }
    bool incoming_wins = !(n_in > 0 && best.d2 < dist_in);  // strict '<' at ikd_Tree.cpp:507
    if (!incoming_wins && n_in <= 1) return;                 // the single existing point stays (ikd_Tree.cpp:515)
    // pass 2: remove the voxel's points (all, or all but the surviving existing one)
    int removed = 0;
    for (int b = head; b >= 0; b = m.buckets[b].next) {
        Bucket *B = &m.buckets[b];
        unsigned msk = B->mask;
        unsigned clr = 0u;
        for (int s = 0; s < kBucketSlots; s++) {
            if (!((msk >> s) & 1u)) continue;
            float4 e = B->pts[s];
            if (mnx <= e.x && mxx > e.x && mny <= e.y && mxy > e.y && mnz <= e.z && mxz > e.z) {
                if (!incoming_wins && (b * 8 + s + 1) == best.id) continue;
                clr |= 1u << s;
            }
        }
        if (clr) {
            atomicAnd(&B->mask, ~clr);
            removed += __popc(clr);
        }
    }
    if (removed) atomicSub(m.n_live, removed);
    if (incoming_wins) add_flag[i] = 1;
}

// ------------------------------------------------------------------ box delete / export
struct BoxSet {
    float mn[8][3];
    float mx[8][3];
    int n;
};

__global__ void k_map_delete_boxes(MapView m, BoxSet boxes, int n_buckets, int *__restrict__ deleted) {
    DLT_PDL_WAIT();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int b = i >> 3, s = (i & 7) - 1;
    int hit = 0;
    if (b < n_buckets && s >= 0) {
        Bucket *B = &m.buckets[b];
        if ((B->mask >> s) & 1u) {
            float4 e = B->pts[s];
            for (int k = 0; k < boxes.n; k++) {
                if (boxes.mn[k][0] <= e.x && boxes.mx[k][0] > e.x && boxes.mn[k][1] <= e.y && boxes.mx[k][1] > e.y &&
                    boxes.mn[k][2] <= e.z && boxes.mx[k][2] > e.z) {  // ikd_Tree.cpp:796
                    hit = 1;
                    break;
                }
            }
            if (hit) atomicAnd(&B->mask, ~(1u << s));
        }
    }
    unsigned bal = __ballot_sync(0xffffffffu, hit);
    if ((threadIdx.x & 31) == 0 && bal) {
        int c = __popc(bal);
        atomicAdd(deleted, c);
        atomicSub(m.n_live, c);
    }
}

__global__ void k_map_export(MapView m, int n_buckets, float4 *__restrict__ out, int cap, int *__restrict__ counter) {
    DLT_PDL_WAIT();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int b = i >> 3, s = (i & 7) - 1;
    if (b >= n_buckets || s < 0) return;
    const Bucket *B = &m.buckets[b];
    if ((B->mask >> s) & 1u) {
        int o = atomicAdd(counter, 1);
        if (o < cap) out[o] = B->pts[s];
    }
}

}  // namespace dlt
