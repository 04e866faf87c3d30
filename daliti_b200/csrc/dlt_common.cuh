// daliti_b200/csrc/dlt_common.cuh
//
// Shared device/host definitions for the B200 scan-to-map path: the voxel-hash map
// layout, cell/voxel index arithmetic, the float distance, the 5-point plane fit and the
// small fp64 SO(3) helpers.  Everything that decides a bit-exact result (voxel and cell
// assignment, neighbour distances, the plane fit and its gates) lives here in one place
// and is compiled with -fmad=false so that no multiply-add is contracted: the reference
// is an x86-64 -O3 build without -mfma (eskf_lio/CMakeLists.txt:8).
//
// Reference lines are cited as path:line relative to the DaLiTI tree.
#pragma once
#include <cfloat>
#include <cmath>
#include <cstdint>

#if defined(DLT_EMU)
#include "cuda_emu.h"  // tests/emu: kernel-logic emulator (test infrastructure, never shipped)
#else
#include <cuda_runtime.h>
#endif

#define DLT_HD __host__ __device__ __forceinline__
#define DLT_D __device__ __forceinline__

// Programmatic dependent launch: every kernel of this library is launched with the programmatic-stream-serialization
// attribute (dlt_rt.h), so its launch overlaps the tail of the kernel in front of it in the stream; in return the FIRST
// statement of every kernel waits here until that kernel has completed and its writes are visible.  (Without the
// attribute -- graph capture, cooperative launch -- the instruction returns at once.)
#if defined(DLT_EMU)
#define DLT_PDL_WAIT()
#else
#define DLT_PDL_WAIT() asm volatile("griddepcontrol.wait;" ::: "memory")
#endif

namespace dlt {

constexpr int kK = 5;               // NUM_MATCH_POINTS, laserMapping.cpp:77
constexpr int kBucketSlots = 7;     // points per 128-byte bucket
constexpr unsigned long long kEmptyKey = 0xFFFFFFFFFFFFFFFFull;

// ------------------------------------------------------------------ map layout in HBM
// One bucket = one 128-byte line: 16-byte header + 7 points (x, y, z, intensity).
// A search cell (edge = 2^cell_shift * ds_map) owns a chain of buckets.
struct __align__(16) Bucket {
    unsigned long long key;  // packed cell coordinates
    int next;                // next bucket of the chain, -1 = end
    unsigned mask;           // bit s set <=> pts[s] is a live map point
    float4 pts[kBucketSlots];
};
static_assert(sizeof(Bucket) == 128, "bucket must be one 128-byte line");

// Open-addressing table entry, read with one 16-byte load.
struct __align__(16) Slot {
    unsigned long long key;  // kEmptyKey = free
    int bucket;              // head bucket of the cell
    int pad;
};
static_assert(sizeof(Slot) == 16, "slot");

struct MapView {
    Slot *table;
    unsigned table_mask;  // capacity - 1 (power of two)
    Bucket *buckets;
    int bucket_cap;
    int *n_buckets;  // device counter: buckets allocated so far
    int *n_live;     // device counter: live points
    int *error;      // device sticky error flag (1 = bucket pool exhausted, 2 = table full)
    // per run of 64 consecutive buckets of the pool: the box of the search cells those buckets belong to (cell coordinates,
    // min[3] and max[3] arrays; an unused run has min > max).  Buckets are allocated in the order cells are first touched, so the
    // runs are spatially compact; the nearest-point search of far queries (k_nn1) skips runs whose box is too far away.
    int *pool_box_min, *pool_box_max;
    float ds;        // downsample voxel edge (filter_size_map)
    int cell_shift;  // cell edge = ds * 2^cell_shift
    // spatial sharding across GPUs (shard_count == 1: off)
    int shard_rank, shard_count;
    int tile_shift;  // tile edge = cell edge * 2^tile_shift
};

// ------------------------------------------------------------------ index arithmetic
// Downsample voxel index of a coordinate: floor(x / ds) evaluated in float exactly as
// ikd_Tree.cpp:491 does (`floor(PointToAdd[i].x / downsample_size)`).
DLT_HD int voxel_index(float x, float ds) { return (int)floorf(x / ds); }
DLT_HD int cell_of_voxel(int k, int shift) { return k >> shift; }  // arithmetic shift = floor division

DLT_HD unsigned long long pack_key(int cx, int cy, int cz) {
    return ((unsigned long long)((unsigned)cx & 0x1FFFFFu) << 42) | ((unsigned long long)((unsigned)cy & 0x1FFFFFu) << 21) |
           (unsigned long long)((unsigned)cz & 0x1FFFFFu);
}
DLT_HD void unpack_key(unsigned long long k, int &cx, int &cy, int &cz) {  // inverse of pack_key (21-bit two's complement fields)
    cx = ((int)((unsigned)(k >> 42) << 11)) >> 11;
    cy = ((int)((unsigned)(k >> 21) << 11)) >> 11;
    cz = ((int)((unsigned)k << 11)) >> 11;
}
constexpr int kPoolRun = 64;  // buckets per pool run (k_nn1 streams the pool in tiles of this size)
DLT_HD unsigned hash_key(unsigned long long k) {  // murmur3 finaliser
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdull;
    k ^= k >> 33;
    k *= 0xc4ceb9fe1a85ec53ull;
    k ^= k >> 33;
    return (unsigned)k;
}
// owner rank of a cell under spatial sharding: tiles of 2^tile_shift cells, round-robin.
DLT_HD int tile_owner(int cx, int cy, int cz, int tile_shift, int shard_count) {
    unsigned h = hash_key(pack_key(cx >> tile_shift, cy >> tile_shift, cz >> tile_shift));
    return (int)(h % (unsigned)shard_count);
}

// calc_dist: ikd_Tree.cpp:1682-1688, common_lib.h:244-248 -- ((dx*dx)+(dy*dy))+(dz*dz), fp32
DLT_HD float calc_dist(float ax, float ay, float az, float bx, float by, float bz) {
    float dx = ax - bx, dy = ay - by, dz = az - bz;
    float d = dx * dx + dy * dy;
    d = d + dz * dz;
    return d;
}

// Stated neighbour order: (d2 fp32, x, y, z, id).  The reference's order under exact d2
// ties depends on the tree shape (strict '<' at ikd_Tree.cpp:1088,1099).
struct Cand {
    float d2, x, y, z;
    int id;
};
DLT_HD bool cand_less(const Cand &a, const Cand &b) {
    if (a.d2 != b.d2) return a.d2 < b.d2;
    if (a.x != b.x) return a.x < b.x;
    if (a.y != b.y) return a.y < b.y;
    if (a.z != b.z) return a.z < b.z;
    return a.id < b.id;
}

// ------------------------------------------------------------------ esti_plane (fp32)
// common_lib.h:267-299: solve A n = -1 (A = 5 neighbours x 3) with Eigen 3.3.7
// ColPivHouseholderQR<Matrix<float,5,3>>, normalise, reject when any neighbour is farther
// than `threshold` from the plane.  The operation order is fixed here (pivot on the largest
// running column norm, LAPACK-style norm downdate, reflectors applied column by column,
// back substitution in axpy form) and is the one oracle/oracle.hpp states; reductions over
// 4 or 5 elements use the SSE2 packet order ((a0+a2)+(a1+a3))[+a4].
namespace plane_detail {
template <int N>
DLT_HD float ssq(const float *v) {
    if (N >= 4) {
        float a0 = v[0] * v[0], a1 = v[1] * v[1], a2 = v[2] * v[2], a3 = v[3] * v[3];
        float r = (a0 + a2) + (a1 + a3);
#pragma unroll
        for (int i = 4; i < N; i++) r = r + v[i] * v[i];
        return r;
    } else if (N <= 0) {
        return 0.f;
    } else {
        float r = v[0] * v[0];
#pragma unroll
        for (int i = 1; i < N; i++) r = r + v[i] * v[i];
        return r;
    }
}
DLT_HD void swapf(float &a, float &b) {
    float t = a;
    a = b;
    b = t;
}

template <int K>
DLT_HD void qr_step(float (&a)[3][5], float (&nu)[3], float (&nd)[3], float (&hc)[3], int (&transp)[3], int &nzp,
                    float threshold_helper, float ndt) {
    constexpr int rows = 5, cols = 3, TL = rows - K - 1;
    int big = K;
    float bigv = nu[K];
#pragma unroll
    for (int j = K + 1; j < cols; j++)
        if (nu[j] > bigv) {
            bigv = nu[j];
            big = j;
        }
    float big_sq = bigv * bigv;
    if (nzp == 3 && big_sq < threshold_helper * float(rows - K)) nzp = K;
    transp[K] = big;
#pragma unroll
    for (int j = K + 1; j < cols; j++)
        if (big == j) {
#pragma unroll
            for (int r = 0; r < rows; r++) swapf(a[K][r], a[j][r]);
            swapf(nu[K], nu[j]);
            swapf(nd[K], nd[j]);
        }
    // Householder vector of a[K][K..4]
    float tailSq = ssq<TL>(&a[K][K + 1]);
    float c0 = a[K][K];
    float beta, tau;
    if (tailSq <= FLT_MIN) {
        tau = 0.f;
        beta = c0;
#pragma unroll
        for (int i = 0; i < TL; i++) a[K][K + 1 + i] = 0.f;
    } else {
        beta = sqrtf(c0 * c0 + tailSq);
        if (c0 >= 0.f) beta = -beta;
        float den = c0 - beta;
#pragma unroll
        for (int i = 0; i < TL; i++) a[K][K + 1 + i] = a[K][K + 1 + i] / den;
        tau = (beta - c0) / beta;
    }
    hc[K] = tau;
    a[K][K] = beta;
    if (tau != 0.f) {
#pragma unroll
        for (int j = K + 1; j < cols; j++) {
            float tmp = a[K][K + 1] * a[j][K + 1];
#pragma unroll
            for (int i = 1; i < TL; i++) tmp = tmp + a[K][K + 1 + i] * a[j][K + 1 + i];
            tmp = tmp + a[j][K];
            a[j][K] = a[j][K] - tau * tmp;
#pragma unroll
            for (int i = 0; i < TL; i++) a[j][K + 1 + i] = a[j][K + 1 + i] - tmp * (tau * a[K][K + 1 + i]);
        }
    }
#pragma unroll
    for (int j = K + 1; j < cols; j++) {
        if (nu[j] != 0.f) {
            float temp = fabsf(a[j][K]) / nu[j];
            temp = (1.f + temp) * (1.f - temp);
            temp = temp < 0.f ? 0.f : temp;
            float q = nu[j] / nd[j];
            float temp2 = temp * (q * q);
            if (temp2 <= ndt) {
                nd[j] = sqrtf(ssq<TL>(&a[j][K + 1]));
                nu[j] = nd[j];
            } else {
                nu[j] = nu[j] * sqrtf(temp);
            }
        }
    }
}

template <int K>
DLT_HD void apply_qt(const float (&a)[3][5], const float (&hc)[3], float (&c)[5]) {
    constexpr int TL = 5 - K - 1;
    float tau = hc[K];
    if (tau != 0.f) {
        float tmp = a[K][K + 1] * c[K + 1];
#pragma unroll
        for (int i = 1; i < TL; i++) tmp = tmp + a[K][K + 1 + i] * c[K + 1 + i];
        tmp = tmp + c[K];
        c[K] = c[K] - tau * tmp;
#pragma unroll
        for (int i = 0; i < TL; i++) c[K + 1 + i] = c[K + 1 + i] - tmp * (tau * a[K][K + 1 + i]);
    }
}
}  // namespace plane_detail

// px/py/pz: the 5 neighbours.  Returns the reference's bool; pabcd always written.
DLT_HD bool esti_plane(float (&pabcd)[4], const float (&px)[5], const float (&py)[5], const float (&pz)[5], float threshold) {
    using namespace plane_detail;
    float a[3][5];
#pragma unroll
    for (int j = 0; j < 5; j++) {
        a[0][j] = px[j];
        a[1][j] = py[j];
        a[2][j] = pz[j];
    }
    float hc[3], nu[3], nd[3];
    int transp[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        nd[k] = sqrtf(ssq<5>(a[k]));
        nu[k] = nd[k];
    }
    float maxn = nu[0];
    if (nu[1] > maxn) maxn = nu[1];
    if (nu[2] > maxn) maxn = nu[2];
    float th = maxn * FLT_EPSILON;
    float threshold_helper = (th * th) / 5.0f;
    float ndt = sqrtf(FLT_EPSILON);
    int nzp = 3;
    qr_step<0>(a, nu, nd, hc, transp, nzp, threshold_helper, ndt);
    qr_step<1>(a, nu, nd, hc, transp, nzp, threshold_helper, ndt);
    qr_step<2>(a, nu, nd, hc, transp, nzp, threshold_helper, ndt);
    // column permutation: identity with transpositions (k, transp[k]) applied on the right
    int p0 = 0, p1 = 1, p2 = 2;
    if (transp[0] == 1) {
        int t = p0;
        p0 = p1;
        p1 = t;
    } else if (transp[0] == 2) {
        int t = p0;
        p0 = p2;
        p2 = t;
    }
    if (transp[1] == 2) {
        int t = p1;
        p1 = p2;
        p2 = t;
    }
    float nv0 = 0.f, nv1 = 0.f, nv2 = 0.f;
    if (nzp > 0) {
        float c[5] = {-1.f, -1.f, -1.f, -1.f, -1.f};
        apply_qt<0>(a, hc, c);
        if (nzp > 1) apply_qt<1>(a, hc, c);
        if (nzp > 2) apply_qt<2>(a, hc, c);
        if (nzp > 2) {
            c[2] = c[2] / a[2][2];
            c[0] = c[0] - c[2] * a[2][0];
            c[1] = c[1] - c[2] * a[2][1];
        }
        if (nzp > 1) {
            c[1] = c[1] / a[1][1];
            c[0] = c[0] - c[1] * a[1][0];
        }
        c[0] = c[0] / a[0][0];
        float s0 = c[0], s1 = (nzp > 1) ? c[1] : 0.f, s2 = (nzp > 2) ? c[2] : 0.f;
        // dst[perm[i]] = c[i]
        nv0 = (p0 == 0) ? s0 : (p1 == 0) ? s1 : s2;
        nv1 = (p0 == 1) ? s0 : (p1 == 1) ? s1 : s2;
        nv2 = (p0 == 2) ? s0 : (p1 == 2) ? s1 : s2;
    }
    float n = sqrtf(nv0 * nv0 + (nv1 * nv1 + nv2 * nv2));
    pabcd[0] = nv0 / n;
    pabcd[1] = nv1 / n;
    pabcd[2] = nv2 / n;
    pabcd[3] = (float)(1.0 / (double)n);
    bool ok = true;
#pragma unroll
    for (int j = 0; j < 5; j++) {
        float v = pabcd[0] * px[j] + pabcd[1] * py[j];
        v = v + pabcd[2] * pz[j];
        v = v + pabcd[3];
        if (fabsf(v) > threshold) ok = false;
    }
    return ok;
}

// ------------------------------------------------------------------ fp64 3x3 helpers (row-major)
struct Pose {  // what crosses the boundary per iteration (laserMapping.cpp:836)
    double rot_end[9];
    double pos_end[3];
    double R_L_I[9];
    double T_L_I[3];
};

DLT_HD void mat3_vec(const double *m, double x, double y, double z, double &ox, double &oy, double &oz) {
    ox = m[0] * x + m[1] * y + m[2] * z;
    oy = m[3] * x + m[4] * y + m[5] * z;
    oz = m[6] * x + m[7] * y + m[8] * z;
}
DLT_HD void mat3T_vec(const double *m, double x, double y, double z, double &ox, double &oy, double &oz) {
    ox = m[0] * x + m[3] * y + m[6] * z;
    oy = m[1] * x + m[4] * y + m[7] * z;
    oz = m[2] * x + m[5] * y + m[8] * z;
}
// p_global = rot_end * (R_L_I * p_body + T_L_I) + pos_end, stored back to float (laserMapping.cpp:835-840)
DLT_HD void body_to_world(const Pose &P, float bx, float by, float bz, float &wx, float &wy, float &wz) {
    double lx, ly, lz, gx, gy, gz;
    mat3_vec(P.R_L_I, (double)bx, (double)by, (double)bz, lx, ly, lz);
    lx += P.T_L_I[0];
    ly += P.T_L_I[1];
    lz += P.T_L_I[2];
    mat3_vec(P.rot_end, lx, ly, lz, gx, gy, gz);
    wx = (float)(gx + P.pos_end[0]);
    wy = (float)(gy + P.pos_end[1]);
    wz = (float)(gz + P.pos_end[2]);
}
// Eye3 + sin(ang) K + (1 - cos(ang)) K K with K = skew(unit axis): so3_math.h:41-48, 62-68
DLT_HD void so3_rodrigues(double ax, double ay, double az, double ang, double *R) {
    double s = sin(ang), c1 = 1.0 - cos(ang);
    double K[9] = {0.0, -az, ay, az, 0.0, -ax, -ay, ax, 0.0};
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) {
            double kk = (c1 * K[3 * i]) * K[j] + (c1 * K[3 * i + 1]) * K[3 + j] + (c1 * K[3 * i + 2]) * K[6 + j];
            R[3 * i + j] = ((i == j) ? 1.0 : 0.0) + K[3 * i + j] * s + kk;
        }
}
// R = Exp(ang_vel, dt): so3_math.h:32-52
DLT_HD void so3_exp(const double *w, double dt, double *R) {
    double n = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
    if (n > 0.0000001) {
        so3_rodrigues(w[0] / n, w[1] / n, w[2] / n, n * dt, R);
    } else {
#pragma unroll
        for (int i = 0; i < 9; i++) R[i] = (i % 4 == 0) ? 1.0 : 0.0;
    }
}

}  // namespace dlt
