// daliti_b200/csrc/dlt_scan_kernels.cuh
//
// Per-scan preparation on the device:
//   (1) IMU-propagated per-point deskew   -- ImuProcess::UndistortPcl, backward part,
//       eskf_lio/src/IMU_Processing.hpp:332-370 (the forward IMU integration over ~20
//       samples is sequential and stays on the host, IMU_Processing.hpp:226-308)
//   (2) pcl::VoxelGrid<PointXYZINormal> with leaf = filter_size_surf -- call sites
//       eskf_lio/src/laserMapping.cpp:703,775-776; PCL 1.10 semantics: bounding box,
//       ijk = floor(p * inv_leaf) - min_b, idx = ijk . (1, div_x, div_x*div_y), one centroid
//       per occupied voxel, output ordered by idx.
// The sort PCL uses is replaced by an occupancy bitmap over the idx space plus a prefix
// popcount (rank = output position, so the output order is PCL's ascending idx without
// sorting), and the per-voxel sums are accumulated in 2^-24 fixed point with 64-bit
// integer atomics: order independent, hence bit-reproducible from run to run.
#pragma once
#include "dlt_common.cuh"

namespace dlt {

// 48-byte pcl::PointXYZINormal as three float4
//   q0 = x y z 1 | q1 = normal_x(time ratio) normal_y(ring) normal_z(span) 0 | q2 = intensity curvature - -
constexpr int kRawStride4 = 3;

struct ImuPoseDev {  // eskf_lio/msg/Pose6D.msg
    double offset_time;
    double acc[3], gyr[3], vel[3], pos[3], rot[9];
};

struct ScanScalars {        // device-side scalars of the current scan
    unsigned bbox_min[3];   // order-preserving uint encoding of float
    unsigned bbox_max[3];
    unsigned long long first_key;  // (ordered normal_x << 32) | index: the point std::sort would put first
    int n_down;
    int vox_status;  // 0 ok, 1 leaf too small (PCL passes the cloud through), 2 bitmap capacity exceeded
    int n_words;
    int pad;
    long long vox_cells;  // size of PCL's voxel-index space for this scan's bounding box (what the bitmap has to cover)
};

DLT_HD unsigned float_to_ordered(float f) {
#if defined(__CUDA_ARCH__)
    unsigned u = __float_as_uint(f);
#else
    unsigned u;
    memcpy(&u, &f, 4);
#endif
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
DLT_HD float ordered_to_float(unsigned u) {
    unsigned v = (u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u;
#if defined(__CUDA_ARCH__)
    return __uint_as_float(v);
#else
    float f;
    memcpy(&f, &v, 4);
    return f;
#endif
}

__global__ void k_scan_reset(ScanScalars *sc) {
    DLT_PDL_WAIT();
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        for (int a = 0; a < 3; a++) {
            sc->bbox_min[a] = 0xFFFFFFFFu;
            sc->bbox_max[a] = 0u;
        }
        sc->first_key = 0xFFFFFFFFFFFFFFFFull;
        sc->n_down = 0;
        sc->vox_status = 0;
        sc->n_words = 0;
        sc->vox_cells = 0;
    }
}

// which raw point has the smallest normal_x (ties: lowest index) -- the point the
// reference's std::sort by normal_x (IMU_Processing.hpp:216) leaves at begin()
__global__ void k_scan_first(const float4 *__restrict__ raw, int n, ScanScalars *sc) {
    DLT_PDL_WAIT();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long key = 0xFFFFFFFFFFFFFFFFull;
    if (i < n) {
        float nx = raw[(size_t)i * kRawStride4 + 1].x;
        key = ((unsigned long long)float_to_ordered(nx) << 32) | (unsigned)i;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
        key = other < key ? other : key;
    }
    if ((threadIdx.x & 31) == 0 && key != 0xFFFFFFFFFFFFFFFFull) atomicMin(&sc->first_key, key);
}

DLT_D void compensate_point(const ImuPoseDev &head, const Pose &st, double t, float &x, float &y, float &z) {
    // IMU_Processing.hpp:347-365
    double dt = t - head.offset_time;
    double E[9], Ri[9];
    so3_exp(head.gyr, dt, E);
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) Ri[3 * i + j] = head.rot[3 * i] * E[j] + head.rot[3 * i + 1] * E[3 + j] + head.rot[3 * i + 2] * E[6 + j];
    double Tx = head.pos[0] + head.vel[0] * dt + head.acc[0] * 0.5 * dt * dt - st.pos_end[0];
    double Ty = head.pos[1] + head.vel[1] * dt + head.acc[1] * 0.5 * dt * dt - st.pos_end[1];
    double Tz = head.pos[2] + head.vel[2] * dt + head.acc[2] * 0.5 * dt * dt - st.pos_end[2];
    double lx, ly, lz, gx, gy, gz, ex, ey, ez, ox, oy, oz;
    mat3_vec(st.R_L_I, (double)x, (double)y, (double)z, lx, ly, lz);
    lx += st.T_L_I[0];
    ly += st.T_L_I[1];
    lz += st.T_L_I[2];
    mat3_vec(Ri, lx, ly, lz, gx, gy, gz);
    gx += Tx;
    gy += Ty;
    gz += Tz;
    mat3T_vec(st.rot_end, gx, gy, gz, ex, ey, ez);
    ex -= st.T_L_I[0];
    ey -= st.T_L_I[1];
    ez -= st.T_L_I[2];
    mat3T_vec(st.R_L_I, ex, ey, ez, ox, oy, oz);
    x = (float)ox;
    y = (float)oy;
    z = (float)oz;
}

constexpr int kDeskewBlock = 128;

// One thread per raw point.  The 48-byte records are staged through shared memory with
// coalesced 16-byte loads.  do_deskew == 0: copy through (no IMU poses).  The bounding box
// of the output (VoxelGrid's getMinMax3D) is reduced in the same pass.
__global__ void __launch_bounds__(kDeskewBlock)
    k_scan_deskew(const float4 *__restrict__ raw, int n, const ImuPoseDev *__restrict__ poses, int n_pose, Pose st, int do_deskew,
                  float4 *__restrict__ undist, ScanScalars *sc) {
    DLT_PDL_WAIT();
    __shared__ float4 stage[kDeskewBlock * kRawStride4];
    __shared__ unsigned s_mn[3], s_mx[3];
    const int base = blockIdx.x * kDeskewBlock;
    const int cnt = min(kDeskewBlock, n - base);
    if (threadIdx.x < 3) {
        s_mn[threadIdx.x] = 0xFFFFFFFFu;
        s_mx[threadIdx.x] = 0u;
    }
    for (int k = threadIdx.x; k < cnt * kRawStride4; k += kDeskewBlock) stage[k] = raw[(size_t)base * kRawStride4 + k];
    __syncthreads();
    const int i = base + threadIdx.x;
    bool live = threadIdx.x < cnt;
    float x = 0.f, y = 0.f, z = 0.f, inten = 0.f;
    if (live) {
        float4 q0 = stage[threadIdx.x * kRawStride4 + 0];
        float4 q1 = stage[threadIdx.x * kRawStride4 + 1];
        float4 q2 = stage[threadIdx.x * kRawStride4 + 2];
        x = q0.x;
        y = q0.y;
        z = q0.z;
        inten = q2.x;
        if (do_deskew && n_pose >= 2) {
            float tf = q1.x * q1.z;  // normal_x * normal_z in fp32 (IMU_Processing.hpp:345)
            double t = (double)tf;
            int head = -1;
            for (int j = n_pose - 2; j >= 0; j--)
                if (t > poses[j].offset_time) {
                    head = j;
                    break;
                }
            if (head >= 0) {
                compensate_point(poses[head], st, t, x, y, z);
                // Reference quirk (IMU_Processing.hpp:345-369): once the backward sweep reaches
                // begin() it `break`s without decrementing, so the first point of the sorted
                // cloud is compensated again by every earlier IMU interval.
                if ((unsigned)(sc->first_key & 0xFFFFFFFFull) == (unsigned)i)
                    for (int j = head - 1; j >= 0; j--) compensate_point(poses[j], st, t, x, y, z);
            }
        }
        undist[i] = make_float4(x, y, z, inten);
        atomicMin(&s_mn[0], float_to_ordered(x));
        atomicMin(&s_mn[1], float_to_ordered(y));
        atomicMin(&s_mn[2], float_to_ordered(z));
        atomicMax(&s_mx[0], float_to_ordered(x));
        atomicMax(&s_mx[1], float_to_ordered(y));
        atomicMax(&s_mx[2], float_to_ordered(z));
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        atomicMin(&sc->bbox_min[threadIdx.x], s_mn[threadIdx.x]);
        atomicMax(&sc->bbox_max[threadIdx.x], s_mx[threadIdx.x]);
    }
}

// bounding box only (input already float4, e.g. dlt_scan_set_points)
__global__ void k_scan_bbox(const float4 *__restrict__ pts, int n, ScanScalars *sc) {
    DLT_PDL_WAIT();
    __shared__ unsigned s_mn[3], s_mx[3];
    if (threadIdx.x < 3) {
        s_mn[threadIdx.x] = 0xFFFFFFFFu;
        s_mx[threadIdx.x] = 0u;
    }
    __syncthreads();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        float4 p = pts[i];
        atomicMin(&s_mn[0], float_to_ordered(p.x));
        atomicMin(&s_mn[1], float_to_ordered(p.y));
        atomicMin(&s_mn[2], float_to_ordered(p.z));
        atomicMax(&s_mx[0], float_to_ordered(p.x));
        atomicMax(&s_mx[1], float_to_ordered(p.y));
        atomicMax(&s_mx[2], float_to_ordered(p.z));
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        atomicMin(&sc->bbox_min[threadIdx.x], s_mn[threadIdx.x]);
        atomicMax(&sc->bbox_max[threadIdx.x], s_mx[threadIdx.x]);
    }
}

// ------------------------------------------------------------------ VoxelGrid
struct VoxGrid {  // derived from the bounding box exactly as PCL does
    float inv;
    int min_b[3];
    int div[3];
    int mul1, mul2;
    long long cells;
    int status;
};
DLT_D VoxGrid vox_grid_of(const ScanScalars *sc, float leaf) {
    VoxGrid g;
    g.inv = 1.0f / leaf;
    float mn[3], mx[3];
    for (int a = 0; a < 3; a++) {
        mn[a] = ordered_to_float(sc->bbox_min[a]);
        mx[a] = ordered_to_float(sc->bbox_max[a]);
    }
    long long dx = (long long)((mx[0] - mn[0]) * g.inv) + 1;
    long long dy = (long long)((mx[1] - mn[1]) * g.inv) + 1;
    long long dz = (long long)((mx[2] - mn[2]) * g.inv) + 1;
    g.status = (dx * dy * dz > 2147483647ll) ? 1 : 0;
    for (int a = 0; a < 3; a++) {
        g.min_b[a] = (int)floorf(mn[a] * g.inv);
        int max_b = (int)floorf(mx[a] * g.inv);
        g.div[a] = max_b - g.min_b[a] + 1;
    }
    g.mul1 = g.div[0];
    g.mul2 = g.div[0] * g.div[1];
    g.cells = (long long)g.div[0] * g.div[1] * g.div[2];
    return g;
}
DLT_D unsigned vox_idx_of(const VoxGrid &g, float x, float y, float z) {
    int i0 = (int)(floorf(x * g.inv) - (float)g.min_b[0]);
    int i1 = (int)(floorf(y * g.inv) - (float)g.min_b[1]);
    int i2 = (int)(floorf(z * g.inv) - (float)g.min_b[2]);
    return (unsigned)(i0 + i1 * g.mul1 + i2 * g.mul2);
}

// mark occupied voxels in the bitmap; remember every point's idx
__global__ void k_vox_mark(const float4 *__restrict__ pts, int n, float leaf, ScanScalars *sc, unsigned *__restrict__ bitmap,
                           long long bitmap_bits, unsigned *__restrict__ vidx) {
    DLT_PDL_WAIT();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    VoxGrid g = vox_grid_of(sc, leaf);
    int status = g.status;
    if (!status && g.cells > bitmap_bits) status = 2;
    if (i == 0) {
        sc->vox_status = status;
        sc->n_words = status ? 0 : (int)((g.cells + 31) >> 5);
        sc->vox_cells = g.cells;
    }
    if (i >= n || status) return;
    float4 p = pts[i];
    unsigned idx = vox_idx_of(g, p.x, p.y, p.z);
    vidx[i] = idx;
    atomicOr(&bitmap[idx >> 5], 1u << (idx & 31u));
}

constexpr int kScanBlock = 256;
constexpr int kScanWordsPerBlock = 1024;  // 4 words per thread

// exclusive scan of the block totals by ONE block of NT threads, total -> n_down
template <int NT>
DLT_D void vox_scan_block_totals(ScanScalars *sc, const unsigned *__restrict__ blksum, unsigned *__restrict__ blkoff) {
    __shared__ unsigned warp_tot2[NT / 32];
    __shared__ unsigned carry_s;
    const int n_blk = (sc->n_words + kScanWordsPerBlock - 1) / kScanWordsPerBlock;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0u;
    __syncthreads();
    for (int base = 0; base < n_blk; base += NT) {  // block-uniform
        int i = base + threadIdx.x;
        unsigned mine = (i < n_blk) ? __ldcg(blksum + i) : 0u;
        unsigned incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) warp_tot2[warp] = incl;
        __syncthreads();
        unsigned woff = 0, total = 0;
        for (int k = 0; k < NT / 32; k++) {
            unsigned t = warp_tot2[k];
            if (k < warp) woff += t;
            total += t;
        }
        unsigned carry = carry_s;
        if (i < n_blk) blkoff[i] = carry + woff + incl - mine;
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + total;
        __syncthreads();
    }
    if (threadIdx.x == 0) sc->n_down = sc->vox_status ? 0 : (int)carry_s;
}

// exclusive prefix popcount inside each 1024-word block + per-block totals
// ... and, by the last block to finish, the exclusive scan of those totals (what used to be a second launch)
__global__ void __launch_bounds__(kScanBlock)
    k_vox_scan1(const unsigned *__restrict__ bitmap, ScanScalars *sc, unsigned *__restrict__ wprefix, unsigned *__restrict__ blksum,
                unsigned *__restrict__ blkoff, unsigned *__restrict__ ticket) {
    DLT_PDL_WAIT();
    __shared__ unsigned warp_tot[kScanBlock / 32];
    __shared__ int s_last;
    const int n_words = sc->n_words;
    const int n_blk = (n_words + kScanWordsPerBlock - 1) / kScanWordsPerBlock;
    if (n_blk == 0) {  // nothing marked (status != 0): n_down = 0 / pass-through is decided downstream
        if (blockIdx.x == 0 && threadIdx.x == 0) sc->n_down = 0;
        return;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // the grid is a fixed two blocks per SM (the bitmap's size is only known on the device): every block strides over the
    // 1024-word chunks of the part of the bitmap that this scan's bounding box covers
    for (int chunk = blockIdx.x; chunk < n_blk; chunk += gridDim.x) {  // block-uniform
        const int w0 = chunk * kScanWordsPerBlock + threadIdx.x * 4;
        unsigned c[4];
#pragma unroll
        for (int k = 0; k < 4; k++) c[k] = (w0 + k < n_words) ? (unsigned)__popc(bitmap[w0 + k]) : 0u;
        unsigned mine = c[0] + c[1] + c[2] + c[3];
        unsigned incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        __syncthreads();  // (warp_tot of the previous chunk has been read)
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        unsigned woff = 0, total = 0;
#pragma unroll
        for (int k = 0; k < kScanBlock / 32; k++) {
            unsigned t = warp_tot[k];
            if (k < warp) woff += t;
            total += t;
        }
        unsigned excl = woff + incl - mine;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (w0 + k < n_words) wprefix[w0 + k] = excl;
            excl += c[k];
        }
        if (threadIdx.x == 0) blksum[chunk] = total;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned t = atomicAdd(ticket, 1u);
        s_last = (t == gridDim.x - 1u) ? 1 : 0;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (threadIdx.x == 0) *ticket = 0u;
    vox_scan_block_totals<kScanBlock>(sc, blksum, blkoff);
}

struct VoxAcc {  // per output voxel: 2^-24 fixed-point sums
    long long *sx, *sy, *sz, *si;
    unsigned *cnt;
    unsigned *idx;  // the voxel's idx (so the finaliser can clear its bitmap word)
};
constexpr double kFix = 16777216.0;  // 2^24

__global__ void k_vox_accum(const float4 *__restrict__ pts, int n, const ScanScalars *sc, const unsigned *__restrict__ bitmap,
                            const unsigned *__restrict__ wprefix, const unsigned *__restrict__ blkoff, const unsigned *__restrict__ vidx,
                            VoxAcc acc, int *__restrict__ voxel_of_point) {
    DLT_PDL_WAIT();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || sc->vox_status) return;
    unsigned idx = vidx[i];
    unsigned w = idx >> 5, bit = idx & 31u;
    unsigned below = bitmap[w] & ((1u << bit) - 1u);
    unsigned rank = blkoff[w / kScanWordsPerBlock] + wprefix[w] + (unsigned)__popc(below);
    float4 p = pts[i];
    atomicAdd((unsigned long long *)&acc.sx[rank], (unsigned long long)__double2ll_rn((double)p.x * kFix));
    atomicAdd((unsigned long long *)&acc.sy[rank], (unsigned long long)__double2ll_rn((double)p.y * kFix));
    atomicAdd((unsigned long long *)&acc.sz[rank], (unsigned long long)__double2ll_rn((double)p.z * kFix));
    atomicAdd((unsigned long long *)&acc.si[rank], (unsigned long long)__double2ll_rn((double)p.w * kFix));
    atomicAdd(&acc.cnt[rank], 1u);
    acc.idx[rank] = idx;
    if (voxel_of_point) voxel_of_point[i] = (int)rank;
}

// centroid per voxel; leaves the accumulators, the bitmap and the bounding box reset for the next scan.
// vox_status == 1 (PCL: leaf too small for the data): output = input.
__global__ void k_vox_final(ScanScalars *sc, VoxAcc acc, unsigned *__restrict__ bitmap, const float4 *__restrict__ pts, float4 *__restrict__ down,
                            int n) {
    DLT_PDL_WAIT();
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    const int status = sc->vox_status;
    const int n_down = sc->n_down;
    if (status == 1) {
        if (v < n) down[v] = pts[v];
    } else if (v < n && v < n_down) {
        double c = (double)acc.cnt[v];
        float4 o;
        o.x = (float)((double)acc.sx[v] / kFix / c);
        o.y = (float)((double)acc.sy[v] / kFix / c);
        o.z = (float)((double)acc.sz[v] / kFix / c);
        o.w = (float)((double)acc.si[v] / kFix / c);
        down[v] = o;
        bitmap[acc.idx[v] >> 5] = 0u;
        acc.sx[v] = 0;
        acc.sy[v] = 0;
        acc.sz[v] = 0;
        acc.si[v] = 0;
        acc.cnt[v] = 0u;
    }
    if (v == 0) {  // every consumer of the bounding box (k_vox_mark) ran before this kernel
        if (status == 1) sc->n_down = n;
        for (int a = 0; a < 3; a++) {
            sc->bbox_min[a] = 0xFFFFFFFFu;
            sc->bbox_max[a] = 0u;
        }
        sc->first_key = 0xFFFFFFFFFFFFFFFFull;
    }
}

}  // namespace dlt
