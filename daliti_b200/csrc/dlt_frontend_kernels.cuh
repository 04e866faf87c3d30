// daliti_b200/csrc/dlt_frontend_kernels.cuh
//
// LiDAR front end with feature_enabled = 0: FeatureExtract::cachePointCloud
// (eskf_lio/src/feature_extract.cpp:264-423) -- sensor-specific unpacking of the PointCloud2 records into
// PointXYZINormal with normal_x = time ratio, normal_y = ring, normal_z = sweep span -- fused with
// samplePointCloud (:425-450): keep every point_filter_num-th point, range gate, order preserved.  The 48-byte
// records land where the deskew kernel reads them, so the sensor cloud crosses PCIe once and the sampled cloud never.
//   k_fe_convert   one thread per candidate (every point_filter_num-th record): convert, decide, block-level scan
//   k_fe_offsets   exclusive scan of the block totals (single block), total -> n_out
//   k_fe_scatter   order-preserving write of the kept records
#pragma once
#include "../../include/daliti_b200.h"
#include "dlt_common.cuh"

namespace dlt {

constexpr int kFeBlock = 256;

struct FeParams {
    dlt_cloud_layout lay;
    int sensor;
    int n_points;          // records in the sensor cloud
    int filter_num;        // point_filter_num
    int n_cand;            // ceil(n_points / filter_num)
    float min_range, max_range;
    double timespan;       // as cachePointCloud computes it for this sensor
    double t0;             // RoboSense: timestamp of the first record
    float vel_time0;       // Velodyne: time of the first record
};

DLT_D float fe_load_f32(const unsigned char *p) { return *reinterpret_cast<const float *>(p); }
DLT_D unsigned fe_load_u32(const unsigned char *p) { return *reinterpret_cast<const unsigned *>(p); }
DLT_D unsigned short fe_load_u16(const unsigned char *p) { return *reinterpret_cast<const unsigned short *>(p); }
DLT_D double fe_load_f64(const unsigned char *p) {  // rsPointXYZIRT::timestamp sits at an 8-byte aligned offset only if the layout says so
    unsigned lo = *reinterpret_cast<const unsigned *>(p), hi = *reinterpret_cast<const unsigned *>(p + 4);
    return __longlong_as_double((long long)(((unsigned long long)hi << 32) | lo));
}

// one sensor record -> the PointXYZINormal fields the path uses; false = dropped before sampling (RoboSense NaN, :369-370)
DLT_D bool fe_convert(const FeParams &P, const unsigned char *rec, float &x, float &y, float &z, float &inten, float &nx, float &ny, float &nz) {
    x = fe_load_f32(rec + P.lay.off_x);
    y = fe_load_f32(rec + P.lay.off_y);
    z = fe_load_f32(rec + P.lay.off_z);
    ny = (float)fe_load_u16(rec + P.lay.off_ring);
    if (P.sensor == DLT_SENSOR_VELODYNE) {  // :279-297, the TEST_LIO_SAM_6AXIS_DATA branch the reference compiles
        inten = fe_load_f32(rec + P.lay.off_intensity);
        const float t = fe_load_f32(rec + P.lay.off_time);
        nz = (float)P.timespan;
        nx = (float)(((double)t + P.timespan) / P.timespan);
    } else if (P.sensor == DLT_SENSOR_LIVOX) {  // :306-323
        inten = fe_load_f32(rec + P.lay.off_intensity);
        const float t = fe_load_f32(rec + P.lay.off_time);
        nz = (float)P.timespan;
        nx = (float)((double)t / P.timespan);
    } else if (P.sensor == DLT_SENSOR_OUSTER) {  // :324-347
        inten = fe_load_f32(rec + P.lay.off_intensity);
        const unsigned t = fe_load_u32(rec + P.lay.off_time);
        nz = (float)(P.timespan * (double)1e-9f);
        nx = (float)((double)t / P.timespan);
    } else {  // DLT_SENSOR_ROBOSENSE, :348-383
        inten = (float)rec[P.lay.off_intensity];
        const double ts = fe_load_f64(rec + P.lay.off_time);
        nz = (float)P.timespan;
        nx = (float)((ts - P.t0) / P.timespan);
        if (!(fabsf(x) <= FLT_MAX) || !(fabsf(y) <= FLT_MAX) || !(fabsf(z) <= FLT_MAX)) return false;  // pcl_isfinite
    }
    return true;
}

// RoboSense drops non-finite records BEFORE sampling, so the i % point_filter_num test of samplePointCloud (:432) applies to
// the index among the surviving records; the host passes those indices through `rs_index` (nullptr for the other sensors,
// whose inputCloud keeps every record).
__global__ void __launch_bounds__(kFeBlock)
    k_fe_convert(FeParams P, const unsigned char *__restrict__ cloud, const int *__restrict__ rs_index, float4 *__restrict__ tmp /* [n_cand][3] */,
                 unsigned char *__restrict__ keep, unsigned *__restrict__ local_pos, unsigned *__restrict__ blk_total) {
    DLT_PDL_WAIT();
    __shared__ unsigned warp_tot[kFeBlock / 32];
    const int j = blockIdx.x * kFeBlock + threadIdx.x;
    int k = 0;
    if (j < P.n_cand) {
        const long long i = rs_index ? (long long)rs_index[j] : (long long)j * P.filter_num;
        const unsigned char *rec = cloud + (size_t)i * P.lay.point_step;
        float x, y, z, inten, nx, ny, nz;
        bool ok = fe_convert(P, rec, x, y, z, inten, nx, ny, nz);
        float r2 = x * x + y * y;  // pointDistance, my_utility.h:76-79 (float sum of squares, correctly rounded sqrt)
        r2 = r2 + z * z;
        const float range = sqrtf(r2);
        if (range < P.min_range || range > P.max_range) ok = false;  // :445
        k = ok ? 1 : 0;
        tmp[(size_t)j * 3 + 0] = make_float4(x, y, z, 1.0f);   // PCL_ADD_POINT4D: data[3] = 1
        tmp[(size_t)j * 3 + 1] = make_float4(nx, ny, nz, 0.f);
        tmp[(size_t)j * 3 + 2] = make_float4(inten, 0.f, 0.f, 0.f);
        keep[j] = (unsigned char)k;
    }
    // block-level exclusive scan of the keep flags
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned bal = __ballot_sync(0xffffffffu, k);
    const unsigned before = (unsigned)__popc(bal & ((1u << lane) - 1u));
    if (lane == 0) warp_tot[warp] = (unsigned)__popc(bal);
    __syncthreads();
    unsigned woff = 0, total = 0;
#pragma unroll
    for (int w = 0; w < kFeBlock / 32; w++) {
        const unsigned t = warp_tot[w];
        if (w < warp) woff += t;
        total += t;
    }
    if (j < P.n_cand) local_pos[j] = woff + before;
    if (threadIdx.x == 0) blk_total[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) k_fe_offsets(const unsigned *__restrict__ blk_total, int n_blk, unsigned *__restrict__ blk_off, int *__restrict__ n_out) {
    DLT_PDL_WAIT();
    __shared__ unsigned warp_tot[32];
    __shared__ unsigned carry_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0u;
    __syncthreads();
    for (int base = 0; base < n_blk; base += 1024) {  // block-uniform
        const int i = base + threadIdx.x;
        const unsigned mine = (i < n_blk) ? blk_total[i] : 0u;
        unsigned incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        unsigned woff = 0, total = 0;
        for (int w = 0; w < 32; w++) {
            const unsigned t = warp_tot[w];
            if (w < warp) woff += t;
            total += t;
        }
        const unsigned carry = carry_s;
        if (i < n_blk) blk_off[i] = carry + woff + incl - mine;
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + total;
        __syncthreads();
    }
    if (threadIdx.x == 0) *n_out = (int)carry_s;
}

__global__ void __launch_bounds__(kFeBlock)
    k_fe_scatter(int n_cand, const float4 *__restrict__ tmp, const unsigned char *__restrict__ keep, const unsigned *__restrict__ local_pos,
                 const unsigned *__restrict__ blk_off, float4 *__restrict__ out /* 48-byte records */, int cap) {
    DLT_PDL_WAIT();
    const int j = blockIdx.x * kFeBlock + threadIdx.x;
    if (j >= n_cand || !keep[j]) return;
    const unsigned o = blk_off[blockIdx.x] + local_pos[j];
    if (o >= (unsigned)cap) return;
    out[(size_t)o * 3 + 0] = tmp[(size_t)j * 3 + 0];
    out[(size_t)o * 3 + 1] = tmp[(size_t)j * 3 + 1];
    out[(size_t)o * 3 + 2] = tmp[(size_t)j * 3 + 2];
}

}  // namespace dlt
