// daliti_b200/csrc/dlt_api.cu -- the C ABI declared in include/daliti_b200.h.
//
// One handle = one CUDA stream + the device-resident state of one LiDAR sequence: the
// voxel-hash map, the current scan (raw / undistorted / downsampled), the neighbour sets
// and cached planes of the current IEKF update.  Host buffers in, host buffers out; every
// kernel is in the three *_kernels.cuh headers.  There is no CPU code path here.
#include <cmath>
#include <cstddef>
#include <cstring>
#include <initializer_list>
#include <string>
#include <vector>

#include "../../include/daliti_b200.h"
#include "dlt_frontend_kernels.cuh"
#include "dlt_map_kernels.cuh"
#include "dlt_measure_kernels.cuh"
#include "dlt_loop_kernels.cuh"
#include "dlt_rt.h"
#include "dlt_scan_kernels.cuh"

using namespace dlt;

std::atomic<unsigned long long> dlt::rt::g_launches{0};
thread_local void (*dlt::rt::t_wait_hook)(void *) = nullptr;
thread_local void *dlt::rt::t_wait_ctx = nullptr;
#if !defined(DLT_EMU)
int dlt::rt::g_pdl = 1;
#endif

namespace {
constexpr int kMaxImuPoses = 512;
constexpr int kProfKinds = 9;  // 0 knn, 1 residual, 2 deskew, 3 voxelgrid, 4 insert, 5 far fallback, 6 k_iekf_step, 7 k_knn8 alone, 8 k_eigen6
struct ProfSpan {
    int kind;
    rt::Event a, b;
};
constexpr int kFarChunk = 4096;
constexpr int kFarSlices = 64;
constexpr int kFarGroupsX = 4;
constexpr int kNn1GroupsX = 2;
}  // namespace

struct dlt_handle_s {
    dlt_config cfg;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    // the degeneracy eigen-decomposition runs on a side stream so that it overlaps the map insert kernels
    cudaStream_t aux_stream = nullptr;
    rt::Event ev_fork, ev_join;
    bool have_aux = false, eig_pending = false;
    bool eig_zc = false;                 // the last k_eigen6 publishes into pinned host memory (flag at h_ints + 34)
    unsigned long long eig_seq = 0;
    cudaStream_t eig_stream = nullptr;
    // map_incremental off the critical path (dlt_map_incremental_async): its kernels run on their own stream behind the last
    // evaluation of the measurement model; the main stream only waits for them where the next scan first touches what they
    // touch (ev_classified: the downsampled scan may be overwritten; ev_inserted: the map and the neighbour buffers)
    cudaStream_t ins_stream = nullptr;
    rt::Event ev_loop_done, ev_classified, ev_inserted;
    bool have_ins = false;
    bool cls_pending = false;       // the main stream has not been ordered behind ev_classified yet
    bool ins_pending = false;       // ... behind ev_inserted
    bool ins_uncollected = false;   // the counters of that insert have not been adopted on the host
    int *h_ins_ints = nullptr;      // pinned: d_counters[0..8) as the insert left them
    int last_ins_ds = 0, last_ins_raw = 0;
    std::string err;
    int cap = 0;  // per-point array capacity (max_scan_points)
    int n_sm = 148;

    // map
    MapView map;
    size_t table_cap = 0;
    int *d_unres = nullptr;     // queries the thread-per-query pass could not resolve (input of the warp-per-query pass)
    int *d_counters = nullptr;  // [8] unresolved-after-ring-1 count; [0] n_buckets [1] n_live [2] error [3] deleted [4] export counter [5] far_count [6] ds adds [7] raw adds
    // scan
    int n_raw = 0, n_down = 0;
    bool have_raw = false, have_down = false, have_match = false;
    bool counters_fresh = false;  // h_ints mirrors d_counters (no map mutation since the last read-back)
    bool nfar_known = false;      // h_last_nfar is the far_count of the last match pass
    int h_last_nfar = 0;
    bool eig_valid = false;       // the eigen block of d_result belongs to its normal equations
    bool sc_clean = true;         // the scan scalars (bounding box, first-point key) are reset
    bool vox_retry = false;       // a scan is being re-evaluated after the VoxelGrid bitmap was grown
    bool map_dead = false;        // the bucket pool / table overflowed: refuse everything but dlt_map_build
    float4 *d_raw = nullptr, *d_undist = nullptr, *d_down = nullptr;
    // double-buffered upload (dlt_scan_prefetch): the next scan's records cross PCIe on the copy stream while this one is processed
    float4 *d_raw_next = nullptr;
    cudaStream_t copy_stream = nullptr;
    rt::Event ev_copy;
    bool have_copy = false;
    const void *prefetched_ptr = nullptr;
    int prefetched_n = -1;
    const void *pf_pending_ptr = nullptr;  // registered by dlt_scan_prefetch, issued behind the current scan's own small uploads
    int pf_pending_n = -1;
    ImuPoseDev *d_poses = nullptr;
    ScanScalars *d_sc = nullptr;
    unsigned *d_bitmap = nullptr, *d_wprefix = nullptr, *d_blksum = nullptr, *d_blkoff = nullptr, *d_vidx = nullptr;
    long long bitmap_bits = 0;
    int n_scan_blocks = 0;
    size_t n_pool_runs = 0;  // runs of kPoolRun buckets (MapView::pool_box_*)
    VoxAcc acc;
    int *d_vop = nullptr;
    // measure
    KnnOut knn;
    float4 *d_plane = nullptr, *d_coeff = nullptr;
    unsigned char *d_sel = nullptr, *d_eff = nullptr;
    double *d_partials = nullptr, *d_result = nullptr;
    unsigned *d_ticket = nullptr;
    Cand *d_far_partial = nullptr;
    // insert
    float4 *d_pw = nullptr;
    unsigned char *d_dsflag = nullptr, *d_addflag = nullptr;
    int *d_cellslot = nullptr, *d_vslot = nullptr;
    DsScratch scratch;
    size_t scratch_cap = 0;
    // device-resident iteration loop (dlt_iekf_update)
    IekfDev *d_iekf = nullptr;
    dlt_iekf_block *h_iekf = nullptr;  // pinned
    bool n_down_on_device = false;     // dlt_scan_downsample_async ran: h->n_down is only an estimate until the next read-back
    int n_down_hint = 0;               // feats_down_size of the previous scan (grid sizing for the speculative launches)
    // sharded map_incremental: exchange of the per-point decisions
    dlt_reduce_fn shard_reduce = nullptr;
    void *shard_reduce_ctx = nullptr;
    double *d_flagbuf = nullptr;
    // sharded map: exchange through peer mailboxes (dlt_peer_*, dlt_peer.cuh)
    PeerBox *peer_box = nullptr;       // this rank's mailbox (IPC-exportable allocation)
    size_t peer_bytes = 0;
    rt::ShareBlob peer_blob = {};
    PeerComm *d_peer = nullptr;
    void *peer_maps[DLT_MAX_PEERS] = {};  // mappings opened by dlt_peer_attach (to be closed)
    size_t peer_map_bytes[DLT_MAX_PEERS] = {};
    // (DLT_ZEROCOPY=0 opts out): dlt_measure's result block is stored by k_residual straight into pinned host memory and the
    // host spins on a flag there instead of issuing a device->host copy and synchronising the stream
    int reuse_seen_last = 0, reuse_redo_last = 0, reuse_holdoff = 0;  // reuse_feedback
    double match_pose[24] = {0};  // pose of the last match pass (dlt_measure)
    bool knn_reuse = true;  // rematch passes prove neighbour sets unchanged where they can (DLT_KNN_REUSE=0: always search; A/B switch)
    bool zerocopy = true;  // measured on B200 (round 2): -22 us per C2 scan against the copy + stream synchronisation; DLT_ZEROCOPY=0 switches it off
    unsigned long long zc_seq = 0;
    bool peer_on = false;
    bool peer_detached = false;
    int *h_peer_status = nullptr;         // pinned
    // the iteration loop as a CUDA graph with conditional nodes (built lazily, once: every kernel in it has a fixed grid and
    // takes its sizes from device memory)
#if !defined(DLT_EMU)
    cudaGraph_t loop_graph = nullptr;
    cudaGraphExec_t loop_exec = nullptr;
#endif
    int loop_graph_state = 0;  // 0 not tried, 1 ready, -1 unavailable (stream launches are used instead)
    int use_graph = 0;  // measured slower than stream launches with early-exit kernels: opt-in via DLT_LOOP_GRAPH=1
    int far_hint = 1;                  // unresolved queries of the previous scan: queue the exact-neighbour fallback behind the loop?
    // fused loop kernels (dlt_loop_kernels.cuh): one launch per iteration, or -- cooperative launch -- one per scan
    int loop_fused = 1;                // DLT_LOOP_FUSED=0: the classic three kernels per iteration
    int loop_coop = 0;                 // DLT_LOOP_COOP=1: the persistent cooperative k_iekf_loop instead of k_iekf_iter launches (measured slower on B200)
    int loop_blocks = 0;               // grid of the loop kernels (0 = not sized yet); DLT_LOOP_BLOCKS caps it (several sequences per GPU)
    GridBar *d_bar = nullptr;
    // front end (dlt_frontend_sample): sensor cloud staging, grown on demand
    unsigned char *d_sensor = nullptr;
    size_t sensor_cap = 0;
    float4 *d_fe_tmp = nullptr;
    unsigned char *d_fe_keep = nullptr;
    unsigned *d_fe_pos = nullptr, *d_fe_blk = nullptr, *d_fe_blkoff = nullptr;
    int *d_fe_index = nullptr;
    size_t fe_cap = 0;
    // pinned host staging
    double *h_result = nullptr;
    int *h_ints = nullptr;
    ScanScalars *h_sc = nullptr;
    std::vector<void *> allocs;
    // optional per-kernel timing with CUDA events on the launching stream
    bool prof_on = false;
    std::vector<ProfSpan> spans;
    double prof_ms[kProfKinds] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long prof_n[kProfKinds] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
};

namespace {
struct ProfScope {  // records an event pair around the launches issued inside the scope
    dlt_handle h;
    ProfSpan sp;
    bool on;
    cudaStream_t st;
    ProfScope(dlt_handle h_, int kind, cudaStream_t s = nullptr) : h(h_), on(h_->prof_on), st(s ? s : h_->stream) {
        if (!on) return;
        sp.kind = kind;
        if (rt::event_create(&sp.a) || rt::event_create(&sp.b)) {
            on = false;
            return;
        }
        rt::event_record(sp.a, st);
    }
    ~ProfScope() {
        if (!on) return;
        rt::event_record(sp.b, st);
        h->spans.push_back(sp);
    }
};
}  // namespace

#define DLT_FAIL(h, code, msg)  \
    do {                        \
        (h)->err = (msg);       \
        return (code);          \
    } while (0)
#define DLT_RT(h, expr)                                                                    \
    do {                                                                                   \
        if ((expr) != 0) {                                                                 \
            (h)->err = std::string(#expr) + ": " + rt::last_error();                       \
            return DLT_E_CUDA;                                                             \
        }                                                                                  \
    } while (0)

template <typename T>
static int dalloc(dlt_handle h, T **p, size_t count) {
    void *v = nullptr;
    if (rt::alloc(&v, count * sizeof(T)) != 0) return 1;
    h->allocs.push_back(v);
    *p = static_cast<T *>(v);
    return 0;
}
static inline int div_up(long long a, long long b) { return (int)((a + b - 1) / b); }

static Pose pose_from(const double *p) {
    Pose P;
    for (int i = 0; i < 9; i++) P.rot_end[i] = p[i];
    for (int i = 0; i < 3; i++) P.pos_end[i] = p[9 + i];
    for (int i = 0; i < 9; i++) P.R_L_I[i] = p[12 + i];
    for (int i = 0; i < 3; i++) P.T_L_I[i] = p[21 + i];
    return P;
}

static int map_reset(dlt_handle h) {
    h->counters_fresh = false;
    h->map_dead = false;
    DLT_RT(h, rt::fill(h->map.table, 0xFF, h->table_cap * sizeof(Slot), h->stream));
    DLT_RT(h, rt::fill(h->d_counters, 0, 16 * sizeof(int), h->stream));
    DLT_RT(h, rt::fill(h->map.pool_box_min, 0x7F, 3 * h->n_pool_runs * sizeof(int), h->stream));  // min > max: unused run
    DLT_RT(h, rt::fill(h->map.pool_box_max, 0x80, 3 * h->n_pool_runs * sizeof(int), h->stream));
    h->reuse_seen_last = h->reuse_redo_last = 0;
    return DLT_OK;
}

// one match pass: 8 lanes per query over the 3^3 block, then warp-per-query for what that could not prove exact.
// n_grid sizes the k_knn8 grid (an estimate of n is fine: the kernel strides); with la.ctl the kernels take n, the pose
// and the do_match decision from device memory.  far_count is zeroed by k_knn8, the hand-over counter by k_residual
// (the classic callers without a residual pass clear it themselves).
static int launch_knn(dlt_handle h, const float4 *d_q, int n, int n_grid, int body_frame, const Pose &P, LoopArgs la, bool reuse = false) {
    {
        ProfScope prof8(h, 7);  // the dominant kernel on its own (group 0 spans the whole match pass)
        const int wave8 = h->n_sm * DLT_KNN8_MINBLOCKS;  // what is resident at once: the kernel strides, so a partial second wave never forms
        if (reuse) {  // one query per thread: most neighbour sets of a rematch pass are proven unchanged, the rest goes to k_knn
            int gr = div_up(n_grid, kReuseBlock);
            if (gr > 8 * h->n_sm) gr = 8 * h->n_sm;
            if (gr < 1) gr = 1;
            DLT_LAUNCH(k_knn_reuse, gr, kReuseBlock, h->stream, h->map, d_q, n, P, h->cfg.max_sq_dist, h->knn, h->d_unres, h->d_counters + 8, la,
                       h->d_counters + 12);
        } else {
            int g8 = div_up(n_grid, kKnn8Block / 8);
            if (g8 > wave8) g8 = wave8;
            if (g8 < 1) g8 = 1;
            DLT_LAUNCH(k_knn8, g8, kKnn8Block, h->stream, h->map, d_q, n, body_frame, P, h->cfg.max_sq_dist, h->knn, h->d_unres, h->d_counters + 8, la);
        }
    }
    int grid = div_up(n_grid, kKnnWarps);
    const int cap_grid = h->n_sm * 8;
    if (grid > cap_grid) grid = cap_grid;
    if (grid < 1) grid = 1;
    DLT_LAUNCH(k_knn, grid, kKnnWarps * 32, h->stream, h->map, d_q, n, body_frame, P, h->cfg.max_sq_dist, h->knn, (const int *)h->d_unres,
               (const int *)(h->d_counters + 8), la, reuse ? 1 : 0);
    return DLT_OK;
}

// the side stream's read of d_result must finish before the main stream overwrites it
static int eig_join(dlt_handle h) {
    if (h->eig_pending) {
        DLT_RT(h, rt::stream_wait(h->stream, h->ev_join));
        h->eig_pending = false;
    }
    return DLT_OK;
}

// feats_down_size after dlt_scan_downsample_async: fetch it the first time the host needs it
static int resolve_n_down(dlt_handle h) {
    if (!h->n_down_on_device) return DLT_OK;
    DLT_RT(h, rt::d2h(h->h_sc, h->d_sc, sizeof(ScanScalars), h->stream));
    DLT_RT(h, rt::sync(h->stream));
    h->n_down_on_device = false;
    if (h->h_sc->vox_status == 2) {  // the bitmap was too small for this scan's bounding box: grow it and run the VoxelGrid again
        int nd = 0;
        return dlt_scan_downsample(h, &nd);
    }
    h->n_down = h->h_sc->n_down;
    h->n_down_hint = h->n_down;
    return DLT_OK;
}

// ... or take it from a result block that just came back (k_residual leaves it in R[159])
static int adopt_n_down(dlt_handle h, const double *R) {
    if (!h->n_down_on_device) return DLT_OK;
    h->n_down_on_device = false;
    if (R[159] < 0.0) return -1;  // the occupancy bitmap was too small for this scan: the caller grows it and runs again
    h->n_down = (int)(R[159] + 0.5);
    h->n_down_hint = h->n_down;
    return DLT_OK;
}

// The device counter keeps growing after the bucket pool is exhausted (map_claim / map_append add first, then flag the
// error), so every host-side use of it as a loop bound is clamped to the pool size.
static inline int clamped_buckets(dlt_handle h) {
    const int nb = h->h_ints[0];
    return nb < 0 ? 0 : (nb > h->map.bucket_cap ? h->map.bucket_cap : nb);
}
// Once the pool or the table overflowed the map no longer mirrors the reference tree: every later mutation / update is
// refused until dlt_map_build starts a new map.
static int map_refuse_if_dead(dlt_handle h) {
    if (h->map_dead) DLT_FAIL(h, DLT_E_CAPACITY, "the map overflowed earlier (max_map_points): rebuild it with dlt_map_build");
    return DLT_OK;
}

// The reuse statistics ride back with the map counters once per scan: when more than a quarter of a rematch pass had to be
// searched again (large corrections between the passes) the warp-per-query kernel is the wrong tool for that many queries, so
// the next scans search every pass in full and the reuse is tried again a little later.
static void reuse_feedback(dlt_handle h, const int *c16) {
    const int seen = c16[12] - h->reuse_seen_last, redo = c16[13] - h->reuse_redo_last;
    h->reuse_seen_last = c16[12];
    h->reuse_redo_last = c16[13];
    if (h->reuse_holdoff > 0) h->reuse_holdoff--;
    if (seen > 0 && 4 * (long long)redo > (long long)seen) h->reuse_holdoff = 8;
}

static int map_check_error(dlt_handle h) {
    DLT_RT(h, rt::d2h(h->h_ints, h->d_counters, 16 * sizeof(int), h->stream));
    DLT_RT(h, rt::sync(h->stream));
    h->counters_fresh = true;
    reuse_feedback(h, h->h_ints);
    if (h->h_ints[2] != 0) h->map_dead = true;
    if (h->h_ints[2] == 1) DLT_FAIL(h, DLT_E_CAPACITY, "map bucket pool exhausted (raise max_map_points)");
    if (h->h_ints[2] == 2) DLT_FAIL(h, DLT_E_CAPACITY, "map hash table full (raise max_map_points)");
    return DLT_OK;
}

// ---- asynchronous map_incremental: ordering of the main stream behind the insert stream, adoption of its counters
static inline int ins_wait_classified(dlt_handle h) {  // before anything overwrites the downsampled scan
    if (h->cls_pending) {
        DLT_RT(h, rt::stream_wait(h->stream, h->ev_classified));
        h->cls_pending = false;
    }
    return DLT_OK;
}
static int ins_finish(dlt_handle h) {  // before anything reads or changes the map, the neighbour buffers or the counters
    if (h->ins_pending) {
        DLT_RT(h, rt::stream_wait(h->stream, h->ev_inserted));
        h->ins_pending = false;
        h->cls_pending = false;
    }
    if (h->ins_uncollected) {
        DLT_RT(h, rt::event_sync(h->ev_inserted));
        h->ins_uncollected = false;
        std::memcpy(h->h_ints, h->h_ins_ints, 16 * sizeof(int));
        h->counters_fresh = true;
        reuse_feedback(h, h->h_ints);
        h->last_ins_ds = h->h_ints[6];
        h->last_ins_raw = h->h_ints[7];
        if (h->h_ints[2] != 0) h->map_dead = true;
        if (h->h_ints[2] == 1) DLT_FAIL(h, DLT_E_CAPACITY, "map bucket pool exhausted (raise max_map_points)");
        if (h->h_ints[2] == 2) DLT_FAIL(h, DLT_E_CAPACITY, "map hash table full (raise max_map_points)");
        if (h->peer_on && *h->h_peer_status != 0) DLT_FAIL(h, DLT_E_STATE, "peer exchange timed out (a rank did not take part in map_incremental)");
    }
    return DLT_OK;
}

// the voxel scratch of one insert batch: only has to be clean where the batch can hash to, so it is sized to the batch
// (power of two >= 4 n, or >= 2 n when n is an upper bound of the batch) and cleared with ONE memset (vwin sits right behind vkeys)
static int prepare_scratch(dlt_handle h, int n, bool n_is_upper_bound, DsScratch *out) {
    size_t sc = 1024;
    while (sc < (n_is_upper_bound ? 2 : 4) * (size_t)n) sc <<= 1;
    if (sc > h->scratch_cap) sc = h->scratch_cap;
    DsScratch scr = h->scratch;
    scr.mask = (unsigned)(sc - 1);
    scr.vwin = scr.vkeys + sc;
    DLT_RT(h, rt::fill(scr.vkeys, 0xFF, 2 * sc * sizeof(unsigned long long), h->stream));
    *out = scr;
    return DLT_OK;
}

// claim | (bid, resolve) | append over n device points with per-point flags.  With a gate (device-resident loop) n is
// only the grid size: the kernels read the real count, and whether to run at all, from device memory.  front_done: the
// caller's classification kernel already claimed the cells and placed the bids (scratch `pre`).
static int insert_points(dlt_handle h, const float4 *d_pts, int n, bool any_ds, InsertGate gate = {nullptr, nullptr}, const DsScratch *pre = nullptr) {
    if (n <= 0) return DLT_OK;
    h->counters_fresh = false;
    ProfScope prof(h, 4);
    const int B = 256, G = div_up(n, B);
    if (!pre)
        DLT_LAUNCH(k_map_claim, G, B, h->stream, h->map, d_pts, n, (const unsigned char *)h->d_dsflag, (const unsigned char *)h->d_addflag,
                   h->d_cellslot, h->map.shard_count > 1 ? 1 : 0, gate);
    if (any_ds) {
        DsScratch scr;
        if (pre) {
            scr = *pre;
        } else {
            int rs = prepare_scratch(h, n, gate.go != nullptr, &scr);
            if (rs) return rs;
            DLT_LAUNCH(k_ds_bid, G, B, h->stream, h->map, scr, d_pts, n, (const unsigned char *)h->d_dsflag, h->d_vslot, gate);
        }
        DLT_LAUNCH(k_ds_resolve, G, B, h->stream, h->map, scr, d_pts, n, (const unsigned char *)h->d_dsflag, (const int *)h->d_vslot,
                   (const int *)h->d_cellslot, h->d_addflag, gate);
    }
    DLT_LAUNCH(k_map_append, G, B, h->stream, h->map, d_pts, n, (const unsigned char *)h->d_addflag, (const int *)h->d_cellslot, gate);
    DLT_RT(h, rt::check_launch());
    return DLT_OK;
}

static int add_host_points(dlt_handle h, const float *xyzi, int n, int downsample_on) {
    for (int off = 0; off < n; off += h->cap) {
        int c = n - off < h->cap ? n - off : h->cap;
        DLT_RT(h, rt::h2d(h->d_pw, xyzi + (size_t)off * 4, (size_t)c * sizeof(float4), h->stream));
        DLT_RT(h, rt::fill(h->d_dsflag, downsample_on ? 1 : 0, (size_t)c, h->stream));
        DLT_RT(h, rt::fill(h->d_addflag, downsample_on ? 0 : 1, (size_t)c, h->stream));
        int rc = insert_points(h, h->d_pw, c, downsample_on != 0);
        if (rc) return rc;
        // the staging buffers are reused by the next chunk
        DLT_RT(h, rt::sync(h->stream));
    }
    return map_check_error(h);
}

// grid of the fused loop kernels: what is co-resident on the device (a cooperative launch needs exactly that bound), or
// DLT_LOOP_BLOCKS when several sequences are meant to share the GPU
static int size_loop_grid(dlt_handle h) {
    if (h->loop_blocks > 0) return DLT_OK;
    int per_sm = 2;
#if !defined(DLT_EMU)
    int a = 0, b = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, k_iekf_loop<false>, kLoopBlock, 0) != cudaSuccess ||
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_iekf_loop<true>, kLoopBlock, 0) != cudaSuccess || a < 1 || b < 1) {
        cudaGetLastError();
        h->loop_coop = 0;
    } else {
        per_sm = h->cfg.extrinsic_est_en ? b : a;
    }
    int coop = 0, dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev) != cudaSuccess || !coop) h->loop_coop = 0;
#endif
    int g = per_sm * h->n_sm;
    if (const char *e = std::getenv("DLT_LOOP_BLOCKS")) {
        const int cap = std::atoi(e);
        if (cap > 0 && cap < g) g = cap;
    }
    if (g < 1) g = 1;
    h->loop_blocks = g;
    return DLT_OK;
}

// exact neighbours for the queries the ring search could not resolve
static int run_far(dlt_handle h, int *n_far_out) {
    if (h->nfar_known && h->h_last_nfar == 0) {
        if (n_far_out) *n_far_out = 0;
        return DLT_OK;
    }
    DLT_RT(h, rt::d2h(h->h_ints, h->d_counters, 8 * sizeof(int), h->stream));
    DLT_RT(h, rt::sync(h->stream));
    int nfar = h->h_ints[5], n_buckets = clamped_buckets(h);
    if (n_far_out) *n_far_out = nfar;
    if (nfar <= 0) return DLT_OK;
    ProfScope prof(h, 5);
    for (int off = 0; off < nfar; off += kFarChunk) {
        int c = nfar - off < kFarChunk ? nfar - off : kFarChunk;
        DLT_LAUNCH(k_far_scan, dim3(kFarGroupsX, kFarSlices), kFarWarps * 32, h->stream, h->map, n_buckets, (const float4 *)h->knn.qw,
                   (const int *)h->knn.far_list, off, c, kFarSlices, h->d_far_partial);
        DLT_LAUNCH(k_far_merge, div_up(c, 128), 128, h->stream, (const int *)h->knn.far_list, off, c, kFarSlices,
                   (const Cand *)h->d_far_partial, h->cfg.max_sq_dist, h->knn);
    }
    DLT_RT(h, rt::fill(h->d_counters + 5, 0, sizeof(int), h->stream));
    DLT_RT(h, rt::check_launch());
    h->nfar_known = true;
    h->h_last_nfar = 0;
    return DLT_OK;
}

extern "C" {

void dlt_default_config(dlt_config *c) {
    c->ds_scan = 0.5f;
    c->ds_map = 0.5f;
    c->max_sq_dist = 5.0f;
    c->plane_thr = 0.1f;
    c->extrinsic_est_en = 0;
    c->device = 0;
    c->max_scan_points = 262144;
    c->max_map_points = 4 * 1024 * 1024;
    c->voxel_bitmap_bits = 0;
    c->shard_rank = 0;
    c->shard_count = 1;
    c->shard_tile_shift = 0;
}

int dlt_destroy(dlt_handle h) {
    if (!h) return DLT_E_INVALID;
    rt::set_device(h->cfg.device);
    if (h->own_stream) rt::sync(h->own_stream);
    if (h->ins_stream) rt::sync(h->ins_stream);
    h->cls_pending = h->ins_pending = h->ins_uncollected = false;
    for (void *p : h->allocs) rt::release(p);
    for (void *p : {(void *)h->d_sensor, (void *)h->d_fe_tmp, (void *)h->d_fe_keep, (void *)h->d_fe_pos, (void *)h->d_fe_blk, (void *)h->d_fe_blkoff,
                    (void *)h->d_fe_index})
        rt::release(p);
    rt::pinned_release(h->h_result);
    rt::pinned_release(h->h_ints);
    rt::pinned_release(h->h_sc);
    rt::pinned_release(h->h_iekf);
    rt::pinned_release(h->h_ins_ints);
    if (h->have_ins) {
        rt::event_destroy(h->ev_loop_done);
        rt::event_destroy(h->ev_classified);
        rt::event_destroy(h->ev_inserted);
    }
    rt::stream_destroy(h->ins_stream);
    dlt_peer_detach(h);
    if (h->peer_box) rt::shared_release(h->peer_box, h->peer_bytes, &h->peer_blob);
    rt::release(h->d_peer);
    rt::pinned_release(h->h_peer_status);
#if !defined(DLT_EMU)
    if (h->loop_exec) cudaGraphExecDestroy(h->loop_exec);
    if (h->loop_graph) cudaGraphDestroy(h->loop_graph);
#endif
    if (h->have_copy) {
        rt::sync(h->copy_stream);
        rt::event_destroy(h->ev_copy);
    }
    rt::stream_destroy(h->copy_stream);
    if (h->have_aux) {
        rt::sync(h->aux_stream);
        rt::event_destroy(h->ev_fork);
        rt::event_destroy(h->ev_join);
    }
    rt::stream_destroy(h->aux_stream);
    rt::stream_destroy(h->own_stream);
    delete h;
    return DLT_OK;
}

int dlt_create(const dlt_config *cfg, dlt_handle *out) {
    if (!cfg || !out) return DLT_E_INVALID;
    *out = nullptr;
    if (cfg->ds_scan <= 0.f || cfg->ds_map <= 0.f || cfg->max_scan_points <= 0 || cfg->max_map_points <= 0 || cfg->shard_count < 1 ||
        cfg->shard_rank < 0 || cfg->shard_rank >= cfg->shard_count)
        return DLT_E_INVALID;
    // shard_keeps_cell tests the 8 corners of the +-kShardHalo cube: sufficient only while that 9-cell span touches at most two
    // tiles per axis, i.e. a tile edge of at least 8 cells
    if (cfg->shard_count > 1 && cfg->shard_tile_shift != 0 && cfg->shard_tile_shift < 3) return DLT_E_INVALID;
    if (rt::device_count() <= cfg->device) return DLT_E_NO_DEVICE;  // no CPU path: fail loudly
    if (rt::set_device(cfg->device) != 0) return DLT_E_NO_DEVICE;
    dlt_handle h = new dlt_handle_s();
    h->cfg = *cfg;
    h->cap = cfg->max_scan_points;
    h->n_sm = rt::sm_count();
    bool ok = rt::stream_create(&h->own_stream) == 0;
    h->stream = h->own_stream;
    h->have_copy = rt::stream_create(&h->copy_stream) == 0 && rt::event_create_untimed(&h->ev_copy) == 0;
    if (const char *e = std::getenv("DLT_LOOP_GRAPH")) h->use_graph = (e[0] == '1') ? 1 : 0;  // A/B switch for measurements
    if (const char *e = std::getenv("DLT_ZEROCOPY")) h->zerocopy = (e[0] == '1');            // A/B switch for measurements
    if (const char *e = std::getenv("DLT_KNN_REUSE")) h->knn_reuse = (e[0] == '1');
#if !defined(DLT_EMU)
    if (const char *e = std::getenv("DLT_PDL")) rt::g_pdl = (e[0] == '1') ? 1 : 0;  // process-wide A/B switch for measurements
#endif
    if (const char *e = std::getenv("DLT_LOOP_FUSED")) h->loop_fused = (e[0] == '1');
    if (const char *e = std::getenv("DLT_LOOP_COOP")) h->loop_coop = (e[0] == '1');
    h->have_aux = rt::stream_create(&h->aux_stream) == 0 && rt::event_create_untimed(&h->ev_fork) == 0 && rt::event_create_untimed(&h->ev_join) == 0;
    h->have_ins = rt::stream_create(&h->ins_stream) == 0 && rt::event_create_untimed(&h->ev_loop_done) == 0 &&
                  rt::event_create_untimed(&h->ev_classified) == 0 && rt::event_create_untimed(&h->ev_inserted) == 0;

    // search cell edge = ds_map * 2^shift with 3 edges covering sqrt(max_sq_dist)
    int shift = 1;
    while (3.0 * cfg->ds_map * (double)(1 << shift) < 1.05 * std::sqrt((double)cfg->max_sq_dist) && shift < 12) shift++;
    size_t bucket_cap = (size_t)cfg->max_map_points;
    if (bucket_cap < 4096) bucket_cap = 4096;
    size_t tcap = 1;
    while (tcap < 2 * bucket_cap) tcap <<= 1;
    h->table_cap = tcap;
    h->map.table_mask = (unsigned)(tcap - 1);
    h->map.bucket_cap = (int)bucket_cap;
    h->map.ds = cfg->ds_map;
    h->map.cell_shift = shift;
    h->map.shard_rank = cfg->shard_rank;
    h->map.shard_count = cfg->shard_count;
    h->map.tile_shift = cfg->shard_tile_shift > 0 ? cfg->shard_tile_shift : 5;
    h->bitmap_bits = cfg->voxel_bitmap_bits > 0 ? cfg->voxel_bitmap_bits : (1ll << 27);
    const size_t words = (size_t)((h->bitmap_bits + 31) / 32);
    h->n_scan_blocks = div_up((long long)words, kScanWordsPerBlock);
    const size_t cap = (size_t)h->cap;
    int blocks_max = div_up((long long)cap, kResidBlock);
    if (blocks_max < 16 * h->n_sm) blocks_max = 16 * h->n_sm;  // the fused loop kernels write one partial per block of their grid
    size_t sc_cap = 1;
    while (sc_cap < 4 * cap) sc_cap <<= 1;
    h->scratch.mask = (unsigned)(sc_cap - 1);
    h->scratch_cap = sc_cap;

    h->n_pool_runs = bucket_cap / kPoolRun + 1;
    ok = ok && !dalloc(h, &h->map.pool_box_min, 3 * h->n_pool_runs) && !dalloc(h, &h->map.pool_box_max, 3 * h->n_pool_runs);
    ok = ok && !dalloc(h, &h->map.table, tcap) && !dalloc(h, &h->map.buckets, bucket_cap) && !dalloc(h, &h->d_counters, 16) && !dalloc(h, &h->d_unres, cap) &&
         !dalloc(h, &h->d_raw, cap * kRawStride4) && !dalloc(h, &h->d_undist, cap) && !dalloc(h, &h->d_down, cap) &&
         !dalloc(h, &h->d_poses, kMaxImuPoses) && !dalloc(h, &h->d_sc, 1) && !dalloc(h, &h->d_bitmap, words) &&
         !dalloc(h, &h->d_wprefix, words) && !dalloc(h, &h->d_blksum, (size_t)h->n_scan_blocks) &&
         !dalloc(h, &h->d_blkoff, (size_t)h->n_scan_blocks) && !dalloc(h, &h->d_vidx, cap) && !dalloc(h, &h->acc.sx, cap) &&
         !dalloc(h, &h->acc.sy, cap) && !dalloc(h, &h->acc.sz, cap) && !dalloc(h, &h->acc.si, cap) && !dalloc(h, &h->acc.cnt, cap) &&
         !dalloc(h, &h->acc.idx, cap) && !dalloc(h, &h->d_vop, cap) && !dalloc(h, &h->knn.qw, cap) && !dalloc(h, &h->knn.nbr, cap * kK) &&
         !dalloc(h, &h->knn.nbr_id, cap * kK) && !dalloc(h, &h->knn.nbr_cnt, cap) && !dalloc(h, &h->knn.flags, cap) && !dalloc(h, &h->knn.d6lb, cap) &&
         !dalloc(h, &h->knn.far_list, cap) && !dalloc(h, &h->knn.nn_list, cap) && !dalloc(h, &h->knn.nn_pos, cap) && !dalloc(h, &h->knn.nn_key, cap) && !dalloc(h, &h->d_plane, cap) && !dalloc(h, &h->d_coeff, cap) && !dalloc(h, &h->d_sel, cap) &&
         !dalloc(h, &h->d_eff, cap) && !dalloc(h, &h->d_partials, (size_t)blocks_max * NormalEq<true>::NR) &&
         !dalloc(h, &h->d_result, (size_t)kResultDoubles) && !dalloc(h, &h->d_ticket, 4) &&
         !dalloc(h, &h->d_far_partial, (size_t)kFarChunk * kFarSlices * kK) && !dalloc(h, &h->d_pw, cap) && !dalloc(h, &h->d_dsflag, cap) &&
         !dalloc(h, &h->d_addflag, cap) && !dalloc(h, &h->d_cellslot, cap) && !dalloc(h, &h->d_vslot, cap) &&
         !dalloc(h, &h->scratch.vkeys, 2 * sc_cap) && !dalloc(h, &h->d_iekf, 1) && !dalloc(h, &h->d_bar, 1);
    if (cfg->shard_count > 1) ok = ok && !dalloc(h, &h->d_flagbuf, cap);
    void *p = nullptr;
    ok = ok && rt::pinned_alloc(&p, sizeof(dlt_iekf_block)) == 0;
    h->h_iekf = (dlt_iekf_block *)p;
    ok = ok && rt::pinned_alloc(&p, kResultDoubles * sizeof(double)) == 0;
    h->h_result = (double *)p;
    ok = ok && rt::pinned_alloc(&p, 64 * sizeof(int)) == 0;
    h->h_ints = (int *)p;
    ok = ok && rt::pinned_alloc(&p, sizeof(ScanScalars)) == 0;
    h->h_sc = (ScanScalars *)p;
    ok = ok && rt::pinned_alloc(&p, 16 * sizeof(int)) == 0;
    h->h_ins_ints = (int *)p;
    if (!ok) {
        dlt_destroy(h);
        return DLT_E_CUDA;
    }
    // pinned staging starts from zero: the zero-copy flags in h_ints are compared with sequence numbers that start at 1, and a
    // recycled allocation may still hold an earlier handle's flag (seen as a result block that was never written)
    std::memset(h->h_iekf, 0, sizeof(dlt_iekf_block));
    std::memset(h->h_result, 0, kResultDoubles * sizeof(double));
    std::memset(h->h_ints, 0, 64 * sizeof(int));
    std::memset(h->h_sc, 0, sizeof(ScanScalars));
    std::memset(h->h_ins_ints, 0, 16 * sizeof(int));
    h->map.n_buckets = h->d_counters + 0;
    h->map.n_live = h->d_counters + 1;
    h->map.error = h->d_counters + 2;
    h->knn.far_count = h->d_counters + 5;
    h->knn.nn_count = h->d_counters + 9;  // [9] queries needing their nearest point, [10] ring-search overflows
    // accumulators and bitmap are kept zero between scans by k_vox_final
    bool z = rt::fill(h->d_bitmap, 0, words * 4, h->stream) == 0 && rt::fill(h->acc.sx, 0, cap * 8, h->stream) == 0 &&
             rt::fill(h->acc.sy, 0, cap * 8, h->stream) == 0 && rt::fill(h->acc.sz, 0, cap * 8, h->stream) == 0 &&
             rt::fill(h->acc.si, 0, cap * 8, h->stream) == 0 && rt::fill(h->acc.cnt, 0, cap * 4, h->stream) == 0 &&
             rt::fill(h->d_ticket, 0, 16, h->stream) == 0 && rt::fill(h->d_sel, 0, cap, h->stream) == 0 &&
             rt::fill(h->d_result, 0, kResultDoubles * sizeof(double), h->stream) == 0 && rt::fill(h->d_bar, 0, sizeof(GridBar), h->stream) == 0;
    DLT_LAUNCH(k_scan_reset, 1, 32, h->stream, h->d_sc);
    if (!z || map_reset(h) != DLT_OK || rt::sync(h->stream) != 0) {
        dlt_destroy(h);
        return DLT_E_CUDA;
    }
    *out = h;
    return DLT_OK;
}

const char *dlt_last_error(dlt_handle h) { return h ? h->err.c_str() : "null handle"; }

int dlt_set_stream(dlt_handle h, void *s) {
    if (!h) return DLT_E_INVALID;
    if (int ri = ins_finish(h)) return ri;
    rt::sync(h->stream);
    h->stream = s ? (cudaStream_t)s : h->own_stream;
    return DLT_OK;
}
void *dlt_stream(dlt_handle h) { return h ? (void *)h->stream : nullptr; }
int dlt_sync(dlt_handle h) {
    if (!h) return DLT_E_INVALID;
    if (int ri = ins_finish(h)) return ri;
    DLT_RT(h, rt::sync(h->stream));
    return DLT_OK;
}

// ------------------------------------------------------------------ map
int dlt_map_build(dlt_handle h, const float *xyzi, int n) {
    if (!h || (n > 0 && !xyzi) || n < 0) return DLT_E_INVALID;
    rt::set_device(h->cfg.device);
    if (int ri = ins_finish(h)) return ri;
    int rc = map_reset(h);
    if (rc) return rc;
    h->have_match = false;
    return add_host_points(h, xyzi, n, 0);
}

int dlt_map_build_from_scan(dlt_handle h, const double *pose24) {
    if (!h || !pose24) return DLT_E_INVALID;
    if (!h->have_down) DLT_FAIL(h, DLT_E_STATE, "dlt_map_build_from_scan before a downsampled scan is set");
    rt::set_device(h->cfg.device);
    if (int ri = ins_finish(h)) return ri;
    if (int rn = resolve_n_down(h)) return rn;
    int rc = map_reset(h);
    if (rc) return rc;
    h->have_match = false;
    const int n = h->n_down;
    if (n == 0) return DLT_OK;
    Pose P = pose_from(pose24);
    DLT_LAUNCH(k_scan_to_world, div_up(n, 256), 256, h->stream, (const float4 *)h->d_down, n, P, h->d_pw, h->d_dsflag, h->d_addflag);
    rc = insert_points(h, h->d_pw, n, false);
    if (rc) return rc;
    return map_check_error(h);
}

int dlt_map_add(dlt_handle h, const float *xyzi, int n, int downsample_on) {
    if (!h || (n > 0 && !xyzi) || n < 0) return DLT_E_INVALID;
    rt::set_device(h->cfg.device);
    if (int ri = ins_finish(h)) return ri;
    if (int rd = map_refuse_if_dead(h)) return rd;
    h->have_match = false;
    return add_host_points(h, xyzi, n, downsample_on);
}

int dlt_map_delete_boxes(dlt_handle h, const float *boxes6, int nb, int *deleted) {
    if (!h || nb < 0 || (nb > 0 && !boxes6)) return DLT_E_INVALID;
    if (int rd = map_refuse_if_dead(h)) return rd;
    rt::set_device(h->cfg.device);
    if (int ri = ins_finish(h)) return ri;
    h->counters_fresh = false;
    DLT_RT(h, rt::fill(h->d_counters + 3, 0, sizeof(int), h->stream));
    DLT_RT(h, rt::d2h(h->h_ints, h->d_counters, 8 * sizeof(int), h->stream));
    DLT_RT(h, rt::sync(h->stream));
    int n_buckets = clamped_buckets(h);
    for (int off = 0; off < nb && n_buckets > 0; off += 8) {
        BoxSet bs;
        bs.n = nb - off < 8 ? nb - off : 8;
        for (int k = 0; k < bs.n; k++)
            for (int a = 0; a < 3; a++) {
                bs.mn[k][a] = boxes6[(size_t)(off + k) * 6 + a];
                bs.mx[k][a] = boxes6[(size_t)(off + k) * 6 + 3 + a];
            }
        DLT_LAUNCH(k_map_delete_boxes, div_up((long long)n_buckets * 8, 256), 256, h->stream, h->map, bs, n_buckets, h->d_counters + 3);
    }
    DLT_RT(h, rt::check_launch());
    DLT_RT(h, rt::d2h(h->h_ints, h->d_counters, 8 * sizeof(int), h->stream));
    DLT_RT(h, rt::sync(h->stream));
    if (deleted) *deleted = h->h_ints[3];
    h->have_match = false;
    h->counters_fresh = true;
    return DLT_OK;
}

int dlt_map_valid_count(dlt_handle h, int *n) {
    if (!h || !n) return DLT_E_INVALID;
    rt::set_device(h->cfg.device);
    if (int ri = ins_finish(h)) return ri;
    if (!h->counters_fresh) {
        DLT_RT(h, rt::d2h(h->h_ints, h->d_counters, 8 * sizeof(int), h->stream));
        DLT_RT(h, rt::sync(h->stream));
        h->counters_fresh = true;
    }
    *n = h->h_ints[1];
    return DLT_OK;
}

int dlt_map_export(dlt_handle h, float *xyzi, int cap, int *n) {
    if (!h || !n || cap < 0 || (cap > 0 && !xyzi)) return DLT_E_INVALID;
    rt::set_device(h->cfg.device);
    if (int ri = ins_finish(h)) return ri;
    DLT_RT(h, rt::fill(h->d_counters + 4, 0, sizeof(int), h->stream));
    DLT_RT(h, rt::d2h(h->h_ints, h->d_counters, 8 * sizeof(int), h->stream));
    DLT_RT(h, rt::sync(h->stream));
    int n_buckets = clamped_buckets(h), live = h->h_ints[1];
    *n = live;
    if (cap == 0 || live == 0 || n_buckets == 0) return DLT_OK;
    float4 *d_out = nullptr;
    void *v = nullptr;
    if (rt::alloc(&v, (size_t)cap * sizeof(float4)) != 0) DLT_FAIL(h, DLT_E_CUDA, "export buffer allocation failed");
    d_out = (float4 *)v;
    DLT_LAUNCH(k_map_export, div_up((long long)n_buckets * 8, 256), 256, h->stream, h->map, n_buckets, d_out, cap, h->d_counters + 4);
    int rc = rt::check_launch();
    int m = live < cap ? live : cap;
    if (!rc) rc = rt::d2h(xyzi, d_out, (size_t)m * sizeof(float4), h->stream);
    if (!rc) rc = rt::sync(h->stream);
    rt::release(v);
    if (rc) DLT_FAIL(h, DLT_E_CUDA, std::string("map export: ") + rt::last_error());
    return DLT_OK;
}

int dlt_map_knn(dlt_handle h, const float *q, int nq, float *out_xyzi, float *out_d2, int *out_cnt) {
    if (!h || nq < 0 || (nq > 0 && (!q || !out_xyzi || !out_d2 || !out_cnt))) return DLT_E_INVALID;
    rt::set_device(h->cfg.device);
    if (int ri = ins_finish(h)) return ri;
    h->have_match = false;  // the neighbour buffers of the current scan are reused
    Pose P = {};
    std::vector<float> stage;
    std::vector<float4> nb;
    for (int off = 0; off < nq; off += h->cap) {
        int c = nq - off < h->cap ? nq - off : h->cap;
        stage.assign((size_t)c * 4, 0.f);
        for (int i = 0; i < c; i++) {
            stage[4 * (size_t)i] = q[3 * (size_t)(off + i)];
            stage[4 * (size_t)i + 1] = q[3 * (size_t)(off + i) + 1];
            stage[4 * (size_t)i + 2] = q[3 * (size_t)(off + i) + 2];
        }
        DLT_RT(h, rt::h2d(h->d_pw, stage.data(), (size_t)c * sizeof(float4), h->stream));
        {
            DLT_RT(h, rt::fill(h->d_counters + 8, 0, sizeof(int), h->stream));  // no residual pass here to re-arm it
            LoopArgs la = {nullptr, nullptr, nullptr, 0, 0, 0ull, 0ull, nullptr};
            int rk = launch_knn(h, (const float4 *)h->d_pw, c, c, 0, P, la);
            if (rk) return rk;
        }
        DLT_RT(h, rt::fill(h->d_counters + 8, 0, sizeof(int), h->stream));
        DLT_RT(h, rt::check_launch());
        h->nfar_known = false;
        int rc = run_far(h, nullptr);
        if (rc) return rc;
        nb.resize((size_t)c * kK);
        DLT_RT(h, rt::d2h(nb.data(), h->knn.nbr, (size_t)c * kK * sizeof(float4), h->stream));
        DLT_RT(h, rt::d2h(out_cnt + off, h->knn.nbr_cnt, (size_t)c * sizeof(int), h->stream));
        DLT_RT(h, rt::sync(h->stream));
        for (size_t i = 0; i < (size_t)c * kK; i++) {
            float *o = out_xyzi + ((size_t)off * kK + i) * 4;
            bool ok = nb[i].w >= 0.f;
            o[0] = nb[i].x;
            o[1] = nb[i].y;
            o[2] = nb[i].z;
            o[3] = 0.f;
            out_d2[(size_t)off * kK + i] = ok ? nb[i].w : -1.f;
        }
    }
    return DLT_OK;
}

// ------------------------------------------------------------------ scan
static int scan_deskew_impl(dlt_handle h, const void *pts48, bool on_device, int n_raw, const double *imu_pose22, int n_pose,
                            const double *pose24) {
    if (!h || n_raw < 0 || (n_raw > 0 && !pts48) || n_pose < 0 || (n_pose > 0 && !imu_pose22) || (n_pose >= 2 && !pose24)) return DLT_E_INVALID;
    if (n_raw > h->cap) DLT_FAIL(h, DLT_E_CAPACITY, "scan larger than max_scan_points");
    if (n_pose > kMaxImuPoses) DLT_FAIL(h, DLT_E_CAPACITY, "too many IMU poses");
    rt::set_device(h->cfg.device);
    h->n_raw = n_raw;
    h->have_raw = true;
    h->have_down = false;
    h->have_match = false;
    if (!h->sc_clean) DLT_LAUNCH(k_scan_reset, 1, 32, h->stream, h->d_sc);  // normally k_vox_final already did
    h->sc_clean = false;
    if (n_raw == 0) return DLT_OK;
    const float4 *d_in = h->d_raw;
    if (on_device) {
        d_in = static_cast<const float4 *>(pts48);
    } else if (h->pf_pending_ptr == pts48) {  // registered but never issued (no update ran in between): plain upload
        h->pf_pending_ptr = nullptr;
        h->pf_pending_n = -1;
        DLT_RT(h, rt::h2d(h->d_raw, pts48, (size_t)n_raw * 48, h->stream));
    } else if (h->prefetched_ptr == pts48 && h->prefetched_n == n_raw && h->d_raw_next) {
        // these records were uploaded ahead of time (dlt_scan_prefetch): swap the buffers and order the stream behind the copy
        float4 *t = h->d_raw;
        h->d_raw = h->d_raw_next;
        h->d_raw_next = t;
        h->prefetched_ptr = nullptr;
        h->prefetched_n = -1;
        DLT_RT(h, rt::stream_wait(h->stream, h->ev_copy));
        d_in = h->d_raw;
    } else {
        DLT_RT(h, rt::h2d(h->d_raw, pts48, (size_t)n_raw * 48, h->stream));
    }
    Pose P = {};
    ProfScope prof(h, 2);
    if (n_pose >= 2) {
        static_assert(sizeof(ImuPoseDev) == 22 * sizeof(double), "Pose6D layout");
        DLT_RT(h, rt::h2d(h->d_poses, imu_pose22, (size_t)n_pose * sizeof(ImuPoseDev), h->stream));
        P = pose_from(pose24);
        DLT_LAUNCH(k_scan_first, div_up(n_raw, 256), 256, h->stream, d_in, n_raw, h->d_sc);
    }
    DLT_LAUNCH(k_scan_deskew, div_up(n_raw, kDeskewBlock), kDeskewBlock, h->stream, d_in, n_raw, (const ImuPoseDev *)h->d_poses, n_pose, P,
               n_pose >= 2 ? 1 : 0, h->d_undist, h->d_sc);
    DLT_RT(h, rt::check_launch());
    return DLT_OK;
}

// The copy engine serves host->device transfers in issue order, so a 6 MB prefetch issued ahead of the current scan's own small
// uploads (IMU poses, the IEKF block) would hold them -- and with them the whole update -- back by ~100 us (measured).  The
// prefetch is therefore only registered here and issued right behind those uploads (issue_prefetch).
static int issue_prefetch(dlt_handle h) {
    if (!h->pf_pending_ptr) return DLT_OK;
    const void *pts48 = h->pf_pending_ptr;
    const int n_raw = h->pf_pending_n;
    h->pf_pending_ptr = nullptr;
    h->pf_pending_n = -1;
    if (!h->d_raw_next && dalloc(h, &h->d_raw_next, (size_t)h->cap * kRawStride4) != 0)  // (owned through h->allocs: the two buffers swap roles)
        DLT_FAIL(h, DLT_E_CUDA, "prefetch buffer allocation failed");
    // d_raw_next is never read by kernels in flight (they read d_raw or the caller's device buffer); a previous, unconsumed
    // prefetch is simply overwritten, in copy-stream order
    DLT_RT(h, rt::h2d(h->d_raw_next, pts48, (size_t)n_raw * 48, h->copy_stream));
    DLT_RT(h, rt::event_record(h->ev_copy, h->copy_stream));
    h->prefetched_ptr = pts48;
    h->prefetched_n = n_raw;
    return DLT_OK;
}

int dlt_scan_prefetch(dlt_handle h, const void *pts48, int n_raw) {
    if (!h || n_raw < 0 || (n_raw > 0 && !pts48)) return DLT_E_INVALID;
    if (n_raw > h->cap) DLT_FAIL(h, DLT_E_CAPACITY, "scan larger than max_scan_points");
    if (!h->have_copy || n_raw == 0) return DLT_OK;  // nothing to do: dlt_scan_deskew uploads as usual
    rt::set_device(h->cfg.device);
    h->pf_pending_ptr = pts48;
    h->pf_pending_n = n_raw;
    return DLT_OK;
}

int dlt_scan_deskew(dlt_handle h, const void *pts48, int n_raw, const double *imu_pose22, int n_pose, const double *pose24) {
    return scan_deskew_impl(h, pts48, false, n_raw, imu_pose22, n_pose, pose24);
}
int dlt_scan_deskew_dev(dlt_handle h, const void *pts48_dev, int n_raw, const double *imu_pose22, int n_pose, const double *pose24) {
    return scan_deskew_impl(h, pts48_dev, true, n_raw, imu_pose22, n_pose, pose24);
}

// pcl::VoxelGrid accepts any bounding box whose voxel-index space fits an int (beyond that it passes the cloud through,
// vox_status 1); the occupancy bitmap starts at voxel_bitmap_bits and is grown to what a scan needs (one far outlier
// point, or a small leaf at long range) instead of failing the scan.
static int grow_bitmap(dlt_handle h, long long cells) {
    if (cells <= h->bitmap_bits) return DLT_OK;
    if (cells > 2147483647ll) DLT_FAIL(h, DLT_E_CAPACITY, "VoxelGrid index space beyond PCL's own limit");
    long long bits = h->bitmap_bits;
    while (bits < cells) bits <<= 1;
    const size_t words = (size_t)((bits + 31) / 32);
    const int n_blocks = div_up((long long)words, kScanWordsPerBlock);
    DLT_RT(h, rt::sync(h->stream));
    unsigned *nb = nullptr, *nw = nullptr, *ns = nullptr, *no = nullptr;
    void *v[4] = {nullptr, nullptr, nullptr, nullptr};
    if (rt::alloc(&v[0], words * 4) || rt::alloc(&v[1], words * 4) || rt::alloc(&v[2], (size_t)n_blocks * 4) || rt::alloc(&v[3], (size_t)n_blocks * 4)) {
        for (void *q : v) rt::release(q);
        DLT_FAIL(h, DLT_E_CAPACITY, "VoxelGrid bitmap cannot grow to this scan's bounding box (out of device memory)");
    }
    nb = (unsigned *)v[0], nw = (unsigned *)v[1], ns = (unsigned *)v[2], no = (unsigned *)v[3];
    for (void *old : {(void *)h->d_bitmap, (void *)h->d_wprefix, (void *)h->d_blksum, (void *)h->d_blkoff}) {
        for (size_t k = 0; k < h->allocs.size(); k++)
            if (h->allocs[k] == old) {
                h->allocs.erase(h->allocs.begin() + (long)k);
                break;
            }
        rt::release(old);
    }
    for (void *q : v) h->allocs.push_back(q);
    h->d_bitmap = nb;
    h->d_wprefix = nw;
    h->d_blksum = ns;
    h->d_blkoff = no;
    h->bitmap_bits = bits;
    h->n_scan_blocks = n_blocks;
    DLT_RT(h, rt::fill(h->d_bitmap, 0, words * 4, h->stream));
    return DLT_OK;
}

// the five VoxelGrid kernels, enqueued
static int enqueue_downsample(dlt_handle h) {
    const int n = h->n_raw;
    const int B = 256, G = div_up(n, B);
    ProfScope prof(h, 3);
    // k_vox_final of an earlier downsample of this very scan reset the bounding box: recompute it from the undistorted points
    if (h->sc_clean) DLT_LAUNCH(k_scan_bbox, G, B, h->stream, (const float4 *)h->d_undist, n, h->d_sc);
    DLT_LAUNCH(k_vox_mark, G, B, h->stream, (const float4 *)h->d_undist, n, h->cfg.ds_scan, h->d_sc, h->d_bitmap, h->bitmap_bits, h->d_vidx);
    {
        const int gs = h->n_scan_blocks < 2 * h->n_sm ? h->n_scan_blocks : 2 * h->n_sm;  // the kernel strides over the chunks in use
        DLT_LAUNCH(k_vox_scan1, gs, kScanBlock, h->stream, (const unsigned *)h->d_bitmap, h->d_sc, h->d_wprefix, h->d_blksum, h->d_blkoff, h->d_ticket + 1);
    }
    DLT_LAUNCH(k_vox_accum, G, B, h->stream, (const float4 *)h->d_undist, n, (const ScanScalars *)h->d_sc, (const unsigned *)h->d_bitmap,
               (const unsigned *)h->d_wprefix, (const unsigned *)h->d_blkoff, (const unsigned *)h->d_vidx, h->acc, h->d_vop);
    DLT_LAUNCH(k_vox_final, G, B, h->stream, h->d_sc, h->acc, h->d_bitmap, (const float4 *)h->d_undist, h->d_down, n);
    h->sc_clean = true;
    return DLT_OK;
}

int dlt_scan_downsample(dlt_handle h, int *n_down) {
    if (!h || !n_down) return DLT_E_INVALID;
    if (!h->have_raw) DLT_FAIL(h, DLT_E_STATE, "dlt_scan_downsample before dlt_scan_deskew");
    rt::set_device(h->cfg.device);
    if (int ri = ins_wait_classified(h)) return ri;
    const int n = h->n_raw;
    h->n_down = 0;
    *n_down = 0;
    h->have_match = false;
    h->n_down_on_device = false;
    if (n == 0) {
        h->have_down = true;
        return DLT_OK;
    }
    for (int attempt = 0;; attempt++) {
        enqueue_downsample(h);
        DLT_RT(h, rt::check_launch());
        DLT_RT(h, rt::d2h(h->h_sc, h->d_sc, sizeof(ScanScalars), h->stream));
        DLT_RT(h, rt::sync(h->stream));
        if (h->h_sc->vox_status != 2) break;
        if (attempt > 0) DLT_FAIL(h, DLT_E_CAPACITY, "VoxelGrid bounding box exceeds the occupancy bitmap");
        if (int rg = grow_bitmap(h, h->h_sc->vox_cells)) return rg;
    }
    h->n_down = h->h_sc->n_down;
    h->n_down_hint = h->n_down;
    *n_down = h->n_down;
    h->have_down = true;
    return DLT_OK;
}

int dlt_scan_downsample_async(dlt_handle h) {
    if (!h) return DLT_E_INVALID;
    if (!h->have_raw) DLT_FAIL(h, DLT_E_STATE, "dlt_scan_downsample_async before dlt_scan_deskew");
    rt::set_device(h->cfg.device);
    if (int ri = ins_wait_classified(h)) return ri;
    h->have_match = false;
    h->have_down = true;
    if (h->n_raw == 0) {
        h->n_down = 0;
        h->n_down_on_device = false;
        return DLT_OK;
    }
    enqueue_downsample(h);
    DLT_RT(h, rt::check_launch());
    h->n_down = h->n_raw;  // upper bound until dlt_iekf_update reads feats_down_size back
    h->n_down_on_device = true;
    return DLT_OK;
}

int dlt_scan_get_undistorted(dlt_handle h, float *xyzi, int cap, int *n) {
    if (!h || !n) return DLT_E_INVALID;
    if (!h->have_raw) DLT_FAIL(h, DLT_E_STATE, "no scan");
    *n = h->n_raw;
    int m = h->n_raw < cap ? h->n_raw : cap;
    if (m > 0 && xyzi) {
        DLT_RT(h, rt::d2h(xyzi, h->d_undist, (size_t)m * sizeof(float4), h->stream));
        DLT_RT(h, rt::sync(h->stream));
    }
    return DLT_OK;
}
int dlt_scan_get_down(dlt_handle h, float *xyzi, int cap, int *n) {
    if (!h || !n) return DLT_E_INVALID;
    if (!h->have_down) DLT_FAIL(h, DLT_E_STATE, "no downsampled scan");
    if (int rn = resolve_n_down(h)) return rn;
    *n = h->n_down;
    int m = h->n_down < cap ? h->n_down : cap;
    if (m > 0 && xyzi) {
        DLT_RT(h, rt::d2h(xyzi, h->d_down, (size_t)m * sizeof(float4), h->stream));
        DLT_RT(h, rt::sync(h->stream));
    }
    return DLT_OK;
}
int dlt_scan_set_down(dlt_handle h, const float *xyzi, int n) {
    if (!h || n < 0 || (n > 0 && !xyzi)) return DLT_E_INVALID;
    if (n > h->cap) DLT_FAIL(h, DLT_E_CAPACITY, "scan larger than max_scan_points");
    rt::set_device(h->cfg.device);
    if (int ri = ins_finish(h)) return ri;
    if (n > 0) DLT_RT(h, rt::h2d(h->d_down, xyzi, (size_t)n * sizeof(float4), h->stream));
    h->n_down = n;
    h->have_down = true;
    h->have_match = false;
    return DLT_OK;
}
int dlt_scan_get_voxel_of_point(dlt_handle h, int *slot, int cap) {
    if (!h || !slot) return DLT_E_INVALID;
    if (!h->have_down || !h->have_raw) DLT_FAIL(h, DLT_E_STATE, "no downsampled scan");
    int m = h->n_raw < cap ? h->n_raw : cap;
    if (m > 0) {
        DLT_RT(h, rt::d2h(slot, h->d_vop, (size_t)m * sizeof(int), h->stream));
        DLT_RT(h, rt::sync(h->stream));
    }
    return DLT_OK;
}

// ------------------------------------------------------------------ front end
int dlt_frontend_sample(dlt_handle h, const void *cloud_data, int n_points, const dlt_cloud_layout *lay, int sensor, int point_filter_num,
                        float lidar_min_range, float lidar_max_range, void **pts48_dev, int *n_out, double *timespan_out, double *sweep_span_out,
                        double *stamp_shift_out) {
    if (!h || !lay || !pts48_dev || !n_out || n_points < 0 || (n_points > 0 && !cloud_data) || point_filter_num < 1) return DLT_E_INVALID;
    if (sensor < DLT_SENSOR_VELODYNE || sensor > DLT_SENSOR_ROBOSENSE) return DLT_E_INVALID;
    const int offs[6] = {lay->off_x, lay->off_y, lay->off_z, lay->off_intensity, lay->off_ring, lay->off_time};
    const int t_bytes = sensor == DLT_SENSOR_ROBOSENSE ? 8 : 4;
    if (lay->point_step < 16 || (lay->point_step & 3)) DLT_FAIL(h, DLT_E_INVALID, "point_step must be a multiple of 4");
    for (int k = 0; k < 6; k++) {
        const int width = k == 4 ? 2 : (k == 5 ? t_bytes : (k == 3 && sensor == DLT_SENSOR_ROBOSENSE ? 1 : 4));
        const int align = width >= 4 ? 4 : width;
        if (offs[k] < 0 || offs[k] + width > lay->point_step || (offs[k] % align)) DLT_FAIL(h, DLT_E_INVALID, "field offset outside the record or misaligned");
    }
    rt::set_device(h->cfg.device);
    *pts48_dev = h->d_raw;
    *n_out = 0;
    if (timespan_out) *timespan_out = 0.0;
    if (sweep_span_out) *sweep_span_out = 0.0;
    if (stamp_shift_out) *stamp_shift_out = 0.0;
    if (n_points == 0) return DLT_OK;
    const unsigned char *src = static_cast<const unsigned char *>(cloud_data);
    auto rec = [&](long long i) { return src + (size_t)i * lay->point_step; };
    auto f32 = [](const unsigned char *p) { float v; std::memcpy(&v, p, 4); return v; };
    auto u32 = [](const unsigned char *p) { unsigned v; std::memcpy(&v, p, 4); return v; };
    auto f64 = [](const unsigned char *p) { double v; std::memcpy(&v, p, 8); return v; };
    FeParams P;
    P.lay = *lay;
    P.sensor = sensor;
    P.n_points = n_points;
    P.filter_num = point_filter_num;
    P.min_range = lidar_min_range;
    P.max_range = lidar_max_range;
    P.t0 = 0.0;
    P.vel_time0 = 0.f;
    double timespan = 0.0, ret_timespan = 0.0, shift = 0.0;
    std::vector<int> rs_index;  // RoboSense: indices of the finite records that survive i % point_filter_num
    if (sensor == DLT_SENSOR_VELODYNE) {  // :279: timespan = back().time - points[0].time (float arithmetic), zeroed after use (:299)
        const float tb = f32(rec(n_points - 1) + lay->off_time), t0 = f32(rec(0) + lay->off_time);
        timespan = (double)(tb - t0);
        ret_timespan = 0.0;
    } else if (sensor == DLT_SENSOR_LIVOX) {  // :311
        timespan = (double)f32(rec(n_points - 1) + lay->off_time);
        ret_timespan = timespan;
    } else if (sensor == DLT_SENSOR_OUSTER) {  // :333 (the second to last record), :346
        if (n_points < 2) DLT_FAIL(h, DLT_E_INVALID, "an Ouster cloud needs at least two records");
        timespan = (double)u32(rec(n_points - 2) + lay->off_time);
        ret_timespan = timespan * (double)1e-9f;
    } else {  // :360, :383
        P.t0 = f64(rec(0) + lay->off_time);
        timespan = f64(rec(n_points - 1) + lay->off_time) - P.t0;
        ret_timespan = timespan;
        shift = timespan;
        int kept = 0;
        for (int i = 0; i < n_points; i++) {  // inputCloud->push_back skips non-finite records (:369-370) before the 1-in-N sampling
            const unsigned char *r = rec(i);
            const float x = f32(r + lay->off_x), y = f32(r + lay->off_y), z = f32(r + lay->off_z);
            if (!std::isfinite(x) || !std::isfinite(y) || !std::isfinite(z)) continue;
            if (kept % point_filter_num == 0) rs_index.push_back(i);
            kept++;
        }
    }
    P.timespan = timespan;
    const int n_cand = sensor == DLT_SENSOR_ROBOSENSE ? (int)rs_index.size() : (n_points + point_filter_num - 1) / point_filter_num;
    P.n_cand = n_cand;
    if (timespan_out) *timespan_out = ret_timespan;
    if (stamp_shift_out) *stamp_shift_out = shift;
    if (sweep_span_out) *sweep_span_out = (double)(sensor == DLT_SENSOR_OUSTER ? (float)(timespan * (double)1e-9f) : (float)timespan);
    if (n_cand == 0) return DLT_OK;
    if (n_cand > h->cap) DLT_FAIL(h, DLT_E_CAPACITY, "sampled cloud larger than max_scan_points");
    // staging (grown on demand; the sensor cloud is the only thing that crosses PCIe)
    const size_t bytes = (size_t)n_points * lay->point_step;
    if (bytes > h->sensor_cap) {
        DLT_RT(h, rt::sync(h->stream));
        rt::release(h->d_sensor);
        void *v = nullptr;
        size_t want = bytes + bytes / 4;
        if (rt::alloc(&v, want) != 0) DLT_FAIL(h, DLT_E_CUDA, "sensor cloud staging allocation failed");
        h->d_sensor = (unsigned char *)v;
        h->sensor_cap = want;
    }
    if ((size_t)n_cand > h->fe_cap) {
        DLT_RT(h, rt::sync(h->stream));
        for (void *q : {(void *)h->d_fe_tmp, (void *)h->d_fe_keep, (void *)h->d_fe_pos, (void *)h->d_fe_blk, (void *)h->d_fe_blkoff, (void *)h->d_fe_index}) rt::release(q);
        const size_t c = (size_t)h->cap;
        void *v[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
        const size_t nblk = (c + kFeBlock - 1) / kFeBlock;
        if (rt::alloc(&v[0], c * 3 * sizeof(float4)) || rt::alloc(&v[1], c) || rt::alloc(&v[2], c * 4) || rt::alloc(&v[3], nblk * 4) ||
            rt::alloc(&v[4], nblk * 4) || rt::alloc(&v[5], c * 4))
            DLT_FAIL(h, DLT_E_CUDA, "front end scratch allocation failed");
        h->d_fe_tmp = (float4 *)v[0];
        h->d_fe_keep = (unsigned char *)v[1];
        h->d_fe_pos = (unsigned *)v[2];
        h->d_fe_blk = (unsigned *)v[3];
        h->d_fe_blkoff = (unsigned *)v[4];
        h->d_fe_index = (int *)v[5];
        h->fe_cap = c;
    }
    DLT_RT(h, rt::h2d(h->d_sensor, cloud_data, bytes, h->stream));
    const int *d_index = nullptr;
    if (sensor == DLT_SENSOR_ROBOSENSE) {
        DLT_RT(h, rt::h2d(h->d_fe_index, rs_index.data(), rs_index.size() * sizeof(int), h->stream));
        DLT_RT(h, rt::sync(h->stream));  // rs_index is a local
        d_index = h->d_fe_index;
    }
    const int G = div_up(n_cand, kFeBlock);
    DLT_LAUNCH(k_fe_convert, G, kFeBlock, h->stream, P, (const unsigned char *)h->d_sensor, d_index, h->d_fe_tmp, h->d_fe_keep, h->d_fe_pos, h->d_fe_blk);
    DLT_LAUNCH(k_fe_offsets, 1, 1024, h->stream, (const unsigned *)h->d_fe_blk, G, h->d_fe_blkoff, h->d_counters + 11);
    DLT_LAUNCH(k_fe_scatter, G, kFeBlock, h->stream, n_cand, (const float4 *)h->d_fe_tmp, (const unsigned char *)h->d_fe_keep,
               (const unsigned *)h->d_fe_pos, (const unsigned *)h->d_fe_blkoff, h->d_raw, h->cap);
    DLT_RT(h, rt::check_launch());
    DLT_RT(h, rt::d2h(h->h_ints + 56, h->d_counters + 11, sizeof(int), h->stream));  // (a slot of its own: [32..35] hold the zero-copy flags)
    DLT_RT(h, rt::sync(h->stream));
    *n_out = h->h_ints[56];
    return DLT_OK;
}

int dlt_frontend_read(dlt_handle h, void *pts48, int n) {
    if (!h || n < 0 || (n > 0 && !pts48) || n > h->cap) return DLT_E_INVALID;
    rt::set_device(h->cfg.device);
    if (n > 0) {
        DLT_RT(h, rt::d2h(pts48, h->d_raw, (size_t)n * 48, h->stream));
        DLT_RT(h, rt::sync(h->stream));
    }
    return DLT_OK;
}

// ------------------------------------------------------------------ measurement model
static int measure_dev_impl(dlt_handle h, const double *pose24, int do_match, double *result_dev, bool zc, bool *launched) {
    if (launched) *launched = false;
    if (!h || !pose24 || !result_dev) return DLT_E_INVALID;
    if (!h->have_down) DLT_FAIL(h, DLT_E_STATE, "dlt_measure before a downsampled scan is set");
    if (!do_match && !h->have_match) DLT_FAIL(h, DLT_E_STATE, "dlt_measure(do_match=0) before any match pass");
    if (int rd = map_refuse_if_dead(h)) return rd;
    rt::set_device(h->cfg.device);
    if (int ri = ins_finish(h)) return ri;
    // right behind dlt_scan_downsample_async feats_down_size is still on the device: the kernels read it there (grids from
    // an estimate / an upper bound) and it comes back with the result block, so no synchronisation is spent on it
    if (int rp = issue_prefetch(h)) return rp;  // (this scan's uploads are queued by now)
    const bool dev_n = h->n_down_on_device;
    const int n = h->n_down;
    Pose P = pose_from(pose24);
    if (!dev_n && n == 0) {
        DLT_RT(h, rt::fill(result_dev, 0, kResultDoubles * sizeof(double), h->stream));
        h->have_match = true;
        return DLT_OK;
    }
    if (result_dev == h->d_result) {
        h->eig_valid = false;
        if (int rj = eig_join(h)) return rj;
    }
    LoopArgs la = {nullptr, dev_n ? &h->d_sc->n_down : nullptr, dev_n ? &h->d_sc->vox_status : nullptr, 0, 0, 0ull, 0ull, h->peer_on ? h->d_peer : nullptr};
    int n_grid = n;
    if (dev_n) {
        n_grid = h->n_down_hint > 0 ? (int)(1.25 * h->n_down_hint) + 1024 : h->n_raw;
        if (n_grid > h->n_raw) n_grid = h->n_raw;
    }
    if (do_match) {
        h->nfar_known = false;
        ProfScope prof(h, 0);
        // a rematch pass of the same scan against the same map: neighbour sets that can be proven unchanged are not searched again
        // ... unless the pose moved so far since the pass that found them that few proofs can succeed (the proof needs twice the
        // displacement to fit between the 5th and the 6th neighbour distance, a few centimetres on a 0.5 m map): then search
        bool reuse = h->have_match && h->knn_reuse && h->reuse_holdoff == 0;
        if (reuse) {
            double tr = 0.0, dt2 = 0.0;
            for (int i = 0; i < 3; i++) {
                for (int j = 0; j < 3; j++) tr += h->match_pose[3 * i + j] * pose24[3 * i + j];  // trace(R1^T R2)
                const double d = pose24[9 + i] - h->match_pose[9 + i];
                dt2 += d * d;
            }
            const double c = 0.5 * (tr - 1.0);
            const double ang = c >= 1.0 ? 0.0 : std::acos(c < -1.0 ? -1.0 : c);
            if (std::sqrt(dt2) + ang * 30.0 > 0.06 * h->cfg.ds_map) reuse = false;  // displacement of a point 30 m out: 3 cm on a 0.5 m map
        }
        std::memcpy(h->match_pose, pose24, sizeof(h->match_pose));
        int rk = launch_knn(h, (const float4 *)h->d_down, n, n_grid, 1, P, la, reuse);
        if (rk) return rk;
        h->have_match = true;
    }
    MeasureBufs mb;
    mb.down = h->d_down;
    mb.nbr = h->knn.nbr;
    mb.flags = h->knn.flags;
    mb.plane = h->d_plane;
    mb.coeff = h->d_coeff;
    mb.sel = h->d_sel;
    mb.eff = h->d_eff;
    mb.partials = h->d_partials;
    mb.ticket = h->d_ticket;
    mb.far_count = h->d_counters + 5;
    mb.unres_count = h->d_counters + 8;
    mb.result = result_dev;
    mb.zc_result = nullptr;
    mb.zc_flag = nullptr;
    mb.zc_seq = 0;
    if (zc) {  // (pinned allocations are addressable from the device under unified addressing)
        mb.zc_result = h->h_result;
        mb.zc_flag = reinterpret_cast<unsigned long long *>(h->h_ints + 32);
        mb.zc_seq = ++h->zc_seq;
    }
    int G = div_up(dev_n ? h->n_raw : n, kResidBlock);  // surplus blocks return at once
    if (G < 1) G = 1;
    ProfScope prof(h, 1);
    if (h->cfg.extrinsic_est_en)
        DLT_LAUNCH(k_residual<true>, G, kResidBlock, h->stream, mb, n, do_match ? 1 : 0, P, h->cfg.plane_thr, la);
    else
        DLT_LAUNCH(k_residual<false>, G, kResidBlock, h->stream, mb, n, do_match ? 1 : 0, P, h->cfg.plane_thr, la);
    DLT_RT(h, rt::check_launch());
    if (launched) *launched = true;
    return DLT_OK;
}

int dlt_measure_dev(dlt_handle h, const double *pose24, int do_match, double *result_dev) {
    return measure_dev_impl(h, pose24, do_match, result_dev, false, nullptr);
}

// ------------------------------------------------------------------ the iteration loop on the device
#if !defined(DLT_EMU)
// WHILE(cond_while) { IF(cond_match) { k_knn8; k_knn }  k_residual (+ fused solve / control step) }
// Both conditions default to 1 at every launch (iteration 0 always matches) and are then driven by the fused step.
static bool build_loop_graph(dlt_handle h, const MeasureBufs &mb) {
    cudaGraph_t g = nullptr;
    cudaGraphConditionalHandle hw = 0, hm = 0;
    cudaStream_t cs = h->own_stream;
    bool capturing = false;
    auto fail = [&]() {
        if (capturing) {
            cudaGraph_t junk = nullptr;
            cudaStreamEndCapture(cs, &junk);
        }
        cudaGetLastError();
        if (g) cudaGraphDestroy(g);
        return false;
    };
    if (cudaGraphCreate(&g, 0) != cudaSuccess) return fail();
    if (cudaGraphConditionalHandleCreate(&hw, g, 1, cudaGraphCondAssignDefault) != cudaSuccess) return fail();
    if (cudaGraphConditionalHandleCreate(&hm, g, 1, cudaGraphCondAssignDefault) != cudaSuccess) return fail();
    cudaGraphNodeParams wp = {};
    wp.type = cudaGraphNodeTypeConditional;
    wp.conditional.handle = hw;
    wp.conditional.type = cudaGraphCondTypeWhile;
    wp.conditional.size = 1;
    cudaGraphNode_t wnode = nullptr;
    if (cudaGraphAddNode(&wnode, g, nullptr, 0, &wp) != cudaSuccess) return fail();
    cudaGraph_t wbody = wp.conditional.phGraph_out[0];
    cudaGraphNodeParams ip = {};
    ip.type = cudaGraphNodeTypeConditional;
    ip.conditional.handle = hm;
    ip.conditional.type = cudaGraphCondTypeIf;
    ip.conditional.size = 1;
    cudaGraphNode_t inode = nullptr;
    if (cudaGraphAddNode(&inode, wbody, nullptr, 0, &ip) != cudaSuccess) return fail();
    cudaGraph_t ibody = ip.conditional.phGraph_out[0];

    LoopArgs la = {h->d_iekf, &h->d_sc->n_down, &h->d_sc->vox_status, 1, 1, (unsigned long long)hw, (unsigned long long)hm, nullptr};
    Pose P = {};
    const unsigned long long launches_before = rt::g_launches.load();
    // fixed grids: two full waves of k_knn8 (no stride pass up to 24 * SMs * 16 queries), surplus blocks return at once
    const int g8 = 24 * h->n_sm, gk = 8 * h->n_sm, gr = div_up(h->cap, kResidBlock);
    if (cudaStreamBeginCaptureToGraph(cs, ibody, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal) != cudaSuccess) return fail();
    capturing = true;
    k_knn8<<<g8, kKnn8Block, 0, cs>>>(h->map, (const float4 *)h->d_down, 0, 1, P, h->cfg.max_sq_dist, h->knn, h->d_unres, h->d_counters + 8, la);
    k_knn<<<gk, kKnnWarps * 32, 0, cs>>>(h->map, (const float4 *)h->d_down, 0, 1, P, h->cfg.max_sq_dist, h->knn, (const int *)h->d_unres,
                                           (const int *)(h->d_counters + 8), la, 0);
    cudaGraph_t out = nullptr;
    if (cudaStreamEndCapture(cs, &out) != cudaSuccess) {
        capturing = false;
        return fail();
    }
    capturing = false;
    if (cudaStreamBeginCaptureToGraph(cs, wbody, &inode, nullptr, 1, cudaStreamCaptureModeThreadLocal) != cudaSuccess) return fail();
    capturing = true;
    if (h->cfg.extrinsic_est_en)
        k_residual<true><<<gr, kResidBlock, 0, cs>>>(mb, 0, 0, P, h->cfg.plane_thr, la);
    else
        k_residual<false><<<gr, kResidBlock, 0, cs>>>(mb, 0, 0, P, h->cfg.plane_thr, la);
    if (cudaStreamEndCapture(cs, &out) != cudaSuccess) {
        capturing = false;
        return fail();
    }
    capturing = false;
    rt::g_launches = launches_before;
    cudaGraphExec_t exec = nullptr;
    if (cudaGraphInstantiate(&exec, g, 0) != cudaSuccess) return fail();
    h->loop_graph = g;
    h->loop_exec = exec;
    return true;
}
#endif

int dlt_iekf_update(dlt_handle h, dlt_iekf_block *blk, dlt_reduce_fn reduce, void *reduce_ctx, double *result_dev) {
    if (!h || !blk) return DLT_E_INVALID;
    if (!h->have_down) DLT_FAIL(h, DLT_E_STATE, "dlt_iekf_update before a downsampled scan is set");
    if (blk->max_iteration < 1 || blk->max_iteration > DLT_IEKF_MAX_ITER) DLT_FAIL(h, DLT_E_INVALID, "max_iteration out of range");
    if (int rd = map_refuse_if_dead(h)) return rd;
    rt::set_device(h->cfg.device);
    if (int ri = ins_finish(h)) return ri;
    if (h->peer_on) reduce = nullptr;  // the sum over ranks happens inside k_residual (peer mailboxes), the solve step stays fused
    const int n_iter = blk->max_iteration;
    // host -> device: everything up to (not including) the out fields
    blk->n_iters = blk->converged = blk->ekf_stop = blk->have_gain = blk->status = 0;
    blk->n_down = blk->n_unresolved = blk->reserved1 = 0;
    blk->iter = blk->rematch_num = blk->rematch_en = blk->done = 0;
    blk->insert_status = 0;
    // the device-side finish needs an unsharded map (a sharded map_incremental exchanges the per-point decisions first)
    // ... or peer mailboxes, through which the kernels exchange them themselves
    const bool finish = blk->finish && (h->map.shard_count <= 1 || h->peer_on);
    const bool sharded_finish = finish && h->map.shard_count > 1;
    blk->finish = finish ? 1 : 0;
    const bool far_q = finish && h->far_hint > 0;
    blk->far_enqueued = far_q ? 1 : 0;
    const size_t up_bytes = offsetof(dlt_iekf_block, blend_state);
    std::memcpy(h->h_iekf, blk, up_bytes);
    DLT_RT(h, rt::h2d(&h->d_iekf->b, h->h_iekf, up_bytes, h->stream));
    if (int rp = issue_prefetch(h)) return rp;  // (this scan's uploads are queued by now)

    // speculative grids: feats_down_size may still be on the device
    int n_grid = h->n_down;
    if (h->n_down_on_device) {
        n_grid = h->n_down_hint > 0 ? (int)(1.25 * h->n_down_hint) + 1024 : h->n_raw;
        if (n_grid > h->n_raw) n_grid = h->n_raw;
    }
    const int n_upper = h->n_down_on_device ? h->n_raw : h->n_down;  // k_residual: one point per thread, surplus blocks exit
    Pose P = {};
    MeasureBufs mb;
    mb.down = h->d_down;
    mb.nbr = h->knn.nbr;
    mb.flags = h->knn.flags;
    mb.plane = h->d_plane;
    mb.coeff = h->d_coeff;
    mb.sel = h->d_sel;
    mb.eff = h->d_eff;
    mb.partials = h->d_partials;
    mb.ticket = h->d_ticket;
    mb.far_count = h->d_counters + 5;
    mb.unres_count = h->d_counters + 8;
    double *res = result_dev ? result_dev : h->d_result;
    mb.result = res;
    mb.zc_result = nullptr;
    mb.zc_flag = nullptr;
    mb.zc_seq = 0;
    // without a reduction over ranks between them the solve step rides in the last block of k_residual
    LoopArgs la = {h->d_iekf, &h->d_sc->n_down, &h->d_sc->vox_status, reduce ? 0 : 1, 0, 0ull, 0ull, h->peer_on ? h->d_peer : nullptr};
    if (!h->n_down_on_device) {  // the scan was set with a host-known size: publish it where the kernels look
        DLT_RT(h, rt::h2d(&h->d_sc->n_down, &h->n_down, sizeof(int), h->stream));
    }
    h->eig_valid = false;
    h->nfar_known = false;
    if (int rj = eig_join(h)) return rj;
    int G = div_up(n_upper, kResidBlock);
    if (G < 1) G = 1;
    bool graph_run = false;
#if !defined(DLT_EMU)
    if (!reduce && !h->peer_on && res == h->d_result && h->use_graph && !h->prof_on) {
        if (h->loop_graph_state == 0) {
            DLT_RT(h, rt::sync(h->own_stream));
            h->loop_graph_state = build_loop_graph(h, mb) ? 1 : -1;
        }
        if (h->loop_graph_state == 1) {
            if (cudaGraphLaunch(h->loop_exec, h->stream) != cudaSuccess) {
                cudaGetLastError();
                h->loop_graph_state = -1;
            } else {
                graph_run = true;
            }
        }
    }
#endif
    // ---- fused loop kernels: match pass + residual pass + solve of an iteration in one launch, or the whole loop in one
    //      cooperative launch with the eigen-decomposition inside (dlt_loop_kernels.cuh).  Not with a reduce callback (the
    //      library collective has to run between the residual pass and the solve) and not while per-kernel profiling is on.
    bool fused_run = false, eig_in_kernel = false;
    if (!graph_run && !reduce && res == h->d_result && h->loop_fused && !h->prof_on) {
        if (int rs = size_loop_grid(h)) return rs;
        LoopK lk;
        lk.m = h->map;
        lk.down = h->d_down;
        lk.knn = h->knn;
        lk.mb = mb;
        lk.dev = h->d_iekf;
        lk.n_ptr = &h->d_sc->n_down;
        lk.vox_ptr = &h->d_sc->vox_status;
        lk.max_sq_dist = h->cfg.max_sq_dist;
        lk.plane_thr = h->cfg.plane_thr;
        lk.peer = h->peer_on ? h->d_peer : nullptr;
        lk.bar = h->d_bar;
        lk.eig_out = h->d_result + kEigOffset;
        lk.max_iter = n_iter;
        // far_count, the add counters, the hand-over counter and the nearest-point counters in one memset (d_counters[5..10])
        DLT_RT(h, rt::fill(h->d_counters + 5, 0, 6 * sizeof(int), h->stream));
        const bool ext = h->cfg.extrinsic_est_en != 0;
#if !defined(DLT_EMU)
        if (h->loop_coop) {
            void *args[] = {(void *)&lk};
            const void *fn = ext ? (const void *)k_iekf_loop<true> : (const void *)k_iekf_loop<false>;
            if (cudaLaunchCooperativeKernel(fn, dim3(h->loop_blocks), dim3(kLoopBlock), args, 0, h->stream) == cudaSuccess) {
                ++rt::g_launches;
                fused_run = eig_in_kernel = true;
            } else {
                cudaGetLastError();
                h->loop_coop = 0;  // (e.g. a device without cooperative launch): one launch per iteration from here on
            }
        }
#endif
        if (!fused_run) {
            for (int it = 0; it < n_iter; it++) {
                if (ext)
                    DLT_LAUNCH(k_iekf_iter<true>, h->loop_blocks, kLoopBlock, h->stream, lk);
                else
                    DLT_LAUNCH(k_iekf_iter<false>, h->loop_blocks, kLoopBlock, h->stream, lk);
            }
            fused_run = true;
        }
    }
    for (int it = 0; it < n_iter && !graph_run && !fused_run; it++) {
        {
            ProfScope prof(h, 0);
            int rk = launch_knn(h, (const float4 *)h->d_down, 0, n_grid, 1, P, la);
            if (rk) return rk;
        }
        {
            ProfScope prof(h, 1);
            if (h->cfg.extrinsic_est_en)
                DLT_LAUNCH(k_residual<true>, G, kResidBlock, h->stream, mb, 0, 0, P, h->cfg.plane_thr, la);
            else
                DLT_LAUNCH(k_residual<false>, G, kResidBlock, h->stream, mb, 0, 0, P, h->cfg.plane_thr, la);
        }
        if (reduce) {
            if (reduce(reduce_ctx, res, kNormalEqDoubles) != 0) DLT_FAIL(h, DLT_E_STATE, "reduce callback failed");
            ProfScope profs(h, 6);
            DLT_LAUNCH(k_iekf_step, 1, kIekfBlock, h->stream, h->d_iekf, (const double *)res, (const int *)&h->d_sc->n_down,
                       (const int *)&h->d_sc->vox_status, h->cfg.extrinsic_est_en ? 12 : 6);
        }
    }
    if (res != h->d_result) DLT_RT(h, rt::d2d(h->d_result, res, kNormalEqDoubles * sizeof(double), h->stream));  // for dlt_degeneracy
    DLT_RT(h, rt::check_launch());
    // degeneracy output: fork the eigen-decomposition of the last normal equations onto the side stream now, so that
    // it runs while the host waits for / digests the block below
    h->have_match = true;
    if (eig_in_kernel) {  // k_iekf_loop left the decomposition of the last iteration's pose block next to the normal equations
        DLT_RT(h, rt::d2h(h->h_result + kEigOffset, h->d_result + kEigOffset, 42 * sizeof(double), h->stream));
        h->eig_valid = true;
        h->eig_pending = false;
        h->eig_zc = false;
    } else if (int re = dlt_degeneracy_begin(h)) {
        return re;
    }
    if (finish) {  // ---- map_incremental (:582-630, 1164-1168) behind the loop, armed by the last k_iekf_step
        IekfDev *ctl = h->d_iekf;
        if (far_q) {
            ProfScope prof(h, 5);
            DLT_LAUNCH(k_nn1_seed, h->n_sm, kSeedBlock, h->stream, h->map, (const int *)h->map.n_buckets, (const float4 *)h->knn.qw,
                       (const int *)h->knn.nn_list, (const int *)h->knn.nn_count, h->knn.nn_key, &ctl->b.insert_status);
            DLT_LAUNCH(k_nn1, dim3(kNn1GroupsX, 8 * h->n_sm), kNn1Block, h->stream, h->map, (const int *)h->map.n_buckets, (const float4 *)h->knn.qw,
                       (const int *)h->knn.nn_list, (const int *)h->knn.nn_count, h->knn.nn_key, &ctl->b.insert_status);
        }
        if (!fused_run) DLT_RT(h, rt::fill(h->d_counters + 6, 0, 2 * sizeof(int), h->stream));  // downsample adds, raw adds (far_count is re-armed by the next k_knn8)
        // unsharded: the classification kernel also claims cells and places the voxel bids.  Sharded: the owners' decisions are
        // exchanged first (k_incr_push / k_incr_pull through the peer mailboxes, gated like everything here), then every rank
        // inserts what falls into its tiles + halo
        FuseInsert fi = {1, h->scratch, h->d_cellslot, h->d_vslot};  // (sharded: the cells are claimed and the bids placed by k_incr_pull)
        if (int rs = prepare_scratch(h, n_upper, true, &fi.sc)) return rs;
        const FuseInsert fi_pull = fi;
        if (sharded_finish) fi.on = 0;
        DLT_LAUNCH(k_incr_classify, div_up(n_upper, 256), 256, h->stream, (const float4 *)h->d_down, 0, P, (const float4 *)h->knn.nbr,
                   (const int *)h->knn.nbr_cnt, (double)h->cfg.ds_map, 0, h->d_pw, h->d_dsflag, h->d_addflag, h->d_counters + 6, la, h->map,
                   (const int *)h->map.n_live, (const unsigned char *)h->knn.flags, (const int *)h->knn.nn_pos, (const unsigned long long *)h->knn.nn_key, fi);
        InsertGate gate = {&h->d_iekf->b.insert_status, &h->d_sc->n_down};
        if (sharded_finish) {
            const int Gx = div_up(n_upper, 256);
            DLT_LAUNCH(k_incr_push, Gx, 256, h->stream, h->d_peer, (const unsigned char *)h->d_dsflag, (const unsigned char *)h->d_addflag,
                       (const unsigned char *)h->knn.flags, 0, (unsigned char)kFlagForeign, (const int *)gate.go, (const int *)gate.n_ptr);
            DLT_RT(h, rt::fill(h->d_counters + 6, 0, 2 * sizeof(int), h->stream));
            DLT_LAUNCH(k_incr_pull, Gx, 256, h->stream, h->d_peer, 0, h->d_dsflag, h->d_addflag, h->d_counters + 6, (const int *)gate.go, (const int *)gate.n_ptr,
                       h->map, (const float4 *)h->d_pw, fi_pull);
        }
        int ri = insert_points(h, h->d_pw, n_upper, true, gate, &fi_pull.sc);
        if (ri) return ri;
        DLT_RT(h, rt::d2h(h->h_ints, h->d_counters, 8 * sizeof(int), h->stream));
    }
    // device -> host: in/out + out fields + the iteration records that can have been written
    const size_t lo = offsetof(dlt_iekf_block, state);
    const size_t hi = offsetof(dlt_iekf_block, iters) + (size_t)n_iter * sizeof(dlt_iekf_iter);
    DLT_RT(h, rt::d2h((char *)h->h_iekf + lo, (const char *)&h->d_iekf->b + lo, hi - lo, h->stream));
    if (h->peer_on) DLT_RT(h, rt::d2h(h->h_peer_status, &h->d_peer->status, sizeof(int), h->stream));
    DLT_RT(h, rt::sync(h->stream));
    if (h->h_iekf->reserved1 == 2 && !h->vox_retry) {  // the occupancy bitmap was too small for this scan: grow it, run the
        int nd = 0;                                   // VoxelGrid again and repeat the update from the caller's untouched block
        if (int rg = dlt_scan_downsample(h, &nd)) return rg;
        h->vox_retry = true;
        const int r2 = dlt_iekf_update(h, blk, reduce, reduce_ctx, result_dev);
        h->vox_retry = false;
        return r2;
    }
    std::memcpy((char *)blk + lo, (const char *)h->h_iekf + lo, hi - lo);
    if (h->peer_on && *h->h_peer_status != 0) DLT_FAIL(h, DLT_E_STATE, "peer exchange timed out (a rank did not take part in this update)");
    if (graph_run) {  // kernels the graph actually ran: one residual pass per iteration, two kNN kernels per match pass
        int n_match = 0;
        for (int k = 0; k < blk->n_iters && k < DLT_IEKF_MAX_ITER; k++) n_match += blk->iters[k].did_match ? 1 : 0;
        rt::g_launches += (unsigned long long)(blk->n_iters + 2 * n_match);
    }
    if (h->n_down_on_device && blk->n_iters == 0) {  // the loop never ran: fetch feats_down_size the plain way
        if (int rn = resolve_n_down(h)) return rn;
        blk->n_down = h->n_down;
    }
    h->n_down = blk->n_down;
    h->n_down_hint = blk->n_down;
    h->n_down_on_device = false;
    h->have_match = blk->n_iters > 0;
    if (finish) {
        if (blk->insert_status == 1) {  // the map changed: neighbour sets are stale, the counters are fresh
            h->have_match = false;
            h->counters_fresh = true;
            blk->n_added_ds = h->h_ints[6];
            blk->n_added_raw = h->h_ints[7];
            for (int i = 0; i < 5; i++) blk->map_counters[i] = h->h_ints[i];
            if (h->h_ints[2] != 0) h->map_dead = true;
            if (h->h_ints[2] == 1) DLT_FAIL(h, DLT_E_CAPACITY, "map bucket pool exhausted (raise max_map_points)");
            if (h->h_ints[2] == 2) DLT_FAIL(h, DLT_E_CAPACITY, "map hash table full (raise max_map_points)");
        } else {
            h->counters_fresh = true;  // nothing was inserted: the read-back still mirrors the counters
            blk->n_added_ds = blk->n_added_raw = 0;
        }
    }
    h->h_last_nfar = blk->n_unresolved;
    h->far_hint = blk->n_unresolved;
    h->nfar_known = blk->n_iters > 0;
    if (blk->reserved1 == 2) DLT_FAIL(h, DLT_E_CAPACITY, "VoxelGrid bounding box exceeds the occupancy bitmap");
    if (blk->status == 1) DLT_FAIL(h, DLT_E_STATE, "H^T H + (P/R)^-1 is singular");
    return DLT_OK;
}

int dlt_measure(dlt_handle h, const double *pose24, int do_match, dlt_measure_out *out) {
    if (!h || !out) return DLT_E_INVALID;
    const bool zc = h->zerocopy && !h->prof_on;  // (with peers the block that is published is the sum over the ranks)
    bool launched = false;
    int rc = measure_dev_impl(h, pose24, do_match, h->d_result, zc, &launched);
    if (rc) return rc;
    if (zc && launched) {  // k_residual's last block stores the block into h_result and then this launch's number into the flag
        const volatile unsigned long long *flag = reinterpret_cast<const volatile unsigned long long *>(h->h_ints + 32);
        const unsigned long long want = h->zc_seq;
        for (unsigned spins = 1; *flag != want; spins++) {
            if (rt::t_wait_hook) rt::t_wait_hook(rt::t_wait_ctx);  // (the multi-sequence driver runs another sequence meanwhile)
#if defined(__x86_64__)
            __builtin_ia32_pause();
#endif
            if ((spins & 0xFFFu) == 0u) {  // a kernel that died never raises the flag: ask the stream now and then
                const int q = rt::stream_query(h->stream);
                if (q == 2) DLT_FAIL(h, DLT_E_CUDA, std::string("stream error while waiting for the result block: ") + rt::last_error());
                if (q == 0 && *flag != want) {  // the kernel returned without publishing: with peers that is an exchange that timed out
                    if (h->peer_on) {
                        DLT_RT(h, rt::d2h(h->h_peer_status, &h->d_peer->status, sizeof(int), h->stream));
                        DLT_RT(h, rt::sync(h->stream));
                        if (*h->h_peer_status != 0) DLT_FAIL(h, DLT_E_STATE, "peer exchange timed out (a rank did not take part in this evaluation)");
                    }
                    DLT_FAIL(h, DLT_E_CUDA, "the stream drained without the result flag");
                }
            }
        }
        std::atomic_thread_fence(std::memory_order_acquire);
    } else {
        DLT_RT(h, rt::d2h(h->h_result, h->d_result, kFetchDoubles * sizeof(double), h->stream));
        if (h->peer_on) DLT_RT(h, rt::d2h(h->h_peer_status, &h->d_peer->status, sizeof(int), h->stream));
        DLT_RT(h, rt::sync(h->stream));
    }
    if (h->peer_on && *h->h_peer_status != 0) DLT_FAIL(h, DLT_E_STATE, "peer exchange timed out (a rank did not take part in this evaluation)");
    const double *R = h->h_result;
    if (int rn = adopt_n_down(h, R)) {
        if (rn > 0) return rn;
        if (h->vox_retry) DLT_FAIL(h, DLT_E_CAPACITY, "VoxelGrid bounding box exceeds the occupancy bitmap");
        int nd = 0;
        if (int rg = dlt_scan_downsample(h, &nd)) return rg;  // grows the bitmap, runs the VoxelGrid again (synchronously)
        h->vox_retry = true;
        const int r2 = dlt_measure(h, pose24, do_match, out);
        h->vox_retry = false;
        return r2;
    }
    for (int i = 0; i < 144; i++) out->HtH[i] = R[i];
    for (int i = 0; i < 12; i++) out->Htr[i] = R[144 + i];
    out->effct_feat_num = (int)(R[156] + 0.5);
    out->total_residual = R[157];
    out->n_down = h->n_down;
    if (do_match) {
        h->h_last_nfar = (h->n_down > 0) ? (int)(R[158] + 0.5) : 0;
        h->nfar_known = true;
    }
    out->n_unresolved = h->nfar_known ? h->h_last_nfar : 0;
    out->reserved = 0;
    return DLT_OK;
}

int dlt_fetch_result(dlt_handle h, const double *result_dev, dlt_measure_out *out) {
    if (!h || !result_dev || !out) return DLT_E_INVALID;
    rt::set_device(h->cfg.device);
    DLT_RT(h, rt::d2h(h->h_result, result_dev, kFetchDoubles * sizeof(double), h->stream));
    DLT_RT(h, rt::sync(h->stream));
    const double *R = h->h_result;
    if (int rn = adopt_n_down(h, R)) {
        if (rn > 0) return rn;
        int nd = 0;  // grow the bitmap for the scans to come; this evaluation (already reduced over the ranks) cannot be redone here
        if (int rg = dlt_scan_downsample(h, &nd)) return rg;
        DLT_FAIL(h, DLT_E_CAPACITY, "VoxelGrid occupancy bitmap was too small for this scan and has been grown: evaluate the scan again");
    }
    for (int i = 0; i < 144; i++) out->HtH[i] = R[i];
    for (int i = 0; i < 12; i++) out->Htr[i] = R[144 + i];
    out->effct_feat_num = (int)(R[156] + 0.5);
    out->total_residual = R[157];
    out->n_down = h->n_down;
    // R[158]: this rank's unresolved-query count (the all-reduce covers the first 158 doubles only)
    out->n_unresolved = (int)(R[158] + 0.5);
    out->reserved = 0;
    if (result_dev != h->d_result) {  // keep a copy so that dlt_degeneracy works on the reduced normal equations
        if (int rj = eig_join(h)) return rj;
        DLT_RT(h, rt::d2d(h->d_result, result_dev, kNormalEqDoubles * sizeof(double), h->stream));
        h->eig_valid = false;
    }
    return DLT_OK;
}

int dlt_effective_points(dlt_handle h, float *xyzi, float *coeff, int cap, int *n) {
    if (!h || !n) return DLT_E_INVALID;
    if (!h->have_match) DLT_FAIL(h, DLT_E_STATE, "no measurement yet");
    if (int rn = resolve_n_down(h)) return rn;
    const int nd = h->n_down;
    std::vector<unsigned char> eff(nd);
    std::vector<float4> pts(nd), cf(nd);
    if (nd > 0) {
        DLT_RT(h, rt::d2h(eff.data(), h->d_eff, (size_t)nd, h->stream));
        DLT_RT(h, rt::d2h(pts.data(), h->d_down, (size_t)nd * sizeof(float4), h->stream));
        DLT_RT(h, rt::d2h(cf.data(), h->d_coeff, (size_t)nd * sizeof(float4), h->stream));
        DLT_RT(h, rt::sync(h->stream));
    }
    int m = 0;
    for (int i = 0; i < nd; i++) {
        if (!eff[i]) continue;
        if (m < cap) {
            if (xyzi) std::memcpy(xyzi + 4 * (size_t)m, &pts[i], 16);
            if (coeff) std::memcpy(coeff + 4 * (size_t)m, &cf[i], 16);
        }
        m++;
    }
    *n = m;
    return DLT_OK;
}

int dlt_get_nearest(dlt_handle h, float *nbr, int *cnt, unsigned char *selected, int cap) {
    if (!h) return DLT_E_INVALID;
    if (!h->have_match) DLT_FAIL(h, DLT_E_STATE, "no match pass yet");
    int rc = run_far(h, nullptr);  // make every neighbour set exact before handing it out
    if (rc) return rc;
    int m = h->n_down < cap ? h->n_down : cap;
    if (m > 0) {
        if (nbr) DLT_RT(h, rt::d2h(nbr, h->knn.nbr, (size_t)m * kK * sizeof(float4), h->stream));
        if (cnt) DLT_RT(h, rt::d2h(cnt, h->knn.nbr_cnt, (size_t)m * sizeof(int), h->stream));
        if (selected) DLT_RT(h, rt::d2h(selected, h->d_sel, (size_t)m, h->stream));
        DLT_RT(h, rt::sync(h->stream));
    }
    return DLT_OK;
}

int dlt_degeneracy_begin(dlt_handle h) {
    if (!h) return DLT_E_INVALID;
    if (!h->have_match && !h->eig_valid) DLT_FAIL(h, DLT_E_STATE, "no measurement yet");
    rt::set_device(h->cfg.device);
    if (!h->eig_valid) {
        cudaStream_t s = h->stream;
        if (h->have_aux) {  // fork: the decomposition overlaps whatever the caller queues next (map_incremental)
            DLT_RT(h, rt::event_record(h->ev_fork, h->stream));
            DLT_RT(h, rt::stream_wait(h->aux_stream, h->ev_fork));
            s = h->aux_stream;
        }
        // zero-copy (as for the result block of dlt_measure): the kernel stores the 42 doubles straight into pinned host memory and
        // then this launch's number into a host flag dlt_degeneracy spins on -- no device->host copy, no stream synchronisation
        const bool zc = h->zerocopy && !h->prof_on;
        if (zc) ++h->eig_seq;
        {
            ProfScope profe(h, 8, s);
            DLT_LAUNCH(k_eigen6, 1, 32, s, (const double *)h->d_result, h->d_result + kEigOffset, zc ? h->h_result + kEigOffset : (double *)nullptr,
                       zc ? reinterpret_cast<unsigned long long *>(h->h_ints + 34) : (unsigned long long *)nullptr, zc ? h->eig_seq : 0ull);
        }
        DLT_RT(h, rt::check_launch());
        if (!zc) DLT_RT(h, rt::d2h(h->h_result + kEigOffset, h->d_result + kEigOffset, 42 * sizeof(double), s));
        h->eig_zc = zc;
        h->eig_stream = s;
        if (h->have_aux) {
            DLT_RT(h, rt::event_record(h->ev_join, h->aux_stream));
            h->eig_pending = true;
        }
        h->eig_valid = true;
    }
    return DLT_OK;
}

int dlt_degeneracy(dlt_handle h, double *eigvals6, double *eigvecs36) {
    if (!h || !eigvals6 || !eigvecs36) return DLT_E_INVALID;
    int rc = dlt_degeneracy_begin(h);
    if (rc) return rc;
    if (h->eig_zc) {
        const volatile unsigned long long *flag = reinterpret_cast<const volatile unsigned long long *>(h->h_ints + 34);
        for (unsigned spins = 1; *flag != h->eig_seq; spins++) {
            if (rt::t_wait_hook) rt::t_wait_hook(rt::t_wait_ctx);
#if defined(__x86_64__)
            __builtin_ia32_pause();
#endif
            if ((spins & 0xFFFu) == 0u) {  // a kernel that died never raises the flag
                const int q = rt::stream_query(h->eig_stream);
                if (q == 2) DLT_FAIL(h, DLT_E_CUDA, std::string("stream error while waiting for the eigen block: ") + rt::last_error());
                if (q == 0 && *flag != h->eig_seq) DLT_FAIL(h, DLT_E_CUDA, "the stream drained without the eigen flag");
            }
        }
        std::atomic_thread_fence(std::memory_order_acquire);
        // (ev_join stays pending: the main stream still orders its next write of d_result behind the kernel, eig_join)
    } else if (h->eig_pending) {
        DLT_RT(h, rt::sync(h->aux_stream));
        h->eig_pending = false;
    } else {
        DLT_RT(h, rt::sync(h->stream));
    }
    for (int i = 0; i < 6; i++) eigvals6[i] = h->h_result[kEigOffset + i];
    for (int i = 0; i < 36; i++) eigvecs36[i] = h->h_result[kEigOffset + 6 + i];
    return DLT_OK;
}

// ------------------------------------------------------------------ map_incremental
namespace {
struct StreamSwap {  // enqueue on another stream for the lifetime of the scope
    dlt_handle h;
    cudaStream_t saved;
    bool on;
    StreamSwap(dlt_handle h_, cudaStream_t s, bool on_) : h(h_), saved(h_->stream), on(on_) {
        if (on) h->stream = s;
    }
    ~StreamSwap() {
        if (on) h->stream = saved;
    }
};
}  // namespace

static int map_incremental_impl(dlt_handle h, const double *pose24, int flg_EKF_inited, int *n_ds, int *n_raw, bool want_async) {
    if (!h || !pose24) return DLT_E_INVALID;
    if (!h->have_down) DLT_FAIL(h, DLT_E_STATE, "dlt_map_incremental before a scan");
    const bool sharded = h->map.shard_count > 1;
    if (sharded && !h->shard_reduce && !h->peer_on)
        DLT_FAIL(h, DLT_E_STATE, "dlt_map_incremental on a sharded map needs dlt_peer_attach or dlt_set_shard_reduce");
    if (int rd = map_refuse_if_dead(h)) return rd;
    rt::set_device(h->cfg.device);
    if (int ri = ins_finish(h)) return ri;
    if (int rn = resolve_n_down(h)) return rn;
    const int n = h->n_down;
    h->last_ins_ds = h->last_ins_raw = 0;
    if (n_ds) *n_ds = 0;
    if (n_raw) *n_raw = 0;
    if (n == 0) return DLT_OK;
    if (h->have_match && !(h->nfar_known && h->h_last_nfar == 0)) {
        // unresolved queries: map_incremental only needs the nearest point of those that saw nothing at all (k_nn1);
        // a ring-search overflow (pathological chains) takes the full exact fallback
        DLT_RT(h, rt::d2h(h->h_ints, h->d_counters, 16 * sizeof(int), h->stream));
        DLT_RT(h, rt::sync(h->stream));
        if (h->h_ints[10] > 0) {
            h->nfar_known = false;
            int rc = run_far(h, nullptr);
            if (rc) return rc;
        } else if (h->h_ints[9] > 0) {
            ProfScope prof(h, 5);
            DLT_LAUNCH(k_nn1_seed, h->n_sm, kSeedBlock, h->stream, h->map, (const int *)h->map.n_buckets, (const float4 *)h->knn.qw,
                       (const int *)h->knn.nn_list, (const int *)h->knn.nn_count, h->knn.nn_key, (int *)nullptr);
            DLT_LAUNCH(k_nn1, dim3(kNn1GroupsX, 8 * h->n_sm), kNn1Block, h->stream, h->map, (const int *)h->map.n_buckets, (const float4 *)h->knn.qw,
                       (const int *)h->knn.nn_list, (const int *)h->knn.nn_count, h->knn.nn_key, (int *)nullptr);
        }
    }
    Pose P = pose_from(pose24);
    // Asynchronous form (unsharded map, every query resolved, so no host decision is pending): the insert kernels go to their own
    // stream behind what the main stream holds now; nothing here waits for them.
    // (a sharded map with attached peers exchanges the owners' decisions through the mailboxes from inside the kernels, so its
    // insert can go the same way; with a reduce CALLBACK the library collective stays on the caller's stream: synchronous)
    const bool async = want_async && (!sharded || (h->have_match && h->peer_on)) && h->have_ins && !h->prof_on &&
                       !(h->have_match && !(h->nfar_known && h->h_last_nfar == 0));
    if (async) {
        DLT_RT(h, rt::event_record(h->ev_loop_done, h->stream));
        DLT_RT(h, rt::stream_wait(h->ins_stream, h->ev_loop_done));
    }
    StreamSwap swap(h, h->ins_stream, async);
    DLT_RT(h, rt::fill(h->d_counters + 6, 0, 2 * sizeof(int), h->stream));
    FuseInsert fi = {0, h->scratch, h->d_cellslot, h->d_vslot};
    const bool peer_pull = sharded && h->have_match && h->peer_on;
    if (!sharded || peer_pull) {  // unsharded: the classification kernel also claims cells and places the voxel bids; with peer
        if (int rs = prepare_scratch(h, n, false, &fi.sc)) return rs;  // mailboxes k_incr_pull does, once it knows every decision
        fi.on = 1;
    }
    const FuseInsert fi_pull = fi;
    if (peer_pull) fi.on = 0;
    DLT_LAUNCH(k_incr_classify, div_up(n, 256), 256, h->stream, (const float4 *)h->d_down, n, P, (const float4 *)h->knn.nbr,
               (const int *)h->knn.nbr_cnt, (double)h->cfg.ds_map, flg_EKF_inited ? 1 : 0, h->d_pw, h->d_dsflag, h->d_addflag, h->d_counters + 6,
               LoopArgs{nullptr, nullptr, nullptr, 0, 0, 0ull, 0ull, nullptr}, h->map, (const int *)h->map.n_live, h->have_match ? (const unsigned char *)h->knn.flags : (const unsigned char *)nullptr,
               (const int *)h->knn.nn_pos, (const unsigned long long *)h->knn.nn_key, fi);
    if (async) DLT_RT(h, rt::event_record(h->ev_classified, h->stream));  // the downsampled scan has been read
    if (sharded && h->have_match && h->peer_on) {  // owners store their decisions straight into every rank's mailbox
        DLT_LAUNCH(k_incr_push, div_up(n, 256), 256, h->stream, h->d_peer, (const unsigned char *)h->d_dsflag, (const unsigned char *)h->d_addflag,
                   (const unsigned char *)h->knn.flags, n, (unsigned char)kFlagForeign, (const int *)nullptr, (const int *)nullptr);
        DLT_RT(h, rt::fill(h->d_counters + 6, 0, 2 * sizeof(int), h->stream));
        DLT_LAUNCH(k_incr_pull, div_up(n, 256), 256, h->stream, h->d_peer, n, h->d_dsflag, h->d_addflag, h->d_counters + 6, (const int *)nullptr,
                   (const int *)nullptr, h->map, (const float4 *)h->d_pw, fi_pull);
    } else if (sharded && h->have_match) {  // owners decide, everybody learns every decision, every rank inserts into its tiles + halo
                                     // (without a match pass every rank already agrees: all points are PointToAdd)
        DLT_LAUNCH(k_incr_pack, div_up(n, 256), 256, h->stream, (const unsigned char *)h->d_dsflag, (const unsigned char *)h->d_addflag, n, h->d_flagbuf);
        DLT_RT(h, rt::check_launch());
        if (h->shard_reduce(h->shard_reduce_ctx, h->d_flagbuf, n) != 0) DLT_FAIL(h, DLT_E_STATE, "shard reduce callback failed");
        DLT_RT(h, rt::fill(h->d_counters + 6, 0, 2 * sizeof(int), h->stream));
        DLT_LAUNCH(k_incr_unpack, div_up(n, 256), 256, h->stream, (const double *)h->d_flagbuf, n, h->d_dsflag, h->d_addflag, h->d_counters + 6);
    }
    int rc = insert_points(h, h->d_pw, n, true, InsertGate{nullptr, nullptr}, fi_pull.on ? &fi_pull.sc : nullptr);
    if (rc) return rc;
    h->have_match = false;  // the map changed: neighbour sets are stale
    if (async) {
        DLT_RT(h, rt::d2h(h->h_ins_ints, h->d_counters, 16 * sizeof(int), h->stream));
        if (h->peer_on) DLT_RT(h, rt::d2h(h->h_peer_status, &h->d_peer->status, sizeof(int), h->stream));
        DLT_RT(h, rt::event_record(h->ev_inserted, h->stream));
        h->cls_pending = h->ins_pending = h->ins_uncollected = true;
        h->counters_fresh = false;
        return DLT_OK;  // counts: dlt_map_incremental_collect
    }
    if (h->peer_on) DLT_RT(h, rt::d2h(h->h_peer_status, &h->d_peer->status, sizeof(int), h->stream));
    rc = map_check_error(h);  // reads the 8 counters back
    if (h->peer_on && *h->h_peer_status != 0) DLT_FAIL(h, DLT_E_STATE, "peer exchange timed out (a rank did not take part in map_incremental)");
    h->last_ins_ds = h->h_ints[6];
    h->last_ins_raw = h->h_ints[7];
    if (n_ds) *n_ds = h->h_ints[6];
    if (n_raw) *n_raw = h->h_ints[7];
    return rc;
}

int dlt_map_incremental(dlt_handle h, const double *pose24, int flg_EKF_inited, int *n_ds, int *n_raw) {
    return map_incremental_impl(h, pose24, flg_EKF_inited, n_ds, n_raw, false);
}
int dlt_map_incremental_async(dlt_handle h, const double *pose24, int flg_EKF_inited) {
    return map_incremental_impl(h, pose24, flg_EKF_inited, nullptr, nullptr, true);
}
int dlt_map_incremental_collect(dlt_handle h, int *n_ds, int *n_raw) {
    if (!h) return DLT_E_INVALID;
    rt::set_device(h->cfg.device);
    const int rc = ins_finish(h);
    if (n_ds) *n_ds = h->last_ins_ds;
    if (n_raw) *n_raw = h->last_ins_raw;
    return rc;
}

double *dlt_result_dev(dlt_handle h) { return h ? h->d_result : nullptr; }

int dlt_set_thread_wait_hook(void (*fn)(void *), void *ctx) {
    rt::t_wait_hook = fn;
    rt::t_wait_ctx = ctx;
    return DLT_OK;
}

int dlt_debug_counters(dlt_handle h, int *out16) {
    if (!h || !out16) return DLT_E_INVALID;
    rt::set_device(h->cfg.device);
    if (int ri = ins_finish(h)) return ri;
    DLT_RT(h, rt::d2h(h->h_ints + 40, h->d_counters, 16 * sizeof(int), h->stream));
    DLT_RT(h, rt::sync(h->stream));
    std::memcpy(out16, h->h_ints + 40, 16 * sizeof(int));
    return DLT_OK;
}

int dlt_set_shard_reduce(dlt_handle h, dlt_reduce_fn reduce, void *ctx) {
    if (!h) return DLT_E_INVALID;
    h->shard_reduce = reduce;
    h->shard_reduce_ctx = ctx;
    return DLT_OK;
}

// ------------------------------------------------------------------ peer mailboxes (sharded map over NVLink peer memory)
int dlt_peer_export(dlt_handle h, unsigned char *blob) {
    if (!h || !blob) return DLT_E_INVALID;
    if (h->cfg.shard_count < 2 || h->cfg.shard_count > DLT_MAX_PEERS) DLT_FAIL(h, DLT_E_STATE, "dlt_peer_export needs 2 <= shard_count <= DLT_MAX_PEERS");
    rt::set_device(h->cfg.device);
    if (int ri = ins_finish(h)) return ri;
    if (!h->peer_box) {
        const size_t dec_cap = ((size_t)h->cfg.max_scan_points + 127) & ~(size_t)127;
        h->peer_bytes = sizeof(PeerBox) + 2 * dec_cap;
        void *p = nullptr;
        if (rt::shared_alloc(&p, h->peer_bytes, &h->peer_blob) != 0) DLT_FAIL(h, DLT_E_CUDA, "peer mailbox allocation failed");
        h->peer_box = (PeerBox *)p;  // zeroed: sequence numbers start at 1, so early posts of a fast peer are never mistaken
    }
    static_assert(sizeof(rt::ShareBlob) == DLT_PEER_BLOB_BYTES, "blob size");
    std::memcpy(blob, &h->peer_blob, sizeof(h->peer_blob));
    return DLT_OK;
}

int dlt_peer_detach(dlt_handle h) {
    if (!h) return DLT_E_INVALID;
    if (!h->peer_on) return DLT_OK;
    rt::set_device(h->cfg.device);
    if (int ri = ins_finish(h)) return ri;
    rt::sync(h->stream);
    for (int r = 0; r < DLT_MAX_PEERS; r++) {
        if (h->peer_maps[r]) rt::shared_close(h->peer_maps[r], h->peer_map_bytes[r]);
        h->peer_maps[r] = nullptr;
    }
    h->peer_on = false;
    h->peer_detached = true;  // the mailbox keeps the old sequence numbers: no second attach on this handle
    return DLT_OK;
}

int dlt_peer_attach(dlt_handle h, const unsigned char *blobs) {
    if (!h || !blobs) return DLT_E_INVALID;
    if (!h->peer_box) DLT_FAIL(h, DLT_E_STATE, "dlt_peer_attach before dlt_peer_export");
    if (h->peer_on || h->peer_detached) DLT_FAIL(h, DLT_E_STATE, "peers are already attached or were detached (attach once per handle)");
    rt::set_device(h->cfg.device);
    if (int ri = ins_finish(h)) return ri;
    h->err.clear();
    const int W = h->cfg.shard_count, me = h->cfg.shard_rank;
    PeerComm pc;
    std::memset(&pc, 0, sizeof(pc));
    pc.world = W;
    pc.rank = me;
    pc.dec_cap = (int)((h->peer_bytes - sizeof(PeerBox)) / 2);
    bool ok = true;
    for (int r = 0; r < W && ok; r++) {
        rt::ShareBlob b;
        std::memcpy(&b, blobs + (size_t)r * DLT_PEER_BLOB_BYTES, sizeof(b));
        if (r == me) {
            if (b.ptr != h->peer_blob.ptr || b.pid != h->peer_blob.pid) {
                h->err = "blob of this rank is not the one dlt_peer_export produced (blobs must be in rank order)";
                ok = false;
            }
            pc.box[r] = h->peer_box;
            continue;
        }
        if (b.magic != rt::kShareMagic) {
            h->err = "cannot map a peer's mailbox: not a dlt_peer_export blob";
            ok = false;
            break;
        }
        if (b.bytes != h->peer_bytes) {
            h->err = "peer mailbox size differs (max_scan_points must be the same on every rank)";
            ok = false;
            break;
        }
        void *p = nullptr;
        bool mapped = false;
        if (rt::shared_open(&b, h->cfg.device, &p, &mapped) != 0) {
            h->err = "cannot map a peer's mailbox (CUDA IPC / peer access unavailable between these devices)";
            ok = false;
            break;
        }
        if (mapped) {
            h->peer_maps[r] = p;
            h->peer_map_bytes[r] = (size_t)b.bytes;
        }
        pc.box[r] = (PeerBox *)p;
    }
    if (ok && !h->d_peer) ok = rt::alloc((void **)&h->d_peer, sizeof(PeerComm)) == 0;
    if (ok && !h->h_peer_status) ok = rt::pinned_alloc((void **)&h->h_peer_status, 64) == 0;
    if (ok) {
        *h->h_peer_status = 0;
        ok = rt::h2d(h->d_peer, &pc, sizeof(pc), h->stream) == 0 && rt::sync(h->stream) == 0;
    }
    if (!ok) {
        for (int r = 0; r < DLT_MAX_PEERS; r++) {
            if (h->peer_maps[r]) rt::shared_close(h->peer_maps[r], h->peer_map_bytes[r]);
            h->peer_maps[r] = nullptr;
        }
        if (h->err.empty()) h->err = "dlt_peer_attach failed";
        return DLT_E_CUDA;
    }
    h->peer_on = true;
    return DLT_OK;
}

// ------------------------------------------------------------------ instrumentation
int dlt_set_profiling(dlt_handle h, int on) {
    if (!h) return DLT_E_INVALID;
    h->prof_on = on != 0;
    return DLT_OK;
}
int dlt_get_profile(dlt_handle h, double *ms8, long long *count8, int reset) {
    if (!h || !ms8 || !count8) return DLT_E_INVALID;
    DLT_RT(h, rt::sync(h->stream));
    for (ProfSpan &sp : h->spans) {
        h->prof_ms[sp.kind] += (double)rt::event_elapsed_ms(sp.a, sp.b);
        h->prof_n[sp.kind] += 1;
        rt::event_destroy(sp.a);
        rt::event_destroy(sp.b);
    }
    h->spans.clear();
    for (int k = 0; k < kProfKinds; k++) {
        if (k < 8) {
            ms8[k] = h->prof_ms[k];
            count8[k] = h->prof_n[k];
        }
        if (reset) {
            h->prof_ms[k] = 0;
            h->prof_n[k] = 0;
        }
    }
    return DLT_OK;
}
unsigned long long dlt_launch_count(void) { return rt::g_launches.load(); }

int dlt_get_iekf_clocks(dlt_handle h, long long *clocks, int n_iter) {
    if (!h || !clocks || n_iter < 0 || n_iter > DLT_IEKF_MAX_ITER) return DLT_E_INVALID;
    DLT_RT(h, rt::d2h(clocks, h->d_iekf->clocks, (size_t)n_iter * 16 * sizeof(long long), h->stream));
    DLT_RT(h, rt::sync(h->stream));
    return DLT_OK;
}

int dlt_get_timeline(dlt_handle h, double *kind_start_end, int cap, int *n) {
    if (!h || !n || cap < 0 || (cap > 0 && !kind_start_end)) return DLT_E_INVALID;
    DLT_RT(h, rt::sync(h->stream));
    if (h->have_aux) DLT_RT(h, rt::sync(h->aux_stream));
    *n = (int)h->spans.size();
    for (int i = 0; i < *n && i < cap; i++) {
        kind_start_end[3 * i] = (double)h->spans[i].kind;
        kind_start_end[3 * i + 1] = (double)rt::event_elapsed_ms(h->spans[0].a, h->spans[i].a);
        kind_start_end[3 * i + 2] = (double)rt::event_elapsed_ms(h->spans[0].a, h->spans[i].b);
    }
    return DLT_OK;
}

}  // extern "C"
