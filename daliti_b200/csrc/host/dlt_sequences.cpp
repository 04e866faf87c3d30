// daliti_b200/csrc/host/dlt_sequences.cpp
//
// dlt_lio_replay_sequences: the per-scan update (eskf_lio/src/laserMapping.cpp:731-1177) of MANY independent LiDAR
// sequences on one GPU, driven from native threads (BASELINE config C5; no reference counterpart -- the reference runs one
// sequence per process).
//
// A single sequence's update is a chain of small dependent kernels with a host decision per IEKF iteration, so one sequence
// cannot fill a B200; many can.  What limits them is the host: a thread that drives one sequence spends most of a scan
// waiting for the device.  Here every worker thread owns several sequences and runs each one's replay
// (on_lidar_msg -> [prefetch scan k+1] -> process_scan k) as a coroutine; wherever the library would wait for the device
// (dlt_set_thread_wait_hook: the result block of an iteration, a stream or event wait) the coroutine hands the thread to the
// next sequence, which enqueues ITS kernels meanwhile.  The sequences are not in lockstep: each advances as fast as its own
// dependency chain allows, on its own CUDA streams.
#include <ucontext.h>

#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "../../../include/daliti_b200_lio.h"

namespace {

constexpr size_t kStackBytes = 512 * 1024;

struct Worker;
struct Coro {
    ucontext_t ctx;
    char *stack = nullptr;
    dlt_lio_seq *seq = nullptr;
    Worker *w = nullptr;
    bool done = false;
};
struct Worker {
    ucontext_t sched;
    std::vector<Coro *> coros;
    int current = -1;
};

void wait_hook(void *ctx) {  // called by the library on this thread while sequence `current` waits for the device
    Worker *w = static_cast<Worker *>(ctx);
    if (w->current < 0 || w->coros.size() < 2) return;  // nobody to run meanwhile: plain spin
    Coro *c = w->coros[(size_t)w->current];
    swapcontext(&c->ctx, &w->sched);
}

int replay_one(dlt_lio_seq *s) {
    s->n_done = 0;
    for (int k = 0; k < s->n_scans; k++) {
        const dlt_lio_seq_scan &sc = s->scans[k];
        for (int m = 0; m < sc.lidar_msgs; m++) dlt_lio_on_lidar_msg(s->h);
        if (s->prefetch && !sc.pts_on_device && k + 1 < s->n_scans && !s->scans[k + 1].pts_on_device) {
            int rp = dlt_lio_prefetch_scan(s->h, s->scans[k + 1].pts48, s->scans[k + 1].n);
            if (rp != 0) return rp;
        }
        dlt_lio_scan_out local;
        dlt_lio_scan_out *out = s->outs ? &s->outs[k] : &local;
        int rc;
        if (sc.pts_on_device)
            rc = dlt_lio_process_scan_dev(s->h, sc.pts48, sc.n, sc.lidar_beg_time, sc.observation_end_time, sc.imu7, sc.n_imu, s->thermal, out);
        else
            rc = dlt_lio_process_scan(s->h, sc.pts48, sc.n, sc.lidar_beg_time, sc.imu7, sc.n_imu, s->thermal, out);
        if (rc != 0) return rc;
        s->n_done = k + 1;
    }
    return dlt_lio_collect_insert(s->h, nullptr, nullptr);  // the last scan's map_incremental (async_insert)
}

void coro_entry(unsigned lo, unsigned hi) {
    Coro *c = reinterpret_cast<Coro *>(((unsigned long long)hi << 32) | (unsigned long long)lo);
    c->seq->rc = replay_one(c->seq);
    c->done = true;
    swapcontext(&c->ctx, &c->w->sched);  // never resumed
}

void run_worker(dlt_lio_seq *seqs, int n_seqs, int first, int stride) {
    Worker w;
    for (int i = first; i < n_seqs; i += stride) {
        Coro *c = new Coro();
        c->seq = &seqs[i];
        c->w = &w;
        c->stack = static_cast<char *>(std::malloc(kStackBytes));
        if (!c->stack) {
            seqs[i].rc = DLT_E_CUDA;
            delete c;
            continue;
        }
        getcontext(&c->ctx);
        c->ctx.uc_stack.ss_sp = c->stack;
        c->ctx.uc_stack.ss_size = kStackBytes;
        c->ctx.uc_link = nullptr;
        const unsigned long long p = reinterpret_cast<unsigned long long>(c);
        makecontext(&c->ctx, reinterpret_cast<void (*)()>(coro_entry), 2, (unsigned)(p & 0xFFFFFFFFull), (unsigned)(p >> 32));
        w.coros.push_back(c);
    }
    dlt_set_thread_wait_hook(wait_hook, &w);
    size_t remaining = w.coros.size();
    while (remaining > 0) {
        for (size_t i = 0; i < w.coros.size(); i++) {
            Coro *c = w.coros[i];
            if (c->done) continue;
            w.current = (int)i;
            swapcontext(&w.sched, &c->ctx);
            if (c->done) remaining--;
        }
    }
    w.current = -1;
    dlt_set_thread_wait_hook(nullptr, nullptr);
    for (Coro *c : w.coros) {
        std::free(c->stack);
        delete c;
    }
}

}  // namespace

extern "C" int dlt_lio_replay_sequences(dlt_lio_seq *seqs, int n_seqs, int n_threads) {
    if (n_seqs < 0 || (n_seqs > 0 && !seqs)) return DLT_E_INVALID;
    for (int i = 0; i < n_seqs; i++) {
        if (!seqs[i].h || seqs[i].n_scans < 0 || (seqs[i].n_scans > 0 && !seqs[i].scans)) return DLT_E_INVALID;
        seqs[i].rc = 0;
        seqs[i].n_done = 0;
    }
    if (n_seqs == 0) return DLT_OK;
    int T = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    if (T < 1) T = 1;
    if (T > n_seqs) T = n_seqs;
    std::vector<std::thread> pool;
    for (int t = 1; t < T; t++) pool.emplace_back(run_worker, seqs, n_seqs, t, T);
    run_worker(seqs, n_seqs, 0, T);  // the calling thread is worker 0
    for (std::thread &th : pool) th.join();
    for (int i = 0; i < n_seqs; i++)
        if (seqs[i].rc != 0) return seqs[i].rc;
    return DLT_OK;
}
