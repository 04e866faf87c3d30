// daliti_b200/csrc/dlt_rt.h -- thin runtime layer: device memory, copies, launches.
// The product build is CUDA (nvcc, sm_100a).  With -DDLT_EMU (tests/emu only) the same
// sources compile as plain C++ against the test emulator so kernel logic can be exercised
// without a GPU; that build is test infrastructure and is never loaded by the package.
#pragma once
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <unistd.h>
#if defined(DLT_EMU)
#include <fcntl.h>
#include <sys/mman.h>
#endif

#include "dlt_common.cuh"

#if defined(DLT_EMU)

namespace dlt {
namespace rt {
extern std::atomic<unsigned long long> g_launches;
// dlt_set_thread_wait_hook: called instead of blocking / spinning while this thread waits for the device
extern thread_local void (*t_wait_hook)(void *);
extern thread_local void *t_wait_ctx;
struct Event {};
inline int event_create(Event *) { return 0; }
inline void event_destroy(Event) {}
inline int event_record(Event, cudaStream_t) { return 0; }
inline float event_elapsed_ms(Event, Event) { return 0.f; }
inline int event_create_untimed(Event *) { return 0; }
inline int stream_wait(cudaStream_t, Event) { return 0; }
inline int event_sync(Event) {
    if (t_wait_hook) t_wait_hook(t_wait_ctx);  // (exercises the multi-sequence driver's switching under the emulator)
    return 0;
}
inline int alloc(void **p, size_t bytes) {
    size_t n = (bytes + 255) & ~(size_t)255;
    if (n == 0) n = 256;
    *p = aligned_alloc(256, n);
    return *p ? 0 : 1;
}
inline void release(void *p) { free(p); }
inline int pinned_alloc(void **p, size_t bytes) { return alloc(p, bytes); }
inline void pinned_release(void *p) { free(p); }
inline int h2d(void *d, const void *s, size_t n, cudaStream_t) {
    std::memcpy(d, s, n);
    return 0;
}
inline int d2h(void *d, const void *s, size_t n, cudaStream_t) {
    std::memcpy(d, s, n);
    return 0;
}
inline int d2d(void *d, const void *s, size_t n, cudaStream_t) {
    std::memmove(d, s, n);
    return 0;
}
inline int fill(void *d, int byte, size_t n, cudaStream_t) {
    std::memset(d, byte, n);
    return 0;
}
inline int sync(cudaStream_t) {
    if (t_wait_hook) t_wait_hook(t_wait_ctx);
    return 0;
}
inline int stream_query(cudaStream_t) { return 0; }  // 0 idle, 1 still busy, 2 error
inline int stream_create(cudaStream_t *s) {
    *s = nullptr;
    return 0;
}
inline void stream_destroy(cudaStream_t) {}
inline int set_device(int) { return 0; }
inline int device_count() { return 1; }
inline int sm_count() { return 4; }
inline const char *last_error() { return "emu"; }
inline int check_launch() { return 0; }

// ---- memory shared with peer processes (dlt_peer_*): POSIX shared memory stands in for CUDA IPC
struct ShareBlob {  // what dlt_peer_export hands out (DLT_PEER_BLOB_BYTES)
    unsigned long long magic;
    long long pid;
    unsigned long long ptr;
    unsigned long long bytes;
    int device;
    int pad;
    char ipc[64];  // cudaIpcMemHandle_t | shm object name
    char pad2[24];
};
static_assert(sizeof(ShareBlob) == 128, "blob size is part of the ABI");
constexpr unsigned long long kShareMagic = 0x444C545045455231ull;  // "DLTPEER1"
inline int shared_alloc(void **p, size_t bytes, ShareBlob *blob) {
    static std::atomic<int> counter{0};
    std::memset(blob, 0, sizeof(*blob));
    std::snprintf(blob->ipc, sizeof(blob->ipc), "/dlt_peer_%ld_%d", (long)getpid(), counter.fetch_add(1));
    int fd = shm_open(blob->ipc, O_CREAT | O_EXCL | O_RDWR, 0600);
    if (fd < 0) return 1;
    if (ftruncate(fd, (off_t)bytes) != 0) {
        close(fd);
        shm_unlink(blob->ipc);
        return 1;
    }
    void *m = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (m == MAP_FAILED) {
        shm_unlink(blob->ipc);
        return 1;
    }
    std::memset(m, 0, bytes);
    *p = m;
    blob->magic = kShareMagic;
    blob->pid = (long long)getpid();
    blob->ptr = (unsigned long long)(uintptr_t)m;
    blob->bytes = bytes;
    return 0;
}
inline void shared_release(void *p, size_t bytes, const ShareBlob *blob) {
    if (p) munmap(p, bytes);
    if (blob->magic == kShareMagic) shm_unlink(blob->ipc);
}
// *mapped: the pointer must be handed back to shared_close
inline int shared_open(const ShareBlob *blob, int /*my_device*/, void **p, bool *mapped) {
    *mapped = false;
    if (blob->magic != kShareMagic) return 1;
    if (blob->pid == (long long)getpid()) {
        *p = (void *)(uintptr_t)blob->ptr;
        return 0;
    }
    int fd = shm_open(blob->ipc, O_RDWR, 0600);
    if (fd < 0) return 1;
    void *m = mmap(nullptr, (size_t)blob->bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (m == MAP_FAILED) return 1;
    *p = m;
    *mapped = true;
    return 0;
}
inline void shared_close(void *p, size_t bytes) {
    if (p) munmap(p, bytes);
}
}  // namespace rt
}  // namespace dlt

#else  // ------------------------------------------------------------------ CUDA

namespace dlt { namespace rt {
extern std::atomic<unsigned long long> g_launches;
extern int g_pdl;  // 1: launch with the programmatic-stream-serialization attribute (DLT_PDL=0 switches it off)
} }
// Every kernel starts with DLT_PDL_WAIT() (dlt_common.cuh), so it may be launched as a programmatic dependent of the
// kernel in front of it: the launch latency then overlaps that kernel's tail instead of following its completion.
#define DLT_LAUNCH(kernel, grid, block, stream_, ...)                                                   \
    do {                                                                                                \
        cudaLaunchConfig_t cfg__ = {};                                                                  \
        cfg__.gridDim = dim3(grid);                                                                     \
        cfg__.blockDim = dim3(block);                                                                   \
        cfg__.dynamicSmemBytes = 0;                                                                     \
        cfg__.stream = (stream_);                                                                       \
        cudaLaunchAttribute attr__[1];                                                                  \
        attr__[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                              \
        attr__[0].val.programmaticStreamSerializationAllowed = 1;                                       \
        cfg__.attrs = attr__;                                                                           \
        cfg__.numAttrs = dlt::rt::g_pdl ? 1 : 0;                                                        \
        cudaLaunchKernelEx(&cfg__, kernel, __VA_ARGS__);                                                \
        ++dlt::rt::g_launches;                                                                          \
    } while (0)

namespace dlt {
namespace rt {
// dlt_set_thread_wait_hook: called instead of blocking / spinning while this thread waits for the device (the multi-sequence
// driver switches to another sequence's update there)
extern thread_local void (*t_wait_hook)(void *);
extern thread_local void *t_wait_ctx;
typedef cudaEvent_t Event;
inline int event_create(Event *e) { return cudaEventCreate(e) == cudaSuccess ? 0 : 1; }
inline void event_destroy(Event e) { cudaEventDestroy(e); }
inline int event_record(Event e, cudaStream_t s) { return cudaEventRecord(e, s) == cudaSuccess ? 0 : 1; }
inline float event_elapsed_ms(Event a, Event b) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    return ms;
}
inline int event_create_untimed(Event *e) { return cudaEventCreateWithFlags(e, cudaEventDisableTiming) == cudaSuccess ? 0 : 1; }
inline int stream_wait(cudaStream_t s, Event e) { return cudaStreamWaitEvent(s, e, 0) == cudaSuccess ? 0 : 1; }
inline int event_sync(Event e) {
    if (t_wait_hook) {
        cudaError_t r;
        while ((r = cudaEventQuery(e)) == cudaErrorNotReady) t_wait_hook(t_wait_ctx);
        return r == cudaSuccess ? 0 : 1;
    }
    return cudaEventSynchronize(e) == cudaSuccess ? 0 : 1;
}
inline int alloc(void **p, size_t bytes) { return cudaMalloc(p, bytes ? bytes : 256) == cudaSuccess ? 0 : 1; }
inline void release(void *p) {
    if (p) cudaFree(p);
}
inline int pinned_alloc(void **p, size_t bytes) { return cudaMallocHost(p, bytes ? bytes : 256) == cudaSuccess ? 0 : 1; }
inline void pinned_release(void *p) {
    if (p) cudaFreeHost(p);
}
inline int h2d(void *d, const void *s, size_t n, cudaStream_t st) {
    return cudaMemcpyAsync(d, s, n, cudaMemcpyHostToDevice, st) == cudaSuccess ? 0 : 1;
}
inline int d2h(void *d, const void *s, size_t n, cudaStream_t st) {
    return cudaMemcpyAsync(d, s, n, cudaMemcpyDeviceToHost, st) == cudaSuccess ? 0 : 1;
}
inline int d2d(void *d, const void *s, size_t n, cudaStream_t st) {
    return cudaMemcpyAsync(d, s, n, cudaMemcpyDeviceToDevice, st) == cudaSuccess ? 0 : 1;
}
inline int fill(void *d, int byte, size_t n, cudaStream_t st) { return cudaMemsetAsync(d, byte, n, st) == cudaSuccess ? 0 : 1; }
inline int sync(cudaStream_t st) {
    if (t_wait_hook) {
        cudaError_t r;
        while ((r = cudaStreamQuery(st)) == cudaErrorNotReady) t_wait_hook(t_wait_ctx);
        return r == cudaSuccess ? 0 : 1;
    }
    return cudaStreamSynchronize(st) == cudaSuccess ? 0 : 1;
}
inline int stream_query(cudaStream_t st) {  // 0 idle, 1 still busy, 2 error
    cudaError_t e = cudaStreamQuery(st);
    return e == cudaSuccess ? 0 : (e == cudaErrorNotReady ? 1 : 2);
}
inline int stream_create(cudaStream_t *s) { return cudaStreamCreateWithFlags(s, cudaStreamNonBlocking) == cudaSuccess ? 0 : 1; }
inline void stream_destroy(cudaStream_t s) {
    if (s) cudaStreamDestroy(s);
}
inline int set_device(int d) { return cudaSetDevice(d) == cudaSuccess ? 0 : 1; }
inline int device_count() {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}
inline int sm_count() {
    int dev = 0, n = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n > 0 ? n : 148;
}
inline const char *last_error() { return cudaGetErrorString(cudaPeekAtLastError()); }
inline int check_launch() { return cudaGetLastError() == cudaSuccess ? 0 : 1; }

// ---- memory shared with peer processes (dlt_peer_*): CUDA IPC across processes, peer access inside one
struct ShareBlob {  // what dlt_peer_export hands out (DLT_PEER_BLOB_BYTES)
    unsigned long long magic;
    long long pid;
    unsigned long long ptr;
    unsigned long long bytes;
    int device;
    int pad;
    char ipc[64];  // cudaIpcMemHandle_t
    char pad2[24];
};
static_assert(sizeof(ShareBlob) == 128, "blob size is part of the ABI");
static_assert(sizeof(cudaIpcMemHandle_t) <= 64, "IPC handle fits the blob");
constexpr unsigned long long kShareMagic = 0x444C545045455231ull;  // "DLTPEER1"
inline int shared_alloc(void **p, size_t bytes, ShareBlob *blob) {
    std::memset(blob, 0, sizeof(*blob));
    if (cudaMalloc(p, bytes) != cudaSuccess) return 1;
    if (cudaMemset(*p, 0, bytes) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) return 1;
    cudaIpcMemHandle_t hd;
    if (cudaIpcGetMemHandle(&hd, *p) != cudaSuccess) {
        cudaGetLastError();
        std::memset(&hd, 0, sizeof(hd));  // same-process peers still work through the raw pointer
    }
    std::memcpy(blob->ipc, &hd, sizeof(hd));
    int dev = 0;
    cudaGetDevice(&dev);
    blob->magic = kShareMagic;
    blob->pid = (long long)getpid();
    blob->ptr = (unsigned long long)(uintptr_t)*p;
    blob->bytes = bytes;
    blob->device = dev;
    return 0;
}
inline void shared_release(void *p, size_t, const ShareBlob *) {
    if (p) cudaFree(p);
}
inline int shared_open(const ShareBlob *blob, int my_device, void **p, bool *mapped) {
    *mapped = false;
    if (blob->magic != kShareMagic) return 1;
    if (blob->pid == (long long)getpid()) {  // same process: the allocation is addressable once peer access is on
        if (blob->device != my_device) {
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, my_device, blob->device) != cudaSuccess || !can) return 1;
            cudaError_t e = cudaDeviceEnablePeerAccess(blob->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return 1;
            cudaGetLastError();
        }
        *p = (void *)(uintptr_t)blob->ptr;
        return 0;
    }
    cudaIpcMemHandle_t hd;
    std::memcpy(&hd, blob->ipc, sizeof(hd));
    if (cudaIpcOpenMemHandle(p, hd, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        return 1;
    }
    *mapped = true;
    return 0;
}
inline void shared_close(void *p, size_t) {
    if (p) cudaIpcCloseMemHandle(p);
}
}  // namespace rt
}  // namespace dlt

#endif
