// daliti_b200/csrc/dlt_rt.h -- thin runtime layer: device memory, copies, launches.
// The product build is CUDA (nvcc, sm_100a).  With -DDLT_EMU (tests/emu only) the same
// sources compile as plain C++ against the test emulator so kernel logic can be exercised
// without a GPU; that build is test infrastructure and is never loaded by the package.
#pragma once
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "dlt_common.cuh"

#if defined(DLT_EMU)

namespace dlt {
namespace rt {
extern std::atomic<unsigned long long> g_launches;
struct Event {};
inline int event_create(Event *) { return 0; }
inline void event_destroy(Event) {}
inline int event_record(Event, cudaStream_t) { return 0; }
inline float event_elapsed_ms(Event, Event) { return 0.f; }
inline int event_create_untimed(Event *) { return 0; }
inline int stream_wait(cudaStream_t, Event) { return 0; }
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
inline int sync(cudaStream_t) { return 0; }
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
}  // namespace rt
}  // namespace dlt

#else  // ------------------------------------------------------------------ CUDA

namespace dlt { namespace rt { extern std::atomic<unsigned long long> g_launches; } }
#define DLT_LAUNCH(kernel, grid, block, stream, ...)                      \
    do {                                                                  \
        kernel<<<(grid), (block), 0, (stream)>>>(__VA_ARGS__);            \
        ++dlt::rt::g_launches;                                            \
    } while (0)

namespace dlt {
namespace rt {
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
inline int sync(cudaStream_t st) { return cudaStreamSynchronize(st) == cudaSuccess ? 0 : 1; }
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
}  // namespace rt
}  // namespace dlt

#endif
