// tests/emu/cuda_emu.cpp -- TEST INFRASTRUCTURE ONLY (see cuda_emu.h).
// Stackful-coroutine scheduler: one coroutine per CUDA thread of the running block.
#include "cuda_emu.h"

#include <vector>

namespace emu {

uint3 g_threadIdx, g_blockIdx;
dim3 g_blockDim, g_gridDim;

extern "C" void emu_switch(void **save_sp, void *load_sp);
asm(".text\n"
    ".globl emu_switch\n"
    ".type emu_switch,@function\n"
    "emu_switch:\n"
    "  pushq %rbp\n"
    "  pushq %rbx\n"
    "  pushq %r12\n"
    "  pushq %r13\n"
    "  pushq %r14\n"
    "  pushq %r15\n"
    "  movq %rsp, (%rdi)\n"
    "  movq %rsi, %rsp\n"
    "  popq %r15\n"
    "  popq %r14\n"
    "  popq %r13\n"
    "  popq %r12\n"
    "  popq %rbx\n"
    "  popq %rbp\n"
    "  ret\n");

static const size_t kStack = 96 * 1024;

struct Co {
    void *sp = nullptr;
    char *stack = nullptr;
    bool done = true;
    uint3 tid;
};
struct Warp {
    unsigned long long in[2][32];
    int aux[2][32];
    unsigned long long out[2][32];
    int arrived = 0;
    int live = 0;
    unsigned live_mask = 0;
    unsigned long long gen = 0;
    int kind[2];
    int width[2];
};

static std::vector<Co> g_co;
static std::vector<Warp> g_warps;
static void *g_sched_sp = nullptr;
static int g_cur = -1;
static const std::function<void()> *g_body = nullptr;
static int g_block_live = 0, g_block_arrived = 0;
static unsigned long long g_block_gen = 0;
static unsigned long long g_progress = 0;

static void yield_to_sched() { emu_switch(&g_co[g_cur].sp, g_sched_sp); }

static void trampoline() {
    (*g_body)();
    Co &c = g_co[g_cur];
    c.done = true;
    int w = g_cur / 32, lane = g_cur % 32;
    g_warps[w].live--;
    g_warps[w].live_mask &= ~(1u << lane);
    g_block_live--;
    g_progress++;
    // a thread that exits counts as arrived for barriers other threads are waiting in
    if (g_block_live > 0 && g_block_arrived == g_block_live) {
        g_block_arrived = 0;
        g_block_gen++;
    }
    if (g_warps[w].live > 0 && g_warps[w].arrived == g_warps[w].live) {
        fprintf(stderr, "cuda_emu: a lane exited while its warp waits in a collective (undefined on the GPU)\n");
        abort();
    }
    yield_to_sched();
    abort();
}

void block_sync() {
    unsigned long long g = g_block_gen;
    g_block_arrived++;
    g_progress++;
    if (g_block_arrived == g_block_live) {
        g_block_arrived = 0;
        g_block_gen++;
        return;
    }
    while (g_block_gen == g) yield_to_sched();
}

static void finish_collective(Warp &w, int b) {
    int kind = w.kind[b], width = w.width[b];
    if (kind == 4) {
        unsigned bal = 0;
        for (int l = 0; l < 32; l++)
            if ((w.live_mask >> l) & 1)
                if (w.in[b][l]) bal |= 1u << l;
        for (int l = 0; l < 32; l++) w.out[b][l] = bal;
    } else if (kind == 5) {
    } else {
        for (int l = 0; l < 32; l++) {
            if (!((w.live_mask >> l) & 1)) continue;
            int base = l & ~(width - 1);
            int src;
            bool ok = true;
            if (kind == 0)
                src = base + (w.aux[b][l] & (width - 1));
            else if (kind == 1) {
                src = l ^ w.aux[b][l];
                ok = (src >= base && src < base + width);
            } else if (kind == 2) {
                src = l + w.aux[b][l];
                ok = src < base + width;
            } else {
                src = l - w.aux[b][l];
                ok = src >= base;
            }
            if (!ok) src = l;
            if (!((w.live_mask >> src) & 1)) {
                fprintf(stderr, "cuda_emu: shuffle reads lane %d which has exited\n", src);
                abort();
            }
            w.out[b][l] = w.in[b][src];
        }
    }
}

unsigned long long warp_collective(int kind, unsigned mask, unsigned long long payload, int aux, int width) {
    int wi = g_cur / 32, lane = g_cur % 32;
    Warp &w = g_warps[wi];
    if ((mask & w.live_mask) != w.live_mask) {
        fprintf(stderr, "cuda_emu: partial-mask collective (mask %08x, live %08x) is not supported\n", mask, w.live_mask);
        abort();
    }
    unsigned long long g = w.gen;
    int b = (int)(g & 1);
    if (w.arrived == 0) {
        w.kind[b] = kind;
        w.width[b] = width;
    } else if (w.kind[b] != kind || w.width[b] != width) {
        fprintf(stderr, "cuda_emu: lanes of one warp are in different collectives (%d vs %d): divergent collective\n",
                w.kind[b], kind);
        abort();
    }
    w.in[b][lane] = payload;
    w.aux[b][lane] = aux;
    w.arrived++;
    g_progress++;
    if (w.arrived == w.live) {
        finish_collective(w, b);
        w.arrived = 0;
        w.gen++;
    } else {
        while (w.gen == g) yield_to_sched();
    }
    return w.out[b][lane];
}

void launch(dim3 grid, dim3 block, const std::function<void()> &body) {
    unsigned nthreads = block.x * block.y * block.z;
    if (g_co.size() < nthreads) {
        size_t old = g_co.size();
        g_co.resize(nthreads);
        for (size_t i = old; i < nthreads; i++) g_co[i].stack = (char *)aligned_alloc(64, kStack);
    }
    g_warps.assign((nthreads + 31) / 32, Warp());
    g_body = &body;
    g_blockDim = block;
    g_gridDim = grid;
    for (unsigned bz = 0; bz < grid.z; bz++)
        for (unsigned by = 0; by < grid.y; by++)
            for (unsigned bx = 0; bx < grid.x; bx++) {
                g_blockIdx = uint3{bx, by, bz};
                for (auto &w : g_warps) {
                    w.arrived = 0;
                    w.live = 0;
                    w.live_mask = 0;
                    w.gen = 0;
                }
                g_block_live = (int)nthreads;
                g_block_arrived = 0;
                g_block_gen = 0;
                for (unsigned t = 0; t < nthreads; t++) {
                    Co &c = g_co[t];
                    c.done = false;
                    c.tid = uint3{t % block.x, (t / block.x) % block.y, t / (block.x * block.y)};
                    uintptr_t top = ((uintptr_t)(c.stack + kStack)) & ~(uintptr_t)15;
                    void **sp = (void **)(top - 64);
                    for (int i = 0; i < 6; i++) sp[i] = nullptr;
                    sp[6] = (void *)&trampoline;
                    sp[7] = nullptr;
                    c.sp = sp;
                    g_warps[t / 32].live++;
                    g_warps[t / 32].live_mask |= 1u << (t % 32);
                }
                int remaining = (int)nthreads;
                while (remaining > 0) {
                    unsigned long long before = g_progress;
                    remaining = 0;
                    for (unsigned t = 0; t < nthreads; t++) {
                        Co &c = g_co[t];
                        if (c.done) continue;
                        g_cur = (int)t;
                        g_threadIdx = c.tid;
                        emu_switch(&g_sched_sp, c.sp);
                        if (!c.done) remaining++;
                    }
                    if (remaining > 0 && g_progress == before) {
                        fprintf(stderr, "cuda_emu: deadlock in block (%u,%u,%u): %d threads blocked\n", bx, by, bz, remaining);
                        abort();
                    }
                }
            }
    g_body = nullptr;
}

}  // namespace emu
