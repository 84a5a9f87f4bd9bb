// simt_runtime.cpp -- SIMT emulator runtime (TEST INFRASTRUCTURE ONLY; see include/cuda_runtime.h).
//
// Execution model.  A launch hands its blocks to a small pool of OS threads.  The CUDA threads of one
// block are coroutines (ucontext) scheduled cooperatively on the OS thread that runs the block: a CUDA
// thread runs until it reaches __syncthreads, a warp collective or the end of the kernel, then the next
// one runs.  __shared__ variables are `static thread_local`, i.e. private to the OS thread and therefore
// to the block it is executing.  __syncthreads is a barrier over the threads that have not left the
// kernel yet; a warp collective is a rendezvous of the live lanes named in its mask (keyed by the mask,
// so disjoint groups of a diverged warp can run their own collectives).  Atomics are real atomics because
// blocks of the same launch do run concurrently.  A block in which no thread can make progress is
// reported as a deadlock.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <sys/mman.h>
#include <ucontext.h>
#include <unistd.h>

#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

thread_local uint3 threadIdx, blockIdx;
thread_local dim3 blockDim, gridDim;

namespace simt {
namespace {

// fatal emulator diagnostics end the process without a core dump
[[noreturn]] void die() {
    fflush(stderr);
    _exit(70);
}

constexpr size_t kStackBytes = 256 * 1024;
constexpr int kMaxThreads = 1024;

enum State { kRunnable, kWaitBlock, kWaitWarp, kDone };

struct Rendezvous {
    unsigned mask, arrived;
    unsigned long long vals[32], result[32];
};

struct WarpState {
    unsigned alive;
    int n_rv;
    Rendezvous rv[40];
};

struct Fiber {
    ucontext_t ctx;
    State state;
    uint3 tid3;
};

struct BlockExec {  // one per pool thread
    ucontext_t sched;
    Fiber* fibers = nullptr;
    char* stacks = nullptr;
    WarpState* warps = nullptr;
    void* dyn = nullptr;
    size_t dyn_cap = 0;
    int n = 0, cur = -1, alive = 0;
    int b_arrived = 0, b_or = 0, b_or_result = 0;
    const std::function<void()>* body = nullptr;
    const char* name = "";
    int order[kMaxThreads];
    unsigned long long rng = 0x9e3779b97f4a7c15ull;
};

thread_local BlockExec* t_exec = nullptr;
int g_schedule = 0;  // 0 in thread order, 1 reversed, 2 shuffled before every scheduling round

struct Pool {
    std::mutex mu;
    std::condition_variable cv_job, cv_done;
    std::vector<std::thread> workers;
    unsigned long long job_id = 0;
    int active = 0;
    dim3 grid, block;
    size_t dyn_bytes = 0;
    const std::function<void()>* body = nullptr;
    const char* name = "";
    std::atomic<unsigned long long> next_block{0};
    unsigned long long n_blocks = 0;
};
// leaked on purpose: detached workers still wait on its condition variables when the process exits
Pool& g = *new Pool;

void to_scheduler(BlockExec* e) { swapcontext(&e->fibers[e->cur].ctx, &e->sched); }

void release_block_barrier(BlockExec* e) {
    e->b_arrived = 0;
    e->b_or_result = e->b_or;
    e->b_or = 0;
    for (int i = 0; i < e->n; ++i)
        if (e->fibers[i].state == kWaitBlock) e->fibers[i].state = kRunnable;
}

void complete_rendezvous(BlockExec* e, int warp, Rendezvous& r) {
    memcpy(r.result, r.vals, sizeof r.result);
    const unsigned was = r.arrived;
    r.arrived = 0;
    for (int l = 0; l < 32; ++l)
        if ((was >> l) & 1u) {
            Fiber& f = e->fibers[warp * 32 + l];
            if (f.state == kWaitWarp) f.state = kRunnable;
        }
}

void fiber_main() {
    BlockExec* e = t_exec;
    (*e->body)();
    // the thread leaves the kernel: barriers and collectives no longer wait for it
    Fiber& f = e->fibers[e->cur];
    f.state = kDone;
    --e->alive;
    WarpState& w = e->warps[e->cur >> 5];
    w.alive &= ~(1u << (e->cur & 31));
    for (int i = 0; i < w.n_rv; ++i) {
        Rendezvous& r = w.rv[i];
        const unsigned need = r.mask & w.alive;
        if (r.arrived && (r.arrived & need) == need) complete_rendezvous(e, e->cur >> 5, r);
    }
    if (e->alive > 0 && e->b_arrived >= e->alive) release_block_barrier(e);
    to_scheduler(e);
    die();  // a finished fiber is never resumed
}

void run_block(BlockExec* e, unsigned bx, unsigned by, unsigned bz) {
    const dim3 block = g.block;
    const int n = e->n;
    blockIdx.x = bx; blockIdx.y = by; blockIdx.z = bz;
    blockDim = block;
    gridDim = g.grid;
    e->alive = n;
    e->b_arrived = 0;
    e->b_or = 0;
    for (int w = 0; w * 32 < n; ++w) {
        const int lanes = n - w * 32 >= 32 ? 32 : n - w * 32;
        e->warps[w].alive = lanes == 32 ? 0xffffffffu : ((1u << lanes) - 1u);
        e->warps[w].n_rv = 0;
    }
    for (int i = 0; i < n; ++i) {
        Fiber& f = e->fibers[i];
        f.state = kRunnable;
        f.tid3.x = i % block.x;
        f.tid3.y = (i / block.x) % block.y;
        f.tid3.z = i / (block.x * block.y);
        getcontext(&f.ctx);
        f.ctx.uc_stack.ss_sp = e->stacks + (size_t)i * kStackBytes;
        f.ctx.uc_stack.ss_size = kStackBytes;
        f.ctx.uc_link = nullptr;
        makecontext(&f.ctx, fiber_main, 0);
    }
    // SIMT_SCHEDULE=reverse | random[:seed] changes the order in which runnable threads are resumed: a kernel
    // whose result depends on that order has a missing barrier (or relies on warp-synchronous execution)
    int* order = e->order;
    for (int i = 0; i < n; ++i) order[i] = g_schedule == 1 ? n - 1 - i : i;
    while (e->alive > 0) {
        bool progressed = false;
        if (g_schedule == 2)
            for (int i = n - 1; i > 0; --i) {
                e->rng = e->rng * 6364136223846793005ull + 1442695040888963407ull;
                const int j = (int)((e->rng >> 33) % (unsigned long long)(i + 1));
                const int t = order[i]; order[i] = order[j]; order[j] = t;
            }
        for (int oi = 0; oi < n; ++oi) {
            const int i = order[oi];
            Fiber& f = e->fibers[i];
            if (f.state != kRunnable) continue;
            progressed = true;
            e->cur = i;
            threadIdx = f.tid3;
            swapcontext(&e->sched, &f.ctx);
        }
        if (!progressed) {
            int wb = 0, ww = 0;
            for (int i = 0; i < n; ++i) {
                wb += e->fibers[i].state == kWaitBlock;
                ww += e->fibers[i].state == kWaitWarp;
            }
            fprintf(stderr, "simt: deadlock in %s, block (%u,%u,%u): %d threads wait at __syncthreads, %d in a warp "
                            "collective, %d alive\n", e->name, bx, by, bz, wb, ww, e->alive);
            die();
        }
    }
    e->cur = -1;
}

void worker_main(int wid) {
    BlockExec* e = new BlockExec;
    e->fibers = new Fiber[kMaxThreads];
    e->warps = new WarpState[kMaxThreads / 32];
    e->stacks = (char*)mmap(nullptr, kStackBytes * kMaxThreads, PROT_READ | PROT_WRITE,
                            MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (e->stacks == (char*)MAP_FAILED) { perror("simt: mmap"); die(); }
    t_exec = e;
    unsigned long long seen = 0;
    if (const char* sch = getenv("SIMT_SCHEDULE"))
        if (!strncmp(sch, "random:", 7)) e->rng ^= strtoull(sch + 7, nullptr, 10) * 0x2545f4914f6cdd1dull;
    e->rng += (unsigned long long)wid * 0x632be59bd9b4e019ull;
    for (;;) {
        {
            std::unique_lock<std::mutex> lk(g.mu);
            g.cv_job.wait(lk, [&] { return g.job_id != seen; });
            seen = g.job_id;
        }
        e->n = (int)(g.block.x * g.block.y * g.block.z);
        e->body = g.body;
        e->name = g.name;
        if (g.dyn_bytes > e->dyn_cap) {
            free(e->dyn);
            if (posix_memalign(&e->dyn, 256, g.dyn_bytes)) die();
            e->dyn_cap = g.dyn_bytes;
        }
        for (;;) {
            const unsigned long long b = g.next_block.fetch_add(1);
            if (b >= g.n_blocks) break;
            const unsigned bx = (unsigned)(b % g.grid.x);
            const unsigned by = (unsigned)((b / g.grid.x) % g.grid.y);
            const unsigned bz = (unsigned)(b / ((unsigned long long)g.grid.x * g.grid.y));
            if (g.dyn_bytes) memset(e->dyn, 0xcd, g.dyn_bytes);
            run_block(e, bx, by, bz);
        }
        {
            std::unique_lock<std::mutex> lk(g.mu);
            if (--g.active == 0) g.cv_done.notify_all();
        }
    }
}

}  // namespace

static void die_public() { die(); }

int lane_id() { return t_exec->cur & 31; }
void* dyn_smem() { return t_exec->dyn; }

void sync_block() { (void)sync_block_or(0); }

int sync_block_or(int pred) {
    BlockExec* e = t_exec;
    if (pred) e->b_or = 1;
    if (++e->b_arrived >= e->alive) {
        release_block_barrier(e);
    } else {
        e->fibers[e->cur].state = kWaitBlock;
        to_scheduler(e);
    }
    return e->b_or_result;
}

void warp_exchange(unsigned mask, unsigned long long v, unsigned long long* out) {
    BlockExec* e = t_exec;
    const int lane = e->cur & 31, warp = e->cur >> 5;
    WarpState& w = e->warps[warp];
    if (!((mask >> lane) & 1u)) {
        fprintf(stderr, "simt: %s: lane %d calls a warp collective whose mask %08x does not name it\n", e->name, lane,
                mask);
        die_public();
    }
    Rendezvous* r = nullptr;
    for (int i = 0; i < w.n_rv; ++i)
        if (w.rv[i].mask == mask) { r = &w.rv[i]; break; }
    if (!r) {
        if (w.n_rv == (int)(sizeof w.rv / sizeof w.rv[0])) {
            fprintf(stderr, "simt: %s: too many distinct collective masks in one warp\n", e->name);
            die_public();
        }
        r = &w.rv[w.n_rv++];
        r->mask = mask;
        r->arrived = 0;
    }
    if ((r->arrived >> lane) & 1u) {
        fprintf(stderr, "simt: %s: lane %d re-enters a collective (mask %08x) that is still pending\n", e->name, lane,
                mask);
        die_public();
    }
    r->vals[lane] = v;
    r->arrived |= 1u << lane;
    const unsigned need = mask & w.alive;
    if ((r->arrived & need) == need) {
        complete_rendezvous(e, warp, *r);
    } else {
        e->fibers[e->cur].state = kWaitWarp;
        to_scheduler(e);
    }
    memcpy(out, r->result, sizeof r->result);
}

void launch(dim3 grid, dim3 block, size_t dyn_bytes, const std::function<void()>& body, const char* name) {
    const unsigned long long n = (unsigned long long)block.x * block.y * block.z;
    if (n == 0 || n > kMaxThreads || grid.x == 0 || grid.y == 0 || grid.z == 0 || grid.y > 65535 || grid.z > 65535 ||
        dyn_bytes > 232448) {
        fprintf(stderr, "simt: invalid launch configuration for %s: grid %ux%ux%u block %ux%ux%u smem %zu\n", name,
                grid.x, grid.y, grid.z, block.x, block.y, block.z, dyn_bytes);
        die_public();
    }
    static std::mutex launch_mu;  // launches are serialised, like work on one stream
    std::lock_guard<std::mutex> guard(launch_mu);
    static const bool trace = getenv("SIMT_TRACE") != nullptr;  // kernel names to stderr (which variant ran?)
    if (trace) fprintf(stderr, "simt: launch %s grid %ux%ux%u block %ux%ux%u smem %zu\n", name, grid.x, grid.y, grid.z,
                       block.x, block.y, block.z, dyn_bytes);
    if (g.workers.empty()) {
        if (const char* sch = getenv("SIMT_SCHEDULE")) {
            if (!strncmp(sch, "reverse", 7)) g_schedule = 1;
            else if (!strncmp(sch, "random", 6)) g_schedule = 2;
        }
        unsigned hw = std::thread::hardware_concurrency();
        const char* env = getenv("SIMT_WORKERS");
        int nw = env ? atoi(env) : (int)(hw ? (hw > 8 ? 8 : hw) : 4);
        if (nw < 1) nw = 1;
        for (int i = 0; i < nw; ++i) {
            g.workers.emplace_back(worker_main, i);
            g.workers.back().detach();
        }
    }
    std::unique_lock<std::mutex> lk(g.mu);
    g.grid = grid;
    g.block = block;
    g.dyn_bytes = dyn_bytes;
    g.body = &body;
    g.name = name;
    g.n_blocks = (unsigned long long)grid.x * grid.y * grid.z;
    g.next_block.store(0);
    g.active = (int)g.workers.size();
    ++g.job_id;
    g.cv_job.notify_all();
    g.cv_done.wait(lk, [&] { return g.active == 0; });
}

}  // namespace simt
