// tests/emu/syg_emu.cpp -- TEST INFRASTRUCTURE ONLY (see syg_emu.h).
#include "syg_emu.h"

thread_local dim3 threadIdx, blockIdx, blockDim, gridDim;

namespace sygemu {

thread_local BlockCtx* g_blk = nullptr;
bool g_banks = false;
std::mutex g_bank_mu;
std::map<int, std::pair<long, long>> g_bank_stats;

static void set_thread(int tid) {
    BlockCtx* b = g_blk;
    b->cur = tid;
    unsigned bx = blockDim.x, by = blockDim.y;
    threadIdx.x = tid % bx;
    threadIdx.y = (tid / bx) % by;
    threadIdx.z = tid / (bx * by);
}

void yield() {
    BlockCtx* b = g_blk;
    int me = b->cur;
    swapcontext(&b->fibers[me].ctx, &b->sched);
}

static void trampoline() {
    BlockCtx* b = g_blk;
    int me = b->cur;
    (*b->body)();
    b->fibers[me].done = true;
    b->alive--;
    // a thread that exits counts as arrived for any barrier in flight
    if (b->alive > 0 && b->bar_count >= b->alive) {
        b->bar_count = 0;
        b->bar_gen++;
    }
    swapcontext(&b->fibers[me].ctx, &b->sched);
}

void block_barrier() {
    BlockCtx* b = g_blk;
    unsigned gen = b->bar_gen;
    if (++b->bar_count >= b->alive) {
        b->bar_count = 0;
        b->bar_gen++;
        return;
    }
    while (b->bar_gen == gen) yield();
}

void warp_barrier(unsigned mask) {
    BlockCtx* b = g_blk;
    int tid = b->cur;
    int wbase = (tid / 32) * 32;
    WarpState& w = b->warps[tid / 32];
    int expected = 0;
    for (int l = 0; l < 32; ++l)
        if (((mask >> l) & 1u) && wbase + l < b->nthreads && !b->fibers[wbase + l].done) expected++;
    unsigned gen = w.gen;
    if (++w.count >= expected) {
        w.count = 0;
        w.gen++;
        return;
    }
    while (w.gen == gen) yield();
}

void smem_access(const void* p, int bytes, int site) {
    if (!g_banks) return;
    BlockCtx* b = g_blk;
    int tid = b->cur;
    Fiber& f = b->fibers[tid];
    int occ = f.occ[site]++;
    b->bank.push_back(BankRec{tid / 32, site, occ, tid % 32, (uintptr_t)p, bytes});
}

static void bank_flush(BlockCtx& b) {
    if (b.bank.empty()) return;
    std::map<std::tuple<int, int, int>, std::vector<const BankRec*>> groups;
    for (auto& r : b.bank) groups[{r.warp, r.site, r.occ}].push_back(&r);
    std::lock_guard<std::mutex> lk(g_bank_mu);
    for (auto& kv : groups) {
        // wavefronts = max over banks of the number of distinct 4-byte words touched in that bank
        std::map<int, std::vector<uintptr_t>> per_bank;
        for (auto* r : kv.second)
            for (int o = 0; o < r->bytes; o += 4) {
                uintptr_t word = (r->addr + o) >> 2;
                auto& v = per_bank[(int)(word % 32)];
                if (std::find(v.begin(), v.end(), word) == v.end()) v.push_back(word);
            }
        long wf = 0;
        for (auto& pb : per_bank) wf = std::max<long>(wf, (long)pb.second.size());
        // 8/16-byte accesses need >= 2/4 wavefronts for a full warp
        auto& st = g_bank_stats[std::get<1>(kv.first)];
        st.first += 1;
        st.second += wf;
    }
    b.bank.clear();
}

void bank_report(FILE* f) {
    std::lock_guard<std::mutex> lk(g_bank_mu);
    for (auto& kv : g_bank_stats)
        std::fprintf(f, "smem site line %d: %ld warp-accesses, %.2f wavefronts/access\n", kv.first, kv.second.first,
                     (double)kv.second.second / (double)std::max<long>(1, kv.second.first));
}
void bank_reset() {
    std::lock_guard<std::mutex> lk(g_bank_mu);
    g_bank_stats.clear();
}

static void run_block(BlockCtx& b, dim3 grid, dim3 block, unsigned bid, std::function<void()>& body, bool reverse) {
    g_blk = &b;
    gridDim = grid;
    blockDim = block;
    blockIdx.x = bid % grid.x;
    blockIdx.y = (bid / grid.x) % grid.y;
    blockIdx.z = bid / (grid.x * grid.y);
    int n = b.nthreads;
    b.alive = n;
    b.bar_count = 0;
    b.body = &body;
    for (auto& w : b.warps) { w.count = 0; }
    for (int t = 0; t < n; ++t) {
        Fiber& f = b.fibers[t];
        f.done = false;
        f.occ.clear();
        getcontext(&f.ctx);
        f.ctx.uc_stack.ss_sp = f.stack;
        f.ctx.uc_stack.ss_size = kStack;
        f.ctx.uc_link = &b.sched;
        makecontext(&f.ctx, (void (*)())trampoline, 0);
    }
    while (b.alive > 0) {
        for (int i = 0; i < n; ++i) {
            int t = reverse ? n - 1 - i : i;
            if (b.fibers[t].done) continue;
            set_thread(t);
            swapcontext(&b.sched, &b.fibers[t].ctx);
        }
    }
    bank_flush(b);
}

void launch_impl(dim3 grid, dim3 block, size_t smem, std::function<void()> body) {
    const char* eb = std::getenv("SYG_EMU_BANKS");
    g_banks = eb && eb[0] == '1';
    const char* er = std::getenv("SYG_EMU_REVERSE");
    bool reverse = er && er[0] == '1';
    unsigned nblocks = grid.x * grid.y * grid.z;
    int nthreads = (int)(block.x * block.y * block.z);
    unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    const char* et = std::getenv("SYG_EMU_THREADS");
    unsigned nworkers = et ? (unsigned)std::max(1, std::atoi(et)) : hw;
    nworkers = std::min(nworkers, std::max(1u, nblocks));
    std::atomic<unsigned> next{0};
    auto worker = [&]() {
        BlockCtx b;
        b.nthreads = nthreads;
        b.fibers.resize(nthreads);
        b.warps.resize((nthreads + 31) / 32);
        for (auto& f : b.fibers) f.stack = (char*)std::malloc(kStack);
        // dynamic shared memory with guard zones on both sides: filled with 0xFF (a NaN pattern as float, so a stray read poisons
        // the result and the parity tests notice) and checked after every block (a stray write aborts)
        constexpr size_t kGuard = 4096;
        const size_t span = ((smem + 15) / 16) * 16;
        unsigned char* raw = (unsigned char*)std::aligned_alloc(1024, ((kGuard + span + kGuard + 1023) / 1024) * 1024);
        std::memset(raw, 0xFF, kGuard + span + kGuard);
        b.dyn = raw + kGuard;
        std::function<void()> local_body = body;
        for (;;) {
            unsigned bid = next.fetch_add(1);
            if (bid >= nblocks) break;
            run_block(b, grid, block, bid, local_body, reverse);
            for (size_t i = 0; i < kGuard; ++i) {
                if (raw[i] != 0xFF || raw[kGuard + span + i] != 0xFF) {
                    std::fprintf(stderr, "sygemu: block %u wrote outside its %zu bytes of dynamic shared memory (guard byte %zu)\n", bid, smem, i);
                    std::abort();
                }
            }
        }
        for (auto& f : b.fibers) std::free(f.stack);
        std::free(raw);
        g_blk = nullptr;
    };
    if (nworkers <= 1) {
        worker();
    } else {
        std::vector<std::thread> th;
        for (unsigned i = 0; i < nworkers; ++i) th.emplace_back(worker);
        for (auto& t : th) t.join();
    }
}

}  // namespace sygemu
