// sygnals_b200/csrc/syg_async.cuh
//
// Asynchronous bulk copies (the TMA engine's non-tensor form, cp.async.bulk -> SASS UBLKCP) and the mbarrier objects that track
// them, as thin wrappers over PTX for sm_100a -- plus CPU stand-ins for the g++ -DSYG_EMU test build (tests/emu), where a copy
// completes at issue and a wait yields to the other fibers.
//
//   producer (one thread)                                   consumers (every thread)
//     mbar_expect_tx(bar, bytes)   arm: 1 arrival + bytes       mbar_wait(bar, parity)   spin on try_wait (acquire)
//     bulk_g2s(dst, src, bytes, bar)  the engine signals bar        ... read dst from shared memory ...
//
// Constraints of cp.async.bulk: source, destination and size are multiples of 16 bytes.
#pragma once

#include "syg_platform.h"

namespace sygdev {

#ifndef SYG_EMU

SYG_DEVICE SYG_INLINE unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

SYG_DEVICE SYG_INLINE void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
// makes the initialised barriers visible to the async proxy (the copy engine); follow with a CTA barrier
SYG_DEVICE SYG_INLINE void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

// one arrival of the calling thread + `bytes` of pending transaction bytes for the current phase
SYG_DEVICE SYG_INLINE void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}

SYG_DEVICE SYG_INLINE bool mbar_try_wait(unsigned long long* bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_addr(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// blocks until the phase with the given parity has completed (its arrivals are in and its bytes have landed)
SYG_DEVICE SYG_INLINE void mbar_wait(unsigned long long* bar, unsigned parity) {
    while (!mbar_try_wait(bar, parity)) {}
}

// global -> shared bulk copy through the TMA engine; completion is signalled on `bar` as `bytes` transaction bytes
SYG_DEVICE SYG_INLINE void bulk_g2s(void* smem_dst, const void* gmem_src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}

// shared -> global bulk copy (bulk-group completion): writes made by ordinary stores must be fenced into the async proxy first
SYG_DEVICE SYG_INLINE void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
SYG_DEVICE SYG_INLINE void bulk_s2g(void* gmem_dst, const void* smem_src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_addr(smem_src)), "r"(bytes) : "memory");
}
SYG_DEVICE SYG_INLINE void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// the committed groups have finished READING shared memory (the source may be overwritten); the writes may still be in flight
SYG_DEVICE SYG_INLINE void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
SYG_DEVICE SYG_INLINE void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

#else  // ------------------------------------------------------------------------------------------------ CPU emulator build

// mbarrier image: low word = pending arrivals | phase << 31, high word = pending transaction bytes; `count` restored per phase
struct EmuBar { unsigned pend_phase; int tx; };
static_assert(sizeof(EmuBar) == sizeof(unsigned long long), "mbarrier image");
inline unsigned& emu_bar_count(unsigned long long* bar) {          // arrival count per barrier, kept beside the block context
    static thread_local std::map<void*, unsigned> counts;
    return counts[bar];
}
inline void emu_bar_settle(unsigned long long* bar) {
    EmuBar* b = reinterpret_cast<EmuBar*>(bar);
    if ((b->pend_phase & 0x7fffffffu) == 0 && b->tx == 0)
        b->pend_phase = ((b->pend_phase ^ 0x80000000u) & 0x80000000u) | emu_bar_count(bar);
}
inline void mbar_init(unsigned long long* bar, unsigned count) {
    EmuBar* b = reinterpret_cast<EmuBar*>(bar);
    b->pend_phase = count;
    b->tx = 0;
    emu_bar_count(bar) = count;
}
inline void mbar_init_fence() {}
inline void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    EmuBar* b = reinterpret_cast<EmuBar*>(bar);
    b->tx += (int)bytes;
    b->pend_phase -= 1;
    emu_bar_settle(bar);
}
inline bool mbar_try_wait(unsigned long long* bar, unsigned parity) {
    return (reinterpret_cast<EmuBar*>(bar)->pend_phase >> 31) != (parity & 1u);
}
inline void mbar_wait(unsigned long long* bar, unsigned parity) {
    while (!mbar_try_wait(bar, parity)) ::sygemu::yield();
}
inline void bulk_g2s(void* smem_dst, const void* gmem_src, unsigned bytes, unsigned long long* bar) {
    if (((uintptr_t)smem_dst | (uintptr_t)gmem_src | bytes) & 15u) { std::fprintf(stderr, "emu: misaligned bulk copy\n"); std::abort(); }
    std::memcpy(smem_dst, gmem_src, bytes);
    reinterpret_cast<EmuBar*>(bar)->tx -= (int)bytes;
    emu_bar_settle(bar);
}
inline void fence_async_smem() {}
inline void bulk_s2g(void* gmem_dst, const void* smem_src, unsigned bytes) {
    if (((uintptr_t)gmem_dst | (uintptr_t)smem_src | bytes) & 15u) { std::fprintf(stderr, "emu: misaligned bulk copy\n"); std::abort(); }
    std::memcpy(gmem_dst, smem_src, bytes);
}
inline void bulk_commit() {}
inline void bulk_wait_read_all() {}
inline void bulk_wait_all() {}

#endif

}  // namespace sygdev
