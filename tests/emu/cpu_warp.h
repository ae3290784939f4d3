// cpu_warp.h — TEST INFRASTRUCTURE: a lock-step 32-lane warp emulator on ucontext fibres.
//
// Lets circuitsimulator_b200/csrc/bg_device.cuh (the product's warp-level device code) be
// compiled by g++ and executed on the CPU so that the same source can be checked against
// the oracle where there is no GPU.  Lanes run round-robin; every warp collective
// (__shfl_sync, __ballot_sync, __reduce_xor_sync, ...) is a yield point, and the emulator
// checks that all 32 lanes reach the same collective (convergence), which is exactly the
// contract the *_sync intrinsics require on the device.
// Never linked into the product library.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <ucontext.h>
#include <functional>

namespace emu {
enum Kind { K_NONE = 0, K_SHFL, K_SHFLX, K_BALLOT, K_REDXOR, K_SYNC };
struct Warp {
    ucontext_t main_ctx, fib[32];
    char* stacks[32];
    int cur;
    bool done[32];
    uint32_t in[2][32];
    int kind[2][32];
    int phase[32];
    std::function<void()> body;
};
extern thread_local Warp* g_warp;
void run(const std::function<void()>& body);     // runs body on 32 lanes in lock-step

inline void yield_(int kind, uint32_t v) {
    Warp* w = g_warp;
    const int l = w->cur, ph = w->phase[l] & 1;
    w->in[ph][l] = v; w->kind[ph][l] = kind;
    swapcontext(&w->fib[l], &w->main_ctx);
    for (int j = 0; j < 32; j++)
        if (w->kind[ph][j] != kind) { fprintf(stderr, "emu: divergent collective (lane %d kind %d vs lane %d kind %d)\n", l, kind, j, w->kind[ph][j]); abort(); }
    w->phase[l]++;
}
inline const uint32_t* last_inputs() { Warp* w = g_warp; return w->in[(w->phase[w->cur] - 1) & 1]; }
}  // namespace emu

static inline int bg_lane() { return emu::g_warp->cur; }

static inline uint32_t __shfl_sync(uint32_t, uint32_t v, int src) { emu::yield_(emu::K_SHFL, v); return emu::last_inputs()[src & 31]; }
static inline uint32_t __shfl_xor_sync(uint32_t, uint32_t v, int m) { int l = bg_lane(); emu::yield_(emu::K_SHFLX, v); return emu::last_inputs()[(l ^ m) & 31]; }
static inline uint32_t __ballot_sync(uint32_t, bool p) {
    emu::yield_(emu::K_BALLOT, p ? 1u : 0u);
    const uint32_t* in = emu::last_inputs(); uint32_t r = 0;
    for (int j = 0; j < 32; j++) r |= (in[j] & 1u) << j;
    return r;
}
static inline uint32_t __reduce_xor_sync(uint32_t, uint32_t v) {
    emu::yield_(emu::K_REDXOR, v);
    const uint32_t* in = emu::last_inputs(); uint32_t r = 0;
    for (int j = 0; j < 32; j++) r ^= in[j];
    return r;
}
static inline void __syncwarp() { emu::yield_(emu::K_SYNC, 0); }
static inline int __popc(uint32_t x) { return __builtin_popcount(x); }
static inline int __popcll(uint64_t x) { return __builtin_popcountll(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __ffsll(long long x) { return __builtin_ffsll(x); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
