// emu_lib.cpp — TEST INFRASTRUCTURE: runs the product's warp-level device code
// (circuitsimulator_b200/csrc/bg_device.cuh, bg_philox.cuh) on the CPU warp emulator.
#include "cpu_warp.h"
#include "bg_device.cuh"
#include "bg_warp_ops.cuh"
#include <string.h>

namespace emu {
thread_local Warp* g_warp = nullptr;
static void trampoline() {
    Warp* w = g_warp;
    w->body();
    w->done[w->cur] = true;
    swapcontext(&w->fib[w->cur], &w->main_ctx);
}
void run(const std::function<void()>& body) {
    static thread_local Warp* w = nullptr;
    const size_t STK = 1 << 18;
    if (!w) { w = new Warp(); for (int i = 0; i < 32; i++) w->stacks[i] = (char*)malloc(STK); }
    g_warp = w;
    w->body = body;
    for (int i = 0; i < 32; i++) {
        w->done[i] = false; w->phase[i] = 0;
        getcontext(&w->fib[i]);
        w->fib[i].uc_stack.ss_sp = w->stacks[i];
        w->fib[i].uc_stack.ss_size = STK;
        w->fib[i].uc_link = &w->main_ctx;
        makecontext(&w->fib[i], (void (*)())trampoline, 0);
    }
    for (int j = 0; j < 32; j++) { w->kind[0][j] = w->kind[1][j] = K_NONE; }
    while (true) {
        int ndone = 0;
        for (int i = 0; i < 32; i++) {
            if (w->done[i]) { ndone++; continue; }
            w->cur = i;
            swapcontext(&w->main_ctx, &w->fib[i]);
            if (w->done[i]) ndone++;
        }
        if (ndone == 32) break;
        if (ndone != 0) { fprintf(stderr, "emu: lanes finished at different collectives\n"); abort(); }
    }
}
}  // namespace emu

using namespace bg;

extern "C" {

// <b|a> through the generic path
void emu_inner_product(const bg_state* a, const bg_state* b, int32_t* epm) {
    const int n = a->n;
    emu::run([&]() {
        int e, p, m;
        if (n <= 32) warp_inner_product<1>(a, b, e, p, m); else warp_inner_product<2>(a, b, e, p, m);
        if (bg_lane() == 0) { epm[0] = e; epm[1] = p; epm[2] = m; }
    });
}

// theta (optionally projected by P) against every term of the decomposition
// epm: chi x 3; returns alive flag; npf = number of 2^-1/2 factors; kout = dim after projection
int emu_terms(const bg_state* theta, const bg_projector* P, int project, int exact, int t, int nterms,
              const uint64_t* terms, int32_t* epm, int* npf_out, int* k_out, long long* zw_out) {
    int alive = 1;
    emu::run([&]() {
        if (t <= 32) warp_sample_terms<1>(theta, P, project, exact, t, nterms, terms, epm, &alive, npf_out, k_out, zw_out);
        else warp_sample_terms<2>(theta, P, project, exact, t, nterms, terms, epm, &alive, npf_out, k_out, zw_out);
    });
    return alive;
}

int emu_measure_pauli(bg_state* st, uint64_t* A, int m, uint64_t zeta, uint64_t xi) {
    int res = 0;
    const int n = st->n;
    emu::run([&]() {
        int r;
        if (n <= 32) r = warp_measure_pauli<1>(st, A, m, zeta, xi); else r = warp_measure_pauli<2>(st, A, m, zeta, xi);
        if (bg_lane() == 0) res = r;
    });
    return res;
}

void emu_random_state(int n, uint64_t seed, uint32_t bin, uint64_t sample, const double* cdf, bg_state* out, uint64_t* A) {
    emu::run([&]() {
        if (n <= 32) warp_random_state<1>(n, seed, bin, sample, cdf, out, A);
        else warp_random_state<2>(n, seed, bin, sample, cdf, out, A);
    });
}

}  // extern "C"
