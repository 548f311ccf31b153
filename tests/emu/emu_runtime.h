// tests/emu/emu_runtime.h — the fibre scheduler behind tests/emu/cuda_runtime.h. TEST INFRASTRUCTURE ONLY.
// Include once, in the harness's translation unit.
#pragma once
#include <cuda_runtime.h>

namespace emu {

Cta* g_cta = nullptr;
unsigned long long g_collectives = 0, g_events = 0;
unsigned g_rcp_seed = 12345u;
int g_order_mode = -1;
unsigned g_order_seed = 1u;

void trampoline() {
  Cta* c = g_cta;
#ifdef EMU_ASAN
  __sanitizer_finish_switch_fiber(nullptr, &c->sched_stack, &c->sched_stack_size);      // first entry: learn the scheduler's stack
#endif
  c->entry(c->arg);
  c->cur->done = true;
#ifdef EMU_ASAN
  __sanitizer_start_switch_fiber(nullptr, c->sched_stack, c->sched_stack_size);         // nullptr: this fibre never resumes
#endif
  swapcontext(&c->cur->ctx, &c->sched);
}

void run_grid(dim3 grid, dim3 block, void (*entry)(void*), void* arg, size_t dyn_smem_bytes) {
  const int nthreads = (int)(block.x * block.y * block.z);
  const size_t stack_bytes = 256 << 10;
  Cta cta;
  cta.fibres.resize(nthreads);
  for (auto& f : cta.fibres) f.stack = (char*)malloc(stack_bytes);
  cta.bdim = block; cta.gdim = grid; cta.entry = entry; cta.arg = arg;
  std::vector<char> dyn(dyn_smem_bytes + 16);
  cta.dyn_smem = dyn.data() + (16 - (reinterpret_cast<uintptr_t>(dyn.data()) & 15)) % 16;
  Cta* const outer = g_cta;
  g_cta = &cta;
  if (g_order_mode < 0) {
    const char* v = getenv("VH_EMU_ORDER");
    g_order_mode = !v ? 0 : (!strncmp(v, "reverse", 7) ? 1 : (!strncmp(v, "random", 6) ? 2 : 0));
    if (g_order_mode == 2 && strchr(v, ':')) g_order_seed = (unsigned)atoi(strchr(v, ':') + 1) * 2654435761u + 1u;
  }
  std::vector<int> order(nthreads);
  for (int t = 0; t < nthreads; t++) order[t] = t;
  for (unsigned bz = 0; bz < grid.z; bz++) for (unsigned by = 0; by < grid.y; by++) for (unsigned bx = 0; bx < grid.x; bx++) {
    cta.bid = uint3{bx, by, bz};
    cta.warps.assign((nthreads + 31) / 32, Group());
    for (size_t w = 0; w < cta.warps.size(); w++) cta.warps[w].size = std::min(32, nthreads - (int)w * 32);
    cta.all = Group(); cta.all.size = nthreads;
    for (int t = 0; t < nthreads; t++) {
      Fibre& f = cta.fibres[t];
      f.done = false; f.warp_parity = 0;
      f.tid = uint3{(unsigned)t % block.x, ((unsigned)t / block.x) % block.y, (unsigned)t / (block.x * block.y)};
      getcontext(&f.ctx);
      f.ctx.uc_stack.ss_sp = f.stack; f.ctx.uc_stack.ss_size = stack_bytes; f.ctx.uc_link = nullptr;
      makecontext(&f.ctx, (void (*)())trampoline, 0);
    }
    int live = nthreads;
    while (live > 0) {
      const unsigned long long events_before = g_events;
      // Scheduling order of a round. The default runs the threads in index order, which can hide a missing barrier (the
      // writer happens to run before the reader). VH_EMU_ORDER=reverse or =random:<seed> runs every round in another order:
      // results that depend on the order point at unsynchronised communication through shared or global memory.
      if (g_order_mode == 1) { for (int t = 0; t < nthreads; t++) order[t] = nthreads - 1 - t; }
      else if (g_order_mode == 2) { for (int t = nthreads - 1; t > 0; t--) { g_order_seed = g_order_seed * 1664525u + 1013904223u; std::swap(order[t], order[(g_order_seed >> 8) % (unsigned)(t + 1)]); } }
      for (int oi = 0; oi < nthreads; oi++) {
        const int t = order[oi];
        Fibre& f = cta.fibres[t];
        if (f.done) continue;
        cta.cur = &f;
#ifdef EMU_ASAN
        void* fake = nullptr;
        __sanitizer_start_switch_fiber(&fake, f.stack, stack_bytes);
        swapcontext(&cta.sched, &f.ctx);
        __sanitizer_finish_switch_fiber(fake, nullptr, nullptr);
#else
        swapcontext(&cta.sched, &f.ctx);
#endif
        if (f.done) { live--; g_events++; }
      }
      if (g_events == events_before) {   // a whole round in which no barrier opened and no thread finished
        fprintf(stderr, "emu: deadlock — %d threads wait at a collective their warp/CTA cannot complete (divergent collective?)\n", live);
        abort();
      }
    }
  }
  g_cta = outer;
  for (auto& f : cta.fibres) free(f.stack);
}

}  // namespace emu
