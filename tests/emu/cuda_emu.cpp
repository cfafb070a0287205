// cuda_emu.cpp -- fibre scheduler of the CPU block emulator (see cuda_emu.h).  TEST INFRASTRUCTURE ONLY.
#include "cuda_emu.h"

thread_local a1emu::Dim3 threadIdx, blockIdx, blockDim, gridDim;

// minimal x86-64 System V context switch: callee-saved registers + stack pointer
extern "C" void a1emu_switch(void** save_sp, void* load_sp);
asm(R"(
.text
.globl a1emu_switch
.type a1emu_switch,@function
a1emu_switch:
  pushq %rbp
  pushq %rbx
  pushq %r12
  pushq %r13
  pushq %r14
  pushq %r15
  movq %rsp, (%rdi)
  movq %rsi, %rsp
  popq %r15
  popq %r14
  popq %r13
  popq %r12
  popq %rbx
  popq %rbp
  ret
.size a1emu_switch,.-a1emu_switch
)");

namespace a1emu {

thread_local Block* g_blk = nullptr;
thread_local int g_f32 = 0;

void yield_to_scheduler() {
  Block& b = *g_blk;
  a1emu_switch(&b.fib[b.cur].sp, b.sched_sp);
}

static void trampoline() {
  Block& b = *g_blk;
  b.body();
  b.fib[b.cur].done = true;
  ++b.progress;
  for (;;) yield_to_scheduler();
}

static constexpr size_t STACK_BYTES = 512 * 1024;

void run_block(Dim3 block_idx, Dim3 grid_dim, int nthreads, size_t smem_bytes, int order_mode, const std::function<void()>& body,
               unsigned long* n_collectives, unsigned long* n_mma) {
  Block b;
  b.nthreads = nthreads;
  b.body = body;
  b.order_mode = order_mode;
  // shared memory is NOT zero on a GPU: poison it (NaN) so that any read of a never-written word shows up in the results
  static const bool poison = std::getenv("A1EMU_NO_POISON") == nullptr;
  b.smem.assign(smem_bytes / 8 + 2, poison ? std::nan("") : 0.0);
  if (const char* r = std::getenv("A1EMU_POISON_RANGE")) {   // debugging aid: "lo:hi" (doubles) -- poison only that window
    long lo = 0, hi = 0;
    if (std::sscanf(r, "%ld:%ld", &lo, &hi) == 2)
      for (long i = 0; i < (long)b.smem.size(); ++i) b.smem[i] = (i >= lo && i < hi) ? std::nan("") : 0.0;
  }
  b.fib.resize(nthreads);
  b.warps.resize((nthreads + 31) / 32);
  for (size_t w = 0; w < b.warps.size(); ++w) b.warps[w].nlanes = std::min(32, nthreads - 32 * (int)w);
  std::vector<char> stacks((size_t)nthreads * STACK_BYTES + 64);
  for (int t = 0; t < nthreads; ++t) {
    char* top = stacks.data() + (size_t)(t + 1) * STACK_BYTES;
    uintptr_t sp = ((uintptr_t)top) & ~(uintptr_t)15;
    sp -= 64;                                  // rsp0: 16-byte aligned; [0..48) six registers, [48] entry, [56] dummy return
    void** frame = (void**)sp;
    for (int i = 0; i < 6; ++i) frame[i] = nullptr;
    frame[6] = (void*)&trampoline;
    frame[7] = nullptr;
    b.fib[t].sp = (void*)sp;
  }
  Block* saved = g_blk;
  g_blk = &b;
  blockIdx = block_idx;
  gridDim = grid_dim;
  blockDim = Dim3{(unsigned)nthreads, 1, 1};
  std::vector<int> order(nthreads);
  int remaining = nthreads;
  while (remaining > 0) {
    for (int t = 0; t < nthreads; ++t) order[t] = (order_mode == 1) ? nthreads - 1 - t : t;
    if (order_mode == 2) {
      for (int t = nthreads - 1; t > 0; --t) {
        b.rng ^= b.rng << 13; b.rng ^= b.rng >> 7; b.rng ^= b.rng << 17;
        std::swap(order[t], order[(int)(b.rng % (uint64_t)(t + 1))]);
      }
    }
    const unsigned long before = b.progress;
    remaining = 0;
    for (int k = 0; k < nthreads; ++k) {
      const int t = order[k];
      if (b.fib[t].done) continue;
      b.cur = t;
      threadIdx = Dim3{(unsigned)t, 0, 0};
      a1emu_switch(&b.sched_sp, b.fib[t].sp);
      if (!b.fib[t].done) ++remaining;
    }
    if (remaining > 0 && b.progress == before) {
      std::fprintf(stderr, "a1emu: deadlock -- %d threads wait at a collective that not all lanes reach (divergent barrier)\n", remaining);
      std::abort();
    }
  }
  if (n_collectives) *n_collectives += b.n_collectives;
  if (n_mma) *n_mma += b.n_mma;
  g_blk = saved;
}

}  // namespace a1emu
