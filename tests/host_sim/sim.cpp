// tests/host_sim/sim.cpp -- TEST ONLY.  Host functional simulation of libfp8fq.so's kernels: this translation unit is
// fp8_quantization_b200/csrc/fp8fq_kernels.cu itself, compiled by g++ against tests/host_sim/shim/cuda_runtime.h with
// -DFP8FQ_HOST_SIM, plus the cooperative scheduler that runs a grid on one OS thread.  Built by oracle/Makefile into
// oracle/_build/libfp8fq_sim.so and loaded by tests/test_host_sim.py only -- "device" pointers are host pointers.  The
// product (fp8_quantization_b200) never loads it: there is no CPU path in the product.
#include "shim/cuda_runtime.h"

namespace fp8fq_sim {
namespace {

constexpr size_t kStackBytes = 64 * 1024;
constexpr size_t kDynSmemFloats = 64 * 1024;  // 256 KB >= any dynamic shared memory request of the library

enum State { RUNNABLE, AT_CTA_BARRIER, AT_WARP_BARRIER, DONE };

// Context switch.  glibc's swapcontext saves and restores the signal mask with two system calls per switch, which
// dominates the run time of the barrier-heavy kernels; on x86-64 a fiber is therefore just a saved stack pointer and the
// switch saves / restores the callee-saved registers (System V ABI: rbx, rbp, r12-r15, plus the SSE / x87 control words).
// Other architectures use ucontext.
#if defined(__x86_64__) && !defined(FP8FQ_SIM_UCONTEXT)
#define FP8FQ_SIM_ASM_SWITCH 1
extern "C" void fp8fq_sim_switch(void** save_sp, void* load_sp);
asm(R"(
    .text
    .hidden fp8fq_sim_switch
    .globl fp8fq_sim_switch
    .type fp8fq_sim_switch,@function
fp8fq_sim_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    subq $8, %rsp
    stmxcsr (%rsp)
    fnstcw 4(%rsp)
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    ldmxcsr (%rsp)
    fldcw 4(%rsp)
    addq $8, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
    .size fp8fq_sim_switch,.-fp8fq_sim_switch
)");
struct Fiber {
  void* sp;
  State st;
};
void* g_sched_sp = nullptr;
#else
struct Fiber {
  ucontext_t ctx;
  State st;
};
ucontext_t g_sched;
#endif

alignas(16) float g_dyn_smem[kDynSmemFloats];
std::vector<Fiber> g_fibers;
std::vector<char> g_stacks;
std::vector<float> g_warp_slots;  // [warps][32]
const std::function<void()>* g_body = nullptr;
int g_cur = -1;
int64_t g_launches = 0, g_ctas = 0;

#ifdef FP8FQ_SIM_ASM_SWITCH
void yield_to_scheduler() { fp8fq_sim_switch(&g_fibers[g_cur].sp, g_sched_sp); }

// first activation of a fiber "returns" here; it never returns itself (there is no caller frame): after the body it
// marks the fiber done and switches to the scheduler for good
__attribute__((no_sanitize_address)) void trampoline() {
  (*g_body)();
  g_fibers[g_cur].st = DONE;
  yield_to_scheduler();
  abort();  // a finished fiber is never resumed
}

void init_fiber(Fiber& f, char* stack, size_t bytes) {
  // stack image popped by fp8fq_sim_switch: [mxcsr | x87 cw][r15][r14][r13][r12][rbx][rbp][return address]
  uintptr_t top = (reinterpret_cast<uintptr_t>(stack) + bytes) & ~uintptr_t(15);
  uint64_t* sp = reinterpret_cast<uint64_t*>(top) - 1;  // slot the ABI expects above the return address (alignment)
  *sp = 0;
  *--sp = reinterpret_cast<uint64_t>(&trampoline);       // after `ret`: rsp = top - 8, i.e. 8 mod 16 as at a call
  for (int i = 0; i < 6; ++i) *--sp = 0;                 // rbp, rbx, r12-r15
  uint32_t csr[2];
  asm volatile("stmxcsr %0" : "=m"(csr[0]));
  uint16_t cw;
  asm volatile("fnstcw %0" : "=m"(cw));
  csr[1] = cw;
  --sp;
  memcpy(sp, csr, 8);
  f.sp = sp;
  f.st = RUNNABLE;
}

void resume(Fiber& f) { fp8fq_sim_switch(&g_sched_sp, f.sp); }
#else
void yield_to_scheduler() { swapcontext(&g_fibers[g_cur].ctx, &g_sched); }

void trampoline() {
  (*g_body)();
  g_fibers[g_cur].st = DONE;
  // returning resumes uc_link = the scheduler
}

void init_fiber(Fiber& f, char* stack, size_t bytes) {
  getcontext(&f.ctx);
  f.ctx.uc_stack.ss_sp = stack;
  f.ctx.uc_stack.ss_size = bytes;
  f.ctx.uc_link = &g_sched;
  makecontext(&f.ctx, trampoline, 0);
  f.st = RUNNABLE;
}

void resume(Fiber& f) { swapcontext(&g_sched, &f.ctx); }
#endif

void warp_barrier() {
  g_fibers[g_cur].st = AT_WARP_BARRIER;
  yield_to_scheduler();
}

void run_cta(unsigned nthreads) {
  const unsigned nwarps = (nthreads + 31) / 32;
  g_fibers.resize(nthreads);
  if (g_stacks.size() < (size_t)nthreads * kStackBytes) g_stacks.resize((size_t)nthreads * kStackBytes);
  g_warp_slots.assign((size_t)nwarps * 32, 0.0f);
  for (unsigned t = 0; t < nthreads; ++t) init_fiber(g_fibers[t], g_stacks.data() + (size_t)t * kStackBytes, kStackBytes);
  unsigned alive = nthreads;
  while (alive > 0) {
    bool progressed = false;
    for (unsigned t = 0; t < nthreads; ++t) {
      if (g_fibers[t].st != RUNNABLE) continue;
      g_cur = (int)t;
      threadIdx = uint3{t % blockDim.x, (t / blockDim.x) % blockDim.y, t / (blockDim.x * blockDim.y)};
      resume(g_fibers[t]);
      progressed = true;
      if (g_fibers[t].st == DONE) --alive;
    }
    // release the barriers every live participant has reached (exited threads do not take part, as on the GPU)
    unsigned at_cta = 0, live = 0;
    for (unsigned t = 0; t < nthreads; ++t) {
      if (g_fibers[t].st == AT_CTA_BARRIER) ++at_cta;
      if (g_fibers[t].st != DONE) ++live;
    }
    bool released = false;
    if (live > 0 && at_cta == live) {
      for (unsigned t = 0; t < nthreads; ++t)
        if (g_fibers[t].st == AT_CTA_BARRIER) g_fibers[t].st = RUNNABLE;
      released = true;
    }
    for (unsigned w = 0; w < nwarps; ++w) {
      unsigned at_w = 0, live_w = 0;
      const unsigned t0 = w * 32, t1 = t0 + 32 < nthreads ? t0 + 32 : nthreads;
      for (unsigned t = t0; t < t1; ++t) {
        if (g_fibers[t].st == AT_WARP_BARRIER) ++at_w;
        if (g_fibers[t].st != DONE) ++live_w;
      }
      if (live_w > 0 && at_w == live_w) {
        for (unsigned t = t0; t < t1; ++t)
          if (g_fibers[t].st == AT_WARP_BARRIER) g_fibers[t].st = RUNNABLE;
        released = true;
      }
    }
    if (!progressed && !released && alive > 0) {
      fprintf(stderr, "fp8fq_sim: deadlock in block (%u,%u): %u threads alive, %u at __syncthreads\n", blockIdx.x,
              blockIdx.y, alive, at_cta);
      abort();
    }
  }
  g_cur = -1;
}

}  // namespace

float* dynamic_smem() { return g_dyn_smem; }

void not_cooperative(const char* what) {
  fprintf(stderr, "fp8fq_sim: %s reached in a kernel that was launched as barrier-free (launch_impl's `pdl` kernels); "
                  "launch it cooperatively\n", what);
  abort();
}

void cta_barrier() {
  if (g_cur < 0) not_cooperative("__syncthreads");
  g_fibers[g_cur].st = AT_CTA_BARRIER;
  yield_to_scheduler();
}

float warp_exchange(float v, int lane_xor) {
  if (g_cur < 0) not_cooperative("__shfl_xor_sync");
  const int warp = g_cur >> 5, lane = g_cur & 31;
  float* slots = g_warp_slots.data() + (size_t)warp * 32;
  slots[lane] = v;
  warp_barrier();                                    // every live lane has published its value
  const int src = lane ^ lane_xor;
  const int src_thread = (warp << 5) + src;
  // a lane outside the block (partial last warp) or already exited: the GPU returns the caller's own value
  const bool src_live = src_thread < (int)g_fibers.size() && g_fibers[src_thread].st != DONE;
  const float r = src_live ? slots[src] : v;
  warp_barrier();                                    // every live lane has read before the slots are reused
  return r;
}

void run_grid(dim3 grid, dim3 block, size_t smem, bool cooperative, const std::function<void()>& thread_body) {
  if (g_cur != -1) {
    fprintf(stderr, "fp8fq_sim: nested launch\n");
    abort();
  }
  if (smem > sizeof(g_dyn_smem)) {
    fprintf(stderr, "fp8fq_sim: %zu bytes of dynamic shared memory requested\n", smem);
    abort();
  }
  const unsigned nthreads = block.x * block.y * block.z;
  if (nthreads == 0 || nthreads > 1024 || grid.x == 0 || grid.y == 0 || grid.z == 0) {
    fprintf(stderr, "fp8fq_sim: invalid launch configuration\n");
    abort();
  }
  ++g_launches;
  g_body = &thread_body;
  gridDim = grid;
  blockDim = block;
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        blockIdx = uint3{bx, by, bz};
        // shared memory is uninitialised on the GPU: poison the dynamic part so that a read-before-write shows up
        if (smem) memset(g_dyn_smem, 0xff, smem);
        if (cooperative) {
          run_cta(nthreads);
        } else {
          g_cur = -2;   // plain mode: a barrier aborts
          for (unsigned t = 0; t < nthreads; ++t) {
            threadIdx = uint3{t % blockDim.x, (t / blockDim.x) % blockDim.y, t / (blockDim.x * blockDim.y)};
            thread_body();
          }
          g_cur = -1;
        }
        ++g_ctas;
      }
  g_body = nullptr;
}

int64_t launches() { return g_launches; }
int64_t ctas_run() { return g_ctas; }

}  // namespace fp8fq_sim

// the library itself -- the same source nvcc compiles
#include "../../fp8_quantization_b200/csrc/fp8fq_kernels.cu"

extern "C" {
// test hooks of the simulation (not part of include/fp8fq.h)
void fp8fq_sim_set_sm_count(int n) {
  fp8fq_sim::g_sm_count = n;
  for (auto& c : g_sms) c.store(0);   // the library caches the attribute per device: make it read again
}
void fp8fq_sim_report_pinned(int yes) { fp8fq_sim::g_report_pinned = yes != 0; }
int64_t fp8fq_sim_launches(void) { return fp8fq_sim::launches(); }
int64_t fp8fq_sim_ctas(void) { return fp8fq_sim::ctas_run(); }
}
