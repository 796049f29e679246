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

struct Fiber {
  ucontext_t ctx;
  State st;
};

alignas(16) float g_dyn_smem[kDynSmemFloats];
std::vector<Fiber> g_fibers;
std::vector<char> g_stacks;
std::vector<float> g_warp_slots;  // [warps][32]
ucontext_t g_sched;
const std::function<void()>* g_body = nullptr;
int g_cur = -1;
int64_t g_launches = 0, g_ctas = 0;

void yield_to_scheduler() { swapcontext(&g_fibers[g_cur].ctx, &g_sched); }

void trampoline() {
  (*g_body)();
  g_fibers[g_cur].st = DONE;
  // returning resumes uc_link = the scheduler
}

void warp_barrier() {
  g_fibers[g_cur].st = AT_WARP_BARRIER;
  yield_to_scheduler();
}

void run_cta(unsigned nthreads) {
  const unsigned nwarps = (nthreads + 31) / 32;
  g_fibers.resize(nthreads);
  if (g_stacks.size() < (size_t)nthreads * kStackBytes) g_stacks.resize((size_t)nthreads * kStackBytes);
  g_warp_slots.assign((size_t)nwarps * 32, 0.0f);
  for (unsigned t = 0; t < nthreads; ++t) {
    Fiber& f = g_fibers[t];
    getcontext(&f.ctx);
    f.ctx.uc_stack.ss_sp = g_stacks.data() + (size_t)t * kStackBytes;
    f.ctx.uc_stack.ss_size = kStackBytes;
    f.ctx.uc_link = &g_sched;
    makecontext(&f.ctx, trampoline, 0);
    f.st = RUNNABLE;
  }
  unsigned alive = nthreads;
  while (alive > 0) {
    bool progressed = false;
    for (unsigned t = 0; t < nthreads; ++t) {
      if (g_fibers[t].st != RUNNABLE) continue;
      g_cur = (int)t;
      threadIdx = uint3{t % blockDim.x, (t / blockDim.x) % blockDim.y, t / (blockDim.x * blockDim.y)};
      swapcontext(&g_sched, &g_fibers[t].ctx);
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

void cta_barrier() {
  g_fibers[g_cur].st = AT_CTA_BARRIER;
  yield_to_scheduler();
}

float warp_exchange(float v, int lane_xor) {
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

void run_grid(dim3 grid, dim3 block, size_t smem, const std::function<void()>& thread_body) {
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
        run_cta(nthreads);
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
void fp8fq_sim_set_sm_count(int n) { fp8fq_sim::g_sm_count = n; }
void fp8fq_sim_report_pinned(int yes) { fp8fq_sim::g_report_pinned = yes != 0; }
int64_t fp8fq_sim_launches(void) { return fp8fq_sim::launches(); }
int64_t fp8fq_sim_ctas(void) { return fp8fq_sim::ctas_run(); }
}
