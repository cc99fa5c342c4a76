// vv_stream.cuh -- persistent weight-stream kernel (tcgen05 / TMEM MMA fed by TMA) and its grid-wide synchronisation primitives.
#pragma once
#include "vv_kernels.cuh"

namespace vv {

struct GridBar { unsigned count; unsigned pad0[31]; unsigned gen; unsigned pad1[31]; };   // arrival counter and generation on separate 128 B lines

VV_DEVINL unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
VV_DEVINL unsigned ld_relaxed_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
VV_DEVINL void st_release_u32(unsigned* p, unsigned v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
VV_DEVINL void st_relaxed_u32(unsigned* p, unsigned v) { asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
VV_DEVINL unsigned atom_add_acqrel_u32(unsigned* p, unsigned v) {
  unsigned r;
  asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], %2;" : "=r"(r) : "l"(p), "r"(v) : "memory");
  return r;
}

// sense-free generation barrier across all CTAs of a cooperative launch.  bar.sync orders the CTA's writes before
// thread 0's gpu-scope release (cumulativity), the last arriver resets the counter and bumps the generation.
VV_DEVINL void grid_barrier(GridBar* gb, unsigned nctas) {
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned gen = ld_relaxed_u32(&gb->gen);
    const unsigned prev = atom_add_acqrel_u32(&gb->count, 1u);
    if (prev == nctas - 1) {
      st_relaxed_u32(&gb->count, 0u);
      st_release_u32(&gb->gen, gen + 1);
    } else {
      while (ld_acquire_u32(&gb->gen) == gen) { }
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(256) barrier_bench_kernel(GridBar* gb, int iters, float* sink) {
  float acc = 0.f;
  for (int i = 0; i < iters; ++i) {
    grid_barrier(gb, gridDim.x);
    acc += 1.f;
  }
  if (threadIdx.x == 0 && blockIdx.x == 0) *sink = acc;
}


}  // namespace vv
