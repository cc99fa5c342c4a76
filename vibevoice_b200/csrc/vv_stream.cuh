// vv_stream.cuh -- persistent weight-stream kernel: every M <= 8..64-row linear of the generation loop as a tcgen05 / TMEM MMA whose
// weight tiles arrive by TMA (cp.async.bulk.tensor, 128-byte swizzle) through a deep shared-memory ring that keeps running ACROSS the
// grid-wide dependencies between stages.
//
// Why this shape (profiles/r01_*): one speech frame is a chain of ~450 dependent skinny linears (28 LM layers x 4, 30 diffusion
// steps x 9, codec front).  As one kernel per stage each of them pays launch + cold pipeline + drain (~5 us) around a 1-9 us weight
// stream, and HBM idles in between: 0.33 of the roofline.  Here ONE cooperative grid (one CTA per SM) runs a whole program of stages:
//
//   warp 0   producer   walks the program's static schedule and issues the TMA loads of this CTA's weight tiles (128 rows x 64 k, bf16,
//                       16 KB) into an S-stage ring.  Weights never depend on activations, so it runs ahead of the grid barriers: while
//                       the other warps synchronise / stage activations for stage i, the ring already fills with tiles of stage i, i+1...
//   warp 1   MMA        one elected thread: per tile 4 x tcgen05.mma.cta_group::1.kind::f16 (M = 128 weight rows, N = 16..64 activation
//                       rows, K = 16), fp32 accumulators in TMEM; tcgen05.commit hands the ring slot back to the producer.
//   warps 2-5 workers   grid barrier -> prologue: activations from L2, RMSNorm / AdaLN / SwiGLU / GELU / solver update applied, split into
//                       bf16 hi + lo (x = hi + lo to 2^-17) and written as the MMA's B operand in the canonical K-major SWIZZLE_128B
//                       layout (rows [0, nB/2) = hi, [nB/2, nB) = lo, so ONE MMA yields both partial products) -> epilogue: tcgen05.ld
//                       of the accumulators, hi + lo, bias / gate / gamma scaling, fp32 atomics (red.global.add) into the output.
//
// Work split ("stream-K"): a stage with R = ceil(N/128) row tiles and KB = ceil(K/64) k-blocks has U = R*KB tile units, dealt out as
// contiguous ranges [c*U/G, (c+1)*U/G) to the G CTAs in row-tile-major order: every SM streams the same number of bytes (+-1 tile) for
// ANY shape, a CTA owns <= 3 (row tile, k range) segments, each with its own TMEM accumulator, and partial sums meet in the output through
// atomics.  That is why epilogues are linear (bias, scaling, residual) and the non-linearities (SwiGLU, GELU) live in the NEXT stage's
// prologue.
//
// Reference arithmetic: the stages are the same linears as the stand-alone kernels in vv_kernels.cuh (see the anchors there).
#pragma once
#include "vv_kernels.cuh"

namespace vv {

// ---------------------------------------------------------------------------------------------------------------------------------
// grid-wide synchronisation
// ---------------------------------------------------------------------------------------------------------------------------------
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
VV_DEVINL void red_add_release_u32(unsigned* p, unsigned v) { asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

// sense-free generation barrier across all CTAs of a cooperative launch (used by barrier_bench_kernel)
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

// ---------------------------------------------------------------------------------------------------------------------------------
// program representation (built on the host in vv_runtime.cu, read-only on the device)
// ---------------------------------------------------------------------------------------------------------------------------------
constexpr int ST_THREADS = 192;          // warp 0 producer, warp 1 MMA issuer, warps 2..5 workers
constexpr int ST_WORKERS = 128;
constexpr int ST_TILE = 16384;           // one weight tile: 128 rows x 64 bf16
constexpr int ST_MAXSEG = 8;             // (row tile, k range) segments a CTA may own in one stage
constexpr int ST_MAX_STAGES = 12;

enum SKind { SK_GEMV = 0, SK_NOP = 1 };
enum SPro { SP_NONE = 0, SP_RMSNORM = 1, SP_ADALN = 2, SP_SWIGLU = 3, SP_GELU = 4, SP_DPM = 5, SP_SILU = 6 };
enum SAlpha { SA_ONE = 0, SA_GATE = 1 /* alpha[m][n], row stride lda */, SA_GAMMA = 2 /* alpha[n] */ };

// CFG + DPM-Solver++ update of step `step` (same arithmetic as dpm_update_proj_kernel), evaluated in the prologue of the stage that
// projects the new latent (noisy_images_proj): B-operand row m = z'[m mod B].
struct SDpm {
  const float* z_in; float* z_out; const float* x0_in; float* x0_out; const float* v; const float* noise; const DpmCoef* coef;
  const float* cfg_p; const float* step_noise; float* latent_out;
  int step, B;
};

struct SOp {
  int kind;
  int sync_before;          // wait until every CTA has finished the previous stage (grid barrier) before touching activations
  int M, N, K;              // activation rows, weight rows (outputs), reduction length (K % 8 == 0)
  int nB;                   // MMA N: 16, 32 or 64 (rows [0,nB/2) = hi, [nB/2,nB) = lo)
  unsigned long long tmap;  // device address of the CUtensorMap of W [N][K] bf16 (box 64 x 128, SWIZZLE_128B)
  int pro;
  const float* x; long long ldx;         // activations, row stride in floats (SP_SWIGLU: interleaved gate/up sums, row length 2K)
  const float* pro_w; float pro_eps;     // norm weight [K] (may be null for SP_ADALN)
  const float* pro_shift; const float* pro_scale; long long pro_ld;
  const SDpm* dpm;
  float* y; long long ldy;               // y[m][n] += alpha * (acc + bias[n] if the segment starts at k = 0);  store != 0: y = ... (KB == 1 only)
  const float* bias;
  int alpha_kind; const float* alpha; long long lda;
  int store;
  float* init_dst; long long init_n;     // optional: zero-fill job (a buffer a LATER stage accumulates into), spread over the grid
};

struct SParams {
  const SOp* ops; int n_ops;
  unsigned* bar_count;      // zeroed by the host before every launch
  unsigned* diag;           // host-mapped: [0] = error code, [1..7] = where (watchdog)
  int n_stages;             // ring depth
  int b_bytes;              // bytes of the activation-operand region
};

// ---------------------------------------------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------------------------------------------
VV_DEVINL void st_die(unsigned* diag, unsigned code, unsigned a, unsigned b, unsigned c) {
  if (diag) {
    diag[1] = blockIdx.x; diag[2] = threadIdx.x; diag[3] = a; diag[4] = b; diag[5] = c;
    __threadfence_system();
    diag[0] = code;
    __threadfence_system();
  }
  __trap();
}
// bounded mbarrier wait: a lost arrival must end in a diagnosable trap, never in a hung GPU
VV_DEVINL void mbar_wait_wd(unsigned long long* bar, unsigned parity, unsigned* diag, unsigned code, unsigned a, unsigned b) {
  const unsigned addr = smem_u32(bar);
  unsigned ok = 0;
  long long t0 = 0;
  for (unsigned spins = 0;; ++spins) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    if (ok) return;
    if ((spins & 1023u) == 1023u) {
      const long long t = clock64();
      if (t0 == 0) t0 = t;
      else if (t - t0 > 6000000000ll) st_die(diag, code, a, b, parity);
    }
  }
}
VV_DEVINL void tma_load_2d(void* smem_dst, unsigned long long tmap, int c0, int c1, unsigned long long* bar, unsigned long long policy) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;"
               ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar)), "l"(policy) : "memory");
}
VV_DEVINL float ldcg1(const float* p) { return __ldcg(p); }
VV_DEVINL float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
VV_DEVINL void red_add_f32(float* p, float v) { asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory"); }
VV_DEVINL void worker_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// this CTA's unit range of a stage: units are (row tile, k-block) pairs in row-tile-major order
VV_DEVINL void st_part(const SOp& op, long long& u0, long long& u1, int& KB) {
  KB = (op.K + 63) >> 6;
  const long long U = (long long)((op.N + 127) >> 7) * KB;
  u0 = U * (long long)blockIdx.x / (long long)gridDim.x;
  u1 = U * (long long)(blockIdx.x + 1) / (long long)gridDim.x;
}

// ---------------------------------------------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(ST_THREADS, 1) stream_kernel(SParams P) {
  extern __shared__ unsigned char st_raw[];
  __shared__ unsigned long long full_bar[ST_MAX_STAGES], empty_bar[ST_MAX_STAGES];
  __shared__ unsigned long long b_ready, acc_full;
  __shared__ unsigned tmem_base_s;
  __shared__ float s_red[4][8];
  __shared__ float s_inv[64];
  __shared__ float s_z[8 * 64];
  const unsigned raw_addr = smem_u32(st_raw);
  unsigned char* sm = st_raw + ((1024u - (raw_addr & 1023u)) & 1023u);      // 1024 B aligned (swizzle atom)
  unsigned char* ring = sm;
  unsigned char* breg = sm + (size_t)P.n_stages * ST_TILE;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int NS = P.n_stages;
  const unsigned G = gridDim.x;

  if (tid == 0) {
    for (int i = 0; i < NS; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(&b_ready, 1);
    mbar_init(&acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned tmem = tmem_base_s;

  if (warp == 0) {
    // =============================== producer: weights only, never waits on activations ===============================
    if (lane == 0) {
      unsigned long long policy;
      asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
      unsigned it = 0;
      for (int oi = 0; oi < P.n_ops; ++oi) {
        const SOp& op = P.ops[oi];
        if (op.kind != SK_GEMV) continue;
        long long u0, u1; int KB;
        st_part(op, u0, u1, KB);
        const unsigned long long tmap = op.tmap;
        int rt = (int)(u0 / KB), kb = (int)(u0 % KB);
        for (long long u = u0; u < u1; ++u, ++it) {
          const unsigned slot = it % (unsigned)NS, ph = (it / (unsigned)NS) & 1u;
          mbar_wait_wd(&empty_bar[slot], ph ^ 1u, P.diag, 1u, (unsigned)oi, it);
          mbar_expect_tx(&full_bar[slot], (unsigned)ST_TILE);
          tma_load_2d(ring + (size_t)slot * ST_TILE, tmap, kb * 64, rt * 128, &full_bar[slot], policy);
          if (++kb == KB) { kb = 0; ++rt; }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    if (lane == 0) {
      unsigned it = 0, gi = 0;
      for (int oi = 0; oi < P.n_ops; ++oi) {
        const SOp& op = P.ops[oi];
        if (op.kind != SK_GEMV) continue;
        long long u0, u1; int KB;
        st_part(op, u0, u1, KB);
        if (u0 == u1) continue;
        const int nB = op.nB;
        // instruction descriptor: D = f32, A = B = bf16, both K-major, N = nB, M = 128
        const unsigned idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(nB >> 3) << 17) | ((unsigned)(128 >> 4) << 24);
        const int kb_first = (int)(u0 % KB), rt_first = (int)(u0 / KB);
        mbar_wait_wd(&b_ready, gi & 1u, P.diag, 2u, (unsigned)oi, gi);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        int rt = rt_first, kb = kb_first;
        bool fresh = true;                              // first k-block of a segment overwrites its accumulator
        for (long long u = u0; u < u1; ++u, ++it) {
          const unsigned slot = it % (unsigned)NS, ph = (it / (unsigned)NS) & 1u;
          mbar_wait_wd(&full_bar[slot], ph, P.diag, 3u, (unsigned)oi, it);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          int jloc = kb - kb_first; if (jloc < 0) jloc += KB;
          const unsigned long long da = umma_desc_sw128(smem_u32(ring + (size_t)slot * ST_TILE));
          const unsigned long long db = umma_desc_sw128(smem_u32(breg + (size_t)jloc * (size_t)(nB * 128)));
          const unsigned dcol = tmem + (unsigned)((rt - rt_first) * nB);
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) tc5_mma(dcol, da + 2 * k4, db + 2 * k4, idesc, (fresh && k4 == 0) ? 0u : 1u);
          tc5_commit(&empty_bar[slot]);                 // slot is free once these MMAs have read it
          fresh = false;
          if (++kb == KB) { kb = 0; ++rt; fresh = true; }
        }
        tc5_commit(&acc_full);                          // accumulators of this stage are complete
        ++gi;
      }
    }
  } else {
    // =============================== workers: barrier, prologue (B operand), epilogue ===============================
    const int wt = tid - 64;                  // 0..127
    const int wq = warp & 3;                  // TMEM lane quadrant this warp may read
    const int ww = warp - 2;
    unsigned gi = 0, bar_target = 0;
    for (int oi = 0; oi < P.n_ops; ++oi) {
      const SOp& op = P.ops[oi];
      if (op.sync_before) {
        bar_target += G;
        worker_sync();                         // every worker's global writes of the previous stage are issued ...
        if (wt == 0) {
          __threadfence();                     // ... and ordered before this CTA's arrival (cumulativity through bar.sync)
          red_add_release_u32(P.bar_count, 1u);
          long long t0 = 0;
          for (unsigned spins = 0; ld_acquire_u32(P.bar_count) < bar_target; ++spins) {
            if ((spins & 255u) == 255u) {
              const long long t = clock64();
              if (t0 == 0) t0 = t;
              else if (t - t0 > 6000000000ll) st_die(P.diag, 4u, (unsigned)oi, bar_target, ld_acquire_u32(P.bar_count));
            }
          }
        }
        worker_sync();
      }
      if (op.init_dst) {
        for (long long i = (long long)blockIdx.x * ST_WORKERS + wt; i < op.init_n; i += (long long)G * ST_WORKERS) op.init_dst[i] = 0.f;
      }
      if (op.kind != SK_GEMV) continue;
      long long u0, u1; int KB;
      st_part(op, u0, u1, KB);
      if (u0 == u1) continue;
      const int M = op.M, K = op.K, N = op.N, nB = op.nB, half = nB >> 1;
      const int units = (int)(u1 - u0);
      const int count = units < KB ? units : KB;          // k-blocks of activations this CTA needs (contiguous mod KB from kb_first)
      const int kb_first = (int)(u0 % KB), rt_first = (int)(u0 / KB);
      const int pro = op.pro;
      // ---------------- row statistics (full rows) ----------------
      if (pro == SP_RMSNORM || pro == SP_ADALN) {
        for (int m0 = 0; m0 < M; m0 += 8) {
          float ss[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) ss[j] = 0.f;
          for (int q = wt; q < (K >> 2); q += ST_WORKERS) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              if (m0 + j < M) {
                const float4 v = ldcg4(op.x + (long long)(m0 + j) * op.ldx + 4 * q);
                ss[j] += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
              }
            }
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) { ss[j] = warp_sum(ss[j]); if (lane == 0) s_red[ww][j] = ss[j]; }
          worker_sync();
          if (wt < 8 && m0 + wt < M) s_inv[m0 + wt] = rsqrtf((s_red[0][wt] + s_red[1][wt] + s_red[2][wt] + s_red[3][wt]) / (float)K + op.pro_eps);
          worker_sync();
        }
      } else if (pro == SP_DPM) {
        // z' for every sample (same arithmetic as dpm_update_proj_kernel); CTA 0... every CTA recomputes, the first unit owner publishes
        const SDpm d = *op.dpm;
        for (int i = wt; i < d.B * 64; i += ST_WORKERS) {
          const int b = i >> 6, e = i & 63;
          float zn, x0 = 0.f;
          if (d.step < 0) {
            zn = ldcg1(d.noise + i);
          } else {
            const DpmCoef c = d.coef[d.step];
            const float cfg = *d.cfg_p;
            const float vc = ldcg1(d.v + (size_t)b * 64 + e), vu = ldcg1(d.v + (size_t)(d.B + b) * 64 + e);
            const float vv_ = vu + cfg * (vc - vu);
            const float zo = ldcg1(d.z_in + i);
            x0 = c.a0 * zo - c.s0 * vv_;
            zn = c.ks * zo - c.kx * x0;
            if (c.order == 2) zn -= 0.5f * c.kx * (c.rinv * (x0 - ldcg1(d.x0_in + i)));
            if (d.step_noise) zn += c.kn * ldcg1(d.step_noise + ((size_t)d.step * d.B + b) * 64 + e);
          }
          s_z[i] = zn;
          if (u0 == 0) {                       // exactly one CTA owns unit 0 of the stage
            d.z_out[i] = zn; d.x0_out[i] = x0;
            if (d.latent_out) d.latent_out[i] = zn;
          }
        }
        worker_sync();
      }
      // ---------------- B operand: 16-byte chunks of 8 consecutive k for one activation row ----------------
      {
        const int total = M * count * 8;
        for (int c = wt; c < total; c += ST_WORKERS) {
          const int m = c / (count * 8), r = c - m * (count * 8);
          const int jloc = r >> 3, ch = r & 7;
          int kb = kb_first + jloc; if (kb >= KB) kb -= KB;
          const int k = kb * 64 + ch * 8;
          float v[8];
          if (k < K) {
            if (pro == SP_SWIGLU) {
              const float* xr = op.x + (long long)m * op.ldx + 2 * k;
              const float4 a0 = ldcg4(xr), a1 = ldcg4(xr + 4), a2 = ldcg4(xr + 8), a3 = ldcg4(xr + 12);
              v[0] = silu_f(a0.x) * a0.y; v[1] = silu_f(a0.z) * a0.w; v[2] = silu_f(a1.x) * a1.y; v[3] = silu_f(a1.z) * a1.w;
              v[4] = silu_f(a2.x) * a2.y; v[5] = silu_f(a2.z) * a2.w; v[6] = silu_f(a3.x) * a3.y; v[7] = silu_f(a3.z) * a3.w;
            } else if (pro == SP_DPM) {
              const float* zr = s_z + (m % op.dpm->B) * 64 + k;
#pragma unroll
              for (int j = 0; j < 8; ++j) v[j] = zr[j];
            } else {
              const float* xr = op.x + (long long)m * op.ldx + k;
              const float4 a0 = ldcg4(xr), a1 = ldcg4(xr + 4);
              v[0] = a0.x; v[1] = a0.y; v[2] = a0.z; v[3] = a0.w; v[4] = a1.x; v[5] = a1.y; v[6] = a1.z; v[7] = a1.w;
              if (pro == SP_RMSNORM) {
                const float inv = s_inv[m];
                const float4 w0 = *reinterpret_cast<const float4*>(op.pro_w + k), w1 = *reinterpret_cast<const float4*>(op.pro_w + k + 4);
                v[0] *= inv * w0.x; v[1] *= inv * w0.y; v[2] *= inv * w0.z; v[3] *= inv * w0.w;
                v[4] *= inv * w1.x; v[5] *= inv * w1.y; v[6] *= inv * w1.z; v[7] *= inv * w1.w;
              } else if (pro == SP_ADALN) {
                const float inv = s_inv[m];
                float w[8];
                if (op.pro_w) {
                  const float4 w0 = *reinterpret_cast<const float4*>(op.pro_w + k), w1 = *reinterpret_cast<const float4*>(op.pro_w + k + 4);
                  w[0] = w0.x; w[1] = w0.y; w[2] = w0.z; w[3] = w0.w; w[4] = w1.x; w[5] = w1.y; w[6] = w1.z; w[7] = w1.w;
                } else {
#pragma unroll
                  for (int j = 0; j < 8; ++j) w[j] = 1.f;
                }
                const long long o = (long long)m * op.pro_ld + k;
                const float4 s0 = ldcg4(op.pro_scale + o), s1 = ldcg4(op.pro_scale + o + 4);
                const float4 h0 = ldcg4(op.pro_shift + o), h1 = ldcg4(op.pro_shift + o + 4);
                const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
                const float sh[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = v[j] * inv * w[j] * (1.f + sc[j]) + sh[j];
              } else if (pro == SP_GELU) {
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = gelu_erf_f(v[j]);
              } else if (pro == SP_SILU) {
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = silu_f(v[j]);
              }
            }
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = 0.f;        // k >= K: the weight tile is zero-filled there, keep 0 * x finite
          }
          float h[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) h[j] = __bfloat162float(__float2bfloat16_rn(v[j]));
          const uint4 hv = make_uint4(pack_bf16(h[0], h[1]), pack_bf16(h[2], h[3]), pack_bf16(h[4], h[5]), pack_bf16(h[6], h[7]));
          const uint4 lv = make_uint4(pack_bf16(v[0] - h[0], v[1] - h[1]), pack_bf16(v[2] - h[2], v[3] - h[3]),
                                      pack_bf16(v[4] - h[4], v[5] - h[5]), pack_bf16(v[6] - h[6], v[7] - h[7]));
          unsigned char* blk = breg + (size_t)jloc * (size_t)(nB * 128);
          const int rl = half + m;
          *reinterpret_cast<uint4*>(blk + (m >> 3) * 1024 + (m & 7) * 128 + ((ch ^ (m & 7)) << 4)) = hv;
          *reinterpret_cast<uint4*>(blk + (rl >> 3) * 1024 + (rl & 7) * 128 + ((ch ^ (rl & 7)) << 4)) = lv;
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // generic-proxy smem writes -> visible to the tensor core
      worker_sync();
      if (wt == 0) mbar_arrive(&b_ready);
      // ---------------- epilogue ----------------
      mbar_wait_wd(&acc_full, gi & 1u, P.diag, 5u, (unsigned)oi, gi);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int rt_last = (int)((u1 - 1) / KB);
      for (int rt = rt_first; rt <= rt_last; ++rt) {
        const bool from0 = (rt > rt_first) || kb_first == 0;            // this segment holds k-block 0 of its row tile -> it adds the bias
        const int n = rt * 128 + wq * 32 + lane;
        const unsigned tcol = tmem + ((unsigned)(wq * 32) << 16) + (unsigned)((rt - rt_first) * nB);
        for (int m0 = 0; m0 < half; m0 += 8) {
          unsigned rh[8], rl[8];
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                       : "=r"(rh[0]), "=r"(rh[1]), "=r"(rh[2]), "=r"(rh[3]), "=r"(rh[4]), "=r"(rh[5]), "=r"(rh[6]), "=r"(rh[7]) : "r"(tcol + (unsigned)m0));
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                       : "=r"(rl[0]), "=r"(rl[1]), "=r"(rl[2]), "=r"(rl[3]), "=r"(rl[4]), "=r"(rl[5]), "=r"(rl[6]), "=r"(rl[7]) : "r"(tcol + (unsigned)(half + m0)));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (n < N) {
            const float bias = (from0 && op.bias) ? op.bias[n] : 0.f;
            const float gam = (op.alpha_kind == SA_GAMMA) ? op.alpha[n] : 1.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int m = m0 + j;
              if (m < M) {
                float val = __uint_as_float(rh[j]) + __uint_as_float(rl[j]) + bias;
                if (op.alpha_kind == SA_GATE) val *= ldcg1(op.alpha + (long long)m * op.lda + n);
                else val *= gam;
                float* yp = op.y + (long long)m * op.ldy + n;
                if (op.store) *yp = val; else red_add_f32(yp, val);
              }
            }
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      ++gi;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

}  // namespace vv
