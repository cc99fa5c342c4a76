// vv_mega.cuh -- persistent "program" kernel: one cooperative launch executes a whole dependent chain of
// stages (all 28 LLM layers, or all N diffusion steps) with device-side grid barriers between stages instead
// of kernel boundaries.  Rationale (profiles/r01_*): at M = 2 rows every stage is a ~5-20 us weight stream; as
// separate kernels each one pays launch + pipeline-drain + cold activation round trips (~6 us), 800 times a frame.
// Inside one resident grid the barrier costs ~1 us, and each CTA requests the first weights of the next stage
// BEFORE it waits on the barrier, so HBM keeps streaming across the dependency.
//
// Stage kinds mirror the stand-alone kernels in vv_kernels.cuh (same math, same work split, grid-strided):
//   OP_GEMV     gemv_kernel            OP_ATTN    rope_append + attn_partial (fused; the split that owns the
//   OP_COMBINE  attn_combine                      newest token rotates/stores K,V itself)
//   OP_FINAL    final RMSNorm + lm_head/argmax    OP_DPM  dpm_update_proj
#pragma once
#include "vv_kernels.cuh"

namespace vv {

enum OpKind { OP_GEMV = 0, OP_ATTN = 1, OP_COMBINE = 2, OP_FINAL = 3, OP_DPM = 4 };

struct AttnOp {
  const float* qkv;      // [M, (nq + 2 nkv) * HD] fp32, bias added, NOT yet rotated
  KvView kv;
  float* part_acc; float* part_ml;
  float* attn_out;       // [M, nq*HD] (combine)
  const float* inv_freq;
  int nsplit, M;
  float scale;
};
struct FinalOp {
  const float* h; const float* norm_w; float* hidden; const bf16* w_valid; const int* valid_ids; float* logits; int* tokens;
  int M, B, H, n_valid; float eps;
};
struct DpmOp {
  const float* z_in; float* z_out; const float* x0_in; float* x0_out; const float* v; const float* noise; const DpmCoef* coef;
  const bf16* w_noisy; float* xout; float* latent_out;
  int step, B, H, do_proj; float cfg;
};
struct Op {
  int kind;
  int barrier_before;
  GemvP g;
  AttnOp a;
  FinalOp f;
  DpmOp d;
};

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

VV_DEVINL float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }

VV_DEVINL void epi_store_cg(const GemvP& p, int m, int n, float v) {
  switch (p.epi) {
    case EPI_RESID: v += __ldcg(p.res + (long long)m * p.ldres + n); break;
    case EPI_GATED_RESID: v = __ldcg(p.res + (long long)m * p.ldres + n) + __ldcg(p.epi_a + (long long)m * p.epi_lda + n) * v; break;
    case EPI_GAMMA_RESID: v = __ldcg(p.res + (long long)m * p.ldres + n) + p.epi_a[n] * v; break;
    case EPI_GELU: v = gelu_erf_f(v); break;
    case EPI_SILU: v = silu_f(v); break;
    default: break;
  }
  p.y[(long long)m * p.ldy + n] = v;
}

// ---------------------------------------------------------------------------------------------------------------
// GEMV stage (same algorithm as gemv_kernel; activations are read with ld.global.cg because other SMs wrote them
// during this launch and L1 is not coherent).
// ---------------------------------------------------------------------------------------------------------------
template <int MB>
VV_DEVINL void gemv_stage(const GemvP& p, float* smem_f, GridBar* gb, bool barrier_before, float* s_inv, float* s_part) {
  const int K = p.K, N = p.N;
  const int Kp = (K + 255) & ~255;
  float* xs = smem_f;
  float* red = smem_f + MB * Kp;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int WK = p.WK, WR = 8 / WK;
  const int wr = warp / WK, wk = warp % WK;
  const int ntasks = (N + 4 * WR - 1) / (4 * WR);
  const int nchunks = (K + 255) >> 8;
  const int klane = lane * 8;

  auto load_chunk = [&](uint4 (&wv)[4], const bf16* const (&wrow)[4], int c) {
    if ((c << 8) + klane < K) {
#pragma unroll
      for (int r = 0; r < 4; ++r) wv[r] = ldg_stream(wrow[r] + (c << 8));
    } else {
#pragma unroll
      for (int r = 0; r < 4; ++r) wv[r] = make_uint4(0u, 0u, 0u, 0u);
    }
  };
  auto set_rows = [&](const bf16* (&wrow)[4], int task) {
    const int r0 = (task * WR + wr) * 4;
#pragma unroll
    for (int r = 0; r < 4; ++r) wrow[r] = p.W + (size_t)min(r0 + r, N - 1) * K + klane;
  };

  for (int m0 = 0; m0 < p.M; m0 += MB) {
    int task = blockIdx.x;
    const bf16* wrow[4];
    uint4 cur[4], nxt[4];
    int c = wk;
    if (task < ntasks) {
      set_rows(wrow, task);
      if (c < nchunks) load_chunk(cur, wrow, c);          // weights of this stage are requested before the barrier
    }
    if (m0 == 0 && barrier_before) grid_barrier(gb, gridDim.x); else __syncthreads();
    const bool need_inv = (p.pro == PRO_RMSNORM || p.pro == PRO_ADALN);
    const int K4 = K >> 2;
    if (need_inv) {
      for (int m = 0; m < MB; ++m) {
        float ss = 0.f;
        if (m0 + m < p.M) {
          const float* xr = p.x + p.xmap.off(m0 + m);
          for (int q = tid; q < K4; q += 256) { const float4 v = ldcg4(xr + 4 * q); ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w; }
        }
        ss = warp_sum(ss);
        if (lane == 0) s_part[warp] = ss;
        __syncthreads();
        if (tid == 0) {
          float t = 0.f;
          for (int i = 0; i < 8; ++i) t += s_part[i];
          s_inv[m] = rsqrtf(t / (float)K + p.pro_eps);
        }
        __syncthreads();
      }
    }
    for (int m = 0; m < MB; ++m) {
      const bool valid = (m0 + m < p.M);
      const float* xr = p.x + (valid ? p.xmap.off(m0 + m) : 0);
      const float inv = need_inv ? s_inv[m] : 1.f;
      for (int q = tid; q < (Kp >> 2); q += 256) {
        const int k = q << 2;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid && k < K) {
          v = ldcg4(xr + k);
          if (p.pro == PRO_RMSNORM) {
            const float4 w = *reinterpret_cast<const float4*>(p.pro_w + k);
            v.x *= inv * w.x; v.y *= inv * w.y; v.z *= inv * w.z; v.w *= inv * w.w;
          } else if (p.pro == PRO_ADALN) {
            float4 w = make_float4(1.f, 1.f, 1.f, 1.f);
            if (p.pro_w) w = *reinterpret_cast<const float4*>(p.pro_w + k);
            const long long o = (long long)(m0 + m) * p.pro_ld + k;
            const float4 sc = ldcg4(p.pro_scale + o);
            const float4 sh = ldcg4(p.pro_shift + o);
            v.x = v.x * inv * w.x * (1.f + sc.x) + sh.x; v.y = v.y * inv * w.y * (1.f + sc.y) + sh.y;
            v.z = v.z * inv * w.z * (1.f + sc.z) + sh.z; v.w = v.w * inv * w.w * (1.f + sc.w) + sh.w;
          } else if (p.pro == PRO_SILU) {
            v.x = silu_f(v.x); v.y = silu_f(v.y); v.z = silu_f(v.z); v.w = silu_f(v.w);
          }
        }
        *reinterpret_cast<float4*>(xs + m * Kp + xs_pos(k)) = v;
      }
    }
    __syncthreads();

    int parity = 0;
    while (task < ntasks) {
      float acc[4][MB];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int m = 0; m < MB; ++m) acc[r][m] = 0.f;
      while (c < nchunks) {
        const int cn = c + WK;
        if (cn < nchunks) load_chunk(nxt, wrow, cn);
        float xv[MB][8];
#pragma unroll
        for (int m = 0; m < MB; ++m) {
          const float4 a = *reinterpret_cast<const float4*>(xs + m * Kp + (c << 8) + (lane << 2));
          const float4 b = *reinterpret_cast<const float4*>(xs + m * Kp + (c << 8) + 128 + (lane << 2));
          xv[m][0] = a.x; xv[m][1] = a.y; xv[m][2] = a.z; xv[m][3] = a.w;
          xv[m][4] = b.x; xv[m][5] = b.y; xv[m][6] = b.z; xv[m][7] = b.w;
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          float wf[8];
          bf16x8_unpack(cur[r], wf);
#pragma unroll
          for (int m = 0; m < MB; ++m)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[r][m] = fmaf(wf[j], xv[m][j], acc[r][m]);
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) cur[r] = nxt[r];
        c = cn;
      }
      const int this_task = task;
      task += gridDim.x;
      c = wk;
      if (task < ntasks) {
        set_rows(wrow, task);
        if (c < nchunks) load_chunk(cur, wrow, c);
      }
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int m = 0; m < MB; ++m) acc[r][m] = warp_sum(acc[r][m]);
      float* rbuf = red + parity * (8 * 4 * MB);
      if (lane == 0) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int m = 0; m < MB; ++m) rbuf[warp * (4 * MB) + r * MB + m] = acc[r][m];
      }
      __syncthreads();
      if (p.epi == EPI_SWIGLU) {
        if (tid < WR * 2 * MB) {
          const int q = tid / (2 * MB), pr = (tid / MB) % 2, m = tid % MB;
          const int n0 = (this_task * WR + q) * 4 + pr * 2;
          if (n0 + 1 < N && m0 + m < p.M) {
            float g = 0.f, u = 0.f;
            for (int s = 0; s < WK; ++s) {
              g += rbuf[(q * WK + s) * (4 * MB) + (pr * 2) * MB + m];
              u += rbuf[(q * WK + s) * (4 * MB) + (pr * 2 + 1) * MB + m];
            }
            if (p.bias) { g += p.bias[n0]; u += p.bias[n0 + 1]; }
            p.y[(long long)(m0 + m) * p.ldy + (n0 >> 1)] = silu_f(g) * u;
          }
        }
      } else {
        if (tid < WR * 4 * MB) {
          const int q = tid / (4 * MB), r = (tid / MB) % 4, m = tid % MB;
          const int n = (this_task * WR + q) * 4 + r;
          if (n < N && m0 + m < p.M) {
            float v = 0.f;
            for (int s = 0; s < WK; ++s) v += rbuf[(q * WK + s) * (4 * MB) + r * MB + m];
            if (p.bias) v += p.bias[n];
            epi_store_cg(p, m0 + m, n, v);
          }
        }
      }
      parity ^= 1;
    }
    __syncthreads();   // xs / red are re-staged by the next M block or the next stage
  }
}

// ---------------------------------------------------------------------------------------------------------------
// attention stage: a CTA is two independent 128-thread groups; each group owns work items (split, kv head, row).
// RoPE of q (and of the new k) is applied here; the split that contains the newest token stores its K/V.
// ---------------------------------------------------------------------------------------------------------------
constexpr int ATT_GROUP_SMEM = 2 * ATT_TILE * (HD + 8) * 2 * 2 + ATT_MAXG * HD * 4 + ATT_MAXG * ATT_TILE * 4 + 2 * HD * 2;   // Ks,Vs,qs,ps,knew/vnew

VV_DEVINL void group_sync(int id) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }

VV_DEVINL void attn_stage(const AttnOp& a, unsigned char* smem_g, int gtid, int group, int ngroups, int bar_id) {
  typedef bf16 (*TileP)[ATT_TILE][HD + 8];
  TileP Ks = reinterpret_cast<TileP>(smem_g);
  TileP Vs = reinterpret_cast<TileP>(smem_g + 2 * ATT_TILE * (HD + 8) * 2);
  float (*qs)[HD] = reinterpret_cast<float (*)[HD]>(smem_g + 4 * ATT_TILE * (HD + 8) * 2);
  float (*ps)[ATT_TILE] = reinterpret_cast<float (*)[ATT_TILE]>(smem_g + 4 * ATT_TILE * (HD + 8) * 2 + ATT_MAXG * HD * 4);
  bf16* knew = reinterpret_cast<bf16*>(smem_g + 4 * ATT_TILE * (HD + 8) * 2 + ATT_MAXG * HD * 4 + ATT_MAXG * ATT_TILE * 4);
  bf16* vnew = knew + HD;
  const KvView& kv = a.kv;
  const int lane = gtid & 31, warp = gtid >> 5;
  const int G = kv.q_heads / kv.kv_heads;
  const int nq = kv.q_heads, nkv = kv.kv_heads;
  const int total = a.M * nkv * a.nsplit;
  for (int item = group; item < total; item += ngroups) {
    const int s = item % a.nsplit, g = (item / a.nsplit) % nkv, m = item / (a.nsplit * nkv);
    if (!kv.row_mode[m]) continue;
    const int pos = kv.kv_len[m];
    const int L = pos + 1;
    const int ntiles = (L + ATT_TILE - 1) / ATT_TILE;
    const int tps = (ntiles + a.nsplit - 1) / a.nsplit;
    const int t_begin = s * tps, t_end = min(ntiles, (s + 1) * tps);
    const bool owner = (ntiles - 1 >= t_begin) && (ntiles - 1 < t_end);
    const float* row = a.qkv + (size_t)m * (nq + 2 * nkv) * HD;
    group_sync(bar_id);                          // previous item's readers of qs / tiles are done
    // q heads of this kv group, rotated
    for (int i = gtid; i < G * (HD / 2); i += 128) {
      const int h = i / (HD / 2), d = i % (HD / 2);
      float sn, cs;
      sincosf((float)pos * a.inv_freq[d], &sn, &cs);
      const float x1 = __ldcg(row + (g * G + h) * HD + d), x2 = __ldcg(row + (g * G + h) * HD + d + HD / 2);
      qs[h][d] = x1 * cs - x2 * sn;
      qs[h][d + HD / 2] = x2 * cs + x1 * sn;
    }
    if (owner) {                                 // newest token: rotate k, round to bf16, publish to the paged pool
      const int page = kv.page_table[(size_t)m * kv.max_pages + pos / KV_PAGE];
      const size_t o = (((size_t)page * nkv + g) * KV_PAGE + (pos % KV_PAGE)) * HD;
      if (gtid < HD / 2) {
        const int d = gtid;
        float sn, cs;
        sincosf((float)pos * a.inv_freq[d], &sn, &cs);
        const float x1 = __ldcg(row + (nq + g) * HD + d), x2 = __ldcg(row + (nq + g) * HD + d + HD / 2);
        const bf16 k1 = __float2bfloat16_rn(x1 * cs - x2 * sn), k2 = __float2bfloat16_rn(x2 * cs + x1 * sn);
        knew[d] = k1; knew[d + HD / 2] = k2;
        kv.kpool[o + d] = k1; kv.kpool[o + d + HD / 2] = k2;
      } else {
        for (int d = gtid - HD / 2; d < HD; d += 64) {
          const bf16 vv_ = __float2bfloat16_rn(__ldcg(row + (nq + nkv + g) * HD + d));
          vnew[d] = vv_;
          kv.vpool[o + d] = vv_;
        }
      }
    }
    auto prefetch = [&](int t, int buf) {
      const int tok0 = t * ATT_TILE;
      const int page = kv.page_table[(size_t)m * kv.max_pages + tok0 / KV_PAGE];
      const size_t base = (((size_t)page * nkv + g) * KV_PAGE + (tok0 % KV_PAGE)) * HD;
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int idx = gtid + it * 128;
        const int r = idx >> 4, c = (idx & 15) * 8;
        cp_async16(&Ks[buf][r][c], kv.kpool + base + (size_t)r * HD + c, 16);
        cp_async16(&Vs[buf][r][c], kv.vpool + base + (size_t)r * HD + c, (tok0 + r < L) ? 16 : 0);
      }
    };
    if (t_begin < t_end) prefetch(t_begin, 0);
    cp_async_commit();

    float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
    float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
    for (int t = t_begin; t < t_end; ++t) {
      const int buf = (t - t_begin) & 1;
      const int tok0 = t * ATT_TILE;
      if (t + 1 < t_end) prefetch(t + 1, buf ^ 1);
      cp_async_commit();
      cp_async_wait<1>();
      group_sync(bar_id);
      if (owner && t == ntiles - 1) {            // splice the fresh K/V row over whatever the pool held
        const int r = pos - tok0;
        Ks[buf][r][gtid] = knew[gtid];
        Vs[buf][r][gtid] = vnew[gtid];
        group_sync(bar_id);
      }
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int h = warp + 4 * hh;
        if (h >= G) continue;
        float sc = 0.f;
#pragma unroll
        for (int c = 0; c < HD; c += 8) {
          float kf[8];
          bf16x8_unpack(*reinterpret_cast<const uint4*>(&Ks[buf][lane][c]), kf);
          const float4 qa = *reinterpret_cast<const float4*>(&qs[h][c]);
          const float4 qb = *reinterpret_cast<const float4*>(&qs[h][c + 4]);
          sc = fmaf(kf[0], qa.x, sc); sc = fmaf(kf[1], qa.y, sc); sc = fmaf(kf[2], qa.z, sc); sc = fmaf(kf[3], qa.w, sc);
          sc = fmaf(kf[4], qb.x, sc); sc = fmaf(kf[5], qb.y, sc); sc = fmaf(kf[6], qb.z, sc); sc = fmaf(kf[7], qb.w, sc);
        }
        sc = (tok0 + lane < L) ? sc * a.scale : -INFINITY;
        const float mt = warp_max(sc);
        const float mn = fmaxf(m_run[hh], mt);
        const float pj = __expf(sc - mn);
        const float corr = __expf(m_run[hh] - mn);
        l_run[hh] = l_run[hh] * corr + warp_sum(pj);
        m_run[hh] = mn;
        ps[h][lane] = pj;
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[hh][j] *= corr;
#pragma unroll 8
        for (int tt = 0; tt < ATT_TILE; ++tt) {
          const float pv = ps[h][tt];
          const uint2 v2 = *reinterpret_cast<const uint2*>(&Vs[buf][tt][lane * 4]);
          acc[hh][0] = fmaf(pv, __uint_as_float(v2.x << 16), acc[hh][0]);
          acc[hh][1] = fmaf(pv, __uint_as_float(v2.x & 0xffff0000u), acc[hh][1]);
          acc[hh][2] = fmaf(pv, __uint_as_float(v2.y << 16), acc[hh][2]);
          acc[hh][3] = fmaf(pv, __uint_as_float(v2.y & 0xffff0000u), acc[hh][3]);
        }
        __syncwarp();
      }
      group_sync(bar_id);
    }
    cp_async_wait<0>();
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const int h = warp + 4 * hh;
      if (h >= G) continue;
      const size_t o = ((size_t)m * nq + g * G + h) * a.nsplit + s;
      *reinterpret_cast<float4*>(a.part_acc + o * HD + lane * 4) = make_float4(acc[hh][0], acc[hh][1], acc[hh][2], acc[hh][3]);
      if (lane == 0) { a.part_ml[o * 2] = m_run[hh]; a.part_ml[o * 2 + 1] = l_run[hh]; }
    }
  }
}

VV_DEVINL void combine_stage(const AttnOp& a, unsigned char* smem_g, int gtid, int group, int ngroups, int bar_id) {
  float* wsh = reinterpret_cast<float*>(smem_g);       // [nsplit]
  float* red = wsh + 512;
  const int nq = a.kv.q_heads;
  const int total = a.M * nq;
  for (int item = group; item < total; item += ngroups) {
    const int h = item % nq, m = item / nq;
    if (!a.kv.row_mode[m]) continue;
    const size_t o = ((size_t)m * nq + h) * a.nsplit;
    group_sync(bar_id);
    float mx = -INFINITY;
    for (int s = gtid; s < a.nsplit; s += 128) mx = fmaxf(mx, __ldcg(a.part_ml + (o + s) * 2));
    mx = warp_max(mx);
    if ((gtid & 31) == 0) red[gtid >> 5] = mx;
    group_sync(bar_id);
    mx = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
    group_sync(bar_id);
    float den = 0.f;
    for (int s = gtid; s < a.nsplit; s += 128) {
      const float ms = __ldcg(a.part_ml + (o + s) * 2);
      const float w = (ms == -INFINITY) ? 0.f : __expf(ms - mx);
      wsh[s] = w;
      den = fmaf(w, __ldcg(a.part_ml + (o + s) * 2 + 1), den);
    }
    den = warp_sum(den);
    if ((gtid & 31) == 0) red[gtid >> 5] = den;
    group_sync(bar_id);
    den = red[0] + red[1] + red[2] + red[3];
    float num = 0.f;
#pragma unroll 8
    for (int s = 0; s < a.nsplit; ++s) num = fmaf(wsh[s], __ldcg(a.part_acc + (o + s) * HD + gtid), num);
    a.attn_out[((size_t)m * nq + h) * HD + gtid] = num / den;
  }
}

VV_DEVINL void final_stage(const FinalOp& f, float* smem_f) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float* red = smem_f;           // [8][9]
  for (int m = blockIdx.x; m < f.M; m += gridDim.x) {
    const float* hr = f.h + (size_t)m * f.H;
    float ss = 0.f;
    for (int k = tid; k < f.H; k += 256) { const float v = __ldcg(hr + k); ss += v * v; }
    ss = warp_sum(ss);
    __syncthreads();
    if (lane == 0) red[warp] = ss;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i];
    const float inv = rsqrtf(t / (float)f.H + f.eps);
    float acc[8];
#pragma unroll
    for (int v = 0; v < 8; ++v) acc[v] = 0.f;
    for (int k = tid; k < f.H; k += 256) {
      const float x = __ldcg(hr + k) * inv * f.norm_w[k];
      f.hidden[(size_t)m * f.H + k] = x;
      if (m < f.B) {
#pragma unroll
        for (int v = 0; v < 8; ++v) if (v < f.n_valid) acc[v] = fmaf(__bfloat162float(f.w_valid[(size_t)v * f.H + k]), x, acc[v]);
      }
    }
    if (m < f.B) {
      __syncthreads();
#pragma unroll
      for (int v = 0; v < 8; ++v) { acc[v] = warp_sum(acc[v]); if (lane == 0) red[16 + warp * 8 + v] = acc[v]; }
      __syncthreads();
      if (tid == 0) {
        int best = 0; float bv = -INFINITY;
        for (int v = 0; v < f.n_valid; ++v) {
          float tt = 0.f;
          for (int w = 0; w < 8; ++w) tt += red[16 + w * 8 + v];
          f.logits[(size_t)m * f.n_valid + v] = tt;
          if (tt > bv) { bv = tt; best = v; }
        }
        f.tokens[m] = f.valid_ids[best];
      }
    }
  }
}

VV_DEVINL void dpm_stage(const DpmOp& d, float* smem_f) {
  const int tid = threadIdx.x;
  float* zs = smem_f;            // [B][64]
  __syncthreads();
  for (int i = tid; i < d.B * 64; i += 256) {
    const int b = i >> 6, j = i & 63;
    float zn, x0 = 0.f;
    if (d.step < 0) {
      zn = d.noise[b * 64 + j];
    } else {
      const DpmCoef c = d.coef[d.step];
      const float vc = __ldcg(d.v + (size_t)b * 64 + j), vu = __ldcg(d.v + (size_t)(d.B + b) * 64 + j);
      const float vv_ = vu + d.cfg * (vc - vu);
      const float zo = __ldcg(d.z_in + i);
      x0 = c.a0 * zo - c.s0 * vv_;
      zn = c.ks * zo - c.kx * x0;
      if (c.order == 2) zn -= 0.5f * c.kx * (c.rinv * (x0 - __ldcg(d.x0_in + i)));
    }
    zs[i] = zn;
    if (blockIdx.x == 0) {
      d.z_out[i] = zn;
      d.x0_out[i] = x0;
      if (d.latent_out) d.latent_out[i] = zn;
    }
  }
  __syncthreads();
  if (!d.do_proj) return;
  for (int idx = blockIdx.x * 256 + tid; idx < d.B * d.H; idx += gridDim.x * 256) {
    const int b = idx / d.H, n = idx % d.H;
    const uint4* wr = reinterpret_cast<const uint4*>(d.w_noisy + (size_t)n * 64);
    float acc = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      float wf[8];
      bf16x8_unpack(wr[c], wf);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc = fmaf(wf[j], zs[b * 64 + c * 8 + j], acc);
    }
    d.xout[(size_t)b * d.H + n] = acc;
    d.xout[(size_t)(d.B + b) * d.H + n] = acc;
  }
}

// ---------------------------------------------------------------------------------------------------------------
template <int MB>
__global__ void __launch_bounds__(256) program_kernel(const Op* __restrict__ ops, int n_ops, GridBar* gb) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ float s_inv[MB];
  __shared__ float s_part[8];
  pdl_trigger();
  pdl_wait();
  const int tid = threadIdx.x;
  for (int i = 0; i < n_ops; ++i) {
    const Op& op = ops[i];
    if (op.kind == OP_GEMV) {
      gemv_stage<MB>(op.g, reinterpret_cast<float*>(smem_raw), gb, op.barrier_before != 0, s_inv, s_part);
      continue;
    }
    if (op.barrier_before) grid_barrier(gb, gridDim.x);
    switch (op.kind) {
      case OP_ATTN: {
        const int grp = tid >> 7;
        attn_stage(op.a, smem_raw + grp * ATT_GROUP_SMEM, tid & 127, blockIdx.x * 2 + grp, gridDim.x * 2, 1 + grp);
        __syncthreads();
        break;
      }
      case OP_COMBINE: {
        const int grp = tid >> 7;
        combine_stage(op.a, smem_raw + grp * ATT_GROUP_SMEM, tid & 127, blockIdx.x * 2 + grp, gridDim.x * 2, 1 + grp);
        __syncthreads();
        break;
      }
      case OP_FINAL: final_stage(op.f, reinterpret_cast<float*>(smem_raw)); break;
      case OP_DPM: dpm_stage(op.d, reinterpret_cast<float*>(smem_raw)); break;
      default: break;
    }
  }
}

}  // namespace vv
