// vv_kernels.cuh -- hand-written sm_100a kernels of the VibeVoice generation hot path.
//
// Everything in the per-frame loop runs at M <= 2B rows against bf16 weight matrices, i.e. it is
// HBM-bandwidth bound: the kernels below stream each weight byte exactly once with
// 16-byte coalesced loads, keep activations in fp32 (registers / shared memory), accumulate in
// fp32 and fuse the surrounding norm / modulation / activation / residual work into the GEMV
// prologue and epilogue so no activation round-trips through HBM more than once.
//
// Reference anchors (under /root/reference/vibevoice/modular unless noted):
//   gemv prologues RMSNORM/ADALN      modular_vibevoice_diffusion_head.py:31-45, 158-161, 184-188;
//                                     transformers Qwen2RMSNorm (modeling_qwen2.py:249-266)
//   gemv epilogues SWIGLU/GATED_RESID modular_vibevoice_diffusion_head.py:116-123, 158-161
//   gemv epilogues GELU/GAMMA_RESID   modular_vibevoice_tokenizer.py:592-596, 670-682
//   rope_append / attn_*              transformers Qwen2Attention (modeling_qwen2.py:116-174)
//   dpm_update_proj                   modeling_vibevoice_inference.py:703-709 + schedule/dpm_solver.py:581-584, 669-677, 738-764
//   assemble_window / dwconv_res      modular_vibevoice_tokenizer.py:327-382 (streaming SConv1d), 786-794
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#define VV_DEVINL __device__ __forceinline__

namespace vv {

typedef __nv_bfloat16 bf16;

enum Prologue { PRO_NONE = 0, PRO_RMSNORM = 1, PRO_ADALN = 2, PRO_SILU = 3 };
enum Epilogue { EPI_NONE = 0, EPI_RESID = 2, EPI_GATED_RESID = 3, EPI_GAMMA_RESID = 4, EPI_SWIGLU = 5, EPI_GELU = 6, EPI_SILU = 7 };

// row m of a logical [M, K] operand lives at base + (m / T) * bs + (m % T) * rs; rows may overlap
// (rs < K) -- that is how causal / strided / transposed convolutions read their input window.
struct RowMap {
  int T;
  long long bs;
  long long rs;
  VV_DEVINL long long off(int m) const { return (long long)(m / T) * bs + (long long)(m % T) * rs; }
};

static inline RowMap dense_rows(long long ld) { RowMap r; r.T = 1 << 30; r.bs = 0; r.rs = ld; return r; }

struct GemvP {
  const bf16* W;        // [N, K] row-major, K % 8 == 0
  const float* bias;    // [N] or null (added before the epilogue op)
  const float* x;       // fp32 activations
  RowMap xmap;
  float* y;             // [M, ldy]
  int ldy;
  int M, N, K;
  int pro;
  const float* pro_w;       // [K] norm weight (may be null for ADALN without affine)
  float pro_eps;
  const float* pro_shift;   // ADALN: [M, pro_ld]
  const float* pro_scale;
  long long pro_ld;
  int epi;
  const float* epi_a;       // GATED_RESID: gate [M, epi_lda]; GAMMA_RESID: gamma [N]
  long long epi_lda;
  const float* res;         // residual [M, ldres]
  int ldres;
  int WK;                   // warps splitting K inside a CTA (1,2,4,8); WR = 8 / WK row-quads per task
};

// ---------------------------------------------------------------------------------------------
VV_DEVINL float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
VV_DEVINL float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
VV_DEVINL float silu_f(float x) { return x / (1.0f + __expf(-x)); }
VV_DEVINL float gelu_erf_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

VV_DEVINL uint4 ldg_stream(const void* p) {  // weights are read once: bypass L1 allocation
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
VV_DEVINL void bf16x8_unpack(const uint4& u, float* f) {
  f[0] = __uint_as_float(u.x << 16); f[1] = __uint_as_float(u.x & 0xffff0000u);
  f[2] = __uint_as_float(u.y << 16); f[3] = __uint_as_float(u.y & 0xffff0000u);
  f[4] = __uint_as_float(u.z << 16); f[5] = __uint_as_float(u.z & 0xffff0000u);
  f[6] = __uint_as_float(u.w << 16); f[7] = __uint_as_float(u.w & 0xffff0000u);
}
VV_DEVINL float bf16_bits_to_f(unsigned short b) { return __uint_as_float(((unsigned)b) << 16); }

// position of element k inside the staged activation row: chunks of 256, inside a chunk the 8
// values a lane consumes are split into two float4 planes so both LDS.128 are bank-conflict free.
VV_DEVINL int xs_pos(int k) {
  int c = k >> 8, kk = k & 255;
  int lane = kk >> 3, j = kk & 7;
  return (c << 8) + ((j >> 2) << 7) + (lane << 2) + (j & 3);
}

VV_DEVINL void epi_store(const GemvP& p, int m, int n, float v) {
  // n indexes the weight row; bias already added
  switch (p.epi) {
    case EPI_RESID: v += p.res[(long long)m * p.ldres + n]; break;
    case EPI_GATED_RESID: v = p.res[(long long)m * p.ldres + n] + p.epi_a[(long long)m * p.epi_lda + n] * v; break;
    case EPI_GAMMA_RESID: v = p.res[(long long)m * p.ldres + n] + p.epi_a[n] * v; break;
    case EPI_GELU: v = gelu_erf_f(v); break;
    case EPI_SILU: v = silu_f(v); break;
    default: break;
  }
  p.y[(long long)m * p.ldy + n] = v;
}

// ---------------------------------------------------------------------------------------------
// GEMV: y[m, n] = epi( sum_k W[n,k] * pro(x)[m,k] + bias[n] ),  M <= 16 (blocks of MB rows).
// CTA = 8 warps arranged as WR row-quads x WK k-splits.  The staged activation block lives in
// shared memory (fp32), weights stream from HBM straight into registers (16 B / lane / load).
// ---------------------------------------------------------------------------------------------
template <int MB>
__global__ void __launch_bounds__(256) gemv_kernel(GemvP p) {
  extern __shared__ __align__(16) float smem_f[];
  const int K = p.K, N = p.N;
  const int Kp = (K + 255) & ~255;
  float* xs = smem_f;                   // [MB][Kp]
  float* red = smem_f + MB * Kp;        // [2][8 warps][4*MB]
  __shared__ float s_inv[MB];
  __shared__ float s_part[8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int WK = p.WK, WR = 8 / WK;
  const int wr = warp / WK, wk = warp % WK;
  const int ntasks = (N + 4 * WR - 1) / (4 * WR);
  const int nchunks_full = K >> 8;
  const bool has_tail = (K & 255) != 0;

  for (int m0 = 0; m0 < p.M; m0 += MB) {
    __syncthreads();
    // ---- stage pro(x) for rows m0..m0+MB-1 ----
    const bool need_inv = (p.pro == PRO_RMSNORM || p.pro == PRO_ADALN);
    if (need_inv) {
      for (int m = 0; m < MB; ++m) {
        float ss = 0.f;
        if (m0 + m < p.M) {
          const float* xr = p.x + p.xmap.off(m0 + m);
          for (int k = tid; k < K; k += 256) { float v = xr[k]; ss += v * v; }
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
      for (int k = tid; k < Kp; k += 256) {
        float v = 0.f;
        if (valid && k < K) {
          v = xr[k];
          if (p.pro == PRO_RMSNORM) v = v * inv * p.pro_w[k];
          else if (p.pro == PRO_ADALN) {
            float w = p.pro_w ? p.pro_w[k] : 1.f;
            long long o = (long long)(m0 + m) * p.pro_ld + k;
            v = v * inv * w * (1.f + p.pro_scale[o]) + p.pro_shift[o];
          } else if (p.pro == PRO_SILU) v = silu_f(v);
        }
        xs[m * Kp + xs_pos(k)] = v;
      }
    }
    __syncthreads();

    int parity = 0;
    for (int task = blockIdx.x; task < ntasks; task += gridDim.x, parity ^= 1) {
      const int r0 = (task * WR + wr) * 4;
      const bf16* wrow[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) wrow[r] = p.W + (size_t)min(r0 + r, N - 1) * K + lane * 8;
      float acc[4][MB];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int m = 0; m < MB; ++m) acc[r][m] = 0.f;

      if (r0 < N) {
#pragma unroll 2
        for (int c = wk; c < nchunks_full; c += WK) {
          uint4 wv[4];
#pragma unroll
          for (int r = 0; r < 4; ++r) wv[r] = ldg_stream(wrow[r] + (c << 8));
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
            bf16x8_unpack(wv[r], wf);
#pragma unroll
            for (int m = 0; m < MB; ++m)
#pragma unroll
              for (int j = 0; j < 8; ++j) acc[r][m] = fmaf(wf[j], xv[m][j], acc[r][m]);
          }
        }
        if (has_tail && wk == (nchunks_full % WK)) {
          const int c = nchunks_full;
          if ((c << 8) + lane * 8 < K) {
            uint4 wv[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) wv[r] = ldg_stream(wrow[r] + (c << 8));
#pragma unroll
            for (int r = 0; r < 4; ++r) {
              float wf[8];
              bf16x8_unpack(wv[r], wf);
#pragma unroll
              for (int m = 0; m < MB; ++m) {
                const float4 a = *reinterpret_cast<const float4*>(xs + m * Kp + (c << 8) + (lane << 2));
                const float4 b = *reinterpret_cast<const float4*>(xs + m * Kp + (c << 8) + 128 + (lane << 2));
                acc[r][m] = fmaf(wf[0], a.x, acc[r][m]); acc[r][m] = fmaf(wf[1], a.y, acc[r][m]);
                acc[r][m] = fmaf(wf[2], a.z, acc[r][m]); acc[r][m] = fmaf(wf[3], a.w, acc[r][m]);
                acc[r][m] = fmaf(wf[4], b.x, acc[r][m]); acc[r][m] = fmaf(wf[5], b.y, acc[r][m]);
                acc[r][m] = fmaf(wf[6], b.z, acc[r][m]); acc[r][m] = fmaf(wf[7], b.w, acc[r][m]);
              }
            }
          }
        }
      }
      // ---- reduce: lanes -> lane 0; k-split warps -> shared ----
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
      // ---- epilogue: one thread per (row-quad, r, m) ----
      if (p.epi == EPI_SWIGLU) {
        if (tid < WR * 2 * MB) {
          const int q = tid / (2 * MB), pr = (tid / MB) % 2, m = tid % MB;
          const int n0 = (task * WR + q) * 4 + pr * 2;
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
          const int n = (task * WR + q) * 4 + r;
          if (n < N && m0 + m < p.M) {
            float v = 0.f;
            for (int s = 0; s < WK; ++s) v += rbuf[(q * WK + s) * (4 * MB) + r * MB + m];
            if (p.bias) v += p.bias[n];
            epi_store(p, m0 + m, n, v);
          }
        }
      }
      // double-buffered `red`: the next task writes the other half, the one after is fenced by
      // the next __syncthreads, so no trailing barrier is needed here.
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Tiled GEMM for M > 16 (codec stages with many time steps and narrow channels).
// C[M,N] = A[M,K] (fp32, RowMap) * W[N,K]^T (bf16); 64x64 tile, BK = 16, 4x4 micro-tile / thread.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gemm_tiled_kernel(GemvP p) {
  __shared__ float As[16][64 + 4];
  __shared__ float Ws[16][64 + 4];
  const int tid = threadIdx.x;
  const int bm = blockIdx.y * 64, bn = blockIdx.x * 64;
  const int tx = tid & 15, ty = tid >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int lr = tid >> 2, lk = (tid & 3) * 4;   // 64 rows x 4 k-quads
  const int am = bm + lr, wn = bn + lr;
  const float* arow = (am < p.M) ? p.x + p.xmap.off(am) : nullptr;
  const bf16* wrow = (wn < p.N) ? p.W + (size_t)wn * p.K : nullptr;
  for (int k0 = 0; k0 < p.K; k0 += 16) {
    float a[4] = {0.f, 0.f, 0.f, 0.f}, w[4] = {0.f, 0.f, 0.f, 0.f};
    if (arow) {
#pragma unroll
      for (int j = 0; j < 4; ++j) if (k0 + lk + j < p.K) a[j] = arow[k0 + lk + j];
    }
    if (wrow) {
#pragma unroll
      for (int j = 0; j < 4; ++j) if (k0 + lk + j < p.K) w[j] = __bfloat162float(wrow[k0 + lk + j]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) { As[lk + j][lr] = a[j]; Ws[lk + j][lr] = w[j]; }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float av[4], wv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { av[i] = As[kk][ty * 4 + i]; wv[i] = Ws[kk][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = bm + ty * 4 + i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = bn + tx * 4 + j;
      if (n >= p.N) continue;
      float v = acc[i][j] + (p.bias ? p.bias[n] : 0.f);
      epi_store(p, m, n, v);
    }
  }
}

// thread-per-output small-K product with fp32 weights (encoder stem conv 1->32 k7, decoder head conv 32->1 k7)
__global__ void conv_naive_kernel(const float* __restrict__ W, const float* __restrict__ bias, const float* __restrict__ x,
                                  RowMap xmap, float* __restrict__ y, int M, int N, int K) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)M * N) return;
  const int m = (int)(idx / N), n = (int)(idx % N);
  const float* xr = x + xmap.off(m);
  const float* wr = W + (size_t)n * K;
  float acc = bias ? bias[n] : 0.f;
  for (int k = 0; k < K; ++k) acc = fmaf(wr[k], xr[k], acc);
  y[idx] = acc;
}

// y[m,:] = rmsnorm(x[m,:]) * w   (one warp per row)
__global__ void rows_norm_kernel(const float* __restrict__ x, const float* __restrict__ w, float* __restrict__ y,
                                 int M, int C, float eps) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const float* xr = x + (size_t)row * C;
  float ss = 0.f;
  for (int c = lane; c < C; c += 32) { float v = xr[c]; ss += v * v; }
  ss = warp_sum(ss);
  const float inv = rsqrtf(ss / (float)C + eps);
  for (int c = lane; c < C; c += 32) y[(size_t)row * C + c] = xr[c] * inv * (w ? w[c] : 1.f);
}

// ---------------------------------------------------------------------------------------------
// Streaming conv state.  win[b] = [hist[b] (ctx rows) ; f(src[b]) (T rows)], and the last ctx rows
// of the window are staged into hist_next[b]; `advance_kernel` commits hist_next -> hist for the
// rows that actually took this frame (a-8: VibeVoiceTokenizerStreamingCache, tokenizer.py:193-256).
// f = affine (alpha*x+beta) or RMSNorm*w.
// ---------------------------------------------------------------------------------------------
__global__ void assemble_window_kernel(const float* __restrict__ src, const float* __restrict__ hist,
                                       float* __restrict__ win, float* __restrict__ hist_next, int B, int T, int ctx, int C,
                                       const float* __restrict__ norm_w, float eps, float alpha, float beta) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const int R = ctx + T;
  if (row >= B * R) return;
  const int b = row / R, j = row % R;
  float* wr = win + ((size_t)b * R + j) * C;
  float* hn = (j >= T) ? hist_next + ((size_t)b * ctx + (j - T)) * C : nullptr;
  if (j < ctx) {
    const float* hr = hist + ((size_t)b * ctx + j) * C;
    for (int c = lane; c < C; c += 32) { float v = hr[c]; wr[c] = v; if (hn) hn[c] = v; }
  } else {
    const float* sr = src + ((size_t)b * T + (j - ctx)) * C;
    float inv = 1.f;
    if (norm_w) {
      float ss = 0.f;
      for (int c = lane; c < C; c += 32) { float v = sr[c]; ss += v * v; }
      ss = warp_sum(ss);
      inv = rsqrtf(ss / (float)C + eps);
    }
    for (int c = lane; c < C; c += 32) {
      float v = norm_w ? sr[c] * inv * norm_w[c] : sr[c] * alpha + beta;
      wr[c] = v;
      if (hn) hn[c] = v;
    }
  }
}

// out = x + gamma * (bias + sum_j w[j][c] * win[t+j][c])   (depthwise causal conv k=7 + layer scale + residual)
__global__ void dwconv_res_kernel(const float* __restrict__ x, const float* __restrict__ win, const float* __restrict__ w /*[7][C]*/,
                                  const float* __restrict__ bias, const float* __restrict__ gamma, float* __restrict__ out,
                                  int B, int T, int C) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * T * C) return;
  const int c = (int)(idx % C);
  const long long bt = idx / C;
  const int t = (int)(bt % T), b = (int)(bt / T);
  const float* wp = win + ((size_t)b * (T + 6) + t) * C + c;
  float acc = bias[c];
#pragma unroll
  for (int j = 0; j < 7; ++j) acc = fmaf(w[j * C + c], wp[(size_t)j * C], acc);
  out[idx] = x[idx] + gamma[c] * acc;
}

struct StateSeg { float* hist; float* next; int n; };   // n floats per batch row

__global__ void advance_kernel(const StateSeg* __restrict__ segs, const int* __restrict__ active) {
  const int b = blockIdx.y;
  if (active && !active[b]) return;
  const StateSeg s = segs[blockIdx.x];
  float* d = s.hist + (size_t)b * s.n;
  const float* a = s.next + (size_t)b * s.n;
  for (int i = threadIdx.x; i < s.n; i += blockDim.x) d[i] = a[i];
}
__global__ void state_zero_kernel(const StateSeg* __restrict__ segs, const int* __restrict__ rows) {
  const int b = rows[blockIdx.y];
  const StateSeg s = segs[blockIdx.x];
  float* d = s.hist + (size_t)b * s.n;
  for (int i = threadIdx.x; i < s.n; i += blockDim.x) d[i] = 0.f;
}

// ---------------------------------------------------------------------------------------------
// LLM decode step pieces
// ---------------------------------------------------------------------------------------------
constexpr int KV_PAGE = 64;     // tokens per page
constexpr int HD = 128;         // head_dim of both shipped models (configs/*.json)

struct KvView {
  bf16* kpool; bf16* vpool;     // this layer: [n_pages][kv_heads][KV_PAGE][HD]
  const int* page_table;        // [n_seq][max_pages]
  int max_pages;
  const int* kv_len;            // [n_seq] committed length
  const int* row_mode;          // [n_seq] 0 skip / 1 run
  int kv_heads, q_heads;
};

// qkv [M, (q_heads + 2 kv_heads) * HD] fp32 (bias added) -> q_rot fp32, K/V (bf16) appended at kv_len[m]
__global__ void rope_append_kernel(const float* __restrict__ qkv, float* __restrict__ q_rot, KvView kv,
                                   const float* __restrict__ inv_freq /*[HD/2]*/) {
  const int m = blockIdx.x;
  if (!kv.row_mode[m]) return;
  const int pos = kv.kv_len[m];
  const int nq = kv.q_heads, nkv = kv.kv_heads;
  const float* row = qkv + (size_t)m * (nq + 2 * nkv) * HD;
  const int page = kv.page_table[(size_t)m * kv.max_pages + pos / KV_PAGE];
  const int slot = pos % KV_PAGE;
  for (int i = threadIdx.x; i < (nq + nkv) * (HD / 2); i += blockDim.x) {
    const int h = i / (HD / 2), d = i % (HD / 2);
    const float ang = (float)pos * inv_freq[d];
    float sn, cs;
    sincosf(ang, &sn, &cs);
    const float x1 = row[h * HD + d], x2 = row[h * HD + d + HD / 2];
    const float o1 = x1 * cs - x2 * sn, o2 = x2 * cs + x1 * sn;
    if (h < nq) {
      q_rot[((size_t)m * nq + h) * HD + d] = o1;
      q_rot[((size_t)m * nq + h) * HD + d + HD / 2] = o2;
    } else {
      bf16* kp = kv.kpool + (((size_t)page * nkv + (h - nq)) * KV_PAGE + slot) * HD;
      kp[d] = __float2bfloat16_rn(o1);
      kp[d + HD / 2] = __float2bfloat16_rn(o2);
    }
  }
  for (int i = threadIdx.x; i < nkv * HD; i += blockDim.x) {
    const int h = i / HD, d = i % HD;
    bf16* vp = kv.vpool + (((size_t)page * nkv + h) * KV_PAGE + slot) * HD;
    vp[d] = __float2bfloat16_rn(row[(nq + nkv + h) * HD + d]);
  }
}

// split-KV partial attention: CTA = (split, kv head, sequence); 4 warps; 32-token tiles staged in smem.
constexpr int ATT_TILE = 32;
constexpr int ATT_MAXG = 8;     // q heads per kv head (6 for 1.5B, 7 for 7B)
__global__ void __launch_bounds__(128) attn_partial_kernel(const float* __restrict__ q_rot, KvView kv, float* __restrict__ part_acc,
                                                           float* __restrict__ part_ml, int nsplit, float scale) {
  const int s = blockIdx.x, g = blockIdx.y, m = blockIdx.z;
  if (!kv.row_mode[m]) return;
  const int G = kv.q_heads / kv.kv_heads;
  const int L = kv.kv_len[m] + 1;
  const int ntiles = (L + ATT_TILE - 1) / ATT_TILE;
  const int tps = (ntiles + nsplit - 1) / nsplit;
  const int t_begin = s * tps, t_end = min(ntiles, (s + 1) * tps);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  __shared__ __align__(16) bf16 Ks[ATT_TILE][HD + 8];
  __shared__ __align__(16) bf16 Vs[ATT_TILE][HD + 8];
  __shared__ __align__(16) float qs[ATT_MAXG][HD];
  __shared__ float ps[ATT_MAXG][ATT_TILE];

  for (int i = tid; i < G * HD; i += 128) qs[i / HD][i % HD] = q_rot[((size_t)m * kv.q_heads + g * G) * HD + i];

  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};

  for (int t = t_begin; t < t_end; ++t) {
    const int tok0 = t * ATT_TILE;
    const int page = kv.page_table[(size_t)m * kv.max_pages + tok0 / KV_PAGE];
    const size_t base = (((size_t)page * kv.kv_heads + g) * KV_PAGE + (tok0 % KV_PAGE)) * HD;
    __syncthreads();
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int idx = tid + it * 128;
      const int r = idx >> 4, c = (idx & 15) * 8;
      uint4 kk = ldg_stream(kv.kpool + base + (size_t)r * HD + c);
      uint4 vv = (tok0 + r < L) ? ldg_stream(kv.vpool + base + (size_t)r * HD + c) : make_uint4(0, 0, 0, 0);
      *reinterpret_cast<uint4*>(&Ks[r][c]) = kk;
      *reinterpret_cast<uint4*>(&Vs[r][c]) = vv;
    }
    __syncthreads();
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const int h = warp + 4 * hh;
      if (h >= G) continue;
      float sc = 0.f;
#pragma unroll
      for (int c = 0; c < HD; c += 8) {
        float kf[8];
        bf16x8_unpack(*reinterpret_cast<const uint4*>(&Ks[lane][c]), kf);
        const float4 qa = *reinterpret_cast<const float4*>(&qs[h][c]);
        const float4 qb = *reinterpret_cast<const float4*>(&qs[h][c + 4]);
        sc = fmaf(kf[0], qa.x, sc); sc = fmaf(kf[1], qa.y, sc); sc = fmaf(kf[2], qa.z, sc); sc = fmaf(kf[3], qa.w, sc);
        sc = fmaf(kf[4], qb.x, sc); sc = fmaf(kf[5], qb.y, sc); sc = fmaf(kf[6], qb.z, sc); sc = fmaf(kf[7], qb.w, sc);
      }
      sc = (tok0 + lane < L) ? sc * scale : -INFINITY;
      const float mt = warp_max(sc);
      const float mn = fmaxf(m_run[hh], mt);           // finite: every tile in range has >= 1 valid token
      const float pj = __expf(sc - mn);
      const float corr = __expf(m_run[hh] - mn);        // exp(-inf) = 0 on the first tile
      l_run[hh] = l_run[hh] * corr + warp_sum(pj);
      m_run[hh] = mn;
      ps[h][lane] = pj;
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[hh][j] *= corr;
#pragma unroll 8
      for (int tt = 0; tt < ATT_TILE; ++tt) {
        const float pv = ps[h][tt];
        const uint2 v2 = *reinterpret_cast<const uint2*>(&Vs[tt][lane * 4]);
        acc[hh][0] = fmaf(pv, __uint_as_float(v2.x << 16), acc[hh][0]);
        acc[hh][1] = fmaf(pv, __uint_as_float(v2.x & 0xffff0000u), acc[hh][1]);
        acc[hh][2] = fmaf(pv, __uint_as_float(v2.y << 16), acc[hh][2]);
        acc[hh][3] = fmaf(pv, __uint_as_float(v2.y & 0xffff0000u), acc[hh][3]);
      }
      __syncwarp();
    }
  }
#pragma unroll
  for (int hh = 0; hh < 2; ++hh) {
    const int h = warp + 4 * hh;
    if (h >= G) continue;
    const size_t o = ((size_t)m * kv.q_heads + g * G + h) * nsplit + s;
    float4 a = make_float4(acc[hh][0], acc[hh][1], acc[hh][2], acc[hh][3]);
    *reinterpret_cast<float4*>(part_acc + o * HD + lane * 4) = a;
    if (lane == 0) { part_ml[o * 2] = m_run[hh]; part_ml[o * 2 + 1] = l_run[hh]; }
  }
}

__global__ void __launch_bounds__(128) attn_combine_kernel(const float* __restrict__ part_acc, const float* __restrict__ part_ml,
                                                           const int* __restrict__ row_mode, float* __restrict__ out, int q_heads,
                                                           int nsplit) {
  const int h = blockIdx.x, m = blockIdx.y, d = threadIdx.x;
  if (!row_mode[m]) return;
  const size_t o = ((size_t)m * q_heads + h) * nsplit;
  float mx = -INFINITY;
  for (int s = 0; s < nsplit; ++s) mx = fmaxf(mx, part_ml[(o + s) * 2]);
  float num = 0.f, den = 0.f;
  for (int s = 0; s < nsplit; ++s) {
    const float ms = part_ml[(o + s) * 2];
    if (ms == -INFINITY) continue;
    const float w = __expf(ms - mx);
    num = fmaf(w, part_acc[(o + s) * HD + d], num);
    den = fmaf(w, part_ml[(o + s) * 2 + 1], den);
  }
  out[((size_t)m * q_heads + h) * HD + d] = num / den;
}

__global__ void embed_gather_kernel(const bf16* __restrict__ table, const int* __restrict__ tokens, float* __restrict__ out, int H) {
  const int r = blockIdx.x;
  const bf16* row = table + (size_t)tokens[r] * H;
  for (int k = threadIdx.x; k < H; k += blockDim.x) out[(size_t)r * H + k] = __bfloat162float(row[k]);
}

// logits over the valid ids + constrained argmax (VibeVoiceTokenConstraintProcessor + argmax, :53-66, :498)
__global__ void __launch_bounds__(256) lm_head_argmax_kernel(const float* __restrict__ hidden, const bf16* __restrict__ w_valid,
                                                             const int* __restrict__ valid_ids, int n_valid, int H,
                                                             float* __restrict__ logits, int* __restrict__ tokens) {
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  __shared__ float red[8][8];
  float acc[8];
#pragma unroll
  for (int v = 0; v < 8; ++v) acc[v] = 0.f;
  const float* hr = hidden + (size_t)b * H;
  for (int k = tid; k < H; k += 256) {
    const float x = hr[k];
#pragma unroll
    for (int v = 0; v < 8; ++v) if (v < n_valid) acc[v] = fmaf(__bfloat162float(w_valid[(size_t)v * H + k]), x, acc[v]);
  }
#pragma unroll
  for (int v = 0; v < 8; ++v) { acc[v] = warp_sum(acc[v]); if (lane == 0) red[warp][v] = acc[v]; }
  __syncthreads();
  if (tid == 0) {
    int best = 0; float bv = -INFINITY;
    for (int v = 0; v < n_valid; ++v) {
      float t = 0.f;
      for (int w = 0; w < 8; ++w) t += red[w][v];
      logits[(size_t)b * n_valid + v] = t;
      if (t > bv) { bv = t; best = v; }     // strict > : lowest id wins ties (valid_ids ascending)
    }
    tokens[b] = valid_ids[best];
  }
}

__global__ void kv_commit_kernel(int* __restrict__ kv_len, const int* __restrict__ advance, int n) {
  const int i = threadIdx.x;
  if (i < n) kv_len[i] += advance[i];
}

// ---------------------------------------------------------------------------------------------
// diffusion sampler glue
// ---------------------------------------------------------------------------------------------
// c_all[i][r][:] = silu(condp[r][:] + temb[i][:])   for all steps i (sample-independent t-embedding)
__global__ void head_cond_prep_kernel(const float* __restrict__ condp, const float* __restrict__ temb, float* __restrict__ c_all,
                                      int n_steps, int R, int H) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)n_steps * R * H) return;
  const int k = (int)(idx % H);
  const int r = (int)((idx / H) % R);
  const int i = (int)(idx / ((long long)H * R));
  c_all[idx] = silu_f(condp[(size_t)r * H + k] + temb[(size_t)i * H + k]);
}

struct DpmCoef { float a0, s0, ks, kx, rinv; int order; };

// Step `i` CFG + DPM-Solver++(2M) update of z from the head output v of step i, then (optionally)
// the projection x = noisy_images_proj(z') for the next head evaluation, rows b and B+b.
//   v = v_u + s (v_c - v_u); x0 = a0 z - s0 v; z' = ks z - kx x0 [- 0.5 kx rinv (x0 - x0_prev)]
__global__ void __launch_bounds__(256) dpm_update_proj_kernel(float* __restrict__ z, float* __restrict__ x0_prev,
                                                              const float* __restrict__ v, const float* __restrict__ noise,
                                                              const DpmCoef* __restrict__ coef, int step, float cfg,
                                                              const bf16* __restrict__ w_noisy /*[H][64]*/, float* __restrict__ xout,
                                                              float* __restrict__ latent_out, int B, int H, int do_proj) {
  const int b = blockIdx.x, tid = threadIdx.x;
  __shared__ float zs[64];
  if (tid < 64) {
    float zn;
    if (step < 0) {
      zn = noise[b * 64 + tid];                           // z_0 = CPU-RNG noise (:701)
    } else {
      const DpmCoef c = coef[step];
      const float vc = v[(size_t)b * 64 + tid], vu = v[(size_t)(B + b) * 64 + tid];
      const float vv = vu + cfg * (vc - vu);
      const float zo = z[b * 64 + tid];
      const float x0 = c.a0 * zo - c.s0 * vv;
      zn = c.ks * zo - c.kx * x0;
      if (c.order == 2) zn -= 0.5f * c.kx * (c.rinv * (x0 - x0_prev[b * 64 + tid]));
      x0_prev[b * 64 + tid] = x0;
    }
    z[b * 64 + tid] = zn;
    zs[tid] = zn;
    if (latent_out) latent_out[b * 64 + tid] = zn;
  }
  __syncthreads();
  if (!do_proj) return;
  for (int n = tid; n < H; n += 256) {
    const uint4* wr = reinterpret_cast<const uint4*>(w_noisy + (size_t)n * 64);
    float acc = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      float wf[8];
      bf16x8_unpack(wr[c], wf);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc = fmaf(wf[j], zs[c * 8 + j], acc);
    }
    xout[(size_t)b * H + n] = acc;
    xout[(size_t)(B + b) * H + n] = acc;
  }
}

// embeds[b] = active[b] ? e_new[b] : embeds[b];  embeds[B+b] = embeds[b]  (negative stream is fed the same input, :579-581)
__global__ void select_embeds_kernel(float* __restrict__ embeds, const float* __restrict__ e_new, const int* __restrict__ active,
                                     int B, int H) {
  const int b = blockIdx.x;
  const bool a = active[b] != 0;
  for (int k = threadIdx.x; k < H; k += blockDim.x) {
    const float v = a ? e_new[(size_t)b * H + k] : embeds[(size_t)b * H + k];
    embeds[(size_t)b * H + k] = v;
    embeds[(size_t)(B + b) * H + k] = v;
  }
}

// ---------------------------------------------------------------------------------------------
// load-time repack kernels (run once in vv_finalize_weights)
// ---------------------------------------------------------------------------------------------
__global__ void cvt_f32_to_bf16_kernel(const float* __restrict__ s, bf16* __restrict__ d, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) d[i] = __float2bfloat16_rn(s[i]);
}
__global__ void cvt_f16_to_bf16_kernel(const __half* __restrict__ s, bf16* __restrict__ d, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) d[i] = __float2bfloat16_rn(__half2float(s[i]));
}
__global__ void cvt_bf16_to_f32_kernel(const bf16* __restrict__ s, float* __restrict__ d, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) d[i] = __bfloat162float(s[i]);
}
__global__ void cvt_f16_to_f32_kernel(const __half* __restrict__ s, float* __restrict__ d, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) d[i] = __half2float(s[i]);
}
// out[2j] = a[j], out[2j+1] = b[j]   (rows of length K)
__global__ void interleave_rows_kernel(const bf16* __restrict__ a, const bf16* __restrict__ b, bf16* __restrict__ out, size_t rows, size_t K) {
  const size_t n = rows * K;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / K, k = i % K;
    out[(2 * r) * K + k] = a[i];
    out[(2 * r + 1) * K + k] = b[i];
  }
}
// Conv1d weight [Co][Ci][k] -> window-GEMV form [Co][j*Ci + ci]
__global__ void repack_conv_kernel(const bf16* __restrict__ w, bf16* __restrict__ out, int Co, int Ci, int k) {
  const size_t n = (size_t)Co * Ci * k;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int j = (int)(i % k); const int ci = (int)((i / k) % Ci); const size_t co = i / ((size_t)k * Ci);
    out[co * ((size_t)k * Ci) + (size_t)j * Ci + ci] = w[i];
  }
}
// ConvTranspose1d weight [Ci][Co][k=2s] -> [(j*Co + co)][half*Ci + ci], half 0 = previous frame (tap j+s), half 1 = current frame (tap j)
__global__ void repack_convtr_kernel(const bf16* __restrict__ w, bf16* __restrict__ out, int Ci, int Co, int s) {
  const int k = 2 * s;
  const size_t n = (size_t)Ci * Co * k;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int tap = (int)(i % k); const int co = (int)((i / k) % Co); const int ci = (int)(i / ((size_t)k * Co));
    const int j = tap % s, half = (tap >= s) ? 0 : 1;
    out[((size_t)j * Co + co) * (2 * (size_t)Ci) + (size_t)half * Ci + ci] = w[i];
  }
}
// depthwise [C][1][7] fp32 -> [7][C];  generic small conv [Co][Ci][k] fp32 -> [Co][j*Ci+ci]
__global__ void repack_dw_kernel(const float* __restrict__ w, float* __restrict__ out, int C) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < C * 7) { const int c = i / 7, j = i % 7; out[j * C + c] = w[i]; }
}
__global__ void repack_conv_f32_kernel(const float* __restrict__ w, float* __restrict__ out, int Co, int Ci, int k) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < Co * Ci * k) { const int j = i % k, ci = (i / k) % Ci, co = i / (k * Ci); out[co * (k * Ci) + j * Ci + ci] = w[i]; }
}
__global__ void tile_bias_kernel(const float* __restrict__ b, float* __restrict__ out, int Co, int s) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < Co * s) out[i] = b[i % Co];
}
__global__ void gather_rows_kernel(const bf16* __restrict__ table, const int* __restrict__ ids, bf16* __restrict__ out, int H) {
  const bf16* r = table + (size_t)ids[blockIdx.x] * H;
  for (int k = threadIdx.x; k < H; k += blockDim.x) out[(size_t)blockIdx.x * H + k] = r[k];
}
// sinusoidal timestep features (diffusion_head.py:66-88): [n_steps][256] = [cos(t f_j) | sin(t f_j)]
__global__ void timestep_feat_kernel(const float* __restrict__ t, const float* __restrict__ freqs /*[128]*/, float* __restrict__ out, int n_steps) {
  const int i = blockIdx.x, j = threadIdx.x;   // 256 threads
  if (i >= n_steps) return;
  const float a = t[i] * freqs[j & 127];
  out[i * 256 + j] = (j < 128) ? cosf(a) : sinf(a);
}
// KV hand-off from a prefill: src [n_tokens][kv_heads][HD] bf16 -> pages
__global__ void kv_write_kernel(const bf16* __restrict__ k, const bf16* __restrict__ v, bf16* __restrict__ kpool, bf16* __restrict__ vpool,
                                const int* __restrict__ page_row, int kv_heads, long long pos0, long long n_tokens) {
  const long long t = blockIdx.x;
  if (t >= n_tokens) return;
  const long long pos = pos0 + t;
  const int page = page_row[pos / KV_PAGE];
  const int slot = (int)(pos % KV_PAGE);
  for (int i = threadIdx.x; i < kv_heads * HD; i += blockDim.x) {
    const int h = i / HD, d = i % HD;
    const size_t o = (((size_t)page * kv_heads + h) * KV_PAGE + slot) * HD + d;
    kpool[o] = k[(size_t)t * kv_heads * HD + i];
    vpool[o] = v[(size_t)t * kv_heads * HD + i];
  }
}

}  // namespace vv
