// vv_runtime.cu -- native runtime behind include/vibevoice_b200.h: weight registry + repacker,
// paged KV allocator, streaming-codec state slab, per-frame kernel programs (CUDA-graph cached).
//
// The reference owns none of this (it is eager PyTorch + HF DynamicCache + python dict caches:
// modeling_vibevoice_inference.py:367-695, modular_vibevoice_tokenizer.py:193-256); SURVEY 8b
// sketches the C ABI this file implements.
#include "../../include/vibevoice_b200.h"
#include "vv_kernels.cuh"
#include "vv_stream.cuh"

#include <cuda.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <initializer_list>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <set>
#include <string>
#include <vector>

using namespace vv;

static thread_local std::string g_err;
static int fail(int code, const char* fmt, ...) {
  char buf[2048];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}
#define CK(expr)                                                                                     \
  do {                                                                                               \
    cudaError_t _e = (expr);                                                                         \
    if (_e != cudaSuccess) return fail(VV_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
  } while (0)
#define CKL() CK(cudaGetLastError())
#define RET(expr)            \
  do {                       \
    int _r = (expr);         \
    if (_r < 0) return _r;   \
  } while (0)

struct RawTensor {
  void* p = nullptr;
  std::vector<int64_t> shape;
  bool is_f32 = false;
  size_t numel = 0;
};

struct Block {            // Block1D (tokenizer.py:620-684)
  int C = 0;
  float *norm_w = nullptr, *dw_w = nullptr, *dw_b = nullptr, *gamma = nullptr, *ffn_norm_w = nullptr, *b1 = nullptr, *b2 = nullptr,
        *ffn_gamma = nullptr;
  bf16 *w1 = nullptr, *w2 = nullptr;
  float *hist = nullptr, *next = nullptr;   // [B][6][C]
};
struct ConvL {             // stem / downsample / upsample / head conv in window-GEMV form
  int Cin = 0, Cout = 0, k = 0, stride = 1, ctx = 0, N = 0, K = 0;
  bf16* w = nullptr; float* wf = nullptr; float* bias = nullptr;
  float *hist = nullptr, *next = nullptr;   // [B][ctx][Cin]
};
struct Codec {
  std::vector<ConvL> convs;                  // index i = layer before stage i ; last = head
  std::vector<std::vector<Block>> stages;
  std::vector<int> T;                        // frames per stage (per 1 latent frame)
  std::vector<int> C;
  StateSeg* segs_dev = nullptr; int n_segs = 0;
  int64_t weight_bytes = 0;
};
struct LmLayer { bf16 *wqkv, *wo, *wgu, *wdown; float *bqkv, *ln1, *ln2; };
struct HeadLayer { bf16 *wgu, *wdown; float* norm; };

struct GraphEntry { cudaGraphExec_t exec = nullptr; int64_t launches = 0; };

struct vv_ctx {
  vv_model_desc d;
  int device = 0;
  bool finalized = false;
  bool use_graphs = true;
  bool use_pdl = true;
  bool use_mma_attn = true;
  bool use_splitk = true;
  bool fuse_rope = true;
  bf16* s_planes = nullptr; size_t planes_elems = 0;
  int mma_min_rows = 9;     // M >= this -> tensor-core GEMM (all prologues/epilogues), below -> weight-streaming GEMV (measured: at M = 8 the GEMV streams weights 1.7x faster)
  int use_tc5 = 1;          // tcgen05/TMEM GEMM: 0 off, 1 auto (wide GEMMs), 2 every M > 8 GEMM (VV_TC5)
  bool ring_rms = true;     // RMSNorm folded into gemm_mma_ring_kernel (row scale in the epilogue) instead of rows_norm_kernel (VV_NO_RING_RMS=1 -> off)
  int codec_mma_min_rows = 9;   // codec GEMMs with at least this many rows use the tensor-core ring kernel (VV_CODEC_MMA_MIN_ROWS)
  bool mma_ring = true;     // 6-stage cp.async ring for both GEMM operands (gemm_mma_ring_kernel); VV_NO_MMA_RING=1 -> old 1-ahead kernel
  bf16* head_slab = nullptr; size_t head_slab_bytes = 0; size_t l2_persist_bytes = 0; size_t l2_window_max = 0;
  int wr_tasks_min = 296;
  int wr_force = 0;
  int gemv_grid_cap = 0;
  int sm_count = 148;
  std::map<std::string, RawTensor> raw;
  std::set<std::string> expected;
  std::vector<void*> allocs;
  float speech_scale = NAN, speech_bias = NAN;
  // LM
  std::vector<LmLayer> lm;
  float* lm_norm = nullptr; bf16* embed = nullptr; const bf16* lm_head_w = nullptr; bf16* head_valid = nullptr; int* valid_ids_dev = nullptr; float* inv_freq = nullptr;
  int Nqkv = 0;
  // head
  bf16 *h_noisy = nullptr, *h_cond = nullptr, *h_t0 = nullptr, *h_t2 = nullptr, *h_mod = nullptr, *h_final = nullptr;
  std::vector<HeadLayer> head;
  int n_steps = 0; float* temb = nullptr; DpmCoef* coef_dev = nullptr; float* tfreqs = nullptr;
  std::vector<DpmCoef> coef_host; int coef_version = 0;
  bool sde = false; const float* step_noise = nullptr;   // sde-dpmsolver++: per-step variance noise [n_steps][B][64] (vv_set_step_noise)
  // connectors
  bf16 *ca_fc1 = nullptr, *ca_fc2 = nullptr, *cs_fc1 = nullptr, *cs_fc2 = nullptr;
  float *ca_b1 = nullptr, *ca_b2 = nullptr, *ca_n = nullptr, *cs_b1 = nullptr, *cs_b2 = nullptr, *cs_n = nullptr;
  Codec dec, enc;
  int64_t wbytes[6] = {0, 0, 0, 0, 0, 0};
  // KV
  int64_t n_pages = 0; int max_pages = 0; bf16 *kpool = nullptr, *vpool = nullptr;
  int* page_table_dev = nullptr; int* kv_len_dev = nullptr; int* row_mode_dev = nullptr;
  std::vector<int64_t> kv_len_host; std::vector<std::vector<int>> seq_pages; std::vector<int> free_pages;
  // scratch
  float* s_lgu = nullptr;   // LM gate/up raw sums [2B][2I] (weight-stream path)
  float *s_h = nullptr, *s_qkv = nullptr, *s_qrot = nullptr, *s_attn = nullptr, *s_act = nullptr, *s_pacc = nullptr, *s_pml = nullptr;
  int nsplit = 128;
  float *s_condp = nullptr, *s_call = nullptr, *s_mod = nullptr, *s_hx = nullptr, *s_hg = nullptr, *s_v = nullptr, *s_z = nullptr,
        *s_x0 = nullptr, *s_tfeat = nullptr, *s_t1 = nullptr, *s_hgu = nullptr;
  float *s_xa = nullptr, *s_xb = nullptr, *s_u = nullptr, *s_win = nullptr, *s_xn = nullptr;
  float *s_e = nullptr, *s_c1 = nullptr, *s_feat = nullptr, *s_audio = nullptr, *s_latent = nullptr;
  int* s_tok = nullptr;
  std::map<std::string, GraphEntry> graphs;
  GridBar* gridbar = nullptr;
  // weight-stream programs (vv_stream.cuh)
  struct StreamProg { SOp* ops = nullptr; int n_ops = 0; CUtensorMap* tmaps = nullptr; int n_stages = 0; int b_bytes = 0; int smem = 0; int gemv_ops = 0; int variant = 0; };
  std::map<std::string, StreamProg> sprogs;
  unsigned* st_bar = nullptr;            // grid-barrier counter of the stream kernel
  unsigned* st_diag_host = nullptr; unsigned* st_diag_dev = nullptr;   // host-mapped watchdog record
  int st_inflight = 4;                   // VV_STREAM_INFLIGHT: TMA tiles (16 KB) a CTA keeps in flight
  float* dec_front_x = nullptr;                  // where the streamed decoder front leaves its rows (same for every program: fixed structure)
  float *s_cx = nullptr, *s_cu = nullptr;        // codec rows / FFN hidden sums of the stream path (two buffers each)
  int use_stream = 15;                    // VV_STREAM bit 0: sampler, bit 1: LM linears, bit 2: + LM attention (whole decoder stack as one launch)
                                         // through the weight-stream kernel; 0 -> kernel-per-stage everywhere
  float* s_rope = nullptr;                        // [2B][64][2] cos / sin of the current positions
  float *s_pacc2 = nullptr, *s_pml2 = nullptr;   // attention partials of the stream path: [2B][kv_heads][SMs][8][128] / [..][8][2]
  std::map<const bf16*, bf16*> tiled; size_t tiled_bytes = 0;   // tile-major copies of the weights the stream kernel reads
  long long* st_trace2 = nullptr;
  long long* st_trace = nullptr; int st_trace_ops = 0; int st_trace_cta = 0; int st_trace_last_ops = 0;   // VV_STREAM_TRACE=<cta>: per-stage clock stamps of one CTA
  float* cfg_dev = nullptr; float cfg_last = NAN;   // CFG scale lives in device memory so captured graphs do not depend on its value
  int64_t launches = 0;
  std::map<long long, int> occ_cache;
};

struct L {  // launcher
  vv_ctx* c; cudaStream_t s;
  const void* win_base = nullptr;   // optional L2 access-policy window for this launch (persisting hits inside it)
  size_t win_bytes = 0;
  float win_ratio = 0.f;
};

// every hot-path kernel goes through here: programmatic dependent launch (PDL) lets kernel N+1 be scheduled and run its
// weight-only prologue while kernel N drains; inside CUDA-graph capture these become programmatic dependency edges.
template <typename... KArgs, typename... Args>
static cudaError_t launch_k(const L& l, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = l.s;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (l.c->use_pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (l.win_bytes) {
    attr[na].id = cudaLaunchAttributeAccessPolicyWindow;
    attr[na].val.accessPolicyWindow.base_ptr = const_cast<void*>(l.win_base);
    attr[na].val.accessPolicyWindow.num_bytes = l.win_bytes;
    attr[na].val.accessPolicyWindow.hitRatio = l.win_ratio;
    attr[na].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr[na].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  l.c->launches++;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

// ------------------------------------------------------------------------------------------------
template <class T>
static int dmalloc(vv_ctx* c, T** p, size_t n, bool zero = true) {
  void* q = nullptr;
  cudaError_t e = cudaMalloc(&q, std::max<size_t>(n, 1) * sizeof(T));
  if (e != cudaSuccess) return fail(VV_ERR_NOMEM, "cudaMalloc(%zu bytes): %s", n * sizeof(T), cudaGetErrorString(e));
  if (zero) cudaMemset(q, 0, std::max<size_t>(n, 1) * sizeof(T));
  c->allocs.push_back(q);
  *p = (T*)q;
  return 0;
}

static int gemv_smem_bytes(int MB, int K) { int Kp = (K + 255) & ~255; return (MB * Kp + 2 * 8 * 4 * MB) * 4; }

template <int MB>
static int launch_gemv_t(const L& l, GemvP& p, int grid, int smem) {
  static bool attr_set[8] = {false, false, false, false, false, false, false, false};
  (void)attr_set;
  CK(cudaFuncSetAttribute(gemv_kernel<MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CK(launch_k(l, gemv_kernel<MB>, dim3(grid), dim3(256), smem, p));
  return 0;
}

template <int MB>
static int gemv_occupancy(vv_ctx* c, int smem) {
  long long key = ((long long)MB << 32) | (unsigned)smem;
  auto it = c->occ_cache.find(key);
  if (it != c->occ_cache.end()) return it->second;
  int occ = 1;
  cudaFuncSetAttribute(gemv_kernel<MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, gemv_kernel<MB>, 256, smem) != cudaSuccess || occ < 1) occ = 1;
  c->occ_cache[key] = occ;
  return occ;
}

// y = epi(W * pro(x) + bias); dispatches GEMV (M <= 16) or the tiled GEMM.
static int linear(const L& l, GemvP p) {
  if (p.K % 8 != 0) return fail(VV_ERR_INVALID, "linear: K=%d not a multiple of 8", p.K);
  if (((uintptr_t)p.x & 15) || (p.xmap.rs & 3) || (p.xmap.bs & 3)) return fail(VV_ERR_INVALID, "linear: activation rows must be 16-byte aligned");
  // tcgen05/TMEM path: wide GEMMs that fill the chip with 128-row weight tiles (the all-steps AdaLN modulation GEMM:
  // [N_steps*2B, H] x [(3L+2)H, H]^T = 168 CTAs on 1.5B); mode 2 forces it for every M > 8 GEMM (tests)
  const int tc5_ctas = ((p.N + T5_BM - 1) / T5_BM) * ((p.M + T5_BN - 1) / T5_BN);
  if (l.c->use_tc5 && p.M > 8 && p.pro == PRO_NONE && p.epi != EPI_SWIGLU && (l.c->use_tc5 == 2 || tc5_ctas >= 96)) {
    if ((size_t)p.M * p.K > l.c->planes_elems) return fail(VV_ERR_INVALID, "tc5: activation planes scratch too small (%d x %d)", p.M, p.K);
    bf16* hi = l.c->s_planes;
    bf16* lo = l.c->s_planes + l.c->planes_elems;
    const long long n4 = (long long)p.M * (p.K >> 2);
    CK(launch_k(l, split_bf16_kernel, dim3((unsigned)std::min<long long>((n4 + 255) / 256, 1184)), dim3(256), 0, p.x, p.xmap, hi, lo, p.M, p.K));
    CK(cudaFuncSetAttribute(gemm_tc5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, T5_SMEM));
    CK(launch_k(l, gemm_tc5_kernel, dim3((p.N + T5_BM - 1) / T5_BM, (p.M + T5_BN - 1) / T5_BN), dim3(128), (size_t)T5_SMEM, p, (const bf16*)hi, (const bf16*)lo));
    return 0;
  }
  if (p.M >= l.c->mma_min_rows) {
    dim3 grid((p.N + MM_BN - 1) / MM_BN, (p.M + MM_BM - 1) / MM_BM);
    const bool inplace_res = (p.epi == EPI_RESID || p.epi == EPI_GAMMA_RESID || p.epi == EPI_GATED_RESID) && p.res == p.y && p.ldres == p.ldy;
    const int nk = (p.K + MM_BK - 1) / MM_BK;
    if (l.c->use_splitk && inplace_res && nk >= 8 && (int)(grid.x * grid.y) < l.c->sm_count) {
      int z = std::min(std::min(nk / 2, 16), (2 * l.c->sm_count) / (int)(grid.x * grid.y));
      grid.z = std::max(z, 1);
    }
    const bool ring_rms = p.pro == PRO_RMSNORM && grid.z == 1 && p.K <= MR_MAXK_NORM && p.K % 4 == 0;
    if (l.c->mma_ring && (p.pro == PRO_NONE || ring_rms) && p.epi != EPI_SWIGLU) {
      // the combinations the codec passes use get their own instantiation (FFN1: folded RMSNorm + GELU; FFN2 / transposed convs: plain or
      // gamma-residual, split-K or not); anything else runs the run-time-switched one
      void (*fn)(GemvP) = gemm_mma_ring_kernel<-1, -1>;
      if (!getenv("VV_RING_GENERIC")) {
        if (ring_rms && p.epi == EPI_GELU) fn = gemm_mma_ring_kernel<1, EPI_GELU>;
        else if (ring_rms && p.epi == EPI_NONE) fn = gemm_mma_ring_kernel<1, EPI_NONE>;
        else if (!ring_rms && p.epi == EPI_GAMMA_RESID) fn = gemm_mma_ring_kernel<0, EPI_GAMMA_RESID>;
        else if (!ring_rms && p.epi == EPI_NONE) fn = gemm_mma_ring_kernel<0, EPI_NONE>;
        else if (!ring_rms && p.epi == EPI_GELU) fn = gemm_mma_ring_kernel<0, EPI_GELU>;
        else if (!ring_rms && p.epi == EPI_RESID) fn = gemm_mma_ring_kernel<0, EPI_RESID>;
      }
      CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, MR_SMEM_NORM));
      CK(launch_k(l, fn, dim3(grid), dim3(128), (size_t)(ring_rms ? MR_SMEM_NORM : MR_SMEM), p));
      return 0;
    }
    CK(launch_k(l, gemm_mma_kernel, dim3(grid), dim3(128), 0, p));
    return 0;
  }
  int MB = p.M <= 1 ? 1 : (p.M <= 2 ? 2 : (p.M <= 4 ? 4 : 8));
  while (MB > 1 && gemv_smem_bytes(MB, p.K) > 200 * 1024) MB >>= 1;
  if (gemv_smem_bytes(MB, p.K) > 200 * 1024) return fail(VV_ERR_INVALID, "gemv: K=%d too large", p.K);
  int WR = 8;
  while (WR > 1 && (p.N + 4 * WR - 1) / (4 * WR) < l.c->wr_tasks_min) WR >>= 1;
  const int nchunks = (p.K + 255) / 256;
  while (WR < 8 && 8 / WR > nchunks) WR <<= 1;      // never more k-split warps than 256-element chunks
  if (l.c->wr_force) WR = l.c->wr_force;
  p.WK = 8 / WR;
  const int ntasks = (p.N + 4 * WR - 1) / (4 * WR);
  const int smem = gemv_smem_bytes(MB, p.K);
  int occ = MB == 1 ? gemv_occupancy<1>(l.c, smem) : MB == 2 ? gemv_occupancy<2>(l.c, smem) : MB == 4 ? gemv_occupancy<4>(l.c, smem)
                                                                                                          : gemv_occupancy<8>(l.c, smem);
  int grid = std::min(ntasks, l.c->sm_count * occ);
  if (l.c->gemv_grid_cap) grid = std::min(grid, l.c->gemv_grid_cap);
  switch (MB) {
    case 1: return launch_gemv_t<1>(l, p, grid, smem);
    case 2: return launch_gemv_t<2>(l, p, grid, smem);
    case 4: return launch_gemv_t<4>(l, p, grid, smem);
    default: return launch_gemv_t<8>(l, p, grid, smem);
  }
}

static GemvP mk(const bf16* W, const float* bias, const float* x, long long ldx, float* y, int ldy, int M, int N, int K) {
  GemvP p;
  memset(&p, 0, sizeof p);
  p.W = W; p.bias = bias; p.x = x; p.xmap = dense_rows(ldx); p.y = y; p.ldy = ldy; p.M = M; p.N = N; p.K = K;
  p.pro = PRO_NONE; p.epi = EPI_NONE; p.WK = 1;
  return p;
}

// ------------------------------------------------------------------------------------------------
// weight-stream programs (vv_stream.cuh): host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = (EncodeTiledFn)p;
  }
  return fn;
}
// tile-major weight copy [n_tiles][128][64] bf16 -> 2-D tensor map over [n_tiles*128][64], box = one 16 KB tile, 128-byte swizzle
static int make_weight_tmap(const bf16* T, long long n_tiles, CUtensorMap* out) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return fail(VV_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  const int N = (int)n_tiles, K = 64;
  const cuuint64_t dims[2] = {64, (cuuint64_t)n_tiles * 128};
  const cuuint64_t strides[1] = {128};
  const cuuint32_t box[2] = {64, 128};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)T, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(VV_ERR_CUDA, "cuTensorMapEncodeTiled([%d tiles x %d]) failed with %d", N, K, (int)r);
  return 0;
}
// tile-major copy of W [N][K] (cached per weight pointer; `fresh` = never cache: the caller's buffer may be re-used with other contents)
static int tiled_weight(vv_ctx* c, const bf16* W, int N, int K, bool fresh, bf16** out, long long* n_tiles) {
  if (K % 8 || ((uintptr_t)W & 15)) return fail(VV_ERR_INVALID, "stream: weight [%d x %d] must have K %% 8 == 0 and a 16-byte aligned base", N, K);
  const long long KB = (K + 63) / 64, R = (N + 127) / 128;
  *n_tiles = R * KB;
  if (!fresh) {
    auto it = c->tiled.find(W);
    if (it != c->tiled.end()) { *out = it->second; return 0; }
  }
  bf16* T = nullptr;
  RET(dmalloc(c, &T, (size_t)(*n_tiles) * 8192, false));
  const long long n_chunks = *n_tiles * 1024;
  tile_pack_kernel<<<(unsigned)std::min<long long>((n_chunks + 255) / 256, 148 * 32), 256>>>(W, T, N, K, (int)KB, n_chunks);
  CKL();
  CK(cudaDeviceSynchronize());
  if (!fresh) c->tiled[W] = T;
  c->tiled_bytes += (size_t)(*n_tiles) * 16384;
  *out = T;
  return 0;
}

struct StreamBuilder {
  vv_ctx* c;
  std::vector<SOp> ops;
  std::vector<CUtensorMap> tmaps;
  std::vector<int> tmap_of;        // op -> tensor map index (or -1)
  explicit StreamBuilder(vv_ctx* c_) : c(c_) {}
  SOp& push(int kind, bool sync) {
    SOp o;
    memset(&o, 0, sizeof o);
    o.kind = kind; o.sync_before = sync ? 1 : 0; o.nB = 16;
    ops.push_back(o); tmap_of.push_back(-1);
    return ops.back();
  }
  // y[m][n] (+)= alpha * (W x'[m] + bias); x' = pro(x)
  bool fresh_weights = false;
  std::vector<bf16*> owned;         // tile-major copies made with fresh_weights (freed by the caller)
  int kv_tmap = -1;                 // index of the K-pool tensor map (V-pool map follows) for SK_ATTN stages
  std::vector<int> needs_kv;        // ops whose att.tmap_k / tmap_v must be patched
  int use_kv_pool() {
    if (kv_tmap >= 0) return 0;
    const auto& d = c->d;
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return fail(VV_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t rows = (cuuint64_t)d.num_layers * c->n_pages * d.num_kv_heads * KV_PAGE;
    const cuuint64_t dims[2] = {(cuuint64_t)d.head_dim, rows};
    const cuuint64_t strides[1] = {(cuuint64_t)d.head_dim * 2};
    const cuuint32_t box[2] = {64, KV_PAGE};
    const cuuint32_t estr[2] = {1, 1};
    for (bf16* pool : {c->kpool, c->vpool}) {
      CUtensorMap tm;
      CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)pool, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return fail(VV_ERR_CUDA, "cuTensorMapEncodeTiled(KV pool) failed with %d", (int)r);
      tmaps.push_back(tm);
    }
    kv_tmap = (int)tmaps.size() - 2;
    return 0;
  }
  void fill_att(SAtt* a, int layer) {
    const auto& d = c->d;
    const size_t per_layer = (size_t)c->n_pages * d.num_kv_heads * KV_PAGE * d.head_dim;
    memset(a, 0, sizeof *a);
    a->hd = d.head_dim;
    a->qkv = c->s_qkv;
    a->kv.kpool = c->kpool + per_layer * layer; a->kv.vpool = c->vpool + per_layer * layer;
    a->kv.page_table = c->page_table_dev; a->kv.max_pages = c->max_pages; a->kv.kv_len = c->kv_len_dev; a->kv.row_mode = c->row_mode_dev;
    a->kv.kv_heads = d.num_kv_heads; a->kv.q_heads = d.num_q_heads;
    a->part_acc = c->s_pacc2; a->part_ml = c->s_pml2; a->inv_freq = c->inv_freq; a->scale = 1.0f / sqrtf((float)d.head_dim);
    a->row_base = (unsigned)((size_t)layer * c->n_pages * d.num_kv_heads * KV_PAGE);
    a->rope_cs = c->s_rope;
  }
  int attn(int layer, int M) {
    RET(use_kv_pool());
    SOp& o = push(SK_ATTN, true);
    o.M = M;
    fill_att(&o.att, layer);
    needs_kv.push_back((int)ops.size() - 1);
    return 0;
  }
  int gemv(const bf16* W, const float* bias, const float* x, long long ldx, float* y, long long ldy, int M, int N, int K, bool sync, SOp** out) {
    if (M < 1 || M > 32) return fail(VV_ERR_INVALID, "stream gemv: M=%d outside [1,32]", M);
    CUtensorMap tm;
    bf16* T; long long n_tiles;
    RET(tiled_weight(c, W, N, K, fresh_weights, &T, &n_tiles));
    if (fresh_weights) owned.push_back(T);
    RET(make_weight_tmap(T, n_tiles, &tm));
    SOp& o = push(SK_GEMV, sync);
    o.M = M; o.N = N; o.K = K; o.nB = M <= 8 ? 16 : (M <= 16 ? 32 : 64);
    o.x = x; o.ldx = ldx; o.y = y; o.ldy = ldy; o.bias = bias; o.pro = SP_NONE; o.alpha_kind = SA_ONE;
    tmaps.push_back(tm);
    tmap_of.back() = (int)tmaps.size() - 1;
    *out = &ops.back();
    return 0;
  }
  void nop(bool sync, float* init_dst, long long init_n) {
    SOp& o = push(SK_NOP, sync);
    o.init_dst = init_dst; o.init_n = init_n;
  }
};

// stream_kernel instantiations: one per program family, each compiled with only the stage kinds / prologues / scalings the family uses
// (instruction-cache footprint, see the template's comment); a program runs on the first variant whose feature set covers it.
constexpr unsigned sfeat(std::initializer_list<int> pros, std::initializer_list<int> kinds, std::initializer_list<int> alphas) {
  unsigned f = 0;
  for (int x : pros) f |= 1u << x;
  for (int x : kinds) f |= 1u << (16 + x);
  for (int x : alphas) f |= 1u << (24 + x);
  return f;
}
constexpr unsigned SF_SAMP = sfeat({SP_NONE, SP_ADALN, SP_SWIGLU, SP_DPM}, {SK_GEMV, SK_NOP}, {SA_ONE, SA_GATE});
constexpr unsigned SF_LM = sfeat({SP_NONE, SP_RMSNORM, SP_SWIGLU, SP_COMBINE}, {SK_GEMV, SK_NOP, SK_ATTN}, {SA_ONE});
constexpr unsigned SF_CODEC = sfeat({SP_NONE, SP_WINDOW, SP_RMSNORM, SP_GELU}, {SK_GEMV, SK_NOP, SK_MIX}, {SA_ONE, SA_GAMMA});
constexpr unsigned SF_HD128 = 1u << 30;          // every attention stage has head_dim 128 (compile-time loop bounds)
constexpr unsigned SF_LM128 = SF_LM | SF_HD128;
constexpr unsigned SF_NB16 = 1u << 29;           // every linear stage has a 16-row activation operand (compile-time nB)
constexpr unsigned SF_ALL = 0xffffffffu & ~SF_HD128 & ~SF_NB16;
typedef void (*StreamFn)(SParams);
static const struct { unsigned feat; StreamFn fn; StreamFn fn_trace; const char* name; } STREAM_VARIANTS[] = {
#define SVAR(f, name) {f, stream_kernel<f, false>, stream_kernel<f, true>, name}
  SVAR(SF_SAMP | SF_NB16, "sampler (16-row operand)"), SVAR(SF_SAMP, "sampler"),
  SVAR(SF_LM128 | SF_NB16, "lm (head_dim 128, 16-row operand)"), SVAR(SF_LM128, "lm (head_dim 128)"), SVAR(SF_LM, "lm"),
  SVAR(SF_CODEC | SF_NB16, "codec (16-row operand)"), SVAR(SF_CODEC, "codec"), SVAR(SF_ALL, "all")};
#undef SVAR
constexpr int N_STREAM_VARIANTS = 8;

static int finish_stream(StreamBuilder& b, vv_ctx::StreamProg* pr) {
  vv_ctx* c = b.c;
  const int G = c->sm_count;
  int b_bytes = 2048;
  for (const SOp& o : b.ops) {
    if (o.kind == SK_MIX && (o.cod.T_out > 8 || o.K % 4 || o.K > 4096)) return fail(VV_ERR_INVALID, "stream: mixer stage handles T <= 8, C <= 4096");
    if (o.kind == SK_ATTN) b_bytes = std::max(b_bytes, 32768);        // Q tile, new K/V row and the warp-merge buffers live in the operand region
    if (o.kind != SK_GEMV) continue;
    const long long KB = (o.K + 63) / 64, R = (o.N + 127) / 128, U = R * KB;
    const long long per = (U + G - 1) / G;
    const long long count = std::min(per, KB);
    b_bytes = std::max<long long>(b_bytes, count * o.nB * 128);
    const long long segs = (per + KB - 1) / KB + 1;
    if (segs > ST_MAXSEG || segs * o.nB > 512) return fail(VV_ERR_INVALID, "stream: stage [%d x %d] needs %lld accumulators per CTA", o.N, o.K, segs);
    if (o.store && KB != 1) return fail(VV_ERR_INVALID, "stream: store epilogue needs K <= 64");
    if (o.pro == SP_WINDOW && (o.cod.cin % 8)) return fail(VV_ERR_INVALID, "stream: window prologue needs a channel count that is a multiple of 8");
    if (o.pro == SP_COMBINE) {
      const long long nh = count / 2 + 2;
      if ((long long)o.M * nh > 128 || KB != (c->d.head_dim / 64) * c->d.num_q_heads)
        return fail(VV_ERR_INVALID, "stream: attention-merge prologue does not fit (M=%d, %lld k-blocks per CTA)", o.M, count);
      b_bytes = std::max<long long>(b_bytes, ((count * o.nB * 128 + 1023) & ~1023ll) + o.M * nh * G * 4 + 16 + std::max<long long>(2048, o.M * count * 256) +
                                                 o.M * count * 256);
    }
    if (U * (G + 1) >= (1ll << 32)) return fail(VV_ERR_INVALID, "stream: stage [%d x %d] has too many tiles for 32-bit scheduling", o.N, o.K);
  }
  unsigned feat = 0;
  bool hd128 = true, nb16 = true;
  for (const SOp& o : b.ops) {
    if (o.kind == SK_GEMV && o.nB != 16) nb16 = false;
    feat |= (1u << o.pro) | (1u << (16 + o.kind)) | (1u << (24 + o.alpha_kind));
    if ((o.kind == SK_ATTN || o.pro == SP_COMBINE || o.rope_rows > 0) && o.att.hd != 128) hd128 = false;
  }
  pr->variant = N_STREAM_VARIANTS - 1;
  if (!getenv("VV_STREAM_GENERIC"))
    for (int v = 0; v < N_STREAM_VARIANTS; ++v)
      if ((feat & ~STREAM_VARIANTS[v].feat) == 0 && (hd128 || !(STREAM_VARIANTS[v].feat & SF_HD128)) && (nb16 || !(STREAM_VARIANTS[v].feat & SF_NB16)) &&
          !(getenv("VV_STREAM_NO_NB16") && (STREAM_VARIANTS[v].feat & SF_NB16))) { pr->variant = v; break; }
  if (getenv("VV_VERBOSE")) fprintf(stderr, "[vv] stream program: %zu stages, features %08x -> kernel variant '%s'\n", b.ops.size(), feat, STREAM_VARIANTS[pr->variant].name);
  cudaFuncAttributes fa;
  CK(cudaFuncGetAttributes(&fa, STREAM_VARIANTS[pr->variant].fn));
  const int max_dyn = 232448 - (int)fa.sharedSizeBytes - 256;
  int ns = (max_dyn - 1024 - b_bytes) / ST_TILE;
  ns = std::min(ns, ST_MAX_STAGES);
  if (getenv("VV_STREAM_STAGES")) ns = std::min(ns, atoi(getenv("VV_STREAM_STAGES")));
  if (ns < 2) return fail(VV_ERR_INVALID, "stream: activation operand of %d bytes leaves no room for the weight ring", b_bytes);
  pr->n_stages = ns; pr->b_bytes = b_bytes; pr->smem = ns * ST_TILE + b_bytes + 1024;
  pr->n_ops = (int)b.ops.size();
  RET(dmalloc(c, &pr->tmaps, std::max<size_t>(b.tmaps.size(), 1), false));
  RET(dmalloc(c, &pr->ops, b.ops.size(), false));
  for (int i : b.needs_kv) {
    b.ops[i].att.tmap_k = (unsigned long long)(uintptr_t)(pr->tmaps + b.kv_tmap);
    b.ops[i].att.tmap_v = (unsigned long long)(uintptr_t)(pr->tmaps + b.kv_tmap + 1);
  }
  for (size_t i = 0; i < b.ops.size(); ++i) {
    if (b.tmap_of[i] >= 0) b.ops[i].tmap = (unsigned long long)(uintptr_t)(pr->tmaps + b.tmap_of[i]);
    pr->gemv_ops += b.ops[i].kind == SK_GEMV;
  }
  if (!b.tmaps.empty()) CK(cudaMemcpy(pr->tmaps, b.tmaps.data(), b.tmaps.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(pr->ops, b.ops.data(), b.ops.size() * sizeof(SOp), cudaMemcpyHostToDevice));
  CK(cudaFuncSetAttribute(STREAM_VARIANTS[pr->variant].fn, cudaFuncAttributeMaxDynamicSharedMemorySize, max_dyn));
  CK(cudaFuncSetAttribute(STREAM_VARIANTS[pr->variant].fn_trace, cudaFuncAttributeMaxDynamicSharedMemorySize, max_dyn));
  return 0;
}

static int launch_stream(const L& l, const vv_ctx::StreamProg& pr, int op_begin = 0, int op_count = -1) {
  vv_ctx* c = l.c;
  if (op_count < 0) op_count = pr.n_ops - op_begin;
  CK(cudaMemsetAsync(c->st_bar, 0, sizeof(unsigned), l.s));
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.gridDim = dim3(c->sm_count); cfg.blockDim = dim3(ST_THREADS); cfg.dynamicSmemBytes = pr.smem; cfg.stream = l.s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;      // all CTAs must be co-resident: they synchronise through a grid barrier
  attr[0].val.cooperative = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  SParams P;
  P.ops = pr.ops + op_begin; P.n_ops = op_count; P.bar_count = c->st_bar; P.diag = c->st_diag_dev; P.n_stages = pr.n_stages; P.b_bytes = pr.b_bytes;
  P.max_inflight = std::max(1, std::min(pr.n_stages, c->st_inflight));
  P.kv_len = c->kv_len_dev; P.row_mode = c->row_mode_dev; P.n_seq = 2 * c->d.max_batch; P.kv_heads = c->d.num_kv_heads;
  P.trace = (c->st_trace && op_count == pr.n_ops && pr.n_ops <= c->st_trace_ops) ? c->st_trace : nullptr; P.trace_cta = c->st_trace_cta;
  P.trace2 = P.trace ? c->st_trace2 : nullptr;
  if (P.trace) c->st_trace_last_ops = pr.n_ops;
  c->launches++;
  CK(cudaLaunchKernelEx(&cfg, P.trace ? STREAM_VARIANTS[pr.variant].fn_trace : STREAM_VARIANTS[pr.variant].fn, P));
  return 0;
}

// ------------------------------------------------------------------------------------------------
// expected tensor names (same as vibevoice_b200/synth.py::param_specs, i.e. the HF checkpoint keys)
// ------------------------------------------------------------------------------------------------
static std::string S(const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  return buf;
}
static void block_names(std::set<std::string>& e, const std::string& p) {
  for (const char* s : {"norm.weight", "mixer.conv.conv.conv.weight", "mixer.conv.conv.conv.bias", "gamma", "ffn_norm.weight",
                        "ffn.linear1.weight", "ffn.linear1.bias", "ffn.linear2.weight", "ffn.linear2.bias", "ffn_gamma"})
    e.insert(p + "." + s);
}
static void build_expected(vv_ctx* c) {
  auto& e = c->expected;
  const auto& d = c->d;
  const std::string lm = "model.language_model";
  e.insert(lm + ".embed_tokens.weight");
  e.insert(lm + ".norm.weight");
  for (int l = 0; l < d.num_layers; ++l) {
    std::string q = S("%s.layers.%d", lm.c_str(), l);
    for (const char* s : {"self_attn.q_proj.weight", "self_attn.q_proj.bias", "self_attn.k_proj.weight", "self_attn.k_proj.bias",
                          "self_attn.v_proj.weight", "self_attn.v_proj.bias", "self_attn.o_proj.weight", "mlp.gate_proj.weight",
                          "mlp.up_proj.weight", "mlp.down_proj.weight", "input_layernorm.weight", "post_attention_layernorm.weight"})
      e.insert(q + "." + s);
  }
  if (!d.tie_word_embeddings) e.insert("lm_head.weight");
  const std::string h = "model.prediction_head";
  for (const char* s : {"noisy_images_proj.weight", "cond_proj.weight", "t_embedder.mlp.0.weight", "t_embedder.mlp.2.weight",
                        "final_layer.linear.weight", "final_layer.adaLN_modulation.1.weight"})
    e.insert(h + "." + s);
  for (int l = 0; l < d.head_layers; ++l)
    for (const char* s : {"ffn.gate_proj.weight", "ffn.up_proj.weight", "ffn.down_proj.weight", "norm.weight", "adaLN_modulation.1.weight"})
      e.insert(S("%s.layers.%d.%s", h.c_str(), l, s));
  for (const char* cn : {"model.acoustic_connector", "model.semantic_connector"})
    for (const char* s : {"fc1.weight", "fc1.bias", "norm.weight", "fc2.weight", "fc2.bias"}) e.insert(std::string(cn) + "." + s);
  const std::string dp = "model.acoustic_tokenizer.decoder", ep = "model.semantic_tokenizer.encoder";
  for (int i = 0; i < d.n_stages; ++i) {
    if (i == 0) { e.insert(dp + ".upsample_layers.0.0.conv.conv.weight"); e.insert(dp + ".upsample_layers.0.0.conv.conv.bias"); }
    else { e.insert(S("%s.upsample_layers.%d.0.convtr.convtr.weight", dp.c_str(), i)); e.insert(S("%s.upsample_layers.%d.0.convtr.convtr.bias", dp.c_str(), i)); }
    for (int j = 0; j < d.dec_depths[i]; ++j) block_names(e, S("%s.stages.%d.%d", dp.c_str(), i, j));
    e.insert(S("%s.downsample_layers.%d.0.conv.conv.weight", ep.c_str(), i));
    e.insert(S("%s.downsample_layers.%d.0.conv.conv.bias", ep.c_str(), i));
    for (int j = 0; j < d.enc_depths[i]; ++j) block_names(e, S("%s.stages.%d.%d", ep.c_str(), i, j));
  }
  for (const char* s : {"head.conv.conv.weight", "head.conv.conv.bias"}) { e.insert(dp + "." + s); e.insert(ep + "." + s); }
}

// ------------------------------------------------------------------------------------------------
extern "C" int vv_abi_version(void) { return VV_ABI_VERSION; }
extern "C" const char* vv_last_error(void) { return g_err.c_str(); }

extern "C" int vv_create(const vv_model_desc* desc, int device, vv_ctx** out) {
  if (!desc || !out) return fail(VV_ERR_INVALID, "vv_create: null argument");
  if (desc->head_dim != 128 && desc->head_dim != 64) return fail(VV_ERR_INVALID, "head_dim %d unsupported (64 or 128)", desc->head_dim);
  if (desc->max_batch < 1 || desc->max_batch > 8) return fail(VV_ERR_INVALID, "max_batch must be in [1,8]");
  if (desc->n_stages < 2 || desc->n_stages > 8) return fail(VV_ERR_INVALID, "n_stages out of range");
  if (desc->num_q_heads % desc->num_kv_heads || desc->num_q_heads / desc->num_kv_heads > ATT_MAXG)
    return fail(VV_ERR_INVALID, "GQA group size unsupported");
  if (desc->latent_size != 64 || desc->acoustic_vae_dim != 64) return fail(VV_ERR_INVALID, "latent size must be 64");
  if (desc->n_valid_ids < 1 || desc->n_valid_ids > 8) return fail(VV_ERR_INVALID, "n_valid_ids out of range");
  CK(cudaSetDevice(device));
  vv_ctx* c = new vv_ctx();
  c->d = *desc;
  c->device = device;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  c->sm_count = prop.multiProcessorCount;
  const char* ng = getenv("VV_NO_GRAPH");
  c->use_graphs = !(ng && ng[0] == '1');
  const char* na = getenv("VV_SCALAR_ATTN");
  c->use_mma_attn = !(na && na[0] == '1');
  c->wr_tasks_min = c->sm_count;      // measured (tools/bench_gemv.py): one task per SM beats two for the N=1536 shapes, neutral elsewhere
  if (getenv("VV_WR_TASKS_MIN")) c->wr_tasks_min = atoi(getenv("VV_WR_TASKS_MIN"));
  if (getenv("VV_WR_FORCE")) c->wr_force = atoi(getenv("VV_WR_FORCE"));
  if (getenv("VV_GEMV_GRID_CAP")) c->gemv_grid_cap = atoi(getenv("VV_GEMV_GRID_CAP"));
  if (getenv("VV_NO_FUSE_ROPE")) c->fuse_rope = false;
  if (getenv("VV_STREAM")) c->use_stream = atoi(getenv("VV_STREAM"));
  if (getenv("VV_STREAM_INFLIGHT")) c->st_inflight = atoi(getenv("VV_STREAM_INFLIGHT"));
  if (getenv("VV_TC5")) c->use_tc5 = atoi(getenv("VV_TC5"));
  if (getenv("VV_MMA_MIN_ROWS")) c->mma_min_rows = atoi(getenv("VV_MMA_MIN_ROWS"));
  if (getenv("VV_NO_MMA_RING")) c->mma_ring = false;
  if (getenv("VV_NO_RING_RMS")) c->ring_rms = false;
  if (getenv("VV_CODEC_MMA_MIN_ROWS")) c->codec_mma_min_rows = atoi(getenv("VV_CODEC_MMA_MIN_ROWS"));
  else if (getenv("VV_MMA_MIN_ROWS")) c->codec_mma_min_rows = c->mma_min_rows;
  const char* ns = getenv("VV_NO_SPLITK");
  c->use_splitk = !(ns && ns[0] == '1');
  const char* np = getenv("VV_NO_PDL");
  c->use_pdl = !(np && np[0] == '1');
  build_expected(c);
  *out = c;
  return 0;
}

extern "C" void vv_destroy(vv_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  for (auto& g : c->graphs) if (g.second.exec) cudaGraphExecDestroy(g.second.exec);
  for (auto& r : c->raw) if (r.second.p) cudaFree(r.second.p);
  for (void* p : c->allocs) cudaFree(p);
  if (c->st_diag_host) cudaFreeHost(c->st_diag_host);
  delete c;
}

static bool stored_as_f32(const std::string& name, const int64_t* shape, int ndim) {
  if (ndim <= 1) return true;
  if (name.find("mixer.conv.conv.conv.weight") != std::string::npos) return true;                 // depthwise [C,1,7]
  if (ndim == 3 && (shape[0] == 1 || shape[1] == 1)) return true;                                  // 1->32 stem, 32->1 head
  return false;
}

extern "C" int vv_load_tensor(vv_ctx* c, const char* name_, const void* data, int dtype, const int64_t* shape, int ndim) {
  if (!c || !name_ || !data) return fail(VV_ERR_INVALID, "vv_load_tensor: null argument");
  if (c->finalized) return fail(VV_ERR_STATE, "vv_load_tensor after vv_finalize_weights");
  std::string name(name_);
  CK(cudaSetDevice(c->device));
  if (name == "model.speech_scaling_factor" || name == "model.speech_bias_factor") {
    float v; char tmp[4];
    CK(cudaMemcpy(tmp, data, dtype == VV_DT_F32 ? 4 : 2, cudaMemcpyDefault));
    if (dtype == VV_DT_F32) memcpy(&v, tmp, 4);
    else if (dtype == VV_DT_BF16) { unsigned u = ((unsigned)(*(unsigned short*)tmp)) << 16; memcpy(&v, &u, 4); }
    else v = __half2float(*(__half*)tmp);
    (name == "model.speech_scaling_factor" ? c->speech_scale : c->speech_bias) = v;
    return 0;
  }
  if (!c->expected.count(name)) {
    if (name.rfind("model.acoustic_tokenizer.encoder.", 0) == 0 || name.find("fix_std") != std::string::npos ||
        name.find("rotary_emb") != std::string::npos || (name == "lm_head.weight" && c->d.tie_word_embeddings))
      return 1;   // not on this path
    return fail(VV_ERR_INVALID, "vv_load_tensor: unknown tensor '%s'", name.c_str());
  }
  size_t n = 1;
  for (int i = 0; i < ndim; ++i) n *= (size_t)shape[i];
  RawTensor t;
  t.shape.assign(shape, shape + ndim);
  t.numel = n;
  t.is_f32 = stored_as_f32(name, shape, ndim);
  const size_t esz = dtype == VV_DT_F32 ? 4 : 2;
  void* dst = nullptr;
  CK(cudaMalloc(&dst, n * (t.is_f32 ? 4 : 2)));
  t.p = dst;
  const bool same = (t.is_f32 && dtype == VV_DT_F32) || (!t.is_f32 && dtype == VV_DT_BF16);
  if (same) {
    CK(cudaMemcpy(dst, data, n * esz, cudaMemcpyDefault));
  } else {
    void* tmp = nullptr;
    CK(cudaMalloc(&tmp, n * esz));
    CK(cudaMemcpy(tmp, data, n * esz, cudaMemcpyDefault));
    const int grid = (int)std::min<size_t>((n + 255) / 256, 65535);
    if (t.is_f32 && dtype == VV_DT_BF16) cvt_bf16_to_f32_kernel<<<grid, 256>>>((const bf16*)tmp, (float*)dst, n);
    else if (t.is_f32 && dtype == VV_DT_F16) cvt_f16_to_f32_kernel<<<grid, 256>>>((const __half*)tmp, (float*)dst, n);
    else if (!t.is_f32 && dtype == VV_DT_F32) cvt_f32_to_bf16_kernel<<<grid, 256>>>((const float*)tmp, (bf16*)dst, n);
    else cvt_f16_to_bf16_kernel<<<grid, 256>>>((const __half*)tmp, (bf16*)dst, n);
    CKL();
    CK(cudaDeviceSynchronize());
    cudaFree(tmp);
  }
  auto it = c->raw.find(name);
  if (it != c->raw.end() && it->second.p) cudaFree(it->second.p);
  c->raw[name] = t;
  return 0;
}

extern "C" int vv_set_speech_factors(vv_ctx* c, float s, float b) {
  if (!c) return fail(VV_ERR_INVALID, "null ctx");
  c->speech_scale = s; c->speech_bias = b;
  return 0;
}

// ---- repack helpers ------------------------------------------------------------------------------
static int need(vv_ctx* c, const std::string& name, RawTensor** t, std::vector<int64_t> shape) {
  auto it = c->raw.find(name);
  if (it == c->raw.end()) return fail(VV_ERR_STATE, "missing tensor %s", name.c_str());
  if (it->second.shape != shape) {
    std::string a, b;
    for (auto v : it->second.shape) a += std::to_string(v) + ",";
    for (auto v : shape) b += std::to_string(v) + ",";
    return fail(VV_ERR_INVALID, "tensor %s has shape [%s] expected [%s]", name.c_str(), a.c_str(), b.c_str());
  }
  *t = &it->second;
  return 0;
}
static int take_bf16(vv_ctx* c, const std::string& name, std::vector<int64_t> shape, bf16** out, int64_t* bytes) {
  RawTensor* t;
  RET(need(c, name, &t, shape));
  if (t->is_f32) return fail(VV_ERR_INVALID, "%s stored as f32, expected bf16", name.c_str());
  *out = (bf16*)t->p;
  c->allocs.push_back(t->p);
  t->p = nullptr;
  if (bytes) *bytes += (int64_t)t->numel * 2;
  return 0;
}
static int take_f32(vv_ctx* c, const std::string& name, std::vector<int64_t> shape, float** out, int64_t* bytes) {
  RawTensor* t;
  RET(need(c, name, &t, shape));
  if (!t->is_f32) return fail(VV_ERR_INVALID, "%s stored as bf16, expected f32", name.c_str());
  *out = (float*)t->p;
  c->allocs.push_back(t->p);
  t->p = nullptr;
  if (bytes) *bytes += (int64_t)t->numel * 2;   // algorithmic bytes are quoted at bf16
  return 0;
}
static void drop(vv_ctx* c, const std::string& name) {
  auto it = c->raw.find(name);
  if (it != c->raw.end() && it->second.p) { cudaFree(it->second.p); it->second.p = nullptr; }
}

static int build_block(vv_ctx* c, const std::string& p, int C, Block* b, int64_t* bytes, std::vector<StateSeg>* segs) {
  b->C = C;
  RET(take_f32(c, p + ".norm.weight", {C}, &b->norm_w, bytes));
  RawTensor* t;
  RET(need(c, p + ".mixer.conv.conv.conv.weight", &t, {C, 1, 7}));
  RET(dmalloc(c, &b->dw_w, (size_t)7 * C));
  repack_dw_kernel<<<(C * 7 + 255) / 256, 256>>>((const float*)t->p, b->dw_w, C);
  CKL();
  *bytes += (int64_t)C * 7 * 2;
  RET(take_f32(c, p + ".mixer.conv.conv.conv.bias", {C}, &b->dw_b, bytes));
  RET(take_f32(c, p + ".gamma", {C}, &b->gamma, bytes));
  RET(take_f32(c, p + ".ffn_norm.weight", {C}, &b->ffn_norm_w, bytes));
  RET(take_bf16(c, p + ".ffn.linear1.weight", {4 * C, C}, &b->w1, bytes));
  RET(take_f32(c, p + ".ffn.linear1.bias", {4 * C}, &b->b1, bytes));
  RET(take_bf16(c, p + ".ffn.linear2.weight", {C, 4 * C}, &b->w2, bytes));
  RET(take_f32(c, p + ".ffn.linear2.bias", {C}, &b->b2, bytes));
  RET(take_f32(c, p + ".ffn_gamma", {C}, &b->ffn_gamma, bytes));
  const int B = c->d.max_batch;
  RET(dmalloc(c, &b->hist, (size_t)B * 6 * C));
  RET(dmalloc(c, &b->next, (size_t)B * 6 * C));
  segs->push_back({b->hist, b->next, 6 * C});
  return 0;
}

// Conv1d [Co][Ci][k] -> window GEMV weight [Co][k*Ci]
static int build_conv(vv_ctx* c, const std::string& p, int Ci, int Co, int k, int stride, ConvL* L_, int64_t* bytes,
                      std::vector<StateSeg>* segs) {
  L_->Cin = Ci; L_->Cout = Co; L_->k = k; L_->stride = stride; L_->ctx = (k - 1) - (stride - 1); L_->N = Co; L_->K = k * Ci;
  RawTensor* t;
  RET(need(c, p + ".weight", &t, {Co, Ci, k}));
  const size_t n = (size_t)Co * Ci * k;
  if (t->is_f32) {
    RET(dmalloc(c, &L_->wf, n));
    repack_conv_f32_kernel<<<(int)((n + 255) / 256), 256>>>((const float*)t->p, L_->wf, Co, Ci, k);
  } else {
    RET(dmalloc(c, &L_->w, n));
    repack_conv_kernel<<<(int)std::min<size_t>((n + 255) / 256, 65535), 256>>>((const bf16*)t->p, L_->w, Co, Ci, k);
  }
  CKL();
  *bytes += (int64_t)n * 2;
  RET(take_f32(c, p + ".bias", {Co}, &L_->bias, bytes));
  const int B = c->d.max_batch;
  RET(dmalloc(c, &L_->hist, (size_t)B * L_->ctx * Ci));
  RET(dmalloc(c, &L_->next, (size_t)B * L_->ctx * Ci));
  segs->push_back({L_->hist, L_->next, L_->ctx * Ci});
  return 0;
}
// ConvTranspose1d [Ci][Co][2s] -> [(j,co)][(half,ci)], bias tiled over j
static int build_convtr(vv_ctx* c, const std::string& p, int Ci, int Co, int s, ConvL* L_, int64_t* bytes, std::vector<StateSeg>* segs) {
  L_->Cin = Ci; L_->Cout = Co; L_->k = 2 * s; L_->stride = s; L_->ctx = 1; L_->N = s * Co; L_->K = 2 * Ci;
  RawTensor* t;
  RET(need(c, p + ".weight", &t, {Ci, Co, 2 * s}));
  if (t->is_f32) return fail(VV_ERR_INVALID, "%s: unexpected f32 convtr weight", p.c_str());
  const size_t n = (size_t)Ci * Co * 2 * s;
  RET(dmalloc(c, &L_->w, n));
  repack_convtr_kernel<<<(int)std::min<size_t>((n + 255) / 256, 65535), 256>>>((const bf16*)t->p, L_->w, Ci, Co, s);
  CKL();
  *bytes += (int64_t)n * 2;
  RawTensor* bt;
  RET(need(c, p + ".bias", &bt, {Co}));
  RET(dmalloc(c, &L_->bias, (size_t)s * Co));
  tile_bias_kernel<<<(s * Co + 255) / 256, 256>>>((const float*)bt->p, L_->bias, Co, s);
  CKL();
  *bytes += (int64_t)Co * 2;
  const int B = c->d.max_batch;
  RET(dmalloc(c, &L_->hist, (size_t)B * Ci));
  RET(dmalloc(c, &L_->next, (size_t)B * Ci));
  segs->push_back({L_->hist, L_->next, Ci});
  return 0;
}

static int upload_segs(vv_ctx* c, Codec* k, const std::vector<StateSeg>& segs) {
  k->n_segs = (int)segs.size();
  RET(dmalloc(c, &k->segs_dev, segs.size()));
  CK(cudaMemcpy(k->segs_dev, segs.data(), segs.size() * sizeof(StateSeg), cudaMemcpyHostToDevice));
  return 0;
}

extern "C" int vv_finalize_weights(vv_ctx* c) {
  if (!c) return fail(VV_ERR_INVALID, "null ctx");
  if (c->finalized) return fail(VV_ERR_STATE, "already finalized");
  CK(cudaSetDevice(c->device));
  {
    std::string missing; int nm = 0;
    for (auto& n : c->expected) if (!c->raw.count(n)) { if (nm < 8) missing += n + " "; ++nm; }
    if (nm) return fail(VV_ERR_STATE, "%d tensors missing, e.g. %s", nm, missing.c_str());
    if (std::isnan(c->speech_scale) || std::isnan(c->speech_bias))
      return fail(VV_ERR_STATE, "speech_scaling_factor / speech_bias_factor are NaN (random-init checkpoints: call vv_set_speech_factors)");
  }
  const auto& d = c->d;
  const int H = d.hidden_size, I = d.intermediate_size, nq = d.num_q_heads * d.head_dim, nkv = d.num_kv_heads * d.head_dim, B = d.max_batch;
  const int M2 = 2 * B;
  // ---------------- LM ----------------
  const std::string lm = "model.language_model";
  c->Nqkv = nq + 2 * nkv;
  c->lm.resize(d.num_layers);
  int64_t* wb = &c->wbytes[0];
  for (int l = 0; l < d.num_layers; ++l) {
    LmLayer& y = c->lm[l];
    std::string q = S("%s.layers.%d", lm.c_str(), l);
    RawTensor *tq, *tk, *tv, *bq, *bk, *bv, *tg, *tu;
    RET(need(c, q + ".self_attn.q_proj.weight", &tq, {nq, H}));
    RET(need(c, q + ".self_attn.k_proj.weight", &tk, {nkv, H}));
    RET(need(c, q + ".self_attn.v_proj.weight", &tv, {nkv, H}));
    RET(need(c, q + ".self_attn.q_proj.bias", &bq, {nq}));
    RET(need(c, q + ".self_attn.k_proj.bias", &bk, {nkv}));
    RET(need(c, q + ".self_attn.v_proj.bias", &bv, {nkv}));
    RET(dmalloc(c, &y.wqkv, (size_t)c->Nqkv * H, false));
    RET(dmalloc(c, &y.bqkv, (size_t)c->Nqkv, false));
    CK(cudaMemcpy(y.wqkv, tq->p, (size_t)nq * H * 2, cudaMemcpyDeviceToDevice));
    CK(cudaMemcpy(y.wqkv + (size_t)nq * H, tk->p, (size_t)nkv * H * 2, cudaMemcpyDeviceToDevice));
    CK(cudaMemcpy(y.wqkv + (size_t)(nq + nkv) * H, tv->p, (size_t)nkv * H * 2, cudaMemcpyDeviceToDevice));
    CK(cudaMemcpy(y.bqkv, bq->p, (size_t)nq * 4, cudaMemcpyDeviceToDevice));
    CK(cudaMemcpy(y.bqkv + nq, bk->p, (size_t)nkv * 4, cudaMemcpyDeviceToDevice));
    CK(cudaMemcpy(y.bqkv + nq + nkv, bv->p, (size_t)nkv * 4, cudaMemcpyDeviceToDevice));
    *wb += (int64_t)c->Nqkv * H * 2 + (int64_t)c->Nqkv * 2;
    for (const char* s : {"q", "k", "v"}) { drop(c, q + ".self_attn." + s + "_proj.weight"); drop(c, q + ".self_attn." + s + "_proj.bias"); }
    RET(take_bf16(c, q + ".self_attn.o_proj.weight", {H, nq}, &y.wo, wb));
    RET(need(c, q + ".mlp.gate_proj.weight", &tg, {I, H}));
    RET(need(c, q + ".mlp.up_proj.weight", &tu, {I, H}));
    RET(dmalloc(c, &y.wgu, (size_t)2 * I * H, false));
    interleave_rows_kernel<<<4096, 256>>>((const bf16*)tg->p, (const bf16*)tu->p, y.wgu, (size_t)I, (size_t)H);
    CKL();
    CK(cudaDeviceSynchronize());
    *wb += (int64_t)2 * I * H * 2;
    drop(c, q + ".mlp.gate_proj.weight"); drop(c, q + ".mlp.up_proj.weight");
    RET(take_bf16(c, q + ".mlp.down_proj.weight", {H, I}, &y.wdown, wb));
    RET(take_f32(c, q + ".input_layernorm.weight", {H}, &y.ln1, wb));
    RET(take_f32(c, q + ".post_attention_layernorm.weight", {H}, &y.ln2, wb));
  }
  RET(take_f32(c, lm + ".norm.weight", {H}, &c->lm_norm, wb));
  RET(take_bf16(c, lm + ".embed_tokens.weight", {d.vocab_size, H}, &c->embed, nullptr));
  {
    RET(dmalloc(c, &c->valid_ids_dev, 8));
    CK(cudaMemcpy(c->valid_ids_dev, d.valid_ids, sizeof(int) * d.n_valid_ids, cudaMemcpyHostToDevice));
    RET(dmalloc(c, &c->head_valid, (size_t)8 * H));
    const bf16* table = c->embed;
    if (!d.tie_word_embeddings) {
      bf16* lmh;
      RET(take_bf16(c, "lm_head.weight", {d.vocab_size, H}, &lmh, nullptr));
      table = lmh;
    }
    c->lm_head_w = table;
    gather_rows_kernel<<<d.n_valid_ids, 256>>>(table, c->valid_ids_dev, c->head_valid, H);
    CKL();
    *wb += (int64_t)d.n_valid_ids * H * 2;
    const int hd = d.head_dim;
    RET(dmalloc(c, &c->inv_freq, hd / 2));
    std::vector<float> f(hd / 2);
    for (int i = 0; i < hd / 2; ++i) f[i] = 1.0f / powf(d.rope_theta, (float)(2 * i) / (float)hd);
    CK(cudaMemcpy(c->inv_freq, f.data(), sizeof(float) * (hd / 2), cudaMemcpyHostToDevice));
  }
  // ---------------- diffusion head ----------------
  {
    const std::string h = "model.prediction_head";
    const int F = d.head_ffn_dim, LH = d.head_layers;
    int64_t* hb = &c->wbytes[1];
    RET(take_bf16(c, h + ".noisy_images_proj.weight", {H, 64}, &c->h_noisy, hb));
    RET(take_bf16(c, h + ".cond_proj.weight", {H, H}, &c->h_cond, &c->wbytes[2]));
    RET(take_bf16(c, h + ".t_embedder.mlp.0.weight", {H, 256}, &c->h_t0, nullptr));
    RET(take_bf16(c, h + ".t_embedder.mlp.2.weight", {H, H}, &c->h_t2, nullptr));
    RET(take_bf16(c, h + ".final_layer.linear.weight", {64, H}, &c->h_final, hb));
    const size_t modrows = (size_t)(3 * LH + 2) * H;
    RET(dmalloc(c, &c->h_mod, modrows * H, false));
    c->head.resize(LH);
    // the per-step weights of all head layers live in ONE slab so a single L2 access-policy window can keep (part of) them
    // resident across the N diffusion steps (they are re-read N times per frame; everything else streams once)
    const size_t per_layer_el = (size_t)3 * F * H;
    c->head_slab_bytes = per_layer_el * LH * 2;
    RET(dmalloc(c, &c->head_slab, per_layer_el * LH, false));
    for (int l = 0; l < LH; ++l) {
      std::string q = S("%s.layers.%d", h.c_str(), l);
      RawTensor *tg, *tu, *tm, *td;
      RET(need(c, q + ".ffn.gate_proj.weight", &tg, {F, H}));
      RET(need(c, q + ".ffn.up_proj.weight", &tu, {F, H}));
      c->head[l].wgu = c->head_slab + per_layer_el * l;
      c->head[l].wdown = c->head[l].wgu + (size_t)2 * F * H;
      interleave_rows_kernel<<<4096, 256>>>((const bf16*)tg->p, (const bf16*)tu->p, c->head[l].wgu, (size_t)F, (size_t)H);
      CKL();
      CK(cudaDeviceSynchronize());
      *hb += (int64_t)2 * F * H * 2;
      drop(c, q + ".ffn.gate_proj.weight"); drop(c, q + ".ffn.up_proj.weight");
      RET(need(c, q + ".ffn.down_proj.weight", &td, {H, F}));
      CK(cudaMemcpy(c->head[l].wdown, td->p, (size_t)H * F * 2, cudaMemcpyDeviceToDevice));
      *hb += (int64_t)H * F * 2;
      drop(c, q + ".ffn.down_proj.weight");
      RET(take_f32(c, q + ".norm.weight", {H}, &c->head[l].norm, hb));
      RET(need(c, q + ".adaLN_modulation.1.weight", &tm, {3 * H, H}));
      CK(cudaMemcpy(c->h_mod + (size_t)l * 3 * H * H, tm->p, (size_t)3 * H * H * 2, cudaMemcpyDeviceToDevice));
      *hb += (int64_t)3 * H * H * 2;
      drop(c, q + ".adaLN_modulation.1.weight");
    }
    {
      int maxp = 0, maxw = 0;
      cudaDeviceGetAttribute(&maxp, cudaDevAttrMaxPersistingL2CacheSize, c->device);
      cudaDeviceGetAttribute(&maxw, cudaDevAttrMaxAccessPolicyWindowSize, c->device);
      size_t want = 0;     // measured neutral on B200 (79 MB max set-aside, sampler is latency- not bandwidth-bound): opt-in via VV_L2_PERSIST_MB
      if (getenv("VV_L2_PERSIST_MB")) want = (size_t)atoi(getenv("VV_L2_PERSIST_MB")) << 20;
      c->l2_persist_bytes = std::min<size_t>(want, (size_t)maxp);
      c->l2_window_max = (size_t)maxw;
      if (c->l2_persist_bytes) CK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, c->l2_persist_bytes));
      if (getenv("VV_VERBOSE")) fprintf(stderr, "[vv] L2 persisting max %d MB, window max %d MB, using %zu MB for a %zu MB head slab\n", maxp >> 20,
                                        maxw >> 20, c->l2_persist_bytes >> 20, c->head_slab_bytes >> 20);
    }
    RawTensor* tm;
    RET(need(c, h + ".final_layer.adaLN_modulation.1.weight", &tm, {2 * H, H}));
    CK(cudaMemcpy(c->h_mod + (size_t)LH * 3 * H * H, tm->p, (size_t)2 * H * H * 2, cudaMemcpyDeviceToDevice));
    *hb += (int64_t)2 * H * H * 2;
    drop(c, h + ".final_layer.adaLN_modulation.1.weight");
    const int NS = std::max(d.max_diffusion_steps, 1);
    RET(dmalloc(c, &c->temb, (size_t)NS * H));
    RET(dmalloc(c, &c->coef_dev, (size_t)NS));
    RET(dmalloc(c, &c->tfreqs, 128));
    std::vector<float> fr(128);
    for (int j = 0; j < 128; ++j) fr[j] = expf((-9.210340371976184f * (float)j) / 128.0f);
    CK(cudaMemcpy(c->tfreqs, fr.data(), 128 * 4, cudaMemcpyHostToDevice));
    RET(dmalloc(c, &c->s_tfeat, (size_t)NS * 256));
    RET(dmalloc(c, &c->s_t1, (size_t)NS * H));
    RET(dmalloc(c, &c->s_condp, (size_t)M2 * H));
    RET(dmalloc(c, &c->s_call, (size_t)NS * M2 * H));
    RET(dmalloc(c, &c->s_mod, (size_t)NS * M2 * modrows));
    RET(dmalloc(c, &c->cfg_dev, 4));
    RET(dmalloc(c, &c->gridbar, 2));
    RET(dmalloc(c, &c->st_bar, 32));
    if (getenv("VV_STREAM_TRACE")) {
      c->st_trace_cta = atoi(getenv("VV_STREAM_TRACE"));
      c->st_trace_ops = 4096;
      RET(dmalloc(c, &c->st_trace, (size_t)c->st_trace_ops * ST_TRACE));
      RET(dmalloc(c, &c->st_trace2, (size_t)c->st_trace_ops * c->sm_count * 2));
    }
    CK(cudaHostAlloc((void**)&c->st_diag_host, 64, cudaHostAllocMapped));
    memset(c->st_diag_host, 0, 64);
    CK(cudaHostGetDevicePointer((void**)&c->st_diag_dev, c->st_diag_host, 0));
    RET(dmalloc(c, &c->s_hx, (size_t)M2 * H));
    RET(dmalloc(c, &c->s_hg, (size_t)M2 * F));
    RET(dmalloc(c, &c->s_hgu, (size_t)2 * M2 * 2 * F));
    RET(dmalloc(c, &c->s_v, (size_t)M2 * 64));
    RET(dmalloc(c, &c->s_z, (size_t)2 * B * 64));
    RET(dmalloc(c, &c->s_x0, (size_t)2 * B * 64));
  }
  // ---------------- connectors ----------------
  {
    int64_t* cb = &c->wbytes[5];
    const std::string a = "model.acoustic_connector", s = "model.semantic_connector";
    RET(take_bf16(c, a + ".fc1.weight", {H, d.acoustic_vae_dim}, &c->ca_fc1, cb));
    RET(take_f32(c, a + ".fc1.bias", {H}, &c->ca_b1, cb));
    RET(take_f32(c, a + ".norm.weight", {H}, &c->ca_n, cb));
    RET(take_bf16(c, a + ".fc2.weight", {H, H}, &c->ca_fc2, cb));
    RET(take_f32(c, a + ".fc2.bias", {H}, &c->ca_b2, cb));
    RET(take_bf16(c, s + ".fc1.weight", {H, d.semantic_vae_dim}, &c->cs_fc1, cb));
    RET(take_f32(c, s + ".fc1.bias", {H}, &c->cs_b1, cb));
    RET(take_f32(c, s + ".norm.weight", {H}, &c->cs_n, cb));
    RET(take_bf16(c, s + ".fc2.weight", {H, H}, &c->cs_fc2, cb));
    RET(take_f32(c, s + ".fc2.bias", {H}, &c->cs_b2, cb));
    RET(dmalloc(c, &c->s_e, (size_t)B * H));
    RET(dmalloc(c, &c->s_c1, (size_t)B * H));
    RET(dmalloc(c, &c->s_feat, (size_t)B * d.semantic_vae_dim));
    RET(dmalloc(c, &c->s_audio, (size_t)B * 3200 * 4));
    RET(dmalloc(c, &c->s_latent, (size_t)B * 64));
  }
  // ---------------- codec decoder (tokenizer.py:823-912) ----------------
  int hop = 1;
  for (int i = 0; i < d.n_stages - 1; ++i) hop *= d.dec_ratios[i];
  size_t max_tc = 0, max_win = 0;
  {
    Codec& k = c->dec;
    int64_t* kb = &c->wbytes[3];
    std::vector<StateSeg> segs;
    const std::string p = "model.acoustic_tokenizer.decoder";
    const int ns = d.n_stages, nf = d.dec_n_filters;
    k.convs.resize(ns + 1); k.stages.resize(ns); k.T.resize(ns); k.C.resize(ns);
    int T = 1;
    for (int i = 0; i < ns; ++i) {
      const int C = nf << (ns - 1 - i);
      if (i == 0) RET(build_conv(c, p + ".upsample_layers.0.0.conv.conv", d.acoustic_vae_dim, C, 7, 1, &k.convs[0], kb, &segs));
      else { RET(build_convtr(c, S("%s.upsample_layers.%d.0.convtr.convtr", p.c_str(), i), C * 2, C, d.dec_ratios[i - 1], &k.convs[i], kb, &segs)); T *= d.dec_ratios[i - 1]; }
      k.T[i] = T; k.C[i] = C;
      max_tc = std::max(max_tc, (size_t)T * C);
      max_win = std::max(max_win, (size_t)(T + 8) * C * 2);
      k.stages[i].resize(d.dec_depths[i]);
      for (int j = 0; j < d.dec_depths[i]; ++j) RET(build_block(c, S("%s.stages.%d.%d", p.c_str(), i, j), C, &k.stages[i][j], kb, &segs));
    }
    RET(build_conv(c, p + ".head.conv.conv", nf, 1, 7, 1, &k.convs[ns], kb, &segs));
    if (T != hop) return fail(VV_ERR_INVALID, "decoder hop mismatch");
    RET(upload_segs(c, &k, segs));
    k.weight_bytes = *kb;
  }
  // ---------------- semantic encoder (tokenizer.py:694-774) ----------------
  {
    Codec& k = c->enc;
    int64_t* kb = &c->wbytes[4];
    std::vector<StateSeg> segs;
    const std::string p = "model.semantic_tokenizer.encoder";
    const int ns = d.n_stages, nf = d.enc_n_filters;
    k.convs.resize(ns + 1); k.stages.resize(ns); k.T.resize(ns); k.C.resize(ns);
    int T = hop;
    for (int i = 0; i < ns; ++i) {
      const int C = nf << i;
      if (i == 0) RET(build_conv(c, p + ".downsample_layers.0.0.conv.conv", 1, C, 7, 1, &k.convs[0], kb, &segs));
      else {
        const int r = d.enc_ratios[ns - 1 - i];     // TokenizerEncoder reverses the ratio list (tokenizer.py:701)
        RET(build_conv(c, S("%s.downsample_layers.%d.0.conv.conv", p.c_str(), i), C / 2, C, 2 * r, r, &k.convs[i], kb, &segs));
        if (T % r) return fail(VV_ERR_INVALID, "encoder ratio mismatch");
        T /= r;
      }
      k.T[i] = T; k.C[i] = C;
      max_tc = std::max(max_tc, (size_t)T * C);
      max_win = std::max(max_win, (size_t)(T + 16) * C * 2);
      k.stages[i].resize(d.enc_depths[i]);
      for (int j = 0; j < d.enc_depths[i]; ++j) RET(build_block(c, S("%s.stages.%d.%d", p.c_str(), i, j), C, &k.stages[i][j], kb, &segs));
    }
    if (T != 1) return fail(VV_ERR_INVALID, "encoder hop mismatch");
    RET(build_conv(c, p + ".head.conv.conv", nf << (ns - 1), d.semantic_vae_dim, 7, 1, &k.convs[ns], kb, &segs));
    RET(upload_segs(c, &k, segs));
    k.weight_bytes = *kb;
  }
  max_win = std::max(max_win, (size_t)(hop + 8) * 64);
  RET(dmalloc(c, &c->s_xa, (size_t)B * max_tc));
  RET(dmalloc(c, &c->s_xb, (size_t)B * max_tc));
  RET(dmalloc(c, &c->s_xn, (size_t)B * max_tc));
  RET(dmalloc(c, &c->s_u, (size_t)B * max_tc * 4));
  RET(dmalloc(c, &c->s_win, (size_t)B * max_win));
  c->planes_elems = std::max<size_t>((size_t)B * max_tc * 4, (size_t)std::max(d.max_diffusion_steps, 1) * M2 * H);
  RET(dmalloc(c, &c->s_planes, 2 * c->planes_elems));
  // ---------------- LM scratch ----------------
  RET(dmalloc(c, &c->s_h, (size_t)M2 * H));
  RET(dmalloc(c, &c->s_qkv, (size_t)M2 * c->Nqkv));
  RET(dmalloc(c, &c->s_qrot, (size_t)M2 * nq));
  RET(dmalloc(c, &c->s_attn, (size_t)M2 * nq));
  RET(dmalloc(c, &c->s_act, (size_t)M2 * I));
  RET(dmalloc(c, &c->s_lgu, (size_t)M2 * 2 * I));
  RET(dmalloc(c, &c->s_rope, (size_t)M2 * HD));
  RET(dmalloc(c, &c->s_cx, (size_t)2 * B * 8192));
  RET(dmalloc(c, &c->s_cu, (size_t)2 * B * 32768));
  RET(dmalloc(c, &c->s_pacc2, (size_t)M2 * d.num_kv_heads * c->sm_count * 8 * HD));
  RET(dmalloc(c, &c->s_pml2, (size_t)M2 * d.num_kv_heads * c->sm_count * 8 * 2));
  RET(dmalloc(c, &c->s_pacc, (size_t)M2 * d.num_q_heads * c->nsplit * HD));
  RET(dmalloc(c, &c->s_pml, (size_t)M2 * d.num_q_heads * c->nsplit * 2));
  RET(dmalloc(c, &c->s_tok, 64));
  RET(dmalloc(c, &c->kv_len_dev, 16));
  RET(dmalloc(c, &c->row_mode_dev, 16));
  {
    int ones[16];
    for (int i = 0; i < 16; ++i) ones[i] = 1;
    CK(cudaMemcpy(c->row_mode_dev, ones, sizeof ones, cudaMemcpyHostToDevice));
  }
  c->kv_len_host.assign(M2, 0);
  c->seq_pages.assign(M2, {});
  for (auto& r : c->raw) if (r.second.p) { cudaFree(r.second.p); r.second.p = nullptr; }
  CK(cudaDeviceSynchronize());
  c->finalized = true;
  return 0;
}

extern "C" int64_t vv_weight_bytes(vv_ctx* c, int which) { return (c && which >= 0 && which < 6) ? c->wbytes[which] : -1; }
extern "C" int64_t vv_launch_count(vv_ctx* c) { return c ? c->launches : -1; }

// ------------------------------------------------------------------------------------------------
// graph cache: every per-frame program is captured once per distinct argument tuple and replayed
// ------------------------------------------------------------------------------------------------
template <class F>
static int run_cached(vv_ctx* c, const std::string& key, cudaStream_t s, F&& enqueue) {
  if (!c->use_graphs || s == nullptr) { L l{c, s}; return enqueue(l); }
  auto it = c->graphs.find(key);
  if (it == c->graphs.end()) {
    const int64_t before = c->launches;
    CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    L l{c, s};
    int r = enqueue(l);
    cudaGraph_t g = nullptr;
    cudaError_t e = cudaStreamEndCapture(s, &g);
    if (r < 0) { if (g) cudaGraphDestroy(g); return r; }
    if (e != cudaSuccess) return fail(VV_ERR_CUDA, "graph capture failed (%s): %s", key.c_str(), cudaGetErrorString(e));
    GraphEntry ge;
    ge.launches = c->launches - before;
    c->launches = before;
    e = cudaGraphInstantiate(&ge.exec, g, 0);
    cudaGraphDestroy(g);
    if (e != cudaSuccess) return fail(VV_ERR_CUDA, "graph instantiate failed: %s", cudaGetErrorString(e));
    it = c->graphs.emplace(key, ge).first;
  }
  CK(cudaGraphLaunch(it->second.exec, s));
  c->launches += it->second.launches;
  return 0;
}

// ------------------------------------------------------------------------------------------------
// KV pages
// ------------------------------------------------------------------------------------------------
template <class T>
static void dfree(vv_ctx* c, T** p) {
  if (!*p) return;
  auto it = std::find(c->allocs.begin(), c->allocs.end(), (void*)*p);
  if (it != c->allocs.end()) c->allocs.erase(it);
  cudaFree(*p);
  *p = nullptr;
}
static void drop_graphs(vv_ctx* c, std::initializer_list<const char*> prefixes) {
  for (auto it = c->graphs.begin(); it != c->graphs.end();) {
    bool hit = false;
    for (const char* pre : prefixes) hit = hit || it->first.rfind(pre, 0) == 0;
    if (hit) { cudaGraphExecDestroy(it->second.exec); it = c->graphs.erase(it); }
    else ++it;
  }
}

// (Re-)size the page pool.  Calling it again drops every sequence (lengths 0, all pages free) and re-allocates the pool, so a
// long-lived service can grow the cache between generate() calls; captured LM graphs bake the pool pointers and are re-captured.
extern "C" int vv_kv_init(vv_ctx* c, int64_t n_pages) {
  if (!c || !c->finalized) return fail(VV_ERR_STATE, "vv_kv_init before vv_finalize_weights");
  if (n_pages < 2 * c->d.max_batch) return fail(VV_ERR_INVALID, "vv_kv_init: need at least one page per sequence (%d)", 2 * c->d.max_batch);
  CK(cudaSetDevice(c->device));
  if (c->kpool) {
    CK(cudaDeviceSynchronize());
    dfree(c, &c->kpool); dfree(c, &c->vpool); dfree(c, &c->page_table_dev);
    drop_graphs(c, {"lm:", "lmr:", "frame:"});
    for (auto it = c->sprogs.begin(); it != c->sprogs.end();) {
      if (it->first.rfind("lmf:", 0) == 0) { dfree(c, &it->second.ops); dfree(c, &it->second.tmaps); it = c->sprogs.erase(it); }
      else ++it;
    }
    std::fill(c->kv_len_host.begin(), c->kv_len_host.end(), 0);
    for (auto& pg : c->seq_pages) pg.clear();
  }
  const auto& d = c->d;
  const size_t per_layer = (size_t)n_pages * d.num_kv_heads * KV_PAGE * d.head_dim;
  RET(dmalloc(c, &c->kpool, per_layer * d.num_layers));
  RET(dmalloc(c, &c->vpool, per_layer * d.num_layers));
  c->n_pages = n_pages;
  c->max_pages = (int)n_pages;
  const size_t nt = (size_t)2 * d.max_batch * c->max_pages;
  RET(dmalloc(c, &c->page_table_dev, nt));
  c->free_pages.clear();
  for (int i = (int)n_pages - 1; i >= 0; --i) c->free_pages.push_back(i);
  return 0;
}
extern "C" int64_t vv_kv_pages_free(vv_ctx* c) { return c ? (int64_t)c->free_pages.size() : -1; }
extern "C" int64_t vv_kv_pages_total(vv_ctx* c) { return c ? c->n_pages : -1; }
extern "C" int64_t vv_kv_len(vv_ctx* c, int seq) { return (c && seq >= 0 && seq < (int)c->kv_len_host.size()) ? c->kv_len_host[seq] : -1; }

// page-table entries travel BY VALUE in the launch arguments: pages return to the free list when a sequence shrinks, so an entry can be
// rewritten while copies of its previous value are still queued on the stream -- an asynchronous copy out of a host mirror would race
struct PageVals { int v[32]; };
__global__ void page_table_set_kernel(int* dst, PageVals pv, int n) { if ((int)threadIdx.x < n) dst[threadIdx.x] = pv.v[threadIdx.x]; }

extern "C" int vv_kv_reserve(vv_ctx* c, int seq, int64_t n_tokens, void* stream) {
  if (!c || !c->kpool) return fail(VV_ERR_STATE, "KV pool not initialised");
  if (seq < 0 || seq >= 2 * c->d.max_batch) return fail(VV_ERR_INVALID, "bad seq %d", seq);
  auto& pg = c->seq_pages[seq];
  const int64_t needp = (n_tokens + KV_PAGE - 1) / KV_PAGE;
  if (needp > c->max_pages) return fail(VV_ERR_NOMEM, "KV page pool too small (seq %d needs %lld pages of %d)", seq, (long long)needp, c->max_pages);
  size_t first = pg.size();
  while ((int64_t)pg.size() < needp) {
    if (c->free_pages.empty()) return fail(VV_ERR_NOMEM, "KV page pool exhausted (seq %d needs %lld pages)", seq, (long long)needp);
    pg.push_back(c->free_pages.back());
    c->free_pages.pop_back();
  }
  while (first < pg.size()) {
    PageVals pv;
    const int n = (int)std::min<size_t>(32, pg.size() - first);
    for (int i = 0; i < 32; ++i) pv.v[i] = i < n ? pg[first + i] : 0;
    page_table_set_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(c->page_table_dev + (size_t)seq * c->max_pages + first, pv, n);
    CKL();
    c->launches++;
    first += n;
  }
  return 0;
}
// pages beyond the new length go back to the free list (a negative stream restarts at every <speech_start>; a server reuses rows)
static void release_pages(vv_ctx* c, int seq, int64_t len) {
  auto& pg = c->seq_pages[seq];
  const size_t keep = (size_t)((len + KV_PAGE - 1) / KV_PAGE);
  while (pg.size() > keep) { c->free_pages.push_back(pg.back()); pg.pop_back(); }
}

struct Lens { int v[16]; };
__global__ void kv_set_all_kernel(int* kv_len, Lens l, int n) { if (threadIdx.x < n) kv_len[threadIdx.x] = l.v[threadIdx.x]; }

static int push_lens(vv_ctx* c, cudaStream_t s) {
  Lens l;
  const int n = 2 * c->d.max_batch;
  for (int i = 0; i < 16; ++i) l.v[i] = i < n ? (int)c->kv_len_host[i] : 0;
  kv_set_all_kernel<<<1, 32, 0, s>>>(c->kv_len_dev, l, n);
  CKL();
  c->launches++;
  return 0;
}
extern "C" int vv_kv_set_len(vv_ctx* c, int seq, int64_t len, void* stream) {
  if (!c || !c->kpool) return fail(VV_ERR_STATE, "KV pool not initialised");
  if (seq < 0 || seq >= 2 * c->d.max_batch) return fail(VV_ERR_INVALID, "bad seq %d", seq);
  if (len < 0 || len > (int64_t)c->seq_pages[seq].size() * KV_PAGE) return fail(VV_ERR_INVALID, "vv_kv_set_len: %lld outside the reserved range of seq %d", (long long)len, seq);
  c->kv_len_host[seq] = len;
  release_pages(c, seq, len);
  return push_lens(c, (cudaStream_t)stream);
}
extern "C" int vv_kv_commit(vv_ctx* c, const int32_t* adv, void* stream) {
  if (!c || !c->kpool) return fail(VV_ERR_STATE, "KV pool not initialised");
  for (int i = 0; i < 2 * c->d.max_batch; ++i) c->kv_len_host[i] += adv[i] ? 1 : 0;
  return push_lens(c, (cudaStream_t)stream);
}
// drop the entry at position `pos` of a sequence: the last committed entry moves into its place (all layers), the length shrinks by one.
// Attention does not depend on the order of the cached entries (keys are stored rotated), so this is how a sequence forgets an OLDER entry --
// the reference's cache shifting with refresh_negative=False hides one (modeling_vibevoice_inference.py:599-624).
__global__ void kv_move_kernel(bf16* kpool, bf16* vpool, const int* page_row, int kv_heads, int hd, size_t per_layer, int src, int dst) {
  const int layer = blockIdx.x;
  const int sp = page_row[src / KV_PAGE], dp = page_row[dst / KV_PAGE];
  for (int i = threadIdx.x; i < kv_heads * hd; i += blockDim.x) {
    const int h = i / hd, d = i % hd;
    const size_t so = per_layer * layer + (((size_t)sp * kv_heads + h) * KV_PAGE + (src % KV_PAGE)) * hd + d;
    const size_t dofs = per_layer * layer + (((size_t)dp * kv_heads + h) * KV_PAGE + (dst % KV_PAGE)) * hd + d;
    kpool[dofs] = kpool[so];
    vpool[dofs] = vpool[so];
  }
}
extern "C" int vv_kv_delete_slot(vv_ctx* c, int seq, int64_t pos, void* stream) {
  if (!c || !c->kpool) return fail(VV_ERR_STATE, "KV pool not initialised");
  if (seq < 0 || seq >= 2 * c->d.max_batch) return fail(VV_ERR_INVALID, "bad seq %d", seq);
  const int64_t len = c->kv_len_host[seq];
  if (pos < 0 || pos >= len) return fail(VV_ERR_INVALID, "vv_kv_delete_slot: position %lld outside [0,%lld)", (long long)pos, (long long)len);
  const auto& d = c->d;
  if (pos != len - 1) {
    const size_t per_layer = (size_t)c->n_pages * d.num_kv_heads * KV_PAGE * d.head_dim;
    kv_move_kernel<<<d.num_layers, 256, 0, (cudaStream_t)stream>>>(c->kpool, c->vpool, c->page_table_dev + (size_t)seq * c->max_pages, d.num_kv_heads, d.head_dim, per_layer,
                                                                  (int)(len - 1), (int)pos);
    CKL();
    c->launches++;
  }
  c->kv_len_host[seq] = len - 1;
  release_pages(c, seq, len - 1 + 1);     // keep the page of the next speculative entry
  return push_lens(c, (cudaStream_t)stream);
}
extern "C" int vv_set_row_mode(vv_ctx* c, const int32_t* rm, void* stream) {
  if (!c || !c->finalized) return fail(VV_ERR_STATE, "not finalized");
  Lens l;
  for (int i = 0; i < 16; ++i) l.v[i] = i < 2 * c->d.max_batch ? rm[i] : 0;
  kv_set_all_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(c->row_mode_dev, l, 2 * c->d.max_batch);
  CKL();
  c->launches++;
  return 0;
}
extern "C" int vv_set_rope_inv_freq(vv_ctx* c, const float* f, int n) {
  if (!c || !c->finalized) return fail(VV_ERR_STATE, "not finalized");
  if (n != c->d.head_dim / 2) return fail(VV_ERR_INVALID, "inv_freq must have %d entries", c->d.head_dim / 2);
  CK(cudaMemcpy(c->inv_freq, f, sizeof(float) * n, cudaMemcpyHostToDevice));
  return 0;
}
extern "C" int vv_kv_write(vv_ctx* c, int seq, int layer, int64_t pos0, int64_t n_tokens, const void* k, const void* v, void* stream) {
  if (!c || !c->kpool) return fail(VV_ERR_STATE, "KV pool not initialised");
  RET(vv_kv_reserve(c, seq, pos0 + n_tokens, stream));
  const auto& d = c->d;
  const size_t per_layer = (size_t)c->n_pages * d.num_kv_heads * KV_PAGE * d.head_dim;
  kv_write_kernel<<<(unsigned)n_tokens, 128, 0, (cudaStream_t)stream>>>((const bf16*)k, (const bf16*)v, c->kpool + per_layer * layer,
                                                                      c->vpool + per_layer * layer,
                                                                      c->page_table_dev + (size_t)seq * c->max_pages, d.num_kv_heads, d.head_dim, pos0, n_tokens);
  CKL();
  c->launches++;
  return 0;
}

// ------------------------------------------------------------------------------------------------
// a-3: LM decode
// ------------------------------------------------------------------------------------------------
static int enqueue_lm_head(const L& l, const float* hidden, float* logits, int32_t* tokens) {
  vv_ctx* c = l.c;
  CK(launch_k(l, lm_head_argmax_kernel, dim3(c->d.max_batch), dim3(256), 0, hidden, c->head_valid, c->valid_ids_dev, c->d.n_valid_ids, c->d.hidden_size, logits, tokens));
  return 0;
}

// The linears of decoder layers [li0, li1) as one weight-stream program (vv_stream.cuh), launched in pieces around the attention kernels:
//   [zero s_qkv | QKV(li0)]  attention(li0)  [O(li0) GU(li0) DN(li0) QKV(li0+1)]  attention(li0+1)  ...  [O GU DN](li1-1)
// QKV: RMSNorm prologue, bias; O: accumulates into the residual stream; GU: RMSNorm prologue -> raw gate/up sums; DN: SwiGLU prologue,
// accumulates into the residual stream.  Zero-fill jobs ride on the O stage (s_qkv and the gate/up buffer are dead at that point).
static int lm_stream_prog(vv_ctx* c, int li0, int li1, const vv_ctx::StreamProg** out) {
  char key[64];
  snprintf(key, sizeof key, "lm:%d:%d", li0, li1);
  auto it = c->sprogs.find(key);
  if (it != c->sprogs.end()) { *out = &it->second; return 0; }
  const auto& d = c->d;
  const int H = d.hidden_size, I = d.intermediate_size, M = 2 * d.max_batch, nq = d.num_q_heads * d.head_dim;
  StreamBuilder b(c);
  b.nop(false, c->s_qkv, (long long)M * c->Nqkv);
  auto qkv = [&](int li) -> int {
    const LmLayer& y = c->lm[li];
    SOp* o;
    RET(b.gemv(y.wqkv, y.bqkv, c->s_h, H, c->s_qkv, c->Nqkv, M, c->Nqkv, H, true, &o));
    o->pro = SP_RMSNORM; o->pro_w = y.ln1; o->pro_eps = d.rms_norm_eps;
    return 0;
  };
  RET(qkv(li0));
  for (int li = li0; li < li1; ++li) {
    const LmLayer& y = c->lm[li];
    SOp* o;
    RET(b.gemv(y.wo, nullptr, c->s_attn, nq, c->s_h, H, M, H, nq, false, &o));
    o->init_dst = c->s_qkv; o->init_n = (long long)M * c->Nqkv;
    o->init2_dst = c->s_lgu; o->init2_n = (long long)M * 2 * I;
    RET(b.gemv(y.wgu, nullptr, c->s_h, H, c->s_lgu, 2 * I, M, 2 * I, H, true, &o));
    o->pro = SP_RMSNORM; o->pro_w = y.ln2; o->pro_eps = d.rms_norm_eps;
    RET(b.gemv(y.wdown, nullptr, c->s_lgu, 2 * I, c->s_h, H, M, H, I, true, &o));
    o->pro = SP_SWIGLU;
    if (li + 1 < li1) RET(qkv(li + 1));
  }
  vv_ctx::StreamProg pr;
  RET(finish_stream(b, &pr));
  it = c->sprogs.emplace(key, pr).first;
  *out = &it->second;
  return 0;
}

// All of decoder layers [li0, li1) as ONE weight-stream program: per layer QKV -> attention (K/V pages through the ring) -> O (prologue merges
// the attention partials) -> gate/up -> down.  5 grid barriers per layer, no kernel boundary inside the stack.
static int lm_stream_prog_full(vv_ctx* c, int li0, int li1, const vv_ctx::StreamProg** out) {
  char key[64];
  snprintf(key, sizeof key, "lmf:%d:%d", li0, li1);
  auto it = c->sprogs.find(key);
  if (it != c->sprogs.end()) { *out = &it->second; return 0; }
  const auto& d = c->d;
  const int H = d.hidden_size, I = d.intermediate_size, M = 2 * d.max_batch, nq = d.num_q_heads * d.head_dim;
  StreamBuilder b(c);
  b.nop(false, c->s_qkv, (long long)M * c->Nqkv);
  for (int li = li0; li < li1; ++li) {
    const LmLayer& y = c->lm[li];
    SOp* o;
    RET(b.gemv(y.wqkv, y.bqkv, c->s_h, H, c->s_qkv, c->Nqkv, M, c->Nqkv, H, true, &o));
    o->pro = SP_RMSNORM; o->pro_w = y.ln1; o->pro_eps = d.rms_norm_eps;
    if (li == li0) { b.fill_att(&o->att, li); o->rope_rows = M; }        // positions are the same for every layer of this call
    RET(b.attn(li, M));
    RET(b.gemv(y.wo, nullptr, nullptr, 0, c->s_h, H, M, H, nq, true, &o));
    o->pro = SP_COMBINE;
    b.fill_att(&o->att, li);
    b.needs_kv.push_back((int)b.ops.size() - 1);
    o->init_dst = c->s_qkv; o->init_n = (long long)M * c->Nqkv;
    o->init2_dst = c->s_lgu; o->init2_n = (long long)M * 2 * I;
    RET(b.gemv(y.wgu, nullptr, c->s_h, H, c->s_lgu, 2 * I, M, 2 * I, H, true, &o));
    o->pro = SP_RMSNORM; o->pro_w = y.ln2; o->pro_eps = d.rms_norm_eps;
    RET(b.gemv(y.wdown, nullptr, c->s_lgu, 2 * I, c->s_h, H, M, H, I, true, &o));
    o->pro = SP_SWIGLU;
  }
  vv_ctx::StreamProg pr;
  RET(finish_stream(b, &pr));
  it = c->sprogs.emplace(key, pr).first;
  *out = &it->second;
  return 0;
}

// decoder layers [li0, li1) over the residual stream s_h (rows = 2B sequences; rows with row_mode 0 neither read nor append KV)
static int enqueue_lm_layers(const L& l, int li0, int li1, const vv_ctx::StreamProg* sprog = nullptr) {
  vv_ctx* c = l.c;
  const auto& d = c->d;
  const int H = d.hidden_size, I = d.intermediate_size, M = 2 * d.max_batch, nq = d.num_q_heads * HD;
  const size_t per_layer = (size_t)c->n_pages * d.num_kv_heads * KV_PAGE * HD;
  const float scale = 1.0f / sqrtf((float)HD);
  if (sprog && (c->use_stream & 4)) return launch_stream(l, *sprog);      // whole stack, attention included
  if (d.head_dim != HD) return fail(VV_ERR_INVALID, "head_dim %d runs through the weight-stream path only (VV_STREAM bit 2)", d.head_dim);
  if (sprog) RET(launch_stream(l, *sprog, 0, 2));
  for (int li = li0; li < li1; ++li) {
    const LmLayer& y = c->lm[li];
    GemvP p = mk(y.wqkv, y.bqkv, c->s_h, H, c->s_qkv, c->Nqkv, M, c->Nqkv, H);
    p.pro = PRO_RMSNORM; p.pro_w = y.ln1; p.pro_eps = d.rms_norm_eps;
    if (!sprog) RET(linear(l, p));
    KvView kv;
    kv.kpool = c->kpool + per_layer * li; kv.vpool = c->vpool + per_layer * li;
    kv.page_table = c->page_table_dev; kv.max_pages = c->max_pages; kv.kv_len = c->kv_len_dev; kv.row_mode = c->row_mode_dev;
    kv.kv_heads = d.num_kv_heads; kv.q_heads = d.num_q_heads;
    if (c->use_mma_attn && c->fuse_rope) {
      CK(cudaFuncSetAttribute(attn_partial_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AT2_SMEM));
      CK(launch_k(l, attn_partial_mma_kernel, dim3(c->nsplit, d.num_kv_heads, M), dim3(128), (size_t)AT2_SMEM, c->s_qkv, 1, c->inv_freq, kv, c->s_pacc,
                  c->s_pml, c->nsplit, scale));
    } else {
      CK(launch_k(l, rope_append_kernel, dim3(M), dim3(256), 0, c->s_qkv, c->s_qrot, kv, c->inv_freq));
      if (c->use_mma_attn) {
        CK(cudaFuncSetAttribute(attn_partial_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AT2_SMEM));
        CK(launch_k(l, attn_partial_mma_kernel, dim3(c->nsplit, d.num_kv_heads, M), dim3(128), (size_t)AT2_SMEM, c->s_qrot, 0, c->inv_freq, kv, c->s_pacc,
                    c->s_pml, c->nsplit, scale));
      } else {
        CK(launch_k(l, attn_partial_kernel, dim3(c->nsplit, d.num_kv_heads, M), dim3(128), 0, c->s_qrot, kv, c->s_pacc, c->s_pml, c->nsplit, scale));
      }
    }
    CK(launch_k(l, attn_combine_kernel, dim3(d.num_q_heads, M), dim3(128), 0, c->s_pacc, c->s_pml, c->row_mode_dev, c->s_attn, d.num_q_heads, c->nsplit));
    if (sprog) { RET(launch_stream(l, *sprog, 2 + 4 * (li - li0), li + 1 < li1 ? 4 : 3)); continue; }
    p = mk(y.wo, nullptr, c->s_attn, nq, c->s_h, H, M, H, nq);
    p.epi = EPI_RESID; p.res = c->s_h; p.ldres = H;
    RET(linear(l, p));
    p = mk(y.wgu, nullptr, c->s_h, H, c->s_act, I, M, 2 * I, H);
    p.pro = PRO_RMSNORM; p.pro_w = y.ln2; p.pro_eps = d.rms_norm_eps; p.epi = EPI_SWIGLU;
    RET(linear(l, p));
    p = mk(y.wdown, nullptr, c->s_act, I, c->s_h, H, M, H, I);
    p.epi = EPI_RESID; p.res = c->s_h; p.ldres = H;
    RET(linear(l, p));
  }
  return 0;
}

static int enqueue_final_norm(const L& l, float* hidden) {
  vv_ctx* c = l.c;
  const auto& d = c->d;
  const int H = d.hidden_size, M = 2 * d.max_batch;
  if (H >= 512) CK(launch_k(l, rows_norm_block_kernel, dim3(M), dim3(256), 0, c->s_h, c->lm_norm, hidden, H, d.rms_norm_eps));
  else CK(launch_k(l, rows_norm_kernel, dim3((M + 7) / 8), dim3(256), 0, c->s_h, c->lm_norm, hidden, M, H, d.rms_norm_eps));
  return 0;
}

static int enqueue_lm_decode(const L& l, const float* embeds, float* hidden, float* logits, int32_t* tokens, const vv_ctx::StreamProg* sprog) {
  vv_ctx* c = l.c;
  const int H = c->d.hidden_size, M = 2 * c->d.max_batch;
  CK(cudaMemcpyAsync(c->s_h, embeds, (size_t)M * H * 4, cudaMemcpyDeviceToDevice, l.s));
  RET(enqueue_lm_layers(l, 0, c->d.num_layers, sprog));
  RET(enqueue_final_norm(l, hidden));
  return enqueue_lm_head(l, hidden, logits, tokens);
}

// streaming-0.5B: the Qwen2 stack is split into a lower text-only stack without final norm and an upper "TTS LM" stack
// (modeling_vibevoice_streaming.py:134-146); each stack keeps its own KV sequences (lengths differ: the upper stack also sees the speech
// positions).  One call runs layers [begin, end) for the rows enabled by vv_set_row_mode, appends their K/V speculatively at kv_len (commit
// with vv_kv_commit as for vv_lm_decode) and returns the residual stream -- normalised with the model's final norm iff final_norm != 0.
static int enqueue_lm_range(const L& l, const float* embeds, int li0, int li1, int final_norm, float* hidden, const vv_ctx::StreamProg* sprog) {
  vv_ctx* c = l.c;
  const int H = c->d.hidden_size, M = 2 * c->d.max_batch;
  CK(cudaMemcpyAsync(c->s_h, embeds, (size_t)M * H * 4, cudaMemcpyDeviceToDevice, l.s));
  RET(enqueue_lm_layers(l, li0, li1, sprog));
  if (final_norm) return enqueue_final_norm(l, hidden);
  CK(cudaMemcpyAsync(hidden, c->s_h, (size_t)M * H * 4, cudaMemcpyDeviceToDevice, l.s));
  return 0;
}

extern "C" int vv_lm_decode(vv_ctx* c, const float* embeds, float* hidden, float* logits, int32_t* tokens, void* stream) {
  if (!c || !c->kpool) return fail(VV_ERR_STATE, "vv_lm_decode: KV pool not initialised");
  CK(cudaSetDevice(c->device));
  for (int s = 0; s < 2 * c->d.max_batch; ++s) RET(vv_kv_reserve(c, s, c->kv_len_host[s] + 1, stream));
  char key[256];
  snprintf(key, sizeof key, "lm:%p:%p:%p:%p", (const void*)embeds, (void*)hidden, (void*)logits, (void*)tokens);
  const vv_ctx::StreamProg* sprog = nullptr;      // built outside stream capture
  if (c->use_stream & 4) RET(lm_stream_prog_full(c, 0, c->d.num_layers, &sprog));
  else if (c->use_stream & 2) RET(lm_stream_prog(c, 0, c->d.num_layers, &sprog));
  return run_cached(c, key, (cudaStream_t)stream, [&](const L& l) { return enqueue_lm_decode(l, embeds, hidden, logits, tokens, sprog); });
}
extern "C" int vv_lm_head(vv_ctx* c, const float* hidden, float* logits, int32_t* tokens, void* stream) {
  if (!c || !c->finalized) return fail(VV_ERR_STATE, "not finalized");
  L l{c, (cudaStream_t)stream};
  return enqueue_lm_head(l, hidden, logits, tokens);
}
// Full-vocabulary logits for the positive rows: only needed when the caller installs its own LogitsProcessor objects or samples with
// top-k / top-p warpers, which rank the whole vocabulary BEFORE the token constraint (modeling_vibevoice_inference.py:310-319, 488-490).
// One GEMV over the (tied) embedding / lm_head matrix; the default path never calls this (it computes the <= 5 surviving logits only).
extern "C" int vv_lm_logits_full(vv_ctx* c, const float* hidden, float* logits_out, void* stream) {
  if (!c || !c->finalized) return fail(VV_ERR_STATE, "not finalized");
  CK(cudaSetDevice(c->device));
  L l{c, (cudaStream_t)stream};
  GemvP p = mk(c->lm_head_w, nullptr, hidden, c->d.hidden_size, logits_out, c->d.vocab_size, c->d.max_batch, c->d.vocab_size, c->d.hidden_size);
  return linear(l, p);
}
struct Toks { int v[16]; };
__global__ void embed_gather_val_kernel(const bf16* __restrict__ table, Toks t, float* __restrict__ out, int H) {
  const bf16* row = table + (size_t)t.v[blockIdx.x] * H;
  for (int k = threadIdx.x; k < H; k += blockDim.x) out[(size_t)blockIdx.x * H + k] = __bfloat162float(row[k]);
}
extern "C" int vv_lm_decode_range(vv_ctx* c, const float* embeds, int layer_begin, int layer_end, int final_norm, float* hidden, void* stream) {
  if (!c || !c->kpool) return fail(VV_ERR_STATE, "vv_lm_decode_range: KV pool not initialised");
  if (layer_begin < 0 || layer_end > c->d.num_layers || layer_begin >= layer_end) return fail(VV_ERR_INVALID, "bad layer range [%d,%d)", layer_begin, layer_end);
  CK(cudaSetDevice(c->device));
  for (int s = 0; s < 2 * c->d.max_batch; ++s) RET(vv_kv_reserve(c, s, c->kv_len_host[s] + 1, stream));
  char key[256];
  snprintf(key, sizeof key, "lmr:%p:%p:%d:%d:%d", (const void*)embeds, (void*)hidden, layer_begin, layer_end, final_norm);
  const vv_ctx::StreamProg* sprog = nullptr;
  if (c->use_stream & 4) RET(lm_stream_prog_full(c, layer_begin, layer_end, &sprog));
  else if (c->use_stream & 2) RET(lm_stream_prog(c, layer_begin, layer_end, &sprog));
  return run_cached(c, key, (cudaStream_t)stream, [&](const L& l) { return enqueue_lm_range(l, embeds, layer_begin, layer_end, final_norm, hidden, sprog); });
}

extern "C" int vv_embed_tokens(vv_ctx* c, const int32_t* tokens_host, int n, float* out, void* stream) {
  if (!c || !c->finalized) return fail(VV_ERR_STATE, "not finalized");
  if (n < 1 || n > 16) return fail(VV_ERR_INVALID, "vv_embed_tokens: n must be in [1,16]");
  Toks t;
  for (int i = 0; i < 16; ++i) {
    t.v[i] = i < n ? tokens_host[i] : 0;
    if (t.v[i] < 0 || t.v[i] >= c->d.vocab_size) return fail(VV_ERR_INVALID, "token id %d out of range", t.v[i]);
  }
  embed_gather_val_kernel<<<n, 256, 0, (cudaStream_t)stream>>>(c->embed, t, out, c->d.hidden_size);
  CKL();
  c->launches++;
  return 0;
}

// ------------------------------------------------------------------------------------------------
// a-4: diffusion sampler
// ------------------------------------------------------------------------------------------------
static int set_diffusion_steps(vv_ctx* c, int n_steps, const float* timesteps, const float* coef, int ncol, void* stream);
extern "C" int vv_set_diffusion_steps(vv_ctx* c, int n_steps, const float* timesteps, const float* coef, void* stream) {
  return set_diffusion_steps(c, n_steps, timesteps, coef, 6, stream);
}
extern "C" int vv_set_diffusion_steps_sde(vv_ctx* c, int n_steps, const float* timesteps, const float* coef7, void* stream) {
  return set_diffusion_steps(c, n_steps, timesteps, coef7, 7, stream);
}
extern "C" int vv_set_step_noise(vv_ctx* c, const float* step_noise) {
  if (!c) return fail(VV_ERR_INVALID, "null ctx");
  if (c->step_noise != step_noise) {            // captured graphs hold the old pointer
    drop_graphs(c, {"tail:", "diff:", "frame:"});
  }
  c->step_noise = step_noise;
  return 0;
}
static int set_diffusion_steps(vv_ctx* c, int n_steps, const float* timesteps, const float* coef, int ncol, void* stream) {
  if (!c || !c->finalized) return fail(VV_ERR_STATE, "not finalized");
  if (n_steps < 1 || n_steps > c->d.max_diffusion_steps) return fail(VV_ERR_INVALID, "n_steps %d outside [1,%d]", n_steps, c->d.max_diffusion_steps);
  CK(cudaSetDevice(c->device));
  cudaStream_t s = (cudaStream_t)stream;
  const int H = c->d.hidden_size;
  std::vector<DpmCoef> cf(n_steps);
  for (int i = 0; i < n_steps; ++i) {
    cf[i].a0 = coef[i * ncol + 0]; cf[i].s0 = coef[i * ncol + 1]; cf[i].ks = coef[i * ncol + 2]; cf[i].kx = coef[i * ncol + 3];
    cf[i].rinv = coef[i * ncol + 4]; cf[i].order = (int)coef[i * ncol + 5]; cf[i].kn = ncol == 7 ? coef[i * ncol + 6] : 0.f;
  }
  c->sde = (ncol == 7);
  c->coef_host = cf;
  c->coef_version++;
  CK(cudaStreamSynchronize(s));
  CK(cudaMemcpy(c->coef_dev, cf.data(), sizeof(DpmCoef) * n_steps, cudaMemcpyHostToDevice));
  float* tdev = c->s_t1;   // reuse as staging for the timesteps (n floats) before it is overwritten below
  CK(cudaMemcpy(tdev, timesteps, sizeof(float) * n_steps, cudaMemcpyHostToDevice));
  timestep_feat_kernel<<<n_steps, 256, 0, s>>>(tdev, c->tfreqs, c->s_tfeat, n_steps);
  CKL();
  CK(cudaStreamSynchronize(s));
  L l{c, s};
  GemvP p = mk(c->h_t0, nullptr, c->s_tfeat, 256, c->s_t1, H, n_steps, H, 256);
  p.epi = EPI_SILU;
  RET(linear(l, p));
  p = mk(c->h_t2, nullptr, c->s_t1, H, c->temb, H, n_steps, H, H);
  RET(linear(l, p));
  CK(cudaStreamSynchronize(s));
  c->n_steps = n_steps;
  // programs captured with another step count are stale
  drop_graphs(c, {"tail:", "diff:", "frame:"});
  return 0;
}

// head ops of one diffusion step (shared by the kernel-per-stage path and the program builder)
static void head_step_gemvs(vv_ctx* c, int i, std::vector<GemvP>* out) {
  const auto& d = c->d;
  const int H = d.hidden_size, F = d.head_ffn_dim, M = 2 * d.max_batch, LH = d.head_layers;
  const int modld = (3 * LH + 2) * H;
  const float* mod = c->s_mod + (size_t)i * M * modld;
  for (int li = 0; li < LH; ++li) {
    const HeadLayer& hl = c->head[li];
    GemvP p = mk(hl.wgu, nullptr, c->s_hx, H, c->s_hg, F, M, 2 * F, H);
    p.pro = PRO_ADALN; p.pro_w = hl.norm; p.pro_eps = d.head_rms_eps;
    p.pro_shift = mod + (size_t)li * 3 * H; p.pro_scale = mod + (size_t)li * 3 * H + H; p.pro_ld = modld;
    p.epi = EPI_SWIGLU;
    out->push_back(p);
    p = mk(hl.wdown, nullptr, c->s_hg, F, c->s_hx, H, M, H, F);
    p.epi = EPI_GATED_RESID; p.epi_a = mod + (size_t)li * 3 * H + 2 * H; p.epi_lda = modld; p.res = c->s_hx; p.ldres = H;
    out->push_back(p);
  }
  GemvP p = mk(c->h_final, nullptr, c->s_hx, H, c->s_v, 64, M, 64, H);
  p.pro = PRO_ADALN; p.pro_w = nullptr; p.pro_eps = d.head_rms_eps;
  p.pro_shift = mod + (size_t)LH * 3 * H; p.pro_scale = mod + (size_t)LH * 3 * H + H; p.pro_ld = modld;
  out->push_back(p);
}
__global__ void set_float_kernel(float* p, float v) { *p = v; }
// the CFG scale is read from device memory by the solver kernels, so one captured graph serves every value (a service with a
// user-controlled cfg_scale would otherwise capture and keep one frame-tail graph per distinct float)
static int set_cfg(vv_ctx* c, float cfg, cudaStream_t s) {
  if (memcmp(&cfg, &c->cfg_last, sizeof(float)) == 0) return 0;
  set_float_kernel<<<1, 1, 0, s>>>(c->cfg_dev, cfg);
  CKL();
  c->launches++;
  c->cfg_last = cfg;
  return 0;
}

// The N-step sampler as ONE weight-stream program (vv_stream.cuh): per step 4 x (gate/up with AdaLN prologue -> raw sums; down with SwiGLU
// prologue, gated-residual epilogue) + final layer + noisy_images_proj whose prologue is the CFG / DPM-Solver++ update.  10 grid barriers
// per step instead of 10 kernels, and the TMA ring keeps streaming head weights across all of them.
static int sampler_stream_prog(vv_ctx* c, const float* noise, float* latent_out, const vv_ctx::StreamProg** out) {
  char key[256];
  snprintf(key, sizeof key, "samp:%p:%p:%d:%d:%p:%d", (const void*)noise, (void*)latent_out, c->n_steps, (int)c->sde, (const void*)c->step_noise,
           c->coef_version);       // solver coefficients are baked into the program
  auto it = c->sprogs.find(key);
  if (it != c->sprogs.end()) { *out = &it->second; return 0; }
  const auto& d = c->d;
  const int H = d.hidden_size, F = d.head_ffn_dim, B = d.max_batch, M = 2 * B, LH = d.head_layers, N = c->n_steps;
  const long long modld = (long long)(3 * LH + 2) * H;
  StreamBuilder b(c);
  auto dpm = [&](int i) {
    SDpm o;
    memset(&o, 0, sizeof o);
    if (i < 0) { o.z_in = c->s_z + B * 64; o.z_out = c->s_z; o.x0_in = c->s_x0 + B * 64; o.x0_out = c->s_x0; }
    else {
      o.z_in = c->s_z + (size_t)(i & 1) * B * 64; o.z_out = c->s_z + (size_t)((i + 1) & 1) * B * 64;
      o.x0_in = c->s_x0 + (size_t)(i & 1) * B * 64; o.x0_out = c->s_x0 + (size_t)((i + 1) & 1) * B * 64;
    }
    o.v = c->s_v; o.noise = noise; o.cfg_p = c->cfg_dev; o.step_noise = c->sde ? c->step_noise : nullptr;
    o.latent_out = (i == N - 1) ? latent_out : nullptr; o.step = i; o.B = B;
    if (i >= 0) o.c = c->coef_host[i];
    return o;
  };
  auto proj = [&](int i, bool sync) -> int {       // x = noisy_images_proj(z'), z' = solver update of step i (i = -1: the initial noise)
    SOp* o;
    RET(b.gemv(c->h_noisy, nullptr, nullptr, 0, c->s_hx, H, M, H, 64, sync, &o));
    o->pro = SP_DPM; o->store = 1; o->dpm = dpm(i);
    return 0;
  };
  float* gu[2] = {c->s_hgu, c->s_hgu + (size_t)M * 2 * F};
  RET(proj(-1, false));
  b.ops.back().init_dst = gu[0]; b.ops.back().init_n = (long long)M * 2 * F;
  for (int i = 0; i < N; ++i) {
    const float* mod = c->s_mod + (size_t)i * M * modld;
    for (int li = 0; li < LH; ++li) {
      const HeadLayer& hl = c->head[li];
      SOp* o;
      RET(b.gemv(hl.wgu, nullptr, c->s_hx, H, gu[li & 1], 2 * F, M, 2 * F, H, true, &o));
      o->pro = SP_ADALN; o->pro_w = hl.norm; o->pro_eps = d.head_rms_eps;
      o->pro_shift = mod + (size_t)li * 3 * H; o->pro_scale = mod + (size_t)li * 3 * H + H; o->pro_ld = modld;
      o->init_dst = gu[(li + 1) & 1]; o->init_n = (long long)M * 2 * F;           // the other buffer: its reader (down of li-1) is done
      RET(b.gemv(hl.wdown, nullptr, gu[li & 1], 2 * F, c->s_hx, H, M, H, F, true, &o));
      o->pro = SP_SWIGLU; o->alpha_kind = SA_GATE; o->alpha = mod + (size_t)li * 3 * H + 2 * H; o->lda = modld;
      if (li == 0) { o->init_dst = c->s_v; o->init_n = (long long)M * 64; }        // final layer of this step accumulates into s_v
    }
    SOp* o;
    RET(b.gemv(c->h_final, nullptr, c->s_hx, H, c->s_v, 64, M, 64, H, true, &o));
    o->pro = SP_ADALN; o->pro_w = nullptr; o->pro_eps = d.head_rms_eps;
    o->pro_shift = mod + (size_t)LH * 3 * H; o->pro_scale = mod + (size_t)LH * 3 * H + H; o->pro_ld = modld;
    RET(proj(i, true));
  }
  vv_ctx::StreamProg pr;
  RET(finish_stream(b, &pr));
  it = c->sprogs.emplace(key, pr).first;
  *out = &it->second;
  return 0;
}

static int enqueue_diffusion(const L& l, const float* cond, const float* noise, float* latent_out, const vv_ctx::StreamProg* sprog) {
  vv_ctx* c = l.c;
  const auto& d = c->d;
  const int H = d.hidden_size, B = d.max_batch, M = 2 * B, LH = d.head_layers, N = c->n_steps;
  if (N < 1) return fail(VV_ERR_STATE, "vv_set_diffusion_steps not called");
  const int modld = (3 * LH + 2) * H;
  if (c->sde && !c->step_noise) return fail(VV_ERR_STATE, "sde-dpmsolver++ needs vv_set_step_noise before sampling");
  GemvP p = mk(c->h_cond, nullptr, cond, H, c->s_condp, H, M, H, H);
  RET(linear(l, p));
  {
    const long long n = (long long)N * M * H;
    CK(launch_k(l, head_cond_prep_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, c->s_condp, c->temb, c->s_call, N, M, H));
  }
  // AdaLN modulation of ALL steps and layers in one tensor-core GEMM: c_all [N*M, H] x W_mod^T -> [N*M, (3L+2)H].
  // (the reference recomputes Linear(silu(c)) inside every head call, diffusion_head.py:159, 185; c depends only on (cond, t_i))
  p = mk(c->h_mod, nullptr, c->s_call, H, c->s_mod, modld, N * M, modld, H);
  RET(linear(l, p));
  if (sprog) return launch_stream(l, *sprog);
  CK(launch_k(l, dpm_update_proj_kernel, dim3(B, (H + 255) / 256), dim3(256), 0, c->s_z + B * 64, c->s_z, c->s_x0 + B * 64, c->s_x0, c->s_v, noise,
              c->coef_dev, -1, (const float*)c->cfg_dev, c->h_noisy, c->s_hx, nullptr, B, H, 1, (const float*)nullptr));
  L lh = l;
  if (c->l2_persist_bytes && c->head_slab_bytes) {
    lh.win_base = c->head_slab;
    lh.win_bytes = std::min(c->head_slab_bytes, c->l2_window_max);
    lh.win_ratio = std::min(1.0f, (float)c->l2_persist_bytes / (float)lh.win_bytes);
  }
  for (int i = 0; i < N; ++i) {
    std::vector<GemvP> g;
    head_step_gemvs(c, i, &g);
    for (auto& q : g) RET(linear(lh, q));
    const bool last = (i == N - 1);
    const float* z_in = c->s_z + (size_t)(i & 1) * B * 64; float* z_out = c->s_z + (size_t)((i + 1) & 1) * B * 64;
    const float* x0_in = c->s_x0 + (size_t)(i & 1) * B * 64; float* x0_out = c->s_x0 + (size_t)((i + 1) & 1) * B * 64;
    CK(launch_k(l, dpm_update_proj_kernel, dim3(B, last ? 1 : (H + 255) / 256), dim3(256), 0, z_in, z_out, x0_in, x0_out, (const float*)c->s_v, noise,
                (const DpmCoef*)c->coef_dev, i, (const float*)c->cfg_dev, (const bf16*)c->h_noisy, c->s_hx, last ? latent_out : (float*)nullptr, B, H,
                last ? 0 : 1, c->sde ? c->step_noise : (const float*)nullptr));
  }
  return 0;
}

extern "C" int vv_diffusion_sample(vv_ctx* c, const float* cond, const float* noise, const int32_t* active, float cfg, float* latent_out,
                                   void* stream) {
  (void)active;   // rows are independent; inactive rows are computed and ignored (static shapes keep the program graph-replayable)
  if (!c || !c->finalized) return fail(VV_ERR_STATE, "not finalized");
  CK(cudaSetDevice(c->device));
  RET(set_cfg(c, cfg, (cudaStream_t)stream));
  char key[256];
  snprintf(key, sizeof key, "diff:%p:%p:%p", (const void*)cond, (const void*)noise, (void*)latent_out);
  const vv_ctx::StreamProg* sprog = nullptr;      // built outside stream capture (it allocates and copies)
  if (c->use_stream & 1) RET(sampler_stream_prog(c, noise, latent_out, &sprog));
  return run_cached(c, key, (cudaStream_t)stream, [&](const L& l) { return enqueue_diffusion(l, cond, noise, latent_out, sprog); });
}

// ------------------------------------------------------------------------------------------------
// a-5 / a-6: streaming codec
// ------------------------------------------------------------------------------------------------
static int assemble(const L& l, const float* src, const float* hist, float* win, float* next, int B, int T, int ctx, int C,
                    const float* norm_w, float eps, float alpha, float beta) {
  const int rows = B * (ctx + T);
  if (C >= 512) CK(launch_k(l, assemble_window_block_kernel, dim3(rows), dim3(256), 0, src, hist, win, next, B, T, ctx, C, norm_w, eps, alpha, beta));
  else CK(launch_k(l, assemble_window_kernel, dim3((rows + 7) / 8), dim3(256), 0, src, hist, win, next, B, T, ctx, C, norm_w, eps, alpha, beta));
  return 0;
}

// one Block1D over x [B,T,C] (in `xin`), result in `xout` (may not alias xin)
static int enqueue_block(const L& l, const Block& b, const float* xin, float* xout, int B, int T, float eps) {
  vv_ctx* c = l.c;
  const int C = b.C, M = B * T;
  RET(assemble(l, xin, b.hist, c->s_win, b.next, B, T, 6, C, b.norm_w, eps, 1.f, 0.f));
  {
    const long long n = (long long)M * C;
    CK(launch_k(l, dwconv_res_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, xin, c->s_win, b.dw_w, b.dw_b, b.gamma, xout, B, T, C));
  }
  GemvP p;
  if (M < c->mma_min_rows || (c->ring_rms && c->mma_ring && C <= MR_MAXK_NORM && c->use_tc5 != 2)) {
    p = mk(b.w1, b.b1, xout, C, c->s_u, 4 * C, M, 4 * C, C);
    p.pro = PRO_RMSNORM; p.pro_w = b.ffn_norm_w; p.pro_eps = eps; p.epi = EPI_GELU;
    RET(linear(l, p));
  } else {
    if (C >= 512) CK(launch_k(l, rows_norm_block_kernel, dim3(M), dim3(256), 0, xout, b.ffn_norm_w, c->s_xn, C, eps));
    else CK(launch_k(l, rows_norm_kernel, dim3((M + 7) / 8), dim3(256), 0, xout, b.ffn_norm_w, c->s_xn, M, C, eps));
    p = mk(b.w1, b.b1, c->s_xn, C, c->s_u, 4 * C, M, 4 * C, C);
    p.epi = EPI_GELU;
    RET(linear(l, p));
  }
  p = mk(b.w2, b.b2, c->s_u, 4 * C, xout, C, M, C, 4 * C);
  p.epi = EPI_GAMMA_RESID; p.epi_a = b.ffn_gamma; p.res = xout; p.ldres = C;
  return linear(l, p);
}

static int conv_apply(const L& l, const ConvL& cv, const float* win, float* y, int B, int T_out, int T_in) {
  // rows (b,t) read window rows [t*stride, t*stride + k) of a [ctx+T_in, Cin] window
  RowMap xm; xm.T = T_out; xm.bs = (long long)(cv.ctx + T_in) * cv.Cin; xm.rs = (long long)cv.stride * cv.Cin;
  const int M = B * T_out;
  if (cv.wf) {
    const long long n = (long long)M * cv.N;
    if (cv.K >= 64) CK(launch_k(l, conv_warp_kernel, dim3((unsigned)((n + 7) / 8)), dim3(256), 0, cv.wf, cv.bias, win, xm, y, M, cv.N, cv.K));
    else CK(launch_k(l, conv_naive_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, cv.wf, cv.bias, win, xm, y, M, cv.N, cv.K));
    return 0;
  }
  GemvP p = mk(cv.w, cv.bias, win, 0, y, cv.N, M, cv.N, cv.K);
  p.xmap = xm;
  return linear(l, p);
}

struct ScopedMinRows {     // the codec stages may use a different GEMV/GEMM row threshold than the LM and the sampler
  vv_ctx* c; int saved;
  ScopedMinRows(vv_ctx* c_, int v) : c(c_), saved(c_->mma_min_rows) { c->mma_min_rows = v; }
  ~ScopedMinRows() { c->mma_min_rows = saved; }
};

// ---- codec stages with B*T <= 8 rows (96 % of the codec's weights: 2048- and 1024-wide blocks, stem / up- / down-sampling convolutions
// next to them) as weight-stream programs (vv_stream.cuh: SP_WINDOW, SP_MIXER): two stages per Block1D instead of five kernels ------------
// Buffers: rows ping-pong between X[0] / X[1], FFN hidden sums between U[0] / U[1].  A block reads X[a], its owner CTA stores x1 into X[a^1]
// and the second linear accumulates there; a convolution accumulates into a ZEROED buffer -- zero-fill jobs ride on the stage after the
// buffer's last reader.
static bool codec_stream_stage(const vv_ctx* c, const Codec& k, int i) {
  const long long rows = (long long)c->d.max_batch * k.T[i];
  return (c->use_stream & 8) && k.T[i] <= 8 && rows <= 32;
}
static long long codec_x_floats(const vv_ctx* c) { return (long long)c->d.max_batch * 8192; }
static long long codec_u_floats(const vv_ctx* c) { return (long long)c->d.max_batch * 32768; }
static void window_op(SOp* o, const ConvL& cv, const float* src, int T_in, int T_out, int row_stride, float alpha, float beta) {
  o->pro = SP_WINDOW;
  SCodec& w = o->cod;
  memset(&w, 0, sizeof w);
  w.hist = cv.hist; w.next = cv.next; w.src = src; w.ctx = cv.ctx; w.T_in = T_in; w.T_out = T_out; w.stride = row_stride; w.cin = cv.Cin;
  w.alpha = alpha; w.beta = beta;
}
struct CodecBufs {
  float* X[2]; float* U[2];
  int cur = 0, ub = 0;
  explicit CodecBufs(vv_ctx* c) { X[0] = c->s_cx; X[1] = c->s_cx + codec_x_floats(c); U[0] = c->s_cu; U[1] = c->s_cu + codec_u_floats(c); }
};
// all blocks of one stage; rows enter in X[cur] and leave in X[cur]; if `zero_for_next` the buffer the FOLLOWING convolution accumulates
// into (= the input of the stage's last block) is zero-filled on that block's second linear
static int stage_ops(StreamBuilder& b, vv_ctx* c, CodecBufs& cb, const std::vector<Block>& blocks, int B, int T, bool zero_for_next, float* extra_zero,
                     long long extra_n) {
  for (size_t j = 0; j < blocks.size(); ++j) {
    const Block& blk = blocks[j];
    const int C = blk.C, M = B * T;
    SOp* o;
    {                                                          // mixer stage: x (X[cur]) -> x1 (X[cur^1]) + next history, channels over CTAs
      SOp& mx = b.push(SK_MIX, true);
      mx.M = M; mx.K = C; mx.x = cb.X[cb.cur]; mx.ldx = C;
      SCodec& w = mx.cod;
      memset(&w, 0, sizeof w);
      w.hist = blk.hist; w.next = blk.next; w.T_out = T; w.norm_w = blk.norm_w; w.dw_w = blk.dw_w; w.dw_b = blk.dw_b; w.gamma = blk.gamma;
      w.x1_out = cb.X[cb.cur ^ 1]; w.eps = c->d.codec_eps;
    }
    RET(b.gemv(blk.w1, blk.b1, cb.X[cb.cur ^ 1], C, cb.U[cb.ub], 4 * C, M, 4 * C, C, true, &o));
    o->pro = SP_RMSNORM; o->pro_w = blk.ffn_norm_w; o->pro_eps = c->d.codec_eps;
    o->init_dst = cb.U[cb.ub ^ 1]; o->init_n = codec_u_floats(c);          // hidden-sum buffer of the NEXT block (its reader finished two stages ago)
    RET(b.gemv(blk.w2, blk.b2, cb.U[cb.ub], 4 * C, cb.X[cb.cur ^ 1], C, M, C, 4 * C, true, &o));
    o->pro = SP_GELU; o->alpha_kind = SA_GAMMA; o->alpha = blk.ffn_gamma;
    if (j + 1 == blocks.size()) {
      if (zero_for_next) { o->init_dst = cb.X[cb.cur]; o->init_n = codec_x_floats(c); }
      if (extra_zero) { o->init2_dst = extra_zero; o->init2_n = extra_n; }
    }
    cb.cur ^= 1; cb.ub ^= 1;
  }
  return 0;
}

// decoder front: stem conv + the leading stages with <= 8 rows; *n_front = stages covered, *out_x = where the last one leaves its rows
static int dec_front_prog(vv_ctx* c, const float* latent, const vv_ctx::StreamProg** out, int* n_front, float** out_x) {
  Codec& k = c->dec;
  const auto& d = c->d;
  const int B = d.max_batch;
  int nf = 0;
  while (nf < d.n_stages - 1 && codec_stream_stage(c, k, nf)) ++nf;
  *n_front = nf; *out = nullptr; *out_x = nullptr;
  if (nf == 0) return 0;
  char key[96];
  snprintf(key, sizeof key, "decf:%p", (const void*)latent);
  {
    auto hit = c->sprogs.find(key);
    if (hit != c->sprogs.end()) { *out = &hit->second; *out_x = c->dec_front_x; return 0; }
  }
  CodecBufs cb(c);
  StreamBuilder b(c);
  b.nop(false, cb.X[0], codec_x_floats(c));
  b.ops.back().init2_dst = cb.U[0]; b.ops.back().init2_n = codec_u_floats(c);
  SOp* o;
  RET(b.gemv(k.convs[0].w, k.convs[0].bias, nullptr, 0, cb.X[0], k.convs[0].N, B, k.convs[0].N, k.convs[0].K, true, &o));
  window_op(o, k.convs[0], latent, 1, 1, 1, 1.0f / c->speech_scale, -c->speech_bias);
  for (int i = 0; i < nf; ++i) {
    if (i > 0) {                                  // transposed conv into stage i: row (b, t) = [previous frame | frame t] -> s * Co outputs
      const ConvL& cv = k.convs[i];
      const int Tin = k.T[i - 1];
      RET(b.gemv(cv.w, cv.bias, nullptr, 0, cb.X[cb.cur ^ 1], cv.N, B * Tin, cv.N, cv.K, true, &o));
      window_op(o, cv, cb.X[cb.cur], Tin, Tin, 1, 1.f, 0.f);
      cb.cur ^= 1;
    }
    RET(stage_ops(b, c, cb, k.stages[i], B, k.T[i], i + 1 < nf, nullptr, 0));
  }
  *out_x = c->dec_front_x = cb.X[cb.cur];
  auto it = c->sprogs.find(key);
  if (it == c->sprogs.end()) {
    vv_ctx::StreamProg pr;
    RET(finish_stream(b, &pr));
    it = c->sprogs.emplace(key, pr).first;
  }
  *out = &it->second;
  return 0;
}

// encoder back: the trailing stages with <= 8 rows, each behind its strided conv, + the head conv.  `xin` = rows entering the conv of the
// first covered stage (produced by the kernel-per-stage path).
static int enc_back_first(const vv_ctx* c) {
  const Codec& k = c->enc;
  int f = c->d.n_stages;
  while (f > 1 && codec_stream_stage(c, k, f - 1)) --f;
  return f;
}
static int enc_back_prog(vv_ctx* c, const float* xin, float* feat, const vv_ctx::StreamProg** out) {
  Codec& k = c->enc;
  const auto& d = c->d;
  const int B = d.max_batch, ns = d.n_stages, f = enc_back_first(c);
  *out = nullptr;
  if (f >= ns) return 0;
  char key[96];
  snprintf(key, sizeof key, "encb:%p:%p", (const void*)xin, (void*)feat);
  auto it = c->sprogs.find(key);
  if (it != c->sprogs.end()) { *out = &it->second; return 0; }
  CodecBufs cb(c);
  StreamBuilder b(c);
  b.nop(false, cb.X[0], codec_x_floats(c));
  b.ops.back().init2_dst = cb.U[0]; b.ops.back().init2_n = codec_u_floats(c);
  SOp* o;
  for (int i = f; i < ns; ++i) {
    const ConvL& cv = k.convs[i];
    const int Tin = k.T[i - 1], Tout = k.T[i];
    float* dst = (i == f) ? cb.X[0] : cb.X[cb.cur ^ 1];
    RET(b.gemv(cv.w, cv.bias, nullptr, 0, dst, cv.N, B * Tout, cv.N, cv.K, true, &o));      // strided conv: window rows [t*r, t*r + 2r)
    window_op(o, cv, i == f ? xin : cb.X[cb.cur], Tin, Tout, cv.stride, 1.f, 0.f);
    if (i > f) cb.cur ^= 1;
    const bool last = (i + 1 == ns);
    RET(stage_ops(b, c, cb, k.stages[i], B, Tout, !last, last ? feat : nullptr, (long long)B * d.semantic_vae_dim));
  }
  const ConvL& hd = k.convs[ns];
  RET(b.gemv(hd.w, hd.bias, nullptr, 0, feat, hd.N, B, hd.N, hd.K, true, &o));
  window_op(o, hd, cb.X[cb.cur], 1, 1, 1, 1.f, 0.f);
  vv_ctx::StreamProg pr;
  RET(finish_stream(b, &pr));
  it = c->sprogs.emplace(key, pr).first;
  *out = &it->second;
  return 0;
}

static int enqueue_decode(const L& l, const float* latent, const int32_t* active, float* audio, const vv_ctx::StreamProg* front = nullptr,
                          int n_front = 0, float* front_x = nullptr) {
  vv_ctx* c = l.c;
  ScopedMinRows scoped(c, c->codec_mma_min_rows);
  const auto& d = c->d;
  Codec& k = c->dec;
  const int B = d.max_batch, ns = d.n_stages;
  float *xa = c->s_xa, *xb = c->s_xb;
  if (front) {
    RET(launch_stream(l, *front));              // stem + stages [0, n_front) through the weight-stream kernel
    xa = front_x;
  } else {
    n_front = 0;
    // stem: window over the last 7 (un-scaled) latent frames; un-scaling latent/scale - bias (:636) is folded in
    RET(assemble(l, latent, k.convs[0].hist, c->s_win, k.convs[0].next, B, 1, 6, 64, nullptr, 0.f, 1.0f / c->speech_scale, -c->speech_bias));
    RET(conv_apply(l, k.convs[0], c->s_win, xa, B, 1, 1));
  }
  for (int i = n_front; i < ns; ++i) {
    if (i > 0) {
      const ConvL& cv = k.convs[i];
      const int Tin = k.T[i - 1];
      RET(assemble(l, xa, cv.hist, c->s_win, cv.next, B, Tin, 1, cv.Cin, nullptr, 0.f, 1.f, 0.f));
      RowMap xm; xm.T = Tin; xm.bs = (long long)(1 + Tin) * cv.Cin; xm.rs = cv.Cin;
      float* dst = (xa == c->s_xa) ? c->s_xb : c->s_xa;
      GemvP p = mk(cv.w, cv.bias, c->s_win, 0, dst, cv.N, B * Tin, cv.N, cv.K);
      p.xmap = xm;
      RET(linear(l, p));
      xa = dst; xb = (xa == c->s_xa) ? c->s_xb : c->s_xa;
    }
    for (const Block& b : k.stages[i]) { RET(enqueue_block(l, b, xa, xb, B, k.T[i], d.codec_eps)); std::swap(xa, xb); }
  }
  const ConvL& hd = k.convs[ns];
  const int T = k.T[ns - 1];
  RET(assemble(l, xa, hd.hist, c->s_win, hd.next, B, T, 6, hd.Cin, nullptr, 0.f, 1.f, 0.f));
  RET(conv_apply(l, hd, c->s_win, audio, B, T, T));
  CK(launch_k(l, advance_kernel, dim3(k.n_segs, B, ADV_SLICES), dim3(256), 0, k.segs_dev, active));
  return 0;
}

// stages [0, n_stream_first) of the semantic encoder through the kernel-per-stage path; returns where their rows are (the stream program of
// the remaining stages was built against that pointer)
static float* encode_front_out(vv_ctx* c, int first) {
  // the ping-pong below is deterministic: stem -> s_xa, every conv and every block swaps
  const Codec& k = c->enc;
  bool in_a = true;
  for (int i = 0; i < first; ++i) {
    if (i > 0) in_a = !in_a;
    if (k.stages[i].size() & 1) in_a = !in_a;
  }
  return in_a ? c->s_xa : c->s_xb;
}
static int enqueue_encode(const L& l, const float* audio, const int32_t* active, float* feat, const vv_ctx::StreamProg* back = nullptr) {
  vv_ctx* c = l.c;
  ScopedMinRows scoped(c, c->codec_mma_min_rows);
  const auto& d = c->d;
  Codec& k = c->enc;
  const int B = d.max_batch, ns = d.n_stages;
  const int stop = back ? enc_back_first(c) : ns;
  float *xa = c->s_xa, *xb = c->s_xb;
  int hop = k.T[0];
  RET(assemble(l, audio, k.convs[0].hist, c->s_win, k.convs[0].next, B, hop, 6, 1, nullptr, 0.f, 1.f, 0.f));
  RET(conv_apply(l, k.convs[0], c->s_win, xa, B, hop, hop));
  for (int i = 0; i < stop; ++i) {
    if (i > 0) {
      const ConvL& cv = k.convs[i];
      const int Tin = k.T[i - 1];
      RET(assemble(l, xa, cv.hist, c->s_win, cv.next, B, Tin, cv.ctx, cv.Cin, nullptr, 0.f, 1.f, 0.f));
      RET(conv_apply(l, cv, c->s_win, xb, B, k.T[i], Tin));
      std::swap(xa, xb);
    }
    for (const Block& b : k.stages[i]) { RET(enqueue_block(l, b, xa, xb, B, k.T[i], d.codec_eps)); std::swap(xa, xb); }
  }
  if (back) {
    if (xa != encode_front_out(c, stop)) return fail(VV_ERR_STATE, "encoder hand-off buffer mismatch");
    RET(launch_stream(l, *back));
  } else {
    const ConvL& hd = k.convs[ns];
    RET(assemble(l, xa, hd.hist, c->s_win, hd.next, B, 1, 6, hd.Cin, nullptr, 0.f, 1.f, 0.f));
    RET(conv_apply(l, hd, c->s_win, feat, B, 1, 1));
  }
  CK(launch_k(l, advance_kernel, dim3(k.n_segs, B, ADV_SLICES), dim3(256), 0, k.segs_dev, active));
  return 0;
}

static int enqueue_connect(const L& l, const float* latent, const float* sem, const int32_t* active, float* embeds) {
  vv_ctx* c = l.c;
  const auto& d = c->d;
  const int H = d.hidden_size, B = d.max_batch;
  GemvP p = mk(c->ca_fc1, c->ca_b1, latent, 64, c->s_c1, H, B, H, 64);
  RET(linear(l, p));
  p = mk(c->ca_fc2, c->ca_b2, c->s_c1, H, c->s_e, H, B, H, H);
  p.pro = PRO_RMSNORM; p.pro_w = c->ca_n; p.pro_eps = 1e-6f;
  RET(linear(l, p));
  p = mk(c->cs_fc1, c->cs_b1, sem, d.semantic_vae_dim, c->s_c1, H, B, H, d.semantic_vae_dim);
  RET(linear(l, p));
  p = mk(c->cs_fc2, c->cs_b2, c->s_c1, H, c->s_e, H, B, H, H);
  p.pro = PRO_RMSNORM; p.pro_w = c->cs_n; p.pro_eps = 1e-6f; p.epi = EPI_RESID; p.res = c->s_e; p.ldres = H;
  RET(linear(l, p));
  CK(launch_k(l, select_embeds_kernel, dim3(B), dim3(256), 0, embeds, c->s_e, active, B, H));
  return 0;
}

extern "C" int vv_codec_decode_frame(vv_ctx* c, const float* latent, const int32_t* active, float* audio_out, void* stream) {
  if (!c || !c->finalized) return fail(VV_ERR_STATE, "not finalized");
  CK(cudaSetDevice(c->device));
  char key[256];
  snprintf(key, sizeof key, "dec:%p:%p:%p", (const void*)latent, (const void*)active, (void*)audio_out);
  const vv_ctx::StreamProg* front = nullptr; int nf = 0; float* fx = nullptr;
  RET(dec_front_prog(c, latent, &front, &nf, &fx));
  return run_cached(c, key, (cudaStream_t)stream, [&](const L& l) { return enqueue_decode(l, latent, active, audio_out, front, nf, fx); });
}
extern "C" int vv_semantic_encode_frame(vv_ctx* c, const float* audio, const int32_t* active, float* feat_out, void* stream) {
  if (!c || !c->finalized) return fail(VV_ERR_STATE, "not finalized");
  CK(cudaSetDevice(c->device));
  char key[256];
  snprintf(key, sizeof key, "enc:%p:%p:%p", (const void*)audio, (const void*)active, (void*)feat_out);
  const vv_ctx::StreamProg* back = nullptr;
  RET(enc_back_prog(c, encode_front_out(c, enc_back_first(c)), feat_out, &back));
  return run_cached(c, key, (cudaStream_t)stream, [&](const L& l) { return enqueue_encode(l, audio, active, feat_out, back); });
}
extern "C" int vv_connect(vv_ctx* c, const float* latent, const float* sem, const int32_t* active, float* embeds, void* stream) {
  if (!c || !c->finalized) return fail(VV_ERR_STATE, "not finalized");
  CK(cudaSetDevice(c->device));
  L l{c, (cudaStream_t)stream};
  return enqueue_connect(l, latent, sem, active, embeds);
}
extern "C" int vv_frame_tail(vv_ctx* c, const float* hidden, const float* noise, const int32_t* active, float cfg, float* latent_out,
                             float* audio_out, float* embeds, void* stream) {
  if (!c || !c->finalized) return fail(VV_ERR_STATE, "not finalized");
  CK(cudaSetDevice(c->device));
  char key[320];
  RET(set_cfg(c, cfg, (cudaStream_t)stream));
  snprintf(key, sizeof key, "tail:%p:%p:%p:%p:%p:%p", (const void*)hidden, (const void*)noise, (const void*)active, (void*)latent_out,
           (void*)audio_out, (void*)embeds);
  const vv_ctx::StreamProg* sprog = nullptr;
  if (c->use_stream & 1) RET(sampler_stream_prog(c, noise, latent_out, &sprog));
  const vv_ctx::StreamProg *front = nullptr, *back = nullptr; int nf = 0; float* fx = nullptr;
  RET(dec_front_prog(c, latent_out, &front, &nf, &fx));
  RET(enc_back_prog(c, encode_front_out(c, enc_back_first(c)), c->s_feat, &back));
  return run_cached(c, key, (cudaStream_t)stream, [&](const L& l) {
    RET(enqueue_diffusion(l, hidden, noise, latent_out, sprog));
    RET(enqueue_decode(l, latent_out, active, audio_out, front, nf, fx));
    RET(enqueue_encode(l, audio_out, active, c->s_feat, back));
    return enqueue_connect(l, latent_out, c->s_feat, active, embeds);
  });
}

struct Rows { int v[16]; };
__global__ void state_zero_val_kernel(const StateSeg* __restrict__ segs, Rows r) {
  const int b = r.v[blockIdx.y];
  const StateSeg s = segs[blockIdx.x];
  float* dd = s.hist + (size_t)b * s.n;
  for (int i = threadIdx.x; i < s.n; i += blockDim.x) dd[i] = 0.f;
}
extern "C" int vv_codec_state_zero(vv_ctx* c, const int32_t* rows_host, int n, void* stream) {
  if (!c || !c->finalized) return fail(VV_ERR_STATE, "not finalized");
  if (n < 1) return 0;
  if (n > c->d.max_batch) return fail(VV_ERR_INVALID, "too many rows");
  Rows r;
  for (int i = 0; i < 16; ++i) r.v[i] = i < n ? rows_host[i] : 0;
  for (int i = 0; i < n; ++i) if (r.v[i] < 0 || r.v[i] >= c->d.max_batch) return fail(VV_ERR_INVALID, "row %d out of range", r.v[i]);
  cudaStream_t s = (cudaStream_t)stream;
  state_zero_val_kernel<<<dim3(c->dec.n_segs, n), 256, 0, s>>>(c->dec.segs_dev, r);
  CKL();
  state_zero_val_kernel<<<dim3(c->enc.n_segs, n), 256, 0, s>>>(c->enc.segs_dev, r);
  CKL();
  c->launches += 2;
  return 0;
}
extern "C" int vv_codec_state_reset(vv_ctx* c, void* stream) {
  if (!c || !c->finalized) return fail(VV_ERR_STATE, "not finalized");
  int rows[8];
  for (int i = 0; i < c->d.max_batch; ++i) rows[i] = i;
  return vv_codec_state_zero(c, rows, c->d.max_batch, stream);
}

extern "C" int vv_debug_gemv(vv_ctx* c, const void* w, const float* bias, const float* x, float* y, int M, int N, int K, int prologue,
                             const float* pro_w, float eps, int epilogue, void* stream) {
  if (!c) return fail(VV_ERR_INVALID, "null ctx");
  CK(cudaSetDevice(c->device));
  L l{c, (cudaStream_t)stream};
  const int ldy = epilogue == EPI_SWIGLU ? N / 2 : N;
  GemvP p = mk((const bf16*)w, bias, x, K, y, ldy, M, N, K);
  p.pro = prologue; p.pro_w = pro_w; p.pro_eps = eps; p.epi = epilogue;
  if (epilogue == EPI_RESID) { p.res = y; p.ldres = N; }
  return linear(l, p);
}

// one linear through the weight-stream kernel (tests): y = [y +] alpha * (W pro(x) + bias); pro = SPro, alpha_kind = SAlpha
extern "C" int vv_debug_stream_gemv(vv_ctx* c, const void* w, const float* bias, const float* x, float* y, int M, int N, int K, int pro,
                                    const float* pro_w, float eps, int alpha_kind, const float* alpha, int accumulate, void* stream) {
  if (!c || !c->finalized) return fail(VV_ERR_STATE, "not finalized");
  CK(cudaSetDevice(c->device));
  StreamBuilder b(c);
  b.fresh_weights = true;
  if (!accumulate) b.nop(false, y, (long long)M * N);
  SOp* o;
  RET(b.gemv((const bf16*)w, bias, x, pro == SP_SWIGLU ? 2LL * K : (long long)K, y, N, M, N, K, !accumulate, &o));
  o->pro = pro; o->pro_w = pro_w; o->pro_eps = eps; o->alpha_kind = alpha_kind; o->alpha = alpha; o->lda = N;
  vv_ctx::StreamProg pr;
  RET(finish_stream(b, &pr));
  L l{c, (cudaStream_t)stream};
  RET(launch_stream(l, pr));
  CK(cudaStreamSynchronize((cudaStream_t)stream));
  dfree(c, &pr.ops); dfree(c, &pr.tmaps);
  for (bf16* t : b.owned) dfree(c, &t);
  if (c->st_diag_host[0]) return fail(VV_ERR_CUDA, "stream kernel watchdog: code %u cta %u thread %u a %u b %u c %u", c->st_diag_host[0], c->st_diag_host[1],
                                      c->st_diag_host[2], c->st_diag_host[3], c->st_diag_host[4], c->st_diag_host[5]);
  return 0;
}
// clock stamps of the last traced stream launch (VV_STREAM_TRACE=<cta>): out [n_ops][12] int64; returns the number of stages, 0 if tracing is off.
// Stage kinds / shapes are appended per stage in out_meta [n_ops][4] = {kind, N, K, prologue}.
extern "C" int vv_stream_trace_read2(vv_ctx* c, long long* out, int max_ops) {     // [n_ops][sm_count][2] barrier arrival / release (ns)
  if (!c || !c->st_trace2) return 0;
  CK(cudaDeviceSynchronize());
  const int n = std::min(max_ops, c->st_trace_last_ops);
  CK(cudaMemcpy(out, c->st_trace2, (size_t)n * c->sm_count * 2 * sizeof(long long), cudaMemcpyDeviceToHost));
  return c->sm_count;
}
extern "C" int vv_stream_trace_read(vv_ctx* c, long long* out, int* out_meta, int max_ops, const char* prog_prefix) {
  if (!c || !c->st_trace) return 0;
  CK(cudaDeviceSynchronize());
  const int n = std::min(max_ops, c->st_trace_last_ops);
  CK(cudaMemcpy(out, c->st_trace, (size_t)n * ST_TRACE * sizeof(long long), cudaMemcpyDeviceToHost));
  for (auto& kv : c->sprogs) {
    if (kv.first.rfind(prog_prefix, 0) != 0 || kv.second.n_ops != c->st_trace_last_ops) continue;
    std::vector<SOp> ops(kv.second.n_ops);
    CK(cudaMemcpy(ops.data(), kv.second.ops, ops.size() * sizeof(SOp), cudaMemcpyDeviceToHost));
    for (int i = 0; i < n; ++i) { out_meta[4 * i] = ops[i].kind; out_meta[4 * i + 1] = ops[i].N; out_meta[4 * i + 2] = ops[i].K; out_meta[4 * i + 3] = ops[i].pro; }
    break;
  }
  return n;
}
// tcgen05.mma 128 x nB x 16 rate micro-benchmark (vv_stream.cuh: mma_rate_kernel); cycles_out = median over CTAs of the loop's SM cycles
extern "C" int vv_debug_mma_rate(vv_ctx* c, int n, int nB, int mode, int nacc, int ctas, long long* cycles_out) {
  if (!c || !c->finalized) return fail(VV_ERR_STATE, "not finalized");
  CK(cudaSetDevice(c->device));
  if (nacc < 1 || nacc * nB > 512 || nB > 256) return fail(VV_ERR_INVALID, "nacc * nB must be <= 512, nB <= 256");
  long long* d = nullptr;
  CK(cudaMalloc(&d, sizeof(long long) * 2 * ctas));
  CK(cudaMemset(d, 0, sizeof(long long) * 2 * ctas));
  const int smem = 8 * ST_TILE + 32768 + 1024;
  CK(cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  mma_rate_kernel<<<ctas, 128, smem>>>(n, nB, mode, nacc, d);
  CKL();
  CK(cudaDeviceSynchronize());
  std::vector<long long> h(2 * ctas);
  CK(cudaMemcpy(h.data(), d, sizeof(long long) * 2 * ctas, cudaMemcpyDeviceToHost));
  cudaFree(d);
  std::sort(h.begin(), h.begin() + ctas);
  std::sort(h.begin() + ctas, h.end());
  cycles_out[0] = h[ctas / 2];              // whole loop incl. completion
  cycles_out[1] = h[ctas + ctas / 2];       // issue loop only
  return 0;
}
// watchdog record of the last stream kernel that trapped: out[0] = code (0 = none), out[1..5] = cta, thread, stage, iteration, extra
extern "C" int vv_stream_diag(vv_ctx* c, unsigned* out6) {
  if (!c || !c->st_diag_host) return fail(VV_ERR_STATE, "no context");
  for (int i = 0; i < 6; ++i) out6[i] = c->st_diag_host[i];
  return 0;
}

// time `iters` grid barriers of a cooperative launch with `per_sm` CTAs per SM (debug / profiling aid)
extern "C" int vv_debug_barrier_bench(vv_ctx* c, int iters, int per_sm, float* ms_out) {
  if (!c || !c->finalized) return fail(VV_ERR_STATE, "not finalized");
  CK(cudaSetDevice(c->device));
  cudaStream_t s;
  CK(cudaStreamCreate(&s));
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.gridDim = dim3(c->sm_count * per_sm); cfg.blockDim = dim3(256); cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative; attr[0].val.cooperative = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  float* sink = c->s_v;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaLaunchKernelEx(&cfg, barrier_bench_kernel, c->gridbar, 10, sink));
  CK(cudaEventRecord(e0, s));
  CK(cudaLaunchKernelEx(&cfg, barrier_bench_kernel, c->gridbar, iters, sink));
  CK(cudaEventRecord(e1, s));
  CK(cudaStreamSynchronize(s));
  CK(cudaEventElapsedTime(ms_out, e0, e1));
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaStreamDestroy(s);
  return 0;
}
