// libstz.so — C ABI (include/stz.h) of the B200-native StyleTTS-ZS inference hot path.
// Host orchestration: weight upload/conversion, workspace arena, conditioning prep, the denoiser
// evaluation loop captured as one CUDA graph per (B, T, P, evals, sampler), the duration
// predictor, and the host-buffer end-to-end entry point.  Kernels live in the *.cuh files.
#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/stz.h"
#include "attention.cuh"
#include "elementwise.cuh"
#include "gemm.cuh"
#include "gemm2.cuh"
#include "gemm_ln3.cuh"
#include "philox.cuh"
#include "predictor.cuh"
#include "predictor_tc.cuh"
#include "regulator.cuh"
#include "stz_layout.h"

using namespace stz;
typedef __nv_bfloat16 bf16;

static std::string g_create_error;

// ------------------------------------------------------------------------------------------
// handle
// ------------------------------------------------------------------------------------------
struct Workspace {
  int B = 0, T = 0, P = 0, E = 0, noise_slices = 0;
  size_t mod_rows = 0;        // rows of `mod` ([n_mod] floats each): max over calls of (hoisted evaluations x 2B)
  bool mod_hoisted = false;   // this call's AdaLN modulations of ALL evaluations live in `mod` ([E][2B][n_mod])
  char* base = nullptr;
  size_t bytes = 0;
  // conditioning
  bf16 *text_bf, *prompt_bf, *ctx_text, *ctx_prompt, *kv_text, *kv_prompt, *cvec;
  float *pool_text, *pool_prompt, *pt, *pp, *ctx_pre, *tfeat, *t1, *temb, *coef;
  float *gfeat, *g1, *gemb;   // guidance-scale embedding of the guidance-conditioned student
  // denoiser
  float *mod, *x, *xmid, *h, *noise;
  bf16 *xin, *u, *u3, *qkv, *att, *ffh;
  // predictor
  float *sq, *sk, *sv, *sa, *stok, *G, *xa, *xb, *gb, *skv;
  bf16 *pa, *ps3, *pq, *pstyle3, *psa3;
  int *lens, *perm;
  uint8_t* tile_needed;   // predictor: per 128-row tile of [B*T], 1 = holds a valid token
  uint8_t* tile_needed_ctx;   // sampler: the same for the context rows [B*T text ; B*P prompt] (prompt tiles always 1)
  // graph-baked mask staging (device-pointer calls)
  uint8_t *st_tmask, *st_pmask;
  // host-call staging (stz_synthesize_host / _submit): one set per pipeline slot, so the H2D copies of call i+1 run
  // while call i computes
  struct HostStage {
    float *text, *prompt, *noise, *style;
    uint8_t *tmask, *pmask;
    int32_t* dur;
  } hs[2];
};

constexpr int STZ_MAX_CHAINS = 8;
constexpr int STZ_GUARD_BYTE = 0xA5;
// The sigma schedule is known before the loop, so the AdaLN modulations c[e] · Wmod^T of every evaluation can be one GEMM
// (M = E * 2B) instead of E small ones at the head of each evaluation's dependency chain.  E * 2B * n_mod floats: 78 MB at
// cfg2 (E = 4), 620 MB at cfg3 (64 teacher evaluations, B = 32); beyond 1 GB the per-evaluation GEMM is kept.
constexpr size_t STZ_HOIST_MOD_MAX_BYTES = (size_t)1 << 30;
static bool hoist_mod(const stz_config& c, int B, int E) {
  const size_t bytes = (size_t)E * 2 * B * (9 * c.n_layers + 2) * c.d_model * sizeof(float);
  return bytes <= STZ_HOIST_MOD_MAX_BYTES;
}

struct stz_handle {
  stz_config cfg;
  int device = 0;
  std::string err;
  std::vector<WeightEntry> layout;
  std::map<std::string, size_t> off;
  size_t n_floats = 0;
  float* w32 = nullptr;   // fp32 blob on device
  bf16* wbf = nullptr;    // the same blob rounded to bf16 (same offsets): tensor-core operands
  float *ctx_text_b = nullptr, *ctx_prompt_b = nullptr;  // bias + type embedding
  bf16* kv_null = nullptr;                               // [1, L*2d] K/V of the null-prompt token
  bf16* wkv_all = nullptr;                               // [L*2d, d] every layer's context K/V projection, stacked (one GEMM per call)
  float* bkv_all = nullptr;                              // [L*2d]
  bf16 *w_in3 = nullptr, *w_out3 = nullptr;              // split-bf16 [hi | hi | lo] input / output projection weights
  float* whhT = nullptr;                                 // [n_lstm][2][h][4h]  (k-major: generic kernel)
  float* whh = nullptr;                                  // [n_lstm][2][4h][h]  (row-major: cluster kernel)
  float* lstm_b = nullptr;                               // [n_lstm][2][4h] = b_ih + b_hh
  // predictor GEMMs on tcgen05 at fp32-grade precision: split-bf16 weights [hi | hi | lo] (K tripled)
  bool pred_tc = false;
  bf16 *wq3 = nullptr, *wkv3 = nullptr, *wo3 = nullptr, *wih3 = nullptr, *wada3 = nullptr;
  float* b_kv = nullptr;
  // prosody heads: shared BiLSTM + two small heads behind the length regulator
  bf16 *wih3_pros = nullptr, *wh1_3 = nullptr;
  float *whh_pros = nullptr, *lstm_b_pros = nullptr;
  struct ProsodyWs {        // sized by (B, F_max), grown on demand; separate from the sampler / predictor arena
    char* base = nullptr;
    size_t bytes = 0;
    int B = 0, F = 0, T = 0;
    float *frames, *G, *y;
    bf16* pa;
    int *flens, *perm, *dur;
    uint8_t* needed;
  } pws;
  const float* last_d_enc = nullptr;   // predict_duration_impl: the duration encoder's output (input of the final BiLSTM)
  Workspace ws;
  cudaStream_t stream = nullptr;      // internal stream (create-time work, host entry point, capture)
  // host entry point: prompt / noise H2D and the style D2H run on a second stream, overlapping the text-side
  // conditioning prep and the duration predictor
  cudaStream_t copy_stream = nullptr;   // H2D
  cudaStream_t out_stream = nullptr;    // D2H
  struct HostSlot {
    cudaEvent_t ev_text = nullptr, ev_prompt = nullptr, ev_noise = nullptr, ev_style = nullptr, ev_dur = nullptr, ev_done = nullptr;
    bool busy = false;
  } slot[2];
  cudaEvent_t cur_ev_prompt = nullptr, cur_ev_noise = nullptr;   // the running host call's pending H2D events
  bool wait_prompt = false, wait_noise = false;
  // noise == NULL: draw it on the device (philox.cuh) from this seed; utterance b of a call is global utterance noise_first_utt + b
  bool noise_seeded = false;
  uint64_t noise_seed = 0, noise_first_utt = 0;
  std::vector<unsigned long long> noise_ids;    // explicit global utterance indices (empty: first_utt + b)
  unsigned long long* noise_ids_dev = nullptr;
  size_t noise_ids_cap = 0;
  // calls share one workspace: a call enqueued on a different stream than the previous one first waits for it
  cudaStream_t last_stream = nullptr;
  cudaEvent_t last_ev = nullptr;
  bool has_last = false;
  // captured evaluation loops, keyed by (B, T bucket, P, evaluations, sampler kind + mask flags); least-recently-used
  // entries are evicted beyond max_graphs
  struct GraphEntry { cudaGraphExec_t exec; int launches; uint64_t last_use; int fuse_mode; };
  std::map<std::tuple<int, int, int, int, int>, GraphEntry> graphs;
  uint64_t graph_clock = 0;
  int64_t graph_captures = 0;   // captures since creation (stz_graph_count)
  int max_graphs = 32;
  int t_buckets = 1;            // round the text length up to a bucket (with masks) so that free-form T reuses graphs
  // own bounds checking (compute-sanitizer is closed on this pool): "guard_bytes" > 0 lays a poisoned gap after every buffer
  // of the workspace arenas; stz_debug_check_guards counts the gap bytes a kernel has overwritten
  int guard_bytes = 0;
  std::vector<std::pair<size_t, size_t>> guards, guards_pros;    // (offset, length) of the gaps
  int last_fuse_mode = 0, last_T = 0;   // what the last sample_style call dispatched (stz_get_option: tests assert the benched kernel ran)
  int use_graph = 1, gemm_impl = 0, lstm_impl = 0, lstm_nb = 0, pred_gemm_impl = 0, fuse_ln = 3;
  int chains = 1;      // independent utterance chains (parallel graph branches) of the evaluation loop
  cudaStream_t chain_stream[STZ_MAX_CHAINS] = {};
  cudaEvent_t fork_ev = nullptr, join_ev[STZ_MAX_CHAINS] = {};
  int attn_ctas = 0;   // resident-key tcgen05 attention: 4 = one unit in flight per CTA, four CTAs per SM (attention_tc4_kernel); 2 = attention_tc2_kernel; 0 = by unit count
  int attn_impl = 0;   // 0 = tcgen05 + TMA kernels (resident keys, streaming for long text; mma.sync streaming beyond their shapes), 2 = always the mma.sync streaming kernel
  int attn_box2 = 1;       // knob "attn_box2": one TMA box per attention operand (both branches) | one box per branch
  int gln_tile_rows = 0;   // knob "gln_tile_rows": 0 = heuristic (gemmln3_tile_rows), else forced rows per CTA pair of the fused kernel
  int use_pdl = 1, gemm_bn = 0, gemm_cluster = 0;   // launch knobs (copied into the thread-local launch context by every entry point)
  int ablate = 0;   // tools/ablate.py: bit mask of kernel families skipped inside run_eval (timing attribution only; results are wrong)
  int64_t launches = 0;
  int cur_launches = 0;  // launches issued since the counter was last sampled (capture bookkeeping)
  int tap_eval = -1, tap_layer = -1, tap_stage = -1;
  float* tap_buf = nullptr;
  bool capturing = false;
  // profile mode (bench.py roofline leg): CUDA-event pair around every launch, eager execution
  int profile = 0;
  struct ProfRec { int cls; cudaEvent_t a, b; double work; };
  std::vector<ProfRec> prof;
};

// Kernel classes of the profile buckets (include/stz.h: stz_profile_read).
enum ProfClass : int { PC_GEMM_TC = 0, PC_ATTN = 1, PC_LN = 2, PC_LINEAR_F32 = 3, PC_LSTM = 4, PC_PRED_EW = 5, PC_OTHER = 6, PC_COUNT = 7 };

struct ProfScope {
  stz_handle* H;
  cudaStream_t st;
  int idx = -1;
  ProfScope(stz_handle* H_, cudaStream_t st_, int cls, double work) : H(H_), st(st_) {
    if (!H || !H->profile || H->capturing) return;
    stz_handle::ProfRec r{cls, nullptr, nullptr, work};
    if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return;
    cudaEventRecord(r.a, st);
    H->prof.push_back(r);
    idx = (int)H->prof.size() - 1;
  }
  ~ProfScope() {
    if (idx >= 0) cudaEventRecord(H->prof[idx].b, st);
  }
};

static int fail(stz_handle* H, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (H) H->err = buf; else g_create_error = buf;
  return code;
}

#define CK(H, call)                                                                                   \
  do {                                                                                                \
    cudaError_t e_ = (call);                                                                          \
    if (e_ != cudaSuccess)                                                                            \
      return fail(H, STZ_E_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
  } while (0)

#define KCHECK(H)                                                                                     \
  do {                                                                                                \
    ++(H)->cur_launches;                                                                              \
    cudaError_t e_ = cudaGetLastError();                                                              \
    if (e_ != cudaSuccess)                                                                            \
      return fail(H, STZ_E_CUDA, "%s:%d kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e_)); \
  } while (0)

#define RET(expr)            \
  do {                       \
    int rc_ = (expr);        \
    if (rc_ != 0) return rc_; \
  } while (0)

// Launch-time knobs of the handle whose call is running on this host thread (set by every entry point that launches:
// LaunchScope).  Thread-local, so handles driven from different host threads never see each other's settings.
struct LaunchCtx { int use_pdl = 1, gemm_bn = 0, gemm_cluster = 0; };
static thread_local LaunchCtx tl_launch;

static void drop_graphs(stz_handle* H) {
  for (auto& g : H->graphs) cudaGraphExecDestroy(g.second.exec);
  H->graphs.clear();
}

// Text-length buckets of the captured evaluation loop: 32, 64, 96, 128, 192, 256, 384, 512, then multiples of 256.
// A call with T tokens runs at the bucket's length with the extra key positions masked (padding invariance is a tested
// property of the path), so a serving loop with free-form T captures at most ~10 graphs per (B, P, steps).
static int bucket_T(int T) {
  static const int b[] = {32, 64, 96, 128, 192, 256, 384, 512};
  for (int v : b) if (T <= v) return v;
  return (T + 255) / 256 * 256;
}
static bool runs_graphed(const stz_handle* H) { return H->use_graph && H->tap_buf == nullptr && !H->profile; }
static int effective_T(const stz_handle* H, int T) { return runs_graphed(H) && H->t_buckets ? bucket_T(T) : T; }

// Every entry point that launches kernels installs its handle's launch knobs for the calling thread.
struct LaunchScope {
  explicit LaunchScope(const stz_handle* H) {
    tl_launch = LaunchCtx();
    if (H) { tl_launch.use_pdl = H->use_pdl; tl_launch.gemm_bn = H->gemm_bn; tl_launch.gemm_cluster = H->gemm_cluster; }
  }
};

static int order_after_previous_call(stz_handle* H, cudaStream_t st) {
  if (H->has_last && st != H->last_stream) CK(H, cudaStreamWaitEvent(st, H->last_ev, 0));
  return 0;
}
static int mark_call_end(stz_handle* H, cudaStream_t st) {
  CK(H, cudaEventRecord(H->last_ev, st));
  H->last_stream = st;
  H->has_last = true;
  return 0;
}


// ------------------------------------------------------------------------------------------
// Kernel launch with programmatic dependent launch (PDL): the next kernel's CTAs are scheduled and run their
// prologue while the previous kernel drains; every kernel launched this way executes griddepcontrol.wait
// before touching global memory (ptx.cuh: pdl_wait), so completion stays transitive along the stream.
// ------------------------------------------------------------------------------------------
template <typename... KArgs, typename... Args>
static void launch_kcp(bool pdl, int cluster_x, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[2];
  int n = 0;
  if (pdl) {
    at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  if (cluster_x > 1) {   // runtime cluster shape (kernels without a compile-time __cluster_dims__)
    at[n].id = cudaLaunchAttributeClusterDimension;
    at[n].val.clusterDim.x = cluster_x; at[n].val.clusterDim.y = 1; at[n].val.clusterDim.z = 1;
    ++n;
  }
  cfg.attrs = at;
  cfg.numAttrs = n;
  cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
template <typename... KArgs, typename... Args>
static void launch_kc(int cluster_x, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  launch_kcp(tl_launch.use_pdl != 0, cluster_x, kern, grid, block, smem, st, static_cast<Args&&>(args)...);
}
template <typename... KArgs, typename... Args>
static void launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  launch_kc(1, kern, grid, block, smem, st, static_cast<Args&&>(args)...);
}
// Full (non-programmatic) dependency on everything before it in the stream, whatever the handle's PDL setting.
template <typename... KArgs, typename... Args>
static void launch_k_nopdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  launch_kcp(false, 1, kern, grid, block, smem, st, static_cast<Args&&>(args)...);
}

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }
static inline int ew_grid(size_t n, int per_block = 256) {
  size_t g = (n + per_block - 1) / per_block;
  return (int)(g < 1 ? 1 : (g > 148 * 16 ? 148 * 16 : g));
}

// ------------------------------------------------------------------------------------------
// TMA tensor maps (driver entry point fetched at run time: no link-time libcuda dependency)
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;

static int load_encode() {
  if (g_encode) return 0;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) return -1;
  g_encode = (EncodeTiledFn)fn;
  return 0;
}

// bf16 row-major [rows, cols] with row stride ld (elements); box = box_rows x 64 columns, 128B swizzle.
static int make_tmap(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstr[1] = {ld * sizeof(bf16)};
  cuuint32_t box[2] = {GEMM_BK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : (int)r;
}

// Output map of the staged epilogues: row-major [rows, cols], row stride ld (elements), box = 32 rows x row_bytes
// (128 B, 128B swizzle: gemm_ln kernels; 64 B, 64B swizzle: gemm2's double-buffered staging).  Rows >= `rows` are
// clipped by the TMA unit.
static int make_tmap_out(CUtensorMap* m, void* base, bool is_bf16, uint64_t rows, uint64_t cols, uint64_t ld, int row_bytes = 128,
                         int box_rows = 32) {
  const uint64_t es = is_bf16 ? 2 : 4;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstr[1] = {ld * es};
  cuuint32_t box[2] = {(cuuint32_t)(row_bytes / es), (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(m, is_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, gdim, gstr,
                        box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : (int)r;
}

// General tiled map over bf16 data with 128B swizzle: dims / box innermost first, strides (bytes) for dims 1..rank-1.
static int make_tmap_nd(CUtensorMap* m, const void* base, int rank, const cuuint64_t* gdim, const cuuint64_t* gstride_bytes,
                        const cuuint32_t* box) {
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstride_bytes, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : (int)r;
}

// ------------------------------------------------------------------------------------------
// GEMM launchers
// ------------------------------------------------------------------------------------------
constexpr int GEMM_BN = 128;   // N granularity of every GEMM

// ---- persistent, TMEM double-buffered, TMA-store epilogue (gemm2.cuh) ------------------------------
static int g_num_sms = 148;
static int g_lstm_max_clusters = 15;   // co-resident 8-CTA clusters of lstm_tc_kernel (queried at init)

static int pick_bn(int M, int N) {
  const int bn_override = tl_launch.gemm_bn;   // tuning knob ("gemm_bn"): 0 = heuristic
  if (bn_override && N % bn_override == 0) return bn_override;
  int best = 0;
  long best_cost = 0;
  for (int bn : {256, 192, 128}) {
    if (N % bn) continue;
    const long tiles = (long)(N / bn) * cdiv(M, GEMM_BM);
    // rounds of tiles x (columns per tile + a per-tile overhead worth ~64 columns: prologue / exposed epilogue; measured
    // on the context K/V GEMM, M 7296 x N 8192: BN 256 52 us vs BN 128 75 us although BN 128 needs fewer column-rounds)
    const long cost = (long)cdiv(tiles, g_num_sms) * (bn + 64);
    if (best == 0 || cost < best_cost) { best = bn; best_cost = cost; }
  }
  return best;
}


template <int BN, int EPI>
static int launch_gemm2_bn(stz_handle* H, cudaStream_t st, const bf16* A, int lda, int a_rows, const bf16* W, const GemmParams& p) {
  const int tiles_m = cdiv(p.M, GEMM_BM), tiles_n = p.N / BN;
  // CTA pairs (cta_group::2, 256 x BN tiles) relieve the shared-memory bandwidth bound of the single-CTA kernel;
  // used when there is at least one full round of pair tiles
#ifdef STZ_EXPERIMENTS   // knob "gemm_cluster": CTA pairs (cta_group::2, 256 x BN tiles); measured on par with single-CTA tiles, not in the product build
  const bool pair = tl_launch.gemm_cluster && g2_staged<EPI>() && tiles_m >= 2 && tiles_m * tiles_n > g_num_sms;
#else
  constexpr bool pair = false;
#endif
  CUtensorMap ta, tb, tc;
  memset(&tc, 0, sizeof tc);
  if (make_tmap(&ta, A, (uint64_t)a_rows, (uint64_t)p.K, (uint64_t)lda, GEMM_BM) ||
      make_tmap(&tb, W, (uint64_t)p.N, (uint64_t)p.K, (uint64_t)p.K, pair ? BN / 2 : BN))
    return fail(H, STZ_E_CUDA, "cuTensorMapEncodeTiled failed (M=%d N=%d K=%d)", p.M, p.N, p.K);
  if (g2_staged<EPI>() && make_tmap_out(&tc, p.out, g2_out_bf16<EPI>(), (uint64_t)p.M, (uint64_t)p.N, (uint64_t)p.ldo, 64))
    return fail(H, STZ_E_CUDA, "cuTensorMapEncodeTiled (output) failed (M=%d N=%d ldo=%d)", p.M, p.N, p.ldo);
  ProfScope ps(H, st, PC_GEMM_TC, 2.0 * p.M * p.N * p.K);
#ifdef STZ_EXPERIMENTS
  if (pair) {
    const int units = cdiv(tiles_m, 2) * tiles_n, max_clusters = g_num_sms / 2;
    const int clusters = units < max_clusters ? units : max_clusters;
    launch_kc(2, gemm2_kernel<BN, EPI, 2>, 2 * clusters, G2_THREADS, g2_smem_bytes_cm<BN, 2>(), st, ta, tb, tc, p);
  } else
#endif
  {
    const int tiles = tiles_m * tiles_n;
    const int grid = tiles < g_num_sms ? tiles : g_num_sms;
    launch_k(gemm2_kernel<BN, EPI, 1>, grid, G2_THREADS, g2_smem_bytes<BN>(), st, ta, tb, tc, p);
  }
  if (H) { KCHECK(H); } else if (cudaGetLastError() != cudaSuccess) return STZ_E_CUDA;
  return 0;
}

template <int EPI>
static int launch_gemm2(stz_handle* H, cudaStream_t st, const bf16* A, int lda, int a_rows, const bf16* W, const GemmParams& p) {
  switch (pick_bn(p.M, p.N)) {
    case 256: return launch_gemm2_bn<256, EPI>(H, st, A, lda, a_rows, W, p);
    case 192: return launch_gemm2_bn<192, EPI>(H, st, A, lda, a_rows, W, p);
    case 128: return launch_gemm2_bn<128, EPI>(H, st, A, lda, a_rows, W, p);
  }
  return fail(H, STZ_E_SHAPE, "gemm N=%d is not a multiple of 128", p.N);
}

// ---- fused GEMM + residual/pos + AdaLN (gemm_ln3.cuh): N = d_model = 512, residual tile staged in the operand ring ----------
// Rows per CTA pair of gemmln3_kernel: the smallest multiple of 8 (the swizzle atom) with which the row blocks still fill
// ONE wave of pairs — cfg2's 6400 rows: 88-row blocks on 73 pairs = 146 SMs instead of 128-row blocks on 50 pairs = 100 SMs.
// The MMAs keep M = 128; what shrinks with the rows is the shared-memory-bound epilogue (whole 32-row warps drop out below
// 97 / 65 / 33 rows) and the TMA traffic of the residual / operand tiles.  Measured (tools/ab_tile_rows2.py,
// profiles/r02_ab_tile_rows.txt): B = 64: 88 = 96 rows, -2 % vs 128; B = 72: 104 rows -1.7 % vs 128; B = 48: 72 rows -0.6 % vs 96,
// -4.7 % vs 128.  More row blocks than pairs (two waves and more): 128 rows.
static int gemmln3_tile_rows(int M) {
  const int pairs = g_num_sms / 2;
  if (cdiv(M, GEMM_BM) > pairs) return GEMM_BM;
  const int tr = (cdiv(M, pairs) + 7) / 8 * 8;
  return tr < 8 ? 8 : (tr > GEMM_BM ? GEMM_BM : tr);
}

template <int MODE>
static int launch_gemmln3(stz_handle* H, cudaStream_t st, const bf16* A, int lda, int a_rows, const bf16* W, bf16* u,
                          const GemmLnParams& p_in) {
  GemmLnParams p = p_in;
  if (p.tile_rows == 0) p.tile_rows = H->gln_tile_rows > 0 ? H->gln_tile_rows : gemmln3_tile_rows(p.M);
  if (p.tile_rows < 8 || p.tile_rows > GEMM_BM || p.tile_rows % 8) return fail(H, STZ_E_ARG, "gemmln3 tile_rows %d", p.tile_rows);
  if (p.K % (GEMM_BK * GLN3_STAGES)) return fail(H, STZ_E_SHAPE, "gemmln3 needs K %% %d == 0", GEMM_BK * GLN3_STAGES);
  if ((p.single ? 1 : 2) * ((GEMM_BM - 1 + p.rows_per_utt - 1) / p.rows_per_utt + 1) > GLN3_MAX_SEQ)
    return fail(H, STZ_E_SHAPE, "gemmln3: a 128-row tile spans too many sequences (rows_per_utt %d)", p.rows_per_utt);
  CUtensorMap ta, tb, tu, th;
  const int TR = p.tile_rows;
  if (make_tmap(&ta, A, (uint64_t)a_rows, (uint64_t)p.K, (uint64_t)lda, TR) ||
      make_tmap(&tb, W, (uint64_t)GLN_N, (uint64_t)p.K, (uint64_t)p.K, GLN3_BN) ||
      make_tmap_out(&th, p.h, false, (uint64_t)p.M, (uint64_t)GLN_N, (uint64_t)GLN_N, 128, TR) ||
      make_tmap_out(&tu, u, true, (uint64_t)p.M, (uint64_t)(p.split3 ? 3 * GLN_N : GLN_N), (uint64_t)(p.split3 ? 3 * GLN_N : GLN_N), 128, TR))
    return fail(H, STZ_E_CUDA, "cuTensorMapEncodeTiled failed (gemmln3 M=%d K=%d)", p.M, p.K);
  ProfScope ps(H, st, PC_GEMM_TC, 2.0 * p.M * GLN_N * p.K);
  launch_kc(2, gemmln3_kernel<MODE>, 2 * cdiv(p.M, TR), GLN_THREADS, GLN3_SMEM_BYTES, st, ta, tb, tu, th, p);
  KCHECK(H);
  return 0;
}

template <int BN, int EPI>
static cudaError_t set_gemm2_attr() {
  cudaError_t e = cudaFuncSetAttribute(gemm2_kernel<BN, EPI, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, g2_smem_bytes<BN>());
#ifdef STZ_EXPERIMENTS
  if (e != cudaSuccess || !g2_staged<EPI>()) return e;
  return cudaFuncSetAttribute(gemm2_kernel<BN, EPI, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, g2_smem_bytes_cm<BN, 2>());
#else
  return e;
#endif
}
template <int EPI>
static cudaError_t set_gemm2_attrs() {
  cudaError_t e;
  if ((e = set_gemm2_attr<256, EPI>()) != cudaSuccess) return e;
  if ((e = set_gemm2_attr<192, EPI>()) != cudaSuccess) return e;
  return set_gemm2_attr<128, EPI>();
}

static constexpr int dur_head2_smem(int vpl) { return (32 * 128 * vpl + 8 * 32) * 4; }

// Per-device opt-in to > 48 KB dynamic shared memory; done once per handle, never during graph capture.
static cudaError_t init_kernel_attrs() {
  cudaError_t e;
  if ((e = set_gemm2_attrs<EPI_F32>()) != cudaSuccess) return e;
  if ((e = set_gemm2_attrs<EPI_F32_POS>()) != cudaSuccess) return e;
  if ((e = set_gemm2_attrs<EPI_BF16>()) != cudaSuccess) return e;
  if ((e = set_gemm2_attrs<EPI_GELU_BF16>()) != cudaSuccess) return e;
  if ((e = set_gemm2_attrs<EPI_GATE_RES>()) != cudaSuccess) return e;
  if ((e = set_gemm2_attrs<EPI_SAMPLER>()) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(gemmln3_kernel<GLN_RES>, cudaFuncAttributeMaxDynamicSharedMemorySize, GLN3_SMEM_BYTES)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(gemmln3_kernel<GLN_POS>, cudaFuncAttributeMaxDynamicSharedMemorySize, GLN3_SMEM_BYTES)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM_BYTES)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(attention_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATC_SMEM_BYTES)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(attention_tc4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATC4_SMEM_BYTES)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(attention_tcs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATS_SMEM_BYTES)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(lstm_tc_kernel<false, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, lt_smem_bytes(false, 16))) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(lstm_tc_kernel<true, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, lt_smem_bytes(true, 16))) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(lstm_tc_kernel<true, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, lt_smem_bytes(true, 8))) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(lstm_tc_kernel<true, 24>, cudaFuncAttributeMaxDynamicSharedMemorySize, lt_smem_bytes(true, 24))) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(style_pool_attn2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (64 * (2 * 256 + 1) + 4) * 4)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(dur_head2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, dur_head2_smem(1))) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(dur_head2_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, dur_head2_smem(2))) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(dur_head2_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, dur_head2_smem(4))) != cudaSuccess) return e;
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && sms > 0)
    g_num_sms = sms;
  {  // co-resident 8-CTA clusters of the BiLSTM recurrence (its wave model: lstm_pick_nb)
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(LC_CS * 64, 2); cfg.blockDim = dim3(lt_threads(16)); cfg.dynamicSmemBytes = lt_smem_bytes(true, 16);
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = LC_CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, lstm_tc_kernel<true, 16>, &cfg) == cudaSuccess && n > 0) g_lstm_max_clusters = n;
    else (void)cudaGetLastError();
  }
  return cudaSuccess;
}

template <int EPI>
static int launch_gemm_simt(stz_handle* H, cudaStream_t st, const bf16* A, int lda, const bf16* W, const GemmParams& p) {
  dim3 grid(p.N / 32, cdiv(p.M, 128));
  gemm_simt_kernel<EPI><<<grid, 128, 0, st>>>(A, lda, W, p);
  if (H) { KCHECK(H); } else if (cudaGetLastError() != cudaSuccess) return STZ_E_CUDA;
  return 0;
}

template <int EPI>
static int gemm(stz_handle* H, cudaStream_t st, int impl, const bf16* A, int lda, int a_rows, const bf16* W,
                const GemmParams& p) {
  if (p.N % GEMM_BN != 0 || p.K % GEMM_BK != 0 || p.M <= 0)
    return fail(H, STZ_E_SHAPE, "gemm shape M=%d N=%d K=%d unsupported (N %% 128, K %% 64)", p.M, p.N, p.K);
  if (impl == 0) return launch_gemm2<EPI>(H, st, A, lda, a_rows, W, p);
  return launch_gemm_simt<EPI>(H, st, A, lda, W, p);
}

// fused residual GEMM + AdaLN (gemm_ln3.cuh)
template <int MODE>
static int launch_gemmln(stz_handle* H, cudaStream_t st, const bf16* A, int lda, int a_rows, const bf16* W, bf16* u,
                         const GemmLnParams& p) {
  return launch_gemmln3<MODE>(H, st, A, lda, a_rows, W, u, p);
}

static int linear_f32(stz_handle* H, cudaStream_t st, int act, const float* X1, int ld1, int K1, const float* X2, int ld2,
                      int K2, const float* W, const float* b, float* Y, int ldy, int M, int N) {
  ProfScope ps(H, st, PC_LINEAR_F32, 2.0 * M * N * (K1 + K2));
  if (K2 == 0 && M <= 512 && K1 % 128 == 0 && N % 8 == 0 && ld1 % 4 == 0) {   // tiny-M: warp-per-output kernel
    const int warps = M * (N / 8);
    if (act == ACT_SILU) launch_k(small_linear_kernel<ACT_SILU>, cdiv(warps, 8), 256, 0, st, X1, ld1, W, b, Y, ldy, M, N, K1);
    else launch_k(small_linear_kernel<ACT_NONE>, cdiv(warps, 8), 256, 0, st, X1, ld1, W, b, Y, ldy, M, N, K1);
    KCHECK(H);
    return 0;
  }
  dim3 grid(cdiv(N, 64), cdiv(M, 64));
  if (act == ACT_SILU)
    launch_k(linear_f32_kernel<ACT_SILU>, grid, 256, 0, st, X1, ld1, K1, X2, ld2, K2, W, b, Y, ldy, M, N);
  else
    launch_k(linear_f32_kernel<ACT_NONE>, grid, 256, 0, st, X1, ld1, K1, X2, ld2, K2, W, b, Y, ldy, M, N);
  KCHECK(H);
  return 0;
}

static int ln_mod(stz_handle* H, cudaStream_t st, const float* h, int rows, int D, const float* mod, int n_mod,
                  int shift_off, int scale_off, int rows_per_utt, bf16* out, int split3 = 0, int single = 0) {
  dim3 grid(cdiv(rows, 8));
  ProfScope ps(H, st, PC_LN, (double)rows * D * 6.0);  // fp32 in + bf16 out
  switch (D / 128) {
    case 1: launch_k(ln_mod_kernel<1>, grid, 256, 0, st, h, rows, mod, n_mod, shift_off, scale_off, rows_per_utt, out, split3, single); break;
    case 2: launch_k(ln_mod_kernel<2>, grid, 256, 0, st, h, rows, mod, n_mod, shift_off, scale_off, rows_per_utt, out, split3, single); break;
    case 4: launch_k(ln_mod_kernel<4>, grid, 256, 0, st, h, rows, mod, n_mod, shift_off, scale_off, rows_per_utt, out, split3, single); break;
    case 8: launch_k(ln_mod_kernel<8>, grid, 256, 0, st, h, rows, mod, n_mod, shift_off, scale_off, rows_per_utt, out, split3, single); break;
    default: return fail(H, STZ_E_SHAPE, "d_model %d unsupported by ln_mod", D);
  }
  KCHECK(H);
  return 0;
}

// ------------------------------------------------------------------------------------------
// helpers over the handle
// ------------------------------------------------------------------------------------------
static const float* W32(const stz_handle* H, const std::string& name) { return H->w32 + H->off.at(name); }
static const bf16* WBF(const stz_handle* H, const std::string& name) { return H->wbf + H->off.at(name); }

static int check_config(const stz_config& c) {
  if (c.n_heads <= 0 || c.d_model != c.n_heads * 64) return -1;            // d_head == 64 (attention.cuh)
  if (c.d_model % 128 || c.d_ff % 128 || c.d_style % 128) return -1;       // N tiles of 128
  if (c.d_text % 128 || c.d_prompt % 128 || c.d_model % 64 || c.d_ff % 64 || c.d_style % 64) return -1;  // K tiles of 64; cast_pool: 128 columns per CTA
  if (c.d_model > 1024 || c.d_hid > 1024 || c.d_hid % 128) return -1;      // warp-per-row kernels
  if (c.n_style < 1 || c.n_style > 64) return -1;                          // 2K query rows <= 128
  if (c.d_hid != c.d_text) return -1;                                      // x0 = text_emb
  if (c.n_sp_heads <= 0 || c.d_sty_tok != c.n_sp_heads * 32) return -1;    // lane = channel
  if (c.d_hid / 2 * 4 > 1024 || (c.d_hid / 2) % 4) return -1;              // lstm block = 4h threads
  if (c.n_lstm < 1 || c.n_layers < 1 || c.max_dur < 1 || c.d_time < 2 || c.d_time % 2) return -1;
  return 0;
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static int ensure_workspace(stz_handle* H, int B, int T, int P, int E, int noise_slices) {
  Workspace& w = H->ws;
  const size_t mod_rows_req = (size_t)(hoist_mod(H->cfg, B, E) ? E : 1) * 2 * B;
  if (w.base && B <= w.B && T <= w.T && P <= w.P && E <= w.E && noise_slices <= w.noise_slices && mod_rows_req <= w.mod_rows) return 0;
  // grow monotonically; all cached graphs point into the old arena (stz_reserve sizes it once, up front)
  drop_graphs(H);
  if (w.base) { CK(H, cudaDeviceSynchronize()); CK(H, cudaFree(w.base)); w.base = nullptr; }
  B = B > w.B ? B : w.B; T = T > w.T ? T : w.T; P = P > w.P ? P : w.P; E = E > w.E ? E : w.E;
  noise_slices = noise_slices > w.noise_slices ? noise_slices : w.noise_slices;
  const stz_config& c = H->cfg;
  const size_t d = c.d_model, Ds = c.d_style, L = c.n_layers, K = c.n_style, n_mod = (9 * L + 2) * d;
  const size_t BT = (size_t)B * T, BP = (size_t)B * P, BK = (size_t)B * K, R = 2 * BK, NS = 2 * (size_t)B;
  const size_t ds = c.d_sty_tok, dh = c.d_hid, h8 = 4 * dh;  // 8h = 4 * d_hid
  size_t off = 0;
  std::vector<std::pair<void**, size_t>> plan;
  H->guards.clear();
  auto want = [&](void** p, size_t bytes) {
    plan.push_back({p, off});
    const size_t end = off + bytes;
    off = align_up(end + (size_t)H->guard_bytes, 1024);
    if (H->guard_bytes > 0) H->guards.push_back({end, off - end});
  };
#define WANT(field, count, type) want((void**)&w.field, (size_t)(count) * sizeof(type))
  WANT(text_bf, BT * c.d_text, bf16); WANT(prompt_bf, BP * c.d_prompt, bf16);
  // context tokens [text rows ; prompt rows] and their per-layer K/V are contiguous (one LN, one K/V GEMM per call);
  // ctx_prompt / kv_prompt are derived per call from the actual B*T (sample_style_impl)
  WANT(ctx_text, (BT + BP + 128) * d, bf16);
  WANT(kv_text, (BT + BP) * L * 2 * d, bf16);
  WANT(cvec, (size_t)E * NS * d + 128 * d, bf16);  // + one tile of slack rows for the last eval's TMA box
  WANT(pool_text, (size_t)B * c.d_text, float); WANT(pool_prompt, (size_t)B * c.d_prompt, float);
  WANT(pt, (size_t)B * d, float); WANT(pp, (size_t)B * d, float);
  WANT(ctx_pre, (BT + BP) * d, float);
  WANT(tfeat, (size_t)E * c.d_time, float); WANT(t1, (size_t)E * d, float); WANT(temb, (size_t)E * d, float);
  WANT(coef, (size_t)E * 8, float);
  WANT(gfeat, c.d_time, float); WANT(g1, d, float); WANT(gemb, d, float);
  const size_t mod_rows = mod_rows_req > w.mod_rows ? (mod_rows_req > NS ? mod_rows_req : NS) : (w.mod_rows > NS ? w.mod_rows : NS);
  WANT(mod, mod_rows * n_mod, float);   // few-step samplers: every evaluation's modulations at once
  WANT(x, BK * Ds, float); WANT(xmid, BK * Ds, float); WANT(h, R * d, float);
  WANT(noise, (size_t)noise_slices * BK * Ds, float);
  WANT(xin, R * 3 * Ds, bf16); WANT(u, R * d, bf16); WANT(u3, R * 3 * d, bf16); WANT(qkv, R * 3 * d, bf16); WANT(att, R * d, bf16);
  WANT(ffh, R * c.d_ff, bf16);
  WANT(sq, BT * ds, float); WANT(sk, BK * ds, float); WANT(sv, BK * ds, float); WANT(sa, BT * ds, float);
  WANT(stok, BT * ds, float); WANT(G, BT * h8, float); WANT(xa, BT * dh, float); WANT(xb, BT * dh, float);
  WANT(gb, BT * 2 * dh, float); WANT(lens, B, int); WANT(perm, B, int); WANT(tile_needed, BT / 128 + 2, uint8_t); WANT(tile_needed_ctx, (BT + BP) / 128 + 4, uint8_t);
  WANT(skv, BK * 2 * ds, float);
  WANT(pa, (BT + 128) * 3 * (dh + ds), bf16); WANT(ps3, (BT + 128) * 3 * ds, bf16); WANT(pq, (BT + 128) * 3 * c.d_text, bf16);
  WANT(pstyle3, (BK + 128) * 3 * Ds, bf16); WANT(psa3, (BT + 128) * 3 * ds, bf16);
  WANT(st_tmask, BT, uint8_t); WANT(st_pmask, BP, uint8_t);
  for (int sl = 0; sl < 2; ++sl) {
    WANT(hs[sl].text, BT * c.d_text, float); WANT(hs[sl].prompt, BP * c.d_prompt, float);
    WANT(hs[sl].noise, (size_t)(noise_slices > 1 ? noise_slices : 1) * BK * Ds, float); WANT(hs[sl].style, BK * Ds, float);
    WANT(hs[sl].tmask, BT, uint8_t); WANT(hs[sl].pmask, BP, uint8_t); WANT(hs[sl].dur, BT, int32_t);
  }
#undef WANT
  cudaError_t e = cudaMalloc(&w.base, off);
  if (e != cudaSuccess) {
    w = Workspace();
    return fail(H, STZ_E_NOMEM, "workspace of %zu bytes: %s", off, cudaGetErrorString(e));
  }
  for (auto& pr : plan) *pr.first = w.base + pr.second;
  if (H->guard_bytes > 0) CK(H, cudaMemsetAsync(w.base, STZ_GUARD_BYTE, off, H->stream));
  w.bytes = off; w.B = B; w.T = T; w.P = P; w.E = E; w.noise_slices = noise_slices; w.mod_rows = mod_rows;
  CK(H, cudaMemsetAsync(w.cvec, 0, ((size_t)E * NS * d + 128 * d) * sizeof(bf16), H->stream));
  CK(H, cudaStreamSynchronize(H->stream));
  return 0;
}

// ------------------------------------------------------------------------------------------
// sampler schedule (a-1, a-2, a-6) — fp64 on the host, uploaded as fp32 tables
// ------------------------------------------------------------------------------------------
struct EvalPlan {
  std::vector<double> sigma;   // sigma fed to the denoiser at eval e
  std::vector<float> coef;     // [E][8]
  std::vector<float> tfeat;    // [E][d_time]
  std::vector<float> gfeat;    // [d_time] features of the guidance scale (guidance-conditioned student)
  double sigma0 = 0, cin0 = 0;
};

static std::vector<double> karras(int n, double smin, double smax, double rho) {
  std::vector<double> s(n);
  if (n == 1) { s[0] = smax; return s; }
  const double a = pow(smax, 1.0 / rho), b = pow(smin, 1.0 / rho);
  for (int i = 0; i < n; ++i) s[i] = pow(a + (double)i / (n - 1) * (b - a), rho);
  return s;
}

static void precond(double sigma, double sd, double* cskip, double* cout_, double* cin) {
  const double s2 = sigma * sigma, d2 = sd * sd;
  *cskip = d2 / (s2 + d2);
  *cout_ = sigma * sd / sqrt(s2 + d2);
  *cin = 1.0 / sqrt(s2 + d2);
}

static EvalPlan make_plan(const stz_config& c, int steps, int kind, float cfg_scale) {
  EvalPlan pl;
  const double sd = c.sigma_data;
  auto push = [&](double sigma, double cx, double cm, double cF, double cn, double sigma_next_in, int dest) {
    double cin_next = 0.0;
    if (sigma_next_in > 0.0) { double a, b; precond(sigma_next_in, sd, &a, &b, &cin_next); }
    pl.sigma.push_back(sigma);
    const float row[8] = {(float)cx, (float)cm, (float)cF, (float)cn, (float)cin_next, cfg_scale, (float)dest, 0.f};
    pl.coef.insert(pl.coef.end(), row, row + 8);
  };
  if (kind == STZ_SAMPLER_STUDENT || kind == STZ_SAMPLER_GUIDED) {   // the same Euler steps; guided: F is the single-branch network output
    std::vector<double> s = karras(steps, c.sigma_min, c.sigma_max, c.rho);
    s.push_back(0.0);
    for (int i = 0; i < steps; ++i) {
      double cskip, cout_, cin;
      precond(s[i], sd, &cskip, &cout_, &cin);
      const double r = (s[i + 1] - s[i]) / s[i];
      // x' = x + (x - D) r,  D = cskip x + cout F   ->   x' = (1 + (1 - cskip) r) x - cout r F
      push(s[i], 1.0 + (1.0 - cskip) * r, 0.0, -cout_ * r, 0.0, s[i + 1], 0);
    }
  } else {
    std::vector<double> s = karras(steps + 1, c.sigma_min, c.sigma_max, c.rho);
    for (int i = 0; i < steps; ++i) {
      const double sg = s[i], sn = s[i + 1];
      const double sup = sqrt(sn * sn * (sg * sg - sn * sn) / (sg * sg));
      const double sdown = sqrt(fmax(sn * sn - sup * sup, 0.0));
      const double smid = 0.5 * (sg + sdown);
      double cskip, cout_, cin;
      precond(sg, sd, &cskip, &cout_, &cin);
      const double r1 = (smid - sg) / sg;
      // x_mid = x + (x - D(x)) r1
      push(sg, 1.0 + (1.0 - cskip) * r1, 0.0, -cout_ * r1, 0.0, smid, 1);
      precond(smid, sd, &cskip, &cout_, &cin);
      const double r2 = (sdown - sg) / smid;
      // x' = x + (x_mid - D(x_mid)) r2 + sigma_up * noise_{i+1}
      push(smid, 1.0, (1.0 - cskip) * r2, -cout_ * r2, sup, sn, 0);
    }
  }
  pl.sigma0 = pl.sigma[0];
  { double a, b; precond(pl.sigma0, sd, &a, &b, &pl.cin0); }
  const int half = c.d_time / 2;
  {  // sinusoidal features of omega / 4 (same frequencies as the time features)
    const double cg = (double)cfg_scale / 4.0;
    for (int i = 0; i < half; ++i) pl.gfeat.push_back((float)sin(cg * exp(log(100.0) * i / (half > 1 ? half - 1 : 1))));
    for (int i = 0; i < half; ++i) pl.gfeat.push_back((float)cos(cg * exp(log(100.0) * i / (half > 1 ? half - 1 : 1))));
  }
  for (double sg : pl.sigma) {
    const double cn = log(sg) / 4.0;
    for (int i = 0; i < half; ++i) pl.tfeat.push_back((float)sin(cn * exp(log(100.0) * i / (half > 1 ? half - 1 : 1))));
    for (int i = 0; i < half; ++i) pl.tfeat.push_back((float)cos(cn * exp(log(100.0) * i / (half > 1 ? half - 1 : 1))));
  }
  return pl;
}

// Host-only: the schedule / coefficient tables of one call (no device needed) — `-m "not gpu"` tests check them against
// oracle/schedule.py.  Returns the number of denoiser evaluations E; any output pointer may be NULL.
//   sigma_out [E] fp64, coef_out [E][8] fp32 = (c_x, c_mid, c_F, c_noise, c_in(next), cfg_scale, dest, 0),
//   tfeat_out [E][d_time] fp32, init_out [2] fp64 = (sigma_0, c_in(sigma_0)).
extern "C" int stz_debug_plan(const stz_config* cfg, int steps, int sampler_kind, float cfg_scale, double* sigma_out,
                              float* coef_out, float* tfeat_out, double* init_out) {
  if (!cfg || steps < 1 || steps > 1024 || sampler_kind < STZ_SAMPLER_STUDENT || sampler_kind > STZ_SAMPLER_GUIDED) return STZ_E_ARG;
  const EvalPlan pl = make_plan(*cfg, steps, sampler_kind, cfg_scale);
  if (sigma_out) std::copy(pl.sigma.begin(), pl.sigma.end(), sigma_out);
  if (coef_out) std::copy(pl.coef.begin(), pl.coef.end(), coef_out);
  if (tfeat_out) std::copy(pl.tfeat.begin(), pl.tfeat.end(), tfeat_out);
  if (init_out) { init_out[0] = pl.sigma0; init_out[1] = pl.cin0; }
  return (int)pl.sigma.size();
}

// ------------------------------------------------------------------------------------------
// create / destroy
// ------------------------------------------------------------------------------------------
extern "C" int stz_abi_version(void) { return STZ_ABI_VERSION; }

extern "C" size_t stz_weights_nfloats(const stz_config* cfg) {
  if (!cfg) return 0;
  size_t total = 0;
  build_layout(*cfg, &total);
  return total;
}

extern "C" int64_t stz_weight_offset(const stz_config* cfg, const char* name) {
  if (!cfg || !name) return -1;
  size_t total = 0;
  for (const WeightEntry& e : build_layout(*cfg, &total))
    if (e.name == name) return (int64_t)e.off;
  return -1;
}

extern "C" const char* stz_last_error(const stz_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

extern "C" void stz_destroy(stz_handle* H) {
  if (!H) return;
  cudaSetDevice(H->device);
  cudaDeviceSynchronize();
  drop_graphs(H);
  cudaFree(H->ws.base); cudaFree(H->w32); cudaFree(H->wbf); cudaFree(H->ctx_text_b); cudaFree(H->ctx_prompt_b);
  cudaFree(H->wq3); cudaFree(H->wkv3); cudaFree(H->wo3); cudaFree(H->wih3); cudaFree(H->wada3); cudaFree(H->b_kv);
  cudaFree(H->wkv_all); cudaFree(H->bkv_all);
  cudaFree(H->wih3_pros); cudaFree(H->wh1_3); cudaFree(H->whh_pros); cudaFree(H->lstm_b_pros); cudaFree(H->pws.base);
  for (auto& sl : H->slot)
    for (cudaEvent_t e : {sl.ev_text, sl.ev_prompt, sl.ev_noise, sl.ev_style, sl.ev_dur, sl.ev_done})
      if (e) cudaEventDestroy(e);
  if (H->out_stream) cudaStreamDestroy(H->out_stream);
  if (H->fork_ev) cudaEventDestroy(H->fork_ev);
  for (int i = 1; i < STZ_MAX_CHAINS; ++i) {
    if (H->join_ev[i]) cudaEventDestroy(H->join_ev[i]);
    if (H->chain_stream[i]) cudaStreamDestroy(H->chain_stream[i]);
  }
  if (H->copy_stream) cudaStreamDestroy(H->copy_stream);
  cudaFree(H->noise_ids_dev);
  cudaFree(H->kv_null); cudaFree(H->w_in3); cudaFree(H->w_out3); cudaFree(H->whhT); cudaFree(H->whh); cudaFree(H->lstm_b);
  if (H->last_ev) cudaEventDestroy(H->last_ev);
  if (H->stream) cudaStreamDestroy(H->stream);
  delete H;
}

__global__ void transpose_whh_kernel(const float* __restrict__ w, float* __restrict__ wT, int rows, int cols) {
  // w [rows = 4h, cols = h]  ->  wT [cols, rows]
  const size_t n = (size_t)rows * cols;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / cols), c = (int)(i % cols);
    wT[(size_t)c * rows + r] = w[i];
  }
}

static int create_impl(stz_handle* H, const float* weights_host) {
  const stz_config& c = H->cfg;
  const int d = c.d_model, L = c.n_layers, h = c.d_hid / 2;
  CK(H, cudaStreamCreateWithFlags(&H->stream, cudaStreamNonBlocking));
  CK(H, cudaEventCreateWithFlags(&H->last_ev, cudaEventDisableTiming));
  CK(H, cudaStreamCreateWithFlags(&H->copy_stream, cudaStreamNonBlocking));
  CK(H, cudaStreamCreateWithFlags(&H->out_stream, cudaStreamNonBlocking));
  for (auto& sl : H->slot)
    for (cudaEvent_t* e : {&sl.ev_text, &sl.ev_prompt, &sl.ev_noise, &sl.ev_style, &sl.ev_dur, &sl.ev_done})
      CK(H, cudaEventCreateWithFlags(e, cudaEventDisableTiming));
  CK(H, cudaEventCreateWithFlags(&H->fork_ev, cudaEventDisableTiming));
  for (int i = 1; i < STZ_MAX_CHAINS; ++i) {
    CK(H, cudaStreamCreateWithFlags(&H->chain_stream[i], cudaStreamNonBlocking));
    CK(H, cudaEventCreateWithFlags(&H->join_ev[i], cudaEventDisableTiming));
  }
  CK(H, init_kernel_attrs());
  cudaStream_t st = H->stream;
  CK(H, cudaMalloc(&H->w32, H->n_floats * sizeof(float)));
  CK(H, cudaMalloc(&H->wbf, H->n_floats * sizeof(bf16)));
  CK(H, cudaMemcpyAsync(H->w32, weights_host, H->n_floats * sizeof(float), cudaMemcpyHostToDevice, st));
  f32_to_bf16_kernel<<<ew_grid(H->n_floats), 256, 0, st>>>(H->w32, H->wbf, H->n_floats);
  KCHECK(H);
  // fp32-grade input / output projections: split-bf16 weights (see split3_weights_kernel)
  CK(H, cudaMalloc(&H->w_in3, (size_t)d * 3 * c.d_style * sizeof(bf16)));
  CK(H, cudaMalloc(&H->w_out3, (size_t)c.d_style * 3 * d * sizeof(bf16)));
  split3_weights_kernel<<<ew_grid((size_t)d * c.d_style), 256, 0, st>>>(W32(H, "in.w"), H->w_in3, d, c.d_style); KCHECK(H);
  split3_weights_kernel<<<ew_grid((size_t)d * c.d_style), 256, 0, st>>>(W32(H, "out.w"), H->w_out3, c.d_style, d); KCHECK(H);
  // ctx biases with the token-type embedding folded in
  CK(H, cudaMalloc(&H->ctx_text_b, d * sizeof(float)));
  CK(H, cudaMalloc(&H->ctx_prompt_b, d * sizeof(float)));
  add_vec_kernel<<<1, 256, 0, st>>>(W32(H, "ctx_text.b"), W32(H, "type_emb"), H->ctx_text_b, d, d); KCHECK(H);
  add_vec_kernel<<<1, 256, 0, st>>>(W32(H, "ctx_prompt.b"), W32(H, "type_emb") + d, H->ctx_prompt_b, d, d); KCHECK(H);
  // null-prompt context token: LN(null_tok + e1) -> per-layer K/V
  float* tmp = nullptr; bf16* nullc = nullptr;
  CK(H, cudaMalloc(&tmp, d * sizeof(float)));
  CK(H, cudaMalloc(&nullc, 128 * d * sizeof(bf16)));  // padded to one M tile
  CK(H, cudaMemsetAsync(nullc, 0, 128 * d * sizeof(bf16), st));
  CK(H, cudaMalloc(&H->kv_null, (size_t)L * 2 * d * sizeof(bf16)));
  add_vec_kernel<<<1, 256, 0, st>>>(W32(H, "null_tok"), W32(H, "type_emb") + d, tmp, d, d); KCHECK(H);
  RET(ln_mod(H, st, tmp, 1, d, nullptr, 0, 0, 0, 1, nullc));
  CK(H, cudaMalloc(&H->wkv_all, (size_t)L * 2 * d * d * sizeof(bf16)));
  CK(H, cudaMalloc(&H->bkv_all, (size_t)L * 2 * d * sizeof(float)));
  for (int l = 0; l < L; ++l) {
    const std::string p = "l" + std::to_string(l) + ".kv2.";
    CK(H, cudaMemcpyAsync(H->wkv_all + (size_t)l * 2 * d * d, WBF(H, p + "w"), (size_t)2 * d * d * sizeof(bf16), cudaMemcpyDeviceToDevice, st));
    CK(H, cudaMemcpyAsync(H->bkv_all + (size_t)l * 2 * d, W32(H, p + "b"), (size_t)2 * d * sizeof(float), cudaMemcpyDeviceToDevice, st));
    GemmParams gp{};
    gp.M = 1; gp.N = 2 * d; gp.K = d; gp.bias = W32(H, p + "b"); gp.out = H->kv_null + (size_t)l * 2 * d; gp.ldo = L * 2 * d;
    RET(gemm<EPI_BF16>(H, st, 1, nullc, d, 128, WBF(H, p + "w"), gp));  // M = 1: CUDA-core kernel, create time only
  }
  // predictor: recurrent weights k-major, biases summed
  CK(H, cudaMalloc(&H->whhT, (size_t)c.n_lstm * 2 * h * 4 * h * sizeof(float)));
  CK(H, cudaMalloc(&H->lstm_b, (size_t)c.n_lstm * 2 * 4 * h * sizeof(float)));
  CK(H, cudaMalloc(&H->whh, (size_t)c.n_lstm * 2 * 4 * h * h * sizeof(float)));
  for (int l = 0; l < c.n_lstm; ++l)
    for (int dr = 0; dr < 2; ++dr) {
      const std::string p = "lstm" + std::to_string(l) + (dr ? ".r." : ".f.");
      CK(H, cudaMemcpyAsync(H->whh + ((size_t)l * 2 + dr) * 4 * h * h, W32(H, p + "w_hh"), (size_t)4 * h * h * sizeof(float),
                            cudaMemcpyDeviceToDevice, st));
      transpose_whh_kernel<<<ew_grid((size_t)4 * h * h), 256, 0, st>>>(W32(H, p + "w_hh"),
                                                                        H->whhT + ((size_t)l * 2 + dr) * h * 4 * h, 4 * h, h);
      KCHECK(H);
      add_vec_kernel<<<ew_grid(4 * h), 256, 0, st>>>(W32(H, p + "b_ih"), W32(H, p + "b_hh"),
                                                     H->lstm_b + ((size_t)l * 2 + dr) * 4 * h, 4 * h, 4 * h);
      KCHECK(H);
    }
  {  // split-bf16 predictor weights
    const size_t ds = c.d_sty_tok, dh = c.d_hid, Ds = c.d_style, kin = dh + ds;
    H->pred_tc = ds % 128 == 0 && dh % 128 == 0 && kin % 64 == 0 && c.d_text % 64 == 0 && Ds % 64 == 0;
    if (H->pred_tc) {
      auto split = [&](const float* src, bf16* dst, size_t N, size_t K) -> int {
        split3_weights_kernel<<<ew_grid(N * K), 256, 0, st>>>(src, dst, N, K);
        KCHECK(H);
        return 0;
      };
      CK(H, cudaMalloc(&H->wq3, ds * 3 * c.d_text * sizeof(bf16)));
      CK(H, cudaMalloc(&H->wkv3, 2 * ds * 3 * Ds * sizeof(bf16)));
      CK(H, cudaMalloc(&H->wo3, ds * 3 * ds * sizeof(bf16)));
      CK(H, cudaMalloc(&H->wih3, (size_t)c.n_lstm * 8 * h * 3 * kin * sizeof(bf16)));
      CK(H, cudaMalloc(&H->wada3, (size_t)(c.n_lstm > 1 ? c.n_lstm - 1 : 1) * 2 * dh * 3 * ds * sizeof(bf16)));
      CK(H, cudaMalloc(&H->b_kv, 2 * ds * sizeof(float)));
      RET(split(W32(H, "sp.q.w"), H->wq3, ds, c.d_text));
      RET(split(W32(H, "sp.k.w"), H->wkv3, ds, Ds));
      RET(split(W32(H, "sp.v.w"), H->wkv3 + ds * 3 * Ds, ds, Ds));
      RET(split(W32(H, "sp.o.w"), H->wo3, ds, ds));
      CK(H, cudaMemcpyAsync(H->b_kv, W32(H, "sp.k.b"), ds * sizeof(float), cudaMemcpyDeviceToDevice, st));
      CK(H, cudaMemcpyAsync(H->b_kv + ds, W32(H, "sp.v.b"), ds * sizeof(float), cudaMemcpyDeviceToDevice, st));
      // prosody heads: shared BiLSTM (both directions' input projections in one GEMM) + the two heads' hidden layer
      CK(H, cudaMalloc(&H->wih3_pros, (size_t)8 * h * 3 * kin * sizeof(bf16)));
      CK(H, cudaMalloc(&H->wh1_3, (size_t)dh * 3 * kin * sizeof(bf16)));
      CK(H, cudaMalloc(&H->whh_pros, (size_t)2 * 4 * h * h * sizeof(float)));
      CK(H, cudaMalloc(&H->lstm_b_pros, (size_t)2 * 4 * h * sizeof(float)));
      for (int dr = 0; dr < 2; ++dr) {
        const std::string p = std::string("pros.lstm.") + (dr ? "r." : "f.");
        RET(split(W32(H, p + "w_ih"), H->wih3_pros + (size_t)dr * 4 * h * 3 * kin, 4 * h, kin));
        CK(H, cudaMemcpyAsync(H->whh_pros + (size_t)dr * 4 * h * h, W32(H, p + "w_hh"), (size_t)4 * h * h * sizeof(float),
                              cudaMemcpyDeviceToDevice, st));
        add_vec_kernel<<<ew_grid(4 * h), 256, 0, st>>>(W32(H, p + "b_ih"), W32(H, p + "b_hh"), H->lstm_b_pros + (size_t)dr * 4 * h, 4 * h, 4 * h);
        KCHECK(H);
      }
      RET(split(W32(H, "pros.h1.w"), H->wh1_3, dh, kin));
      for (int l = 0; l < c.n_lstm; ++l) {
        for (int dr = 0; dr < 2; ++dr)
          RET(split(W32(H, "lstm" + std::to_string(l) + (dr ? ".r.w_ih" : ".f.w_ih")),
                    H->wih3 + ((size_t)l * 8 * h + (size_t)dr * 4 * h) * 3 * kin, 4 * h, kin));
        if (l < c.n_lstm - 1)
          RET(split(W32(H, "adaln" + std::to_string(l) + ".w"), H->wada3 + (size_t)l * 2 * dh * 3 * ds, 2 * dh, ds));
      }
    }
  }
  CK(H, cudaStreamSynchronize(st));
  cudaFree(tmp); cudaFree(nullc);
  return 0;
}

extern "C" int stz_create(const stz_config* cfg, const float* weights_host, size_t nfloats, int device,
                          stz_handle** out) {
  if (!cfg || !weights_host || !out) return fail(nullptr, STZ_E_ARG, "null argument");
  *out = nullptr;
  if (check_config(*cfg)) return fail(nullptr, STZ_E_SHAPE, "configuration outside kernel support (see check_config)");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev)
    return fail(nullptr, STZ_E_DEVICE, "CUDA device %d not available (count %d): no CPU fallback", device, ndev);
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major != 10 || prop.minor != 0)
    return fail(nullptr, STZ_E_DEVICE, "device %d is sm_%d%d; this library is sm_100a only", device, prop.major, prop.minor);
  if (cudaSetDevice(device) != cudaSuccess) return fail(nullptr, STZ_E_DEVICE, "cudaSetDevice(%d) failed", device);
  if (load_encode()) return fail(nullptr, STZ_E_DEVICE, "cuTensorMapEncodeTiled entry point not found");
  stz_handle* H = new stz_handle();
  H->cfg = *cfg;
  H->device = device;
  H->layout = build_layout(*cfg, &H->n_floats);
  for (const WeightEntry& e : H->layout) H->off[e.name] = e.off;
  if (nfloats != H->n_floats) {
    fail(nullptr, STZ_E_ARG, "weight blob has %zu floats, layout needs %zu", nfloats, H->n_floats);
    delete H;
    return STZ_E_ARG;
  }
  LaunchScope ls(H);
  int rc = create_impl(H, weights_host);
  if (rc != 0) {
    g_create_error = H->err;
    stz_destroy(H);
    return rc;
  }
  H->launches = 0; H->cur_launches = 0;
  *out = H;
  return 0;
}

extern "C" int64_t stz_launch_count(const stz_handle* h) { return h ? h->launches : 0; }

extern "C" int stz_graph_count(const stz_handle* h, int64_t* captures_total) {
  if (!h) return STZ_E_ARG;
  if (captures_total) *captures_total = h->graph_captures;
  return (int)h->graphs.size();
}

// Sizes the workspace arena for the largest call the caller will make, so that no later call reallocates it (a
// reallocation synchronises the device and invalidates every captured graph).
extern "C" int stz_reserve(stz_handle* H, int max_B, int max_T, int max_P, int max_steps, int sampler_kind) {
  if (!H) return STZ_E_ARG;
  if (max_B < 1 || max_T < 1 || max_P < 1 || max_steps < 1 || max_steps > 1024) return fail(H, STZ_E_ARG, "bad sizes");
  if (sampler_kind < STZ_SAMPLER_STUDENT || sampler_kind > STZ_SAMPLER_GUIDED) return fail(H, STZ_E_ARG, "bad sampler kind %d", sampler_kind);
  if (cudaSetDevice(H->device) != cudaSuccess) return fail(H, STZ_E_DEVICE, "cudaSetDevice failed");
  const int E = sampler_kind == STZ_SAMPLER_TEACHER ? 2 * max_steps : max_steps;
  const int slices = sampler_kind == STZ_SAMPLER_TEACHER ? max_steps + 1 : 1;
  for (auto& sl : H->slot) if (sl.busy) return fail(H, STZ_E_ARG, "stz_reserve with a host call in flight");
  return ensure_workspace(H, max_B, H->t_buckets ? bucket_T(max_T) : max_T, max_P, E, slices);
}

extern "C" int stz_set_option(stz_handle* H, const char* key, int value) {
  if (!H || !key) return STZ_E_ARG;
  if (!strcmp(key, "use_graph")) H->use_graph = value;
  else if (!strcmp(key, "gemm_impl")) {
    if (H->gemm_impl != value) drop_graphs(H);
    H->gemm_impl = value;
  } else if (!strcmp(key, "lstm_impl") || !strcmp(key, "lstm_nb")) {
    if (!strcmp(key, "lstm_nb") && value != 0 && value != 8 && value != 16 && value != 24) return fail(H, STZ_E_ARG, "lstm_nb must be 0, 8, 16 or 24");
    drop_graphs(H);
    (key[5] == 'i' ? H->lstm_impl : H->lstm_nb) = value;
  }
  else if (!strcmp(key, "pred_gemm_impl")) H->pred_gemm_impl = value;
  else if (!strcmp(key, "fuse_ln") || !strcmp(key, "gln_tile_rows") || !strcmp(key, "attn_box2") || !strcmp(key, "attn_ctas") || !strcmp(key, "ablate") || !strcmp(key, "attn_impl") || !strcmp(key, "chains") ||
           !strcmp(key, "gemm_bn") || !strcmp(key, "use_pdl") || !strcmp(key, "gemm_cluster")) {   // baked into captured graphs
    drop_graphs(H);
    if (!strcmp(key, "chains")) H->chains = value;
    else if (!strcmp(key, "ablate")) H->ablate = value;
    else if (!strcmp(key, "attn_impl")) H->attn_impl = value;
    else if (!strcmp(key, "attn_ctas")) H->attn_ctas = value;
    else if (!strcmp(key, "gemm_bn")) H->gemm_bn = value;
    else if (!strcmp(key, "use_pdl")) H->use_pdl = value;
    else if (!strcmp(key, "gemm_cluster")) H->gemm_cluster = value;
    else if (!strcmp(key, "gln_tile_rows")) H->gln_tile_rows = value;
    else if (!strcmp(key, "attn_box2")) H->attn_box2 = value;
    else H->fuse_ln = value;
  } else if (!strcmp(key, "t_buckets")) H->t_buckets = value;
  else if (!strcmp(key, "guard_bytes")) {    // re-plan both arenas with (or without) poisoned gaps
    if (value < 0) return fail(H, STZ_E_ARG, "guard_bytes must be >= 0");
    for (auto& sl : H->slot) if (sl.busy) return fail(H, STZ_E_ARG, "guard_bytes with a host call in flight");
    if (cudaSetDevice(H->device) != cudaSuccess) return fail(H, STZ_E_DEVICE, "cudaSetDevice failed");
    CK(H, cudaDeviceSynchronize());
    drop_graphs(H);
    if (H->ws.base) { CK(H, cudaFree(H->ws.base)); H->ws = Workspace(); }
    if (H->pws.base) { CK(H, cudaFree(H->pws.base)); H->pws = stz_handle::ProsodyWs(); }
    H->guard_bytes = value;
  }
  else if (!strcmp(key, "max_graphs")) {
    if (value < 1) return fail(H, STZ_E_ARG, "max_graphs must be >= 1");
    H->max_graphs = value;
  } else if (!strcmp(key, "profile")) {
    H->profile = value;
    for (auto& r : H->prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    H->prof.clear();
  } else return fail(H, STZ_E_ARG, "unknown option %s", key);
  return 0;
}

extern "C" int stz_get_option(const stz_handle* H, const char* key, int* value) {
  if (!H || !key || !value) return STZ_E_ARG;
  if (!strcmp(key, "use_graph")) *value = H->use_graph;
  else if (!strcmp(key, "t_buckets")) *value = H->t_buckets;
  else if (!strcmp(key, "max_graphs")) *value = H->max_graphs;
  else if (!strcmp(key, "use_pdl")) *value = H->use_pdl;
  else if (!strcmp(key, "gemm_impl")) *value = H->gemm_impl;
  else if (!strcmp(key, "gemm_bn")) *value = H->gemm_bn;
  else if (!strcmp(key, "fuse_ln")) *value = H->fuse_ln;
  else if (!strcmp(key, "gln_tile_rows")) *value = H->gln_tile_rows;
  else if (!strcmp(key, "attn_impl")) *value = H->attn_impl;
  else if (!strcmp(key, "attn_ctas")) *value = H->attn_ctas;
  else if (!strcmp(key, "chains")) *value = H->chains;
  else if (!strcmp(key, "lstm_impl")) *value = H->lstm_impl;
  else if (!strcmp(key, "lstm_nb")) *value = H->lstm_nb;
  else if (!strcmp(key, "pred_gemm_impl")) *value = H->pred_gemm_impl;
  else if (!strcmp(key, "profile")) *value = H->profile;
  else if (!strcmp(key, "guard_bytes")) *value = H->guard_bytes;
  else if (!strcmp(key, "last_fuse_mode")) *value = H->last_fuse_mode;
  else if (!strcmp(key, "last_T")) *value = H->last_T;
  else return STZ_E_ARG;
  return 0;
}

__global__ void guard_check_kernel(const uint8_t* __restrict__ p, size_t n, unsigned long long* __restrict__ bad) {
  unsigned long long local = 0;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x)
    local += p[i] != STZ_GUARD_BYTE ? 1 : 0;
  if (local) atomicAdd(bad, local);
}

// Counts the guard-gap bytes of both workspace arenas that no longer hold the poison pattern (synchronises the device).
// Returns the number of gaps checked or a negative status; *bad_bytes = 0 means no kernel wrote outside its buffers.
extern "C" int stz_debug_check_guards(stz_handle* H, long long* bad_bytes) {
  if (!H || !bad_bytes) return STZ_E_ARG;
  if (cudaSetDevice(H->device) != cudaSuccess) return fail(H, STZ_E_DEVICE, "cudaSetDevice failed");
  CK(H, cudaDeviceSynchronize());
  unsigned long long* bad = nullptr;
  CK(H, cudaMalloc(&bad, sizeof *bad));
  CK(H, cudaMemset(bad, 0, sizeof *bad));
  int n = 0;
  auto scan = [&](const char* base, const std::vector<std::pair<size_t, size_t>>& gaps) {
    if (!base) return;
    for (auto& g : gaps) {
      if (g.second == 0) continue;
      guard_check_kernel<<<(unsigned)((g.second + 255) / 256 < 64 ? (g.second + 255) / 256 : 64), 256>>>((const uint8_t*)base + g.first, g.second, bad);
      ++n;
    }
  };
  scan(H->ws.base, H->guards);
  scan(H->pws.base, H->guards_pros);
  unsigned long long hb = 0;
  cudaError_t e = cudaMemcpy(&hb, bad, sizeof hb, cudaMemcpyDeviceToHost);
  cudaFree(bad);
  if (e != cudaSuccess) return fail(H, STZ_E_CUDA, "guard check -> %s", cudaGetErrorString(e));
  *bad_bytes = (long long)hb;
  return n;
}

// BiLSTM recurrence on tcgen05 (predictor_tc.cuh).  Sequences per cluster = the NB in {8, 16, 24} that minimises
// waves x step time.  Only g_lstm_max_clusters (15 on a B200: one GPC holds a single cluster of 8) clusters run at a time, and
// with the machine full a step at NB = 24 measured 1.66 x a step at NB = 16 (the DSMEM exchange grows with NB and clusters
// sharing a GPC share its bandwidth): NB = 24 pays where it saves a whole wave — B = 121 .. 168 (16 .. 22 clusters = two
// waves at NB = 16, one at NB = 24: predictor -10 % at B = 128) — and loses elsewhere (B = 256: +6 %).  NB = 8 only for
// B <= 16: beyond 4 clusters it measured slower (profiles/r02_ab_lstm_forms.txt).
// lstm_impl 3 = W_hh operand in shared memory (A/B baseline).
static int lstm_pick_nb(int B) {
  if (B <= 16) return 8;
  int best = 16;
  long best_cost = -1;
  for (int nb : {16, 24}) {
    const long clusters = 2L * cdiv(B, nb), waves = (clusters + g_lstm_max_clusters - 1) / g_lstm_max_clusters;
    const long cost = waves * (nb == 24 ? 166 : 100);
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = nb; }
  }
  return best;
}
static void launch_lstm_tc(stz_handle* H, cudaStream_t st, const float* G, const float* whh, const int* lens, const int* perm,
                           float* out, int B, int T) {
  if (H->lstm_impl == 3) {
    launch_k(lstm_tc_kernel<false, 16>, dim3(cdiv(B, 16) * LC_CS, 2), lt_threads(16), lt_smem_bytes(false, 16), st, G, whh, lens, perm, out, B, T);
    return;
  }
  const int nb = H->lstm_nb != 0 ? H->lstm_nb : lstm_pick_nb(B);
  if (nb == 8) launch_k(lstm_tc_kernel<true, 8>, dim3(cdiv(B, 8) * LC_CS, 2), lt_threads(8), lt_smem_bytes(true, 8), st, G, whh, lens, perm, out, B, T);
  else if (nb == 24) launch_k(lstm_tc_kernel<true, 24>, dim3(cdiv(B, 24) * LC_CS, 2), lt_threads(24), lt_smem_bytes(true, 24), st, G, whh, lens, perm, out, B, T);
  else launch_k(lstm_tc_kernel<true, 16>, dim3(cdiv(B, 16) * LC_CS, 2), lt_threads(16), lt_smem_bytes(true, 16), st, G, whh, lens, perm, out, B, T);
}

// Max co-resident 8-CTA clusters of the BiLSTM recurrence kernel on the current device (diagnostic).
extern "C" int stz_debug_max_lstm_clusters(void) {
  if (init_kernel_attrs() != cudaSuccess) return -2;
  return g_lstm_max_clusters;
}

extern "C" int stz_profile_read(stz_handle* H, int cls, double* ms, double* work, int64_t* launches) {
  if (!H || cls < 0 || cls >= PC_COUNT) return STZ_E_ARG;
  if (cudaSetDevice(H->device) != cudaSuccess) return fail(H, STZ_E_DEVICE, "cudaSetDevice failed");
  CK(H, cudaDeviceSynchronize());
  double t = 0, w = 0;
  int64_t n = 0;
  for (auto& r : H->prof) {
    if (r.cls != cls) continue;
    float e = 0.f;
    CK(H, cudaEventElapsedTime(&e, r.a, r.b));
    t += e; w += r.work; ++n;
  }
  if (ms) *ms = t;
  if (work) *work = w;
  if (launches) *launches = n;
  return 0;
}

extern "C" int stz_debug_set_gemm_trace(stz_handle* H, long long* trace_dev) {
  if (!H) return STZ_E_ARG;
  if (cudaSetDevice(H->device) != cudaSuccess) return fail(H, STZ_E_DEVICE, "cudaSetDevice failed");
  CK(H, cudaMemcpyToSymbol(g_gemm_trace, &trace_dev, sizeof trace_dev));
  return 0;
}

#ifdef STZ_TRACE   // experiment flags of the timeline build only (tools/gemm_trace.py); not part of include/stz.h
extern "C" int stz_debug_set_gemm_dbg(stz_handle* H, int flags) {
  if (!H) return STZ_E_ARG;
  if (cudaSetDevice(H->device) != cudaSuccess) return fail(H, STZ_E_DEVICE, "cudaSetDevice failed");
  CK(H, cudaMemcpyToSymbol(g_gemm_dbg, &flags, sizeof flags));
  return 0;
}
#endif

extern "C" int stz_debug_set_lstm_trace(stz_handle* H, long long* trace_dev) {
  if (!H) return STZ_E_ARG;
  if (cudaSetDevice(H->device) != cudaSuccess) return fail(H, STZ_E_DEVICE, "cudaSetDevice failed");
  CK(H, cudaMemcpyToSymbol(g_lstm_trace, &trace_dev, sizeof trace_dev));
  return 0;
}

extern "C" int stz_debug_set_att_trace(stz_handle* H, long long* trace_dev) {
  if (!H) return STZ_E_ARG;
  if (cudaSetDevice(H->device) != cudaSuccess) return fail(H, STZ_E_DEVICE, "cudaSetDevice failed");
  CK(H, cudaMemcpyToSymbol(g_att_trace, &trace_dev, sizeof trace_dev));
  return 0;
}

extern "C" int stz_debug_set_tap(stz_handle* H, int eval, int layer, int stage, float* tap_dev) {
  if (!H) return STZ_E_ARG;
  H->tap_eval = eval; H->tap_layer = layer; H->tap_stage = stage; H->tap_buf = tap_dev;
  return 0;
}

// ------------------------------------------------------------------------------------------
// sample_style
// ------------------------------------------------------------------------------------------
static int tap(stz_handle* H, cudaStream_t st, int e, int l, int s, size_t R) {
  if (H->tap_buf && !H->capturing && H->tap_eval == e && H->tap_layer == l && H->tap_stage == s)
    CK(H, cudaMemcpyAsync(H->tap_buf, H->ws.h, R * H->cfg.d_model * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return 0;
}

__global__ void empty_pdl_kernel(int) { pdl_sync(); }

static int attention(stz_handle* H, cudaStream_t st, const AttnParams& ap, int B) {
  if (H->ablate & 512) {   // timing attribution: what a dependent kernel node costs when it does nothing
    launch_k(empty_pdl_kernel, 148, 128, 0, st, 0);
    KCHECK(H);
    return 0;
  }
  dim3 grid(H->cfg.n_heads, B);
  int n_keys = 0;
  for (int i = 0; i < ap.nseg; ++i) n_keys += ap.seg[i].n;
  ProfScope ps(H, st, PC_ATTN, 4.0 * B * H->cfg.n_heads * ap.n_q * (double)n_keys * ATT_DH);
  const int n_style = ap.single ? ap.n_q : ap.n_q / 2, dm = H->cfg.d_model;
  const int nbr = ap.single ? 1 : 2;     // branches interleaved in the rows of one utterance
  const bool self = ap.nseg == 1 && ap.seg[0].rule == KEY_SAME_BRANCH;
  const bool cross3 = ap.nseg == 3 && ap.seg[0].rule == KEY_ALL && ap.seg[1].rule == KEY_COND && ap.seg[2].rule == KEY_UNCOND && ap.seg[2].n == 1;
  const int T8 = cross3 ? (ap.seg[0].n + 7) / 8 * 8 : 0, P8 = cross3 ? (ap.seg[1].n + 7) / 8 * 8 : 0;
  if (H->attn_impl == 0 && ap.n_q <= 128 && n_style <= 64 && (self || (cross3 && T8 + P8 + 1 <= 128))) {
    // tcgen05 attention, TMA-staged operands (attention_tc2_kernel)
    CUtensorMap tq, tt, tp, tn;
    bool box2 = false;
    memset(&tp, 0, sizeof tp); memset(&tn, 0, sizeof tn);
    const int qcols = self ? (int)(ap.seg[0].v - ap.q) + dm : dm;
    {  // Q (and, for self-attention, K / V): (column, branch, token) view of the R-layout buffer
      // (column, token, branch) view, box 64 x 64 x branches: one TMA instruction lands both branches de-interleaved
      const cuuint64_t gd2[3] = {(cuuint64_t)qcols, (cuuint64_t)B * n_style, (cuuint64_t)nbr};
      const cuuint64_t gs2[2] = {(cuuint64_t)ap.ldq * 2 * nbr, (cuuint64_t)ap.ldq * 2};
      const cuuint32_t bx2[3] = {64, 64, (cuuint32_t)nbr};
      box2 = H->attn_box2 && make_tmap_nd(&tq, ap.q, 3, gd2, gs2, bx2) == 0;
      if (!box2) {   // (column, branch, token) view, one box per branch
        const cuuint64_t gd[3] = {(cuuint64_t)qcols, (cuuint64_t)nbr, (cuuint64_t)B * n_style};
        const cuuint64_t gs[2] = {(cuuint64_t)ap.ldq * 2, (cuuint64_t)ap.ldq * 2 * nbr};
        const cuuint32_t bx[3] = {64, 1, 64};
        if (make_tmap_nd(&tq, ap.q, 3, gd, gs, bx)) return fail(H, STZ_E_CUDA, "cuTensorMapEncodeTiled (attention Q) failed");
      }
    }
    tt = tq;
    // self-attention with Q | K | V side by side in one buffer: a 4-D view (column, token, branch, part) whose box
    // 64 x 64 x 2 x 3 lands all three operand tiles with ONE TMA instruction (48 KB; TMA issues serialise per SM)
    bool box3 = false;
    if (self && box2 && nbr == 2 && (int)(ap.seg[0].k - ap.q) == dm && (int)(ap.seg[0].v - ap.q) == 2 * dm) {
      const cuuint64_t gd4[4] = {(cuuint64_t)dm, (cuuint64_t)B * n_style, 2, 3};
      const cuuint64_t gs4[3] = {(cuuint64_t)ap.ldq * 2 * 2, (cuuint64_t)ap.ldq * 2, (cuuint64_t)dm * 2};
      const cuuint32_t bx4[4] = {64, 64, 2, 3};
      box3 = make_tmap_nd(&tt, ap.q, 4, gd4, gs4, bx4) == 0;
      if (!box3) tt = tq;
    }
    AttnTcParams tp_{};
    tp_.out = ap.out; tp_.ldo = ap.ldo; tp_.n_q = ap.n_q; tp_.n_heads = H->cfg.n_heads; tp_.n_units = B * H->cfg.n_heads;
    tp_.self = self ? 1 : 0; tp_.n_style = n_style; tp_.scale_log2 = ap.scale_log2; tp_.single = ap.single; tp_.box2 = box3 ? 2 : (box2 ? 1 : 0);
    if (self) {
      tp_.col_k = (int)(ap.seg[0].k - ap.q); tp_.col_v = (int)(ap.seg[0].v - ap.q);
    } else {
      const AttnSeg* sg = ap.seg;
      const CUtensorMap* maps[3] = {&tt, &tp, &tn};
      for (int i = 0; i < 3; ++i) {   // K | V of this layer = 2 * d_model columns starting at seg.k
        const cuuint64_t rows = i == 2 ? 1 : (cuuint64_t)B * sg[i].n;
        const cuuint64_t gd[2] = {(cuuint64_t)2 * dm, rows};
        const cuuint64_t gs[1] = {(cuuint64_t)sg[i].ld * 2};
        const cuuint32_t bx[2] = {64, (cuuint32_t)sg[i].n};
        if ((int)(sg[i].v - sg[i].k) != dm || make_tmap_nd(const_cast<CUtensorMap*>(maps[i]), sg[i].k, 2, gd, gs, bx))
          return fail(H, STZ_E_CUDA, "cuTensorMapEncodeTiled (attention K/V segment %d) failed", i);
      }
      tp_.T = sg[0].n; tp_.P = sg[1].n; tp_.T8 = T8; tp_.P8 = P8; tp_.col_k = 0; tp_.col_v = dm;
      tp_.tmask = sg[0].mask; tp_.pmask = sg[1].mask;
    }
    const int units = tp_.n_units;
    // one wave of 4 CTAs/SM but not of 2 CTAs/SM (38 <= B <= 74 at 8 heads): attention_tc4_kernel measured -2 % on the cfg2
    // step and -3.6 % on the guided student; outside that range attention_tc2_kernel (prefetching, 8 warps) is 1-3 % faster
    const int ctas = H->attn_ctas != 0 ? H->attn_ctas : (units > 2 * g_num_sms && units <= 4 * g_num_sms ? 4 : 2);
    if (ctas == 4) launch_k(attention_tc4_kernel, units < 4 * g_num_sms ? units : 4 * g_num_sms, 128, ATC4_SMEM_BYTES, st, tq, tt, tp, tn, tp_);
    else launch_k(attention_tc2_kernel, units < 2 * g_num_sms ? units : 2 * g_num_sms, 256, ATC_SMEM_BYTES, st, tq, tt, tp, tn, tp_);
  } else if (H->attn_impl == 0 && cross3 && ap.n_q <= 128 && n_style <= 64 && P8 + 1 <= 128) {
    // long text: streaming tcgen05 attention over 128-key blocks (attention_tcs_kernel)
    CUtensorMap tq, tt, tp, tn;
    bool box2 = false;
    {
      const cuuint64_t gd2[3] = {(cuuint64_t)dm, (cuuint64_t)B * n_style, (cuuint64_t)nbr};
      const cuuint64_t gs2[2] = {(cuuint64_t)ap.ldq * 2 * nbr, (cuuint64_t)ap.ldq * 2};
      const cuuint32_t bx2[3] = {64, 64, (cuuint32_t)nbr};
      box2 = H->attn_box2 && make_tmap_nd(&tq, ap.q, 3, gd2, gs2, bx2) == 0;
      if (!box2) {
        const cuuint64_t gd[3] = {(cuuint64_t)dm, (cuuint64_t)nbr, (cuuint64_t)B * n_style};
        const cuuint64_t gs[2] = {(cuuint64_t)ap.ldq * 2, (cuuint64_t)ap.ldq * 2 * nbr};
        const cuuint32_t bx[3] = {64, 1, 64};
        if (make_tmap_nd(&tq, ap.q, 3, gd, gs, bx)) return fail(H, STZ_E_CUDA, "cuTensorMapEncodeTiled (attention Q) failed");
      }
    }
    const AttnSeg* sg = ap.seg;
    CUtensorMap* maps[3] = {&tt, &tp, &tn};
    for (int i = 0; i < 3; ++i) {
      const cuuint64_t rows = i == 2 ? 1 : (cuuint64_t)B * sg[i].n;
      const cuuint64_t gd[2] = {(cuuint64_t)2 * dm, rows};
      const cuuint64_t gs[1] = {(cuuint64_t)sg[i].ld * 2};
      const cuuint32_t bx[2] = {64, (cuuint32_t)(i == 0 ? 128 : sg[i].n)};     // text: fixed 128-row blocks
      if ((int)(sg[i].v - sg[i].k) != dm || make_tmap_nd(maps[i], sg[i].k, 2, gd, gs, bx))
        return fail(H, STZ_E_CUDA, "cuTensorMapEncodeTiled (attention K/V segment %d) failed", i);
    }
    AttnTcParams tp_{};
    tp_.out = ap.out; tp_.ldo = ap.ldo; tp_.n_q = ap.n_q; tp_.n_heads = H->cfg.n_heads; tp_.n_units = B * H->cfg.n_heads;
    tp_.self = 0; tp_.n_style = n_style; tp_.scale_log2 = ap.scale_log2; tp_.single = ap.single; tp_.box2 = box2 ? 1 : 0;
    tp_.T = sg[0].n; tp_.P = sg[1].n; tp_.T8 = T8; tp_.P8 = P8; tp_.col_k = 0; tp_.col_v = dm;
    tp_.tmask = sg[0].mask; tp_.pmask = sg[1].mask;
    const int units = tp_.n_units;
    launch_k(attention_tcs_kernel, units < 2 * g_num_sms ? units : 2 * g_num_sms, 256, ATS_SMEM_BYTES, st, tq, tt, tp, tn, tp_);
  } else {
    if (ap.single) return fail(H, STZ_E_SHAPE, "the guidance-conditioned student needs the tcgen05 attention kernels (n_style <= 64, prompt <= 127 tokens)");
    launch_k(attention_kernel, grid, 2 * cdiv(ap.n_q / 2, 16) * 32, ATT_SMEM_BYTES, st, ap);   // ceil(K / 16) warps per CFG branch
  }
  KCHECK(H);
  return 0;
}

// One denoiser evaluation + fused guidance/sampler update (a-4, a-5, a-6) for utterances [b0, b0 + nb) of a batch of
// Btot.  Utterances never interact, so the evaluation loop of a batch may run as several independent chains
// (sub-ranges) on parallel graph branches: one chain's launch gaps, prologues and tails are filled by the others.
// nbr = 2: the CFG pair layout (row = (b*K + k)*2 + branch, conditional / unconditional batched in one launch); nbr = 1: the
// guidance-conditioned student (row = b*K + k, one branch, the guidance scale enters through the conditioning vector).
static int run_eval(stz_handle* H, cudaStream_t st, int e, int Btot, int b0, int nb, int T, int P, const uint8_t* tmask,
                    const uint8_t* pmask, int nbr) {
  const stz_config& c = H->cfg;
  const Workspace& w = H->ws;
  const int d = c.d_model, L = c.n_layers, K = c.n_style, Ds = c.d_style, n_mod = (9 * L + 2) * d;
  const int R = nbr * nb * K, NS = nbr * nb, impl = H->gemm_impl, single = nbr == 1 ? 1 : 0;
  const size_t r0 = (size_t)nbr * b0 * K, s0 = (size_t)nbr * b0;     // first activation row / first sequence of this chain
  float* mod = w.mod + s0 * n_mod;
  float* h = w.h + r0 * d;
  bf16 *u = w.u + r0 * d, *u3 = w.u3 + r0 * 3 * d, *qkv = w.qkv + r0 * 3 * d, *att = w.att + r0 * d, *ffh = w.ffh + r0 * c.d_ff;
  bf16* xin = w.xin + r0 * 3 * Ds;
  if (tmask) tmask += (size_t)b0 * T;
  if (pmask) pmask += (size_t)b0 * P;
  const float scale_log2 = 1.4426950408889634f / sqrtf((float)(d / c.n_heads));
  GemmParams base{};
  base.mod = mod; base.n_mod = n_mod; base.rows_per_utt = nbr * K; base.n_style = K; base.single = single;
  // ablation bits: 1 self-attn, 2 cross-attn, 4 ln_mod, 8 qkv, 16 attention out-projections, 32 q2, 64 ff1, 128 ff2, 256 mod,
  // 512 empty dependent nodes, 1024 cross-attention of every layer reads layer 0's K/V (L2-resident)
  const int ab = H->ablate;

  if (w.mod_hoisted) {   // all evaluations' modulations were computed by one GEMM before the loop (sample_style_impl)
    mod = w.mod + ((size_t)e * nbr * Btot + s0) * n_mod;
    base.mod = mod;
  } else if (!(ab & 256)) {  // AdaLN modulations of this eval: mod[NS, n_mod] = c[e] · Wmod^T + b
    GemmParams p = base;
    p.M = NS; p.N = n_mod; p.K = d; p.a_row0 = e * nbr * Btot + (int)s0; p.bias = W32(H, "mod.b"); p.out = mod; p.ldo = n_mod;
    RET(gemm<EPI_F32>(H, st, impl, w.cvec, d, w.E * 2 * w.B + 128, WBF(H, "mod.w"), p));
  }
  // GEMM + residual + AdaLN in one kernel.  Product: gemmln3_kernel (fuse_ln = 3) whenever its shape constraints hold
  // (d_model 512, every contraction length a multiple of 256, a 128-row tile spanning <= 8 sequences); otherwise, and for
  // fuse_ln = 0, the GEMM + ln_mod_kernel pair.
  int fuse_mode = (impl == 0 && d == GLN_N) ? H->fuse_ln : 0;
  if (fuse_mode != 3 && fuse_mode != 4) fuse_mode = 0;
  if (fuse_mode != 0 && (d % 256 || c.d_ff % 256 || (3 * Ds) % 256 || nbr * ((GEMM_BM - 1 + nbr * K - 1) / (nbr * K) + 1) > GLN3_MAX_SEQ)) fuse_mode = 0;
  // The fused kernel wins where its CTA pairs fill ONE wave: 37 .. 74 row tiles (B = 47 .. 94 at K = 50: +7 .. 12 %, with
  // 96-row tiles where they fit).  Below, the separate LayerNorm kernel is cheap and the fused kernel's long serial epilogue
  // loses 2 .. 9 %; above (two and more waves of pairs) the persistent GEMM + LayerNorm pair wins 1 .. 5 % since its MMA issue
  // loop went warp-uniform (tools/ab_tile_rows.py, profiles/r02_ab_tile_rows.txt).  fuse_ln = 4 forces the fused kernel.
  {
    const int row_tiles = cdiv(R, GEMM_BM);
    if (fuse_mode == 3 && (row_tiles < 37 || row_tiles > g_num_sms / 2)) fuse_mode = 0;
  }
  if (fuse_mode == 4) fuse_mode = 3;
  const bool fused = fuse_mode != 0;
  H->last_fuse_mode = fuse_mode;
  GemmLnParams lb{};
  lb.M = R; lb.h = h; lb.mod = mod; lb.n_mod = n_mod; lb.rows_per_utt = nbr * K; lb.pos = W32(H, "pos"); lb.n_style = K; lb.single = single;
  // residual GEMM of a sub-layer followed by the AdaLN of the next one: (gate, shift, scale) offsets into mod
  auto res_ln = [&](const bf16* A, int Kc, const bf16* Wt, const float* bias, int gate_off, int ln_off, bool last) -> int {
    if (fused) {
      GemmLnParams p = lb;
      p.K = Kc; p.bias = bias; p.gate_off = gate_off; p.shift_off = ln_off; p.scale_off = ln_off + d; p.split3 = last ? 1 : 0;
      return launch_gemmln<GLN_RES>(H, st, A, Kc, R, Wt, last ? u3 : u, p);
    }
    GemmParams p = base;
    p.M = R; p.N = d; p.K = Kc; p.bias = bias; p.out = h; p.ldo = d; p.gate_off = gate_off;
    if (!(ab & (Kc == d ? 16 : 128))) RET(gemm<EPI_GATE_RES>(H, st, impl, A, Kc, R, Wt, p));
    if (ab & 4) return 0;
    return ln_mod(H, st, h, R, d, mod, n_mod, ln_off, ln_off + d, nbr * K, last ? u3 : u, last ? 1 : 0, single);
  };
  {  // h = x_in · Win^T + b + pos, u = AdaLN_1 of layer 0
    if (fused) {
      GemmLnParams p = lb;
      p.K = 3 * Ds; p.bias = W32(H, "in.b"); p.shift_off = 0; p.scale_off = d; p.split3 = 0;
      RET(launch_gemmln<GLN_POS>(H, st, xin, 3 * Ds, R, H->w_in3, u, p));
    } else {
      GemmParams p = base;
      p.M = R; p.N = d; p.K = 3 * Ds; p.bias = W32(H, "in.b"); p.out = h; p.ldo = d; p.pos = W32(H, "pos");
      RET(gemm<EPI_F32_POS>(H, st, impl, xin, 3 * Ds, R, H->w_in3, p));
      RET(ln_mod(H, st, h, R, d, mod, n_mod, 0, d, nbr * K, u, 0, single));
    }
  }
  for (int l = 0; l < L; ++l) {
    const std::string pf = "l" + std::to_string(l) + ".";
    const int mo = 9 * l * d;
    // --- self-attention (u holds AdaLN_1(h))
    if (!(ab & 8)) {
      GemmParams p = base;
      p.M = R; p.N = 3 * d; p.K = d; p.bias = W32(H, pf + "qkv.b"); p.out = qkv; p.ldo = 3 * d;
      RET(gemm<EPI_BF16>(H, st, impl, u, d, R, WBF(H, pf + "qkv.w"), p));
    }
    if (!(ab & 1)) {
      AttnParams ap{};
      ap.q = qkv; ap.ldq = 3 * d; ap.out = att; ap.ldo = d; ap.n_q = nbr * K; ap.nseg = 1; ap.scale_log2 = scale_log2; ap.single = single;
      ap.seg[0] = AttnSeg{qkv + d, qkv + 2 * d, 3 * d, nbr * K, nbr * K, nullptr, KEY_SAME_BRANCH};
      RET(attention(H, st, ap, nb));
    }
    RET(res_ln(att, d, WBF(H, pf + "o.w"), W32(H, pf + "o.b"), mo + 2 * d, mo + 3 * d, false));
    if (nb == Btot) RET(tap(H, st, e, l, 0, R));
    // --- cross-attention over [text ; prompt | null]  (u holds AdaLN_2(h))
    if (!(ab & 32)) {
      GemmParams p = base;
      p.M = R; p.N = d; p.K = d; p.bias = W32(H, pf + "q2.b"); p.out = qkv; p.ldo = d;
      RET(gemm<EPI_BF16>(H, st, impl, u, d, R, WBF(H, pf + "q2.w"), p));
    }
    if (!(ab & 2)) {
      AttnParams ap{};
      const int ldkv = L * 2 * d;
      const int lk = (ab & 1024) ? 0 : l;   // attribution: every layer reads layer 0's K/V (L2-resident) -> cost of the HBM misses
      const bf16* kt = w.kv_text + ((size_t)b0 * T * L + lk) * 2 * d;
      const bf16* kp = w.kv_prompt + ((size_t)b0 * P * L + lk) * 2 * d;
      ap.q = qkv; ap.ldq = d; ap.out = att; ap.ldo = d; ap.n_q = nbr * K; ap.nseg = 3; ap.scale_log2 = scale_log2; ap.single = single;
      ap.seg[0] = AttnSeg{kt, kt + d, ldkv, T, T, tmask, KEY_ALL};
      ap.seg[1] = AttnSeg{kp, kp + d, ldkv, P, P, pmask, KEY_COND};
      ap.seg[2] = AttnSeg{H->kv_null + l * 2 * d, H->kv_null + l * 2 * d + d, ldkv, 1, 0, nullptr, KEY_UNCOND};
      RET(attention(H, st, ap, nb));
    }
    RET(res_ln(att, d, WBF(H, pf + "o2.w"), W32(H, pf + "o2.b"), mo + 5 * d, mo + 6 * d, false));
    if (nb == Btot) RET(tap(H, st, e, l, 1, R));
    // --- FFN  (u holds AdaLN_3(h)); its residual GEMM also produces AdaLN_1 of the next layer / the final AdaLN
    if (!(ab & 64)) {
      GemmParams p = base;
      p.M = R; p.N = c.d_ff; p.K = d; p.bias = W32(H, pf + "ff1.b"); p.out = ffh; p.ldo = c.d_ff;
      RET(gemm<EPI_GELU_BF16>(H, st, impl, u, d, R, WBF(H, pf + "ff1.w"), p));
    }
    RET(res_ln(ffh, c.d_ff, WBF(H, pf + "ff2.w"), W32(H, pf + "ff2.b"), mo + 8 * d, 9 * (l + 1) * d, l == L - 1));
    if (nb == Btot) RET(tap(H, st, e, l, 2, R));
  }
  {  // F = u · Wout^T + b, then CFG combine + sampler update + next input in the epilogue
    GemmParams p = base;
    p.M = R; p.N = Ds; p.K = 3 * d; p.bias = W32(H, "out.b"); p.ldo = Ds;
    p.x = w.x + (size_t)b0 * K * Ds; p.xmid = w.xmid + (size_t)b0 * K * Ds; p.xin = xin; p.coef = w.coef + (size_t)e * 8;
    // teacher: eval 2i+1 adds sigma_up * noise slice i+1 (slice 0 seeded the state)
    p.noise = w.noise + ((size_t)((e >> 1) + 1) * Btot + b0) * K * Ds;
    p.tap = (H->tap_buf && !H->capturing && H->tap_eval == e && H->tap_layer == L && nb == Btot) ? H->tap_buf : nullptr;
    RET(gemm<EPI_SAMPLER>(H, st, impl, u3, 3 * d, R, H->w_out3, p));
  }
  return 0;
}

static int sample_style_impl(stz_handle* H, const float* text, const uint8_t* tmask, const float* prompt,
                             const uint8_t* pmask, const float* noise, int B, int T, int P, int steps, float cfg_scale,
                             int kind, float* out, cudaStream_t st) {
  const stz_config& c = H->cfg;
  if (!text || !prompt || !out) return fail(H, STZ_E_ARG, "null tensor argument");
  if (!noise && !H->noise_seeded) return fail(H, STZ_E_ARG, "noise is NULL and no seed was set (stz_set_noise_seed)");
  if (B < 1 || T < 1 || P < 1 || steps < 1 || steps > 1024) return fail(H, STZ_E_ARG, "bad sizes B=%d T=%d P=%d steps=%d", B, T, P, steps);
  if (kind != STZ_SAMPLER_STUDENT && kind != STZ_SAMPLER_TEACHER && kind != STZ_SAMPLER_GUIDED) return fail(H, STZ_E_ARG, "bad sampler kind %d", kind);
  const int E = kind == STZ_SAMPLER_TEACHER ? 2 * steps : steps;
  const int slices = kind == STZ_SAMPLER_TEACHER ? steps + 1 : 1;
  const int nbr = kind == STZ_SAMPLER_GUIDED ? 1 : 2;      // the guidance-conditioned student runs ONE branch
  const bool graphed = runs_graphed(H);
  // length bucket: the call runs with T = bucket and a key-padding mask over the extra positions
  const int T_src = T;
  T = effective_T(H, T);
  RET(ensure_workspace(H, B, T, P, E, slices));
  RET(order_after_previous_call(H, st));
  Workspace& w = H->ws;
  H->last_T = T;
  if (T != T_src) {
    launch_k(pad_mask_kernel, ew_grid((size_t)B * T), 256, 0, st, tmask, w.st_tmask, B, T_src, T); KCHECK(H);
    tmask = w.st_tmask;
  }
  const int d = c.d_model, L = c.n_layers, K = c.n_style, Ds = c.d_style;
  const int NS = nbr * B, impl = H->gemm_impl;
  const size_t BK = (size_t)B * K;
  H->cur_launches = 0;

  // ---- schedule tables ----------------------------------------------------------------
  EvalPlan pl = make_plan(c, steps, kind, cfg_scale);
  CK(H, cudaMemcpyAsync(w.coef, pl.coef.data(), pl.coef.size() * sizeof(float), cudaMemcpyHostToDevice, st));
  CK(H, cudaMemcpyAsync(w.tfeat, pl.tfeat.data(), pl.tfeat.size() * sizeof(float), cudaMemcpyHostToDevice, st));

  // ---- conditioning prep (a-3): once per call ---------------------------------------------
  // text side first: with the host entry point the prompt / noise H2D copies are still in flight on the copy stream
  w.ctx_prompt = w.ctx_text + (size_t)B * T * d;
  w.kv_prompt = w.kv_text + (size_t)B * T * L * 2 * d;
  const int rows_all = B * (T + P);
  launch_k(cast_pool_kernel, dim3(B, c.d_text / 128), 256, 0, st, text, tmask, w.text_bf, w.pool_text, T, c.d_text, T_src); KCHECK(H);
  RET(linear_f32(H, st, ACT_NONE, w.pool_text, c.d_text, c.d_text, nullptr, 0, 0, W32(H, "ptext.w"), W32(H, "ptext.b"), w.pt, d, B, d));
  RET(linear_f32(H, st, ACT_SILU, w.tfeat, c.d_time, c.d_time, nullptr, 0, 0, W32(H, "time.w1"), W32(H, "time.b1"), w.t1, d, E, d));
  RET(linear_f32(H, st, ACT_NONE, w.t1, d, d, nullptr, 0, 0, W32(H, "time.w2"), W32(H, "time.b2"), w.temb, d, E, d));
  // Long padded text (T a multiple of 128, so GEMM row tiles coincide with the streaming attention's key blocks): row
  // tiles that are padding throughout are skipped by the context GEMMs — the attention never visits those key blocks.
  const uint8_t* ctx_needed = nullptr;
  if (tmask != nullptr && impl == 0 && T % 128 == 0 && T > 128) {
    launch_k(tile_needed_kernel, cdiv(cdiv(B * T, 128), 8), 256, 0, st, tmask, w.tile_needed_ctx, B * T, T / 128); KCHECK(H);
    CK(H, cudaMemsetAsync(w.tile_needed_ctx + B * T / 128, 1, cdiv(B * P, 128) + 1, st));
    ctx_needed = w.tile_needed_ctx;
  }
  {
    GemmParams p{};
    p.M = B * T; p.N = d; p.K = c.d_text; p.bias = H->ctx_text_b; p.out = w.ctx_pre; p.ldo = d; p.tile_needed = ctx_needed;
    RET(gemm<EPI_F32>(H, st, impl, w.text_bf, c.d_text, B * T, WBF(H, "ctx_text.w"), p));
  }
  if (H->wait_prompt) { CK(H, cudaStreamWaitEvent(st, H->cur_ev_prompt, 0)); H->wait_prompt = false; }
  launch_k(cast_pool_kernel, dim3(B, c.d_prompt / 128), 256, 0, st, prompt, pmask, w.prompt_bf, w.pool_prompt, P, c.d_prompt, P); KCHECK(H);
  RET(linear_f32(H, st, ACT_NONE, w.pool_prompt, c.d_prompt, c.d_prompt, nullptr, 0, 0, W32(H, "pprompt.w"), W32(H, "pprompt.b"), w.pp, d, B, d));
  const float* gemb = nullptr;
  if (nbr == 1) {   // guidance-scale embedding g = W_g2 SiLU(W_g1 feat(omega) + b_g1) + b_g2, added to the conditioning vector
    CK(H, cudaMemcpyAsync(w.gfeat, pl.gfeat.data(), pl.gfeat.size() * sizeof(float), cudaMemcpyHostToDevice, st));
    RET(linear_f32(H, st, ACT_SILU, w.gfeat, c.d_time, c.d_time, nullptr, 0, 0, W32(H, "gs.w1"), W32(H, "gs.b1"), w.g1, d, 1, d));
    RET(linear_f32(H, st, ACT_NONE, w.g1, d, d, nullptr, 0, 0, W32(H, "gs.w2"), W32(H, "gs.b2"), w.gemb, d, 1, d));
    gemb = w.gemb;
  }
  launch_k(cvec_kernel, ew_grid((size_t)E * NS * d), 256, 0, st, w.temb, w.pt, w.pp, W32(H, "null_pp"), w.cvec, E, NS, d, nbr == 1 ? 1 : 0, gemb); KCHECK(H);
  w.mod_hoisted = hoist_mod(c, B, E) && !(H->ablate & 256);
  if (w.mod_hoisted) {   // AdaLN modulations of every evaluation: mod[E * NS, n_mod] = c · Wmod^T + b
    const int n_mod = (9 * L + 2) * d;
    GemmParams p{};
    p.M = E * NS; p.N = n_mod; p.K = d; p.bias = W32(H, "mod.b"); p.out = w.mod; p.ldo = n_mod;
    RET(gemm<EPI_F32>(H, st, impl, w.cvec, d, w.E * 2 * w.B + 128, WBF(H, "mod.w"), p));
  }
  {
    GemmParams p{};
    p.M = B * P; p.N = d; p.K = c.d_prompt; p.bias = H->ctx_prompt_b; p.out = w.ctx_pre + (size_t)B * T * d; p.ldo = d;
    RET(gemm<EPI_F32>(H, st, impl, w.prompt_bf, c.d_prompt, B * P, WBF(H, "ctx_prompt.w"), p));
  }
  // context tokens [text ; prompt]: one LayerNorm, then every layer's K/V in ONE GEMM (N = L * 2d)
  RET(ln_mod(H, st, w.ctx_pre, rows_all, d, nullptr, 0, 0, 0, 1, w.ctx_text));
  {
    GemmParams q{};
    q.M = rows_all; q.N = L * 2 * d; q.K = d; q.bias = H->bkv_all; q.out = w.kv_text; q.ldo = L * 2 * d; q.tile_needed = ctx_needed;
    RET(gemm<EPI_BF16>(H, st, impl, w.ctx_text, d, rows_all, H->wkv_all, q));
  }
  if (H->wait_noise) { CK(H, cudaStreamWaitEvent(st, H->cur_ev_noise, 0)); H->wait_noise = false; }
  // ---- sampler state --------------------------------------------------------------------
  if (!noise) {   // every slice drawn on the device, bit-identical to oracle/philox.py
    const unsigned long long* ids = nullptr;
    if (!H->noise_ids.empty()) {
      if ((int)H->noise_ids.size() != B) return fail(H, STZ_E_ARG, "stz_set_noise_utterances gave %zu indices, the batch has %d utterances", H->noise_ids.size(), B);
      if (H->noise_ids_cap < (size_t)B) {
        CK(H, cudaStreamSynchronize(st));
        if (H->noise_ids_dev) CK(H, cudaFree(H->noise_ids_dev));
        H->noise_ids_dev = nullptr; H->noise_ids_cap = 0;
        CK(H, cudaMalloc(&H->noise_ids_dev, (size_t)B * sizeof(unsigned long long)));
        H->noise_ids_cap = B;
      }
      CK(H, cudaMemcpyAsync(H->noise_ids_dev, H->noise_ids.data(), (size_t)B * sizeof(unsigned long long), cudaMemcpyHostToDevice, st));
      ids = H->noise_ids_dev;
    }
    launch_k(philox_normal_kernel, ew_grid((size_t)slices * BK * Ds / 4), 256, 0, st, w.noise, (uint32_t)H->noise_seed,
             (uint32_t)(H->noise_seed >> 32), (unsigned long long)H->noise_first_utt, ids, slices, B, K * Ds / 4); KCHECK(H);
    noise = w.noise;
  } else if (kind == STZ_SAMPLER_TEACHER) {
    CK(H, cudaMemcpyAsync(w.noise, noise, (size_t)slices * BK * Ds * sizeof(float), cudaMemcpyDeviceToDevice, st));
  }
  {
    // Launched WITHOUT programmatic serialization: a full dependency on everything before it.  Kernels of the evaluation
    // loop read the call's constants (context K/V, modulations) before their griddepcontrol.wait (attention_tc2_kernel's
    // early K/V boxes); this launch guarantees the conditioning prep has completed before any of them can start, also
    // for tiny grids where a whole chain of waiting kernels is co-resident.
    launch_k_nopdl(init_state_kernel, ew_grid(BK * Ds / 4), 256, 0, st, noise, w.x, w.xin, BK, Ds, (float)pl.sigma0, (float)pl.cin0, nbr);
    KCHECK(H);
  }
  H->launches += H->cur_launches;
  H->cur_launches = 0;

  // ---- the evaluation loop: one CUDA graph per (B, T, P, E, kind) ----------------------------
  if (graphed) {
    auto key = std::make_tuple(B, T, P, E, kind * 4 + (tmask ? 1 : 0) + (pmask ? 2 : 0));
    auto it = H->graphs.find(key);
    // the graph bakes in the mask pointers: masks are copied into library-owned staging first
    const uint8_t* tm = nullptr; const uint8_t* pm = nullptr;
    if (tmask == w.st_tmask) tm = w.st_tmask;   // the bucket mask was built there
    else if (tmask) { CK(H, cudaMemcpyAsync(w.st_tmask, tmask, (size_t)B * T, cudaMemcpyDeviceToDevice, st)); tm = w.st_tmask; }
    if (pmask) { CK(H, cudaMemcpyAsync(w.st_pmask, pmask, (size_t)B * P, cudaMemcpyDeviceToDevice, st)); pm = w.st_pmask; }
    if (it == H->graphs.end()) {
      cudaGraph_t graph = nullptr;
      cudaGraphExec_t exec = nullptr;
      CK(H, cudaStreamSynchronize(st));
      CK(H, cudaStreamBeginCapture(H->stream, cudaStreamCaptureModeThreadLocal));
      H->capturing = true;
      int rc = 0;
      // independent utterance chains on parallel graph branches (fork / join with events inside the capture)
      int nch = H->chains < 1 ? 1 : (H->chains > STZ_MAX_CHAINS ? STZ_MAX_CHAINS : H->chains);
      if (nch > B) nch = B;
      for (int ch = 1; ch < nch && rc == 0; ++ch) {
        if (cudaEventRecord(H->fork_ev, H->stream) != cudaSuccess || cudaStreamWaitEvent(H->chain_stream[ch], H->fork_ev, 0) != cudaSuccess)
          rc = fail(H, STZ_E_CUDA, "graph fork failed");
      }
      for (int e = 0; e < E && rc == 0; ++e)
        for (int ch = 0; ch < nch && rc == 0; ++ch) {
          const int b0 = (int)((long long)B * ch / nch), b1 = (int)((long long)B * (ch + 1) / nch);
          rc = run_eval(H, ch == 0 ? H->stream : H->chain_stream[ch], e, B, b0, b1 - b0, T, P, tm, pm, nbr);
        }
      for (int ch = 1; ch < nch && rc == 0; ++ch) {
        if (cudaEventRecord(H->join_ev[ch], H->chain_stream[ch]) != cudaSuccess || cudaStreamWaitEvent(H->stream, H->join_ev[ch], 0) != cudaSuccess)
          rc = fail(H, STZ_E_CUDA, "graph join failed");
      }
      H->capturing = false;
      cudaError_t ce = cudaStreamEndCapture(H->stream, &graph);
      if (rc != 0) { if (graph) cudaGraphDestroy(graph); return rc; }
      if (ce != cudaSuccess) return fail(H, STZ_E_CUDA, "cudaStreamEndCapture -> %s", cudaGetErrorString(ce));
      ce = cudaGraphInstantiate(&exec, graph, 0);
      cudaGraphDestroy(graph);
      if (ce != cudaSuccess) return fail(H, STZ_E_CUDA, "cudaGraphInstantiate -> %s", cudaGetErrorString(ce));
      while ((int)H->graphs.size() >= H->max_graphs) {   // evict the least recently used graphs (the stream was synchronised above)
        auto victim = H->graphs.begin();
        for (auto jt = H->graphs.begin(); jt != H->graphs.end(); ++jt)
          if (jt->second.last_use < victim->second.last_use) victim = jt;
        cudaGraphExecDestroy(victim->second.exec);
        H->graphs.erase(victim);
      }
      it = H->graphs.emplace(key, stz_handle::GraphEntry{exec, H->cur_launches, 0, H->last_fuse_mode}).first;
      ++H->graph_captures;
      H->cur_launches = 0;
    }
    it->second.last_use = ++H->graph_clock;
    H->last_fuse_mode = it->second.fuse_mode;
    CK(H, cudaGraphLaunch(it->second.exec, st));
    H->launches += it->second.launches;
  } else {
    for (int e = 0; e < E; ++e) RET(run_eval(H, st, e, B, 0, B, T, P, tmask, pmask, nbr));
    H->launches += H->cur_launches;
    H->cur_launches = 0;
  }
  CK(H, cudaMemcpyAsync(out, w.x, BK * Ds * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return mark_call_end(H, st);
}

extern "C" int stz_sample_style(stz_handle* H, const float* text_emb_dev, const uint8_t* text_mask_dev,
                                const float* prompt_feats_dev, const uint8_t* prompt_mask_dev, const float* noise_dev,
                                int B, int T, int P, int steps, float cfg_scale, int sampler_kind, float* out_style_dev,
                                void* cuda_stream) {
  if (!H) return STZ_E_ARG;
  if (cudaSetDevice(H->device) != cudaSuccess) return fail(H, STZ_E_DEVICE, "cudaSetDevice failed");
  LaunchScope ls(H);
  return sample_style_impl(H, text_emb_dev, text_mask_dev, prompt_feats_dev, prompt_mask_dev, noise_dev, B, T, P, steps,
                           cfg_scale, sampler_kind, out_style_dev, (cudaStream_t)cuda_stream);
}

// ------------------------------------------------------------------------------------------
// predict_duration (a-8 .. a-11)
// ------------------------------------------------------------------------------------------
static int predict_duration_impl(stz_handle* H, const float* text, const uint8_t* tmask, const float* style, int B, int T,
                                 int32_t* out_dur, float* out_presum, cudaStream_t st) {
  const stz_config& c = H->cfg;
  if (!text || !style || !out_dur) return fail(H, STZ_E_ARG, "null tensor argument");
  if (B < 1 || T < 1) return fail(H, STZ_E_ARG, "bad sizes B=%d T=%d", B, T);
  RET(ensure_workspace(H, B, T, 1, 1, 1));
  RET(order_after_previous_call(H, st));
  Workspace& w = H->ws;
  const int ds = c.d_sty_tok, dh = c.d_hid, h = dh / 2, K = c.n_style, Ds = c.d_style;
  const int BT = B * T, BK = B * K;
  H->cur_launches = 0;
  launch_k(lens_perm_kernel, 1, 1024, 0, st, tmask, w.lens, w.perm, B, T); KCHECK(H);
  const bool tc = H->pred_tc && H->pred_gemm_impl == 0;   // split-bf16 tcgen05 GEMMs vs fp32 CUDA-core GEMMs
  const int kin = dh + ds, impl = H->gemm_impl;
  auto split_rows = [&](const float* src, int ld, int Kc, bf16* dst, int ldd, int segK, int off, size_t M) -> int {
    ProfScope ps(H, st, PC_PRED_EW, (double)M * Kc * 10.0);
    launch_k(split3_rows_kernel, ew_grid(M * (Kc / 4)), 256, 0, st, src, ld, Kc, dst, ldd, segK, off, M);
    KCHECK(H);
    return 0;
  };
  // with text masks, 128-row tiles of [B*T] that are padding throughout are skipped by the token-row GEMMs (their rows feed
  // nothing: the recurrence, AdaLN and the duration head all honour the mask)
  const uint8_t* needed = nullptr;
  if (tmask != nullptr && tc && impl == 0) {
    launch_k(tile_needed_kernel, cdiv(cdiv(BT, 128), 8), 256, 0, st, tmask, w.tile_needed, BT, 0); KCHECK(H);
    needed = w.tile_needed;
  }
  auto gemm3 = [&](const bf16* A, int K3, int rows, const bf16* W3, const float* bias, float* out, int N, bool token_rows = true) -> int {
    GemmParams p{};
    p.M = rows; p.N = N; p.K = K3; p.bias = bias; p.out = out; p.ldo = N;
    p.tile_needed = token_rows ? needed : nullptr;
    return gemm<EPI_F32>(H, st, impl, A, K3, rows, W3, p);
  };
  // a-8: per-token style summary
  auto style_pool = [&](const float* kp, const float* vp, int ldkv) -> int {
    ProfScope ps(H, st, PC_PRED_EW, (double)BT * ds * 8.0);
    if (ds <= 256 && K <= 64)   // product kernel: K / V of the utterance staged in shared memory
      launch_k(style_pool_attn2_kernel, dim3(cdiv(T, SP_TOK), B), 256, (size_t)(K * (2 * ds + 1) + 4) * 4, st, w.sq, kp, vp, ldkv, w.sa, T, K, ds, 1.0f / sqrtf(32.0f));
    else
      launch_k(style_pool_attn_kernel, cdiv(BT, 8), 256, 0, st, w.sq, kp, vp, ldkv, w.sa, BT, T, K, ds, 1.0f / sqrtf(32.0f));
    KCHECK(H);
    return 0;
  };
  if (tc) {
    RET(split_rows(text, c.d_text, c.d_text, w.pq, 3 * c.d_text, c.d_text, 0, BT));
    RET(split_rows(text, c.d_text, c.d_text, w.pa, 3 * kin, kin, 0, BT));          // x part of layer 0's [x | s_tok]
    RET(split_rows(style, Ds, Ds, w.pstyle3, 3 * Ds, Ds, 0, BK));
    RET(gemm3(w.pq, 3 * c.d_text, BT, H->wq3, W32(H, "sp.q.b"), w.sq, ds));
    RET(gemm3(w.pstyle3, 3 * Ds, BK, H->wkv3, H->b_kv, w.skv, 2 * ds, false));   // style-code rows, not token rows
    RET(style_pool(w.skv, w.skv + ds, 2 * ds));
    RET(split_rows(w.sa, ds, ds, w.psa3, 3 * ds, ds, 0, BT));
    RET(gemm3(w.psa3, 3 * ds, BT, H->wo3, W32(H, "sp.o.b"), w.stok, ds));
    RET(split_rows(w.stok, ds, ds, w.pa, 3 * kin, kin, dh, BT));                    // s_tok part, shared by all layers
    RET(split_rows(w.stok, ds, ds, w.ps3, 3 * ds, ds, 0, BT));
  } else {
    RET(linear_f32(H, st, ACT_NONE, text, c.d_text, c.d_text, nullptr, 0, 0, W32(H, "sp.q.w"), W32(H, "sp.q.b"), w.sq, ds, BT, ds));
    RET(linear_f32(H, st, ACT_NONE, style, Ds, Ds, nullptr, 0, 0, W32(H, "sp.k.w"), W32(H, "sp.k.b"), w.sk, ds, BK, ds));
    RET(linear_f32(H, st, ACT_NONE, style, Ds, Ds, nullptr, 0, 0, W32(H, "sp.v.w"), W32(H, "sp.v.b"), w.sv, ds, BK, ds));
    RET(style_pool(w.sk, w.sv, ds));
    RET(linear_f32(H, st, ACT_NONE, w.sa, ds, ds, nullptr, 0, 0, W32(H, "sp.o.w"), W32(H, "sp.o.b"), w.stok, ds, BT, ds));
  }
  // a-9: (BiLSTM + AdaLN) x (n_lstm - 1) + BiLSTM
  const float* x = text;
  float* bufs[2] = {w.xa, w.xb};
  constexpr int NB = 8;
  const size_t lstm_smem = (size_t)NB * h * 5 * sizeof(float);
  for (int l = 0; l < c.n_lstm; ++l) {
    if (l == c.n_lstm - 1) H->last_d_enc = x;   // the duration encoder's output: what the prosody heads regulate
    if (tc) {  // G = [x | s_tok] W_ih^T + (b_ih + b_hh), both directions in one GEMM (N = 8h)
      RET(gemm3(w.pa, 3 * kin, BT, H->wih3 + (size_t)l * 8 * h * 3 * kin, H->lstm_b + (size_t)l * 8 * h, w.G, 8 * h));
    } else {
      for (int dr = 0; dr < 2; ++dr) {
        const std::string p = "lstm" + std::to_string(l) + (dr ? ".r." : ".f.");
        RET(linear_f32(H, st, ACT_NONE, x, dh, dh, w.stok, ds, ds, W32(H, p + "w_ih"), H->lstm_b + ((size_t)l * 2 + dr) * 4 * h,
                       w.G + (size_t)dr * 4 * h, 8 * h, BT, 4 * h));
      }
    }
    float* xo = bufs[l & 1];
    {
      ProfScope ps(H, st, PC_LSTM, 2.0 * BT * 2.0 * h * 4.0 * h);
      if (h == LC_H && (H->lstm_impl == 0 || H->lstm_impl == 3)) {  // product path (3: W_hh in shared memory instead of tensor memory, A/B): recurrent product on tcgen05 (split-bf16), cluster of 8 CTAs, DSMEM exchange
        const float* whh = H->whh + (size_t)l * 2 * 4 * h * h;
        launch_lstm_tc(H, st, w.G, whh, w.lens, w.perm, xo, B, T);
      } else {
        lstm_rec_kernel<NB><<<dim3(cdiv(B, NB), 2), 4 * h, lstm_smem, st>>>(w.G, H->whhT + (size_t)l * 2 * h * 4 * h, w.lens, xo, B, T, h);
      }
      KCHECK(H);
    }
    if (l < c.n_lstm - 1) {
      const std::string p = "adaln" + std::to_string(l) + ".";
      if (tc) RET(gemm3(w.ps3, 3 * ds, BT, H->wada3 + (size_t)l * 2 * dh * 3 * ds, W32(H, p + "b"), w.gb, 2 * dh));
      else RET(linear_f32(H, st, ACT_NONE, w.stok, ds, ds, nullptr, 0, 0, W32(H, p + "w"), W32(H, p + "b"), w.gb, 2 * dh, BT, 2 * dh));
      dim3 grid(cdiv(BT, 8));
      bf16* a3 = tc ? w.pa : nullptr;
      ProfScope ps(H, st, PC_PRED_EW, (double)BT * dh * 16.0);
      switch (dh / 128) {
        case 1: launch_k(adaln_pred_kernel<1>, grid, 256, 0, st, xo, w.gb, tmask, BT, a3, 3 * kin, kin); break;
        case 2: launch_k(adaln_pred_kernel<2>, grid, 256, 0, st, xo, w.gb, tmask, BT, a3, 3 * kin, kin); break;
        case 4: launch_k(adaln_pred_kernel<4>, grid, 256, 0, st, xo, w.gb, tmask, BT, a3, 3 * kin, kin); break;
        case 8: launch_k(adaln_pred_kernel<8>, grid, 256, 0, st, xo, w.gb, tmask, BT, a3, 3 * kin, kin); break;
        default: return fail(H, STZ_E_SHAPE, "d_hid %d unsupported", dh);
      }
      KCHECK(H);
    }
    x = xo;
  }
  {  // a-10
    dim3 grid(cdiv(BT, 8));
    const dim3 grid2(cdiv(BT, 32));
    ProfScope ps(H, st, PC_PRED_EW, (double)BT * dh * 4.0);
    switch (dh / 128) {
      case 1: launch_k(dur_head2_kernel<1>, grid2, 256, dur_head2_smem(1), st, x, W32(H, "dur.w"), W32(H, "dur.b"), tmask, out_dur, out_presum, BT, c.max_dur); break;
      case 2: launch_k(dur_head2_kernel<2>, grid2, 256, dur_head2_smem(2), st, x, W32(H, "dur.w"), W32(H, "dur.b"), tmask, out_dur, out_presum, BT, c.max_dur); break;
      case 4: launch_k(dur_head2_kernel<4>, grid2, 256, dur_head2_smem(4), st, x, W32(H, "dur.w"), W32(H, "dur.b"), tmask, out_dur, out_presum, BT, c.max_dur); break;
      case 8: launch_k(dur_head_kernel<8>, grid, 256, 0, st, x, W32(H, "dur.w"), W32(H, "dur.b"), tmask, out_dur, out_presum, BT, c.max_dur); break;
      default: return fail(H, STZ_E_SHAPE, "d_hid %d unsupported", dh);
    }
    KCHECK(H);
  }
  H->launches += H->cur_launches;
  H->cur_launches = 0;
  return mark_call_end(H, st);
}

extern "C" int stz_predict_duration(stz_handle* H, const float* text_emb_dev, const uint8_t* text_mask_dev,
                                    const float* style_dev, int B, int T, int32_t* out_dur_dev, float* out_presum_dev,
                                    void* cuda_stream) {
  if (!H) return STZ_E_ARG;
  if (cudaSetDevice(H->device) != cudaSuccess) return fail(H, STZ_E_DEVICE, "cudaSetDevice failed");
  LaunchScope ls(H);
  return predict_duration_impl(H, text_emb_dev, text_mask_dev, style_dev, B, T, out_dur_dev, out_presum_dev,
                               (cudaStream_t)cuda_stream);
}

// ------------------------------------------------------------------------------------------
// length regulator: the step right after predict_duration
// ------------------------------------------------------------------------------------------
extern "C" int stz_regulate_length(stz_handle* H, const float* feats_dev, const int32_t* dur_dev, int B, int T, int C, int F_max,
                                   float* out_frames_dev, int32_t* out_frame_lens_dev, int32_t* out_frame_tok_dev,
                                   void* cuda_stream) {
  if (!H) return STZ_E_ARG;
  if (!feats_dev || !dur_dev || !out_frames_dev || !out_frame_lens_dev) return fail(H, STZ_E_ARG, "null tensor argument");
  if (B < 1 || T < 1 || F_max < 1) return fail(H, STZ_E_ARG, "bad sizes B=%d T=%d F_max=%d", B, T, F_max);
  if (T > LR_MAX_T || C < 4 || C % 4) return fail(H, STZ_E_SHAPE, "length regulator supports T <= %d, C %% 4 == 0", LR_MAX_T);
  if (cudaSetDevice(H->device) != cudaSuccess) return fail(H, STZ_E_DEVICE, "cudaSetDevice failed");
  cudaStream_t st = (cudaStream_t)cuda_stream;
  LaunchScope ls(H);
  H->cur_launches = 0;
  launch_k(length_regulate_kernel, dim3(cdiv(F_max, LR_FRAMES), B), LR_THREADS, 0, st, feats_dev, dur_dev, out_frames_dev,
           out_frame_lens_dev, out_frame_tok_dev, T, C, F_max, (const float*)nullptr, 0);
  KCHECK(H);
  H->launches += H->cur_launches;
  H->cur_launches = 0;
  return 0;
}

// ------------------------------------------------------------------------------------------
// prosody heads (second half): F0 and energy curves over the length-regulated frames
// ------------------------------------------------------------------------------------------
static int ensure_prosody_ws(stz_handle* H, int B, int F, int T) {
  stz_handle::ProsodyWs& w = H->pws;
  if (w.base && B <= w.B && F <= w.F && T <= w.T) return 0;
  if (w.base) { CK(H, cudaDeviceSynchronize()); CK(H, cudaFree(w.base)); w.base = nullptr; }
  B = B > w.B ? B : w.B; F = F > w.F ? F : w.F; T = T > w.T ? T : w.T;
  const stz_config& c = H->cfg;
  const size_t BF = (size_t)B * F, kin = c.d_hid + c.d_sty_tok, h8 = 4 * (size_t)c.d_hid;
  size_t off = 0;
  std::vector<std::pair<void**, size_t>> plan;
  H->guards_pros.clear();
  auto want = [&](void** p, size_t bytes) {
    plan.push_back({p, off});
    const size_t end = off + bytes;
    off = align_up(end + (size_t)H->guard_bytes, 1024);
    if (H->guard_bytes > 0) H->guards_pros.push_back({end, off - end});
  };
  want((void**)&w.frames, BF * kin * sizeof(float));
  want((void**)&w.pa, (BF + 128) * 3 * kin * sizeof(bf16));
  want((void**)&w.G, BF * h8 * sizeof(float));
  want((void**)&w.y, BF * c.d_hid * sizeof(float));
  want((void**)&w.flens, (size_t)B * sizeof(int)); want((void**)&w.perm, (size_t)B * sizeof(int));
  want((void**)&w.dur, (size_t)B * T * sizeof(int));
  want((void**)&w.needed, BF / 128 + 2);
  cudaError_t e = cudaMalloc(&w.base, off);
  if (e != cudaSuccess) {
    w = stz_handle::ProsodyWs();
    return fail(H, STZ_E_NOMEM, "prosody workspace of %zu bytes: %s", off, cudaGetErrorString(e));
  }
  for (auto& pr : plan) *pr.first = w.base + pr.second;
  if (H->guard_bytes > 0) { CK(H, cudaMemset(w.base, STZ_GUARD_BYTE, off)); }
  w.bytes = off; w.B = B; w.F = F; w.T = T;
  return 0;
}

extern "C" int stz_predict_prosody(stz_handle* H, const float* text_emb_dev, const uint8_t* text_mask_dev, const float* style_dev,
                                   const int32_t* dur_in_dev, int B, int T, int F_max, float* out_f0_dev, float* out_energy_dev,
                                   int32_t* out_frame_lens_dev, int32_t* out_dur_dev, void* cuda_stream) {
  if (!H) return STZ_E_ARG;
  if (!text_emb_dev || !style_dev || !out_f0_dev || !out_energy_dev || !out_frame_lens_dev) return fail(H, STZ_E_ARG, "null tensor argument");
  if (B < 1 || T < 1 || F_max < 1) return fail(H, STZ_E_ARG, "bad sizes B=%d T=%d F_max=%d", B, T, F_max);
  const stz_config& c = H->cfg;
  const int dh = c.d_hid, ds = c.d_sty_tok, h = dh / 2, kin = dh + ds, dp = dh / 2;
  if (!H->pred_tc || h != LC_H || T > LR_MAX_T || dp % 128 || dp > 512)
    return fail(H, STZ_E_SHAPE, "prosody heads need the tensor-core predictor configuration (d_hid 512, d_sty_tok %% 128 == 0) and T <= %d", LR_MAX_T);
  if (cudaSetDevice(H->device) != cudaSuccess) return fail(H, STZ_E_DEVICE, "cudaSetDevice failed");
  LaunchScope ls(H);
  cudaStream_t st = (cudaStream_t)cuda_stream;
  RET(ensure_prosody_ws(H, B, F_max, T));
  stz_handle::ProsodyWs& w = H->pws;
  // the duration predictor's forward: durations (unless given) and, in the workspace, d_enc and s_tok
  int32_t* dur_pred = out_dur_dev ? out_dur_dev : w.dur;
  RET(predict_duration_impl(H, text_emb_dev, text_mask_dev, style_dev, B, T, dur_pred, nullptr, st));
  const int32_t* dur = dur_in_dev ? dur_in_dev : dur_pred;
  const size_t BF = (size_t)B * F_max;
  H->cur_launches = 0;
  // frames = regulate([d_enc | s_tok], dur)
  launch_k(length_regulate_kernel, dim3(cdiv(F_max, LR_FRAMES), B), LR_THREADS, 0, st, H->last_d_enc, dur, w.frames, w.flens,
           (int32_t*)nullptr, T, dh, F_max, (const float*)H->ws.stok, ds); KCHECK(H);
  launch_k(perm_from_lens_kernel, 1, 1024, 0, st, (const int*)w.flens, w.perm, B); KCHECK(H);
  launch_k(frame_tile_needed_kernel, cdiv(cdiv((long long)BF, 128), 256), 256, 0, st, (const int*)w.flens, w.needed, B, F_max); KCHECK(H);
  CK(H, cudaMemcpyAsync(out_frame_lens_dev, w.flens, (size_t)B * sizeof(int), cudaMemcpyDeviceToDevice, st));
  auto split_rows = [&](const float* src, int ld, int Kc, int off) -> int {
    ProfScope ps(H, st, PC_PRED_EW, (double)BF * Kc * 10.0);
    launch_k(split3_rows_kernel, ew_grid(BF * (Kc / 4)), 256, 0, st, src, ld, Kc, w.pa, 3 * kin, kin, off, BF);
    KCHECK(H);
    return 0;
  };
  auto gemm3 = [&](const bf16* W3, const float* bias, float* out, int N) -> int {
    GemmParams p{};
    p.M = (int)BF; p.N = N; p.K = 3 * kin; p.bias = bias; p.out = out; p.ldo = N; p.tile_needed = w.needed;
    return gemm<EPI_F32>(H, st, H->gemm_impl, w.pa, 3 * kin, (int)BF, W3, p);
  };
  // shared BiLSTM over the frames (packed by frame count)
  RET(split_rows(w.frames, kin, kin, 0));
  RET(gemm3(H->wih3_pros, H->lstm_b_pros, w.G, 8 * h));
  {
    ProfScope ps(H, st, PC_LSTM, 2.0 * BF * 2.0 * h * 4.0 * h);
    launch_lstm_tc(H, st, w.G, H->whh_pros, w.flens, w.perm, w.y, B, F_max);
    KCHECK(H);
  }
  // z = [y | s_frame] W_h1^T + b_h1 (the s_frame columns of the operand are still in place), then the two heads
  RET(split_rows(w.y, dh, dh, 0));
  RET(gemm3(H->wh1_3, W32(H, "pros.h1.b"), w.G, dh));
  {
    ProfScope ps(H, st, PC_PRED_EW, (double)BF * (dh * 4.0 + 8.0));
    const dim3 grid(cdiv((long long)BF, 8));
    switch (dp / 128) {
      case 1: launch_k(prosody_head_kernel<1>, grid, 256, 0, st, (const float*)w.G, W32(H, "pros.f0.w"), W32(H, "pros.f0.b"), W32(H, "pros.en.w"), W32(H, "pros.en.b"), (const int*)w.flens, out_f0_dev, out_energy_dev, B, F_max); break;
      case 2: launch_k(prosody_head_kernel<2>, grid, 256, 0, st, (const float*)w.G, W32(H, "pros.f0.w"), W32(H, "pros.f0.b"), W32(H, "pros.en.w"), W32(H, "pros.en.b"), (const int*)w.flens, out_f0_dev, out_energy_dev, B, F_max); break;
      case 4: launch_k(prosody_head_kernel<4>, grid, 256, 0, st, (const float*)w.G, W32(H, "pros.f0.w"), W32(H, "pros.f0.b"), W32(H, "pros.en.w"), W32(H, "pros.en.b"), (const int*)w.flens, out_f0_dev, out_energy_dev, B, F_max); break;
      default: return fail(H, STZ_E_SHAPE, "d_hid %d unsupported by the prosody heads", dh);
    }
    KCHECK(H);
  }
  H->launches += H->cur_launches;
  H->cur_launches = 0;
  return mark_call_end(H, st);
}

// ------------------------------------------------------------------------------------------
// on-device noise
// ------------------------------------------------------------------------------------------
extern "C" int stz_set_noise_seed(stz_handle* H, uint64_t seed, uint64_t first_utterance) {
  if (!H) return STZ_E_ARG;
  H->noise_seeded = true;
  H->noise_seed = seed;
  H->noise_first_utt = first_utterance;
  H->noise_ids.clear();
  return 0;
}

extern "C" int stz_set_noise_utterances(stz_handle* H, uint64_t seed, const uint64_t* utterance_ids, int n) {
  if (!H) return STZ_E_ARG;
  if (!utterance_ids || n < 1) return fail(H, STZ_E_ARG, "utterance_ids must hold n >= 1 indices");
  H->noise_seeded = true;
  H->noise_seed = seed;
  H->noise_first_utt = 0;
  H->noise_ids.assign(utterance_ids, utterance_ids + n);
  return 0;
}

extern "C" int stz_philox_normal(uint64_t seed, uint64_t first_utterance, int slices, int B, int n_per_utt, float* out_dev,
                                 int device, void* cuda_stream) {
  if (!out_dev) return fail(nullptr, STZ_E_ARG, "null argument");
  if (slices < 1 || B < 1 || n_per_utt < 4 || n_per_utt % 4) return fail(nullptr, STZ_E_SHAPE, "slices, B >= 1 and n_per_utt %% 4 == 0 required");
  if (cudaSetDevice(device) != cudaSuccess) return fail(nullptr, STZ_E_DEVICE, "cudaSetDevice(%d) failed", device);
  const size_t groups = (size_t)slices * B * (n_per_utt / 4);
  const size_t blocks = (groups + 255) / 256;
  philox_normal_kernel<<<(unsigned)(blocks < 148 * 16 ? blocks : 148 * 16), 256, 0, (cudaStream_t)cuda_stream>>>(
      out_dev, (uint32_t)seed, (uint32_t)(seed >> 32), (unsigned long long)first_utterance, nullptr, slices, B, n_per_utt / 4);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(nullptr, STZ_E_CUDA, "philox_normal_kernel -> %s", cudaGetErrorString(e));
  return 0;
}

// ------------------------------------------------------------------------------------------
// host-buffer end-to-end entry point
// ------------------------------------------------------------------------------------------
static int host_wait_slot(stz_handle* H, int slot) {
  if (!H->slot[slot].busy) return 0;
  H->slot[slot].busy = false;
  CK(H, cudaEventSynchronize(H->slot[slot].ev_done));
  return 0;
}

// Two-slot pipeline: the H2D copies run on copy_stream, compute on the internal stream, the D2H copies on out_stream, and
// every slot has its own device staging — so between _submit(slot) and _wait(slot) the caller may _submit the other slot,
// whose input copies then overlap this slot's compute.  Host buffers of a slot stay owned by the library until its _wait.
extern "C" int stz_synthesize_host_submit(stz_handle* H, int slot, const float* text_emb, const uint8_t* text_mask,
                                          const float* prompt_feats, const uint8_t* prompt_mask, const float* noise, int B,
                                          int T, int P, int steps, float cfg_scale, int sampler_kind, float* out_style,
                                          int32_t* out_dur) {
  if (!H) return STZ_E_ARG;
  if (slot < 0 || slot > 1) return fail(H, STZ_E_ARG, "slot must be 0 or 1");
  if (!text_emb || !prompt_feats || !out_style) return fail(H, STZ_E_ARG, "null tensor argument");
  if (!noise && !H->noise_seeded) return fail(H, STZ_E_ARG, "noise is NULL and no seed was set (stz_set_noise_seed)");
  if (B < 1 || T < 1 || P < 1 || steps < 1) return fail(H, STZ_E_ARG, "bad sizes");
  if (cudaSetDevice(H->device) != cudaSuccess) return fail(H, STZ_E_DEVICE, "cudaSetDevice failed");
  const stz_config& c = H->cfg;
  LaunchScope ls(H);
  const int E = sampler_kind == STZ_SAMPLER_TEACHER ? 2 * steps : steps;
  const int slices = sampler_kind == STZ_SAMPLER_TEACHER ? steps + 1 : 1;
  RET(host_wait_slot(H, slot));                       // resubmitting a slot first retires its previous call
  {
    const Workspace& w0 = H->ws;
    const size_t mod_rows_req = (size_t)(hoist_mod(c, B, E) ? E : 1) * 2 * B;
    const bool grows = !w0.base || B > w0.B || effective_T(H, T) > w0.T || P > w0.P || E > w0.E || slices > w0.noise_slices || mod_rows_req > w0.mod_rows;
    if (grows) RET(host_wait_slot(H, slot ^ 1));      // the arena is reallocated: nothing may be in flight
  }
  RET(ensure_workspace(H, B, effective_T(H, T), P, E, slices));
  Workspace& w = H->ws;
  Workspace::HostStage& hs = w.hs[slot];
  stz_handle::HostSlot& sl = H->slot[slot];
  cudaStream_t st = H->stream, cs = H->copy_stream, os = H->out_stream;
  const size_t BK = (size_t)B * c.n_style, BT = (size_t)B * T, BP = (size_t)B * P;
  // copy stream: masks + text (gate the whole call), then prompt and noise (consumed after the text-side prep / at state
  // initialisation: sample_style_impl waits on their events)
  uint8_t *tm = nullptr, *pm = nullptr;
  if (text_mask) { tm = hs.tmask; CK(H, cudaMemcpyAsync(tm, text_mask, BT, cudaMemcpyHostToDevice, cs)); }
  if (prompt_mask) { pm = hs.pmask; CK(H, cudaMemcpyAsync(pm, prompt_mask, BP, cudaMemcpyHostToDevice, cs)); }
  CK(H, cudaMemcpyAsync(hs.text, text_emb, BT * c.d_text * sizeof(float), cudaMemcpyHostToDevice, cs));
  CK(H, cudaEventRecord(sl.ev_text, cs));
  CK(H, cudaMemcpyAsync(hs.prompt, prompt_feats, BP * c.d_prompt * sizeof(float), cudaMemcpyHostToDevice, cs));
  CK(H, cudaEventRecord(sl.ev_prompt, cs));
  if (noise) {
    CK(H, cudaMemcpyAsync(hs.noise, noise, (size_t)slices * BK * c.d_style * sizeof(float), cudaMemcpyHostToDevice, cs));
    CK(H, cudaEventRecord(sl.ev_noise, cs));
  }
  sl.busy = true;
  auto bail = [&](int rc) { cudaStreamSynchronize(cs); cudaStreamSynchronize(st); cudaStreamSynchronize(os); sl.busy = false; return rc; };
  if (cudaStreamWaitEvent(st, sl.ev_text, 0) != cudaSuccess) return bail(fail(H, STZ_E_CUDA, "cudaStreamWaitEvent failed"));
  H->cur_ev_prompt = sl.ev_prompt; H->cur_ev_noise = sl.ev_noise;
  H->wait_prompt = true;
  H->wait_noise = noise != nullptr;
  int rc = sample_style_impl(H, hs.text, tm, hs.prompt, pm, noise ? hs.noise : nullptr, B, T, P, steps, cfg_scale, sampler_kind, hs.style, st);
  H->wait_prompt = H->wait_noise = false;
  if (rc != 0) return bail(rc);
  // style D2H overlaps the duration predictor
  if (cudaEventRecord(sl.ev_style, st) != cudaSuccess || cudaStreamWaitEvent(os, sl.ev_style, 0) != cudaSuccess ||
      cudaMemcpyAsync(out_style, hs.style, BK * c.d_style * sizeof(float), cudaMemcpyDeviceToHost, os) != cudaSuccess)
    return bail(fail(H, STZ_E_CUDA, "style D2H failed"));
  if (out_dur) {
    rc = predict_duration_impl(H, hs.text, tm, hs.style, B, T, hs.dur, nullptr, st);
    if (rc != 0) return bail(rc);
    if (cudaEventRecord(sl.ev_dur, st) != cudaSuccess || cudaStreamWaitEvent(os, sl.ev_dur, 0) != cudaSuccess ||
        cudaMemcpyAsync(out_dur, hs.dur, BT * sizeof(int32_t), cudaMemcpyDeviceToHost, os) != cudaSuccess)
      return bail(fail(H, STZ_E_CUDA, "duration D2H failed"));
  }
  if (cudaEventRecord(sl.ev_done, os) != cudaSuccess) return bail(fail(H, STZ_E_CUDA, "cudaEventRecord failed"));
  return 0;
}

extern "C" int stz_synthesize_host_wait(stz_handle* H, int slot) {
  if (!H) return STZ_E_ARG;
  if (slot < 0 || slot > 1) return fail(H, STZ_E_ARG, "slot must be 0 or 1");
  if (cudaSetDevice(H->device) != cudaSuccess) return fail(H, STZ_E_DEVICE, "cudaSetDevice failed");
  return host_wait_slot(H, slot);
}

extern "C" int stz_synthesize_host(stz_handle* H, const float* text_emb, const uint8_t* text_mask,
                                   const float* prompt_feats, const uint8_t* prompt_mask, const float* noise, int B, int T,
                                   int P, int steps, float cfg_scale, int sampler_kind, float* out_style, int32_t* out_dur) {
  if (!H) return STZ_E_ARG;
  RET(stz_synthesize_host_submit(H, 0, text_emb, text_mask, prompt_feats, prompt_mask, noise, B, T, P, steps, cfg_scale,
                                 sampler_kind, out_style, out_dur));
  return stz_synthesize_host_wait(H, 0);
}

// ------------------------------------------------------------------------------------------
// unit-test entry point
// ------------------------------------------------------------------------------------------
extern "C" int stz_op_gemm_bf16(const void* A, const void* W, const float* bias, float* C, int M, int N, int K, int impl,
                                int device, void* cuda_stream) {
  if (!A || !W || !C) return fail(nullptr, STZ_E_ARG, "null argument");
  if (cudaSetDevice(device) != cudaSuccess) return fail(nullptr, STZ_E_DEVICE, "cudaSetDevice(%d) failed", device);
  if (load_encode()) return fail(nullptr, STZ_E_DEVICE, "cuTensorMapEncodeTiled entry point not found");
  if (N % GEMM_BN || K % GEMM_BK || M < 1) return fail(nullptr, STZ_E_SHAPE, "N %% 128, K %% 64 required");
  if (init_kernel_attrs() != cudaSuccess) return fail(nullptr, STZ_E_CUDA, "cudaFuncSetAttribute failed");
  LaunchScope ls(nullptr);
  GemmParams p{};
  p.M = M; p.N = N; p.K = K; p.bias = bias; p.out = C; p.ldo = N;
  cudaStream_t st = (cudaStream_t)cuda_stream;
  int rc = impl == 0 ? launch_gemm2<EPI_F32>(nullptr, st, (const bf16*)A, K, M, (const bf16*)W, p)
                     : launch_gemm_simt<EPI_F32>(nullptr, st, (const bf16*)A, K, (const bf16*)W, p);
  if (rc != 0 && g_create_error.empty()) g_create_error = "gemm launch failed";
  return rc;
}

// Unit-test entry for the fused attention kernels.  Self-attention (kv_text == NULL): qkv [2*B*K, 3d] bf16 in R layout.
// Cross-attention: q = first d columns of `qkv` with row stride ldq, keys = [text ; prompt | null] with one layer's
// K | V per row: kv_text [B*T, 2d], kv_prompt [B*P, 2d], kv_null [1, 2d].  out [2*B*K, d] bf16.
// impl: 0 tcgen05 + TMA, 1 mma.sync resident keys, 2 mma.sync streaming, 3 tcgen05 + cp.async.
extern "C" int stz_op_attention(stz_handle* H, const void* qkv, int ldq, const void* kv_text, const void* kv_prompt,
                                const void* kv_null, const uint8_t* tmask, const uint8_t* pmask, int B, int T, int P,
                                void* out, int impl, void* cuda_stream) {
  if (!H || !qkv || !out || B < 1) return STZ_E_ARG;
  if (cudaSetDevice(H->device) != cudaSuccess) return fail(H, STZ_E_DEVICE, "cudaSetDevice failed");
  const stz_config& c = H->cfg;
  LaunchScope ls(H);
  const int d = c.d_model, K = c.n_style;
  const bf16* q = (const bf16*)qkv;
  AttnParams ap{};
  ap.q = q; ap.ldq = ldq; ap.out = (bf16*)out; ap.ldo = d; ap.n_q = 2 * K;
  ap.scale_log2 = 1.4426950408889634f / sqrtf((float)(d / c.n_heads));
  if (!kv_text) {
    ap.nseg = 1;
    ap.seg[0] = AttnSeg{q + d, q + 2 * d, ldq, 2 * K, 2 * K, nullptr, KEY_SAME_BRANCH};
  } else {
    if (!kv_prompt || !kv_null || T < 1 || P < 1) return fail(H, STZ_E_ARG, "cross-attention needs text, prompt and null K/V");
    const bf16 *kt = (const bf16*)kv_text, *kp = (const bf16*)kv_prompt, *kn = (const bf16*)kv_null;
    ap.nseg = 3;
    ap.seg[0] = AttnSeg{kt, kt + d, 2 * d, T, T, tmask, KEY_ALL};
    ap.seg[1] = AttnSeg{kp, kp + d, 2 * d, P, P, pmask, KEY_COND};
    ap.seg[2] = AttnSeg{kn, kn + d, 2 * d, 1, 0, nullptr, KEY_UNCOND};
  }
  const int saved = H->attn_impl;
  H->attn_impl = impl;
  const int rc = attention(H, (cudaStream_t)cuda_stream, ap, B);
  H->attn_impl = saved;
  H->cur_launches = 0;
  return rc;
}

// ------------------------------------------------------------------------------------------
// bench.py roofline leg: one GEMM shape of the denoiser, `iters` back-to-back launches on the handle's
// stream between two CUDA events (PDL-chained exactly like the evaluation loop), average microseconds.
// epi: 2 = bf16 out, 3 = GELU bf16 out, 4 = gated residual (TMA reduce-add into an fp32 buffer).
// ------------------------------------------------------------------------------------------
extern "C" int stz_bench_gemm(stz_handle* H, int M, int N, int K, int epi, int iters, double* avg_us) {
  if (!H || !avg_us || M < 1 || iters < 1) return STZ_E_ARG;
  if (N % 128 || K % 64) return fail(H, STZ_E_SHAPE, "N %% 128, K %% 64 required");
  if (cudaSetDevice(H->device) != cudaSuccess) return fail(H, STZ_E_DEVICE, "cudaSetDevice failed");
  cudaStream_t st = H->stream;
  LaunchScope ls(H);
  bf16 *A = nullptr, *W = nullptr, *Cb = nullptr;
  float *Cf = nullptr, *bias = nullptr, *mod = nullptr;
  const int rpu = 2 * H->cfg.n_style, n_seq = 2 * cdiv(M, rpu) + 2;
  CK(H, cudaMalloc(&A, (size_t)(M + 128) * K * sizeof(bf16)));
  CK(H, cudaMalloc(&W, (size_t)N * K * sizeof(bf16)));
  CK(H, cudaMalloc(&Cb, (size_t)M * N * sizeof(bf16)));
  CK(H, cudaMalloc(&Cf, (size_t)M * N * sizeof(float)));
  CK(H, cudaMalloc(&bias, (size_t)N * sizeof(float)));
  CK(H, cudaMalloc(&mod, (size_t)n_seq * 3 * N * sizeof(float)));
  // random operands (a zero-filled problem flatters the tensor pipe: no operand toggling, lower power); the gate of the
  // residual forms is kept small so that hundreds of accumulating launches stay finite
  fill_random_bf16_kernel<<<ew_grid((size_t)(M + 128) * K), 256, 0, st>>>(A, (size_t)(M + 128) * K, 1.0f, 1u);
  fill_random_bf16_kernel<<<ew_grid((size_t)N * K), 256, 0, st>>>(W, (size_t)N * K, 1.0f / sqrtf((float)K), 2u);
  fill_random_f32_kernel<<<ew_grid((size_t)M * N), 256, 0, st>>>(Cf, (size_t)M * N, 1.0f, 3u);
  fill_random_f32_kernel<<<ew_grid((size_t)N), 256, 0, st>>>(bias, (size_t)N, 0.1f, 4u);
  fill_random_f32_kernel<<<ew_grid((size_t)n_seq * 3 * N), 256, 0, st>>>(mod, (size_t)n_seq * 3 * N, 1e-3f, 5u);
  CK(H, cudaGetLastError());
  GemmParams p{};
  p.M = M; p.N = N; p.K = K; p.bias = bias; p.ldo = N; p.mod = mod; p.n_mod = N; p.gate_off = 0; p.rows_per_utt = rpu;
  p.n_style = H->cfg.n_style;
  auto once = [&]() -> int {
    switch (epi) {
      case 2: p.out = Cb; return gemm<EPI_BF16>(H, st, 0, A, K, M, W, p);
      case 3: p.out = Cb; return gemm<EPI_GELU_BF16>(H, st, 0, A, K, M, W, p);
      case 4: p.out = Cf; return gemm<EPI_GATE_RES>(H, st, 0, A, K, M, W, p);
      case 6: {   // the product form of the residual GEMMs: GEMM + gated residual + AdaLN (gemmln3_kernel), N = d_model
        if (N != GLN_N) return fail(H, STZ_E_SHAPE, "fused residual GEMM needs N = %d", GLN_N);
        GemmLnParams q{};
        q.M = M; q.K = K; q.bias = bias; q.h = Cf; q.mod = mod; q.n_mod = 3 * N; q.gate_off = 0; q.shift_off = N; q.scale_off = 2 * N;
        q.rows_per_utt = rpu; q.pos = nullptr; q.n_style = H->cfg.n_style; q.split3 = 0;
        return launch_gemmln<GLN_RES>(H, st, A, K, M, W, Cb, q);
      }
      default: return fail(H, STZ_E_ARG, "epi %d not benchable", epi);
    }
  };
  const int saved_profile = H->profile;
  H->profile = 0;
  int rc = 0;
  for (int i = 0; i < 3 && rc == 0; ++i) rc = once();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0, st);
  for (int i = 0; i < iters && rc == 0; ++i) rc = once();
  cudaEventRecord(e1, st);
  cudaError_t ce = cudaStreamSynchronize(st);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(A); cudaFree(W); cudaFree(Cb); cudaFree(Cf); cudaFree(bias); cudaFree(mod);
  H->profile = saved_profile;
  H->cur_launches = 0;
  if (rc != 0) return rc;
  if (ce != cudaSuccess) return fail(H, STZ_E_CUDA, "bench gemm -> %s", cudaGetErrorString(ce));
  *avg_us = (double)ms * 1e3 / iters;
  return 0;
}

// ------------------------------------------------------------------------------------------
// unit-test entry points for the fused epilogues of the product GEMM kernels (tests/test_gpu_kernels.py compare each
// with a plain PyTorch fp32 restatement of the same op)
// ------------------------------------------------------------------------------------------
extern "C" int stz_op_gemm_epi(stz_handle* H, const void* A, const void* W, const float* bias, int M, int N, int K, int epi,
                               void* out, const float* mod, int n_mod, int gate_off, const float* pos, void* cuda_stream) {
  if (!H || !A || !W || !out) return STZ_E_ARG;
  if (cudaSetDevice(H->device) != cudaSuccess) return fail(H, STZ_E_DEVICE, "cudaSetDevice failed");
  LaunchScope ls(H);
  cudaStream_t st = (cudaStream_t)cuda_stream;
  GemmParams p{};
  p.M = M; p.N = N; p.K = K; p.bias = bias; p.out = out; p.ldo = N; p.mod = mod; p.n_mod = n_mod; p.gate_off = gate_off;
  p.rows_per_utt = 2 * H->cfg.n_style; p.n_style = H->cfg.n_style; p.pos = pos;
  int rc;
  switch (epi) {
    case EPI_F32: rc = gemm<EPI_F32>(H, st, 0, (const bf16*)A, K, M, (const bf16*)W, p); break;
    case EPI_F32_POS:
      if (!pos) return fail(H, STZ_E_ARG, "EPI_F32_POS needs pos");
      rc = gemm<EPI_F32_POS>(H, st, 0, (const bf16*)A, K, M, (const bf16*)W, p); break;
    case EPI_BF16: rc = gemm<EPI_BF16>(H, st, 0, (const bf16*)A, K, M, (const bf16*)W, p); break;
    case EPI_GELU_BF16: rc = gemm<EPI_GELU_BF16>(H, st, 0, (const bf16*)A, K, M, (const bf16*)W, p); break;
    case EPI_GATE_RES:
      if (!mod) return fail(H, STZ_E_ARG, "EPI_GATE_RES needs mod");
      rc = gemm<EPI_GATE_RES>(H, st, 0, (const bf16*)A, K, M, (const bf16*)W, p); break;
    default: return fail(H, STZ_E_ARG, "epi %d is not a plain epilogue (see stz_op_gemm_sampler / stz_op_gemm_ln)", epi);
  }
  H->cur_launches = 0;
  return rc;
}

extern "C" int stz_op_gemm_sampler(stz_handle* H, const void* A, const void* W, const float* bias, int M, int N, int K,
                                   float* x, float* xmid, const float* noise, const float* coef_dev, void* xin_out,
                                   float* tap, void* cuda_stream) {
  if (!H || !A || !W || !x || !xmid || !noise || !coef_dev || !xin_out) return STZ_E_ARG;
  if (M % 2) return fail(H, STZ_E_SHAPE, "the sampler epilogue pairs rows: M must be even");
  if (cudaSetDevice(H->device) != cudaSuccess) return fail(H, STZ_E_DEVICE, "cudaSetDevice failed");
  LaunchScope ls(H);
  GemmParams p{};
  p.M = M; p.N = N; p.K = K; p.bias = bias; p.ldo = N; p.rows_per_utt = 2 * H->cfg.n_style; p.n_style = H->cfg.n_style;
  p.x = x; p.xmid = xmid; p.noise = noise; p.coef = coef_dev; p.xin = (bf16*)xin_out; p.tap = tap;
  const int rc = gemm<EPI_SAMPLER>(H, (cudaStream_t)cuda_stream, 0, (const bf16*)A, K, M, (const bf16*)W, p);
  H->cur_launches = 0;
  return rc;
}

extern "C" int stz_op_gemm_ln(stz_handle* H, const void* A, const void* W, const float* bias, int M, int K, int mode,
                              float* h, const float* mod, int n_mod, int gate_off, int shift_off, int scale_off,
                              const float* pos, int split3, void* u_out, void* cuda_stream) {
  if (!H || !A || !W || !bias || !h || !mod || !u_out) return STZ_E_ARG;
  if (H->cfg.d_model != GLN_N) return fail(H, STZ_E_SHAPE, "the fused GEMM + AdaLN kernel needs d_model = %d", GLN_N);
  if (mode == GLN_POS && !pos) return fail(H, STZ_E_ARG, "GLN_POS needs pos");
  if (cudaSetDevice(H->device) != cudaSuccess) return fail(H, STZ_E_DEVICE, "cudaSetDevice failed");
  LaunchScope ls(H);
  GemmLnParams p{};
  p.M = M; p.K = K; p.bias = bias; p.h = h; p.mod = mod; p.n_mod = n_mod; p.gate_off = gate_off; p.shift_off = shift_off;
  p.scale_off = scale_off; p.rows_per_utt = 2 * H->cfg.n_style; p.pos = pos; p.n_style = H->cfg.n_style; p.split3 = split3 ? 1 : 0;
  cudaStream_t st = (cudaStream_t)cuda_stream;
  int rc;
  if (mode == GLN_RES) rc = launch_gemmln<GLN_RES>(H, st, (const bf16*)A, K, M, (const bf16*)W, (bf16*)u_out, p);
  else if (mode == GLN_POS) rc = launch_gemmln<GLN_POS>(H, st, (const bf16*)A, K, M, (const bf16*)W, (bf16*)u_out, p);
  else return fail(H, STZ_E_ARG, "mode must be 0 (residual) or 1 (positional)");
  H->cur_launches = 0;
  return rc;
}
