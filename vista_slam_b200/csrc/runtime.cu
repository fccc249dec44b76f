// STA model runtime: packed weights, workspace and the kernel sequence of the forward pass,
// exported through the C ABI declared in include/sta_b200.h.
//
// Mirrors (behaviour, not code) vista_slam/sta_model/sta_model.py:
//   _encode_image :163-174, _decode_stereo :177-244, forward :247-291,
//   DPT head heads/dpt_head.py:34-66 + heads/dpt_block.py, pose head heads/pose_head.py:109-119.
#include <stdint.h>
#include <string.h>

#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/sta_b200.h"
#include "host_util.h"
#include "ops.h"

namespace sta {

// ---------------------------------------------------------------------------
// weight packing kernels (run once per tensor at load time)
// ---------------------------------------------------------------------------
// dst[r * ldd + c] = bf16(src[r * cols + c]).  Split-precision mode (common.cuh): a weight row of logical width kw
// is stored as (hi | hi | lo), each part kw wide (ldd = 3 * kw).
__device__ __forceinline__ void put_weight(__nv_bfloat16* d, float v, int split, long long kw) {
  const __nv_bfloat16 hi = __float2bfloat16(v);
  d[0] = hi;
  if (split) {
    d[kw] = hi;
    d[2 * kw] = __float2bfloat16(v - __bfloat162float(hi));
  }
}
__global__ void pack_linear_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int rows, int cols,
                                   long long ldd, int split) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<long long>(rows) * cols) return;
  const int r = static_cast<int>(idx / cols), c = static_cast<int>(idx % cols);
  put_weight(dst + r * ldd + c, src[idx], split, ldd / 3);
}
// Conv2d weight [Cout][Cin][3][3] -> [Cout][(kh*3+kw)][Cin_pad]  (split mode: [..][3 * Cin_pad] = hi | hi | lo per tap)
__global__ void pack_conv3_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int Cout, int Cin,
                                  int Cin_pad, int split) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<long long>(Cout) * Cin * 9) return;
  const int tap = static_cast<int>(idx % 9);
  const int ci = static_cast<int>((idx / 9) % Cin);
  const int co = static_cast<int>(idx / (9LL * Cin));
  put_weight(dst + (static_cast<long long>(co) * 9 + tap) * (split ? 3 * Cin_pad : Cin_pad) + ci, src[idx], split, Cin_pad);
}
// ConvTranspose2d weight [Cin][Cout][k][k] -> [(kh*k+kw)*Cout_pad + co][Cin_pad]
__global__ void pack_convT_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int Cin, int Cout,
                                  int k, int Cin_pad, int Cout_pad, int split) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<long long>(Cin) * Cout * k * k) return;
  const int kk = static_cast<int>(idx % (k * k));
  const int co = static_cast<int>((idx / (k * k)) % Cout);
  const int ci = static_cast<int>(idx / (static_cast<long long>(k) * k * Cout));
  put_weight(dst + (static_cast<long long>(kk) * Cout_pad + co) * (split ? 3 * Cin_pad : Cin_pad) + ci, src[idx], split,
             Cin_pad);
}
// [rows][cols] fp32 -> [cols][rows] fp32
__global__ void transpose_f32_kernel(const float* __restrict__ src, float* __restrict__ dst, int rows, int cols) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * cols) return;
  const int r = idx / cols, c = idx % cols;
  dst[c * rows + r] = src[idx];
}
// decoder positions: [2B][N+1][2] int32 from two int64 [B][N][2] arrays, pose token at (-1,-1)
__global__ void build_dec_pos_kernel(const long long* __restrict__ p1, const long long* __restrict__ p2, int B, int N,
                                     int* __restrict__ out) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = 2LL * B * (N + 1);
  if (idx >= total) return;
  const int t = static_cast<int>(idx % (N + 1));
  const int s = static_cast<int>(idx / (N + 1));
  int y = -1, x = -1;
  if (t > 0) {
    const long long* p = (s < B) ? p1 + (static_cast<long long>(s) * N + (t - 1)) * 2
                                 : p2 + (static_cast<long long>(s - B) * N + (t - 1)) * 2;
    y = static_cast<int>(p[0]);
    x = static_cast<int>(p[1]);
  }
  out[2 * idx] = y;
  out[2 * idx + 1] = x;
}
__global__ void pos_to_int64_kernel(const int* __restrict__ in, long long n, long long* __restrict__ out) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx < n) out[idx] = in[idx];
}
// LayerNorm with fp32 output (all rows) -- only used by the per-layer-output compatibility path.
__global__ void __launch_bounds__(256)
layernorm_f32_kernel(const float* __restrict__ x, int rows, int C, float eps, const float* __restrict__ g,
                     const float* __restrict__ b, float* __restrict__ out) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = x + static_cast<long long>(row) * C;
  float sum = 0.f;
  for (int c = lane; c < C; c += 32) sum += xr[c];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / C;
  float sq = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float d = xr[c] - mean;
    sq += d * d;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float rstd = rsqrtf(sq / C + eps);
  float* orow = out + static_cast<long long>(row) * C;
  for (int c = lane; c < C; c += 32) orow[c] = (xr[c] - mean) * rstd * g[c] + b[c];
}

// ---------------------------------------------------------------------------
// model
// ---------------------------------------------------------------------------
struct Lin {
  bf16* w = nullptr;  // [N][K]
  float* b = nullptr; // [N] or null
  int N = 0, K = 0;   // K is the PHYSICAL reduction length: 3 x the logical one in split-precision mode
};
struct LNp {
  float *g = nullptr, *b = nullptr;
};
struct EncBlock {
  LNp n1, n2;
  Lin qkv, proj, fc1, fc2;
};
struct DecBlock {
  LNp n1, n2, n3, ny;
  Lin qkv, proj, cq, ckv, cproj, fc1, fc2;
};
struct Rcu {
  Lin c1, c2;  // 3x3 256 -> 256
};
struct Refine {
  Rcu r1, r2;
  Lin out;  // 1x1
};

enum PackKind { PK_F32, PK_LINEAR, PK_CONV3, PK_CONVT, PK_TRANSPOSE_F32, PK_IGNORE };
struct Slot {
  PackKind kind;
  std::vector<int64_t> shape;  // expected source shape
  void* dst = nullptr;
  long long ldd = 0;            // PK_LINEAR: destination row stride
  int cin_pad = 0, cout_pad = 0, k = 0;
  bool loaded = false;
};

struct Workspace {
  char* base = nullptr;
  size_t bytes = 0;
  size_t off = 0;
  int nimg = 0, h = 0, w = 0;
  int bmul = 1;  // bf16 activations are 3x wide in split-precision mode
  template <typename T>
  T* take(size_t n) {
    off = (off + 255) & ~static_cast<size_t>(255);
    T* p = reinterpret_cast<T*>(base + off);
    off += n * sizeof(T);
    return p;
  }
  bf16* takeb(size_t n) { return take<bf16>(n * bmul); }  // bf16 activation buffer with n LOGICAL elements
};

}  // namespace sta

using namespace sta;

struct StaModel {
  // Precision of the tensor-core operands: 0 = bf16 (production), 1 = split-precision "x3" parity mode -- every bf16
  // activation is (hi | lo | hi), every bf16 weight (hi | hi | lo) (common.cuh), the same gemm_tc_kernel mainloop runs
  // over 3K, attention runs in fp32 on the CUDA cores.  ~17 significant operand bits instead of 8; ~3.5x slower.
  int split = 0;
  int bmul() const { return split ? 3 : 1; }
  // ---- weights ----
  char* arena = nullptr;
  size_t arena_bytes = 0, arena_off = 0;
  std::unordered_map<std::string, Slot> slots;
  int missing = 0;

  float* pose_tok = nullptr;
  Lin patch, dec_embed;
  EncBlock enc[24];
  DecBlock dec[12];
  LNp dec_norm;
  // DPT
  Lin act0, act0T, act1, act1T, act2, act3, act3c;
  Lin rn[4];
  Refine ref[4];  // ref[0] = refinenet1 ... ref[3] = refinenet4
  Lin head0, head2;
  float *head4_w = nullptr, *head4_b = nullptr;
  PoseHeadWeights pose = {};

  // ---- workspace ----
  Workspace ws;
  float* stage = nullptr;  // staging for host->device weight uploads
  size_t stage_bytes = 0;
  char* io = nullptr;      // device staging for sta_forward_pairs_host
  size_t io_bytes = 0;
  static constexpr int kMaxHostChunks = 32;
  cudaStream_t s_in = nullptr, s_out = nullptr;  // copy streams of the host entry point
  static constexpr int kMaxParts = 4;  // DPT parts per view in the host entry point (D2H of a part overlaps the next part)
  cudaEvent_t ev_in[kMaxHostChunks] = {}, ev_done[kMaxHostChunks] = {}, ev_part[kMaxHostChunks][2 * kMaxParts] = {},
              ev_start = nullptr;
  int64_t launches = 0;
  int max_pairs_per_chunk = 16;
  int rp_K = 0, rp_H = 0, rp_W = 0;  // state between sta_regress_pairs_begin and _finish
  // The handle owns ONE workspace / io buffer / graph set: calls are ordered on whatever stream they are given.  When the
  // caller switches streams, the new stream first waits for the work this handle enqueued on the previous one.
  cudaStream_t last_stream = nullptr;
  bool has_last_stream = false;
  cudaEvent_t ev_switch = nullptr;

  // ---- CUDA-graph replay of the launch-bound small-batch entry points (SLAM mode) ----
  struct GraphEntry {
    std::vector<long long> key;
    cudaGraphExec_t exec = nullptr;  // null: seen once (ran eagerly), capture on the next call
    int64_t launches = 0;
    bool unsupported = false;  // capture / instantiate failed once: stay eager for this key
  };
  float* splitk_ws = nullptr;  // fp32 partial tiles of the split-K route for small problems (gemm.cu)
  static constexpr size_t kSplitKBytes = static_cast<size_t>(48) << 20;
  std::vector<GraphEntry> graphs;
  cudaStream_t s_cap = nullptr;  // capture stream (the caller's stream may be the legacy default stream)
  int graph_mode = -1;           // env STA_CUDA_GRAPHS (default on)
  int64_t graph_replays = 0;

  // ---- optional per-kernel-family timing (CUDA events on the launch stream) ----
  bool prof_on = false;
  struct ProfRec { int cat; cudaEvent_t e0, e1; };
  std::vector<ProfRec> prof_recs;
  std::vector<cudaEvent_t> prof_pool;
  double prof_ms[4] = {0, 0, 0, 0};
  int64_t prof_cnt[4] = {0, 0, 0, 0};
  double prof_flops[4] = {0, 0, 0, 0};
  cudaEvent_t prof_event() {
    cudaEvent_t e;
    if (!prof_pool.empty()) {
      e = prof_pool.back();
      prof_pool.pop_back();
    } else {
      cudaEventCreate(&e);
    }
    return e;
  }

  template <typename T>
  T* alloc(size_t n) {
    arena_off = (arena_off + 255) & ~static_cast<size_t>(255);
    T* p = reinterpret_cast<T*>(arena + arena_off);
    arena_off += n * sizeof(T);
    return p;
  }
};

namespace {

constexpr float kLnEps = 1e-6f;
constexpr int kEncDim = 1024, kDecDim = 768, kEncHeads = 16, kDecHeads = 12;

void add_slot(StaModel* m, const std::string& name, PackKind kind, std::vector<int64_t> shape, void* dst,
              long long ldd = 0, int cin_pad = 0, int cout_pad = 0, int k = 0) {
  Slot s;
  s.kind = kind;
  s.shape = std::move(shape);
  s.dst = dst;
  s.ldd = ldd;
  s.cin_pad = cin_pad;
  s.cout_pad = cout_pad;
  s.k = k;
  m->slots[name] = s;
  if (kind != PK_IGNORE) m->missing++;
}

// linear [N][K] (+bias) registered under prefix.weight / prefix.bias; optional zero padding of N
void reg_linear(StaModel* m, const std::string& prefix, Lin* L, int N, int K, bool bias = true, int N_pad = 0,
                std::vector<int64_t> wshape = {}) {
  const int Np = N_pad > 0 ? N_pad : N;
  L->N = Np;
  L->K = K * m->bmul();
  L->w = m->alloc<bf16>(static_cast<size_t>(Np) * L->K);
  if (wshape.empty()) wshape = {N, K};
  add_slot(m, prefix + ".weight", PK_LINEAR, wshape, L->w, L->K);
  if (bias) {
    L->b = m->alloc<float>(Np);
    add_slot(m, prefix + ".bias", PK_F32, {N}, L->b);
  }
}
void reg_ln(StaModel* m, const std::string& prefix, LNp* p, int C) {
  p->g = m->alloc<float>(C);
  p->b = m->alloc<float>(C);
  add_slot(m, prefix + ".weight", PK_F32, {C}, p->g);
  add_slot(m, prefix + ".bias", PK_F32, {C}, p->b);
}
void reg_conv3(StaModel* m, const std::string& prefix, Lin* L, int Cout, int Cin, bool bias, int Cin_pad = 0) {
  const int Cp = Cin_pad > 0 ? Cin_pad : Cin;
  L->N = Cout;
  L->K = 9 * Cp * m->bmul();
  L->w = m->alloc<bf16>(static_cast<size_t>(Cout) * L->K);
  add_slot(m, prefix + ".weight", PK_CONV3, {Cout, Cin, 3, 3}, L->w, 0, Cp);
  if (bias) {
    L->b = m->alloc<float>(Cout);
    add_slot(m, prefix + ".bias", PK_F32, {Cout}, L->b);
  }
}
void reg_convT(StaModel* m, const std::string& prefix, Lin* L, int C, int k, int C_pad) {
  L->N = k * k * C_pad;
  L->K = C_pad * m->bmul();
  L->w = m->alloc<bf16>(static_cast<size_t>(L->N) * L->K);
  add_slot(m, prefix + ".weight", PK_CONVT, {C, C, k, k}, L->w, 0, C_pad, C_pad, k);
  L->b = m->alloc<float>(C_pad);  // indexed by output channel
  add_slot(m, prefix + ".bias", PK_F32, {C}, L->b);
}
void reg_f32(StaModel* m, const std::string& name, float** dst, std::vector<int64_t> shape) {
  size_t n = 1;
  for (auto d : shape) n *= static_cast<size_t>(d);
  *dst = m->alloc<float>(n);
  add_slot(m, name, PK_F32, shape, *dst);
}

int build_registry(StaModel* m) {
  reg_f32(m, "init_pose_token", &m->pose_tok, {1, 1, kDecDim});
  reg_linear(m, "patch_embed.proj", &m->patch, kEncDim, 768, true, 0, {kEncDim, 3, 16, 16});
  for (int i = 0; i < 24; ++i) {
    const std::string p = "enc_blocks." + std::to_string(i) + ".";
    EncBlock& b = m->enc[i];
    reg_ln(m, p + "norm1", &b.n1, kEncDim);
    reg_linear(m, p + "attn.qkv", &b.qkv, 3 * kEncDim, kEncDim);
    reg_linear(m, p + "attn.proj", &b.proj, kEncDim, kEncDim);
    reg_ln(m, p + "norm2", &b.n2, kEncDim);
    reg_linear(m, p + "mlp.fc1", &b.fc1, 4 * kEncDim, kEncDim);
    reg_linear(m, p + "mlp.fc2", &b.fc2, kEncDim, 4 * kEncDim);
  }
  add_slot(m, "enc_norm.weight", PK_IGNORE, {kEncDim}, nullptr);  // never applied (normalize=False everywhere)
  add_slot(m, "enc_norm.bias", PK_IGNORE, {kEncDim}, nullptr);
  reg_linear(m, "decoder_embed", &m->dec_embed, kDecDim, kEncDim);
  for (int i = 0; i < 12; ++i) {
    const std::string p = "dec_block." + std::to_string(i) + ".";
    DecBlock& b = m->dec[i];
    reg_ln(m, p + "norm1", &b.n1, kDecDim);
    reg_linear(m, p + "attn.qkv", &b.qkv, 3 * kDecDim, kDecDim);
    reg_linear(m, p + "attn.proj", &b.proj, kDecDim, kDecDim);
    reg_linear(m, p + "cross_attn.projq", &b.cq, kDecDim, kDecDim);
    // projk and projv are fused into one [1536][768] matrix (both act on norm_y(y))
    b.ckv.N = 2 * kDecDim;
    b.ckv.K = kDecDim * m->bmul();
    b.ckv.w = m->alloc<bf16>(static_cast<size_t>(2) * kDecDim * b.ckv.K);
    b.ckv.b = m->alloc<float>(2 * kDecDim);
    add_slot(m, p + "cross_attn.projk.weight", PK_LINEAR, {kDecDim, kDecDim}, b.ckv.w, b.ckv.K);
    add_slot(m, p + "cross_attn.projk.bias", PK_F32, {kDecDim}, b.ckv.b);
    add_slot(m, p + "cross_attn.projv.weight", PK_LINEAR, {kDecDim, kDecDim},
             b.ckv.w + static_cast<size_t>(kDecDim) * b.ckv.K, b.ckv.K);
    add_slot(m, p + "cross_attn.projv.bias", PK_F32, {kDecDim}, b.ckv.b + kDecDim);
    reg_linear(m, p + "cross_attn.proj", &b.cproj, kDecDim, kDecDim);
    reg_ln(m, p + "norm2", &b.n2, kDecDim);
    reg_ln(m, p + "norm3", &b.n3, kDecDim);
    reg_linear(m, p + "mlp.fc1", &b.fc1, 4 * kDecDim, kDecDim);
    reg_linear(m, p + "mlp.fc2", &b.fc2, kDecDim, 4 * kDecDim);
    reg_ln(m, p + "norm_y", &b.ny, kDecDim);
  }
  reg_ln(m, "dec_norm", &m->dec_norm, kDecDim);

  const std::string d = "downstream_head_pts.dpt.";
  const int rn_cin[4] = {96, 192, 384, 768};
  const int rn_cin_pad[4] = {128, 192, 384, 768};
  for (int i = 0; i < 4; ++i) {
    reg_conv3(m, d + "scratch.layer_rn." + std::to_string(i), &m->rn[i], 256, rn_cin[i], false, rn_cin_pad[i]);
    // the same Parameter is also registered as scratch.layer{i+1}_rn (dpt_block.py:33-75)
    add_slot(m, d + "scratch.layer" + std::to_string(i + 1) + "_rn.weight", PK_IGNORE, {256, rn_cin[i], 3, 3}, nullptr);
  }
  for (int i = 0; i < 4; ++i) {
    const std::string p = d + "scratch.refinenet" + std::to_string(i + 1) + ".";
    Refine& r = m->ref[i];
    reg_linear(m, p + "out_conv", &r.out, 256, 256, true, 0, {256, 256, 1, 1});
    if (i == 3) {  // refinenet4 is called with one input: resConfUnit1 is never used (dpt_block.py:189-197)
      for (const char* c : {"conv1", "conv2"}) {
        add_slot(m, p + "resConfUnit1." + c + ".weight", PK_IGNORE, {256, 256, 3, 3}, nullptr);
        add_slot(m, p + "resConfUnit1." + c + ".bias", PK_IGNORE, {256}, nullptr);
      }
    } else {
      reg_conv3(m, p + "resConfUnit1.conv1", &r.r1.c1, 256, 256, true);
      reg_conv3(m, p + "resConfUnit1.conv2", &r.r1.c2, 256, 256, true);
    }
    reg_conv3(m, p + "resConfUnit2.conv1", &r.r2.c1, 256, 256, true);
    reg_conv3(m, p + "resConfUnit2.conv2", &r.r2.c2, 256, 256, true);
  }
  reg_conv3(m, d + "head.0", &m->head0, 128, 256, true);
  reg_conv3(m, d + "head.2", &m->head2, 128, 128, true);
  m->head4_w = m->alloc<float>(128 * 4);
  add_slot(m, d + "head.4.weight", PK_TRANSPOSE_F32, {4, 128, 1, 1}, m->head4_w);
  reg_f32(m, d + "head.4.bias", &m->head4_b, {4});
  reg_linear(m, d + "act_postprocess.0.0", &m->act0, 96, kEncDim, true, 128, {96, kEncDim, 1, 1});
  reg_convT(m, d + "act_postprocess.0.1", &m->act0T, 96, 4, 128);
  reg_linear(m, d + "act_postprocess.1.0", &m->act1, 192, kDecDim, true, 0, {192, kDecDim, 1, 1});
  reg_convT(m, d + "act_postprocess.1.1", &m->act1T, 192, 2, 192);
  reg_linear(m, d + "act_postprocess.2.0", &m->act2, 384, kDecDim, true, 0, {384, kDecDim, 1, 1});
  reg_linear(m, d + "act_postprocess.3.0", &m->act3, 768, kDecDim, true, 0, {768, kDecDim, 1, 1});
  reg_conv3(m, d + "act_postprocess.3.1", &m->act3c, 768, 768, true);

  float* tmp = nullptr;
  auto regp = [&](const std::string& name, const float** dst, std::vector<int64_t> shape) {
    reg_f32(m, name, &tmp, shape);
    *dst = tmp;
  };
  m->pose.ln_g = m->dec_norm.g;
  m->pose.ln_b = m->dec_norm.b;
  regp("head_pose_s.mlp.0.weight", &m->pose.w0, {512, kDecDim});
  regp("head_pose_s.mlp.0.bias", &m->pose.b0, {512});
  regp("head_pose_s.mlp.2.weight", &m->pose.w1, {512, 512});
  regp("head_pose_s.mlp.2.bias", &m->pose.b1, {512});
  regp("head_pose_s.mlp.4.weight", &m->pose.w2, {512, 512});
  regp("head_pose_s.mlp.4.bias", &m->pose.b2, {512});
  regp("head_pose_s.fc_t.weight", &m->pose.wt, {3, 512});
  regp("head_pose_s.fc_t.bias", &m->pose.bt, {3});
  regp("head_pose_s.fc_conf.0.weight", &m->pose.wc, {1, 512});
  regp("head_pose_s.fc_conf.0.bias", &m->pose.bc, {1});
  regp("head_pose_s.fc_rot.weight", &m->pose.wr, {9, 512});
  regp("head_pose_s.fc_rot.bias", &m->pose.br, {9});
  return 0;
}

// ---------------------------------------------------------------------------
// launch helpers (all increment the model's launch counter)
// ---------------------------------------------------------------------------
struct Ctx {
  StaModel* m;
  cudaStream_t st;
};

enum ProfCat { PROF_GEMM = 0, PROF_CONV = 1, PROF_ATTN = 2, PROF_LN = 3 };
// RAII scope: records an event pair around one launch when profiling is enabled
struct ProfScope {
  StaModel* m;
  cudaStream_t st;
  int cat;
  cudaEvent_t e0 = nullptr;
  ProfScope(const Ctx& c, int cat_, double flops) : m(c.m), st(c.st), cat(cat_) {
    if (!m->prof_on) return;
    m->prof_flops[cat] += flops;
    e0 = m->prof_event();
    cudaEventRecord(e0, st);
  }
  ~ProfScope() {
    if (!e0) return;
    cudaEvent_t e1 = m->prof_event();
    cudaEventRecord(e1, st);
    m->prof_recs.push_back({cat, e0, e1});
  }
};

int gemm(const Ctx& c, int amode, int epi, const bf16* A, long long lda, const Lin& L, GemmParams p) {
  GemmLaunch g;
  g.amode = amode;
  g.epi = epi;
  g.A = A;
  g.lda = lda;
  g.splitk_ws = c.m->splitk_ws;
  g.splitk_ws_bytes = c.m->splitk_ws ? StaModel::kSplitKBytes : 0;
  g.Wt = L.w;
  g.ldw = L.K;
  p.N = L.N;
  p.K = L.K;
  if (!p.bias) p.bias = L.b;
  p.split = c.m->split;
  if (p.split && (epi == EPI_BF16 || epi == EPI_GELU || epi == EPI_ROPE)) p.ldo *= 3;  // (hi | lo | hi) output rows
  g.p = p;
  c.m->launches++;
  const double rows = (amode == A_CONV3) ? static_cast<double>(p.nimg) * p.H * p.W : static_cast<double>(p.M);
  ProfScope ps(c, amode == A_CONV3 ? PROF_CONV : PROF_GEMM, 2.0 * rows * L.N * L.K);
  return launch_gemm(g, c.st);
}
// plain linear on rows
int linear(const Ctx& c, int epi, const bf16* A, int M, const Lin& L, void* out, const void* resid = nullptr,
           const int* pos = nullptr, int rope_cols = 0, int rowmap_n = 0) {
  GemmParams p = {};
  p.M = M;
  p.out = out;
  p.ldo = L.N;
  p.resid = resid;
  p.rowmap_n = rowmap_n;
  if (epi == EPI_ROPE) {
    p.pos = pos;
    p.rope_cols = rope_cols;
    p.rope_tab = rope_table(&p.rope_max_pos);
    if (!p.rope_tab) {
      set_last_error("failed to build the RoPE table");
      return 1;
    }
    p.rope_smem_rows = 64;  // positions -1..62 from shared memory, larger ones from the global table
  }
  return gemm(c, A_LINEAR, epi, A, L.K, L, p);
}
int conv3(const Ctx& c, const bf16* in, int nimg, int H, int W, int Cin, const Lin& L, bf16* out, bf16* out_relu,
          const bf16* resid, const bf16* resid2, int relu_main) {
  GemmParams p = {};
  p.nimg = nimg;
  p.H = H;
  p.W = W;
  p.Cin = Cin * c.m->bmul();  // physical channels of the NHWC input
  p.out = out;
  p.out2 = out_relu;
  p.ldo = L.N;
  p.resid = resid;
  p.resid2 = resid2;
  p.relu_main = relu_main;
  return gemm(c, A_CONV3, EPI_BF16, in, 0, L, p);
}
int ln(const Ctx& c, const float* x, int rows, int C, const LNp& a, bf16* out1, const LNp* b2 = nullptr,
       bf16* out2 = nullptr, int drop_first_of = 0) {
  c.m->launches++;
  ProfScope ps(c, PROF_LN, 0.0);
  return launch_layernorm(x, rows, C, kLnEps, a.g, a.b, out1, b2 ? b2->g : nullptr, b2 ? b2->b : nullptr, out2,
                          drop_first_of, c.st, c.m->split);
}
int attn(const Ctx& c, const bf16* q, long long ldq, int qc, const bf16* k, long long ldk, int kc, const bf16* v,
         long long ldv, int vc, bf16* out, long long ldo, int batch, int heads, int nq, int nk, int shift,
         int split_first_row = 0) {
  AttnLaunch a;
  const int bm = c.m->bmul();  // the ld* arguments are logical widths
  a.q = q; a.ldq = ldq * bm; a.q_col0 = qc;
  a.k = k; a.ldk = ldk * bm; a.k_col0 = kc;
  a.v = v; a.ldv = ldv * bm; a.v_col0 = vc;
  a.out = out; a.ldo = ldo * bm;
  a.batch = batch; a.heads = heads; a.nq = nq; a.nk = nk;
  a.kv_batch_shift = shift;
  a.split_first_row = split_first_row;
  a.split = c.m->split;
  a.scale = 0.125f;  // head_dim ** -0.5, sta_blocks.py:86
  c.m->launches++;
  ProfScope ps(c, PROF_ATTN, 4.0 * batch * heads * static_cast<double>(nq) * nk * 64);
  return launch_attention(a, c.st);
}

#define RUN(expr)            \
  do {                       \
    int _rc = (expr);        \
    if (_rc) return _rc;     \
  } while (0)

// ---------------------------------------------------------------------------
// workspace
// ---------------------------------------------------------------------------
struct EncBufs {
  bf16 *patches, *lnb, *qkv, *att, *hid;
  int* pos;
};
struct DecBufs {
  float* xd;
  bf16 *enc_bf16, *ln1, *lny, *qkv, *att, *qc, *kvc, *hid, *hook[3];
  int* pos;
};
struct DptBufs {
  bf16 *a0, *l1, *a1, *l2, *l3, *a3, *col4, *l4;
  bf16 *r_raw[4], *r_relu[4];
  bf16 *t1, *t2, *s_raw, *s_relu, *o, *up, *path;  // sized for the largest level
  bf16 *hc1, *hup;
};

size_t ws_need(int nimg, int h, int w, int bmul = 1) {
  if (bmul > 1) return ws_need(nimg, h, w) * bmul;  // split-precision mode: bf16 buffers are 3x wide (fp32 ones over-counted)
  const size_t N = static_cast<size_t>(h) * w, T = nimg * N, Td = nimg * (N + 1);
  size_t b = 0;
  auto add = [&](size_t n) { b += ((n + 255) / 256) * 256 + 256; };
  // encoder
  add(T * 768 * 2); add(T * 1024 * 2); add(T * 3072 * 2); add(T * 1024 * 2); add(T * 4096 * 2); add(T * 8);
  add(T * 1024 * 4);  // x (fp32 residual stream)
  // decoder
  add(Td * 768 * 4); add(T * 1024 * 2); add(Td * 768 * 2 * 2); add(Td * 2304 * 2); add(Td * 768 * 2 * 2);
  add(Td * 1536 * 2); add(Td * 3072 * 2); add(T * 768 * 2 * 3); add(Td * 8);
  // dpt
  const size_t h4 = (h + 1) / 2, w4 = (w + 1) / 2;
  const size_t P1 = 16 * N, P2 = 4 * N, P3 = N, P4 = h4 * w4;
  add(T * 128 * 2); add(nimg * P1 * 128 * 2); add(T * 192 * 2); add(nimg * P2 * 192 * 2); add(T * 384 * 2);
  add(T * 768 * 2); add(nimg * P4 * 6912 * 2); add(nimg * P4 * 768 * 2);
  const size_t P[4] = {P1, P2, P3, P4};
  for (int i = 0; i < 4; ++i) { add(nimg * P[i] * 256 * 2); add(nimg * P[i] * 256 * 2); }
  for (int i = 0; i < 5; ++i) add(nimg * P1 * 256 * 2);       // t1, t2, s_raw, s_relu, o
  add(nimg * 64 * N * 256 * 2); add(nimg * 64 * N * 256 * 2);  // up, path (up to 8h x 8w)
  add(nimg * 64 * N * 128 * 2); add(nimg * 256 * N * 128 * 2); // hc1, hup (full res)
  // static outputs of the graph-replayed keyframe step: pts3d, conf, depth (fp32, full res), poses, K, reduction scratch
  add(nimg * 256 * N * 5 * 4); add(nimg * 32 * 4); add(pointmap_scratch_bytes(nimg));
  // gated keyframe step: compact copies of the surviving edges' hooks (sta_regress_pairs_finish)
  add(T * 1024 * 2); add(T * 768 * 2 * 3); add(nimg * 32 * 4);
  return b + (1 << 20);
}

void drop_graphs(StaModel* m);

int ensure_ws(StaModel* m, int nimg, int h, int w) {
  const size_t need = ws_need(nimg, h, w, m->bmul());
  m->ws.bmul = m->bmul();
  if (m->ws.bytes < need) {
    STA_CHECK_CUDA(cudaDeviceSynchronize());
    drop_graphs(m);  // they point into the old workspace
    if (m->ws.base) STA_CHECK_CUDA(cudaFree(m->ws.base));
    m->ws.base = nullptr;
    m->ws.bytes = 0;
    STA_CHECK_CUDA(cudaMalloc(&m->ws.base, need));
    m->ws.bytes = need;
  }
  m->ws.off = 0;
  m->ws.nimg = nimg;
  m->ws.h = h;
  m->ws.w = w;
  return 0;
}

// ---------------------------------------------------------------------------
// Small-batch calls (one keyframe, a handful of edges) are launch-bound: ~170-400 kernels of a few microseconds each.
// run_graphed() runs `body(ctx)` eagerly the first time a key is seen (lazy initialisation such as
// cudaFuncSetAttribute / the RoPE table happens there), captures it into a CUDA graph on the second call (on a private
// stream -- the caller's may be the legacy default stream -- with the programmatic-dependent-launch edges kept) and
// replays the instantiated graph on the caller's stream from then on.  `body` must only touch memory whose
// addresses are a function of the key (workspace offsets for that shape, weights): user pointers are handled by
// the caller outside the graph.  Graphs die with the workspace they point into.
// ---------------------------------------------------------------------------
constexpr long long kGraphMaxTokens = 8192;  // only launch-bound problem sizes

void drop_graphs(StaModel* m) {
  for (auto& g : m->graphs)
    if (g.exec) cudaGraphExecDestroy(g.exec);
  m->graphs.clear();
}

bool graphs_enabled(StaModel* m) {
  if (m->graph_mode < 0) {
    const char* e = getenv("STA_CUDA_GRAPHS");
    m->graph_mode = (e && e[0] == '0') ? 0 : 1;
  }
  return m->graph_mode == 1 && !m->prof_on;
}

template <typename Body>
int run_graphed(StaModel* m, cudaStream_t st, const std::vector<long long>& key, Body&& body) {
  if (!graphs_enabled(m)) return body(Ctx{m, st});
  StaModel::GraphEntry* ent = nullptr;
  for (auto& g : m->graphs)
    if (g.key == key) ent = &g;
  if (!ent) {
    if (m->graphs.size() >= 32) drop_graphs(m);
    m->graphs.emplace_back();
    m->graphs.back().key = key;
    return body(Ctx{m, st});  // first sighting: eager
  }
  if (ent->unsupported) return body(Ctx{m, st});
  if (!ent->exec) {
    if (!m->s_cap) STA_CHECK_CUDA(cudaStreamCreateWithFlags(&m->s_cap, cudaStreamNonBlocking));
    const int64_t l0 = m->launches;
    STA_CHECK_CUDA(cudaStreamBeginCapture(m->s_cap, cudaStreamCaptureModeThreadLocal));
    const int rc = body(Ctx{m, m->s_cap});
    cudaGraph_t graph = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(m->s_cap, &graph);
    if (rc != 0) {
      if (graph) cudaGraphDestroy(graph);
      cudaGetLastError();
      return rc;
    }
    cudaGraphExec_t exec = nullptr;
    if (ce != cudaSuccess || graph == nullptr || cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess) {
      if (graph) cudaGraphDestroy(graph);
      cudaGetLastError();  // clear; fall back to eager launches for this key
      ent->unsupported = true;
      m->launches = l0;
      return body(Ctx{m, st});
    }
    cudaGraphDestroy(graph);
    ent->exec = exec;
    ent->launches = m->launches - l0;
    m->launches = l0;
  }
  STA_CHECK_CUDA(cudaGraphLaunch(ent->exec, st));
  m->launches += ent->launches;
  m->graph_replays++;
  return 0;
}

// ---------------------------------------------------------------------------
// encoder over nimg images whose patches (e.patches) and positions (e.pos) are prepared.
// x: fp32 residual stream [nimg*N][1024], provided by the caller (user buffer or workspace);
// on return it holds the un-normalised encoder features (normalize=False, sta_model.py:172).
// ---------------------------------------------------------------------------
int run_encoder(const Ctx& c, int nimg, int N, float* x, EncBufs& e) {
  StaModel* m = c.m;
  const int T = nimg * N;
  RUN(linear(c, EPI_F32, e.patches, T, m->patch, x));
  for (int l = 0; l < 24; ++l) {
    const EncBlock& b = m->enc[l];
    RUN(ln(c, x, T, kEncDim, b.n1, e.lnb));
    RUN(linear(c, EPI_ROPE, e.lnb, T, b.qkv, e.qkv, nullptr, e.pos, 2 * kEncDim));
    RUN(attn(c, e.qkv, 3 * kEncDim, 0, e.qkv, 3 * kEncDim, kEncDim, e.qkv, 3 * kEncDim, 2 * kEncDim, e.att, kEncDim,
             nimg, kEncHeads, N, N, 0));
    RUN(linear(c, EPI_F32, e.att, T, b.proj, x, x));
    RUN(ln(c, x, T, kEncDim, b.n2, e.lnb));
    RUN(linear(c, EPI_GELU, e.lnb, T, b.fc1, e.hid));
    RUN(linear(c, EPI_F32, e.hid, T, b.fc2, x, x));
  }
  return 0;
}

EncBufs take_enc(Workspace& ws, int nimg, int N) {
  const size_t T = static_cast<size_t>(nimg) * N;
  EncBufs e;
  e.patches = ws.takeb(T * 768);
  e.lnb = ws.takeb(T * 1024);
  e.qkv = ws.takeb(T * 3072);
  e.att = ws.takeb(T * 1024);
  e.hid = ws.takeb(T * 4096);
  e.pos = ws.take<int>(T * 2);
  return e;
}
DecBufs take_dec(Workspace& ws, int S, int N) {
  const size_t T = static_cast<size_t>(S) * N, Td = static_cast<size_t>(S) * (N + 1);
  DecBufs d;
  d.xd = ws.take<float>(Td * 768);
  d.enc_bf16 = ws.takeb(T * 1024);
  d.ln1 = ws.takeb(Td * 768);
  d.lny = ws.takeb(Td * 768);
  d.qkv = ws.takeb(Td * 2304);
  d.att = ws.takeb(Td * 768);
  d.qc = ws.takeb(Td * 768);
  d.kvc = ws.takeb(Td * 1536);
  d.hid = ws.takeb(Td * 3072);
  for (int i = 0; i < 3; ++i) d.hook[i] = ws.takeb(T * 768);
  d.pos = ws.take<int>(Td * 2);
  return d;
}

// ---------------------------------------------------------------------------
// symmetric decoder over S = 2B samples (samples [0,B) = view 1, [B,2B) = view 2).
// d.enc_bf16 (bf16 encoder features of all S samples) and d.pos must be filled.
// outs: optional 2 x 13 fp32 per-layer outputs (compat path), may be null.
// ---------------------------------------------------------------------------
int run_decoder(const Ctx& c, int B, int N, DecBufs& d, float* const* out1, float* const* out2) {
  StaModel* m = c.m;
  const int S = 2 * B, M = N + 1, T = S * N, Td = S * M;
  const long long half = static_cast<long long>(B) * M * kDecDim;
  auto emit = [&](int idx) -> int {
    if (out1 && out1[idx]) { m->launches++; RUN(launch_copy_f32(d.xd, out1[idx], half, c.st)); }
    if (out2 && out2[idx]) { m->launches++; RUN(launch_copy_f32(d.xd + half, out2[idx], half, c.st)); }
    return 0;
  };
  RUN(linear(c, EPI_F32, d.enc_bf16, T, m->dec_embed, d.xd, nullptr, nullptr, 0, N));
  m->launches++;
  RUN(launch_fill_pose_token(d.xd, m->pose_tok, S, M, kDecDim, c.st));
  RUN(emit(0));
  // the pose token (query row 0) goes to the attention kernel's spare warps when that saves a pair of query tiles
  // The pose token (query row 0) can go to a separate small kernel so that the tiled kernel runs 3 instead of 4
  // query-tile pairs per (head, sample).  Measured on B200 the attention time drops (7.9 -> 7.5 ms per cfg-2 step) but
  // the 24 extra launches cost more than that end to end (34.0 vs 33.4 ms), so it is off unless STA_ATTN_SPLIT_POSE=1.
  static int split_env = -1;
  if (split_env < 0) {
    const char* e = getenv("STA_ATTN_SPLIT_POSE");
    split_env = (e && e[0] == '1') ? 1 : 0;
  }
  const int split_pose = (split_env && (M - 1 + 255) / 256 < (M + 255) / 256) ? 1 : 0;
  for (int l = 0; l < 12; ++l) {
    const DecBlock& b = m->dec[l];
    // self-attention on norm1(x); norm_y(x) is what the partner view cross-attends to
    RUN(ln(c, d.xd, Td, kDecDim, b.n1, d.ln1, &b.ny, d.lny));
    RUN(linear(c, EPI_ROPE, d.ln1, Td, b.qkv, d.qkv, nullptr, d.pos, 2 * kDecDim));
    RUN(attn(c, d.qkv, 3 * kDecDim, 0, d.qkv, 3 * kDecDim, kDecDim, d.qkv, 3 * kDecDim, 2 * kDecDim, d.att, kDecDim, S,
             kDecHeads, M, M, 0, split_pose));
    RUN(linear(c, EPI_F32, d.att, Td, b.proj, d.xd, d.xd));
    // cross-attention: q from norm2(x), k/v from norm_y(partner input) -- kv sample = (s + B) % 2B
    RUN(linear(c, EPI_ROPE, d.lny, Td, b.ckv, d.kvc, nullptr, d.pos, kDecDim));
    RUN(ln(c, d.xd, Td, kDecDim, b.n2, d.ln1));
    RUN(linear(c, EPI_ROPE, d.ln1, Td, b.cq, d.qc, nullptr, d.pos, kDecDim));
    RUN(attn(c, d.qc, kDecDim, 0, d.kvc, 2 * kDecDim, 0, d.kvc, 2 * kDecDim, kDecDim, d.att, kDecDim, S, kDecHeads, M, M,
             B, split_pose));
    RUN(linear(c, EPI_F32, d.att, Td, b.cproj, d.xd, d.xd));
    // MLP
    RUN(ln(c, d.xd, Td, kDecDim, b.n3, d.ln1));
    RUN(linear(c, EPI_GELU, d.ln1, Td, b.fc1, d.hid));
    RUN(linear(c, EPI_F32, d.hid, Td, b.fc2, d.xd, d.xd));
    if (l + 1 == 6 || l + 1 == 9) {  // DPT hooks [0, 7, 10, 13] -> decoder outputs 6, 9, 12 (dpt_head.py:112)
      m->launches++;
      RUN(launch_cast_f32_bf16(d.xd, d.hook[l + 1 == 6 ? 0 : 1], Td, kDecDim, M, c.st, m->split));
    }
    if (l + 1 < 12) RUN(emit(l + 1));
  }
  // dec_norm on the last output; bf16 copy without the pose token feeds the DPT head
  RUN(ln(c, d.xd, Td, kDecDim, m->dec_norm, d.hook[2], nullptr, nullptr, M));
  if ((out1 && out1[12]) || (out2 && out2[12])) {
    // compat path: fp32 LayerNorm-ed last layer
    const int rows_half = B * M;
    if (out1 && out1[12]) {
      m->launches++;
      layernorm_f32_kernel<<<(rows_half + 7) / 8, 256, 0, c.st>>>(d.xd, rows_half, kDecDim, kLnEps, m->dec_norm.g,
                                                                  m->dec_norm.b, out1[12]);
    }
    if (out2 && out2[12]) {
      m->launches++;
      layernorm_f32_kernel<<<(rows_half + 7) / 8, 256, 0, c.st>>>(d.xd + half, rows_half, kDecDim, kLnEps,
                                                                  m->dec_norm.g, m->dec_norm.b, out2[12]);
    }
    STA_CHECK_CUDA(cudaGetLastError());
  }
  return 0;
}

// ---------------------------------------------------------------------------
// DPT head over nimg images.  hooks: bf16 token matrices [nimg*N][1024 | 768 | 768 | 768].
// ---------------------------------------------------------------------------
int run_dpt(const Ctx& c, Workspace& ws, int nimg, int h, int w, const bf16* hook0, const bf16* hook1,
            const bf16* hook2, const bf16* hook3, float* pts3d, float* conf) {
  StaModel* m = c.m;
  const int N = h * w, T = nimg * N;
  const int h4 = (h + 1) / 2, w4 = (w + 1) / 2;
  const int LH[4] = {4 * h, 2 * h, h, h4}, LW[4] = {4 * w, 2 * w, w, w4};
  DptBufs b;
  b.a0 = ws.takeb(static_cast<size_t>(T) * 128);
  b.l1 = ws.takeb(static_cast<size_t>(nimg) * 16 * N * 128);
  b.a1 = ws.takeb(static_cast<size_t>(T) * 192);
  b.l2 = ws.takeb(static_cast<size_t>(nimg) * 4 * N * 192);
  b.l3 = ws.takeb(static_cast<size_t>(T) * 384);
  b.a3 = ws.takeb(static_cast<size_t>(T) * 768);
  b.col4 = ws.takeb(static_cast<size_t>(nimg) * h4 * w4 * 6912);
  b.l4 = ws.takeb(static_cast<size_t>(nimg) * h4 * w4 * 768);
  for (int i = 0; i < 4; ++i) {
    b.r_raw[i] = ws.takeb(static_cast<size_t>(nimg) * LH[i] * LW[i] * 256);
    b.r_relu[i] = ws.takeb(static_cast<size_t>(nimg) * LH[i] * LW[i] * 256);
  }
  const size_t big = static_cast<size_t>(nimg) * 16 * N * 256;
  b.t1 = ws.takeb(big);
  b.t2 = ws.takeb(big);
  b.s_raw = ws.takeb(big);
  b.s_relu = ws.takeb(big);
  b.o = ws.takeb(big);
  b.up = ws.takeb(static_cast<size_t>(nimg) * 64 * N * 256);
  b.path = ws.takeb(static_cast<size_t>(nimg) * 64 * N * 256);
  b.hc1 = ws.takeb(static_cast<size_t>(nimg) * 64 * N * 128);
  b.hup = ws.takeb(static_cast<size_t>(nimg) * 256 * N * 128);

  // ---- act_postprocess (dpt_block.py:356-410) ----
  RUN(linear(c, EPI_BF16, hook0, T, m->act0, b.a0));
  {
    GemmParams p = {};
    p.M = T; p.out = b.l1; p.ps_k = 4; p.ps_cout = 128; p.ps_h = h; p.ps_w = w;
    RUN(gemm(c, A_LINEAR, EPI_PIXSHUF, b.a0, m->act0T.K, m->act0T, p));
  }
  RUN(linear(c, EPI_BF16, hook1, T, m->act1, b.a1));
  {
    GemmParams p = {};
    p.M = T; p.out = b.l2; p.ps_k = 2; p.ps_cout = 192; p.ps_h = h; p.ps_w = w;
    RUN(gemm(c, A_LINEAR, EPI_PIXSHUF, b.a1, m->act1T.K, m->act1T, p));
  }
  RUN(linear(c, EPI_BF16, hook2, T, m->act2, b.l3));
  RUN(linear(c, EPI_BF16, hook3, T, m->act3, b.a3));
  m->launches++;
  RUN(launch_im2col_3x3_s2(b.a3, b.col4, nimg, h, w, 768 * m->bmul(), c.st));  // pure channel-vector copy
  RUN(linear(c, EPI_BF16, b.col4, nimg * h4 * w4, m->act3c, b.l4));
  // ---- layer_rn: 3x3 conv to 256 channels, no bias (dpt_block.py:33-75); raw + relu copies ----
  const bf16* lin[4] = {b.l1, b.l2, b.l3, b.l4};
  const int lc[4] = {128, 192, 384, 768};
  for (int i = 0; i < 4; ++i)
    RUN(conv3(c, lin[i], nimg, LH[i], LW[i], lc[i], m->rn[i], b.r_raw[i], b.r_relu[i], nullptr, nullptr, 0));

  // ---- refinenet4 .. refinenet1 (dpt_block.py:189-218, dpt_head.py:58-61) ----
  const bf16* path = nullptr;  // output of the previous (coarser) fusion block at this level's resolution
  for (int lvl = 3; lvl >= 0; --lvl) {
    const Refine& r = m->ref[lvl];
    const int Hh = LH[lvl], Ww = LW[lvl];
    const bf16* s_raw;
    const bf16* s_relu;
    if (lvl == 3) {
      s_raw = b.r_raw[3];
      s_relu = b.r_relu[3];
    } else {
      // res = RCU1(layer); s = path + res
      RUN(conv3(c, b.r_relu[lvl], nimg, Hh, Ww, 256, r.r1.c1, b.t1, nullptr, nullptr, nullptr, 1));
      RUN(conv3(c, b.t1, nimg, Hh, Ww, 256, r.r1.c2, b.s_raw, b.s_relu, b.r_raw[lvl], path, 0));
      s_raw = b.s_raw;
      s_relu = b.s_relu;
    }
    // out = RCU2(s)
    RUN(conv3(c, s_relu, nimg, Hh, Ww, 256, r.r2.c1, b.t2, nullptr, nullptr, nullptr, 1));
    RUN(conv3(c, b.t2, nimg, Hh, Ww, 256, r.r2.c2, b.o, nullptr, s_raw, nullptr, 0));
    // bilinear x2 (align_corners=True), cropped to the next level's size for refinenet4 (dpt_head.py:58)
    // The reference applies out_conv (1x1) AFTER the upsample (dpt_block.py:215-217).  Both are linear and the
    // bilinear weights sum to one, so conv1x1(up(x)) == up(conv1x1(x)) exactly in real arithmetic; doing the
    // 1x1 at the low resolution costs a quarter of the FLOPs and HBM bytes.
    int OH = 2 * Hh, OW = 2 * Ww;
    if (lvl == 3) { OH = LH[2]; OW = LW[2]; }
    RUN(linear(c, EPI_BF16, b.o, nimg * Hh * Ww, r.out, b.up));
    m->launches++;
    RUN(launch_upsample2x(b.up, b.path, nimg, Hh, Ww, 256, OH, OW, c.st, m->split));
    path = b.path;
  }
  // ---- head (dpt_block.py:318-324) + postprocess (postprocess.py:10-62) ----
  const int H8 = 8 * h, W8 = 8 * w;
  RUN(conv3(c, b.path, nimg, H8, W8, 256, m->head0, b.hc1, nullptr, nullptr, nullptr, 0));
  m->launches++;
  RUN(launch_upsample2x(b.hc1, b.hup, nimg, H8, W8, 128, 2 * H8, 2 * W8, c.st, m->split));
  {
    GemmParams p = {};
    p.nimg = nimg; p.H = 2 * H8; p.W = 2 * W8; p.Cin = 128 * m->bmul();
    p.head_w = m->head4_w; p.head_b = m->head4_b; p.pts3d = pts3d; p.conf = conf;
    RUN(gemm(c, A_CONV3, EPI_HEAD, b.hup, 0, m->head2, p));
  }
  return 0;
}

int check_ready(StaModel* m) {
  if (!m) {
    set_last_error("null model handle");
    return 2;
  }
  if (m->missing != 0) {
    char buf[128];
    snprintf(buf, sizeof(buf), "model is missing %d state-dict tensors", m->missing);
    set_last_error(buf);
    return 2;
  }
  return 0;
}

// DPT parts per view in the host entry point.  Two halves measured best on B200 (r02: exposed copy 2.0 ms per 16-pair step;
// four parts of 4 images shorten the D2H tail but the smaller DPT batches cost more than that: 3.0 ms).
int host_dpt_parts(int B) { return B >= 2 ? 2 : 1; }

// one stream per handle at a time: order a call on `st` after everything this handle enqueued on its previous stream
int enter_stream(StaModel* m, cudaStream_t st) {
  if (m->has_last_stream && m->last_stream != st) {
    if (!m->ev_switch) STA_CHECK_CUDA(cudaEventCreateWithFlags(&m->ev_switch, cudaEventDisableTiming));
    STA_CHECK_CUDA(cudaEventRecord(m->ev_switch, m->last_stream));
    STA_CHECK_CUDA(cudaStreamWaitEvent(st, m->ev_switch, 0));
  }
  m->last_stream = st;
  m->has_last_stream = true;
  return 0;
}

int forward_chunk(const Ctx& c, const void* img1, const void* img2, int img_is_bf16, int B, int H, int W,
                  float* pts3d, float* conf, float* pose, float* pose_conf, int B_total,
                  cudaEvent_t* ev_parts = nullptr) {
  StaModel* m = c.m;
  const int h = H / 16, w = W / 16, N = h * w, S = 2 * B, M = N + 1;
  RUN(ensure_ws(m, S, h, w));
  Workspace& ws = m->ws;
  float* x = ws.take<float>(static_cast<size_t>(S) * N * kEncDim);
  EncBufs e = take_enc(ws, S, N);
  DecBufs d = take_dec(ws, S, N);
  // encode both views as one batch of 2B images: [view 1 batch ; view 2 batch]
  m->launches += 3;
  const size_t bm = m->bmul();  // physical / logical width of the bf16 activation rows
  RUN(launch_patch_im2col(img1, img_is_bf16, B, H, W, e.patches, c.st, m->split));
  RUN(launch_patch_im2col(img2, img_is_bf16, B, H, W, e.patches + static_cast<size_t>(B) * N * 768 * bm, c.st, m->split));
  RUN(launch_make_positions(e.pos, S, h, w, 0, c.st));
  RUN(run_encoder(c, S, N, x, e));
  m->launches += 2;
  RUN(launch_cast_f32_bf16(x, d.enc_bf16, static_cast<long long>(S) * N, kEncDim, 0, c.st, m->split));
  RUN(launch_make_positions(d.pos, S, h, w, 1, c.st));
  RUN(run_decoder(c, B, N, d, nullptr, nullptr));
  // pose heads on the (dec_norm-ed) pose tokens of both views
  {
    // outputs are laid out [2][B_total]: view 1 block then view 2 block
    m->launches += 2;
    RUN(launch_pose_head(d.xd, static_cast<long long>(M) * kDecDim, B, 1, kLnEps, m->pose, pose, pose_conf, c.st));
    RUN(launch_pose_head(d.xd + static_cast<long long>(B) * M * kDecDim, static_cast<long long>(M) * kDecDim, B, 1,
                         kLnEps, m->pose, pose + static_cast<long long>(B_total) * 16, pose_conf + B_total, c.st));
  }
  // DPT heads: view 1 images then view 2 images (outputs are [2][B_total] blocks).  The host entry point passes events:
  // each view is then processed in `halves` parts (host_dpt_parts) and an event is recorded after every part, so that the
  // D2H copy of one part runs under the head of the next and only the last quarter of the outputs is exposed.
  const long long px = static_cast<long long>(H) * W;
  const int halves = ev_parts ? host_dpt_parts(B) : 1;
  for (int v = 0; v < 2; ++v) {
    for (int hf = 0; hf < halves; ++hf) {
      const int i0 = hf * (B / halves), nb = (hf == halves - 1) ? B - i0 : B / halves;
      const size_t save = ws.off;
      const size_t tok0 = (static_cast<size_t>(v) * B + i0) * N;
      RUN(run_dpt(c, ws, nb, h, w, d.enc_bf16 + tok0 * 1024 * bm, d.hook[0] + tok0 * 768 * bm, d.hook[1] + tok0 * 768 * bm,
                  d.hook[2] + tok0 * 768 * bm, pts3d + (static_cast<long long>(v) * B_total + i0) * px * 3,
                  conf + (static_cast<long long>(v) * B_total + i0) * px));
      ws.off = save;
      if (ev_parts) STA_CHECK_CUDA(cudaEventRecord(ev_parts[v * halves + hf], c.st));
    }
  }
  return 0;
}

}  // namespace

// ===========================================================================
// C ABI
// ===========================================================================
extern "C" {

const char* sta_last_error(void) { return get_last_error(); }
int sta_version(void) { return 1; }

int sta_device_synchronize(void) {
  STA_CHECK_CUDA(cudaDeviceSynchronize());
  return 0;
}

int sta_create(StaModel** out) { return sta_create_ex(out, STA_PRECISION_BF16); }

int sta_create_ex(StaModel** out, int precision) {
  if (precision != STA_PRECISION_BF16 && precision != STA_PRECISION_X3) {
    set_last_error("sta_create_ex: unknown precision (0 = bf16, 1 = split-precision x3 parity mode)");
    return 2;
  }
  if (!out) {
    set_last_error("sta_create: null output pointer");
    return 2;
  }
  *out = nullptr;
  int dev = 0;
  STA_CHECK_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  STA_CHECK_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10) {
    char buf[160];
    snprintf(buf, sizeof(buf), "sta_b200 needs a Blackwell sm_100 GPU (found %s, sm_%d%d); there is no fallback path",
             prop.name, prop.major, prop.minor);
    set_last_error(buf);
    return 3;
  }
  StaModel* m = new StaModel();
  m->split = (precision == STA_PRECISION_X3) ? 1 : 0;
  // 438.5 M params: ~877 MB bf16 + fp32 vectors + padding (3x the bf16 part in split-precision mode)
  m->arena_bytes = static_cast<size_t>(m->split ? 2800 : 960) << 20;
  cudaError_t e = cudaMalloc(&m->arena, m->arena_bytes);
  if (e != cudaSuccess) {
    set_last_error(std::string("cudaMalloc(weight arena) failed: ") + cudaGetErrorString(e));
    delete m;
    return 1;
  }
  cudaMemset(m->arena, 0, m->arena_bytes);
  if (cudaMalloc(&m->splitk_ws, StaModel::kSplitKBytes) != cudaSuccess) {
    cudaGetLastError();
    m->splitk_ws = nullptr;  // the split-K route is simply not taken
  }
  build_registry(m);
  if (m->arena_off > m->arena_bytes) {
    set_last_error("internal: weight arena too small");
    cudaFree(m->arena);
    delete m;
    return 1;
  }
  *out = m;
  return 0;
}

void sta_destroy(StaModel* m) {
  if (!m) return;
  cudaDeviceSynchronize();
  if (m->arena) cudaFree(m->arena);
  if (m->ws.base) cudaFree(m->ws.base);
  if (m->stage) cudaFree(m->stage);
  if (m->io) cudaFree(m->io);
  if (m->splitk_ws) cudaFree(m->splitk_ws);
  drop_graphs(m);
  if (m->s_cap) cudaStreamDestroy(m->s_cap);
  if (m->ev_switch) cudaEventDestroy(m->ev_switch);
  if (m->s_in) {
    cudaStreamDestroy(m->s_in);
    cudaStreamDestroy(m->s_out);
    for (int i = 0; i < StaModel::kMaxHostChunks; ++i) {
      cudaEventDestroy(m->ev_in[i]);
      cudaEventDestroy(m->ev_done[i]);
      for (int k = 0; k < 2 * StaModel::kMaxParts; ++k) cudaEventDestroy(m->ev_part[i][k]);
    }
    cudaEventDestroy(m->ev_start);
  }
  for (auto& r : m->prof_recs) {
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  for (auto e : m->prof_pool) cudaEventDestroy(e);
  delete m;
}

int sta_missing_tensors(StaModel* m) { return m ? m->missing : -1; }
int64_t sta_launch_count(StaModel* m) { return m ? m->launches : 0; }
int64_t sta_device_bytes(StaModel* m) {
  return m ? static_cast<int64_t>(m->arena_bytes + m->ws.bytes + m->stage_bytes + m->io_bytes) : 0;
}
int sta_profile(StaModel* m, int enable) {
  if (!m) return 2;
  m->prof_on = enable != 0;
  return 0;
}
int sta_profile_read(StaModel* m, double* ms4, int64_t* counts4, double* flops4) {
  if (!m || !ms4 || !counts4 || !flops4) {
    set_last_error("sta_profile_read: null argument");
    return 2;
  }
  STA_CHECK_CUDA(cudaDeviceSynchronize());
  for (auto& r : m->prof_recs) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess) {
      m->prof_ms[r.cat] += ms;
      m->prof_cnt[r.cat]++;
    }
    m->prof_pool.push_back(r.e0);
    m->prof_pool.push_back(r.e1);
  }
  m->prof_recs.clear();
  for (int i = 0; i < 4; ++i) {
    ms4[i] = m->prof_ms[i];
    counts4[i] = m->prof_cnt[i];
    flops4[i] = m->prof_flops[i];
    m->prof_ms[i] = 0;
    m->prof_cnt[i] = 0;
    m->prof_flops[i] = 0;
  }
  return 0;
}
int sta_weight_arena(StaModel* m, void** ptr, int64_t* bytes) {
  if (!m || !ptr || !bytes) {
    set_last_error("sta_weight_arena: null argument");
    return 2;
  }
  *ptr = m->arena;
  *bytes = static_cast<int64_t>((m->arena_off + 255) & ~static_cast<size_t>(255));
  return 0;
}
int sta_mark_all_loaded(StaModel* m) {  // after the arena was filled by a broadcast from another rank
  if (!m) return 2;
  for (auto& kv : m->slots) kv.second.loaded = true;
  m->missing = 0;
  return 0;
}

int sta_load_tensor(StaModel* m, const char* name, const float* data, const int64_t* shape, int ndim,
                    int data_on_device) {
  if (!m || !name || !data) {
    set_last_error("sta_load_tensor: null argument");
    return 2;
  }
  auto it = m->slots.find(name);
  if (it == m->slots.end()) {
    set_last_error(std::string("unexpected key in state_dict: ") + name);
    return 4;
  }
  Slot& s = it->second;
  bool ok = static_cast<int>(s.shape.size()) == ndim;
  size_t numel = 1;
  for (int i = 0; ok && i < ndim; ++i) {
    ok = (shape[i] == s.shape[i]);
    numel *= static_cast<size_t>(shape[i]);
  }
  if (!ok) {
    set_last_error(std::string("size mismatch for ") + name);
    return 5;
  }
  if (s.kind == PK_IGNORE) return 0;
  const float* src = data;
  if (!data_on_device) {
    if (m->stage_bytes < numel * sizeof(float)) {
      STA_CHECK_CUDA(cudaDeviceSynchronize());
      if (m->stage) STA_CHECK_CUDA(cudaFree(m->stage));
      m->stage = nullptr;
      m->stage_bytes = 0;
      size_t want = numel * sizeof(float);
      if (want < (static_cast<size_t>(32) << 20)) want = static_cast<size_t>(32) << 20;
      STA_CHECK_CUDA(cudaMalloc(&m->stage, want));
      m->stage_bytes = want;
    }
    STA_CHECK_CUDA(cudaMemcpy(m->stage, data, numel * sizeof(float), cudaMemcpyHostToDevice));
    src = m->stage;
  }
  const int threads = 256;
  const int blocks = static_cast<int>((numel + threads - 1) / threads);
  switch (s.kind) {
    case PK_F32:
      STA_CHECK_CUDA(cudaMemcpy(s.dst, src, numel * sizeof(float), cudaMemcpyDeviceToDevice));
      break;
    case PK_LINEAR: {
      const int rows = static_cast<int>(s.shape[0]);
      const int cols = static_cast<int>(numel / rows);
      pack_linear_kernel<<<blocks, threads>>>(src, static_cast<bf16*>(s.dst), rows, cols, s.ldd, m->split);
      break;
    }
    case PK_CONV3:
      pack_conv3_kernel<<<blocks, threads>>>(src, static_cast<bf16*>(s.dst), static_cast<int>(s.shape[0]),
                                             static_cast<int>(s.shape[1]), s.cin_pad, m->split);
      break;
    case PK_CONVT:
      pack_convT_kernel<<<blocks, threads>>>(src, static_cast<bf16*>(s.dst), static_cast<int>(s.shape[0]),
                                             static_cast<int>(s.shape[1]), s.k, s.cin_pad, s.cout_pad, m->split);
      break;
    case PK_TRANSPOSE_F32:
      transpose_f32_kernel<<<blocks, threads>>>(src, static_cast<float*>(s.dst), static_cast<int>(s.shape[0]),
                                                static_cast<int>(numel / s.shape[0]));
      break;
    default:
      break;
  }
  STA_CHECK_CUDA(cudaGetLastError());
  STA_CHECK_CUDA(cudaDeviceSynchronize());  // the staging buffer is reused by the next call
  if (!s.loaded) {
    s.loaded = true;
    m->missing--;
  }
  return 0;
}

int sta_encode(StaModel* m, const void* img_dev, int img_is_bf16, int B, int H, int W, float* feat_out_dev,
               int64_t* pos_out_dev, void* stream) {
  RUN(check_ready(m));
  STA_REQUIRE(B > 0 && H % 16 == 0 && W % 16 == 0 && H > 0 && W > 0, "image size must be a positive multiple of 16");
  STA_REQUIRE(H / 16 <= 1024 && W / 16 <= 1024, "token grid exceeds the RoPE table");
  Ctx c{m, static_cast<cudaStream_t>(stream)};
  RUN(enter_stream(m, c.st));
  const int h = H / 16, w = W / 16, N = h * w;
  RUN(ensure_ws(m, B, h, w));
  EncBufs e = take_enc(m->ws, B, N);
  m->launches += 1;
  RUN(launch_patch_im2col(img_dev, img_is_bf16, B, H, W, e.patches, c.st, m->split));
  // launch-bound sizes: the encoder proper is replayed as a CUDA graph on workspace buffers, the features are
  // then copied to the caller's tensor (which is long-lived: slam.py:145 keeps it for the whole run)
  const bool graphed = static_cast<long long>(B) * N <= kGraphMaxTokens;
  float* x = graphed ? m->ws.take<float>(static_cast<size_t>(B) * N * kEncDim) : feat_out_dev;
  auto body = [&](const Ctx& cc) -> int {
    m->launches += 1;
    RUN(launch_make_positions(e.pos, B, h, w, 0, cc.st));
    return run_encoder(cc, B, N, x, e);
  };
  if (graphed) {
    RUN(run_graphed(m, c.st, {1, B, H, W}, body));
    STA_CHECK_CUDA(cudaMemcpyAsync(feat_out_dev, x, static_cast<size_t>(B) * N * kEncDim * sizeof(float),
                                   cudaMemcpyDeviceToDevice, c.st));
  } else {
    RUN(body(c));
  }
  if (pos_out_dev) {
    const long long n = 2LL * B * N;
    m->launches++;
    pos_to_int64_kernel<<<static_cast<int>((n + 255) / 256), 256, 0, c.st>>>(
        e.pos, n, reinterpret_cast<long long*>(pos_out_dev));
    STA_CHECK_CUDA(cudaGetLastError());
  }
  return 0;
}

int sta_decode(StaModel* m, const float* feat1_dev, const float* feat2_dev, const int64_t* pos1_dev,
               const int64_t* pos2_dev, int B, int N, float* const* out1_dev, float* const* out2_dev, void* stream) {
  RUN(check_ready(m));
  STA_REQUIRE(B > 0 && N > 0, "empty batch");
  Ctx c{m, static_cast<cudaStream_t>(stream)};
  RUN(enter_stream(m, c.st));
  // workspace sized as an (N x 1) token grid: only token counts matter for the decoder buffers
  RUN(ensure_ws(m, 2 * B, N, 1));
  DecBufs d = take_dec(m->ws, 2 * B, N);
  m->launches += 3;
  RUN(launch_cast_f32_bf16(feat1_dev, d.enc_bf16, static_cast<long long>(B) * N, kEncDim, 0, c.st, m->split));
  RUN(launch_cast_f32_bf16(feat2_dev, d.enc_bf16 + static_cast<size_t>(B) * N * kEncDim * m->bmul(),
                           static_cast<long long>(B) * N, kEncDim, 0, c.st, m->split));
  {
    const long long total = 2LL * B * (N + 1);
    build_dec_pos_kernel<<<static_cast<int>((total + 255) / 256), 256, 0, c.st>>>(
        reinterpret_cast<const long long*>(pos1_dev), reinterpret_cast<const long long*>(pos2_dev), B, N, d.pos);
    STA_CHECK_CUDA(cudaGetLastError());
  }
  return run_decoder(c, B, N, d, out1_dev, out2_dev);
}

int sta_head_pose(StaModel* m, const float* tok_dev, int B, float* pose_out_dev, float* conf_out_dev, void* stream) {
  RUN(check_ready(m));
  m->launches++;
  return launch_pose_head(tok_dev, kDecDim, B, 0, kLnEps, m->pose, pose_out_dev, conf_out_dev,
                          static_cast<cudaStream_t>(stream));
}

int sta_head_pts(StaModel* m, const float* enc_feat_dev, const float* dec6_dev, const float* dec9_dev,
                 const float* dec12_dev, int B, int H, int W, float* pts3d_out_dev, float* conf_out_dev,
                 void* stream) {
  RUN(check_ready(m));
  STA_REQUIRE(B > 0 && H % 16 == 0 && W % 16 == 0, "image size must be a multiple of 16");
  Ctx c{m, static_cast<cudaStream_t>(stream)};
  RUN(enter_stream(m, c.st));
  const int h = H / 16, w = W / 16, N = h * w;
  RUN(ensure_ws(m, B, h, w));
  Workspace& ws = m->ws;
  const size_t T = static_cast<size_t>(B) * N;
  bf16* k0 = ws.takeb(T * 1024);
  bf16* k1 = ws.takeb(T * 768);
  bf16* k2 = ws.takeb(T * 768);
  bf16* k3 = ws.takeb(T * 768);
  m->launches += 4;
  RUN(launch_cast_f32_bf16(enc_feat_dev, k0, T, 1024, 0, c.st, m->split));
  RUN(launch_cast_f32_bf16(dec6_dev, k1, T, 768, 0, c.st, m->split));
  RUN(launch_cast_f32_bf16(dec9_dev, k2, T, 768, 0, c.st, m->split));
  RUN(launch_cast_f32_bf16(dec12_dev, k3, T, 768, 0, c.st, m->split));
  return run_dpt(c, ws, B, h, w, k0, k1, k2, k3, pts3d_out_dev, conf_out_dev);
}

int64_t sta_graph_replays(StaModel* m) { return m ? m->graph_replays : 0; }

int sta_regress_pairs(StaModel* m, const float* feat_i_dev, const float* feat_j_dev, int K, int H, int W,
                      float* pose_out_dev, float* pose_conf_out_dev, float* pts3d_out_dev, float* conf_out_dev,
                      float* intri_out_dev, float* depth_out_dev, float* conf_mean_out_dev, void* scratch, void* stream) {
  RUN(check_ready(m));
  STA_REQUIRE(K > 0 && K <= m->max_pairs_per_chunk, "edge batch must be in [1, 16]");
  STA_REQUIRE(H % 16 == 0 && W % 16 == 0 && H > 0 && W > 0, "image size must be a positive multiple of 16");
  STA_REQUIRE(H / 16 <= 1024 && W / 16 <= 1024, "token grid exceeds the RoPE table");
  STA_REQUIRE(feat_i_dev && feat_j_dev && pose_out_dev && pose_conf_out_dev && pts3d_out_dev && conf_out_dev,
              "null pointer");
  Ctx c{m, static_cast<cudaStream_t>(stream)};
  RUN(enter_stream(m, c.st));
  const int h = H / 16, w = W / 16, N = h * w, S = 2 * K, M = N + 1;
  RUN(ensure_ws(m, S, h, w));
  Workspace& ws = m->ws;
  DecBufs d = take_dec(ws, S, N);
  const bool consumers = intri_out_dev || depth_out_dev || conf_mean_out_dev;
  if (consumers)
    STA_REQUIRE(scratch != nullptr && intri_out_dev != nullptr,
                "the pointmap consumers need the scratch buffer and intri_out (depth_out / conf_mean_out are optional)");
  const long long px = static_cast<long long>(H) * W;
  // Launch-bound sizes: everything after the input casts is replayed as a CUDA graph that writes workspace
  // buffers; the results are then copied to the caller's tensors.  Large batches write the caller's tensors directly.
  const bool graphed = static_cast<long long>(S) * M <= kGraphMaxTokens;
  float *o_pose = pose_out_dev, *o_pconf = pose_conf_out_dev, *o_pts = pts3d_out_dev, *o_conf = conf_out_dev;
  float *o_intri = intri_out_dev, *o_depth = depth_out_dev, *o_cmean = conf_mean_out_dev;
  void* o_scratch = scratch;
  if (graphed) {
    o_pose = ws.take<float>(static_cast<size_t>(S) * 16);
    o_pconf = ws.take<float>(S);
    o_pts = ws.take<float>(static_cast<size_t>(S) * px * 3);
    o_conf = ws.take<float>(static_cast<size_t>(S) * px);
    if (consumers) {
      o_intri = ws.take<float>(static_cast<size_t>(K) * 9);
      o_depth = depth_out_dev ? ws.take<float>(static_cast<size_t>(S) * px) : nullptr;
      o_cmean = conf_mean_out_dev ? ws.take<float>(S) : nullptr;
      o_scratch = ws.take<double>(pointmap_scratch_bytes(S) / sizeof(double));
    }
  }
  m->launches += 2;
  const size_t bm = m->bmul();
  RUN(launch_cast_f32_bf16(feat_i_dev, d.enc_bf16, static_cast<long long>(K) * N, kEncDim, 0, c.st, m->split));
  RUN(launch_cast_f32_bf16(feat_j_dev, d.enc_bf16 + static_cast<size_t>(K) * N * kEncDim * bm, static_cast<long long>(K) * N,
                           kEncDim, 0, c.st, m->split));
  auto body = [&](const Ctx& cc) -> int {
    m->launches += 3;
    RUN(launch_make_positions(d.pos, S, h, w, 1, cc.st));
    RUN(run_decoder(cc, K, N, d, nullptr, nullptr));
    RUN(launch_pose_head(d.xd, static_cast<long long>(M) * kDecDim, K, 1, kLnEps, m->pose, o_pose, o_pconf, cc.st));
    RUN(launch_pose_head(d.xd + static_cast<long long>(K) * M * kDecDim, static_cast<long long>(M) * kDecDim, K, 1,
                         kLnEps, m->pose, o_pose + static_cast<long long>(K) * 16, o_pconf + K, cc.st));
    for (int v = 0; v < 2; ++v) {
      const size_t save = ws.off;
      const size_t tok0 = static_cast<size_t>(v) * K * N;
      RUN(run_dpt(cc, ws, K, h, w, d.enc_bf16 + tok0 * 1024 * bm, d.hook[0] + tok0 * 768 * bm, d.hook[1] + tok0 * 768 * bm,
                  d.hook[2] + tok0 * 768 * bm, o_pts + static_cast<long long>(v) * K * px * 3,
                  o_conf + static_cast<long long>(v) * K * px));
      ws.off = save;
    }
    if (consumers) {
      m->launches += 2;
      RUN(launch_pointmap_consumers(o_pts, o_conf, S, H, W, 2, o_intri, o_depth, o_cmean, o_scratch, cc.st));
    }
    return 0;
  };
  if (!graphed) return body(c);
  RUN(run_graphed(m, c.st, {2, K, H, W, consumers ? 1 : 0, depth_out_dev ? 1 : 0, conf_mean_out_dev ? 1 : 0}, body));
  auto copy = [&](float* dst, const float* src, size_t n) -> int {
    if (dst) STA_CHECK_CUDA(cudaMemcpyAsync(dst, src, n * sizeof(float), cudaMemcpyDeviceToDevice, c.st));
    return 0;
  };
  RUN(copy(pose_out_dev, o_pose, static_cast<size_t>(S) * 16));
  RUN(copy(pose_conf_out_dev, o_pconf, S));
  RUN(copy(pts3d_out_dev, o_pts, static_cast<size_t>(S) * px * 3));
  RUN(copy(conf_out_dev, o_conf, static_cast<size_t>(S) * px));
  if (consumers) {
    RUN(copy(intri_out_dev, o_intri, static_cast<size_t>(K) * 9));
    RUN(copy(depth_out_dev, o_depth, static_cast<size_t>(S) * px));
    RUN(copy(conf_mean_out_dev, o_cmean, S));
  }
  return 0;
}

// ---------------------------------------------------------------------------
// Gated keyframe step (SURVEY.md 8(f) rank 1 with the early-out of slam.py:169-170): phase 1 runs the decoder and the
// pose heads for all K candidate edges and leaves the decoder hooks in the workspace; the caller reads the K pose
// confidences ONCE, decides which edges survive (`conf >= thres or i - j == 1`) and phase 2 runs the DPT heads and the
// pointmap consumers for the survivors only.  Nothing else may run on this handle between the two phases.
// ---------------------------------------------------------------------------
int sta_regress_pairs_begin(StaModel* m, const float* feat_i_dev, const float* feat_j_dev, int K, int H, int W,
                            float* pose_out_dev, float* pose_conf_out_dev, void* stream) {
  RUN(check_ready(m));
  m->rp_K = 0;
  STA_REQUIRE(K > 0 && K <= m->max_pairs_per_chunk, "edge batch must be in [1, 16]");
  STA_REQUIRE(H % 16 == 0 && W % 16 == 0 && H > 0 && W > 0, "image size must be a positive multiple of 16");
  STA_REQUIRE(H / 16 <= 1024 && W / 16 <= 1024, "token grid exceeds the RoPE table");
  STA_REQUIRE(feat_i_dev && feat_j_dev && pose_out_dev && pose_conf_out_dev, "null pointer");
  Ctx c{m, static_cast<cudaStream_t>(stream)};
  RUN(enter_stream(m, c.st));
  const int h = H / 16, w = W / 16, N = h * w, S = 2 * K, M = N + 1;
  RUN(ensure_ws(m, S, h, w));
  Workspace& ws = m->ws;
  DecBufs d = take_dec(ws, S, N);
  float* o_pose = ws.take<float>(static_cast<size_t>(S) * 16);
  float* o_pconf = ws.take<float>(S);
  const size_t bm = m->bmul();
  m->launches += 2;
  RUN(launch_cast_f32_bf16(feat_i_dev, d.enc_bf16, static_cast<long long>(K) * N, kEncDim, 0, c.st, m->split));
  RUN(launch_cast_f32_bf16(feat_j_dev, d.enc_bf16 + static_cast<size_t>(K) * N * kEncDim * bm, static_cast<long long>(K) * N,
                           kEncDim, 0, c.st, m->split));
  auto body = [&](const Ctx& cc) -> int {
    m->launches += 3;
    RUN(launch_make_positions(d.pos, S, h, w, 1, cc.st));
    RUN(run_decoder(cc, K, N, d, nullptr, nullptr));
    RUN(launch_pose_head(d.xd, static_cast<long long>(M) * kDecDim, K, 1, kLnEps, m->pose, o_pose, o_pconf, cc.st));
    RUN(launch_pose_head(d.xd + static_cast<long long>(K) * M * kDecDim, static_cast<long long>(M) * kDecDim, K, 1, kLnEps,
                         m->pose, o_pose + static_cast<long long>(K) * 16, o_pconf + K, cc.st));
    return 0;
  };
  if (static_cast<long long>(S) * M <= kGraphMaxTokens) RUN(run_graphed(m, c.st, {3, K, H, W}, body));
  else RUN(body(c));
  STA_CHECK_CUDA(cudaMemcpyAsync(pose_out_dev, o_pose, static_cast<size_t>(S) * 16 * sizeof(float), cudaMemcpyDeviceToDevice, c.st));
  STA_CHECK_CUDA(cudaMemcpyAsync(pose_conf_out_dev, o_pconf, static_cast<size_t>(S) * sizeof(float), cudaMemcpyDeviceToDevice, c.st));
  m->rp_K = K;
  m->rp_H = H;
  m->rp_W = W;
  return 0;
}

int sta_regress_pairs_finish(StaModel* m, const int* edge_idx_host, int n_sel, float* pts3d_out_dev, float* conf_out_dev,
                             float* intri_out_dev, float* depth_out_dev, float* conf_mean_out_dev, void* scratch,
                             void* stream) {
  RUN(check_ready(m));
  STA_REQUIRE(m->rp_K > 0, "sta_regress_pairs_finish without a preceding sta_regress_pairs_begin on this handle");
  const int K = m->rp_K, H = m->rp_H, W = m->rp_W;
  STA_REQUIRE(n_sel >= 0 && n_sel <= K, "more selected edges than candidates");
  if (n_sel == 0) return 0;
  STA_REQUIRE(edge_idx_host && pts3d_out_dev && conf_out_dev, "null pointer");
  for (int k = 0; k < n_sel; ++k) STA_REQUIRE(edge_idx_host[k] >= 0 && edge_idx_host[k] < K, "edge index out of range");
  const bool consumers = intri_out_dev || depth_out_dev || conf_mean_out_dev;
  if (consumers) STA_REQUIRE(scratch != nullptr && intri_out_dev != nullptr, "the pointmap consumers need scratch and intri_out");
  Ctx c{m, static_cast<cudaStream_t>(stream)};
  RUN(enter_stream(m, c.st));
  const int h = H / 16, w = W / 16, N = h * w, S = 2 * K, S2 = 2 * n_sel;
  RUN(ensure_ws(m, S, h, w));  // same shape as phase 1: no reallocation, and the same take sequence gives the same buffers
  Workspace& ws = m->ws;
  DecBufs d = take_dec(ws, S, N);
  ws.take<float>(static_cast<size_t>(S) * 16);
  ws.take<float>(S);
  const size_t bm = m->bmul();
  const long long px = static_cast<long long>(H) * W;
  // compact copies of the survivors' hooks: [view][edge] order, view 0 = the (i -> j) direction
  bf16* sel[4];
  const size_t width[4] = {1024 * bm, 768 * bm, 768 * bm, 768 * bm};
  const bf16* src[4] = {d.enc_bf16, d.hook[0], d.hook[1], d.hook[2]};
  for (int q = 0; q < 4; ++q) sel[q] = ws.take<bf16>(static_cast<size_t>(S2) * N * width[q]);
  float* o_pts = ws.take<float>(static_cast<size_t>(S2) * px * 3);
  float* o_conf = ws.take<float>(static_cast<size_t>(S2) * px);
  float* o_intri = consumers ? ws.take<float>(static_cast<size_t>(n_sel) * 9) : nullptr;
  float* o_depth = depth_out_dev ? ws.take<float>(static_cast<size_t>(S2) * px) : nullptr;
  float* o_cmean = conf_mean_out_dev ? ws.take<float>(S2) : nullptr;
  void* o_scratch = consumers ? ws.take<double>(pointmap_scratch_bytes(S2) / sizeof(double)) : nullptr;
  STA_REQUIRE(ws.off <= ws.bytes, "internal: workspace too small for the gated keyframe step");
  for (int v = 0; v < 2; ++v)
    for (int k = 0; k < n_sel; ++k)
      for (int q = 0; q < 4; ++q) {
        const size_t rows = static_cast<size_t>(N) * width[q];
        STA_CHECK_CUDA(cudaMemcpyAsync(sel[q] + (static_cast<size_t>(v) * n_sel + k) * rows,
                                       src[q] + (static_cast<size_t>(v) * K + edge_idx_host[k]) * rows, rows * sizeof(bf16),
                                       cudaMemcpyDeviceToDevice, c.st));
      }
  auto body = [&](const Ctx& cc) -> int {
    for (int v = 0; v < 2; ++v) {
      const size_t save = ws.off;
      const size_t tok0 = static_cast<size_t>(v) * n_sel * N;
      RUN(run_dpt(cc, ws, n_sel, h, w, sel[0] + tok0 * width[0], sel[1] + tok0 * width[1], sel[2] + tok0 * width[2],
                  sel[3] + tok0 * width[3], o_pts + static_cast<long long>(v) * n_sel * px * 3,
                  o_conf + static_cast<long long>(v) * n_sel * px));
      ws.off = save;
    }
    if (consumers) {
      m->launches += 2;
      RUN(launch_pointmap_consumers(o_pts, o_conf, S2, H, W, 2, o_intri, o_depth, o_cmean, o_scratch, cc.st));
    }
    return 0;
  };
  if (static_cast<long long>(S) * (N + 1) <= kGraphMaxTokens)
    RUN(run_graphed(m, c.st, {4, K, n_sel, H, W, consumers ? 1 : 0, depth_out_dev ? 1 : 0, conf_mean_out_dev ? 1 : 0}, body));
  else
    RUN(body(c));
  auto copy = [&](float* dst, const float* srcp, size_t n) -> int {
    if (dst) STA_CHECK_CUDA(cudaMemcpyAsync(dst, srcp, n * sizeof(float), cudaMemcpyDeviceToDevice, c.st));
    return 0;
  };
  RUN(copy(pts3d_out_dev, o_pts, static_cast<size_t>(S2) * px * 3));
  RUN(copy(conf_out_dev, o_conf, static_cast<size_t>(S2) * px));
  if (consumers) {
    RUN(copy(intri_out_dev, o_intri, static_cast<size_t>(n_sel) * 9));
    RUN(copy(depth_out_dev, o_depth, static_cast<size_t>(S2) * px));
    RUN(copy(conf_mean_out_dev, o_cmean, S2));
  }
  return 0;
}

int sta_forward_pairs(StaModel* m, const void* img1_dev, const void* img2_dev, int img_is_bf16, int B, int H, int W,
                      float* pts3d_out_dev, float* conf_out_dev, float* pose_out_dev, float* pose_conf_out_dev,
                      void* stream) {
  RUN(check_ready(m));
  STA_REQUIRE(B > 0 && H % 16 == 0 && W % 16 == 0 && H > 0 && W > 0, "image size must be a positive multiple of 16");
  STA_REQUIRE(H / 16 <= 1024 && W / 16 <= 1024, "token grid exceeds the RoPE table");
  Ctx c{m, static_cast<cudaStream_t>(stream)};
  RUN(enter_stream(m, c.st));
  const size_t esz = img_is_bf16 ? 2 : 4;
  const long long px = static_cast<long long>(H) * W;
  for (int b0 = 0; b0 < B; b0 += m->max_pairs_per_chunk) {
    const int nb = (B - b0 < m->max_pairs_per_chunk) ? (B - b0) : m->max_pairs_per_chunk;
    const char* i1 = static_cast<const char*>(img1_dev) + static_cast<size_t>(b0) * 3 * px * esz;
    const char* i2 = static_cast<const char*>(img2_dev) + static_cast<size_t>(b0) * 3 * px * esz;
    RUN(forward_chunk(c, i1, i2, img_is_bf16, nb, H, W, pts3d_out_dev + static_cast<long long>(b0) * px * 3,
                      conf_out_dev + static_cast<long long>(b0) * px, pose_out_dev + static_cast<long long>(b0) * 16,
                      pose_conf_out_dev + b0, B));
  }
  return 0;
}

int sta_forward_pairs_host(StaModel* m, const void* img1_host, const void* img2_host, int img_is_bf16, int B, int H,
                           int W, float* pts3d_out_host, float* conf_out_host, float* pose_out_host,
                           float* pose_conf_out_host, void* stream) {
  RUN(check_ready(m));
  STA_REQUIRE(B > 0 && H % 16 == 0 && W % 16 == 0 && H > 0 && W > 0, "image size must be a positive multiple of 16");
  STA_REQUIRE(H / 16 <= 1024 && W / 16 <= 1024, "token grid exceeds the RoPE table");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  RUN(enter_stream(m, st));
  const size_t esz = img_is_bf16 ? 2 : 4;
  const size_t px = static_cast<size_t>(H) * W;
  const size_t img_bytes = static_cast<size_t>(B) * 3 * px * esz;
  const size_t pts_bytes = 2 * static_cast<size_t>(B) * px * 3 * sizeof(float);
  const size_t conf_bytes = 2 * static_cast<size_t>(B) * px * sizeof(float);
  const size_t pose_bytes = 2 * static_cast<size_t>(B) * 16 * sizeof(float);
  const size_t pconf_bytes = 2 * static_cast<size_t>(B) * sizeof(float);
  auto up = [](size_t v) { return (v + 255) & ~static_cast<size_t>(255); };
  const size_t need = 2 * up(img_bytes) + up(pts_bytes) + up(conf_bytes) + up(pose_bytes) + up(pconf_bytes);
  if (m->io_bytes < need) {
    STA_CHECK_CUDA(cudaDeviceSynchronize());
    if (m->io) STA_CHECK_CUDA(cudaFree(m->io));
    m->io = nullptr;
    m->io_bytes = 0;
    STA_CHECK_CUDA(cudaMalloc(&m->io, need));
    m->io_bytes = need;
  }
  if (!m->s_in) {
    STA_CHECK_CUDA(cudaStreamCreateWithFlags(&m->s_in, cudaStreamNonBlocking));
    STA_CHECK_CUDA(cudaStreamCreateWithFlags(&m->s_out, cudaStreamNonBlocking));
    for (int i = 0; i < StaModel::kMaxHostChunks; ++i) {
      STA_CHECK_CUDA(cudaEventCreateWithFlags(&m->ev_in[i], cudaEventDisableTiming));
      STA_CHECK_CUDA(cudaEventCreateWithFlags(&m->ev_done[i], cudaEventDisableTiming));
      for (int k = 0; k < 2 * StaModel::kMaxParts; ++k)
        STA_CHECK_CUDA(cudaEventCreateWithFlags(&m->ev_part[i][k], cudaEventDisableTiming));
    }
    STA_CHECK_CUDA(cudaEventCreateWithFlags(&m->ev_start, cudaEventDisableTiming));
  }
  char* p = m->io;
  char* d_img1 = p; p += up(img_bytes);
  char* d_img2 = p; p += up(img_bytes);
  float* d_pts = reinterpret_cast<float*>(p); p += up(pts_bytes);
  float* d_conf = reinterpret_cast<float*>(p); p += up(conf_bytes);
  float* d_pose = reinterpret_cast<float*>(p); p += up(pose_bytes);
  float* d_pconf = reinterpret_cast<float*>(p);

  // Pipeline over chunks of pairs: the H2D copy of chunk c+1 and the D2H copy of chunk c-1 overlap the compute
  // of chunk c (separate copy streams, pinned host memory makes them truly asynchronous).
  // one chunk of up to 16 pairs measured best on B200 (448 vs 433 pairs/s for two chunks of 8 at B = 16: the exposed
  // H2D of a 16-pair chunk costs less than running the GEMMs at half the tile count); larger batches pipeline chunks
  int cp = B;
  {
    static int chunk_env = -1;  // STA_HOST_CHUNK=n overrides the pairs per pipelined chunk (tuning knob)
    if (chunk_env < 0) {
      const char* e = getenv("STA_HOST_CHUNK");
      chunk_env = e ? atoi(e) : 0;
    }
    if (chunk_env > 0) cp = chunk_env < B ? chunk_env : B;
  }
  if (cp > m->max_pairs_per_chunk) cp = m->max_pairs_per_chunk;
  int nchunks = (B + cp - 1) / cp;
  if (nchunks > StaModel::kMaxHostChunks) {
    nchunks = StaModel::kMaxHostChunks;
    cp = (B + nchunks - 1) / nchunks;
    STA_REQUIRE(cp <= m->max_pairs_per_chunk, "batch too large for the host entry point; split it");
    nchunks = (B + cp - 1) / cp;
  }
  // the copy streams must not start before previously enqueued work on `st` that may still use the io buffers
  STA_CHECK_CUDA(cudaEventRecord(m->ev_start, st));
  STA_CHECK_CUDA(cudaStreamWaitEvent(m->s_in, m->ev_start, 0));
  STA_CHECK_CUDA(cudaStreamWaitEvent(m->s_out, m->ev_start, 0));
  const size_t img_pair = 3 * px * esz;
  for (int c = 0; c < nchunks; ++c) {
    const int b0 = c * cp, nb = (B - b0 < cp) ? (B - b0) : cp;
    STA_CHECK_CUDA(cudaMemcpyAsync(d_img1 + b0 * img_pair, static_cast<const char*>(img1_host) + b0 * img_pair,
                                   nb * img_pair, cudaMemcpyHostToDevice, m->s_in));
    STA_CHECK_CUDA(cudaMemcpyAsync(d_img2 + b0 * img_pair, static_cast<const char*>(img2_host) + b0 * img_pair,
                                   nb * img_pair, cudaMemcpyHostToDevice, m->s_in));
    STA_CHECK_CUDA(cudaEventRecord(m->ev_in[c], m->s_in));
  }
  Ctx cx{m, st};
  for (int c = 0; c < nchunks; ++c) {
    const int b0 = c * cp, nb = (B - b0 < cp) ? (B - b0) : cp;
    STA_CHECK_CUDA(cudaStreamWaitEvent(st, m->ev_in[c], 0));
    RUN(forward_chunk(cx, d_img1 + b0 * img_pair, d_img2 + b0 * img_pair, img_is_bf16, nb, H, W,
                      d_pts + static_cast<size_t>(b0) * px * 3, d_conf + static_cast<size_t>(b0) * px,
                      d_pose + static_cast<size_t>(b0) * 16, d_pconf + b0, B, m->ev_part[c]));
    STA_CHECK_CUDA(cudaEventRecord(m->ev_done[c], st));
    const int halves = host_dpt_parts(nb);  // mirrors forward_chunk
    for (int v = 0; v < 2; ++v) {
      for (int hf = 0; hf < halves; ++hf) {
        const int i0 = hf * (nb / halves), ni = (hf == halves - 1) ? nb - i0 : nb / halves;
        STA_CHECK_CUDA(cudaStreamWaitEvent(m->s_out, m->ev_part[c][v * halves + hf], 0));
        const size_t o = static_cast<size_t>(v) * B + b0 + i0;
        STA_CHECK_CUDA(cudaMemcpyAsync(pts3d_out_host + o * px * 3, d_pts + o * px * 3, ni * px * 3 * sizeof(float),
                                       cudaMemcpyDeviceToHost, m->s_out));
        STA_CHECK_CUDA(cudaMemcpyAsync(conf_out_host + o * px, d_conf + o * px, ni * px * sizeof(float),
                                       cudaMemcpyDeviceToHost, m->s_out));
      }
    }
    STA_CHECK_CUDA(cudaStreamWaitEvent(m->s_out, m->ev_done[c], 0));
    for (int v = 0; v < 2; ++v) {
      const size_t o = static_cast<size_t>(v) * B + b0;
      STA_CHECK_CUDA(cudaMemcpyAsync(pose_out_host + o * 16, d_pose + o * 16, nb * 16 * sizeof(float),
                                     cudaMemcpyDeviceToHost, m->s_out));
      STA_CHECK_CUDA(cudaMemcpyAsync(pose_conf_out_host + o, d_pconf + o, nb * sizeof(float), cudaMemcpyDeviceToHost,
                                     m->s_out));
    }
  }
  STA_CHECK_CUDA(cudaStreamSynchronize(m->s_out));
  STA_CHECK_CUDA(cudaStreamSynchronize(st));
  return 0;
}

// ---------------------------------------------------------------------------
// op-level entry points (parity tests drive the same kernels the model uses)
// ---------------------------------------------------------------------------
int sta_op_gemm(const StaGemmDesc* d, void* stream) {
  if (!d) {
    set_last_error("sta_op_gemm: null descriptor");
    return 2;
  }
  GemmLaunch g;
  g.amode = d->conv3x3 ? A_CONV3 : A_LINEAR;
  g.epi = d->epi;
  g.A = static_cast<const bf16*>(d->A);
  g.lda = d->lda;
  g.Wt = static_cast<const bf16*>(d->W);
  g.ldw = d->ldw;
  g.splitk_ws = d->splitk_ws;
  g.splitk_ws_bytes = d->splitk_ws ? static_cast<size_t>(d->splitk_ws_bytes) : 0;
  GemmParams p = {};
  p.M = d->M; p.N = d->N; p.K = d->K;
  p.nimg = d->nimg; p.H = d->H; p.W = d->Wd; p.Cin = d->Cin;
  p.bias = d->bias;
  p.out = d->out; p.ldo = d->ldo; p.out2 = d->out2;
  p.resid = d->resid; p.resid2 = d->resid2;
  p.relu_main = d->relu_main;
  p.rowmap_n = d->rowmap_n;
  p.pos = d->pos; p.rope_cols = d->rope_cols;
  if (d->epi == EPI_ROPE) {
    p.rope_tab = rope_table(&p.rope_max_pos);
    if (!p.rope_tab) {
      set_last_error("failed to build the RoPE table");
      return 1;
    }
    p.rope_smem_rows = 64;
  }
  p.ps_k = d->ps_k; p.ps_cout = d->ps_cout; p.ps_h = d->ps_h; p.ps_w = d->ps_w;
  p.head_w = d->head_w; p.head_b = d->head_b; p.pts3d = d->pts3d; p.conf = d->conf;
  p.split = d->split_precision;
  g.p = p;
  return launch_gemm(g, static_cast<cudaStream_t>(stream));
}

int sta_op_attention(const void* q, int64_t ldq, int q_col0, const void* k, int64_t ldk, int k_col0, const void* v,
                     int64_t ldv, int v_col0, void* out, int64_t ldo, int batch, int heads, int nq, int nk,
                     int kv_batch_shift, float scale, int split_first_row, void* stream) {
  AttnLaunch a;
  a.q = static_cast<const bf16*>(q); a.ldq = ldq; a.q_col0 = q_col0;
  a.k = static_cast<const bf16*>(k); a.ldk = ldk; a.k_col0 = k_col0;
  a.v = static_cast<const bf16*>(v); a.ldv = ldv; a.v_col0 = v_col0;
  a.out = static_cast<bf16*>(out); a.ldo = ldo;
  a.batch = batch; a.heads = heads; a.nq = nq; a.nk = nk;
  a.kv_batch_shift = kv_batch_shift;
  a.scale = scale;
  a.split_first_row = split_first_row;
  return launch_attention(a, static_cast<cudaStream_t>(stream));
}

int sta_op_layernorm(const float* x, int rows, int C, float eps, const float* g1, const float* b1, void* out1_bf16,
                     const float* g2, const float* b2, void* out2_bf16, int drop_first_of, void* stream) {
  return launch_layernorm(x, rows, C, eps, g1, b1, static_cast<bf16*>(out1_bf16), g2, b2,
                          static_cast<bf16*>(out2_bf16), drop_first_of, static_cast<cudaStream_t>(stream));
}
int sta_op_patch_im2col(const void* img, int img_is_bf16, int B, int H, int W, void* out_bf16, void* stream) {
  return launch_patch_im2col(img, img_is_bf16, B, H, W, static_cast<bf16*>(out_bf16), static_cast<cudaStream_t>(stream));
}
int sta_op_upsample2x(const void* in_bf16, void* out_bf16, int nimg, int H, int W, int C, void* stream) {
  return launch_upsample2x(static_cast<const bf16*>(in_bf16), static_cast<bf16*>(out_bf16), nimg, H, W, C, 2 * H, 2 * W,
                           static_cast<cudaStream_t>(stream));
}
int sta_op_im2col_3x3_s2(const void* in_bf16, void* out_bf16, int nimg, int H, int W, int C, void* stream) {
  return launch_im2col_3x3_s2(static_cast<const bf16*>(in_bf16), static_cast<bf16*>(out_bf16), nimg, H, W, C,
                              static_cast<cudaStream_t>(stream));
}
int sta_op_cast_f32_bf16(const float* in, void* out_bf16, int64_t rows, int C, int drop_first_of, void* stream) {
  return launch_cast_f32_bf16(in, static_cast<bf16*>(out_bf16), rows, C, drop_first_of,
                              static_cast<cudaStream_t>(stream));
}
int sta_op_rope2d(void* tokens_bf16, const int64_t* pos, int B, int N, int H, void* stream) {
  return launch_rope2d(static_cast<bf16*>(tokens_bf16), reinterpret_cast<const long long*>(pos), B, N, H,
                       static_cast<cudaStream_t>(stream));
}

// ---------------------------------------------------------------------------
// pointmap consumers (pointmap.cu)
// ---------------------------------------------------------------------------
size_t sta_pointmap_scratch_bytes(int V) { return pointmap_scratch_bytes(V); }
int sta_pointmap_consumers(const float* pts3d, const float* conf, int V, int H, int W, int shared, float* K_out,
                           float* depth_out, float* conf_mean_out, void* scratch, void* stream) {
  return launch_pointmap_consumers(pts3d, conf, V, H, W, shared, K_out, depth_out, conf_mean_out, scratch,
                                   static_cast<cudaStream_t>(stream));
}
int sta_depth_scale(const float* Di, const float* Dj, const float* ci, const float* cj, int64_t n, float* out2,
                    void* scratch, void* stream) {
  return launch_depth_scale(Di, Dj, ci, cj, static_cast<long long>(n), out2, scratch, static_cast<cudaStream_t>(stream));
}

// ---------------------------------------------------------------------------
// Sim(3) pose-graph LM step (pose_graph.cu)
// ---------------------------------------------------------------------------
size_t sta_pose_graph_scratch_bytes(int num_nodes, int num_edges, int num_opt) {
  return pose_graph_scratch_bytes(num_nodes, num_edges, num_opt);
}
int sta_pose_graph_lm_step(const float* nodes, int num_nodes, const int64_t* edges, const float* meas, const float* weights,
                           int num_edges, const int64_t* opt_idx, int num_opt, double damping, double dmin, double dmax,
                           float* nodes_out, double* info_out, void* scratch, void* stream) {
  return launch_pose_graph_lm_step(nodes, num_nodes, reinterpret_cast<const long long*>(edges), meas, weights, num_edges,
                                   reinterpret_cast<const long long*>(opt_idx), num_opt, damping, dmin, dmax, nodes_out,
                                   info_out, scratch, static_cast<cudaStream_t>(stream));
}

// ---------------------------------------------------------------------------
// image preprocessing (preprocess.cu)
// ---------------------------------------------------------------------------
int sta_preprocess_shape(int H, int W, int res_w, int res_h, int w_edge, int h_edge, int* out_hw) {
  return launch_preprocess_rgb8(nullptr, H, W, res_w, res_h, w_edge, h_edge, nullptr, nullptr, nullptr, out_hw, 1, nullptr);
}
int sta_preprocess_geometry(int H, int W, int res_w, int res_h, int w_edge, int h_edge, int* out10) {
  if (!out10) {
    set_last_error("sta_preprocess_geometry: null output");
    return 2;
  }
  return preprocess_geometry(H, W, res_w, res_h, w_edge, h_edge, out10);
}
int sta_preprocess_coeffs(int in_size, int out_size, int* ksize, int* bounds, int* kk, int64_t kk_capacity) {
  return preprocess_coeffs(in_size, out_size, ksize, bounds, kk, static_cast<long long>(kk_capacity));
}
int sta_preprocess_rgb8(const uint8_t* rgb_dev, int H, int W, int res_w, int res_h, int w_edge, int h_edge,
                        float* rgb_out_dev, float* gray_out_dev, uint8_t* u8_out_dev, void* stream) {
  return launch_preprocess_rgb8(rgb_dev, H, W, res_w, res_h, w_edge, h_edge, rgb_out_dev, gray_out_dev, u8_out_dev,
                                nullptr, 0, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
