// Persistent warp-specialised tcgen05 GEMM / implicit-GEMM 3x3 convolution for sm_100a.
//
//   D[M,N] = A[M,K] * W[N,K]^T  (bf16 operands, fp32 accumulation in TMEM)
//
// Roles: warp 0 = TMA producer, warp 1 = MMA issuer (one elected thread issues tcgen05.mma),
// warp 2 = TMEM allocator, warps 4.. = 8 or 16 epilogue warps
// (TMEM -> registers -> fused epilogue -> global).  The accumulator is double
// buffered in TMEM (2 x BN columns) so the epilogue of tile i overlaps the
// mainloop of tile i+1.  Operand tiles are 128B-swizzled K-major [rows][64]
// boxes written by TMA; the 3x3 convolution uses a 4-D NHWC tensor map and
// shifts the box per filter tap (out-of-bounds = zero padding), so no im2col
// buffer exists.
//
// Reference ops covered (vista_slam/sta_model): nn.Linear in blocks/sta_blocks.py:67-70,
// 88-90,179-183; PatchEmbed conv sta_blocks.py:262 (after im2col); the DPT convs of
// heads/dpt_block.py:33-68,93-111,174-182,319-323,356-403.
#pragma once
#include "common.cuh"

namespace sta {

// A_CONV3:  3x3 convolution, one TMA box (8 x 16 pixels x 64 channels, shifted by the tap) per filter tap and channel chunk.
// A_CONV3H: 3x3 convolution, halo-staged: ONE box of 18 x 16 pixels x 64 channels per channel chunk feeds all nine taps --
//           a tap is a row-shifted view of the staged tile (tile 16 rows x 8 pixels, 8-row descriptor groups 2048 bytes
//           apart, descriptor base offset = the tap's column shift).  Shared-memory write traffic of the A operand drops from
//           9 x 16 KB to 36 KB per channel chunk; the N = 128 head convolutions were bound by exactly that traffic.
enum AMode : int { A_LINEAR = 0, A_CONV3 = 1, A_CONV3H = 2 };
__host__ __device__ constexpr bool is_conv(int amode) { return amode != A_LINEAR; }
enum Epi : int {
  EPI_BF16 = 0,     // out_bf16 = [relu](acc + bias [+ resid_bf16] [+ resid2_bf16]); optional relu copy in out2
  EPI_GELU = 1,     // out_bf16 = gelu_erf(acc + bias)
  EPI_F32 = 2,      // out_f32[row'] = acc + bias [+ resid_f32[row']]   (row' optionally skips one pose-token row per sample)
  EPI_ROPE = 3,     // out_bf16 = rope2d(acc + bias) on columns < rope_cols, plain on the rest
  EPI_PIXSHUF = 4,  // ConvTranspose2d(k = stride): scatter to the k x k sub-pixels, NHWC bf16
  EPI_HEAD = 5,     // relu(acc + bias) -> 1x1 conv (128 -> 4) -> pointmap/confidence post-process (fp32 out)
};

struct GemmParams {
  int M, N, K;
  // A_CONV3 geometry (input and output NHWC, same H x W)
  int nimg, H, W, Cin, tiles_h, tiles_w;
  // epilogue
  const float* bias;
  void* out;
  long long ldo;  // output row stride in elements (also residual row stride)
  void* out2;
  const void* resid;
  const void* resid2;
  int relu_main;
  int rowmap_n;  // EPI_F32: if > 0, out row = (r / n) * (n + 1) + 1 + r % n
  // EPI_ROPE
  const float* rope_tab;  // [(pos + 1)][16] x (cos, sin)
  const int* pos;         // [M][2] (y, x)
  int rope_cols;
  int rope_max_pos;
  int rope_smem_rows;  // > 0: table rows for positions [-1, rope_smem_rows - 2] are staged in shared memory
  // EPI_PIXSHUF
  int ps_k, ps_cout, ps_h, ps_w;
  // EPI_HEAD
  const float* head_w;  // [128][4]
  const float* head_b;  // [4]
  float* pts3d;         // [pixels][3]
  float* conf;          // [pixels]
  // split-K (TMA epilogue, EPI_F32 only): a work item is (m, n, ks); split ks accumulates k-blocks
  // [ks * nkb / ksplit, (ks + 1) * nkb / ksplit) and stores its fp32 partial tile at output row ks * split_rows + row
  int ksplit;      // >= 1
  int split_rows;  // rows of one partial slice (multiple of 128)
  int c_reduce;         // TMA epilogue, EPI_F32: out += result (in-place fp32 residual stream) via bulk reduce-add
  // split-precision parity mode (common.cuh): the operands are already (hi | lo | hi) x (hi | hi | lo) expanded along K
  // (K here is the physical 3K); bf16 outputs / bf16 skip tensors use the (hi | lo | hi) row layout with logical
  // width N (ldo = 3N).  fp32 outputs are unaffected.
  int split;
  // Wave-quantisation tail (TMA epilogue, CTA pair, BN = 256, EPI_F32, no split-K): the last tail_r work items, which would
  // run as a partial wave on a few CTA pairs, are each issued as TWO 256 x 128 half tiles (see launch_gemm).
  int tail_first;  // first tile index of the tail
  int tail_r;      // number of tail tiles (0 = off)
  long long* dbg;       // optional clock64 trace (env STA_GEMM_TRACE), else null
};

// CG = 1: one CTA per 128 x BN tile (tcgen05 cta_group::1).
// CG = 2: a CTA pair (cluster of 2 on one TPC) per 256 x BN tile (cta_group::2): each CTA stages its own
//         128 rows of A and BN/2 rows of B, so per-SM shared-memory traffic per MMA is halved for B.
// TMA = true: the epilogue writes through shared memory + TMA tensor stores (epilogue_tile_tma), else the
// transposing per-warp staging path (epilogue_tile).
// A_CONV3H: pixels per staged halo row.  10 = the 8-pixel tile plus its two halo columns (dense box, 8-row descriptor groups
// 1280 B apart -- groups then start at any 128-byte multiple, which the absolute-address swizzle permits); 16 = padded rows
// (groups 2048 B apart), 60 % more bytes per box.
constexpr int kHaloPitch = 10;
constexpr int kRopeLd = 36;  // floats per staged sin/cos table row (32 + 4 pad: conflict-free row-per-lane reads)
template <int BN, int CG, int EW, int EPI, bool TMA, bool HALO = false>
struct GemmCfg {
  static constexpr int BM = 128;
  static constexpr int BK = 64;
  static constexpr uint32_t A_BYTES = BM * BK * 2;
  static constexpr uint32_t B_BYTES = (BN / CG) * BK * 2;
  static constexpr uint32_t BAR_BYTES = 256;
  static constexpr int EPI_WARPS = EW;  // 8 (mainloop-bound launches) or 16 (epilogue-heavy: 4 warps per scheduler)
  static constexpr int PARTS = EW / 4;  // column parts of a tile, one per epilogue warp of a lane quarter
  static constexpr uint32_t EPI_SMEM_BYTES = (EPI == EPI_HEAD) ? 3 * 128 * 4 * sizeof(float) : 0;  // partial sums
  // per-warp staging: [32][36] fp32 transpose buffer, or one 4 KB SWIZZLE_128B box for the TMA store
  static constexpr uint32_t STG_WARP_BYTES = TMA ? 4096 : 32 * 36 * sizeof(float);
  static constexpr uint32_t STG_BYTES = EPI_WARPS * STG_WARP_BYTES;
  static constexpr int ROPE_SMEM_ROWS = 64;  // positions -1 .. 62
  static constexpr uint32_t ROPE_SMEM_BYTES = (EPI == EPI_ROPE) ? ROPE_SMEM_ROWS * kRopeLd * sizeof(float) : 0;
  static constexpr uint32_t MAX_SMEM = 227 * 1024;
  // halo-staged convolution (A_CONV3H): the ring holds B tiles only; two 18 x 16 x 64-channel input tiles besides it
  static constexpr uint32_t HALO_TX_BYTES = 18 * kHaloPitch * 128;                 // bytes one box delivers
  static constexpr uint32_t HALO_BYTES = (HALO_TX_BYTES + 1023u) & ~1023u;        // stage stride (swizzle-pattern aligned)
  static constexpr int HALO_STAGES = kHaloPitch == 16 ? 2 : 3;
  static constexpr uint32_t HALO_TOTAL = HALO ? HALO_STAGES * HALO_BYTES : 0;
  static constexpr uint32_t STAGE_BYTES = HALO ? B_BYTES : A_BYTES + B_BYTES;
  static constexpr int STAGES_FIT =
      (MAX_SMEM - BAR_BYTES - EPI_SMEM_BYTES - STG_BYTES - ROPE_SMEM_BYTES - HALO_TOTAL) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_FIT > 8 ? 8 : STAGES_FIT;
  // layout: operand ring | halo tiles | staging (1024-byte aligned) | barriers | EPI_HEAD partials | RoPE table
  static constexpr uint32_t OFF_HALO = STAGES * STAGE_BYTES;
  static constexpr uint32_t OFF_STG = OFF_HALO + HALO_TOTAL;
  static constexpr uint32_t OFF_BAR = OFF_STG + STG_BYTES;
  static constexpr uint32_t OFF_EPI = OFF_BAR + BAR_BYTES;
  static constexpr uint32_t OFF_ROPE = OFF_EPI + EPI_SMEM_BYTES;
  static constexpr uint32_t SMEM_BYTES = OFF_ROPE + ROPE_SMEM_BYTES;
  static constexpr int THREADS = 128 + 32 * EPI_WARPS;
  static constexpr int EPI_THREADS = 32 * EPI_WARPS;
  static_assert(STAGES <= 8 && STAGES >= 3, "barrier area holds at most 8 stages");
  static_assert(OFF_STG % 1024 == 0 && OFF_HALO % 1024 == 0 && STG_WARP_BYTES % 16 == 0, "staging alignment");
};

// ---------------------------------------------------------------------------
// Epilogue for one 128 x BN accumulator tile; executed by the 16 epilogue warps (4 per scheduler, so that the
// global-load / TMEM-load latencies of one warp are covered by the others).
// warp (quarter, part): TMEM lanes [32*quarter, +32), columns [part*BN/4, +BN/4).
//
// tcgen05.ld (32x32b) hands every thread one accumulator ROW (32 consecutive columns).  Storing
// that way would make each warp store touch 32 different cache lines, so every 32x32 fp32 chunk
// is transposed through a per-warp shared-memory staging buffer ([32][36] floats, conflict-free
// for 128-bit accesses) into a COALESCED mapping -- lane -> (row = 4*it + lane/8, 4 columns at
// (lane%8)*4) -- in which bias, residuals, activation and the global stores happen: a warp-wide
// access covers 4 rows x 128 contiguous bytes (fp32) or 4 rows x 64 bytes (bf16).
// ---------------------------------------------------------------------------
constexpr int kStageLd = 36;                        // floats per staged row (32 + 4 pad)
constexpr int kStageFloatsPerWarp = 32 * kStageLd;  // 4608 B per epilogue warp

template <int AMODE>
__device__ __forceinline__ void tile_row(const GemmParams& p, int m_tile, int r_local, long long& orow, bool& valid) {
  if constexpr (AMODE == A_CONV3) {
    const int tpi = p.tiles_h * p.tiles_w;
    const int n = m_tile / tpi;
    const int t = m_tile - n * tpi;
    const int th = t / p.tiles_w;
    const int tw = t - th * p.tiles_w;
    const int h = th * 8 + (r_local >> 4);
    const int w = tw * 16 + (r_local & 15);
    valid = (h < p.H) && (w < p.W) && (n < p.nimg);
    orow = (static_cast<long long>(n) * p.H + h) * p.W + w;
  } else if constexpr (AMODE == A_CONV3H) {  // 16 rows x 8 pixels per 128-row tile
    const int tpi = p.tiles_h * p.tiles_w;
    const int n = m_tile / tpi;
    const int t = m_tile - n * tpi;
    const int th = t / p.tiles_w;
    const int tw = t - th * p.tiles_w;
    const int h = th * 16 + (r_local >> 3);
    const int w = tw * 8 + (r_local & 7);
    valid = (h < p.H) && (w < p.W) && (n < p.nimg);
    orow = (static_cast<long long>(n) * p.H + h) * p.W + w;
  } else {
    const int r = m_tile * 128 + r_local;
    valid = r < p.M;
    orow = r;
  }
}

__device__ __forceinline__ void stage_store32(float* stg, int lane, const float* v) {
  float4* d = reinterpret_cast<float4*>(stg + lane * kStageLd);
#pragma unroll
  for (int j = 0; j < 8; ++j) d[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
}

__device__ __forceinline__ uint2 pack4_bf16(const float4& v) {
  return make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
}
__device__ __forceinline__ void add_bf16x4(float4& v, uint2 q) {
  v.x += bf16_lo(q.x); v.y += bf16_hi(q.x); v.z += bf16_lo(q.y); v.w += bf16_hi(q.y);
}
// bf16 store of 4 values at p; in split mode also the residual part at p + w and the second hi copy at p + 2w
__device__ __forceinline__ void store4_bf16(__nv_bfloat16* p, const float4& v, int split, long long w) {
  const uint2 hi = pack4_bf16(v);
  *reinterpret_cast<uint2*>(p) = hi;
  if (split) {
    *reinterpret_cast<uint2*>(p + w) = make_uint2(pack_bf16x2_resid(v.x, v.y), pack_bf16x2_resid(v.z, v.w));
    *reinterpret_cast<uint2*>(p + 2 * w) = hi;
  }
}

template <int BN, int AMODE, int EPI, int EW>
__device__ __forceinline__ void epilogue_tile(const GemmParams& p, uint32_t taddr, int m_tile, int n_tile,
                                              int quarter, int part, float* epi_smem, float* stg,
                                              const float* rope_s, uint64_t* tfull_bar, uint32_t tfull_phase) {
  // Everything that does not depend on the accumulator (row bookkeeping, positions, residual prefetch) is done
  // BEFORE waiting for the tile's MMAs, so its global-memory latency hides behind the mainloop.
  constexpr int PARTS = EW / 4;
  constexpr int CH = BN / PARTS;  // columns per epilogue warp
  const int lane = threadIdx.x & 31;
  const int colbase = n_tile * BN + part * CH;

  if constexpr (EPI == EPI_HEAD) {
    // head.2 epilogue: ReLU -> head.4 (1x1, 128 -> 4) -> postprocess (dpt_block.py:319-323,
    // postprocess.py:10-62).  Row mapping: each thread owns 32 of the 128 channels of one pixel; the four
    // column parts are combined through shared memory.  Nothing but 16 bytes per pixel is written.
    static_assert(BN == 128, "EPI_HEAD expects the 128-channel head");
    mbar_wait(tfull_bar, tfull_phase);
    tc_fence_after();
    const int r_local = quarter * 32 + lane;
    long long orow;
    bool valid;
    tile_row<AMODE>(p, m_tile, r_local, orow, valid);
    float psum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
    for (int c = 0; c < CH; c += 32) {
      const int col = colbase + c;
      uint32_t acc[32];
      tmem_ld32(taddr + c, acc);
      tmem_ld_wait();
      const float4* wp = reinterpret_cast<const float4*>(p.head_w) + col;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        float v = fmaxf(__uint_as_float(acc[i]) + __ldg(p.bias + col + i), 0.0f);
        float4 w4 = __ldg(wp + i);
        psum[0] = fmaf(v, w4.x, psum[0]);
        psum[1] = fmaf(v, w4.y, psum[1]);
        psum[2] = fmaf(v, w4.z, psum[2]);
        psum[3] = fmaf(v, w4.w, psum[3]);
      }
    }
    if (part != 0)
      reinterpret_cast<float4*>(epi_smem)[(part - 1) * 128 + r_local] = make_float4(psum[0], psum[1], psum[2], psum[3]);
    named_bar_sync(1, 32 * EW);
    if (part == 0) {
      float4 o = reinterpret_cast<float4*>(epi_smem)[r_local];
#pragma unroll
      for (int q = 1; q < PARTS - 1; ++q) {
        const float4 t = reinterpret_cast<float4*>(epi_smem)[q * 128 + r_local];
        o.x += t.x; o.y += t.y; o.z += t.z; o.w += t.w;
      }
      const float x = psum[0] + o.x + __ldg(p.head_b + 0);
      const float y = psum[1] + o.y + __ldg(p.head_b + 1);
      const float z = psum[2] + o.z + __ldg(p.head_b + 2);
      const float cf = psum[3] + o.w + __ldg(p.head_b + 3);
      if (valid) {
        const float d = sqrtf(x * x + y * y + z * z);
        const float dc = fmaxf(d, 1e-8f);
        const float e = expm1f(d);
        p.pts3d[orow * 3 + 0] = (x / dc) * e;
        p.pts3d[orow * 3 + 1] = (y / dc) * e;
        p.pts3d[orow * 3 + 2] = (z / dc) * e;
        p.conf[orow] = 1.0f + expf(cf);
      }
    }
    named_bar_sync(1, 32 * EW);
    return;
  } else {
    // ---- coalesced mapping bookkeeping: this lane touches rows 4*it + (lane>>3), it = 0..7 ----
    const int crow = lane >> 3;
    const int cc = (lane & 7) * 4;  // first of this lane's 4 columns inside the 32-column chunk
    long long orow_c[8];
    uint32_t vmask = 0;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      bool v;
      tile_row<AMODE>(p, m_tile, quarter * 32 + it * 4 + crow, orow_c[it], v);
      vmask |= (v ? 1u : 0u) << it;
      if constexpr (EPI == EPI_F32) {
        if (p.rowmap_n > 0) {
          const long long s = orow_c[it] / p.rowmap_n;
          orow_c[it] = s * (p.rowmap_n + 1) + 1 + (orow_c[it] - s * p.rowmap_n);
        }
      }
      if constexpr (EPI == EPI_PIXSHUF) {
        // token -> (n, h, w); keep ((n*ps_h + h)*k) * (ps_w*k) + w*k  (pixel index of sub-pixel (0,0))
        const int hw = p.ps_h * p.ps_w;
        const int n = static_cast<int>(orow_c[it] / hw);
        const int rem = static_cast<int>(orow_c[it] - static_cast<long long>(n) * hw);
        const int h = rem / p.ps_w;
        const int w = rem - h * p.ps_w;
        orow_c[it] = (static_cast<long long>(n) * p.ps_h + h) * p.ps_k * (p.ps_w * p.ps_k) + w * p.ps_k;
      }
    }

    // EPI_ROPE: (y, x) token positions of this lane's 8 rows
    [[maybe_unused]] int pos_y[8], pos_x[8];
    if constexpr (EPI == EPI_ROPE) {
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        int py = 0, px = 0;
        if ((vmask >> it) & 1u) {
          const int2 pp = *reinterpret_cast<const int2*>(p.pos + 2 * orow_c[it]);
          py = pp.x;
          px = pp.y;
          if (py < -1 || py > p.rope_max_pos || px < -1 || px > p.rope_max_pos)
            device_fatal("token position outside the RoPE table");
        }
        pos_y[it] = py;
        pos_x[it] = px;
      }
    }

    // flush one staged 32-column chunk (global columns [col, col+32)) in the coalesced mapping
    auto flush = [&](int col) {
      if (col >= p.N) return;
      float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
      long long pix_off = 0;
      int co = 0;
      if constexpr (EPI == EPI_PIXSHUF) {
        const int kk = col / p.ps_cout;
        co = col - kk * p.ps_cout;
        const int kh = kk / p.ps_k;
        const int kw = kk - kh * p.ps_k;
        pix_off = static_cast<long long>(kh) * (p.ps_w * p.ps_k) + kw;
        b4 = __ldg(reinterpret_cast<const float4*>(p.bias + co + cc));
      } else {
        if (p.bias) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col + cc));
      }
      // Residual loads are issued ahead of the stores, four rows at a time: the in-place residual stream
      // (resid == out) would otherwise force load -> store -> load ordering (possible aliasing) and serialise
      // ~1 us DRAM round trips.
#pragma unroll
      for (int hb = 0; hb < 2; ++hb) {
        [[maybe_unused]] float4 rf[4];
        [[maybe_unused]] uint2 ra[4], rb[4];
        if constexpr (EPI == EPI_F32) {
          if (p.resid) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int it = hb * 4 + k;
              rf[k] = ((vmask >> it) & 1u)
                          ? *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.resid) +
                                                             orow_c[it] * p.ldo + col + cc)
                          : make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
        }
        if constexpr (EPI == EPI_BF16) {
          if (p.resid) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int it = hb * 4 + k;
              ra[k] = ((vmask >> it) & 1u)
                          ? *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(p.resid) +
                                                            orow_c[it] * p.ldo + col + cc)
                          : make_uint2(0u, 0u);
            }
          }
          if (p.resid2) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int it = hb * 4 + k;
              rb[k] = ((vmask >> it) & 1u)
                          ? *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(p.resid2) +
                                                            orow_c[it] * p.ldo + col + cc)
                          : make_uint2(0u, 0u);
            }
          }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          // straight-line body: only the global loads / stores are predicated on row validity, so the
          // compiler can interleave the four rows of a batch
          const int it = hb * 4 + k;
          const bool ok = (vmask >> it) & 1u;
          float4 v = *reinterpret_cast<const float4*>(stg + (it * 4 + crow) * kStageLd + cc);
          v.x += b4.x; v.y += b4.y; v.z += b4.z; v.w += b4.w;
          if constexpr (EPI == EPI_ROPE) {
            // 2-D RoPE (pos_embed/pos_embed.py:149-185, curope/kernels.cu:17-82).  A 32-column chunk is exactly
            // one half of a 64-wide head: the y-half (rotated by pos_y) or the x-half (pos_x).  Inside a half the
            // pairs are (i, i+16): the partner columns live in lane ^ 4 of the same row group.
            if (col < p.rope_cols) {  // uniform per chunk: the V columns are not rotated
              const float4 w = make_float4(__shfl_xor_sync(0xffffffffu, v.x, 4), __shfl_xor_sync(0xffffffffu, v.y, 4),
                                           __shfl_xor_sync(0xffffffffu, v.z, 4), __shfl_xor_sync(0xffffffffu, v.w, 4));
              const int pz = ((col >> 5) & 1) ? pos_x[it] : pos_y[it];
              const int ps = min(pz + 1, p.rope_smem_rows - 1);  // clamped shared-memory row (always safe to read)
              const float4* tb = reinterpret_cast<const float4*>(rope_s + ps * kRopeLd) + ((cc & 15) >> 1);
              float4 t0 = tb[0];  // (cos f, sin f, cos f+1, sin f+1)
              float4 t1 = tb[1];  // (cos f+2, sin f+2, cos f+3, sin f+3)
              if (pz + 1 >= p.rope_smem_rows) {  // beyond the staged rows (very large images): global table
                const float4* tg =
                    reinterpret_cast<const float4*>(p.rope_tab + static_cast<long long>(pz + 1) * 32) + ((cc & 15) >> 1);
                t0 = __ldg(tg);
                t1 = __ldg(tg + 1);
              }
              const float sg = (cc & 16) ? 1.0f : -1.0f;  // lower 16: u*cos - w*sin ; upper 16: u*cos + w*sin
              v.x = fmaf(sg * w.x, t0.y, v.x * t0.x);
              v.y = fmaf(sg * w.y, t0.w, v.y * t0.z);
              v.z = fmaf(sg * w.z, t1.y, v.z * t1.x);
              v.w = fmaf(sg * w.w, t1.w, v.w * t1.z);
            }
          }
          if constexpr (EPI == EPI_F32) {
            float* op = reinterpret_cast<float*>(p.out) + orow_c[it] * p.ldo + col + cc;
            if (p.resid) { v.x += rf[k].x; v.y += rf[k].y; v.z += rf[k].z; v.w += rf[k].w; }
            if (ok) *reinterpret_cast<float4*>(op) = v;
          } else if constexpr (EPI == EPI_PIXSHUF) {
            __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.out) +
                                (orow_c[it] + pix_off) * (p.split ? 3 * p.ps_cout : p.ps_cout) + co + cc;
            if (ok) store4_bf16(op, v, p.split, p.ps_cout);
          } else {
            const long long off = orow_c[it] * p.ldo + col + cc;
            if constexpr (EPI == EPI_BF16) {
              if (p.resid) {
                add_bf16x4(v, ra[k]);
                if (p.split && ok)  // residual (lo) part of the skip tensor
                  add_bf16x4(v, *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(p.resid) + off + p.N));
              }
              if (p.resid2) {
                add_bf16x4(v, rb[k]);
                if (p.split && ok)
                  add_bf16x4(v, *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(p.resid2) + off + p.N));
              }
              if (p.relu_main) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
            }
            if constexpr (EPI == EPI_GELU) { v.x = gelu_erf(v.x); v.y = gelu_erf(v.y); v.z = gelu_erf(v.z); v.w = gelu_erf(v.w); }
            if (p.out && ok) store4_bf16(reinterpret_cast<__nv_bfloat16*>(p.out) + off, v, p.split, p.N);
            if constexpr (EPI == EPI_BF16) {
              if (p.out2 && ok) {
                v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
                store4_bf16(reinterpret_cast<__nv_bfloat16*>(p.out2) + off, v, p.split, p.N);
              }
            }
          }
        }
      }
    };

    // L2 prefetch of the residual rows this lane will add (fp32 residual stream / bf16 skip tensors)
    if constexpr (EPI == EPI_F32 || EPI == EPI_BF16) {
      if (p.resid) {
        constexpr int esz = (EPI == EPI_F32) ? 4 : 2;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          if ((vmask >> it) & 1u) {
#pragma unroll
            for (int c = 0; c < CH; c += 32) {
              if (colbase + c < p.N) {
                const char* a = reinterpret_cast<const char*>(p.resid) + (orow_c[it] * p.ldo + colbase + c + cc) * esz;
                asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
                if constexpr (EPI == EPI_BF16) {
                  if (p.resid2) {
                    const char* a2 = reinterpret_cast<const char*>(p.resid2) + (orow_c[it] * p.ldo + colbase + c + cc) * esz;
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(a2));
                  }
                }
              }
            }
          }
        }
      }
    }
    mbar_wait(tfull_bar, tfull_phase);
    tc_fence_after();
#pragma unroll 1
    for (int c = 0; c < CH; c += 32) {
      uint32_t acc[32];
      tmem_ld32(taddr + c, acc);
      tmem_ld_wait();
      stage_store32(stg, lane, reinterpret_cast<const float*>(acc));
      __syncwarp();
      flush(colbase + c);
      __syncwarp();
    }
  }
}

// ---------------------------------------------------------------------------
// TMA-store epilogue for the wide linear layers (BN = 256; EPI_BF16 / EPI_GELU / EPI_ROPE / EPI_F32).
//
// The math stays in the tcgen05.ld mapping -- one accumulator ROW per thread, 32 consecutive columns per load --
// so bias, GELU and the RoPE rotation (whose (i, i+16) partners sit in the same thread) need no shuffles.  The
// results go to a per-warp 32-row x 128-byte shared-memory box in the SWIZZLE_128B layout (conflict-free 16-byte
// stores) and leave with ONE bulk tensor store per box issued by lane 0: no per-lane global stores, rows >= M are
// clipped by the tensor map, and the in-place fp32 residual stream (x += proj(..)) becomes a bulk reduce-add
// performed by the L2, so the epilogue never loads the residual.  The TMEM load of chunk c+1 is in flight while
// chunk c is processed, and the accumulator stage is released as soon as the last load has landed.
// ---------------------------------------------------------------------------
template <int BN, int EPI, int EW, typename Release>
__device__ __forceinline__ void epilogue_tile_tma(const GemmParams& p, const CUtensorMap* tmC, uint32_t taddr, int m_tile,
                                                  int colbase, int nch, int row_off, int quarter, int part, uint8_t* cbuf,
                                                  const float* rope_s,
                                                  uint64_t* tfull_bar, uint32_t tfull_phase, Release&& release,
                                                  long long* trace) {
  constexpr int PARTS = EW / 4;
  constexpr int CH = BN / PARTS;  // columns per epilogue warp of a full tile
  constexpr int NCH = CH / 32;    // 32-column TMEM loads per full tile; nch <= NCH is what this tile has (half tiles: NCH / 2)
  // 8 warps (168 registers): the next chunk's TMEM load is in flight while the current one is processed;
  // 16 warps (96 registers) rely on the four warps per scheduler instead
  constexpr bool PF = (EW == 8);
  constexpr bool OUT_F32 = (EPI == EPI_F32);
  const int lane = threadIdx.x & 31;
  (void)part;
  const int row0 = m_tile * 128 + quarter * 32;  // first row of this warp's box (token index)
  const int srow0 = row0 + row_off;              // ... in the output tensor (split-K partial slices)
  const uint32_t cb = smem_u32(cbuf);
  const uint32_t srow = cb + lane * 128;  // this thread's row inside the box
  const uint32_t sx = lane & 7;           // 128B swizzle: 16-byte chunk j lives at j ^ (row & 7)

  [[maybe_unused]] int py = 0, px = 0;
  if constexpr (EPI == EPI_ROPE) {
    if (row0 + lane < p.M) {
      const int2 pp = *reinterpret_cast<const int2*>(p.pos + 2 * static_cast<long long>(row0 + lane));
      py = pp.x;
      px = pp.y;
      if (py < -1 || py > p.rope_max_pos || px < -1 || px > p.rope_max_pos)
        device_fatal("token position outside the RoPE table");
    }
  }

  // Split-precision parity mode, bf16 outputs: the tile is drained twice -- pass 0 stores hi = bf16(v) at columns
  // [col, ..) and [2N + col, ..), pass 1 recomputes v from TMEM and stores lo = bf16(v - hi) at [N + col, ..) -- so the
  // production register footprint is unchanged.  The accumulator stage is released in the last pass only.
  const int npass = (!OUT_F32 && p.split) ? 2 : 1;
  const bool has_bias = p.bias != nullptr;
#pragma unroll 1
  for (int pass = 0; pass < npass; ++pass) {
  // bias of chunk c is loaded one chunk ahead (warp-uniform addresses: L1 broadcast), so its latency never sits
  // between the TMEM load and the math
  float4 bcur[8];
#pragma unroll
  for (int j = 0; j < 8; ++j)
    bcur[j] = has_bias ? __ldg(reinterpret_cast<const float4*>(p.bias + colbase) + j) : make_float4(0.f, 0.f, 0.f, 0.f);

  if (pass == 0) {
    mbar_wait(tfull_bar, tfull_phase);
    tc_fence_after();
    if (trace) trace[0] = clock64();
  }
  uint32_t acc[PF ? 2 : 1][32];
  tmem_ld32(taddr, acc[0]);
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    if (c >= nch) break;
    tmem_ld_wait();
    float* v = reinterpret_cast<float*>(acc[PF ? (c & 1) : 0]);
    const int col = colbase + c * 32;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      v[4 * j] += bcur[j].x; v[4 * j + 1] += bcur[j].y; v[4 * j + 2] += bcur[j].z; v[4 * j + 3] += bcur[j].w;
    }
    if (c + 1 < nch) {
      if constexpr (PF) tmem_ld32(taddr + (c + 1) * 32, acc[(c + 1) & 1]);
      if (has_bias) {
#pragma unroll
        for (int j = 0; j < 8; ++j) bcur[j] = __ldg(reinterpret_cast<const float4*>(p.bias + col + 32) + j);
      }
    } else if (pass == npass - 1) {
      release();
    }
    if (trace) trace[1 + 3 * c] = clock64();
    if constexpr (EPI == EPI_GELU) {
#pragma unroll
      for (int i = 0; i < 32; i += 2) gelu_erf2(v[i], v[i + 1]);
    }
    if constexpr (EPI == EPI_BF16) {
      if (p.relu_main) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.0f);
      }
    }
    if constexpr (EPI == EPI_ROPE) {
      // 2-D RoPE (pos_embed/pos_embed.py:149-185, curope/kernels.cu:17-82): a 32-column chunk is one half of a
      // 64-wide head -- the y-half (rotated by pos_y) or the x-half (pos_x); inside it the pairs are (i, i + 16).
      if (col < p.rope_cols) {  // warp-uniform: the V columns are not rotated
        const int ps = (((col >> 5) & 1) ? px : py) + 1;
        const float4* tb = (ps < p.rope_smem_rows)
                               ? reinterpret_cast<const float4*>(rope_s + ps * kRopeLd)
                               : reinterpret_cast<const float4*>(p.rope_tab + static_cast<long long>(ps) * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 t = tb[j];  // (cos f, sin f, cos f+1, sin f+1), f = 2j
          const float a0 = v[2 * j], b0 = v[2 * j + 16], a1 = v[2 * j + 1], b1 = v[2 * j + 17];
          v[2 * j] = fmaf(-b0, t.y, a0 * t.x);
          v[2 * j + 16] = fmaf(a0, t.y, b0 * t.x);
          v[2 * j + 1] = fmaf(-b1, t.w, a1 * t.z);
          v[2 * j + 17] = fmaf(a1, t.w, b1 * t.z);
        }
      }
    }
    if (trace) trace[2 + 3 * c] = clock64();
    // ---- registers -> swizzled box -> bulk tensor store ----
    if constexpr (OUT_F32) {
      // one box per chunk: 32 rows x 32 fp32 columns
      if (lane == 0) tma_store_wait_read();
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 8; ++j)
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(srow + ((j ^ sx) << 4)), "f"(v[4 * j]),
                     "f"(v[4 * j + 1]), "f"(v[4 * j + 2]), "f"(v[4 * j + 3])
                     : "memory");
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        if (p.c_reduce) tma_reduce_add_2d(tmC, cbuf, col, srow0);
        else tma_store_2d(tmC, cbuf, col, srow0);
        tma_store_commit();
      }
    } else {
      // one box per two chunks: 32 rows x 64 bf16 columns
      const int sub = c & 1;
      if (sub == 0) {
        if (lane == 0) tma_store_wait_read();
        __syncwarp();
      }
      if (pass == 1) {  // residual part: lo = v - bf16(v)
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = bf16_resid(v[i]);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j)
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(srow + (((sub * 4 + j) ^ sx) << 4)),
                     "r"(pack_bf16x2(v[8 * j], v[8 * j + 1])), "r"(pack_bf16x2(v[8 * j + 2], v[8 * j + 3])),
                     "r"(pack_bf16x2(v[8 * j + 4], v[8 * j + 5])), "r"(pack_bf16x2(v[8 * j + 6], v[8 * j + 7]))
                     : "memory");
      if (sub == 1) {
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(tmC, cbuf, col - 32 + pass * p.N, srow0);
          if (p.split && pass == 0) tma_store_2d(tmC, cbuf, col - 32 + 2 * p.N, srow0);  // second hi copy
          tma_store_commit();
        }
      }
    }
    if constexpr (!PF) {
      if (c + 1 < nch) tmem_ld32(taddr + (c + 1) * 32, acc[0]);  // the registers are free again
    }
    if (trace) trace[3 + 3 * c] = clock64();
  }
  }  // pass
}

template <int BN, int AMODE, int EPI, int CG, int EW, bool TMA>
__global__ void __launch_bounds__(GemmCfg<BN, CG, EW, EPI, TMA>::THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, const GemmParams p) {
  constexpr bool HALO = (AMODE == A_CONV3H);
  using Cfg = GemmCfg<BN, CG, EW, EPI, TMA, HALO>;
  static_assert(!TMA || (AMODE == A_LINEAR && (BN == 256 || BN == 128) &&
                         (EPI == EPI_BF16 || EPI == EPI_GELU || EPI == EPI_F32 || EPI == EPI_ROPE)),
                "TMA-store epilogue: wide linear layers only");
  constexpr int STAGES = Cfg::STAGES;
  constexpr uint32_t A_BYTES = Cfg::A_BYTES;
  constexpr uint32_t B_BYTES = Cfg::B_BYTES;

  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sB = smem + (HALO ? 0u : STAGES * A_BYTES);
  [[maybe_unused]] uint8_t* sH = smem + Cfg::OFF_HALO;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* tfull = bars + 2 * STAGES;
  uint64_t* tempty = tfull + 2;
  [[maybe_unused]] uint64_t* hfull = tempty + 2;   // halo tiles (A_CONV3H), up to 3 stages
  [[maybe_unused]] uint64_t* hempty = tempty + 5;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 8);
  static_assert(Cfg::HALO_STAGES <= 3 && (2 * STAGES + 12) * 8 + 4 <= Cfg::BAR_BYTES, "barrier area");
  [[maybe_unused]] float* epi_smem = reinterpret_cast<float*>(smem + Cfg::OFF_EPI);
  uint8_t* stg_all = smem + Cfg::OFF_STG;
  [[maybe_unused]] float* rope_s = reinterpret_cast<float*>(smem + Cfg::OFF_ROPE);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = (CG == 2) ? cluster_ctarank() : 0u;
  const bool leader = (cta_rank == 0);

  // 128-row tiles; a work item covers CG consecutive ones
  const int m_tiles128 = is_conv(AMODE) ? p.nimg * p.tiles_h * p.tiles_w : (p.M + 127) / 128;
  const int m_tiles = (m_tiles128 + CG - 1) / CG;
  const int n_tiles = (p.N + BN - 1) / BN;
  const int ksplit = TMA ? p.ksplit : 1;
  const int num_tiles = m_tiles * n_tiles * ksplit;  // work items (m, n, ks), ks fastest
  const int cpb = is_conv(AMODE) ? (p.Cin / 64) : 1;  // 64-channel chunks per filter tap
  const int nkb = is_conv(AMODE) ? 9 * cpb : (p.K + 63) / 64;
  const int first_tile = blockIdx.x / CG;
  const int tile_step = gridDim.x / CG;
  // work items: tiles [0, num_tiles), the last tail_r of them as two half tiles each (TMA epilogue / CTA pair / BN = 256 only)
  constexpr bool TAIL = TMA && CG == 2 && BN == 256 && EPI == EPI_F32;
  const int tail_r = TAIL ? p.tail_r : 0;
  const int num_work = num_tiles + tail_r;
  auto work_tile = [&](int w, int& half) {
    half = -1;
    if (TAIL && tail_r > 0 && w >= p.tail_first) {
      const int q = w - p.tail_first;
      half = q & 1;
      return p.tail_first + (q >> 1);
    }
    return w;
  };

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0) device_fatal("dynamic shared memory is not 1024-byte aligned");
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);   // the leader's producer arrives once (its expect_tx covers both CTAs' bytes)
      mbar_init(&empty[s], 1);  // one (multicast) tcgen05.commit
    }
    mbar_init(&tfull[0], 1);
    mbar_init(&tfull[1], 1);
    mbar_init(&tempty[0], Cfg::EPI_WARPS * CG);  // every epilogue warp of the pair arrives on the leader's barrier
    mbar_init(&tempty[1], Cfg::EPI_WARPS * CG);
    if constexpr (HALO) {
      for (int s = 0; s < Cfg::HALO_STAGES; ++s) {
        mbar_init(&hfull[s], 1);
        mbar_init(&hempty[s], 1);
      }
    }
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if constexpr (TMA) tma_prefetch_desc(&tmC);
  }
  if constexpr (EPI == EPI_ROPE) {
    // the sin/cos table is a constant of the library (not produced by the previous kernel): stage the rows for
    // small positions in shared memory before the PDL wait
    if (warp >= 4) {
      const int nfl = p.rope_smem_rows * 32;
      for (int i = (threadIdx.x - 128) * 4; i < nfl; i += Cfg::EPI_THREADS * 4)
        *reinterpret_cast<float4*>(rope_s + (i >> 5) * kRopeLd + (i & 31)) =
            __ldg(reinterpret_cast<const float4*>(p.rope_tab + i));
    }
  }
  if (warp == 2) {
    if constexpr (CG == 2) {
      tmem_alloc_cg2(tmem_slot, 2 * BN);
      tmem_relinquish_cg2();
    } else {
      tmem_alloc(tmem_slot, 2 * BN);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();  // CTA-level barrier also for CG == 2: orders the barrier init / TMEM-slot write for this CTA's readers
                    // (and is what compute-sanitizer racecheck models; barrier.cluster alone is reported as a hazard)
  if constexpr (CG == 2) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();               // previous kernel's outputs (A operand, residuals) are complete and visible
  pdl_launch_dependents();  // the next kernel may start its prologue

  if (warp == 0) {
    // ===================== TMA producer (one per CTA) =====================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      [[maybe_unused]] int hstage = 0;
      [[maybe_unused]] uint32_t hphase = 0;
      for (int w = first_tile; w < num_work; w += tile_step) {
        int half;
        const int tile = work_tile(w, half);
        const int ks = tile % ksplit, mn = tile / ksplit;
        const int m_tile = (mn / n_tiles) * CG + static_cast<int>(cta_rank);  // this CTA's 128-row tile
        const int n_tile = mn % n_tiles;
        const int kb0 = ks * nkb / ksplit, kb1 = (ks + 1) * nkb / ksplit;
        // half tile: this CTA's 64 weight rows are the first 64 of the (full-size) box it loads
        const int b_row0 = (half < 0) ? n_tile * BN + static_cast<int>(cta_rank) * (BN / CG)
                                      : n_tile * BN + half * (BN / 2) + static_cast<int>(cta_rank) * (BN / 4);
        int cn = 0, ch0 = 0, cw0 = 0;
        if constexpr (is_conv(AMODE)) {
          const int tpi = p.tiles_h * p.tiles_w;
          cn = m_tile / tpi;  // >= nimg for the padding tile of an odd tile count: TMA zero-fills
          const int t = m_tile - cn * tpi;
          const int th = t / p.tiles_w;
          ch0 = th * (HALO ? 16 : 8);
          cw0 = (t - th * p.tiles_w) * (HALO ? 8 : 16);
        }
        if constexpr (HALO) {
          // channel chunk outermost: one halo tile, then the nine taps' weight tiles
          for (int cc = 0; cc < cpb; ++cc) {
            mbar_wait(&hempty[hstage], hphase ^ 1);
            if (CG == 1 || leader) mbar_arrive_expect_tx(&hfull[hstage], CG * Cfg::HALO_TX_BYTES);
            if constexpr (CG == 2)
              tma_load_4d_cg2(sH + hstage * Cfg::HALO_BYTES, &tmA, &hfull[hstage], cc * 64, cw0 - 1, ch0 - 1, cn);
            else
              tma_load_4d(sH + hstage * Cfg::HALO_BYTES, &tmA, &hfull[hstage], cc * 64, cw0 - 1, ch0 - 1, cn);
            if (++hstage == Cfg::HALO_STAGES) { hstage = 0; hphase ^= 1; }
            for (int tap = 0; tap < 9; ++tap) {
              mbar_wait(&empty[stage], phase ^ 1);
              if (CG == 1 || leader) mbar_arrive_expect_tx(&full[stage], CG * B_BYTES);
              if constexpr (CG == 2)
                tma_load_2d_cg2(sB + stage * B_BYTES, &tmB, &full[stage], (tap * cpb + cc) * 64, b_row0);
              else
                tma_load_2d(sB + stage * B_BYTES, &tmB, &full[stage], (tap * cpb + cc) * 64, b_row0);
              if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
          }
        } else {
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          if (CG == 1 || leader) mbar_arrive_expect_tx(&full[stage], CG * (A_BYTES + B_BYTES));
          if constexpr (AMODE == A_CONV3) {
            const int tap = kb / cpb;
            const int cc = kb - tap * cpb;
            const int kh = tap / 3;
            const int kw = tap - kh * 3;
            if constexpr (CG == 2)
              tma_load_4d_cg2(sA + stage * A_BYTES, &tmA, &full[stage], cc * 64, cw0 + kw - 1, ch0 + kh - 1, cn);
            else
              tma_load_4d(sA + stage * A_BYTES, &tmA, &full[stage], cc * 64, cw0 + kw - 1, ch0 + kh - 1, cn);
          } else {
            if constexpr (CG == 2)
              tma_load_2d_cg2(sA + stage * A_BYTES, &tmA, &full[stage], kb * 64, m_tile * 128);
            else
              tma_load_2d(sA + stage * A_BYTES, &tmA, &full[stage], kb * 64, m_tile * 128);
          }
          if constexpr (CG == 2)
            tma_load_2d_cg2(sB + stage * B_BYTES, &tmB, &full[stage], kb * 64, b_row0);
          else
            tma_load_2d(sB + stage * B_BYTES, &tmB, &full[stage], kb * 64, b_row0);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader && elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(128 * CG, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      [[maybe_unused]] int hstage = 0;
      [[maybe_unused]] uint32_t hphase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int w = first_tile; w < num_work; w += tile_step) {
        int half;
        const int tile = work_tile(w, half);
        const uint32_t idesc_t = (half < 0) ? idesc : make_idesc_bf16(128 * CG, BN / 2, 0, 0);
        const int tcount = (w - first_tile) / tile_step;
        const bool trc = p.dbg && blockIdx.x == 0 && tcount < 40;
        if (trc) p.dbg[256 + 4 * tcount + 0] = clock64();
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        if (trc) p.dbg[256 + 4 * tcount + 1] = clock64();
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        const int ks = tile % ksplit;
        const int kb0 = ks * nkb / ksplit, kb1 = (ks + 1) * nkb / ksplit;
        if constexpr (HALO) {
          for (int cc = 0; cc < cpb; ++cc) {
            mbar_wait(&hfull[hstage], hphase);
            const uint32_t hbase = smem_u32(sH + hstage * Cfg::HALO_BYTES);
            for (int tap = 0; tap < 9; ++tap) {
              mbar_wait(&full[stage], phase);
              tc_fence_after();
              const int kh = tap / 3, kw = tap - kh * 3;
              // output pixel (y, x) of the 16 x 8 tile reads halo pixel (y + kh, x + kw); halo rows are kHaloPitch pixels
              const uint64_t adesc = make_smem_desc_sw128_rows(hbase + (kh * kHaloPitch + kw) * 128, kHaloPitch * 128);
              const uint64_t bdesc = make_smem_desc_sw128(smem_u32(sB + stage * B_BYTES));
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                if constexpr (CG == 2)
                  umma_bf16_cg2(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (cc > 0 || tap > 0 || k != 0) ? 1u : 0u);
                else
                  umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (cc > 0 || tap > 0 || k != 0) ? 1u : 0u);
              }
              if constexpr (CG == 2) umma_commit_cg2(&empty[stage]); else umma_commit(&empty[stage]);
              if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            if constexpr (CG == 2) umma_commit_cg2(&hempty[hstage]); else umma_commit(&hempty[hstage]);
            if (++hstage == Cfg::HALO_STAGES) { hstage = 0; hphase ^= 1; }
          }
        } else {
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint64_t adesc = make_smem_desc_sw128(smem_u32(sA + stage * A_BYTES));
          const uint64_t bdesc = make_smem_desc_sw128(smem_u32(sB + stage * B_BYTES));
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            // advance 16 K-elements = 32 bytes inside the 128B swizzle atom (encoded >> 4)
            if constexpr (CG == 2)
              umma_bf16_cg2(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc_t, (kb > kb0 || k != 0) ? 1u : 0u);
            else
              umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc_t, (kb > kb0 || k != 0) ? 1u : 0u);
          }
          if constexpr (CG == 2) umma_commit_cg2(&empty[stage]); else umma_commit(&empty[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        }
        if constexpr (CG == 2) umma_commit_cg2(&tfull[acc]); else umma_commit(&tfull[acc]);
        if (trc) p.dbg[256 + 4 * tcount + 2] = clock64();
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ============ epilogue warps (every CTA drains its own 128 accumulator rows) ============
    const int quarter = warp & 3;
    const int part = (warp - 4) >> 2;  // which quarter of the tile's columns
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int w = first_tile; w < num_work; w += tile_step) {
      int half;
      const int tile = work_tile(w, half);
      const int ks = tile % ksplit, mn = tile / ksplit;
      const int m_tile = (mn / n_tiles) * CG + static_cast<int>(cta_rank);
      const int n_tile = mn % n_tiles;
      const int tcount = (w - first_tile) / tile_step;
      const bool trc = p.dbg && blockIdx.x == 0 && warp == 4 && lane == 0 && tcount < 40;
      if (trc) p.dbg[4 * tcount + 0] = clock64();
      if (trc) p.dbg[4 * tcount + 1] = clock64();
      // columns of this epilogue warp: a quarter-of-the-parts slice of the full tile, or of the 128-column half tile
      const int cols_w = (half < 0 ? BN : BN / 2) / Cfg::PARTS;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN + part * cols_w;
      // hand the accumulator stage back to the MMA warp (called once all of this warp's TMEM loads have landed)
      auto release = [&]() {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (CG == 1 || leader) mbar_arrive(&tempty[acc]); else mbar_arrive_remote(&tempty[acc], 0);
        }
      };
      if constexpr (TMA) {
        epilogue_tile_tma<BN, EPI, EW>(p, &tmC, taddr, m_tile, n_tile * BN + (half < 0 ? 0 : half * (BN / 2)) + part * cols_w,
                                       cols_w / 32, ks * p.split_rows, quarter, part,
                                       stg_all + (warp - 4) * Cfg::STG_WARP_BYTES, rope_s, &tfull[acc], acc_phase, release,
                                       (trc && tcount < 16) ? p.dbg + 512 + 16 * tcount : nullptr);
      } else {
        epilogue_tile<BN, AMODE, EPI, EW>(p, taddr, m_tile, n_tile, quarter, part, epi_smem,
                                          reinterpret_cast<float*>(stg_all + (warp - 4) * Cfg::STG_WARP_BYTES), rope_s,
                                          &tfull[acc], acc_phase);
        release();
      }
      if (trc) p.dbg[4 * tcount + 2] = clock64();
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if constexpr (TMA) {
      if (lane == 0) tma_store_wait_all();  // the bulk stores read this CTA's shared memory
    }
  }

  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    if constexpr (CG == 2) tmem_dealloc_cg2(tmem_base, 2 * BN); else tmem_dealloc(tmem_base, 2 * BN);
  }
}

}  // namespace sta
