// Persistent warp-specialised tcgen05 GEMM / implicit-GEMM 3x3 convolution for sm_100a.
//
//   D[M,N] = A[M,K] * W[N,K]^T  (bf16 operands, fp32 accumulation in TMEM)
//
// Roles (384 threads): warp 0 = TMA producer, warp 1 = MMA issuer (one elected
// thread issues tcgen05.mma), warp 2 = TMEM allocator, warps 4..11 = epilogue
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

enum AMode : int { A_LINEAR = 0, A_CONV3 = 1 };
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
  // EPI_PIXSHUF
  int ps_k, ps_cout, ps_h, ps_w;
  // EPI_HEAD
  const float* head_w;  // [128][4]
  const float* head_b;  // [4]
  float* pts3d;         // [pixels][3]
  float* conf;          // [pixels]
};

// CG = 1: one CTA per 128 x BN tile (tcgen05 cta_group::1).
// CG = 2: a CTA pair (cluster of 2 on one TPC) per 256 x BN tile (cta_group::2): each CTA stages its own
//         128 rows of A and BN/2 rows of B, so per-SM shared-memory traffic per MMA is halved for B.
template <int BN, int CG>
struct GemmCfg {
  static constexpr int BM = 128;
  static constexpr int BK = 64;
  static constexpr uint32_t A_BYTES = BM * BK * 2;
  static constexpr uint32_t B_BYTES = (BN / CG) * BK * 2;
  static constexpr uint32_t BAR_BYTES = 256;
  static constexpr uint32_t EPI_SMEM_BYTES = 128 * 4 * sizeof(float);  // EPI_HEAD partial sums
  static constexpr uint32_t STG_BYTES = 8 * 32 * 36 * sizeof(float);   // 8 epilogue warps x [32][36] fp32 staging
  static constexpr uint32_t MAX_SMEM = 227 * 1024;
  static constexpr int STAGES_FIT = (MAX_SMEM - BAR_BYTES - EPI_SMEM_BYTES - STG_BYTES) / (A_BYTES + B_BYTES);
  static constexpr int STAGES = STAGES_FIT > 8 ? 8 : STAGES_FIT;  // 3 / 5 (CG=1), 5 / 7 (CG=2)
  static constexpr uint32_t SMEM_BYTES = STAGES * (A_BYTES + B_BYTES) + BAR_BYTES + EPI_SMEM_BYTES + STG_BYTES;
  static constexpr int THREADS = 384;
  static constexpr int EPI_THREADS = 256;
  static_assert(STAGES <= 8, "barrier area holds at most 8 stages");
};

// ---------------------------------------------------------------------------
// Epilogue for one 128 x BN accumulator tile; executed by the 8 epilogue warps.
// warp (quarter, half): TMEM lanes [32*quarter, +32), columns [half*BN/2, +BN/2).
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

template <int BN, int AMODE, int EPI>
__device__ __forceinline__ void epilogue_tile(const GemmParams& p, uint32_t taddr, int m_tile, int n_tile,
                                              int quarter, int half, float* epi_smem, float* stg) {
  constexpr int CH = BN / 2;
  const int lane = threadIdx.x & 31;
  const int colbase = n_tile * BN + half * CH;

  if constexpr (EPI == EPI_HEAD) {
    // head.2 epilogue: ReLU -> head.4 (1x1, 128 -> 4) -> postprocess (dpt_block.py:319-323,
    // postprocess.py:10-62).  Row mapping: each thread owns 64 of the 128 channels of one pixel; the two
    // column halves are combined through shared memory.  Nothing but 16 bytes per pixel is written.
    static_assert(BN == 128, "EPI_HEAD expects the 128-channel head");
    const int r_local = quarter * 32 + lane;
    long long orow;
    bool valid;
    tile_row<AMODE>(p, m_tile, r_local, orow, valid);
    float part[4] = {0.f, 0.f, 0.f, 0.f};
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
        part[0] = fmaf(v, w4.x, part[0]);
        part[1] = fmaf(v, w4.y, part[1]);
        part[2] = fmaf(v, w4.z, part[2]);
        part[3] = fmaf(v, w4.w, part[3]);
      }
    }
    if (half == 1) reinterpret_cast<float4*>(epi_smem)[r_local] = make_float4(part[0], part[1], part[2], part[3]);
    named_bar_sync(1, 256);
    if (half == 0) {
      float4 o = reinterpret_cast<float4*>(epi_smem)[r_local];
      const float x = part[0] + o.x + __ldg(p.head_b + 0);
      const float y = part[1] + o.y + __ldg(p.head_b + 1);
      const float z = part[2] + o.z + __ldg(p.head_b + 2);
      const float cf = part[3] + o.w + __ldg(p.head_b + 3);
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
    named_bar_sync(1, 256);
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
      } else if constexpr (EPI != EPI_ROPE) {  // RoPE adds its bias before rotating (row mapping)
        if (p.bias) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col + cc));
      }
      // Residual loads are issued up front: the in-place residual stream (resid == out) would otherwise force
      // load -> store -> load ordering (possible aliasing) and serialise ~800-cycle DRAM round trips.
      [[maybe_unused]] float4 rf[8];
      [[maybe_unused]] uint2 ra[8], rb[8];
      if constexpr (EPI == EPI_F32) {
        if (p.resid) {
#pragma unroll
          for (int it = 0; it < 8; ++it)
            rf[it] = ((vmask >> it) & 1u)
                         ? *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.resid) +
                                                            orow_c[it] * p.ldo + col + cc)
                         : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      if constexpr (EPI == EPI_BF16) {
        if (p.resid) {
#pragma unroll
          for (int it = 0; it < 8; ++it)
            ra[it] = ((vmask >> it) & 1u)
                         ? *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(p.resid) +
                                                           orow_c[it] * p.ldo + col + cc)
                         : make_uint2(0u, 0u);
        }
        if (p.resid2) {
#pragma unroll
          for (int it = 0; it < 8; ++it)
            rb[it] = ((vmask >> it) & 1u)
                         ? *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(p.resid2) +
                                                           orow_c[it] * p.ldo + col + cc)
                         : make_uint2(0u, 0u);
        }
      }
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        float4 v = *reinterpret_cast<const float4*>(stg + (it * 4 + crow) * kStageLd + cc);
        if (!((vmask >> it) & 1u)) continue;
        v.x += b4.x; v.y += b4.y; v.z += b4.z; v.w += b4.w;
        if constexpr (EPI == EPI_F32) {
          float* op = reinterpret_cast<float*>(p.out) + orow_c[it] * p.ldo + col + cc;
          if (p.resid) { v.x += rf[it].x; v.y += rf[it].y; v.z += rf[it].z; v.w += rf[it].w; }
          *reinterpret_cast<float4*>(op) = v;
        } else if constexpr (EPI == EPI_PIXSHUF) {
          __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.out) + (orow_c[it] + pix_off) * p.ps_cout + co + cc;
          *reinterpret_cast<uint2*>(op) = pack4_bf16(v);
        } else {
          const long long off = orow_c[it] * p.ldo + col + cc;
          if constexpr (EPI == EPI_BF16) {
            if (p.resid) add_bf16x4(v, ra[it]);
            if (p.resid2) add_bf16x4(v, rb[it]);
            if (p.relu_main) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
          }
          if constexpr (EPI == EPI_GELU) { v.x = gelu_erf(v.x); v.y = gelu_erf(v.y); v.z = gelu_erf(v.z); v.w = gelu_erf(v.w); }
          if (p.out) *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.out) + off) = pack4_bf16(v);
          if constexpr (EPI == EPI_BF16) {
            if (p.out2) {
              v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
              *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.out2) + off) = pack4_bf16(v);
            }
          }
        }
      }
    };

    if constexpr (EPI == EPI_ROPE) {
      static_assert(CH % 64 == 0, "RoPE epilogue works on whole 64-wide heads");
      // row mapping: bias + 2-D RoPE (rotate-half inside each 32-wide half of the head; reference:
      // pos_embed/pos_embed.py:149-185, curope/kernels.cu:17-82), then staged out like everything else
      long long orow_r;
      bool valid_r;
      tile_row<AMODE>(p, m_tile, quarter * 32 + lane, orow_r, valid_r);
      int py = 0, px = 0;
      if (valid_r) {
        py = p.pos[2 * orow_r + 0];
        px = p.pos[2 * orow_r + 1];
        if (py < -1 || py > p.rope_max_pos || px < -1 || px > p.rope_max_pos)
          device_fatal("token position outside the RoPE table");
      }
      const float4* ty = reinterpret_cast<const float4*>(p.rope_tab + static_cast<long long>(py + 1) * 32);
      const float4* tx = reinterpret_cast<const float4*>(p.rope_tab + static_cast<long long>(px + 1) * 32);
#pragma unroll 1
      for (int c = 0; c < CH; c += 64) {
        const int col = colbase + c;
        uint32_t a0[32], a1[32];
        tmem_ld32(taddr + c, a0);
        tmem_ld32(taddr + c + 32, a1);
        tmem_ld_wait();
        float v[64];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          v[i] = __uint_as_float(a0[i]);
          v[32 + i] = __uint_as_float(a1[i]);
        }
        if (col < p.N) {
          if (p.bias) {
            const float4* bp = reinterpret_cast<const float4*>(p.bias + col);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              float4 b = __ldg(bp + j);
              v[4 * j + 0] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
            }
          }
          if (col < p.rope_cols) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float4 t = __ldg(ty + j);  // (cos_{2j}, sin_{2j}, cos_{2j+1}, sin_{2j+1})
              float u0 = v[2 * j], w0 = v[2 * j + 16];
              v[2 * j] = u0 * t.x - w0 * t.y;
              v[2 * j + 16] = w0 * t.x + u0 * t.y;
              float u1 = v[2 * j + 1], w1 = v[2 * j + 17];
              v[2 * j + 1] = u1 * t.z - w1 * t.w;
              v[2 * j + 17] = w1 * t.z + u1 * t.w;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float4 t = __ldg(tx + j);
              float u0 = v[32 + 2 * j], w0 = v[32 + 2 * j + 16];
              v[32 + 2 * j] = u0 * t.x - w0 * t.y;
              v[32 + 2 * j + 16] = w0 * t.x + u0 * t.y;
              float u1 = v[32 + 2 * j + 1], w1 = v[32 + 2 * j + 17];
              v[32 + 2 * j + 1] = u1 * t.z - w1 * t.w;
              v[32 + 2 * j + 17] = w1 * t.z + u1 * t.w;
            }
          }
        }
        stage_store32(stg, lane, v);
        __syncwarp();
        flush(col);
        __syncwarp();
        stage_store32(stg, lane, v + 32);
        __syncwarp();
        flush(col + 32);
        __syncwarp();
      }
    } else {
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
}

template <int BN, int AMODE, int EPI, int CG>
__global__ void __launch_bounds__(384, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  using Cfg = GemmCfg<BN, CG>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr uint32_t A_BYTES = Cfg::A_BYTES;
  constexpr uint32_t B_BYTES = Cfg::B_BYTES;

  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * (A_BYTES + B_BYTES));
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* tfull = bars + 2 * STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float* epi_smem = reinterpret_cast<float*>(smem + STAGES * (A_BYTES + B_BYTES) + Cfg::BAR_BYTES);
  float* stg_all = reinterpret_cast<float*>(smem + STAGES * (A_BYTES + B_BYTES) + Cfg::BAR_BYTES + Cfg::EPI_SMEM_BYTES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = (CG == 2) ? cluster_ctarank() : 0u;
  const bool leader = (cta_rank == 0);

  // 128-row tiles; a work item covers CG consecutive ones
  const int m_tiles128 = (AMODE == A_CONV3) ? p.nimg * p.tiles_h * p.tiles_w : (p.M + 127) / 128;
  const int m_tiles = (m_tiles128 + CG - 1) / CG;
  const int n_tiles = (p.N + BN - 1) / BN;
  const int num_tiles = m_tiles * n_tiles;
  const int cpb = (AMODE == A_CONV3) ? (p.Cin / 64) : 1;  // 64-channel chunks per filter tap
  const int nkb = (AMODE == A_CONV3) ? 9 * cpb : (p.K + 63) / 64;
  const int first_tile = blockIdx.x / CG;
  const int tile_step = gridDim.x / CG;

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0) device_fatal("dynamic shared memory is not 1024-byte aligned");
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);   // the leader's producer arrives once (its expect_tx covers both CTAs' bytes)
      mbar_init(&empty[s], 1);  // one (multicast) tcgen05.commit
    }
    mbar_init(&tfull[0], 1);
    mbar_init(&tfull[1], 1);
    mbar_init(&tempty[0], 8 * CG);  // 8 epilogue warps per CTA, all arriving on the leader's barrier
    mbar_init(&tempty[1], 8 * CG);
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
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
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();               // previous kernel's outputs (A operand, residuals) are complete and visible
  pdl_launch_dependents();  // the next kernel may start its prologue

  if (warp == 0) {
    // ===================== TMA producer (one per CTA) =====================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = first_tile; tile < num_tiles; tile += tile_step) {
        const int m_tile = (tile / n_tiles) * CG + static_cast<int>(cta_rank);  // this CTA's 128-row tile
        const int n_tile = tile % n_tiles;
        const int b_row0 = n_tile * BN + static_cast<int>(cta_rank) * (BN / CG);
        int cn = 0, ch0 = 0, cw0 = 0;
        if constexpr (AMODE == A_CONV3) {
          const int tpi = p.tiles_h * p.tiles_w;
          cn = m_tile / tpi;  // >= nimg for the padding tile of an odd tile count: TMA zero-fills
          const int t = m_tile - cn * tpi;
          const int th = t / p.tiles_w;
          ch0 = th * 8;
          cw0 = (t - th * p.tiles_w) * 16;
        }
        for (int kb = 0; kb < nkb; ++kb) {
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
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader && elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(128 * CG, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = first_tile; tile < num_tiles; tile += tile_step) {
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint64_t adesc = make_smem_desc_sw128(smem_u32(sA + stage * A_BYTES));
          const uint64_t bdesc = make_smem_desc_sw128(smem_u32(sB + stage * B_BYTES));
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            // advance 16 K-elements = 32 bytes inside the 128B swizzle atom (encoded >> 4)
            if constexpr (CG == 2)
              umma_bf16_cg2(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            else
              umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          if constexpr (CG == 2) umma_commit_cg2(&empty[stage]); else umma_commit(&empty[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if constexpr (CG == 2) umma_commit_cg2(&tfull[acc]); else umma_commit(&tfull[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ============ epilogue warps (every CTA drains its own 128 accumulator rows) ============
    const int quarter = warp & 3;
    const int half = (warp - 4) >> 2;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = first_tile; tile < num_tiles; tile += tile_step) {
      const int m_tile = (tile / n_tiles) * CG + static_cast<int>(cta_rank);
      const int n_tile = tile % n_tiles;
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN + half * (BN / 2);
      epilogue_tile<BN, AMODE, EPI>(p, taddr, m_tile, n_tile, quarter, half, epi_smem,
                                    stg_all + (warp - 4) * kStageFloatsPerWarp);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CG == 1 || leader) mbar_arrive(&tempty[acc]); else mbar_arrive_remote(&tempty[acc], 0);
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
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
