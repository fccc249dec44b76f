// Flash-attention forward (head_dim 64, non-causal) on tcgen05 for sm_100a.
//
// Replaces, for one (sample, head, 128-query tile) per CTA:
//   * xformers memory_efficient_attention at vista_slam/sta_model/blocks/sta_blocks.py:143
//     (encoder / decoder self-attention), and
//   * the materialised softmax(q k^T * scale) v of CrossAttention, sta_blocks.py:201-205
//     (K/V taken from the partner sample: kv_batch_shift).
//
// Q, K, V are read straight out of the projection GEMM outputs with 3-D TMA
// tensor maps (64 head columns x tokens x samples); nothing is re-laid-out.
//   S = Q K^T        tcgen05.mma 128x128x64, K-major A and B, fp32 in TMEM
//   P = exp2(S*c - m) in registers (one query row per thread), written as bf16
//       into a 128B-swizzled K-major smem tile
//   O_j = P V_j      tcgen05.mma 128x64x128, V consumed MN-major exactly as TMA wrote it
//   running (m, l, O) rescaling in fp32 registers.
// Two CTAs fit per SM (112 KB smem, 256 TMEM columns) so one CTA's softmax
// overlaps the other's MMAs.
#include "common.cuh"
#include "host_util.h"
#include "ops.h"

namespace sta {

namespace {

constexpr int ATT_THREADS = 192;
constexpr uint32_t TILE_BYTES = 128 * 64 * 2;  // 16 KB: [128 rows][64 bf16]
constexpr uint32_t ATT_SMEM = TILE_BYTES * (1 + 2 + 2 + 2) + 256;

struct AttnParams {
  int nq, nk, batch, kv_batch_shift;
  float scale_log2;
  __nv_bfloat16* out;
  long long ldo;
};

__global__ void __launch_bounds__(ATT_THREADS, 2)
attention_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, const AttnParams p, int q_col0, int k_col0,
                     int v_col0) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sK = smem + TILE_BYTES;       // 2 stages
  uint8_t* sV = smem + 3 * TILE_BYTES;   // 2 stages
  uint8_t* sP = smem + 5 * TILE_BYTES;   // 2 sub-tiles of 64 keys
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 7 * TILE_BYTES);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;   // [2]
  uint64_t* v_full = bars + 3;   // [2]
  uint64_t* k_empty = bars + 5;  // [2]
  uint64_t* v_empty = bars + 7;  // [2]
  uint64_t* s_full = bars + 9;
  uint64_t* p_full = bars + 10;
  uint64_t* o_full = bars + 11;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128;
  const int head = blockIdx.y;
  const int b = blockIdx.z;
  const int kvb = (b + p.kv_batch_shift) % p.batch;
  const int T = (p.nk + 127) / 128;

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0) device_fatal("dynamic shared memory is not 1024-byte aligned");
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, 4);
    mbar_init(o_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base;
  const uint32_t tmem_O = tmem_base + 128;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmK);
      tma_prefetch_desc(&tmV);
      mbar_arrive_expect_tx(q_full, TILE_BYTES);
      tma_load_3d(sQ, &tmQ, q_full, q_col0 + head * 64, q0, b);
      for (int j = 0; j < T; ++j) {
        const int st = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        mbar_wait(&k_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&k_full[st], TILE_BYTES);
        tma_load_3d(sK + st * TILE_BYTES, &tmK, &k_full[st], k_col0 + head * 64, j * 128, kvb);
        mbar_wait(&v_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&v_full[st], TILE_BYTES);
        tma_load_3d(sV + st * TILE_BYTES, &tmV, &v_full[st], v_col0 + head * 64, j * 128, kvb);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idesc_o = make_idesc_bf16(128, 64, 0, 1);  // B (= V) is MN-major
      const uint64_t qdesc = make_smem_desc_sw128(smem_u32(sQ));
      const uint64_t pdesc0 = make_smem_desc_sw128(smem_u32(sP));
      const uint64_t pdesc1 = make_smem_desc_sw128(smem_u32(sP + TILE_BYTES));
      mbar_wait(q_full, 0);
      // S(0)
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      {
        const uint64_t kdesc = make_smem_desc_sw128(smem_u32(sK));
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem_S, qdesc + 2 * k, kdesc + 2 * k, idesc_s, k != 0);
        umma_commit(&k_empty[0]);
        umma_commit(s_full);
      }
      for (int j = 0; j < T; ++j) {
        const int st = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        // O_j = P_j V_j
        mbar_wait(p_full, j & 1);
        mbar_wait(&v_full[st], ph);
        tc_fence_after();
        const uint64_t vdesc = make_smem_desc_sw128(smem_u32(sV + st * TILE_BYTES));
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint64_t pd = (k < 4 ? pdesc0 : pdesc1) + 2 * (k & 3);
          // V advances 16 keys = 16 rows x 128 B = 2048 B per step
          umma_bf16(tmem_O, pd, vdesc + (2048 >> 4) * k, idesc_o, k != 0);
        }
        umma_commit(&v_empty[st]);
        umma_commit(o_full);
        // S(j+1) may start as soon as S(j) has been consumed (= p_full(j))
        if (j + 1 < T) {
          const int st1 = (j + 1) & 1;
          const uint32_t ph1 = ((j + 1) >> 1) & 1;
          mbar_wait(&k_full[st1], ph1);
          tc_fence_after();
          const uint64_t kdesc = make_smem_desc_sw128(smem_u32(sK + st1 * TILE_BYTES));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(tmem_S, qdesc + 2 * k, kdesc + 2 * k, idesc_s, k != 0);
          umma_commit(&k_empty[st1]);
          umma_commit(s_full);
        }
      }
    }
  } else {
    // ===================== softmax warps (one query row per thread) =====================
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    float m = -INFINITY, l = 0.f;
    float o_acc[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) o_acc[i] = 0.f;
    uint8_t* prow = sP + r * 128;
    const int rx = r & 7;

    for (int j = 0; j < T; ++j) {
      const int nvalid = p.nk - j * 128;  // >= 1
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      // ---- pass 1: row max ----
      float mx = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < 128; c += 32) {
        uint32_t s[32];
        tmem_ld32(tmem_S + lane_addr + c, s);
        tmem_ld_wait();
        if (c + 32 <= nvalid) {
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(s[i]));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c + i < nvalid) mx = fmaxf(mx, __uint_as_float(s[i]));
        }
      }
      const float m_new = fmaxf(m, mx * p.scale_log2);
      const float alpha = ex2_approx(m - m_new);  // first tile: exp2(-inf) = 0
      // ---- fold O(j-1) and rescale ----
      if (j > 0) {
        mbar_wait(o_full, (j - 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < 64; c += 32) {
          uint32_t o[32];
          tmem_ld32(tmem_O + lane_addr + c, o);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o_acc[c + i] = (o_acc[c + i] + __uint_as_float(o[i])) * alpha;
        }
      }
      // ---- pass 2: P = exp2(S*c - m_new), write bf16 to swizzled smem ----
      float rowsum = 0.f;
#pragma unroll 1
      for (int c = 0; c < 128; c += 32) {
        uint32_t s[32];
        tmem_ld32(tmem_S + lane_addr + c, s);
        tmem_ld_wait();
        float pv[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float e = ex2_approx(fmaf(__uint_as_float(s[i]), p.scale_log2, -m_new));
          if (c + i >= nvalid) e = 0.f;
          pv[i] = e;
          rowsum += e;
        }
        uint8_t* sub = prow + (c >> 6) * TILE_BYTES;
        const int chunk0 = (c & 63) >> 3;  // 16-byte chunk index inside the 128-byte row
#pragma unroll
        for (int jc = 0; jc < 4; ++jc) {
          uint4 q;
          q.x = pack_bf16x2(pv[8 * jc + 0], pv[8 * jc + 1]);
          q.y = pack_bf16x2(pv[8 * jc + 2], pv[8 * jc + 3]);
          q.z = pack_bf16x2(pv[8 * jc + 4], pv[8 * jc + 5]);
          q.w = pack_bf16x2(pv[8 * jc + 6], pv[8 * jc + 7]);
          *reinterpret_cast<uint4*>(sub + (((chunk0 + jc) ^ rx) << 4)) = q;
        }
      }
      l = l * alpha + rowsum;
      m = m_new;
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }
    // ---- last O tile, normalise, store ----
    mbar_wait(o_full, (T - 1) & 1);
    tc_fence_after();
    const float inv_l = 1.0f / l;
    const int qrow = q0 + r;
    __nv_bfloat16* orow = p.out + (static_cast<long long>(b) * p.nq + qrow) * p.ldo + head * 64;
#pragma unroll
    for (int c = 0; c < 64; c += 32) {
      uint32_t o[32];
      tmem_ld32(tmem_O + lane_addr + c, o);
      tmem_ld_wait();
      if (qrow < p.nq) {
#pragma unroll
        for (int jc = 0; jc < 4; ++jc) {
          uint4 q;
          q.x = pack_bf16x2((o_acc[c + 8 * jc + 0] + __uint_as_float(o[8 * jc + 0])) * inv_l,
                            (o_acc[c + 8 * jc + 1] + __uint_as_float(o[8 * jc + 1])) * inv_l);
          q.y = pack_bf16x2((o_acc[c + 8 * jc + 2] + __uint_as_float(o[8 * jc + 2])) * inv_l,
                            (o_acc[c + 8 * jc + 3] + __uint_as_float(o[8 * jc + 3])) * inv_l);
          q.z = pack_bf16x2((o_acc[c + 8 * jc + 4] + __uint_as_float(o[8 * jc + 4])) * inv_l,
                            (o_acc[c + 8 * jc + 5] + __uint_as_float(o[8 * jc + 5])) * inv_l);
          q.w = pack_bf16x2((o_acc[c + 8 * jc + 6] + __uint_as_float(o[8 * jc + 6])) * inv_l,
                            (o_acc[c + 8 * jc + 7] + __uint_as_float(o[8 * jc + 7])) * inv_l);
          *reinterpret_cast<uint4*>(orow + c + 8 * jc) = q;
        }
      }
    }
    tc_fence_before();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

int make_qkv_map(CUtensorMap* m, const bf16* base, long long ld, int ntok, int batch) {
  uint64_t dims[3] = {(uint64_t)ld, (uint64_t)ntok, (uint64_t)batch};
  uint64_t strides[2] = {(uint64_t)ld * 2, (uint64_t)ld * 2 * (uint64_t)ntok};
  uint32_t box[3] = {64, 128, 1};
  return make_tmap_bf16(m, base, 3, dims, strides, box);
}

}  // namespace

int launch_attention(const AttnLaunch& a, cudaStream_t stream) {
  STA_REQUIRE(a.batch > 0 && a.heads > 0 && a.nq > 0 && a.nk > 0, "empty attention problem");
  STA_REQUIRE(a.ldq % 8 == 0 && a.ldk % 8 == 0 && a.ldv % 8 == 0 && a.ldo % 8 == 0, "row strides must be 16B multiples");
  STA_REQUIRE(a.q_col0 % 8 == 0 && a.k_col0 % 8 == 0 && a.v_col0 % 8 == 0, "column offsets must be 16B multiples");
  static bool attr_set = false;
  if (!attr_set) {
    STA_CHECK_CUDA(cudaFuncSetAttribute(attention_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)ATT_SMEM));
    attr_set = true;
  }
  CUtensorMap tmQ, tmK, tmV;
  if (make_qkv_map(&tmQ, a.q, a.ldq, a.nq, a.batch)) return 1;
  if (make_qkv_map(&tmK, a.k, a.ldk, a.nk, a.batch)) return 1;
  if (make_qkv_map(&tmV, a.v, a.ldv, a.nk, a.batch)) return 1;
  AttnParams p;
  p.nq = a.nq;
  p.nk = a.nk;
  p.batch = a.batch;
  p.kv_batch_shift = a.kv_batch_shift;
  p.scale_log2 = a.scale * 1.4426950408889634f;
  p.out = a.out;
  p.ldo = a.ldo;
  dim3 grid((a.nq + 127) / 128, a.heads, a.batch);
  attention_fwd_kernel<<<grid, ATT_THREADS, ATT_SMEM, stream>>>(tmQ, tmK, tmV, p, a.q_col0, a.k_col0, a.v_col0);
  STA_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace sta
