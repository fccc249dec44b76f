// Micro-probe (not part of the product): what one / two warps per SM sub-partition reach on the MUFU.EX2 pipe, alone and in
// the instruction mix of the attention softmax (FFMA -> EX2 -> FADD row sum -> F2FP pack).  Prints clocks per warp-level EX2.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probe/mufu_probe tools/probe/mufu_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_bf16.h>
__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t pack(float a, float b) { uint32_t r; asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a)); return r; }

template <int MODE>
__global__ void probe(float* out, long long* clk, int iters, float scale, float m) {
  float s[128];
#pragma unroll
  for (int i = 0; i < 128; ++i) s[i] = -0.01f * (float)((threadIdx.x * 7 + i * 13) & 255);
  float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
  uint32_t x = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {  // EX2 only, 128 independent chains
#pragma unroll
      for (int i = 0; i < 128; ++i) s[i] = ex2(s[i]);
    } else if (MODE == 1) {  // the softmax mix
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        float e[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) e[i] = ex2(fmaf(s[8 * c + i], scale, -m));
        acc0 += e[0] + e[4]; acc1 += e[1] + e[5]; acc2 += e[2] + e[6]; acc3 += e[3] + e[7];
        x ^= pack(e[0], e[1]) ^ pack(e[2], e[3]) ^ pack(e[4], e[5]) ^ pack(e[6], e[7]);
      }
      m += 1e-7f;
    } else if (MODE == 2) {  // mix without the packs
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        float e[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) e[i] = ex2(fmaf(s[8 * c + i], scale, -m));
        acc0 += e[0] + e[4]; acc1 += e[1] + e[5]; acc2 += e[2] + e[6]; acc3 += e[3] + e[7];
      }
      m += 1e-7f;
    } else if (MODE == 3) {  // EX2 + pack only
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        float e[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) e[i] = ex2(s[8 * c + i]);
#pragma unroll
        for (int i = 0; i < 8; ++i) s[8 * c + i] = -e[i];
        x ^= pack(e[0], e[1]) ^ pack(e[2], e[3]) ^ pack(e[4], e[5]) ^ pack(e[6], e[7]);
      }
    } else if (MODE == 5 || MODE == 6 || MODE == 7) {  // 5: mix without the row sum; 6: row sum + truncating PRMT pack; 7: PRMT pack only
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        float e[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) e[i] = ex2(fmaf(s[8 * c + i], scale, -m));
        if (MODE == 6) { acc0 += e[0] + e[4]; acc1 += e[1] + e[5]; acc2 += e[2] + e[6]; acc3 += e[3] + e[7]; }
        if (MODE == 5) {
          x ^= pack(e[0], e[1]) ^ pack(e[2], e[3]) ^ pack(e[4], e[5]) ^ pack(e[6], e[7]);
        } else {
          x ^= __byte_perm(__float_as_uint(e[0]), __float_as_uint(e[1]), 0x7632) ^
               __byte_perm(__float_as_uint(e[2]), __float_as_uint(e[3]), 0x7632) ^
               __byte_perm(__float_as_uint(e[4]), __float_as_uint(e[5]), 0x7632) ^
               __byte_perm(__float_as_uint(e[6]), __float_as_uint(e[7]), 0x7632);
        }
      }
      m += 1e-7f;
    } else if (MODE == 8) {  // packed fp32: FFMA2 for the scale-subtract, FADD2 for the row sums
      const float2 sc2 = make_float2(scale, scale), nm2 = make_float2(-m, -m);
      const uint64_t scp = *reinterpret_cast<const uint64_t*>(&sc2), nmp = *reinterpret_cast<const uint64_t*>(&nm2);
      uint64_t a01 = 0, a23 = 0;
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        float e[8];
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
          const float2 sv = make_float2(s[8 * c + i], s[8 * c + i + 1]);
          uint64_t t;
          asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(t) : "l"(*reinterpret_cast<const uint64_t*>(&sv)), "l"(scp), "l"(nmp));
          const float2 tv = *reinterpret_cast<float2*>(&t);
          e[i] = ex2(tv.x);
          e[i + 1] = ex2(tv.y);
        }
#pragma unroll
        for (int i = 0; i < 8; i += 4) {
          const float2 p0 = make_float2(e[i], e[i + 1]), p1 = make_float2(e[i + 2], e[i + 3]);
          asm("add.rn.f32x2 %0, %0, %1;" : "+l"(a01) : "l"(*reinterpret_cast<const uint64_t*>(&p0)));
          asm("add.rn.f32x2 %0, %0, %1;" : "+l"(a23) : "l"(*reinterpret_cast<const uint64_t*>(&p1)));
        }
        x ^= pack(e[0], e[1]) ^ pack(e[2], e[3]) ^ pack(e[4], e[5]) ^ pack(e[6], e[7]);
      }
      const float2 r01 = *reinterpret_cast<float2*>(&a01), r23 = *reinterpret_cast<float2*>(&a23);
      acc0 += r01.x; acc1 += r01.y; acc2 += r23.x; acc3 += r23.y;
      m += 1e-7f;
    } else if (MODE == 4) {  // 3/4 MUFU + 1/4 polynomial on the FMA pipe, mix as MODE 1
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        float e[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float t = fmaf(s[8 * c + i], scale, -m);
          if ((i & 3) == 3) {
            const float tc = fmaxf(t, -126.f);
            const float r = tc + 12582912.f;  // round to nearest integer
            const float n = r - 12582912.f;
            const float f = tc - n;
            float p = fmaf(0.0551716685f, f, 0.242611125f);
            p = fmaf(p, f, 0.693260968f);
            p = fmaf(p, f, 0.999928057f);
            e[i] = __int_as_float(__float_as_int(p) + (__float_as_int(r) << 23));
          } else {
            e[i] = ex2(t);
          }
        }
        acc0 += e[0] + e[4]; acc1 += e[1] + e[5]; acc2 += e[2] + e[6]; acc3 += e[3] + e[7];
        x ^= pack(e[0], e[1]) ^ pack(e[2], e[3]) ^ pack(e[4], e[5]) ^ pack(e[6], e[7]);
      }
      m += 1e-7f;
    }
  }
  const long long t1 = clock64();
  float r = acc0 + acc1 + acc2 + acc3 + __uint_as_float(x & 0x3f800000u);
#pragma unroll
  for (int i = 0; i < 128; ++i) r += s[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int threads, float mufu_frac) {
  float* out; long long* clk;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&clk, 148 * 8);
  const int iters = 2000;
  probe<MODE><<<148, threads>>>(out, clk, 10, 1.0f, 0.5f);
  probe<MODE><<<148, threads>>>(out, clk, iters, 1.0f, 0.5f);
  cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < 148; ++i) avg += (double)h[i]; avg /= 148;
  const double warps_per_smsp = threads / 128.0;
  printf("%-28s threads %4d: %7.2f clk per 128-element row per warp; %5.2f clk per warp-EX2 per SMSP (8.0 = pipe peak)  [%s]\n", name, threads,
         avg / iters, avg / iters / (128.0 * mufu_frac) / warps_per_smsp, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out); cudaFree(clk);
}
int main() {
  for (int t : {128, 256}) {
    run<0>("ex2 only", t, 1.f);
    run<3>("ex2 + pack", t, 1.f);
    run<2>("ffma + ex2 + rowsum", t, 1.f);
    run<1>("ffma + ex2 + rowsum + pack", t, 1.f);
    run<4>("same, 1/4 on the FMA pipe", t, 0.75f);
    run<5>("ffma + ex2 + pack (no row sum)", t, 1.f);
    run<6>("ffma + ex2 + rowsum + prmt", t, 1.f);
    run<7>("ffma + ex2 + prmt", t, 1.f);
    run<8>("ffma2 + ex2 + fadd2 rowsum + pack", t, 1.f);
  }
  return 0;
}
