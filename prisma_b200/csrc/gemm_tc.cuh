// prisma_b200 -- the tcgen05 "shifted-row GEMM" core (sm_100a).
//
//   D[m, n] = sum_{t < taps} sum_{c < Cin} A[m + tap_off[t], c] * W[n, t * kchunks*64 + c]        (fp16 x fp16 -> fp32)
//
// taps == 1, tap_off = {0}  : a plain GEMM (ViT linears, 1x1 convs, transposed convs with k == s, RAFT correlation)
// taps == kh*kw             : an implicit-GEMM convolution over a zero-bordered NHWC activation whose pixels are
//                             flattened to rows; tap (ky,kx) is the same 2-D TMA box shifted by (ky-ph)*Wp + (kx-pw)
//                             rows.  Border rows are computed and discarded (waste 2/(W+2)); no im2col buffer, no 4-D
//                             tensor map; TMA zero-fills out-of-range rows (either sign).
//
// Kernel anatomy (persistent, warp-specialised, one CTA per SM):
//   warp 0   : TMA producer  (A 128x64 and W BNx64 fp16 tiles, 128B swizzle, STAGES-deep mbarrier ring)
//   warp 1   : MMA issuer    (one lane: tcgen05.mma cta_group::1 kind::f16, M=128, N=BN, K=16; fp32 accum in TMEM,
//                             two accumulator stages so the epilogue of tile i overlaps the mainloop of tile i+1)
//   warps 2-5: epilogue      (tcgen05.ld 32x32b -> registers -> bias / GELU / ReLU / LayerScale / residuals ->
//                             fp32 and/or fp16 stores with the row mapping of the consumer's layout)
#pragma once
#include "common.cuh"

namespace prisma {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_THREADS = 192;
constexpr int GEMM_MAX_TAPS = 49;

enum RowMap : int {
  ROW_LINEAR = 0,   // dst row = m                                   (valid: m < M)
  ROW_PADDED = 1,   // m indexes a zero-bordered [Hp][Wp] image; only interior pixels are stored (borders stay zero)
  ROW_TOK2PAD = 2,  // m = y*W + x (dense tokens)  -> dst row (y+1)*out_wp + x+1
  ROW_SHUFFLE = 3,  // ConvTranspose k == s: m = y*W + x, n = (dy*s+dx)*cout + co -> dst row (y*s+dy+1)*out_wp + x*s+dx+1
};

struct GemmEpilogue {
  const float* bias = nullptr;   // [N]
  const float* gamma = nullptr;  // [N]  LayerScale
  int act = 0;                   // 0 none, 1 exact-erf GELU, 2 ReLU
  const float* res_f32 = nullptr;  // + residual (fp32), indexed by dst row
  int res_f32_ld = 0;
  const __half* res_a = nullptr;  // + residual (fp16), indexed by dst row
  int res_a_ld = 0;
  const __half* res_b = nullptr;
  int res_b_ld = 0;
  float* out_f32 = nullptr;
  int out_f32_ld = 0;
  __half* out_f16 = nullptr;
  int out_f16_ld = 0;
  __half* out_f16_relu = nullptr;  // relu(result) copy, for consumers that take relu(x) as operand and x as skip
  int out_f16_relu_ld = 0;
  int row_map = ROW_LINEAR;
  int img_rows = 0;  // ROW_PADDED: rows per image (Hp*Wp); 0 = single image
  int in_w = 0;      // ROW_PADDED: Wp ; ROW_TOK2PAD / ROW_SHUFFLE: W
  int in_h = 0;      // ROW_PADDED: Hp ; ROW_TOK2PAD / ROW_SHUFFLE: H (rows per image = H*W)
  int out_wp = 0;    // destination padded width
  int out_img_rows = 0;  // destination rows per image
  int sub = 1;       // ROW_PADDED: keep every sub-th pixel (stride-2 conv evaluated at stride 1)
  int shuf_s = 1;
  int shuf_cout = 0;
  // fused DPT output head (dpt.py:96-100): depth = relu(head_b + sum_j head_w[j] * relu(acc[j] + bias[j])), N == 32,
  // written dense fp32 [H][W] (ROW_PADDED input geometry)
  const float* head_w = nullptr;
  float head_b = 0.f;
  float* head_out = nullptr;
};

struct GemmArgs {
  int M;  // rows of the output space
  int N;  // output columns
  int taps;
  int kchunks;  // 64-wide K blocks per tap
  int tap_off[GEMM_MAX_TAPS];
  GemmEpilogue ep;
};

template <int BN>
struct GemmCfg {
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
  static constexpr int B_BYTES = BN * GEMM_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN >= 256) ? 4 : (BN >= 128 ? 6 : 8);
  static constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64 ? 64 : (2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512)));
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

#ifdef __CUDACC__
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

// One 8-column group of one accumulator row -> epilogue math -> stores.
__device__ __forceinline__ void epilogue_store8(const GemmEpilogue& ep, float* v, long long drow, int n) {
  if (ep.bias) {
    const float4 b0 = *reinterpret_cast<const float4*>(ep.bias + n);
    const float4 b1 = *reinterpret_cast<const float4*>(ep.bias + n + 4);
    v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
    v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
  }
  if (ep.act == 1) {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = gelu_erf(v[j]);
  } else if (ep.act == 2) {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.0f);
  }
  if (ep.gamma) {
    const float4 g0 = *reinterpret_cast<const float4*>(ep.gamma + n);
    const float4 g1 = *reinterpret_cast<const float4*>(ep.gamma + n + 4);
    v[0] *= g0.x; v[1] *= g0.y; v[2] *= g0.z; v[3] *= g0.w;
    v[4] *= g1.x; v[5] *= g1.y; v[6] *= g1.z; v[7] *= g1.w;
  }
  int col = n;
  if (ep.row_map == ROW_SHUFFLE) col = n % ep.shuf_cout;
  if (ep.res_f32) {
    const float* p = ep.res_f32 + drow * ep.res_f32_ld + col;
    const float4 r0 = *reinterpret_cast<const float4*>(p);
    const float4 r1 = *reinterpret_cast<const float4*>(p + 4);
    v[0] += r0.x; v[1] += r0.y; v[2] += r0.z; v[3] += r0.w;
    v[4] += r1.x; v[5] += r1.y; v[6] += r1.z; v[7] += r1.w;
  }
  if (ep.res_a) {
    const uint4 r = *reinterpret_cast<const uint4*>(ep.res_a + drow * ep.res_a_ld + col);
    const __half2* h = reinterpret_cast<const __half2*>(&r);
#pragma unroll
    for (int j = 0; j < 4; ++j) { const float2 f = __half22float2(h[j]); v[2 * j] += f.x; v[2 * j + 1] += f.y; }
  }
  if (ep.res_b) {
    const uint4 r = *reinterpret_cast<const uint4*>(ep.res_b + drow * ep.res_b_ld + col);
    const __half2* h = reinterpret_cast<const __half2*>(&r);
#pragma unroll
    for (int j = 0; j < 4; ++j) { const float2 f = __half22float2(h[j]); v[2 * j] += f.x; v[2 * j + 1] += f.y; }
  }
  if (ep.out_f32) {
    float* p = ep.out_f32 + drow * ep.out_f32_ld + col;
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
  if (ep.out_f16) {
    uint4 o;
    o.x = pack_half2(v[0], v[1]); o.y = pack_half2(v[2], v[3]); o.z = pack_half2(v[4], v[5]); o.w = pack_half2(v[6], v[7]);
    *reinterpret_cast<uint4*>(ep.out_f16 + drow * ep.out_f16_ld + col) = o;
  }
  if (ep.out_f16_relu) {
    uint4 o;
    o.x = pack_half2(fmaxf(v[0], 0.f), fmaxf(v[1], 0.f)); o.y = pack_half2(fmaxf(v[2], 0.f), fmaxf(v[3], 0.f));
    o.z = pack_half2(fmaxf(v[4], 0.f), fmaxf(v[5], 0.f)); o.w = pack_half2(fmaxf(v[6], 0.f), fmaxf(v[7], 0.f));
    *reinterpret_cast<uint4*>(ep.out_f16_relu + drow * ep.out_f16_relu_ld + col) = o;
  }
}

template <int BN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ GemmArgs args) {
  using Cfg = GemmCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * Cfg::A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* tfull = bars + 2 * STAGES;
  uint64_t* tempty = bars + 2 * STAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tiles_m = (args.M + GEMM_BM - 1) / GEMM_BM;
  const int tiles_n = (args.N + BN - 1) / BN;
  const int num_tiles = tiles_m * tiles_n;
  const int num_kb = args.taps * args.kchunks;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], 4); }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile % tiles_m) * GEMM_BM;
        const int n0 = (tile / tiles_m) * BN;
        int tap = 0, chunk = 0;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full[stage], Cfg::STAGE_BYTES);
          tma_load_2d(sA + stage * Cfg::A_BYTES, &tmA, &full[stage], chunk * GEMM_BK, m0 + args.tap_off[tap]);
          tma_load_2d(sB + stage * Cfg::B_BYTES, &tmB, &full[stage], kb * GEMM_BK, n0);
          if (++chunk == args.kchunks) { chunk = 0; ++tap; }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_f16(GEMM_BM, BN);
      int stage = 0; uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(&tempty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint64_t adesc = make_sdesc_sw128(smem_u32(sA + stage * Cfg::A_BYTES));
          const uint64_t bdesc = make_sdesc_sw128(smem_u32(sB + stage * Cfg::B_BYTES));
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) {
            // advance 16 fp16 = 32 B along K inside the 128 B swizzled row: +2 in the (addr >> 4) field
            umma_f16(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull[as]);
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps (2..5)
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    const GemmEpilogue& ep = args.ep;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int m0 = (tile % tiles_m) * GEMM_BM;
      const int n0 = (tile / tiles_m) * BN;
      const int m = m0 + quarter * 32 + lane;
      // ---- row mapping
      bool valid = m < args.M;
      long long drow = m;
      int ty = 0, tx = 0;  // token coordinates for ROW_SHUFFLE
      long long img_base = 0;
      if (ep.row_map == ROW_PADDED) {
        int img = 0, r = m;
        if (ep.img_rows > 0) { img = m / ep.img_rows; r = m - img * ep.img_rows; }
        const int y = r / ep.in_w, x = r - y * ep.in_w;
        valid = valid && y >= 1 && y <= ep.in_h - 2 && x >= 1 && x <= ep.in_w - 2;
        if (ep.sub > 1) {
          valid = valid && ((y - 1) % ep.sub == 0) && ((x - 1) % ep.sub == 0);
          drow = (long long)img * ep.out_img_rows + (long long)((y - 1) / ep.sub + 1) * ep.out_wp + (x - 1) / ep.sub + 1;
        }
      } else if (ep.row_map == ROW_TOK2PAD || ep.row_map == ROW_SHUFFLE) {
        const int per = ep.in_w * ep.in_h;
        const int img = m / per, r = m - img * per;
        ty = r / ep.in_w; tx = r - ty * ep.in_w;
        img_base = (long long)img * ep.out_img_rows;
        drow = img_base + (long long)(ty + 1) * ep.out_wp + tx + 1;
      }
      mbar_wait(&tfull[as], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + as * BN;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        if (n0 + c0 >= args.N) break;  // warp-uniform
        uint32_t r[32];
        tmem_ld32(taddr + c0, r);
        tmem_ld_wait();
        if (valid && ep.head_w != nullptr) {
          float acc = ep.head_b;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            acc = fmaf(fmaxf(__uint_as_float(r[j]) + __ldg(ep.bias + j), 0.f), __ldg(ep.head_w + j), acc);
          const int y = m / ep.in_w, x = m - y * ep.in_w;
          ep.head_out[(size_t)(y - 1) * (ep.in_w - 2) + (x - 1)] = fmaxf(acc, 0.f);
        } else if (valid) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int n = n0 + c0 + g * 8;
            if (n < args.N) {
              float v[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[g * 8 + j]);
              long long dr = drow;
              if (ep.row_map == ROW_SHUFFLE) {
                const int q = n / ep.shuf_cout;
                const int dy = q / ep.shuf_s, dx = q - dy * ep.shuf_s;
                dr = img_base + (long long)(ty * ep.shuf_s + dy + 1) * ep.out_wp + tx * ep.shuf_s + dx + 1;
              }
              epilogue_store8(ep, v, dr, n);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[as]);
    }
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}
#endif  // __CUDACC__

// Host-side prepared launch: tensor maps encoded once, replayed every frame (also inside CUDA graphs).
struct GemmLaunch {
  CUtensorMap tmA, tmB;
  GemmArgs args;
  int bn = 128;
  int grid = 1;
  double flops = 0;  // algorithmic 2*M*N*K (excluding border/padding waste)
};

// A: fp16 [a_rows][a_cols] with pitch a_pitch (elements);  W: fp16 [w_rows >= round_up(N, bn)][taps*kchunks*64]
int gemm_prepare(GemmLaunch* out, const __half* A, long long a_rows, int a_cols, int a_pitch, const __half* W,
                 int w_rows, int M, int N, int taps, const int* tap_off, const GemmEpilogue& ep, int num_sms,
                 int force_bn = 0);
int gemm_run(const GemmLaunch& g, cudaStream_t stream);
int gemm_pick_bn(int M, int N, int num_sms);

}  // namespace prisma
