// prisma_b200 -- ViT self-attention on tcgen05 (sm_100a): softmax(q k^T) v, non-causal, head_dim 64
// (dinov2/layers/attention.py:49-62; the 1/sqrt(64) scale is folded into the q rows of the qkv weights, exactly).
//
// One CTA per (128-query tile, head, image); 2 CTAs co-reside per SM so one CTA's softmax overlaps the other's MMAs.
//   warp 0   : TMA producer  (Q once; K/V 128x64 fp16 tiles in a 2-stage ring, 128B swizzle, straight out of the
//                             [tokens][3*D] qkv matrix -- no head split/transposes ever materialise)
//   warp 1   : MMA issuer    (S = Q K^T  -> TMEM cols [0,128);  O += P V -> TMEM cols [128,192): A = P read from TMEM
//                             cols [192,256) (fp16 pairs written there by the softmax warps with tcgen05.st), B = V
//                             from smem as an MN-major operand.  P never goes through shared memory: the SM's 128 B/clk
//                             smem port is then only used by TMA fills and the Q/K/V operand reads)
//   warps 2-9: softmax       (thread = half a query row -- warps w and w+4 share a TMEM lane quarter and take the
//                             column halves, so 4 warps per scheduler keep the MUFU pipe fed: tcgen05.ld S, online max/sum in fp32, P -> fp16 -> swizzled smem
//                             as the A operand of the PV MMA).  O accumulates in TMEM across all KV tiles; the
//                             running-max rescale is lazy: O (and l) are only rescaled -- tcgen05.ld/st of this
//                             warp's 32 lanes -- when a row's max grew by more than 2^8, so P stays <= 256 in fp16
//                             and the common path has no per-tile accumulator traffic at all.
#include "attention.cuh"

namespace prisma {

constexpr int ATT_BQ = 128, ATT_BKV = 128, ATT_HD = 64;
constexpr int ATT_THREADS = 320;  // TMA warp + MMA warp + 8 softmax warps (two per TMEM lane quarter: column halves)
constexpr int ATT_SOFTMAX_WARP0 = 2;
constexpr int ATT_TILE_BYTES = 128 * 64 * 2;  // 16 KB
// 7 tiles + barriers = 114,816 B: two CTAs (+1 KB reserved each) fit the 228 KB of an SM; no alignment slack, the
// dynamic smem base is declared 1024-aligned (128B-swizzle atoms are 1 KB) and checked at kernel entry.
// Q + 2 x (K, V) + barriers + row-max exchange; P never touches shared memory (it is the TMEM A operand of the PV MMA)
constexpr int ATT_SMEM = ATT_TILE_BYTES * (1 + 2 + 2) + 128 /*barriers*/ + 1024 /*row-max exchange, two parities*/;

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(ATT_THREADS, 2)
attention_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ AttnArgs args) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) {
    if (threadIdx.x == 0) printf("prisma: attention smem base not 1024-aligned\n");
    __trap();
  }
  uint8_t* sQ = smem;
  uint8_t* sK = smem + ATT_TILE_BYTES;      // 2 stages
  uint8_t* sV = smem + 3 * ATT_TILE_BYTES;  // 2 stages
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 5 * ATT_TILE_BYTES);
  uint64_t* q_full = bars + 0;
  uint64_t* kv_full = bars + 1;   // [2]
  uint64_t* kv_empty = bars + 3;  // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* p_full = bars + 6;
  uint64_t* pv_done = bars + 7;
  uint64_t* s_free = bars + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);
  __half* s_xmax = reinterpret_cast<__half*>(smem + 5 * ATT_TILE_BYTES + 128);  // [2 parities][2 halves][128 rows]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T = args.tokens, D = args.D;
  const int q0 = blockIdx.x * ATT_BQ;
  const int head = blockIdx.y;
  const int row_base = blockIdx.z * T;  // first row of this image in the qkv matrix
  const int n_kv = (T + ATT_BKV - 1) / ATT_BKV;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmQKV);
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
    mbar_init(s_full, 1);
    mbar_init(p_full, 8);
    mbar_init(pv_done, 1);
    mbar_init(s_free, 8);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base;
  const uint32_t tmem_O = tmem_base + 128;
  const uint32_t tmem_P = tmem_base + 192;  // P as packed fp16 pairs: 64 columns = 128 keys

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, ATT_TILE_BYTES);
      tma_load_2d(sQ, &tmQKV, q_full, head * ATT_HD, row_base + q0);
      for (int j = 0; j < n_kv; ++j) {
        const int st = j & 1;
        mbar_wait(&kv_empty[st], ((j >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&kv_full[st], 2 * ATT_TILE_BYTES);
        tma_load_2d(sK + st * ATT_TILE_BYTES, &tmQKV, &kv_full[st], D + head * ATT_HD, row_base + j * ATT_BKV);
        tma_load_2d(sV + st * ATT_TILE_BYTES, &tmQKV, &kv_full[st], 2 * D + head * ATT_HD, row_base + j * ATT_BKV);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_f16(128, 128, 0, 0);
      constexpr uint32_t idesc_o = make_idesc_f16(128, 64, 0, 1);  // B (= V) is MN-major
      const uint64_t qdesc = make_sdesc_sw128(smem_u32(sQ));
      mbar_wait(q_full, 0);
      mbar_wait(&kv_full[0], 0);
      tc_fence_after();
      {
        const uint64_t kdesc = make_sdesc_sw128(smem_u32(sK));
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(tmem_S, qdesc + 2 * k, kdesc + 2 * k, idesc_s, k != 0);
        umma_commit(s_full);
      }
      for (int j = 0; j < n_kv; ++j) {
        const int st = j & 1;
        if (j + 1 < n_kv) {
          // S(j+1) as soon as the softmax warps hold S(j) in registers: it overlaps the whole softmax of tile j
          const int st1 = (j + 1) & 1;
          mbar_wait(s_free, j & 1);
          mbar_wait(&kv_full[st1], ((j + 1) >> 1) & 1);
          tc_fence_after();
          const uint64_t kdesc = make_sdesc_sw128(smem_u32(sK + st1 * ATT_TILE_BYTES));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(tmem_S, qdesc + 2 * k, kdesc + 2 * k, idesc_s, k != 0);
          umma_commit(s_full);
        }
        mbar_wait(p_full, j & 1);  // P(j) in smem, any lazy rescale of O done
        tc_fence_after();
        const uint32_t vbase = smem_u32(sV + st * ATT_TILE_BYTES);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint64_t vdesc = make_sdesc_sw128(vbase + k * 2048);
          umma_f16_ts(tmem_O, tmem_P + k * 8, vdesc, idesc_o, (j | k) != 0);  // A = P from TMEM; O accumulates over all tiles
        }
        umma_commit(pv_done);
        umma_commit(&kv_empty[st]);
      }
    }
  } else if (warp >= ATT_SOFTMAX_WARP0) {
    const int quarter = warp & 3;
    const int half = (warp - ATT_SOFTMAX_WARP0) >> 2;   // which 64 score columns (= which K-slab of P, which 32 columns of O)
    const int r = quarter * 32 + lane;  // query row inside the tile == TMEM lane
    const uint32_t lane_sel = static_cast<uint32_t>(quarter * 32) << 16;
    const float LOG2E = 1.4426950408889634f;
    float m_used = -INFINITY, l_part = 0.f;  // m_used: the max P / O are currently scaled by; l_part: my half's row sum
    const uint32_t bar_id = 1 + quarter;  // named barrier of the two warps that share this lane quarter
#ifdef PRISMA_ATTN_PROFILE
    long long t_wait = 0, t_p1 = 0, t_p2 = 0, t_ld = 0, t_max = 0, t_pvw = 0, t_exp = 0, t_tot = clock64();
    const bool prof = args.dbg != nullptr && threadIdx.x == 64 && blockIdx.x == 1 && blockIdx.y == 1;
#define ATT_CLK(x) long long x = clock64()
#else
#define ATT_CLK(x)
#endif

    for (int j = 0; j < n_kv; ++j) {
      const int valid = min(64, max(0, T - j * ATT_BKV - half * 64));  // valid columns of my half (0..64)
      ATT_CLK(c0_);
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      ATT_CLK(c1_);
      // ---- my half of the S row (64 fp32) into registers with one wait; then the TMEM buffer is free for S(j+1)
      uint32_t v[64];
      tmem_ld32(tmem_S + lane_sel + half * 64, v);
      tmem_ld32(tmem_S + lane_sel + half * 64 + 32, v + 32);
      tmem_ld_wait();
      ATT_CLK(c1a_);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_free);
      float mx = -INFINITY;
      if (valid == 64) {  // four FMNMX3 chains
        float m0 = fmax3(__uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2]));
        float m1 = fmax3(__uint_as_float(v[3]), __uint_as_float(v[4]), __uint_as_float(v[5]));
        float m2 = fmax3(__uint_as_float(v[6]), __uint_as_float(v[7]), __uint_as_float(v[8]));
        float m3 = fmax3(__uint_as_float(v[9]), __uint_as_float(v[10]), __uint_as_float(v[11]));
#pragma unroll
        for (int i = 12; i < 60; i += 8) {
          m0 = fmax3(m0, __uint_as_float(v[i]), __uint_as_float(v[i + 1]));
          m1 = fmax3(m1, __uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
          m2 = fmax3(m2, __uint_as_float(v[i + 4]), __uint_as_float(v[i + 5]));
          m3 = fmax3(m3, __uint_as_float(v[i + 6]), __uint_as_float(v[i + 7]));
        }
        m0 = fmax3(m0, __uint_as_float(v[60]), __uint_as_float(v[61]));
        m1 = fmax3(m1, __uint_as_float(v[62]), __uint_as_float(v[63]));
        mx = fmaxf(fmax3(m0, m1, m2), m3);
      } else {
#pragma unroll
        for (int i = 0; i < 64; ++i) if (i < valid) mx = fmaxf(mx, __uint_as_float(v[i]));
      }
      // ---- row max across the two halves: exchanged as fp16 rounded UP (both threads then use the identical value,
      // which is all the online softmax needs; >= the true max, so p <= 1 up to the lazy-rescale slack)
      ATT_CLK(c1b_);
      // Buffers alternate with the tile parity, so one barrier per tile suffices: a buffer is rewritten two tiles later,
      // after the barrier of the tile in between, which the partner only reaches once it has read this one.
      __half* xm = s_xmax + (j & 1) * 256;
      xm[half * 128 + r] = __float2half_ru(fmaxf(mx, -60000.f));
      asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
      mx = fmaxf(__half2float(xm[r]), __half2float(xm[128 + r]));
      // ---- lazy rescale (identical decision in both warps of the quarter; each owns 32 columns of O)
      if (j == 0) {
        m_used = mx;
      } else {
        const bool need = (mx - m_used) * LOG2E > 8.0f;
        if (__any_sync(0xffffffffu, need)) {
          const float m_new = need ? mx : m_used;
          const float alpha = ex2_approx((m_used - m_new) * LOG2E);  // 1 for rows that keep their max
          mbar_wait(pv_done, (j - 1) & 1);                            // no PV may be in flight on O
          tc_fence_after();
#pragma unroll 1
          for (int hh = 0; hh < 2; ++hh) {  // 16 columns at a time: keeps the S row in registers
            uint32_t o[16];
            tmem_ld16(tmem_O + lane_sel + half * 32 + hh * 16, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st16(tmem_O + lane_sel + half * 32 + hh * 16, o);
          }
          tmem_st_wait();
          l_part *= alpha;
          m_used = m_new;
        }
      }
      const float mscaled = m_used * LOG2E;
      ATT_CLK(c2_);
      // ---- p = exp(s - m_used), partial row sum, P -> fp16 -> swizzled smem (K-slab `half` of the PV A operand)
      float sum0 = 0.f, sum1 = 0.f, sum2 = 0.f, sum3 = 0.f;
      // two chunks of 32 probabilities -> 16 packed fp16 pairs -> 16 TMEM columns each (keeps the live set small)
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        uint32_t pk[16];
        if (valid == 64) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            float e[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) e[i] = ex2_approx(fmaf(__uint_as_float(v[hh * 32 + c * 8 + i]), LOG2E, -mscaled));
            sum0 += e[0] + e[4]; sum1 += e[1] + e[5]; sum2 += e[2] + e[6]; sum3 += e[3] + e[7];
            pk[c * 4 + 0] = pack_half2(e[0], e[1]); pk[c * 4 + 1] = pack_half2(e[2], e[3]);
            pk[c * 4 + 2] = pack_half2(e[4], e[5]); pk[c * 4 + 3] = pack_half2(e[6], e[7]);
          }
        } else {
#pragma unroll
          for (int c = 0; c < 4; ++c) {  // last KV tile only; fully unrolled so v[] stays in registers
            float e[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float x = ex2_approx(fmaf(__uint_as_float(v[hh * 32 + c * 8 + i]), LOG2E, -mscaled));
              e[i] = (hh * 32 + c * 8 + i < valid) ? x : 0.f;
            }
            sum0 += e[0] + e[4]; sum1 += e[1] + e[5]; sum2 += e[2] + e[6]; sum3 += e[3] + e[7];
            pk[c * 4 + 0] = pack_half2(e[0], e[1]); pk[c * 4 + 1] = pack_half2(e[2], e[3]);
            pk[c * 4 + 2] = pack_half2(e[4], e[5]); pk[c * 4 + 3] = pack_half2(e[6], e[7]);
          }
        }
        // P(j-1) must have been consumed by PV(j-1) before the buffer is rewritten (issued a whole softmax ago)
        if (hh == 0 && j > 0) { mbar_wait(pv_done, (j - 1) & 1); tc_fence_after(); }
        tmem_st16(tmem_P + lane_sel + half * 32 + hh * 16, pk);
      }
      l_part += (sum0 + sum1) + (sum2 + sum3);
      ATT_CLK(c2b_);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
#ifdef PRISMA_ATTN_PROFILE
      long long c3_ = clock64();
      t_wait += c1_ - c0_; t_p1 += c2_ - c1_; t_p2 += c3_ - c2_;
      t_ld += c1a_ - c1_; t_max += c1b_ - c1a_; t_exp += c2b_ - c2_;
#endif
    }
#ifdef PRISMA_ATTN_PROFILE
    if (prof) { args.dbg[0] = t_wait; args.dbg[1] = t_p1; args.dbg[2] = t_p2; args.dbg[3] = clock64() - t_tot; args.dbg[4] = n_kv;
                args.dbg[5] = t_ld; args.dbg[6] = t_max; args.dbg[7] = t_pvw; args.dbg[8] = t_exp; }
#endif
    // ---- O is complete once the last PV retires; the K stages are then idle and carry the row-sum exchange
    mbar_wait(pv_done, (n_kv - 1) & 1);
    tc_fence_after();
    float* s_l = reinterpret_cast<float*>(sK);  // [2 parities][2 halves][128 rows]; the K stages are idle by now
    s_l[half * 128 + r] = l_part;
    uint32_t ov[32];
    tmem_ld32(tmem_O + lane_sel + half * 32, ov);
    tmem_ld_wait();
    asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
    const float inv = 1.0f / (s_l[r] + s_l[128 + r]);
    // ---- normalise and store my 32 columns: out[row][head*64 + half*32 + d]
    const int q = q0 + r;
    if (q < T) {
      __half* dst = args.out + (size_t)(row_base + q) * args.out_ld + head * ATT_HD + half * 32;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint4 o;
        o.x = pack_half2(__uint_as_float(ov[g * 8 + 0]) * inv, __uint_as_float(ov[g * 8 + 1]) * inv);
        o.y = pack_half2(__uint_as_float(ov[g * 8 + 2]) * inv, __uint_as_float(ov[g * 8 + 3]) * inv);
        o.z = pack_half2(__uint_as_float(ov[g * 8 + 4]) * inv, __uint_as_float(ov[g * 8 + 5]) * inv);
        o.w = pack_half2(__uint_as_float(ov[g * 8 + 6]) * inv, __uint_as_float(ov[g * 8 + 7]) * inv);
        *reinterpret_cast<uint4*>(dst + g * 8) = o;
      }
    }
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

int attention_prepare(AttnLaunch* out, const __half* qkv, __half* o, int batch, int tokens, int heads, int D) {
  PRISMA_CHECK(D == heads * ATT_HD, "attention: head_dim must be 64");
  out->args.tokens = tokens;
  out->args.heads = heads;
  out->args.D = D;
  out->args.batch = batch;
  out->args.out = o;
  out->args.out_ld = D;
  out->args.dbg = nullptr;
  PRISMA_TRY(make_tmap_2d_f16(&out->tm, qkv, (uint64_t)3 * D, (uint64_t)batch * tokens, (uint64_t)3 * D, 64, 128));
  out->grid = dim3(ceil_div(tokens, ATT_BQ), heads, batch);
  out->flops = 4.0 * batch * heads * (double)tokens * tokens * ATT_HD;
  return 0;
}

int attention_run(const AttnLaunch& a, cudaStream_t s) {
  static bool attr_set = false;
  if (!attr_set) {
    PRISMA_CUDA_OK(cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
    attr_set = true;
  }
  attention_kernel<<<a.grid, ATT_THREADS, ATT_SMEM, s>>>(a.tm, a.args);
  PRISMA_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace prisma
