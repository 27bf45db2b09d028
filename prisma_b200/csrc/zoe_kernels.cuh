// prisma_b200 -- ZoeDepth metric head (depth_anything --metric) pointwise kernels (see zoe_kernels.cu).
#pragma once
#include "common.cuh"

namespace prisma {

int zoe_preprocess(const uint8_t* rgb, int H, int W, float* out_chw, int h, int w, cudaStream_t s);
// out[P][C] fp16 = a[P][C] (f32) + bilinear(align_corners=True) of prev [Hp*Wp][C] (f32) at the (H, W) grid
int zoe_embed_add(const float* a, int H, int W, int C, const float* prev, int Hp, int Wp, __half* out, cudaStream_t s);
// AttractorLayerUnnormed (inv attractor, mean): b_new = b + mean_k dx/(1 + 300 dx^2), dx = A_k - b, b = bilinear_ac(b_prev)
int zoe_attractor(const float* A, int lda, int n_attr, const float* b_prev, int Hp, int Wp, int H, int W, int bins, float* b_out,
                  cudaStream_t s);
// concat operand of ConditionalLogBinomial.mlp: [act 32 | rel depth | bilinear_ac(embedding 128) | zeros] fp16, pitch 192
int zoe_concat(const __half* act32, const float* rel, const float* emb, int He, int We, int H, int W, __half* out,
               cudaStream_t s);
// log-binomial softmax over the bins x interpolated bin centres -> metric depth
int zoe_final(const float* pt, const float* centers, int Hc, int Wc, int H, int W, int bins, float min_temp, float max_temp,
              float* out, cudaStream_t s);
// PIL Image.resize((W, H)) of a mode "F" image, default BICUBIC (Pillow Resample.c, a = -0.5): horizontal then vertical
int pil_bicubic_resize_f32(const float* in, int ih, int iw, float* tmp, float* out, int oh, int ow, cudaStream_t s);

}  // namespace prisma
