// prisma_b200 -- SOLOv2 (mask_mmdet band) pointwise / reduction / decode kernels (see solo_kernels.cu).
#pragma once
#include "common.cuh"

namespace prisma {

struct SoloCand {       // one candidate (grid cell, class) above score_thr, solov2_head.py:694-703
  float score;          // cls score, later x maskness
  int flat;             // cell * num_classes + class: the reference's nonzero() order, used as the stable tie-break
  float area;           // sum of (p > mask_thr)
  float stride;
};

int solo_preprocess(const uint8_t* rgb, int H, int W, int nh, int nw, int hp, int wp, float* out_chw, uint8_t* resized_u8,
                    cudaStream_t s);
int maxpool3s2_f16(const __half* in, int H, int W, int C, __half* out, int Ho, int Wo, cudaStream_t s);
int nearest_add_f16(__half* fine, int Hf, int Wf, const __half* coarse, int Hc, int Wc, int C, cudaStream_t s);
int subsample2_f16(const __half* in, int H, int W, int C, __half* out, int Ho, int Wo, cudaStream_t s);
int gn_partial_floats(int HW, int C);
int groupnorm_relu_f16(const float* x, int H, int W, int C, int groups, const float* gamma, const float* beta, float* part,
                       float* stats, __half* out_padded, __half* out_dense, cudaStream_t s);
// bilinear, align_corners=False, padded NHWC fp16 -> padded NHWC fp16 (pitch Cdst >= Csrc); coord: append the
// generate_coordinate channels (x, y in [-1,1] of the SOURCE grid, interpolated like the features) at [Csrc, Csrc+2);
// accumulate: dst += value
int resize_bilinear_f16(const __half* src, int Hs, int Ws, int Csrc, __half* dst, int Hd, int Wd, int Cdst, int coord,
                        int accumulate, cudaStream_t s, int dst_frame_w = 0);  // dst_frame_w: row pitch of dst if wider than Wd
// The five head levels share their tower weights, so they run as ONE stack of F x F frames (F = the largest grid, 40): level
// l occupies the top-left S_l x S_l of frame l, everything else stays zero (which is exactly the convs' zero padding).
struct GridSizes { int s[8]; };
// GroupNorm(32)+ReLU of a stacked conv output x: dense fp32 [B][F*F][C]; statistics over the valid S_l x S_l pixels of each
// frame; writes the valid pixels of the stacked zero-bordered fp16 map [B][(F+2)^2][C].  One block per (frame, group).
int groupnorm_relu_grid_f16(const float* x, int B, int F, GridSizes S, int C, int groups, const float* gamma, const float* beta,
                            __half* out_padded, cudaStream_t s);
// decode (solov2_head.py:582-766, matrix_nms.py)
int solo_candidates(const float* cls_logits, int S, int cell0, int num_classes, float score_thr, float stride, SoloCand* cand,
                    int* count, int cap, cudaStream_t s, int frame = 0);  // frame: row pitch (in cells) of the logits, 0 = S
int solo_sort_candidates(const SoloCand* cand, int* count, int cap, SoloCand* sorted, cudaStream_t s);  // nonzero() order
int solo_gather_kernels(const SoloCand* cand, const int* count, int cap, const float* const* lvl_kernels,
                        const int* lvl_cell0, int levels, int num_classes, int C, __half* out, cudaStream_t s,
                        const int* lvl_S = nullptr, int frame = 0);  // lvl_S/frame: kernels stored in F x F frames
int solo_mask_stats(const __half* masks, int HW, float mask_thr, SoloCand* cand, const int* count, int cap,
                    cudaStream_t s);
int solo_rank(const SoloCand* cand, const int* count, int cap, int nms_pre, int* top, int* n_top,
              cudaStream_t s);
int solo_binarize(const __half* masks, int HW, float mask_thr, const int* top, const int* n_top, int nms_pre, __half* bin,
                  cudaStream_t s);
int solo_matrix_nms(const float* inter, int ld, const SoloCand* cand, const int* top, const int* n_top, int num_classes,
                    float sigma, float filter_thr, int max_num, int* keep, float* keep_score, int* keep_label, int* n_keep,
                    cudaStream_t s);
int solo_final_masks(const __half* masks, int fh, int fw, int h, int w, int H, int W, float mask_thr, const int* keep,
                     const float* keep_score, const int* keep_label, const int* n_keep, int max_num, float confidence,
                     uint8_t* inst_masks, uint8_t* union_mask, cudaStream_t s);

// ---- fp32-class head (solo_exact.cu): split maps = zero-bordered NHWC rows of 2 C floats [hi | lo]; fp32 mask predictions
int solo_mask_stats(const float* masks, int HW, float mask_thr, SoloCand* cand, const int* count, int cap, cudaStream_t s);
int solo_binarize(const float* masks, int HW, float mask_thr, const int* top, const int* n_top, int nms_pre, __half* bin,
                  cudaStream_t s);
int solo_final_masks(const float* masks, int fh, int fw, int h, int w, int H, int W, float mask_thr, const int* keep,
                     const float* keep_score, const int* keep_label, const int* n_keep, int max_num, float confidence,
                     uint8_t* inst_masks, uint8_t* union_mask, cudaStream_t s);
int solo_dense_to_split(const float* src, int H, int W, int Csrc, float* dst, int Cdst, cudaStream_t s);
int solo_f16map_to_split(const __half* src, int H, int W, int C, float* dst, int Cdst, cudaStream_t s);
int groupnorm_relu_split(const float* x, int H, int W, int C, int groups, const float* gamma, const float* beta, float* part,
                         float* stats, float* out_split, float* out_w3, cudaStream_t s);
int groupnorm_relu_grid_split(const float* x, int B, int F, GridSizes S, int C, int groups, const float* gamma, const float* beta,
                              float* out_split, cudaStream_t s);
int resize_bilinear_split(const float* src, int Hs, int Ws, int Csrc, int Csrc_pitch, float* dst, int Hd, int Wd, int Cdst, int coord,
                          int accumulate, cudaStream_t s, int dst_frame_w = 0);
// fp32-class backbone ("-exact" variants): dense fp32 NHWC maps between the 3xTF32 convs
int solo_im2col_stem_split(const float* x_chw, int H, int W, float* out, cudaStream_t s);  // -> [H/2 * W/2][hi(192) | lo(192)]
int maxpool3s2_dense(const float* in, int H, int W, int C, float* out, float* out_split, int Ho, int Wo, cudaStream_t s);
int nearest_add_dense(float* fine, int Hf, int Wf, const float* coarse, int Hc, int Wc, int C, cudaStream_t s);
int subsample2_dense(const float* in, int H, int W, int C, float* out, int Ho, int Wo, cudaStream_t s);
int solo_gather_kernels_split(const SoloCand* cand, const int* count, int cap, const float* const* lvl_kernels,
                              const int* lvl_cell0, int levels, int num_classes, int C, float* out, cudaStream_t s,
                              const int* lvl_S = nullptr, int frame = 0);

}  // namespace prisma

namespace prisma {
// getSDF (bands/mask_mmdet.py:64-69): exact Euclidean signed distance of the union mask (snowy.generate_sdf = udf(mask) -
// udf(~mask), Felzenszwalb exact EDT), remapped to the green channel: 255 * (1 - clip(((sdf + 127)/255 - 0.25) * 2, 0, 1)).
// mask: H x W u8 (non-zero = inside), d_scratch: 2 * H * W ints, green: H x W u8.  All on the device.
int mask_sdf_green(const uint8_t* mask, int H, int W, int* d_scratch, uint8_t* green, cudaStream_t s);
}  // namespace prisma
