// Internal launch interface between api.cu (C ABI, argument checking, dispatch) and the kernel files.
#pragma once
#include "ofa_common.cuh"

namespace ofa {

// ---- simt_kernels.cu ----------------------------------------------------------------------------
int launch_active_filter(const float* w7, int kmax, const float* m75, const float* m53, int transform_on,
                         int ks, int C, float* out, cudaStream_t st);
int launch_active_filter_chunked(const float* w7, int kmax, const float* m75, const float* m53, int transform_on,
                                 int ks, int C, int flip, float* out, cudaStream_t st);   // [C / 64][ks * ks][64]
int launch_dw_simt(const TV& x, const TV& y, const float* w7, int kmax, const float* m75,
                   const float* m53, int transform_on, int ks, int flip, const Epi& epi, cudaStream_t st);
int launch_conv_simt(const TV& x, const TV& y, const float* w, long long w_so, long long w_si,
                     long long w_sh, long long w_sw, int cin, int cout, int ks, int flip, int store,
                     const Epi& epi, cudaStream_t st);
int launch_pack_weight(const float* w, long long w_so, long long w_si, long long w_sh, long long w_sw,
                       int cin, int cout, int ks, int cin_pad, int cout_pad, int store, int f16, void* out,
                       cudaStream_t st);
void keep_async_pool_resident();   // raise the default mempool's release threshold once per device (cudaMallocAsync scratch)
int launch_pack_weights_multi(const OfaPackJob* jobs_device, int njobs, cudaStream_t st);
int launch_pack_weights4(const OfaPackJob* jobs_host, int njobs, cudaStream_t st);   // <= 4 jobs by value, one launch
int launch_affine_act(const TV& x, const TV& y, const Epi& epi, int store, cudaStream_t st);
int launch_bn_stats(const TV& x, float* mean, float* var, cudaStream_t st);
int launch_bn_stats_update(const TV& x, float* mean, float* var, float* rm, float* rv, float momentum,
                           long long* num_batches_tracked, cudaStream_t st);
int launch_bn_update_running(const float* mean, const float* var, long long count, float* rm, float* rv,
                             float momentum, int C, long long* num_batches_tracked, cudaStream_t st);
int launch_bn_bwd_reduce(const TV& x, const TV& dy, const float* gamma, const float* beta,
                         const float* mean, const float* var, float eps, int act, float* sum_dz,
                         float* sum_dz_xhat, cudaStream_t st);
int launch_bn_bwd_apply(const TV& x, const TV& dy, const TV& dx, const float* gamma, const float* beta,
                        const float* mean, const float* var, float eps, int act, int training,
                        const float* sum_dz, const float* sum_dz_xhat, cudaStream_t st);
int launch_dw_bwd_filter(const TV& x, const TV& dy, int ks, float* dw, cudaStream_t st);
int launch_active_filter_bwd(const float* w7, int kmax, const float* m75, const float* m53,
                             int transform_on, int ks, int C, const float* dwa, float* dw7, float* dm75,
                             float* dm53, cudaStream_t st);
int launch_conv_bwd_weight(const TV& x, const TV& dy, float* dw, long long w_so, long long w_si,
                           long long w_sh, long long w_sw, int cin, int cout, int ks, cudaStream_t st);

int launch_adam_step(const OfaAdamTensor* table, const int* chunks, int n_tensors, int n_chunks,
                     const float* const* grads, int* steps, float lr, float beta1, float beta2, float eps,
                     cudaStream_t st);
int launch_psnr_y_sse(const TV& a, const TV& b, long long* sse, cudaStream_t st);

// ---- dw_fast.cu : NHWC bf16 depthwise, smem halo tiles --------------------------------------------
bool dw_fast_supported(const OfaTensor4* x, const OfaTensor4* y, int ks, const OfaEpilogue* epi);
int launch_dw_fast(const OfaTensor4* x, const OfaTensor4* y, const float* w7, int kmax, const float* m75,
                   const float* m53, int transform_on, int ks, int flip, const OfaEpilogue* epi, cudaStream_t st);

// ---- conv_tc.cu : NHWC bf16 implicit GEMM on tcgen05 / TMEM / TMA --------------------------------
bool conv_tc_supported(const OfaConvArgs* a);
int launch_conv_tc(const OfaConvArgs* a, cudaStream_t st);

// ---- conv_thin.cu : the thin ends (64 -> 3 on tcgen05 with kx folded into N; 3 -> 64 on CUDA cores) ---
bool conv_out_rows_supported(const OfaConvArgs* a);
int launch_conv_out_rows(const OfaConvArgs* a, cudaStream_t st);
bool conv_stem_supported(const OfaConvArgs* a);
// conv_stem_tc.cu : the same stems as an im2col GEMM on tcgen05 (NHWC 16-bit output, 64 channels)
bool conv_stem_tc_supported(const OfaConvArgs* a);
int launch_conv_stem_tc(const OfaConvArgs* a, cudaStream_t st);
int launch_conv_stem(const OfaConvArgs* a, cudaStream_t st);

// ---- wgrad_tc.cu : dense-conv weight gradient on tcgen05 (pixels are K; both operands MN-major) ---------
bool wgrad_tc_supported(const OfaTensor4* x, const OfaTensor4* dy, int cin, int cout, int ks);
int launch_wgrad_tc(const OfaTensor4* x, const OfaTensor4* dy, float* dw, long long w_so, long long w_si,
                    long long w_sh, long long w_sw, int cin, int cout, int ks, cudaStream_t st);

// ---- mbconv_planar.cu : MBConv block on channel-planar 16-bit intermediates (tcgen05) ---------------
bool mbconv_planar_supported(const OfaMBConvArgs* a);
bool mbconv_planar_preferred(const OfaMBConvArgs* a);   // tile-fill heuristic of OFA_IMPL_AUTO
int launch_pack_block_weights(const float* w_exp, long long e_so, long long e_si, const float* w_proj,
                              long long p_so, long long p_si, int cin, int mid, int cout, int mid_pad, int trunk_f16,
                              int f16, void* wexp_p, void* wproj_p, cudaStream_t st);
int launch_expand_planar(const void* x, void* y, const void* wexp_p, int N, int HW, int mid, int trunk_f16, int f16,
                         const OfaBn* bn, int act, cudaStream_t st);
int launch_dw_planar(const void* x, void* y, int N, int C, int H, int W, const float* w7, int kmax, const float* m75,
                     const float* m53, int transform_on, int ks, int f16, const OfaBn* bn, int act, cudaStream_t st);
int launch_project_planar(const void* x, const void* res, void* y, const void* wproj_p, int N, int HW, int mid,
                          int trunk_f16, int f16, const OfaBn* bn, cudaStream_t st);

// ---- mbconv_band.cu : the same three stages as ONE launch around an L2-resident region ring ---------------
bool mbconv_band_supported(const OfaMBConvArgs* a);
bool mbconv_band_preferred(const OfaMBConvArgs* a);
long long mbconv_band_workspace_bytes(int W, int mid, int n_regions_max);
int launch_mbconv_band(const OfaMBConvArgs* a, const void* wexp_p, const void* wproj_p, int f16, void* ws,
                       long long ws_bytes, cudaStream_t st);

// ---- data_prep.cu : SR data preparation (Pillow-exact bicubic resampling, crop / flip / rotate, ToTensor) ----
int resample_ksize(int in_size, int out_size);
int resample_build_table(int in_size, int out_size, int32_t* bounds_host, int32_t* kk_host);
int launch_bicubic_resize_u8(const uint8_t* src, int N, int H, int W, int oh, int ow, const int32_t* bounds_h,
                             const int32_t* kk_h, int ksize_h, const int32_t* bounds_v, const int32_t* kk_v,
                             int ksize_v, uint8_t* tmp, uint8_t* out_u8, float* out_f32, cudaStream_t st);
int launch_sr_augment_u8(const uint8_t* src, long long sample_stride, int N, int W, const int32_t* params, int S,
                         uint8_t* out_u8, float* out_f32, cudaStream_t st);

// ---- mbv3_ops.cu : squeeze-and-excite, sliced linear, strided depthwise (exact fp32 CUDA-core kernels) ----
int launch_linear_fwd(const float* x, long long ldx, const float* w, long long ldw, const float* bias, int N, int in,
                      int out, int act, float* y, long long ldy, cudaStream_t st);
int launch_linear_bwd_data(const float* dz, long long lddz, const float* w, long long ldw, int N, int in, int out,
                           float* dx, long long lddx, cudaStream_t st);
int launch_linear_bwd_weight(const float* dz, long long lddz, const float* x, long long ldx, int N, int in, int out,
                             float* dw, long long lddw, float* db, cudaStream_t st);
int launch_act_bwd_from_output(const float* dy, const float* y, int act, float* dz, long long total, cudaStream_t st);
int launch_plane_reduce(const TV& x, const TV* dy, float scale, float* out, cudaStream_t st);
int launch_channel_scale(const TV& x, const TV& y, const float* s, const float* add, cudaStream_t st);
int launch_dw_strided_fwd(const TV& x, const TV& y, const float* f, int ks, int stride, const Epi& epi, cudaStream_t st);
int launch_dw_strided_bwd_data(const TV& dy, const TV& dx, const float* f, int ks, int stride, cudaStream_t st);
int launch_dw_strided_bwd_filter(const TV& x, const TV& dy, int ks, int stride, float* df, cudaStream_t st);

}  // namespace ofa
