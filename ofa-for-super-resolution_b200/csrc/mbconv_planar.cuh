// Geometry constants and small device helpers shared by the planar MBConv kernels (mbconv_planar.cu: one kernel per
// stage; mbconv_band.cu: the three stages as role-specialised CTAs of ONE launch around an L2-resident ring).
#pragma once
#include "ofa_common.cuh"
#include "sm100_ptx.cuh"

namespace ofa {
namespace {

__device__ __forceinline__ void bn_fold(const float* gamma, const float* beta, const float* mean, const float* var,
                                        float eps, int c, float& scale, float& shift) {
  const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
  const float m = mean ? mean[c] : 0.f, rstd = var ? rsqrtf(var[c] + eps) : 1.f;
  scale = g * rstd;
  shift = b - m * scale;
}

// ---- expand ----
constexpr int EX_NPIX = 256;                 // pixels per tile = UMMA N
constexpr int EX_X_STAGES = 3;
constexpr int EX_SBUFS = 2;                  // store staging buffers per epilogue warp (TMA stores in flight)
constexpr int EX_X_BYTES = EX_NPIX * 128;    // 32 KiB
constexpr int EX_EPI_WARPS = 8;
constexpr int EX_THREADS = 64 + 32 * EX_EPI_WARPS;
constexpr int EX_SBUF_BYTES = 32 * 128;      // per-warp store staging: 32 channels x 64 pixels
constexpr int EX_MAX_MT = 3;

// ---- depthwise (banded-Toeplitz tiles) ----
constexpr int DW_TW = 112;                    // valid output columns per tile (224 bytes: TMA-storable)
constexpr int DW_TH = 128;                    // output rows per tile = UMMA M
constexpr int DW_XPAD = 8;
constexpr int DW_CHUNKS = 8;                  // 16-column K chunks per tile row
constexpr int DW_ACC_COLS = 16 * (DW_CHUNKS - 1) + 32;   // 144
constexpr int DW_ACC_STAGES = 3;
constexpr int DW_A_STAGES = 4;
constexpr int DW_ATOM_STRIDE = 17408;         // (128 + 6) * 128 rounded up to 1024
constexpr int DW_A_STRIDE = 2 * DW_ATOM_STRIDE;
constexpr int DW_BFIRST_BYTES = (DW_ACC_COLS / 8) * 256;   // 4608: N = 144 zero-extended matrix
constexpr int DW_B_BYTES = DW_BFIRST_BYTES + 7 * 1024;
constexpr int DW_OUT_BYTES = DW_TH * DW_TW * 2;            // 28672
constexpr int DW_EPI_WARPS = 8;               // 4 lane quarters x 2 column halves
constexpr int DW_THREADS = (4 + DW_EPI_WARPS) * 32;   // TMA, MMA issuer 0, filter builder, 8 epilogue warps, MMA issuer 1
constexpr int DW_ISSUER1_WARP = 3 + DW_EPI_WARPS;     // warp 11

// offset of element (n = accumulator column, k = input column within the chunk) in a K-major, unswizzled
// B tile: 8 x 16-byte core matrices, K-halves 128 bytes apart, 8-column groups 256 bytes apart
__device__ __forceinline__ int dw_b_off(int n, int k) { return (n >> 3) * 256 + (k >> 3) * 128 + (n & 7) * 16 + (k & 7) * 2; }

// ---- project ----
constexpr int PJ_MPIX = 128;                  // pixels per tile = UMMA M
constexpr int PJ_A_STAGES = 9;
constexpr int PJ_A_BYTES = 16384;             // 2 boxes of 64 pixels x 64 channels
constexpr int PJ_MAX_KC = 6;
constexpr int PJ_ACC_STAGES = 4;
constexpr int PJ_R_BYTES = PJ_MPIX * 128;     // residual / output staging tile
constexpr int PJ_THREADS = 6 * 32;            // TMA, MMA, 4 epilogue warps

}  // namespace
}  // namespace ofa
