// Dense convolution (1x1 channel-sliced point convs and k x k ConvLayers) as an implicit GEMM on the
// 5th-gen tensor cores:  D[pixel, cout] = sum_{tap, cin} X[pixel + tap, cin] * W[tap][cout][cin].
//
//   * activations NHWC bf16; a CTA tile is TH x TW = 8 x 16 = 128 pixels (the UMMA M dimension);
//   * per K block (one filter tap x 64 input channels) TMA loads the shifted 8x16x64 activation box
//     (zero fill outside the image = the conv padding) and the [BN x 64] weight box, both with the
//     128-byte swizzle, into a STAGES-deep mbarrier ring;
//   * one elected thread issues tcgen05.mma (M=128, N=BN, K=16) x 4 per K block, accumulating in TMEM;
//     two accumulator stages let the epilogue of tile i overlap the main loop of tile i+1;
//   * 4 epilogue warps read TMEM (tcgen05.ld), apply the folded BatchNorm scale/shift, the
//     activation and the residual / long skip, and store plain, PixelShuffle(2) or
//     PixelUnshuffle(2) layouts (bf16 NHWC vectorised, or any strided fp32 / bf16 view);
//   * persistent: grid = #SMs, tiles strided over CTAs.
#include "ofa_common.cuh"
#include "kernels.h"
#include "sm100_ptx.cuh"

#include <cudaTypedefs.h>
#include <string.h>

namespace ofa {

// --------------------------------------------------------------------------------------------------
// tensor-map encoding via the runtime's driver entry point lookup
// --------------------------------------------------------------------------------------------------
int encode_tmap(CUtensorMap* out, CUtensorMapDataType dt, uint32_t rank, void* base, const uint64_t* dims,
                const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swz) {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !p)
      return fail(OFA_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }
  uint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(out, dt, rank, base, dims, strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(OFA_ERR_CUDA, "cuTensorMapEncodeTiled failed (CUresult %d)", (int)r);
  return OFA_OK;
}

namespace {

constexpr int TH = 8, TW = 16;         // spatial tile = 128 pixels = UMMA M
constexpr int BK = 64;                 // channels per K block (128 bytes of bf16 = one swizzle row)
constexpr int A_BYTES = TH * TW * BK * 2;  // 16 KiB
constexpr int NUM_THREADS = 192;       // warp 0 TMA, warp 1 MMA, warps 2..5 epilogue
constexpr int MAX_COUT = 512;

struct ConvTcParams {
  int N, H, W;          // input (= conv resolution) extents
  int cin, cout, ks;
  int BN;               // output channels per CTA tile (multiple of 16, <= 256)
  int n_splits;         // cout_pad / BN
  int tiles_h, tiles_w;
  int stages;
  int tmem_cols;        // power of two >= 2*BN
  int store;
  int act;
  int vec_store;        // 1: y is channel-innermost bf16 and 16-channel groups are 32-byte aligned
  const float* gamma; const float* beta; const float* mean; const float* var; float eps;
  TV y;
  TV res;
};

// packed-row index o' of the weight -> original conv output channel o (see pack order below)
__device__ __forceinline__ int packed_to_conv_channel(int op, int cout, int store) {
  if (store == OFA_STORE_PIXELSHUFFLE2) {
    int q = cout >> 2;
    return 4 * (op % q) + op / q;
  }
  return op;
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
               const ConvTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [stages x A][stages x B][scale MAX_COUT][shift MAX_COUT][barriers][tmem ptr]
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int B_BYTES = p.BN * BK * 2;
  uint8_t* sA = smem;
  uint8_t* sB = smem + (size_t)p.stages * A_BYTES;
  float* s_scale = reinterpret_cast<float*>(sB + (size_t)p.stages * B_BYTES);
  float* s_shift = s_scale + MAX_COUT;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_shift + MAX_COUT);
  uint64_t* empty_bar = full_bar + 8;
  uint64_t* tfull_bar = empty_bar + 8;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // folded BN for all (packed-order) output channels of this layer
  for (int op = threadIdx.x; op < p.n_splits * p.BN; op += NUM_THREADS) {
    float sc = 0.f, sh = 0.f;
    if (op < p.cout) {
      int o = packed_to_conv_channel(op, p.cout, p.store);
      float g = p.gamma ? p.gamma[o] : 1.f;
      float b = p.beta ? p.beta[o] : 0.f;
      float m = p.mean ? p.mean[o] : 0.f;
      float rstd = p.var ? rsqrtf(p.var[o] + p.eps) : 1.f;
      sc = g * rstd;
      sh = b - m * sc;
    }
    s_scale[op] = sc;
    s_shift[op] = sh;
  }
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_x);
    ptx::prefetch_tmap(&tmap_w);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&tfull_bar[a], 1);
      ptx::mbar_init(&tempty_bar[a], 4);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_ptr, (uint32_t)p.tmem_cols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int tiles_per_img = p.tiles_h * p.tiles_w;
  const int num_tiles = p.N * tiles_per_img * p.n_splits;
  const int taps = p.ks * p.ks;
  const int kcs = p.cin / BK;
  const int num_k = taps * kcs;
  const int R = p.ks / 2;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        int split = t % p.n_splits;
        int sp = t / p.n_splits;
        int n = sp / tiles_per_img;
        int r = sp - n * tiles_per_img;
        int h0 = (r / p.tiles_w) * TH, w0 = (r % p.tiles_w) * TW;
        for (int kb = 0; kb < num_k; ++kb) {
          int tap = kb / kcs, kc = kb - tap * kcs;
          int ky = tap / p.ks, kx = tap - ky * p.ks;
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          ptx::mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)(A_BYTES + B_BYTES));
          ptx::tma_load_4d(sA + (size_t)stage * A_BYTES, &tmap_x, &full_bar[stage], kc * BK, w0 + kx - R,
                           h0 + ky - R, n);
          ptx::tma_load_3d(sB + (size_t)stage * B_BYTES, &tmap_w, &full_bar[stage], kc * BK, split * p.BN, tap);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = ptx::umma_idesc_bf16(128, p.BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        ptx::mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.BN);
        for (int kb = 0; kb < num_k; ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after();
          const uint64_t da = ptx::umma_desc_sw128(ptx::smem_u32(sA + (size_t)stage * A_BYTES), 1024);
          const uint64_t db = ptx::umma_desc_sw128(ptx::smem_u32(sB + (size_t)stage * B_BYTES), 1024);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // advance 16 bf16 = 32 bytes along K inside the 128-byte swizzle row
            ptx::umma_bf16(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) ? 1u : 0u);
          }
          ptx::umma_commit(&empty_bar[stage]);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        ptx::umma_commit(&tfull_bar[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int quarter = warp & 3;              // TMEM lane quarter this warp may read
    const int row = quarter * 32 + lane;       // pixel index inside the tile
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      int split = t % p.n_splits;
      int sp = t / p.n_splits;
      int n = sp / tiles_per_img;
      int r = sp - n * tiles_per_img;
      int h = (r / p.tiles_w) * TH + row / TW, w = (r % p.tiles_w) * TW + row % TW;
      const bool pix_ok = h < p.H && w < p.W;
      ptx::mbar_wait(&tfull_bar[acc], acc_phase);
      ptx::tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * p.BN);
      for (int j = 0; j < p.BN; j += 16) {
        uint32_t v[16];
        ptx::tmem_ld16(t_addr + (uint32_t)j, v);
        ptx::tmem_ld_wait();
        const int op0 = split * p.BN + j;  // packed-order channel of v[0]
        if (pix_ok && op0 < p.cout) {
          float f[16];
#pragma unroll
          for (int i = 0; i < 16; ++i)
            f[i] = apply_act(fmaf(__uint_as_float(v[i]), s_scale[op0 + i], s_shift[op0 + i]), p.act);
          // packed channel op -> stored (channel, h, w)
          int oc0, oh, ow;
          if (p.store == OFA_STORE_PIXELSHUFFLE2) {
            int q = p.cout >> 2;
            int s = op0 / q;  // sub-pixel; a 16-group never straddles sub-pixels (q % 16 == 0)
            oc0 = op0 - s * q; oh = 2 * h + (s >> 1); ow = 2 * w + (s & 1);
          } else if (p.store == OFA_STORE_PIXELUNSHUFFLE2) {
            oc0 = 4 * op0 + 2 * (h & 1) + (w & 1); oh = h >> 1; ow = w >> 1;
          } else {
            oc0 = op0; oh = h; ow = w;
          }
          const int cstep = (p.store == OFA_STORE_PIXELUNSHUFFLE2) ? 4 : 1;
          const bool full16 = op0 + 16 <= p.cout;
          if (p.vec_store && full16 && cstep == 1) {
            const long long o = p.y.off(n, oc0, oh, ow);
            if (p.res.ptr) {
              const uint4* rp = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.res.ptr) +
                                                               p.res.off(n, oc0, oh, ow));
              uint4 r0 = rp[0], r1 = rp[1];
              const uint32_t rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                f[2 * i] += __uint_as_float(rr[i] << 16);
                f[2 * i + 1] += __uint_as_float(rr[i] & 0xffff0000u);
              }
            }
            uint32_t pk[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              __nv_bfloat162 b2 = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
              pk[i] = *reinterpret_cast<uint32_t*>(&b2);
            }
            uint4* yp = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.y.ptr) + o);
            yp[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            yp[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              if (op0 + i < p.cout) {
                int oc = oc0 + i * cstep;
                float val = f[i];
                if (p.res.ptr) val += p.res.ld(p.res.off(n, oc, oh, ow));
                p.y.st(p.y.off(n, oc, oh, ow), val);
              }
            }
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tempty_bar[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

int pick_bn(int cout) {
  int cp = (cout + 15) / 16 * 16;
  if (cp <= 256) return cp;
  // split N over CTAs: largest divisor-friendly tile
  if (cp % 192 == 0) return 192;
  if (cp % 128 == 0) return 128;
  if (cp % 256 == 0) return 256;
  return 0;
}

}  // namespace

bool conv_tc_supported(const OfaConvArgs* a) {
  if (!a->w_bf16) return false;
  if (a->flip) return false;
  if (a->x.dtype != OFA_BF16 || !is_nhwc_dense(&a->x)) return false;
  if ((reinterpret_cast<uintptr_t>(a->x.ptr) & 15) || (reinterpret_cast<uintptr_t>(a->w_bf16) & 15)) return false;
  if (a->cin % BK != 0 || a->cin_pad != a->cin) return false;
  if (a->ks > 7) return false;
  int bn = pick_bn(a->cout);
  if (bn == 0) return false;
  int cp = (a->cout + 15) / 16 * 16;
  if (a->cout_pad < cp || a->cout_pad % bn != 0 || a->cout_pad > MAX_COUT) return false;
  if (a->store == OFA_STORE_PIXELSHUFFLE2 && (a->cout % 64 != 0)) return false;
  if (a->x.n <= 0 || a->x.h <= 0 || a->x.w <= 0) return false;
  return true;
}

int launch_conv_tc(const OfaConvArgs* a, cudaStream_t st) {
  ConvTcParams p;
  memset(&p, 0, sizeof(p));
  p.N = a->x.n; p.H = a->x.h; p.W = a->x.w;
  p.cin = a->cin; p.cout = a->cout; p.ks = a->ks;
  p.BN = pick_bn(a->cout);
  p.n_splits = a->cout_pad / p.BN;
  p.tiles_h = (p.H + TH - 1) / TH;
  p.tiles_w = (p.W + TW - 1) / TW;
  p.store = a->store;
  p.act = a->epi.act;
  p.gamma = a->epi.gamma; p.beta = a->epi.beta; p.mean = a->epi.mean; p.var = a->epi.var; p.eps = a->epi.eps;
  p.y = make_tv(&a->y);
  p.res = a->epi.residual ? make_tv(a->epi.residual) : null_tv();
  int cols = 32;
  while (cols < 2 * p.BN) cols <<= 1;
  p.tmem_cols = cols;
  const int b_bytes = p.BN * BK * 2;
  const int fixed = 1024 /*align slack*/ + 2 * MAX_COUT * 4 + 256;
  int stages = (227 * 1024 - fixed) / (A_BYTES + b_bytes);
  if (stages > 8) stages = 8;
  if (stages < 2) return fail(OFA_ERR_UNSUPPORTED, "conv_tc: tile does not fit shared memory");
  p.stages = stages;
  const size_t smem = (size_t)fixed + (size_t)stages * (A_BYTES + b_bytes);

  // vectorised bf16 store: channel-innermost y (and residual), 16-channel groups 32-byte aligned
  bool vec = a->y.dtype == OFA_BF16 && a->y.sc == 1 && (reinterpret_cast<uintptr_t>(a->y.ptr) & 15) == 0 &&
             a->y.sw % 8 == 0 && a->y.sh % 8 == 0 && a->y.sn % 8 == 0 && a->store != OFA_STORE_PIXELUNSHUFFLE2;
  if (a->epi.residual) {
    const OfaTensor4* r = a->epi.residual;
    vec = vec && r->dtype == OFA_BF16 && r->sc == 1 && (reinterpret_cast<uintptr_t>(r->ptr) & 15) == 0 &&
          r->sw % 8 == 0 && r->sh % 8 == 0 && r->sn % 8 == 0;
  }
  p.vec_store = vec ? 1 : 0;

  CUtensorMap tx, tw;
  {
    uint64_t dims[4] = {(uint64_t)p.cin, (uint64_t)p.W, (uint64_t)p.H, (uint64_t)p.N};
    uint64_t strides[3] = {(uint64_t)p.cin * 2, (uint64_t)p.W * p.cin * 2, (uint64_t)p.H * p.W * p.cin * 2};
    uint32_t box[4] = {BK, TW, TH, 1};
    int rc = encode_tmap(&tx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, a->x.ptr, dims, strides, box,
                         CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)a->cin_pad, (uint64_t)a->cout_pad, (uint64_t)(a->ks * a->ks)};
    uint64_t strides[2] = {(uint64_t)a->cin_pad * 2, (uint64_t)a->cin_pad * a->cout_pad * 2};
    uint32_t box[3] = {BK, (uint32_t)p.BN, 1};
    int rc = encode_tmap(&tw, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(a->w_bf16), dims, strides, box,
                         CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  static thread_local size_t smem_set = 0;
  if (smem > smem_set) {
    OFA_CUDA(cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    smem_set = 227 * 1024;
  }
  const int num_tiles = p.N * p.tiles_h * p.tiles_w * p.n_splits;
  int grid = sm_count();
  if (grid > num_tiles) grid = num_tiles;
  conv_tc_kernel<<<grid, NUM_THREADS, smem, st>>>(tx, tw, p);
  return check_launch("conv_tc_kernel");
}

}  // namespace ofa
