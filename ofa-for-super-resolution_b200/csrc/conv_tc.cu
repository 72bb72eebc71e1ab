// Dense convolution (1x1 channel-sliced point convs and k x k ConvLayers) as an implicit GEMM on the
// 5th-gen tensor cores:  D[pixel, cout] = sum_{tap, cin} X[pixel + tap, cin] * W[tap][cout][cin].
//
//   * activations NHWC bf16.  A CTA owns a 16 x 16 pixel tile = two UMMA M-tiles of 16 rows x 8 cols
//     (an M-tile's 8-row group = one image row of 8 pixels).
//   * HALO REUSE: per 64-channel chunk the (16+k-1) x (16+k-1) x 64 activation halo arrives with ONE TMA
//     load (128-byte swizzle; out-of-bounds zero fill = the conv's padding) and every filter tap is fed
//     from it: the tap's operand is the same shared-memory tile addressed through a UMMA descriptor
//     whose start address is shifted by (ky * halo_w + kx) rows and whose 8-row-group stride is the
//     halo row pitch.  (The swizzle XOR is a function of the absolute smem address, so TMA's write
//     pattern and tcgen05's read pattern agree for any 128-byte-aligned start: probed in
//     csrc/experiments/umma_probe.cu.)  L2 -> SM activation traffic drops k*k-fold.
//   * weights [tap][cout][cin] bf16: resident in shared memory for the whole persistent CTA when they
//     fit (all 1x1 convs, thin outputs), else streamed per tap through an mbarrier ring shared by both
//     M-tiles.
//   * one elected thread issues tcgen05.mma (M=128, N=BN<=128, K=16); accumulators live in TMEM, two
//     stages, so the epilogue of one (tile, N-split) overlaps the MMAs of the next.  When Cin = 64 the
//     N-splits of a tile are looped INSIDE the CTA, re-using the resident halo.
//   * 8 epilogue warps: tcgen05.ld -> folded BatchNorm scale/shift -> activation -> + residual / long
//     skip -> plain, PixelShuffle(2) or PixelUnshuffle(2) store (vectorised bf16 NHWC, or any view).
//   * persistent: grid = #SMs.
#include "ofa_common.cuh"
#include "kernels.h"
#include "sm100_ptx.cuh"

#include <cudaTypedefs.h>
#include <stdlib.h>
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

constexpr int MT_ROWS = 16, MT_COLS = 8;   // one UMMA M-tile: 16 image rows x 8 pixels
constexpr int MTX = 2;                     // M-tiles side by side -> CTA tile 16 x 16 pixels
constexpr int TILE_H = MT_ROWS, TILE_W = MT_COLS * MTX;
constexpr int BK = 64;                     // channels per K chunk (128 bytes = one swizzle row)
constexpr int NUM_EPI_WARPS = 8;
constexpr int NUM_THREADS = 64 + 32 * NUM_EPI_WARPS;   // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue
constexpr int MAX_COUT = 512;
constexpr int MAX_A_STAGES = 4, MAX_B_STAGES = 8;
constexpr int B_RESIDENT_LIMIT = 64 * 1024;
constexpr int STAGE_BYTES = 32 * 64;         // per epilogue warp: 32 pixel rows x 32 bf16 channels

struct ConvTcParams {
  int N, H, W;
  int cin, cout, ks;
  int kcs;               // cin / 64
  int BN;                // accumulator columns per M-tile (multiple of 16, <= 128)
  int n_splits;          // cout_pad / BN
  int inner_splits;      // splits looped inside a work item (kcs == 1), else 1
  int cout_pad;
  int tiles_h, tiles_w;
  int halo_w, halo_h;
  int a_bytes;           // one halo block (one 64-channel chunk)
  int a_stages, b_stages;
  int b_resident;
  int bres_rows;         // weight rows per TMA box when resident
  int tmem_cols;
  int store, act, vec_store;
  int f16;               // 16-bit storage format of x, the packed weights, y and the residual: 1 = fp16, 0 = bf16
  const float* gamma; const float* beta; const float* mean; const float* var; float eps;
  TV y;
  TV res;
  unsigned long long* trace;   // debug (OFA_CONV_TC_TRACE=1): per-CTA clock stamps of the pipeline stages, else nullptr
};

__device__ __forceinline__ int packed_to_conv_channel(int op, int cout, int store) {
  if (store == OFA_STORE_PIXELSHUFFLE2) {
    int q = cout >> 2;
    return 4 * (op % q) + op / q;
  }
  return op;
}

// f[i] = act(acc[i] * scale[c0 + i] + shift[c0 + i]) for N consecutive channels (c0 % 4 == 0): scale / shift come as
// 16-byte shared-memory broadcasts and the activation is ONE uniform branch per vector, not a switch per element
// (the epilogue is on the critical path of the tensor-bound 64 -> 256 conv: a fourth switch case cost it 19 %)
template <int N>
__device__ __forceinline__ void affine_act_vec(const uint32_t* v, const float* s_scale, const float* s_shift, int c0,
                                               int act, float* f) {
  const float4* sc4 = reinterpret_cast<const float4*>(s_scale + c0);
  const float4* sh4 = reinterpret_cast<const float4*>(s_shift + c0);
#pragma unroll
  for (int q = 0; q < N / 4; ++q) {
    const float4 a = sc4[q], b = sh4[q];
    f[4 * q + 0] = fmaf(__uint_as_float(v[4 * q + 0]), a.x, b.x);
    f[4 * q + 1] = fmaf(__uint_as_float(v[4 * q + 1]), a.y, b.y);
    f[4 * q + 2] = fmaf(__uint_as_float(v[4 * q + 2]), a.z, b.z);
    f[4 * q + 3] = fmaf(__uint_as_float(v[4 * q + 3]), a.w, b.w);
  }
  if (act == OFA_ACT_RELU6) {
#pragma unroll
    for (int i = 0; i < N; ++i) f[i] = fminf(fmaxf(f[i], 0.f), 6.f);
  } else if (act == OFA_ACT_RELU) {
#pragma unroll
    for (int i = 0; i < N; ++i) f[i] = fmaxf(f[i], 0.f);
  } else if (act == OFA_ACT_HSWISH) {
#pragma unroll
    for (int i = 0; i < N; ++i) f[i] = f[i] * fminf(fmaxf(f[i] + 3.f, 0.f), 6.f) * (1.f / 6.f);
  }
}

// folded BN + activation + residual + store of 16 consecutive (packed-order) output channels of one pixel
__device__ __forceinline__ void emit16(const ConvTcParams& p, const float* s_scale, const float* s_shift,
                                       const uint32_t* v, int op0, int n, int h, int w) {
  if (op0 >= p.cout) return;
  float f[16];
  affine_act_vec<16>(v, s_scale, s_shift, op0, p.act, f);
  int oc0, oh, ow;
  if (p.store == OFA_STORE_PIXELSHUFFLE2) {
    int q = p.cout >> 2;
    int sub = op0 / q;  // a 16-group never straddles sub-pixels (q % 16 == 0)
    oc0 = op0 - sub * q; oh = 2 * h + (sub >> 1); ow = 2 * w + (sub & 1);
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
        const float2 r2 = unpack16(rr[i], p.f16);
        f[2 * i] += r2.x;
        f[2 * i + 1] += r2.y;
      }
    }
    uint32_t pk[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) pk[i] = pack16(f[2 * i], f[2 * i + 1], p.f16);
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

__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
               const ConvTcParams p) {
  const long long t_entry = clock64();
  pdl_wait();
#define OFA_TRACE(slot) do { if (p.trace && blockIdx.x < 16) p.trace[blockIdx.x * 16 + (slot)] = (unsigned long long)clock64(); } while (0)
  if (p.trace && blockIdx.x < 16 && threadIdx.x == 0) { p.trace[blockIdx.x * 16 + 0] = (unsigned long long)t_entry; OFA_TRACE(1); }
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space
  const int taps = p.ks * p.ks;
  const int B_BYTES = p.BN * BK * 2;
  const int a_stride = (p.a_bytes + 1023) & ~1023;
  uint8_t* sA = smem;
  uint8_t* sB = sA + (size_t)p.a_stages * a_stride;
  const size_t b_total = p.b_resident ? (size_t)taps * p.kcs * p.cout_pad * 128 : (size_t)p.b_stages * B_BYTES;
  float* s_scale = reinterpret_cast<float*>(sB + b_total);
  float* s_shift = s_scale + MAX_COUT;
  uint8_t* s_stage = reinterpret_cast<uint8_t*>(s_shift + MAX_COUT);   // NUM_EPI_WARPS x 2 KiB store staging
  uint64_t* a_full = reinterpret_cast<uint64_t*>(s_stage + NUM_EPI_WARPS * STAGE_BYTES);
  uint64_t* a_empty = a_full + MAX_A_STAGES;
  uint64_t* b_full = a_empty + MAX_A_STAGES;
  uint64_t* b_empty = b_full + MAX_B_STAGES;
  uint64_t* tfull = b_empty + MAX_B_STAGES;
  uint64_t* tempty = tfull + 2;
  uint64_t* bres_bar = tempty + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bres_bar + 1);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // provably warp-uniform
  const int lane = threadIdx.x & 31;

  for (int op = threadIdx.x; op < p.cout_pad; op += NUM_THREADS) {
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
    for (int s = 0; s < MAX_A_STAGES; ++s) { ptx::mbar_init(&a_full[s], 1); ptx::mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < MAX_B_STAGES; ++s) { ptx::mbar_init(&b_full[s], 1); ptx::mbar_init(&b_empty[s], 1); }
    for (int a = 0; a < 2; ++a) { ptx::mbar_init(&tfull[a], 1); ptx::mbar_init(&tempty[a], NUM_EPI_WARPS); }
    ptx::mbar_init(bres_bar, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_ptr, (uint32_t)p.tmem_cols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);   // warp-uniform by construction
  if (threadIdx.x == 0) OFA_TRACE(2);                                   // prologue done

  const int tiles_per_img = p.tiles_h * p.tiles_w;
  const int outer_splits = p.n_splits / p.inner_splits;
  const int num_work = p.N * tiles_per_img * outer_splits;
  const int R = p.ks / 2;
  const int acc_cols = MTX * p.BN;   // TMEM columns of one accumulator stage

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      if (p.b_resident) {
        ptx::mbar_arrive_expect_tx(bres_bar, (uint32_t)(taps * p.kcs * p.cout_pad * 128));
        for (int tk = 0; tk < taps * p.kcs; ++tk) {
          int tap = tk / p.kcs, kc = tk - tap * p.kcs;
          for (int r0 = 0; r0 < p.cout_pad; r0 += p.bres_rows)   // box rows divide cout_pad exactly
            ptx::tma_load_3d(sB + ((size_t)tk * p.cout_pad + r0) * 128, &tmap_w, bres_bar, kc * BK, r0, tap);
        }
      }
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      for (int wi = blockIdx.x; wi < num_work; wi += gridDim.x) {
        const int split0 = wi % outer_splits;
        const int sp = wi / outer_splits;
        const int n = sp / tiles_per_img;
        const int r = sp - n * tiles_per_img;
        const int h0 = (r / p.tiles_w) * TILE_H, w0 = (r % p.tiles_w) * TILE_W;
        for (int s = 0; s < p.inner_splits; ++s) {
          for (int kc = 0; kc < p.kcs; ++kc) {
            if (s == 0) {
              ptx::mbar_wait(&a_empty[as], aph ^ 1);
              ptx::mbar_arrive_expect_tx(&a_full[as], (uint32_t)p.a_bytes);
              ptx::tma_load_4d(sA + (size_t)as * a_stride, &tmap_x, &a_full[as], kc * BK, w0 - R, h0 - R, n);
              if (++as == p.a_stages) { as = 0; aph ^= 1; }
            }
            if (!p.b_resident) {
              const int row0 = (split0 * p.inner_splits + s) * p.BN;
              for (int tap = 0; tap < taps; ++tap) {
                ptx::mbar_wait(&b_empty[bs], bph ^ 1);
                ptx::mbar_arrive_expect_tx(&b_full[bs], (uint32_t)B_BYTES);
                ptx::tma_load_3d(sB + (size_t)bs * B_BYTES, &tmap_w, &b_full[bs], kc * BK, row0, tap);
                if (++bs == p.b_stages) { bs = 0; bph ^= 1; }
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The whole warp walks the loops (warp-uniform control flow, so descriptors live in uniform
    // registers); only the tcgen05 instructions are predicated on one lane.
    const uint32_t idesc = ptx::umma_idesc_f16(128, p.BN, p.f16 ? 0 : 1, p.f16 ? 0 : 1, 0, 0);
    const uint32_t sbo_a = (uint32_t)p.halo_w * 128;
    const uint32_t sB_addr = ptx::smem_u32(sB);
    int as = 0, bs = 0, acc = 0;
    uint32_t aph = 0, bph = 0, accph = 0;
    if (p.b_resident) ptx::mbar_wait(bres_bar, 0);
    if (lane == 0) OFA_TRACE(3);                                        // resident weights landed
    for (int wi = blockIdx.x; wi < num_work; wi += gridDim.x) {
      const int split0 = wi % outer_splits;
      const int as_item = as;           // first A stage of this work item
      const uint32_t aph_item = aph;
      for (int s = 0; s < p.inner_splits; ++s) {
        ptx::mbar_wait(&tempty[acc], accph ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * acc_cols);
        int a_cur = as_item;
        uint32_t a_cur_ph = aph_item;
        const int row0 = (split0 * p.inner_splits + s) * p.BN;
        uint32_t first = 0;             // 0 until the first MMA of this accumulator has been issued
        for (int kc = 0; kc < p.kcs; ++kc) {
          if (s == 0) ptx::mbar_wait(&a_full[a_cur], a_cur_ph);
          if (lane == 0 && wi == blockIdx.x && s == 0) OFA_TRACE(4 + (kc < 6 ? kc : 5));   // K chunk kc of the first item landed
          ptx::tc_fence_after();
          const uint32_t a_base = ptx::smem_u32(sA) + (uint32_t)(a_cur * a_stride);
          uint32_t b_res = sB_addr + (uint32_t)((kc * p.cout_pad + row0) * 128);   // resident: tap-major blocks
          const uint32_t b_res_step = (uint32_t)(p.kcs * p.cout_pad * 128);
          for (int ky = 0; ky < p.ks; ++ky) {
            uint32_t a_row = a_base + (uint32_t)(ky * p.halo_w * 128);
            for (int kx = 0; kx < p.ks; ++kx, a_row += 128, b_res += b_res_step) {
              uint32_t b_addr = b_res;
              if (!p.b_resident) {
                ptx::mbar_wait(&b_full[bs], bph);
                ptx::tc_fence_after();
                b_addr = sB_addr + (uint32_t)(bs * B_BYTES);
              }
              const uint64_t db = ptx::umma_desc_sw128(b_addr, 1024);
#pragma unroll
              for (int m = 0; m < MTX; ++m) {
                const uint64_t da = ptx::umma_desc_sw128(a_row + (uint32_t)(m * MT_COLS * 128), sbo_a);
#pragma unroll
                for (int k = 0; k < BK / 16; ++k)
                  ptx::umma_elect(d_tmem + (uint32_t)(m * p.BN), da + (uint64_t)(k * 2), db + (uint64_t)(k * 2),
                                      idesc, first | (uint32_t)k);
              }
              first = 1;
              if (!p.b_resident) {
                ptx::umma_commit_elect(&b_empty[bs]);
                if (++bs == p.b_stages) { bs = 0; bph ^= 1; }
              }
            }
          }
          if (s == p.inner_splits - 1) ptx::umma_commit_elect(&a_empty[a_cur]);
          if (++a_cur == p.a_stages) { a_cur = 0; a_cur_ph ^= 1; }
        }
        if (s == p.inner_splits - 1) { as = a_cur; aph = a_cur_ph; }
        ptx::umma_commit_elect(&tfull[acc]);
        if (lane == 0 && wi == blockIdx.x && s == 0) OFA_TRACE(10);      // first accumulator's MMAs issued
        if (++acc == 2) { acc = 0; accph ^= 1; }
      }
    }
  } else {
    // ===================== epilogue (warps 2..9) =====================
    const int ew = warp - 2;
    const int quarter = warp & 3;               // TMEM lane quarter this warp may read
    const int m = ew >> 2;                      // M-tile handled by this warp group (MTX == 2)
    const int row = quarter * 32 + lane;        // row of the M-tile = pixel (row / 8, row % 8)
    int acc = 0;
    uint32_t accph = 0;
    for (int wi = blockIdx.x; wi < num_work; wi += gridDim.x) {
      const int split0 = wi % outer_splits;
      const int sp = wi / outer_splits;
      const int n = sp / tiles_per_img;
      const int r = sp - n * tiles_per_img;
      const int h = (r / p.tiles_w) * TILE_H + row / MT_COLS;
      const int w = (r % p.tiles_w) * TILE_W + m * MT_COLS + row % MT_COLS;
      const bool pix_ok = h < p.H && w < p.W;
      for (int s = 0; s < p.inner_splits; ++s) {
        const int col0 = (split0 * p.inner_splits + s) * p.BN;
        ptx::mbar_wait(&tfull[acc], accph);
        if (threadIdx.x == 64 && wi == blockIdx.x) OFA_TRACE(11 + (s < 3 ? s : 2));   // accumulator s complete
        ptx::tc_fence_after();
        const uint32_t t_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * acc_cols + m * p.BN);
        uint8_t* stage = s_stage + ew * STAGE_BYTES;
        const int h_t0 = (r / p.tiles_w) * TILE_H, w_t0 = (r % p.tiles_w) * TILE_W + m * MT_COLS;
        for (int j = 0; j < p.BN; j += 32) {
          uint32_t v[32];
          const bool two = j + 16 < p.BN;
          ptx::tmem_ld16(t_addr + (uint32_t)j, v);
          if (two) ptx::tmem_ld16(t_addr + (uint32_t)(j + 16), v + 16);
          ptx::tmem_ld_wait();
          const int op0 = col0 + j;
          // (a 32-channel block must stay inside ONE PixelShuffle sub-pixel group of cout / 4 channels: cout = 64 or 192
          // have groups of 16 / 48 channels and take the 16-channel emit16 path)
          if (p.vec_store && j + 32 <= p.BN && op0 + 32 <= p.cout &&
              (p.store != OFA_STORE_PIXELSHUFFLE2 || ((p.cout >> 2) & 31) == 0)) {
            // ---- coalesced path: thread = pixel row computes, then the warp re-reads the 32 x 64-byte
            // block through a swizzled staging tile so each store instruction writes 8 x 64 contiguous bytes
            int sub = 0, oc0 = op0;
            if (p.store == OFA_STORE_PIXELSHUFFLE2) { const int q = p.cout >> 2; sub = op0 / q; oc0 = op0 - sub * q; }
            float f[32];
            affine_act_vec<32>(v, s_scale, s_shift, op0, p.act, f);
            if (p.res.ptr && pix_ok) {
              int oh = h, ow = w;
              if (p.store == OFA_STORE_PIXELSHUFFLE2) { oh = 2 * h + (sub >> 1); ow = 2 * w + (sub & 1); }
              const uint4* rp = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.res.ptr) +
                                                               p.res.off(n, oc0, oh, ow));
#pragma unroll
              for (int q4 = 0; q4 < 4; ++q4) {
                const uint4 rv = rp[q4];
                const uint32_t rr[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const float2 r2 = unpack16(rr[i], p.f16);
                  f[q4 * 8 + 2 * i] += r2.x;
                  f[q4 * 8 + 2 * i + 1] += r2.y;
                }
              }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              uint32_t pk[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) pk[i] = pack16(f[k * 8 + 2 * i], f[k * 8 + 2 * i + 1], p.f16);
              *reinterpret_cast<uint4*>(stage + lane * 64 + ((k ^ ((lane >> 1) & 3)) << 4)) =
                  make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
            __syncwarp();
#pragma unroll
            for (int it = 0; it < 4; ++it) {
              const int rr = it * 8 + (lane >> 2), piece = lane & 3;
              const uint4 val = *reinterpret_cast<const uint4*>(stage + rr * 64 + ((piece ^ ((rr >> 1) & 3)) << 4));
              const int R2 = quarter * 32 + rr;
              const int hh = h_t0 + R2 / MT_COLS, ww = w_t0 + R2 % MT_COLS;
              if (hh < p.H && ww < p.W) {
                int oh = hh, ow = ww;
                if (p.store == OFA_STORE_PIXELSHUFFLE2) { oh = 2 * hh + (sub >> 1); ow = 2 * ww + (sub & 1); }
                *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.y.ptr) + p.y.off(n, oc0 + piece * 8, oh, ow)) = val;
              }
            }
            __syncwarp();
          } else if (pix_ok) {
            emit16(p, s_scale, s_shift, v, op0, n, h, w);
            if (two) emit16(p, s_scale, s_shift, v + 16, op0 + 16, n, h, w);
          }
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&tempty[acc]);
        if (++acc == 2) { acc = 0; accph ^= 1; }
      }
    }
  }

  if (threadIdx.x == 64) OFA_TRACE(14);                                 // epilogue warp 2 done
  ptx::tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) OFA_TRACE(15);
#undef OFA_TRACE
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

// accumulator width per M-tile: largest multiple of 16 that is <= 128 and divides cout_pad
int pick_bn(int cout_pad) {
  for (int bn = 128; bn >= 16; bn -= 16)
    if (cout_pad % bn == 0) return bn;
  return 0;
}

}  // namespace

bool conv_tc_supported(const OfaConvArgs* a) {
  if (!a->w_bf16) return false;
  if (a->flip) return false;
  if (!is_16bit(a->x.dtype) || !is_nhwc_dense(&a->x)) return false;
  if ((reinterpret_cast<uintptr_t>(a->x.ptr) & 15) || (reinterpret_cast<uintptr_t>(a->w_bf16) & 15)) return false;
  if (a->cin % BK != 0 || a->cin_pad != a->cin) return false;
  if (a->ks > 7 || !(a->ks & 1)) return false;
  if (a->cout_pad % 16 != 0 || a->cout_pad < a->cout || a->cout_pad > MAX_COUT) return false;
  if (a->store == OFA_STORE_PIXELSHUFFLE2 && (a->cout % 64 != 0)) return false;
  if (a->x.n <= 0 || a->x.h <= 0 || a->x.w <= 0) return false;
  return true;
}

// A 1x1 conv has no spatial structure: when the image edges do not fill the 16 x 16 pixel tiles (the 24 x 24 LR
// training patches use 56 % of 2 x 2 tiles each), the dense NHWC tensors are re-viewed as ONE image of P / 16 rows
// by 16 pixels, which tiles exactly.
static bool flatten_pointwise(const OfaConvArgs* a, OfaConvArgs* flat) {
  if (a->ks != 1 || a->store != OFA_STORE_PLAIN) return false;
  if (a->x.h % TILE_H == 0 && a->x.w % TILE_W == 0) return false;
  const long long P = (long long)a->x.n * a->x.h * a->x.w;
  if (P % TILE_W != 0 || P / TILE_W >= (1ll << 31)) return false;
  if (!is_nhwc_dense(&a->x) || !is_nhwc_dense(&a->y)) return false;
  if (a->epi.residual && !is_nhwc_dense(a->epi.residual)) return false;
  *flat = *a;
  auto reshape = [&](OfaTensor4& t) {
    t.n = 1; t.h = (int32_t)(P / TILE_W); t.w = TILE_W;
    t.sc = 1; t.sw = t.c; t.sh = (int64_t)TILE_W * t.c; t.sn = P * t.c;
  };
  reshape(flat->x);
  reshape(flat->y);
  return true;
}

int launch_conv_tc(const OfaConvArgs* a_in, cudaStream_t st) {
  OfaConvArgs a_flat;
  OfaTensor4 res_flat;
  const OfaConvArgs* a = a_in;
  if (flatten_pointwise(a_in, &a_flat)) {
    if (a_in->epi.residual) {
      res_flat = *a_in->epi.residual;
      res_flat.n = 1; res_flat.h = a_flat.x.h; res_flat.w = TILE_W;
      res_flat.sc = 1; res_flat.sw = res_flat.c; res_flat.sh = (int64_t)TILE_W * res_flat.c;
      res_flat.sn = (int64_t)a_flat.x.h * TILE_W * res_flat.c;
      a_flat.epi.residual = &res_flat;
    }
    a = &a_flat;
  }
  ConvTcParams p;
  memset(&p, 0, sizeof(p));
  p.N = a->x.n; p.H = a->x.h; p.W = a->x.w;
  p.cin = a->cin; p.cout = a->cout; p.ks = a->ks;
  p.kcs = a->cin / BK;
  p.cout_pad = a->cout_pad;
  p.BN = pick_bn(a->cout_pad);
  p.n_splits = a->cout_pad / p.BN;
  p.inner_splits = (p.kcs == 1) ? p.n_splits : 1;
  p.tiles_h = (p.H + TILE_H - 1) / TILE_H;
  p.tiles_w = (p.W + TILE_W - 1) / TILE_W;
  p.halo_w = TILE_W + p.ks - 1;
  p.halo_h = TILE_H + p.ks - 1;
  p.a_bytes = p.halo_w * p.halo_h * 128;
  p.store = a->store;
  p.act = a->epi.act;
  p.f16 = a->x.dtype == OFA_F16 ? 1 : 0;
  p.gamma = a->epi.gamma; p.beta = a->epi.beta; p.mean = a->epi.mean; p.var = a->epi.var; p.eps = a->epi.eps;
  p.y = make_tv(&a->y);
  p.res = a->epi.residual ? make_tv(a->epi.residual) : null_tv();
  int cols = 32;
  while (cols < 2 * MTX * p.BN) cols <<= 1;
  p.tmem_cols = cols;

  const int taps = p.ks * p.ks;
  const int b_bytes = p.BN * BK * 2;
  const long long b_all = (long long)taps * p.kcs * p.cout_pad * 128;
  p.b_resident = b_all <= B_RESIDENT_LIMIT ? 1 : 0;
  const int a_stride = (p.a_bytes + 1023) & ~1023;
  const int fixed = 1024 /*align slack*/ + 2 * MAX_COUT * 4 + NUM_EPI_WARPS * STAGE_BYTES + 512;
  const int budget = 227 * 1024 - fixed;
  if (p.b_resident) {
    p.b_stages = 0;
    int as = (int)((budget - b_all) / a_stride);
    if (as > MAX_A_STAGES) as = MAX_A_STAGES;
    if (as < 1) return fail(OFA_ERR_UNSUPPORTED, "conv_tc: halo tile does not fit shared memory");
    p.a_stages = as;
  } else {
    // at least 2 halo stages (prefetch the next tile), the rest to the weight ring
    int as = (p.kcs > 1) ? 3 : 2;
    int bs = (budget - as * a_stride) / b_bytes;
    if (bs > MAX_B_STAGES) bs = MAX_B_STAGES;
    if (bs < 2) { as = 1; bs = (budget - as * a_stride) / b_bytes; if (bs > MAX_B_STAGES) bs = MAX_B_STAGES; }
    if (bs < 2) return fail(OFA_ERR_UNSUPPORTED, "conv_tc: tile does not fit shared memory");
    p.a_stages = as;
    p.b_stages = bs;
  }
  const size_t smem = (size_t)fixed + (size_t)p.a_stages * a_stride +
                      (p.b_resident ? (size_t)b_all : (size_t)p.b_stages * b_bytes);

  bool vec = a->y.dtype == a->x.dtype && a->y.sc == 1 && (reinterpret_cast<uintptr_t>(a->y.ptr) & 15) == 0 &&
             a->y.sw % 8 == 0 && a->y.sh % 8 == 0 && a->y.sn % 8 == 0 && a->store != OFA_STORE_PIXELUNSHUFFLE2;
  if (a->epi.residual) {
    const OfaTensor4* r = a->epi.residual;
    vec = vec && r->dtype == a->x.dtype && r->sc == 1 && (reinterpret_cast<uintptr_t>(r->ptr) & 15) == 0 &&
          r->sw % 8 == 0 && r->sh % 8 == 0 && r->sn % 8 == 0;
  }
  p.vec_store = vec ? 1 : 0;

  CUtensorMap tx, tw;
  {
    uint64_t dims[4] = {(uint64_t)p.cin, (uint64_t)p.W, (uint64_t)p.H, (uint64_t)p.N};
    uint64_t strides[3] = {(uint64_t)p.cin * 2, (uint64_t)p.W * p.cin * 2, (uint64_t)p.H * p.W * p.cin * 2};
    uint32_t box[4] = {BK, (uint32_t)p.halo_w, (uint32_t)p.halo_h, 1};
    int rc = encode_tmap(&tx, p.f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, a->x.ptr,
                         dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)a->cin_pad, (uint64_t)a->cout_pad, (uint64_t)taps};
    uint64_t strides[2] = {(uint64_t)a->cin_pad * 2, (uint64_t)a->cin_pad * a->cout_pad * 2};
    p.bres_rows = a->cout_pad <= 256 ? a->cout_pad : p.BN;
    uint32_t rows = p.b_resident ? (uint32_t)p.bres_rows : (uint32_t)p.BN;
    uint32_t box[3] = {BK, rows, 1};
    int rc = encode_tmap(&tw, p.f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3,
                         const_cast<void*>(a->w_bf16), dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  static int trace_on = -1;
  if (trace_on < 0) { const char* e = getenv("OFA_CONV_TC_TRACE"); trace_on = (e && e[0] == '1') ? 1 : 0; }
  p.trace = nullptr;
  if (trace_on) {
    static unsigned long long* tbuf = nullptr;
    if (!tbuf) cudaMalloc(reinterpret_cast<void**>(&tbuf), 16 * 16 * sizeof(unsigned long long));
    cudaMemsetAsync(tbuf, 0, 16 * 16 * sizeof(unsigned long long), st);
    p.trace = tbuf;
  }
  static unsigned char attr_done[64] = {0};
  if (once_per_device(attr_done))
    OFA_CUDA(cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  const int num_work = p.N * p.tiles_h * p.tiles_w * (p.n_splits / p.inner_splits);
  int grid = sm_count();
  if (grid > num_work) grid = num_work;
  launch_pdl(conv_tc_kernel, dim3(grid), dim3(NUM_THREADS), smem, st, tx, tw, p);
  if (trace_on) {
    unsigned long long h[16 * 16];
    cudaStreamSynchronize(st);
    cudaMemcpy(h, p.trace, sizeof(h), cudaMemcpyDeviceToHost);
    for (int b = 0; b < 2; ++b) {
      const unsigned long long* t = h + b * 16;
      fprintf(stderr, "conv_tc trace cin %d cout %d ks %d P %d cta %d (clk from entry): pdl_wait %lld prologue %lld weights %lld A0 %lld A1 %lld A5 %lld mma_issued %lld acc0 %lld acc1 %lld acc2 %lld epi_done %lld end %lld\n",
              p.cin, p.cout, p.ks, p.N * p.H * p.W, b, (long long)(t[1] - t[0]), (long long)(t[2] - t[0]), (long long)(t[3] - t[0]), (long long)(t[4] - t[0]),
              t[5] ? (long long)(t[5] - t[0]) : -1ll, t[9] ? (long long)(t[9] - t[0]) : -1ll, (long long)(t[10] - t[0]), (long long)(t[11] - t[0]),
              t[12] ? (long long)(t[12] - t[0]) : -1ll, t[13] ? (long long)(t[13] - t[0]) : -1ll, (long long)(t[14] - t[0]), (long long)(t[15] - t[0]));
    }
  }
  return check_launch("conv_tc_kernel");
}

}  // namespace ofa
