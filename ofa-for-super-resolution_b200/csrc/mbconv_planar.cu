// Inference MBConv block (dynamic_layers.py:70-84 + proxyless_nets.py:44-51) as three tcgen05 kernels
// around a CHANNEL-PLANAR 16-bit intermediate:
//
//   trunk x [P,64] NHWC bf16 --expand--> mid1 [N][M][H][W] --depthwise--> mid2 [N][M][H][W] --project--> y [P,64]
//
//   * expand  (a3): D[c_out, pixel] = W[c_out, :64] . x[pixel, :64]^T.  The weight slice is the UMMA A
//     operand (M = 128 output channels), 256 pixels are N.  An accumulator lane is therefore one
//     CHANNEL and its columns are consecutive pixels: the epilogue (folded BN + ReLU6, scale/shift in
//     registers) emits channel-planar rows, written back with swizzled TMA stores.
//   * depthwise (a1 + a2): per channel plane the ks x ks filter is a sum over dy of banded Toeplitz
//     products  out[y, :] += in[y + dy, :] . T_dy,  T_dy[x_in, x_out] = f[dy][x_in - x_out].
//     A = the plane tile [128 + ks - 1 rows][64 columns] exactly as TMA lands it (K-major, 128-byte
//     swizzle; out-of-bounds zero fill = the conv's padding); the dy shift is a descriptor start
//     address shifted by dy rows.  Each 16-column K chunk touches a 32-column output window, so a
//     tile costs 4 * ks MMAs of M128 x N32 x K16 and the kernel is bound by HBM, not by CUDA-core
//     FMA throughput (the 7x7 SIMT kernel was).  The Toeplitz B tiles are built in shared memory by
//     a helper warp from the on-the-fly 7->5->3 transformed filter (dynamic_op.py:46-71).
//   * project (a4 + a8): D[pixel, c_out] = mid2[:, pixel]^T . W[c_out, :]^T with the planar tensor as
//     an MN-major A operand (TMA boxes of 64 pixels x 64 channels); folded BN + the residual (loaded
//     by TMA into the same staging tile the TMA store later reads) in the epilogue.
//
// The intermediates are fp16 (values are ReLU6-clamped to [0, 6], where fp16 carries 3 more mantissa
// bits than bf16) or bf16; accumulation is fp32 in TMEM.  Persistent CTAs, one per SM.
#include "ofa_common.cuh"
#include "kernels.h"
#include "sm100_ptx.cuh"
#include "mbconv_planar.cuh"

#include <string.h>

namespace ofa {
namespace {

// ==================================================================================================
// weight packing for the block: expand -> bf16 [mt*128][64] (A operand), project -> 16-bit [64][mid]
// ==================================================================================================
__global__ void pack_block_weights_kernel(const float* __restrict__ w_exp, long long e_so, long long e_si,
                                          const float* __restrict__ w_proj, long long p_so, long long p_si,
                                          int cin, int mid, int cout, int mid_pad, int trunk_f16, int f16,
                                          uint16_t* __restrict__ out_exp, uint16_t* __restrict__ out_proj) {
  const int n_exp = mid_pad * cin, n_proj = cout * mid;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_exp + n_proj; i += gridDim.x * blockDim.x) {
    if (i < n_exp) {
      const int o = i / cin, ci = i - o * cin;
      out_exp[i] = cvt16(o < mid ? w_exp[o * e_so + ci * e_si] : 0.f, trunk_f16);
    } else {
      const int j = i - n_exp;
      const int o = j / mid, ci = j - o * mid;
      out_proj[j] = cvt16(w_proj[o * p_so + ci * p_si], f16);
    }
  }
}

// ==================================================================================================
// (1) expand: NHWC bf16 trunk -> planar 16-bit, folded BN + activation
// ==================================================================================================
struct ExpandParams {
  int N, HW, mid, mt, f16, xf16, act;
  const float* gamma; const float* beta; const float* mean; const float* var; float eps;
  int tiles_per_img;
};

// F16: fp16 (1) or bf16 (0) output; RELU6: the activation is ReLU6 (else p.act at run time)
template <int F16, int RELU6>
__global__ void __launch_bounds__(EX_THREADS, 1)
expand_planar_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w,
                     const __grid_constant__ CUtensorMap tm_y, const ExpandParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space
  uint8_t* sW = smem;                                              // mt x 16 KiB, resident
  uint8_t* sX = sW + EX_MAX_MT * 16384;                            // pixel-tile ring
  uint8_t* sS = sX + EX_X_STAGES * EX_X_BYTES;                     // 8 warps x EX_SBUFS x 4 KiB staging
  uint64_t* x_full = reinterpret_cast<uint64_t*>(sS + EX_EPI_WARPS * EX_SBUFS * EX_SBUF_BYTES);
  uint64_t* x_empty = x_full + EX_X_STAGES;
  uint64_t* tfull = x_empty + EX_X_STAGES;
  uint64_t* tempty = tfull + 2;
  uint64_t* w_bar = tempty + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(w_bar + 1);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tm_x); ptx::prefetch_tmap(&tm_w); ptx::prefetch_tmap(&tm_y);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < EX_X_STAGES; ++s) { ptx::mbar_init(&x_full[s], 1); ptx::mbar_init(&x_empty[s], 1); }
    for (int a = 0; a < 2; ++a) { ptx::mbar_init(&tfull[a], 1); ptx::mbar_init(&tempty[a], EX_EPI_WARPS); }
    ptx::mbar_init(w_bar, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 2) { ptx::tmem_alloc(tmem_ptr, 512); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);
  const int num_tiles = p.N * p.tiles_per_img;

  if (warp == 0) {
    if (lane == 0) {
      // Programmatic dependent launch: everything above (barriers, TMEM, tensor-map fetches) ran while the previous
      // kernel was still draining on other SMs.  The packed weights may come from the launch just before this one
      // (blocks that pack per call), so they too are loaded only after the wait.
      pdl_wait();
      ptx::mbar_arrive_expect_tx(w_bar, (uint32_t)(p.mt * 16384));
      for (int m = 0; m < p.mt; ++m) ptx::tma_load_3d(sW + m * 16384, &tm_w, w_bar, 0, m * 128, 0);
      int s = 0; uint32_t ph = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int n = t / p.tiles_per_img, p0 = (t - n * p.tiles_per_img) * EX_NPIX;
        ptx::mbar_wait(&x_empty[s], ph ^ 1);
        ptx::mbar_arrive_expect_tx(&x_full[s], EX_X_BYTES);
        ptx::tma_load_3d(sX + s * EX_X_BYTES, &tm_x, &x_full[s], 0, p0, n);
        if (++s == EX_X_STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    const int xfmt = p.xf16 ? 0 : 1;
    const uint32_t idesc = ptx::umma_idesc_f16(128, EX_NPIX, xfmt, xfmt, 0, 0);
    const uint32_t sW_addr = ptx::smem_u32(sW), sX_addr = ptx::smem_u32(sX);
    int s = 0, acc = 0; uint32_t ph = 0, accph = 0;
    ptx::mbar_wait(w_bar, 0);
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      ptx::mbar_wait(&x_full[s], ph);
      ptx::tc_fence_after();
      for (int m = 0; m < p.mt; ++m) {
        ptx::mbar_wait(&tempty[acc], accph ^ 1);
        ptx::tc_fence_after();
        const uint64_t da = ptx::umma_desc_sw128(sW_addr + (uint32_t)(m * 16384), 1024);
        const uint64_t db = ptx::umma_desc_sw128(sX_addr + (uint32_t)(s * EX_X_BYTES), 1024);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          ptx::umma_elect(tmem_base + (uint32_t)(acc * 256), da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc,
                              (uint32_t)k);
        if (m == p.mt - 1) ptx::umma_commit_elect(&x_empty[s]);
        ptx::umma_commit_elect(&tfull[acc]);
        if (++acc == 2) { acc = 0; accph ^= 1; }
      }
      if (++s == EX_X_STAGES) { s = 0; ph ^= 1; }
    }
  } else {
    const int ew = warp - 2;
    const int quarter = warp & 3;
    const int half = ew >> 2;                       // which 128 accumulator columns
    uint8_t* sbuf = sS + ew * EX_SBUFS * EX_SBUF_BYTES;
    int acc = 0, sb = 0; uint32_t accph = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int n = t / p.tiles_per_img, p0 = (t - n * p.tiles_per_img) * EX_NPIX;
      for (int m = 0; m < p.mt; ++m) {
        const int c_warp = m * 128 + quarter * 32;
        const int c = c_warp + lane;
        float scale = 0.f, shift = 0.f;
        if (c < p.mid) bn_fold(p.gamma, p.beta, p.mean, p.var, p.eps, c, scale, shift);
        ptx::mbar_wait(&tfull[acc], accph);
        ptx::tc_fence_after();
        const uint32_t t_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * 256 + half * 128);
#pragma unroll 1
        for (int bx = 0; bx < 2; ++bx) {
          const int px = p0 + half * 128 + bx * 64;
          const bool live = (px < p.HW) && (c_warp < p.mid);      // warp-uniform
          if (live) {
            uint32_t v[64];
            ptx::tmem_ld16(t_addr + (uint32_t)(bx * 64), v);
            ptx::tmem_ld16(t_addr + (uint32_t)(bx * 64 + 16), v + 16);
            ptx::tmem_ld16(t_addr + (uint32_t)(bx * 64 + 32), v + 32);
            ptx::tmem_ld16(t_addr + (uint32_t)(bx * 64 + 48), v + 48);
            ptx::tmem_ld_wait();
            // (The TMA engine takes one 128-byte row request every ~8.6 clocks -- 15 B/clk per SM; UTMACMDFLUSH holds 26 %
            // of the stall samples, profiles/r2_ncu_expand_planar.csv -- and that, not HBM, bounds this kernel.  Tried:
            // storing every other sub-block straight from registers, a lane's 128 contiguous bytes as eight 16-byte
            // st.global: 0.096 -> 0.150 ms, the scattered half-sector writes are far worse.)
            if (lane == 0) ptx::tma_store_wait_read<EX_SBUFS - 1>();   // staging buffer `sb` is free again
            __syncwarp();
            uint8_t* dst = sbuf + sb * EX_SBUF_BYTES + lane * 128;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              uint32_t pk[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                float a = fmaf(__uint_as_float(v[j * 8 + 2 * i]), scale, shift);
                float b = fmaf(__uint_as_float(v[j * 8 + 2 * i + 1]), scale, shift);
                if (RELU6) {
                  pk[i] = pack16_relu6(a, b, F16);
                } else {
                  a = apply_act(a, p.act); b = apply_act(b, p.act);
                  pk[i] = pack16(a, b, F16);
                }
              }
              *reinterpret_cast<uint4*>(dst + ((j ^ (lane & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
            ptx::fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              ptx::tma_store_3d(&tm_y, sbuf + sb * EX_SBUF_BYTES, px, c_warp, n);
              ptx::tma_store_commit();
            }
            if (++sb == EX_SBUFS) sb = 0;
          }
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&tempty[acc]);
        if (++acc == 2) { acc = 0; accph ^= 1; }
      }
    }
    if (lane == 0) ptx::tma_store_wait_all<0>();
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem_base, 512); }
}

// ==================================================================================================
// (2) depthwise ks x ks on channel planes as banded-Toeplitz tcgen05 MMAs
// ==================================================================================================
// Tile geometry.  TMA needs the box start 16-byte aligned in the innermost dimension, so a tile's input
// window starts DW_XPAD = 8 columns left of its first output column (x0 is a multiple of 8) and spans two
// 128-byte swizzle atoms = 128 input columns, of which 112 produce outputs:
//     out[y0 + r][x0 + xo] = sum_{dy,dx} f[dy][dx] * tile[r + dy][xo + (8 - R) + dx],   R = ks / 2
// Input column chunk j (16 columns) only reaches output columns 16j - 8 - R .. 16j + 7 + R, i.e. the 32
// accumulator columns starting at 16j when output column xo lives in accumulator column xo + 16:
//     B_dy[k][n] = f[dy][k - n + 8 + R]   (zero outside the band), the same matrix for every chunk.
// The very first MMA of a tile (dy = 0, chunk 0) uses the same matrix zero-extended to all 144 accumulator
// columns with accumulate = 0: it initialises the accumulator, every later MMA accumulates.
// (Tried: 5 input stages with a single output staging tile -- 0.192 ms instead of 0.179 at C2: the epilogue then waits
// for the TMA engine to get to the previous store behind the queued prefetch loads.)
constexpr int DWP_A_STAGES = 4;

struct DwPlanarParams {
  int NC, C, H, W, ks, kmax, transform_on, f16, act;
  const float* filt;   // [C][ks * ks] active filters (fp32), derived once per launch by active_filter_kernel
  int filt_from_kernel; // 1: `filt` is the output of the launch preceding this one (else the parameter itself)
  const float* gamma; const float* beta; const float* mean; const float* var; float eps;
  int tiles_x, tiles_y;
  int total_tiles;
  int short_last;   // the last tile row has <= 64 valid rows: it runs as an M = 64 MMA tile
};

// tile index -> plane, tile origin, and whether it is a short (M = 64) tile
__device__ __forceinline__ bool dw_decode(const DwPlanarParams& p, int t, int& pc, int& y0, int& x0) {
  const unsigned tx = (unsigned)t % (unsigned)p.tiles_x;
  const unsigned r = (unsigned)t / (unsigned)p.tiles_x;
  const unsigned ty = r % (unsigned)p.tiles_y;
  pc = (int)(r / (unsigned)p.tiles_y);
  y0 = (int)ty * DW_TH;
  x0 = (int)tx * DW_TW;
  return p.short_last && (int)ty == p.tiles_y - 1;
}
// KS: kernel size; F16: fp16 (1) or bf16 (0) storage; RELU6: the activation is ReLU6 (else p.act at run time)
template <int KS, int F16, int RELU6>
__global__ void __launch_bounds__(DW_THREADS, 1)
dw_planar_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_xs,
                 const __grid_constant__ CUtensorMap tm_y, const DwPlanarParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space
  uint8_t* sA = smem;
  uint8_t* sO = sA + DWP_A_STAGES * DW_A_STRIDE;                   // 2 output staging tiles (1024-aligned)
  uint8_t* sB = sO + 2 * DW_OUT_BYTES;                             // 2 filter buffers
  uint16_t* s_rows = reinterpret_cast<uint16_t*>(sB + 2 * DW_B_BYTES);   // [KS][64] zero-padded 16-bit filter rows
  uint64_t* a_full = reinterpret_cast<uint64_t*>(s_rows + 512);
  uint64_t* a_empty = a_full + DWP_A_STAGES;
  uint64_t* tfull = a_empty + DWP_A_STAGES;
  uint64_t* tempty = tfull + DW_ACC_STAGES;
  uint64_t* b_full = tempty + DW_ACC_STAGES;
  uint64_t* b_empty = b_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(b_empty + 2);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) { ptx::prefetch_tmap(&tm_x); ptx::prefetch_tmap(&tm_xs); ptx::prefetch_tmap(&tm_y); }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < DWP_A_STAGES; ++s) { ptx::mbar_init(&a_full[s], 1); ptx::mbar_init(&a_empty[s], 1); }
    for (int a = 0; a < DW_ACC_STAGES; ++a) { ptx::mbar_init(&tfull[a], 1); ptx::mbar_init(&tempty[a], DW_EPI_WARPS); }
    // both MMA issuer warps release a filter buffer (each commits at every plane switch)
    for (int b = 0; b < 2; ++b) { ptx::mbar_init(&b_full[b], 1); ptx::mbar_init(&b_empty[b], 2); }
    ptx::fence_barrier_init();
  }
  if (warp == 2) { ptx::tmem_alloc(tmem_ptr, 512); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);

  const int t_begin = (int)((long long)p.total_tiles * blockIdx.x / gridDim.x);
  const int t_end = (int)((long long)p.total_tiles * (blockIdx.x + 1) / gridDim.x);
  constexpr int R = KS >> 1;
  constexpr uint32_t atom_bytes = (uint32_t)((DW_TH + KS - 1) * 128);
  constexpr uint32_t atom_bytes_short = (uint32_t)((DW_TH / 2 + KS - 1) * 128);

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      pdl_wait();                                   // the planes are the previous kernel's output
      int s = 0; uint32_t ph = 0;
      for (int t = t_begin; t < t_end; ++t) {
        int pc, y0, x0;
        const bool shrt = dw_decode(p, t, pc, y0, x0);
        const CUtensorMap* tm = shrt ? &tm_xs : &tm_x;
        ptx::mbar_wait(&a_empty[s], ph ^ 1);
        ptx::mbar_arrive_expect_tx(&a_full[s], 2 * (shrt ? atom_bytes_short : atom_bytes));
        ptx::tma_load_3d(sA + s * DW_A_STRIDE, tm, &a_full[s], x0 - DW_XPAD, y0 - R, pc);
        ptx::tma_load_3d(sA + s * DW_A_STRIDE + DW_ATOM_STRIDE, tm, &a_full[s], x0 - DW_XPAD + 64, y0 - R, pc);
        if (++s == DWP_A_STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1 || warp == DW_ISSUER1_WARP) {
    // ===================== MMA issuers =====================
    // ONE warp can issue a tcgen05.mma only every ~51 clocks (measured, csrc/experiments/umma_rate_probe.cu), but an
    // M128 x N32 x K16 MMA occupies the tensor pipe for 40 (its 5 KB of shared-memory operands at 128 B/clk): with a
    // single issuer the kernel ran at ~65 clocks per MMA.  Two issuer warps take ALTERNATE tiles -- every tile's MMAs still
    // come from one thread, in order, into that tile's own accumulator stage -- and their streams interleave in the pipe.
    // Both walk the whole tile list (stage / phase bookkeeping stays in step) and both arrive on b_empty at every plane
    // switch; a commit with no MMAs of its own outstanding arrives at once.
    const int issuer = warp == 1 ? 0 : 1;
    constexpr int fmt = F16 ? 0 : 1;
    constexpr uint32_t idesc32_tall = ptx::umma_idesc_f16(128, 32, fmt, fmt, 0, 0);
    constexpr uint32_t idesc_first_tall = ptx::umma_idesc_f16(128, DW_ACC_COLS, fmt, fmt, 0, 0);
    // M = 64 variants (short tiles): half the A rows are read; the accumulator rows land in lanes
    // 0-15 of each 32-lane quarter (row m -> lane 32 * (m / 16) + m % 16)
    constexpr uint32_t idesc32_short = ptx::umma_idesc_f16(64, 32, fmt, fmt, 0, 0);
    constexpr uint32_t idesc_first_short = ptx::umma_idesc_f16(64, DW_ACC_COLS, fmt, fmt, 0, 0);
    const uint32_t sA_addr = ptx::smem_u32(sA), sB_addr = ptx::smem_u32(sB);
    int s = 0, acc = 0, bi = 0, cur_pc = -1;
    uint32_t ph = 0, accph = 0, bph = 0;
    for (int t = t_begin; t < t_end; ++t) {
      int pc, y0, x0;
      const bool shrt = dw_decode(p, t, pc, y0, x0);
      const uint32_t idesc32 = shrt ? idesc32_short : idesc32_tall;
      const uint32_t idesc_first = shrt ? idesc_first_short : idesc_first_tall;
      if (pc != cur_pc) {
        if (cur_pc >= 0) {
          ptx::umma_commit_elect(&b_empty[bi]);     // all MMAs reading the old filter tiles are done
          if (++bi == 2) { bi = 0; bph ^= 1; }
        }
        ptx::mbar_wait(&b_full[bi], bph);
        cur_pc = pc;
      }
      if (((t - t_begin) & 1) != issuer) {          // the other issuer's tile
        if (++s == DWP_A_STAGES) { s = 0; ph ^= 1; }
        if (++acc == DW_ACC_STAGES) { acc = 0; accph ^= 1; }
        continue;
      }
      ptx::mbar_wait(&a_full[s], ph);
      ptx::mbar_wait(&tempty[acc], accph ^ 1);
      ptx::tc_fence_after();
      const uint32_t d0 = tmem_base + (uint32_t)(acc * DW_ACC_COLS);
      // one descriptor per operand per tile; every MMA adds a compile-time constant to its low word
      const uint64_t da0 = ptx::umma_desc_sw128(sA_addr + (uint32_t)(s * DW_A_STRIDE), 1024);
      const uint64_t dbf = ptx::umma_desc(sB_addr + (uint32_t)(bi * DW_B_BYTES), 128, 256, 0);
      const uint64_t db0 = dbf + (uint64_t)(DW_BFIRST_BYTES >> 4);
      // ragged right edge: input chunks past the reach of the last valid output column are skipped (the
      // first MMA has zero-initialised every accumulator column)
      const int valid_w = min(DW_TW, p.W - x0);
      const int nch = min(DW_CHUNKS, (valid_w + DW_XPAD + R + 15) >> 4);
#pragma unroll
      for (int dy = 0; dy < KS; ++dy) {
#pragma unroll
        for (int j = 0; j < DW_CHUNKS; ++j) {
          // chunk j: atom j / 4, 32 bytes per chunk inside the 128-byte swizzled row, shifted down dy rows
          const uint64_t da = da0 + (uint64_t)(((j >> 2) * DW_ATOM_STRIDE + dy * 128 + (j & 3) * 32) >> 4);
          if (dy == 0 && j == 0)
            ptx::umma_elect(d0, da, dbf, idesc_first, 0u);
          else if (j < nch)
            ptx::umma_elect(d0 + (uint32_t)(16 * j), da, db0 + (uint64_t)((dy * 1024) >> 4), idesc32, 1u);
        }
      }
      ptx::umma_commit_elect(&a_empty[s]);
      ptx::umma_commit_elect(&tfull[acc]);
      if (++s == DWP_A_STAGES) { s = 0; ph ^= 1; }
      if (++acc == DW_ACC_STAGES) { acc = 0; accph ^= 1; }
    }
  } else if (warp == 2) {
    // ===================== filter builder: active filter -> Toeplitz B tiles =====================
    // Element (n, k) of a tile is f[dy][k - n + dx0]: only rows n_lo..n_hi can be non-zero, everything else is zeroed
    // ONCE; a row's two 16-byte K-halves are 8 consecutive entries of the zero-padded 16-bit filter row s_rows[dy].
    // The whole warp builds (round 1: lane 0 derived the filter with the 25 x 25 / 9 x 9 mat-vecs per plane, ~6 us, which
    // is why planes of fewer than ~45 tiles ran slowly); the transformed filters now come from one tiny launch.
    int bi = 0, cur_pc = -1; uint32_t bph = 0;
    constexpr int dx0 = DW_XPAD + R;
    constexpr int PADL = 16, ROWLEN = 64;
    constexpr int n_lo = dx0 - KS + 1, n_hi = dx0 + 15, NR = n_hi - n_lo + 1;
    for (int i = lane; i < (2 * DW_B_BYTES) / 16; i += 32)
      *reinterpret_cast<uint4*>(sB + 16 * i) = make_uint4(0u, 0u, 0u, 0u);
    for (int i = lane; i < KS * ROWLEN; i += 32) s_rows[i] = (uint16_t)0;
    // a lane owns up to two (row n, K-half) vectors of EVERY tile: source / destination offsets are loop invariants
    int src_off[2], dst_off[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int q = lane + 32 * h;
      const int n = n_lo + (q >> 1), kh = q & 1;
      src_off[h] = q < NR * 2 ? kh * 8 - n + dx0 + PADL : -1;
      dst_off[h] = dw_b_off(n, kh * 8);
    }
    __syncwarp();
    if (p.filt_from_kernel) pdl_wait();             // transformed filters: written by the launch just before this one
    for (int t = t_begin; t < t_end; ++t) {
      int pc, y0, x0;
      dw_decode(p, t, pc, y0, x0);
      if (pc == cur_pc) continue;
      cur_pc = pc;
      const float* f = p.filt + (size_t)(pc % p.C) * KS * KS;
      const float t0 = lane < KS * KS ? f[lane] : 0.f;                 // issued before the wait: latency overlaps it
      const float t1 = lane + 32 < KS * KS ? f[lane + 32] : 0.f;
      ptx::mbar_wait(&b_empty[bi], bph ^ 1);
      if (lane < KS * KS) s_rows[(lane / KS) * ROWLEN + PADL + lane % KS] = cvt16(t0, F16);
      if (lane + 32 < KS * KS) s_rows[((lane + 32) / KS) * ROWLEN + PADL + (lane + 32) % KS] = cvt16(t1, F16);
      __syncwarp();
      uint8_t* b0 = sB + bi * DW_B_BYTES;
      // tile 0: the N = 144 first matrix (dy = 0); tiles 1..KS: the N = 32 matrices of dy = 0..KS-1
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (src_off[h] >= 0) {
#pragma unroll
          for (int tt = 0; tt <= KS; ++tt) {
            const int dy = tt == 0 ? 0 : tt - 1;
            const uint16_t* src = s_rows + dy * ROWLEN + src_off[h];
            uint32_t w[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) w[q] = (uint32_t)src[2 * q] | ((uint32_t)src[2 * q + 1] << 16);
            uint8_t* dst = (tt == 0 ? b0 : b0 + DW_BFIRST_BYTES + dy * 1024) + dst_off[h];
            *reinterpret_cast<uint4*>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
          }
        }
      }
      ptx::fence_proxy_async();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&b_full[bi]);
      if (++bi == 2) { bi = 0; bph ^= 1; }
    }
  } else {
    // ===================== epilogue (warps 3..10): lane quarter x 56-column half =====================
    const int quarter = warp & 3;
    const int hf = (warp - 3) >> 2;
    const bool issuer = (warp == 3 && lane == 0);
    int acc = 0, ob = 0; uint32_t accph = 0;
    for (int t = t_begin; t < t_end; ++t) {
      int pc, y0, x0;
      const bool shrt = dw_decode(p, t, pc, y0, x0);
      const int row = shrt ? quarter * 16 + lane : quarter * 32 + lane;   // accumulator lane -> tile row
      const bool live = !shrt || lane < 16;
      float scale, shift;
      bn_fold(p.gamma, p.beta, p.mean, p.var, p.eps, pc % p.C, scale, shift);
      ptx::mbar_wait(&tfull[acc], accph);
      ptx::tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * DW_ACC_COLS + 16 + hf * 56);
      uint32_t v[56];
      ptx::tmem_ld16(t_addr, v);
      ptx::tmem_ld16(t_addr + 16, v + 16);
      ptx::tmem_ld16(t_addr + 32, v + 32);
      ptx::tmem_ld8(t_addr + 48, v + 48);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();                            // accumulator drained: hand it back to the MMA warp
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tempty[acc]);
      if (++acc == DW_ACC_STAGES) { acc = 0; accph ^= 1; }
      uint32_t pk[28];
#pragma unroll
      for (int i = 0; i < 28; ++i) {
        float a = fmaf(__uint_as_float(v[2 * i]), scale, shift);
        float b = fmaf(__uint_as_float(v[2 * i + 1]), scale, shift);
        if (RELU6) {
          pk[i] = pack16_relu6(a, b, F16);
        } else {
          a = apply_act(a, p.act); b = apply_act(b, p.act);
          pk[i] = pack16(a, b, F16);
        }
      }
      if (issuer) ptx::tma_store_wait_read<1>();       // the store that last read staging tile `ob` is done
      ptx::named_bar_sync(1, 32 * DW_EPI_WARPS);
      uint8_t* dst = sO + ob * DW_OUT_BYTES + row * (DW_TW * 2) + hf * 112;
      if (live) {
#pragma unroll
        for (int j = 0; j < 7; ++j)
          *reinterpret_cast<uint4*>(dst + j * 16) = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
      }
      ptx::fence_proxy_async();
      ptx::named_bar_sync(1, 32 * DW_EPI_WARPS);
      if (issuer) {
        ptx::tma_store_3d(&tm_y, sO + ob * DW_OUT_BYTES, x0, y0, pc);
        ptx::tma_store_commit();
      }
      ob ^= 1;
    }
    if (issuer) ptx::tma_store_wait_all<0>();
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem_base, 512); }
}

// ==================================================================================================
// (3) project: planar 16-bit -> NHWC bf16 trunk, folded BN + residual
// ==================================================================================================
struct ProjectParams {
  int N, HW, mid, kcs, f16, yf16, has_res;
  const float* gamma; const float* beta; const float* mean; const float* var; float eps;
  int tiles_per_img;
};

__global__ void __launch_bounds__(PJ_THREADS, 1)
project_planar_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_w,
                      const __grid_constant__ CUtensorMap tm_r, const __grid_constant__ CUtensorMap tm_y,
                      const ProjectParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space
  uint8_t* sA = smem;
  uint8_t* sW = sA + PJ_A_STAGES * PJ_A_BYTES;                     // kcs x 8 KiB resident
  uint8_t* sR = sW + PJ_MAX_KC * 8192;                             // 2 residual / output tiles
  float* s_scale = reinterpret_cast<float*>(sR + 2 * PJ_R_BYTES);
  float* s_shift = s_scale + 64;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(s_shift + 64);
  uint64_t* a_empty = a_full + PJ_A_STAGES;
  uint64_t* tfull = a_empty + PJ_A_STAGES;
  uint64_t* tempty = tfull + PJ_ACC_STAGES;
  uint64_t* r_full = tempty + PJ_ACC_STAGES;
  uint64_t* r_empty = r_full + 2;
  uint64_t* w_bar = r_empty + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(w_bar + 1);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  if (threadIdx.x < 64) {
    float sc, sh;
    bn_fold(p.gamma, p.beta, p.mean, p.var, p.eps, threadIdx.x, sc, sh);
    s_scale[threadIdx.x] = sc;
    s_shift[threadIdx.x] = sh;
  }
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tm_a); ptx::prefetch_tmap(&tm_w); ptx::prefetch_tmap(&tm_r); ptx::prefetch_tmap(&tm_y);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < PJ_A_STAGES; ++s) { ptx::mbar_init(&a_full[s], 1); ptx::mbar_init(&a_empty[s], 1); }
    for (int a = 0; a < PJ_ACC_STAGES; ++a) { ptx::mbar_init(&tfull[a], 1); ptx::mbar_init(&tempty[a], 4); }
    for (int b = 0; b < 2; ++b) { ptx::mbar_init(&r_full[b], 1); ptx::mbar_init(&r_empty[b], 1); }
    ptx::mbar_init(w_bar, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 2) { ptx::tmem_alloc(tmem_ptr, 256); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);
  const int num_tiles = p.N * p.tiles_per_img;

  if (warp == 0) {
    if (lane == 0) {
      pdl_wait();                                   // the planes, the trunk and possibly the packed weights: earlier launches
      ptx::mbar_arrive_expect_tx(w_bar, (uint32_t)(p.kcs * 8192));
      for (int kc = 0; kc < p.kcs; ++kc) ptx::tma_load_3d(sW + kc * 8192, &tm_w, w_bar, kc * 64, 0, 0);
      int s = 0, rb = 0; uint32_t ph = 0, rph = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int n = t / p.tiles_per_img, p0 = (t - n * p.tiles_per_img) * PJ_MPIX;
        for (int kc = 0; kc < p.kcs; ++kc) {
          ptx::mbar_wait(&a_empty[s], ph ^ 1);
          ptx::mbar_arrive_expect_tx(&a_full[s], PJ_A_BYTES);
          ptx::tma_load_3d(sA + s * PJ_A_BYTES, &tm_a, &a_full[s], p0, kc * 64, n);
          ptx::tma_load_3d(sA + s * PJ_A_BYTES + 8192, &tm_a, &a_full[s], p0 + 64, kc * 64, n);
          if (++s == PJ_A_STAGES) { s = 0; ph ^= 1; }
        }
        // The residual / output staging tile AFTER the operand loads: it is freed by the epilogue of the PREVIOUS tile,
        // and waiting for it first (round 1) kept the producer from running ahead -- tile t's loads could only start
        // once tile t-1 was loaded, multiplied and drained: one HBM latency (~2 us) per tile, 8000 clocks per tile.
        ptx::mbar_wait(&r_empty[rb], rph ^ 1);
        if (p.has_res) {
          ptx::mbar_arrive_expect_tx(&r_full[rb], PJ_R_BYTES);
          ptx::tma_load_3d(sR + rb * PJ_R_BYTES, &tm_r, &r_full[rb], 0, p0, n);
        } else {
          ptx::mbar_arrive(&r_full[rb]);
        }
        if (++rb == 2) { rb = 0; rph ^= 1; }
      }
    }
  } else if (warp == 1) {
    const int fmt = p.f16 ? 0 : 1;
    const uint32_t idesc = ptx::umma_idesc_f16(128, 64, fmt, fmt, /*A MN-major*/ 1, 0);
    const uint32_t sA_addr = ptx::smem_u32(sA), sW_addr = ptx::smem_u32(sW);
    int s = 0, acc = 0; uint32_t ph = 0, accph = 0;
    ptx::mbar_wait(w_bar, 0);
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      ptx::mbar_wait(&tempty[acc], accph ^ 1);
      ptx::tc_fence_after();
      const uint32_t d0 = tmem_base + (uint32_t)(acc * 64);
      for (int kc = 0; kc < p.kcs; ++kc) {
        ptx::mbar_wait(&a_full[s], ph);
        ptx::tc_fence_after();
        const uint32_t a0 = sA_addr + (uint32_t)(s * PJ_A_BYTES);
        const uint64_t db = ptx::umma_desc_sw128(sW_addr + (uint32_t)(kc * 8192), 1024);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          // 16 channels = two 8-deep K groups (1024 bytes each); the two 64-pixel atoms are 8192 bytes apart
          const uint64_t da = ptx::umma_desc(a0 + (uint32_t)(k * 2048), 8192, 1024, 2);
          ptx::umma_elect(d0, da, db + (uint64_t)(k * 2), idesc, (uint32_t)(kc | k));
        }
        ptx::umma_commit_elect(&a_empty[s]);
        if (++s == PJ_A_STAGES) { s = 0; ph ^= 1; }
      }
      ptx::umma_commit_elect(&tfull[acc]);
      if (++acc == PJ_ACC_STAGES) { acc = 0; accph ^= 1; }
    }
  } else {
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const bool issuer = (warp == 2 && lane == 0);
    int acc = 0, rb = 0, prev_rb = -1; uint32_t accph = 0, rph = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int n = t / p.tiles_per_img, p0 = (t - n * p.tiles_per_img) * PJ_MPIX;
      ptx::mbar_wait(&tfull[acc], accph);
      ptx::tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * 64);
      uint32_t v[64];
      ptx::tmem_ld16(t_addr, v);
      ptx::tmem_ld16(t_addr + 16, v + 16);
      ptx::tmem_ld16(t_addr + 32, v + 32);
      ptx::tmem_ld16(t_addr + 48, v + 48);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tempty[acc]);
      if (++acc == PJ_ACC_STAGES) { acc = 0; accph ^= 1; }

      ptx::mbar_wait(&r_full[rb], rph);
      uint8_t* tile = sR + rb * PJ_R_BYTES + row * 128;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        uint4* q = reinterpret_cast<uint4*>(tile + ((j ^ (row & 7)) << 4));
        uint32_t rr[4] = {0u, 0u, 0u, 0u};
        if (p.has_res) { const uint4 r4 = *q; rr[0] = r4.x; rr[1] = r4.y; rr[2] = r4.z; rr[3] = r4.w; }
        uint32_t pk[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int ch = j * 8 + 2 * i;
          const float2 r2 = unpack16(rr[i], p.yf16);   // rr == 0 -> (0, 0) in either format
          const float a = fmaf(__uint_as_float(v[ch]), s_scale[ch], s_shift[ch]) + r2.x;
          const float b = fmaf(__uint_as_float(v[ch + 1]), s_scale[ch + 1], s_shift[ch + 1]) + r2.y;
          pk[i] = pack16(a, b, p.yf16);
        }
        *q = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
      ptx::fence_proxy_async();
      ptx::named_bar_sync(1, 128);
      if (issuer) {
        // the previous tile's store must have finished reading its staging tile before the producer refills it
        if (prev_rb >= 0) { ptx::tma_store_wait_read<0>(); ptx::mbar_arrive(&r_empty[prev_rb]); }
        ptx::tma_store_3d(&tm_y, sR + rb * PJ_R_BYTES, 0, p0, n);
        ptx::tma_store_commit();
        prev_rb = rb;
      }
      if (++rb == 2) { rb = 0; rph ^= 1; }
    }
    if (issuer) ptx::tma_store_wait_all<0>();
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem_base, 256); }
}

}  // namespace

// ==================================================================================================
// host side
// ==================================================================================================
bool mbconv_planar_supported(const OfaMBConvArgs* a) {
  const OfaTensor4& x = a->x;
  if (a->cin != 64 || a->cout != 64) return false;
  if (a->mid % 64 != 0 || a->mid < 64 || a->mid > 384) return false;
  if (x.w % 8 != 0 || x.n <= 0 || x.h <= 0 || x.w <= 0) return false;
  if ((long long)x.h * x.w >= (1ll << 31) || (long long)x.n * a->mid >= (1ll << 31)) return false;
  if ((reinterpret_cast<uintptr_t>(x.ptr) & 15) || (reinterpret_cast<uintptr_t>(a->y.ptr) & 15)) return false;
  if (reinterpret_cast<uintptr_t>(a->ws) & 15) return false;
  return true;
}

// OFA_IMPL_AUTO's choice between the planar path and the three NHWC kernels (OFA_IMPL_FAST forces planar).  The planar
// depthwise works per channel PLANE: its Toeplitz filter matrices are rebuilt for every plane (~6 us, hidden behind the
// previous plane's tiles only when a plane has several tiles) and its tiles are 128 (64 for a short tail) x 112
// pixels.  Batches of small images therefore run below the frame rates.  Round 1: 64 x 384 planes of 48 x 48 took
// 1.07 ms per block against ~0.25 ms on the NHWC kernels (rule: >= 8192 pixels).  Round 2 (filters precomputed per
// launch, whole-warp Toeplitz build): 0.33 ms per block at 48 x 48, and the X4 teacher forward on 64 x 48 x 48 planes is
// 7.28 ms planar against 7.55 ms NHWC (tools/bench_teacher_x4.py), while 24 x 24 planes (8 % tile fill) are 2.3x slower
// planar.  Rule: planes of at least 2304 pixels (48 x 48) that fill at least a quarter of their tiles.
bool mbconv_planar_preferred(const OfaMBConvArgs* a) {
  const OfaTensor4& x = a->x;
  const int tail = x.h % DW_TH;
  const long long rows = (long long)(x.h / DW_TH) * DW_TH + (tail == 0 ? 0 : tail <= DW_TH / 2 ? DW_TH / 2 : DW_TH);
  const long long cols = (long long)((x.w + DW_TW - 1) / DW_TW) * DW_TW;
  const long long area = (long long)x.h * x.w;
  return area >= 2304 && 4 * area >= rows * cols;
}

static CUtensorMapDataType dt16(int f16) {
  return f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
}

int launch_pack_block_weights(const float* w_exp, long long e_so, long long e_si, const float* w_proj,
                              long long p_so, long long p_si, int cin, int mid, int cout, int mid_pad, int trunk_f16,
                              int f16, void* wexp_p, void* wproj_p, cudaStream_t st) {
  const int total = mid_pad * cin + cout * mid;
  pack_block_weights_kernel<<<(total + 255) / 256, 256, 0, st>>>(
      w_exp, e_so, e_si, w_proj, p_so, p_si, cin, mid, cout, mid_pad, trunk_f16, f16,
      reinterpret_cast<uint16_t*>(wexp_p), reinterpret_cast<uint16_t*>(wproj_p));
  return check_launch("pack_block_weights_kernel");
}

// x: NHWC bf16 [N,H,W,64];  y: planar [N][mid][H*W] 16-bit
int launch_expand_planar(const void* x, void* y, const void* wexp_p, int N, int HW, int mid, int trunk_f16, int f16,
                         const OfaBn* bn, int act, cudaStream_t st) {
  ExpandParams p;
  memset(&p, 0, sizeof(p));
  p.N = N; p.HW = HW; p.mid = mid; p.mt = (mid + 127) / 128; p.f16 = f16; p.xf16 = trunk_f16; p.act = act;
  p.gamma = bn->gamma; p.beta = bn->beta; p.mean = bn->mean; p.var = bn->var; p.eps = bn->eps;
  p.tiles_per_img = (HW + EX_NPIX - 1) / EX_NPIX;
  CUtensorMap tx, tw, ty;
  int rc;
  {
    uint64_t dims[3] = {64, (uint64_t)HW, (uint64_t)N};
    uint64_t str[2] = {128, (uint64_t)HW * 128};
    uint32_t box[3] = {64, EX_NPIX, 1};
    if ((rc = encode_tmap(&tx, dt16(trunk_f16), 3, const_cast<void*>(x), dims, str, box,
                          CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  }
  {
    uint64_t dims[3] = {64, (uint64_t)p.mt * 128, 1};
    uint64_t str[2] = {128, (uint64_t)p.mt * 128 * 128};
    uint32_t box[3] = {64, 128, 1};
    if ((rc = encode_tmap(&tw, dt16(trunk_f16), 3, const_cast<void*>(wexp_p), dims, str, box,
                          CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)HW, (uint64_t)mid, (uint64_t)N};
    uint64_t str[2] = {(uint64_t)HW * 2, (uint64_t)HW * mid * 2};
    uint32_t box[3] = {64, 32, 1};
    if ((rc = encode_tmap(&ty, dt16(f16), 3, y, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  }
  const size_t smem = 1024 + EX_MAX_MT * 16384 + EX_X_STAGES * EX_X_BYTES + EX_EPI_WARPS * EX_SBUFS * EX_SBUF_BYTES + 256;
  int grid = sm_count();
  const int tiles = N * p.tiles_per_img;
  if (grid > tiles) grid = tiles;
  const int relu6 = act == OFA_ACT_RELU6 ? 1 : 0;
#define OFA_EX_LAUNCH(F16_, R6_)                                                                                  \
  do {                                                                                                            \
    OFA_CUDA(cudaFuncSetAttribute(expand_planar_kernel<F16_, R6_>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                  (int)smem));                                                                    \
    launch_pdl(expand_planar_kernel<F16_, R6_>, dim3(grid), dim3(EX_THREADS), smem, st, tx, tw, ty, p);                               \
  } while (0)
  if (f16 && relu6) OFA_EX_LAUNCH(1, 1);
  else if (f16) OFA_EX_LAUNCH(1, 0);
  else if (relu6) OFA_EX_LAUNCH(0, 1);
  else OFA_EX_LAUNCH(0, 0);
#undef OFA_EX_LAUNCH
  return check_launch("expand_planar_kernel");
}

// x, y: planar [N*C][H][W] 16-bit
int launch_dw_planar(const void* x, void* y, int N, int C, int H, int W, const float* w7, int kmax, const float* m75,
                     const float* m53, int transform_on, int ks, int f16, const OfaBn* bn, int act, cudaStream_t st) {
  DwPlanarParams p;
  memset(&p, 0, sizeof(p));
  p.NC = N * C; p.C = C; p.H = H; p.W = W; p.ks = ks; p.kmax = kmax; p.transform_on = transform_on; p.f16 = f16;
  p.act = act;
  // the transformed filters (centre crop + learned 7->5->3 matrices, dynamic_op.py:46-71) once per launch
  // ... unless the active kernel size IS the stored one: then the [C][ks * ks] active filter is the parameter itself
  // (the max sub-network of the headline frame: one 5 us launch per block less)
  float* filt = nullptr;
  p.filt_from_kernel = ks == kmax ? 0 : 1;
  if (ks == kmax) {
    p.filt = w7;
  } else {
    keep_async_pool_resident();
    OFA_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&filt), (size_t)C * ks * ks * sizeof(float), st));
    const int rc0 = launch_active_filter(w7, kmax, m75, m53, transform_on, ks, C, filt, st);
    if (rc0) { cudaFreeAsync(filt, st); return rc0; }
    p.filt = filt;
  }
  if (bn) { p.gamma = bn->gamma; p.beta = bn->beta; p.mean = bn->mean; p.var = bn->var; p.eps = bn->eps; }
  p.tiles_x = (W + DW_TW - 1) / DW_TW;
  p.tiles_y = (H + DW_TH - 1) / DW_TH;
  const long long total = (long long)p.NC * p.tiles_x * p.tiles_y;
  if (total >= (1ll << 31)) return fail(OFA_ERR_UNSUPPORTED, "dw_planar: too many tiles");
  p.total_tiles = (int)total;
  const int rem = H % DW_TH;
  p.short_last = (rem > 0 && rem <= DW_TH / 2) ? 1 : 0;
  CUtensorMap tx, txs, ty;
  int rc;
  uint64_t dims[3] = {(uint64_t)W, (uint64_t)H, (uint64_t)p.NC};
  uint64_t str[2] = {(uint64_t)W * 2, (uint64_t)H * W * 2};
  {
    uint32_t box[3] = {64, (uint32_t)(DW_TH + ks - 1), 1};
    if ((rc = encode_tmap(&tx, dt16(f16), 3, const_cast<void*>(x), dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)))
      return rc;
    box[1] = (uint32_t)(DW_TH / 2 + ks - 1);
    if ((rc = encode_tmap(&txs, dt16(f16), 3, const_cast<void*>(x), dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)))
      return rc;
  }
  {
    uint32_t box[3] = {DW_TW, DW_TH, 1};
    if ((rc = encode_tmap(&ty, dt16(f16), 3, y, dims, str, box, CU_TENSOR_MAP_SWIZZLE_NONE))) return rc;
  }
  const size_t smem = 1024 + DWP_A_STAGES * DW_A_STRIDE + 2 * DW_OUT_BYTES + 2 * DW_B_BYTES + 1024 + 256;
  long long grid = sm_count();
  if (grid > p.total_tiles) grid = p.total_tiles;
  const int relu6 = act == OFA_ACT_RELU6 ? 1 : 0;
#define OFA_DW_LAUNCH(KS_, F16_, R6_)                                                                            \
  do {                                                                                                           \
    OFA_CUDA(cudaFuncSetAttribute(dw_planar_kernel<KS_, F16_, R6_>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                  (int)smem));                                                                   \
    launch_pdl(dw_planar_kernel<KS_, F16_, R6_>, dim3((unsigned)grid), dim3(DW_THREADS), smem, st, tx, txs, ty, p);                      \
  } while (0)
#define OFA_DW_LAUNCH_KS(KS_)                                           \
  do {                                                                  \
    if (f16 && relu6) OFA_DW_LAUNCH(KS_, 1, 1);                         \
    else if (f16) OFA_DW_LAUNCH(KS_, 1, 0);                             \
    else if (relu6) OFA_DW_LAUNCH(KS_, 0, 1);                           \
    else OFA_DW_LAUNCH(KS_, 0, 0);                                      \
  } while (0)
  if (ks == 3) OFA_DW_LAUNCH_KS(3);
  else if (ks == 5) OFA_DW_LAUNCH_KS(5);
  else OFA_DW_LAUNCH_KS(7);
#undef OFA_DW_LAUNCH_KS
#undef OFA_DW_LAUNCH
  const int rc_launch = check_launch("dw_planar_kernel");
  if (filt) cudaFreeAsync(filt, st);
  return rc_launch;
}

// x: planar [N][mid][H*W] 16-bit; res / y: NHWC bf16 [N,H*W,64]
int launch_project_planar(const void* x, const void* res, void* y, const void* wproj_p, int N, int HW, int mid,
                          int trunk_f16, int f16, const OfaBn* bn, cudaStream_t st) {
  ProjectParams p;
  memset(&p, 0, sizeof(p));
  p.N = N; p.HW = HW; p.mid = mid; p.kcs = mid / 64; p.f16 = f16; p.yf16 = trunk_f16; p.has_res = res ? 1 : 0;
  p.gamma = bn->gamma; p.beta = bn->beta; p.mean = bn->mean; p.var = bn->var; p.eps = bn->eps;
  p.tiles_per_img = (HW + PJ_MPIX - 1) / PJ_MPIX;
  CUtensorMap ta, tw, tr, ty;
  int rc;
  {
    uint64_t dims[3] = {(uint64_t)HW, (uint64_t)mid, (uint64_t)N};
    uint64_t str[2] = {(uint64_t)HW * 2, (uint64_t)HW * mid * 2};
    uint32_t box[3] = {64, 64, 1};
    if ((rc = encode_tmap(&ta, dt16(f16), 3, const_cast<void*>(x), dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)))
      return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)mid, 64, 1};
    uint64_t str[2] = {(uint64_t)mid * 2, (uint64_t)mid * 64 * 2};
    uint32_t box[3] = {64, 64, 1};
    if ((rc = encode_tmap(&tw, dt16(f16), 3, const_cast<void*>(wproj_p), dims, str, box,
                          CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  }
  {
    uint64_t dims[3] = {64, (uint64_t)HW, (uint64_t)N};
    uint64_t str[2] = {128, (uint64_t)HW * 128};
    uint32_t box[3] = {64, PJ_MPIX, 1};
    if ((rc = encode_tmap(&tr, dt16(trunk_f16), 3, const_cast<void*>(res ? res : y), dims, str, box,
                          CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    if ((rc = encode_tmap(&ty, dt16(trunk_f16), 3, y, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  }
  const size_t smem = 1024 + PJ_A_STAGES * PJ_A_BYTES + PJ_MAX_KC * 8192 + 2 * PJ_R_BYTES + 128 * 4 + 256;
  OFA_CUDA(cudaFuncSetAttribute(project_planar_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int grid = sm_count();
  const int tiles = N * p.tiles_per_img;
  if (grid > tiles) grid = tiles;
  launch_pdl(project_planar_kernel, dim3(grid), dim3(PJ_THREADS), smem, st, ta, tw, tr, ty, p);
  return check_launch("project_planar_kernel");
}

}  // namespace ofa
