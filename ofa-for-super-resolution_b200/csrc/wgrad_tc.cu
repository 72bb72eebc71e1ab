// Weight gradient of the dense convs (autograd of dynamic_op.py:104-112 and layers.py:135-147) on tcgen05:
//
//     dW[co][ci][ky][kx] += sum_{n,y,x} dY[n,y,x,co] * X[n, y+ky-R, x+kx-R, ci]
//
// is, per filter tap, a GEMM whose K dimension is the PIXELS.  Both operands are NHWC 16-bit tensors, i.e.
// "pixel rows of contiguous channels": exactly MN-major UMMA operands as TMA lands them (one 128-byte
// swizzle atom = 64 channels x 8 pixels).  One side of every conv of the SR nets has 64 channels (the trunk):
// that side is N (one atom), the other side (64 ... 384 channels, padded to 128-row tiles by TMA zero fill)
// is M, so D[m-channel lane][64 columns] is one (tap, M-tile) "unit" of 64 TMEM columns.
//   * a pixel tile is 16 rows x 8 columns; a K = 16 step is two tile rows of 8 pixels (descriptor SBO =
//     the row pitch of each operand's box, so the activation operand may be a halo box);
//   * filter taps are shifted descriptor starts into ONE halo box of X (as in conv_tc);
//   * a CTA owns up to 8 units (512 TMEM columns) and a contiguous range of pixel tiles; it accumulates in
//     TMEM over its whole range and adds its fp32 partial into the strided weight-gradient slice with
//     red.global.add (grid = unit groups x pixel splits);
//   * "swap" mode (cout = 64: the 1x1 project convs, the stem): M = input channels from X, N = output channels
//     from dY, and the tap shift is applied to dY (with the opposite sign) instead of X.
#include "ofa_common.cuh"
#include "kernels.h"
#include "sm100_ptx.cuh"

#include <string.h>

namespace ofa {
namespace {

constexpr int WG_TH = 16, WG_TW = 8;               // pixel tile
constexpr int WG_ATOM_BYTES = WG_TH * WG_TW * 128;  // 16 KiB: 64 channels x 128 pixels
constexpr int WG_M_SLOTS = 4;                       // up to 2 M-tiles x 2 atoms per stage
constexpr int WG_N_BYTES = 32 * 1024;               // halo box of the 64-channel side (ks <= 5: 20 x 12 x 128 B)
constexpr int WG_STAGE_BYTES = WG_M_SLOTS * WG_ATOM_BYTES + WG_N_BYTES;
constexpr int WG_STAGES = 2;
constexpr int WG_MAX_UNITS = 8;
constexpr int WG_THREADS = 6 * 32;                  // TMA, MMA, 4 epilogue warps

struct WgradParams {
  int N, H, W, ks, f16, swap;
  int m_ch;                 // channels on the M side (cout, or cin in swap mode)
  int mt_total;             // ceil(m_ch / 128)
  int units_total;          // taps * mt_total
  int groups, splits;
  int tiles_h, tiles_w, total_tiles;
  float* dw; long long w_so, w_si, w_sh, w_sw;
};

// unit u -> (tap, mt); a group holds units [g*gs, g*gs+gs) where gs is chosen so that a group spans <= 2 M-tiles
__device__ __forceinline__ void wg_unit(const WgradParams& p, int u, int& tap, int& mt) {
  tap = u / p.mt_total;
  mt = u - tap * p.mt_total;
}

__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tm_m, const __grid_constant__ CUtensorMap tm_n,
                const WgradParams p, const int group_size) {
  pdl_wait();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + WG_STAGES * WG_STAGE_BYTES);
  uint64_t* empty = full + WG_STAGES;
  uint64_t* done = empty + WG_STAGES;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(done + 1);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) { ptx::prefetch_tmap(&tm_m); ptx::prefetch_tmap(&tm_n); }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < WG_STAGES; ++s) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], 1); }
    ptx::mbar_init(done, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 2) { ptx::tmem_alloc(tmem_ptr, 512); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);

  const int g = blockIdx.x % p.groups, split = blockIdx.x / p.groups;
  const int u0 = g * group_size;
  const int nu = min(group_size, p.units_total - u0);
  int tap0, mt0, tapl, mtl;
  wg_unit(p, u0, tap0, mt0);
  wg_unit(p, u0 + nu - 1, tapl, mtl);
  // M-tiles touched by this group: a contiguous range when the group lies inside one tap, else all of them
  const int mt_lo = (tap0 == tapl) ? mt0 : 0;
  const int mt_hi = (tap0 == tapl) ? mtl : p.mt_total - 1;
  const int n_mt = mt_hi - mt_lo + 1;                       // <= 2 by construction of group_size
  const int t_begin = (int)((long long)p.total_tiles * split / p.splits);
  const int t_end = (int)((long long)p.total_tiles * (split + 1) / p.splits);
  const int R = p.ks >> 1;
  const int halo_w = WG_TW + 2 * R, halo_h = WG_TH + 2 * R;
  const uint32_t n_bytes = (uint32_t)(halo_w * halo_h * 128);
  const int per_img = p.tiles_h * p.tiles_w;

  if (warp == 0) {
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      for (int t = t_begin; t < t_end; ++t) {
        const int n = t / per_img, r = t - n * per_img;
        const int y0 = (r / p.tiles_w) * WG_TH, x0 = (r % p.tiles_w) * WG_TW;
        uint8_t* st = smem + s * WG_STAGE_BYTES;
        ptx::mbar_wait(&empty[s], ph ^ 1);
        ptx::mbar_arrive_expect_tx(&full[s], (uint32_t)(n_mt * 2 * WG_ATOM_BYTES) + n_bytes);
        for (int i = 0; i < n_mt; ++i)
          for (int a = 0; a < 2; ++a)
            ptx::tma_load_4d(st + (i * 2 + a) * WG_ATOM_BYTES, &tm_m, &full[s], (mt_lo + i) * 128 + a * 64, x0, y0, n);
        ptx::tma_load_4d(st + WG_M_SLOTS * WG_ATOM_BYTES, &tm_n, &full[s], 0, x0 - R, y0 - R, n);
        if (++s == WG_STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    const int fmt = p.f16 ? 0 : 1;
    const uint32_t idesc = ptx::umma_idesc_f16(128, 64, fmt, fmt, /*A MN-major*/ 1, /*B MN-major*/ 1);
    const uint32_t smem_addr = ptx::smem_u32(smem);
    int s = 0; uint32_t ph = 0;
    uint32_t acc_flag = 0;
    for (int t = t_begin; t < t_end; ++t) {
      ptx::mbar_wait(&full[s], ph);
      ptx::tc_fence_after();
      const uint32_t st = smem_addr + (uint32_t)(s * WG_STAGE_BYTES);
      const uint32_t nb = st + (uint32_t)(WG_M_SLOTS * WG_ATOM_BYTES);
      for (int i = 0; i < nu; ++i) {
        int tap, mt;
        wg_unit(p, u0 + i, tap, mt);
        const int ky = tap / p.ks, kx = tap - ky * p.ks;
        const uint32_t a_base = st + (uint32_t)((mt - mt_lo) * 2 * WG_ATOM_BYTES);
#pragma unroll
        for (int j = 0; j < WG_TH / 2; ++j) {
          // K = 16 pixels: tile rows 2j and 2j+1 (8 pixels each); M side: dense box, rows 1024 B apart, the
          // two 64-channel atoms 16 KiB apart; N side: halo box, rows halo_w * 128 B apart, shifted by the tap
          const uint64_t da = ptx::umma_desc(a_base + (uint32_t)(j * 2 * WG_TW * 128), WG_ATOM_BYTES, WG_TW * 128, 2);
          const uint64_t db = ptx::umma_desc(nb + (uint32_t)(((2 * j + ky) * halo_w + kx) * 128), 0,
                                             (uint32_t)(halo_w * 128), 2);
          ptx::umma_elect(tmem_base + (uint32_t)(i * 64), da, db, idesc, acc_flag | (uint32_t)j);
        }
      }
      acc_flag = 1;
      ptx::umma_commit_elect(&empty[s]);
      if (++s == WG_STAGES) { s = 0; ph ^= 1; }
    }
    ptx::umma_commit_elect(done);
  } else if (t_begin < t_end) {
    // ===================== epilogue: add this CTA's partial into the weight-gradient slice =====================
    const int quarter = warp & 3;
    const int m_local = quarter * 32 + lane;
    const bool vec4 = !p.swap && p.ks == 1 && p.w_si == 1 && (p.w_so & 3) == 0 &&
                      (reinterpret_cast<unsigned long long>(p.dw) & 15ull) == 0;
    ptx::mbar_wait(done, 0);
    ptx::tc_fence_after();
    for (int i = 0; i < nu; ++i) {
      int tap, mt;
      wg_unit(p, u0 + i, tap, mt);
      const int ky = tap / p.ks, kx = tap - ky * p.ks;
      const int mch = mt * 128 + m_local;
      const uint32_t t_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(i * 64);
#pragma unroll 1
      for (int c0 = 0; c0 < 64; c0 += 16) {
        uint32_t v[16];
        ptx::tmem_ld16(t_addr + (uint32_t)c0, v);
        ptx::tmem_ld_wait();
        if (mch < p.m_ch && vec4) {
          // 1x1 convs, M = cout: a thread's 64 columns are 64 consecutive input channels of one dW row -- four
          // 16-byte vector reductions per 16 columns instead of sixteen scalar ones (the scalar form is 32 separate
          // L2 atomic transactions per warp instruction and held ~45 % of this kernel's time)
          float* row = p.dw + (long long)mch * p.w_so + c0;
#pragma unroll
          for (int c = 0; c < 16; c += 4)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(row + c), "f"(__uint_as_float(v[c])),
                         "f"(__uint_as_float(v[c + 1])), "f"(__uint_as_float(v[c + 2])), "f"(__uint_as_float(v[c + 3]))
                         : "memory");
        } else if (mch < p.m_ch) {
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            const int nch = c0 + c;
            // swap mode shifts dY instead of X: dW[..][a][b] pairs X[q] with dY[q - (a-R, b-R)], i.e. the unit's tap
            // (ky, kx) is the gradient of the filter tap (ks-1-ky, ks-1-kx)
            const int co = p.swap ? nch : mch, ci = p.swap ? mch : nch;
            const int oy = p.swap ? p.ks - 1 - ky : ky, ox = p.swap ? p.ks - 1 - kx : kx;
            atomicAdd(p.dw + co * p.w_so + ci * p.w_si + oy * p.w_sh + ox * p.w_sw, __uint_as_float(v[c]));
          }
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem_base, 512); }
}

}  // namespace

bool wgrad_tc_supported(const OfaTensor4* x, const OfaTensor4* dy, int cin, int cout, int ks) {
  if (!is_16bit(x->dtype) || dy->dtype != x->dtype) return false;
  if (!is_nhwc_dense(x) || !is_nhwc_dense(dy)) return false;
  if ((reinterpret_cast<uintptr_t>(x->ptr) & 15) || (reinterpret_cast<uintptr_t>(dy->ptr) & 15)) return false;
  if (x->n <= 0 || x->h <= 0 || x->w <= 0) return false;
  if (ks != 1 && ks != 3 && ks != 5) return false;
  if (cin == 64 && cout % 8 == 0 && cout >= 8 && cout <= 384) return true;      // M = cout (dY), N = cin (X halo)
  if (cout == 64 && cin % 8 == 0 && cin >= 8 && cin <= 384 && (ks == 1 || cin <= 256)) return true;   // swap: M = cin (X), N = cout (dY halo)
  return false;
}

int launch_wgrad_tc(const OfaTensor4* x, const OfaTensor4* dy, float* dw, long long w_so, long long w_si,
                    long long w_sh, long long w_sw, int cin, int cout, int ks, cudaStream_t st) {
  WgradParams p;
  memset(&p, 0, sizeof(p));
  p.N = x->n; p.H = x->h; p.W = x->w; p.ks = ks;
  if (ks == 1 && (p.H % WG_TH != 0 || p.W % WG_TW != 0)) {
    // a 1x1 conv has no spatial structure and the pixels are only the reduction dimension: re-view the dense tensors
    // as one image of P / 8 rows by 8 pixels, which the 16 x 8 pixel tiles cover exactly (24 x 24 patches: 6 tiles
    // of which a quarter is padding -> 4.5)
    const long long P = (long long)x->n * x->h * x->w;
    if (P % WG_TW == 0 && P / WG_TW < (1ll << 31)) { p.N = 1; p.H = (int)(P / WG_TW); p.W = WG_TW; }
  }
  p.f16 = x->dtype == OFA_F16 ? 1 : 0;
  p.swap = (cin == 64) ? 0 : 1;
  p.m_ch = p.swap ? cin : cout;
  p.mt_total = (p.m_ch + 127) / 128;
  p.units_total = ks * ks * p.mt_total;
  // group size: up to 8 units, spanning at most 2 M-tiles (the M side of a stage holds 2 M-tiles)
  int gs;
  if (p.mt_total <= 2) gs = WG_MAX_UNITS / p.mt_total * p.mt_total;   // whole taps: 8 (mt 1), 8 (mt 2)
  else gs = 2;                                                        // mt_total == 3 (1x1 only in these nets)
  if (p.mt_total > 2 && ks != 1) return fail(OFA_ERR_UNSUPPORTED, "wgrad_tc: k x k with more than 256 M channels");
  if (gs > p.units_total) gs = p.units_total;
  p.groups = (p.units_total + gs - 1) / gs;
  p.tiles_h = (p.H + WG_TH - 1) / WG_TH;
  p.tiles_w = (p.W + WG_TW - 1) / WG_TW;
  const long long tiles = (long long)p.N * p.tiles_h * p.tiles_w;
  if (tiles >= (1ll << 30)) return fail(OFA_ERR_UNSUPPORTED, "wgrad_tc: too many tiles");
  p.total_tiles = (int)tiles;
  int splits = sm_count() / p.groups;
  if (splits < 1) splits = 1;
  if (splits > p.total_tiles) splits = p.total_tiles;
  p.splits = splits;
  p.dw = dw; p.w_so = w_so; p.w_si = w_si; p.w_sh = w_sh; p.w_sw = w_sw;

  const OfaTensor4* tm_src = p.swap ? x : dy;     // M side
  const OfaTensor4* tn_src = p.swap ? dy : x;     // N side (64 channels)
  const CUtensorMapDataType dt = p.f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  const int R = ks / 2;
  CUtensorMap tmm, tmn;
  int rc;
  {
    const uint64_t C = (uint64_t)tm_src->c;
    uint64_t dims[4] = {C, (uint64_t)p.W, (uint64_t)p.H, (uint64_t)p.N};
    uint64_t str[3] = {C * 2, (uint64_t)p.W * C * 2, (uint64_t)p.H * p.W * C * 2};
    uint32_t box[4] = {64, WG_TW, WG_TH, 1};
    if ((rc = encode_tmap(&tmm, dt, 4, tm_src->ptr, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  }
  {
    uint64_t dims[4] = {64, (uint64_t)p.W, (uint64_t)p.H, (uint64_t)p.N};
    uint64_t str[3] = {128, (uint64_t)p.W * 128, (uint64_t)p.H * p.W * 128};
    uint32_t box[4] = {64, (uint32_t)(WG_TW + 2 * R), (uint32_t)(WG_TH + 2 * R), 1};
    if ((rc = encode_tmap(&tmn, dt, 4, tn_src->ptr, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  }
  const size_t smem = 1024 + WG_STAGES * WG_STAGE_BYTES + 128;
  static unsigned char attr_done[64] = {0};
  if (once_per_device(attr_done))
    OFA_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  launch_pdl(wgrad_tc_kernel, dim3(p.groups * p.splits), dim3(WG_THREADS), smem, st, tmm, tmn, p, gs);
  return check_launch("wgrad_tc_kernel");
}

}  // namespace ofa
