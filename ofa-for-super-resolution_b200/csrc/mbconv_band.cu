// Inference MBConv block (dynamic_layers.py:70-84 + proxyless_nets.py:44-51) as ONE launch whose expanded
// intermediates never reach HBM.
//
// The three stages of mbconv_planar.cu -- expand (tcgen05 GEMM, weights as A), banded-Toeplitz depthwise (tcgen05),
// project (tcgen05 GEMM, planar tensor as MN-major A) -- run CONCURRENTLY as role-specialised persistent CTAs of one
// grid (blockIdx < nE: expand, < nE + nD: depthwise, else project).  The frame is cut into REGIONS of 128 rows x
// (112 * RT) columns (RT depthwise tiles wide).  Region r flows
//
//     trunk x --E--> t1 ring slot r % S --D--> t2 ring slot r % S --P--> trunk y (+ x)
//
// through two small ring buffers in global memory (S slots of one region x all M channels each: ~13 + 11 MB per slot
// at RT = 1, M = 384).  The rings are re-written every S regions, so their lines are overwritten while still dirty in
// the 126 MB L2 and (almost) never written back: per block the DRAM traffic is the trunk in + out instead of
// 1.79 GB (measured: profiles/r2_ncu_mbconv_band_*.csv).  Measured L2 bandwidth through the TMA path on this part
// (profiles/r2_l2_bw_probe_tma.txt): 16.8 TB/s reads / 7.7 TB/s writes while the working set stays below ~64-96 MB,
// against 7.2 / 6.3 TB/s from HBM.
//
// Stage hand-over is by region: device-scope counters e_done / d_done / p_done[r] count finished E tile-warps, D tiles
// and P tiles; a consumer's TMA producer lane spins (acquire) until the count reaches the region's total, then issues
// its loads.  Ring slots are recycled the same way (E(r) waits for D(r - S), D(r) waits for P(r - S)).  Every CTA of
// the grid must be resident (cooperative launch guarantees it), dependencies only point to lower region numbers, so
// the schedule cannot deadlock; all spins are bounded and trap.
//
// The expand stage recomputes the depthwise halo (ks - 1 rows, 16 columns per 112) and writes ZEROS for halo pixels
// outside the image: that is the depthwise conv's zero padding (the stand-alone kernel gets it from TMA's
// out-of-bounds fill on an image-shaped tensor map; a ring slot has no image shape).
#include "ofa_common.cuh"
#include "kernels.h"
#include "sm100_ptx.cuh"
#include "mbconv_planar.cuh"

#include <stdlib.h>
#include <string.h>

namespace ofa {
namespace {

constexpr int BD_THREADS = 352;          // 11 warps: TMA producer, MMA issuer, filter builder, 8 epilogue warps
constexpr int BD_T1H = 136;              // ring slot rows of t1 (128 + 6 halo rows, rounded up to the 4-row E tiles)
constexpr int BD_BAR_OFF = 222720;       // barrier block at a fixed offset behind the largest role's buffers
constexpr int BD_SMEM = 1024 + BD_BAR_OFF + 512;
constexpr int BD_PJ_STAGES = 8;          // project role: operand stages (the stand-alone kernel has 9; here D's buffers set the size)

struct BandParams {
  int N, H, W, mid, mt, kcs, xf16;
  int RT, rw;                            // region width in depthwise tiles / in columns
  int n_bands, n_cgs, n_regions, S;
  int nE, nD, nP;
  int short_last;                        // the last band has <= 64 rows: its depthwise tiles run as M = 64 tiles
  int has_res;
  int part_px;                           // rw % 64: width of the narrower y store of an interior region's last box
  const float* filt;                     // [mid][KS * KS] active depthwise filters (fp32)
  const float *g1, *b1, *m1, *v1; float eps1;
  const float *g2, *b2, *m2, *v2; float eps2;
  const float *g3, *b3, *m3, *v3; float eps3;
  unsigned* e_done; unsigned* d_done; unsigned* p_done;
  unsigned long long* stats;             // optional [grid][4]: total clocks, clocks spent in region hand-over waits (x2), units
};

struct Region {
  int n, y0, gx0, rows, cols, tiles_x, shrt, slot;
  int e_tr, e_tc, e_tiles;               // expand: 4-row x 64-column tiles covering the slot area D reads
  int d_tiles;                           // depthwise: mid * tiles_x
  int bpr, p_tiles;                      // project: 64-pixel boxes per region row; tiles of two boxes
};

template <int KS>
__device__ __forceinline__ Region region_info(const BandParams& p, int r) {
  Region g;
  const int per_img = p.n_bands * p.n_cgs;
  g.n = r / per_img;
  const int rem = r - g.n * per_img;
  const int band = rem / p.n_cgs, cg = rem - band * p.n_cgs;
  g.y0 = band * DW_TH;
  g.gx0 = cg * p.rw;
  g.rows = min(DW_TH, p.H - g.y0);
  g.cols = min(p.rw, p.W - g.gx0);
  g.tiles_x = (g.cols + DW_TW - 1) / DW_TW;
  g.shrt = (p.short_last && band == p.n_bands - 1) ? 1 : 0;
  g.slot = r % p.S;
  g.e_tr = ((g.shrt ? DW_TH / 2 : DW_TH) + KS - 1 + 3) >> 2;
  g.e_tc = 2 * g.tiles_x;
  g.e_tiles = g.e_tr * g.e_tc;
  g.d_tiles = p.mid * g.tiles_x;
  g.bpr = (g.cols + 63) >> 6;
  g.p_tiles = (g.rows * g.bpr + 1) >> 1;
  return g;
}

// first unit of a region that CTA `rank` (of `n` CTAs of its role) owns when units are dealt round-robin across regions
__device__ __forceinline__ int first_unit(long long base, int rank, int n) {
  const int off = (int)(base % n);
  return rank >= off ? rank - off : rank - off + n;
}

__device__ __forceinline__ long long wait_count(const unsigned* flag, unsigned expect) {
  unsigned v;
  uint32_t spins = 0;
  const long long t0 = clock64();
  for (;;) {
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
    if (v >= expect) break;
    __nanosleep(100);
    if (++spins > (1u << 24)) {
      printf("ofa: mbconv band kernel: region hand-over timed out (block %d thread %d: %u < %u)\n", (int)blockIdx.x,
             (int)threadIdx.x, v, expect);
      __trap();
    }
  }
  asm volatile("fence.proxy.async;" ::: "memory");      // the TMA loads that follow must observe the producer's stores
  return clock64() - t0;
}
// all TMA stores this thread issued have completed (caller waited on its bulk groups): publish `n` finished units
__device__ __forceinline__ void signal_count(unsigned* flag, unsigned n) {
  asm volatile("fence.proxy.async;" ::: "memory");
  __threadfence();
  atomicAdd(flag, n);
}

__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(ptx::smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}

// ==================================================================================================
// role E: expand 64 -> mid on region + halo, folded BN + ReLU6, zero outside the image, -> t1 ring
// ==================================================================================================
template <int KS, int F16>
__device__ __forceinline__ void role_expand(uint8_t* smem, uint64_t* bars, uint32_t tmem_base, int warp, int lane,
                                            int rank, const CUtensorMap* tm_x, const CUtensorMap* tm_w,
                                            const CUtensorMap* tm_t1s, const BandParams& p) {
  constexpr int R = KS >> 1;
  uint8_t* sW = smem;
  uint8_t* sX = sW + EX_MAX_MT * 16384;
  uint8_t* sS = sX + EX_X_STAGES * EX_X_BYTES;
  uint64_t* x_full = bars;
  uint64_t* x_empty = x_full + EX_X_STAGES;
  uint64_t* tfull = x_empty + EX_X_STAGES;
  uint64_t* tempty = tfull + 2;
  uint64_t* w_bar = tempty + 2;

  if (warp == 0) {
    if (lane == 0) {
      ptx::mbar_arrive_expect_tx(w_bar, (uint32_t)(p.mt * 16384));
      for (int m = 0; m < p.mt; ++m) ptx::tma_load_3d(sW + m * 16384, tm_w, w_bar, 0, m * 128, 0);
      int s = 0; uint32_t ph = 0;
      long long base = 0;
      for (int r = 0; r < p.n_regions; ++r) {
        const Region g = region_info<KS>(p, r);
        for (int u = first_unit(base, rank, p.nE); u < g.e_tiles; u += p.nE) {
          const int tr = u / g.e_tc, tc = u - tr * g.e_tc;
          ptx::mbar_wait(&x_empty[s], ph ^ 1);
          ptx::mbar_arrive_expect_tx(&x_full[s], EX_X_BYTES);
          ptx::tma_load_4d(sX + s * EX_X_BYTES, tm_x, &x_full[s], 0, g.gx0 - DW_XPAD + 64 * tc, g.y0 - R + 4 * tr, g.n);
          if (++s == EX_X_STAGES) { s = 0; ph ^= 1; }
        }
        base += g.e_tiles;
      }
    }
  } else if (warp == 1) {
    const int xfmt = p.xf16 ? 0 : 1;
    const uint32_t idesc = ptx::umma_idesc_f16(128, EX_NPIX, xfmt, xfmt, 0, 0);
    const uint32_t sW_addr = ptx::smem_u32(sW), sX_addr = ptx::smem_u32(sX);
    int s = 0, acc = 0; uint32_t ph = 0, accph = 0;
    ptx::mbar_wait(w_bar, 0);
    long long base = 0;
    for (int r = 0; r < p.n_regions; ++r) {
      const Region g = region_info<KS>(p, r);
      for (int u = first_unit(base, rank, p.nE); u < g.e_tiles; u += p.nE) {
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
      base += g.e_tiles;
    }
  } else if (warp >= 3) {
    const int ew = warp - 3;
    const int quarter = warp & 3;
    const int half = ew >> 2;                       // which 128 accumulator columns = tile rows 2 * half, 2 * half + 1
    uint8_t* sbuf = sS + ew * EX_SBUFS * EX_SBUF_BYTES;
    int acc = 0, sb = 0; uint32_t accph = 0;
    long long base = 0, waited = 0, units = 0;
    for (int r = 0; r < p.n_regions; ++r) {
      const Region g = region_info<KS>(p, r);
      if (r >= p.S) {                               // ring slot reuse: the depthwise stage is done with region r - S
        if (lane == 0) waited += wait_count(&p.d_done[r - p.S], (unsigned)region_info<KS>(p, r - p.S).d_tiles);
        __syncwarp();
      }
      unsigned cnt = 0;
      for (int u = first_unit(base, rank, p.nE); u < g.e_tiles; u += p.nE) {
        const int tr = u / g.e_tc, tc = u - tr * g.e_tc;
        const int ix0 = g.gx0 - DW_XPAD + 64 * tc;  // image column of the tile's first pixel
        // columns of this 64-pixel row segment that lie inside the image
        const int lo = max(0, -ix0), hi = min(64, p.W - ix0);
        for (int m = 0; m < p.mt; ++m) {
          const int c_warp = m * 128 + quarter * 32;
          const int c = c_warp + lane;
          float scale = 0.f, shift = 0.f;
          if (c < p.mid) bn_fold(p.g1, p.b1, p.m1, p.v1, p.eps1, c, scale, shift);
          ptx::mbar_wait(&tfull[acc], accph);
          ptx::tc_fence_after();
          const uint32_t t_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * 256 + half * 128);
          if (c_warp < p.mid) {                     // warp-uniform
#pragma unroll 1
            for (int bx = 0; bx < 2; ++bx) {
              const int sr = 4 * tr + 2 * half + bx;            // slot row
              const int iy = g.y0 - R + sr;
              const bool row_in = iy >= 0 && iy < p.H && hi > lo;
              uint32_t v[64];
              if (row_in) {
                ptx::tmem_ld16(t_addr + (uint32_t)(bx * 64), v);
                ptx::tmem_ld16(t_addr + (uint32_t)(bx * 64 + 16), v + 16);
                ptx::tmem_ld16(t_addr + (uint32_t)(bx * 64 + 32), v + 32);
                ptx::tmem_ld16(t_addr + (uint32_t)(bx * 64 + 48), v + 48);
                ptx::tmem_ld_wait();
              }
              if (lane == 0) ptx::tma_store_wait_read<EX_SBUFS - 1>();   // staging buffer `sb` is free again
              __syncwarp();
              uint8_t* dst = sbuf + sb * EX_SBUF_BYTES + lane * 128;
              if (!row_in) {
#pragma unroll
                for (int j = 0; j < 8; ++j) *reinterpret_cast<uint4*>(dst + (j << 4)) = make_uint4(0u, 0u, 0u, 0u);
              } else {
                const bool edge = lo > 0 || hi < 64;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  uint32_t pk[4];
#pragma unroll
                  for (int i = 0; i < 4; ++i) {
                    float a = fmaf(__uint_as_float(v[j * 8 + 2 * i]), scale, shift);
                    float b = fmaf(__uint_as_float(v[j * 8 + 2 * i + 1]), scale, shift);
                    if (edge) {
                      const int px = j * 8 + 2 * i;
                      if (px < lo || px >= hi) a = 0.f;
                      if (px + 1 < lo || px + 1 >= hi) b = 0.f;
                    }
                    pk[i] = pack16_relu6(a, b, F16);
                  }
                  *reinterpret_cast<uint4*>(dst + ((j ^ (lane & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                }
              }
              ptx::fence_proxy_async();
              __syncwarp();
              if (lane == 0) {
                tma_store_4d(tm_t1s, sbuf + sb * EX_SBUF_BYTES, 64 * tc, sr, c_warp, g.slot);
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
        ++cnt;
      }
      if (lane == 0) {
        ptx::tma_store_wait_all<0>();
        if (cnt) signal_count(&p.e_done[r], cnt);
      }
      __syncwarp();
      units += cnt;
      base += g.e_tiles;
    }
    if (p.stats && warp == 3 && lane == 0) {
      p.stats[blockIdx.x * 4 + 1] = (unsigned long long)waited;
      p.stats[blockIdx.x * 4 + 3] = (unsigned long long)units;
    }
  }
}

// ==================================================================================================
// role D: depthwise ks x ks on the channel planes of a t1 ring slot -> t2 ring slot (banded-Toeplitz MMAs)
// ==================================================================================================
template <int KS, int F16>
__device__ __forceinline__ void role_depthwise(uint8_t* smem, uint64_t* bars, uint32_t tmem_base, int warp, int lane,
                                               int rank, const CUtensorMap* tm_t1l, const CUtensorMap* tm_t1ls,
                                               const CUtensorMap* tm_t2s, const BandParams& p) {
  constexpr int R = KS >> 1;
  constexpr uint32_t atom_bytes = (uint32_t)((DW_TH + KS - 1) * 128);
  constexpr uint32_t atom_bytes_short = (uint32_t)((DW_TH / 2 + KS - 1) * 128);
  uint8_t* sA = smem;
  uint8_t* sO = sA + DW_A_STAGES * DW_A_STRIDE;
  uint8_t* sB = sO + 2 * DW_OUT_BYTES;
  uint16_t* s_filt = reinterpret_cast<uint16_t*>(sB + 2 * DW_B_BYTES);    // 2 x [KS][32] zero-padded filter rows
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + DW_A_STAGES;
  uint64_t* tfull = a_empty + DW_A_STAGES;
  uint64_t* tempty = tfull + DW_ACC_STAGES;
  uint64_t* b_full = tempty + DW_ACC_STAGES;
  uint64_t* b_empty = b_full + 2;

  // a CTA's unit is a PLANE of a region (tiles_x tiles that share their Toeplitz filter tiles)
  if (warp == 0) {
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      long long base = 0, wclk = 0;
      for (int r = 0; r < p.n_regions; ++r) {
        const Region g = region_info<KS>(p, r);
        bool waited = false;
        for (int c = first_unit(base, rank, p.nD); c < p.mid; c += p.nD) {
          if (!waited) { wclk += wait_count(&p.e_done[r], (unsigned)(g.e_tiles * EX_EPI_WARPS)); waited = true; }
          for (int tx = 0; tx < g.tiles_x; ++tx) {
            const CUtensorMap* tm = g.shrt ? tm_t1ls : tm_t1l;
            ptx::mbar_wait(&a_empty[s], ph ^ 1);
            ptx::mbar_arrive_expect_tx(&a_full[s], 2 * (g.shrt ? atom_bytes_short : atom_bytes));
            ptx::tma_load_4d(sA + s * DW_A_STRIDE, tm, &a_full[s], DW_TW * tx, 0, c, g.slot);
            ptx::tma_load_4d(sA + s * DW_A_STRIDE + DW_ATOM_STRIDE, tm, &a_full[s], DW_TW * tx + 64, 0, c, g.slot);
            if (++s == DW_A_STAGES) { s = 0; ph ^= 1; }
          }
        }
        base += p.mid;
      }
      if (p.stats) p.stats[blockIdx.x * 4 + 1] = (unsigned long long)wclk;
    }
  } else if (warp == 1) {
    constexpr int fmt = F16 ? 0 : 1;
    constexpr uint32_t idesc32_tall = ptx::umma_idesc_f16(128, 32, fmt, fmt, 0, 0);
    constexpr uint32_t idesc_first_tall = ptx::umma_idesc_f16(128, DW_ACC_COLS, fmt, fmt, 0, 0);
    constexpr uint32_t idesc32_short = ptx::umma_idesc_f16(64, 32, fmt, fmt, 0, 0);
    constexpr uint32_t idesc_first_short = ptx::umma_idesc_f16(64, DW_ACC_COLS, fmt, fmt, 0, 0);
    const uint32_t sA_addr = ptx::smem_u32(sA), sB_addr = ptx::smem_u32(sB);
    int s = 0, acc = 0, bi = 0; uint32_t ph = 0, accph = 0, bph = 0;
    bool first_plane = true;
    long long base = 0;
    for (int r = 0; r < p.n_regions; ++r) {
      const Region g = region_info<KS>(p, r);
      const uint32_t idesc32 = g.shrt ? idesc32_short : idesc32_tall;
      const uint32_t idesc_first = g.shrt ? idesc_first_short : idesc_first_tall;
      for (int c = first_unit(base, rank, p.nD); c < p.mid; c += p.nD) {
        if (!first_plane) {
          ptx::umma_commit_elect(&b_empty[bi]);       // all MMAs reading the previous plane's filter tiles are done
          if (++bi == 2) { bi = 0; bph ^= 1; }
        }
        first_plane = false;
        ptx::mbar_wait(&b_full[bi], bph);
        for (int tx = 0; tx < g.tiles_x; ++tx) {
          ptx::mbar_wait(&a_full[s], ph);
          ptx::mbar_wait(&tempty[acc], accph ^ 1);
          ptx::tc_fence_after();
          const uint32_t d0 = tmem_base + (uint32_t)(acc * DW_ACC_COLS);
          const uint64_t da0 = ptx::umma_desc_sw128(sA_addr + (uint32_t)(s * DW_A_STRIDE), 1024);
          const uint64_t dbf = ptx::umma_desc(sB_addr + (uint32_t)(bi * DW_B_BYTES), 128, 256, 0);
          const uint64_t db0 = dbf + (uint64_t)(DW_BFIRST_BYTES >> 4);
          const int valid_w = min(DW_TW, g.cols - DW_TW * tx);
          const int nch = min(DW_CHUNKS, (valid_w + DW_XPAD + R + 15) >> 4);
#pragma unroll
          for (int dy = 0; dy < KS; ++dy) {
#pragma unroll
            for (int j = 0; j < DW_CHUNKS; ++j) {
              const uint64_t da = da0 + (uint64_t)(((j >> 2) * DW_ATOM_STRIDE + dy * 128 + (j & 3) * 32) >> 4);
              if (dy == 0 && j == 0)
                ptx::umma_elect(d0, da, dbf, idesc_first, 0u);
              else if (j < nch)
                ptx::umma_elect(d0 + (uint32_t)(16 * j), da, db0 + (uint64_t)((dy * 1024) >> 4), idesc32, 1u);
            }
          }
          ptx::umma_commit_elect(&a_empty[s]);
          ptx::umma_commit_elect(&tfull[acc]);
          if (++s == DW_A_STAGES) { s = 0; ph ^= 1; }
          if (++acc == DW_ACC_STAGES) { acc = 0; accph ^= 1; }
        }
      }
      base += p.mid;
    }
  } else if (warp == 2) {
    // filter builder: the plane's active filter (precomputed, fp32 in global memory) -> Toeplitz B tiles.  Element (n, k) of
    // a tile is f[dy][k - n + dx0]: only rows n_lo..n_hi can be non-zero, everything else is zeroed ONCE; a row's two
    // 16-byte K-halves are 8 consecutive entries of the zero-padded 16-bit filter row `rows[dy]`.
    int bi = 0; uint32_t bph = 0;
    constexpr int dx0 = DW_XPAD + R;
    constexpr int PADL = 16, ROWLEN = 64;
    constexpr int n_lo = dx0 - KS + 1, n_hi = dx0 + 15, NR = n_hi - n_lo + 1;
    uint16_t* rows = s_filt;                        // [KS][ROWLEN]: rows[dy][PADL + dx] = f[dy][dx], zero elsewhere
    for (int i = lane; i < (2 * DW_B_BYTES) / 16; i += 32)
      *reinterpret_cast<uint4*>(sB + 16 * i) = make_uint4(0u, 0u, 0u, 0u);
    for (int i = lane; i < KS * ROWLEN; i += 32) rows[i] = (uint16_t)0;
    __syncwarp();
    long long base = 0;
    for (int r = 0; r < p.n_regions; ++r) {
      for (int c = first_unit(base, rank, p.nD); c < p.mid; c += p.nD) {
        const float* f = p.filt + (size_t)c * KS * KS;
        const float t0 = lane < KS * KS ? f[lane] : 0.f;                 // issued before the wait: latency overlaps it
        const float t1 = lane + 32 < KS * KS ? f[lane + 32] : 0.f;
        ptx::mbar_wait(&b_empty[bi], bph ^ 1);
        if (lane < KS * KS) rows[(lane / KS) * ROWLEN + PADL + lane % KS] = cvt16(t0, F16);
        if (lane + 32 < KS * KS) rows[((lane + 32) / KS) * ROWLEN + PADL + (lane + 32) % KS] = cvt16(t1, F16);
        __syncwarp();
        uint8_t* b0 = sB + bi * DW_B_BYTES;
        // tile 0: the N = 144 first matrix (dy = 0); tiles 1..KS: the N = 32 matrices of dy = 0..KS-1
        for (int i = lane; i < (KS + 1) * NR * 2; i += 32) {
          const int tt = i / (NR * 2), rem = i - tt * (NR * 2);
          const int n = n_lo + (rem >> 1), kh = rem & 1;
          const int dy = tt == 0 ? 0 : tt - 1;
          const uint16_t* src = rows + dy * ROWLEN + (kh * 8 - n + dx0 + PADL);
          uint32_t w[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) w[q] = (uint32_t)src[2 * q] | ((uint32_t)src[2 * q + 1] << 16);
          uint8_t* dst = (tt == 0 ? b0 : b0 + DW_BFIRST_BYTES + dy * 1024) + dw_b_off(n, kh * 8);
          *reinterpret_cast<uint4*>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
        }
        ptx::fence_proxy_async();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&b_full[bi]);
        if (++bi == 2) { bi = 0; bph ^= 1; }
      }
      base += p.mid;
    }
  } else {
    const int quarter = warp & 3;
    const int hf = (warp - 3) >> 2;
    const bool issuer = (warp == 3 && lane == 0);
    int acc = 0, ob = 0; uint32_t accph = 0;
    long long base = 0, wclk = 0, units = 0;
    for (int r = 0; r < p.n_regions; ++r) {
      const Region g = region_info<KS>(p, r);
      const int row = g.shrt ? quarter * 16 + lane : quarter * 32 + lane;
      const bool live = !g.shrt || lane < 16;
      bool slot_ok = (r < p.S);
      unsigned cnt = 0;
      for (int c = first_unit(base, rank, p.nD); c < p.mid; c += p.nD) {
        float scale, shift;
        bn_fold(p.g2, p.b2, p.m2, p.v2, p.eps2, c, scale, shift);
        for (int tx = 0; tx < g.tiles_x; ++tx) {
          ptx::mbar_wait(&tfull[acc], accph);
          ptx::tc_fence_after();
          const uint32_t t_addr =
              tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * DW_ACC_COLS + 16 + hf * 56);
          uint32_t v[56];
          ptx::tmem_ld16(t_addr, v);
          ptx::tmem_ld16(t_addr + 16, v + 16);
          ptx::tmem_ld16(t_addr + 32, v + 32);
          ptx::tmem_ld8(t_addr + 48, v + 48);
          ptx::tmem_ld_wait();
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&tempty[acc]);
          if (++acc == DW_ACC_STAGES) { acc = 0; accph ^= 1; }
          uint32_t pk[28];
#pragma unroll
          for (int i = 0; i < 28; ++i) {
            float a = fmaf(__uint_as_float(v[2 * i]), scale, shift);
            float b = fmaf(__uint_as_float(v[2 * i + 1]), scale, shift);
            pk[i] = pack16_relu6(a, b, F16);
          }
          if (issuer) ptx::tma_store_wait_read<1>();
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
            if (!slot_ok) {                           // ring slot reuse: the project stage is done with region r - S
              wclk += wait_count(&p.p_done[r - p.S], (unsigned)region_info<KS>(p, r - p.S).p_tiles);
              slot_ok = true;
            }
            tma_store_4d(tm_t2s, sO + ob * DW_OUT_BYTES, DW_TW * tx, 0, c, g.slot);
            ptx::tma_store_commit();
          }
          ob ^= 1;
          ++cnt;
        }
      }
      if (issuer) {
        ptx::tma_store_wait_all<0>();
        if (cnt) signal_count(&p.d_done[r], cnt);
      }
      units += cnt;
      base += p.mid;
    }
    if (p.stats && issuer) {
      p.stats[blockIdx.x * 4 + 2] = (unsigned long long)wclk;
      p.stats[blockIdx.x * 4 + 3] = (unsigned long long)units;
    }
  }
}

// ==================================================================================================
// role P: project mid -> 64 from a t2 ring slot, folded BN + residual, -> trunk y
// ==================================================================================================
template <int KS, int F16>
__device__ __forceinline__ void role_project(uint8_t* smem, uint64_t* bars, uint32_t tmem_base, int warp, int lane,
                                             int rank, const CUtensorMap* tm_t2l, const CUtensorMap* tm_w,
                                             const CUtensorMap* tm_r, const CUtensorMap* tm_y,
                                             const CUtensorMap* tm_yp, const BandParams& p) {
  uint8_t* sA = smem;
  uint8_t* sW = sA + BD_PJ_STAGES * PJ_A_BYTES;
  uint8_t* sR = sW + PJ_MAX_KC * 8192;
  float* s_scale = reinterpret_cast<float*>(sR + 2 * PJ_R_BYTES);
  float* s_shift = s_scale + 64;
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + BD_PJ_STAGES;
  uint64_t* tfull = a_empty + BD_PJ_STAGES;
  uint64_t* tempty = tfull + PJ_ACC_STAGES;
  uint64_t* r_full = tempty + PJ_ACC_STAGES;
  uint64_t* r_empty = r_full + 2;
  uint64_t* w_bar = r_empty + 2;

  if (warp == 0) {
    if (lane == 0) {
      ptx::mbar_arrive_expect_tx(w_bar, (uint32_t)(p.kcs * 8192));
      for (int kc = 0; kc < p.kcs; ++kc) ptx::tma_load_3d(sW + kc * 8192, tm_w, w_bar, kc * 64, 0, 0);
      int s = 0, rb = 0; uint32_t ph = 0, rph = 0;
      long long base = 0, wclk = 0, units = 0;
      for (int r = 0; r < p.n_regions; ++r) {
        const Region g = region_info<KS>(p, r);
        bool waited = false;
        for (int t = first_unit(base, rank, p.nP); t < g.p_tiles; t += p.nP) {
          ++units;
          if (!waited) { wclk += wait_count(&p.d_done[r], (unsigned)g.d_tiles); waited = true; }
          int brow[2], bcol[2];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int b = 2 * t + h;
            brow[h] = b / g.bpr;
            bcol[h] = 64 * (b - brow[h] * g.bpr);
          }
          ptx::mbar_wait(&r_empty[rb], rph ^ 1);
          if (p.has_res) {
            ptx::mbar_arrive_expect_tx(&r_full[rb], PJ_R_BYTES);
#pragma unroll
            for (int h = 0; h < 2; ++h)
              ptx::tma_load_3d(sR + rb * PJ_R_BYTES + h * 8192, tm_r, &r_full[rb], 0, g.gx0 + bcol[h],
                               g.n * p.H + g.y0 + brow[h]);
          } else {
            ptx::mbar_arrive(&r_full[rb]);
          }
          if (++rb == 2) { rb = 0; rph ^= 1; }
          for (int kc = 0; kc < p.kcs; ++kc) {
            ptx::mbar_wait(&a_empty[s], ph ^ 1);
            ptx::mbar_arrive_expect_tx(&a_full[s], PJ_A_BYTES);
#pragma unroll
            for (int h = 0; h < 2; ++h)
              ptx::tma_load_4d(sA + s * PJ_A_BYTES + h * 8192, tm_t2l, &a_full[s], bcol[h], min(brow[h], DW_TH - 1),
                               kc * 64, g.slot);
            if (++s == BD_PJ_STAGES) { s = 0; ph ^= 1; }
          }
        }
        base += g.p_tiles;
      }
      if (p.stats) {
        p.stats[blockIdx.x * 4 + 1] = (unsigned long long)wclk;
        p.stats[blockIdx.x * 4 + 3] = (unsigned long long)units;
      }
    }
  } else if (warp == 1) {
    constexpr int fmt = F16 ? 0 : 1;
    const uint32_t idesc = ptx::umma_idesc_f16(128, 64, fmt, fmt, /*A MN-major*/ 1, 0);
    const uint32_t sA_addr = ptx::smem_u32(sA), sW_addr = ptx::smem_u32(sW);
    int s = 0, acc = 0; uint32_t ph = 0, accph = 0;
    ptx::mbar_wait(w_bar, 0);
    long long base = 0;
    for (int r = 0; r < p.n_regions; ++r) {
      const Region g = region_info<KS>(p, r);
      for (int t = first_unit(base, rank, p.nP); t < g.p_tiles; t += p.nP) {
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
            const uint64_t da = ptx::umma_desc(a0 + (uint32_t)(k * 2048), 8192, 1024, 2);
            ptx::umma_elect(d0, da, db + (uint64_t)(k * 2), idesc, (uint32_t)(kc | k));
          }
          ptx::umma_commit_elect(&a_empty[s]);
          if (++s == BD_PJ_STAGES) { s = 0; ph ^= 1; }
        }
        ptx::umma_commit_elect(&tfull[acc]);
        if (++acc == PJ_ACC_STAGES) { acc = 0; accph ^= 1; }
      }
      base += g.p_tiles;
    }
  } else if (warp >= 3 && warp < 7) {
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const bool issuer = (warp == 3 && lane == 0);
    int acc = 0, rb = 0, prev_rb = -1; uint32_t accph = 0, rph = 0;
    long long base = 0;
    for (int r = 0; r < p.n_regions; ++r) {
      const Region g = region_info<KS>(p, r);
      unsigned cnt = 0;
      for (int t = first_unit(base, rank, p.nP); t < g.p_tiles; t += p.nP) {
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
            const float2 r2 = unpack16(rr[i], p.xf16);
            const float a = fmaf(__uint_as_float(v[ch]), s_scale[ch], s_shift[ch]) + r2.x;
            const float b = fmaf(__uint_as_float(v[ch + 1]), s_scale[ch + 1], s_shift[ch + 1]) + r2.y;
            pk[i] = pack16(a, b, p.xf16);
          }
          *q = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
        ptx::fence_proxy_async();
        ptx::named_bar_sync(1, 128);
        if (issuer) {
          if (prev_rb >= 0) { ptx::tma_store_wait_read<0>(); ptx::mbar_arrive(&r_empty[prev_rb]); }
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int b = 2 * t + h;
            const int brow = b / g.bpr, bcol = 64 * (b - brow * g.bpr);
            if (brow < g.rows) {
              // an interior region's last box of a row is only rw % 64 pixels wide (its neighbour owns the rest); at the
              // right image edge the full box is clipped by TMA
              const bool part = (g.cols - bcol < 64) && (g.gx0 + bcol + 64 <= p.W);
              ptx::tma_store_3d(part ? tm_yp : tm_y, sR + rb * PJ_R_BYTES + h * 8192, 0, g.gx0 + bcol,
                                g.n * p.H + g.y0 + brow);
            }
          }
          ptx::tma_store_commit();
          prev_rb = rb;
        }
        if (++rb == 2) { rb = 0; rph ^= 1; }
        ++cnt;
      }
      // every A tile of this CTA's share of region r has been consumed (its accumulators were complete): the t2 slot
      // may be recycled once all project CTAs said so
      if (issuer && cnt) signal_count(&p.p_done[r], cnt);
      base += g.p_tiles;
    }
    if (issuer) ptx::tma_store_wait_all<0>();
  }
}

template <int KS, int F16>
__global__ void __launch_bounds__(BD_THREADS, 1)
mbconv_band_kernel(const __grid_constant__ CUtensorMap tm_x4, const __grid_constant__ CUtensorMap tm_we,
                   const __grid_constant__ CUtensorMap tm_t1s, const __grid_constant__ CUtensorMap tm_t1l,
                   const __grid_constant__ CUtensorMap tm_t1ls, const __grid_constant__ CUtensorMap tm_t2s,
                   const __grid_constant__ CUtensorMap tm_t2l, const __grid_constant__ CUtensorMap tm_wp,
                   const __grid_constant__ CUtensorMap tm_r, const __grid_constant__ CUtensorMap tm_y,
                   const __grid_constant__ CUtensorMap tm_yp, const BandParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BD_BAR_OFF);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 48);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const int role = (int)blockIdx.x < p.nE ? 0 : ((int)blockIdx.x < p.nE + p.nD ? 1 : 2);

  if (warp == 0 && lane == 0) {
    if (role == 0) { ptx::prefetch_tmap(&tm_x4); ptx::prefetch_tmap(&tm_we); ptx::prefetch_tmap(&tm_t1s); }
    else if (role == 1) { ptx::prefetch_tmap(&tm_t1l); ptx::prefetch_tmap(&tm_t1ls); ptx::prefetch_tmap(&tm_t2s); }
    else { ptx::prefetch_tmap(&tm_t2l); ptx::prefetch_tmap(&tm_wp); ptx::prefetch_tmap(&tm_r); ptx::prefetch_tmap(&tm_y);
           ptx::prefetch_tmap(&tm_yp); }
  }
  if (warp == 1 && lane == 0) {
    if (role == 0) {
      uint64_t* x_full = bars; uint64_t* x_empty = x_full + EX_X_STAGES; uint64_t* tfull = x_empty + EX_X_STAGES;
      uint64_t* tempty = tfull + 2; uint64_t* w_bar = tempty + 2;
      for (int s = 0; s < EX_X_STAGES; ++s) { ptx::mbar_init(&x_full[s], 1); ptx::mbar_init(&x_empty[s], 1); }
      for (int a = 0; a < 2; ++a) { ptx::mbar_init(&tfull[a], 1); ptx::mbar_init(&tempty[a], EX_EPI_WARPS); }
      ptx::mbar_init(w_bar, 1);
    } else if (role == 1) {
      uint64_t* a_full = bars; uint64_t* a_empty = a_full + DW_A_STAGES; uint64_t* tfull = a_empty + DW_A_STAGES;
      uint64_t* tempty = tfull + DW_ACC_STAGES; uint64_t* b_full = tempty + DW_ACC_STAGES; uint64_t* b_empty = b_full + 2;
      for (int s = 0; s < DW_A_STAGES; ++s) { ptx::mbar_init(&a_full[s], 1); ptx::mbar_init(&a_empty[s], 1); }
      for (int a = 0; a < DW_ACC_STAGES; ++a) { ptx::mbar_init(&tfull[a], 1); ptx::mbar_init(&tempty[a], DW_EPI_WARPS); }
      for (int b = 0; b < 2; ++b) { ptx::mbar_init(&b_full[b], 1); ptx::mbar_init(&b_empty[b], 1); }
    } else {
      uint64_t* a_full = bars; uint64_t* a_empty = a_full + BD_PJ_STAGES; uint64_t* tfull = a_empty + BD_PJ_STAGES;
      uint64_t* tempty = tfull + PJ_ACC_STAGES; uint64_t* r_full = tempty + PJ_ACC_STAGES; uint64_t* r_empty = r_full + 2;
      uint64_t* w_bar = r_empty + 2;
      for (int s = 0; s < BD_PJ_STAGES; ++s) { ptx::mbar_init(&a_full[s], 1); ptx::mbar_init(&a_empty[s], 1); }
      for (int a = 0; a < PJ_ACC_STAGES; ++a) { ptx::mbar_init(&tfull[a], 1); ptx::mbar_init(&tempty[a], 4); }
      for (int b = 0; b < 2; ++b) { ptx::mbar_init(&r_full[b], 1); ptx::mbar_init(&r_empty[b], 1); }
      ptx::mbar_init(w_bar, 1);
    }
    ptx::fence_barrier_init();
  }
  if (role == 2 && threadIdx.x < 64) {
    float* s_scale = reinterpret_cast<float*>(smem + BD_PJ_STAGES * PJ_A_BYTES + PJ_MAX_KC * 8192 + 2 * PJ_R_BYTES);
    float sc, sh;
    bn_fold(p.g3, p.b3, p.m3, p.v3, p.eps3, threadIdx.x, sc, sh);
    s_scale[threadIdx.x] = sc;
    s_scale[64 + threadIdx.x] = sh;
  }
  if (warp == 2) { ptx::tmem_alloc(tmem_ptr, 512); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);
  const long long t_start = clock64();

  if (role == 0)
    role_expand<KS, F16>(smem, bars, tmem_base, warp, lane, (int)blockIdx.x, &tm_x4, &tm_we, &tm_t1s, p);
  else if (role == 1)
    role_depthwise<KS, F16>(smem, bars, tmem_base, warp, lane, (int)blockIdx.x - p.nE, &tm_t1l, &tm_t1ls, &tm_t2s, p);
  else
    role_project<KS, F16>(smem, bars, tmem_base, warp, lane, (int)blockIdx.x - p.nE - p.nD, &tm_t2l, &tm_wp, &tm_r, &tm_y,
                          &tm_yp, p);

  ptx::tc_fence_before();
  __syncthreads();
  if (p.stats && threadIdx.x == 0) p.stats[blockIdx.x * 4] = (unsigned long long)(clock64() - t_start);
  if (warp == 2) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem_base, 512); }
}

CUtensorMapDataType dt16(int f16) { return f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16; }

int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}

}  // namespace

// Region / ring geometry for a frame (also used by ofa_mbconv_workspace_bytes).  RT = region width in depthwise tiles.
static void band_geometry(int W, int mid, int* RT, int* S, long long* t1_slot, long long* t2_slot) {
  int rt = env_int("OFA_BAND_RT", 1);
  if (rt < 1) rt = 1;
  if (rt > 2) rt = 2;
  if (W <= DW_TW) rt = 1;
  int s = env_int("OFA_BAND_SLOTS", rt == 1 ? 3 : 2);
  if (s < 2) s = 2;
  if (s > 4) s = 4;
  *RT = rt; *S = s;
  *t1_slot = (long long)mid * BD_T1H * (128 * rt) * 2;
  *t2_slot = (long long)mid * DW_TH * (128 * rt) * 2;
}

long long mbconv_band_workspace_bytes(int W, int mid, int n_regions_max) {
  int RT, S; long long t1, t2;
  band_geometry(W, mid, &RT, &S, &t1, &t2);
  return S * (t1 + t2) + (long long)mid * 49 * 4 + 3ll * n_regions_max * 4 + 4096;
}

bool mbconv_band_supported(const OfaMBConvArgs* a) {
  if (!mbconv_planar_supported(a)) return false;
  if (a->act != OFA_ACT_RELU6) return false;
  if (a->ks != 3 && a->ks != 5 && a->ks != 7) return false;
  if (env_int("OFA_BAND_DISABLE", 0)) return false;
  return true;
}

// OFA_IMPL_AUTO never picks this kernel: measured at C2 (B200) it is SLOWER than the three stand-alone kernels
// (1.08 ms against 0.39 ms per block), although its intermediates stay in L2 -- each stage is bounded per SM (tcgen05
// issue rate, epilogue, TMA request rate), not by HBM, so removing the HBM round trip buys nothing while the region
// hand-over adds halo recompute and pipeline drains.  OFA_BAND_ENABLE=1 or OFA_IMPL_BAND select it.
bool mbconv_band_preferred(const OfaMBConvArgs* a) {
  const long long area = (long long)a->x.h * a->x.w;
  return env_int("OFA_BAND_ENABLE", 0) && mbconv_planar_preferred(a) && area >= 128ll * 448 && a->x.w >= 224 &&
         a->x.h >= 96;
}

int launch_mbconv_band(const OfaMBConvArgs* a, const void* wexp_p, const void* wproj_p, int f16, void* ws,
                       long long ws_bytes, cudaStream_t st) {
  const int N = a->x.n, H = a->x.h, W = a->x.w, mid = a->mid, ks = a->ks;
  const int tf16 = a->x.dtype == OFA_F16 ? 1 : 0;
  BandParams p;
  memset(&p, 0, sizeof(p));
  int RT, S; long long t1_slot, t2_slot;
  band_geometry(W, mid, &RT, &S, &t1_slot, &t2_slot);
  p.N = N; p.H = H; p.W = W; p.mid = mid; p.mt = (mid + 127) / 128; p.kcs = mid / 64; p.xf16 = tf16;
  p.RT = RT; p.rw = DW_TW * RT; p.S = S;
  p.n_bands = (H + DW_TH - 1) / DW_TH;
  p.n_cgs = (W + p.rw - 1) / p.rw;
  const long long nreg = (long long)N * p.n_bands * p.n_cgs;
  if (nreg >= (1ll << 24)) return fail(OFA_ERR_UNSUPPORTED, "mbconv band: too many regions");
  p.n_regions = (int)nreg;
  const int rem = H % DW_TH;
  p.short_last = (rem > 0 && rem <= DW_TH / 2) ? 1 : 0;
  p.has_res = a->add_residual ? 1 : 0;
  p.part_px = p.rw % 64;
  p.g1 = a->bn_exp.gamma; p.b1 = a->bn_exp.beta; p.m1 = a->bn_exp.mean; p.v1 = a->bn_exp.var; p.eps1 = a->bn_exp.eps;
  p.g2 = a->bn_dw.gamma; p.b2 = a->bn_dw.beta; p.m2 = a->bn_dw.mean; p.v2 = a->bn_dw.var; p.eps2 = a->bn_dw.eps;
  p.g3 = a->bn_proj.gamma; p.b3 = a->bn_proj.beta; p.m3 = a->bn_proj.mean; p.v3 = a->bn_proj.var; p.eps3 = a->bn_proj.eps;

  // workspace: [t1 ring][t2 ring][active filters][counters]
  const long long need = S * (t1_slot + t2_slot) + (long long)mid * 49 * 4 + 3ll * p.n_regions * 4 + 4096;
  if (ws_bytes < need) return fail(OFA_ERR_ARG, "mbconv band: workspace too small (%lld < %lld)", ws_bytes, need);
  char* w8 = reinterpret_cast<char*>(ws);
  void* t1 = w8;
  void* t2 = w8 + S * t1_slot;
  float* filt = reinterpret_cast<float*>(w8 + S * (t1_slot + t2_slot));
  unsigned* flags = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(filt) + (((long long)mid * 49 * 4 + 255) / 256) * 256);
  p.filt = filt;
  p.e_done = flags; p.d_done = flags + p.n_regions; p.p_done = flags + 2 * p.n_regions;
  int rc;
  if ((rc = launch_active_filter(a->w_dw, a->kmax, a->m75, a->m53, a->transform_on, ks, mid, filt, st))) return rc;
  // OFA_BAND_ONLY = 1 | 2 | 3 (bring-up / profiling): run ONE role on every SM with all hand-over counters pre-satisfied
  // (results are garbage; the time is that stage's L2-resident throughput)
  const int only = env_int("OFA_BAND_ONLY", 0);
  OFA_CUDA(cudaMemsetAsync(flags, only ? 0x7f : 0, 3ull * p.n_regions * 4, st));

  // role split of the SMs: proportional to each stage's per-SM cost (shared-memory-port time of D, L2 write time of E)
  const int sms = sm_count();
  int nE = env_int("OFA_BAND_NE", 0), nP = env_int("OFA_BAND_NP", 0);
  if (nE <= 0 || nP <= 0) {
    // per-pixel cost model in "SM clocks per pixel per channel" (measured)
    const double cE = 0.062 * 1.2, cD = (ks == 7 ? 0.27 : ks == 5 ? 0.21 : 0.16), cP = 0.046;
    const double tot = cE + cD + cP;
    nE = (int)(sms * cE / tot + 0.5);
    nP = (int)(sms * cP / tot + 0.5);
  }
  if (nE < 1) nE = 1;
  if (nP < 1) nP = 1;
  int nD = sms - nE - nP;
  if (nD < 1) return fail(OFA_ERR_UNSUPPORTED, "mbconv band: not enough SMs");
  if (nD > mid) nD = mid;
  if (only == 1) { nE = sms; nD = 0; nP = 0; }
  if (only == 2) { nE = 0; nD = sms < mid ? sms : mid; nP = 0; }
  if (only == 3) { nE = 0; nD = 0; nP = sms; }
  p.nE = nE; p.nD = nD; p.nP = nP;

  CUtensorMap tx4, twe, tt1s, tt1l, tt1ls, tt2s, tt2l, twp, tr, ty, typ;
  const uint64_t T1W = 128ull * RT, T2W = 128ull * RT;
  {
    uint64_t dims[4] = {64, (uint64_t)W, (uint64_t)H, (uint64_t)N};
    uint64_t str[3] = {128, (uint64_t)W * 128, (uint64_t)H * W * 128};
    uint32_t box[4] = {64, 64, 4, 1};
    if ((rc = encode_tmap(&tx4, dt16(tf16), 4, const_cast<void*>(a->x.ptr), dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)))
      return rc;
  }
  {
    uint64_t dims[3] = {64, (uint64_t)p.mt * 128, 1};
    uint64_t str[2] = {128, (uint64_t)p.mt * 128 * 128};
    uint32_t box[3] = {64, 128, 1};
    if ((rc = encode_tmap(&twe, dt16(tf16), 3, const_cast<void*>(wexp_p), dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)))
      return rc;
  }
  {
    uint64_t dims[4] = {T1W, BD_T1H, (uint64_t)mid, (uint64_t)S};
    uint64_t str[3] = {T1W * 2, T1W * 2 * BD_T1H, (uint64_t)t1_slot};
    uint32_t box[4] = {64, 1, 32, 1};
    if ((rc = encode_tmap(&tt1s, dt16(f16), 4, t1, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    uint32_t boxl[4] = {64, (uint32_t)(DW_TH + ks - 1), 1, 1};
    if ((rc = encode_tmap(&tt1l, dt16(f16), 4, t1, dims, str, boxl, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    boxl[1] = (uint32_t)(DW_TH / 2 + ks - 1);
    if ((rc = encode_tmap(&tt1ls, dt16(f16), 4, t1, dims, str, boxl, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  }
  {
    uint64_t dims[4] = {T2W, DW_TH, (uint64_t)mid, (uint64_t)S};
    uint64_t str[3] = {T2W * 2, T2W * 2 * DW_TH, (uint64_t)t2_slot};
    uint32_t box[4] = {DW_TW, DW_TH, 1, 1};
    if ((rc = encode_tmap(&tt2s, dt16(f16), 4, t2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_NONE))) return rc;
    uint32_t boxl[4] = {64, 1, 64, 1};
    if ((rc = encode_tmap(&tt2l, dt16(f16), 4, t2, dims, str, boxl, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)mid, 64, 1};
    uint64_t str[2] = {(uint64_t)mid * 2, (uint64_t)mid * 64 * 2};
    uint32_t box[3] = {64, 64, 1};
    if ((rc = encode_tmap(&twp, dt16(f16), 3, const_cast<void*>(wproj_p), dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)))
      return rc;
  }
  {
    uint64_t dims[3] = {64, (uint64_t)W, (uint64_t)H * N};
    uint64_t str[2] = {128, (uint64_t)W * 128};
    uint32_t box[3] = {64, 64, 1};
    if ((rc = encode_tmap(&tr, dt16(tf16), 3, const_cast<void*>(a->x.ptr), dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)))
      return rc;
    if ((rc = encode_tmap(&ty, dt16(tf16), 3, a->y.ptr, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    box[1] = (uint32_t)(p.part_px ? p.part_px : 64);
    if ((rc = encode_tmap(&typ, dt16(tf16), 3, a->y.ptr, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  }

  unsigned long long* stats = nullptr;
  const int want_stats = env_int("OFA_BAND_STATS", 0);
  if (want_stats) {
    OFA_CUDA(cudaMallocManaged(&stats, (size_t)(nE + nD + nP) * 4 * sizeof(unsigned long long)));
    memset(stats, 0, (size_t)(nE + nD + nP) * 4 * sizeof(unsigned long long));
  }
  p.stats = stats;
  void* args[] = {&tx4, &twe, &tt1s, &tt1l, &tt1ls, &tt2s, &tt2l, &twp, &tr, &ty, &typ, &p};
  const int grid = nE + nD + nP;
#define OFA_BAND_LAUNCH(KS_, F16_)                                                                                 \
  do {                                                                                                             \
    OFA_CUDA(cudaFuncSetAttribute(mbconv_band_kernel<KS_, F16_>, cudaFuncAttributeMaxDynamicSharedMemorySize,      \
                                  BD_SMEM));                                                                       \
    OFA_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(&mbconv_band_kernel<KS_, F16_>), dim3(grid), \
                                         dim3(BD_THREADS), args, (size_t)BD_SMEM, st));                            \
  } while (0)
  if (ks == 3) { if (f16) OFA_BAND_LAUNCH(3, 1); else OFA_BAND_LAUNCH(3, 0); }
  else if (ks == 5) { if (f16) OFA_BAND_LAUNCH(5, 1); else OFA_BAND_LAUNCH(5, 0); }
  else { if (f16) OFA_BAND_LAUNCH(7, 1); else OFA_BAND_LAUNCH(7, 0); }
#undef OFA_BAND_LAUNCH
  if (want_stats) {
    OFA_CUDA(cudaStreamSynchronize(st));
    const char* names[3] = {"E", "D", "P"};
    const int lo[3] = {0, nE, nE + nD}, hi[3] = {nE, nE + nD, nE + nD + nP};
    for (int k = 0; k < 3; ++k) {
      double tot = 0, w1 = 0, w2 = 0, un = 0, mx = 0;
      for (int b = lo[k]; b < hi[k]; ++b) {
        tot += (double)stats[b * 4]; w1 += (double)stats[b * 4 + 1]; w2 += (double)stats[b * 4 + 2];
        un += (double)stats[b * 4 + 3];
        if ((double)stats[b * 4] > mx) mx = (double)stats[b * 4];
      }
      const int n = hi[k] - lo[k];
      fprintf(stderr, "band stats role %s: %3d CTAs  avg %.0f clk (max %.0f)  wait-in %.0f  wait-slot %.0f  units/CTA %.1f  "
              "busy clk/unit %.0f\n", names[k], n, tot / n, mx, w1 / n, w2 / n, un / n,
              un > 0 ? (tot - w1 - w2) / un : 0.0);
    }
    fprintf(stderr, "band stats: regions %d  RT %d  S %d  ks %d  mid %d\n", p.n_regions, RT, S, ks, mid);
    cudaFree(stats);
  }
  return check_launch("mbconv_band_kernel");
}

}  // namespace ofa
