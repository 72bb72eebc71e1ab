// Hardware probe (not part of the library): can a tcgen05 shared-memory descriptor address a SHIFTED
// window of a 128B-swizzled halo tile that TMA wrote?  That is what lets a k x k convolution load
// its activation halo once and feed every filter tap from shared memory.
//
//   halo tile : [20 rows][16 cols][64 ch] bf16 = rows of 128 B, TMA SWIZZLE_128B, 1024-B aligned
//   window    : 16 rows x 8 cols at (ky, kx): group g (8 pixels of row g) starts at
//               base + ((g + ky) * 16 + kx) * 128  ->  SBO = 2048, start = base + (ky*16 + kx)*128
//   variants  : base_offset field = 0, or (start >> 7) & 7
// Prints the max abs error of D = A_window * W^T against a host reference for each variant / tap.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe umma_probe.cu ../conv_tc.cu ... (see run_probe.sh)
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "../sm100_ptx.cuh"

using namespace ofa;

constexpr int HR = 20, HC = 16, CH = 64, NOUT = 64;
constexpr int NTAPS = 6;
__constant__ int c_taps[NTAPS][2];

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t sbo, uint32_t base_off) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(base_off & 7) << 49;
  d |= (uint64_t)2 << 61;
  return d;
}

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap tmx, const __grid_constant__ CUtensorMap tmw, float* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                              // 20*16*128 = 40960
  uint8_t* sB = smem + HR * HC * 128;              // 64*128 = 8192
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + NOUT * 128);
  uint64_t* mbar = bar + 1;
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    ptx::mbar_init(bar, 1);
    ptx::mbar_init(mbar, 1);
    ptx::fence_barrier_init();
    ptx::mbar_arrive_expect_tx(bar, HR * HC * 128 + NOUT * 128);
    ptx::tma_load_3d(sA, &tmx, bar, 0, 0, 0);
    ptx::tma_load_3d(sB, &tmw, bar, 0, 0, 0);
  }
  if (warp == 0) {
    ptx::tmem_alloc(tptr, 64);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tptr;
  ptx::mbar_wait(bar, 0);
  uint32_t phase = 0;
  for (int variant = 0; variant < 2; ++variant) {
    for (int t = 0; t < NTAPS; ++t) {
      const int ky = c_taps[t][0], kx = c_taps[t][1];
      if (threadIdx.x == 0) {
        ptx::tc_fence_after();
        const uint32_t a_addr = ptx::smem_u32(sA) + (uint32_t)((ky * HC + kx) * 128);
        const uint32_t bo = variant ? ((a_addr >> 7) & 7) : 0;
        const uint64_t da = make_desc(a_addr, 2048, bo);
        const uint64_t db = make_desc(ptx::smem_u32(sB), 1024, 0);
        const uint32_t idesc = ptx::umma_idesc_bf16(128, NOUT);
        for (int k = 0; k < 4; ++k) ptx::umma_bf16(tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, k ? 1u : 0u);
        ptx::umma_commit(mbar);
      }
      ptx::mbar_wait(mbar, phase);
      phase ^= 1;
      ptx::tc_fence_after();
      const int row = warp * 32 + lane;
      float* o = out + (((size_t)variant * NTAPS + t) * 128 + row) * NOUT;
      for (int j = 0; j < NOUT; j += 16) {
        uint32_t v[16];
        ptx::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + j, v);
        ptx::tmem_ld_wait();
        for (int i = 0; i < 16; ++i) o[j + i] = __uint_as_float(v[i]);
      }
      ptx::tc_fence_before();
      __syncthreads();
    }
  }
  if (warp == 0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, 64);
  }
}

static int encode(CUtensorMap* out, void* base, const uint64_t* dims, const uint64_t* strides, const uint32_t* box,
                  int rank) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) return 1;
  auto fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  uint32_t es[5] = {1, 1, 1, 1, 1};
  return fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS;
}

int main() {
  const int GH = 24, GW = 16;
  std::vector<__nv_bfloat16> hx((size_t)GH * GW * CH), hw((size_t)NOUT * CH);
  std::vector<float> fx(hx.size()), fw(hw.size());
  srand(1);
  for (size_t i = 0; i < hx.size(); ++i) { hx[i] = __float2bfloat16((rand() % 2001 - 1000) / 1000.f); fx[i] = __bfloat162float(hx[i]); }
  for (size_t i = 0; i < hw.size(); ++i) { hw[i] = __float2bfloat16((rand() % 2001 - 1000) / 4000.f); fw[i] = __bfloat162float(hw[i]); }
  int taps[NTAPS][2] = {{0, 0}, {0, 1}, {1, 3}, {2, 5}, {4, 4}, {3, 7}};
  cudaMemcpyToSymbol(c_taps, taps, sizeof(taps));
  __nv_bfloat16 *dx, *dw;
  float* dout;
  cudaMalloc(&dx, hx.size() * 2); cudaMalloc(&dw, hw.size() * 2);
  cudaMalloc(&dout, sizeof(float) * 2 * NTAPS * 128 * NOUT);
  cudaMemcpy(dx, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dw, hw.data(), hw.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap tmx, tmw;
  {
    uint64_t dims[3] = {CH, GW, GH};
    uint64_t str[2] = {CH * 2, (uint64_t)GW * CH * 2};
    uint32_t box[3] = {CH, HC, HR};
    if (encode(&tmx, dx, dims, str, box, 3)) { printf("encode x failed\n"); return 1; }
  }
  {
    uint64_t dims[3] = {CH, NOUT, 1};
    uint64_t str[2] = {CH * 2, (uint64_t)NOUT * CH * 2};
    uint32_t box[3] = {CH, NOUT, 1};
    if (encode(&tmw, dw, dims, str, box, 3)) { printf("encode w failed\n"); return 1; }
  }
  const int smem = HR * HC * 128 + NOUT * 128 + 64 + 1024;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe_kernel<<<1, 128, smem>>>(tmx, tmw, dout);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 2; }
  std::vector<float> ho((size_t)2 * NTAPS * 128 * NOUT);
  cudaMemcpy(ho.data(), dout, ho.size() * 4, cudaMemcpyDeviceToHost);
  for (int v = 0; v < 2; ++v)
    for (int t = 0; t < NTAPS; ++t) {
      double maxerr = 0, maxref = 0;
      for (int r = 0; r < 128; ++r) {
        int g = r / 8, c = r % 8;
        for (int n = 0; n < NOUT; ++n) {
          double acc = 0;
          for (int k = 0; k < CH; ++k)
            acc += (double)fx[((size_t)(g + taps[t][0]) * GW + (c + taps[t][1])) * CH + k] * fw[(size_t)n * CH + k];
          double d = fabs(acc - ho[(((size_t)v * NTAPS + t) * 128 + r) * NOUT + n]);
          if (d > maxerr) maxerr = d;
          if (fabs(acc) > maxref) maxref = fabs(acc);
        }
      }
      printf("variant %d (base_offset %s) tap (%d,%d): max err %.5f (max |ref| %.3f) %s\n", v, v ? "=(addr>>7)&7" : "=0",
             taps[t][0], taps[t][1], maxerr, maxref, maxerr < 1e-2 * maxref ? "OK" : "MISMATCH");
    }
  return 0;
}
