// Hardware probe (not part of the library): HBM read bandwidth through TMA for the access patterns of the planar MBConv
// kernels -- is the ~4.2-4.8 TB/s they reach a property of their 128-256-byte row granularity?
//   pattern A: depthwise input window as the kernel loads it: 2 boxes of [134 rows x 64 cols] 16-bit from a plane with
//              a 1920-byte row pitch (960-pixel rows), tiles walked left to right, top to bottom, plane after plane
//   pattern B: the same bytes when a plane is stored as 128-column strips (row pitch 256 B): a window is ONE contiguous
//              34 KB run
//   pattern C: project's operand: [64 pixels x 64 planes] boxes, planes 1 MB apart
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hbm_pattern_probe hbm_pattern_probe.cu && ./hbm_pattern_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>
#include <stdio.h>
#include <stdlib.h>
#include "../sm100_ptx.cuh"
using namespace ofa;

static int encode(CUtensorMap* out, void* base, int rank, const uint64_t* dims, const uint64_t* strides, const uint32_t* box,
                  CUtensorMapSwizzle swz, CUtensorMapL2promotion promo = CU_TENSOR_MAP_L2_PROMOTION_L2_256B) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) return 1;
  auto fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  uint32_t es[5] = {1, 1, 1, 1, 1};
  return fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, rank, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
            promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS;
}

constexpr int STAGES = 6;
constexpr int STAGE_BYTES = 36864;

// each CTA: contiguous range of `total` tiles; tile t -> (c2 = t / per2, c1 = (t % per2) / per1 * step1, c0 = t % per1 * step0)
__global__ void __launch_bounds__(64, 1) walk(const __grid_constant__ CUtensorMap tm, int total, int per1, int per2, int step0,
                                              int step1, int nbox, int box_dx, uint32_t box_bytes) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t full[STAGES];
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) ptx::mbar_init(&full[s], 1);
    ptx::fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  const int t0 = (int)((long long)total * blockIdx.x / gridDim.x), t1 = (int)((long long)total * (blockIdx.x + 1) / gridDim.x);
  uint32_t ph[STAGES] = {0};
  int issued = 0;
  for (int t = t0; t < t1; ++t) {
    const int s = issued % STAGES;
    if (issued >= STAGES) { ptx::mbar_wait(&full[s], ph[s]); ph[s] ^= 1; }
    const int c2 = t / per2, r = t - c2 * per2;
    const int c1 = (r / per1) * step1, c0 = (r % per1) * step0;
    ptx::mbar_arrive_expect_tx(&full[s], box_bytes * nbox);
    for (int b = 0; b < nbox; ++b)
      ptx::tma_load_3d(smem + s * STAGE_BYTES + b * (STAGE_BYTES / 2), &tm, &full[s], c0 + b * box_dx, c1, c2);
    ++issued;
  }
  const int outstanding = issued < STAGES ? issued : STAGES;
  for (int k = 0; k < outstanding; ++k) {
    const int s = (issued - outstanding + k) % STAGES;
    ptx::mbar_wait(&full[s], ph[s]); ph[s] ^= 1;
  }
}

int main() {
  const int H = 540, W = 960, C = 384;
  const size_t bytes = (size_t)C * H * 1152 * 2;     // room for the strip layout (9 strips x 128 cols)
  uint8_t* buf;
  cudaMalloc(&buf, bytes);
  cudaMemset(buf, 1, bytes);
  uint8_t* flush;
  cudaMalloc(&flush, 512 << 20);
  const int smem = STAGES * STAGE_BYTES + 1024;
  cudaFuncSetAttribute(walk, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto run = [&](const char* name, const CUtensorMap& tm, int total, int per1, int per2, int step0, int step1, int nbox,
                 int box_dx, uint32_t box_bytes) {
    for (int rep = 0; rep < 3; ++rep) {
      cudaMemsetAsync(flush, rep, 512 << 20);
      cudaEventRecord(e0);
      walk<<<148, 64, smem>>>(tm, total, per1, per2, step0, step1, nbox, box_dx, box_bytes);
      cudaEventRecord(e1);
      cudaError_t err = cudaDeviceSynchronize();
      if (err != cudaSuccess) { printf("error %s\n", cudaGetErrorString(err)); exit(1); }
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (rep == 2) printf("%-62s %8.1f GB/s  (%.3f ms, %.0f MB)\n", name, (double)total * nbox * box_bytes / ms / 1e6, ms,
                           (double)total * nbox * box_bytes / 1e6);
    }
  };
  {  // A: image layout, 2 boxes of 64 cols x 134 rows
    CUtensorMap tm;
    uint64_t dims[3] = {(uint64_t)W, (uint64_t)H, (uint64_t)C};
    uint64_t str[2] = {(uint64_t)W * 2, (uint64_t)H * W * 2};
    uint32_t box[3] = {64, 134, 1};
    if (encode(&tm, buf, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
    run("A  dw window, image layout (row pitch 1920 B), 2 x [64 x 134]", tm, C * 5 * 9, 9, 45, 112, 128, 2, 64, 134 * 128);
  }
  {  // B: strip layout: plane = 9 strips of [540 rows x 128 cols]; window = rows of 256 B, pitch 256 B
    CUtensorMap tm;
    uint64_t dims[3] = {128, (uint64_t)H, (uint64_t)C * 9};
    uint64_t str[2] = {256, (uint64_t)H * 256};
    uint32_t box[3] = {64, 134, 1};
    if (encode(&tm, buf, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
    // tile t -> strip-plane index c2 = t / 5 ... walk rows within a strip first
    run("B  dw window, strip layout (row pitch 256 B), 2 x [64 x 134]", tm, C * 9 * 5, 1, 5, 0, 128, 2, 64, 134 * 128);
  }
  {  // B2: strip layout, one no-swizzle box of 128 cols x 134 rows
    CUtensorMap tm;
    uint64_t dims[3] = {128, (uint64_t)H, (uint64_t)C * 9};
    uint64_t str[2] = {256, (uint64_t)H * 256};
    uint32_t box[3] = {128, 134, 1};
    if (encode(&tm, buf, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_NONE)) return 1;
    run("B2 dw window, strip layout, 1 x [128 x 134] (no swizzle)", tm, C * 9 * 5, 1, 5, 0, 128, 1, 0, 134 * 256);
  }
  {  // C: project operand: [64 px x 64 planes] boxes over planes of H*W pixels
    CUtensorMap tm;
    uint64_t dims[3] = {(uint64_t)H * W, (uint64_t)C, 1};
    uint64_t str[2] = {(uint64_t)H * W * 2, (uint64_t)H * W * C * 2};
    uint32_t box[3] = {64, 64, 1};
    if (encode(&tm, buf, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
    // tile t: pixel block (t / 6) * 64... walk the 6 channel chunks of a 64-pixel block, then the next block
    run("C  project operand, [64 px x 64 planes], planes 1 MB apart", tm, (H * W / 64) * 6, H * W / 64, (H * W / 64) * 6, 64, 64,
        1, 0, 8192);
  }
  const uint64_t HW = (uint64_t)H * W;
  for (int promo = 0; promo < 3; ++promo) {
    const CUtensorMapL2promotion pr = promo == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE
                                                 : promo == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
    char name[128];
    {  // C again with each promotion
      CUtensorMap tm;
      uint64_t dims[3] = {HW, (uint64_t)C, 1};
      uint64_t str[2] = {HW * 2, HW * C * 2};
      uint32_t box[3] = {64, 64, 1};
      if (encode(&tm, buf, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B, pr)) return 1;
      snprintf(name, sizeof(name), "C  [64 px x 64 planes] SW128, promotion %d", promo == 0 ? 0 : promo == 1 ? 128 : 256);
      run(name, tm, (H * W / 64) * 6, H * W / 64, (H * W / 64) * 6, 64, 64, 1, 0, 8192);
    }
    {  // C2: 128 px x 32 planes (no swizzle): 256 B per plane
      CUtensorMap tm;
      uint64_t dims[3] = {HW, (uint64_t)C, 1};
      uint64_t str[2] = {HW * 2, HW * C * 2};
      uint32_t box[3] = {128, 32, 1};
      if (encode(&tm, buf, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_NONE, pr)) return 1;
      snprintf(name, sizeof(name), "C2 [128 px x 32 planes] no swizzle, promotion %d", promo == 0 ? 0 : promo == 1 ? 128 : 256);
      run(name, tm, (H * W / 128) * 12, H * W / 128, (H * W / 128) * 12, 128, 32, 1, 0, 8192);
    }
    {  // C3: 256 px x 16 planes: 512 B per plane
      CUtensorMap tm;
      uint64_t dims[3] = {HW, (uint64_t)C, 1};
      uint64_t str[2] = {HW * 2, HW * C * 2};
      uint32_t box[3] = {256, 16, 1};
      if (encode(&tm, buf, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_NONE, pr)) return 1;
      snprintf(name, sizeof(name), "C3 [256 px x 16 planes] no swizzle, promotion %d", promo == 0 ? 0 : promo == 1 ? 128 : 256);
      run(name, tm, (H * W / 256) * 24, H * W / 256, (H * W / 256) * 24, 256, 16, 1, 0, 8192);
    }
  }
  {  // C4: project's real order: for one 128-pixel tile all 6 channel chunks, 2 boxes each; tiles dealt round-robin
    CUtensorMap tm;
    uint64_t dims[3] = {HW, (uint64_t)C, 1};
    uint64_t str[2] = {HW * 2, HW * C * 2};
    uint32_t box[3] = {64, 64, 1};
    if (encode(&tm, buf, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
    // walk: t -> kc = t % 6 (c1 = kc * 64), pixel tile = t / 6 (c0 = tile * 128), 2 boxes 64 px apart
    run("C4 project order: per 128-px tile, 6 chunks x 2 boxes", tm, (H * W / 128) * 6, 6, 6 * (H * W / 128), 0, 0, 2, 64, 8192);
  }
  return 0;
}
