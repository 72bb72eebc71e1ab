// L2 vs HBM bandwidth through the TMA path (cp.async.bulk), as the planar MBConv stages would see it when their
// intermediates live in an L2-resident ring (round-2 design probe; not part of the library).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2_bw_probe l2_bw_probe.cu && ./l2_bw_probe
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory");
}
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_store(void* dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}

constexpr int CHUNK = 16384;
constexpr int STAGES = 8;

// mode 0: read only; 1: write only; 2: read chunk then write it to the second half (copy)
__global__ void __launch_bounds__(128, 1) probe(uint8_t* buf, size_t bytes, int passes, int mode) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[STAGES];
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  const size_t half = bytes / 2;
  const size_t span = (mode == 2) ? half : bytes;
  const size_t nchunks = span / CHUNK;
  uint32_t ph[STAGES] = {0};
  int issued = 0;
  for (int p = 0; p < passes; ++p) {
    for (size_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
      const int s = issued % STAGES;
      if (mode == 1) {
        if (issued >= STAGES) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(STAGES - 1) : "memory");
        bulk_store(buf + c * CHUNK, smem + s * CHUNK, CHUNK);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      } else {
        if (issued >= STAGES) {
          // stage s was filled STAGES issues ago: wait for it, (copy mode) store it out, then refill
          mbar_wait(&full[s], ph[s]); ph[s] ^= 1;
          if (mode == 2) {
            // the chunk that landed there
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          }
        }
        if (mode == 2 && issued >= STAGES) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(STAGES - 1) : "memory");
        mbar_expect(&full[s], CHUNK);
        bulk_load(smem + s * CHUNK, buf + c * CHUNK, CHUNK, &full[s]);
        if (mode == 2) {
          // write the PREVIOUS stage's data (already landed or not -- bandwidth probe only, data is irrelevant)
          bulk_store(buf + half + c * CHUNK, smem + ((s + 1) % STAGES) * CHUNK, CHUNK);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
      ++issued;
    }
  }
  if (mode != 1) {
    const int outstanding = issued < STAGES ? issued : STAGES;
    for (int k = 0; k < outstanding; ++k) {
      const int s = (issued - outstanding + k) % STAGES;
      mbar_wait(&full[s], ph[s]); ph[s] ^= 1;
    }
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const size_t maxb = (size_t)2 << 30;
  uint8_t* buf;
  cudaMalloc(&buf, maxb);
  cudaMemset(buf, 1, maxb);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, STAGES * CHUNK);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const char* names[3] = {"read ", "write", "copy "};
  for (int mode = 0; mode < 3; ++mode) {
    for (size_t mb : {8, 16, 24, 32, 48, 64, 96, 128, 256, 2048}) {
      const size_t bytes = mb << 20;
      const size_t target = (size_t)8 << 30;                 // bytes moved per measurement
      int passes = (int)(target / bytes); if (passes < 1) passes = 1;
      probe<<<sms, 128, STAGES * CHUNK>>>(buf, bytes, 2, mode);   // warm (fills L2)
      cudaEventRecord(e0);
      probe<<<sms, 128, STAGES * CHUNK>>>(buf, bytes, passes, mode);
      cudaEventRecord(e1);
      cudaError_t err = cudaDeviceSynchronize();
      if (err != cudaSuccess) { printf("error %s\n", cudaGetErrorString(err)); return 1; }
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      const double moved = (double)bytes * passes;           // copy mode: half read + half written = `bytes` per pass
      printf("%s working set %5zu MB : %8.1f GB/s (%d passes, %.3f ms)\n", names[mode], mb, moved / ms / 1e6, passes, ms);
    }
  }
  // is the L2 write (read) rate a per-SM or a chip-wide limit?  32 MB working set, fewer CTAs
  for (int mode = 0; mode < 2; ++mode) {
    for (int ctas : {8, 16, 32, 64, 100, sms}) {
      const size_t bytes = (size_t)32 << 20;
      const int passes = 64;
      probe<<<ctas, 128, STAGES * CHUNK>>>(buf, bytes, 2, mode);
      cudaEventRecord(e0);
      probe<<<ctas, 128, STAGES * CHUNK>>>(buf, bytes, passes, mode);
      cudaEventRecord(e1);
      cudaDeviceSynchronize();
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      const double moved = (double)bytes * passes;
      printf("%s 32 MB, %3d CTAs : %8.1f GB/s total, %6.1f GB/s per SM\n", names[mode], ctas, moved / ms / 1e6,
             moved / ms / 1e6 / ctas);
    }
  }
  return 0;
}
