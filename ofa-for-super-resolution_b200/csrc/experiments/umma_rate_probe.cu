// Hardware probe (not part of the library): cost of ONE tcgen05.mma kind::f16 M x N x K16 with both operands in shared
// memory, issued back to back by one thread -- the depthwise kernel's unit of work is M128 x N32 x K16 (7/32 useful
// columns), 56 per tile.  Is it paced by the tensor pipe (M*N/256 clocks), by the operand fetch (bytes / 128 B/clk) or
// by a fixed per-instruction cost?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_rate_probe umma_rate_probe.cu && ./umma_rate_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include "../sm100_ptx.cuh"
using namespace ofa;

// mode 0: every MMA reads the same A chunk; mode 1: A walks through a 4-stage x 34 KB ring as the depthwise kernel does
// (dy row shifts + 32-byte K chunks), B = 1 KB no-swizzle tile (N = 32) or N/8 * 256 bytes in general
template <int M, int N>
__global__ void __launch_bounds__(128, 1) rate_kernel(long long* out, int reps, int mode) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;                         // 4 x 34816
  uint8_t* sB = smem + 4 * 34816;             // up to 8 KB
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + 16384);
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 1);
  for (int i = threadIdx.x; i < (4 * 34816 + 16384) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { ptx::mbar_init(bar, 1); ptx::fence_barrier_init(); }
  if (threadIdx.x < 32) { ptx::tmem_alloc(tptr, 512); ptx::tmem_relinquish(); }
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tptr;
  if (threadIdx.x == 0) {
    const uint32_t idesc = ptx::umma_idesc_f16(M, N, 0, 0, 0, 0);
    const uint32_t a0 = ptx::smem_u32(sA), b0 = ptx::smem_u32(sB);
    const uint64_t db = ptx::umma_desc(b0, 128, 256, 0);
    // warm-up
    for (int i = 0; i < 64; ++i) ptx::umma_bf16(tmem, ptx::umma_desc_sw128(a0, 1024), db, idesc, 1u);
    ptx::umma_commit(bar);
    ptx::mbar_wait(bar, 0);
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      const uint32_t stage = (uint32_t)(r & 3) * 34816u;
#pragma unroll
      for (int dy = 0; dy < 7; ++dy) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t off = mode ? stage + (uint32_t)((j >> 2) * 17408 + dy * 128 + (j & 3) * 32) : 0u;
          // mode 2: as mode 1 but consecutive MMAs hit DISJOINT accumulator columns (even chunks, then odd chunks)
          // mode 3: every MMA its own 32-column accumulator (8 independent ones)
          const int jj = mode == 2 ? ((j & 3) * 2 + (j >> 2)) : j;
          const uint32_t off2 = mode >= 2 ? stage + (uint32_t)((jj >> 2) * 17408 + dy * 128 + (jj & 3) * 32) : off;
          const uint32_t col = mode == 0 ? 0u : mode == 3 ? (uint32_t)(N * j) % 512u : (uint32_t)(16 * jj);
          ptx::umma_bf16(tmem + col, ptx::umma_desc_sw128(a0 + off2, 1024), db + (uint64_t)(mode ? ((dy * 1024) >> 4) : 0),
                         idesc, 1u);
        }
      }
    }
    const long long t1 = clock64();
    ptx::umma_commit(bar);
    ptx::mbar_wait(bar, 1);
    const long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  __syncthreads();
  if (threadIdx.x < 32) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem, 512); }
}

// tight issue: descriptors are loop-invariant registers, `nwarps` warps (1, 2 or 4) each issue reps * 56 / nwarps MMAs
template <int M, int N>
__global__ void __launch_bounds__(128, 1) tight_kernel(long long* out, int reps, int nwarps, int indep) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + 4 * 34816;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + 16384);
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 1);
  for (int i = threadIdx.x; i < (4 * 34816 + 16384) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { ptx::mbar_init(bar, nwarps); ptx::fence_barrier_init(); }
  if (threadIdx.x < 32) { ptx::tmem_alloc(tptr, 512); ptx::tmem_relinquish(); }
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tptr;
  const int warp = threadIdx.x >> 5;
  long long t0 = 0;
  uint64_t* bar2 = bar + 2;
  if (threadIdx.x == 0) {
    ptx::mbar_init(bar2, 1);
    ptx::fence_barrier_init();
    // zero-initialise all 512 columns: accumulate = 0 with ... a plain MMA per 128-column block, then subtract below
    const uint32_t idesc0 = ptx::umma_idesc_f16(M, 128, 0, 0, 0, 0);
    const uint64_t da = ptx::umma_desc_sw128(ptx::smem_u32(sA), 1024);
    const uint64_t db = ptx::umma_desc(ptx::smem_u32(sB), 128, 256, 0);
    for (int c = 0; c < 512; c += 128) ptx::umma_bf16(tmem + c, da, db, idesc0, 0u);
    ptx::umma_commit(bar2);
    ptx::mbar_wait(bar2, 0);
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  if (warp < nwarps && (threadIdx.x & 31) == 0) {
    const uint32_t idesc = ptx::umma_idesc_f16(M, N, 0, 0, 0, 0);
    const uint64_t da = ptx::umma_desc_sw128(ptx::smem_u32(sA) + warp * 34816, 1024);
    const uint64_t db = ptx::umma_desc(ptx::smem_u32(sB), 128, 256, 0);
    const uint32_t d = tmem + (indep ? (uint32_t)(warp * 128) : 0u);
    t0 = clock64();
    for (int r = 0; r < reps * 56 / nwarps / 8; ++r) {
#pragma unroll
      for (int u = 0; u < 8; ++u)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(d), "l"(da), "l"(db), "r"(idesc) : "memory");
    }
    ptx::umma_commit(bar);
  }
  if (threadIdx.x == 0) {
    ptx::mbar_wait(bar, 0);
    out[0] = clock64() - t0;
  }
  __syncthreads();
  ptx::tc_fence_after();
  // verify: every operand element is 1.0, so each MMA adds exactly 16 to its N columns (init added 16 as well)
  {
    const int per_warp = reps * 56 / nwarps / 8 * 8;
    long long bad = 0;
    for (int c = 0; c < N; c += 16) {
      uint32_t v[16];
      const int col = c;
      ptx::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)col, v);
      ptx::tmem_ld_wait();
      const float expect = indep ? (warp == 0 ? 16.f * (1 + per_warp) : -1.f) : 16.f * (1 + (float)per_warp * nwarps);
      if (expect > 0 && (M == 128 || (threadIdx.x & 31) < 16))
        for (int i = 0; i < 16; ++i) if (__uint_as_float(v[i]) != expect) ++bad;
    }
    if (bad && blockIdx.x == 0) atomicAdd((unsigned long long*)&out[1], (unsigned long long)bad);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem, 512); }
}

template <int M, int N>
void run_tight(long long* dout) {
  const int smem = 4 * 34816 + 16384 + 64 + 1024;
  cudaFuncSetAttribute(tight_kernel<M, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int indep = 0; indep < 2; ++indep)
    for (int nw : {1, 2, 4}) {
      const int reps = 200;
      cudaMemset(dout, 0, 16);
      tight_kernel<M, N><<<148, 128, smem>>>(dout, reps, nw, indep);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
      long long h[2];
      cudaMemcpy(h, dout, sizeof(h), cudaMemcpyDeviceToHost);
      printf("tight M %3d N %3d  %d issuing warp(s), %s accumulators : %.1f clk/MMA   wrong accumulator values (block 0): %lld\n",
             M, N, nw, indep ? "separate" : "one shared", (double)h[0] / (reps * 56), h[1]);
    }
}

template <int M, int N>
void run(long long* dout, int grid) {
  const int smem = 4 * 34816 + 16384 + 64 + 1024;
  cudaFuncSetAttribute(rate_kernel<M, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int mode = 0; mode < 4; ++mode) {
    const int reps = 200;
    rate_kernel<M, N><<<grid, 128, smem>>>(dout, reps, mode);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
    long long h[2];
    cudaMemcpy(h, dout, sizeof(h), cudaMemcpyDeviceToHost);
    printf("M %3d N %3d grid %3d mode %d : issue %.1f clk/MMA, complete %.1f clk/MMA  (tensor floor %d, operand bytes %d -> %d clk at 128 B/clk)\n",
           M, N, grid, mode, (double)h[0] / (reps * 56), (double)h[1] / (reps * 56), (M < 128 ? 128 : M) * N / 256,
           M * 32 + N * 32, (M * 32 + N * 32) / 128);
  }
}

int main() {
  long long* dout;
  cudaMalloc(&dout, 64);
  run_tight<128, 32>(dout);
  run_tight<128, 64>(dout);
  run_tight<128, 128>(dout);
  for (int grid : {148}) { if (grid) break;
    run<128, 16>(dout, grid);
    run<128, 32>(dout, grid);
    run<128, 64>(dout, grid);
    run<128, 128>(dout, grid);
    run<64, 32>(dout, grid);
  }
  return 0;
}
