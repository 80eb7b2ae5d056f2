// Micro-benchmark 2: fully unrolled 8-MMA GEMM groups (like the kernels' GEMMs), TF32 M=128.
// Variants: one thread issues (divergent context) vs the whole warp in uniform control flow with one elected lane.
#include "tc_ptx.cuh"
using namespace nlam;

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

template <int N, bool TS, bool UNIFORM>
__global__ void __launch_bounds__(128, 1) bench(int n_gemm, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  __shared__ uint32_t tmem_ptr;
  __shared__ __align__(8) unsigned long long bar;
  const int tid = threadIdx.x;
  const int warp = __shfl_sync(0xffffffff, tid >> 5, 0);
  for (int i = tid; i < 96 * 1024 / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 0.001f * (i & 255);
  if (warp == 0) {
    if (tid == 0) {
      mbar_init(smem_u32(&bar), 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffff, tmem_ptr, 0);
  const bool issuer = UNIFORM ? (warp == 0) : (tid == 0);
  if (issuer) {
    constexpr uint32_t idesc = umma_idesc_tf32(128, N);
    const uint64_t da = umma_desc(sbase);
    const uint64_t db = umma_desc(sbase + 65536);
    long long t0 = clock64();
    for (int g = 0; g < n_gemm; ++g) {
      const uint32_t dd = tmem + (uint32_t)((g & 1) * N);
      if (!UNIFORM || elect_one()) {
#pragma unroll
        for (int jj = 0; jj < 2; ++jj)
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            if (TS) umma_tf32_ts(dd, tmem + 256 + (uint32_t)(jj * 32 + kk * 8), db + (uint64_t)((jj * 8192) >> 4) + 2 * kk, idesc, (uint32_t)((jj | kk) != 0));
            else umma_tf32(dd, da + (uint64_t)((jj * 16384) >> 4) + 2 * kk, db + (uint64_t)((jj * 8192) >> 4) + 2 * kk, idesc, (uint32_t)((jj | kk) != 0));
          }
      }
    }
    long long t1 = clock64();
    if (!UNIFORM || elect_one()) umma_commit(smem_u32(&bar));
    if (UNIFORM) __syncwarp();
    mbar_wait(smem_u32(&bar), 0);
    long long t2 = clock64();
    if (blockIdx.x == 0 && tid == 0) {
      out[0] = t1 - t0;
      out[1] = t2 - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

template <int N, bool TS, bool UNIFORM>
void run(long long* d, const char* name) {
  cudaFuncSetAttribute(bench<N, TS, UNIFORM>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  const int n = 64;
  bench<N, TS, UNIFORM><<<148, 128, 96 * 1024>>>(n, d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("error %s (%s)\n", cudaGetErrorString(e), name);
    exit(1);
  }
  long long h[2];
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("%-28s N=%3d  issue %7.1f cycles/MMA   total %7.1f cycles/MMA  (%6.0f per 8-MMA GEMM)\n", name, N, (double)h[0] / (8 * n),
         (double)h[1] / (8 * n), (double)h[1] / n);
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  run<64, false, false>(d, "SS one-thread");
  run<64, true, false>(d, "TS one-thread");
  run<64, false, true>(d, "SS warp-uniform+elect");
  run<64, true, true>(d, "TS warp-uniform+elect");
  run<128, false, true>(d, "SS warp-uniform+elect");
  run<128, true, true>(d, "TS warp-uniform+elect");
  run<256, false, true>(d, "SS warp-uniform+elect");
  run<256, true, true>(d, "TS warp-uniform+elect");
  run<32, false, true>(d, "SS warp-uniform+elect");
  return 0;
}
