// Micro-benchmark: cycles per tcgen05.mma kind::tf32 instruction (M=128, K=8) as a function of N,
// operand form (SS: A from shared memory, TS: A from TMEM) — one CTA per SM, back-to-back issue.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I neural_lam_b200/csrc -I include scripts/micro/mma_bench.cu -o gpurun_out/mma_bench
#include "tc_ptx.cuh"
using namespace nlam;

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

// whole warp runs the loop in uniform control flow; one elected lane issues (operands stay in uniform registers)
__global__ void __launch_bounds__(128, 1) bench_warp(int N, int ts_form, int n_mma, int M, int n_acc, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  __shared__ uint32_t tmem_ptr;
  __shared__ __align__(8) unsigned long long bar;
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffff, tid >> 5, 0);
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
  if (warp == 0) {
    const uint32_t idesc = umma_idesc_tf32(M, N);
    const uint64_t da = umma_desc(sbase);
    const uint64_t db = umma_desc(sbase + 32768);
    if (elect_one()) {
      for (int i = 0; i < 8; ++i) umma_tf32(tmem, da + 2 * (i & 3), db + 2 * (i & 3), idesc, 1u);
      umma_commit(smem_u32(&bar));
    }
    __syncwarp();
    mbar_wait(smem_u32(&bar), 0);
    tc_fence_after();
    long long t0 = clock64();
    int acc = 0;
    for (int i = 0; i < n_mma; ++i) {
      const uint32_t dd = tmem + (uint32_t)(acc * N);
      if (++acc == n_acc) acc = 0;
      if (elect_one()) {
        if (ts_form == 2) {
          const uint32_t id16 = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                       "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                       ::"r"(dd), "l"(da + 2 * (i & 3)), "l"(db + 2 * (i & 3)), "r"(id16), "r"(1u) : "memory");
        } else if (ts_form == 3) {
          const uint32_t id16 = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                       "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                       ::"r"(dd), "r"(tmem + 256 + 8 * (i & 7)), "l"(db + 2 * (i & 3)), "r"(id16), "r"(1u) : "memory");
        } else if (ts_form) umma_tf32_ts(dd, tmem + 256 + 8 * (i & 7), db + 2 * (i & 3), idesc, 1u);
        else umma_tf32(dd, da + 2 * (i & 3), db + 2 * (i & 3), idesc, 1u);
      }
    }
    long long t1 = clock64();
    if (elect_one()) umma_commit(smem_u32(&bar));
    __syncwarp();
    mbar_wait(smem_u32(&bar), 1);
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

__global__ void __launch_bounds__(128, 1) bench(int N, int ts_form, int n_mma, int M, int n_acc, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  __shared__ uint32_t tmem_ptr;
  __shared__ __align__(8) unsigned long long bar;
  const int tid = threadIdx.x, warp = tid >> 5;
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
  const uint32_t tmem = tmem_ptr;
  if (tid == 0) {
    const uint32_t idesc = umma_idesc_tf32(M, N);
    const uint64_t da = umma_desc(sbase);
    const uint64_t db = umma_desc(sbase + 32768);
    // warm
    for (int i = 0; i < 8; ++i) umma_tf32(tmem, da + 2 * (i & 3), db + 2 * (i & 3), idesc, 1u);
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    tc_fence_after();
    long long t0 = clock64();
    int acc = 0;
    for (int i = 0; i < n_mma; ++i) {
      const uint32_t dd = tmem + (uint32_t)(acc * N);
      if (++acc == n_acc) acc = 0;
      if (ts_form) umma_tf32_ts(dd, tmem + 256 + 8 * (i & 7), db + 2 * (i & 3), idesc, 1u);
      else umma_tf32(dd, da + 2 * (i & 3), db + 2 * (i & 3), idesc, 1u);
    }
    long long t1 = clock64();
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 1);
    long long t2 = clock64();
    if (blockIdx.x == 0) {
      out[0] = t1 - t0;
      out[1] = t2 - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  const int n = 512;
  printf("form  M   N  n_acc  issue_cycles/mma   total_cycles/mma\n");
  cudaFuncSetAttribute(bench_warp, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  for (int mode = 1; mode < 2; ++mode)
  for (int M : {128})
    for (int ts = 0; ts < 4; ++ts)
      for (int N : {32, 64, 128, 256})
       for (int n_acc : {1, 2, 4}) {
        if (ts && M == 64) continue;
        if (n_acc * N > 256) continue;
        if (mode) bench_warp<<<148, 128, 96 * 1024>>>(N, ts, n, M, n_acc, d);
        else bench<<<148, 128, 96 * 1024>>>(N, ts, n, M, n_acc, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
          printf("error %s (M=%d N=%d ts=%d)\n", cudaGetErrorString(e), M, N, ts);
          return 1;
        }
        long long h[2];
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("%s %s  %3d %3d %3d   %8.1f   %8.1f\n", mode ? "warp-uniform" : "one-thread  ", ts == 0 ? "tf32 SS" : ts == 1 ? "tf32 TS" : ts == 2 ? "f16 SS" : "f16 TS", M, N, n_acc, (double)h[0] / n, (double)h[1] / n);
      }
  return 0;
}
