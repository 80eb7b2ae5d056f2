// Generic tcgen05 Linear kernel for hidden widths the fused H = 64 kernels do not cover (H = 128 / 256: BASELINE
// configs 3-5), and for the layer variants they do not cover at H = 64 (PropagationNet):
//
//     Y[r, :] = epi( [X0 | X1][r, :] · Wᵀ + bias + A0[idx0[r], :] + A1[idx1[r], :] )
//     epi: none | SiLU | LayerNorm (two passes over the TMEM accumulator), then "+ post[idxp[r]]" (PropagationNet's
//     x_j + edge_mlp(...)), a second output before the residual (the message m next to e' = e + m) and "+ residual".
//
// One CTA tile = 128 rows x N (N = the full output width <= 256 = one UMMA instruction shape M128 x N), K streamed
// in 32-column chunks through a ring of shared-memory stages: the X chunk (16 KB) and the W chunk (N x 128 B, L2
// resident) arrive by TMA, `tcgen05.mma kind::tf32` accumulates into one of TWO TMEM accumulator buffers (2 x N <= 512
// columns), so that the epilogue of tile i (128 threads, thread = row, `tcgen05.ld` 32 columns at a time) overlaps the
// GEMM of tile i + 1.  Warp roles: 0 = TMA producer, 1 = MMA issue, 2-5 = epilogue.
// An MLP (reference utils/networks.py:27-40) is two launches with the hidden activations in a workspace; an
// InteractionNet (reference gnn_layers.py:110-157) is: node projections of the split first edge Linear, edge layer 1
// (gathered projections added in the epilogue, SiLU), edge layer 2 (LayerNorm, m and e' = e + m), CSR segment sum,
// node layers 1 / 2.  At H = 256 the path is tensor-pipe bound, which is what this kernel is for; the
// fully fused H = 64 kernels (tc3-tc6) stay the path of the BASELINE metric.
#include "tc_ptx.cuh"

namespace nlam {

namespace g7 {
constexpr int THREADS = 192;
constexpr int KC = 32;                 // K columns per chunk (128 bytes: one swizzle row)
constexpr uint32_t A_BYTES = 128 * 128;  // 128 rows x 128 B
}  // namespace g7

struct LinParams {
  int n_tiles;       // row tiles per batch
  int B;
  long long n_rows;
  int kc0, kc1;      // chunks of source 0 / source 1
  int a_batched[2];
  int w_batched;
  int n_real;        // real output width (<= N; < N only for direct narrow stores)
  const float* bias;
  int act;           // 0 none, 1 SiLU
  const float* gamma;
  const float* beta;
  float eps;
  // pre-activation gathered addends (rows of width N)
  const float* add[2];
  const int32_t* add_idx[2];
  long long add_bs[2];
  // post-epilogue gathered addend (PropagationNet)
  const float* post;
  const int32_t* post_idx;
  long long post_bs;
  // residual (row r of the residual tensor; batch stride 0 = broadcast)
  const float* res;
  long long res_bs;
  float* out;        // (B, n_rows, n_real)
  float* out2;       // optional: value before the residual
  int stages;
};

template <int N>
__global__ void __launch_bounds__(g7::THREADS, 1)
tc_linear_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                 const __grid_constant__ CUtensorMap tmW, const LinParams p) {
  using namespace g7;
  constexpr uint32_t W_BYTES = N * 128;
  constexpr uint32_t STG_BYTES = A_BYTES + W_BYTES;
  constexpr uint32_t TM_COLS = (2 * N <= 32) ? 32 : (2 * N <= 64) ? 64 : (2 * N <= 128) ? 128 : (2 * N <= 256) ? 256 : 512;
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int NSTG = p.stages;
  const uint32_t off_misc = (uint32_t)NSTG * STG_BYTES;
  const uint32_t mb = sbase + off_misc;
  const uint32_t bar_full = mb;             // [8]
  const uint32_t bar_empty = mb + 64;       // [8]
  const uint32_t bar_acc_full = mb + 128;   // [2]
  const uint32_t bar_acc_empty = mb + 144;  // [2] 128 arrivals
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + off_misc + 160);
  float* sprm = reinterpret_cast<float*>(smem + off_misc + 256);  // bias | gamma | beta (N each)
  if ((sbase & 1023u) != 0) {
    if (tid == 0) printf("nlam tc_linear: dynamic shared memory not 1024-byte aligned\n");
    __trap();
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < NSTG; ++s) {
        mbar_init(bar_full + 8 * s, 1);
        mbar_init(bar_empty + 8 * s, 1);
      }
      for (int s = 0; s < 2; ++s) {
        mbar_init(bar_acc_full + 8 * s, 1);
        mbar_init(bar_acc_empty + 8 * s, 128);
      }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                 "r"(TM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA0) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA1) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
  }
  for (int i = tid; i < N; i += THREADS) {
    sprm[i] = (p.bias && i < p.n_real) ? p.bias[i] : 0.f;
    sprm[N + i] = p.gamma ? p.gamma[i] : 1.f;
    sprm[2 * N + i] = p.beta ? p.beta[i] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr_smem, 0);
  const long long n_work = (long long)p.n_tiles * p.B;
  const int kc_total = p.kc0 + p.kc1;

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      const uint64_t pol_stream = policy_evict_first();
      const uint64_t pol_keep = policy_evict_last();
      long long c = 0;  // chunk counter over the whole kernel
      for (long long w = blockIdx.x; w < n_work; w += gridDim.x) {
        const int b = (int)(w / p.n_tiles), tile = (int)(w - (long long)b * p.n_tiles);
        for (int kc = 0; kc < kc_total; ++kc, ++c) {
          const int s = (int)(c % NSTG);
          mbar_wait(bar_empty + 8 * s, (uint32_t)(((c / NSTG) & 1) ^ 1));
          const uint32_t full = bar_full + 8 * s;
          mbar_expect_tx(full, STG_BYTES);
          const uint32_t sa = sbase + s * STG_BYTES;
          if (kc < p.kc0)
            tma_load_3d(sa, &tmA0, full, kc * KC, tile * 128, p.a_batched[0] ? b : 0, pol_stream);
          else
            tma_load_3d(sa, &tmA1, full, (kc - p.kc0) * KC, tile * 128, p.a_batched[1] ? b : 0, pol_stream);
          tma_load_3d(sa + A_BYTES, &tmW, full, kc * KC, 0, p.w_batched ? b : 0, pol_keep);
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issue ===============================
    const uint32_t idesc = umma_idesc_tf32(128, N);
    long long c = 0;
    int tl = 0;  // local tile counter
    for (long long w = blockIdx.x; w < n_work; w += gridDim.x, ++tl) {
      const int buf = tl & 1;
      if (lane == 0) mbar_wait(bar_acc_empty + 8 * buf, (uint32_t)(((tl >> 1) & 1) ^ 1));
      __syncwarp();
      tc_fence_after();
      const uint32_t dd = tmem_base + buf * N;
      for (int kc = 0; kc < kc_total; ++kc, ++c) {
        const int s = (int)(c % NSTG);
        if (lane == 0) mbar_wait(bar_full + 8 * s, (uint32_t)((c / NSTG) & 1));
        __syncwarp();
        tc_fence_after();
        const uint64_t da = umma_desc(sbase + s * STG_BYTES);
        const uint64_t db = umma_desc(sbase + s * STG_BYTES + A_BYTES);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_tf32(dd, da + 2 * k, db + 2 * k, idesc, (uint32_t)((kc | k) != 0));
          umma_commit(bar_empty + 8 * s);
          if (kc == kc_total - 1) umma_commit(bar_acc_full + 8 * buf);
        }
        __syncwarp();
      }
    }
  } else {
    // =============================== epilogue (thread = row) ===============================
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;
    const uint32_t t_lane = ((uint32_t)(q * 32)) << 16;
    int tl = 0;
    for (long long w = blockIdx.x; w < n_work; w += gridDim.x, ++tl) {
      const int b = (int)(w / p.n_tiles), tile = (int)(w - (long long)b * p.n_tiles);
      const int buf = tl & 1;
      const long long r = (long long)tile * 128 + row;
      const bool valid = r < p.n_rows;
      // row bases of the optional epilogue operands (loads issued before the accumulator wait)
      const float* a0 = nullptr;
      const float* a1 = nullptr;
      const float* pp = nullptr;
      const float* rr = nullptr;
      if (valid) {
        if (p.add[0]) a0 = p.add[0] + (long long)b * p.add_bs[0] + (long long)(p.add_idx[0] ? __ldg(p.add_idx[0] + r) : r) * N;
        if (p.add[1]) a1 = p.add[1] + (long long)b * p.add_bs[1] + (long long)(p.add_idx[1] ? __ldg(p.add_idx[1] + r) : r) * N;
        if (p.post) pp = p.post + (long long)b * p.post_bs + (long long)(p.post_idx ? __ldg(p.post_idx + r) : r) * N;
        if (p.res) rr = p.res + (long long)b * p.res_bs + r * N;
      }
      if (lane == 0) mbar_wait(bar_acc_full + 8 * buf, (uint32_t)((tl >> 1) & 1));
      __syncwarp();
      tc_fence_after();
      const uint32_t acc = tmem_base + buf * N + t_lane;
      float mu = 0.f, rstd = 1.f;
      if (p.gamma) {  // LayerNorm statistics: first pass over the accumulator row
        float sm = 0.f, sq = 0.f;
        for (int c = 0; c < N / 32; ++c) {
          float v[32];
          tmem_ld32(acc + 32 * c, v);
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float x = v[i] + sprm[32 * c + i];
            sm += x;
            sq = fmaf(x, x, sq);
          }
        }
        mu = sm * (1.0f / N);
        rstd = rsqrtf(fmaxf(sq * (1.0f / N) - mu * mu, 0.f) + p.eps);
      }
      float* orow = p.out + ((long long)b * p.n_rows + r) * p.n_real;
      float* o2row = p.out2 ? p.out2 + ((long long)b * p.n_rows + r) * N : nullptr;
      for (int c = 0; c < N / 32; ++c) {
        float v[32];
        tmem_ld32(acc + 32 * c, v);
        if (c == N / 32 - 1) {  // the accumulator buffer is in registers: hand it back to the MMA warp
          tc_fence_before();
          mbar_arrive(bar_acc_empty + 8 * buf);
        }
        if (!valid) continue;
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] += sprm[32 * c + i];
        if (a0) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 x = __ldg(reinterpret_cast<const float4*>(a0 + 32 * c) + i);
            v[4 * i] += x.x; v[4 * i + 1] += x.y; v[4 * i + 2] += x.z; v[4 * i + 3] += x.w;
          }
        }
        if (a1) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 x = __ldg(reinterpret_cast<const float4*>(a1 + 32 * c) + i);
            v[4 * i] += x.x; v[4 * i + 1] += x.y; v[4 * i + 2] += x.z; v[4 * i + 3] += x.w;
          }
        }
        if (p.act) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = silu_fast(v[i]);
        }
        if (p.gamma) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = fmaf((v[i] - mu) * rstd, sprm[N + 32 * c + i], sprm[2 * N + 32 * c + i]);
        }
        if (pp) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 x = __ldg(reinterpret_cast<const float4*>(pp + 32 * c) + i);
            v[4 * i] += x.x; v[4 * i + 1] += x.y; v[4 * i + 2] += x.z; v[4 * i + 3] += x.w;
          }
        }
        if (o2row) {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            reinterpret_cast<float4*>(o2row + 32 * c)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        }
        if (rr) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 x = __ldg(reinterpret_cast<const float4*>(rr + 32 * c) + i);
            v[4 * i] += x.x; v[4 * i + 1] += x.y; v[4 * i + 2] += x.z; v[4 * i + 3] += x.w;
          }
        }
        if (p.n_real == N) {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            reinterpret_cast<float4*>(orow + 32 * c)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (32 * c + i < p.n_real) orow[32 * c + i] = v[i];
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------------------ host
static int npad(int n) { return n <= 32 ? 32 : n <= 64 ? 64 : n <= 128 ? 128 : 256; }

bool tc_linear_shape_ok(int k0, int k1, int n_out) {
  if (n_out < 1 || n_out > 256) return false;
  if (k0 < 32 || k0 % 32 != 0 || k1 % 32 != 0 || (long long)k0 + k1 > (1LL << 26)) return false;
  return true;
}

int tc_linear(const LinearCall& c, cudaStream_t st) {
  using namespace g7;
  const int N = npad(c.n_out);
  NLAM_REQUIRE(tc_linear_shape_ok(c.k0, c.k1, c.n_out), NLAM_E_UNSUPPORTED, "tc_linear: unsupported shape K=%d+%d N=%d", c.k0,
               c.k1, c.n_out);
  NLAM_REQUIRE(c.n_out == N || (!c.out2 && !c.res && !c.post && !c.add[0] && !c.add[1] && !c.gamma), NLAM_E_UNSUPPORTED,
               "tc_linear: narrow outputs support bias / activation only");
  NLAM_REQUIRE(aligned16(c.x0) && aligned16(c.w) && (!c.x1 || aligned16(c.x1)) && aligned16(c.out) && c.ldw % 4 == 0,
               NLAM_E_INVALID, "tc_linear: pointers must be 16-byte aligned");
  CUtensorMap ma0, ma1, mw;
  const bool b0 = c.x0_bs != 0 && c.B > 1, b1 = c.x1 && c.x1_bs != 0 && c.B > 1;
  const uint64_t pitch0 = c.x0_pitch ? (uint64_t)c.x0_pitch : (uint64_t)c.k0;
  NLAM_REQUIRE(pitch0 % 4 == 0 && c.x0_bs % 4 == 0 && c.w_bs % 4 == 0, NLAM_E_INVALID, "tc_linear: pitches must be multiples of 4");
  int rc = make_map(&ma0, c.x0, (uint64_t)c.k0, (uint64_t)c.n_rows, b0 ? (uint64_t)c.B : 1, pitch0,
                    b0 ? (uint64_t)c.x0_bs : (uint64_t)c.n_rows * pitch0, 128, true);
  if (rc) return rc;
  if (c.x1) {
    rc = make_map(&ma1, c.x1, (uint64_t)c.k1, (uint64_t)c.n_rows, b1 ? (uint64_t)c.B : 1, (uint64_t)c.k1,
                  b1 ? (uint64_t)c.x1_bs : (uint64_t)c.n_rows * c.k1, 128, true);
    if (rc) return rc;
  } else {
    ma1 = ma0;
  }
  // W: (n_out rows, k0 + k1 columns) slice of a row-major matrix with row pitch ldw; rows past n_out read as zero
  const bool wb = c.w_bs != 0 && c.B > 1;
  rc = make_map(&mw, c.w, (uint64_t)(c.w_cols ? c.w_cols : c.k0 + c.k1), (uint64_t)c.n_out, wb ? (uint64_t)c.B : 1, (uint64_t)c.ldw,
                wb ? (uint64_t)c.w_bs : (uint64_t)c.n_out * c.ldw, (uint32_t)N, true);
  if (rc) return rc;
  LinParams p;
  memset(&p, 0, sizeof(p));
  p.n_tiles = (int)((c.n_rows + 127) / 128);
  p.B = c.B;
  p.n_rows = c.n_rows;
  p.kc0 = c.k0 / KC;
  p.kc1 = c.x1 ? c.k1 / KC : 0;
  p.a_batched[0] = b0;
  p.a_batched[1] = b1;
  p.w_batched = wb;
  p.n_real = c.n_out;
  p.bias = c.bias;
  p.act = c.act;
  p.gamma = c.gamma;
  p.beta = c.beta;
  p.eps = c.eps;
  for (int i = 0; i < 2; ++i) {
    p.add[i] = c.add[i];
    p.add_idx[i] = c.add_idx[i];
    p.add_bs[i] = c.B > 1 ? c.add_bs[i] : 0;
  }
  p.post = c.post;
  p.post_idx = c.post_idx;
  p.post_bs = c.B > 1 ? c.post_bs : 0;
  p.res = c.res;
  p.res_bs = c.B > 1 ? c.res_bs : 0;
  p.out = c.out;
  p.out2 = c.out2;
  const uint32_t stg = A_BYTES + (uint32_t)N * 128;
  p.stages = (int)std::min<uint32_t>(8, (220 * 1024 - 4096) / stg);
  const size_t smem = (size_t)p.stages * stg + 256 + 3 * N * sizeof(float) + 64;
  const long long n_work = (long long)p.n_tiles * p.B;
  NLAM_REQUIRE(n_work >= 1 && n_work < (1LL << 40), NLAM_E_INVALID, "tc_linear: bad work size");
  const int grid = (int)std::min<long long>(n_work, num_sms());
  double bytes = 4.0 * c.n_rows * ((double)c.k0 * (b0 ? c.B : 1) + (c.x1 ? (double)c.k1 * (b1 ? c.B : 1) : 0.0)) +
                 4.0 * c.n_rows * c.B * c.n_out * (c.out2 ? 2 : 1) + 4.0 * c.n_out * (c.k0 + c.k1);
  for (int i = 0; i < 2; ++i)
    if (c.add[i]) bytes += 4.0 * c.n_rows * c.B * N;
  if (c.res) bytes += 4.0 * c.n_rows * N * (c.res_bs != 0 ? c.B : 1);
  if (c.post) bytes += 4.0 * c.n_rows * c.B * N;
  int dev = 0;
  NLAM_CUDA_OK(cudaGetDevice(&dev));
#define NLAM_LAUNCH_LIN(NN)                                                                                          \
  {                                                                                                                  \
    static unsigned mask = 0;                                                                                        \
    if (!(mask & (1u << (dev & 31)))) {                                                                              \
      NLAM_CUDA_OK(cudaFuncSetAttribute(tc_linear_kernel<NN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024)); \
      mask |= 1u << (dev & 31);                                                                                      \
    }                                                                                                                \
    ProfScope ps("tc_linear_kernel<" #NN ">", st, bytes);                                                            \
    tc_linear_kernel<NN><<<grid, THREADS, smem, st>>>(ma0, ma1, mw, p);                                              \
  }
  if (N == 32) NLAM_LAUNCH_LIN(32)
  else if (N == 64) NLAM_LAUNCH_LIN(64)
  else if (N == 128) NLAM_LAUNCH_LIN(128)
  else NLAM_LAUNCH_LIN(256)
#undef NLAM_LAUNCH_LIN
  count_launch();
  NLAM_CUDA_OK(cudaGetLastError());
  return NLAM_OK;
}

// Two-Linear MLP (Linear, SiLU, Linear[, LayerNorm]) over up to two dense row blocks, residual optional: two launches,
// hidden activations in `ws` (n_rows * B * hidden floats).
bool tc_mlp2_supported(const NlamMlp* m, const NlamRowSrc* srcs, int n_src, const NlamRowSrc* res, const NlamRowSrc* res2) {
  if (m->n_linear != 2 || res2) return false;
  const int H = m->out_dim[0], no = m->out_dim[1];
  if (!(H == 128 || H == 256) || no < 1 || no > 256) return false;
  if (m->ln_gamma && no != npad(no)) return false;
  if (n_src < 1 || n_src > 2) return false;
  int k = 0;
  for (int s = 0; s < n_src; ++s) {
    if (srcs[s].idx || srcs[s].dim % 32 != 0 || srcs[s].dim < 32 || !aligned16(srcs[s].ptr) || srcs[s].bstride % 4 != 0) return false;
    k += srcs[s].dim;
  }
  if (k != m->in_dim || k > 1024) return false;
  if (res && (res->idx || res->dim != no || no != npad(no))) return false;
  return aligned16(m->w[0]) && aligned16(m->w[1]);
}

int tc_mlp2(const NlamMlp* m, const NlamRowSrc* srcs, int n_src, const NlamRowSrc* res, float* out, int64_t n_rows, int B,
            cudaStream_t st, float* ws) {
  const int H = m->out_dim[0], no = m->out_dim[1];
  LinearCall c;
  memset(&c, 0, sizeof(c));
  c.x0 = srcs[0].ptr;
  c.x0_bs = srcs[0].bstride;
  c.k0 = srcs[0].dim;
  if (n_src == 2) {
    c.x1 = srcs[1].ptr;
    c.x1_bs = srcs[1].bstride;
    c.k1 = srcs[1].dim;
  }
  // batch-broadcast inputs: the hidden layer is computed once and broadcast
  bool any_batched = false;
  for (int s = 0; s < n_src; ++s) any_batched |= (srcs[s].bstride != 0 && B > 1);
  const int Bh = any_batched ? B : 1;
  c.w = m->w[0];
  c.ldw = m->in_dim;
  c.bias = m->b[0];
  c.n_out = H;
  c.act = 1;
  c.n_rows = n_rows;
  c.B = Bh;
  c.out = ws;
  int rc = tc_linear(c, st);
  if (rc) return rc;
  LinearCall d;
  memset(&d, 0, sizeof(d));
  d.x0 = ws;
  d.x0_bs = Bh > 1 ? (int64_t)n_rows * H : 0;
  d.k0 = H;
  d.w = m->w[1];
  d.ldw = H;
  d.bias = m->b[1];
  d.n_out = no;
  d.gamma = m->ln_gamma;
  d.beta = m->ln_beta;
  d.eps = m->ln_eps;
  if (res) {
    d.res = res->ptr;
    d.res_bs = res->bstride;
  }
  d.n_rows = n_rows;
  d.B = B;
  d.out = out;
  return tc_linear(d, st);
}

// ---- narrow / concatenated inputs (grid embedder: prev | prev_prev | forcing | static, reference graph/base.py:275-283)
// at H = 128 / 256: the sources are packed into one zero-padded dense (rows, Kp) block first, then the two Linear launches
// (a source with a row index gathers: row r of the block is row idx[r] of the source — the [e | x_src | x_dst] rows of an
// edge MLP whose per-chunk weights rule out the fused edge kernels: SplitMLPs, reference gnn_layers.py:274-324)
struct PackIdx {
  const int32_t* i[4];
};
__global__ void pack_rows_kernel(const float* s0, const float* s1, const float* s2, const float* s3, int d0, int d1, int d2,
                                 int d3, long long bs0, long long bs1, long long bs2, long long bs3, float* out, int kp,
                                 long long n_rows, int B, PackIdx ix) {
  const long long total = (long long)B * n_rows * kp;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % kp);
    const long long r = (i / kp) % n_rows;
    const int b = (int)(i / ((long long)kp * n_rows));
    float v = 0.f;
    int cc = c;
    if (cc < d0) v = s0[b * bs0 + (ix.i[0] ? (long long)ix.i[0][r] : r) * d0 + cc];
    else if ((cc -= d0) < d1) v = s1[b * bs1 + (ix.i[1] ? (long long)ix.i[1][r] : r) * d1 + cc];
    else if ((cc -= d1) < d2) v = s2[b * bs2 + (ix.i[2] ? (long long)ix.i[2][r] : r) * d2 + cc];
    else if ((cc -= d2) < d3) v = s3[b * bs3 + (ix.i[3] ? (long long)ix.i[3][r] : r) * d3 + cc];
    out[i] = v;
  }
}

bool tc_mlp2_packed_supported(const NlamMlp* m, const NlamRowSrc* srcs, int n_src, const NlamRowSrc* res, const NlamRowSrc* res2) {
  if (m->n_linear != 2 || res2 || n_src < 1 || n_src > 4) return false;
  const int H = m->out_dim[0], no = m->out_dim[1];
  if (!(H == 64 || H == 128 || H == 256) || no < 1 || no > 256 || (m->ln_gamma && no != npad(no))) return false;
  if (res && (res->dim != no || no != npad(no) || res->bstride % 4 != 0 || !aligned16(res->ptr))) return false;
  int k = 0;
  for (int s = 0; s < n_src; ++s) k += srcs[s].dim;  // sources may gather (row index)
  return k == m->in_dim && k % 4 == 0 && k <= 1024 && aligned16(m->w[0]) && aligned16(m->w[1]);
}

size_t tc_mlp2_packed_workspace_floats(const NlamMlp* m, int64_t n_rows, int B) {
  const int kp = (m->in_dim + 31) / 32 * 32;
  return (size_t)n_rows * B * (kp + m->out_dim[0]);
}

int tc_mlp2_packed(const NlamMlp* m, const NlamRowSrc* srcs, int n_src, const NlamRowSrc* res, float* out, int64_t n_rows, int B,
                   cudaStream_t st, float* ws) {
  const int kp = (m->in_dim + 31) / 32 * 32, H = m->out_dim[0], no = m->out_dim[1];
  float* packed = ws;
  float* hid = ws + (size_t)n_rows * B * kp;
  const float* sp[4] = {nullptr, nullptr, nullptr, nullptr};
  int d[4] = {0, 0, 0, 0};
  long long bs[4] = {0, 0, 0, 0};
  PackIdx ix = {{nullptr, nullptr, nullptr, nullptr}};
  for (int s = 0; s < n_src; ++s) {
    sp[s] = srcs[s].ptr;
    d[s] = srcs[s].dim;
    bs[s] = B > 1 ? srcs[s].bstride : 0;
    ix.i[s] = srcs[s].idx;
  }
  const long long total = (long long)B * n_rows * kp;
  {
    ProfScope ps("pack_rows_kernel", st, 4.0 * total + 4.0 * (double)B * n_rows * m->in_dim);
    pack_rows_kernel<<<(int)std::min<long long>((total + 255) / 256, 148 * 16), 256, 0, st>>>(
        sp[0], sp[1], sp[2], sp[3], d[0], d[1], d[2], d[3], bs[0], bs[1], bs[2], bs[3], packed, kp, n_rows, B, ix);
  }
  count_launch();
  NLAM_CUDA_OK(cudaGetLastError());
  LinearCall c;
  memset(&c, 0, sizeof(c));
  c.x0 = packed; c.x0_bs = (int64_t)n_rows * kp; c.k0 = kp; c.w = m->w[0]; c.ldw = m->in_dim; c.w_cols = m->in_dim;
  c.bias = m->b[0]; c.n_out = H; c.act = 1; c.n_rows = n_rows; c.B = B; c.out = hid;
  int rc = tc_linear(c, st);
  if (rc) return rc;
  memset(&c, 0, sizeof(c));
  c.x0 = hid; c.x0_bs = (int64_t)n_rows * H; c.k0 = H; c.w = m->w[1]; c.ldw = H; c.bias = m->b[1]; c.n_out = no;
  c.gamma = m->ln_gamma; c.beta = m->ln_beta; c.eps = m->ln_eps; c.n_rows = n_rows; c.B = B; c.out = out;
  if (res && res->idx) {  // gathered residual (PropagationNet: + x_src)
    c.post = res->ptr;
    c.post_idx = res->idx;
    c.post_bs = B > 1 ? res->bstride : 0;
  } else if (res) {
    c.res = res->ptr;
    c.res_bs = B > 1 ? res->bstride : 0;
  }
  return tc_linear(c, st);
}

}  // namespace nlam
// zero-padded concatenation of up to four row blocks: out[b, r, :kp] = [s0 | s1 | s2 | s3 | 0...]
extern "C" int nlam_pack_rows(const float* s0, const float* s1, const float* s2, const float* s3, int d0, int d1, int d2, int d3,
                              int64_t bs0, int64_t bs1, int64_t bs2, int64_t bs3, float* out, int kp, int64_t n_rows, int B,
                              void* stream) {
  NLAM_REQUIRE(s0 && out && kp >= d0 + d1 + d2 + d3 && n_rows >= 0 && B >= 1, NLAM_E_INVALID, "nlam_pack_rows: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const long long total = (long long)B * n_rows * kp;
  if (total == 0) return NLAM_OK;
  {
    nlam::ProfScope ps("pack_rows_kernel", st, 8.0 * total);
    nlam::pack_rows_kernel<<<(int)std::min<long long>((total + 255) / 256, 148 * 16), 256, 0, st>>>(
        s0, s1, s2, s3, d0, d1, d2, d3, bs0, bs1, bs2, bs3, out, kp, n_rows, B, nlam::PackIdx{{nullptr, nullptr, nullptr, nullptr}});
  }
  nlam::count_launch();
  NLAM_CUDA_OK(cudaGetLastError());
  return NLAM_OK;
}
namespace nlam {

// ---- InteractionNet / PropagationNet (reference gnn_layers.py:110-157, :231-249) from the generic Linear kernel ----
bool tc_inet_gen_supported(const NlamGraph* g, const NlamMlp* em, const NlamMlp* am, int flags, const float* send,
                           int64_t send_bs, const float* rec, int64_t rec_bs, const float* edge, int64_t edge_bs) {
  if (!g || em->n_linear != 2 || am->n_linear != 2 || !em->ln_gamma || !am->ln_gamma) return false;
  const int H = em->out_dim[1];
  const bool prop = flags & NLAM_PROPAGATION;
  if (!(H == 128 || H == 256 || (H == 64 && prop))) return false;
  if (em->out_dim[0] != H || am->out_dim[0] != H || am->out_dim[1] != H || em->in_dim != 3 * H || am->in_dim != 2 * H) return false;
  if (!(aligned16(send) && aligned16(rec) && aligned16(edge) && send_bs % 4 == 0 && rec_bs % 4 == 0 && edge_bs % 4 == 0))
    return false;
  for (int l = 0; l < 2; ++l)
    if (!aligned16(em->w[l]) || !aligned16(am->w[l])) return false;
  return true;
}

static size_t rup64(size_t n) { return (n + 63) / 64 * 64; }

size_t tc_inet_gen_workspace_floats(const NlamGraph* g, int B, int H) {
  // Ps | Pr | edge hidden | messages | node hidden
  return rup64((size_t)B * g->n_send * H) + rup64((size_t)B * g->n_rec * H) + 2 * rup64((size_t)B * g->n_edges * H) +
         rup64((size_t)B * g->n_rec * H);
}

int tc_inet_gen(const NlamGraph* g, const NlamMlp* em, const NlamMlp* am, const float* send, int64_t send_bs, const float* rec,
                int64_t rec_bs, const float* edge, int64_t edge_bs, float* rec_out, float* edge_out, float* aggr, int B, int flags,
                float* ws, cudaStream_t st) {
  const int H = em->out_dim[1];
  const bool prop = flags & NLAM_PROPAGATION;
  const int mean = (flags & (NLAM_AGGR_MEAN | NLAM_PROPAGATION)) ? 1 : 0;
  const int64_t ns = g->n_send, nr = g->n_rec, E = g->n_edges;
  const int Bs = (send_bs == 0 || B == 1) ? 1 : B, Br = (rec_bs == 0 || B == 1) ? 1 : B;
  float* Ps = ws;
  float* Pr = Ps + rup64((size_t)B * ns * H);
  float* hid = Pr + rup64((size_t)B * nr * H);
  float* msg = hid + rup64((size_t)B * E * H);
  float* nhid = msg + rup64((size_t)B * E * H);
  const float* w1 = em->w[0];  // (H, 3H): columns [e | sender | receiver]
  LinearCall c;
  // 1 / 2: node projections of the split first Linear
  memset(&c, 0, sizeof(c));
  c.x0 = send; c.x0_bs = send_bs; c.k0 = H; c.w = w1 + H; c.ldw = 3 * H; c.n_out = H; c.n_rows = ns; c.B = Bs; c.out = Ps;
  int rc = tc_linear(c, st);
  if (rc) return rc;
  memset(&c, 0, sizeof(c));
  c.x0 = rec; c.x0_bs = rec_bs; c.k0 = H; c.w = w1 + 2 * H; c.ldw = 3 * H; c.bias = em->b[0]; c.n_out = H; c.n_rows = nr; c.B = Br;
  c.out = Pr;
  rc = tc_linear(c, st);
  if (rc) return rc;
  // 3: edge layer 1: SiLU(W1e·e + Ps[src] + Pr[dst])
  memset(&c, 0, sizeof(c));
  c.x0 = edge; c.x0_bs = edge_bs; c.k0 = H; c.w = w1; c.ldw = 3 * H; c.n_out = H; c.act = 1; c.n_rows = E; c.B = B; c.out = hid;
  c.add[0] = Ps; c.add_idx[0] = g->src; c.add_bs[0] = Bs > 1 ? ns * H : 0;
  c.add[1] = Pr; c.add_idx[1] = g->dst; c.add_bs[1] = Br > 1 ? nr * H : 0;
  rc = tc_linear(c, st);
  if (rc) return rc;
  // 4: edge layer 2: m = LN(W2·h + b2) (+ x_j for PropagationNet); e' = e + m
  memset(&c, 0, sizeof(c));
  c.x0 = hid; c.x0_bs = E * H; c.k0 = H; c.w = em->w[1]; c.ldw = H; c.bias = em->b[1]; c.n_out = H; c.gamma = em->ln_gamma;
  c.beta = em->ln_beta; c.eps = em->ln_eps; c.n_rows = E; c.B = B;
  if (prop) { c.post = send; c.post_idx = g->src; c.post_bs = send_bs; }
  if (edge_out) { c.out = edge_out; c.out2 = msg; c.res = edge; c.res_bs = edge_bs; }
  else c.out = msg;
  rc = tc_linear(c, st);
  if (rc) return rc;
  // 5: CSR segmented sum (deterministic)
  rc = nlam_segment_sum(g->rowptr, nullptr, nr, msg, E * H, aggr, nr * H, B, H, mean, st);
  if (rc) return rc;
  // 6: node update
  NlamRowSrc srcs[2] = {{rec, nullptr, rec_bs, H, 0}, {aggr, nullptr, (int64_t)nr * H, H, 0}};
  NlamRowSrc res = prop ? srcs[1] : srcs[0];
  return tc_mlp2(am, srcs, 2, &res, rec_out, nr, B, st, nhid);
}

}  // namespace nlam
