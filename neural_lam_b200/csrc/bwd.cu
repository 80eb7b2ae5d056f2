// Building blocks of the hand-written BACKWARD of the path (reference: autograd through gnn_layers.py:110-157 and
// utils/networks.py:27-40; tests/test_gnn_layers.py section F).  The dense products of the backward —
//   dX = dY · W          (Linear with the transposed weight)
//   dW = dYᵀ · X         (reduction over the rows: split-K partial products on the tensor cores + a deterministic
//                         reduction of the partials)
// run on the generic tcgen05 Linear kernel (tc7.cu); this file holds the elementwise / reduction / layout kernels
// around them: SiLU forward / backward, LayerNorm forward / backward (with the dγ / dβ column reductions), column
// sums (bias gradients), zero-padded transposes (the K-major operands of the dW products), the reduction of partial
// products, and the gather-add that assembles the message gradient g_m[e] = g_e'[e] + g_aggr[dst(e)] / deg.
// All reductions are two-stage and ordered: results are bit-reproducible run to run (no floating-point atomics).
#include "common.cuh"

namespace nlam {

static int grid_1d(long long total, int block, int cap = 148 * 8) {
  long long g = (total + block - 1) / block;
  return (int)std::max<long long>(1, std::min<long long>(g, cap));
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }

__global__ void silu_fwd_kernel(const float* __restrict__ z, float* __restrict__ h, long long n4) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 v = reinterpret_cast<const float4*>(z)[i];
    v.x *= sigmoidf_(v.x);
    v.y *= sigmoidf_(v.y);
    v.z *= sigmoidf_(v.z);
    v.w *= sigmoidf_(v.w);
    reinterpret_cast<float4*>(h)[i] = v;
  }
}

__device__ __forceinline__ float dsilu(float z) {
  const float s = sigmoidf_(z);
  return s * (1.0f + z * (1.0f - s));
}

__global__ void silu_bwd_kernel(const float* __restrict__ gh, const float* __restrict__ z, float* __restrict__ gz, long long n4) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 g = reinterpret_cast<const float4*>(gh)[i];
    const float4 v = reinterpret_cast<const float4*>(z)[i];
    reinterpret_cast<float4*>(gz)[i] = make_float4(g.x * dsilu(v.x), g.y * dsilu(v.y), g.z * dsilu(v.z), g.w * dsilu(v.w));
  }
}

// one warp per row; H <= 256 (8 values per lane)
template <int VPL>
__device__ __forceinline__ void row_stats(const float (&v)[VPL], int H, float eps, float& mu, float& rstd) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) s += v[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  mu = s / H;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const float d = v[i] - mu;
    q += d * d;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  rstd = rsqrtf(q / H + eps);
}

template <int VPL>
__global__ void ln_fwd_kernel(const float* __restrict__ y, const float* __restrict__ gamma, const float* __restrict__ beta,
                              float eps, float* __restrict__ out, long long rows) {
  constexpr int H = VPL * 32;
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp; r < rows; r += n_warps) {
    float v[VPL];
#pragma unroll
    for (int i = 0; i < VPL; ++i) v[i] = y[r * H + lane + 32 * i];
    float mu, rstd;
    row_stats<VPL>(v, H, eps, mu, rstd);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c = lane + 32 * i;
      out[r * H + c] = (v[i] - mu) * rstd * gamma[c] + beta[c];
    }
  }
}

// g_y = rstd * (γ g − mean(γ g) − x̂ · mean(γ g x̂)); per-warp partial dγ = Σ g x̂, dβ = Σ g  ->  part[warp][2H]
template <int VPL>
__global__ void ln_bwd_kernel(const float* __restrict__ g, const float* __restrict__ y, const float* __restrict__ gamma,
                              float eps, float* __restrict__ gy, float* __restrict__ part, long long rows) {
  constexpr int H = VPL * 32;
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  float dg[VPL], db[VPL], gm[VPL];
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    dg[i] = 0.f;
    db[i] = 0.f;
    gm[i] = gamma[lane + 32 * i];
  }
  for (long long r = warp; r < rows; r += n_warps) {
    float v[VPL], gg[VPL];
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      v[i] = y[r * H + lane + 32 * i];
      gg[i] = g[r * H + lane + 32 * i];
    }
    float mu, rstd;
    row_stats<VPL>(v, H, eps, mu, rstd);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const float xh = (v[i] - mu) * rstd;
      const float gx = gg[i] * gm[i];
      s1 += gx;
      s2 += gx * xh;
      dg[i] += gg[i] * xh;
      db[i] += gg[i];
      v[i] = xh;
      gg[i] = gx;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    s1 /= H;
    s2 /= H;
#pragma unroll
    for (int i = 0; i < VPL; ++i) gy[r * H + lane + 32 * i] = rstd * (gg[i] - s1 - v[i] * s2);
  }
  if (warp < n_warps) {
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      part[warp * 2 * H + lane + 32 * i] = dg[i];
      part[warp * 2 * H + H + lane + 32 * i] = db[i];
    }
  }
}

// per-warp column sums of g (rows x C, C <= 1024): part[warp][C]
__global__ void colsum_kernel(const float* __restrict__ g, long long rows, int C, float* __restrict__ part) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (int c0 = 0; c0 < C; c0 += 32) {
    const int c = c0 + lane;
    float s = 0.f;
    if (c < C)
      for (long long r = warp; r < rows; r += n_warps) s += g[r * C + c];
    if (c < C) part[warp * C + c] = s;
  }
}

// out[r][c] (+)= Σ_p part[p][r][c]   (ordered: deterministic)
__global__ void reduce_partials_kernel(const float* __restrict__ part, int P, long long R, int C, float* __restrict__ out,
                                       long long out_pitch, int accumulate) {
  const long long total = R * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / C;
    const int c = (int)(i - r * C);
    float s = 0.f;
    for (int p = 0; p < P; ++p) s += part[(long long)p * total + i];
    float* o = out + r * out_pitch + c;
    *o = accumulate ? *o + s : s;
  }
}

// XT[c][r] = X[r][c] for r < rows, 0 for rows <= r < rows_pad   (32 x 32 tiles through shared memory)
__global__ void transpose_pad_kernel(const float* __restrict__ x, long long rows, int C, long long x_pitch, float* __restrict__ xt,
                                     long long rows_pad) {
  __shared__ float tile[32][33];
  const long long tiles_r = (rows_pad + 31) / 32;
  const int tiles_c = (C + 31) / 32;
  for (long long t = blockIdx.x; t < tiles_r * tiles_c; t += gridDim.x) {
    const long long tr = t / tiles_c;
    const int tc = (int)(t - tr * tiles_c);
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8 threads
    for (int j = ty; j < 32; j += 8) {
      const long long r = tr * 32 + j;
      const int c = tc * 32 + tx;
      tile[j][tx] = (r < rows && c < C) ? x[r * x_pitch + c] : 0.f;
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
      const int c = tc * 32 + j;
      const long long r = tr * 32 + tx;
      if (c < C && r < rows_pad) xt[(long long)c * rows_pad + r] = tile[tx][j];
    }
    __syncthreads();
  }
}

// out[b, e, :] = (a ? a[b, e, :] : 0) + scale_e * v[b, idx[e], :],  scale_e = 1 or 1 / max(deg(idx[e]), 1)
__global__ void add_gather_kernel(const float* __restrict__ a, const float* __restrict__ v, const int32_t* __restrict__ idx,
                                  const int32_t* __restrict__ deg_ptr, long long n_e, long long n_v, int H4, int B,
                                  float* __restrict__ out) {
  const long long total = (long long)B * n_e * H4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % H4);
    const long long e = (i / H4) % n_e;
    const long long b = i / ((long long)H4 * n_e);
    const int j = idx[e];
    float sc = 1.f;
    if (deg_ptr) sc = 1.0f / (float)max(deg_ptr[j + 1] - deg_ptr[j], 1);
    float4 x = reinterpret_cast<const float4*>(v)[(b * n_v + j) * H4 + c];
    x.x *= sc; x.y *= sc; x.z *= sc; x.w *= sc;
    if (a) {
      const float4 y = reinterpret_cast<const float4*>(a)[i];
      x.x += y.x; x.y += y.y; x.z += y.z; x.w += y.w;
    }
    reinterpret_cast<float4*>(out)[i] = x;
  }
}

}  // namespace nlam

using namespace nlam;

extern "C" int nlam_silu(const float* z, const float* gh, float* out, int64_t n, void* stream) {
  NLAM_REQUIRE(z && out && n % 4 == 0, NLAM_E_INVALID, "nlam_silu: bad arguments");
  if (n == 0) return NLAM_OK;
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope ps(gh ? "silu_bwd_kernel" : "silu_fwd_kernel", st, 4.0 * n * (gh ? 3 : 2));
  if (gh) silu_bwd_kernel<<<grid_1d(n / 4, 256), 256, 0, st>>>(gh, z, out, n / 4);
  else silu_fwd_kernel<<<grid_1d(n / 4, 256), 256, 0, st>>>(z, out, n / 4);
  count_launch();
  NLAM_CUDA_OK(cudaGetLastError());
  return NLAM_OK;
}

extern "C" int nlam_layernorm_fwd(const float* y, const float* gamma, const float* beta, float eps, float* out, int64_t rows,
                                  int H, void* stream) {
  NLAM_REQUIRE(y && gamma && beta && out && (H == 32 || H == 64 || H == 128 || H == 256), NLAM_E_UNSUPPORTED,
               "nlam_layernorm_fwd: width %d unsupported", H);
  if (rows == 0) return NLAM_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = grid_1d(rows * 32, 256);
  ProfScope ps("ln_fwd_kernel", st, 8.0 * rows * H);
  if (H == 32) ln_fwd_kernel<1><<<grid, 256, 0, st>>>(y, gamma, beta, eps, out, rows);
  else if (H == 64) ln_fwd_kernel<2><<<grid, 256, 0, st>>>(y, gamma, beta, eps, out, rows);
  else if (H == 128) ln_fwd_kernel<4><<<grid, 256, 0, st>>>(y, gamma, beta, eps, out, rows);
  else ln_fwd_kernel<8><<<grid, 256, 0, st>>>(y, gamma, beta, eps, out, rows);
  count_launch();
  NLAM_CUDA_OK(cudaGetLastError());
  return NLAM_OK;
}

extern "C" size_t nlam_bwd_scratch_floats(int C) { return (size_t)(148 * 8 * 8 + 2) * 2 * (size_t)C; }

extern "C" int nlam_layernorm_bwd(const float* g, const float* y, const float* gamma, float eps, float* gy, float* dgamma,
                                  float* dbeta, int64_t rows, int H, float* scratch, void* stream) {
  NLAM_REQUIRE(g && y && gamma && gy && dgamma && dbeta && scratch && (H == 32 || H == 64 || H == 128 || H == 256),
               NLAM_E_UNSUPPORTED, "nlam_layernorm_bwd: width %d unsupported", H);
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = grid_1d(std::max<int64_t>(rows, 1) * 32, 256);
  const int n_warps = grid * 8;
  {
    ProfScope ps("ln_bwd_kernel", st, 12.0 * rows * H);
    if (H == 32) ln_bwd_kernel<1><<<grid, 256, 0, st>>>(g, y, gamma, eps, gy, scratch, rows);
    else if (H == 64) ln_bwd_kernel<2><<<grid, 256, 0, st>>>(g, y, gamma, eps, gy, scratch, rows);
    else if (H == 128) ln_bwd_kernel<4><<<grid, 256, 0, st>>>(g, y, gamma, eps, gy, scratch, rows);
    else ln_bwd_kernel<8><<<grid, 256, 0, st>>>(g, y, gamma, eps, gy, scratch, rows);
  }
  count_launch();
  // partials (n_warps, 2H): columns [0,H) -> dgamma, [H,2H) -> dbeta
  reduce_partials_kernel<<<grid_1d(2 * H, 128), 128, 0, st>>>(scratch, n_warps, 1, 2 * H, scratch + (size_t)n_warps * 2 * H, 2 * H, 0);
  count_launch();
  NLAM_CUDA_OK(cudaMemcpyAsync(dgamma, scratch + (size_t)n_warps * 2 * H, H * sizeof(float), cudaMemcpyDeviceToDevice, st));
  NLAM_CUDA_OK(cudaMemcpyAsync(dbeta, scratch + (size_t)n_warps * 2 * H + H, H * sizeof(float), cudaMemcpyDeviceToDevice, st));
  NLAM_CUDA_OK(cudaGetLastError());
  return NLAM_OK;
}

extern "C" int nlam_colsum(const float* g, int64_t rows, int C, float* out, float* scratch, void* stream) {
  NLAM_REQUIRE(g && out && scratch && C >= 1 && C <= 1024, NLAM_E_INVALID, "nlam_colsum: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = grid_1d(std::max<int64_t>(rows, 1) * 32, 256, 148 * 4);
  const int n_warps = grid * 8;
  {
    ProfScope ps("colsum_kernel", st, 4.0 * rows * C);
    colsum_kernel<<<grid, 256, 0, st>>>(g, rows, C, scratch);
  }
  count_launch();
  reduce_partials_kernel<<<grid_1d(C, 128), 128, 0, st>>>(scratch, n_warps, 1, C, out, C, 0);
  count_launch();
  NLAM_CUDA_OK(cudaGetLastError());
  return NLAM_OK;
}

extern "C" int nlam_reduce_partials(const float* part, int P, int64_t R, int C, float* out, int64_t out_pitch, int accumulate,
                                    void* stream) {
  NLAM_REQUIRE(part && out && P >= 1 && R >= 1 && C >= 1, NLAM_E_INVALID, "nlam_reduce_partials: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  {
    ProfScope ps("reduce_partials_kernel", st, 4.0 * P * R * C);
    reduce_partials_kernel<<<grid_1d(R * C, 256), 256, 0, st>>>(part, P, R, C, out, out_pitch, accumulate);
  }
  count_launch();
  NLAM_CUDA_OK(cudaGetLastError());
  return NLAM_OK;
}

extern "C" int nlam_transpose_pad(const float* x, int64_t rows, int C, int64_t x_pitch, float* xt, int64_t rows_pad, void* stream) {
  NLAM_REQUIRE(x && xt && rows >= 0 && rows_pad >= rows && C >= 1, NLAM_E_INVALID, "nlam_transpose_pad: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const long long tiles = ((rows_pad + 31) / 32) * ((C + 31) / 32);
  {
    ProfScope ps("transpose_pad_kernel", st, 4.0 * C * (rows + rows_pad));
    transpose_pad_kernel<<<(int)std::min<long long>(tiles, 148 * 16), 256, 0, st>>>(x, rows, C, x_pitch, xt, rows_pad);
  }
  count_launch();
  NLAM_CUDA_OK(cudaGetLastError());
  return NLAM_OK;
}

extern "C" int nlam_add_gather(const float* a, const float* v, const int32_t* idx, const int32_t* deg_ptr, int64_t n_e,
                               int64_t n_v, int H, int B, float* out, void* stream) {
  NLAM_REQUIRE(v && idx && out && H % 4 == 0, NLAM_E_INVALID, "nlam_add_gather: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const long long total = (long long)B * n_e * (H / 4);
  if (total == 0) return NLAM_OK;
  {
    ProfScope ps("add_gather_kernel", st, 4.0 * B * n_e * H * (a ? 3 : 2));
    add_gather_kernel<<<grid_1d(total, 256), 256, 0, st>>>(a, v, idx, deg_ptr, n_e, n_v, H / 4, B, out);
  }
  count_launch();
  NLAM_CUDA_OK(cudaGetLastError());
  return NLAM_OK;
}

// Generic tcgen05 Linear through the C ABI (LinearCall fields flattened; see common.cuh / tc7.cu)
extern "C" int nlam_linear(const float* x0, int64_t x0_bs, int k0, int64_t x0_pitch, const float* x1, int64_t x1_bs, int k1,
                           const float* w, int ldw, int w_cols, int64_t w_bs, const float* bias, int n_out, int act,
                           const float* gamma, const float* beta, float eps, const float* add0, const int32_t* add0_idx,
                           int64_t add0_bs, const float* add1, const int32_t* add1_idx, int64_t add1_bs, const float* post,
                           const int32_t* post_idx, int64_t post_bs, const float* res, int64_t res_bs, int64_t n_rows, int B,
                           float* out, float* out2, void* stream) {
  NLAM_REQUIRE(x0 && w && out && n_rows >= 1 && B >= 1, NLAM_E_INVALID, "nlam_linear: bad arguments");
  LinearCall c;
  memset(&c, 0, sizeof(c));
  c.x0 = x0; c.x0_bs = x0_bs; c.k0 = k0; c.x0_pitch = x0_pitch; c.x1 = x1; c.x1_bs = x1_bs; c.k1 = k1;
  c.w = w; c.ldw = ldw; c.w_cols = w_cols; c.w_bs = w_bs; c.bias = bias; c.n_out = n_out; c.act = act;
  c.gamma = gamma; c.beta = beta; c.eps = eps;
  c.add[0] = add0; c.add_idx[0] = add0_idx; c.add_bs[0] = add0_bs;
  c.add[1] = add1; c.add_idx[1] = add1_idx; c.add_bs[1] = add1_bs;
  c.post = post; c.post_idx = post_idx; c.post_bs = post_bs; c.res = res; c.res_bs = res_bs;
  c.n_rows = n_rows; c.B = B; c.out = out; c.out2 = out2;
  return tc_linear(c, (cudaStream_t)stream);
}
