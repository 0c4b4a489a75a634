// Exact-fp32 CUDA-core GEMM for the MLP layers: the parity back end (DLRM_GEMM_SIMT_FP32) and
// the any-shape path (K=13, N=1, dim 2 ... of the reference's CLI defaults).  Replaces
// aten::addmm + ReLU/Sigmoid modules (dlrm_s_pytorch.py:399-405) and their autograd.
//
//   C[i,j] = epilogue( sum_l A(i,l) * B(j,l) ),  A(i,l) = A[i*sa_i + l*sa_l],  B likewise.
// 64x64x16 tiles, 256 threads, 4x4 register micro-tile, fp32 FFMA, fp32 accumulate.
#include "common.cuh"

namespace dlrm {

constexpr int BM = 64, BN = 64, BK = 16, LDS_PAD = 4;

struct GemmArgs {
  const float* A; long long sa_i, sa_l;
  const float* B; long long sb_j, sb_l;
  float* C; long long ldc;
  const float* bias;          // [N] or null
  const float* mask; long long ldm;  // activation values for act'() or null
  int act;                    // forward activation applied to C
  int mask_act;               // which act' to apply with mask
  long long M, N, K;
};

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == DLRM_ACT_RELU) return fmaxf(v, 0.f);
  if (act == DLRM_ACT_SIGMOID) return 1.0f / (1.0f + expf(-v));
  return v;
}
__device__ __forceinline__ float act_grad(float y, int act) {
  if (act == DLRM_ACT_RELU) return y > 0.f ? 1.f : 0.f;
  if (act == DLRM_ACT_SIGMOID) return (1.0f - y) * y;
  return 1.f;
}

// load a [64 rows x 16 l] tile into smem as S[l][row]
template <bool LCONTIG>
__device__ __forceinline__ void load_tile(float (*S)[BM + LDS_PAD], const float* P, long long s_r,
                                          long long s_l, long long r0, long long l0, long long R,
                                          long long L, bool vec_ok, int t) {
  if (LCONTIG) {
    const int row = t >> 2, lc = (t & 3) * 4;
    const long long r = r0 + row, l = l0 + lc;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (r < R) {
      const float* p = P + r * s_r + l;
      if (vec_ok && l + 3 < L) {
        const float4 q = *reinterpret_cast<const float4*>(p);
        v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) if (l + e < L) v[e] = p[e];
      }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) S[lc + e][row] = v[e];
  } else {
    const int lr = t >> 4, rc = (t & 15) * 4;
    const long long l = l0 + lr, r = r0 + rc;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (l < L) {
      const float* p = P + l * s_l + r * s_r;
      if (vec_ok && s_r == 1 && r + 3 < R) {
        const float4 q = *reinterpret_cast<const float4*>(p);
        v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) if (r + e < R) v[e] = p[e * s_r];
      }
    }
    *reinterpret_cast<float4*>(&S[lr][rc]) = make_float4(v[0], v[1], v[2], v[3]);
  }
}

template <bool AL, bool BL>
__global__ void __launch_bounds__(256) sgemm_kernel(const GemmArgs g, bool a_vec, bool b_vec) {
  __shared__ __align__(16) float As[2][BK][BM + LDS_PAD];
  __shared__ __align__(16) float Bs[2][BK][BN + LDS_PAD];
  const int t = threadIdx.x;
  const int ty = t >> 4, tx = t & 15;
  const long long i0 = (long long)blockIdx.y * BM, j0 = (long long)blockIdx.x * BN;
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;

  const int nk = (int)((g.K + BK - 1) / BK);
  load_tile<AL>(As[0], g.A, g.sa_i, g.sa_l, i0, 0, g.M, g.K, a_vec, t);
  load_tile<BL>(Bs[0], g.B, g.sb_j, g.sb_l, j0, 0, g.N, g.K, b_vec, t);
  __syncthreads();
  for (int kb = 0; kb < nk; ++kb) {
    const int cur = kb & 1;
    if (kb + 1 < nk) {
      load_tile<AL>(As[cur ^ 1], g.A, g.sa_i, g.sa_l, i0, (long long)(kb + 1) * BK, g.M, g.K, a_vec, t);
      load_tile<BL>(Bs[cur ^ 1], g.B, g.sb_j, g.sb_l, j0, (long long)(kb + 1) * BK, g.N, g.K, b_vec, t);
    }
#pragma unroll
    for (int l = 0; l < BK; ++l) {
      const float4 a = *reinterpret_cast<const float4*>(&As[cur][l][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[cur][l][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) acc[x][y] = fmaf(av[x], bv[y], acc[x][y]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int x = 0; x < 4; ++x) {
    const long long i = i0 + ty * 4 + x;
    if (i >= g.M) continue;
#pragma unroll
    for (int y = 0; y < 4; ++y) {
      const long long j = j0 + tx * 4 + y;
      if (j >= g.N) continue;
      float v = acc[x][y];
      if (g.bias) v += g.bias[j];
      v = apply_act(v, g.act);
      if (g.mask) v *= act_grad(g.mask[i * g.ldm + j], g.mask_act);
      g.C[i * g.ldc + j] = v;
    }
  }
}

static int launch_sgemm(const GemmArgs& g, cudaStream_t st) {
  if (g.M <= 0 || g.N <= 0) return 0;
  const bool AL = g.sa_l == 1, BL = g.sb_l == 1;
  const bool a_vec = aligned16(g.A) && ((AL ? g.sa_i : g.sa_l) % 4 == 0);
  const bool b_vec = aligned16(g.B) && ((BL ? g.sb_j : g.sb_l) % 4 == 0);
  dim3 grid((unsigned)((g.N + BN - 1) / BN), (unsigned)((g.M + BM - 1) / BM));
  if (grid.y > 65535) return set_error("sgemm: M=%lld too large", g.M);
  if (AL && BL) sgemm_kernel<true, true><<<grid, 256, 0, st>>>(g, a_vec, b_vec);
  else if (AL && !BL) sgemm_kernel<true, false><<<grid, 256, 0, st>>>(g, a_vec, b_vec);
  else if (!AL && BL) sgemm_kernel<false, true><<<grid, 256, 0, st>>>(g, a_vec, b_vec);
  else sgemm_kernel<false, false><<<grid, 256, 0, st>>>(g, a_vec, b_vec);
  DLRM_CHECK_LAUNCH("sgemm_kernel");
  return 0;
}

// dbias[n] = sum_m dY[m,n]; fixed reduction order (deterministic)
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ dY, long long ld,
                                                     float* __restrict__ out, long long M,
                                                     long long N) {
  __shared__ float red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const long long n = (long long)blockIdx.x * 32 + tx;
  float s = 0.f;
  if (n < N)
    for (long long m = ty; m < M; m += 8) s += dY[m * ld + n];
  red[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && n < N) {
    float a = red[0][tx];
#pragma unroll
    for (int r = 1; r < 8; ++r) a += red[r][tx];
    out[n] = a;
  }
}

int simt_linear_fwd(const float* X, long long ldx, const float* W, long long ldw, const float* bias,
                    float* Y, long long ldy, long long M, long long N, long long K, int act,
                    cudaStream_t st) {
  GemmArgs g{};
  g.A = X; g.sa_i = ldx; g.sa_l = 1;
  g.B = W; g.sb_j = ldw; g.sb_l = 1;
  g.C = Y; g.ldc = ldy; g.bias = bias; g.mask = nullptr; g.ldm = 0; g.act = act; g.mask_act = 0;
  g.M = M; g.N = N; g.K = K;
  return launch_sgemm(g, st);
}

int simt_linear_dgrad(const float* dY, long long lddy, const float* W, long long ldw,
                      const float* Xact, long long ldxa, int act_prev, float* dX, long long lddx,
                      long long M, long long N, long long K, cudaStream_t st) {
  GemmArgs g{};  // dX[m,k] = sum_n dY[m,n] W[n,k]
  g.A = dY; g.sa_i = lddy; g.sa_l = 1;
  g.B = W; g.sb_j = 1; g.sb_l = ldw;
  g.C = dX; g.ldc = lddx; g.bias = nullptr;
  g.mask = (act_prev != DLRM_ACT_NONE) ? Xact : nullptr; g.ldm = ldxa; g.act = DLRM_ACT_NONE;
  g.mask_act = act_prev;
  g.M = M; g.N = K; g.K = N;
  return launch_sgemm(g, st);
}

int simt_linear_wgrad(const float* dY, long long lddy, const float* X, long long ldx, float* dW,
                      long long lddw, float* dbias, long long M, long long N, long long K,
                      cudaStream_t st) {
  GemmArgs g{};  // dW[n,k] = sum_m dY[m,n] X[m,k]
  g.A = dY; g.sa_i = 1; g.sa_l = lddy;
  g.B = X; g.sb_j = 1; g.sb_l = ldx;
  g.C = dW; g.ldc = lddw; g.bias = nullptr; g.mask = nullptr; g.act = DLRM_ACT_NONE;
  g.M = N; g.N = K; g.K = M;
  if (int rc = launch_sgemm(g, st)) return rc;
  if (dbias && N > 0) {
    colsum_kernel<<<(unsigned)((N + 31) / 32), 256, 0, st>>>(dY, lddy, dbias, M, N);
    DLRM_CHECK_LAUNCH("colsum_kernel");
  }
  return 0;
}

}  // namespace dlrm
