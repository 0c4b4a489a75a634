// interact_features (dlrm_s_pytorch.py:483-504), "dot" op, forward and backward, CUDA-core fp32.
// Fuses torch.cat -> torch.bmm -> Z[:, li, lj] -> torch.cat into one pass over T:
// the operand T[b] = [x; ly_0; ...; ly_{T-1}] is read IN PLACE from the buffer the bottom MLP and
// the gather wrote, the strict-lower-triangle flatten happens in the epilogue, and x is copied
// into R[:, 0:D] on the way.
#include <cuda_bf16.h>

#include "common.cuh"

namespace dlrm {

// ------------------------------------------------------------------------------------------
// forward: 3x3 register blocks of the Gram matrix, lower-triangular blocks only.
// smem per sample: T rows padded to Fp = 3*ceil(F/3) rows of (D+1) floats (conflict-free
// column walks) + the flattened interactions (coalesced write-out).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(192) interact_fwd_kernel(const float* __restrict__ T, long long ldt,
                                                           float* __restrict__ R, long long ldr,
                                                           long long batch, int F, int D, int itself,
                                                           int spb) {
  extern __shared__ __align__(128) float smem[];
  const int nb = (F + 2) / 3;
  const int Fp = nb * 3;
  const int LD = D + 1;
  const int tps = nb * (nb + 1) / 2;                 // threads (blocks) per sample
  const int npairs = itself ? F * (F + 1) / 2 : F * (F - 1) / 2;
  const int per_sample = Fp * LD + npairs;
  const long long s0 = (long long)blockIdx.x * spb;
  const int ns = (int)min((long long)spb, batch - s0);

  // load T (zero-padded rows), copy x to R[:, 0:D]
  for (int e = threadIdx.x; e < ns * Fp * D; e += blockDim.x) {
    const int s = e / (Fp * D);
    const int rem = e - s * Fp * D;
    const int f = rem / D, d = rem - f * D;
    float v = 0.f;
    if (f < F) v = T[(s0 + s) * ldt + (long long)f * D + d];
    smem[s * per_sample + f * LD + d] = v;
    if (f == 0) R[(s0 + s) * ldr + d] = v;
  }
  __syncthreads();

  for (int item = threadIdx.x; item < ns * tps; item += blockDim.x) {
    const int s = item / tps;
    const int q = item - s * tps;
    // q -> (a, c) with c <= a:  a = floor((sqrt(8q+1)-1)/2)
    int a = (int)((sqrtf(8.f * q + 1.f) - 1.f) * 0.5f);
    while ((a + 1) * (a + 2) / 2 <= q) ++a;
    while (a * (a + 1) / 2 > q) --a;
    const int c = q - a * (a + 1) / 2;
    const float* Ts = smem + s * per_sample;
    const float* ra = Ts + (3 * a) * LD;
    const float* rc = Ts + (3 * c) * LD;
    float z[3][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
    for (int d = 0; d < D; ++d) {
      const float a0 = ra[d], a1 = ra[LD + d], a2 = ra[2 * LD + d];
      const float c0 = rc[d], c1 = rc[LD + d], c2 = rc[2 * LD + d];
      z[0][0] = fmaf(a0, c0, z[0][0]); z[0][1] = fmaf(a0, c1, z[0][1]); z[0][2] = fmaf(a0, c2, z[0][2]);
      z[1][0] = fmaf(a1, c0, z[1][0]); z[1][1] = fmaf(a1, c1, z[1][1]); z[1][2] = fmaf(a1, c2, z[1][2]);
      z[2][0] = fmaf(a2, c0, z[2][0]); z[2][1] = fmaf(a2, c1, z[2][1]); z[2][2] = fmaf(a2, c2, z[2][2]);
    }
    float* Zs = smem + s * per_sample + Fp * LD;
#pragma unroll
    for (int x = 0; x < 3; ++x)
#pragma unroll
      for (int y = 0; y < 3; ++y) {
        const int i = 3 * a + x, j = 3 * c + y;
        if (i < F && (j < i || (itself && j == i))) {
          const int p = itself ? i * (i + 1) / 2 + j : i * (i - 1) / 2 + j;
          Zs[p] = z[x][y];
        }
      }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < ns * npairs; e += blockDim.x) {
    const int s = e / npairs, p = e - s * npairs;
    R[(s0 + s) * ldr + D + p] = smem[s * per_sample + Fp * LD + p];
  }
}

// ------------------------------------------------------------------------------------------
// forward v2 (dim % 4 == 0): the F x D operand of each sample is one contiguous block of T, so
// it is brought into shared memory by ONE bulk-async copy (cp.async.bulk + mbarrier complete_tx,
// the 1-D TMA path) instead of per-thread loads; rows stay unpadded and every lane walks the
// feature dimension rotated by its lane id, which keeps the 32 lanes on 32 different banks.
// Optionally writes R as the (hi, lo) bf16 operand pair of the first top-MLP GEMM.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(192) interact_fwd2_kernel(const float* __restrict__ T, long long ldt,
                                                            float* __restrict__ R, long long ldr,
                                                            __nv_bfloat16* __restrict__ Rh,
                                                            __nv_bfloat16* __restrict__ Rl, long long ldrb,
                                                            long long batch, int F, int D, int itself,
                                                            int spb) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ __align__(128) float smem[];
  const int nb = (F + 2) / 3;
  const int Fp = nb * 3;
  const int tps = nb * (nb + 1) / 2;
  const int npairs = itself ? F * (F + 1) / 2 : F * (F - 1) / 2;
  const int per_sample = Fp * D + ((npairs + 3) & ~3);
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(smem + (size_t)spb * per_sample);
  const long long s0 = (long long)blockIdx.x * spb;
  const int ns = (int)min((long long)spb, batch - s0);
  const unsigned bar_a = (unsigned)__cvta_generic_to_shared(bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const unsigned bytes = (unsigned)(F * D * 4);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(bytes * ns) : "memory");
    for (int s = 0; s < ns; ++s) {
      const unsigned dst = (unsigned)__cvta_generic_to_shared(smem + (size_t)s * per_sample);
      const float* src = T + (s0 + s) * ldt;
      asm volatile(
          "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
          "l"(src), "r"(bytes), "r"(bar_a)
          : "memory");
    }
  }
  // zero the padding rows F..Fp-1 while the copy is in flight
  for (int e = threadIdx.x; e < ns * (Fp - F) * D; e += blockDim.x) {
    const int s = e / ((Fp - F) * D);
    const int rem = e - s * (Fp - F) * D;
    smem[(size_t)s * per_sample + F * D + rem] = 0.f;
  }
  __syncthreads();  // barrier init visible to all threads + padding written
  {
    unsigned ok = 0;
    int tries = 0;
    while (true) {
      asm volatile(
          "{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.u32 %0, 1, 0, P1;\n\t}"
          : "=r"(ok)
          : "r"(bar_a), "r"(0u)
          : "memory");
      if (ok) break;
      if (++tries > (1 << 22)) __trap();
    }
  }
  const int lane = threadIdx.x & 31;
  for (int item = threadIdx.x; item < ns * tps; item += blockDim.x) {
    const int s = item / tps;
    const int q = item - s * tps;
    int a = (int)((sqrtf(8.f * q + 1.f) - 1.f) * 0.5f);
    while ((a + 1) * (a + 2) / 2 <= q) ++a;
    while (a * (a + 1) / 2 > q) --a;
    const int c = q - a * (a + 1) / 2;
    const float* Ts = smem + (size_t)s * per_sample;
    const float* ra = Ts + (3 * a) * D;
    const float* rc = Ts + (3 * c) * D;
    float z[3][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
    int d = lane % D;
    for (int it = 0; it < D; ++it) {
      const float a0 = ra[d], a1 = ra[D + d], a2 = ra[2 * D + d];
      const float c0 = rc[d], c1 = rc[D + d], c2 = rc[2 * D + d];
      z[0][0] = fmaf(a0, c0, z[0][0]); z[0][1] = fmaf(a0, c1, z[0][1]); z[0][2] = fmaf(a0, c2, z[0][2]);
      z[1][0] = fmaf(a1, c0, z[1][0]); z[1][1] = fmaf(a1, c1, z[1][1]); z[1][2] = fmaf(a1, c2, z[1][2]);
      z[2][0] = fmaf(a2, c0, z[2][0]); z[2][1] = fmaf(a2, c1, z[2][1]); z[2][2] = fmaf(a2, c2, z[2][2]);
      if (++d == D) d = 0;
    }
    float* Zs = smem + (size_t)s * per_sample + Fp * D;
#pragma unroll
    for (int x = 0; x < 3; ++x)
#pragma unroll
      for (int y = 0; y < 3; ++y) {
        const int i = 3 * a + x, j = 3 * c + y;
        if (i < F && (j < i || (itself && j == i))) {
          const int p = itself ? i * (i + 1) / 2 + j : i * (i - 1) / 2 + j;
          Zs[p] = z[x][y];
        }
      }
  }
  __syncthreads();
  const int ncols = D + npairs;
  for (int e = threadIdx.x; e < ns * ncols; e += blockDim.x) {
    const int s = e / ncols, cidx = e - s * ncols;
    const float v = cidx < D ? smem[(size_t)s * per_sample + cidx]
                             : smem[(size_t)s * per_sample + Fp * D + (cidx - D)];
    if (R) R[(s0 + s) * ldr + cidx] = v;
    if (Rh) {
      const __nv_bfloat16 hb = __float2bfloat16_rn(v);
      Rh[(s0 + s) * ldrb + cidx] = hb;
      if (Rl) Rl[(s0 + s) * ldrb + cidx] = __float2bfloat16_rn(v - __bfloat162float(hb));
    }
  }
}

// ------------------------------------------------------------------------------------------
// backward: dT[i][d] = sum_j S[i][j] T[j][d] (+ dR[d] for i == 0),  S = dZ + dZ^T.
// One thread owns a column d of one sample: T[:, d] lives in registers, S rows are read as
// broadcast float4 from shared memory.
// ------------------------------------------------------------------------------------------
// ROUTE: feature i's gradient rows go to route.base[i] + sample * route.ld[i] instead of dT -- on a sharded
// run base[1 + t] points into the buffer of the rank that owns table t (NVLink peer store, 512-byte
// rows), so the gradient exchange rides on this kernel's stores and needs no all-to-all.
// A feature may have several destinations: a row-split table keeps a shard on every rank and each shard owner
// needs the gradient rows of all samples (placement.py), so its feature is stored to every rank.
constexpr int ROUTE_MAX_DST = 128;
struct FeatRoute {
  float* base[ROUTE_MAX_DST];
  long long ld[ROUTE_MAX_DST];
  unsigned char first[66];     // destinations [first[i], first[i + 1]) belong to feature i
  float emb_scale;             // factor on the gradient rows of features >= 1 (see dlrm_b200_interact_bwd_p2p)
};

struct NoRoute {};

template <int MAXF, typename Route>
__global__ void __launch_bounds__(128) interact_bwd_kernel(const float* __restrict__ T, long long ldt,
                                                           const float* __restrict__ dR, long long lddr,
                                                           float* __restrict__ dT, long long lddt,
                                                           long long batch, int F, int D, int itself,
                                                           int mask0, int spb,
                                                           __nv_bfloat16* __restrict__ g0h,
                                                           __nv_bfloat16* __restrict__ g0l, long long ldg0,
                                                           const __grid_constant__ Route route) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr bool ROUTE = sizeof(Route) > 8;
  extern __shared__ __align__(128) float smem[];
  const int F4 = (F + 3) & ~3;
  const long long s0 = (long long)blockIdx.x * spb;
  const int ns = (int)min((long long)spb, batch - s0);
  // S[s][i][j], row stride F4, zero padded
  for (int e = threadIdx.x; e < ns * F * F4; e += blockDim.x) {
    const int s = e / (F * F4);
    const int rem = e - s * F * F4;
    const int i = rem / F4, j = rem - i * F4;
    float v = 0.f;
    if (j < F) {
      const float* g = dR + (s0 + s) * lddr + D;
      if (i == j) {
        if (itself) v = 2.f * g[i * (i + 1) / 2 + i];
      } else {
        const int hi = i > j ? i : j, lo = i > j ? j : i;
        v = g[itself ? hi * (hi + 1) / 2 + lo : hi * (hi - 1) / 2 + lo];
      }
    }
    smem[e] = v;
  }
  __syncthreads();
  for (int item = threadIdx.x; item < ns * D; item += blockDim.x) {
    const int s = item / D, d = item - s * D;
    const float* Tb = T + (s0 + s) * ldt + d;
    float t[MAXF];
#pragma unroll
    for (int j = 0; j < MAXF; ++j) t[j] = (j < F) ? Tb[(long long)j * D] : 0.f;
    const float* Ss = smem + s * F * F4;
    float* out = dT + (s0 + s) * lddt + d;
    for (int i = 0; i < F; ++i) {
      float acc = (i == 0) ? dR[(s0 + s) * lddr + d] : 0.f;
#pragma unroll
      for (int j4 = 0; j4 < MAXF / 4; ++j4) {
        if (j4 * 4 < F) {
          const float4 sv = *reinterpret_cast<const float4*>(Ss + i * F4 + j4 * 4);
          acc = fmaf(sv.x, t[j4 * 4 + 0], acc);
          acc = fmaf(sv.y, t[j4 * 4 + 1], acc);
          acc = fmaf(sv.z, t[j4 * 4 + 2], acc);
          acc = fmaf(sv.w, t[j4 * 4 + 3], acc);
        }
      }
      if (i == 0 && mask0 == DLRM_ACT_RELU) acc = (t[0] > 0.f) ? acc : 0.f;
      if (i == 0 && mask0 == DLRM_ACT_SIGMOID) acc *= (1.0f - t[0]) * t[0];
      if constexpr (ROUTE) {
        const float a_ = i ? acc * route.emb_scale : acc;
        for (int q = route.first[i]; q < route.first[i + 1]; ++q) route.base[q][(s0 + s) * route.ld[q] + d] = a_;
      } else {
        out[(long long)i * D] = acc;
      }
      if (i == 0 && g0h) {  // feature 0 = gradient into the bottom MLP: also as a (hi, lo) bf16 pair
        const __nv_bfloat16 hb = __float2bfloat16_rn(acc);
        g0h[(s0 + s) * ldg0 + d] = hb;
        if (g0l) g0l[(s0 + s) * ldg0 + d] = __float2bfloat16_rn(acc - __bfloat162float(hb));
      }
    }
  }
}


// Two adjacent columns per thread (float2): every S value read from shared memory feeds two FMAs, and a
// warp moves 256 contiguous bytes per feature row.  Needs an even D and 8-byte aligned rows.
template <int MAXF, typename Route>
__global__ void __launch_bounds__(128) interact_bwd2_kernel(const float* __restrict__ T, long long ldt,
                                                            const float* __restrict__ dR, long long lddr,
                                                            float* __restrict__ dT, long long lddt,
                                                            long long batch, int F, int D, int itself,
                                                            int mask0, int spb,
                                                            __nv_bfloat16* __restrict__ g0h,
                                                            __nv_bfloat16* __restrict__ g0l, long long ldg0,
                                                            const __grid_constant__ Route route) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr bool ROUTE = sizeof(Route) > 8;
  extern __shared__ __align__(128) float smem[];
  const int F4 = (F + 3) & ~3;
  const int D2 = D >> 1;
  const long long s0 = (long long)blockIdx.x * spb;
  const int ns = (int)min((long long)spb, batch - s0);
  // S[s][i][j] = dZ + dZ^T, row stride F4, zero padded; one (s, i) row per loop trip, j = inner lanes
  for (int row = threadIdx.x / F4; row < ns * F; row += blockDim.x / F4) {
    const int j = threadIdx.x % F4;
    if (threadIdx.x / F4 >= blockDim.x / F4) break;
    const int s = row / F, i = row - s * F;
    float v = 0.f;
    if (j < F) {
      const float* g = dR + (s0 + s) * lddr + D;
      if (i == j) {
        if (itself) v = 2.f * g[i * (i + 1) / 2 + i];
      } else {
        const int hi = i > j ? i : j, lo = i > j ? j : i;
        v = g[itself ? hi * (hi + 1) / 2 + lo : hi * (hi - 1) / 2 + lo];
      }
    }
    smem[row * F4 + j] = v;
  }
  __syncthreads();
  for (int item = threadIdx.x; item < ns * D2; item += blockDim.x) {
    const int s = item / D2, d = (item - s * D2) * 2;
    const float* Tb = T + (s0 + s) * ldt + d;
    float2 t[MAXF];
#pragma unroll
    for (int j = 0; j < MAXF; ++j)
      t[j] = (j < F) ? *reinterpret_cast<const float2*>(Tb + (long long)j * D) : make_float2(0.f, 0.f);
    const float* Ss = smem + s * F * F4;
    for (int i = 0; i < F; ++i) {
      float2 acc = make_float2(0.f, 0.f);
      if (i == 0) acc = *reinterpret_cast<const float2*>(dR + (s0 + s) * lddr + d);
#pragma unroll
      for (int j4 = 0; j4 < MAXF / 4; ++j4) {
        if (j4 * 4 < F) {
          const float4 sv = *reinterpret_cast<const float4*>(Ss + i * F4 + j4 * 4);
          acc.x = fmaf(sv.x, t[j4 * 4 + 0].x, acc.x); acc.y = fmaf(sv.x, t[j4 * 4 + 0].y, acc.y);
          acc.x = fmaf(sv.y, t[j4 * 4 + 1].x, acc.x); acc.y = fmaf(sv.y, t[j4 * 4 + 1].y, acc.y);
          acc.x = fmaf(sv.z, t[j4 * 4 + 2].x, acc.x); acc.y = fmaf(sv.z, t[j4 * 4 + 2].y, acc.y);
          acc.x = fmaf(sv.w, t[j4 * 4 + 3].x, acc.x); acc.y = fmaf(sv.w, t[j4 * 4 + 3].y, acc.y);
        }
      }
      if (i == 0 && mask0 == DLRM_ACT_RELU) {
        acc.x = (t[0].x > 0.f) ? acc.x : 0.f;
        acc.y = (t[0].y > 0.f) ? acc.y : 0.f;
      }
      if (i == 0 && mask0 == DLRM_ACT_SIGMOID) {
        acc.x *= (1.0f - t[0].x) * t[0].x;
        acc.y *= (1.0f - t[0].y) * t[0].y;
      }
      if constexpr (ROUTE) {
        const float2 a_ = i ? make_float2(acc.x * route.emb_scale, acc.y * route.emb_scale) : acc;
        for (int q = route.first[i]; q < route.first[i + 1]; ++q)
          *reinterpret_cast<float2*>(route.base[q] + (s0 + s) * route.ld[q] + d) = a_;
      } else {
        *reinterpret_cast<float2*>(dT + (s0 + s) * lddt + (long long)i * D + d) = acc;
      }
      if (i == 0 && g0h) {  // feature 0 = gradient into the bottom MLP: also as a (hi, lo) bf16 pair
        const __nv_bfloat162 hb = __floats2bfloat162_rn(acc.x, acc.y);
        *reinterpret_cast<__nv_bfloat162*>(g0h + (s0 + s) * ldg0 + d) = hb;
        if (g0l)
          *reinterpret_cast<__nv_bfloat162*>(g0l + (s0 + s) * ldg0 + d) =
              __floats2bfloat162_rn(acc.x - __low2float(hb), acc.y - __high2float(hb));
      }
    }
  }
}

}  // namespace dlrm

extern "C" int dlrm_b200_interact_fwd_ex(const float* T, int64_t ldt, float* R, int64_t ldr, void* R_hi,
                                         void* R_lo, int64_t ld_rb, int64_t batch, int num_features,
                                         int dim, int itself, void* stream) {
  using namespace dlrm;
  if (batch == 0) return 0;
  if (num_features < 1 || dim < 1) return set_error("interact_fwd: F=%d D=%d", num_features, dim);
  if (!T || (!R && !R_hi)) return set_error("interact_fwd: NULL pointer");
  const int F = num_features, D = dim;
  const int nb = (F + 2) / 3, Fp = nb * 3, tps = nb * (nb + 1) / 2;
  const int npairs = itself ? F * (F + 1) / 2 : F * (F - 1) / 2;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool bulk = (D % 4 == 0) && (ldt % 4 == 0) && aligned16(T);
  if (bulk) {
    const size_t per_sample = ((size_t)Fp * D + ((npairs + 3) & ~3)) * sizeof(float);
    int spb = 192 / tps;
    if (spb < 1) spb = 1;
    while (spb > 1 && spb * per_sample + 16 > 100 * 1024) --spb;
    const size_t smem = spb * per_sample + 16;
    if (smem <= 200 * 1024) {
      static thread_local bool configured = false;
      if (smem > 48 * 1024 && !configured) {
        DLRM_CUDA(cudaFuncSetAttribute(interact_fwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        configured = true;
      }
      const long long grid = (batch + spb - 1) / spb;
      (void)launch_chain(interact_fwd2_kernel, dim3((unsigned)grid), dim3(192), smem, st, T, (long long)ldt, R,
                         (long long)ldr, static_cast<__nv_bfloat16*>(R_hi), static_cast<__nv_bfloat16*>(R_lo),
                         (long long)ld_rb, (long long)batch, F, D, itself, spb);
      DLRM_CHECK_LAUNCH("interact_fwd2_kernel");
      return 0;
    }
  }
  if (R_hi) return set_error("interact_fwd: bf16 output needs dim %% 4 == 0 and aligned T");
  const size_t per_sample = ((size_t)Fp * (D + 1) + npairs) * sizeof(float);
  if (per_sample > 200 * 1024)
    return set_error("interact_fwd: F=%d D=%d needs %zu B of shared memory per sample (max 200 KB)",
                     F, D, per_sample);
  int spb = 192 / tps;
  if (spb < 1) spb = 1;
  while (spb > 1 && spb * per_sample > 56 * 1024) --spb;
  const size_t smem = spb * per_sample;
  static thread_local size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    DLRM_CUDA(cudaFuncSetAttribute(interact_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   200 * 1024));
    configured = 200 * 1024;
  }
  const long long grid = (batch + spb - 1) / spb;
  interact_fwd_kernel<<<(unsigned)grid, 192, smem, st>>>(T, ldt, R, ldr, batch, F, D, itself, spb);
  DLRM_CHECK_LAUNCH("interact_fwd_kernel");
  return 0;
}

extern "C" int dlrm_b200_interact_fwd(const float* T, int64_t ldt, float* R, int64_t ldr,
                                      int64_t batch, int num_features, int dim, int itself,
                                      void* stream) {
  return dlrm_b200_interact_fwd_ex(T, ldt, R, ldr, nullptr, nullptr, 0, batch, num_features, dim, itself, stream);
}

namespace dlrm {
static int interact_bwd_launch(const float* T, int64_t ldt, const float* dR, int64_t lddr, float* dT,
                               int64_t lddt, const FeatRoute* route, int64_t batch, int num_features,
                               int dim, int itself, int mask_feature0, void* g0_hi, void* g0_lo,
                               int64_t ld_g0, void* stream) {
  if (batch == 0) return 0;
  const int F = num_features, D = dim;
  if (F < 1 || D < 1) return set_error("interact_bwd: F=%d D=%d", F, D);
  if (F > 64) return set_error("interact_bwd: num_features=%d > 64 not supported yet", F);
  if (!T || !dR || (!dT && !route)) return set_error("interact_bwd: NULL pointer");
  const int F4 = (F + 3) & ~3;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  __nv_bfloat16* gh = static_cast<__nv_bfloat16*>(g0_hi);
  __nv_bfloat16* gl = static_cast<__nv_bfloat16*>(g0_lo);
  // float2 variant: even D, 8-byte aligned rows everywhere
  auto even8 = [](const void* p, long long ld) { return (reinterpret_cast<uintptr_t>(p) & 7) == 0 && (ld & 1) == 0; };
  bool two = get_tunable(TUNE_INTERACT_BWD_COLS) != 1 && (D % 2 == 0) && even8(T, ldt) && even8(dR, lddr) && (route || even8(dT, lddt)) &&
             (!gh || ((reinterpret_cast<uintptr_t>(gh) & 3) == 0 && (ld_g0 & 1) == 0)) &&
             (!gl || (reinterpret_cast<uintptr_t>(gl) & 3) == 0);
  if (route)
    for (int q = 0; q < route->first[F]; ++q) two = two && even8(route->base[q], route->ld[q]);
  const int cols = two ? D / 2 : D;
  int spb = 128 / cols;
  if (spb < 1) spb = 1;
  while (spb > 1 && (size_t)spb * F * F4 * sizeof(float) > 48 * 1024) --spb;
  const size_t smem = (size_t)spb * F * F4 * sizeof(float);
  if (smem > 48 * 1024) return set_error("interact_bwd: shared memory %zu too large", smem);
  const long long grid = (batch + spb - 1) / spb;
#define DLRM_IB(KERNEL, MAXF)                                                                                 \
  if (route)                                                                                                  \
    (void)launch_chain(KERNEL<MAXF, FeatRoute>, dim3((unsigned)grid), dim3(128), smem, st, T, (long long)ldt, \
                       dR, (long long)lddr, dT, (long long)lddt, (long long)batch, F, D, itself,               \
                       mask_feature0, spb, gh, gl, (long long)ld_g0, *route);                                  \
  else                                                                                                        \
    (void)launch_chain(KERNEL<MAXF, NoRoute>, dim3((unsigned)grid), dim3(128), smem, st, T, (long long)ldt,   \
                       dR, (long long)lddr, dT, (long long)lddt, (long long)batch, F, D, itself,               \
                       mask_feature0, spb, gh, gl, (long long)ld_g0, NoRoute{})
  if (two) {
    if (F <= 8) { DLRM_IB(interact_bwd2_kernel, 8); }
    else if (F <= 32) { DLRM_IB(interact_bwd2_kernel, 32); }
    else { DLRM_IB(interact_bwd2_kernel, 64); }
  } else {
    if (F <= 8) { DLRM_IB(interact_bwd_kernel, 8); }
    else if (F <= 32) { DLRM_IB(interact_bwd_kernel, 32); }
    else { DLRM_IB(interact_bwd_kernel, 64); }
  }
#undef DLRM_IB
  DLRM_CHECK_LAUNCH("interact_bwd_kernel");
  return 0;
}
}  // namespace dlrm

extern "C" int dlrm_b200_interact_bwd_ex(const float* T, int64_t ldt, const float* dR, int64_t lddr,
                                         float* dT, int64_t lddt, int64_t batch, int num_features,
                                         int dim, int itself, int mask_feature0, void* g0_hi, void* g0_lo,
                                         int64_t ld_g0, void* stream) {
  return dlrm::interact_bwd_launch(T, ldt, dR, lddr, dT, lddt, nullptr, batch, num_features, dim, itself,
                                   mask_feature0, g0_hi, g0_lo, ld_g0, stream);
}

extern "C" int dlrm_b200_interact_bwd_p2p(const float* T, int64_t ldt, const float* dR, int64_t lddr,
                                          void* const* feat_dst, const int64_t* feat_ld, const int* feat_first,
                                          float emb_grad_scale, int64_t batch, int num_features, int dim, int itself,
                                          int mask_feature0, void* g0_hi, void* g0_lo, int64_t ld_g0, void* stream) {
  using namespace dlrm;
  if (!feat_dst || !feat_ld || !feat_first) return set_error("interact_bwd_p2p: NULL route");
  if (num_features > 64) return set_error("interact_bwd_p2p: num_features=%d > 64", num_features);
  const int ndst = feat_first[num_features];
  if (feat_first[0] != 0 || ndst < num_features || ndst > ROUTE_MAX_DST)
    return set_error("interact_bwd_p2p: %d destinations for %d features (max %d)", ndst, num_features, ROUTE_MAX_DST);
  FeatRoute r = {};
  r.emb_scale = emb_grad_scale;
  for (int i = 0; i <= num_features; ++i) {
    if (i && feat_first[i] <= feat_first[i - 1]) return set_error("interact_bwd_p2p: feature %d has no destination", i - 1);
    r.first[i] = (unsigned char)feat_first[i];
  }
  for (int q = 0; q < ndst; ++q) {
    if (!feat_dst[q]) return set_error("interact_bwd_p2p: feat_dst[%d] is NULL", q);
    r.base[q] = static_cast<float*>(feat_dst[q]);
    r.ld[q] = feat_ld[q];
  }
  return interact_bwd_launch(T, ldt, dR, lddr, nullptr, 0, &r, batch, num_features, dim, itself,
                             mask_feature0, g0_hi, g0_lo, ld_g0, stream);
}

extern "C" int dlrm_b200_interact_bwd(const float* T, int64_t ldt, const float* dR, int64_t lddr,
                                      float* dT, int64_t lddt, int64_t batch, int num_features,
                                      int dim, int itself, int mask_feature0, void* stream) {
  return dlrm_b200_interact_bwd_ex(T, ldt, dR, lddr, dT, lddt, batch, num_features, dim, itself,
                                   mask_feature0, nullptr, nullptr, 0, stream);
}
