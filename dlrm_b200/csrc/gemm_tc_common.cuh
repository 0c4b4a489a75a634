// Shared pieces of the tcgen05 GEMM kernels (gemm_tc.cu: one tile per CTA; gemm_chain.cu: persistent
// tile-dataflow kernel): tile constants, the problem description, PTX wrappers (mbarrier, TMA, tcgen05),
// the UMMA shared-memory descriptor and the plan object the C ABI hands out.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"

namespace dlrm {

constexpr int TC_BM = 128;
constexpr int TC_BK = 64;  // 64 bf16 = 128 bytes = one swizzle row

struct TcArgs {
  long long M, N, K;
  int x3;
  int a_mn, b_mn;  // operand majorness (0 = K-major, 1 = MN-major)
  int kb_per_split, num_kb;
  int act;
  int mask_act;
  const __nv_bfloat16* mask_hi;
  const __nv_bfloat16* mask_lo;
  long long ldmask;
  float* out_f32;
  long long ld_f32, slab_stride;
  __nv_bfloat16* out_hi;
  __nv_bfloat16* out_lo;
  long long ld_out;
  __nv_bfloat16* outT_hi;
  __nv_bfloat16* outT_lo;
  long long ld_outT;
  float* out_col;
  long long col_index, col_slab_stride;
  const float* bias;   // optional fp32 [N], added to the accumulator before the activation
};

// ------------------------------------------------------------------------------------------ PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  // try_wait suspends for a bounded time per call; a protocol bug must trap (after 2 s of wall clock),
  // not hang the GPU
  uint32_t ok = 0;
  unsigned long long t0 = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P1;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) break;
    const unsigned long long now = globaltimer_ns();
    if (t0 == 0) t0 = now;
    else if (now - t0 > 2000000000ull) __trap();
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int x, int y) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(x), "r"(y)
      : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// UMMA shared-memory descriptor (cute::UMMA::SmemDescriptor): start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version=1 [46,48), layout SWIZZLE_128B=2 [61,64).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__device__ __forceinline__ float apply_act_tc(float v, int act) {
  if (act == DLRM_ACT_RELU) return fmaxf(v, 0.f);
  if (act == DLRM_ACT_SIGMOID) return 1.0f / (1.0f + expf(-v));
  return v;
}


// ------------------------------------------------------------------------------------------ epilogue
// One accumulator tile (128 rows x bn columns in tensor memory) -> global memory, shared by the per-layer
// kernel (gemm_tc.cu) and the persistent chain kernel (gemm_chain.cu).
//
// tcgen05.ld hands every thread ONE ROW (32 consecutive columns per load).  Storing from that layout makes
// each warp-wide 16-byte store hit 32 different rows (32 half-written sectors per instruction); measured with
// the chain kernel's per-task timeline (tools/chain_timeline.py, profiles/r2_chain_timeline_before.txt) that
// cost 10-19 us per 128 x 128 tile -- more than the tile's MMAs.  Here every 32 x 32 chunk is transposed
// through a per-warp shared-memory tile (padded rows: conflict-free 16-byte accesses) so that a warp store
// covers whole row segments: 8 rows x 64 B for the bf16 (hi, lo) operands, 4 rows x 128 B for fp32; the
// activation-gradient mask is read the same way.  Chunks that are ragged (N tail, the diverted bias-gradient
// column) or whose rows are not 16-byte aligned use element-wise but still row-contiguous accesses.
constexpr int TC_EPI_ROW_F32 = 20;                 // floats per staged fp32 half-row (16 + 4 pad)
constexpr int TC_EPI_ROW_BF16 = 40;                // bf16 per staged bf16 row (32 + 8 pad)
constexpr int TC_EPI_WARP_BYTES = 32 * TC_EPI_ROW_BF16 * 2;  // 2560 B per epilogue warp (fp32: two 16-column passes)
// Epilogue warps per CTA: two per TMEM lane quadrant (warps w and w + 4 may both read quadrant w % 4) split the
// 32-column chunks of a tile.  The epilogue is a long dependent instruction stream per warp (tcgen05.ld -> convert
// -> shared-memory transpose -> store) with ONE warp per scheduler, i.e. issue-latency bound: a second warp per
// scheduler halves it (chain timeline: 4-10 us -> 3-6 us per tile).
constexpr int TC_EPI_WARPS = 8;
constexpr int TC_THREADS = 64 + 32 * TC_EPI_WARPS;       // warp 0: TMA, warp 1: MMA, then the epilogue warps
constexpr int TC_EPI_BYTES = TC_EPI_WARPS * TC_EPI_WARP_BYTES;

// 32 x 32 bf16 chunk, one row per thread in `mine` -> global rows [mrow0, mrow0 + 32) x columns [nb, nb + 32)
__device__ __forceinline__ void tc_epi_store_bf16(const __nv_bfloat16 (&mine)[32], __nv_bfloat16* dst, long long ld,
                                                  long long mrow0, long long nb, long long M, long long N, bool vec,
                                                  int lane, __nv_bfloat16* sb) {
#pragma unroll
  for (int q = 0; q < 4; ++q)
    *reinterpret_cast<uint4*>(sb + lane * TC_EPI_ROW_BF16 + q * 8) = reinterpret_cast<const uint4*>(mine)[q];
  __syncwarp();
  if (vec) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {     // 8 rows x 64 B per warp store
      const int row = i * 8 + (lane >> 2), seg = lane & 3;
      if (mrow0 + row < M)
        *reinterpret_cast<uint4*>(dst + (mrow0 + row) * ld + nb + seg * 8) =
            *reinterpret_cast<const uint4*>(sb + row * TC_EPI_ROW_BF16 + seg * 8);
    }
  } else {
    const long long col = nb + lane;
    const int nrow = (int)min(32ll, M - mrow0);
    if (col < N)
      for (int row = 0; row < nrow; ++row) dst[(mrow0 + row) * ld + col] = sb[row * TC_EPI_ROW_BF16 + lane];
  }
  __syncwarp();
}

// c0 / cstep: this warp handles the 32-column chunks c0, c0 + cstep, ... (two warps per TMEM lane quadrant split
// the chunks of a tile between them in the chain kernel; the per-layer kernel passes 0, 1).
__device__ __forceinline__ void tc_epilogue_tile(const TcArgs& g, int bn, int m0, int n0, int bz, uint32_t tmem_acc,
                                                 int quad, int lane, uint8_t* stage_warp, int c0 = 0, int cstep = 1) {
  // Every field is copied into a register ONCE: `g` lives in kernel-parameter space (indexed at run time in the
  // chain kernel) and the tcgen05 / mbarrier asm statements clobber memory, so reading g.act or g.bias inside the
  // element loops costs a constant-bank load + a branch PER ELEMENT (measured: 3-7 us per 32 x 32 chunk).
  const long long M = g.M, N = g.N;
  const int act = g.act, mask_act = g.mask_act;
  const __nv_bfloat16* const mask_hi = g.mask_hi;
  const __nv_bfloat16* const mask_lo = g.mask_lo;
  const long long ldmask = g.ldmask, ld_f32 = g.ld_f32, ld_out = g.ld_out, ld_outT = g.ld_outT;
  __nv_bfloat16* const out_hi = g.out_hi;
  __nv_bfloat16* const out_lo = g.out_lo;
  __nv_bfloat16* const outT_hi = g.outT_hi;
  __nv_bfloat16* const outT_lo = g.outT_lo;
  const float* const bias = g.bias;
  const long long col_index = g.col_index;
  float* const of32 = g.out_f32 ? g.out_f32 + (long long)bz * g.slab_stride : nullptr;
  float* const ocol = g.out_col ? g.out_col + (long long)bz * g.col_slab_stride : nullptr;

  const long long mrow0 = (long long)m0 + quad * 32;   // first row of this warp
  const long long m = mrow0 + lane;
  const bool m_ok = m < M;
  float* sf = reinterpret_cast<float*>(stage_warp);
  __nv_bfloat16* sb = reinterpret_cast<__nv_bfloat16*>(stage_warp);
  const bool f32_vec = of32 && (ld_f32 & 3) == 0 && (reinterpret_cast<uintptr_t>(of32) & 15) == 0;
  const bool bf_vec = out_hi && (ld_out & 7) == 0 && (reinterpret_cast<uintptr_t>(out_hi) & 15) == 0 &&
                      (!out_lo || (reinterpret_cast<uintptr_t>(out_lo) & 15) == 0);
  const bool mask_vec = mask_act != DLRM_ACT_NONE && (ldmask & 7) == 0 &&
                        (reinterpret_cast<uintptr_t>(mask_hi) & 15) == 0 &&
                        (!mask_lo || (reinterpret_cast<uintptr_t>(mask_lo) & 15) == 0);
  const bool bias_vec = bias && (reinterpret_cast<uintptr_t>(bias) & 15) == 0;
#pragma unroll 1
  for (int c = c0; c < bn / 32; c += cstep) {
    const long long nb = (long long)n0 + c * 32;
    if (nb >= N) break;
    uint32_t r[32];
    tmem_ld32(tmem_acc + ((uint32_t)(quad * 32) << 16) + (uint32_t)(c * 32), r);
    float v[32];
    const bool full = nb + 32 <= N;
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
    if (bias) {       // nn.Linear bias: the same 32 values for every row -> 8 broadcast 16-byte loads
      if (full && bias_vec) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + nb) + q);
          v[4 * q] += b4.x; v[4 * q + 1] += b4.y; v[4 * q + 2] += b4.z; v[4 * q + 3] += b4.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += (nb + j < N) ? __ldg(bias + nb + j) : 0.f;
      }
    }
    if (act == DLRM_ACT_RELU) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
    } else if (act == DLRM_ACT_SIGMOID) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = 1.0f / (1.0f + expf(-v[j]));
    }
    // ---------------------------------------------------------------- activation-gradient mask
    if (mask_act != DLRM_ACT_NONE) {
      if (full && mask_vec) {
        const int passes = (mask_act == DLRM_ACT_SIGMOID && mask_lo) ? 2 : 1;
        for (int pass = 0; pass < passes; ++pass) {
          const __nv_bfloat16* src = pass ? mask_lo : mask_hi;
#pragma unroll
          for (int i = 0; i < 4; ++i) {       // 8 rows x 64 B per warp load
            const int row = i * 8 + (lane >> 2), seg = lane & 3;
            uint4 t = make_uint4(0, 0, 0, 0);
            if (mrow0 + row < M) t = *reinterpret_cast<const uint4*>(src + (mrow0 + row) * ldmask + nb + seg * 8);
            *reinterpret_cast<uint4*>(sb + row * TC_EPI_ROW_BF16 + seg * 8) = t;
          }
          __syncwarp();
          __align__(16) __nv_bfloat16 y[32];
#pragma unroll
          for (int q = 0; q < 4; ++q)
            reinterpret_cast<uint4*>(y)[q] = *reinterpret_cast<const uint4*>(sb + lane * TC_EPI_ROW_BF16 + q * 8);
          __syncwarp();
          if (mask_act == DLRM_ACT_RELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __bfloat162float(y[j]) > 0.f ? v[j] : 0.f;
          } else if (passes == 1) {
#pragma unroll
            for (int j = 0; j < 32; ++j) { const float yy = __bfloat162float(y[j]); v[j] *= (1.0f - yy) * yy; }
          } else if (pass == 0) {
#pragma unroll
            for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__bfloat162float(y[j]));   // keep hi, add lo next pass
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float yy = __uint_as_float(r[j]) + __bfloat162float(y[j]);
              v[j] *= (1.0f - yy) * yy;
            }
          }
        }
      } else if (m_ok) {
        for (int j = 0; j < 32; ++j) {
          if (nb + j < N) {
            const long long o = m * ldmask + nb + j;
            float y = __bfloat162float(mask_hi[o]);
            if (mask_act == DLRM_ACT_RELU) {
              v[j] = y > 0.f ? v[j] : 0.f;
            } else {
              if (mask_lo) y += __bfloat162float(mask_lo[o]);
              v[j] *= (1.0f - y) * y;
            }
          }
        }
      }
    }
    // ---------------------------------------------------------------- fp32 output (+ diverted column)
    if (of32) {
      const bool vec = full && f32_vec && (!ocol || nb + 32 <= col_index);
      const int nrow = (int)min(32ll, M - mrow0);
#pragma unroll
      for (int h = 0; h < 2; ++h) {           // two 16-column passes through the 2560-byte staging tile
#pragma unroll
        for (int q = 0; q < 4; ++q)
          *reinterpret_cast<float4*>(sf + lane * TC_EPI_ROW_F32 + q * 4) =
              make_float4(v[16 * h + 4 * q], v[16 * h + 4 * q + 1], v[16 * h + 4 * q + 2], v[16 * h + 4 * q + 3]);
        __syncwarp();
        if (vec) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {       // 8 rows x 64 B per warp store
            const int row = i * 8 + (lane >> 2), seg = lane & 3;
            if (row < nrow)
              *reinterpret_cast<float4*>(of32 + (mrow0 + row) * ld_f32 + nb + 16 * h + seg * 4) =
                  *reinterpret_cast<const float4*>(sf + row * TC_EPI_ROW_F32 + seg * 4);
          }
        } else {                              // 2 rows x 16 consecutive floats per store
          const long long col = nb + 16 * h + (lane & 15);
          const bool to_col = ocol && col == col_index;
          const bool to_out = col < N && (!ocol || col < col_index);
          for (int row = lane >> 4; row < nrow; row += 2) {
            const float x = sf[row * TC_EPI_ROW_F32 + (lane & 15)];
            if (to_col) ocol[mrow0 + row] = x;
            else if (to_out) of32[(mrow0 + row) * ld_f32 + col] = x;
          }
        }
        __syncwarp();
      }
    }
    // ---------------------------------------------------------------- (hi, lo) bf16 operand outputs
    if (out_hi || outT_hi) {
      __align__(16) __nv_bfloat16 hi[32], lo[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        hi[j] = __float2bfloat16_rn(v[j]);
        lo[j] = __float2bfloat16_rn(v[j] - __bfloat162float(hi[j]));
      }
      if (out_hi) {
        tc_epi_store_bf16(hi, out_hi, ld_out, mrow0, nb, M, N, full && bf_vec, lane, sb);
        if (out_lo) tc_epi_store_bf16(lo, out_lo, ld_out, mrow0, nb, M, N, full && bf_vec, lane, sb);
      }
      if (outT_hi && m_ok) {                  // transposed copy: consecutive lanes = consecutive rows = contiguous
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          if (full || nb + j < N) {
            const long long o = (nb + j) * ld_outT + m;
            outT_hi[o] = hi[j];
            if (outT_lo) outT_lo[o] = lo[j];
          }
        }
      }
    }
  }
}

struct TcPlan {
  CUtensorMap tmAh, tmAl, tmBh, tmBl;
  TcArgs args;
  int bn, stages, splits;
  size_t smem;
  dim3 grid;
};

}  // namespace dlrm
