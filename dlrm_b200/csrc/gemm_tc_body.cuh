// Body of the tcgen05 GEMM CTA (see gemm_tc.cu for the design).  Textually included by the kernels
// that share it; the including kernel provides, as macros or locals:
//   TCB_BX / TCB_BY / TCB_BZ        tile coordinates (n tile, m tile, k split)
//   TCB_MAP_AH / _AL / _BH / _BL    const CUtensorMap* of the four operand maps (kernel-parameter space)
//   g       const TcArgs&           problem description
//   stages  int                     depth of the operand ring
//   BN      template int            tile width
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr uint32_t A_BYTES = TC_BM * TC_BK * 2;  // 16 KB
  constexpr uint32_t B_BYTES = BN * TC_BK * 2;
  const uint32_t stage_bytes = (g.x3 ? 2u : 1u) * (A_BYTES + B_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)stages * stage_bytes);
  // bars[0..stages) full, [stages..2*stages) empty, [2*stages] accumulator ready
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * stages + 1);

  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = TCB_BY * TC_BM, n0 = TCB_BX * BN;
  const int kb0 = TCB_BZ * g.kb_per_split;
  const int kb1 = min(g.num_kb, kb0 + g.kb_per_split);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bar_base = smem_u32(bars);

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(bar_base + 8 * s, 1);
      mbar_init(bar_base + 8 * (stages + s), 1);
    }
    mbar_init(bar_base + 8 * (2 * stages), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    // allocate BN fp32 accumulator columns (power of two >= 32)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)(BN < 32 ? 32 : BN))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();   // barriers and TMEM are set up; global memory is first touched below

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(bar_base + 8 * (stages + stage), phase ^ 1);
        const uint32_t full = bar_base + 8 * stage;
        mbar_expect_tx(full, stage_bytes);
        uint32_t dst = smem_base + stage * stage_bytes;
        const int k0 = kb * TC_BK;
        // A tile(s)
        for (int part = 0; part < (g.x3 ? 2 : 1); ++part) {
          const CUtensorMap* map = part ? TCB_MAP_AL : TCB_MAP_AH;
          if (!g.a_mn) {
            tma_load_2d(dst, map, full, k0, m0);                       // box {64 k, 128 m}
          } else {
            tma_load_2d(dst, map, full, m0, k0);                       // box {64 m, 64 k} x 2
            tma_load_2d(dst + A_BYTES / 2, map, full, m0 + 64, k0);
          }
          dst += A_BYTES;
        }
        for (int part = 0; part < (g.x3 ? 2 : 1); ++part) {
          const CUtensorMap* map = part ? TCB_MAP_BL : TCB_MAP_BH;
          if (!g.b_mn) {
            tma_load_2d(dst, map, full, k0, n0);                       // box {64 k, BN n}
          } else {
#pragma unroll
            for (int h = 0; h < BN / 64; ++h) tma_load_2d(dst + h * 8192, map, full, n0 + 64 * h, k0);
          }
          dst += B_BYTES;
        }
        if (++stage == stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    // instruction descriptor: D=f32 (1<<4), A=B=bf16 (1<<7, 1<<10), majorness bits 15/16,
    // N>>3 at [17,23), M>>4 at [24,29)
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)g.a_mn << 15) |
                           ((uint32_t)g.b_mn << 16) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
    int stage = 0;
    uint32_t phase = 0;
    uint32_t accum = 0;
    for (int kb = kb0; kb < kb1; ++kb) {
      mbar_wait(bar_base + 8 * stage, phase);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (lane == 0) {
        const uint32_t sa_hi = smem_base + stage * stage_bytes;
        const uint32_t sa_lo = sa_hi + A_BYTES;
        const uint32_t sb_hi = sa_hi + (g.x3 ? 2u : 1u) * A_BYTES;
        const uint32_t sb_lo = sb_hi + B_BYTES;
#pragma unroll
        for (int k = 0; k < TC_BK / 16; ++k) {
          // K-major SW128: rows of 128 B, 8-row groups 1024 B apart (SBO), k-step = +32 B.
          // MN-major SW128: [64 k rows][64 mn] boxes: k-groups 1024 B apart (SBO), 64-wide mn blocks
          //                 8192 B apart (LBO), k-step (16 rows) = +2048 B.
          const uint32_t a_off = g.a_mn ? k * 2048u : k * 32u;
          const uint32_t b_off = g.b_mn ? k * 2048u : k * 32u;
          const uint32_t a_lbo = g.a_mn ? 8192u : 16u, b_lbo = g.b_mn ? 8192u : 16u;
          const uint64_t ah = make_smem_desc(sa_hi + a_off, a_lbo, 1024);
          const uint64_t bh = make_smem_desc(sb_hi + b_off, b_lbo, 1024);
          if (g.x3) {
            const uint64_t al = make_smem_desc(sa_lo + a_off, a_lbo, 1024);
            const uint64_t bl = make_smem_desc(sb_lo + b_off, b_lbo, 1024);
            umma_bf16(tmem_base, al, bh, idesc, accum);
            umma_bf16(tmem_base, ah, bl, idesc, 1u);
            umma_bf16(tmem_base, ah, bh, idesc, 1u);
          } else {
            umma_bf16(tmem_base, ah, bh, idesc, accum);
          }
          accum = 1u;
        }
        umma_commit(bar_base + 8 * (stages + stage));              // smem stage free when MMAs retire
        if (kb == kb1 - 1) umma_commit(bar_base + 8 * (2 * stages));  // accumulator complete
      }
      __syncwarp();
      if (++stage == stages) { stage = 0; phase ^= 1; }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..5)
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access
    mbar_wait(bar_base + 8 * (2 * stages), 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const long long m = (long long)m0 + quad * 32 + lane;
    const bool m_ok = m < g.M;
    float* of32 = g.out_f32 ? g.out_f32 + (long long)TCB_BZ * g.slab_stride : nullptr;
    float* ocol = g.out_col ? g.out_col + (long long)TCB_BZ * g.col_slab_stride : nullptr;
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t r[32];
      tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(c * 32), r);
      const long long nb = (long long)n0 + c * 32;
      if (nb >= g.N) break;
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = apply_act_tc(__uint_as_float(r[j]), g.act);
      const bool full = nb + 32 <= g.N;
      if (g.mask_act != DLRM_ACT_NONE && m_ok) {
        if (full && (g.ldmask & 7) == 0) {
          // 32 bf16 of this row = 4 x 16-byte loads (hi), + 4 (lo) for sigmoid'
          __align__(16) __nv_bfloat16 yh[32], yl[32];
          const uint4* ph = reinterpret_cast<const uint4*>(g.mask_hi + m * g.ldmask + nb);
#pragma unroll
          for (int q = 0; q < 4; ++q) reinterpret_cast<uint4*>(yh)[q] = ph[q];
          if (g.mask_act == DLRM_ACT_RELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __bfloat162float(yh[j]) > 0.f ? v[j] : 0.f;
          } else {
            if (g.mask_lo) {
              const uint4* pl = reinterpret_cast<const uint4*>(g.mask_lo + m * g.ldmask + nb);
#pragma unroll
              for (int q = 0; q < 4; ++q) reinterpret_cast<uint4*>(yl)[q] = pl[q];
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              float y = __bfloat162float(yh[j]);
              if (g.mask_lo) y += __bfloat162float(yl[j]);
              v[j] *= (1.0f - y) * y;
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (full || nb + j < g.N) {
              const long long o = m * g.ldmask + nb + j;
              float y = __bfloat162float(g.mask_hi[o]);
              if (g.mask_act == DLRM_ACT_RELU) {
                v[j] = y > 0.f ? v[j] : 0.f;
              } else {
                if (g.mask_lo) y += __bfloat162float(g.mask_lo[o]);
                v[j] *= (1.0f - y) * y;
              }
            }
          }
        }
      }
      if (of32 && m_ok) {
        float* p = of32 + m * g.ld_f32 + nb;
        const bool colsplit = ocol != nullptr && g.col_index >= nb && g.col_index < nb + 32;
        if (full && !colsplit && (g.ld_f32 & 3) == 0) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(p + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (nb + j < g.N) {
              if (ocol && nb + j == g.col_index) ocol[m] = v[j];
              else if (!ocol || nb + j < g.col_index) p[j] = v[j];
            }
          }
        }
      }
      if (g.out_hi || g.outT_hi) {
        __nv_bfloat16 hi[32], lo[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          hi[j] = __float2bfloat16_rn(v[j]);
          lo[j] = __float2bfloat16_rn(v[j] - __bfloat162float(hi[j]));
        }
        if (g.out_hi && m_ok) {
          __nv_bfloat16* ph = g.out_hi + m * g.ld_out + nb;
          __nv_bfloat16* pl = g.out_lo ? g.out_lo + m * g.ld_out + nb : nullptr;
          if (full) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              *reinterpret_cast<uint4*>(ph + j) = *reinterpret_cast<const uint4*>(&hi[j]);
              if (pl) *reinterpret_cast<uint4*>(pl + j) = *reinterpret_cast<const uint4*>(&lo[j]);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (nb + j < g.N) { ph[j] = hi[j]; if (pl) pl[j] = lo[j]; }
          }
        }
        if (g.outT_hi && m_ok) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (full || nb + j < g.N) {
              const long long o = (nb + j) * g.ld_outT + m;
              g.outT_hi[o] = hi[j];
              if (g.outT_lo) g.outT_lo[o] = lo[j];
            }
          }
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"((uint32_t)(BN < 32 ? 32 : BN))
                 : "memory");
  }
