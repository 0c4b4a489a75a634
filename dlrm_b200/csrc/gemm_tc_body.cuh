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
    // ------------------------------------------------------------------ epilogue (warps 2..)
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access
    mbar_wait(bar_base + 8 * (2 * stages), 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint8_t* stage_warp = smem + (size_t)stages * stage_bytes + 256 + (size_t)(warp - 2) * TC_EPI_WARP_BYTES;
    tc_epilogue_tile(g, BN, m0, n0, TCB_BZ, tmem_base, quad, lane, stage_warp, (warp - 2) >> 2, TC_EPI_WARPS / 4);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"((uint32_t)(BN < 32 ? 32 : BN))
                 : "memory");
  }
