// Persistent tile-dataflow tcgen05 kernel: a whole MLP chain (forward layers, or the dgrad chain with
// all weight-gradient GEMMs beside it) in ONE launch.
//
// Why: at batch 2048 every layer of the DLRM MLPs is a 32-144-tile GEMM whose math takes 1-3 us; as one
// launch per layer each paid launch + prologue (TMEM allocation, barrier init, descriptor fetch) +
// pipeline fill + drain, 10-30 us apiece, 17 launches per step (profiles/r1_final_launches_tc_step.csv:
// 278 of 540 us).  Here the tiles of ALL layers of a chain are one topologically ordered task list:
//
//   * grid = one CTA per SM (or fewer); warp 0 claims the next task from a device-side queue
//     (atomicAdd, claimed one task ahead so the round trip is hidden) and publishes it to the other
//     warps through a 4-deep shared-memory ring guarded by mbarriers;
//   * a task whose A operand is produced inside this launch (layer l+1 reads layer l's activations;
//     a weight-gradient GEMM reads the dgrad chain's gz) spins on per-(producer, 128-row block)
//     completion counters (ld.acquire.gpu), then issues a async-proxy fence before its TMA loads.  Tasks
//     are claimed in topological order and a claimed task depends only on lower-numbered tasks, which are
//     owned by resident CTAs: no deadlock whatever the number of resident CTAs;
//   * warp 1 issues tcgen05.mma into one of 4 TMEM accumulators (512 columns allocated once), warps 2-5
//     run the epilogue of tile i (tcgen05.ld, activation / activation-gradient mask, fp32 and (hi,lo) bf16
//     stores) while the MMAs of tile i+1 run; after their stores they bump the completion counter
//     (fence + named barrier + release reduction);
//   * the last CTA to leave zeroes queue + counters, so a CUDA-graph replay needs no memset node.
//
// The CTA program per tile (TMA boxes, UMMA descriptors, epilogue) is the one of gemm_tc_body.cuh, so the
// results are bit-identical to the per-layer launches.
#include <string.h>

#include "gemm_tc_common.cuh"

namespace dlrm {

constexpr int CH_MAX_PROBLEMS = 16;
constexpr int CH_ACC_STAGES = 4;    // 4 x 128 TMEM columns
constexpr int CH_RING = 4;

struct ChProblem {
  CUtensorMap m[4];   // A_hi, A_lo, B_hi, B_lo
  TcArgs a;
  int bn;
  int gx, gy, gz;     // n tiles, m tiles, k splits
  int task_begin;
  int dep;            // producing problem or -1
  int dep_on_k;       // 0: A rows = this task's m tile; 1: A rows = this task's k range
  int dep_target;     // tiles per 128-row block of the producer
  int ctr_base;       // this problem's counters: ctr[2 + ctr_base + m_tile]
  int signal;         // some later problem depends on this one
};

// Task order.  A group of consecutive single-split problems with the same number of m tiles (the layers of a
// forward MLP, the dgrad chain) is enumerated M-TILE MAJOR: (m, problem, n tile).  Layer l+1 of m tile 0 is then
// claimed right after layer l of m tile 0 and can start as soon as those few tiles are done, while other CTAs
// work on the next m tiles: the layers pipeline across the batch instead of running as one wave per layer.
// Any other problem is a group of its own, enumerated (k split, m, n).
struct ChGroup {
  int task_begin, p0, np, per_m;
};

struct ChParams {
  ChProblem p[CH_MAX_PROBLEMS];
  ChGroup grp[CH_MAX_PROBLEMS];
  int ngroups;
  int n, total, stages;
  uint32_t stage_bytes;
  int* ctr;           // [0] task queue, [1] exit count, [2...] completion counters
  int n_ctr;
  // optional per-task timeline (globaltimer ns), 8 words per task: claim, deps ready, last TMA issued,
  // first operands landed, last MMA issued, accumulator ready, epilogue + signal done, SM id
  unsigned long long* trace;
};

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.b32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add(int* p, int v) {
  asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

struct ChTask {
  int pi, bx, by, bz;
};
__device__ __forceinline__ ChTask ch_decode(const ChParams& P, int t) {
  ChTask k;
  int gi = 0;
  while (gi + 1 < P.ngroups && t >= P.grp[gi + 1].task_begin) ++gi;
  const ChGroup& G = P.grp[gi];
  const int local = t - G.task_begin;
  if (G.np == 1) {
    const ChProblem& Q = P.p[G.p0];
    k.pi = G.p0;
    k.bx = local % Q.gx;
    k.by = (local / Q.gx) % Q.gy;
    k.bz = local / (Q.gx * Q.gy);
  } else {
    k.by = local / G.per_m;
    int r = local - k.by * G.per_m;
    k.pi = G.p0;
    while (r >= P.p[k.pi].gx) { r -= P.p[k.pi].gx; ++k.pi; }
    k.bx = r;
    k.bz = 0;
  }
  return k;
}

constexpr int CH_EPI_WARPS = TC_EPI_WARPS;
constexpr int CH_THREADS = TC_THREADS;

__global__ void __launch_bounds__(CH_THREADS, 1) gemm_chain_kernel(const __grid_constant__ ChParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int stages = P.stages;
  const uint32_t stage_bytes = P.stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)stages * stage_bytes);
  // bars: full[stages] | empty[stages] | acc_full[4] | acc_empty[4] | task_full[4] | task_empty[4]
  int* ring = reinterpret_cast<int*>(bars + 2 * stages + 2 * CH_ACC_STAGES + 2 * CH_RING);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ring + CH_RING);
  uint8_t* epi_stage = smem + (size_t)stages * stage_bytes + 512;   // barriers + ring + slot < 512 B

  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bar_full = smem_u32(bars);
  const uint32_t bar_empty = bar_full + 8 * stages;
  const uint32_t bar_acc_full = bar_empty + 8 * stages;
  const uint32_t bar_acc_empty = bar_acc_full + 8 * CH_ACC_STAGES;
  const uint32_t bar_task_full = bar_acc_empty + 8 * CH_ACC_STAGES;
  const uint32_t bar_task_empty = bar_task_full + 8 * CH_RING;

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int s = 0; s < CH_ACC_STAGES; ++s) {
      mbar_init(bar_acc_full + 8 * s, 1);
      mbar_init(bar_acc_empty + 8 * s, CH_EPI_WARPS);     // lane 0 of each epilogue warp
    }
    for (int s = 0; s < CH_RING; ++s) {
      mbar_init(bar_task_full + 8 * s, 1);
      mbar_init(bar_task_empty + 8 * s, 1 + CH_EPI_WARPS);    // MMA warp + epilogue warps
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();   // barriers and TMEM are set up; global memory is first touched below

  if (warp == 0) {
    // ------------------------------------------------------------------ task claim + TMA producer
    if (lane == 0) {
      int stage = 0, slot = 0;
      uint32_t phase = 0, tph = 0;
      int t = atomicAdd(P.ctr, 1);
      while (true) {
        mbar_wait(bar_task_empty + 8 * slot, tph ^ 1);
        ring[slot] = t;
        mbar_arrive(bar_task_full + 8 * slot);
        if (++slot == CH_RING) { slot = 0; tph ^= 1; }
        if (t >= P.total) break;
        if (P.trace) P.trace[(size_t)t * 8 + 0] = globaltimer_ns();
        const int tnext = atomicAdd(P.ctr, 1);    // claimed one task ahead: the round trip hides behind this task
        const ChTask k = ch_decode(P, t);
        const ChProblem& Q = P.p[k.pi];
        const TcArgs& g = Q.a;
        const uint32_t A_BYTES = TC_BM * TC_BK * 2;
        const uint32_t B_BYTES = (uint32_t)Q.bn * TC_BK * 2;
        const uint32_t tx = (g.x3 ? 2u : 1u) * (A_BYTES + B_BYTES);
        const int m0 = k.by * TC_BM, n0 = k.bx * Q.bn;
        const int kb0 = k.bz * g.kb_per_split;
        const int kb1 = min(g.num_kb, kb0 + g.kb_per_split);
        if (Q.dep >= 0) {
          long long r0, r1;
          if (Q.dep_on_k) { r0 = (long long)kb0 * TC_BK; r1 = min((long long)kb1 * TC_BK, g.K); }
          else { r0 = m0; r1 = min((long long)m0 + TC_BM, g.M); }
          const int* cbase = P.ctr + 2 + P.p[Q.dep].ctr_base;
          for (int c = (int)(r0 / TC_BM); c < (int)((r1 + TC_BM - 1) / TC_BM); ++c) {
            unsigned long long t0 = 0;
            while (ld_acquire_gpu(cbase + c) < Q.dep_target) {
              __nanosleep(40);
              const unsigned long long now = globaltimer_ns();
              if (t0 == 0) t0 = now;
              else if (now - t0 > 2000000000ull) __trap();   // a protocol bug must trap, not hang the GPU
            }
          }
          asm volatile("fence.proxy.async;" ::: "memory");   // producer's generic-proxy stores -> our TMA reads
        }
        if (P.trace) P.trace[(size_t)t * 8 + 1] = globaltimer_ns();
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(bar_empty + 8 * stage, phase ^ 1);
          const uint32_t full = bar_full + 8 * stage;
          mbar_expect_tx(full, tx);
          uint32_t dst = smem_base + stage * stage_bytes;
          const int k0 = kb * TC_BK;
          for (int part = 0; part < (g.x3 ? 2 : 1); ++part) {
            const CUtensorMap* map = &Q.m[part];
            if (!g.a_mn) {
              tma_load_2d(dst, map, full, k0, m0);
            } else {
              tma_load_2d(dst, map, full, m0, k0);
              tma_load_2d(dst + A_BYTES / 2, map, full, m0 + 64, k0);
            }
            dst += A_BYTES;
          }
          for (int part = 0; part < (g.x3 ? 2 : 1); ++part) {
            const CUtensorMap* map = &Q.m[2 + part];
            if (!g.b_mn) {
              tma_load_2d(dst, map, full, k0, n0);
            } else {
              for (int h = 0; h < Q.bn / 64; ++h) tma_load_2d(dst + h * 8192, map, full, n0 + 64 * h, k0);
            }
            dst += B_BYTES;
          }
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
        if (P.trace) P.trace[(size_t)t * 8 + 2] = globaltimer_ns();
        t = tnext;
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    int stage = 0, slot = 0, as = 0;
    uint32_t phase = 0, tph = 0, aph = 0;
    while (true) {
      mbar_wait(bar_task_full + 8 * slot, tph);
      const int t = ring[slot];
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_task_empty + 8 * slot);
      if (++slot == CH_RING) { slot = 0; tph ^= 1; }
      if (t >= P.total) break;
      const ChTask k = ch_decode(P, t);
      const ChProblem& Q = P.p[k.pi];
      const TcArgs& g = Q.a;
      const uint32_t A_BYTES = TC_BM * TC_BK * 2;
      const uint32_t B_BYTES = (uint32_t)Q.bn * TC_BK * 2;
      const int kb0 = k.bz * g.kb_per_split;
      const int kb1 = min(g.num_kb, kb0 + g.kb_per_split);
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)g.a_mn << 15) |
                             ((uint32_t)g.b_mn << 16) | ((uint32_t)(Q.bn >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      const uint32_t tmem_acc = tmem_base + (uint32_t)(as * 128);
      mbar_wait(bar_acc_empty + 8 * as, aph ^ 1);     // epilogue has drained this accumulator
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      uint32_t accum = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(bar_full + 8 * stage, phase);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (lane == 0) {
          if (P.trace && kb == kb0) P.trace[(size_t)t * 8 + 3] = globaltimer_ns();
          const uint32_t sa_hi = smem_base + stage * stage_bytes;
          const uint32_t sa_lo = sa_hi + A_BYTES;
          const uint32_t sb_hi = sa_hi + (g.x3 ? 2u : 1u) * A_BYTES;
          const uint32_t sb_lo = sb_hi + B_BYTES;
#pragma unroll
          for (int kk = 0; kk < TC_BK / 16; ++kk) {
            const uint32_t a_off = g.a_mn ? kk * 2048u : kk * 32u;
            const uint32_t b_off = g.b_mn ? kk * 2048u : kk * 32u;
            const uint32_t a_lbo = g.a_mn ? 8192u : 16u, b_lbo = g.b_mn ? 8192u : 16u;
            const uint64_t ah = make_smem_desc(sa_hi + a_off, a_lbo, 1024);
            const uint64_t bh = make_smem_desc(sb_hi + b_off, b_lbo, 1024);
            if (g.x3) {
              const uint64_t al = make_smem_desc(sa_lo + a_off, a_lbo, 1024);
              const uint64_t bl = make_smem_desc(sb_lo + b_off, b_lbo, 1024);
              umma_bf16(tmem_acc, al, bh, idesc, accum);
              umma_bf16(tmem_acc, ah, bl, idesc, 1u);
              umma_bf16(tmem_acc, ah, bh, idesc, 1u);
            } else {
              umma_bf16(tmem_acc, ah, bh, idesc, accum);
            }
            accum = 1u;
          }
          umma_commit(bar_empty + 8 * stage);
          if (kb == kb1 - 1) {
            umma_commit(bar_acc_full + 8 * as);
            if (P.trace) P.trace[(size_t)t * 8 + 4] = globaltimer_ns();
          }
        }
        __syncwarp();
        if (++stage == stages) { stage = 0; phase ^= 1; }
      }
      if (++as == CH_ACC_STAGES) { as = 0; aph ^= 1; }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..9)
    const int quad = warp & 3;            // TMEM lane quadrant this warp may access
    const int half = (warp - 2) >> 2;     // which of the quadrant's two warps: even / odd 32-column chunks
    int slot = 0, as = 0;
    uint32_t tph = 0, aph = 0;
    while (true) {
      mbar_wait(bar_task_full + 8 * slot, tph);
      const int t = ring[slot];
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_task_empty + 8 * slot);
      if (++slot == CH_RING) { slot = 0; tph ^= 1; }
      if (t >= P.total) break;
      const ChTask k = ch_decode(P, t);
      const ChProblem& Q = P.p[k.pi];
      mbar_wait(bar_acc_full + 8 * as, aph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (P.trace && warp == 2 && lane == 0) P.trace[(size_t)t * 8 + 5] = globaltimer_ns();
      tc_epilogue_tile(Q.a, Q.bn, k.by * TC_BM, k.bx * Q.bn, k.bz, tmem_base + (uint32_t)(as * 128), quad, lane,
                       epi_stage + (size_t)(warp - 2) * TC_EPI_WARP_BYTES, half, CH_EPI_WARPS / 4);
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_acc_empty + 8 * as);
      if (Q.signal) {
        // grid-sync idiom: every epilogue thread has issued its stores (barrier), then ONE thread makes them
        // visible GPU-wide (cumulative fence) and publishes the tile with a release reduction
        asm volatile("bar.sync 1, %0;" ::"n"(32 * CH_EPI_WARPS) : "memory");
        if (warp == 2 && lane == 0) {
          __threadfence();
          asm volatile("fence.proxy.async;" ::: "memory");
          red_release_gpu_add(P.ctr + 2 + Q.ctr_base + k.by, 1);
        }
      }
      if (P.trace && warp == 2 && lane == 0) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        P.trace[(size_t)t * 8 + 6] = globaltimer_ns();
        P.trace[(size_t)t * 8 + 7] = smid;
      }
      if (++as == CH_ACC_STAGES) { as = 0; aph ^= 1; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
  // the last CTA to leave resets queue + counters for the next launch / graph replay
  if (threadIdx.x == 0) {
    __threadfence();
    const int left = atomicAdd(P.ctr + 1, 1);
    if (left == (int)gridDim.x - 1) {
      __threadfence();
      for (int i = 0; i < P.n_ctr; ++i) P.ctr[i] = 0;
      __threadfence();
    }
  }
}

struct Chain {
  ChParams P;
  unsigned grid;
  size_t smem;
};

}  // namespace dlrm

extern "C" int dlrm_b200_gemm_chain_create(void* const* plans, const int* dep, const int* dep_on_k, int n,
                                           int32_t* counters, int64_t counters_len, void** chain_out) {
  using namespace dlrm;
  if (!plans || !dep || !dep_on_k || !counters || !chain_out) return set_error("gemm_chain_create: NULL argument");
  if (n < 1 || n > CH_MAX_PROBLEMS) return set_error("gemm_chain_create: n=%d (1..%d)", n, CH_MAX_PROBLEMS);
  Chain* c = new Chain();
  memset(&c->P, 0, sizeof(c->P));
  ChParams& P = c->P;
  long long total = 0, nctr = 0;
  uint32_t stage_bytes = 0;
  for (int i = 0; i < n; ++i) {
    const TcPlan* p = static_cast<const TcPlan*>(plans[i]);
    if (!p) { delete c; return set_error("gemm_chain_create: plan %d is NULL", i); }
    if (p->bn != 32 && p->bn != 64 && p->bn != 128) { delete c; return set_error("gemm_chain_create: plan %d tile_n=%d", i, p->bn); }
    ChProblem& Q = P.p[i];
    Q.m[0] = p->tmAh; Q.m[1] = p->tmAl; Q.m[2] = p->tmBh; Q.m[3] = p->tmBl;
    Q.a = p->args;
    Q.bn = p->bn;
    Q.gx = (int)p->grid.x; Q.gy = (int)p->grid.y; Q.gz = (int)p->grid.z;
    Q.task_begin = (int)total;
    total += (long long)Q.gx * Q.gy * Q.gz;
    Q.ctr_base = (int)nctr;
    nctr += Q.gy;
    Q.dep = dep[i];
    Q.dep_on_k = dep_on_k[i] ? 1 : 0;
    Q.signal = 0;
    if (Q.dep >= i) { delete c; return set_error("gemm_chain_create: plan %d depends on plan %d (must be earlier)", i, Q.dep); }
    if (Q.dep >= 0) {
      ChProblem& Dp = P.p[Q.dep];
      if (Dp.gz != 1) { delete c; return set_error("gemm_chain_create: plan %d is split-K and cannot be a producer", Q.dep); }
      Dp.signal = 1;
      Q.dep_target = Dp.gx;
      // the rows this problem reads must be rows the producer writes
      const long long rows = Q.dep_on_k ? Q.a.K : Q.a.M;
      if (rows > Dp.a.M) { delete c; return set_error("gemm_chain_create: plan %d reads %lld rows, producer %d has %lld", i, rows, Q.dep, (long long)Dp.a.M); }
    }
    const uint32_t sb = (uint32_t)((Q.a.x3 ? 2 : 1) * (TC_BM * TC_BK * 2 + Q.bn * TC_BK * 2));
    stage_bytes = sb > stage_bytes ? sb : stage_bytes;
  }
  // groups (see ChGroup): maximal runs of single-split problems with equal m tiles, m-tile major
  {
    int gi = 0, i = 0;
    long long t0 = 0;
    const bool mmajor = get_tunable(TUNE_CHAIN_ORDER) == 2;   // measured slower (blocking claims): opt-in
    while (i < n) {
      int j = i + 1;
      // only row-linked problems (a task depends on ITS m tile of the producer) may be interleaved by m tile:
      // a problem that reduces over the batch (dep_on_k) needs ALL m tiles of its producer first
      if (mmajor && P.p[i].gz == 1 && !P.p[i].dep_on_k)
        while (j < n && P.p[j].gz == 1 && !P.p[j].dep_on_k && P.p[j].gy == P.p[i].gy) ++j;
      ChGroup& G = P.grp[gi++];
      G.task_begin = (int)t0; G.p0 = i; G.np = j - i; G.per_m = 0;
      for (int q = i; q < j; ++q) {
        G.per_m += P.p[q].gx;
        t0 += (long long)P.p[q].gx * P.p[q].gy * P.p[q].gz;
      }
      i = j;
    }
    P.ngroups = gi;
  }
  if (total <= 0 || total >= (1ll << 30)) { delete c; return set_error("gemm_chain_create: %lld tasks", total); }
  if (counters_len < 2 + nctr) { delete c; return set_error("gemm_chain_create: counters_len=%lld < %lld", (long long)counters_len, 2 + nctr); }
  P.n = n;
  P.total = (int)total;
  P.stage_bytes = stage_bytes;
  int budget_kb = get_tunable(TUNE_GEMM_SMEM_KB);
  if (budget_kb <= 0 || budget_kb > 200) budget_kb = 200;
  int stages = (int)(((size_t)budget_kb * 1024) / stage_bytes);
  if (stages > 8) stages = 8;
  if (stages < 2) stages = 2;
  P.stages = stages;
  P.ctr = counters;
  P.n_ctr = (int)(2 + nctr);
  c->smem = (size_t)stages * stage_bytes + 512 + CH_EPI_WARPS * TC_EPI_WARP_BYTES + 1024;   // ring | barriers, task ring | staging
  int dev = 0, sms = 0;
  DLRM_CUDA(cudaGetDevice(&dev));
  DLRM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  c->grid = (unsigned)(total < sms ? total : sms);
  *chain_out = c;
  return 0;
}

extern "C" int dlrm_b200_gemm_chain_info(void* chain, int* tasks, int* ctas, int* stages, int* smem_bytes) {
  using namespace dlrm;
  if (!chain) return set_error("gemm_chain_info: NULL chain");
  Chain* c = static_cast<Chain*>(chain);
  if (tasks) *tasks = c->P.total;
  if (ctas) *ctas = (int)c->grid;
  if (stages) *stages = c->P.stages;
  if (smem_bytes) *smem_bytes = (int)c->smem;
  return 0;
}

extern "C" int dlrm_b200_gemm_chain_set_trace(void* chain, uint64_t* trace) {
  using namespace dlrm;
  if (!chain) return set_error("gemm_chain_set_trace: NULL chain");
  static_cast<Chain*>(chain)->P.trace = reinterpret_cast<unsigned long long*>(trace);
  return 0;
}

extern "C" int dlrm_b200_gemm_chain_run(void* chain, void* stream) {
  using namespace dlrm;
  if (!chain) return set_error("gemm_chain_run: NULL chain");
  Chain* c = static_cast<Chain*>(chain);
  static bool configured = false;
  if (!configured) {
    DLRM_CUDA(cudaFuncSetAttribute(gemm_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  (void)launch_chain(gemm_chain_kernel, dim3(c->grid), dim3(CH_THREADS), c->smem, static_cast<cudaStream_t>(stream), c->P);
  DLRM_CHECK_LAUNCH("gemm_chain_kernel");
  return 0;
}

extern "C" int dlrm_b200_gemm_chain_destroy(void* chain) {
  delete static_cast<dlrm::Chain*>(chain);
  return 0;
}
