// tcgen05 / TMEM / TMA GEMM for the MLP layers (forward, dgrad, wgrad) on sm_100a.
//
//   D[M,N] = sum_k A(m,k) * B(n,k)        bf16 operands, fp32 accumulation in tensor memory.
//
// Precision modes
//   BF16   : one MMA per k-step (operands rounded to bf16).
//   BF16X3 : fp32-grade result from bf16 tensor cores.  Every fp32 operand x is stored as the pair
//            hi = bf16(x), lo = bf16(x - hi); the kernel issues hi*hi + hi*lo + lo*hi per k-step
//            (the dropped lo*lo and residual terms are <= 2^-16 relative per product), which keeps
//            the logits within 1e-5 of the reference's fp32 CPU forward (BASELINE.json north_star;
//            measured in tests/test_gpu_gemm_tc.py).  4 operand tiles per stage instead of 2.
//
// Structure (one 128 x BN output tile per CTA, optional split-K over gridDim.z):
//   warp 0     : TMA producer  -- cp.async.bulk.tensor.2d into a 128B-swizzled smem ring, mbarrier
//                complete_tx signalling.
//   warp 1     : allocates TMEM, one elected lane issues tcgen05.mma (cta_group::1, kind::f16,
//                UMMA 128 x BN x 16), tcgen05.commit releases smem stages / publishes the accumulator.
//   warps 2..5 : epilogue -- tcgen05.ld (32x32b.x32) of their TMEM lane quadrant, fused
//                activation / activation-gradient mask, then fp32 and/or (hi,lo) bf16 stores in
//                normal and transposed layout (the operand layouts of the next GEMMs).
// Operands may be K-major ([rows, K] with K contiguous) or MN-major ([K, rows] with rows
// contiguous, e.g. dY^T read straight from dY): the UMMA descriptors and instruction descriptor
// carry the majorness, so no transposed copy is needed for MN-major inputs.
#include <cuda.h>
#include <string.h>
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
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  // try_wait suspends for a bounded time per call; a protocol bug must trap, not hang the GPU
  uint32_t ok = 0;
  int tries = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P1;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) break;
    if (++tries > (1 << 24)) __trap();
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

// ------------------------------------------------------------------------------------------ kernel
// smem per stage: A_hi [A_lo] B_hi [B_lo]; every tile 1024-byte aligned.
template <int BN>
__global__ void __launch_bounds__(192, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
               const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl,
               const TcArgs g, int stages) {
#define TCB_BX blockIdx.x
#define TCB_BY blockIdx.y
#define TCB_BZ blockIdx.z
#define TCB_MAP_AH (&tmAh)
#define TCB_MAP_AL (&tmAl)
#define TCB_MAP_BH (&tmBh)
#define TCB_MAP_BL (&tmBl)
#include "gemm_tc_body.cuh"
#undef TCB_BX
#undef TCB_BY
#undef TCB_BZ
#undef TCB_MAP_AH
#undef TCB_MAP_AL
#undef TCB_MAP_BH
#undef TCB_MAP_BL
}

// Several independent problems in ONE launch (same tile width): CTAs [cta_begin, next cta_begin) of the
// linear grid belong to entry p and are numbered n tile fastest, then m tile, then k split -- the same
// CTA program as above, so results are bit-identical to separate launches.  Meant for the weight-gradient
// GEMMs of an MLP, which are mutually independent and individually too small to fill the GPU.
constexpr int TC_MAX_GROUP = 4;

struct TcGroupEntry {
  CUtensorMap m[4];   // A_hi, A_lo, B_hi, B_lo
  TcArgs a;
  int stages;
  int gx, gy, gz;
  int cta_begin;
};

struct TcGroup {
  TcGroupEntry e[TC_MAX_GROUP];
  int n;
};

template <int BN>
__global__ void __launch_bounds__(192, 1) gemm_tc_group_kernel(const __grid_constant__ TcGroup P) {
  int p = 0;
  while (p + 1 < P.n && (int)blockIdx.x >= P.e[p + 1].cta_begin) ++p;
  const TcGroupEntry& E = P.e[p];
  const TcArgs& g = E.a;
  const int stages = E.stages;
  const int local = (int)blockIdx.x - E.cta_begin;
  const int tcb_bx = local % E.gx;
  const int tcb_by = (local / E.gx) % E.gy;
  const int tcb_bz = local / (E.gx * E.gy);
#define TCB_BX tcb_bx
#define TCB_BY tcb_by
#define TCB_BZ tcb_bz
#define TCB_MAP_AH (&E.m[0])
#define TCB_MAP_AL (&E.m[1])
#define TCB_MAP_BH (&E.m[2])
#define TCB_MAP_BL (&E.m[3])
#include "gemm_tc_body.cuh"
#undef TCB_BX
#undef TCB_BY
#undef TCB_BZ
#undef TCB_MAP_AH
#undef TCB_MAP_AL
#undef TCB_MAP_BH
#undef TCB_MAP_BL
}

// ------------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D bf16 tensor [outer, inner] with `ld` elements between outer rows; box {box_inner, box_outer}
static int make_map(CUtensorMap* map, const void* ptr, long long inner, long long outer, long long ld,
                    int box_inner, int box_outer) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return set_error("gemm_tc: cuTensorMapEncodeTiled not available");
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (ld * 2) % 16)
    return set_error("gemm_tc: operand pointer/ld not 16-byte aligned (ld=%lld)", ld);
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error("gemm_tc: cuTensorMapEncodeTiled failed (%d)", (int)r);
  return 0;
}

struct TcPlan {
  CUtensorMap tmAh, tmAl, tmBh, tmBl;
  TcArgs args;
  int bn, stages, splits;
  size_t smem;
  dim3 grid;
};

template <int BN>
static int launch_tc(const TcPlan& p, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    DLRM_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  (void)launch_chain(gemm_tc_kernel<BN>, p.grid, dim3(192), p.smem, st, p.tmAh, p.tmAl, p.tmBh, p.tmBl, p.args, p.stages);
  DLRM_CHECK_LAUNCH("gemm_tc_kernel");
  return 0;
}

}  // namespace dlrm

extern "C" int dlrm_b200_gemm_tc_plan_create(const dlrm_gemm_tc_desc_t* d, void** plan_out) {
  using namespace dlrm;
  if (!d || !plan_out) return set_error("gemm_tc_plan_create: NULL argument");
  if (d->M <= 0 || d->N <= 0 || d->K <= 0) return set_error("gemm_tc: empty problem M=%lld N=%lld K=%lld", (long long)d->M, (long long)d->N, (long long)d->K);
  if (!d->A_hi || !d->B_hi || (d->mode_x3 && (!d->A_lo || !d->B_lo)))
    return set_error("gemm_tc: NULL operand");
  TcPlan* p = new TcPlan();
  TcArgs& a = p->args;
  a.M = d->M; a.N = d->N; a.K = d->K;
  a.x3 = d->mode_x3 ? 1 : 0;
  a.a_mn = d->a_mn_major ? 1 : 0;
  a.b_mn = d->b_mn_major ? 1 : 0;
  a.act = d->act; a.mask_act = d->mask_act;
  a.mask_hi = static_cast<const __nv_bfloat16*>(d->mask_hi);
  a.mask_lo = static_cast<const __nv_bfloat16*>(d->mask_lo);
  a.ldmask = d->ldmask;
  a.out_f32 = d->out_f32; a.ld_f32 = d->ld_f32; a.slab_stride = d->slab_stride;
  a.out_hi = static_cast<__nv_bfloat16*>(d->out_hi); a.out_lo = static_cast<__nv_bfloat16*>(d->out_lo);
  a.ld_out = d->ld_out;
  a.outT_hi = static_cast<__nv_bfloat16*>(d->outT_hi); a.outT_lo = static_cast<__nv_bfloat16*>(d->outT_lo);
  a.ld_outT = d->ld_outT;
  a.out_col = d->out_col; a.col_index = d->col_index; a.col_slab_stride = d->col_slab_stride;
  if (a.mask_act != DLRM_ACT_NONE && !a.mask_hi) { delete p; return set_error("gemm_tc: mask_act without mask_hi"); }
  if (a.out_hi && (a.ld_out % 8)) { delete p; return set_error("gemm_tc: ld_out must be a multiple of 8"); }
  // tile width: keep >= ~64 CTAs when N is small
  const long long mt = (d->M + TC_BM - 1) / TC_BM;
  int bn = 128;
  if (d->tile_n == 32 || d->tile_n == 64 || d->tile_n == 128) bn = d->tile_n;
  else {
    while (bn > 32 && mt * ((d->N + bn - 1) / bn) < 96) bn >>= 1;
    if (d->N <= 32) bn = 32; else if (d->N <= 64 && bn > 64) bn = 64;
  }
  if (a.b_mn && bn < 64) bn = 64;  // MN-major boxes are 64 wide
  // operand-ring budget: 200 KB = deepest pipeline, one CTA per SM; ~100 KB lets two CTAs (of this or of
  // a concurrent GEMM on another stream) share an SM, which hides the latency of these small problems
  int budget_kb = get_tunable(TUNE_GEMM_SMEM_KB);
  if (budget_kb <= 0 || budget_kb > 200) budget_kb = 200;
  if (budget_kb < 48) budget_kb = 48;
  if (!(d->tile_n == 128) && bn == 128 &&
      (size_t)budget_kb * 1024 < 2 * (size_t)(a.x3 ? 2 : 1) * (TC_BM * TC_BK * 2 + 128 * TC_BK * 2))
    bn = 64;                       // two stages of a 128-wide tile would not fit the budget
  p->bn = bn;
  a.num_kb = (int)((d->K + TC_BK - 1) / TC_BK);
  int splits = d->split_k > 1 ? d->split_k : 1;
  if (splits > a.num_kb) splits = a.num_kb;
  a.kb_per_split = (a.num_kb + splits - 1) / splits;
  splits = (a.num_kb + a.kb_per_split - 1) / a.kb_per_split;  // no empty split
  p->splits = splits;
  if (splits > 1 && (a.out_hi || a.outT_hi || a.act != DLRM_ACT_NONE || a.mask_act != DLRM_ACT_NONE)) {
    delete p; return set_error("gemm_tc: split-K only supports fp32 slab outputs");
  }
  const size_t stage_bytes = (size_t)(a.x3 ? 2 : 1) * (TC_BM * TC_BK * 2 + bn * TC_BK * 2);
  int stages = (int)(((size_t)budget_kb * 1024) / stage_bytes);
  if (stages > 8) stages = 8;
  if (stages < 2) stages = 2;
  if (stages > a.kb_per_split) stages = a.kb_per_split < 2 ? 2 : a.kb_per_split;
  p->stages = stages;
  p->smem = stages * stage_bytes + (2 * stages + 1) * 8 + 16 + 1024;
  p->grid = dim3((unsigned)((d->N + bn - 1) / bn), (unsigned)mt, (unsigned)splits);
  int rc = 0;
  // operand maps.  K-major: tensor [rows, K]; MN-major: tensor [K, rows].
  if (!a.a_mn) {
    rc |= make_map(&p->tmAh, d->A_hi, d->K, d->M, d->lda, TC_BK, TC_BM);
    rc |= make_map(&p->tmAl, a.x3 ? d->A_lo : d->A_hi, d->K, d->M, d->lda, TC_BK, TC_BM);
  } else {
    rc |= make_map(&p->tmAh, d->A_hi, d->M, d->K, d->lda, 64, TC_BK);
    rc |= make_map(&p->tmAl, a.x3 ? d->A_lo : d->A_hi, d->M, d->K, d->lda, 64, TC_BK);
  }
  if (!a.b_mn) {
    rc |= make_map(&p->tmBh, d->B_hi, d->K, d->N, d->ldb, TC_BK, bn);
    rc |= make_map(&p->tmBl, a.x3 ? d->B_lo : d->B_hi, d->K, d->N, d->ldb, TC_BK, bn);
  } else {
    rc |= make_map(&p->tmBh, d->B_hi, d->N, d->K, d->ldb, 64, TC_BK);
    rc |= make_map(&p->tmBl, a.x3 ? d->B_lo : d->B_hi, d->N, d->K, d->ldb, 64, TC_BK);
  }
  if (rc) { delete p; return -1; }
  *plan_out = p;
  return 0;
}

extern "C" int dlrm_b200_gemm_tc_plan_info(void* plan, int* tile_n, int* stages, int* splits, int* ctas) {
  using namespace dlrm;
  if (!plan) return set_error("gemm_tc_plan_info: NULL plan");
  TcPlan* p = static_cast<TcPlan*>(plan);
  if (tile_n) *tile_n = p->bn;
  if (stages) *stages = p->stages;
  if (splits) *splits = p->splits;
  if (ctas) *ctas = (int)(p->grid.x * p->grid.y * p->grid.z);
  return 0;
}

extern "C" int dlrm_b200_gemm_tc_run(void* plan, void* stream) {
  using namespace dlrm;
  if (!plan) return set_error("gemm_tc_run: NULL plan");
  TcPlan* p = static_cast<TcPlan*>(plan);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (p->bn == 128) return launch_tc<128>(*p, st);
  if (p->bn == 64) return launch_tc<64>(*p, st);
  return launch_tc<32>(*p, st);
}

namespace dlrm {
template <int BN>
static int launch_tc_group(const TcGroup& G, unsigned total, size_t smem, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    DLRM_CUDA(cudaFuncSetAttribute(gemm_tc_group_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  (void)launch_chain(gemm_tc_group_kernel<BN>, dim3(total), dim3(192), smem, st, G);
  DLRM_CHECK_LAUNCH("gemm_tc_group_kernel");
  return 0;
}
}  // namespace dlrm

extern "C" int dlrm_b200_gemm_tc_run_group(void* const* plans, int num_plans, void* stream) {
  using namespace dlrm;
  if (!plans || num_plans < 1 || num_plans > TC_MAX_GROUP)
    return set_error("gemm_tc_run_group: num_plans=%d (1..%d)", num_plans, TC_MAX_GROUP);
  TcGroup G;
  memset(&G, 0, sizeof(G));
  G.n = num_plans;
  long long total = 0;
  size_t smem = 0;
  int bn = 0;
  for (int i = 0; i < num_plans; ++i) {
    const TcPlan* p = static_cast<const TcPlan*>(plans[i]);
    if (!p) return set_error("gemm_tc_run_group: plan %d is NULL", i);
    if (i == 0) bn = p->bn;
    if (p->bn != bn) return set_error("gemm_tc_run_group: plans must share the tile width (%d vs %d)", p->bn, bn);
    TcGroupEntry& e = G.e[i];
    e.m[0] = p->tmAh; e.m[1] = p->tmAl; e.m[2] = p->tmBh; e.m[3] = p->tmBl;
    e.a = p->args;
    e.stages = p->stages;
    e.gx = (int)p->grid.x; e.gy = (int)p->grid.y; e.gz = (int)p->grid.z;
    e.cta_begin = (int)total;
    total += (long long)p->grid.x * p->grid.y * p->grid.z;
    smem = p->smem > smem ? p->smem : smem;
  }
  if (total <= 0 || total >= (1ll << 31)) return set_error("gemm_tc_run_group: %lld CTAs", total);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (bn == 32) return launch_tc_group<32>(G, (unsigned)total, smem, st);
  if (bn == 64) return launch_tc_group<64>(G, (unsigned)total, smem, st);
  return launch_tc_group<128>(G, (unsigned)total, smem, st);
}

extern "C" int dlrm_b200_gemm_tc_plan_destroy(void* plan) {
  delete static_cast<dlrm::TcPlan*>(plan);
  return 0;
}
