/*
 * dlrm_b200.h -- C ABI of libdlrm_b200.so: the B200 (sm_100a) kernels behind the
 * DLRM_Net forward/backward hot path of facebookresearch/dlrm.
 *
 * The reference has NO native interface for this path: every op below replaces an
 * ATen call made from dlrm_s_pytorch.py (cited per entry point).  A maintainer of
 * the reference binds these with ctypes; dlrm_b200/_lib.py is
 * exactly that binding.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no torch types.
 *   - every pointer is a DEVICE pointer on the current CUDA device unless marked
 *     [host]; the caller (PyTorch) owns all memory, the library allocates nothing
 *     and keeps no pointer beyond the call (except inside explicit handles).
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).
 *     All work is enqueued asynchronously on it; no hidden synchronisation, so
 *     every entry point is legal inside CUDA-graph stream capture.
 *   - return 0 on success, <0 on error; dlrm_b200_last_error() gives the text
 *     (thread-local).  Unsupported shapes are errors, never silent fallbacks.
 *   - float = IEEE fp32.  Index/offset tensors are int64 (idx_bytes=8) or int32
 *     (idx_bytes=4), both arrays of one call having the same width, exactly as
 *     nn.EmbeddingBag accepts them; they are consumed bit-for-bit, never copied.
 */
#ifndef DLRM_B200_H_
#define DLRM_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DLRM_B200_ABI_VERSION 1
#define DLRM_B200_MAX_TABLES_PER_CALL 64 /* larger T: split into several calls */
#define DLRM_B200_MAX_PEERS 8            /* GPUs of one NVSwitch box */

/* activations of create_mlp (dlrm_s_pytorch.py:237-241) */
enum { DLRM_ACT_NONE = 0, DLRM_ACT_RELU = 1, DLRM_ACT_SIGMOID = 2 };
/* loss functions (dlrm_s_pytorch.py:385-393) */
enum { DLRM_LOSS_MSE = 0, DLRM_LOSS_BCE = 1, DLRM_LOSS_WBCE = 2 };
/* sparse optimizers: torch.optim.SGD (dlrm_s_pytorch.py:1343) / optim/rwsadagrad.py */
enum { DLRM_OPT_SGD = 0, DLRM_OPT_RWSADAGRAD = 1 };
/* GEMM back ends */
enum { DLRM_GEMM_SIMT_FP32 = 0, DLRM_GEMM_TC_BF16X3 = 1, DLRM_GEMM_TC_BF16 = 2 };

int dlrm_b200_abi_version(void);
const char* dlrm_b200_last_error(void);
/* Reads and clears the device error word of the current device (synchronises `stream`): bit 0 = an embedding
 * index outside its table since the last check.  The only hidden state of the library: 256 bytes of device
 * memory per GPU, allocated on first use. */
int dlrm_b200_check_device_errors(void* stream);
/* sm count / compute capability of `device`; error unless cc >= 10.0 */
int dlrm_b200_device_info(int device, int* sm_count, int* cc_major, int* cc_minor);
/* Experiment knobs of the kernels (ids: csrc/common.cuh `enum Tunable`; dlrm_b200/_lib.py TUNE; env DLRM_TUNE).
 * Process-wide, read at launch time; 0 restores the default of every knob.  Not part of the reference's surface. */
int dlrm_b200_set_tunable(int id, int value);

/* ------------------------------------------------------------------------------------------
 * apply_emb  (dlrm_s_pytorch.py:407-462: one nn.EmbeddingBag(mode="sum") call per table)
 * ONE launch for all tables:  out[b, k, :] = sum_{j in bag(k,b)} rw_k[idx_k[j]] * W_k[idx_k[j], :]
 * accumulated sequentially in index order (bit-identical to the reference CPU kernel when
 * row_weights == NULL).  Bag b of table k is idx[off[b] .. off[b+1]) and the last bag runs to
 * nnz (EmbeddingBag without include_last_offset) unless include_last != 0, in which case
 * offsets has batch+1 entries and `nnz` is ignored (graph-replay friendly).  Empty bag -> 0.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  const float* weight;      /* [rows, dim] row-major, 16-byte aligned when dim % 4 == 0 */
  const void* indices;      /* [nnz]   int64 / int32 */
  const void* offsets;      /* [batch] (or [batch+1] when include_last) same type */
  const float* row_weights; /* NULL, or [rows]: v_W_l[k] (weighted pooling, :425-428) */
  int64_t nnz;
  int64_t rows;             /* rows of the WHOLE table: an index outside [0, rows) sets the device error word
                             * (dlrm_b200_check_device_errors) and is read as row 0 / skipped; 0 = unchecked */
  int64_t ld;               /* row stride of `weight` in floats; 0 = dim (dense rows) */
  /* Sharded placement (dlrm_b200/placement.py).  All zero = the call-level layout out[b, k, :].
   * out_stride > 0: the pooled row of bag b goes to out (or the owner's peer buffer) + b_local*out_stride + out_off
   *   -- a whole table lands in feature slot 1+t of the interaction operand, a row-split shard in its slab of
   *   the partial-sum area (dlrm_b200_emb_reduce_partials adds the slabs).
   * row_n > 0: `weight` holds rows [row_lo, row_lo + row_n) of the table; indices outside the range belong
   *   to another shard and are skipped (a partial sum over this shard's rows). */
  int64_t out_off, out_stride;
  int64_t row_lo, row_n;
} dlrm_emb_fwd_table_t;

int dlrm_b200_emb_bag_fwd(const dlrm_emb_fwd_table_t* tables /*[host]*/, int num_tables, int dim,
                          int64_t batch, int idx_bytes, int include_last,
                          float* out, int64_t out_stride_sample, int64_t out_stride_table,
                          void* stream);

/* ------------------------------------------------------------------------------------------
 * Embedding backward fused with the sparse optimizer
 * (autograd _embedding_bag_backward, dlrm_s_pytorch.py:1613 + optimizer.step() :1620:
 *  optim/rwsadagrad.py:117-143 or torch.optim.SGD sparse add).
 *
 * Step 1, dlrm_b200_emb_bwd_link: depends on the indices only (can run on a side stream
 * during the forward pass).  Threads every (table,row) occurrence of the batch onto a
 * per-row list: prev = atomicExch(&head[row], pos+1); next[pos] = prev.  `head` is an int32
 * array over all rows of the table, zero on entry and zero again after step 2.
 * Step 2, dlrm_b200_emb_bwd_update: one warp per bag; the list head (= unique owner of a row)
 * sums dY over the row's occurrences in ascending position (== grad.coalesce()), then
 *   RWSAdagrad: momentum[row] += mean_d(g^2); W[row] -= lr * g / (sqrt(momentum[row]) + eps)
 *   SGD:        W[row] -= lr * g
 * `lr` is the already-decayed clr of optim/rwsadagrad.py:115.
 * dY[b, k, :] is read at dY + b*dy_stride_sample + k*dy_stride_table.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  float* weight;        /* [rows, dim] updated in place */
  float* momentum;      /* [rows] (RWSAdagrad) or NULL (SGD) */
  int32_t* head;        /* [rows] zero-initialised scratch, self-cleaning */
  const void* indices;  /* as in forward */
  const void* offsets;
  int64_t nnz;
  int64_t rows;
  int64_t pair_base;    /* first slot of this table in next[] / (sum of nnz of earlier tables) */
  int64_t ld;           /* row stride of `weight` in floats; 0 = dim */
  int64_t mom_stride;   /* elements between consecutive rows' accumulators in `momentum`; 0 = 1.
                         * ld = dim + 4 with momentum = weight + dim and mom_stride = ld keeps the
                         * row-wise Adagrad accumulator in the SAME DRAM burst as its row: the update then
                         * costs one activation per row instead of two (measured: profiles/). */
  /* Sharded placement.  use_dy_off != 0: the gradient row of (bag b, this table) is at dY(b) + dy_off instead
   * of dY(b) + k*dy_stride_table.  row_n > 0: `weight`/`momentum`/`head` hold rows [row_lo, row_lo + row_n)
   * of the table; occurrences of other rows are another shard's.  head == NULL: this table is not linked and
   * not updated by dlrm_b200_emb_bwd_update (tiny tables: dlrm_b200_emb_bwd_small_update). */
  int64_t use_dy_off, dy_off;
  int64_t row_lo, row_n;
  /* elements between consecutive rows' list heads; 0 = 1.  head = (int32*)weight + dim + 1 with head_stride = ld
   * (and momentum = weight + dim, mom_stride = ld, ld = dim + 4) keeps BOTH per-row words inside the row's own
   * DRAM page: the update is bound by the RATE of random DRAM accesses (~10 G/s measured: the gather's 512-byte
   * rows and the update's 4-byte words cost the same), so 6 accesses per occurrence become 2-3. */
  int64_t head_stride;
} dlrm_emb_bwd_table_t;

/* Optional duplicate filter (dlrm_emb_dedup_t): at 1e6-row tables almost every row of a batch is
 * touched once, and the 4-byte random accesses to head[] cost as much HBM time as the 512-byte rows.
 * With a filter the training gather only bumps a counter in an L2-sized hashed array
 * (fire-and-forget RED), dlrm_b200_emb_bwd_classify() -- right after the gather, while the
 * counters are L2-hot -- marks the occurrences whose counter is > 1 as suspects (hash collisions
 * only add false suspects) and threads ONLY those onto the per-row lists; the update kernel then
 * treats unflagged occurrences as sole owners of their row without touching head[] or link[].x.
 *   filter   : uint32 [2^log2_size + 1], zeroed by the caller before every training gather
 *              (the last element is the suspect counter)
 *   flags    : uint8  [nnz capacity]   suspects : int32 [nnz capacity]
 * dedup == NULL everywhere: every occurrence is linked (the original scheme). */
typedef struct {
  uint32_t* filter;
  int32_t log2_size;
  uint8_t* flags;
  int32_t* suspects;
} dlrm_emb_dedup_t;

/* Training forward: the gather of dlrm_b200_emb_bag_fwd AND step 1 (link) in the same launch --
 * the index of every occurrence is already in a register, so linking costs one atomicExch. */
int dlrm_b200_emb_bag_fwd_train(const dlrm_emb_fwd_table_t* tables /*[host]*/,
                                const dlrm_emb_bwd_table_t* train /*[host]*/, int num_tables, int dim,
                                int64_t batch, int idx_bytes, int include_last, int32_t* next,
                                float* out, int64_t out_stride_sample, int64_t out_stride_table,
                                const dlrm_emb_dedup_t* dedup /*[host] or NULL*/, void* stream);

/* After a filtered training gather: flag suspects and link them (two small launches). */
int dlrm_b200_emb_bwd_classify(const dlrm_emb_bwd_table_t* tables /*[host]*/, int num_tables,
                               int64_t batch, int idx_bytes, int include_last, int32_t* next,
                               const dlrm_emb_dedup_t* dedup /*[host]*/, void* stream);

int dlrm_b200_emb_bwd_link(const dlrm_emb_bwd_table_t* tables /*[host]*/, int num_tables,
                           int64_t batch, int idx_bytes, int include_last,
                           int32_t* next /*[total nnz]*/, void* stream);

int dlrm_b200_emb_bwd_update(const dlrm_emb_bwd_table_t* tables /*[host]*/, int num_tables, int dim,
                             int64_t batch, int idx_bytes, int include_last,
                             const int32_t* next, const float* dY, int64_t dy_stride_sample,
                             int64_t dy_stride_table, int optimizer, float lr, float eps,
                             const dlrm_emb_dedup_t* dedup /*[host] or NULL*/, void* stream);

/* ------------------------------------------------------------------------------------------
 * Table-wise sharded runs (replaces extend_distributed.alltoall, extend_distributed.py:389-486, and
 * the butterfly shuffle of parallel_forward, dlrm_s_pytorch.py:693-699): the exchange is fused into
 * the kernels through peer-mapped memory (cudaIpc / NVLink), no staging buffer, no collective.
 *   fwd : this rank pools ITS tables for the GLOBAL batch; bag b is stored into rank (b / batch_local)'s
 *         buffer: peer_out[d] + (b % batch_local) * out_stride_sample + k * out_stride_table.
 *   bwd : the dY row of global bag b is loaded from peer_dY[b / batch_local] with the same strides.
 * The caller synchronises the ranks (a barrier after fwd, before bwd).  train may be NULL (inference).
 * ------------------------------------------------------------------------------------------ */
int dlrm_b200_emb_bag_fwd_p2p(const dlrm_emb_fwd_table_t* tables /*[host]*/,
                              const dlrm_emb_bwd_table_t* train /*[host] or NULL*/, int num_tables, int dim,
                              int64_t batch_global, int idx_bytes, int include_last, int32_t* next,
                              float* const* peer_out /*[host][world]*/, int world, int64_t batch_local,
                              int64_t out_stride_sample, int64_t out_stride_table,
                              const dlrm_emb_dedup_t* dedup /*[host] or NULL*/, void* stream);
int dlrm_b200_emb_bwd_update_p2p(const dlrm_emb_bwd_table_t* tables /*[host]*/, int num_tables, int dim,
                                 int64_t batch_global, int idx_bytes, int include_last, const int32_t* next,
                                 const float* const* peer_dY /*[host][world]*/, int world,
                                 int64_t batch_local, int64_t dy_stride_sample, int64_t dy_stride_table,
                                 int optimizer, float lr, float eps,
                                 const dlrm_emb_dedup_t* dedup /*[host] or NULL*/, void* stream);

/* NCCL-free cross-GPU steps over the same peer mappings (graph-capturable):
 *   barrier        : peer_sig[r] = int32[world] on rank r (zero-initialised once); epoch = device int32.
 *   allreduce_mean : peer_grad[r] = dense-gradient arena of rank r (n floats); every rank ends with
 *                    the mean over ranks (DDP semantics), summed in rank order.  The caller places a
 *                    barrier before and after. */
/* kernels running on `device` may dereference memory of `peer_device` (cudaDeviceEnablePeerAccess);
 * needed once per peer before passing IPC-mapped peer pointers to the entry points above. */
int dlrm_b200_enable_peer_access(int device, int peer_device);
/* Export the cudaMalloc allocation that holds device pointer `ptr` of this process: a 64-byte
 * cudaIpcMemHandle_t and ptr's byte offset inside the allocation (send both to the peer process). */
int dlrm_b200_ipc_export(const void* ptr, void* handle64_out /*[host, 64 bytes]*/, int64_t* offset_out /*[host]*/);
/* Open such a handle in another process with `device` current, so that kernels launched on `device`
 * can dereference the mapping (base_out + offset = the peer's ptr).  One open per allocation. */
int dlrm_b200_ipc_open(const void* handle64 /*[host]*/, int device, void** base_out /*[host]*/);
int dlrm_b200_ipc_close(void* base);
int dlrm_b200_p2p_barrier(void* const* peer_sig /*[host][world]*/, int rank, int world, int32_t* epoch,
                          void* stream);
int dlrm_b200_p2p_allreduce_mean(void* const* peer_grad /*[host][world]*/, int rank, int world, int64_t n,
                                 void* stream);

/* ------------------------------------------------------------------------------------------
 * apply_mlp layer (dlrm_s_pytorch.py:399-405: nn.Linear -> addmm, + ReLU / Sigmoid modules)
 *   fwd  : Y[M,N]  = act(X[M,K] W[N,K]^T + bias[N])
 *   dgrad: dX[M,K] = (dY[M,N] W[N,K]) * act'(Xact[M,K])   (Xact = output of the previous
 *          layer, NULL / DLRM_ACT_NONE when the input is not an activation)
 *   wgrad: dW[N,K] = dY[M,N]^T X[M,K];  dbias[N] = sum_m dY[m,n]
 * ld* are row strides in elements.  backend = DLRM_GEMM_*.
 * ------------------------------------------------------------------------------------------ */
int dlrm_b200_linear_fwd(const float* X, int64_t ldx, const float* W, int64_t ldw, const float* bias,
                         float* Y, int64_t ldy, int64_t M, int64_t N, int64_t K, int act,
                         int backend, void* stream);
int dlrm_b200_linear_dgrad(const float* dY, int64_t lddy, const float* W, int64_t ldw,
                           const float* Xact, int64_t ldxa, int act_prev,
                           float* dX, int64_t lddx, int64_t M, int64_t N, int64_t K,
                           int backend, void* stream);
int dlrm_b200_linear_wgrad(const float* dY, int64_t lddy, const float* X, int64_t ldx,
                           float* dW, int64_t lddw, float* dbias, int64_t M, int64_t N, int64_t K,
                           int backend, void* stream);

/* ------------------------------------------------------------------------------------------
 * interact_features (dlrm_s_pytorch.py:483-515), op == "dot":
 *   T[b] = [x[b]; ly_0[b]; ...] (F x D, read in place at T + b*ldt, feature stride D)
 *   R[b, 0:D] = x[b];  R[b, D + tri(i,j)] = <T[b,i], T[b,j]>  for j < i (+ diagonal if itself),
 *   row-major strict lower triangle (1,0),(2,0),(2,1),...  (cat/bmm/index/cat fused: K3-K6).
 * bwd: dT[b] = (dZ + dZ^T) T[b] (+ dR[b,0:D] on feature 0), with dZ scattered from dR[b, D:];
 *   feature 0 of dT is multiplied by act'(x), act = mask_feature0 (DLRM_ACT_*: the bottom MLP's
 *   last activation; DLRM_ACT_NONE = no mask).
 * op == "cat" is a pure layout (R == T flattened): no kernel, handled by strides on the host.
 * ------------------------------------------------------------------------------------------ */
int dlrm_b200_interact_fwd(const float* T, int64_t ldt, float* R, int64_t ldr, int64_t batch,
                           int num_features, int dim, int itself, void* stream);
/* _ex variants: additionally emit the (hi, lo) bf16 operand pair consumed by the tcgen05 GEMMs
 * (R for the first top-MLP layer; feature 0 of dT for the bottom MLP's backward).  R may be NULL
 * when only the bf16 pair is wanted. */
int dlrm_b200_interact_fwd_ex(const float* T, int64_t ldt, float* R, int64_t ldr, void* R_hi, void* R_lo,
                              int64_t ld_rb, int64_t batch, int num_features, int dim, int itself,
                              void* stream);
int dlrm_b200_interact_bwd_ex(const float* T, int64_t ldt, const float* dR, int64_t lddr, float* dT,
                              int64_t lddt, int64_t batch, int num_features, int dim, int itself,
                              int mask_feature0, void* g0_hi, void* g0_lo, int64_t ld_g0, void* stream);
/* Sharded variant: feature i's gradient rows are stored at feat_dst[q] + sample * feat_ld[q] (floats) for
 * every destination q in [feat_first[i], feat_first[i+1]) instead of dT (at most 128 destinations in all).
 * On a sharded run the destinations of feature 1 + t point into the receive buffers of the rank(s) storing
 * table t (peer-mapped over NVLink; a row-split table has one on every rank), which replaces the backward
 * all-to-all of dlrm_s_pytorch.py:545-560 / extend_distributed.py:alltoall backward; feature 0 stays local.
 * emb_grad_scale multiplies the rows of features >= 1: 1 = the reference's distributed semantics (every rank's
 * loss is the mean over ITS batch slice and the embedding gradients of the ranks are SUMMED, i.e. world x the
 * single-process gradient); 1/world = the gradient of the global mean loss (equals a single-process run). */
int dlrm_b200_interact_bwd_p2p(const float* T, int64_t ldt, const float* dR, int64_t lddr,
                               void* const* feat_dst /*[host][ndst]*/, const int64_t* feat_ld /*[host][ndst]*/,
                               const int* feat_first /*[host][F+1]*/, float emb_grad_scale, int64_t batch,
                               int num_features, int dim,
                               int itself, int mask_feature0, void* g0_hi, void* g0_lo, int64_t ld_g0,
                               void* stream);
int dlrm_b200_interact_bwd(const float* T, int64_t ldt, const float* dR, int64_t lddr,
                           float* dT, int64_t lddt, int64_t batch, int num_features, int dim,
                           int itself, int mask_feature0, void* stream);

/* ------------------------------------------------------------------------------------------
 * loss_fn_wrap (dlrm_s_pytorch.py:148-156; MSELoss/BCELoss(mean), wbce) + clamp (:607-610)
 * + backward through the loss, the clamp and the last activation:
 *   z = clamp(p, thr, 1-thr) iff 0 < thr < 1;  loss = mean(l(z,t) [* ws[t]])
 *   gz[i] = dloss/dz[i] * [thr <= p <= 1-thr] * act'(p[i])     (act = last layer's activation)
 * n = batch * outputs (contiguous).  loss_out: 1 float; gz may be NULL (inference).
 * `scratch` >= 1024 floats.  Deterministic (fixed reduction tree).
 * ------------------------------------------------------------------------------------------ */
int dlrm_b200_loss_fwd_bwd(const float* p, const float* target, const float* loss_ws /*[2] or NULL*/,
                           int64_t n, int loss_kind, float loss_threshold, int last_act,
                           float* loss_out, float* gz, float* scratch, void* stream);

/* gz = gy * [thr <= y <= 1-thr] * act'(y): entry of the backward pass when the loss is computed
 * OUTSIDE the library (autograd of DLRM_Net.forward: `E.backward()` hands us dE/d(clamped p)).
 * clamp_threshold outside (0,1) disables the clamp mask. */
int dlrm_b200_act_bwd(const float* gy, const float* y, float* gz, int64_t n, int act,
                      float clamp_threshold, void* stream);

/* ------------------------------------------------------------------------------------------
 * Dense parameters of optimizer.step(): flat arenas (all bot/top W and b, contiguous).
 *   SGD:        p -= lr * g
 *   RWSAdagrad dense branch (optim/rwsadagrad.py:145-148): s += g*g; p -= lr * g / (sqrt(s)+eps)
 * ------------------------------------------------------------------------------------------ */
/* ------------------------------------------------------------------------------------------
 * Fused head for a top MLP whose last layer has one output: Linear(K->1) + act + clamp + loss
 * + d(loss) + wgrad/bias-grad/dgrad of that layer in ONE launch (see csrc/head.cu).
 * target == NULL: inference (p only).  dW == NULL: loss (+ gz) only.  gprev (fp32) and/or
 * gprev_hi/lo (bf16 pair) receive gz * w * act_prev'(h).  scratch: dlrm_b200_head_scratch_bytes()
 * bytes, zero-initialised once.
 * ------------------------------------------------------------------------------------------ */
int64_t dlrm_b200_head_scratch_bytes(int64_t batch, int64_t K);
int dlrm_b200_head_fused(const float* h, int64_t ldh, const float* w, const float* bias,
                         const float* target, const float* loss_ws, int64_t batch, int64_t K,
                         int act_last, int act_prev, int loss_kind, float loss_threshold,
                         float* p, float* loss_out, float* gz, float* dW, float* db,
                         float* gprev, int64_t ld_gprev, void* gprev_hi, void* gprev_lo,
                         int64_t ld_gprev_bf16, void* scratch, void* stream);

int dlrm_b200_dense_update(float* param, const float* grad, float* state /*NULL for SGD*/,
                           int64_t n, int optimizer, float lr, float eps, void* stream);


/* ------------------------------------------------------------------------------------------
 * tcgen05 GEMM back end of apply_mlp and its autograd (aten::addmm, dlrm_s_pytorch.py:399-405):
 *   D[M,N] = sum_k A(m,k) * B(n,k),  bf16 operands staged by TMA, fp32 accumulation in TMEM.
 * Operands are (hi, lo) bf16 pairs of fp32 values (hi = bf16(x), lo = bf16(x - hi)); mode_x3
 * computes hi*hi + hi*lo + lo*hi (fp32-grade), otherwise hi*hi only.  An operand is K-major
 * ([rows, K], ld = row stride) or MN-major ([K, rows]); ld in elements, multiple of 8.
 * Bias: forward layers add `bias` (fp32) in the epilogue; the weight-gradient GEMMs get the bias gradient
 * for free from a constant-1 column of the activations ([dW | db] = gz^T [X | 1]).
 * Epilogue (all optional): act(); multiply by act'(y), y = mask_hi + mask_lo [M, ldmask];
 * fp32 store (split-K: one slab per split, reduced by dense_update_pack); one column diverted to
 * out_col (bias gradient); (hi, lo) bf16 stores in normal [M, ld_out] and transposed [N, ld_outT]
 * layout.  A plan holds the TMA descriptors: create once per buffer set, run every step.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  const void* A_hi; const void* A_lo; int64_t lda; int a_mn_major;
  const void* B_hi; const void* B_lo; int64_t ldb; int b_mn_major;
  int64_t M, N, K;
  int mode_x3;
  int split_k;      /* >= 1 */
  int tile_n;       /* 0 = auto, or 32 / 64 / 128 */
  int act;          /* DLRM_ACT_* applied to the accumulator */
  int mask_act;     /* DLRM_ACT_*: multiply by act'(y) */
  const void* mask_hi; const void* mask_lo; int64_t ldmask;
  float* out_f32; int64_t ld_f32; int64_t slab_stride;
  void* out_hi; void* out_lo; int64_t ld_out;
  void* outT_hi; void* outT_lo; int64_t ld_outT;
  float* out_col; int64_t col_index; int64_t col_slab_stride;
  const float* bias;  /* optional fp32 [N] added to the accumulator before act (nn.Linear bias in fp32) */
} dlrm_gemm_tc_desc_t;

int dlrm_b200_gemm_tc_plan_create(const dlrm_gemm_tc_desc_t* desc /*[host]*/, void** plan);
int dlrm_b200_gemm_tc_plan_info(void* plan, int* tile_n, int* stages, int* splits, int* ctas);
int dlrm_b200_gemm_tc_run(void* plan, void* stream);
int dlrm_b200_gemm_tc_plan_destroy(void* plan);

/* A whole MLP chain in ONE persistent launch (apply_mlp: `layers(x)`, dlrm_s_pytorch.py:399-405, and its
 * autograd: the dgrad chain with every weight-gradient GEMM beside it, :1613).  The plans' output tiles form
 * one topologically ordered task list (plan order; inside a plan: n tile, m tile, k split); one CTA per SM
 * claims tasks from a device-side queue and runs TMA -> tcgen05.mma -> epilogue with up to 4 accumulators
 * in tensor memory, so the epilogue of a tile overlaps the MMAs of the next.  dep[i] >= 0 names the plan
 * that produces plan i's A operand inside this launch: a task waits until every 128-row block of that
 * producer covering its A rows (dep_on_k[i] = 0: the task's own m tile; = 1: the task's k range, i.e. a
 * weight-gradient GEMM reducing over the batch) is complete (per-block counters, release/acquire at GPU
 * scope + async-proxy fence before the TMA reads).  Results are bit-identical to running the plans one by
 * one.  counters: caller-owned, zero-initialised int32[counters_len], counters_len >= 2 + sum of the
 * plans' m tiles; the kernel leaves it zeroed.  Plans must outlive the chain. */
int dlrm_b200_gemm_chain_create(void* const* plans /*[host][n]*/, const int* dep /*[host][n]*/,
                                const int* dep_on_k /*[host][n]*/, int n, int32_t* counters,
                                int64_t counters_len, void** chain);
int dlrm_b200_gemm_chain_info(void* chain, int* tasks, int* ctas, int* stages, int* smem_bytes);
/* Optional per-task timeline for tools/chain_timeline.py: trace = device uint64[8 * tasks] (globaltimer ns: task
 * claimed, dependencies ready, last TMA issued, first operands landed, last MMA issued, accumulator ready,
 * epilogue + signal done; word 7 = SM id), or NULL to switch it off (default). */
int dlrm_b200_gemm_chain_set_trace(void* chain, uint64_t* trace);
int dlrm_b200_gemm_chain_run(void* chain, void* stream);
int dlrm_b200_gemm_chain_destroy(void* chain);

/* fp32 [M,N] (row stride ldx) -> (hi, lo) bf16 [M, ld_out]; lo may be NULL */
int dlrm_b200_split_bf16(const float* X, int64_t ldx, int64_t M, int64_t N, void* hi, void* lo,
                         int64_t ld_out, void* stream);

/* Dense optimizer step fused with the split-K slab reduction of dW/db and the refresh of the
 * [N, K+1] = [W | bias] (hi, lo) operand copies.  optimizer = DLRM_OPT_*, -1 (pack only) or
 * -2 (only fold the slabs into slab 0, for a cross-rank all-reduce before the update). */
typedef struct {
  float* W; float* b; float* sW; float* sb;
  const float* dW; const float* db;
  void* pack_hi; void* pack_lo;
  int64_t slab_stride; int64_t N; int64_t K; int64_t ld_pack; int64_t num_slabs;
} dlrm_dense_layer_t;
int dlrm_b200_dense_update_pack(const dlrm_dense_layer_t* layers /*[host]*/, int num_layers, int optimizer,
                                float lr, float eps, void* stream);

/* ------------------------------------------------------------------------------------------
 * Sharded placement (dlrm_b200/placement.py): the pieces around the gather / update kernels.
 * ------------------------------------------------------------------------------------------ */
/* Row-split table, remote-read forward: the rank that owns the samples pools each bag itself, in index order,
 * reading every row from the rank that stores it (shard_weight[s] = base of rows [s*rows_per_shard, ...), a
 * peer-mapped pointer for s != own rank): 512-byte NVLink loads inside the gather kernel instead of partial sums
 * + reduction.  out[b*out_stride + out_off .. +dim) for the `batch` bags described by offsets / indices. */
typedef struct {
  const float* shard_weight[DLRM_B200_MAX_PEERS];
  int32_t num_shards;
  int64_t rows_per_shard;
  int64_t rows;
  int64_t ld;
  const void* indices; const void* offsets; int64_t nnz;
  int64_t out_off, out_stride;
} dlrm_emb_remote_table_t;
int dlrm_b200_emb_bag_fwd_remote(const dlrm_emb_remote_table_t* tables /*[host]*/, int num_tables, int dim,
                                 int64_t batch, int idx_bytes, int include_last, float* out, void* stream);

/* Tiny tables (a few to a few hundred rows, hit thousands of times per step at MLPerf batch sizes): dense
 * two-pass coalesce + row update instead of the per-row list walk (csrc/emb_small.cu).  Same semantics as
 * dlrm_b200_emb_bwd_update (grad.coalesce() + optim/rwsadagrad.py:117-143 / sparse SGD), deterministic.
 * Every table needs use_dy_off; `scratch` holds the per-chunk partial sums
 * (dlrm_b200_emb_bwd_small_scratch_bytes(total rows of the call, dim, batch) bytes). */
int64_t dlrm_b200_emb_bwd_small_scratch_bytes(int64_t total_small_rows, int dim, int64_t batch);
int dlrm_b200_emb_bwd_small_update(const dlrm_emb_bwd_table_t* tables /*[host]*/, int num_tables, int dim,
                                   int64_t batch, int idx_bytes, int include_last, const float* dY,
                                   const float* const* peer_dY /*[host][world] or NULL*/, int world,
                                   int64_t batch_local, int64_t dy_stride_sample, int optimizer, float lr,
                                   float eps, float* scratch, int64_t scratch_bytes, void* stream);
/* Row-split tables: T[b, slot_feature[s], :] = sum of the slabs [slot_first[s], slot_first[s+1]) of
 * partial ([slab][batch][dim]), in slab order. */
int dlrm_b200_emb_reduce_partials(const float* partial, float* T, int64_t ldt, int64_t batch, int dim,
                                  const int* slot_feature /*[host]*/, const int* slot_first /*[host][n+1]*/,
                                  int num_slots, void* stream);
/* n <= 64 contiguous blocks (16-byte aligned pointers and sizes) copied in ONE launch; destinations may be
 * peer-mapped (the index exchange of a sharded step). */
int dlrm_b200_block_copy(const void* const* src /*[host]*/, void* const* dst /*[host]*/,
                         const int64_t* nbytes /*[host]*/, int n, void* stream);
/* Device-side synthetic batch of the MLPerf multi-hot distribution (torchrec_dlrm/multi_hot.py:80-127):
 * out[k] = [batch, hot[k]] indices (idx_bytes wide) of table table_ids[k] for global samples
 * [sample0, sample0 + batch) of `step`; optional dense features X [batch, m_den] and rounded targets [batch].
 * Bit-identical to dlrm_b200/mlperf.py (host). */
int dlrm_b200_gen_multihot(void* const* out /*[host]*/, const int64_t* rows /*[host]*/, const int* hot /*[host]*/,
                           const int* table_ids /*[host]*/, int num_tables, int idx_bytes, uint64_t seed,
                           uint64_t step, int64_t sample0, int64_t batch, float* X, float* target, int m_den,
                           void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DLRM_B200_H_ */
