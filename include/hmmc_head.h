/*
 * hmmc_head.h -- C ABI of libhmmc_head.so, the B200 (sm_100a) implementation of the
 * HMMC hierarchical-matching contrastive head.
 *
 * The reference (cheetah003/HMMC) is pure Python / PyTorch: it has no FFI layer, its
 * "interface" for this path is a handful of Python methods.  Each entry point below
 * names the reference function (file:line, relative to the reference root) whose
 * arithmetic it replaces; hmmc_b200/ (the Python host side) keeps the reference's
 * method names and signatures and calls these through ctypes.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host;
 *   - no allocation, no global state except the last-error string: outputs and
 *     workspaces are caller-allocated (sizes from the *_workspace_bytes helpers);
 *   - every launch goes to `stream` (a cudaStream_t passed as void*), nothing
 *     synchronises;
 *   - return value 0 = ok, negative = error (text from hmmc_last_error());
 *   - matrices are row-major; "ld" = leading dimension in elements.
 *
 * Precision modes (`prec`): how the big contractions are carried out
 *   HMMC_PREC_FP32   CUDA-core fp32 FMA (reference-grade, slow, no tensor cores)
 *   HMMC_PREC_BF16   tcgen05 bf16 x bf16 -> fp32 (fast; ~1e-3 on logits)
 *   HMMC_PREC_BF16X3 tcgen05 with operands split hi+lo in bf16, three products
 *                    hi*hi + hi*lo + lo*hi accumulated in fp32 (~1e-6, fp32-parity)
 */
#ifndef HMMC_HEAD_H
#define HMMC_HEAD_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HMMC_PREC_FP32 0
#define HMMC_PREC_BF16 1
#define HMMC_PREC_BF16X3 2

#define HMMC_OK 0
#define HMMC_ERR_ARG (-1)
#define HMMC_ERR_CUDA (-2)
#define HMMC_ERR_UNSUPPORTED (-3)
#define HMMC_ERR_WORKSPACE (-4)

/* positive-key layouts of hmmc_infonce_queue_fwd_bwd */
#define HMMC_POS_PAIR 0          /* contrastive_loss(q, k, queue): key row = query row              */
#define HMMC_POS_FRAME_NEIGHBOUR 1 /* frame_self_loss: query (n,f) vs keys (n,f+1) and (n,f-1)       */
#define HMMC_POS_ONE_TO_FRAMES 2   /* frame_cross_loss 1st term: query n vs keys (n,0..F-1)          */
#define HMMC_POS_FRAMES_TO_ONE 3   /* frame_cross_loss 2nd term: query (n,f) vs key n                */

const char* hmmc_last_error(void);
int hmmc_version(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
unsigned long long hmmc_launch_count(void);
/* 0 when the current device is sm_100 (B200); HMMC_ERR_UNSUPPORTED otherwise. */
int hmmc_device_check(void);

/* ------------------------------------------------------------------ operands */

/* Row L2-normalise x[R,D] and emit the GEMM operand planes.
 *   eps > 0 : x / max(||x||, eps)   (F.normalize, modules/modeling.py:289,291,250-258)
 *   eps = 0 : x / ||x||             (loose_similarity, modules/modeling.py:211,214)
 * Any of the outputs may be NULL.  packed is bf16 [R, planes*D] (plane 0 = hi,
 * plane 1 = lo = bf16(x_hat - hi)), row stride ld_packed elements. */
int hmmc_rownorm_pack(const float* x, int64_t R, int D, int64_t ldx, float eps, int planes,
                      float* xhat, float* inv_norm, void* packed, int64_t ld_packed, void* stream);

/* C[M,N] = alpha * sum_k A[m*sam + k*sak] * B[n*sbn + k*sbk], CUDA-core fp32. */
int hmmc_gemm_f32(const float* A, int64_t sam, int64_t sak, const float* B, int64_t sbn, int64_t sbk,
                  float* C, int64_t ldc, int M, int N, int K, float alpha, void* stream);

/* C[M,N] = alpha * A[M,:] . B[N,:]  on tcgen05; A, B are bf16 plane-packed operands
 * ([rows, planes*K], from hmmc_rownorm_pack).  planes = 1 -> one product, 2 -> three
 * products (BF16X3).  K % 64 == 0; M, N arbitrary (edge tiles are masked). */
int hmmc_umma_gemm_nt(const void* A, int64_t lda, const void* B, int64_t ldb, float* C, int64_t ldc,
                      int M, int N, int K, int planes, float alpha, void* stream);
/* The same contraction with the tile configuration named explicitly (measurement aid, tools/gemm_bench.py):
 * tiling 0 = as hmmc_umma_gemm_nt chooses, 128 / 256 = single-CTA kernel with 128 x tiling tiles,
 * 512 = CTA-pair kernel (tcgen05 cta_group::2, 256 x 256 tiles). */
int hmmc_umma_gemm_nt_tiled(const void* A, int64_t lda, const void* B, int64_t ldb, float* C, int64_t ldc,
                            int M, int N, int K, int planes, float alpha, int tiling, void* stream);

/* ------------------------------------------------------- pre-train head (MoCo) */

/* A negative queue as the kernels see it.  `dk` is the reference's buffer
 * (modules/modeling.py:138-149): fp32 [D, Kq], unit-norm columns.  pack_kd / pack_dk
 * are derived bf16 operand copies kept in step by hmmc_queue_pack / hmmc_enqueue_norm:
 *   pack_kd  [Kq, planes*D ]  (B operand of  S = q_hat . Q)
 *   pack_dk  [D,  planes*Kq]  (B operand of  U = E . Q^T)
 * They may be NULL when only HMMC_PREC_FP32 is used. */
typedef struct {
  float* dk;
  void* pack_kd;
  void* pack_dk;
  int32_t D;
  int32_t Kq;
  int32_t planes;
  int32_t reserved;
} hmmc_queue;

/* (Re)build pack_kd / pack_dk from dk (after init or load_state_dict). */
int hmmc_queue_pack(const hmmc_queue* q, void* stream);

size_t hmmc_infonce_workspace_bytes(int64_t R, int D, int Kq, int prec);

/* Fused InfoNCE-vs-queue forward + backward for one query block.
 * Replaces contrastive_loss (modules/modeling.py:286-313) and, through `pos_mode`,
 * the loops of frame_self_loss (:315-323) and frame_cross_loss (:325-332):
 *   loss_out[0] += weight * sum_terms mean_n [ log(e^{l+} + sum_j e^{l_nj}) - l+ ]
 *   dq          =  d(that sum)/dq          (fp32 [R,D]; only q receives a gradient)
 * q is [b*Fq, D] (row (n,f) = n*Fq+f), keys is [b*Fk, D]; both raw (un-normalised).
 * dq may be NULL (forward only). */
int hmmc_infonce_queue_fwd_bwd(const float* q, const float* keys, int pos_mode, int b, int Fq, int Fk, int D,
                               const hmmc_queue* queue, float temperature, float weight, int prec,
                               float* loss_out, float* dq, void* workspace, size_t workspace_bytes,
                               void* stream);

/* The whole pre-train head loss of BirdPreTrainedModel.forward (modules/modeling.py:385-400,424;
 * dataset != "bird"), forward and backward in four launches (normalise + pack, S-GEMM with the
 * InfoNCE epilogue, U-GEMM, finish):
 *   losses_out[0] = w_fam*FAM + w_vtm*VTM + w_ftm*FTM,  losses_out[1..3] = FAM, VTM, FTM
 *   FAM = frame_self_loss(frame_pred, frame_proj_k, q_frame_proj)
 *   VTM = contrastive_loss(v_fea, title_fea_k, q_title) + contrastive_loss(title_fea, v_fea_k, q_v)
 *   FTM = frame_cross_loss(frame_fea, frame_fea_k, q_frame_cross, title_fea, title_fea_k, q_title)
 * and d(losses_out[0])/d{v_fea, title_fea, frame_fea, frame_pred} (all four NULL = forward only).
 * [b,D] / [b,F,D] fp32 contiguous tensors. */
typedef struct {
  const float* v_fea;
  const float* title_fea;
  const float* frame_fea;
  const float* frame_pred;
  const float* v_fea_k;
  const float* title_fea_k;
  const float* frame_fea_k;
  const float* frame_proj_k;
  float* d_v_fea;
  float* d_title_fea;
  float* d_frame_fea;
  float* d_frame_pred;
} hmmc_pretrain_io;
size_t hmmc_pretrain_head_workspace_bytes(int b, int F, int D, int K, int prec);
int hmmc_pretrain_head_fwd_bwd(const hmmc_pretrain_io* io, int b, int F, int D, const hmmc_queue* q_v,
                               const hmmc_queue* q_title, const hmmc_queue* q_frame_proj,
                               const hmmc_queue* q_frame_cross, float temperature, float w_fam, float w_vtm,
                               float w_ftm, int use_frame_fea, int prec, float* losses_out, void* workspace,
                               size_t workspace_bytes, void* stream);
/* The same call with an explicit schedule (NULL = the call above):
 *   phase 1 = normalise the queries and run the two GEMM passes against the queues (needs only the query
 *             tensors and the gradient buffers; the key pointers of `io` may be NULL),
 *   phase 2 = positives, losses, gradients (needs the keys), on the SAME arguments and an untouched workspace,
 *   phase 0 = both.
 * In the reference the queries exist before `_momentum_update()` and the key encoders run
 * (modules/modeling.py:340-377), so phase 1 can execute beside them on another stream.
 * reserved_sms: SMs the persistent GEMM grids of THIS call leave free (0 = use them all) -- for a
 *             bandwidth-bound kernel (the momentum update) or a collective running beside them; a
 *             persistent CTA that cannot become resident would hold its tiles back.  Per call: the
 *             library keeps no scheduling state, so concurrent callers (eval threads, side streams)
 *             cannot disturb each other.
 * queues_released: NULL or a cudaEvent_t recorded on `stream` right after the last kernel that reads the
 * queues; the enqueue of the step (which overwrites queue columns) can then run on another stream that
 * waits for it, beside the rest of the loss. */
typedef struct {
  int phase;
  int reserved_sms;
  void* queues_released;
} hmmc_head_schedule;
int hmmc_pretrain_head_fwd_bwd_sched(const hmmc_pretrain_io* io, int b, int F, int D, const hmmc_queue* q_v,
                                     const hmmc_queue* q_title, const hmmc_queue* q_frame_proj,
                                     const hmmc_queue* q_frame_cross, float temperature, float w_fam, float w_vtm,
                                     float w_ftm, int use_frame_fea, int prec, float* losses_out,
                                     const hmmc_head_schedule* sched, void* workspace, size_t workspace_bytes,
                                     void* stream);

/* _momentum_update (modules/modeling.py:238-242): p_k <- p_k*m + p*(1-m) for a table of
 * tensors, each in its own dtype (0 = fp32, 1 = fp16, 2 = bf16), three separately
 * rounded ops like the reference.  The three tables are device arrays of length n. */
int hmmc_ema_multi(const uint64_t* pk_ptrs, const uint64_t* p_ptrs, const int64_t* numels,
                   const int32_t* dtypes, const int64_t* block_offsets, int n, int64_t total_blocks,
                   float m, float one_minus_m, void* stream);
/* elements handled by one thread block of hmmc_ema_multi (for building block_offsets) */
int hmmc_ema_block_elems(void);

/* SURVEY.md §8(f) N3 — the optimizer step that follows the head's backward:
 *   torch.nn.utils.clip_grad_norm_(model.parameters(), G)   (main_pretrain.py:277,
 *                                                            main_task_retrieval.py:291)
 *   BertAdam.step()                                          (modules/optimization.py:103-168)
 * for a table of n fp32 tensors in three launches.  Tables are device arrays as for
 * hmmc_ema_multi (block_offsets has n+1 entries, built with hmmc_ema_block_elems()).
 * hyper is a device table [n][8] fp32: lr_scheduled, weight_decay, b1, 1-b1, b2, 1-b2, e,
 * max_grad_norm (the per-parameter clip inside step(); <= 0: none).  global_max_norm <= 0: no
 * global clip.  Per element, in the reference's order and rounding:
 *   g <- g*cg*ct;  m <- m*b1 + (1-b1)*g;  v <- v*b2 + (1-b2)*g*g;
 *   u <- m/(sqrt(v)+e) [+ wd*p];  p <- p - lr*u
 * with cg = min(1, G/(||g_all|| + 1e-6)), ct = min(1, max_grad_norm/(||g_t||*cg + 1e-6)).
 * write_back_grads != 0 also stores the clipped gradients (the reference clips p.grad in
 * place and zeroes it right after).  inv_scale: NULL or a device fp32 scalar 1/scale of
 * torch.cuda.amp.GradScaler (--enable_amp, main_pretrain.py:267-284): every gradient element is first
 * replaced by fl(g * inv_scale), the value GradScaler.unscale_ would have stored, so the norms, both clips
 * and the update see the unscaled gradients without a separate pass over them.  norms_out: NULL or
 * device fp32 [n+1] receiving the per-tensor gradient norms and, last, the total norm
 * clip_grad_norm_ returns. */
size_t hmmc_bert_adam_workspace_bytes(int n, int64_t total_blocks);
int hmmc_bert_adam_multi(const uint64_t* p_ptrs, const uint64_t* g_ptrs, const uint64_t* m_ptrs,
                         const uint64_t* v_ptrs, const int64_t* numels, const int32_t* dtypes,
                         const int64_t* block_offsets, int n, int64_t total_blocks, const float* hyper,
                         float global_max_norm, int write_back_grads, const float* inv_scale, float* norms_out,
                         void* workspace, size_t workspace_bytes, void* stream);
/* clip_grad_norm_ on its own: gradients scaled in place by min(1, max_norm/(total + 1e-6));
 * workspace sized by hmmc_bert_adam_workspace_bytes. */
int hmmc_clip_grad_norm_multi(const uint64_t* g_ptrs, const int64_t* numels, const int64_t* block_offsets, int n,
                              int64_t total_blocks, float max_norm, float* norms_out, void* workspace,
                              size_t workspace_bytes, void* stream);

/* _dequeue_and_enqueue (modules/modeling.py:244-284) after the all-gather:
 * L2-normalise (eps 1e-12) the gathered keys and write them as queue columns
 * [ptr, ptr+B) (frame queues: columns [(ptr)*F, (ptr+B)*F), frame index fastest), in
 * dk and in the packed copies; then ptr <- (ptr+B) % K on the device.
 * `gathered` is the all-gather output [W][b][row_elems] with one rank's row =
 * [v | tag | title | frame_fea(F*D) | frame_proj(F*D)]; queues order: v, tag, title,
 * frame_cross (gets frame_fea), frame_proj.  queue_ptr is the int64[1] buffer.
 * ptr_host is the host's copy of the pointer (the kernel takes it by value: no device
 * read, no sync); it also drives the bounds check the reference performs through slice
 * assignment.  The kernel stores (ptr_host + B) % K into queue_ptr.
 * ptr_host = -1: the pointer is read and advanced on the device (CUDA-graph capture / replay,
 * where no host value may be baked into the launch); requires K % B == 0 and a pointer that
 * is a multiple of B, like the reference's own no-wrap condition. */
int hmmc_enqueue_norm(const float* gathered, int W, int b, int F, int D, const hmmc_queue* queues5,
                      int64_t* queue_ptr, int64_t ptr_host, int K, float* scratch /* (3+2F)*W*b floats, or NULL */,
                      int32_t* staged, const int32_t* slot_epoch, int64_t slot_stride, int prenormalised,
                      void* stream);
/* staged: NULL, or the device mark hmmc_pack_rows set when it filled `gathered`'s send buffer (deferred
 * schedule: the keys of step i are exchanged and enqueued beside step i+1).  The enqueue then happens only if
 * the mark is set and clears it, so issuing it twice for the same keys - an eager flush followed by the replay
 * of a captured step that carries the same enqueue - writes them once.  Needs ptr_host < 0.
 * slot_epoch: NULL, or the exchange counter of hmmc_peer_wait: `gathered` is then the base of a two-slot
 * receive buffer and the keys are read from slot (*slot_epoch - 1) & 1, slot_stride elements apart.
 * prenormalised != 0: `gathered` holds unit vectors already (hmmc_pack_rows with norm_dim = D on every rank). */

/* Same, reading the five key tensors in place ([B,D] x3, [B,F,D] x2, contiguous): the
 * single-process case needs no gather and no packed copy. */
int hmmc_enqueue_norm_direct(const float* v_k, const float* tag_k, const float* title_k, const float* frame_fea_k,
                             const float* frame_proj_k, int B, int F, int D, const hmmc_queue* queues5,
                             int64_t* queue_ptr, int64_t ptr_host, int K, float* scratch, void* stream);

/* Key exchange over peer memory (replaces the all-gather of dist_collect for the deferred enqueue; the reference
 * gathers with torch.distributed, modules/modeling.py:25-36 and :249-262).  Every rank of the node owns a receive
 * buffer of two slots [2][W][elems] fp32 and W int32 flags, both mapped into its peers (symmetric memory; the host
 * passes the W base addresses as this process sees them).  hmmc_peer_push_rows copies this rank's `elems` floats
 * into block `rank` of slot (*epoch & 1) of EVERY rank's buffer with plain stores over NVLink and then sets
 * flag[rank] = *epoch + 1 in every rank.  hmmc_peer_wait blocks the stream until all W flags of this rank have
 * reached *epoch + 1 and then increments *epoch.  epoch and done_counter are device words owned by the caller,
 * zero before the first exchange; nothing here depends on host state, so a captured step replays unchanged.
 * Every rank must issue the same sequence of exchanges. */
#define HMMC_MAX_PEERS 16
int hmmc_peer_push_rows(const float* send, int64_t elems, const uint64_t* peer_bufs_host,
                        const uint64_t* peer_flags_host, int W, int rank, int64_t slot_stride, const int32_t* epoch,
                        uint32_t* done_counter, void* stream);
int hmmc_peer_wait(const int32_t* my_flags, int W, int32_t* epoch, void* stream);

/* x_t *= scale[0] (device scalar) for up to 8 fp32 tensors in one launch: applies the upstream
 * gradient to the gradients the fused heads produced together with the loss. */
int hmmc_scale_tensors(const uint64_t* ptrs_host, const int64_t* numels_host, int n, const float* scale,
                       void* stream);

/* gather n row-blocks src_i[rows, width_i] into dst[rows, sum width_i] (the packed
 * send buffer of the key / embedding all-gather) and the inverse.
 * norm_dim > 0: every norm_dim-vector is written L2-normalised (eps 1e-12), bit-identical to what
 * hmmc_enqueue_norm computes from the raw keys; the enqueue is then told `prenormalised`. */
int hmmc_pack_rows(const uint64_t* src_ptrs_host, const int32_t* widths_host, int n, int64_t rows,
                   float* dst, int32_t* staged /* NULL, or a device mark set to 1 */, int norm_dim, void* stream);
int hmmc_unpack_rows(const float* src, const uint64_t* dst_ptrs_host, const int32_t* widths_host, int n,
                     int64_t rows, void* stream);

/* SURVEY.md §8(f) N1 — the projector / predictor MLP applied to the frame features right before the
 * head (modules/modeling.py:788-807 with num_layers = 2, used at :355-377):
 *   y = Linear2(ReLU(BatchNorm1d(Linear1(x))))      x [M, Din], W1 [Dh, Din], W2 [Dout, Dh]
 * forward and backward on the tensor cores (prec bf16 or bf16x3).  The reference converts these
 * MLPs to SyncBatchNorm (:127-129): the batch statistics span all ranks' rows, so each direction
 * is two calls around one exchange the host performs on a device buffer of 2*Dh doubles:
 *   hmmc_mlp_fwd_a  -> *stats_out = (sum h, sum h^2) of this rank   [host: all-reduce SUM]
 *   hmmc_mlp_fwd_b  (count = rows of all ranks; training = 0 uses the running statistics instead;
 *                    training = 1 also updates running_mean / running_var with `momentum`)
 *   hmmc_mlp_bwd_a  -> dW2, db2, *sums_out = (sum dZ, sum dZ*xhat)  [host: all-reduce SUM]
 *   hmmc_mlp_bwd_b  -> dx, dW1, db1, dgamma, dbeta (parameter gradients are this rank's, as with
 *                    SyncBatchNorm + DDP)
 * `ctx` (hmmc_mlp_ctx_bytes) carries everything between the four calls and must stay untouched from
 * fwd_a to bwd_b.  Any gradient pointer may be NULL. */
typedef struct {
  const float* W1;      /* [Dh, Din]  linear_hidden.1.weight */
  const float* b1;      /* [Dh]       linear_hidden.1.bias   */
  const float* gamma;   /* [Dh]       linear_hidden.2.weight */
  const float* beta;    /* [Dh]       linear_hidden.2.bias   */
  const float* W2;      /* [Dout, Dh] linear_out.weight      */
  const float* b2;      /* [Dout]     linear_out.bias        */
  float* running_mean;  /* [Dh] or NULL */
  float* running_var;   /* [Dh] or NULL */
} hmmc_mlp_params;
size_t hmmc_mlp_ctx_bytes(int M, int Din, int Dh, int Dout, int prec, int need_grad);
int hmmc_mlp_fwd_a(const float* x, int M, int Din, int Dh, int Dout, const hmmc_mlp_params* p, int prec, int need_grad,
                   void* ctx, size_t ctx_bytes, double** stats_out, void* stream);
int hmmc_mlp_fwd_b(int M, int Din, int Dh, int Dout, const hmmc_mlp_params* p, float eps, float momentum, double count,
                   int training, int prec, int need_grad, void* ctx, size_t ctx_bytes, float* y, void* stream);
int hmmc_mlp_bwd_a(const float* dy, int M, int Din, int Dh, int Dout, const hmmc_mlp_params* p, int prec, void* ctx,
                   size_t ctx_bytes, float* dW2, float* db2, double** sums_out, void* stream);
int hmmc_mlp_bwd_b(int M, int Din, int Dh, int Dout, const hmmc_mlp_params* p, double count, int prec, void* ctx,
                   size_t ctx_bytes, float* dx, float* dW1, float* db1, float* dgamma, float* dbeta, void* stream);

/* Tail of VisualEncoder.forward (modules/module_cross.py:207-213), the op that produces the video
 * embedding the head consumes:  out[b] = mean_f normalize(temporal[b,f] + original[b,f])  with
 * temporal NULL when the temporal transformer is off.  temporal / original [B,F,D], out [B,D].
 * Backward: dhidden [B,F,D] is the gradient w.r.t. the sum, i.e. w.r.t. BOTH branches. */
int hmmc_visual_tail_fwd(const float* temporal, const float* original, int B, int F, int D, float* out, void* stream);
int hmmc_visual_tail_bwd(const float* temporal, const float* original, const float* dout, int B, int F, int D,
                         float* dhidden, void* stream);

/* ---------------------------------------------------------- fine-tune head (HM) */

/* loose_similarity (modules/modeling.py:207-229), forward.  vis is [Bv*Fv, D]
 * (Fv = 1 for the 2-D case); out is [Bt, Bv*Fv] = [Bt,Bv,Fv] contiguous. */
size_t hmmc_similarity_workspace_bytes(int64_t Bt, int64_t Bv, int Fv, int D, int prec);
int hmmc_loose_similarity_fwd(const float* seq, int64_t Bt, const float* vis, int64_t Bv, int Fv, int D,
                              float scale, int prec, float* out, void* workspace, size_t workspace_bytes,
                              void* stream);
/* backward of the above: dseq [Bt,D], dvis [Bv*Fv,D] from dout [Bt,Bv*Fv]. */
int hmmc_loose_similarity_bwd(const float* seq, int64_t Bt, const float* vis, int64_t Bv, int Fv, int D,
                              float scale, const float* dout, float* dseq, float* dvis, void* workspace,
                              size_t workspace_bytes, void* stream);

/* CrossEn.forward (modules/until_module.py:196-205) + its backward:
 * loss_out[0] = -mean(diag(log_softmax(S, -1))), dS = (softmax - I)/B (NULL to skip). */
int hmmc_cross_en_fwd_bwd(const float* S, int64_t lds, int B, float* loss_out, float* dS, int64_t ldds,
                          float* row_scratch /* B floats */, void* stream);

/* The whole fine-tune head after the gather (BirdModel.forward, modules/modeling.py:702-709
 * with frame_loss :665-673): loss = w_vtm*(CE(S)+CE(S^T)) + w_ftm/F * sum_f (CE(S_f)+CE(S_f^T)).
 * text [B,D], video [B,D], frames [B,F,D]; gradients may be NULL (forward only). */
size_t hmmc_sym_ce_workspace_bytes(int B, int F, int D, int prec);
int hmmc_sym_ce_fwd_bwd(const float* text, const float* video, const float* frames, int B, int F, int D,
                        float scale, float w_vtm, float w_ftm, int prec, float* loss_out, float* dtext,
                        float* dvideo, float* dframes, void* workspace, size_t workspace_bytes, void* stream);
/* Same head on the layout the differentiable all-gather moves (modules/modeling.py:698-700 gathers the
 * three tensors; here they travel as one row [text(D) | video(D) | frames(F*D)] per sample):
 * packed [B, (2+F)*D]; dpacked (NULL: forward only) receives the gradient in the same layout, which is
 * what the gather's backward (reduce-scatter) consumes.  No un-pack / re-pack copies on the
 * tensor-core path. */
size_t hmmc_sym_ce_packed_workspace_bytes(int B, int F, int D, int prec);
int hmmc_sym_ce_packed_fwd_bwd(const float* packed, int B, int F, int D, float scale, float w_vtm, float w_ftm,
                               int prec, float* loss_out, float* dpacked, void* workspace, size_t workspace_bytes,
                               void* stream);

/* ------------------------------------------------------------------------ eval */

/* One (text tile x gallery tile) of _run_on_single_gpu (main_task_retrieval.py:321-357):
 *   sim  [Nt,Nv] = s * t_hat . v_hat
 *   fsim [Nt,Nv] = mean over the top_k frames of s * t_hat . f_hat   (torch.topk + mean)
 * video [Nv,D], frames [Nv,F,D].  Either output may be NULL.  sim == fsim (one buffer): that buffer receives
 * sim + fsim, the sum eval_epoch forms on the host (main_task_retrieval.py:512-513).
 * With 12 frames, top_k <= 4 and a tensor-core precision the call is two launches: one kernel normalises and
 * packs the captions and the gallery ([video | 12 frames] column groups), one GEMM sweep whose epilogue pools the
 * top-k frames in registers and writes each score once (no [Nt, Nv*F] frame-similarity matrix). */
size_t hmmc_sim_topk_workspace_bytes(int64_t Nt, int64_t Nv, int F, int D, int prec);
int hmmc_sim_topk_fwd(const float* text, int64_t Nt, const float* video, const float* frames, int64_t Nv,
                      int F, int D, float scale, int top_k, int prec, float* sim, float* fsim, int64_t ld_out,
                      void* workspace, size_t workspace_bytes, void* stream);

/* Rank counting on a materialised similarity matrix (metrics.py:12-39, 49-86):
 *   t2v[s] = #{ j : sim[s,j] > sim[s,gt[s]] }                       (int32 [Nt])
 *   v2t[j] = #{ g != j : max_{s in g} sim[s,j] > max_{s in group j} sim[s,j] }   (int32 [Nv])
 * gt[s] = index of the video text s belongs to; texts of one video are contiguous
 * and group_start[g] .. group_start[g+1] delimits them (int32 [Nv+1]).  NaN is
 * treated as -inf in v2t (metrics.py:83).  Either output may be NULL. */
int hmmc_rank_count(const float* sim, int64_t lds, int Nt, int Nv, const int32_t* gt,
                    const int32_t* group_start, int32_t* t2v, int32_t* v2t, float* theta_scratch /* Nv floats */,
                    void* stream);

/* ---- large-gallery eval without materialising the matrix (BASELINE config 5) --------------
 *   score[s,j] = scale*t_hat_s.v_hat_j + mean top_k_f scale*t_hat_s.f_hat_jf
 *   t2v_cnt[s] = #{ j != gt(s) : score[s,j] > score[s,gt(s)] }
 *   v2t_cnt[j] = #{ groups g != j : max_{s in g} score[s,j] > theta_j },  theta_j = max_{s in group j} score[s,j]
 * i.e. the multi-sentence ranks of metrics.py:49-86 (a square set = one caption per video), for
 * THIS rank's gallery shard (videos [video_base, video_base + Nv_local)).
 * Packed operands: texts [Nt_pad, planes*D] bf16, Nt_pad % 128 == 0, packed row i = text
 * src_row[i] (-1 = zero padding) with the captions of one video contiguous and never straddling a
 * 128-row tile; gallery [ceil(Nv/16)*16*(1+F), planes*D] with 1+F rows per video.
 * grp[i] = global video id of packed caption i (-1 = padding).
 * Driver (hmmc_b200/retrieval.py): pack -> gt_scores on the diagonal tiles -> (all-reduce) ->
 * theta -> counting sweep -> (all-reduce t2v_cnt). */
int hmmc_eval_fused_supported(int F, int D, int top_k);
int hmmc_eval_pack_text(const float* text, const int32_t* src_row, int64_t rows_pad, int D, int prec, void* out,
                        void* stream);
int hmmc_eval_pack_gallery(const float* video, const float* frames, int64_t Nv, int F, int D, int prec, void* out,
                           void* stream);
int hmmc_eval_gt_scores(const void* text_packed, const void* gallery_packed, int64_t Nt_pad, int64_t Nv_local, int D,
                        int prec, float scale, int top_k, int video_base, const int32_t* grp,
                        const int32_t* diag_tiles /* (m_blk, n_blk) pairs */, int n_diag_tiles, float* gt_score,
                        void* stream);
int hmmc_eval_theta(const float* gt_score, const int32_t* group_start_packed, const int32_t* group_count, int Nv_local,
                    float* theta, void* stream);
int hmmc_eval_fused_rank(const void* text_packed, const void* gallery_packed, int64_t Nt_pad, int64_t Nv_local, int D,
                         int prec, float scale, int top_k, int video_base, const int32_t* grp, const float* gt_score,
                         const float* theta, int32_t* t2v_cnt, int32_t* v2t_cnt, void* stream);

/* tensor_video_to_text_sim (metrics.py:79-86): out[j, g] = max_{s in group g} sim[s, j]
 * (NaN -> -inf); out is [Nv, G] row-major, G < 65536. */
int hmmc_group_max(const float* sim, int64_t lds, int Nv, int G, const int32_t* group_start, float* out,
                   void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HMMC_HEAD_H */
