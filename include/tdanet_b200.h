/*
 * tdanet_b200 -- C ABI of the B200-native TDANet separation hot path.
 *
 * This is the drop-in boundary.  The reference has no FFI seam of its own: the
 * path is a tree of torch.nn modules (look2hear/models/TDANet_best.py:402-525,
 * TDANet.py:788-913, TDANet_mult_tes.py:455-579) plus the PIT loss
 * (look2hear/losses/matrix.py:12-56, pit_wrapper.py:14-131).  Each entry point
 * below names the reference call it replaces.  Signatures are plain pointers
 * and sizes; no torch types.  All pointers are DEVICE pointers unless the name
 * says host; all floating point buffers are fp32, contiguous.
 *
 * Conventions
 *   - every function returns 0 on success, a negative TDANET_E* code otherwise,
 *     never throws, never synchronises the device, allocates no device memory;
 *     kernels are launched on the stream that is passed in;
 *   - the library borrows every buffer for the duration of the enqueued work;
 *   - tdanet_last_error() returns a thread-local message for the last failure.
 *
 * Boundary tensors keep the reference layout (wav [B,T], est [B,n_src,T],
 * weights exactly as in state_dict()).  Inside the workspace activations are
 * stored channels-last, [B, L, C] with C contiguous.
 */
#ifndef TDANET_B200_H
#define TDANET_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define TDANET_API __attribute__((visibility("default")))
#else
#define TDANET_API
#endif

#define TDANET_ABI_VERSION 3
#define TDANET_MAX_DEPTH 8
#define TDANET_MAX_ENC 4

typedef struct CUstream_st* tdanet_stream_t; /* == cudaStream_t */

enum tdanet_error {
  TDANET_OK = 0,
  TDANET_EINVAL = -1,    /* bad shape / configuration                           */
  TDANET_ENOSPACE = -2,  /* workspace too small                                 */
  TDANET_ECUDA = -3,     /* a CUDA runtime / driver call failed                 */
  TDANET_EUNSUPPORTED = -4 /* no sm_100 device, or a size outside kernel limits */
};

enum tdanet_variant {
  TDANET_BEST = 0,    /* look2hear.models.TDANetBest   (TDANet_best.py)     */
  TDANET_FORK = 1,    /* look2hear.models.TDANet       (TDANet.py)          */
  TDANET_MULTRES = 2, /* look2hear.models.TDANetMultRes(TDANet_mult_tes.py) */
  TDANET_ORIGIN = 3   /* look2hear.models.TDANetOrigin / TDANetYang (TDANet_origin.py, TDANet_yang.py):
                         GroupNorm GlobLN, average-pool gather, additive injection, batch-axis attention */
};

enum tdanet_gemm_mode {
  TDANET_GEMM_FP32 = 0,   /* CUDA-core fp32 FMA tiles (exact-parity mode)                 */
  TDANET_GEMM_TF32 = 1,   /* tcgen05 kind::tf32, weights pre-rounded to TF32 (RN)         */
  TDANET_GEMM_TF32X3 = 2  /* tcgen05 kind::tf32, weights split hi+lo (two MMA passes)     */
};

enum tdanet_act_dtype {
  TDANET_ACT_F32 = 0,   /* every intermediate in fp32                                              */
  TDANET_ACT_BF16 = 1   /* proj_1x1 / spp_dw / top-down outputs stored as bf16 (fp32 arithmetic and
                           statistics); needs a tensor-core gemm_mode; the "bf16 mode" tolerance    */
};

/* Static description of one model instance (constructor kwargs of the reference class). */
typedef struct tdanet_config {
  int32_t variant;       /* enum tdanet_variant                                           */
  int32_t out_channels;  /* c : residual-stream width (reference kwarg out_channels)       */
  int32_t in_channels;   /* C : UConvBlock width      (reference kwarg in_channels)        */
  int32_t num_blocks;    /* shared-weight iterations  (Recurrent.iter)                     */
  int32_t depth;         /* upsampling_depth                                               */
  int32_t enc_kernel;    /* K : encoder window in samples (enc_kernel_size*sr/1000)        */
  int32_t enc_stride;    /* S = K/4                                                        */
  int32_t n_basis;       /* encoder channels: K/2+1, or out_channels for MULTRES           */
  int32_t num_sources;
  int32_t enc_convs;     /* 1, or `kernels` for MULTRES (conv k has window (k+1)*K)        */
  int32_t n_head;        /* 8 in every reference model                                     */
  int32_t gemm_mode;     /* enum tdanet_gemm_mode                                          */
  int32_t attn_group;    /* batch items that attend to each other (BEST/FORK: the reference
                            batch being emulated); 0 = the whole batch of the call         */
  int32_t act_dtype;     /* enum tdanet_act_dtype: storage of the large activations             */
  /* Training-mode regularisation (tdanet_forward_train_rng / tdanet_backward only; tdanet_forward is eval).
   * The reference hard-codes both to 0.1 (TDANet_best.py:256-259,335-337; TDANet.py:397-401,573-575).   */
  float dropout;         /* nn.MultiheadAttention(dropout), MultiHeadAttention.dropout, FFN.drop */
  float drop_path;       /* GA.drop_path (per-item stochastic depth on both GA branches)         */
} tdanet_config_t;

/* conv weight (+ optional bias) followed by a GlobLN: ConvNorm / DilatedConvNorm / ConvNormAct */
typedef struct tdanet_convnorm {
  const float* w;     /* Conv1d.weight as stored: [out, in/groups, k]                      */
  const float* b;     /* Conv1d.bias or NULL                                               */
  const float* gamma; /* GlobLN gamma  (GroupNorm.weight in the fork)                      */
  const float* beta;  /* GlobLN beta   (GroupNorm.bias   in the fork)                      */
} tdanet_convnorm_t;

/* LA (TDANet_best.py:266-292): three depthwise ConvNorms */
typedef struct tdanet_la {
  tdanet_convnorm_t local_embedding, global_embedding, global_act;
} tdanet_la_t;

/* fork only: DilatedSeparableConvNorm (TDANet.py:190-228) */
typedef struct tdanet_sepconvnorm {
  const float *dw_w, *dw_b, *pw_w, *pw_b, *gamma, *beta;
} tdanet_sepconvnorm_t;

/* Device pointers to every tensor of state_dict(); the comment is the state_dict key. */
typedef struct tdanet_weights {
  const float* enc_w[TDANET_MAX_ENC];       /* encoder.weight | encoder.conv_list.k.weight         */
  const float *ln_gamma, *ln_beta;          /* ln.gamma/beta | ln.weight/bias                      */
  const float *bottleneck_w, *bottleneck_b; /* bottleneck.*  (NULL for MULTRES)                    */
  tdanet_convnorm_t proj;                   /* sm.unet.proj_1x1.{conv,norm}                        */
  const float* proj_prelu;                  /* sm.unet.proj_1x1.act.weight                         */
  tdanet_convnorm_t spp_dw[TDANET_MAX_DEPTH];      /* sm.unet.spp_dw.k                             */
  tdanet_la_t loc_glo_fus[TDANET_MAX_DEPTH];       /* sm.unet.loc_glo_fus.k          (BEST)        */
  tdanet_sepconvnorm_t conv_pool[TDANET_MAX_DEPTH];/* sm.unet.conv_pool.k            (FORK)        */
  const float *res_w, *res_b;               /* sm.unet.res_conv.*                                  */
  const float* pe;                          /* sm.unet.globalatt.attn.pos_enc.pe [1, pe_rows, C]   */
  const float *ln1_w, *ln1_b;               /* ...attn.attn_in_norm.*                              */
  const float *in_proj_w, *in_proj_b;       /* ...attn.attn.in_proj_{weight,bias}                  */
  const float *out_proj_w, *out_proj_b;     /* ...attn.attn.out_proj.*                             */
  const float *ln2_w, *ln2_b;               /* ...attn.norm.*                                      */
  tdanet_convnorm_t fc1;                    /* sm.unet.globalatt.mlp.fc1                           */
  const float *ffn_dw_w, *ffn_dw_b;         /* sm.unet.globalatt.mlp.dwconv.*                      */
  tdanet_convnorm_t fc2;                    /* sm.unet.globalatt.mlp.fc2                           */
  tdanet_la_t last_layer[TDANET_MAX_DEPTH]; /* sm.unet.last_layer.i                                */
  const float *concat_w, *concat_b, *concat_prelu; /* sm.concat_block.{0.weight,0.bias,1.weight}   */
  const float *mask_prelu, *mask_w, *mask_b;/* mask_net.{0.weight,1.weight,1.bias}                 */
  const float* dec_w;                       /* decoder.weight [n_src*n_basis, n_src, K]            */
  int32_t pe_rows;
  int32_t reserved;
} tdanet_weights_t;

/* ------------------------------------------------------------------ library state */
TDANET_API int tdanet_abi_version(void);
/* sizeof(tdanet_config_t), sizeof(tdanet_weights_t): lets a binding verify its struct mirror */
TDANET_API int tdanet_abi_sizes(size_t* config_bytes, size_t* weights_bytes);
TDANET_API const char* tdanet_last_error(void);
/* number of kernels this library has enqueued so far in this process */
TDANET_API uint64_t tdanet_launch_count(void);
/* Per-launch CUDA-event timing (measurement aid, off by default).  While enabled, every kernel the
 * library enqueues outside stream capture is bracketed by events on its own stream;
 * tdanet_profile_dump synchronises and writes a JSON array [{"kernel","launches","ms"}] aggregated by
 * kernel since the previous dump, returning the length it needs. */
TDANET_API int tdanet_profile_enable(int on);
TDANET_API int tdanet_profile_dump(char* out, size_t cap);
/* 0 if device `dev` can run the library (compute capability 10.x) */
TDANET_API int tdanet_device_supported(int dev);
/* Deterministic statistics (process-wide, off by default; TDANET_DETERMINISTIC=1 in the environment turns it on at
 * load).  The GlobLN sums of tdanet_forward are accumulated across thread blocks with floating-point atomics, so two
 * runs of the same call differ by ~1e-7 (fp32 GEMMs) / up to ~1e-4 (TF32: a sum that moves in its last bit can flip
 * a TF32 rounding) - inside the parity tolerance, but in the way when bisecting a regression.  With the mode on every
 * such sum is accumulated exactly (integer atomics on a fixed-point pair per statistic) and tdanet_forward is
 * bit-reproducible run to run, at the cost of one small launch per producer of a statistic.  The mode enlarges the
 * inference workspace: set it BEFORE tdanet_workspace_bytes.  Training calls (tdanet_forward_train*, tdanet_backward:
 * their weight-gradient accumulation is atomic as well) are not affected. */
TDANET_API int tdanet_set_deterministic(int on);
TDANET_API int tdanet_get_deterministic(void);

/* ------------------------------------------------------------------ model forward
 * Replaces TDANet*.forward(input_wav) (TDANet_best.py:482-521, TDANet.py:869-909,
 * TDANet_mult_tes.py:540-579): pad_input, encoder, GlobLN, bottleneck, Recurrent
 * (num_blocks x shared UConvBlock), mask_net, ReLU mask, decoder, crop.
 *   wav [B, T]  ->  est [B, num_sources, T]
 */
TDANET_API int tdanet_workspace_bytes(const tdanet_config_t* cfg, int batch, int n_samples, size_t* bytes);
TDANET_API int tdanet_forward(const tdanet_config_t* cfg, const tdanet_weights_t* w, const float* wav,
                   int batch, int n_samples, float* est, void* workspace, size_t workspace_bytes,
                   tdanet_stream_t stream);

/* Locate an intermediate of the last enqueued block inside the workspace (tests / profiling).
 * name: "enc" "x0" "u" "proj" "spp0".."sppN" "ga_in" "attn_in" "qkv" "attn_ctx" "attn_out"
 *       "ga_mid" "fc1" "ffn_dw" "fc2" "ga_out" "expanded0".."expandedN" "block_out" "masked"
 * dims = {B, L, C} (channels-last).  Returns TDANET_EINVAL for an unknown name. */
TDANET_API int tdanet_workspace_tensor(const tdanet_config_t* cfg, int batch, int n_samples, const char* name,
                            size_t* byte_offset, int64_t dims[3]);
/* latent lengths L[0..depth-1] for n_samples input samples, and padded length / rest */
TDANET_API int tdanet_latent_lengths(const tdanet_config_t* cfg, int n_samples, int32_t* lengths,
                          int32_t* padded_len, int32_t* rest);

/* ------------------------------------------------------------------ training step (every variant)
 * Replaces the autograd graph of AudioLightningModule.training_step (system/audio_litmodule.py:83-124) and
 * the Trainer's clip + optimiser (audio_train.py:71,187-197; configs/tdanet_lsr2.yml:42-45) for the model
 * path: forward keeping every GlobLN-delimited tensor of every UConvBlock iteration, hand-written backward,
 * global-norm clipping and Adam on flat buffers.
 *
 * Dropout / DropPath (cfg->dropout, cfg->drop_path > 0; drop_path / DropPath TDANet_best.py:7-30, nn.Dropout
 * sites :210,:212,:251, nn.MultiheadAttention(dropout) :241): tdanet_forward_train_rng draws every keep-mask of
 * every UConvBlock iteration up front with Philox4x32-10 (key = rng_state[0], counter = {element group, site |
 * iteration << 8, rng_state[1]}), stores them as bytes in the training workspace ("m_att" "m_ao" "m_f1" "m_f2"
 * "m_dp", elem_bytes 1) and then advances rng_state[1] on the device, so a captured CUDA graph draws fresh masks
 * at every replay; tdanet_backward reads the stored masks.  element kept  <=>  u32 >= floor(p * 2^32); kept
 * elements are scaled by 1/(1-p) like torch.  tdanet_forward_train is the p = 0 form (it fails if p > 0).
 *
 * tdanet_forward_train   same result as tdanet_forward; `workspace` (tdanet_train_workspace_bytes) then holds
 *                        what tdanet_backward needs and must stay untouched until it has run.
 * tdanet_backward        d_est [B, n_src, T] = d loss / d est.  `grads` has the layout of tdanet_weights_t and
 *                        points at the gradient buffer of every parameter (pe: NULL); gradients are ADDED to
 *                        those buffers (zero them first; the 16 iterations of the shared block accumulate).
 *                        Parameters the forward never reads (loc_glo_fus of the last scale) are left alone. */
TDANET_API int tdanet_train_workspace_bytes(const tdanet_config_t* cfg, int batch, int n_samples, size_t* bytes);
TDANET_API int tdanet_forward_train(const tdanet_config_t* cfg, const tdanet_weights_t* w, const float* wav,
                   int batch, int n_samples, float* est, void* workspace, size_t workspace_bytes,
                   tdanet_stream_t stream);
/* rng_state: device uint64[2] = {seed, offset}; offset is incremented by the call (on the stream). */
TDANET_API int tdanet_forward_train_rng(const tdanet_config_t* cfg, const tdanet_weights_t* w, const float* wav,
                   int batch, int n_samples, float* est, void* workspace, size_t workspace_bytes,
                   uint64_t* rng_state, tdanet_stream_t stream);
TDANET_API int tdanet_backward(const tdanet_config_t* cfg, const tdanet_weights_t* w, const tdanet_weights_t* grads,
                   const float* wav, const float* d_est, int batch, int n_samples, void* workspace,
                   size_t workspace_bytes, tdanet_stream_t stream);
/* Locate a tensor of UConvBlock iteration `block` in the training workspace (tests).  Names of
 * tdanet_workspace_tensor plus "bin" "y" "fused0".. "mlogit", the statistics "st_*" (elem_bytes = 8) and, with
 * dropout, the keep-masks (elem_bytes = 1): "m_att" [problems*heads, n, n] (problem = group*L + t), "m_ao" /
 * "m_f2" [B, L, C], "m_f1" [B, L, 2C], "m_dp" [2, B, 1] (attention branch, FFN branch). */
TDANET_API int tdanet_train_workspace_tensor(const tdanet_config_t* cfg, int batch, int n_samples, const char* name,
                   int block, size_t* byte_offset, int64_t dims[3], int32_t* elem_bytes);
/* Weight / bias gradient of a 1x1 Conv1d / Linear in channels-last form: dW[N, K] += G[rows, N]^T A[rows, K],
 * db[N] += column sums of G (db may be NULL).  gemm_mode fp32: CUDA-core fp32; tf32 / tf32x3: TF32 tensor cores
 * (mma.sync) when N and K are multiples of 4. */
TDANET_API int tdanet_wgrad(int gemm_mode, const float* G, const float* A, float* dW, float* db, int rows, int N,
                   int K, tdanet_stream_t stream);
/* sqnorm[0] = sum of squares of grads[0..n) (double, device; sqnorm has 2 elements). */
TDANET_API int tdanet_grad_sqnorm(const float* grads, size_t n, double* sqnorm, tdanet_stream_t stream);
/* clip_grad_norm_(max_grad_norm) (skipped if <= 0) then one torch.optim.Adam step (no weight decay, no amsgrad)
 * on flat buffers; gradients are first multiplied by grad_scale (1/world_size after a sum all-reduce).
 * `step` is a device counter of the steps taken so far; it is incremented. */
TDANET_API int tdanet_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, size_t n,
                   float lr, float beta1, float beta2, float eps, float max_grad_norm, float grad_scale,
                   const double* sqnorm, int32_t* step, tdanet_stream_t stream);

/* ------------------------------------------------------------------ standalone ops
 * D[b, r, :N] = A[b, r, :K] . W[:N, :K]^T (+ bias), rows r < rows_per_item of each of `batch`
 * items; optional per-item sum / sum-of-squares of D accumulated (atomically) into
 * stats[b*2 + {0,1}] (double).  The 1x1 Conv1d / Linear of proj_1x1, res_conv, in_proj,
 * out_proj, fc1, fc2, pw_conv in channels-last form. */
TDANET_API int tdanet_gemm(int gemm_mode, const float* A, const float* W, const float* bias, float* D,
                int batch, int rows_per_item, int N, int K, double* stats,
                void* workspace, size_t workspace_bytes, tdanet_stream_t stream);
TDANET_API size_t tdanet_gemm_workspace_bytes(int N, int K);

/* Replaces PITLossWrapper(PairwiseNegSDR(sdr_type), pit_from="pw_mtx", threshold_byloss)
 * (pit_wrapper.py:29-67, matrix.py:21-56) for n_src = 2 or 3, forward and backward in one call.
 *   est, tgt  [B, n_src, T]
 *   loss      [1]              mean over the kept items of the best-permutation loss
 *   pw        [B, n_src, n_src] pairwise matrix [b, est, tgt]      (may be NULL)
 *   perm      [B, n_src] int32 best permutation: perm[b, j] = estimate matched to target j
 *   grad_est  [B, n_src, T]     d loss / d est                       (may be NULL)
 * sdr_type: 0 snr, 1 sisdr, 2 sdsdr.  scratch: tdanet_pit_loss_scratch_bytes(B, n_src). */
TDANET_API size_t tdanet_pit_loss_scratch_bytes(int batch, int n_src);
TDANET_API int tdanet_pit_loss(const float* est, const float* tgt, int batch, int n_src, int n_samples,
                    int sdr_type, int threshold_byloss, float* loss, float* pw, int32_t* perm,
                    float* grad_est, void* scratch, size_t scratch_bytes, tdanet_stream_t stream);

/* Long-form stitching (audio_test_css.py:108-134).  est [n_streams, n_chunks, 2, seg_len]: every chunk
 * separated on its own (attn_group = 1).  Chunk k > 0 is appended from sample `overlap` on, with its two
 * sources exchanged when that maximises the cosine similarity of its head with the tail of chunk 0.
 *   swap [n_streams, n_chunks] int32 (out), out [n_streams, 2, out_len] with
 *   out_len <= seg_len + (n_chunks-1)*(seg_len-overlap)  (the caller drops the zero padding of the last chunk). */
TDANET_API int tdanet_css_stitch(const float* est, int n_streams, int n_chunks, int seg_len, int overlap,
                                 int out_len, int32_t* swap, float* out, tdanet_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* TDANET_B200_H */
