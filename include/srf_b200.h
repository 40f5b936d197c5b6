/*
 * srf_b200.h -- C-ABI of the B200-native SRF capsule-routing hot path.
 *
 * The reference (sephiroce/srf) has no FFI / plugin layer: its boundary is the
 * Keras model class tfsr/model/sequence_router_naive.py:33 (SequenceRouter), whose
 * `call` (naive:120-193) executes the routing stack as stock TF ops.  The entry
 * points below are what a ctypes binding for that path binds instead; each one
 * cites the reference lines it replaces.  Plain C types only, no exceptions cross
 * the boundary, every pointer that is not marked "host" is a DEVICE pointer to
 * contiguous row-major float32 owned by the caller's tensor framework and only
 * borrowed for the duration of the call (DLPack / data_ptr interchange).
 *
 * Return convention: 0 = ok; negative = invalid argument (shape/dtype/config);
 * positive = CUDA / NCCL error code.  srf_last_error() returns the message.
 * All compute entry points are asynchronous on the given cudaStream_t (passed
 * as void*; NULL = legacy default stream).  One handle per device; a handle may
 * be used from one host thread at a time.
 */
#ifndef SRF_B200_H_
#define SRF_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SRF_B200_VERSION 210 /* major*10000 + minor*100 + patch */

typedef struct srf_handle srf_handle;

/* u_hat arithmetic: how the prediction vectors W.x+b are formed (north_star knob). */
enum {
  SRF_UHAT_FP32 = 0, /* FP32 FFMA on CUDA cores, fp32 weights                     */
  SRF_UHAT_TF32 = 1, /* tensor cores, TF32 operands, fp32 accumulate (tcgen05); fused kernel: u_hat stays in TMEM */
  SRF_UHAT_BF16 = 2, /* tensor cores, TF32 operands, fp32 accumulate, u_hat STORED as bf16 in HBM (two-kernel path) */
  SRF_UHAT_FP32X3 = 3, /* tensor cores, 3 x TF32 split (W_hi x_hi + W_lo x_hi + W_hi x_lo): fp32-class
                         u_hat (rel. error ~1e-6), fp32 storage -- the 1e-4 parity class on tcgen05 */
  SRF_UHAT_F16 = 4    /* tensor cores, FP16 operand images (kind::f16), fp32 accumulate, fused kernel.  FP16
                         carries the same 11-bit significand as TF32, so the rounding class is that of
                         SRF_UHAT_TF32, with 2/3 (d = 20) to 1/2 (d = 8, 16) of the operand bytes streamed
                         per time step; operands are clamped to the finite FP16 range (|x|, |W| <= 65504;
                         magnitudes below 6.1e-5 keep an absolute error <= 3e-8).  Shapes the fused kernel
                         does not take run as SRF_UHAT_TF32 */
};

/*
 * One routing layer = naive:145-191 for one value of the layer index i:
 *   window gather (naive:150-151) -> u_hat = W.x + bias (naive:154-159) ->
 *   SDR scan (naive:162-170, body_context :232-245, pad_body_context :213-229) or
 *   DR (naive:171-185, _loop_body :200-206) -> LayerNorm over O*D + dropout
 *   (naive:188-191) -> optionally the head ln_o(length(.)) (naive:193).
 *
 * Layouts (the reference's einsum layout is canonical, einsum:82-97):
 *   emb          [B,S,H,d]     input capsules
 *   W            [I,O,D,d]     I = (lpad+rpad+1)*H, input capsule index i = w*H + h
 *   bias         [I,O,D]
 *   ln_gamma/beta[O*D]         ln_mid%d; NULL,NULL = skip the LayerNorm (raw capsules out)
 *   dropout_mask [B,S,O,D]     already scaled keep mask (0 or 1/(1-rate)); NULL = inference
 *   head_gamma/beta [O]        ln_output; non-NULL = also emit logits
 *   out_caps     [B,S,O,D]     layer output (after LN and dropout); may be NULL in a
 *                              stack call (library workspace is used)
 *   out_logits   [B,S,O]       required iff head_gamma != NULL
 *   out_raw      [B,S,O,D]     optional: the squashed capsules BEFORE LayerNorm/dropout
 *                              (what the backward pass needs saved); NULL = not written
 */
typedef struct srf_layer_desc {
  const float* emb;
  const float* W;
  const float* bias;
  const float* ln_gamma;
  const float* ln_beta;
  const float* dropout_mask;
  const float* head_gamma;
  const float* head_beta;
  float* out_caps;
  float* out_logits;
  float* out_raw;
  int32_t B, S;        /* utterances, routing frames per utterance                    */
  int32_t H, d;        /* input capsules per frame, input capsule dim                 */
  int32_t O, D;        /* output capsules, output capsule dim                         */
  int32_t lpad, rpad;  /* --model-caps-window-{lpad,rpad}                             */
  int32_t iters;       /* --model-caps-iter (ITER)                                    */
  int32_t sdr;         /* --model-caps-context: 1 = SDR (sequential), 0 = DR          */
  int32_t mask_class0; /* 1 on the last layer: -1e9 on output capsule 0 (naive:174,219) */
  int32_t uhat_mode;   /* SRF_UHAT_*                                                  */
  float ln_eps;        /* Keras LayerNormalization default 1e-3                       */
  float length_eps;    /* 1e-7 (naive:256); the einsum variant uses 1e-9 (einsum:238)  */
  uint64_t weights_version; /* != 0: packed weights cached in the handle under this tag
                               until W/bias pointers, shapes or the tag change;
                               0 = repack on every call                              */
} srf_layer_desc;

/* library version (SRF_B200_VERSION of the built .so) */
int srf_version(void);

/* create / destroy the per-device workspace handle (host pointers).  A handle owns scratch buffers that are
 * reused, regrown and freed in stream order: calls may come on any stream -- a call on another stream than the
 * previous one first waits (event) for everything the handle enqueued there, so streams sharing a handle
 * serialise -- but one host thread at a time; concurrent host threads need one handle each.  The fused inference
 * path of a multi-layer stack reserves the device's persisting-L2 carve-out for the packed weights and every
 * other path gives it back (SRF_L2_WINDOW=0 in the environment of srf_create: never). */
int srf_create(int device, srf_handle** out);
int srf_destroy(srf_handle* h);

/* last error message of this handle (host string, valid until the next call); h may be
 * NULL for errors of srf_create */
const char* srf_last_error(const srf_handle* h);

/*
 * Kernel selection.  uhat_mode TF32 / F16 / FP32X3 run the FUSED routing kernel (routing_fused.cu: u_hat is
 * produced by tcgen05.mma into TMEM and consumed there, an SDR stack is ONE persistent launch over all
 * layers) when the shape is supported (d % 4 == 0, D <= 20, O <= 64, 16-byte aligned emb) and the
 * library's policy expects it to be faster than the two-kernel path (materialised u_hat); BF16 always
 * runs the two-kernel path, FP32 the CUDA-core kernel.  Results differ between the paths only within
 * the mode's rounding class.  Environment (read by srf_create): SRF_NO_FUSED=1 / SRF_FORCE_FUSED=1.
 * Developer switches of the fused path, read at every call (none changes a result beyond the summation
 * grouping of the plan): SRF_FUSED_CAPSTAGE=0/1 (tile-granular W ring with two MMA issuers / capsule-sized
 * stages with one issuer), SRF_FUSED_NO_ALIGN=1 (greedy CTA plan without the TPC alignment),
 * SRF_FUSED_PLAN=c0:c1:... (CTAs per layer of an SDR stack).
 * The fused kernel's waits are bounded: if one ever times out the launch drains and the NEXT call on
 * the handle returns 700 + wait-site code ("a fused routing launch timed out ...").
 */

/* one routing layer, forward.  Replaces naive:145-191 (+193 when the head is requested). */
int srf_route_layer_fwd(srf_handle* h, const srf_layer_desc* layer, void* stream);

/*
 * the whole routing stack, forward: `for i in range(self.enc_num)` of naive:145-191
 * plus the head naive:193.  layers[0].emb is the primary-capsule tensor; for n > 0 a NULL
 * layers[n].emb means "the previous layer's output"; NULL out_caps = library workspace.
 * The B,S of all layers must agree and layers[n].H,d must equal layers[n-1].O,D.
 */
int srf_route_stack_fwd(srf_handle* h, const srf_layer_desc* layers, int32_t n_layers,
                        void* stream);

/*
 * Backward of one routing layer (training mode of naive:145-193; the reference gets it from
 * tf.GradientTape through the tf.while_loop, tfsr/trainer_sr.py:62-71).  `layer` is the SAME
 * descriptor the forward call used (emb, W, bias, LayerNorm/head parameters, dropout mask,
 * knobs; its out_* pointers are ignored).  All gradient outputs ACCUMULATE (+=): zero them
 * before the first call.  Nothing but the layer input and `v_raw` is saved by the forward pass:
 * u_hat is recomputed -- in FP32 from the weights with uhat_mode FP32, by the tcgen05 GEMM
 * (streamed by the sweep, same rounding as the forward) with uhat_mode TF32 / BF16.
 *   v_raw        [B,S,O,D]  forward's out_raw of this layer
 *   d_out        [B,S,O,D]  dL/d(out_caps) from the next layer's d_emb; NULL on the last layer
 *   d_logits     [B,S,O]    dL/d(out_logits); NULL unless the layer carries the head
 *   d_raw        [B,S,O,D]  scratch: dL/d(v_raw) is left here
 *   dW [I,O,D,d], dbias [I,O,D], dgamma/dbeta [O*D] (NULL if no LayerNorm),
 *   dhead_gamma/dhead_beta [O] (NULL if no head), d_emb [B,S,H,d] (NULL = not needed)
 */
typedef struct srf_layer_grads {
  const float* v_raw;
  const float* d_out;
  const float* d_logits;
  float* d_raw;
  float* dW;
  float* dbias;
  float* dgamma;
  float* dbeta;
  float* dhead_gamma;
  float* dhead_beta;
  float* d_emb;
} srf_layer_grads;

int srf_route_layer_bwd(srf_handle* h, const srf_layer_desc* layer, const srf_layer_grads* grads,
                        void* stream);

/*
 * Backward of the whole routing stack (tape.gradient through `for i in range(self.enc_num)`,
 * tfsr/trainer_sr.py:62-71): srf_route_layer_bwd for n = n_layers-1 ... 0.  `layers` are the
 * descriptors of the training forward (srf_route_stack_fwd with out_caps AND out_raw given for every
 * layer, or one srf_route_layer_fwd per layer): a NULL layers[n].emb (n > 0) means layers[n-1].out_caps,
 * the saved input of layer n.  grads[n].v_raw is layers[n].out_raw of that forward; grads[last].d_logits
 * is dL/dlogits; for n < last a NULL grads[n].d_out means grads[n+1].d_emb (which must then be given, as
 * must every d_emb but grads[0]'s).  All gradient outputs accumulate: zero them first.
 */
int srf_route_stack_bwd(srf_handle* h, const srf_layer_desc* layers, const srf_layer_grads* grads,
                        int32_t n_layers, void* stream);

/*
 * Greedy CTC decode (the parity criterion of SURVEY.md 8c; blank = class_n - 1,
 * tfsr/trainer_sr.py:133-134): argmax per routing frame for s < lens[b], collapse repeats, drop
 * blank.  logits [B,S,C]; lens [B] int32 (routing frames); out_ids [B,S] int32 (first out_lens[b]
 * entries valid); out_lens [B] int32.
 */
int srf_ctc_greedy_decode(srf_handle* h, const float* logits, const int32_t* lens, int32_t B,
                          int32_t S, int32_t C, int32_t blank, int32_t* out_ids, int32_t* out_lens,
                          void* stream);

/*
 * CTC loss and its gradient (tf.nn.ctc_loss(labels, logits, label_length, logit_length,
 * logits_time_major=False, blank_index=blank), tfsr/trainer_sr.py:64-66).  logits [B,S,C] are
 * unnormalised (log-softmax is applied inside, as TF does); labels [B,Lmax] int32; in_lens,
 * lab_lens [B] int32.  loss [B] receives the per-utterance negative log-likelihood; d_logits
 * [B,S,C] receives grad_scale * dloss/dlogits (grad_scale = 1/global_batch reproduces
 * tf.nn.compute_average_loss, trainer_sr.py:67-68).  Utterances with no feasible alignment, or with a
 * label outside [0, C) among their lab_lens labels, get loss 0 and a zero gradient (tf.nn.ctc_loss returns
 * inf for the former and fails on the latter).  The per-class posterior sums are accumulated with
 * shared-memory float atomics: d_logits is reproducible to rounding, not bitwise.
 */
int srf_ctc_loss(srf_handle* h, const float* logits, const int32_t* labels, const int32_t* in_lens,
                 const int32_t* lab_lens, int32_t B, int32_t S, int32_t C, int32_t Lmax, int32_t blank,
                 float grad_scale, float* loss, float* d_logits, void* stream);

/*
 * One Adam update on a flat fp32 buffer with tf.keras.optimizers.Adam semantics
 * (tfsr/helper/train_helper.py:58-68; beta1 .9, beta2 .98, eps 1e-9 in egs/conf): step >= 1 is the
 * 1-based update count, lr the already scheduled learning rate (warm-up schedule
 * train_helper.py:32-56, see srf_b200.training.warmup_lr).
 */
int srf_adam_step(srf_handle* h, float* param, const float* grad, float* m, float* v, int64_t n,
                  float lr, float beta1, float beta2, float eps, int64_t step, void* stream);

/*
 * Capsulation front-end, forward (SURVEY.md 8f "next-1"): fbank features -> primary capsules, i.e.
 * everything SequenceRouter.call does BEFORE the routing stack (naive:129-142):
 *   CapsulationLayer (tfsr/model/sequence_router.py:44-82): 2 stages of { two Conv2D(3x3, stride 2,
 *     padding 'same') paths -> Dropout(0.2) each -> maximum -> feat_mask -> BatchNormalization(axis=-1,
 *     eps 1e-3) -> feat_mask }  (stage 0: Cin = 1, stage 1: Cin = C)
 *   -> reshape [B,S,Fq*C] -> Dense(PH) "flatten" (naive:131-132)
 *   [einsum variant only, pos_enc != 0: *= sqrt(PH); += get_pos_enc(S, PH)
 *     (sequence_router_einsum.py:130-131, tfsr/helper/model_helper.py:30-58)]
 *   -> two Conv2D(3x3, stride 1, 'same', 1 -> PD) "encaps" paths -> Dropout(0.2) -> maximum (naive:133)
 *   -> feat_mask(stride^2) (naive:134) -> squash over PD (naive:137, :248-253)
 *   -> LayerNormalization(PH*PD, eps 1e-3) "ln_input" (naive:139-141) -> Dropout(train_inp_dropout)
 *     (naive:142) -> emb [B,S,PH,PD], the layout srf_route_stack_fwd reads.
 * T1 = ceil(T/2), S = ceil(T1/2), F1 = ceil(F/2), Fq = ceil(F1/2) (TF 'same' padding: pad_total =
 * max((out-1)*2 + 3 - in, 0), pad_before = pad_total / 2).  feat_mask zeroes the frames
 * t >= ceil(length / stride^k) (model_helper.py:125-140).
 *
 * Kernels are TF layouts throughout: conv kernels [3,3,Cin,Cout], dense [Fq*C,PH].
 * cnn_kernel[p][s]: path p (0/1) of stage s (the reference indexes conv_layers[p][s],
 * sequence_router.py:76-77).
 *
 * training == 0: inference (Keras training=False): no dropout, BatchNormalization with the moving
 *   statistics bn_mean / bn_var.
 * training != 0: every non-NULL *_dropout tensor is an already scaled keep mask (0 or 1/(1-rate))
 *   multiplied in where the reference applies the Dropout layer; BatchNormalization uses the batch
 *   statistics over (B, time, freq) (biased variance) and updates bn_mean / bn_var IN PLACE with
 *   momentum bn_momentum (moving = moving*m + batch*(1-m)).  Deterministic (no atomics).
 * The backward of the front-end is not part of this library (tf.GradientTape / torch autograd
 * differentiate the host-framework copy); out_emb feeds srf_route_stack_fwd directly.
 */
typedef struct srf_frontend_desc {
  const float* feats;            /* [B,T,F] fbank features                                   */
  const int32_t* lengths;        /* [B] int32 DEVICE: valid fbank frames per utterance       */
  const float* cnn_kernel[2][2]; /* [path][stage] [3,3,Cin,C]                                */
  const float* cnn_bias[2][2];   /* [path][stage] [C]                                        */
  const float* bn_gamma[2];      /* [stage] [C]                                              */
  const float* bn_beta[2];
  float* bn_mean[2];             /* moving mean / variance (read; updated when training)     */
  float* bn_var[2];
  const float* dense_kernel;     /* [Fq*C,PH]                                                */
  const float* dense_bias;       /* [PH]                                                     */
  const float* encaps_kernel[2]; /* [path] [3,3,1,PD]                                        */
  const float* encaps_bias[2];   /* [path] [PD]                                              */
  const float* ln_gamma;         /* ln_input [PH*PD]                                         */
  const float* ln_beta;
  const float* cnn_dropout[2][2]; /* training: [path][stage] keep masks [B,T_s,F_s,C] or NULL */
  const float* encaps_dropout[2]; /* training: [path] [B,S,PH,PD] or NULL                     */
  const float* inp_dropout;       /* training: [B,S,PH,PD] or NULL                            */
  float* out_emb;                /* [B,S,PH,PD]                                              */
  int32_t B, T, F;               /* utterances, fbank frames, feature dim (123)              */
  int32_t C;                     /* --model-conv-filter-num                                  */
  int32_t PH, PD;                /* --model-caps-primary-num / -dim                          */
  int32_t training;
  int32_t pos_enc;               /* 1 = einsum variant's sqrt(PH) scale + positional encoding */
  float bn_eps, bn_momentum;     /* Keras defaults 1e-3, 0.99                                */
  float ln_eps, squash_eps;      /* 1e-3, 1e-7                                               */
} srf_frontend_desc;

int srf_capsulate_fwd(srf_handle* h, const srf_frontend_desc* fe, void* stream);

/*
 * prediction vectors alone: window gather + u_hat = W.x + bias (naive:150-159) for every frame,
 * written in the reference's [B,S,I,O,D] layout (naive:158), fp32.  Uses emb, W, bias, B,S,H,d,
 * O,D, lpad, rpad and uhat_mode of the descriptor: SRF_UHAT_TF32 / SRF_UHAT_BF16 / SRF_UHAT_FP32X3
 * run the tcgen05 GEMM (u_hat kept in fp32 / bf16 / fp32 before the copy-out).  Requires d % 4 == 0.
 */
int srf_uhat_fwd(srf_handle* h, const srf_layer_desc* layer, float* out_uhat, void* stream);

/*
 * per-kernel device timing (measurement aid for bench.py): between srf_profile_begin and
 * srf_profile_end every kernel the handle launches is bracketed by CUDA events on its launch
 * stream.  srf_profile_end synchronises those events and returns, per kernel kind
 * (0 = weight packing, 1 = u_hat GEMM, 2 = routing kernel), the summed milliseconds and the
 * number of launches in ms[3] / launches[3] (host arrays).
 */
int srf_profile_begin(srf_handle* h);
int srf_profile_end(srf_handle* h, float* ms, int32_t* launches);

/* number of kernels this library has launched through the handle since creation
 * (bench.py's gpu_launches claim is read from here) */
int64_t srf_launch_count(const srf_handle* h);

/* name of the u_hat / routing kernel variant the last layer call dispatched to (host string) */
const char* srf_last_kernel(const srf_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* SRF_B200_H_ */
