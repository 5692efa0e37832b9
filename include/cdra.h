/* libcdra — C ABI of the B200-native PPO-update hot path of Luca96/carla-driving-rl-agent.
 *
 * The reference has no native boundary (it is pure Python on TF2/Keras); the functions below are what
 * a drop-in replacement of its numeric layer binds (see INTEGRATION.md for the Python-side stub).
 * Each entry cites the reference code it replaces.  Conventions:
 *   - every function returns 0 on success or a negative CDRA_ERR_* code; cdra_last_error() returns a
 *     thread-local description of the last failure.  No exceptions cross the boundary.
 *   - all tensor arguments are raw DEVICE pointers owned by the caller (torch.Tensor.data_ptr());
 *     `stream` is a cudaStream_t passed as void*.  All work is enqueued asynchronously on it.
 *   - the library allocates nothing after cdra_plan_create; scratch memory is one caller-owned
 *     workspace of cdra_plan_workspace_bytes() bytes that must be zeroed once before first use and then
 *     belongs to the plan: the library keeps launch descriptors in it across calls (re-uploaded only when an
 *     arena pointer changes), so use ONE workspace per plan and do not overwrite or re-create it in between.
 *   - a plan is thread-compatible, not thread-safe: one plan per (process, device).
 */
#ifndef CDRA_H_
#define CDRA_H_
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define CDRA_OK 0
#define CDRA_ERR_BADARG (-1)
#define CDRA_ERR_SHAPE (-2)
#define CDRA_ERR_WORKSPACE (-3)
#define CDRA_ERR_CUDA (-4)
#define CDRA_ERR_NCCL (-5)

#define CDRA_DTYPE_F32 0   /* parity mode: fp32 activations */
#define CDRA_DTYPE_BF16 1  /* perf mode: bf16 activation storage, fp32 accumulate / statistics */

/* arenas (flat fp32 buffers owned by the caller) */
#define CDRA_ARENA_DYN_PARAMS 0   /* dynamics trainable (2,128,450 floats)  */
#define CDRA_ARENA_DYN_STATE 1    /* dynamics BN moving mean/var (16,564)   */
#define CDRA_ARENA_POL_PARAMS 2   /* policy head trainable (270,470)        */
#define CDRA_ARENA_POL_STATE 3    /* policy head BN moving stats (1,664)    */
#define CDRA_ARENA_VAL_PARAMS 4   /* value head trainable (269,828)         */
#define CDRA_ARENA_VAL_STATE 5    /* value head BN moving stats (1,664)     */

typedef struct cdra_config {
    int32_t batch;      /* B: samples per call on this device (SGD minibatch shard)             */
    int32_t height;     /* image H (90)                                                         */
    int32_t width;      /* image W (120)                                                        */
    int32_t dtype;      /* CDRA_DTYPE_*                                                         */
    int32_t image_u8;   /* 1: state_image is uint8 0..255 (scaled by 1/255 on load); 0: float32 */
} cdra_config;

typedef struct cdra_plan cdra_plan_t;

const char* cdra_last_error(void);
int cdra_version(void);

/* Builds the layer graph of core/networks.py:37-56 (dynamics_layers) + core/architectures.py:30-173
 * for a fixed (B, H, W, dtype).  Host-only; does not touch the GPU. */
int cdra_plan_create(const cdra_config* cfg, cdra_plan_t** out);
void cdra_plan_destroy(cdra_plan_t* plan);
size_t cdra_plan_workspace_bytes(const cdra_plan_t* plan);

/* Arena layout introspection (mirrors Keras `model.get_weights()` bookkeeping of core/networks.py:297-310). */
int64_t cdra_arena_size(const cdra_plan_t* plan, int arena);             /* floats */
int cdra_arena_num_tensors(const cdra_plan_t* plan, int arena);
int cdra_arena_tensor(const cdra_plan_t* plan, int arena, int index, char* name, int name_cap,
                      int64_t* offset, int32_t* ndim, int32_t dims[4]);
/* Location of a named intermediate tensor inside the workspace (parity tests read taps through it). */
int cdra_plan_tensor(const cdra_plan_t* plan, const char* name, int64_t* byte_offset, int32_t dims[4],
                     int32_t* elem_size);

/* Parity taps in bf16 perf mode: the tower stores padded, shuffled channel planes; this writes
 * the named tensor ("tower.s2.u3.pw1", "tower.s1.u2.out", "grad:<name>" ...) in the reference's logical layout
 * [4B][H][W][C] as fp32 into `out` (may be NULL to query dims only).  Test infrastructure, not on the hot path. */
int cdra_debug_export(cdra_plan_t* plan, const char* name, void* workspace, float* out, int32_t dims[4], void* stream);

/* Parity aid for the stem backward (max-pool backward + BatchNorm backward + weight gradient of the 3x3 s2 conv,
 * core/architectures.py:159-161): re-runs ONLY that part on the workspace left by a training forward + backward (stem
 * output, pool winners and the pool-output gradient are still there) and writes the stem's gradients into `grads`
 * (dynamics arena layout, zeroed first).  legacy != 0 selects the CUDA-core kernels, 0 the tensor-core band kernel
 * (uint8 frames, bf16 mode), so tests can compare both on identical inputs.  Test infrastructure. */
int cdra_debug_stem_backward(cdra_plan_t* plan, const float* params, const void* image, float* grads, void* workspace,
                             int legacy, void* stream);

/* CARLANetwork.dynamics_predict_train / dynamics_predict (core/networks.py:206-212) on
 * dynamics_layers (core/networks.py:37-56).  image [B,4,H,W,3] (u8 or f32), road [B,4,9],
 * vehicle [B,4,4], navigation [B,4,5] f32; out512 [B,512] f32.  training!=0 uses batch statistics and
 * updates the moving statistics in `state` exactly like 4 sequential Keras BN calls per layer. */
int cdra_dynamics_forward(cdra_plan_t* plan, const float* params, float* state, const void* image,
                          const float* road, const float* vehicle, const float* navigation, int training,
                          float* out512, void* workspace, void* stream);

/* tape.gradient(loss, dynamics.trainable_variables) (core/carla_agent.py:361-365,440-444): consumes
 * d loss / d out512 [B,512] and the activations cdra_dynamics_forward left in `workspace`; writes
 * (overwrites) the flat gradient arena `grads` (same layout as CDRA_ARENA_DYN_PARAMS). */
int cdra_dynamics_backward(cdra_plan_t* plan, const float* params, const void* image, const float* road,
                           const float* vehicle, const float* navigation, const float* d_out512,
                           float* grads, void* workspace, void* stream);

/* CARLAgent.policy_objective (core/carla_agent.py:394-428) on PolicyNetwork.call
 * (core/networks.py:96-137) + its backward.  x512 [B,512]; actions_eval [B,2] (decision D2: the action
 * the new policy's log-prob is evaluated at, clipped to [eps,1-eps] like _clip_actions :139-144);
 * logp_old [B,2]; adv [B]; true_speed/true_sim [B,1].  scalars_out (16 floats): 0 total, 1 loss_policy,
 * 2 loss_entropy(=coef*H), 3 loss_speed, 4 loss_similarity, 5 ratio mean, 6 log_prob mean, 7 entropy,
 * 8 speed mean, 9 similarity mean.  d_x512 [B,512]; grads = policy gradient arena (overwritten).
 * grad_scale multiplies every gradient (1/world_size under data parallelism).
 * actions_jac [B,2,2] (may be NULL): (d a / d alpha, d a / d beta) of each evaluated action when it is a reparameterised
 * sample of the NEW policy -- the reference's `log_prob(clip(sample))` is differentiated through the sample (TFP Beta is
 * FULLY_REPARAMETERIZED: x = g1 / (g1 + g2) with implicit gamma gradients [lib]); NULL treats the action as a constant
 * (the base class's stored-action PPO, rl/agents/ppo.py:324-325). */
/* The dense GEMM behind the GRU projections, the trunk and the control branches (Keras Dense / GRU kernels,
 * core/networks.py:24-66): C[M][N] (=|+=) opA(A) opB(B) (+ bias[n]) on fp32 row-major device matrices; ta: A is stored
 * [K][M]; tb: B is stored [N][K].  tensor_core = 0: fp32 CUDA-core kernel (parity mode); 1: TF32 mma.sync kernel (bf16
 * perf mode).  Exposed for the parity tests. */
int cdra_debug_gemm(int ta, int tb, const float* A, int lda, const float* B, int ldb, float* C, int ldc, const float* bias,
                    int M, int N, int K, int accumulate, int tensor_core, void* stream);

/* Self test of the tcgen05 / TMEM path used by the tower's weight-gradient kernel: C[Mw][Nw] (fp32) = X^T Y for row-major
 * bf16 device matrices X [rows][Mw], Y [rows][Nw]; both operands are staged MN-major in 128-byte-swizzled shared memory
 * and accumulated over 64-row tiles in tensor memory.  Mw in {128, 256}, Nw % 16 == 0, Nw <= 256, (Mw / 128) * Nw <= 512,
 * rows % 64 == 0.  Test infrastructure. */
int cdra_debug_umma_selftest(const void* X, const void* Y, float* C, int rows, int Mw, int Nw, void* stream);
/* Same for K-major operands (the forward / data-gradient product): C[Mw][Nw] = A B^T, A [Mw][Kw], B [Nw][Kw] row-major
 * bf16; Mw in {128, 256}, Nw % 16 == 0, Nw <= 256, (Mw / 128) * Nw <= 512, Kw % 64 == 0, Kw <= 256. */
int cdra_debug_umma_selftest_k(const void* A, const void* B, float* C, int Mw, int Nw, int Kw, void* stream);

/* Kernel-selection switches for A/B parity runs inside one process: key "tc" = 1 | 0 routes the tower's pointwise weight
 * gradient through the tcgen05 / TMEM kernel or the mma.sync kernel (default: tcgen05 unless CDRA_NO_TC is set); key
 * "fwd_tc" does the same for the forward of the plain-output pointwise layers (pw1 of every unit).
 * Returns CDRA_ERR_BADARG for an unknown key.  Test infrastructure. */
int cdra_debug_set(const char* key, int value);

/* Role timeline of the warp-specialised tcgen05 kernels (test / profiling aid): with CDRA_TIMELINE=1 in the environment, block 0
 * of every pwg_fwd_kernel / pw_bwd_fused_kernel / pw_fwd_tc_kernel launch records %globaltimer nanoseconds at its hand-offs (prologue done,
 * griddepcontrol.wait passed, first operand block staged, first accumulator complete, first / last tile stored, statistics
 * flushed, last-CTA finalisation).  Synchronises the device and copies the stamps of the LAST such launches:
 * out64[0..15] forward GEMM family, out64[16..31] fused backward, out64[32..47] pw_fwd_tc_kernel, out64[48..63] fused
 * backward, steady state (tile 20 of block 0, role by role); profiles/pwg_timeline_probe.py names the slots.
 * CDRA_TIMELINE_R=32|64 restricts the fused-backward stamps to launches of that row-tile size. */
int cdra_debug_timeline(uint64_t* out64);

int cdra_policy_head_loss_fwd_bwd(cdra_plan_t* plan, const float* params, float* state, const float* x512,
                                  const float* actions_eval, const float* actions_jac, const float* logp_old, const float* adv,
                                  const float* true_speed, const float* true_sim, float clip_ratio,
                                  float ent_coef, int training, float grad_scale, float* scalars_out,
                                  float* head_out, float* d_x512, float* grads, void* workspace, void* stream);

/* CARLAgent.value_objective (core/carla_agent.py:469-486) on CARLANetwork.value_branch/value_head
 * (core/networks.py:255-275) + backward.  returns_be [B,2] (base, exp).  scalars_out: 0 total, 1 loss_v,
 * 2 loss_speed, 3 loss_similarity, 4 speed mean, 5 similarity mean.  head_out [B,4] = (base, exp, speed, sim). */
int cdra_value_head_loss_fwd_bwd(cdra_plan_t* plan, const float* params, float* state, const float* x512,
                                 const float* returns_be, const float* true_speed, const float* true_sim,
                                 int training, float grad_scale, float* scalars_out, float* head_out,
                                 float* d_x512, float* grads, void* workspace, void* stream);

/* PPOMemory.end_trajectory + compute_returns + compute_advantages (rl/agents/ppo.py:692-727) with
 * utils.gae / rewards_to_go / discount_cumsum / decompose_number / tf_sp_norm
 * (rl/utils.py:57-84,140-151,344-349) for `bs` independent trajectories of length T.
 * rewards [bs][T]; values_be [bs][T][2]; last_value_be [bs][2] (zeros for terminal states);
 * outputs returns_be [bs][T][2], adv [bs][T] (sp-normalised * scale).  gamma / lambda_ are doubles because the
 * reference hands python floats to scipy.signal.lfilter, which filters in float64. */
int cdra_gae(const float* rewards, const float* values_be, const float* last_value_be, double gamma,
             double lambda_, float scale, int bs, int T, float* returns_be_out, float* adv_out, void* stream);

/* utils.clip_gradients (rl/utils.py:120-121; per-tensor tf.clip_by_norm) followed by Keras
 * Adam.apply_gradients (rl/agents/ppo.py:246-250,270-273; core/carla_agent.py:386-388) over a flat
 * arena of `total` floats.  tensor_offsets [n_tensors+1] (device, int64) delimits the tensors; clip_norm<=0 disables
 * clipping.  step is the 1-based Adam iteration.  grad_scale is applied to the gradient first. */
int cdra_clip_adam(float* params, const float* grads, float* m, float* v, const int64_t* tensor_offsets,
                   int n_tensors, int64_t total, float clip_norm, float lr, float beta1, float beta2, float eps,
                   int64_t step, float grad_scale, float* norms_out, void* stream);

/* The per-tensor gradient norms the reference logs after every SGD step (`[tf.norm(g) for g in grads]`,
 * rl/agents/ppo.py:209-210,223-224; core/carla_agent.py:382,461): ONE launch over the flat arena instead of one reduction
 * (and one host round trip) per tensor.  sq_norms_out [n_tensors] receives sum((grad_scale * g)^2) per tensor. */
int cdra_grad_norms(const float* grads, const int64_t* tensor_offsets, int n_tensors, int64_t total, float grad_scale,
                    float* sq_norms_out, void* stream);

/* utils.data_to_batches gather (rl/utils.py:365-393): dst[i] = src[index[i]] for rows of row_bytes. */
int cdra_gather_rows(const void* src, const int64_t* index, int64_t n, int64_t row_bytes, void* dst, void* stream);
/* The same for every tensor of a minibatch (state components, actions, advantages, ...) in ONE launch: srcs / dsts / row_bytes are
 * HOST arrays of n_tensors (<= 12) device pointers / row sizes; all tensors are gathered with the same index vector. */
int cdra_gather_rows_multi(const void* const* srcs, void* const* dsts, const int64_t* row_bytes, int n_tensors, const int64_t* index,
                           int64_t n, void* stream);

/* Data-parallel gradient exchange (nothing like it exists in the single-process reference): one NCCL sum
 * all-reduce of a flat fp32 gradient range per pass, enqueued on `stream` like every kernel of the library so that a whole
 * SGD step stays on one stream / is capturable in a CUDA graph.  libnccl is resolved at run time (dlopen; the copy PyTorch
 * already loaded is reused), so the library itself has no link-time dependency on it.
 *   cdra_comm_unique_id : rank 0 creates the 128-byte NCCL id; the caller broadcasts it to the other ranks by any means
 *   cdra_comm_create    : every rank, on its own device (cudaSetDevice done by the caller); collective
 *   cdra_allreduce_grads: in-place sum over ranks of grads[0 .. count) */
typedef struct cdra_comm cdra_comm_t;
int cdra_comm_unique_id(void* id_out_128_bytes);
int cdra_comm_create(const void* id_128_bytes, int world_size, int rank, cdra_comm_t** out);
void cdra_comm_destroy(cdra_comm_t* comm);
int cdra_allreduce_grads(cdra_comm_t* comm, float* grads, int64_t count, void* stream);

/* On-device image augmentation: the `augment_fn` closure of the reference (core/carla_agent.py:527-579;
 * rl/augmentations/augmentations.py:44-263, rl/augmentations/simclr.py:44-58) over a batch of frames.  The scalars TF
 * draws once per call are passed explicitly (the host mirror draws them); per-pixel draws come from a counter-based hash
 * of (seed, frame, pixel, stream) that oracle/augment.py restates bit for bit.
 *   image   [frames][H][W][3] uint8 (value / 255 is augmented) or fp32
 *   out     [frames][H][W][3] fp32
 *   scratch >= 3 * frames + 2 * ceil(frames / group) 32-bit words (channel means, per-sample min / max keys)
 *   dropout_mask [dropout_size^2] uint8 on the device (1 = keep), may be NULL when dropout_size == 0 */
typedef struct cdra_augment_params {
    uint32_t seed;                       /* stream of the per-pixel hash */
    int32_t jitter;                      /* colour jitter: brightness -> contrast -> saturation -> hue -> clip [0, 1] */
    float brightness, contrast, saturation, hue;    /* delta, factor, factor, delta (fraction of a turn) */
    int32_t blur_size;                   /* 0 | 3 | 5: depthwise SAME convolution with blur_kernel [size][size][3] */
    float blur_kernel[75];
    int32_t salt_pepper; float sp_amount;            /* select p = amount / 10, salt with p = 1/2 */
    int32_t gauss_noise; float gn_amount, gn_std;    /* select p = amount, add clip(N(0, std), 0, 1) */
    int32_t normalize, group; float eps;             /* (x - min) / (max - min + eps) over every `group` consecutive frames */
    int32_t cutout_size, cutout_cell;                /* zero grid cell `cutout_cell` of a size x size grid stretched over the frame */
    int32_t dropout_size;                            /* coarse dropout grid (0 = off) */
} cdra_augment_params;
int cdra_augment(const void* image, int image_u8, int64_t frames, int height, int width, const cdra_augment_params* params,
                 const uint8_t* dropout_mask, float* out, void* scratch, void* stream);

/* Launch accounting (bench.py's `gpu_launches`) and optional per-kernel CUDA-event timing on the launch
 * stream (bench.py's live roofline numbers).  Profiling serialises every launch; never leave it on
 * inside a timed region.  cdra_profile_report writes "name\tcount\ttotal_ms\talgorithmic_bytes\n" lines
 * and returns the number of bytes the full report needs. */
int64_t cdra_launch_count(void);
void cdra_profile_enable(int on);
void cdra_profile_reset(void);
int cdra_profile_report(char* buf, int cap);

#ifdef __cplusplus
}
#endif
#endif /* CDRA_H_ */
