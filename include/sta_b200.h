/*
 * sta_b200.h -- C ABI of libsta_b200.so, the B200 (sm_100a) implementation of the
 * ViSTA-SLAM "Symmetric Two-view Association" (STA) frontend forward pass.
 *
 * The library is the drop-in boundary underneath the reference's Python module
 * vista_slam/sta_model/sta_model.py::SymmetricTwoViewAssociation.  Every model-level
 * entry point below names the reference method it replaces (file:line in
 * zhangganlin/vista-slam).  Conventions:
 *   - plain C: pointers, sizes, opaque handle; no C++/torch types, no exceptions;
 *   - every function returns 0 on success, non-zero on failure; sta_last_error()
 *     then returns a thread-local human-readable message;
 *   - all "dev" pointers are CUDA device pointers on the current device, "host"
 *     pointers are host memory; `stream` is a cudaStream_t passed as void* (NULL =
 *     default stream); calls are asynchronous with respect to the host unless noted;
 *   - the caller owns all input/output buffers, the handle owns weights + workspace;
 *   - the CUDA kernels are the only implementation: there is no CPU fallback.
 */
#ifndef STA_B200_H_
#define STA_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct StaModel StaModel; /* opaque */

/* ---- library ---- */
const char* sta_last_error(void);
int sta_version(void);                 /* ABI version, currently 1 */
int sta_device_synchronize(void);      /* cudaDeviceSynchronize + error check */

/* ---- model lifetime (replaces STA() + load_state_dict + .to(cuda), slam.py:95-106) ---- */
int sta_create(StaModel** out);
/* Same with an explicit operand precision.  STA_PRECISION_BF16 (what sta_create uses, and what every benchmark
 * number is measured with): bf16 tensor-core operands, fp32 accumulation.  STA_PRECISION_X3: split-precision PARITY
 * mode -- every bf16 activation is carried as (hi | lo | hi) and every bf16 weight as (hi | hi | lo), so the same
 * tcgen05 kernels compute a_hi w_hi + a_lo w_hi + a_hi w_lo over K' = 3K (about 17 significant operand bits),
 * and attention runs in fp32 on the CUDA cores.  It exists to check the kernels against the fp32 reference at
 * north_star's tolerance (pointmaps 1e-3, pose 1e-4); ~3x the weight memory, several times slower. */
enum { STA_PRECISION_BF16 = 0, STA_PRECISION_X3 = 1 };
int sta_create_ex(StaModel** out, int precision);
void sta_destroy(StaModel* m);

/* Upload one state-dict tensor (fp32, contiguous, host or device memory) under its
 * reference state_dict name (e.g. "enc_blocks.0.attn.qkv.weight").
 * The library converts/packs it into its own bf16 / fp32 device layout.
 * Names the forward pass does not use (enc_norm.*, refinenet4.resConfUnit1.*, the
 * aliased scratch.layerN_rn.*) are accepted and ignored.  Returns non-zero for an
 * unknown name or a shape mismatch (strict=True semantics, sta_model.py:143). */
int sta_load_tensor(StaModel* m, const char* name, const float* data, const int64_t* shape, int ndim,
                    int data_on_device);
/* Number of tensors still missing before the model can run (0 = ready). */
int sta_missing_tensors(StaModel* m);
/* The packed weights live in ONE device arena.  For multi-GPU runs rank 0 loads the state dict,
 * every rank exposes its arena, the host broadcasts it (torch.distributed / ncclBroadcast over NVLink)
 * and the receiving ranks call sta_mark_all_loaded. */
int sta_weight_arena(StaModel* m, void** dev_ptr_out, int64_t* bytes_out);
int sta_mark_all_loaded(StaModel* m);

/* ---- model-level forward entry points ---- */

/* _encode_image(image, true_shape, normalize=False), sta_model.py:163-174.
 * img_dev: [B,3,H,W] fp32 (img_is_bf16 = 0) or bf16 (1), values in [-1,1].
 * feat_out_dev: [B, (H/16)*(W/16), 1024] fp32.  pos_out_dev: [B, N, 2] int64 (y, x) or NULL. */
int sta_encode(StaModel* m, const void* img_dev, int img_is_bf16, int B, int H, int W, float* feat_out_dev,
               int64_t* pos_out_dev, void* stream);

/* _decode_stereo(feat1, feat2, pos1, pos2), sta_model.py:177-244.
 * feat*_dev: [B,N,1024] fp32; pos*_dev: [B,N,2] int64 (values in [-1,1023]).
 * out1_dev / out2_dev: arrays (host memory) of 13 device pointers, each [B,N+1,768] fp32 or NULL to
 * skip that layer's output; entry 12 is LayerNorm-ed (dec_norm), as in the reference. */
int sta_decode(StaModel* m, const float* feat1_dev, const float* feat2_dev, const int64_t* pos1_dev,
               const int64_t* pos2_dev, int B, int N, float* const* out1_dev, float* const* out2_dev, void* stream);

/* head_pose_s(tok), heads/pose_head.py:109-119.  tok_dev: [B,768] fp32 (already dec_norm-ed).
 * pose_out_dev: [B,4,4] fp32, conf_out_dev: [B] fp32. */
int sta_head_pose(StaModel* m, const float* tok_dev, int B, float* pose_out_dev, float* conf_out_dev, void* stream);

/* head_pts([enc_feat] + [dec_k[:,1:]]_k, true_shape), sta_model.py:135-137 + utils/misc.py:48-76 +
 * heads/dpt_head.py:34-66.  Only the four hooked layers are read:
 *   enc_feat_dev [B,N,1024], dec6_dev, dec9_dev, dec12_dev [B,N,768] fp32 (pose token already dropped).
 * pts3d_out_dev: [B,H,W,3] fp32, conf_out_dev: [B,H,W] fp32. */
int sta_head_pts(StaModel* m, const float* enc_feat_dev, const float* dec6_dev, const float* dec9_dev,
                 const float* dec12_dev, int B, int H, int W, float* pts3d_out_dev, float* conf_out_dev,
                 void* stream);

/* forward(views) for ONE support view over a batch of B pairs (sta_model.py:247-291): 2 encodes +
 * symmetric decode + 2 DPT heads + 2 pose heads, fused fast path (no per-layer outputs).
 * img1 = main view, img2 = support view, each [B,3,H,W].  Outputs, index 0 = main view, 1 = support:
 * pts3d_out_dev [2,B,H,W,3], conf_out_dev [2,B,H,W], pose_out_dev [2,B,4,4], pose_conf_out_dev [2,B]. */
int sta_forward_pairs(StaModel* m, const void* img1_dev, const void* img2_dev, int img_is_bf16, int B, int H, int W,
                      float* pts3d_out_dev, float* conf_out_dev, float* pose_out_dev, float* pose_conf_out_dev,
                      void* stream);

/* Same as sta_forward_pairs but with HOST buffers (pinned or pageable): copies the images to the
 * device, runs the forward, copies the four outputs back and synchronises the stream.
 * This is the end-to-end entry point bench.py times as "e2e". */
int sta_forward_pairs_host(StaModel* m, const void* img1_host, const void* img2_host, int img_is_bf16, int B, int H,
                           int W, float* pts3d_out_host, float* conf_out_host, float* pose_out_host,
                           float* pose_conf_out_host, void* stream);

/* Counters: number of kernel launches issued by this library on behalf of `m` since creation. */
int64_t sta_launch_count(StaModel* m);
/* Bytes of device memory currently held (weights + workspace). */
int64_t sta_device_bytes(StaModel* m);

/* Per-kernel-family device timing with CUDA events on the launch stream (adds two event records per
 * launch while enabled).  Families: 0 = tcgen05 GEMM (linear), 1 = tcgen05 GEMM (implicit 3x3 conv),
 * 2 = attention, 3 = LayerNorm.  sta_profile_read synchronises, returns the accumulated milliseconds,
 * launch counts and executed FLOPs per family since the last read, and resets the counters. */
int sta_profile(StaModel* m, int enable);
int sta_profile_read(StaModel* m, double* ms4, int64_t* counts4, double* flops4);

/* ---- op-level entry points (used by the parity tests; same kernels the model uses) ---- */

enum { STA_EPI_BF16 = 0, STA_EPI_GELU = 1, STA_EPI_F32 = 2, STA_EPI_ROPE = 3, STA_EPI_PIXSHUF = 4, STA_EPI_HEAD = 5 };

typedef struct StaGemmDesc {
  int conv3x3;          /* 0: A is [M][lda] bf16;  1: A is NHWC [nimg][H][W][Cin] bf16 (3x3, pad 1, stride 1) */
  int epi;              /* STA_EPI_* */
  const void* A;
  int64_t lda;
  const void* W;        /* [N][ldw] bf16; for conv3x3 K = 9*Cin ordered (kh, kw, cin) */
  int64_t ldw;
  int M, N, K;
  int nimg, H, Wd, Cin; /* conv3x3 geometry */
  const float* bias;    /* [N] fp32 or NULL */
  void* out;            /* bf16 (EPI_BF16/GELU/ROPE/PIXSHUF) or fp32 (EPI_F32) */
  int64_t ldo;
  void* out2;           /* EPI_BF16: optional relu copy */
  const void* resid;    /* EPI_BF16: bf16 [.][ldo];  EPI_F32: fp32 [.][ldo] */
  const void* resid2;   /* EPI_BF16 only */
  int relu_main;
  int rowmap_n;         /* EPI_F32: insert one skipped row before every n output rows */
  const int32_t* pos;   /* EPI_ROPE: [M][2] int32 (y, x) */
  int rope_cols;        /* EPI_ROPE: columns [0, rope_cols) are rotated */
  int ps_k, ps_cout, ps_h, ps_w; /* EPI_PIXSHUF */
  const float* head_w;  /* EPI_HEAD: [128][4] fp32 */
  const float* head_b;  /* EPI_HEAD: [4] */
  float* pts3d;         /* EPI_HEAD: [pixels][3] */
  float* conf;          /* EPI_HEAD: [pixels] */
  void* splitk_ws;      /* optional fp32 scratch (16-byte aligned): lets small EPI_F32 problems split K across CTAs */
  int64_t splitk_ws_bytes;
  int split_precision;  /* 1: operands are already expanded along K as A (hi|lo|hi) x W (hi|hi|lo) (K = 3x logical);
                           bf16 outputs / skip tensors use (hi|lo|hi) rows of logical width N (ldo >= 3N) */
} StaGemmDesc;

int sta_op_gemm(const StaGemmDesc* d, void* stream);

/* softmax(q k^T * scale) v, head_dim 64.  q/k/v: [batch][n*][ld*] bf16 with head h at columns
 * *_col0 + 64*h; out: [batch][nq][ldo] bf16.  kv sample for query sample b is (b + kv_batch_shift) % batch.
 * split_first_row != 0: query row 0 (the decoder's pose token) is computed by a small SIMT kernel and the
 * tensor-core kernel tiles rows [1, nq) -- same result, avoids a 128-row tile for one row when nq = 128k + 1. */
int sta_op_attention(const void* q, int64_t ldq, int q_col0, const void* k, int64_t ldk, int k_col0, const void* v,
                     int64_t ldv, int v_col0, void* out, int64_t ldo, int batch, int heads, int nq, int nk,
                     int kv_batch_shift, float scale, int split_first_row, void* stream);

int sta_op_layernorm(const float* x, int rows, int C, float eps, const float* g1, const float* b1, void* out1_bf16,
                     const float* g2, const float* b2, void* out2_bf16, int drop_first_of, void* stream);
int sta_op_patch_im2col(const void* img, int img_is_bf16, int B, int H, int W, void* out_bf16, void* stream);
int sta_op_upsample2x(const void* in_bf16, void* out_bf16, int nimg, int H, int W, int C, void* stream);
int sta_op_im2col_3x3_s2(const void* in_bf16, void* out_bf16, int nimg, int H, int W, int C, void* stream);
int sta_op_cast_f32_bf16(const float* in, void* out_bf16, int64_t rows, int C, int drop_first_of, void* stream);
/* In-place 2-D RoPE on tokens [B][N][H][64] bf16 with int64 positions [B][N][2]: the contract of
 * curope.rope_2d (pos_embed/curope/curope.cpp:49-65, kernels.cu:84-108), base 100, F0 = 1. */
int sta_op_rope2d(void* tokens_bf16, const int64_t* pos, int B, int N, int H, void* stream);


/* ---- keyframe step: OnlineSLAM.regress_two_views (vista_slam/slam.py:153-189) for K edges
 * (i, j_k) at once, from cached encoder features: symmetric decoder on the K feature pairs, head_pose_s on both
 * pose tokens, head_pts on both views, then estimate_intrinsic_from_pts3d(shared_intrinsic=True) per edge,
 * depths = pts3d[..., 2] and conf.mean().  Token positions are the regular (y, x) grid of PositionGetter
 * (sta_blocks.py:235-247), which is what slam.py:145 caches.  All outputs are [2][K]...: block 0 = the (i -> j)
 * direction ("ij"), block 1 = "ji".  intri_out [K][3][3], depth_out, conf_mean_out [2][K] may be NULL; scratch as for
 * sta_pointmap_consumers with V = 2K (only needed when intri_out / depth_out / conf_mean_out is given).
 * No host synchronisation: the caller decides when to read rel_pose_conf (slam.py:169). ---- */
int sta_regress_pairs(StaModel* m, const float* feat_i_dev, const float* feat_j_dev, int K, int H, int W,
                      float* pose_out_dev, float* pose_conf_out_dev, float* pts3d_out_dev, float* conf_out_dev,
                      float* intri_out_dev, float* depth_out_dev, float* conf_mean_out_dev, void* scratch, void* stream);

/* The same step in two phases, keeping the reference's early-out (slam.py:169-170: `rel_pose_conf < thres and i - j != 1`
 * skips both DPT heads): _begin runs the symmetric decoder and the pose heads for all K candidate edges
 * (pose_out [2][K][4][4], pose_conf_out [2][K]; block 0 = the i -> j direction) and keeps the decoder hooks in the
 * handle's workspace; the caller reads the K confidences (ONE device-to-host copy per keyframe instead of one per edge),
 * and _finish runs the DPT heads + pointmap consumers only for the n_sel surviving edges edge_idx_host[0..n_sel)
 * (indices into the K candidates; host memory).  Outputs of _finish are compact: pts3d_out [2][n_sel][H][W][3],
 * conf_out [2][n_sel][H][W], intri_out [n_sel][3][3], depth_out / conf_mean_out optional, scratch for V = 2 n_sel.
 * No other call on this handle may come between the two phases. */
int sta_regress_pairs_begin(StaModel* m, const float* feat_i_dev, const float* feat_j_dev, int K, int H, int W,
                            float* pose_out_dev, float* pose_conf_out_dev, void* stream);
int sta_regress_pairs_finish(StaModel* m, const int* edge_idx_host, int n_sel, float* pts3d_out_dev, float* conf_out_dev,
                             float* intri_out_dev, float* depth_out_dev, float* conf_mean_out_dev, void* scratch,
                             void* stream);

/* Launch-bound sizes (<= 8192 tokens per call: one keyframe, a few edges) of sta_encode and sta_regress_pairs are
 * captured into CUDA graphs per (entry point, batch, H, W) on their second use and replayed afterwards
 * (STA_CUDA_GRAPHS=0 disables).  Number of graph replays so far (tests / diagnostics): */
int64_t sta_graph_replays(StaModel* m);

/* ---- pointmap consumers: what OnlineSLAM.regress_two_views / connect_view_i_j compute from
 * the head outputs right after the boundary.  Device pointers, fp32; `scratch` is a caller-owned device buffer of
 * sta_pointmap_scratch_bytes(V) bytes (8-byte aligned).  No host synchronisation. ---- */
size_t sta_pointmap_scratch_bytes(int V);
/* estimate_intrinsic_from_pts3d (vista_slam/utils/slam_utils.py:8-79, call site slam.py:184) fused with
 * depths = pts3d[..., 2] (slam.py:185) and conf.mean() per view (pose_graph.py:41).
 * pts3d [V][H][W][3], conf [V][H][W]; shared = 0 -> K_out [V][3][3]; 1 -> one K_out [3][3] over all views;
 * 2 -> K_out [V/2][3][3], edge e from its two views (e, e + V/2), i.e. shared_intrinsic=True per (ij, ji) pair;
 * depth_out [V][H][W] and conf_mean_out [V] may be NULL. */
int sta_pointmap_consumers(const float* pts3d, const float* conf, int V, int H, int W, int shared, float* K_out,
                           float* depth_out, float* conf_mean_out, void* scratch, void* stream);
/* estimate_scale_with_depth_and_confidence (slam_utils.py:168-190, call site slam.py:224) and
 * scale_conf = (ci * cj).sqrt().mean() (slam.py:227): out2 = {scale, scale_conf}; n elements per map. */
int sta_depth_scale(const float* Di, const float* Dj, const float* ci, const float* cj, int64_t n, float* out2,
                    void* scratch, void* stream);


/* ---- Sim(3) pose-graph Levenberg-Marquardt step: one iteration of the optimisation in
 * OnlineSLAM.pose_graph_optimize (vista_slam/slam.py:108-140) over PoseGraphOpt (vista_slam/pose_graph.py:70-154):
 * residuals r_e = Log(T_e X_i^-1 X_j) of the edges with at least one optimised endpoint, their Jacobians w.r.t. left
 * perturbations X <- Exp(delta) X (what PyPose's LieTensor parameters use), A = J^T W J with W = diag(weights_e) and the LM
 * damping of pp.optim.LM.step (diagonal clamped to [dmin, dmax], then scaled by 1 + damping), a dense Cholesky solve of
 * A d = -J^T W r, and the update.  Device pointers: nodes / nodes_out [num_nodes][8] and meas [num_edges][8] fp32 Sim3 in
 * PyPose's layout (tx ty tz | qx qy qz qw | s), edges [num_edges][2] int64 (i, j), weights [num_edges][7] fp32,
 * opt_idx [num_opt] int64 (the optimised nodes, all others stay fixed).  info_out (device) [4] fp64 = {loss before,
 * loss after, |delta|_2, Cholesky ok (1/0)}; loss = sum r^T W r.  fp64 inside, deterministic, no host synchronisation: the
 * caller reads info_out to accept or reject the step (as pp.optim.LM does).  scratch: sta_pose_graph_scratch_bytes. ---- */
size_t sta_pose_graph_scratch_bytes(int num_nodes, int num_edges, int num_opt);
int sta_pose_graph_lm_step(const float* nodes, int num_nodes, const int64_t* edges, const float* meas, const float* weights,
                           int num_edges, const int64_t* opt_idx, int num_opt, double damping, double dmin, double dmax,
                           float* nodes_out, double* info_out, void* scratch, void* stream);

/* ---- image preprocessing: SLAM_image_only.process_image
 * (vista_slam/datasets/slam_images_only.py:22-34 -> datasets/base/base_view_graph_dataset.py:171-225 ->
 * utils/cropping.py:54-84,102-118 with PIL LANCZOS -> utils/image.py:13) on the device, bit-exact with PIL's 8-bit
 * resampler and torchvision's ToTensor / Normalize(0.5, 0.5) / Grayscale.  rgb [H][W][3] uint8 (RGB order) on the
 * device; res_w >= res_h is the dataset `resolution` (portrait frames get the transposed one, as in the reference).
 * Outputs: rgb_out [3][oh][ow] fp32 in [-1, 1] (the STA input), gray_out [oh][ow] fp32 in [0, 1] or NULL,
 * u8_out [oh][ow][3] (the resized crop itself) or NULL.  sta_preprocess_shape returns (oh, ow) without touching
 * the GPU.  Coefficient tables and the intermediate image are cached per (device, frame geometry); calls for one
 * geometry must not overlap on different streams. ---- */
int sta_preprocess_shape(int H, int W, int res_w, int res_h, int w_edge, int h_edge, int* out_hw);
/* Host-only test hooks (no GPU): the crop/resize geometry {crop l, t, r, b, resized w, h, final l, t, out w, h} and
 * the fixed-point Lanczos windows of one axis (Resample.c precompute_coeffs + normalize_coeffs_8bpc): *ksize taps per
 * output sample, bounds [out_size][2] = (first input sample, tap count), kk [out_size][*ksize] int32 (22-bit). */
int sta_preprocess_geometry(int H, int W, int res_w, int res_h, int w_edge, int h_edge, int* out10);
int sta_preprocess_coeffs(int in_size, int out_size, int* ksize, int* bounds, int* kk, int64_t kk_capacity);
int sta_preprocess_rgb8(const uint8_t* rgb_dev, int H, int W, int res_w, int res_h, int w_edge, int h_edge,
                        float* rgb_out_dev, float* gray_out_dev, uint8_t* u8_out_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* STA_B200_H_ */
