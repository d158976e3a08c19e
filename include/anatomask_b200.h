/*
 * libanatomask_b200 — C ABI of the B200-native (sm_100a) kernels behind AnatoMask's SparK-style masked 3-D ConvNet
 * pre-training step.  This is the drop-in boundary: plain pointers, sizes and a CUDA stream; no torch types.
 *
 * The reference (ricklisz/AnatoMask) has no native layer of its own — every entry point below replaces a
 * PyTorch library call site of the reference's Python modules (P = nnunetv2/training/nnUNetTrainer/variants/
 * pretrain).  The reference-side binding is a ctypes stub; the shipped host mirror is
 * anatomask_b200/{encoder3D,STUNet_head,decoder3D,spark3D,AnatoMask}.py.
 *
 * Conventions
 *   - activations: bf16, channels-last (N, D, H, W, C), C % 8 == 0 (tensor-core paths: C % 16 == 0)
 *   - packed conv weights: bf16 [tap][rows][cols] with `cols` (the contraction channel) contiguous
 *   - `active`: uint8 (N, fd, fh, fw), 1 = visible patch; the patch edge at a given resolution is D / fd
 *   - `active_list` / `active_count`: device work-list of active patch ids (n*L + l) built by
 *     amb_build_active_list — no host synchronisation anywhere on the step path
 *   - every function returns 0 on success or a negative AMB_ERR_* code; amb_last_error() gives the message.
 *     Unsupported shapes fail loudly — there is no CPU fallback and no other-arch dispatch.
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); nothing synchronises the host.
 */
#ifndef ANATOMASK_B200_H
#define ANATOMASK_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AMB_VERSION 100

#define AMB_OK 0
#define AMB_ERR_ARG (-1)       /* bad argument / unsupported shape */
#define AMB_ERR_CUDA (-2)      /* CUDA runtime / driver error      */
#define AMB_ERR_UNSUPPORTED (-3)

/* activation codes for the fused normalise-apply kernels */
#define AMB_ACT_NONE 0
#define AMB_ACT_LRELU 1        /* LeakyReLU(0.01)  — P/STUNet_head.py:84,89 */
#define AMB_ACT_RELU6 2        /* ReLU6            — P/decoder3D.py:21      */

/* convolution operators (all stride/padding variants the path uses) */
#define AMB_OP_CONV 0          /* nn.Conv3d k∈{1,3}, stride∈{1,2}, pad k/2      — P/encoder3D.py:13, P/decoder3D.py:21-22, P/spark3D.py:82 */
#define AMB_OP_CONV_DGRAD 1    /* its input gradient                                                               */
#define AMB_OP_CONVT 2         /* nn.ConvTranspose3d k4 s2 p1                   — P/decoder3D.py:19                */
#define AMB_OP_CONVT_DGRAD 3   /* its input gradient                                                               */

#define AMB_IMPL_AUTO 0        /* tcgen05 implicit GEMM when the shape allows it, CUDA-core gather kernel otherwise */
#define AMB_IMPL_DIRECT 1      /* force the CUDA-core gather kernel (differential testing)                          */
#define AMB_IMPL_TCGEN05 2     /* force a tcgen05 kernel (halo-plane v2 when it applies, else per-tap v1)          */
#define AMB_IMPL_TCGEN05_V1 3  /* force the per-tap tcgen05 kernel (differential testing of v2)                    */

const char* amb_last_error(void);
int amb_version(void);
int amb_sm_arch(void);                 /* 100: the only architecture this library is built for */
long amb_launch_count(void);           /* kernels launched by this library since the last reset */
void amb_reset_launch_count(void);
const char* amb_last_conv_kernel(void); /* name of the kernel that served the last amb_conv / amb_conv_wgrad call */

/* ---- work-list: replaces `_get_active_ex_or_ii(...).nonzero()` (host sync) — P/encoder3D.py:7-10 ------------- */
int amb_build_active_list(const uint8_t* active, int n_patches, int* list, int* count, void* stream);

/* ---- layout / dtype conversion at the module boundary (reference tensors are NCDHW fp32) ---------------------- */
int amb_ncdhw_f32_to_ndhwc_bf16(const float* src, void* dst, int N, int C, int D, int H, int W, void* stream);
int amb_ndhwc_bf16_to_ncdhw_f32(const void* src, float* dst, int N, int C, int D, int H, int W, void* stream);

/* ---- weight packing: dst[t][a][b] (bf16) = src[t*st + a*sa + b*sb] (fp32); and the inverse for weight grads ---- */
int amb_pack_weight(const float* src, void* dst, int T, int A, int B, long st, long sa, long sb, void* stream);
int amb_unpack_wgrad(const float* src, float* dst, int T, int A, int B, long st, long sa, long sb, void* stream);
/* All conv weights of a step in ONE launch (the reference's nn.Conv3d / nn.ConvTranspose3d read their fp32 `weight`
 * parameters in every forward — P/encoder3D.py:13, P/decoder3D.py:18-22, P/spark3D.py:82; here the bf16
 * operand copies are refreshed once per step).  Job kinds (`b_fast`):
 *   1 / 0  stock parameter layout, taps innermost (st == 1): src[(a*B + b)*T + t] / src[(b*A + a)*T + t]; tiles are 4x64
 *          (kind 1) or 16x16 (a, b) positions x 32 taps, tiles_b / tchunks = tiles along b / tap chunks
 *   2      taps-major source in dst order, src[(t*A + a)*B + b] (the engine's parameter arena): plain fp32 -> bf16
 *          conversion, one block per 2048 elements (tiles_b / tchunks unused)
 *   3      taps-major source with a and b swapped, src[(t*B + b)*A + a] (the input-gradient operand of an arena weight):
 *          per-tap transpose, blocks of 32 a x 64 b, tiles_b / tchunks = tiles along b / along a
 * tile_begin = first block of the job.  The table lives in device memory. */
typedef struct amb_pack_job {
    const float* src;
    void* dst;
    int T, A, B, b_fast;
    int tile_begin, tiles_b, tchunks, pad_;
} amb_pack_job;
int amb_pack_weights_batched(const amb_pack_job* jobs_dev, int n_jobs, int total_tiles, void* stream);

/* ---- convolutions ---------------------------------------------------------------------------------------------
 * (N, D, H, W) are the dims of the FINE-resolution tensor of the operator: the input of CONV (output is /stride),
 * the OUTPUT of CONV_DGRAD, the input of CONVT (output is ×2) and the output of CONVT_DGRAD is (N,D,H,W) too.
 *   CONV        x (N,D,H,W,Cin)          → y (N,D/s,H/s,W/s,Cout)      w [k³][Cout][Cin]
 *   CONV_DGRAD  x=dy (N,D/s,..,Cout)     → y=dx (N,D,H,W,Cin)          w [k³][Cin][Cout]
 *   CONVT       x (N,D,H,W,Cin)          → y (N,2D,2H,2W,Cout)         w [64][Cout][Cin]
 *   CONVT_DGRAD x=dy (N,2D,2H,2W,Cout)   → y=dx (N,D,H,W,Cin)          w [64][Cin][Cout]
 * Optional epilogue: + bias[rows]; zero the outputs of inactive patches (`active`, mask grid fd×fh×fw at the
 * OUTPUT resolution); with `active_list` only the output tiles inside active patches are computed at all
 * (the output buffer must have been zeroed once — inactive voxels are never written).
 * `stats` (optional, double[2*rows], accumulated) receives Σy and Σy² per output channel over the written
 * (active) outputs — the producer-side half of SparseInstanceNorm / BatchNorm (P/encoder3D.py:149-158).       */
typedef struct amb_conv_args {
    int op, impl;
    int N, D, H, W;
    int Cin, Cout;              /* channels of x and y as named in the table above (contraction = channels of x) */
    int k, stride;
    const void* x;
    void* y;
    const void* w;
    const float* bias;
    const uint8_t* active;
    int fd, fh, fw;
    const int* active_list;
    const int* active_count;
    double* stats;
    /* optional fused output transform  y = act(acc · ep_scale[r] + bias[r])  applied before the mask: an inference-mode
     * BatchNorm (+ReLU6) folded into the producing convolution (scale = γ/√(σ²+eps), bias = β − μ·scale) — the teacher's
     * decoder in eval mode (P/pretrain_AntoMask.py:422, P/decoder3D.py:19-22).  ep_scale == NULL: scale 1.  `stats`, when
     * given together with it, sees the transformed values. */
    const float* ep_scale;
    int ep_act;                 /* AMB_ACT_* */
    int pad_;
    /* optional fp32 scratch, ZERO-INITIALISED by the caller, of amb_conv_workspace_bytes(a) bytes: layers whose output has
     * too few 128-voxel tiles to fill the GPU (the 8³ / 16³ stages: 16-64 tiles for 148 SMs) split their taps over several
     * CTAs, add fp32 partial sums into it and finish with one small pass (bias / mask / bf16 / Σ,Σ²).  NULL: no split. */
    void* workspace;
    long workspace_bytes;
    void* stream;
} amb_conv_args;
int amb_conv(const amb_conv_args* a);
long amb_conv_workspace_bytes(const amb_conv_args* a);   /* 0 when the call would not split (pointers in `a` are not read) */

/* weight gradient of CONV / CONVT: dw [taps][Cout][Cin] fp32 (accumulated into, caller zeroes), from the
 * operator's input x and output gradient dy, dims as in the table above.  With `active_list` only active
 * patches of dy (OUTPUT resolution) are visited (dy is zero elsewhere).                                      */
typedef struct amb_wgrad_args {
    int op, impl;               /* op: AMB_OP_CONV or AMB_OP_CONVT */
    int N, D, H, W;
    int Cin, Cout;
    int k, stride;
    const void* x;
    const void* dy;
    float* dw;
    int fd, fh, fw;
    const int* active_list;
    const int* active_count;
    void* stream;
} amb_wgrad_args;
int amb_conv_wgrad(const amb_wgrad_args* a);

/* ---- encoder stem (Cin = 1): masked input → conv1 (k3) and conv3 (k1) in one pass — P/STUNet_head.py:96-101 ---- */
int amb_stem_fwd(const float* inp, const uint8_t* active, const int* active_list, const int* active_count,
                 int N, int D, int H, int W, int fd, int fh, int fw, int C,
                 const float* w1, const float* b1, const float* w3, const float* b3,
                 void* out1, void* out3, void* stream);
/* weight/bias grads of the stem: dw1[C][27], db1[C], dw3[C], db3[C] (fp32, accumulated into) */
int amb_stem_wgrad(const float* inp, const uint8_t* active, const int* active_list, const int* active_count,
                   int N, int D, int H, int W, int fd, int fh, int fw, int C,
                   const void* dy1, const void* dy3, float* dw1, float* db1, float* dw3, float* db3, void* stream);

/* ---- reconstruction head: Conv3d(C → 1, k1, bias) — P/decoder3D.py:51,61 -------------------------------------- */
int amb_proj_fwd(const void* x, const float* w, const float* b, float* rec, long voxels, int C, void* stream);
int amb_proj_bwd(const void* x, const float* w, const float* drec, void* dx, float* dw, float* db,
                 long voxels, int C, void* stream);

/* ---- pooled masked norm / BatchNorm (one formula: P/encoder3D.py:17-25,149-158; nn.BatchNorm3d) ---------------
 * Tensor (N, D, H, W, C) with mask grid (fd,fh,fw); `active_list == NULL` → dense (all voxels).
 *   stats     : sums[2C] (double, accumulated) += Σx, Σx² over the visited voxels
 *   finalize  : mean/biased var from sums and n = visited voxels (count·P³ read on device, or N·D·H·W),
 *               scale = γ/√(var+eps), shift = β − mean·scale, saved[2C] = (mean, rstd);
 *               optional running stats update (momentum, unbiased var) and num_batches_tracked += 1
 *   eval      : scale/shift from running stats (teacher forward)
 *   apply     : out = act(scale·x + shift [+ residual]) on visited voxels; with `token` != NULL the tensor is
 *               densified instead: inactive voxels get token[c] (P/spark3D.py:119-122), no activation
 *   bwd_reduce: sums[2C] += Σg, Σg·x̂ with g = dout·act'(scale·x+shift[+res]); dtoken[C] += Σ_inactive dout
 *   bwd_apply : dx = scale·(g − Σg/n − x̂·Σ(g·x̂)/n); dres = g (optional); dgamma = Σg·x̂, dbeta = Σg (finalized)
 */
typedef struct amb_geo {
    int N, D, H, W, C;
    int fd, fh, fw;
    const uint8_t* active;
    const int* active_list;
    const int* active_count;
} amb_geo;
int amb_norm_stats(const amb_geo* g, const void* x, double* sums, void* stream);
int amb_norm_finalize(const amb_geo* g, const double* sums, const float* gamma, const float* beta, float eps,
                      float* scale, float* shift, float* saved, float* running_mean, float* running_var,
                      long* num_batches_tracked, float momentum, const double* n_total, void* stream);
int amb_norm_eval(const float* gamma, const float* beta, const float* running_mean, const float* running_var,
                  float eps, float* scale, float* shift, int C, void* stream);
int amb_norm_apply(const amb_geo* g, const void* x, const float* scale, const float* shift, const void* residual,
                   const float* token, int act, void* out, void* stream);
int amb_norm_bwd_reduce(const amb_geo* g, const void* dout, const void* x, const void* residual,
                        const float* scale, const float* shift, const float* saved, int act, int fill,
                        double* sums, double* dtoken, void* stream);
int amb_norm_bwd_apply(const amb_geo* g, const void* dout, const void* x, const void* residual,
                       const float* scale, const float* shift, const float* saved, const double* sums, int act,
                       int fill, void* dx, void* dres, float* dgamma, float* dbeta, const double* n_total,
                       void* stream);
/* n_total (optional, device): number of voxels the statistics in `sums` were pooled over when they were summed
 * across ranks (SparseSyncBatchNorm3d / nn.SyncBatchNorm, P/encoder3D.py:43, P/decoder3D.py:42-43); NULL = local.
 * amb_count_voxels writes the local count (as double) so that it can ride in the same all-reduce as the sums.   */
int amb_count_voxels(const amb_geo* g, double* out, void* stream);
/* plain elementwise add (decoder skip: x + to_dec[i], P/decoder3D.py:58): out = a + b, n bf16 elements */
int amb_add(const void* a, const void* b, void* out, long n, void* stream);
/* Engine-mode helpers for the masked encoder tensors (the `* mask` of P/encoder3D.py:12-15 without writing the zeros that
 * nothing reads):
 *   amb_zero_shell  : clears the 1-voxel shell around every visible patch where it lies inside masked patches — all a
 *                     3x3x3 consumer that walks the active-patch list can read beyond the visible voxels
 *   amb_add_parity0 : fine[n,2z,2y,2x,:] += coarse[n,z,y,x,:] — the input gradient of the residual block's 1x1 stride-2
 *                     shortcut conv (P/STUNet_head.py:79-80,96-101) accumulated in place; `g` describes the COARSE tensor  */
int amb_zero_shell(const amb_geo* g, void* x, void* stream);
int amb_add_parity0(const amb_geo* g, const void* coarse, void* fine, void* stream);

/* ---- patchify + per-patch (optionally mean/var-normalised) masked MSE — P/spark3D.py:130-138,
 *      P/AnatoMask.py:190-202, teacher score P/pretrain_AntoMask.py:423-425 ------------------------------------
 * inp, rec: fp32 (N,1,D,H,W); patch p=16.  per_patch[N*L] = mean_n (rec − tgt)² · nonactive;
 * loss[0] = Σ per_patch / (Σ nonactive + 1e-8) (written by the last block; no host sync).
 * bwd: drec = dloss · 2 (rec − tgt) / (p³ · (Σnonactive + 1e-8)) on masked patches, 0 on active ones.        */
int amb_patch_loss_fwd(const float* inp, const float* rec, const uint8_t* active, int N, int D, int H, int W,
                       int normalize, float* per_patch, float* loss, float* patch_stats, unsigned int* ticket,
                       void* stream);
int amb_patch_loss_bwd(const float* inp, const float* rec, const uint8_t* active, const float* patch_stats,
                       const float* dloss, int N, int D, int H, int W, float* drec, void* stream);

/* ---- AnatoMask hard-mask generation — P/AnatoMask.py:81-135 ----------------------------------------------------
 * loss_pred (B, L) fp32.  hard[b, :len_loss] = indices of the len_loss largest losses (ascending loss order,
 * i.e. exactly argsort(loss)[L-len_loss:]) — bit-exact in the selected set.
 * mask_out (B, L) uint8: len_keep visible patches drawn uniformly from the non-hard ones with a counter-based
 * device RNG (seed, offset) — "throughput mode"; parity mode replays numpy's shuffle on the host from `hard`.  */
int amb_hard_mask(const float* loss_pred, int B, int L, int len_loss, int len_keep, unsigned long long seed,
                  unsigned long long offset, const unsigned long long* offset_dev /* optional device step counter */,
                  int* hard, int* order /* optional (B,L): full ascending argsort */, uint8_t* mask_out, void* stream);

/* ---- flat-arena optimiser pieces: EMA teacher (timm ModelEma.update), grad-norm clip + AdamW ------------------- */
int amb_ema_update(float* ema, const float* model, long n, double decay, void* stream);   /* ema = ema*d + (1-d)*model, fp32 products rounded separately like torch */
int amb_sumsq(const float* g, long n, double* out, void* stream);                 /* out[0] += Σ g² */
int amb_adamw_step(float* p, const float* g, float* m, float* v, long n, double lr, double beta1, double beta2,
                   double eps, double weight_decay, int step, const double* gnorm_sq, double max_norm,
                   double gscale /* g is multiplied by this first: 1/world after a SUM all-reduce */, void* stream);

/* CUDA-graph-replayable optimiser tail: the per-step scalars are read from device memory,
 * hyper = {ema_decay, 1-ema_decay, lr, beta1, beta2, eps, weight_decay, 1-beta1^t, sqrt(1-beta2^t), max_norm, gscale}. */
int amb_step_dev(float* ema, const float* model, long n_ema, float* p, const float* g, float* m, float* v, long n_live,
                 const float* hyper, const double* gnorm_sq, int do_adamw, int do_ema, void* stream);

/* ---- remaining sparse-layer API (the MedNeXt / ConvNeXt heads) -------------------------------------
 * voxel norm: SparseGroupNorm / SparseConvNeXtLayerNorm — P/encoder3D.py:47-78,193-243.  The reference hands the visible
 * voxels to nn.GroupNorm / nn.LayerNorm as an (N_active, C) matrix, so every voxel is normalised on its own over each of
 * `groups` channel groups (LayerNorm: groups = 1; biased variance) followed by the per-channel affine.  With `active_list`
 * only visible voxels are visited (caller zero-fills `out`); with `active` alone masked voxels are written as zero.
 * Channels per group: 1, 2, 4 or a multiple of 8; C <= 1024.  bwd: dgamma / dbeta fp32[C] are ACCUMULATED into.          */
int amb_voxel_norm_fwd(const amb_geo* g, const void* x, const float* gamma, const float* beta, int groups, float eps,
                       void* out, void* stream);
int amb_voxel_norm_bwd(const amb_geo* g, const void* dout, const void* x, const float* gamma, int groups, float eps,
                       void* dx, float* dgamma, float* dbeta, void* stream);
/* SparseMaxPooling / SparseAvgPooling — P/encoder3D.py:31-36: nn.MaxPool3d (mode 0) / nn.AvgPool3d (mode 1) with cubic
 * kernel k, stride, pad (floor mode), then · mask at the OUTPUT resolution (`active` (N,fd,fh,fw), NULL = none).
 * bwd is the gather form (deterministic); max routes to the first maximum in scan order like torch.                       */
int amb_pool3d_fwd(const void* x, void* y, int N, int D, int H, int W, int C, int k, int stride, int pad, int mode,
                   int count_include_pad, int divisor_override, const uint8_t* active, int fd, int fh, int fw, void* stream);
int amb_pool3d_bwd(const void* x, const void* dy, void* dx, int N, int D, int H, int W, int C, int k, int stride, int pad,
                   int mode, int count_include_pad, int divisor_override, const uint8_t* active, int fd, int fh, int fw,
                   void* stream);
/* SparseAdaptiveAvgPooling(1) — P/encoder3D.py:181-190: mean[n][c] = Σ x·mask / (Σ mask + 1e-6)                          */
int amb_masked_mean_fwd(const amb_geo* g, const void* x, float* mean, void* stream);
int amb_masked_mean_bwd(const amb_geo* g, const float* dmean, void* dx, void* stream);
/* depthwise k³ convolution (groups = C; k in {3,5,7}, stride 1/2, pad k/2) — SparseConvNeXtBlock.dwconv P/encoder3D.py:259,
 * MedNeXtBlock.conv1 P/MedNeXt_head.py:255-262,339-346.  w = the layer's fp32 weight (C,1,k,k,k).
 *   op AMB_OP_CONV       x (N,D,H,W,C) → y (N,D/s,H/s,W/s,C) = (conv + bias) · mask(output resolution)
 *   op AMB_OP_CONV_DGRAD x = dy (output resolution) → y = dx (N,D,H,W,C)
 * wgrad: dw (C,1,k,k,k) fp32 accumulated into; outputs of masked patches are skipped (dy is zero there).                  */
int amb_dwconv3d(int op, const void* x, const float* w, const float* bias, void* y, int N, int D, int H, int W, int C,
                 int k, int stride, const uint8_t* active, int fd, int fh, int fw, void* stream);
int amb_dwconv3d_wgrad(const void* x, const void* dy, float* dw, int N, int D, int H, int W, int C, int k, int stride,
                       const uint8_t* active, int fd, int fh, int fw, void* stream);
/* exact GELU (nn.GELU()): dout == NULL → out = gelu(x); else out = dout · gelu'(x).  n bf16 elements, n % 8 == 0            */
int amb_gelu(const void* x, const void* dout, void* out, long n, void* stream);
/* ConvNeXt tail P/encoder3D.py:270-279: dout == NULL → out = inp + mask·γ_c·x; else out = mask·γ_c·dout (the branch
 * gradient; the skip gradient is dout itself) and dgamma[c] += Σ mask·dout·x.  gamma NULL = 1.                             */
int amb_layer_scale(const amb_geo* g, const void* inp, const void* x, const float* gamma, const void* dout, void* out,
                    float* dgamma, void* stream);

/* ---- device-side input pipeline ---------------------------------------------------------------------
 * Replaces, per sample, nnUNetDataLoader3D.generate_train_batch's crop + zero padding (N/training/dataloading/
 * data_loader_3d.py:22-49) and the transforms of P/pretrain_AntoMask.py:79-113 as the scripts configure them:
 * batchgenerators SpatialTransform (rotation + isotropic scale about the patch centre, order-3 spline = scipy.ndimage.
 * map_coordinates(order=3, mode='constant', cval=0), centre crop when nothing is drawn) and MirrorTransform.
 * A case volume is a single-channel fp32 array (sD, sH, sW) resident in HBM; (lb_z, lb_y, lb_x) is the bounding box's lower
 * corner in case coordinates (may be negative / reach beyond the case: zero padding).
 *   amb_aug_spline_prefilter  cubic B-spline coefficients of the (pD,pH,pW) initial patch (mirror boundary at the patch
 *                             edges, like scipy's spline_filter under mode='constant'); `scratch` is a second patch-sized buffer
 *   amb_aug_resample          out[o] = spline(coef, (o − (O−1)/2)·M + (P/2 − ½)) or cval outside [0, P−1]; matrix9 (host, row-major
 *                             M[i][j], batchgenerators' row-vector convention: rotation R = Rx·Ry·Rz times the scale);
 *                             mirror3 (host) flips the written axis
 *   amb_aug_crop_mirror       no rotation / scale drawn: out = case[lb + o] (zero padded), mirrored — `lb` here is the lower
 *                             corner of the FINAL patch (bbox corner + (P − O) / 2, batchgenerators center_crop_aug)            */
int amb_aug_spline_prefilter(const float* src, int sD, int sH, int sW, int lb_z, int lb_y, int lb_x, float* coef, float* scratch,
                             int pD, int pH, int pW, void* stream);
int amb_aug_resample(const float* coef, int pD, int pH, int pW, const double* matrix9, const int* mirror3, float cval, float* out,
                     int oD, int oH, int oW, void* stream);
int amb_aug_crop_mirror(const float* src, int sD, int sH, int sW, int lb_z, int lb_y, int lb_x, const int* mirror3, float* out,
                        int oD, int oH, int oW, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ANATOMASK_B200_H */
