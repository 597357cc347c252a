/* libaffgw — C ABI of the B200-native (sm_100a) generator / discriminator hot path of AFFGanWriting.
 *
 * The reference has no FFI layer of its own (it is pure PyTorch); the interface these entry points
 * replace is the set of library calls made by the reference's Python classes.  Each function cites the reference
 * call site (paths relative to /root/reference/GAN_word/) whose arithmetic it takes over.
 *
 * Conventions
 *   - plain pointers + sizes, no torch types.  All pointers are DEVICE pointers unless stated otherwise.
 *   - activations are NHWC (channels contiguous); dtype codes: 0 = fp32, 1 = bf16.  Math is fp32 everywhere;
 *     bf16 tensors feed tcgen05 tensor-core MMAs with fp32 (TMEM) accumulation.
 *   - the caller owns every buffer, including workspaces; the library never allocates or retains pointers.
 *   - `stream` is a cudaStream_t passed as void*.  Calls are asynchronous on that stream.
 *   - return 0 on success, negative on failure; affgw_last_error() returns a thread-local message.
 *   - there is no CPU fallback: a missing device or a failed launch is an error.
 */
#ifndef AFFGW_H
#define AFFGW_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define AFFGW_VERSION 111

enum { AFFGW_DT_F32 = 0, AFFGW_DT_BF16 = 1 };
enum { AFFGW_ACT_NONE = 0, AFFGW_ACT_RELU = 1, AFFGW_ACT_LRELU = 2, AFFGW_ACT_TANH = 3 };
enum { AFFGW_PAD_ZERO = 0, AFFGW_PAD_REFLECT = 1, AFFGW_PAD_REPLICATE = 2 };
enum { AFFGW_ALGO_AUTO = 0, AFFGW_ALGO_SIMT = 1, AFFGW_ALGO_TCGEN05 = 2 };
/* packed-weight layouts of the two tcgen05 convolution kernels (affgw_conv_tc_layout tells which one a geometry uses) */
enum { AFFGW_WLAYOUT_IM2COL = 1, AFFGW_WLAYOUT_SHIFT = 2 };
/* 16-bit format of the tensor-core operand planes / packed weights of one convolution (both operands of an MMA share it) */
enum { AFFGW_FMT_BF16 = 0, AFFGW_FMT_F16 = 1 };

/* One convolution = pad -> conv -> (+bias) -> (+addend) -> activation, i.e. the conv part of Conv2dBlock.forward
 * (blocks.py:150-163) with the explicit pad module (blocks.py:113-121), nn.Upsample(scale_factor=2)
 * (modules_tro.py:594) and the activation_first LeakyReLU (blocks.py:151-153) folded into the operand gather. */
typedef struct affgw_conv_desc {
    int32_t N, H, W, Cin;        /* stored input [N,H,W,Cin] (before upsampling)                               */
    int32_t Cout, KH, KW;
    int32_t stride, pad;         /* pad applies to the (upsampled) input                                      */
    int32_t pad_mode;            /* AFFGW_PAD_*                                                               */
    int32_t upsample;            /* 1, or 2 = nearest x2 before padding                                       */
    int32_t Ho, Wo;              /* output extent                                                             */
    int32_t in_pitch, out_pitch; /* elements between consecutive pixels of x / y (>= Cin / Cout)              */
    int32_t pre_act;             /* activation applied to x while gathering (activation_first blocks)         */
    int32_t post_act;            /* activation applied to the result                                          */
    int32_t x_dtype, w_dtype, y_dtype;
    int32_t algo;                /* AFFGW_ALGO_*                                                              */
    /* --- tcgen05 route only (algo = AFFGW_ALGO_TCGEN05) ------------------------------------------------------
     * Operands are bf16 "operand planes" made by affgw_split_planes: [planes][pixels][c_store], c_store % 8 == 0,
     * plane 0 = bf16(v), plane 1 = bf16(v - plane0) (present when passes = 3).  x_dtype = w_dtype = BF16, pre_act is
     * applied by affgw_split_planes (keep it here only for the dgrad fold), in_pitch = c_store of the x planes,
     * out_pitch = pitch of y (forward) or c_store of the dY planes (dgrad / wgrad).                              */
    int32_t passes;              /* 1: a_hi*w_hi; 3: split-bf16 a_hi*w_hi + a_lo*w_hi + a_hi*w_lo                */
    int32_t grad_dtype;          /* dtype of dx (and of x when the fold needs the pre-activation derivative)   */
    int32_t stride_w;            /* column stride when it differs from `stride` (rows): Resnet18.py:43 uses
                                    stride=(2, 1); 0 = same as stride                                          */
    int32_t operand_fmt;         /* AFFGW_FMT_*: bf16 planes (8 significant bits per plane, fp32's range) or fp16 planes
                                    (11 bits per plane).  fp16 weights are packed x 2^8 and fp16
                                    dY planes x a per-tensor power of two (affgw_amax_scale); the kernels undo both exactly */
} affgw_conv_desc;

int affgw_version(void);
const char* affgw_last_error(void);
/* number of kernels this library has launched in the calling process (bench.py's gpu_launches) */
long long affgw_launch_count(void);
/* 1 if the device under the current context is sm_100 */
int affgw_device_ok(void);

/* ---- convolution (replaces nn.Conv2d / nn.Linear: blocks.py:148, vgg_tro_channel3_modi.py:47,
 *      modules_tro.py:222,252-259,272-282) ---------------------------------------------------------------------- */
/* OIHW fp32 parameter -> [Cout][KH][KW][cin_pad] (transpose_flip = 0) or the dgrad operand
 * [Cin][KH][KW][cout_pad] with mirrored taps (transpose_flip = 1), in out_dtype. */
int affgw_pack_weight(const float* w_oihw, void* out, int out_dtype, int Cout, int Cin, int KH, int KW, int i_pad,
                      int transpose_flip, void* stream);
/* same, into the 128B-swizzled shared-memory tile images the tcgen05 kernel bulk-copies (bf16; hi and lo tiles when
 * passes = 3).  i_pad = c_store of the operand planes the weight will meet (x planes, or dY planes for transpose_flip). */
int affgw_pack_weight_tc(const float* w_oihw, void* out, int Cout, int Cin, int KH, int KW, int i_pad, int transpose_flip,
                         int passes, int layout, void* stream);
long long affgw_pack_weight_tc_bytes(int Cout, int Cin, int KH, int KW, int i_pad, int transpose_flip, int passes, int layout);
/* AFFGW_WLAYOUT_* the forward (for_dgrad = 0) or input-gradient (1) convolution described by d runs on; 0 = not
 * supported.  Stride-1 convolutions take the shared-memory-window ("shifted") kernel, the rest the im2col kernel. */
int affgw_conv_tc_layout(const affgw_conv_desc* d, int for_dgrad);
/* Position space of a stride-1 convolution (AFFGW_WLAYOUT_SHIFT): the frame [N][Hp][Wp] of its padded / upsampled input,
 * q = (n*Hp + yp)*Wp + xp.  Operand planes on it are planar: [hi|lo plane][G channel groups][QA positions][8] bf16, with
 * `lead` zero positions in front of q = 0.  affgw_conv_pos_frames gives the frame of the x planes (fx) and of the dY
 * planes (fy: same positions, output channels); affgw_split_positions builds either:
 *   x planes : src = x [N][H][W][pitch], (Hs, Ws) = (H, W), upsample, (oy0, ox0) = (pad, pad), the conv's pad_mode and
 *              pre-activation (activation_first blocks, blocks.py:151-153)
 *   dY planes: src = dY [N][Ho][Wo][Cout], upsample 1, (oy0, ox0) = (0, 0), AFFGW_PAD_ZERO, no activation; `colsum`
 *              (optional, fp32 [Cout], zeroed by the caller) receives sum_m dY[m][c] = the bias gradient
 * and affgw_conv2d_fwd / _dgrad / _wgrad take those planes as their x / dy arguments when affgw_conv_tc_layout says SHIFT. */
typedef struct affgw_pos_frame {
    int32_t N, Hp, Wp, G, lead, reserved;
    int64_t QA;
} affgw_pos_frame;
/* fp16 operand route (decoder convolutions, DESIGN.md "precision"): same calls with an operand format.
 *   affgw_amax_scale          scale2[0] = 2^k with max|x| * 2^k in [2^13, 2^14), scale2[1] = 2^-k (device memory, no host sync);
 *                             workspace8 = 8 bytes of device scratch that must be ZERO on entry and is left zero (one
 *                             buffer per stream serves every call)
 *   affgw_split_positions_fmt planes = fp16(v * *scale_dev) (scale_dev may be NULL = 1); the column sum stays unscaled
 *   affgw_pack_weight_tc_fmt  fp16 tiles of w * 2^8
 *   affgw_conv2d_*_scaled     result multiplied by *inv_scale_dev (the dY planes' 2^-k) and, for fp16 weights, by 2^-8 */
int affgw_amax_scale(const float* x, long long n, float* scale2, void* workspace8, void* stream);
int affgw_split_positions_fmt(const void* src, int dtype, void* planes, const affgw_pos_frame* f, int Hs, int Ws, int C, int pitch,
                              int upsample, int oy0, int ox0, int pad_mode, int pre_act, int passes, float* colsum,
                              int operand_fmt, const float* scale_dev, void* stream);
int affgw_split_planes_fmt(const void* x, int x_dtype, void* planes, long long rows, int C, int pitch, int c_store, int passes,
                           int pre_act, int operand_fmt, const float* scale_dev, void* stream);
int affgw_pack_weight_tc_fmt(const float* w_oihw, void* out, int Cout, int Cin, int KH, int KW, int i_pad, int transpose_flip,
                             int passes, int layout, int operand_fmt, void* stream);
int affgw_conv2d_dgrad_scaled(const void* dy, const void* wt, const void* x, void* dx, void* workspace,
                              const affgw_conv_desc* d, const float* inv_scale_dev, void* stream);
int affgw_conv2d_wgrad_scaled(const void* x, const void* dy, float* dw, void* workspace, const affgw_conv_desc* d,
                              const float* inv_scale_dev, void* stream);
int affgw_conv_pos_frames(const affgw_conv_desc* d, affgw_pos_frame* fx, affgw_pos_frame* fy);
long long affgw_position_planes_bytes(const affgw_pos_frame* f, int passes);
int affgw_split_positions(const void* src, int dtype, void* planes, const affgw_pos_frame* f, int Hs, int Ws, int C, int pitch,
                          int upsample, int oy0, int ox0, int pad_mode, int pre_act, int passes, float* colsum, void* stream);
/* output-channel tile width (the BN template argument) of the kernel that runs the forward (which = 0), input-gradient (1)
 * or weight-gradient (2) convolution of d - lets a profiler name the kernel a call lands on */
int affgw_conv_tc_tile_n(const affgw_conv_desc* d, int which);
/* positions per CTA tile of the position-space forward (which = 0) / dgrad (which = 1) kernel that runs d: 128 x the third
 * template argument of conv_shift_tcgen05_kernel (profiling labels); 128 for every other kernel */
int affgw_conv_tc_tile_m(const affgw_conv_desc* d, int which);
/* enable (1) / disable (0) the shifted kernel, -1 = query only; returns the previous setting (A/B testing) */
int affgw_conv_tc_prefer_shift(int enable);
/* activation tensor [rows][pitch] (fp32 or bf16) -> operand planes [passes == 3 ? 2 : 1][rows][c_store] (bf16), with the
 * activation_first non-linearity (blocks.py:151-153) applied and channels C..c_store-1 zero-filled */
long long affgw_operand_planes_bytes(long long rows, int c_store, int passes);
int affgw_split_planes(const void* x, int x_dtype, void* planes, long long rows, int C, int pitch, int c_store, int passes,
                       int pre_act, void* stream);
int affgw_conv_tc_supported(const affgw_conv_desc* d);   /* 1 if the tcgen05 kernel takes this convolution */

int affgw_conv2d_fwd(const void* x, const void* w_packed, const float* bias, const void* addend, void* y,
                     const affgw_conv_desc* d, void* stream);
/* gradient w.r.t. x of the forward described by d.  dy:[N,Ho,Wo,Cout] (y_dtype; operand planes on the tcgen05 route),
 * w_packed_t from affgw_pack_weight[_tc](..., transpose_flip = 1), x only read when pre_act != NONE.  workspace holds
 * the gradient w.r.t. the padded/upsampled virtual input before it is folded back (reflect halo, x2 nearest).
 * On the tcgen05 route dx is a dense [N,H,W,Cin] tensor of grad_dtype. */
long long affgw_conv2d_dgrad_ws_bytes(const affgw_conv_desc* d);
int affgw_conv2d_dgrad(const void* dy, const void* w_packed_t, const void* x, void* dx, void* workspace,
                       const affgw_conv_desc* d, void* stream);
/* dw (OIHW fp32) += dY^T * gather(x); caller zeroes dw first.  With algo = TCGEN05 x and dy are operand planes and the
 * split-K partials are reduced in `workspace` (affgw_conv2d_wgrad_ws_bytes; 0 = not eligible, use algo = SIMT). */
long long affgw_conv2d_wgrad_ws_bytes(const affgw_conv_desc* d);
int affgw_conv2d_wgrad(const void* x, const void* dy, float* dw_oihw, void* workspace, const affgw_conv_desc* d, void* stream);
/* Single-channel-sided stride-1 stencils on the CUDA cores, fp32 in and out: the 1 -> 16 7x7 stems of DisModel /
 * WriterClaModel (modules_tro.py:125-128,175-178) and the 64 -> 1 7x7 + tanh output convolution of the Decoder
 * (modules_tro.py:600-603).  x [N,H,W,Cin] and y / dy [N,Ho,Wo,Cout] dense NHWC fp32, w the OIHW parameter itself (no
 * packing), pre_act must be NONE.  _supported: 0 = no, 1 = one input channel, 2 = one output channel.  workspace:
 * affgw_conv_thin_ws_bytes(d, for_dgrad) bytes.  wgrad ADDS into dw (caller zeroes it). */
int affgw_conv_thin_supported(const affgw_conv_desc* d);
long long affgw_conv_thin_ws_bytes(const affgw_conv_desc* d, int for_dgrad);
int affgw_conv_thin_fwd(const float* x, const float* w_oihw, const float* bias, float* y, void* workspace,
                        const affgw_conv_desc* d, void* stream);
int affgw_conv_thin_dgrad(const float* dy, const float* w_oihw, float* dx, void* workspace, const affgw_conv_desc* d, void* stream);
int affgw_conv_thin_wgrad(const float* x, const float* dy, float* dw_oihw, const affgw_conv_desc* d, void* stream);
/* out[c] += sum_m a[m][c]  (bias gradient); caller zeroes out */
int affgw_colsum(const void* a, int dtype, float* out, long long M, int C, int pitch, void* stream);

/* ---- normalisation family (replaces nn.InstanceNorm2d vgg_tro_channel3_modi.py:50 / blocks.py:127-128,
 *      F.batch_norm blocks.py:197-204, nn.BatchNorm2d blocks.py:250-281, nn.BatchNorm1d modules_tro.py:275,278,
 *      calc_mean_std blocks.py:227-235).  Tensor view: [G][P][C]. ws: 2*G*C floats. --------------------------- */
int affgw_norm_stats(const void* x, int dtype, float* ws, float* mean, float* rstd, float* var_unbiased, int G, long long P,
                     int C, float eps, int unbiased, void* stream);
/* y = act((x-mean)*rstd*gamma + beta) + residual ; gamma/beta may be NULL; affine_per_group: gamma is [G][C] */
int affgw_norm_apply(const void* x, int dtype, const float* mean, const float* rstd, const float* gamma, const float* beta,
                     const void* residual, void* y, int G, long long P, int C, int act, int affine_per_group, void* stream);
/* s1 = dbeta[G][C], s2 = dgamma[G][C] (always written), dx; batch_stats = 0 for fixed (eval) statistics;
 * unbiased = 1 when the forward statistics used the unbiased variance (get_key) */
int affgw_norm_bwd(const void* dy, const void* x, int dtype, const float* mean, const float* rstd, const float* gamma,
                   const float* beta, float* s1, float* s2, void* dx, int G, long long P, int C, int act,
                   int affine_per_group, int batch_stats, int unbiased, void* stream);
int affgw_bn_update_running(float* running_mean, float* running_var, long long* num_batches_tracked, const float* mean,
                            const float* var_unbiased, int C, float momentum, void* stream);
int affgw_bn_eval_stats(const float* running_mean, const float* running_var, float* mean, float* rstd, int C, float eps,
                        void* stream);

/* ---- pooling / resize (nn.MaxPool2d vgg…:45, modules_tro.py:224; ReflectionPad2d(1)+AvgPool2d(3,2)
 *      modules_tro.py:133-134; F.interpolate nearest blocks.py:214) ------------------------------------------ */
int affgw_maxpool2_fwd(const void* x, void* y, int dtype, int N, int H, int W, int C, void* stream);
int affgw_maxpool2_bwd(const void* dy, const void* x, void* dx, int dtype, int N, int H, int W, int C, void* stream);
int affgw_avgpool3s2_fwd(const void* x, void* y, int dtype, int N, int H, int W, int C, void* stream);
int affgw_avgpool3s2_bwd(const void* dy, void* dx, int dtype, int N, int H, int W, int C, void* stream);
int affgw_resize_nearest_fwd(const void* x, void* y, int dtype, int N, int H, int W, int C, int Ho, int Wo, void* stream);
int affgw_resize_nearest_bwd(const void* dy, void* dx, int dtype, int N, int H, int W, int C, int Ho, int Wo, void* stream);
/* torchvision ResNet encoders (modules_tro.py:464-533, modules_tro2.py:447-516): nn.MaxPool2d(3, 2, 1) of the stem,
 * F.interpolate(mode="bilinear", align_corners=False) of the last map (dx fp32, zeroed by the caller),
 * out = act(a + b) for the residual tails */
int affgw_maxpool3s2_fwd(const void* x, void* y, int dtype, int N, int H, int W, int C, void* stream);
int affgw_maxpool3s2_bwd(const void* dy, const void* x, void* dx, int dtype, int N, int H, int W, int C, void* stream);
/* nn.MaxPool2d(3, stride=(sy, sx), padding=1), sy, sx in {1, 2}: the (2,1) and (1,1) pools of Resnet18.py:45-46;
 * y is [N, (H-1)/sy+1, (W-1)/sx+1, C] */
int affgw_maxpool3_fwd(const void* x, void* y, int dtype, int N, int H, int W, int C, int sy, int sx, void* stream);
int affgw_maxpool3_bwd(const void* dy, const void* x, void* dx, int dtype, int N, int H, int W, int C, int sy, int sx, void* stream);
int affgw_resize_bilinear_fwd(const void* x, void* y, int dtype, int N, int H, int W, int C, int Ho, int Wo, void* stream);
int affgw_resize_bilinear_bwd(const void* dy, float* dx, int dtype, int N, int H, int W, int C, int Ho, int Wo, void* stream);
int affgw_add_act(const void* a, const void* b, void* out, int dtype, long long n, int act, void* stream);

/* ---- iAFF pieces (blocks.py:286-299) ---------------------------------------------------------------------- */
/* y = x*w + r*(1-w), w = sigmoid(xl + xg[n,c]) */
int affgw_gate_fwd(const void* x, const void* r, const void* xl, const void* xg, void* y, int dtype, int N, long long P,
                   int C, void* stream);
int affgw_gate_bwd(const void* dy, const void* x, const void* r, const void* xl, const void* xg, void* dx, void* dr,
                   void* dxl, void* dxg, int dtype, int N, long long P, int C, void* stream);
int affgw_gap_fwd(const void* x, void* out, int dtype, int N, long long P, int C, void* stream);
/* out[n,p,c] = (a ? a[n,p,c] : 0) + v[n,c]*scale   (GAP backward, broadcast adds) */
int affgw_bcast_add(const void* a, const void* v, void* out, int dtype, int N, long long P, int C, float scale, void* stream);
int affgw_add2(const void* a, const void* b, void* out, int dtype, long long n, void* stream);
/* dz = dy * act'(.) evaluated from the activation output y (Conv2dBlock activation, blocks.py:161-162; tanh modules_tro.py:602) */
int affgw_act_bwd(const void* dy, const void* y, void* dz, int dtype, long long n, int act, void* stream);

/* ---- text encoder pieces (modules_tro.py:285-317) --------------------------------------------------------- */
/* err: device int set to 1 on an out-of-range id (label handling is bit-exact or it is an error) */
int affgw_embedding_fwd(const long long* ids, const float* table, void* out, int dtype, long long n_ids, int E, int V,
                        int* err, void* stream);
int affgw_embedding_bwd(const long long* ids, const void* dout, float* dtable, int dtype, long long n_ids, int E, int V,
                        void* stream);
/* chars:[B][ts+1][C] (slot ts = PAD embedding) -> out:[B][H][W][C]; column c takes slot c/reps, columns past ts*reps PAD */
int affgw_text_tile_fwd(const void* chars, void* out, int dtype, int B, int H, int W, int C, int ts, int reps, void* stream);
int affgw_text_tile_bwd(const void* dout, void* dchars, int dtype, int B, int H, int W, int C, int ts, int reps, void* stream);

/* ---- losses (nn.BCEWithLogitsLoss modules_tro.py:145,152-168; nn.CrossEntropyLoss modules_tro.py:193-201) -- */
int affgw_bce_logits_fwd(const void* x, int dtype, float target, float* loss, long long n, void* stream);
int affgw_bce_logits_bwd(const void* x, int dtype, float target, const float* grad_out, void* dx, long long n, void* stream);
int affgw_softmax_ce_fwd(const void* x, int dtype, const long long* y, float* loss, int B, int C, int* err, void* stream);
int affgw_softmax_ce_bwd(const void* x, int dtype, const long long* y, const float* grad_out, void* dx, int B, int C,
                         void* stream);

/* ---- recogniser (RecModel, reference modules_tro.py:610-638) ---------------------------------------------
 * Only the pieces that are not convolutions / GEMMs; the projections run through affgw_conv2d_*.  All tensors fp32, dense.
 * GRU cell (torch.nn.GRU semantics; encoder_vgg.py:700, decoder.py:27): gi [N][>= 3H, row pitch gi_pitch] = W_ih x + b_ih,
 * gh [N][3H] = W_hh h + b_hh, gate order (r, z, n); h_out = (1 - z) n + z h.  The backward call recomputes the gates. */
int affgw_gru_cell_fwd(const float* gi, long long gi_pitch, const float* gh, const float* h, float* h_out, int N, int H, void* stream);
int affgw_gru_cell_bwd(const float* dh_out, const float* gi, long long gi_pitch, const float* gh, const float* h, float* dgi,
                       float* dgh, float* dh, int N, int H, void* stream);
/* y[n][p][c] = x[n][p][c] * m[n][c]: nn.Dropout2d with the caller's keep-mask / (1 - p) (encoder_vgg.py:709); its own backward */
int affgw_scale_nc(const float* x, const float* m, float* y, int N, long long P, int C, void* stream);
int affgw_mul2(const float* a, const float* b, float* y, long long n, void* stream);
/* NHWC feature map [B][H][W][C] <-> sequence [W][B][H*C] (out.permute(3,0,2,1).reshape(-1, B, H*C), encoder_vgg.py:711-713) */
int affgw_map_seq(const float* src, float* dst, int B, int H, int W, int C, int to_seq, void* stream);
/* location attention (attention.py:132-158): energy[n][t] = v . tanh(e[sample[n]][t] + hp[n] + loc[n][t]) + vb.
 * e [B][T][F] is the projected encoder output shared by the hypotheses of a sample (sample[n] = row -> sample index);
 * the backward call ACCUMULATES into de / dhp / dv / dvb (zero them first) and writes dloc. */
int affgw_attn_energy_fwd(const float* e, const long long* sample, const float* hp, const float* loc, const float* v, const float* vb,
                          float* energy, int N, int T, int F, void* stream);
int affgw_attn_energy_bwd(const float* denergy, const float* e, const long long* sample, const float* hp, const float* loc,
                          const float* v, float* de, float* dhp, float* dloc, float* dv, float* dvb, int N, int T, int F, void* stream);
/* attn[n] = softmax_t(energy[n]); ctx[n] = sum_t attn[n][t] enc[sample[n]][t]  (attention.py:139-141, decoder.py:38-40);
 * backward: dattn may be NULL, denc is ACCUMULATED (zero it first).  T <= 64. */
int affgw_attn_ctx_fwd(const float* energy, const float* enc, const long long* sample, float* attn, float* ctx, int N, int T, int F,
                       void* stream);
int affgw_attn_ctx_bwd(const float* dattn, const float* dctx, const float* attn, const float* enc, const long long* sample,
                       float* denergy, float* denc, int N, int T, int F, void* stream);

/* ---- DINOv2 ViT style encoder (BASELINE.json configs[3]: GAN_word/dinomodel.py around a ViT backbone), forward only --------
 * LayerNorm over the last dimension; exact (erf) GELU; y = x + gamma[c] * t (LayerScale + residual; gamma may be NULL = 1);
 * multi-head self-attention of qkv [B][N][3][H][hd] (Linear(D, 3D) output) -> [B][N][H*hd], N <= 128 tokens. */
int affgw_layernorm_fwd(const float* x, const float* w, const float* b, float* y, long long rows, int D, float eps, void* stream);
int affgw_gelu_fwd(const float* x, float* y, long long n, void* stream);
int affgw_scale_residual(const float* x, const float* t, const float* gamma, float* y, long long n, int D, void* stream);
int affgw_attention_fwd(const float* qkv, float* out, int B, int N, int H, int hd, float scale, void* stream);

/* ---- line-level generator (line_generation/model/pure_gen.py) ------------------------------------------
 * depthwise 3x3 binomial blur (1,2,1)x(1,2,1)/16, zero padding, fp32 NHWC (Blur, pure_gen.py:123-136); symmetric, so the same
 * call is its own backward.  PixelNorm over the last dimension (pure_gen.py:306-311). */
int affgw_blur3(const float* x, float* y, int N, int H, int W, int C, void* stream);
int affgw_pixelnorm(const float* x, float* y, int rows, int C, float eps, void* stream);

/* Recogniser loss of the GAN step: crit(log_softmax(x), y) = LabelSmoothing(vocab, PAD, 0.4) over KLDivLoss(reduction='sum')
 * (reference loss_tro.py:8-35 as called at network_tro.py:44-45,92-93).  x [rows][V] fp32 logits, y [rows] int64 targets;
 * rows whose target is pad_idx and the pad_idx column carry no mass; NaN logits give a NaN loss like torch does. */
int affgw_label_smooth_kl_fwd(const float* x, const long long* y, float* loss, int rows, int V, int pad_idx, float smoothing,
                              int* err, void* stream);
int affgw_label_smooth_kl_bwd(const float* x, const long long* y, const float* grad_out, float* dx, int rows, int V, int pad_idx,
                              float smoothing, void* stream);

/* ---- layout / dtype (the boundary: callers hand NCHW fp32 tensors, network_tro.py:30-36) ------------------- */
/* Wire format of the style / target images: grey-level uint8 pixels as cv2 leaves them after the
 * resize (load_data.py:151), right padding = 255.  dst[i] = the reference's normalisation of src[i], bit for bit:
 * float32((float32(1. - src/255.) - 0.5) / 0.5) (load_data.py:152-166).  Both buffers 16-byte aligned, n elements. */
int affgw_u8_to_image(const unsigned char* src, float* dst, long long n, void* stream);
int affgw_nchw_to_nhwc(const float* x, void* y, int dtype, int N, int C, long long HW, int c_pad, void* stream);
int affgw_nhwc_to_nchw(const void* x, float* y, int dtype, int N, int C, long long HW, int c_pad, void* stream);
int affgw_cast(const void* x, int in_dtype, void* y, int out_dtype, long long n, void* stream);
/* out[r] = cat(a[r], b[r]) over rows = N*H*W pixels (torch.cat(dim=1) in GenModel_FC.mix, modules_tro.py:256) and its inverse */
int affgw_concat_channels(const void* a, const void* b, void* out, int dtype, long long rows, int ca, int cb, void* stream);
int affgw_split_channels(const void* in, void* a, void* b, int dtype, long long rows, int ca, int cb, void* stream);

/* ---- data-parallel gradient exchange (replaces the nn.DataParallel reduce at modules_tro.py:341-346) ------- */
/* gather n tensors into one contiguous fp32 bucket (and back) so that one NCCL all-reduce covers the bucket;
 * ptrs / sizes are DEVICE arrays of length n, offsets are element offsets into bucket. scale is applied on unpack. */
int affgw_bucket_pack(const float* const* ptrs, const long long* sizes, const long long* offsets, int n, float* bucket,
                      void* stream);
int affgw_bucket_unpack(float* const* ptrs, const long long* sizes, const long long* offsets, int n, const float* bucket,
                        float scale, void* stream);
/* torch.optim.Adam step (main_run.py:275-278: betas (0.9, 0.999), eps 1e-8, no weight decay, no amsgrad) over a table of n
 * fp32 tensors in ONE launch.  Device tables: params / grads / exp_avg / exp_avg_sq pointers, sizes and
 * dense increasing offsets (offsets[i+1] = offsets[i] + sizes[i]) as for the buckets.  `step` >= 1 is the step count AFTER
 * this call (bias corrections 1 - beta^step are formed on the host in double); grads are multiplied by grad_scale first.
 *   m += (g - m)(1 - beta1);  v = beta2 v + (1 - beta2) g^2;  p -= lr / bc1 * m / (sqrt(v) / sqrt(bc2) + eps)              */
int affgw_adam_step(float* const* params, const float* const* grads, float* const* exp_avg, float* const* exp_avg_sq,
                    const long long* sizes, const long long* offsets, int n, float lr, float beta1, float beta2, float eps,
                    long long step, float grad_scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AFFGW_H */
