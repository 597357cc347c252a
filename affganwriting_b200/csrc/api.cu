// extern "C" surface of libaffgw (see include/affgw.h).  Validates arguments, builds geometry, picks the kernel.
#include <stdarg.h>
#include <string.h>

#include <atomic>

#include "../../include/affgw.h"
#include "common.cuh"
#include "pos_frame.cuh"

// ---- implemented in the other translation units -------------------------------------------------------------
int conv_fwd_simt(const void* x, int x_dt, const void* w, int w_dt, const float* bias, const void* addend, void* y,
                  int y_dt, const ConvGeom& g, cudaStream_t st);
int conv_wgrad_simt(const void* x, int x_dt, const void* dy, int dy_dt, float* dw, const ConvGeom& g, cudaStream_t st);
int conv_fold(const void* dxp, const void* xin, void* dx, int dt, int N, int H, int W, int C, int pad, int pad_mode, int up,
              int pre_act, cudaStream_t st);
int colsum(const void* a, int dt, float* out, long long M, int C, int pitch, cudaStream_t st);
int pack_weight(const float* w, void* out, int out_dt, int Cout, int Cin, int KH, int KW, int ipad, int transpose_flip,
                cudaStream_t st);
// conv_tc.cu
int conv_tc_block_n(int cout);
int conv_tc_ok(const ConvGeom& g);
int conv_fwd_tc(const void* x_planes, long long plane_stride, const void* w_tiles, const float* bias, const void* addend,
                void* y, int y_dt, const ConvGeom& g, int passes, int fmt, const float* alpha_dev, cudaStream_t st);
int pack_weight_tc(const float* w, void* out, int Cout, int Cin, int KH, int KW, int ipad, int transpose_flip, int passes, int fmt,
                   cudaStream_t st);
long long pack_weight_tc_bytes(int Cout, int Cin, int KH, int KW, int ipad, int transpose_flip, int passes);
int split_planes(const void* x, int x_dt, void* planes, long long rows, int C, int pitch, int c_store, int passes, int pre_act,
                 int fmt, const float* scale_dev, cudaStream_t st);
// conv_shift.cu
int shift_block_n(int cout);
int wgrad_shift_block_n(int cout);
int conv_shift_ok(const ConvGeom& g);
void conv_shift_frame(const ConvGeom& g, int channels, PosFrame& f);
int split_positions(const void* src, int dt, void* planes, const PosFrame& f, int Hs, int Ws, int C, int pitch, int up, int oy0,
                    int ox0, int pad_mode, int pre_act, int passes, float* colsum, int fmt, const float* scale_dev, cudaStream_t st);
int amax_scale(const float* x, long long n, float* out2, unsigned* ws, cudaStream_t st);
int conv_pos_tc(const void* x_planes, const PosFrame& f, const void* w_packed, const float* bias, const void* addend, void* y,
                int K, int q_shift, int oy0, int ox0, int OH, int OW, int Cout, int out_pitch, int post_act, int passes,
                int fmt, const float* alpha_dev, cudaStream_t st);
int pack_weight_shift(const float* w, void* out, int Cout, int Cin, int K, int ipad, int transpose_flip, int passes, int fmt,
                      cudaStream_t st);
long long pack_weight_shift_bytes(int Cout, int Cin, int K, int ipad, int transpose_flip, int passes);
// conv_thin.cu
int conv_thin_ok(int Cin, int Cout, int K, int stride, int up);
int shift_tile_mt(int cout, int npl, long long q_last);
int conv_thin_fwd(const float* x, const float* w, const float* bias, float* y, float* scratch, const ConvGeom& g, cudaStream_t st);
int conv_thin_dgrad_frame(const float* dy, const float* w, float* dframe, float* scratch, const ConvGeom& g, cudaStream_t st);
int conv_thin_wgrad(const float* x, const float* dy, float* dw, const ConvGeom& g, cudaStream_t st);
// conv_tc_wgrad.cu
int conv_wgrad_tc_ok(const ConvGeom& g);
long long conv_wgrad_tc_ws_bytes(const ConvGeom& g);
int conv_wgrad_tc(const void* x_planes, long long x_plane, const void* dy_planes, long long dy_plane, float* dw, void* workspace,
                  const ConvGeom& g, int cin_w, int passes, int fmt, const float* alpha_dev, cudaStream_t st);
int conv_wgrad_pos(const void* x_planes, const PosFrame& fx, const void* dy_planes, const PosFrame& fy, float* dw, void* workspace,
                   const ConvGeom& g, int cin_w, int passes, int fmt, const float* alpha_dev, cudaStream_t st);
// norm.cu
int norm_stats(const void* x, int dt, float* ws, float* mean, float* rstd, float* var_unbiased, int G, long long P, int C,
               float eps, int unbiased, cudaStream_t st);
int norm_apply(const void* x, int dt, const float* mean, const float* rstd, const float* gamma, const float* beta,
               const void* residual, void* y, int G, long long P, int C, int act, int affine_per_group, cudaStream_t st);
int norm_bwd(const void* dy, const void* x, int dt, const float* mean, const float* rstd, const float* gamma,
             const float* beta, float* s1, float* s2, void* dx, int G, long long P, int C, int act, int affine_per_group,
             int batch_stats, int unbiased, cudaStream_t st);
int bn_update_running(float* rm, float* rv, long long* nbt, const float* mean, const float* var_unbiased, int C,
                      float momentum, cudaStream_t st);
int bn_eval_stats(const float* rm, const float* rv, float* mean, float* rstd, int C, float eps, cudaStream_t st);
// pointwise.cu
int maxpool2_fwd(const void* x, void* y, int dt, int N, int H, int W, int C, cudaStream_t st);
int maxpool2_bwd(const void* dy, const void* x, void* dx, int dt, int N, int H, int W, int C, cudaStream_t st);
int avgpool3s2_fwd(const void* x, void* y, int dt, int N, int H, int W, int C, cudaStream_t st);
int avgpool3s2_bwd(const void* dy, void* dx, int dt, int N, int H, int W, int C, cudaStream_t st);
int gate_fwd(const void* x, const void* r, const void* xl, const void* xg, void* y, int dt, int N, long long P, int C,
             cudaStream_t st);
int gate_bwd(const void* dy, const void* x, const void* r, const void* xl, const void* xg, void* dx, void* dr, void* dxl,
             void* dxg, int dt, int N, long long P, int C, cudaStream_t st);
int gap_fwd(const void* x, void* out, int dt, int N, long long P, int C, cudaStream_t st);
int bcast_add(const void* a, const void* v, void* out, int dt, int N, long long P, int C, float scale, cudaStream_t st);
int add2(const void* a, const void* b, void* out, int dt, long long n, cudaStream_t st);
int add_act(const void* a, const void* b, void* out, int dt, long long n, int act, cudaStream_t st);
int maxpool3_fwd(const void* x, void* y, int dt, int N, int H, int W, int C, int SY, int SX, cudaStream_t st);
int maxpool3_bwd(const void* dy, const void* x, void* dx, int dt, int N, int H, int W, int C, int SY, int SX, cudaStream_t st);
int resize_bilinear_fwd(const void* x, void* y, int dt, int N, int H, int W, int C, int Ho, int Wo, cudaStream_t st);
int resize_bilinear_bwd(const void* dy, float* dx, int dt, int N, int H, int W, int C, int Ho, int Wo, cudaStream_t st);
int act_bwd(const void* dy, const void* y, void* dz, int dt, long long n, int act, cudaStream_t st);
int resize_nearest_fwd(const void* x, void* y, int dt, int N, int H, int W, int C, int Ho, int Wo, cudaStream_t st);
int resize_nearest_bwd(const void* dy, void* dx, int dt, int N, int H, int W, int C, int Ho, int Wo, cudaStream_t st);
int embedding_fwd(const long long* ids, const float* table, void* out, int dt, long long n_ids, int E, int V, int* err,
                  cudaStream_t st);
int embedding_bwd(const long long* ids, const void* dout, float* dtable, int dt, long long n_ids, int E, int V,
                  cudaStream_t st);
int text_tile_fwd(const void* chars, void* out, int dt, int B, int H, int W, int C, int ts, int reps, cudaStream_t st);
int text_tile_bwd(const void* dout, void* dchars, int dt, int B, int H, int W, int C, int ts, int reps, cudaStream_t st);
int bce_logits_fwd(const void* x, int dt, float target, float* loss, long long n, cudaStream_t st);
int bce_logits_bwd(const void* x, int dt, float target, const float* gout, void* dx, long long n, cudaStream_t st);
int softmax_ce_fwd(const void* x, int dt, const long long* y, float* loss, int B, int C, int* err, cudaStream_t st);
int softmax_ce_bwd(const void* x, int dt, const long long* y, const float* gout, void* dx, int B, int C, cudaStream_t st);
int gru_cell_fwd(const float* gi, long long gi_pitch, const float* gh, const float* h, float* hout, int N, int H, cudaStream_t st);
int gru_cell_bwd(const float* dhout, const float* gi, long long gi_pitch, const float* gh, const float* h, float* dgi, float* dgh,
                 float* dh, int N, int H, cudaStream_t st);
int scale_nc(const float* x, const float* m, float* y, int N, long long P, int C, cudaStream_t st);
int mul2(const float* a, const float* b, float* y, long long n, cudaStream_t st);
int map_seq(const float* src, float* dst, int B, int H, int W, int C, int to_seq, cudaStream_t st);
int attn_energy_fwd(const float* e, const long long* sidx, const float* hp, const float* loc, const float* v, const float* vb,
                    float* energy, int N, int T, int F, cudaStream_t st);
int attn_energy_bwd(const float* denergy, const float* e, const long long* sidx, const float* hp, const float* loc, const float* v,
                    float* de, float* dhp, float* dloc, float* dv, float* dvb, int N, int T, int F, cudaStream_t st);
int attn_ctx_fwd(const float* energy, const float* enc, const long long* sidx, float* attn, float* ctx, int N, int T, int F,
                 cudaStream_t st);
int attn_ctx_bwd(const float* dattn, const float* dctx, const float* attn, const float* enc, const long long* sidx, float* denergy,
                 float* denc, int N, int T, int F, cudaStream_t st);
int layernorm_fwd(const float* x, const float* w, const float* b, float* y, long long rows, int D, float eps, cudaStream_t st);
int gelu_fwd(const float* x, float* y, long long n, cudaStream_t st);
int scale_residual(const float* x, const float* t, const float* gamma, float* y, long long n, int D, cudaStream_t st);
int attention_fwd(const float* qkv, float* out, int B, int N, int H, int hd, float scale, cudaStream_t st);
int blur3(const float* x, float* y, int N, int H, int W, int C, cudaStream_t st);
int pixelnorm(const float* x, float* y, int rows, int C, float eps, cudaStream_t st);
int label_smooth_kl_fwd(const float* x, const long long* y, float* loss, int rows, int V, int pad, float smoothing, int* err,
                        cudaStream_t st);
int label_smooth_kl_bwd(const float* x, const long long* y, const float* gout, float* dx, int rows, int V, int pad, float smoothing,
                        cudaStream_t st);
int nchw_to_nhwc(const float* x, void* y, int dt, int N, int C, long long HW, int cpad, cudaStream_t st);
int u8_to_image(const unsigned char* src, float* dst, long long n, cudaStream_t st);
int nhwc_to_nchw(const void* x, float* y, int dt, int N, int C, long long HW, int cpad, cudaStream_t st);
int cast_dtype(const void* x, int in_dt, void* y, int out_dt, long long n, cudaStream_t st);
int concat2(void* a, void* b, void* out, int dt, long long rows, int ca, int cb, int to_out, cudaStream_t st);

// ---- error string + launch counter --------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void affgw_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void affgw_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

static inline cudaStream_t S(void* s) { return (cudaStream_t)s; }
static inline bool dt_ok(int dt) { return dt == AFFGW_F32 || dt == AFFGW_BF16; }
static inline size_t dt_size(int dt) { return dt == AFFGW_F32 ? 4 : 2; }

static int make_geom(const affgw_conv_desc* d, ConvGeom& g, int zero_insert = 1) {
    AFFGW_CHECK(d != nullptr, "conv: null descriptor");
    AFFGW_CHECK(d->N > 0 && d->H > 0 && d->W > 0 && d->Cin > 0 && d->Cout > 0 && d->KH > 0 && d->KW > 0,
                "conv: non-positive extent");
    AFFGW_CHECK(d->stride >= 1 && d->stride_w >= 0 && d->pad >= 0, "conv: bad stride/pad");
    const int sw = d->stride_w ? d->stride_w : d->stride;
    AFFGW_CHECK(d->upsample == 1 || d->upsample == 2, "conv: upsample must be 1 or 2");
    AFFGW_CHECK(d->pad_mode >= 0 && d->pad_mode <= 2, "conv: bad pad_mode");
    AFFGW_CHECK(dt_ok(d->x_dtype) && dt_ok(d->w_dtype) && dt_ok(d->y_dtype), "conv: bad dtype");
    AFFGW_CHECK(d->in_pitch >= d->Cin && d->out_pitch >= d->Cout, "conv: pitch smaller than channel count");
    g.N = d->N; g.H = d->H; g.W = d->W; g.Cin = d->Cin;
    g.Cout = d->Cout; g.KH = d->KH; g.KW = d->KW;
    g.stride = d->stride; g.pad = d->pad; g.pad_mode = d->pad_mode; g.up = d->upsample; g.zi = zero_insert;
    g.stride_w = sw; g.zi_w = zero_insert;
    g.Ho = d->Ho; g.Wo = d->Wo;
    g.in_pitch = d->in_pitch; g.out_pitch = d->out_pitch;
    g.pre_act = d->pre_act; g.post_act = d->post_act;
    g.Hv = d->H * d->upsample; g.Wv = d->W * d->upsample;
    g.Ktot = d->KH * d->KW * d->Cin;
    g.M = (long long)d->N * d->Ho * d->Wo;
    const int eh = (g.Hv + 2 * d->pad - d->KH) / d->stride + 1, ew = (g.Wv + 2 * d->pad - d->KW) / sw + 1;
    AFFGW_CHECK(eh == d->Ho && ew == d->Wo, "conv: output extent %dx%d does not match the geometry (%dx%d)", d->Ho, d->Wo,
                eh, ew);
    if (d->pad_mode == PAD_REFLECT) AFFGW_CHECK(d->pad < g.Hv && d->pad < g.Wv, "conv: reflect pad >= input extent");
    return 0;
}

extern "C" {

int affgw_version(void) { return AFFGW_VERSION; }
const char* affgw_last_error(void) { return g_err; }
long long affgw_launch_count(void) { return g_launches.load(); }

int affgw_device_ok(void) {
    int dev = 0;
    cudaDeviceProp p;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&p, dev) != cudaSuccess) {
        affgw_set_error("no CUDA device");
        return 0;
    }
    if (p.major != 10) {
        affgw_set_error("device sm_%d%d is not sm_100", p.major, p.minor);
        return 0;
    }
    return 1;
}

int affgw_pack_weight(const float* w, void* out, int out_dtype, int Cout, int Cin, int KH, int KW, int i_pad,
                      int transpose_flip, void* stream) {
    AFFGW_CHECK(w && out && dt_ok(out_dtype), "pack_weight: bad argument");
    AFFGW_CHECK(i_pad >= (transpose_flip ? Cout : Cin), "pack_weight: i_pad too small");
    return pack_weight(w, out, out_dtype, Cout, Cin, KH, KW, i_pad, transpose_flip, S(stream));
}

static inline bool passes_ok(int p) { return p == 1 || p == 3; }
static std::atomic<int> g_prefer_shift{1};

int affgw_conv_tc_prefer_shift(int enable) {
    const int prev = g_prefer_shift.load();
    if (enable >= 0) g_prefer_shift.store(enable ? 1 : 0);
    return prev;
}

long long affgw_pack_weight_tc_bytes(int Cout, int Cin, int KH, int KW, int i_pad, int transpose_flip, int passes, int layout) {
    if (layout == AFFGW_WLAYOUT_SHIFT) return KH == KW ? pack_weight_shift_bytes(Cout, Cin, KH, i_pad, transpose_flip, passes) : -1;
    if (layout != AFFGW_WLAYOUT_IM2COL) return -1;
    return pack_weight_tc_bytes(Cout, Cin, KH, KW, i_pad, transpose_flip, passes);
}
int affgw_pack_weight_tc_fmt(const float* w, void* out, int Cout, int Cin, int KH, int KW, int i_pad, int transpose_flip,
                             int passes, int layout, int operand_fmt, void* stream) {
    AFFGW_CHECK(w && out, "pack_weight_tc: null pointer");
    AFFGW_CHECK(layout == AFFGW_WLAYOUT_IM2COL || layout == AFFGW_WLAYOUT_SHIFT, "pack_weight_tc: bad layout");
    AFFGW_CHECK(operand_fmt == AFFGW_FMT_BF16 || operand_fmt == AFFGW_FMT_F16, "pack_weight_tc: bad operand format");
    if (layout == AFFGW_WLAYOUT_SHIFT) {
        AFFGW_CHECK(KH == KW, "pack_weight_tc: the shifted kernel takes square filters");
        return pack_weight_shift(w, out, Cout, Cin, KH, i_pad, transpose_flip, passes, operand_fmt, S(stream));
    }
    return pack_weight_tc(w, out, Cout, Cin, KH, KW, i_pad, transpose_flip, passes, operand_fmt, S(stream));
}
int affgw_pack_weight_tc(const float* w, void* out, int Cout, int Cin, int KH, int KW, int i_pad, int transpose_flip,
                         int passes, int layout, void* stream) {
    return affgw_pack_weight_tc_fmt(w, out, Cout, Cin, KH, KW, i_pad, transpose_flip, passes, layout, AFFGW_FMT_BF16, stream);
}
long long affgw_operand_planes_bytes(long long rows, int c_store, int passes) {
    if (rows <= 0 || c_store <= 0 || c_store % 8 != 0 || !passes_ok(passes)) return -1;
    return rows * c_store * 2LL * (passes == 3 ? 2 : 1);
}
int affgw_split_planes_fmt(const void* x, int x_dtype, void* planes, long long rows, int C, int pitch, int c_store, int passes,
                           int pre_act, int operand_fmt, const float* scale_dev, void* stream) {
    AFFGW_CHECK(x && planes && dt_ok(x_dtype) && rows > 0 && C > 0 && pitch >= C, "split_planes: bad argument");
    AFFGW_CHECK(pre_act >= 0 && pre_act <= 3, "split_planes: bad activation");
    AFFGW_CHECK(operand_fmt == AFFGW_FMT_BF16 || operand_fmt == AFFGW_FMT_F16, "split_planes: bad operand format");
    AFFGW_CHECK(scale_dev == nullptr || operand_fmt == AFFGW_FMT_F16, "split_planes: a scale is for fp16 planes");
    return split_planes(x, x_dtype, planes, rows, C, pitch, c_store, passes, pre_act, operand_fmt, scale_dev, S(stream));
}
int affgw_split_planes(const void* x, int x_dtype, void* planes, long long rows, int C, int pitch, int c_store, int passes,
                       int pre_act, void* stream) {
    return affgw_split_planes_fmt(x, x_dtype, planes, rows, C, pitch, c_store, passes, pre_act, AFFGW_FMT_BF16, nullptr, stream);
}

// geometry seen by the tcgen05 kernels: channels = the STORED channel count of the operand planes
static int make_tc_geom(const affgw_conv_desc* d, ConvGeom& g) {
    if (int rc = make_geom(d, g)) return rc;
    AFFGW_CHECK(passes_ok(d->passes), "conv (tcgen05): passes must be 1 or 3");
    AFFGW_CHECK(d->operand_fmt == AFFGW_FMT_BF16 || d->operand_fmt == AFFGW_FMT_F16, "conv (tcgen05): operand_fmt must be 0 (bf16) or 1 (fp16)");
    AFFGW_CHECK(d->x_dtype == AFFGW_BF16 && d->w_dtype == AFFGW_BF16, "conv (tcgen05): operands are 16-bit planes / tiles");
    AFFGW_CHECK(d->pre_act == ACT_NONE, "conv (tcgen05): the pre-activation is applied by affgw_split_planes");
    g.Cin = d->in_pitch;
    g.Ktot = g.KH * g.KW * g.Cin;
    return 0;
}

// which tcgen05 kernel family (= which operand-plane and packed-weight layout) the FORWARD geometry g selects; forward,
// dgrad and wgrad of one convolution always use the same family
static int tc_layout(const ConvGeom& g) {
    if (g_prefer_shift.load() && conv_shift_ok(g)) return AFFGW_WLAYOUT_SHIFT;
    return conv_tc_ok(g) > 0 ? AFFGW_WLAYOUT_IM2COL : 0;
}

int affgw_conv_tc_supported(const affgw_conv_desc* d) {
    ConvGeom g;
    if (!d || make_tc_geom(d, g)) return 0;
    return tc_layout(g) != 0;
}

int affgw_conv2d_fwd(const void* x, const void* w, const float* bias, const void* addend, void* y,
                     const affgw_conv_desc* d, void* stream) {
    ConvGeom g;
    if (int rc = make_geom(d, g)) return rc;
    AFFGW_CHECK(x && w && y, "conv2d_fwd: null pointer");
    if (d->algo == AFFGW_ALGO_TCGEN05) {
        if (int rc = make_tc_geom(d, g)) return rc;
        const int lay = tc_layout(g);
        AFFGW_CHECK(lay != 0, "conv2d_fwd: shape not supported by the tcgen05 kernels (stored Cin %d)", g.Cin);
        if (lay == AFFGW_WLAYOUT_SHIFT) {
            AFFGW_CHECK(d->y_dtype == AFFGW_F32, "conv2d_fwd: the position-space kernel writes fp32");
            PosFrame f;
            conv_shift_frame(g, d->Cin, f);
            return conv_pos_tc(x, f, w, bias, addend, y, d->KH, 0, 0, 0, d->Ho, d->Wo, d->Cout, d->out_pitch, d->post_act,
                               d->passes, d->operand_fmt, nullptr, S(stream));
        }
        const long long plane = (long long)d->N * d->H * d->W * d->in_pitch;
        return conv_fwd_tc(x, plane, w, bias, addend, y, d->y_dtype, g, d->passes, d->operand_fmt, nullptr, S(stream));
    }
    return conv_fwd_simt(x, d->x_dtype, w, d->w_dtype, bias, addend, y, d->y_dtype, g, S(stream));
}

// geometry of the dgrad convolution: input dY [N,Ho,Wo,Cout] (zero-inserted by stride), weights [Cin][K][K][Cout]
static int make_dgrad(const affgw_conv_desc* d, affgw_conv_desc& dd, bool& direct, int& Hp, int& Wp) {
    AFFGW_CHECK(d->KH == d->KW, "conv2d_dgrad: square kernels only");
    const bool tc = d->algo == AFFGW_ALGO_TCGEN05;
    if (!tc) AFFGW_CHECK(d->x_dtype == d->y_dtype, "conv2d_dgrad: x and y dtypes must match");
    direct = d->pad_mode == PAD_ZERO && d->upsample == 1 && d->pre_act == ACT_NONE && d->pad <= d->KH - 1;
    Hp = d->H * d->upsample + 2 * d->pad;
    Wp = d->W * d->upsample + 2 * d->pad;
    dd = *d;
    dd.N = d->N; dd.H = d->Ho; dd.W = d->Wo; dd.Cin = d->Cout; dd.Cout = d->Cin;
    dd.stride = 1; dd.stride_w = 0; dd.pad_mode = PAD_ZERO; dd.upsample = 1;
    dd.pad = direct ? d->KH - 1 - d->pad : d->KH - 1;
    dd.Ho = direct ? d->H : Hp;
    dd.Wo = direct ? d->W : Wp;
    dd.in_pitch = d->out_pitch;
    dd.out_pitch = tc ? d->Cin : (direct ? d->in_pitch : d->Cin);
    dd.pre_act = ACT_NONE; dd.post_act = ACT_NONE;
    dd.x_dtype = d->y_dtype; dd.y_dtype = tc ? d->grad_dtype : d->x_dtype;
    return 0;
}

static void dgrad_geom(const affgw_conv_desc* d, const affgw_conv_desc& dd, ConvGeom& g) {
    // built by hand: the virtual input is dY zero-inserted by the forward stride
    g.N = dd.N; g.H = dd.H; g.W = dd.W; g.Cin = dd.Cin; g.Cout = dd.Cout; g.KH = dd.KH; g.KW = dd.KW;
    const int sw = d->stride_w ? d->stride_w : d->stride;
    g.stride = g.stride_w = 1; g.pad = dd.pad; g.pad_mode = PAD_ZERO; g.up = 1; g.zi = d->stride; g.zi_w = sw;
    g.Ho = dd.Ho; g.Wo = dd.Wo; g.in_pitch = dd.in_pitch; g.out_pitch = dd.out_pitch;
    g.pre_act = ACT_NONE; g.post_act = ACT_NONE;
    g.Hv = (dd.H - 1) * d->stride + 1; g.Wv = (dd.W - 1) * sw + 1;
    if (d->algo == AFFGW_ALGO_TCGEN05) g.Cin = dd.in_pitch;       // stored channels of the dY planes
    g.Ktot = g.KH * g.KW * g.Cin;
    g.M = (long long)g.N * g.Ho * g.Wo;
}

int affgw_conv_tc_layout(const affgw_conv_desc* d, int for_dgrad) {
    ConvGeom g;
    if (!d || !passes_ok(d->passes) || d->algo != AFFGW_ALGO_TCGEN05) return 0;
    affgw_conv_desc f = *d;
    f.x_dtype = f.w_dtype = AFFGW_BF16;
    f.pre_act = ACT_NONE;
    if (make_tc_geom(&f, g)) return 0;
    const int lay = tc_layout(g);
    if (lay == AFFGW_WLAYOUT_SHIFT || !for_dgrad) return lay;
    affgw_conv_desc dd;
    bool direct;
    int Hp, Wp;
    if (make_dgrad(d, dd, direct, Hp, Wp)) return 0;
    dgrad_geom(d, dd, g);
    return conv_tc_ok(g) > 0 ? AFFGW_WLAYOUT_IM2COL : 0;
}

// output-channel tile width (the BN template argument) of the kernel that runs d: which = 0 forward, 1 dgrad, 2 wgrad
// positions per CTA tile of the kernel that runs d (which = 0 forward, 1 dgrad; profiling labels: matches the third
// template argument x 128 of conv_shift_tcgen05_kernel), 128 for every other kernel
int affgw_conv_tc_tile_m(const affgw_conv_desc* d, int which) {
    if (!d || which < 0 || which > 1 || affgw_conv_tc_layout(d, which) != AFFGW_WLAYOUT_SHIFT) return 128;
    const long long Hp = (long long)d->H * d->upsample + 2 * d->pad, Wp = (long long)d->W * d->upsample + 2 * d->pad;
    long long q_last;
    if (which == 0) {
        q_last = ((long long)(d->N - 1) * Hp + d->Ho - 1) * Wp + d->Wo - 1;
    } else {
        const bool direct = d->pad_mode == PAD_ZERO && d->upsample == 1 && d->pre_act == ACT_NONE;
        q_last = direct ? ((long long)(d->N - 1) * Hp + d->pad + d->H - 1) * Wp + d->pad + d->W - 1 : (long long)d->N * Hp * Wp - 1;
    }
    return 128 * shift_tile_mt(which == 1 ? d->Cin : d->Cout, d->passes == 3 ? 2 : 1, q_last);
}
int affgw_conv_tc_tile_n(const affgw_conv_desc* d, int which) {
    if (!d || which < 0 || which > 2) return 0;
    const int lay = affgw_conv_tc_layout(d, which == 1);
    if (!lay) return 0;
    if (lay == AFFGW_WLAYOUT_SHIFT) return which == 2 ? wgrad_shift_block_n(d->Cout) : shift_block_n(which == 1 ? d->Cin : d->Cout);
    return conv_tc_block_n(which == 1 ? d->Cin : d->Cout);
}

// forward frame of a position-space convolution (geometry only; dtypes / pre_act of d are not looked at)
static int shift_frames(const affgw_conv_desc* d, ConvGeom& g, PosFrame& fx, PosFrame& fy) {
    affgw_conv_desc f = *d;
    f.x_dtype = f.w_dtype = AFFGW_BF16;
    f.pre_act = ACT_NONE;
    if (int rc = make_tc_geom(&f, g)) return rc;
    AFFGW_CHECK(conv_shift_ok(g), "not a position-space convolution (stride %d, %dx%d filter)", d->stride, d->KH, d->KW);
    conv_shift_frame(g, d->Cin, fx);
    conv_shift_frame(g, d->Cout, fy);
    return 0;
}
static void export_frame(const PosFrame& f, affgw_pos_frame* o) {
    o->N = f.N; o->Hp = f.Hp; o->Wp = f.Wp; o->G = f.G; o->lead = f.lead; o->reserved = 0; o->QA = f.QA;
}
static PosFrame import_frame(const affgw_pos_frame* o) {
    PosFrame f;
    f.N = o->N; f.Hp = o->Hp; f.Wp = o->Wp; f.G = o->G; f.lead = o->lead; f.QA = o->QA;
    return f;
}

int affgw_conv_pos_frames(const affgw_conv_desc* d, affgw_pos_frame* fx, affgw_pos_frame* fy) {
    AFFGW_CHECK(d && fx && fy, "conv_pos_frames: null pointer");
    ConvGeom g;
    PosFrame a, b;
    if (int rc = shift_frames(d, g, a, b)) return rc;
    export_frame(a, fx);
    export_frame(b, fy);
    return 0;
}
long long affgw_position_planes_bytes(const affgw_pos_frame* f, int passes) {
    if (!f || !passes_ok(passes) || f->G <= 0 || f->QA <= 0) return -1;
    return (long long)(passes == 3 ? 2 : 1) * f->G * f->QA * 16;
}
int affgw_amax_scale(const float* x, long long n, float* scale2, void* workspace4, void* stream) {
    AFFGW_CHECK(x && scale2 && workspace4 && n > 0, "amax_scale: bad argument");
    AFFGW_CHECK(((uintptr_t)x & 15) == 0, "amax_scale: the tensor must be 16-byte aligned");
    return amax_scale(x, n, scale2, (unsigned*)workspace4, S(stream));
}
int affgw_split_positions(const void* src, int dtype, void* planes, const affgw_pos_frame* f, int Hs, int Ws, int C, int pitch,
                          int upsample, int oy0, int ox0, int pad_mode, int pre_act, int passes, float* colsum, void* stream) {
    return affgw_split_positions_fmt(src, dtype, planes, f, Hs, Ws, C, pitch, upsample, oy0, ox0, pad_mode, pre_act, passes, colsum,
                                     AFFGW_FMT_BF16, nullptr, stream);
}
int affgw_split_positions_fmt(const void* src, int dtype, void* planes, const affgw_pos_frame* f, int Hs, int Ws, int C, int pitch,
                              int upsample, int oy0, int ox0, int pad_mode, int pre_act, int passes, float* colsum,
                              int operand_fmt, const float* scale_dev, void* stream) {
    AFFGW_CHECK(operand_fmt == AFFGW_FMT_BF16 || operand_fmt == AFFGW_FMT_F16, "split_positions: bad operand format");
    AFFGW_CHECK(scale_dev == nullptr || operand_fmt == AFFGW_FMT_F16, "split_positions: a scale is for fp16 planes");
    AFFGW_CHECK(src && planes && f && dt_ok(dtype) && passes_ok(passes), "split_positions: bad argument");
    AFFGW_CHECK(Hs > 0 && Ws > 0 && C > 0 && pitch >= C && (upsample == 1 || upsample == 2), "split_positions: bad source");
    AFFGW_CHECK(pad_mode >= 0 && pad_mode <= 2 && pre_act >= 0 && pre_act <= 3, "split_positions: bad mode");
    AFFGW_CHECK(C <= 8 * f->G, "split_positions: %d channels do not fit %d groups", C, f->G);
    if (pad_mode == PAD_REFLECT)
        AFFGW_CHECK(oy0 < Hs * upsample && ox0 < Ws * upsample && f->Hp - oy0 - Hs * upsample < Hs * upsample &&
                        f->Wp - ox0 - Ws * upsample < Ws * upsample, "split_positions: reflect pad >= input extent");
    AFFGW_CHECK(colsum == nullptr || (upsample == 1 && pad_mode == PAD_ZERO && pre_act == ACT_NONE),
                "split_positions: the fused column sum is for dY planes (no padding copies, no activation)");
    return split_positions(src, dtype, planes, import_frame(f), Hs, Ws, C, pitch, upsample, oy0, ox0, pad_mode, pre_act, passes,
                           colsum, operand_fmt, scale_dev, S(stream));
}

long long affgw_conv2d_dgrad_ws_bytes(const affgw_conv_desc* d) {
    affgw_conv_desc dd;
    bool direct;
    int Hp, Wp;
    if (!d || make_dgrad(d, dd, direct, Hp, Wp)) return -1;
    const int gdt = d->algo == AFFGW_ALGO_TCGEN05 ? d->grad_dtype : d->x_dtype;
    return direct ? 0 : (long long)d->N * Hp * Wp * d->Cin * (long long)dt_size(gdt);
}

int affgw_conv2d_dgrad(const void* dy, const void* wt, const void* x, void* dx, void* workspace,
                       const affgw_conv_desc* d, void* stream) {
    return affgw_conv2d_dgrad_scaled(dy, wt, x, dx, workspace, d, nullptr, stream);
}
int affgw_conv2d_dgrad_scaled(const void* dy, const void* wt, const void* x, void* dx, void* workspace,
                              const affgw_conv_desc* d, const float* inv_scale_dev, void* stream) {
    AFFGW_CHECK(d && dy && wt && dx, "conv2d_dgrad: null pointer");
    affgw_conv_desc dd;
    bool direct;
    int Hp, Wp;
    if (int rc = make_dgrad(d, dd, direct, Hp, Wp)) return rc;
    ConvGeom g;
    dgrad_geom(d, dd, g);
    void* out = direct ? dx : workspace;
    AFFGW_CHECK(out != nullptr, "conv2d_dgrad: workspace required for this geometry");
    const bool tc = d->algo == AFFGW_ALGO_TCGEN05;
    const int gdt = tc ? d->grad_dtype : d->x_dtype;
    int rc;
    if (tc) {
        AFFGW_CHECK(passes_ok(d->passes) && dt_ok(d->grad_dtype), "conv2d_dgrad: bad passes / grad_dtype");
        AFFGW_CHECK(d->y_dtype == AFFGW_BF16 && d->w_dtype == AFFGW_BF16, "conv2d_dgrad (tcgen05): operands are bf16 planes / tiles");
        if (affgw_conv_tc_layout(d, 1) == AFFGW_WLAYOUT_SHIFT) {
            // position space of the FORWARD convolution: dV[p] = sum_taps dYp[p - (K-1)(Wp+1) + tap] . Wflip[tap]
            AFFGW_CHECK(gdt == AFFGW_F32, "conv2d_dgrad: the position-space kernel writes fp32");
            ConvGeom gf;
            PosFrame fx, fy;
            if (int rc2 = shift_frames(d, gf, fx, fy)) return rc2;
            const int qs = -(d->KH - 1) * (fy.Wp + 1);
            if (direct)
                rc = conv_pos_tc(dy, fy, wt, nullptr, nullptr, dx, d->KH, qs, d->pad, d->pad, d->H, d->W, d->Cin, d->Cin, ACT_NONE,
                                 d->passes, d->operand_fmt, inv_scale_dev, S(stream));
            else
                rc = conv_pos_tc(dy, fy, wt, nullptr, nullptr, workspace, d->KH, qs, 0, 0, Hp, Wp, d->Cin, d->Cin, ACT_NONE,
                                 d->passes, d->operand_fmt, inv_scale_dev, S(stream));
        } else {
            AFFGW_CHECK(conv_tc_ok(g) > 0, "conv2d_dgrad: shape not supported by the tcgen05 kernel");
            const long long plane = (long long)dd.N * dd.H * dd.W * dd.in_pitch;
            rc = conv_fwd_tc(dy, plane, wt, nullptr, nullptr, out, gdt, g, d->passes, d->operand_fmt, inv_scale_dev, S(stream));
        }
    } else {
        rc = conv_fwd_simt(dy, dd.x_dtype, wt, d->w_dtype, nullptr, nullptr, out, dd.y_dtype, g, S(stream));
    }
    if (rc) return rc;
    if (!direct) {
        AFFGW_CHECK(d->pre_act == ACT_NONE || x != nullptr, "conv2d_dgrad: x needed for the pre-activation derivative");
        AFFGW_CHECK(tc || d->in_pitch == d->Cin, "conv2d_dgrad: folded path needs a dense x");
        return conv_fold(workspace, x, dx, gdt, d->N, d->H, d->W, d->Cin, d->pad, d->pad_mode, d->upsample, d->pre_act,
                         S(stream));
    }
    return 0;
}

// tensor-core wgrad sees the STORED channel counts of the x / dY planes; padded channels are dropped when unpacking
static int make_wgrad_tc_geom(const affgw_conv_desc* d, ConvGeom& g) {
    if (int rc = make_geom(d, g)) return rc;
    g.Cin = d->in_pitch;
    g.Ktot = g.KH * g.KW * g.Cin;
    g.pre_act = ACT_NONE;       // already applied to the planes by affgw_split_planes
    return 0;
}

long long affgw_conv2d_wgrad_ws_bytes(const affgw_conv_desc* d) {
    ConvGeom g;
    if (!d || d->algo != AFFGW_ALGO_TCGEN05 || make_wgrad_tc_geom(d, g)) return 0;
    if (affgw_conv_tc_layout(d, 0) == AFFGW_WLAYOUT_SHIFT) return conv_wgrad_tc_ws_bytes(g);
    return conv_wgrad_tc_ok(g) ? conv_wgrad_tc_ws_bytes(g) : 0;
}

int affgw_conv2d_wgrad(const void* x, const void* dy, float* dw, void* workspace, const affgw_conv_desc* d, void* stream) {
    return affgw_conv2d_wgrad_scaled(x, dy, dw, workspace, d, nullptr, stream);
}
int affgw_conv2d_wgrad_scaled(const void* x, const void* dy, float* dw, void* workspace, const affgw_conv_desc* d,
                              const float* inv_scale_dev, void* stream) {
    ConvGeom g;
    if (int rc = make_geom(d, g)) return rc;
    AFFGW_CHECK(x && dy && dw, "conv2d_wgrad: null pointer");
    if (d->algo == AFFGW_ALGO_TCGEN05) {
        if (int rc = make_wgrad_tc_geom(d, g)) return rc;
        AFFGW_CHECK(passes_ok(d->passes), "conv2d_wgrad: passes must be 1 or 3");
        AFFGW_CHECK(d->x_dtype == AFFGW_BF16 && d->y_dtype == AFFGW_BF16, "conv2d_wgrad (tcgen05): operands are bf16 planes");
        AFFGW_CHECK(workspace != nullptr, "conv2d_wgrad: the tcgen05 kernel needs affgw_conv2d_wgrad_ws_bytes() of workspace");
        if (affgw_conv_tc_layout(d, 0) == AFFGW_WLAYOUT_SHIFT) {
            ConvGeom gf;
            PosFrame fx, fy;
            if (int rc2 = shift_frames(d, gf, fx, fy)) return rc2;
            return conv_wgrad_pos(x, fx, dy, fy, dw, workspace, g, d->Cin, d->passes, d->operand_fmt, inv_scale_dev, S(stream));
        }
        AFFGW_CHECK(conv_wgrad_tc_ok(g), "conv2d_wgrad: shape not supported by the tcgen05 kernel");
        const long long xpl = (long long)d->N * d->H * d->W * d->in_pitch;
        const long long ypl = g.M * d->out_pitch;
        return conv_wgrad_tc(x, xpl, dy, ypl, dw, workspace, g, d->Cin, d->passes, d->operand_fmt, inv_scale_dev, S(stream));
    }
    return conv_wgrad_simt(x, d->x_dtype, dy, d->y_dtype, dw, g, S(stream));
}

// ---- single-channel-sided stencils on the CUDA cores (conv_thin.cu): fp32 NHWC tensors and the OIHW parameter directly
static int thin_geom(const affgw_conv_desc* d, ConvGeom& g) {
    if (int rc = make_geom(d, g)) return rc;
    AFFGW_CHECK(d->x_dtype == AFFGW_F32 && d->y_dtype == AFFGW_F32, "conv_thin: fp32 tensors only");
    AFFGW_CHECK(d->KH == d->KW && d->in_pitch == d->Cin && d->out_pitch == d->Cout, "conv_thin: dense square-filter convolution expected");
    AFFGW_CHECK(d->pre_act == ACT_NONE, "conv_thin: no activation-first variant");
    AFFGW_CHECK(conv_thin_ok(d->Cin, d->Cout, d->KH, d->stride_w > d->stride ? d->stride_w : d->stride, d->upsample) != 0,
                "conv_thin: %d -> %d channels, %dx%d, stride %d is not a single-channel-sided stencil", d->Cin, d->Cout, d->KH,
                d->KW, d->stride);
    return 0;
}
int affgw_conv_thin_supported(const affgw_conv_desc* d) {
    ConvGeom g;
    if (!d || d->x_dtype != AFFGW_F32 || d->y_dtype != AFFGW_F32 || d->KH != d->KW || d->in_pitch != d->Cin ||
        d->out_pitch != d->Cout || d->pre_act != ACT_NONE || make_geom(d, g))
        return 0;
    return conv_thin_ok(d->Cin, d->Cout, d->KH, d->stride_w > d->stride ? d->stride_w : d->stride, d->upsample);
}
static long long thin_scratch_bytes(const affgw_conv_desc* d) {
    return (long long)d->KH * d->KW * (d->Cin > d->Cout ? d->Cin : d->Cout) * 4;
}
long long affgw_conv_thin_ws_bytes(const affgw_conv_desc* d, int for_dgrad) {
    if (!affgw_conv_thin_supported(d)) return -1;
    const long long frame = (long long)d->N * (d->H + 2 * d->pad) * (d->W + 2 * d->pad) * d->Cin * 4;
    return thin_scratch_bytes(d) + (for_dgrad ? (frame + 255) / 256 * 256 : 0);
}
int affgw_conv_thin_fwd(const float* x, const float* w_oihw, const float* bias, float* y, void* workspace,
                        const affgw_conv_desc* d, void* stream) {
    ConvGeom g;
    if (int rc = thin_geom(d, g)) return rc;
    AFFGW_CHECK(x && w_oihw && y && workspace, "conv_thin_fwd: null pointer");
    return conv_thin_fwd(x, w_oihw, bias, y, (float*)workspace, g, S(stream));
}
int affgw_conv_thin_dgrad(const float* dy, const float* w_oihw, float* dx, void* workspace, const affgw_conv_desc* d, void* stream) {
    ConvGeom g;
    if (int rc = thin_geom(d, g)) return rc;
    AFFGW_CHECK(dy && w_oihw && dx && workspace, "conv_thin_dgrad: null pointer");
    const long long frame = (long long)d->N * (d->H + 2 * d->pad) * (d->W + 2 * d->pad) * d->Cin * 4;
    float* dframe = (float*)workspace;
    float* scratch = (float*)((char*)workspace + (frame + 255) / 256 * 256);
    if (int rc = conv_thin_dgrad_frame(dy, w_oihw, dframe, scratch, g, S(stream))) return rc;
    return conv_fold(dframe, nullptr, dx, AFFGW_F32, d->N, d->H, d->W, d->Cin, d->pad, d->pad_mode, 1, ACT_NONE, S(stream));
}
int affgw_conv_thin_wgrad(const float* x, const float* dy, float* dw_oihw, const affgw_conv_desc* d, void* stream) {
    ConvGeom g;
    if (int rc = thin_geom(d, g)) return rc;
    AFFGW_CHECK(x && dy && dw_oihw, "conv_thin_wgrad: null pointer");
    return conv_thin_wgrad(x, dy, dw_oihw, g, S(stream));
}

int affgw_colsum(const void* a, int dtype, float* out, long long M, int C, int pitch, void* stream) {
    AFFGW_CHECK(a && out && dt_ok(dtype) && M > 0 && C > 0 && pitch >= C, "colsum: bad argument");
    return colsum(a, dtype, out, M, C, pitch, S(stream));
}

int affgw_norm_stats(const void* x, int dtype, float* ws, float* mean, float* rstd, float* var_unbiased, int G, long long P,
                     int C, float eps, int unbiased, void* stream) {
    AFFGW_CHECK(x && ws && mean && rstd && dt_ok(dtype) && G > 0 && P > 0 && C > 0, "norm_stats: bad argument");
    return norm_stats(x, dtype, ws, mean, rstd, var_unbiased, G, P, C, eps, unbiased, S(stream));
}
int affgw_norm_apply(const void* x, int dtype, const float* mean, const float* rstd, const float* gamma, const float* beta,
                     const void* residual, void* y, int G, long long P, int C, int act, int affine_per_group, void* stream) {
    AFFGW_CHECK(x && y && mean && rstd && dt_ok(dtype) && G > 0 && P > 0 && C > 0, "norm_apply: bad argument");
    AFFGW_CHECK((gamma == nullptr) == (beta == nullptr), "norm_apply: gamma and beta must be given together");
    return norm_apply(x, dtype, mean, rstd, gamma, beta, residual, y, G, P, C, act, affine_per_group, S(stream));
}
int affgw_norm_bwd(const void* dy, const void* x, int dtype, const float* mean, const float* rstd, const float* gamma,
                   const float* beta, float* s1, float* s2, void* dx, int G, long long P, int C, int act,
                   int affine_per_group, int batch_stats, int unbiased, void* stream) {
    AFFGW_CHECK(dy && x && dx && mean && rstd && s1 && s2 && dt_ok(dtype) && G > 0 && P > 0 && C > 0, "norm_bwd: bad argument");
    AFFGW_CHECK(act != ACT_TANH, "norm_bwd: tanh is not a fused norm activation");
    return norm_bwd(dy, x, dtype, mean, rstd, gamma, beta, s1, s2, dx, G, P, C, act, affine_per_group, batch_stats, unbiased, S(stream));
}
int affgw_bn_update_running(float* rm, float* rv, long long* nbt, const float* mean, const float* var_unbiased, int C,
                            float momentum, void* stream) {
    AFFGW_CHECK(rm && rv && mean && var_unbiased && C > 0, "bn_update_running: bad argument");
    return bn_update_running(rm, rv, nbt, mean, var_unbiased, C, momentum, S(stream));
}
int affgw_bn_eval_stats(const float* rm, const float* rv, float* mean, float* rstd, int C, float eps, void* stream) {
    AFFGW_CHECK(rm && rv && mean && rstd && C > 0, "bn_eval_stats: bad argument");
    return bn_eval_stats(rm, rv, mean, rstd, C, eps, S(stream));
}

#define REQ(cond, name) AFFGW_CHECK(cond, name ": bad argument")
int affgw_maxpool2_fwd(const void* x, void* y, int dt, int N, int H, int W, int C, void* s) {
    REQ(x && y && dt_ok(dt) && N > 0 && H > 1 && W > 1 && C > 0, "maxpool2_fwd");
    return maxpool2_fwd(x, y, dt, N, H, W, C, S(s));
}
int affgw_maxpool2_bwd(const void* dy, const void* x, void* dx, int dt, int N, int H, int W, int C, void* s) {
    REQ(dy && x && dx && dt_ok(dt) && N > 0 && H > 1 && W > 1 && C > 0, "maxpool2_bwd");
    return maxpool2_bwd(dy, x, dx, dt, N, H, W, C, S(s));
}
int affgw_avgpool3s2_fwd(const void* x, void* y, int dt, int N, int H, int W, int C, void* s) {
    REQ(x && y && dt_ok(dt) && N > 0 && H > 1 && W > 1 && C > 0, "avgpool3s2_fwd");
    return avgpool3s2_fwd(x, y, dt, N, H, W, C, S(s));
}
int affgw_avgpool3s2_bwd(const void* dy, void* dx, int dt, int N, int H, int W, int C, void* s) {
    REQ(dy && dx && dt_ok(dt) && N > 0 && H > 1 && W > 1 && C > 0, "avgpool3s2_bwd");
    return avgpool3s2_bwd(dy, dx, dt, N, H, W, C, S(s));
}
int affgw_resize_nearest_fwd(const void* x, void* y, int dt, int N, int H, int W, int C, int Ho, int Wo, void* s) {
    REQ(x && y && dt_ok(dt) && N > 0 && H > 0 && W > 0 && C > 0 && Ho > 0 && Wo > 0, "resize_nearest_fwd");
    return resize_nearest_fwd(x, y, dt, N, H, W, C, Ho, Wo, S(s));
}
int affgw_resize_nearest_bwd(const void* dy, void* dx, int dt, int N, int H, int W, int C, int Ho, int Wo, void* s) {
    REQ(dy && dx && dt_ok(dt) && N > 0 && H > 0 && W > 0 && C > 0 && Ho > 0 && Wo > 0, "resize_nearest_bwd");
    return resize_nearest_bwd(dy, dx, dt, N, H, W, C, Ho, Wo, S(s));
}
int affgw_gate_fwd(const void* x, const void* r, const void* xl, const void* xg, void* y, int dt, int N, long long P, int C,
                   void* s) {
    REQ(x && r && xl && xg && y && dt_ok(dt) && N > 0 && P > 0 && C > 0, "gate_fwd");
    return gate_fwd(x, r, xl, xg, y, dt, N, P, C, S(s));
}
int affgw_gate_bwd(const void* dy, const void* x, const void* r, const void* xl, const void* xg, void* dx, void* dr,
                   void* dxl, void* dxg, int dt, int N, long long P, int C, void* s) {
    REQ(dy && x && r && xl && xg && dx && dr && dxl && dxg && dt_ok(dt) && N > 0 && P > 0 && C > 0, "gate_bwd");
    return gate_bwd(dy, x, r, xl, xg, dx, dr, dxl, dxg, dt, N, P, C, S(s));
}
int affgw_gap_fwd(const void* x, void* out, int dt, int N, long long P, int C, void* s) {
    REQ(x && out && dt_ok(dt) && N > 0 && P > 0 && C > 0, "gap_fwd");
    return gap_fwd(x, out, dt, N, P, C, S(s));
}
int affgw_bcast_add(const void* a, const void* v, void* out, int dt, int N, long long P, int C, float scale, void* s) {
    REQ(v && out && dt_ok(dt) && N > 0 && P > 0 && C > 0, "bcast_add");
    return bcast_add(a, v, out, dt, N, P, C, scale, S(s));
}
int affgw_maxpool3s2_fwd(const void* x, void* y, int dt, int N, int H, int W, int C, void* s) {
    REQ(x && y && dt_ok(dt) && N > 0 && H > 1 && W > 1 && C > 0, "maxpool3s2_fwd");
    return maxpool3_fwd(x, y, dt, N, H, W, C, 2, 2, S(s));
}
int affgw_maxpool3s2_bwd(const void* dy, const void* x, void* dx, int dt, int N, int H, int W, int C, void* s) {
    REQ(dy && x && dx && dt_ok(dt) && N > 0 && H > 1 && W > 1 && C > 0, "maxpool3s2_bwd");
    return maxpool3_bwd(dy, x, dx, dt, N, H, W, C, 2, 2, S(s));
}
int affgw_maxpool3_fwd(const void* x, void* y, int dt, int N, int H, int W, int C, int sy, int sx, void* s) {
    REQ(x && y && dt_ok(dt) && N > 0 && H > 1 && W > 1 && C > 0 && (sy == 1 || sy == 2) && (sx == 1 || sx == 2), "maxpool3_fwd");
    return maxpool3_fwd(x, y, dt, N, H, W, C, sy, sx, S(s));
}
int affgw_maxpool3_bwd(const void* dy, const void* x, void* dx, int dt, int N, int H, int W, int C, int sy, int sx, void* s) {
    REQ(dy && x && dx && dt_ok(dt) && N > 0 && H > 1 && W > 1 && C > 0 && (sy == 1 || sy == 2) && (sx == 1 || sx == 2), "maxpool3_bwd");
    return maxpool3_bwd(dy, x, dx, dt, N, H, W, C, sy, sx, S(s));
}
int affgw_resize_bilinear_fwd(const void* x, void* y, int dt, int N, int H, int W, int C, int Ho, int Wo, void* s) {
    REQ(x && y && dt_ok(dt) && N > 0 && H > 0 && W > 0 && C > 0 && Ho > 0 && Wo > 0, "resize_bilinear_fwd");
    return resize_bilinear_fwd(x, y, dt, N, H, W, C, Ho, Wo, S(s));
}
int affgw_resize_bilinear_bwd(const void* dy, float* dx, int dt, int N, int H, int W, int C, int Ho, int Wo, void* s) {
    REQ(dy && dx && dt_ok(dt) && N > 0 && H > 0 && W > 0 && C > 0 && Ho > 0 && Wo > 0, "resize_bilinear_bwd");
    return resize_bilinear_bwd(dy, dx, dt, N, H, W, C, Ho, Wo, S(s));
}
int affgw_add_act(const void* a, const void* b, void* out, int dt, long long n, int act, void* s) {
    REQ(a && b && out && dt_ok(dt) && n > 0 && act >= 0 && act <= 3, "add_act");
    return add_act(a, b, out, dt, n, act, S(s));
}
int affgw_add2(const void* a, const void* b, void* out, int dt, long long n, void* s) {
    REQ(a && b && out && dt_ok(dt) && n > 0, "add2");
    return add2(a, b, out, dt, n, S(s));
}
int affgw_act_bwd(const void* dy, const void* y, void* dz, int dt, long long n, int act, void* s) {
    REQ(dy && y && dz && dt_ok(dt) && n > 0 && act >= 0 && act <= 3, "act_bwd");
    return act_bwd(dy, y, dz, dt, n, act, S(s));
}
int affgw_embedding_fwd(const long long* ids, const float* table, void* out, int dt, long long n_ids, int E, int V, int* err,
                        void* s) {
    REQ(ids && table && out && err && dt_ok(dt) && n_ids > 0 && E > 0 && V > 0, "embedding_fwd");
    return embedding_fwd(ids, table, out, dt, n_ids, E, V, err, S(s));
}
int affgw_embedding_bwd(const long long* ids, const void* dout, float* dtable, int dt, long long n_ids, int E, int V, void* s) {
    REQ(ids && dout && dtable && dt_ok(dt) && n_ids > 0 && E > 0 && V > 0, "embedding_bwd");
    return embedding_bwd(ids, dout, dtable, dt, n_ids, E, V, S(s));
}
int affgw_text_tile_fwd(const void* chars, void* out, int dt, int B, int H, int W, int C, int ts, int reps, void* s) {
    REQ(chars && out && dt_ok(dt) && B > 0 && H > 0 && W > 0 && C > 0 && ts > 0 && reps > 0, "text_tile_fwd");
    return text_tile_fwd(chars, out, dt, B, H, W, C, ts, reps, S(s));
}
int affgw_text_tile_bwd(const void* dout, void* dchars, int dt, int B, int H, int W, int C, int ts, int reps, void* s) {
    REQ(dout && dchars && dt_ok(dt) && B > 0 && H > 0 && W > 0 && C > 0 && ts > 0 && reps > 0, "text_tile_bwd");
    return text_tile_bwd(dout, dchars, dt, B, H, W, C, ts, reps, S(s));
}
int affgw_bce_logits_fwd(const void* x, int dt, float target, float* loss, long long n, void* s) {
    REQ(x && loss && dt_ok(dt) && n > 0, "bce_logits_fwd");
    return bce_logits_fwd(x, dt, target, loss, n, S(s));
}
int affgw_bce_logits_bwd(const void* x, int dt, float target, const float* gout, void* dx, long long n, void* s) {
    REQ(x && gout && dx && dt_ok(dt) && n > 0, "bce_logits_bwd");
    return bce_logits_bwd(x, dt, target, gout, dx, n, S(s));
}
int affgw_softmax_ce_fwd(const void* x, int dt, const long long* y, float* loss, int B, int C, int* err, void* s) {
    REQ(x && y && loss && err && dt_ok(dt) && B > 0 && C > 0, "softmax_ce_fwd");
    return softmax_ce_fwd(x, dt, y, loss, B, C, err, S(s));
}
int affgw_softmax_ce_bwd(const void* x, int dt, const long long* y, const float* gout, void* dx, int B, int C, void* s) {
    REQ(x && y && gout && dx && dt_ok(dt) && B > 0 && C > 0, "softmax_ce_bwd");
    return softmax_ce_bwd(x, dt, y, gout, dx, B, C, S(s));
}
int affgw_gru_cell_fwd(const float* gi, long long gi_pitch, const float* gh, const float* h, float* h_out, int N, int H, void* s) {
    REQ(gi && gh && h && h_out && N > 0 && H > 0 && gi_pitch >= 3LL * H, "gru_cell_fwd");
    return gru_cell_fwd(gi, gi_pitch, gh, h, h_out, N, H, S(s));
}
int affgw_gru_cell_bwd(const float* dh_out, const float* gi, long long gi_pitch, const float* gh, const float* h, float* dgi,
                       float* dgh, float* dh, int N, int H, void* s) {
    REQ(dh_out && gi && gh && h && dgi && dgh && dh && N > 0 && H > 0 && gi_pitch >= 3LL * H, "gru_cell_bwd");
    return gru_cell_bwd(dh_out, gi, gi_pitch, gh, h, dgi, dgh, dh, N, H, S(s));
}
int affgw_scale_nc(const float* x, const float* m, float* y, int N, long long P, int C, void* s) {
    REQ(x && m && y && N > 0 && P > 0 && C > 0, "scale_nc");
    return scale_nc(x, m, y, N, P, C, S(s));
}
int affgw_mul2(const float* a, const float* b, float* y, long long n, void* s) {
    REQ(a && b && y && n > 0, "mul2");
    return mul2(a, b, y, n, S(s));
}
int affgw_map_seq(const float* src, float* dst, int B, int H, int W, int C, int to_seq, void* s) {
    REQ(src && dst && B > 0 && H > 0 && W > 0 && C > 0, "map_seq");
    return map_seq(src, dst, B, H, W, C, to_seq, S(s));
}
int affgw_attn_energy_fwd(const float* e, const long long* sample, const float* hp, const float* loc, const float* v, const float* vb,
                          float* energy, int N, int T, int F, void* s) {
    REQ(e && sample && hp && loc && v && vb && energy && N > 0 && T > 0 && F > 0, "attn_energy_fwd");
    return attn_energy_fwd(e, sample, hp, loc, v, vb, energy, N, T, F, S(s));
}
int affgw_attn_energy_bwd(const float* denergy, const float* e, const long long* sample, const float* hp, const float* loc,
                          const float* v, float* de, float* dhp, float* dloc, float* dv, float* dvb, int N, int T, int F, void* s) {
    REQ(denergy && e && sample && hp && loc && v && de && dhp && dloc && dv && dvb && N > 0 && T > 0 && F > 0, "attn_energy_bwd");
    return attn_energy_bwd(denergy, e, sample, hp, loc, v, de, dhp, dloc, dv, dvb, N, T, F, S(s));
}
int affgw_attn_ctx_fwd(const float* energy, const float* enc, const long long* sample, float* attn, float* ctx, int N, int T, int F,
                       void* s) {
    REQ(energy && enc && sample && attn && ctx && N > 0 && T > 0 && F > 0, "attn_ctx_fwd");
    return attn_ctx_fwd(energy, enc, sample, attn, ctx, N, T, F, S(s));
}
int affgw_attn_ctx_bwd(const float* dattn, const float* dctx, const float* attn, const float* enc, const long long* sample,
                       float* denergy, float* denc, int N, int T, int F, void* s) {
    REQ(dctx && attn && enc && sample && denergy && denc && N > 0 && T > 0 && F > 0, "attn_ctx_bwd");
    return attn_ctx_bwd(dattn, dctx, attn, enc, sample, denergy, denc, N, T, F, S(s));
}
int affgw_layernorm_fwd(const float* x, const float* w, const float* b, float* y, long long rows, int D, float eps, void* s) {
    REQ(x && w && b && y && rows > 0 && D > 0 && eps >= 0.f, "layernorm_fwd");
    return layernorm_fwd(x, w, b, y, rows, D, eps, S(s));
}
int affgw_gelu_fwd(const float* x, float* y, long long n, void* s) {
    REQ(x && y && n > 0, "gelu_fwd");
    return gelu_fwd(x, y, n, S(s));
}
int affgw_scale_residual(const float* x, const float* t, const float* gamma, float* y, long long n, int D, void* s) {
    REQ(x && t && y && n > 0 && D > 0 && n % D == 0, "scale_residual");
    return scale_residual(x, t, gamma, y, n, D, S(s));
}
int affgw_attention_fwd(const float* qkv, float* out, int B, int N, int H, int hd, float scale, void* s) {
    REQ(qkv && out && B > 0 && N > 0 && H > 0 && hd > 0, "attention_fwd");
    return attention_fwd(qkv, out, B, N, H, hd, scale, S(s));
}
int affgw_blur3(const float* x, float* y, int N, int H, int W, int C, void* s) {
    REQ(x && y && x != y && N > 0 && H > 0 && W > 0 && C > 0, "blur3");
    return blur3(x, y, N, H, W, C, S(s));
}
int affgw_pixelnorm(const float* x, float* y, int rows, int C, float eps, void* s) {
    REQ(x && y && rows > 0 && C > 0 && eps >= 0.f, "pixelnorm");
    return pixelnorm(x, y, rows, C, eps, S(s));
}
int affgw_label_smooth_kl_fwd(const float* x, const long long* y, float* loss, int rows, int V, int pad_idx, float smoothing,
                              int* err, void* s) {
    REQ(x && y && loss && err && rows > 0 && V > 2 && pad_idx >= 0 && pad_idx < V && smoothing >= 0.f && smoothing < 1.f,
        "label_smooth_kl_fwd");
    return label_smooth_kl_fwd(x, y, loss, rows, V, pad_idx, smoothing, err, S(s));
}
int affgw_label_smooth_kl_bwd(const float* x, const long long* y, const float* gout, float* dx, int rows, int V, int pad_idx,
                              float smoothing, void* s) {
    REQ(x && y && gout && dx && rows > 0 && V > 2 && pad_idx >= 0 && pad_idx < V, "label_smooth_kl_bwd");
    return label_smooth_kl_bwd(x, y, gout, dx, rows, V, pad_idx, smoothing, S(s));
}
int affgw_u8_to_image(const unsigned char* src, float* dst, long long n, void* s) {
    REQ(src && dst && n > 0, "u8_to_image");
    REQ(((uintptr_t)src & 15) == 0 && ((uintptr_t)dst & 15) == 0, "u8_to_image: buffers must be 16-byte aligned");
    return u8_to_image(src, dst, n, S(s));
}
int affgw_nchw_to_nhwc(const float* x, void* y, int dt, int N, int C, long long HW, int c_pad, void* s) {
    REQ(x && y && dt_ok(dt) && N > 0 && C > 0 && HW > 0 && c_pad >= C, "nchw_to_nhwc");
    return nchw_to_nhwc(x, y, dt, N, C, HW, c_pad, S(s));
}
int affgw_nhwc_to_nchw(const void* x, float* y, int dt, int N, int C, long long HW, int c_pad, void* s) {
    REQ(x && y && dt_ok(dt) && N > 0 && C > 0 && HW > 0 && c_pad >= C, "nhwc_to_nchw");
    return nhwc_to_nchw(x, y, dt, N, C, HW, c_pad, S(s));
}
int affgw_concat_channels(const void* a, const void* b, void* out, int dt, long long rows, int ca, int cb, void* s) {
    REQ(a && b && out && dt_ok(dt) && rows > 0 && ca > 0 && cb > 0, "concat_channels");
    return concat2(const_cast<void*>(a), const_cast<void*>(b), out, dt, rows, ca, cb, 1, S(s));
}
int affgw_split_channels(const void* in, void* a, void* b, int dt, long long rows, int ca, int cb, void* s) {
    REQ(a && b && in && dt_ok(dt) && rows > 0 && ca > 0 && cb > 0, "split_channels");
    return concat2(a, b, const_cast<void*>(in), dt, rows, ca, cb, 0, S(s));
}
int affgw_cast(const void* x, int in_dt, void* y, int out_dt, long long n, void* s) {
    REQ(x && y && dt_ok(in_dt) && dt_ok(out_dt) && n > 0, "cast");
    return cast_dtype(x, in_dt, y, out_dt, n, S(s));
}


}  // extern "C"

// ---- gradient bucket pack / unpack -------------------------------------------------------------------------------
// The bucket is cut into spans of 4096 elements; a block finds the tensor that covers the start of its span by binary search in
// the (dense, increasing) offset table and walks on from there, so the copy is coalesced on both sides and the grid is sized by
// bytes, not by tensor count (one 9 MB weight gradient used to be copied by 32 blocks).
namespace {
template <bool UNPACK>
__global__ void __launch_bounds__(256)
bucket_copy_kernel(float* const* __restrict__ ptrs, const long long* __restrict__ sizes, const long long* __restrict__ offsets, int n,
                   float* __restrict__ bucket, float scale) {
    constexpr int SPAN = 256 * 16;
    const long long total = offsets[n - 1] + sizes[n - 1];
    for (long long s0 = blockIdx.x * (long long)SPAN; s0 < total; s0 += (long long)gridDim.x * SPAN) {
        int lo = 0, hi = n - 1;
        while (lo < hi) {                               // last tensor whose offset is <= s0
            const int mid = (lo + hi + 1) >> 1;
            if (offsets[mid] <= s0) lo = mid; else hi = mid - 1;
        }
        int t = lo;
        long long toff = offsets[t], tend = toff + sizes[t];
        float* p = ptrs[t];
#pragma unroll 4
        for (int k = 0; k < 16; ++k) {
            const long long i = s0 + k * 256 + threadIdx.x;
            if (i >= total) break;
            while (i >= tend) {
                ++t;
                toff = offsets[t];
                tend = toff + sizes[t];
                p = ptrs[t];
            }
            if (UNPACK) p[i - toff] = bucket[i] * scale;
            else bucket[i] = p[i - toff];
        }
    }
}
}  // namespace

extern "C" int affgw_bucket_pack(const float* const* ptrs, const long long* sizes, const long long* offsets, int n,
                                 float* bucket, void* stream) {
    AFFGW_CHECK(ptrs && sizes && offsets && bucket && n > 0 && n <= 65535, "bucket_pack: bad argument");
    bucket_copy_kernel<false><<<148 * 8, 256, 0, S(stream)>>>(const_cast<float* const*>(ptrs), sizes, offsets, n, bucket, 1.f);
    AFFGW_LAUNCH_CHECK("bucket_pack");
    return 0;
}
extern "C" int affgw_bucket_unpack(float* const* ptrs, const long long* sizes, const long long* offsets, int n,
                                   const float* bucket, float scale, void* stream) {
    AFFGW_CHECK(ptrs && sizes && offsets && bucket && n > 0 && n <= 65535, "bucket_unpack: bad argument");
    bucket_copy_kernel<true><<<148 * 8, 256, 0, S(stream)>>>(ptrs, sizes, offsets, n, const_cast<float*>(bucket), scale);
    AFFGW_LAUNCH_CHECK("bucket_unpack");
    return 0;
}


// ---- multi-tensor Adam (torch.optim.Adam, main_run.py:275-278: default betas / eps, no weight decay, no amsgrad) ----------
// One launch steps every tensor of a table: the element range is cut into 4096-element spans like the gradient buckets; the
// bias corrections of this step come from the host (all tensors of a call share the step count).
//   m = m + (g - m)(1 - b1);  v = b2 v + (1 - b2) g^2;  p -= (lr / bc1) * m / (sqrt(v) / sqrt(bc2) + eps)
namespace {
__global__ void __launch_bounds__(256)
adam_step_kernel(float* const* __restrict__ params, const float* const* __restrict__ grads, float* const* __restrict__ exp_avg,
                 float* const* __restrict__ exp_avg_sq, const long long* __restrict__ sizes, const long long* __restrict__ offsets,
                 int n, float step_size, float beta1, float beta2, float eps, float inv_bc2_sqrt, float grad_scale) {
    constexpr int SPAN = 256 * 16;
    const long long total = offsets[n - 1] + sizes[n - 1];
    for (long long s0 = blockIdx.x * (long long)SPAN; s0 < total; s0 += (long long)gridDim.x * SPAN) {
        int lo = 0, hi = n - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (offsets[mid] <= s0) lo = mid; else hi = mid - 1;
        }
        int t = lo;
        long long toff = offsets[t], tend = toff + sizes[t];
#pragma unroll 4
        for (int k = 0; k < 16; ++k) {
            const long long i = s0 + k * 256 + threadIdx.x;
            if (i >= total) break;
            while (i >= tend) {
                ++t;
                toff = offsets[t];
                tend = toff + sizes[t];
            }
            const long long j = i - toff;
            const float g = grads[t][j] * grad_scale;
            float m = exp_avg[t][j], v = exp_avg_sq[t][j];
            m = fmaf(g - m, 1.f - beta1, m);
            v = fmaf(v, beta2, (1.f - beta2) * g * g);
            exp_avg[t][j] = m;
            exp_avg_sq[t][j] = v;
            const float denom = sqrtf(v) * inv_bc2_sqrt + eps;
            params[t][j] -= step_size * (m / denom);
        }
    }
}
}  // namespace

extern "C" int affgw_adam_step(float* const* params, const float* const* grads, float* const* exp_avg, float* const* exp_avg_sq,
                               const long long* sizes, const long long* offsets, int n, float lr, float beta1, float beta2,
                               float eps, long long step, float grad_scale, void* stream) {
    AFFGW_CHECK(params && grads && exp_avg && exp_avg_sq && sizes && offsets && n > 0 && n <= 65535, "adam_step: bad argument");
    AFFGW_CHECK(step >= 1 && lr >= 0.f && beta1 >= 0.f && beta1 < 1.f && beta2 >= 0.f && beta2 < 1.f && eps >= 0.f,
                "adam_step: bad hyper-parameter");
    const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
    adam_step_kernel<<<148 * 8, 256, 0, S(stream)>>>(params, grads, exp_avg, exp_avg_sq, sizes, offsets, n, (float)(lr / bc1), beta1,
                                                      beta2, eps, (float)(1.0 / sqrt(bc2)), grad_scale);
    AFFGW_LAUNCH_CHECK("adam_step");
    return 0;
}
