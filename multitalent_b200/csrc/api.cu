// extern "C" surface of libmtb200.so (see include/mtb200.h).  Thin argument checks + dispatch; no device memory is
// owned here and nothing throws or aborts across the boundary.
#include <stdarg.h>
#include <stdlib.h>

#include "common.cuh"

namespace mtb {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static thread_local const char* g_last_kernel = "";

int check_launch(const char* what) {
  g_last_kernel = what;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: CUDA error %d (%s)", what, (int)e, cudaGetErrorString(e));
    return MTB200_ERR_CUDA;
  }
  return MTB200_OK;
}

bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("MTB200_PDL"); return !e || atoi(e) != 0; }();
  return on;
}

int num_sms() {
  static thread_local int cached_dev = -1, cached = 148;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) cached = n;
    cached_dev = dev;
  }
  return cached;
}

// implemented in the other translation units
int conv_taps_ffma(const mtb200_conv_params& p, cudaStream_t s);
int wgrad_taps_ffma(const mtb200_wgrad_params& p, cudaStream_t s);
int conv_taps_umma(const mtb200_conv_params& p, cudaStream_t s);   // conv_umma.cu
int wgrad_taps_umma(const mtb200_wgrad_params& p, cudaStream_t s); // conv_umma.cu
int umma_available();
int conv_line_umma(const mtb200_conv_params& p, cudaStream_t s);    // conv_line.cu
int wgrad_line_umma(const mtb200_wgrad_params& p, cudaStream_t s);  // wgrad_line.cu
int conv_c1_fwd(const void*, long long, const void*, int, const float*, void*, int, int, int, double*, int, int, int, int,
                int, cudaStream_t);  // conv_c1.cu
int conv_c1_wgrad(const void*, long long, const void*, int, int, int, float*, int, int, int, int, int, int, cudaStream_t);
int colsum(const void* m, int dtype, long long rows, int ldc, int coff, int C, float* out, cudaStream_t s);
int in_finalize(const double*, const float*, const float*, int, int, long long, float, float, float*, float*,
                cudaStream_t);
int in_stats(const void*, int, int, long long, int, int, int, double*, cudaStream_t);
int norm_act(const void*, int, int, void*, int, int, int, int, long long, int, const float*, const void*, int, int,
             const float*, float, cudaStream_t);
int in_bwd_reduce(const void*, int, int, const void*, int, int, int, int, long long, int, const float*, const float*,
                  double*, cudaStream_t);
int in_bwd_apply(const void*, int, int, const void*, int, int, void*, int, int, int, int, long long, int, const float*,
                 const float*, const float*, const double*, float*, float*, cudaStream_t);
int lrelu_bwd(const void*, const void*, void*, int, long long, float, cudaStream_t);
int dcce_stats(const void*, int, int, int, int, const float*, int, long long, double*, double*, cudaStream_t);
int dcce_bwd(const void*, int, int, int, int, const float*, int, long long, const float*, float, const float*, void*, int,
             cudaStream_t);
int residual_bwd(const void*, int, int, const void*, int, int, void*, int, int, int, void*, int, int, int, int, long long, int,
                 float, cudaStream_t);
int mt_loss_stats(const void*, int, int, int, const float*, int, long long, const uint64_t*, const uint64_t*, int,
                  double*, double*, cudaStream_t);
int mt_loss_finalize(const double*, const double*, const uint64_t*, int, int, long long, float, float, float*, float*,
                     cudaStream_t);
int mt_loss_bwd(const void*, int, int, int, const float*, int, long long, const uint64_t*, int, const float*,
                const float*, void*, int, cudaStream_t);
int head_bwd_fused(const mtb200_head_bwd_params& p, cudaStream_t s);  // head_bwd.cu
int head_fwd_stats(const mtb200_head_fwd_params& p, cudaStream_t s);  // head_bwd.cu
int head_aggregate(const mtb200_head_agg_params& p, cudaStream_t s);  // head_bwd.cu
int sw_gather_tile(const float*, int, int, int, int, int, int, int, int, int, int, int, void*, int, int, cudaStream_t);
int sw_aggregate(const void*, int, int, int, int, int, int, int, const float*, float, int, float*, float*, int, int, int,
                 int, int, int, cudaStream_t);
int sw_finalize(float*, const float*, int, long long, long long, const float*, float*, cudaStream_t);
int sumsq(const float*, long long, double*, cudaStream_t);
int sgd_step(float*, const float*, float*, long long, const double*, float, float, float, float, float, int,
             const float*, cudaStream_t);
int loss_scale_update(const double*, float*, float, float, int, cudaStream_t);
int pack_weights(const float*, int, int, int, int, int, void*, int, int, int, int, int, cudaStream_t);
int unpack_wgrad(const float*, int, int, int, int, int, int, int, int, float, int, float*, cudaStream_t);
int pack_weights_batched(const void*, int, int, int, cudaStream_t);
int unpack_wgrad_batched(const void*, int, int, cudaStream_t);
int ncdhw_to_ndhwc(const float*, int, int, long long, void*, int, int, int, int, cudaStream_t);
int ndhwc_to_ncdhw(const void*, int, int, int, int, int, long long, float*, cudaStream_t);
int crop_pad(const float*, int, int, int, int, int, int, int, float*, int, int, int, int, const float*, cudaStream_t);
int resize_nearest(const float*, long long, int, int, int, float*, int, int, int, cudaStream_t);
int resample_probs(const float*, int, int, int, int, int, int, int, int, int, int, void*, int, const float*,
                   unsigned char*, cudaStream_t);

static int validate_taps(int ngroups, const int32_t* begin, int ntaps) {
  MTB_REQUIRE(ngroups >= 1 && ngroups <= MTB200_MAX_GROUPS, "ngroups=%d out of range", ngroups);
  MTB_REQUIRE(ntaps >= 1 && ntaps <= MTB200_MAX_TAPS, "ntaps=%d out of range", ntaps);
  MTB_REQUIRE(begin[0] == 0 && begin[ngroups] == ntaps, "group_tap_begin must span [0, ntaps]");
  for (int g = 0; g < ngroups; ++g) MTB_REQUIRE(begin[g] <= begin[g + 1], "group_tap_begin must be non-decreasing");
  return MTB200_OK;
}

}  // namespace mtb

using namespace mtb;
#define STREAM(s) reinterpret_cast<cudaStream_t>(s)

extern "C" {

int mtb200_version(void) { return MTB200_VERSION; }
const char* mtb200_last_error(void) { return g_err; }
int mtb200_has_tcgen05(void) { return umma_available(); }
const char* mtb200_last_kernel(void) { return g_last_kernel; }

int mtb200_conv_taps(const mtb200_conv_params* p, void* stream) {
  MTB_REQUIRE(p && p->in && p->out && p->w, "conv_taps: null pointer");
  if (int r = validate_taps(p->ngroups, p->group_tap_begin, p->ntaps)) return r;
  MTB_REQUIRE((p->in_split ? (p->in_coff == 0 && p->Cin == 2 * p->in_split && p->in_split <= p->in_ldc)
                           : p->in_coff + p->Cin <= p->in_ldc) &&
                  (p->out_split ? (p->out_coff == 0 && p->Cout == 2 * p->out_split && p->out_split <= p->out_ldc)
                                : p->out_coff + p->Cout <= p->out_ldc),
              "conv_taps: channel slice exceeds ldc (in %d+%d/%d, out %d+%d/%d)", p->in_coff, p->Cin, p->in_ldc,
              p->out_coff, p->Cout, p->out_ldc);
  if (p->in_split || p->out_split) {  // planar halves: the line-streaming tensor-core kernel only
    const int r = (umma_available() && p->dtype != MTB200_F32) ? conv_line_umma(*p, STREAM(stream)) : MTB200_ERR_UNSUPPORTED;
    if (r == MTB200_ERR_UNSUPPORTED) set_error("conv_taps: planar halves (in_split / out_split) outside the line kernel's envelope");
    return r;
  }
  if (p->impl >= 2 || (p->impl == 0 && umma_available() && p->dtype != MTB200_F32)) {
    int r = conv_taps_umma(*p, STREAM(stream));
    if (r != MTB200_ERR_UNSUPPORTED || p->impl >= 2) return r;
  }
  return conv_taps_ffma(*p, STREAM(stream));
}

int mtb200_wgrad_taps(const mtb200_wgrad_params* p, void* stream) {
  MTB_REQUIRE(p && p->x && p->dy && p->dw, "wgrad_taps: null pointer");
  if (int r = validate_taps(p->ngroups, p->group_tap_begin, p->ntaps)) return r;
  if (p->in_split) {  // planar halves: the line-streaming tensor-core kernel only
    const int r = (umma_available() && p->dtype != MTB200_F32) ? wgrad_line_umma(*p, STREAM(stream)) : MTB200_ERR_UNSUPPORTED;
    if (r == MTB200_ERR_UNSUPPORTED) set_error("wgrad_taps: planar halves (in_split) outside the line kernel's envelope");
    return r;
  }
  if (p->impl >= 2 || (p->impl == 0 && umma_available() && p->dtype != MTB200_F32)) {
    int r = wgrad_taps_umma(*p, STREAM(stream));
    if (r != MTB200_ERR_UNSUPPORTED) return r;  // shapes the tensor-core kernel does not cover use the CUDA-core one
  }
  return wgrad_taps_ffma(*p, STREAM(stream));
}

int mtb200_conv_c1_fwd(const void* x, int64_t x_stride, const void* w, int32_t Cin_p, const float* bias, void* out,
                       int32_t out_ldc, int32_t out_coff, int32_t Cout_p, double* stats, int32_t dtype, int32_t B,
                       int32_t D, int32_t H, int32_t W, void* stream) {
  MTB_REQUIRE(x && w && out && x_stride >= 1 && Cin_p >= 1, "conv_c1_fwd: bad arguments");
  MTB_REQUIRE(out_coff + Cout_p <= out_ldc, "conv_c1_fwd: channel slice exceeds ldc");
  return conv_c1_fwd(x, x_stride, w, Cin_p, bias, out, out_ldc, out_coff, Cout_p, stats, dtype, B, D, H, W, STREAM(stream));
}

int mtb200_conv_c1_wgrad(const void* x, int64_t x_stride, const void* dy, int32_t dy_ldc, int32_t dy_coff,
                         int32_t Cout_p, float* dw, int32_t Cin_p, int32_t dtype, int32_t B, int32_t D, int32_t H,
                         int32_t W, void* stream) {
  MTB_REQUIRE(x && dy && dw && x_stride >= 1 && Cin_p >= 1, "conv_c1_wgrad: bad arguments");
  MTB_REQUIRE(dy_coff + Cout_p <= dy_ldc, "conv_c1_wgrad: channel slice exceeds ldc");
  return conv_c1_wgrad(x, x_stride, dy, dy_ldc, dy_coff, Cout_p, dw, Cin_p, dtype, B, D, H, W, STREAM(stream));
}

int mtb200_colsum(const void* m, int32_t dtype, int64_t rows, int32_t ldc, int32_t coff, int32_t C, float* out,
                  void* stream) {
  MTB_REQUIRE(m && out, "colsum: null pointer");
  return colsum(m, dtype, rows, ldc, coff, C, out, STREAM(stream));
}

int mtb200_in_finalize(const double* stats, const float* gamma, const float* beta, int32_t B, int32_t C, int64_t nvox,
                       float eps, float slope, float* xform, float* meanrstd, void* stream) {
  MTB_REQUIRE(stats && gamma && beta && xform && meanrstd && nvox > 0, "in_finalize: bad arguments");
  return in_finalize(stats, gamma, beta, B, C, nvox, eps, slope, xform, meanrstd, STREAM(stream));
}

int mtb200_in_stats(const void* y, int32_t dtype, int32_t B, int64_t nvox, int32_t ldc, int32_t coff, int32_t C,
                    double* stats, void* stream) {
  MTB_REQUIRE(y && stats, "in_stats: null pointer");
  return in_stats(y, dtype, B, nvox, ldc, coff, C, stats, STREAM(stream));
}

int mtb200_norm_act(const void* y, int32_t in_ldc, int32_t in_coff, void* out, int32_t out_ldc, int32_t out_coff,
                    int32_t dtype, int32_t B, int64_t nvox, int32_t C, const float* xform, const void* res,
                    int32_t res_ldc, int32_t res_coff, const float* res_xform, float slope2, void* stream) {
  MTB_REQUIRE(y && out, "norm_act: null pointer");
  return norm_act(y, in_ldc, in_coff, out, out_ldc, out_coff, dtype, B, nvox, C, xform, res, res_ldc, res_coff,
                  res_xform, slope2, STREAM(stream));
}

int mtb200_in_bwd_reduce(const void* dact, int32_t d_ldc, int32_t d_coff, const void* y, int32_t y_ldc, int32_t y_coff,
                         int32_t dtype, int32_t B, int64_t nvox, int32_t C, const float* xform, const float* meanrstd,
                         double* red, void* stream) {
  MTB_REQUIRE(dact && y && xform && meanrstd && red, "in_bwd_reduce: null pointer");
  return in_bwd_reduce(dact, d_ldc, d_coff, y, y_ldc, y_coff, dtype, B, nvox, C, xform, meanrstd, red, STREAM(stream));
}

int mtb200_in_bwd_apply(const void* dact, int32_t d_ldc, int32_t d_coff, const void* y, int32_t y_ldc, int32_t y_coff,
                        void* dy, int32_t dy_ldc, int32_t dy_coff, int32_t dtype, int32_t B, int64_t nvox, int32_t C,
                        const float* xform, const float* meanrstd, const float* gamma, const double* red,
                        float* dgamma, float* dbeta, void* stream) {
  MTB_REQUIRE(dact && y && dy && xform && meanrstd && gamma && red, "in_bwd_apply: null pointer");
  MTB_REQUIRE((dgamma == nullptr) == (dbeta == nullptr), "in_bwd_apply: dgamma/dbeta must both be given or both NULL");
  return in_bwd_apply(dact, d_ldc, d_coff, y, y_ldc, y_coff, dy, dy_ldc, dy_coff, dtype, B, nvox, C, xform, meanrstd,
                      gamma, red, dgamma, dbeta, STREAM(stream));
}

int mtb200_lrelu_bwd(const void* dact, const void* act, void* dv, int32_t dtype, int64_t n, float slope, void* stream) {
  MTB_REQUIRE(dact && act && dv, "lrelu_bwd: null pointer");
  return lrelu_bwd(dact, act, dv, dtype, n, slope, STREAM(stream));
}

int mtb200_residual_bwd(const void* dact, int32_t d_ldc, int32_t d_coff, const void* act, int32_t a_ldc, int32_t a_coff,
                        void* dst0, int32_t ldc0, int32_t coff0, int32_t acc0, void* dst1, int32_t ldc1, int32_t coff1,
                        int32_t acc1, int32_t dtype, int64_t nrows, int32_t C, float slope, void* stream) {
  MTB_REQUIRE(dact && act && (dst0 || dst1), "residual_bwd: null pointer");
  return residual_bwd(dact, d_ldc, d_coff, act, a_ldc, a_coff, dst0, ldc0, coff0, acc0, dst1, ldc1, coff1, acc1, dtype,
                      nrows, C, slope, STREAM(stream));
}

int mtb200_dcce_stats(const void* logits, int32_t dtype, int32_t ldc, int32_t C, int32_t Cp, const float* target, int32_t B,
                      int64_t nvox, double* stats, double* ce_sum, void* stream) {
  MTB_REQUIRE(logits && target && stats && ce_sum, "dcce_stats: null pointer");
  return dcce_stats(logits, dtype, ldc, C, Cp, target, B, nvox, stats, ce_sum, STREAM(stream));
}

int mtb200_dcce_bwd(const void* logits, int32_t dtype, int32_t ldc, int32_t C, int32_t Cp, const float* target, int32_t B,
                    int64_t nvox, const float* coef, float ce_weight, const float* gscale, void* dz, int32_t dz_ldc,
                    void* stream) {
  MTB_REQUIRE(logits && target && coef && gscale && dz, "dcce_bwd: null pointer");
  return dcce_bwd(logits, dtype, ldc, C, Cp, target, B, nvox, coef, ce_weight, gscale, dz, dz_ldc, STREAM(stream));
}

int mtb200_mt_loss_stats(const void* logits, int32_t dtype, int32_t ldc, int32_t C, const float* target, int32_t B,
                         int64_t nvox, const uint64_t* valid_mask, const uint64_t* pos_mask, int32_t n_labels,
                         double* stats, double* hard, void* stream) {
  MTB_REQUIRE(logits && target && valid_mask && pos_mask && stats, "mt_loss_stats: null pointer");
  return mt_loss_stats(logits, dtype, ldc, C, target, B, nvox, valid_mask, pos_mask, n_labels, stats, hard,
                       STREAM(stream));
}

int mtb200_mt_loss_finalize(const double* stats, const double* pooled, const uint64_t* valid_mask, int32_t B, int32_t C,
                            int64_t nvox, float weight, float world_size, float* losses, float* coef, void* stream) {
  MTB_REQUIRE(stats && valid_mask && losses && coef && nvox > 0, "mt_loss_finalize: bad arguments");
  return mt_loss_finalize(stats, pooled, valid_mask, B, C, nvox, weight, world_size, losses, coef, STREAM(stream));
}

int mtb200_mt_loss_bwd(const void* logits, int32_t dtype, int32_t ldc, int32_t C, const float* target, int32_t B,
                       int64_t nvox, const uint64_t* pos_mask, int32_t n_labels, const float* coef, const float* gscale,
                       void* dlogits, int32_t d_ldc, void* stream) {
  MTB_REQUIRE(logits && target && pos_mask && coef && dlogits, "mt_loss_bwd: null pointer");
  return mt_loss_bwd(logits, dtype, ldc, C, target, B, nvox, pos_mask, n_labels, coef, gscale, dlogits, d_ldc,
                     STREAM(stream));
}
int mtb200_head_bwd_fused(const mtb200_head_bwd_params* p, void* stream) {
  MTB_REQUIRE(p && (p->logits || p->w_fwd) && p->target && p->coef && p->pos_mask && p->x && p->w_swap && p->dx && p->dw,
              "head_bwd_fused: null pointer");
  return head_bwd_fused(*p, (cudaStream_t)stream);
}
int mtb200_head_aggregate(const mtb200_head_agg_params* p, void* stream) {
  MTB_REQUIRE(p && p->x && p->w_fwd && p->acc, "head_aggregate: null pointer");
  return head_aggregate(*p, (cudaStream_t)stream);
}
int mtb200_head_fwd_stats(const mtb200_head_fwd_params* p, void* stream) {
  MTB_REQUIRE(p && p->x && p->w_fwd && p->target && p->valid_mask && p->pos_mask && p->stats, "head_fwd_stats: null pointer");
  return head_fwd_stats(*p, (cudaStream_t)stream);
}

int mtb200_sw_gather_tile(const float* vol, int32_t Cin, int32_t X, int32_t Y, int32_t Z, int32_t x0, int32_t y0,
                          int32_t z0, int32_t pd, int32_t ph, int32_t pw, int32_t flip, void* tile, int32_t dtype,
                          int32_t ldc, void* stream) {
  MTB_REQUIRE(vol && tile, "sw_gather_tile: null pointer");
  return sw_gather_tile(vol, Cin, X, Y, Z, x0, y0, z0, pd, ph, pw, flip, tile, dtype, ldc, STREAM(stream));
}

int mtb200_sw_aggregate(const void* logits, int32_t dtype, int32_t ldc, int32_t C, int32_t pd, int32_t ph, int32_t pw,
                        int32_t flip, const float* gauss, float weight, int32_t apply_sigmoid, float* acc, float* nb,
                        int32_t X, int32_t Y, int32_t Z, int32_t x0, int32_t y0, int32_t z0, void* stream) {
  MTB_REQUIRE(logits && acc, "sw_aggregate: null pointer");
  return sw_aggregate(logits, dtype, ldc, C, pd, ph, pw, flip, gauss, weight, apply_sigmoid, acc, nb, X, Y, Z, x0, y0, z0,
                      STREAM(stream));
}

int mtb200_sw_finalize(float* acc, const float* nb, int32_t C, int64_t nvox, const float* class_order, float* seg,
                       void* stream) {
  MTB_REQUIRE(acc && nb, "sw_finalize: null pointer");
  return sw_finalize(acc, nb, C, nvox, nvox, class_order, seg, STREAM(stream));
}
int mtb200_sw_finalize_slab(float* acc, const float* nb, int32_t C, int64_t class_stride, int64_t nvox,
                            const float* class_order, float* seg, void* stream) {
  MTB_REQUIRE(acc && nb && class_stride >= nvox, "sw_finalize_slab: bad arguments");
  return sw_finalize(acc, nb, C, class_stride, nvox, class_order, seg, STREAM(stream));
}

int mtb200_sumsq(const float* g, int64_t n, double* out, void* stream) {
  MTB_REQUIRE(g && out, "sumsq: null pointer");
  return sumsq(g, n, out, STREAM(stream));
}

int mtb200_sgd_step(float* p, const float* g, float* buf, int64_t n, const double* sumsq_, float inv_scale,
                    float max_norm, float lr, float momentum, float weight_decay, int32_t first_step,
                    const float* dyn_scale, void* stream) {
  MTB_REQUIRE(p && g && buf && sumsq_, "sgd_step: null pointer");
  return sgd_step(p, g, buf, n, sumsq_, inv_scale, max_norm, lr, momentum, weight_decay, first_step, dyn_scale,
                  STREAM(stream));
}

int mtb200_loss_scale_update(const double* sumsq_, float* state, float growth_factor, float backoff_factor,
                             int32_t growth_interval, void* stream) {
  MTB_REQUIRE(sumsq_ && state && growth_interval >= 1, "loss_scale_update: bad arguments");
  return loss_scale_update(sumsq_, state, growth_factor, backoff_factor, growth_interval, STREAM(stream));
}

int mtb200_pack_weights(const float* w, int32_t Cout, int32_t Cin, int32_t ntap, int32_t transposed, int32_t swap_io,
                        void* packed, int32_t wdtype, int32_t Cout_p, int32_t Cin_p, int32_t split, int32_t split_p,
                        void* stream) {
  MTB_REQUIRE(w && packed, "pack_weights: null pointer");
  return pack_weights(w, Cout, Cin, ntap, transposed, swap_io, packed, wdtype, Cout_p, Cin_p, split, split_p,
                      STREAM(stream));
}

int mtb200_pack_weights_batched(const mtb200_pack_desc* descs, int32_t n, int32_t total_blocks, int32_t wdtype,
                                void* stream) {
  MTB_REQUIRE(descs || n == 0, "pack_weights_batched: null descriptor table");
  return pack_weights_batched(descs, n, total_blocks, wdtype, STREAM(stream));
}

int mtb200_unpack_wgrad(const float* dw, int32_t Cout, int32_t Cin, int32_t ntap, int32_t transposed, int32_t Cout_p,
                        int32_t Cin_p, int32_t split, int32_t split_p, float scale, int32_t accumulate, float* grad,
                        void* stream) {
  MTB_REQUIRE(dw && grad, "unpack_wgrad: null pointer");
  return unpack_wgrad(dw, Cout, Cin, ntap, transposed, Cout_p, Cin_p, split, split_p, scale, accumulate, grad,
                      STREAM(stream));
}

int mtb200_unpack_wgrad_batched(const mtb200_unpack_desc* descs, int32_t n, int32_t total_blocks, void* stream) {
  MTB_REQUIRE(descs || n == 0, "unpack_wgrad_batched: null descriptor table");
  return unpack_wgrad_batched(descs, n, total_blocks, STREAM(stream));
}

int mtb200_ncdhw_to_ndhwc(const float* src, int32_t B, int32_t C, int64_t nvox, void* dst, int32_t dtype, int32_t ldc,
                          int32_t coff, int32_t Cp, void* stream) {
  MTB_REQUIRE(src && dst, "ncdhw_to_ndhwc: null pointer");
  return ncdhw_to_ndhwc(src, B, C, nvox, dst, dtype, ldc, coff, Cp, STREAM(stream));
}

int mtb200_ndhwc_to_ncdhw(const void* src, int32_t dtype, int32_t ldc, int32_t coff, int32_t B, int32_t C, int64_t nvox,
                          float* dst, void* stream) {
  MTB_REQUIRE(src && dst, "ndhwc_to_ncdhw: null pointer");
  return ndhwc_to_ncdhw(src, dtype, ldc, coff, B, C, nvox, dst, STREAM(stream));
}

int mtb200_crop_pad(const float* src, int32_t C, int32_t X, int32_t Y, int32_t Z, int32_t lbx, int32_t lby, int32_t lbz,
                    float* dst, int32_t pd, int32_t ph, int32_t pw, int32_t edge_mode, const float* pad_values,
                    void* stream) {
  MTB_REQUIRE(src && dst && C >= 1 && X >= 1 && Y >= 1 && Z >= 1, "crop_pad: bad arguments");
  return crop_pad(src, C, X, Y, Z, lbx, lby, lbz, dst, pd, ph, pw, edge_mode, pad_values, STREAM(stream));
}

int mtb200_resize_nearest(const float* src, int64_t NC, int32_t X, int32_t Y, int32_t Z, float* dst, int32_t X2,
                          int32_t Y2, int32_t Z2, void* stream) {
  MTB_REQUIRE(src && dst && X >= 1 && Y >= 1 && Z >= 1 && X2 >= 1 && Y2 >= 1 && Z2 >= 1, "resize_nearest: bad arguments");
  return resize_nearest(src, NC, X, Y, Z, dst, X2, Y2, Z2, STREAM(stream));
}

int mtb200_resample_probs(const float* src, int32_t C, int32_t X, int32_t Y, int32_t Z, int32_t X2, int32_t Y2, int32_t Z2,
                          int32_t order_x, int32_t order_y, int32_t order_z, void* prob, int32_t prob_is_f16,
                          const float* class_order, uint8_t* seg, void* stream) {
  MTB_REQUIRE(src && (prob || seg) && C >= 1, "resample_probs: bad arguments");
  MTB_REQUIRE(order_x >= 0 && order_x <= 1 && order_y >= 0 && order_y <= 1 && order_z >= 0 && order_z <= 1,
              "resample_probs: interpolation orders 0 (nearest) and 1 (linear) are implemented, got %d %d %d", order_x,
              order_y, order_z);
  return resample_probs(src, C, X, Y, Z, X2, Y2, Z2, order_x, order_y, order_z, prob, prob_is_f16, class_order, seg,
                        STREAM(stream));
}

}  // extern "C"
