// process.cpp -- iw_process_image(): the single entry point of the hot path, plus the small host-side steps
// that surround the kernels (background-colour label, result view).
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "iwb_internal.h"

namespace {

// Nearest representable code for one linear sample, scalar host version used only for the three samples of the
// background label (reference: calc_sample_convert_from_linear + get_nearest_valid_colors,
// src/imagew-main.c:570-629, :769-799).
unsigned int label_code(double s, const iw_csdescr *cs, double maxcode) {
  if (s < 0.0) s = 0.0;
  if (s > 1.0) s = 1.0;
  double e = iwb_from_linear(s, cs) * maxcode;
  if (e > maxcode) e = maxcode;
  if (e < 0.0) e = 0.0;
  double lo = floor(e), hi = ceil(e);
  unsigned int ilo = (unsigned int)lo, ihi = (unsigned int)hi;
  if (ilo == ihi) return (unsigned int)(0.5 + lo);
  double lin_lo = iwb_to_linear(((double)ilo) / maxcode, cs);
  double lin_hi = iwb_to_linear(((double)ihi) / maxcode, cs);
  double d_floor = s - lin_lo, d_ceil = lin_hi - s;
  return (unsigned int)(0.5 + (d_ceil <= d_floor ? hi : lo));
}

// reference: iw_process_bkgd_label, src/imagew-main.c:1238-1322
void make_output_label(iw_context *ctx) {
  if (!(ctx->output_profile & (IW_PROFILE_PNG_BKGD | IW_PROFILE_RGB8_BKGD | IW_PROFILE_RGB16_BKGD))) return;
  if (ctx->suppress_out_label) return;
  iw_color clr;
  if (ctx->out_label_valid) clr = ctx->req_out_label;
  else if (ctx->in_label_set) clr = ctx->in_label_lin;
  else return;
  if (ctx->to_grayscale) {
    double g = iwb_gray_of(ctx, clr.c[0], clr.c[1], clr.c[2]);
    clr.c[0] = clr.c[1] = clr.c[2] = g;
  }
  double mx;
  if (ctx->output_profile & IW_PROFILE_RGB8_BKGD) mx = 255.0;
  else if (ctx->output_profile & IW_PROFILE_RGB16_BKGD) mx = 65535.0;
  else if (ctx->img2.bit_depth == 8) mx = 255.0;
  else if (ctx->img2.bit_depth == 16) mx = 65535.0;
  else return;
  for (int k = 0; k < 3; k++) ctx->img2.bkgdlabel.c[k] = ((double)label_code(clr.c[k], &ctx->img2cs, mx)) / mx;
  double a = clr.c[3];
  unsigned int ai = a <= 0.0 ? 0u : a >= 1.0 ? (unsigned int)mx : (unsigned int)(0.5 + a * (unsigned int)mx);
  ctx->img2.bkgdlabel.c[3] = ((double)ai) / mx;
  ctx->img2.has_bkgdlabel = 1;
}

double now_ms() {
  using namespace std::chrono;
  return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

}  // namespace

extern "C" int iwb200_run_host_multi(struct iw_context *ctx, const int *devices, int ndev, const void *h_in, size_t in_bpr, void *h_out,
                                     size_t out_bpr);

extern "C" {

int iw_process_image(struct iw_context *ctx) {
  if (ctx->use_count > 0) {
    iw_set_error(ctx, "Internal: Incorrect attempt to reprocess image");
    return 0;
  }
  ctx->use_count++;

  const char *backend = iwb_option(ctx, "backend");
  if (backend && strcmp(backend, "cuda") != 0) {
    iw_set_errorf(ctx, "Backend \"%s\" is not available: this library only has the CUDA path", backend);
    return 0;
  }
  double t0 = now_ms();
  if (!iwb_prepare(ctx)) return 0;
  if (!ctx->img1.pixels) {
    iw_set_error(ctx, "No input image");
    return 0;
  }

  int device = 0;
  if (const char *d = iwb_option(ctx, "cuda:device")) device = atoi(d);
  // "cuda:devices" = "0,1,2,...": the image is split into output-row bands, one per listed device (pipeline.cu:
  // iwb200_run_host_multi); a single entry is the same as cuda:device
  std::vector<int> devices;
  if (const char *dl = iwb_option(ctx, "cuda:devices")) {
    for (const char *q = dl; *q;) {
      char *end = nullptr;
      const long v = strtol(q, &end, 10);
      if (end == q || v < 0 || v > 1023) { devices.clear(); break; }
      devices.push_back((int)v);
      q = end;
      while (*q == ',' || *q == ' ') q++;
    }
    if (devices.empty() || devices.size() > 64) {
      iw_set_error(ctx, "CUDA backend: cuda:devices must be a comma-separated list of device ordinals");
      return 0;
    }
    if (devices.size() == 1) { device = devices[0]; devices.clear(); }
    else if (iwb_option(ctx, "cuda:band")) {
      iw_set_error(ctx, "CUDA backend: cuda:devices and cuda:band cannot be combined");
      return 0;
    }
  }
  if (!devices.empty()) {
    ctx->img2.bpr = iw_calc_bytesperrow(ctx->img2.width, ctx->img2.bit_depth * ctx->out_nch);
    ctx->img2.pixels = (iw_byte *)iw_malloc_large(ctx, ctx->img2.bpr, (size_t)ctx->img2.height);
    if (!ctx->img2.pixels) return 0;
    ctx->band_row0 = 0;
    ctx->band_row1 = ctx->img2.height;
    ctx->timings.plan_ms = now_ms() - t0;
    if (!iwb200_run_host_multi(ctx, devices.data(), (int)devices.size(), ctx->img1.pixels, ctx->img1.bpr, ctx->img2.pixels, ctx->img2.bpr)) {
      if (!ctx->error_flag) iw_set_error(ctx, "CUDA backend: multi-device run failed");
      return 0;
    }
    return 1;                                                     // (result view and optimiser: inside iwb200_run_host_multi)
  }
  iwb200_job *job = iwb200_job_create(ctx, device);
  if (!job) {
    if (!ctx->error_flag) iw_set_error(ctx, "CUDA backend: could not build the processing plan");
    return 0;
  }
  iwb200_job_info info;
  iwb200_job_get_info(job, &info);

  // optional single band: "cuda:band" = "index/count"
  int r0 = 0, r1 = info.out_height;
  if (const char *b = iwb_option(ctx, "cuda:band")) {
    int idx = 0, cnt = 1;
    if (sscanf(b, "%d/%d", &idx, &cnt) == 2 && cnt >= 1 && idx >= 0 && idx < cnt) {
      if (cnt > 1 && !info.bandable) {
        iw_set_error(ctx, "CUDA backend: error-diffusion dithering cannot be split into row bands");
        iwb200_job_destroy(job);
        return 0;
      }
      r0 = (int)(((long long)info.out_height * idx) / cnt);
      r1 = (int)(((long long)info.out_height * (idx + 1)) / cnt);
      if (r1 <= r0) r1 = r0 + 1 <= info.out_height ? r0 + 1 : r0;
    } else {
      iw_set_error(ctx, "CUDA backend: cuda:band must be \"index/count\"");
      iwb200_job_destroy(job);
      return 0;
    }
  }
  ctx->band_row0 = r0;
  ctx->band_row1 = r1;

  ctx->img2.bpr = info.out_bpr;
  ctx->img2.pixels = (iw_byte *)iw_malloc_large(ctx, ctx->img2.bpr, (size_t)(r1 - r0));
  if (!ctx->img2.pixels) {
    iwb200_job_destroy(job);
    return 0;
  }
  ctx->timings.plan_ms = now_ms() - t0;

  int ok = iwb200_job_run_host(job, ctx->img1.pixels, ctx->img1.bpr, ctx->img2.pixels, ctx->img2.bpr, r0, r1);
  if (!ok) iw_set_errorf(ctx, "CUDA backend: %s", iwb200_job_last_error(job));
  if (ok) {
    iwb_set_result_view(ctx, r0, r1);
    // the post-pass optimiser (src/imagew-opt.c) while the result is still in HBM; a single band of an image is handed
    // back as processed (bands of one image must stay concatenable)
    if (r0 == 0 && r1 == info.out_height) {
      const int band[2] = {r0, r1};
      ok = iwb_optimize_output(ctx, &job, band, 1);
      if (!ok && !ctx->error_flag) iw_set_errorf(ctx, "CUDA backend: output optimisation failed: %s", iwb200_job_last_error(job));
    }
  }
  iwb200_job_destroy(job);
  return ok ? 1 : 0;
}

}  // extern "C"

// what follows the kernels: background-colour label and the result view handed out by iw_get_output_image (the image
// as processed; the optimiser may replace it, pipeline.cu:iwb_optimize_output).  The optimiser's always-on prologue
// (transparent pixels -> black, src/imagew-opt.c:1112-1163) ran on the device.
void iwb_set_result_view(struct iw_context *ctx, int r0, int r1) {
  make_output_label(ctx);
  ctx->res_w = ctx->img2.width;
  ctx->res_h = r1 - r0;
  ctx->res_type = ctx->img2.imgtype;
  ctx->res_depth = ctx->img2.bit_depth;
  ctx->res_bpr = ctx->img2.bpr;
  ctx->res_pixels = ctx->img2.pixels;
  if (ctx->img2.has_bkgdlabel) {
    ctx->res_has_label = 1;
    for (int k = 0; k < 4; k++)
      ctx->res_label[k] = iw_color_get_int_sample(&ctx->img2.bkgdlabel, k, ctx->img2.bit_depth == 8 ? 255 : 65535);
  }
}

extern "C" {

void iwb200_detach_input_image(struct iw_context *ctx) {
  if (ctx) ctx->img1.pixels = nullptr;
}

int iwb200_get_band(struct iw_context *ctx, int *r0, int *r1) {
  if (!ctx) return 0;
  *r0 = ctx->band_row0;
  *r1 = ctx->band_row1;
  return 1;
}

int iwb200_get_timings(struct iw_context *ctx, double *plan_ms, double *h2d_ms, double *kernel_ms, double *d2h_ms) {
  if (!ctx) return 0;
  if (plan_ms) *plan_ms = ctx->timings.plan_ms;
  if (h2d_ms) *h2d_ms = ctx->timings.h2d_ms;
  if (kernel_ms) *kernel_ms = ctx->timings.kernel_ms;
  if (d2h_ms) *d2h_ms = ctx->timings.d2h_ms;
  return 1;
}

}  // extern "C"
