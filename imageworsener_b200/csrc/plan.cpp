// plan.cpp -- the host decision logic that precedes any pixel work.
//
// Restates, step for step and quirk for quirk, what the reference decides in iw_prepare_processing() and its
// helpers (reference: src/imagew-main.c:1463-2248): crop clamp, logical input type (virtual alpha for
// transparent edges), background strategy, input->intermediate->output channel maps, output colour space,
// depth, per-channel max codes / colour counts / dither, automatic filter choice.  A wrong decision here is a
// wrong pixel no matter how good the kernels are, so the structure deliberately mirrors the reference's order
// of evaluation; the data it fills (iw_context decisions block) is this library's own.
#include <cmath>
#include <ctime>

#include "iwb_internal.h"

namespace {

int channeltype_of(int imgtype, int channel) {  // src/imagew-main.c:1463-1486
  switch (imgtype) {
    case IW_IMGTYPE_GRAY: if (channel == 0) return IW_CHANNELTYPE_GRAY; break;
    case IW_IMGTYPE_GRAYA:
      if (channel == 0) return IW_CHANNELTYPE_GRAY;
      if (channel == 1) return IW_CHANNELTYPE_ALPHA;
      break;
    case IW_IMGTYPE_RGB: if (channel >= 0 && channel <= 2) return channel; break;
    case IW_IMGTYPE_RGBA: if (channel >= 0 && channel <= 3) return channel; break;
  }
  return 0;
}

}  // namespace

// Whether the background being composited is itself translucent (src/imagew-main.c:813-828)
bool iwb_bkgd_is_translucent(const iw_context *ctx) {
  if (!ctx->apply_bkgd) return false;
  if (!(ctx->output_profile & IW_PROFILE_TRANSPARENCY)) return false;
  if (ctx->bkgd_strategy == IWB_BKGD_EARLY) return false;
  if (ctx->bkgd_source == IWB_BKGD_SRC_FILE) return ctx->in_label_inputcs.c[3] < 1.0;
  if (ctx->bkgd_source == IWB_BKGD_SRC_REQ) {
    if (ctx->bkgd_checkerboard && ctx->req_bkgd2.c[3] < 1.0) return true;
    return ctx->req_bkgd.c[3] < 1.0;
  }
  return false;
}

namespace {

// Whether, when and from where a background is applied (src/imagew-main.c:1734-1823)
void decide_background(iw_context *ctx) {
  if (!IW_IMGTYPE_HAS_ALPHA(ctx->in_type_logical)) {
    ctx->apply_bkgd = 0;
    return;
  }
  if (ctx->in_label_set && (ctx->use_label_from_file || !ctx->bkgd_valid)) {
    ctx->bkgd_source = IWB_BKGD_SRC_FILE;
  } else if (ctx->bkgd_valid) {
    ctx->bkgd_source = IWB_BKGD_SRC_REQ;
    ctx->bkgd_checkerboard = ctx->req_checkerboard;
  } else {
    ctx->bkgd_source = IWB_BKGD_SRC_NONE;
  }
  if (ctx->bkgd_checkerboard && ctx->check_size < 1) ctx->check_size = 1;
  if (ctx->bkgd_valid) ctx->apply_bkgd = 1;
  if (!(ctx->output_profile & IW_PROFILE_TRANSPARENCY)) {
    if (!ctx->bkgd_valid && !ctx->apply_bkgd)
      iw_warning(ctx, "This image may have transparency, which is incompatible with the output format. A background color will be applied.");
    ctx->apply_bkgd = 1;
  }
  if (ctx->axis[IW_DIMENSION_H].use_offset || ctx->axis[IW_DIMENSION_V].use_offset) {
    // channel offsets + transparency: only "solid colour, before resizing" is supported
    if (!ctx->bkgd_valid && !ctx->apply_bkgd)
      iw_warning(ctx, "This image may have transparency, which is incompatible with a channel offset. A background color will be applied.");
    ctx->apply_bkgd = 1;
    if (ctx->bkgd_checkerboard && ctx->req_checkerboard) {
      iw_warning(ctx, "Checkerboard backgrounds are not supported when using a channel offset.");
      ctx->bkgd_checkerboard = 0;
    }
    ctx->bkgd_strategy = IWB_BKGD_EARLY;
    return;
  }
  if (!ctx->apply_bkgd) return;
  ctx->bkgd_strategy = IWB_BKGD_LATE;  // checkerboards must be late; solid colours are late by choice
}

// The two-stage channel strategy (src/imagew-main.c:1613-1718).  Codes: 1=G 2=GA 3=RGB 4=RGBA.
void decide_strategy(const iw_context *ctx, int *from_to, int *mid_to_out) {
  int in = 0, mid = 0, out = 0;
  switch (ctx->in_type_logical) {
    case IW_IMGTYPE_RGBA: in = 4; mid = ctx->to_grayscale ? 2 : 4; break;
    case IW_IMGTYPE_RGB: in = 3; mid = ctx->to_grayscale ? 1 : 3; break;
    case IW_IMGTYPE_GRAYA: in = 2; mid = ctx->to_grayscale ? 2 : 4; break;
    default: in = 1; mid = ctx->to_grayscale ? 1 : 3; break;
  }
  out = mid;
  if (ctx->apply_bkgd && ctx->bkgd_strategy == IWB_BKGD_EARLY) {
    // background goes in before resizing: the intermediate image has no alpha at all
    if (mid == 4) { mid = 3; out = 3; }
    else if (mid == 2) { mid = 1; out = 1; }
  }
  if (ctx->apply_bkgd && !iwb_bkgd_is_translucent(ctx)) {
    if (mid == 2 && out == 2) out = 1;
    else if (mid == 4 && out == 4) out = 3;
  }
  *from_to = in * 16 + mid;
  *mid_to_out = mid * 16 + out;
}

void set_background_colours(iw_context *ctx) {  // src/imagew-main.c:1546-1611
  if (!ctx->apply_bkgd) return;
  iw_color b1, b2;
  b1.c[0] = 1.0; b1.c[1] = 0.0; b1.c[2] = 1.0; b1.c[3] = 1.0;  // default: magenta
  b2.c[0] = 0.0; b2.c[1] = 0.0; b2.c[2] = 0.0; b2.c[3] = 1.0;
  if (ctx->bkgd_source == IWB_BKGD_SRC_FILE) {
    b1 = ctx->in_label_lin;
    ctx->bkgd_checkerboard = 0;
  } else if (ctx->bkgd_source == IWB_BKGD_SRC_REQ) {
    b1 = ctx->req_bkgd;
    if (ctx->req_checkerboard) b2 = ctx->req_bkgd2;
  }
  if (ctx->bkgd_strategy == IWB_BKGD_EARLY) {
    ctx->bkgd1alpha = 1.0;
  } else {
    ctx->bkgd1alpha = b1.c[3];
    ctx->bkgd2alpha = b2.c[3];
  }
  const bool out_rgb = ctx->img2.imgtype == IW_IMGTYPE_RGB || ctx->img2.imgtype == IW_IMGTYPE_RGBA;
  const bool out_gray = ctx->img2.imgtype == IW_IMGTYPE_GRAY || ctx->img2.imgtype == IW_IMGTYPE_GRAYA;
  if (ctx->bkgd_strategy == IWB_BKGD_LATE && out_rgb) {
    for (int i = 0; i < 3; i++) ctx->out_ci[i].bkgd1 = b1.c[i];
    if (ctx->bkgd_checkerboard)
      for (int i = 0; i < 3; i++) ctx->out_ci[i].bkgd2 = b2.c[i];
  } else if (ctx->bkgd_strategy == IWB_BKGD_LATE && out_gray) {
    ctx->out_ci[0].bkgd1 = iwb_gray_of(ctx, b1.c[0], b1.c[1], b1.c[2]);
    if (ctx->bkgd_checkerboard) ctx->out_ci[0].bkgd2 = iwb_gray_of(ctx, b2.c[0], b2.c[1], b2.c[2]);
  } else if (ctx->bkgd_strategy == IWB_BKGD_EARLY && ctx->img2.imgtype == IW_IMGTYPE_RGB) {
    for (int i = 0; i < 3; i++) ctx->mid_ci[i].early_bkgd = b1.c[i];
  } else if (ctx->bkgd_strategy == IWB_BKGD_EARLY && ctx->img2.imgtype == IW_IMGTYPE_GRAY) {
    ctx->mid_ci[0].early_bkgd = iwb_gray_of(ctx, b1.c[0], b1.c[1], b1.c[2]);
  }
}

void choose_auto_filter(iw_context *ctx, int size_in, int size_out, int d) {  // src/imagew-main.c:1825-1840
  const iwb_axis_request &a = ctx->axis[d];
  if (size_out == size_in && !a.use_offset && !ctx->out_true_valid && a.translate == 0.0)
    iw_set_resize_alg(ctx, d, IW_RESIZETYPE_NULL, 1.0, 0.0, 0.0);
  else
    iw_set_resize_alg(ctx, d, IW_RESIZETYPE_CUBIC, 1.0, 0.0, 0.5);  // Catmull-Rom
}

}  // namespace

int iwb_prepare(iw_context *ctx) {
  const int w = ctx->canvas_w, h = ctx->canvas_h;
  if (ctx->output_profile == 0) {
    iw_set_error(ctx, "Output profile not set");
    return 0;
  }
  if (ctx->randomize) ctx->random_seed = (int)time(nullptr);  // src/imagew-util.c:237-243

  ctx->axis[IW_DIMENSION_H].out_true_size = ctx->out_true_valid ? ctx->out_true_w : (double)w;
  ctx->axis[IW_DIMENSION_V].out_true_size = ctx->out_true_valid ? ctx->out_true_h : (double)h;

  if (!iw_check_image_dimensions(ctx, ctx->img1.width, ctx->img1.height)) return 0;
  if (!iw_check_image_dimensions(ctx, w, h)) return 0;

  if (ctx->to_grayscale) {  // named formulas become weights (src/imagew-main.c:1883-1895)
    if (ctx->grayscale_formula == IW_GSF_STANDARD) {
      ctx->grayscale_formula = IW_GSF_WEIGHTED;
      iw_set_grayscale_weights(ctx, 0.212655, 0.715158, 0.072187);
    } else if (ctx->grayscale_formula == IW_GSF_COMPATIBLE) {
      ctx->grayscale_formula = IW_GSF_WEIGHTED;
      iw_set_grayscale_weights(ctx, 0.299, 0.587, 0.114);
    }
  }

  // ---- logical input type; transparent edges add a virtual alpha channel (src/imagew-main.c:1842-1880)
  ctx->in_type_logical = ctx->img1.imgtype;
  if (ctx->axis[IW_DIMENSION_H].edge_policy == IW_EDGE_POLICY_TRANSPARENT ||
      ctx->axis[IW_DIMENSION_V].edge_policy == IW_EDGE_POLICY_TRANSPARENT) {
    if (ctx->img1.imgtype == IW_IMGTYPE_GRAY) ctx->in_type_logical = IW_IMGTYPE_GRAYA;
    else if (ctx->img1.imgtype == IW_IMGTYPE_RGB) ctx->in_type_logical = IW_IMGTYPE_RGBA;
  }
  ctx->in_nch_phys = iw_imgtype_num_channels(ctx->img1.imgtype);
  ctx->in_nch_logical = iw_imgtype_num_channels(ctx->in_type_logical);
  ctx->in_alpha_index = iwb_imgtype_alpha_index(ctx->in_type_logical);
  ctx->img2.imgtype = ctx->in_type_logical;
  ctx->out_nch = ctx->in_nch_logical;
  ctx->mid_nch = ctx->in_nch_logical;
  for (int i = 0; i < 4; i++) {
    ctx->in_ci[i] = iwb_in_channel();
    ctx->in_ci[i].maxcode = ctx->in_maxcode_req[i];
    ctx->mid_ci[i] = iwb_mid_channel();
    ctx->out_ci[i] = iwb_out_channel();
  }
  for (int i = 0; i < ctx->in_nch_logical; i++) {
    ctx->in_ci[i].channeltype = channeltype_of(ctx->in_type_logical, i);
    ctx->mid_ci[i].channeltype = ctx->in_ci[i].channeltype;
    ctx->mid_ci[i].in_channel = i;
    ctx->out_ci[i].channeltype = ctx->in_ci[i].channeltype;
    if (i >= ctx->in_nch_phys) ctx->in_ci[i].slow_path = 1;  // virtual channel
  }

  ctx->img2.width = w;
  ctx->img2.height = h;

  // ---- crop region.  The fourth clamp assigns the x origin where y was meant; kept (src/imagew-main.c:1952)
  ctx->in_x0 = ctx->crop_x; ctx->in_y0 = ctx->crop_y; ctx->in_w = ctx->crop_w; ctx->in_h = ctx->crop_h;
  if (ctx->in_x0 < 0) ctx->in_x0 = 0;
  if (ctx->in_y0 < 0) ctx->in_y0 = 0;
  if (ctx->in_x0 > ctx->img1.width - 1) ctx->in_x0 = ctx->img1.width - 1;
  if (ctx->in_y0 > ctx->img1.height - 1) ctx->in_x0 = ctx->img1.height - 1;
  if (ctx->in_w < 0) ctx->in_w = ctx->img1.width - ctx->in_x0;
  if (ctx->in_h < 0) ctx->in_h = ctx->img1.height - ctx->in_y0;
  if (ctx->in_w < 1) ctx->in_w = 1;
  if (ctx->in_h < 1) ctx->in_h = 1;
  if (ctx->in_w > (ctx->img1.width - ctx->in_x0)) ctx->in_w = ctx->img1.width - ctx->in_x0;
  if (ctx->in_h > (ctx->img1.height - ctx->in_y0)) ctx->in_h = ctx->img1.height - ctx->in_y0;

  // ---- output colour space (src/imagew-main.c:1960-1980)
  if (ctx->out_cs_valid) {
    ctx->img2cs = ctx->req_out_cs;
    if ((ctx->output_profile & IW_PROFILE_ALWAYSLINEAR) && ctx->img2cs.cstype != IW_CSTYPE_LINEAR) {
      iw_warning(ctx, "Forcing output colorspace to linear; required by the output format.");
      iw_make_linear_csdescr(&ctx->img2cs);
    }
  } else if (ctx->output_profile & IW_PROFILE_ALWAYSLINEAR) {
    iw_make_linear_csdescr(&ctx->img2cs);
  } else {
    iw_make_srgb_csdescr_2(&ctx->img2cs);
  }

  // ---- input max codes (src/imagew-main.c:1982-2007)
  if (ctx->img1.sampletype != IW_SAMPLETYPE_FLOATINGPOINT) {
    ctx->in_maxcode_full = (1 << ctx->img1.bit_depth) - 1;
    for (int i = 0; i < 4; i++) {
      if (ctx->in_ci[i].maxcode <= 0) ctx->in_ci[i].maxcode = ctx->in_maxcode_full;
      if (ctx->in_ci[i].maxcode != ctx->in_maxcode_full) ctx->reduced_input_depths = 1;
    }
  }
  if (ctx->reduced_input_depths || ctx->img1.sampletype == IW_SAMPLETYPE_FLOATINGPOINT)
    for (int i = 0; i < ctx->in_nch_phys; i++) ctx->in_ci[i].slow_path = 1;

  // ---- channel offsets (src/imagew-main.c:2009-2026)
  for (int d = 0; d < 2; d++)
    for (int j = 0; j < 3; j++)
      if (fabs(ctx->axis[d].channel_offset[j]) > 0.00001) ctx->axis[d].use_offset = 1;
  if (ctx->to_grayscale && (ctx->axis[0].use_offset || ctx->axis[1].use_offset)) {
    iw_warning(ctx, "Disabling channel offset, due to grayscale output.");
    ctx->axis[0].use_offset = ctx->axis[1].use_offset = 0;
  }

  decide_background(ctx);

  // per-channel tables: offsets differ per channel, or the virtual-pixel value does (src/imagew-main.c:2030-2045)
  for (int d = 0; d < 2; d++)
    if (ctx->axis[d].use_offset || (ctx->apply_bkgd && ctx->bkgd_strategy == IWB_BKGD_EARLY &&
                                    ctx->axis[d].edge_policy == IW_EDGE_POLICY_TRANSPARENT))
      ctx->axis[d].per_channel_tables = 1;

  int s1, s2;
  decide_strategy(ctx, &s1, &s2);
  switch (s1) {  // input -> intermediate (src/imagew-main.c:2049-2092)
    case 0x44: ctx->mid_type = IW_IMGTYPE_RGBA; break;
    case 0x24:
      ctx->mid_type = IW_IMGTYPE_RGBA;
      ctx->mid_ci[0].in_channel = ctx->mid_ci[1].in_channel = ctx->mid_ci[2].in_channel = 0;
      ctx->mid_ci[3].in_channel = 1;
      break;
    case 0x33: case 0x43: ctx->mid_type = IW_IMGTYPE_RGB; break;
    case 0x13: case 0x23:
      ctx->mid_type = IW_IMGTYPE_RGB;
      ctx->mid_ci[0].in_channel = ctx->mid_ci[1].in_channel = ctx->mid_ci[2].in_channel = 0;
      break;
    case 0x42:
      ctx->mid_type = IW_IMGTYPE_GRAYA;
      ctx->mid_ci[0].to_gray = 1; ctx->mid_ci[0].in_channel = 0; ctx->mid_ci[1].in_channel = 3;
      break;
    case 0x22: ctx->mid_type = IW_IMGTYPE_GRAYA; break;
    case 0x31:
      ctx->mid_type = IW_IMGTYPE_GRAY;
      ctx->mid_ci[0].to_gray = 1; ctx->mid_ci[0].in_channel = 0;
      break;
    case 0x11: ctx->mid_type = IW_IMGTYPE_GRAY; ctx->mid_ci[0].in_channel = 0; break;
    default:
      // GA->G and RGBA->G before resizing have no channel map in the reference either (main.c:2089-2091)
      iw_set_errorf(ctx, "Internal error, unknown strategy %d", s1);
      return 0;
  }
  ctx->mid_nch = iw_imgtype_num_channels(ctx->mid_type);
  ctx->mid_alpha_index = iwb_imgtype_alpha_index(ctx->mid_type);
  for (int i = 0; i < ctx->mid_nch; i++) ctx->mid_ci[i].out_channel = i;
  switch (s2) {  // intermediate -> output (src/imagew-main.c:2102-2126)
    case 0x44: ctx->img2.imgtype = IW_IMGTYPE_RGBA; break;
    case 0x33: ctx->img2.imgtype = IW_IMGTYPE_RGB; break;
    case 0x43: ctx->img2.imgtype = IW_IMGTYPE_RGB; ctx->mid_ci[3].out_channel = -1; break;
    case 0x22: ctx->img2.imgtype = IW_IMGTYPE_GRAYA; break;
    case 0x11: ctx->img2.imgtype = IW_IMGTYPE_GRAY; break;
    case 0x21: ctx->img2.imgtype = IW_IMGTYPE_GRAY; ctx->mid_ci[1].out_channel = -1; break;
    default:
      iw_set_error(ctx, "Internal error");
      return 0;
  }
  ctx->out_nch = iw_imgtype_num_channels(ctx->img2.imgtype);
  for (int i = 0; i < ctx->mid_nch; i++) ctx->mid_ci[i].channeltype = channeltype_of(ctx->mid_type, i);
  for (int i = 0; i < ctx->out_nch; i++) ctx->out_ci[i].channeltype = channeltype_of(ctx->img2.imgtype, i);
  if (IW_IMGTYPE_HAS_ALPHA(ctx->mid_type))
    for (int i = 0; i < ctx->mid_nch; i++)
      if (ctx->mid_ci[i].channeltype != IW_CHANNELTYPE_ALPHA) ctx->mid_ci[i].premultiplied = 1;

  // ---- output depth and sample type (src/imagew-main.c:1514-1541)
  ctx->img2.sampletype = (ctx->output_profile & IW_PROFILE_HDRI) ? IW_SAMPLETYPE_FLOATINGPOINT : IW_SAMPLETYPE_UINT;
  if (ctx->img2.sampletype == IW_SAMPLETYPE_FLOATINGPOINT) {
    ctx->img2.bit_depth = 32;
  } else if (ctx->req_output_depth > 8 && (ctx->output_profile & IW_PROFILE_16BPS)) {
    ctx->img2.bit_depth = 16;
  } else {
    if (ctx->req_output_depth > 8) iw_warning(ctx, "Reducing depth to 8; required by the output format.");
    ctx->img2.bit_depth = 8;
  }

  // ---- per-channel max codes, colour counts, dither (src/imagew-main.c:2145-2203)
  if (ctx->img2.sampletype == IW_SAMPLETYPE_FLOATINGPOINT) {
    bool any = false;
    for (int i = 0; i < IWB_NUM_CHANNELTYPES; i++) any |= ctx->req_color_count[i] != 0;
    if (any) iw_warning(ctx, "Posterization is not supported with floating point output.");
  } else {
    const int full = (1 << ctx->img2.bit_depth) - 1;
    for (int i = 0; i < ctx->out_nch; i++) ctx->out_ci[i].maxcode = full;
    if (ctx->output_profile & IW_PROFILE_REDUCEDBITDEPTHS)
      for (int i = 0; i < ctx->out_nch; i++) {
        int m = ctx->req_output_maxcode[ctx->out_ci[i].channeltype];
        if (m > 0) ctx->out_ci[i].maxcode = m > full ? full : m;
      }
    for (int i = 0; i < ctx->out_nch; i++)
      if (ctx->out_ci[i].maxcode != full) {
        ctx->reduced_out_maxcode = 1;
        ctx->disable_out_tables = 1;
      }
  }
  for (int i = 0; i < ctx->out_nch; i++) {
    iwb_out_channel &o = ctx->out_ci[i];
    o.color_count = ctx->req_color_count[o.channeltype];
    if (o.color_count) {
      if (o.color_count < 2) o.color_count = 2;
      else if (o.color_count > o.maxcode) o.color_count = o.maxcode;
    }
    if (o.color_count == 1 + o.maxcode) o.color_count = 0;
    o.dither_family = ctx->dither_family[o.channeltype];
    o.dither_subtype = ctx->dither_subtype[o.channeltype];
    if (o.dither_family == IW_DITHERFAMILY_ERRDIFF) ctx->uses_errdiff = 1;
  }

  // ---- input background label -> linear (src/imagew-main.c:2209-2215)
  if (ctx->in_label_set) {
    for (int i = 0; i < 3; i++) ctx->in_label_lin.c[i] = iwb_to_linear(ctx->in_label_inputcs.c[i], &ctx->img1cs);
    ctx->in_label_lin.c[3] = ctx->in_label_inputcs.c[3];
  }
  set_background_colours(ctx);

  ctx->img2.rendering_intent = ctx->req_intent == IW_INTENT_UNKNOWN ? ctx->img1.rendering_intent : ctx->req_intent;

  if (ctx->axis[IW_DIMENSION_H].family == IW_RESIZETYPE_AUTO)
    choose_auto_filter(ctx, ctx->in_w, ctx->img2.width, IW_DIMENSION_H);
  if (ctx->axis[IW_DIMENSION_V].family == IW_RESIZETYPE_AUTO)
    choose_auto_filter(ctx, ctx->in_h, ctx->img2.height, IW_DIMENSION_V);
  return 1;
}
