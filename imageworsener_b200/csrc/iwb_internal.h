// iwb_internal.h -- host-side state of libimagew_b200: the context behind the libimagew C API and the
// flat, device-friendly plan that iw_process_image() hands to the CUDA kernels.
//
// Nothing here is shared with oracle/ (test infrastructure); the plan logic is this library's own
// restatement of the reference's decision code (reference: src/imagew-main.c:1463-2248).
#pragma once

#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/imagew_b200.h"

#define IWB_MSG_MAX 200                    // reference: src/imagew-internals.h:16
#define IWB_DEFAULT_MAX_DIMENSION 40000    // reference: src/imagew-internals.h:22
#define IWB_DEFAULT_MAX_MALLOC 2000000000  // reference: src/imagew-internals.h:23
#define IWB_NUM_CHANNELTYPES 5
#define IWB_MAX_OPTIONS 32                 // reference: src/imagew-api.c:998

enum { IWB_BKGD_EARLY = 1, IWB_BKGD_LATE = 2 };
enum { IWB_BKGD_SRC_NONE = 0, IWB_BKGD_SRC_FILE = 1, IWB_BKGD_SRC_REQ = 2 };

// Per-dimension resampling request (reference: struct iw_resize_settings, src/imagew-internals.h:38-50)
struct iwb_axis_request {
  int family = IW_RESIZETYPE_AUTO;
  int edge_policy = IW_EDGE_POLICY_STANDARD;
  int use_offset = 0;
  int per_channel_tables = 0;  // the reference's disable_rrctx_cache
  double param1 = 0.0, param2 = 0.0;
  double blur = 1.0;
  double out_true_size = 0.0;
  double translate = 0.0;
  double channel_offset[3] = {0.0, 0.0, 0.0};
};

struct iwb_in_channel {  // reference: struct iw_channelinfo_in
  int channeltype = 0;
  int slow_path = 0;     // disable_fast_get_sample
  int maxcode = 0;
};
struct iwb_mid_channel {  // reference: struct iw_channelinfo_intermed
  int channeltype = 0;
  int to_gray = 0;
  int in_channel = 0;
  int out_channel = 0;   // -1: resampled but not stored
  double early_bkgd = 0.0;
  int premultiplied = 0; // need_unassoc_alpha_processing
};
struct iwb_out_channel {  // reference: struct iw_channelinfo_out
  int channeltype = 0;
  int dither_family = 0, dither_subtype = 0;
  int color_count = 0;
  int maxcode = 0;
  double bkgd1 = 0.0, bkgd2 = 0.0;
};

struct iwb_option { std::string name, val; };

struct iwb_timings { double plan_ms = 0, h2d_ms = 0, kernel_ms = 0, d2h_ms = 0; };

struct iw_context {
  // ---- hooks / bookkeeping ----
  int caller_api_version = 0;
  int use_count = 0;
  void *userdata = nullptr;
  iw_mallocfn_type mallocfn = nullptr;
  iw_freefn_type freefn = nullptr;
  iw_translatefn_type translate_fn = nullptr;
  iw_warningfn_type warning_fn = nullptr;
  size_t max_malloc = IWB_DEFAULT_MAX_MALLOC;
  int max_width = IWB_DEFAULT_MAX_DIMENSION, max_height = IWB_DEFAULT_MAX_DIMENSION;
  int error_flag = 0;
  char *error_msg = nullptr;
  std::vector<iwb_option> options;

  // ---- requests (setters write these; prepare() reads them) ----
  unsigned int output_profile = 0;
  int req_output_depth = 0;
  int req_output_sample_type = 0;
  int req_output_maxcode[IWB_NUM_CHANNELTYPES] = {0, 0, 0, 0, 0};
  int req_color_count[IWB_NUM_CHANNELTYPES] = {0, 0, 0, 0, 0};
  int dither_family[IWB_NUM_CHANNELTYPES] = {0, 0, 0, 0, 0};
  int dither_subtype[IWB_NUM_CHANNELTYPES] = {0, 0, 0, 0, 0};
  int out_true_valid = 0;
  double out_true_w = 0.0, out_true_h = 0.0;
  int req_intent = 0;
  int out_cs_valid = 0;
  iw_csdescr req_out_cs{};
  int suppress_cslabel = 0;
  int negate_target = 0;
  int bkgd_valid = 0, req_checkerboard = 0;
  iw_color req_bkgd{}, req_bkgd2{};
  int out_label_valid = 0;
  iw_color req_out_label{};
  int use_label_from_file = 0, suppress_out_label = 0;
  int misc_values[64] = {0};  // codec-only integer settings, stored and returned verbatim
  int include_screen = 1;
  iwb_axis_request axis[2];   // indexed by IW_DIMENSION_*
  int to_grayscale = 0;
  int grayscale_formula = IW_GSF_STANDARD;
  double gray_w[3] = {0, 0, 0};
  int no_gamma = 0, intclamp = 0, pref_units = 0;
  int randomize = 0, random_seed = 0;
  unsigned char opt_flags[6] = {0, 1, 1, 1, 1, 1};  // indexed by IW_OPT_*
  int canvas_w = 0, canvas_h = 0;
  int crop_x = 0, crop_y = 0, crop_w = -1, crop_h = -1;
  int check_size = 0, check_origin[2] = {0, 0};
  int in_maxcode_req[4] = {0, 0, 0, 0};

  // ---- input image ----
  iw_image img1{};
  iw_csdescr img1cs{};
  int in_label_set = 0;
  iw_color in_label_inputcs{}, in_label_lin{};

  // ---- decisions (filled by iwb_prepare) ----
  iwb_in_channel in_ci[4];
  int in_type_logical = 0, in_nch_phys = 0, in_nch_logical = 0, in_alpha_index = 0;
  int in_maxcode_full = 0;
  int reduced_input_depths = 0;
  iwb_mid_channel mid_ci[4];
  int mid_type = 0, mid_nch = 0, mid_alpha_index = 0;
  iw_image img2{};
  iw_csdescr img2cs{};
  iwb_out_channel out_ci[4];
  int out_nch = 0;
  int uses_errdiff = 0;
  int apply_bkgd = 0, bkgd_strategy = 0, bkgd_checkerboard = 0, bkgd_source = 0;
  double bkgd1alpha = 0.0, bkgd2alpha = 0.0;
  int disable_out_tables = 0, reduced_out_maxcode = 0;
  int in_x0 = 0, in_y0 = 0, in_w = 0, in_h = 0;  // crop region in logical coordinates

  // ---- result view handed out by iw_get_output_image (the reference's optctx) ----
  int res_w = 0, res_h = 0, res_type = 0, res_depth = 0;
  size_t res_bpr = 0;
  const iw_byte *res_pixels = nullptr;
  int res_has_label = 0;
  unsigned int res_label[4] = {0, 0, 0, 0};
  int band_row0 = 0, band_row1 = 0;
  iwb_timings timings;
  struct iw_zlib_module *zlib_module = nullptr;   // carried for the file-format modules (src/imagew.h:720-724)
  // post-pass optimiser (optimize.h): the optimised image (when it differs from img2), its palette, colour key
  iw_byte *opt_pixels = nullptr;
  iw_palette *palette = nullptr;
  int palette_is_gray = 0;      // the palette is the full gray ramp of its depth (IW_VAL_OUTPUT_PALETTE_GRAYSCALE)
  int res_has_colorkey = 0;
  unsigned int res_colorkey[3] = {0, 0, 0};
};

struct iwb200_job;
void iwb_set_result_view(iw_context *ctx, int r0, int r1);                       // process.cpp
extern "C" int iwb_optimize_output(iw_context *ctx, iwb200_job **jobs, const int *bands, int n);   // pipeline.cu

// ---- page-locked host pool (pipeline.cu) ----
void *iwb_pinned_alloc(size_t n);   // nullptr when unavailable
bool iwb_pinned_free(void *p);      // true if p belonged to the pool

// ---- api.cpp ----
int iwb_imgtype_alpha_index(int t);
const char *iwb_option(const iw_context *ctx, const char *name);

// ---- colour-space scalar maths on the host (tables.cpp) ----
double iwb_to_linear(double v, const iw_csdescr *cs);     // reference: src/imagew-main.c:53-66
double iwb_from_linear(double v, const iw_csdescr *cs);   // reference: src/imagew-main.c:427-446
double iwb_gray_of(const iw_context *ctx, double r, double g, double b);  // src/imagew-main.c:278-305

// ---- plan.cpp ----
int iwb_prepare(iw_context *ctx);  // the decision logic; 1/0

// ---- weights.cpp ----
// One dimension's resampling recipe for one channel, in list order.
struct iwb_axis_table {
  int mode = 0;               // 0: weight list; 1: index map (nearest / null)
  int n_in = 0, n_out = 0;
  double edge_value = 0.0;    // value of a virtual (src<0) sample
  std::vector<int> first;     // [n_out+1] offsets into src/w          (mode 0)
  std::vector<int> src;       // source index per tap; -1 = virtual    (mode 0)
  std::vector<double> w;      // weight per tap                        (mode 0)
  std::vector<int> map;       // [n_out] source index, -2 = zero       (mode 1)
  int min_src(int o0, int o1) const;  // smallest real source index used by outputs [o0,o1); n_in if none
  int max_src(int o0, int o1) const;  // largest; -1 if none
  int max_taps() const;
};
// channeltype selects the per-channel offset and early-background edge value (reference:
// iwpvt_resize_rows_init, src/imagew-resize.c:412-556).  Returns 0 and sets the context error on failure.
int iwb_build_axis_table(iw_context *ctx, int dimension, int channeltype, int n_in, int n_out,
                         iwb_axis_table *out);

// ---- prng (tables.cpp) ----
struct iwb_affine { uint64_t mul, add; };  // t -> mul*t+add (mod p)
void iwb_prng_jump_table(iwb_affine tbl[64]);          // tbl[k] = step^(2^k)
uint64_t iwb_prng_first_state(int seed);               // t_1 (state after the first step), see prng.cuh
