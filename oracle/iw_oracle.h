/* iw_oracle.h -- TEST INFRASTRUCTURE ONLY.  Plain-C, single-threaded restatement of the reference's pixel
 * pipeline (iw_process_image and everything under it) behind one flat parameter struct.  It is the checker the
 * parity tests compare the CUDA path with; it is never linked, imported or called by the product
 * (imageworsener_b200/), which has its own, independently written host logic.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this restatement against (a) the 140 golden vectors taken
 * from the reference's own regression suite and (b) the unmodified reference build (oracle/_ref) on seeded
 * random jobs, byte for byte.
 */
#ifndef IW_ORACLE_H
#define IW_ORACLE_H
#include <stddef.h>
#include <stdint.h>

typedef struct {
  /* input image (reference: struct iw_image, src/imagew.h:377-410) */
  int imgtype, bit_depth, sampletype, width, height; /* width/height: logical (after orientation) */
  size_t bpr;
  const uint8_t *pixels;
  unsigned orient;
  int in_maxcode[4];         /* 0 = full depth */
  int cs_in_type; double cs_in_gamma;
  int in_label_set; double in_label[4];
  /* requests (the public setters of src/imagew-api.c) */
  unsigned profile;
  int out_depth_req;
  int out_maxcode_req[5], color_count_req[5];
  int dither_family[5], dither_subtype[5];
  int canvas_w, canvas_h;
  int out_true_valid; double out_true_w, out_true_h;
  int crop_x, crop_y, crop_w, crop_h;
  int cs_out_valid, cs_out_type; double cs_out_gamma;
  int family[2]; double blur[2], param1[2], param2[2]; int edge[2]; double translate[2];
  double chan_offset[2][3];
  int to_grayscale, gsf; double gray_w[3];
  int no_gamma, intclamp, negate;
  int bkgd_valid; double bkgd[4]; int checker_req; double bkgd2[4]; int check_size, check_ox, check_oy;
  int use_label_from_file, suppress_out_label, out_label_valid; double out_label[4];
  int random_seed;
  int max_width, max_height;
} iwo_params;

typedef struct {
  int imgtype, bit_depth, sampletype, width, height;
  size_t bpr;
  uint8_t *pixels;           /* malloc'd; release with iwo_free_result */
  int cs_out_type; double cs_out_gamma;
  int reduced_maxcolors; unsigned maxcolorcode[5];
  int has_bkgdlabel; double bkgdlabel[4];
  /* last processed channel's FP32 intermediate (input_w x out_h), for intermediate-parity checks */
  float *last_intermediate; int inter_w, inter_h;
} iwo_result;

void iwo_default_params(iwo_params *p);
int iwo_process(const iwo_params *p, iwo_result *r, char *err, int errlen); /* 1 = ok, 0 = error (text in err) */
void iwo_free_result(iwo_result *r);

/* scalar helpers exposed for unit tests */
double iwo_to_linear(double v, int cstype, double gamma);
double iwo_from_linear(double v, int cstype, double gamma);
uint32_t iwo_prng_nth(int seed, uint64_t n);  /* n-th draw (0-based) by direct iteration */
int iwo_weightlist(const iwo_params *p, int dimension, int channeltype, int n_in, int n_out, double out_true,
                   int apply_bkgd_early, double bkgd_value, int *src, int *dst, double *w, int cap);
/* synthetic input and fingerprints (bench.py --impl reference, tests) */
void iwo_synth_xorshift(uint8_t *buf, size_t n, uint64_t seed);
uint64_t iwo_fnv1a64(const uint8_t *buf, size_t n);
#endif
