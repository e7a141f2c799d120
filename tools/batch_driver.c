/* batch_driver.c -- a plain C caller of the libimagew API: a batch of thumbnails (BASELINE cfg4: 1920x1080 RGB8 ->
 * 256x144) from T host threads, one context per image, exactly the calls an application linking the reference
 * library makes (src/imagew.h).  Measures images/s through iw_process_image() on host buffers.
 *
 *   gcc -O2 -pthread -Iinclude tools/batch_driver.c -o tools/batch_driver \
 *       -Limageworsener_b200 -l:libimagew_b200.so -Wl,-rpath,'$ORIGIN/../imageworsener_b200'
 *   tools/batch_driver [images=256] [threads=4] [family: 0 auto | 1 box]
 */
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "imagew_b200.h"

enum { IN_W = 1920, IN_H = 1080, OUT_W = 256, OUT_H = 144 };

static int g_images = 256, g_threads = 4, g_family = 0;
static struct iw_context **g_ctx;
static volatile int g_failed = 0;

static double now_ms(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec * 1e3 + ts.tv_nsec / 1e6;
}

static struct iw_context *make_context(int i) {
  struct iw_context *ctx = iw_create_context(NULL);
  struct iw_image img;
  if (!ctx) return NULL;
  memset(&img, 0, sizeof img);
  img.imgtype = IW_IMGTYPE_RGB;
  img.bit_depth = 8;
  img.sampletype = IW_SAMPLETYPE_UINT;
  img.width = IN_W;
  img.height = IN_H;
  img.bpr = (size_t)IN_W * 3;
  img.pixels = (iw_byte *)iw_malloc_large(ctx, img.bpr, IN_H);     /* the context takes ownership */
  if (!img.pixels) { iw_destroy_context(ctx); return NULL; }
  iwb200_synth_xorshift(img.pixels, img.bpr * IN_H, 0x9E3779B97F4A7C15ull + (uint64_t)(i % 16));
  img.density_code = IW_DENSITY_UNKNOWN;
  iw_set_input_image(ctx, &img);
  iw_set_output_profile(ctx, iw_get_profile_by_fmt(IW_FORMAT_PNG));
  iw_set_output_canvas_size(ctx, OUT_W, OUT_H);
  iw_set_output_depth(ctx, 8);
  if (g_family == 1) {
    iw_set_resize_alg(ctx, IW_DIMENSION_H, IW_RESIZETYPE_BOX, 1.0, 0.0, 0.0);
    iw_set_resize_alg(ctx, IW_DIMENSION_V, IW_RESIZETYPE_BOX, 1.0, 0.0, 0.0);
  }
  for (int o = IW_OPT_GRAYSCALE; o <= IW_OPT_BINARY_TRNS; o++) iw_set_allow_opt(ctx, o, 0);
  return ctx;
}

static int run_one(struct iw_context *ctx) {
  struct iw_image out;
  int ok = iw_process_image(ctx);
  if (ok) {
    iw_get_output_image(ctx, &out);
    ok = out.width == OUT_W && out.height == OUT_H && out.pixels != NULL;
  } else {
    fprintf(stderr, "iw_process_image failed: %s\n", iw_get_errormsg(ctx, (char[200]){0}, 200));
  }
  iw_destroy_context(ctx);
  return ok;
}

static void *worker(void *arg) {
  const int t = (int)(size_t)arg;
  for (int i = t; i < g_images; i += g_threads)
    if (!run_one(g_ctx[i])) g_failed = 1;
  return NULL;
}

int main(int argc, char **argv) {
  if (argc > 1) g_images = atoi(argv[1]);
  if (argc > 2) g_threads = atoi(argv[2]);
  if (argc > 3) g_family = atoi(argv[3]);
  if (g_images < 1 || g_threads < 1 || g_threads > 64) return 2;
  for (int i = 0; i < 8; i++) {            /* warm-up: CUDA context, plan cache, buffer pools */
    struct iw_context *c = make_context(i);
    if (!c || !run_one(c)) { fprintf(stderr, "warm-up failed\n"); return 1; }
  }
  g_ctx = (struct iw_context **)calloc((size_t)g_images, sizeof *g_ctx);
  for (int i = 0; i < g_images; i++)
    if (!(g_ctx[i] = make_context(i))) { fprintf(stderr, "context %d failed\n", i); return 1; }
  pthread_t th[64];
  const double t0 = now_ms();
  for (int t = 0; t < g_threads; t++) pthread_create(&th[t], NULL, worker, (void *)(size_t)t);
  for (int t = 0; t < g_threads; t++) pthread_join(th[t], NULL);
  const double dt = now_ms() - t0;
  printf("%s: %d images, %d threads: %.1f ms = %.4f ms/image, %.0f images/s, %.1f GB/s of input%s\n",
         g_family == 1 ? "cfg4box" : "cfg4auto", g_images, g_threads, dt, dt / g_images, g_images / dt * 1e3,
         (double)g_images * IN_W * IN_H * 3 / dt / 1e6, g_failed ? "  (FAILURES)" : "");
  free(g_ctx);
  return g_failed ? 1 : 0;
}
