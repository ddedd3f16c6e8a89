/*
 * kvz_cuda_encode.c -- minimal libkvazaar host used by the drop-in tests: encodes a raw I420 file through the
 * UNCHANGED public API (kvz_api_get / config_parse / encoder_open / encoder_encode, ref: kvazaar.h:664-829) and,
 * when --cuda is given, binds the cuda strategies into the running library right after encoder_open
 * (kvz_cuda_overlay_install, strategies-cuda-glue.c).  Everything else -- search, RDOQ, CABAC, bitstream -- is the
 * reference's own host code.  Output must be byte-identical with and without --cuda.
 *
 *   kvz_cuda_encode [--cuda[=type,type,...]] in.yuv WxH out.hevc [key=value ...]      (keys as in kvazaar --help)
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "kvazaar.h"

int kvz_cuda_overlay_install(const char *only);

int main(int argc, char **argv)
{
  int use_cuda = 0;
  const char *only = NULL;
  int a = 1;
  if (a < argc && strncmp(argv[a], "--cuda", 6) == 0) { use_cuda = 1; if (argv[a][6] == '=') only = argv[a] + 7; ++a; }
  if (argc - a < 3) { fprintf(stderr, "usage: %s [--cuda[=types]] in.yuv WxH out.hevc [key=value ...]\n", argv[0]); return 2; }
  const char *in = argv[a], *res = argv[a + 1], *out = argv[a + 2];
  int w = 0, h = 0;
  if (sscanf(res, "%dx%d", &w, &h) != 2) { fprintf(stderr, "bad resolution %s\n", res); return 2; }

  const kvz_api *api = kvz_api_get(KVZ_BIT_DEPTH);       /* the 10-bit build (integration/Makefile: kvz_cuda_encode_10b) links libkvazaar_ref_10b.so */
  kvz_config *cfg = api->config_alloc();
  api->config_init(cfg);
  char wh[32];
  snprintf(wh, sizeof(wh), "%d", w); api->config_parse(cfg, "width", wh);
  snprintf(wh, sizeof(wh), "%d", h); api->config_parse(cfg, "height", wh);
  for (int i = a + 3; i < argc; ++i) {
    char *eq = strchr(argv[i], '=');
    if (!eq) { fprintf(stderr, "expected key=value, got %s\n", argv[i]); return 2; }
    *eq = 0;
    if (!api->config_parse(cfg, argv[i], eq + 1)) { fprintf(stderr, "config_parse(%s, %s) failed\n", argv[i], eq + 1); return 2; }
  }
  cfg->enable_logging_output = 0;
  kvz_encoder *enc = api->encoder_open(cfg);
  if (!enc) { fprintf(stderr, "encoder_open failed\n"); return 1; }
  if (use_cuda) {
    const int n = kvz_cuda_overlay_install(only);
    if (n <= 0) { fprintf(stderr, "cuda overlay failed\n"); return 1; }
    fprintf(stderr, "kvz-cuda: %d strategy pointers bound to libkvzcuda\n", n);
  }

  FILE *fi = fopen(in, "rb"), *fo = fopen(out, "wb");
  if (!fi || !fo) { fprintf(stderr, "cannot open files\n"); return 1; }
  const size_t ysz = (size_t)w * h, csz = ysz / 4;
  int frames_in = 0, frames_out = 0, eof = 0;
  for (;;) {
    kvz_picture *pic = NULL;
    if (!eof) {
      pic = api->picture_alloc(w, h);
#if KVZ_BIT_DEPTH == 8
      if (fread(pic->y, 1, ysz, fi) != ysz || fread(pic->u, 1, csz, fi) != csz || fread(pic->v, 1, csz, fi) != csz) {
        api->picture_free(pic); pic = NULL; eof = 1;
      } else ++frames_in;
#else
      /* 8-bit input file into the wider kvz_pixel, scaled to the internal bit depth (what the CLI's reader does for
       * --input-bitdepth 8, ref: src/yuv_io.c) */
      {
        static unsigned char *buf = NULL;
        if (!buf) buf = malloc(ysz + 2 * csz);
        if (fread(buf, 1, ysz + 2 * csz, fi) != ysz + 2 * csz) { api->picture_free(pic); pic = NULL; eof = 1; }
        else {
          for (int r = 0; r < h; ++r) for (int c = 0; c < w; ++c) pic->y[(size_t)r * pic->stride + c] = (kvz_pixel)(buf[(size_t)r * w + c] << (KVZ_BIT_DEPTH - 8));
          for (int r = 0; r < h / 2; ++r) for (int c = 0; c < w / 2; ++c) {
            pic->u[(size_t)r * (pic->stride / 2) + c] = (kvz_pixel)(buf[ysz + (size_t)r * (w / 2) + c] << (KVZ_BIT_DEPTH - 8));
            pic->v[(size_t)r * (pic->stride / 2) + c] = (kvz_pixel)(buf[ysz + csz + (size_t)r * (w / 2) + c] << (KVZ_BIT_DEPTH - 8));
          }
          ++frames_in;
        }
      }
#endif
    }
    kvz_data_chunk *chunks = NULL;
    uint32_t len = 0;
    kvz_picture *rec = NULL;
    kvz_frame_info info;
    if (!api->encoder_encode(enc, pic, &chunks, &len, &rec, NULL, &info)) { fprintf(stderr, "encode failed\n"); return 1; }
    if (pic) api->picture_free(pic);
    if (chunks) {
      for (kvz_data_chunk *c = chunks; c; c = c->next) fwrite(c->data, 1, c->len, fo);
      api->chunk_free(chunks);
      ++frames_out;
    }
    if (rec) api->picture_free(rec);
    if (eof && !chunks) break;
  }
  fclose(fi); fclose(fo);
  api->encoder_close(enc);
  api->config_destroy(cfg);
  fprintf(stderr, "encoded %d frames (%d in)\n", frames_out, frames_in);
  return 0;
}
