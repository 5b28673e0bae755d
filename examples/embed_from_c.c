/* Pure C host of the vpb200 C ABI (no Python, no torch): load a program exported by tools/export_program.py, run it on
 * features [B, T, F] and print the embeddings' checksum.  This is what a C/C++ maintainer of a serving stack would write
 * against include/vpb200.h.
 *
 *   gcc -O2 -I include -I /usr/local/cuda/include examples/embed_from_c.c -o embed_from_c \
 *       -L voiceprintrecognition-pytorch_b200 -lvpb200 -L /usr/local/cuda/lib64 -lcudart -lm \
 *       -Wl,-rpath,$PWD/voiceprintrecognition-pytorch_b200
 *   ./embed_from_c /tmp/ecapa_b8.vpb [feats.f32]      (feats: B*T*F raw float32; default: a deterministic pattern)
 */
#include <cuda_runtime_api.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "vpb200.h"

#define CHECK_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 2; } } while (0)
#define CHECK_VP(h, x) do { int r_ = (x); if (r_ != VP_OK) { fprintf(stderr, "%s -> %d: %s\n", #x, r_, vp_last_error(h)); return 3; } } while (0)

int main(int argc, char** argv) {
  if (argc < 2) { fprintf(stderr, "usage: %s program.vpb [feats.f32]\n", argv[0]); return 1; }
  FILE* f = fopen(argv[1], "rb");
  if (!f) { perror(argv[1]); return 1; }
  char magic[8];
  int32_t hdr[8];
  uint64_t sz[4];
  if (fread(magic, 1, 8, f) != 8 || memcmp(magic, "VPB200P1", 8) != 0 || fread(hdr, 4, 8, f) != 8 || fread(sz, 8, 4, f) != 4) {
    fprintf(stderr, "%s: not a VPB200P1 file\n", argv[1]);
    return 1;
  }
  const int32_t n_ops = hdr[2], B = hdr[3], T = hdr[4], F = hdr[5], D = hdr[6];
  if (hdr[0] != vp_abi_version() || hdr[1] != vp_sizeof_op()) {
    fprintf(stderr, "ABI mismatch: file (abi %d, sizeof(vp_op) %d) vs library (%d, %d)\n", hdr[0], hdr[1], vp_abi_version(), vp_sizeof_op());
    return 1;
  }
  vp_op* ops = (vp_op*)malloc((size_t)n_ops * sizeof(vp_op));
  float* weights = (float*)malloc(sz[3]);
  if (!ops || !weights || fread(ops, sizeof(vp_op), (size_t)n_ops, f) != (size_t)n_ops || fread(weights, 1, sz[3], f) != sz[3]) {
    fprintf(stderr, "%s: truncated\n", argv[1]);
    return 1;
  }
  fclose(f);

  const size_t n_in = (size_t)B * T * F, n_out = (size_t)B * D;
  if (n_in != sz[1] || n_out != sz[2]) { fprintf(stderr, "header sizes disagree\n"); return 1; }
  float* h_feats = (float*)malloc(n_in * sizeof(float));
  if (argc > 2) {
    FILE* g = fopen(argv[2], "rb");
    if (!g || fread(h_feats, sizeof(float), n_in, g) != n_in) { fprintf(stderr, "%s: need %zu floats\n", argv[2], n_in); return 1; }
    fclose(g);
  } else {
    for (size_t i = 0; i < n_in; ++i) h_feats[i] = 2.0f * sinf(0.37f * (float)(i % 9973) + 0.001f * (float)(i / 9973));
  }

  vp_handle* h = NULL;
  int rc = vp_create(0, &h);
  if (rc != VP_OK) { fprintf(stderr, "vp_create -> %d (an sm_100 GPU is required; there is no CPU path)\n", rc); return 3; }
  CHECK_VP(h, vp_weights_load(h, weights, sz[3]));
  vp_program* prog = NULL;
  CHECK_VP(h, vp_program_create(h, ops, n_ops, sz[0], sz[1], sz[2], &prog));

  float *d_feats = NULL, *d_emb = NULL;
  cudaStream_t st;
  CHECK_CUDA(cudaStreamCreate(&st));
  CHECK_CUDA(cudaMalloc((void**)&d_feats, n_in * sizeof(float)));
  CHECK_CUDA(cudaMalloc((void**)&d_emb, n_out * sizeof(float)));
  CHECK_CUDA(cudaMemcpyAsync(d_feats, h_feats, n_in * sizeof(float), cudaMemcpyHostToDevice, st));
  CHECK_VP(h, vp_embed(prog, d_feats, d_emb, st));
  float* h_emb = (float*)malloc(n_out * sizeof(float));
  CHECK_CUDA(cudaMemcpyAsync(h_emb, d_emb, n_out * sizeof(float), cudaMemcpyDeviceToHost, st));
  CHECK_CUDA(cudaStreamSynchronize(st));

  double sum = 0.0, sq = 0.0;
  for (size_t i = 0; i < n_out; ++i) { sum += h_emb[i]; sq += (double)h_emb[i] * h_emb[i]; }
  printf("ops %d launches %d  B %d T %d F %d embd %d\n", n_ops, vp_program_launches(prog), B, T, F, D);
  printf("emb[0][0..3] = %.6f %.6f %.6f %.6f\n", h_emb[0], h_emb[1], h_emb[2], h_emb[3]);
  printf("checksum sum %.6f l2 %.6f\n", sum, sqrt(sq));
  if (argc > 3) {                                  /* optional: dump the embeddings for comparison */
    FILE* o = fopen(argv[3], "wb");
    if (o) { fwrite(h_emb, sizeof(float), n_out, o); fclose(o); }
  }
  vp_program_destroy(prog);
  vp_destroy(h);
  cudaFree(d_feats); cudaFree(d_emb); cudaStreamDestroy(st);
  free(ops); free(weights); free(h_feats); free(h_emb);
  return 0;
}
