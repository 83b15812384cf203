/* C99 host program on the C ABI of include/glasdi_b200.h: no Python, no torch -- plain pointers and sizes.
 * Reads one problem (decoder weights, coefficient draws, initial conditions) from a binary file written by
 * tests/test_gpu_api.py::test_c_host_program_on_the_c_abi, runs the fused greedy step and prints the selected
 * parameter index and the per-parameter maximum standard deviations as text.
 *
 *   gcc -std=c99 abi_smoke.c -I include -I $CUDA/include -L gplasdi_b200/lib -lglasdi_b200 -L $CUDA/lib64 -lcudart
 *
 * File layout (little endian): int32 n_linear, act_id, P, S, n, T; double dt; int32 widths[n_linear + 1];
 * per layer: float W[out * in], float b[out]; double coefs[P * S * n * (n + 1)]; float z0[P * n]. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <cuda_runtime_api.h>

#include "glasdi_b200.h"

#define CHECK(call)                                                              \
  do {                                                                           \
    int rc_ = (call);                                                            \
    if (rc_ != GLASDI_OK) {                                                      \
      fprintf(stderr, "%s -> %d: %s\n", #call, rc_, glasdi_last_error(h));       \
      return 2;                                                                  \
    }                                                                            \
  } while (0)

static int read_exact(FILE* f, void* dst, size_t bytes) { return fread(dst, 1, bytes, f) == bytes ? 0 : 1; }

int main(int argc, char** argv) {
  if (argc < 2) { fprintf(stderr, "usage: %s problem.bin\n", argv[0]); return 1; }
  FILE* f = fopen(argv[1], "rb");
  if (!f) { perror("open"); return 1; }
  int32_t hdr[6];
  double dt;
  if (read_exact(f, hdr, sizeof(hdr)) || read_exact(f, &dt, sizeof(dt))) return 1;
  const int n_linear = hdr[0], act = hdr[1], P = hdr[2], S = hdr[3], n = hdr[4], T = hdr[5];
  int32_t widths[17];
  if (n_linear < 1 || n_linear > 16 || read_exact(f, widths, sizeof(int32_t) * (size_t)(n_linear + 1))) return 1;
  float* W[16];
  float* b[16];
  for (int l = 0; l < n_linear; ++l) {
    const size_t nw = (size_t)widths[l] * widths[l + 1];
    W[l] = (float*)malloc(nw * sizeof(float));
    b[l] = (float*)malloc((size_t)widths[l + 1] * sizeof(float));
    if (read_exact(f, W[l], nw * sizeof(float)) || read_exact(f, b[l], (size_t)widths[l + 1] * sizeof(float))) return 1;
  }
  const size_t nc = (size_t)P * S * n * (n + 1), nz = (size_t)P * n;
  double* coefs = (double*)malloc(nc * sizeof(double));
  float* z0 = (float*)malloc(nz * sizeof(float));
  if (read_exact(f, coefs, nc * sizeof(double)) || read_exact(f, z0, nz * sizeof(float))) return 1;
  fclose(f);

  glasdi_handle* h = NULL;
  if (glasdi_create(&h, 0) != GLASDI_OK) { fprintf(stderr, "glasdi_create failed (no sm_100 device?)\n"); return 3; }
  CHECK(glasdi_set_decoder(h, n_linear, widths, (const float* const*)W, (const float* const*)b, act));

  double* d_coefs; float* d_z0; float* d_ms; int64_t* d_am; float* d_mv;
  if (cudaMalloc((void**)&d_coefs, nc * sizeof(double)) || cudaMalloc((void**)&d_z0, nz * sizeof(float)) ||
      cudaMalloc((void**)&d_ms, (size_t)P * sizeof(float)) || cudaMalloc((void**)&d_am, sizeof(int64_t)) ||
      cudaMalloc((void**)&d_mv, sizeof(float))) return 4;
  cudaMemcpy(d_coefs, coefs, nc * sizeof(double), cudaMemcpyHostToDevice);
  cudaMemcpy(d_z0, z0, nz * sizeof(float), cudaMemcpyHostToDevice);
  CHECK(glasdi_greedy_step(h, d_coefs, d_z0, P, S, n, T, dt, d_ms, d_am, d_mv, NULL /* default stream */));
  CHECK(glasdi_check_numeric(h, NULL));

  float* ms = (float*)malloc((size_t)P * sizeof(float));
  int64_t am = -2;
  cudaMemcpy(ms, d_ms, (size_t)P * sizeof(float), cudaMemcpyDeviceToHost);
  cudaMemcpy(&am, d_am, sizeof(int64_t), cudaMemcpyDeviceToHost);
  printf("index %lld launches %lld\n", (long long)am, (long long)glasdi_launch_count(h));
  for (int p = 0; p < P; ++p) printf("%.9e\n", (double)ms[p]);
  cudaFree(d_coefs); cudaFree(d_z0); cudaFree(d_ms); cudaFree(d_am); cudaFree(d_mv);
  glasdi_destroy(h);
  return 0;
}
