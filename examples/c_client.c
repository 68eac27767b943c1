/* Plain-C client of libcompactb200.so: what a non-Python host (or a cgo / JNI / FFI binding) does with the
 * drop-in boundary declared in include/compactb200.h.  No torch, no C++: C99 + the CUDA runtime C API.
 *
 *   c_client              one BINARY residual round trip on host buffers through cf_host_compress /
 *                         cf_host_decompress (the work of compact_compress + compact_decompress,
 *                         main.py:169-270, :322-388), then the parity properties below
 *   c_client --no-gpu     only the calls that need no device: version, workspace sizes, argument errors
 *   c_client --check F    run the property checker on a dump [N, C as int64 | x | base | payload | recon]
 *                         (tests feed it oracle outputs: the checker itself is checked on the CPU)
 *
 * Properties (size independent, bit exact):
 *   1. sign bit (n, c) of the payload == (fp16(x - base) >= 0)                      fastpath.py:57-72
 *   2. recon == fp16(base +- fp16(U[n] * V[c]))                                     fastpath.py:109-116
 *   3. receiver reconstruction == sender's error-feedback base                      main.py:17-34
 *
 * Build:  gcc -std=c99 -O2 -Iinclude -I/usr/local/cuda/include examples/c_client.c -o c_client \
 *             -Lcompactfusion_b200 -lcompactb200 -L/usr/local/cuda/lib64 -lcudart -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "compactb200.h"

#ifndef CF_CLIENT_NO_CUDA
#include <cuda_runtime_api.h>
#endif

/* IEEE binary16 <-> binary32, round to nearest even */
static float h2f(uint16_t h) {
  const uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
  const uint32_t exp = (h >> 10) & 0x1Fu;
  const uint32_t man = h & 0x3FFu;
  uint32_t bits;
  if (exp == 0) {
    if (man == 0) {
      bits = sign;
    } else { /* subnormal: man * 2^-24 */
      float f = (float)man * (1.0f / 16777216.0f);
      memcpy(&bits, &f, 4);
      bits |= sign;
    }
  } else if (exp == 31) {
    bits = sign | 0x7F800000u | (man << 13);
  } else {
    bits = sign | ((exp + 112u) << 23) | (man << 13);
  }
  float out;
  memcpy(&out, &bits, 4);
  return out;
}

static uint16_t f2h(float f) {
  uint32_t x;
  memcpy(&x, &f, 4);
  const uint16_t sign = (uint16_t)((x >> 16) & 0x8000u);
  const uint32_t absx = x & 0x7FFFFFFFu;
  if (absx >= 0x7F800000u) return (uint16_t)(sign | 0x7C00u | (absx > 0x7F800000u ? 0x200u : 0u)); /* inf / nan */
  if (absx >= 0x477FF000u) return (uint16_t)(sign | 0x7C00u); /* >= 65520 rounds to inf */
  if (absx < 0x33000001u) return sign;                        /* <= 2^-25 rounds to zero (tie to even) */
  int32_t e = (int32_t)(absx >> 23) - 127;
  uint32_t m = (absx & 0x7FFFFFu) | 0x800000u; /* 24-bit significand */
  int shift = (e < -14) ? (13 + (-14 - e)) : 13; /* bits dropped */
  uint32_t kept = m >> shift;
  const uint32_t rem = m & ((1u << shift) - 1u);
  const uint32_t half = 1u << (shift - 1);
  if (rem > half || (rem == half && (kept & 1u))) kept += 1u;
  if (e < -14) return (uint16_t)(sign | kept); /* subnormal (a carry into 0x400 is the smallest normal) */
  /* kept has the implicit bit at 0x400; a carry to 0x800 bumps the exponent */
  uint32_t he = (uint32_t)(e + 15);
  if (kept & 0x800u) {
    kept >>= 1;
    he += 1;
  }
  if (he >= 31) return (uint16_t)(sign | 0x7C00u);
  return (uint16_t)(sign | (he << 10) | (kept & 0x3FFu));
}

/* returns the number of violated properties (prints the first offender of each) */
static int check_binary(const uint16_t* x, const uint16_t* base, const unsigned char* payload, const uint16_t* recon,
                        const uint16_t* new_base, int64_t N, int64_t C) {
  const unsigned char* codes = payload;
  const uint16_t* U = (const uint16_t*)(payload + N * C / 8);
  const uint16_t* V = U + N;
  int bad_sign = 0, bad_recon = 0, bad_ef = 0;
  for (int64_t n = 0; n < N; ++n) {
    const float u = h2f(U[n]);
    for (int64_t c = 0; c < C; ++c) {
      const int64_t i = n * C + c;
      const float d = h2f(f2h(h2f(x[i]) - h2f(base[i])));
      const int bit = (codes[n * (C / 8) + c / 8] >> (c % 8)) & 1;
      if (bit != (d >= 0.0f)) {
        if (!bad_sign) fprintf(stderr, "sign bit (%lld,%lld): payload %d, delta %g\n", (long long)n, (long long)c, bit, d);
        bad_sign = 1;
      }
      const float p = h2f(f2h(u * h2f(V[c])));
      const uint16_t want = f2h(h2f(base[i]) + (bit ? p : -p));
      if (recon[i] != want) {
        if (!bad_recon) fprintf(stderr, "recon (%lld,%lld): 0x%04x, expected 0x%04x\n", (long long)n, (long long)c, recon[i], want);
        bad_recon = 1;
      }
      if (new_base != NULL && new_base[i] != recon[i]) {
        if (!bad_ef) fprintf(stderr, "sender base != receiver recon at (%lld,%lld)\n", (long long)n, (long long)c);
        bad_ef = 1;
      }
    }
  }
  return bad_sign + bad_recon + bad_ef;
}

static int check_file(const char* path) {
  FILE* f = fopen(path, "rb");
  if (!f) { perror(path); return 2; }
  int64_t dims[2];
  if (fread(dims, 8, 2, f) != 2) { fclose(f); return 2; }
  const int64_t N = dims[0], C = dims[1], E = N * C;
  const size_t pbytes = (size_t)(E / 8 + 2 * N + 2 * C);
  uint16_t* x = malloc(2 * E); uint16_t* base = malloc(2 * E); uint16_t* recon = malloc(2 * E);
  unsigned char* payload = malloc(pbytes);
  int ok = fread(x, 2, E, f) == (size_t)E && fread(base, 2, E, f) == (size_t)E && fread(payload, 1, pbytes, f) == pbytes &&
           fread(recon, 2, E, f) == (size_t)E;
  fclose(f);
  if (!ok) { fprintf(stderr, "short file\n"); return 2; }
  const int bad = check_binary(x, base, payload, recon, NULL, N, C);
  printf(bad ? "C_CHECK_FAILED %d\n" : "C_CHECK_OK %d\n", bad);
  free(x); free(base); free(recon); free(payload);
  return bad ? 1 : 0;
}

static int no_gpu_calls(void) {
  if (cf_abi_version() != CF_ABI_VERSION) { fprintf(stderr, "ABI version mismatch\n"); return 1; }
  const size_t ws1 = cf_workspace_bytes(CF_CODEC_BINARY, 4608, 3072, 0, 1);
  const size_t ws2 = cf_workspace_bytes(CF_CODEC_BINARY, 4608, 3072, 0, 2);
  if (ws1 == 0 || ws2 != 2 * ws1) { fprintf(stderr, "workspace sizes: %zu %zu\n", ws1, ws2); return 1; }
  if (cf_host_scratch_bytes(CF_CODEC_BINARY, 544, 3072) < (size_t)(3 * 544 * 3072 * 2)) { fprintf(stderr, "scratch too small\n"); return 1; }
  /* argument errors come back as a status + message, never as a crash */
  if (cf_lse_merge(NULL, NULL, NULL, NULL, NULL, 1, 1, 1, 4, NULL) != CF_ERR_ARG || strlen(cf_last_error()) == 0) return 1;
  if (cf_host_compress(CF_CODEC_BINARY, NULL, NULL, NULL, NULL, 544, 3072, NULL, 0, NULL) == CF_OK) return 1;
  printf("C_NO_GPU_OK abi=%d workspace=%zu\n", cf_abi_version(), ws1);
  return 0;
}

#ifndef CF_CLIENT_NO_CUDA
static int round_trip(void) {
  const int64_t N = 544, C = 3072, E = N * C;
  const size_t pbytes = (size_t)(E / 8 + 2 * N + 2 * C);
  uint16_t* x = malloc(2 * E); uint16_t* base = malloc(2 * E); uint16_t* nb = malloc(2 * E); uint16_t* recon = malloc(2 * E);
  unsigned char* payload = malloc(pbytes);
  uint32_t s = 12345u; /* LCG: activations ~ U(-2, 2), base = x + small residual */
  for (int64_t i = 0; i < E; ++i) {
    s = s * 1664525u + 1013904223u;
    const float a = ((float)(s >> 8) / 16777216.0f - 0.5f) * 4.0f;
    s = s * 1664525u + 1013904223u;
    const float r = ((float)(s >> 8) / 16777216.0f - 0.5f) * 0.5f;
    x[i] = f2h(a);
    base[i] = f2h(a + r);
  }
  void* scratch = NULL;
  const size_t sbytes = cf_host_scratch_bytes(CF_CODEC_BINARY, N, C);
  if (cudaMalloc(&scratch, sbytes) != cudaSuccess) { fprintf(stderr, "cudaMalloc failed\n"); return 1; }
  int rc = cf_host_compress(CF_CODEC_BINARY, x, base, nb, payload, N, C, scratch, sbytes, NULL);
  if (rc != CF_OK) { fprintf(stderr, "cf_host_compress: %d %s\n", rc, cf_last_error()); return 1; }
  rc = cf_host_decompress(CF_CODEC_BINARY, payload, base, recon, N, C, scratch, sbytes, NULL);
  if (rc != CF_OK) { fprintf(stderr, "cf_host_decompress: %d %s\n", rc, cf_last_error()); return 1; }
  cudaFree(scratch);
  const int bad = check_binary(x, base, payload, recon, nb, N, C);
  if (bad)
    printf("C_CLIENT_FAILED %d\n", bad);
  else
    printf("C_CLIENT_OK %lldx%lld payload=%zu bytes (%.1fx smaller)\n", (long long)N, (long long)C, pbytes,
           2.0 * (double)E / (double)pbytes);
  free(x); free(base); free(nb); free(recon); free(payload);
  return bad ? 1 : 0;
}
#endif

int main(int argc, char** argv) {
  if (argc >= 3 && strcmp(argv[1], "--check") == 0) return check_file(argv[2]);
  if (no_gpu_calls() != 0) return 1;
  if (argc >= 2 && strcmp(argv[1], "--no-gpu") == 0) return 0;
#ifndef CF_CLIENT_NO_CUDA
  return round_trip();
#else
  return 0;
#endif
}
