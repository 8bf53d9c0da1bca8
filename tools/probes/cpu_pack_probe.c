// Host-side probe: how fast can the box's CPU strip the 15 padding bytes out of key-major Dpf::Cw rows?
// (Upper bound of any "pack on the CPU, copy half the bytes" variant of the host entry points.)
// gcc -O3 -march=native -fopenmp cpu_pack_probe.c -o cpu_pack_probe
#include <omp.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <immintrin.h>

int main(int argc, char **argv) {
  const size_t ncw = 33, nkeys = (size_t)1 << 20;            // 1.03 GiB of Cw rows
  const size_t in_bytes = nkeys * ncw * 32, out_row = ncw * 16 + 16;
  uint8_t *in = aligned_alloc(4096, in_bytes), *out = aligned_alloc(4096, nkeys * out_row);
  if (!in || !out) return 1;
#pragma omp parallel for schedule(static)
  for (size_t k = 0; k < nkeys; ++k) memset(in + k * ncw * 32, (int)(k & 0xff), ncw * 32);
#pragma omp parallel for schedule(static)
  for (size_t k = 0; k < nkeys; ++k) memset(out + k * out_row, 0, out_row);
  int tlist[] = {1, 2, 4, 8, 16, 32, 64};
  for (int ti = 0; ti < 7; ++ti) {
    const int nt = tlist[ti];
    if (nt > omp_get_num_procs() * 2) break;
    double best = 1e30;
    for (int rep = 0; rep < 3; ++rep) {
      const double t0 = omp_get_wtime();
#pragma omp parallel for schedule(static) num_threads(nt)
      for (size_t k = 0; k < nkeys; ++k) {
        const uint8_t *r = in + k * ncw * 32;
        uint8_t *o = out + k * out_row;
        uint64_t flags = 0;
        for (size_t i = 0; i < ncw; ++i) {
          _mm_stream_si128((__m128i *)(o + 16 * i), _mm_load_si128((const __m128i *)(r + 32 * i)));
          flags |= (uint64_t)(r[32 * i + 16] != 0) << i;
        }
        _mm_stream_si128((__m128i *)(o + 16 * ncw), _mm_set_epi64x(0, (long long)flags));
      }
      const double dt = omp_get_wtime() - t0;
      if (dt < best) best = dt;
    }
    printf("{\"threads\": %d, \"read_gbs\": %.1f, \"ms_per_4.4GB\": %.1f}\n", nt, in_bytes / best / 1e9, 4429185024.0 / (in_bytes / best) * 1e3);
  }
  printf("{\"procs\": %d}\n", omp_get_num_procs());
  return 0;
}
