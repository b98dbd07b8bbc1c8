/* corr_c.c -- TEST INFRASTRUCTURE (oracle): plain-C restatement of the reference's correlation lookup,
 * devo/altcorr/correlation_kernel.cu:82-136 (window dot products, out-of-bounds => 0) followed by the host-side
 * bilinear blend and permute of corr_cuda_forward (:221-232).  One loop nest per CUDA thread index of the reference;
 * OpenMP over edges.  Used only by tests/ (pinned against oracle/corr.py, itself pinned against the reference's compiled
 * extension on the GPU box) and by bench.py's CPU-baseline / --impl reference legs.  float32 in, float32 accumulate,
 * float32 out, like the reference's kernel instantiated for float.
 *
 *   fmap1  [Np][C][P][P]      fmap2 [Nf][C][H][W]      coords [E][2][P][P] (x then y, already divided by the level scale)
 *   out    [E][D1(x-off)][D1(y-off)][P][P],  D1 = 2r+1   (the layout cuda_corr.forward returns after its permute)
 * Build: gcc -O3 -march=native -fopenmp -shared -fPIC oracle/corr_c.c -o oracle/_build/libcorr_oracle.so -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

int corr_oracle_forward(const float* fmap1, const float* fmap2, const float* coords, const int64_t* ii, const int64_t* jj,
                        float* out, int E, int C, int H, int W, int P, int radius) {
  const int D = 2 * radius + 2, D1 = 2 * radius + 1, PP = P * P;
  if (D > 16 || P > 4) return -1;
#pragma omp parallel
  {
    float* vol = (float*)malloc(sizeof(float) * D * D);            /* V[a][b] of one patch pixel (:118-134) */
    float* win = (float*)malloc(sizeof(float) * (size_t)C);
#pragma omp for schedule(static)
    for (int e = 0; e < E; e++) {
      const float* f1 = fmap1 + (size_t)ii[e] * C * PP;
      const float* f2 = fmap2 + (size_t)jj[e] * C * H * W;
      for (int i0 = 0; i0 < P; i0++)
        for (int j0 = 0; j0 < P; j0++) {
          const float x = coords[((size_t)e * 2 + 0) * PP + i0 * P + j0];
          const float y = coords[((size_t)e * 2 + 1) * PP + i0 * P + j0];
          const int fx = (int)floorf(x), fy = (int)floorf(y);
          for (int c = 0; c < C; c++) win[c] = f1[(size_t)c * PP + i0 * P + j0];
          for (int a = 0; a < D; a++)                               /* a: row (y) offset, b: column (x) offset */
            for (int b = 0; b < D; b++) {
              const int i1 = fy + a - radius, j1 = fx + b - radius;
              float s = 0.f;
              if (i1 >= 0 && i1 < H && j1 >= 0 && j1 < W) {
                const float* p2 = f2 + (size_t)i1 * W + j1;
                for (int c = 0; c < C; c++) s += win[c] * p2[(size_t)c * H * W];
              }
              vol[a * D + b] = s;
            }
          const float dx = x - floorf(x), dy = y - floorf(y);       /* :218-230 */
          for (int a = 0; a < D1; a++)
            for (int b = 0; b < D1; b++) {
              const float v = (1 - dx) * (1 - dy) * vol[a * D + b] + dx * (1 - dy) * vol[a * D + b + 1] +
                              (1 - dx) * dy * vol[(a + 1) * D + b] + dx * dy * vol[(a + 1) * D + b + 1];
              out[((((size_t)e * D1 + b) * D1 + a) * P + i0) * P + j0] = v;   /* permute(0,1,3,2,4,5): x-offset first */
            }
        }
    }
    free(vol);
    free(win);
  }
  return 0;
}
