/* ORACLE (test infrastructure only): plain-C restatement of the MaskGitVQGAN vector quantiser search,
 * reference VectorQuantizer.compute_distances + argmin, muse/modeling_maskgit_vqgan.py:303-316,342-348:
 *     d[r][c] = (|z_r|^2 + |e_c|^2) - 2 <z_r, e_c> ;  id[r] = first argmin_c d[r][c]
 * with the arithmetic pinned exactly as csrc/vq.cu pins it, so CUDA == oracle bit-for-bit on any input:
 *     norms and dot products are ascending-k fmaf chains starting from 0,
 *     d = fmaf(-2, dot, (float)(znorm + enorm)), ties resolved to the lowest index.
 * (The reference evaluates the same formula through BLAS addmm, whose summation order is unspecified;
 *  the oracle is pinned to the reference on margin-screened golden inputs, tests/golden/vq_quantizer.pt.)
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC -o _build/libvq_oracle.so vq_oracle.c -lm
 */
#include <math.h>
#include <stdint.h>

static float sqnorm(const float* x, int D) {
  float acc = 0.0f;
  for (int k = 0; k < D; ++k) acc = fmaf(x[k], x[k], acc);
  return acc;
}

/* z: [n, D] row-major, codebook: [ncodes, D]; ids: [n] int64; dmin (nullable): [n] */
void vq_oracle_argmin(const float* z, const float* codebook, int64_t* ids, float* dmin, int n, int ncodes, int D,
                      float* enorm_ws /* [ncodes] scratch */) {
  for (int c = 0; c < ncodes; ++c) enorm_ws[c] = sqnorm(codebook + (int64_t)c * D, D);
  for (int r = 0; r < n; ++r) {
    const float* zr = z + (int64_t)r * D;
    const float zn = sqnorm(zr, D);
    float best = INFINITY;
    int64_t bi = 0;
    for (int c = 0; c < ncodes; ++c) {
      const float* e = codebook + (int64_t)c * D;
      float dot = 0.0f;
      for (int k = 0; k < D; ++k) dot = fmaf(zr[k], e[k], dot);
      const float base = zn + enorm_ws[c];
      const float d = fmaf(-2.0f, dot, base);
      if (d < best) { best = d; bi = c; }
    }
    ids[r] = bi;
    if (dmin) dmin[r] = best;
  }
}
