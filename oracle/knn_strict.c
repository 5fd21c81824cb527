/* TEST INFRASTRUCTURE ONLY - strict fp32 restatement of graph construction G1-G3 (SURVEY.md section 9,
 * GRAPH_SPEC_VERSION 2).  PARITY UNPINNED: the reference repository has no graph code; this file
 * IS the specification of the one thing section 9 left open - the fp32 accumulation order - so that "kNN neighbour
 * indices bit-exact in fp32 (ties broken by lowest index)" (BASELINE.json north_star) is a checkable statement:
 *
 *   G1  ss_i   = fma chain over d = 0..D-1 of p[i][d]*p[i][d]  (one IEEE fp32 fused multiply-add per feature, in order)
 *       n_i    = max(sqrtf(ss_i), 1e-12f)                       (F.normalize's eps, graph_oracle.py l2_normalize)
 *       ph[i][d] = p[i][d] / n_i                                (IEEE fp32 division, as F.normalize divides)
 *   G2  S[i][j] = fma chain over d = 0..D-1 of ph[i][d]*ph[j][d] (same order for every pair => S is exactly symmetric
 *                                                                 and duplicated token rows give bit-equal similarities)
 *   G3  idx[i][:] = the k columns of row i with the largest S, descending, ties -> lowest column (stable sort)
 *
 * Every operation is a single correctly rounded IEEE-754 binary32 operation, so any implementation that follows the
 * order (the CUDA kernel csrc/knn_simt.cu does: FFMA / IEEE division / IEEE sqrt) reproduces idx AND vals bit for bit.
 * Build: oracle/knn_strict.py (gcc -O2 -ffp-contract=off [-mfma]); fmaf() is exact with or without hardware FMA. */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

/* p: (B, Np, D) contiguous fp32.  idx: (B, Np, k) int32, vals: (B, Np, k) fp32, rnorm (may be NULL): (B, Np) = 1 / n_i.
 * Returns 0, or 1 on allocation failure / bad arguments. */
int knn_strict_f32(const float* p, int B, int Np, int D, int k, int32_t* idx, float* vals, float* rnorm) {
  if (B < 0 || Np < 1 || D < 1 || k < 1 || k > Np) return 1;
  float* ph = (float*)malloc(sizeof(float) * (size_t)Np * (size_t)D);
  float* srow = (float*)malloc(sizeof(float) * (size_t)Np);
  if (!ph || !srow) { free(ph); free(srow); return 1; }
  for (int b = 0; b < B; ++b) {
    const float* pb = p + (size_t)b * Np * D;
    for (int i = 0; i < Np; ++i) {                       /* G1 */
      float ss = 0.0f;
      for (int d = 0; d < D; ++d) ss = fmaf(pb[(size_t)i * D + d], pb[(size_t)i * D + d], ss);
      float n = sqrtf(ss);
      if (!(n > 1e-12f)) n = 1e-12f;
      if (rnorm) rnorm[(size_t)b * Np + i] = 1.0f / n;
      for (int d = 0; d < D; ++d) ph[(size_t)i * D + d] = pb[(size_t)i * D + d] / n;
    }
    for (int i = 0; i < Np; ++i) {
      const float* a = ph + (size_t)i * D;
      for (int j = 0; j < Np; ++j) {                     /* G2 */
        const float* c = ph + (size_t)j * D;
        float s = 0.0f;
        for (int d = 0; d < D; ++d) s = fmaf(a[d], c[d], s);
        srow[j] = s;
      }
      int32_t* oi = idx + ((size_t)b * Np + i) * k;      /* G3: sorted insertion, strict '>' keeps the earlier column on ties */
      float* ov = vals + ((size_t)b * Np + i) * k;
      int filled = 0;
      for (int j = 0; j < Np; ++j) {
        const float v = srow[j];
        int pos;
        if (filled < k) pos = filled++;
        else if (v > ov[k - 1]) pos = k - 1;
        else continue;
        while (pos > 0 && v > ov[pos - 1]) { ov[pos] = ov[pos - 1]; oi[pos] = oi[pos - 1]; --pos; }
        ov[pos] = v;
        oi[pos] = j;
      }
    }
  }
  free(ph);
  free(srow);
  return 0;
}
