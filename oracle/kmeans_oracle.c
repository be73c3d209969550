/* TEST INFRASTRUCTURE ONLY -- CPU oracle of the embedding-clustering step.
 *
 * The reference clusters foreground embeddings with
 *     sklearn.cluster.KMeans(n_clusters=n, n_init=35, max_iter=500).fit_predict(X)
 * (/root/reference/code/lib/prediction.py:72-74).  scikit-learn is a third-party dependency
 * that is NOT vendored under /root/reference and the reference pins no version; the image
 * has scikit-learn 1.9.0, whose published algorithm this file restates:
 *   - KMeans.fit: mean-centre X, tol = mean(var(X, axis=0)) * 1e-4, n_init restarts drawing
 *     from ONE RandomState stream, keep the restart with the lowest inertia unless it is the
 *     same clustering            (sklearn/cluster/_kmeans.py:1463-1563, _tolerance :285-292)
 *   - greedy k-means++ seeding with 2+floor(ln k) local trials: first centre by
 *     RandomState.choice, then D^2 sampling by searchsorted on the cumulative potential and
 *     "best candidate = lowest potential"                     (_kmeans.py:180-283)
 *   - Lloyd: argmin_j |c_j|^2 - 2 x.c_j with first-index tie break, centre = mean of members,
 *     empty clusters relocated to the farthest points, stop when labels repeat (strict) else
 *     when sum_j |c_j' - c_j|^2 <= tol, re-run the E-step if not strict, inertia = sum of
 *     squared distances            (_kmeans.py:625-760, _k_means_lloyd.pyx:26-211,
 *                                   _k_means_common.pyx:13-43,167-262)
 *
 * What is pinned and what is not.  scikit-learn's own floating-point sums over points (BLAS sgemm / sgemv / sdot,
 * per-thread partial centre sums) depend on the BLAS kernel of the host CPU and on the OpenMP thread count, so its
 * results are not bit-reproducible across machines.  This oracle fixes ONE arithmetic that a GPU can reproduce bit
 * for bit and that takes scikit-learn's own decisions wherever those are machine independent:
 *   - k-means++ (_kmeans.py:180-283) exactly as scikit-learn computes it for float32 inputs: candidate distances
 *     through the float64 upcast of _euclidean_distances (d = -2 x.y + |x|^2 + |y|^2 in double, cast to fp32,
 *     clamped at 0: sklearn/metrics/pairwise.py _euclidean_distances_upcast), candidates drawn by searchsorted on the
 *     SEQUENTIAL FLOAT32 cumulative sum of closest_dist_sq (numpy's cumsum of a float32 array; its rounding drift,
 *     ~1e-5 of the total, is as large as one point's share, so an exact prefix picks different points), thresholds
 *     u * float32(potential);
 *   - Lloyd per-pair arithmetic in fp32 with a fixed left-to-right fmaf order (the k loop of an sgemm micro-kernel);
 *   - every sum over points whose order scikit-learn leaves to BLAS / OpenMP (column means, potentials, centre
 *     sums, inertia) in exact 64-bit fixed point (order independent: the correctly rounded value of the sum
 *     scikit-learn approximates), scales derived from max|X|, n and C only.
 * Measured on random-init network embeddings (62 k points, k = 16, tools/km_parity_probe.py): 30-34 of 35 restarts
 * draw exactly scikit-learn's seeds (the rest differ by one neighbouring point where BLAS rounding of the potential
 * decides), and Lloyd from identical seeds stops at scikit-learn's iteration in 7 of 8 restarts.
 * The CUDA path must match this oracle EXACTLY (labels, seeds, n_iter, inertia, centres).  The oracle itself is
 * pinned against scikit-learn on golden vectors (tests/golden/kmeans.npz).
 *
 * Build: make -C oracle   (gcc -O2 -ffp-contract=off -mfma: fmaf() is a single rounding).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  int e_m, bits_n, S_mean, S_x, S_d, S_t;
} isa_km_scales;

static int ceil_log2_int(int v) {
  int b = 0;
  while ((1 << b) < v) ++b;
  return b;
}

/* All fixed-point scales derive from max|X| (raw), n and C. */
void isa_km_oracle_scales(float maxabs, int n, int C, isa_km_scales* s) {
  s->e_m = (maxabs > 0.f) ? (ilogbf(maxabs) + 1) : 0;
  int bits = 0;
  while (bits < 31 && (n >> bits) != 0) ++bits;
  s->bits_n = bits;
  const int e_c = s->e_m + 1;
  s->S_mean = 62 - s->e_m - bits;
  s->S_x = 62 - e_c - bits;
  s->S_d = 62 - (2 * e_c + 2 + ceil_log2_int(C)) - bits;
  s->S_t = 62 - 2 * e_c - bits - ceil_log2_int(C);
}

static inline int64_t to_fixed(float v, int S) { return llrint(ldexp((double)v, S)); }

/* squared distance, direct form, fixed fmaf order */
static inline float sqdist(const float* a, const float* b, int C) {
  float acc = 0.f;
  for (int f = 0; f < C; ++f) {
    const float t = a[f] - b[f];
    acc = fmaf(t, t, acc);
  }
  return acc;
}

/* argmin_j fmaf(-2, x.c_j, |c_j|^2), strict <, first index wins */
static inline int nearest(const float* x, const float* centers, const float* csq, int k, int C) {
  int best = 0;
  float bv = 0.f;
  for (int j = 0; j < k; ++j) {
    const float* c = centers + (size_t)j * C;
    float dot = 0.f;
    for (int f = 0; f < C; ++f) dot = fmaf(x[f], c[f], dot);
    const float v = fmaf(-2.f, dot, csq[j]);
    if (j == 0 || v < bv) { bv = v; best = j; }
  }
  return best;
}

/* RandomState.choice(n, p=uniform): cdf[i] = fl64((i+1)/n); searchsorted(u, side='right') */
static int first_center_index(double u, int n) {
  long i = (long)(u * (double)n);
  if (i > n - 1) i = n - 1;
  if (i < 0) i = 0;
  while (i > 0 && (double)i / (double)n > u) --i;
  while (i < n - 1 && (double)(i + 1) / (double)n <= u) ++i;
  return (int)i;
}

/* is l2 a function of l1 (sklearn _is_same_clustering) */
static int same_clustering(const int* l1, const int* l2, int n, int k) {
  int map[256];
  for (int j = 0; j < k; ++j) map[j] = -1;
  for (int i = 0; i < n; ++i) {
    if (map[l1[i]] == -1) map[l1[i]] = l2[i];
    else if (map[l1[i]] != l2[i]) return 0;
  }
  return 1;
}

/* Centres X in place (X: n x C row-major); returns mean[C], tol_abs, scales. status 2 = non-finite input. */
int isa_km_oracle_prepare(float* X, int n, int C, double tol_rel, float* mean, float* tol_abs, isa_km_scales* sc) {
  float maxabs = 0.f;
  for (size_t i = 0; i < (size_t)n * C; ++i) {
    const float a = fabsf(X[i]);
    if (!(a <= 3.0e38f)) return 2;
    if (a > maxabs) maxabs = a;
  }
  isa_km_oracle_scales(maxabs, n, C, sc);
  for (int f = 0; f < C; ++f) {
    int64_t s = 0;
    for (int i = 0; i < n; ++i) s += to_fixed(X[(size_t)i * C + f], sc->S_mean);
    mean[f] = (float)(ldexp((double)s, -sc->S_mean) / (double)n);
  }
  int64_t t = 0;
  for (int i = 0; i < n; ++i)
    for (int f = 0; f < C; ++f) {
      const float v = X[(size_t)i * C + f] - mean[f];
      X[(size_t)i * C + f] = v;
      t += to_fixed(v * v, sc->S_t);
    }
  *tol_abs = (float)(ldexp((double)t, -sc->S_t) / ((double)n * (double)C) * tol_rel);
  return 0;
}

/* |x|^2 of a float32 row in double: row_norms(X_chunk.astype(float64), squared=True) */
static inline double sqnorm64(const float* a, int C) {
  double s = 0.0;
  for (int f = 0; f < C; ++f) s = fma((double)a[f], (double)a[f], s);
  return s;
}

/* squared distance as sklearn's _euclidean_distances computes it for float32 inputs (pairwise.py,
 * _euclidean_distances_upcast): d = -2 * (x . y); d += |x|^2; d += |y|^2 in float64, cast to float32, max(., 0).
 * x = the candidate row (norm cc), y = the point (norm yy). */
static inline float sqdist_upcast(const float* x, double cc, const float* y, double yy, int C) {
  double dot = 0.0;
  for (int f = 0; f < C; ++f) dot = fma((double)x[f], (double)y[f], dot);
  double d = -2.0 * dot;
  d += cc;
  d += yy;
  const float r = (float)d;
  return r > 0.f ? r : 0.f;
}

static inline float pot_to_float(int64_t q, int S) { return (float)ldexp((double)q, -S); }

/* greedy k-means++ on centred X (_kmeans.py:180-283); u = 1 + (k-1)*L uniforms of this restart; xx[n] = sqnorm64 of
 * every row.  idx_out[k]. */
void isa_km_oracle_seed(const float* X, int n, int C, int k, int L, const double* u, const isa_km_scales* sc,
                        int* idx_out, float* closest /* n scratch */, const double* xx) {
  int c0 = first_center_index(u[0], n);
  idx_out[0] = c0;
  int64_t potq = 0;
  for (int i = 0; i < n; ++i) {
    closest[i] = sqdist_upcast(X + (size_t)c0 * C, xx[c0], X + (size_t)i * C, xx[i], C);
    potq += to_fixed(closest[i], sc->S_d);
  }
  float pot = pot_to_float(potq, sc->S_d);           /* current_pot: float32 (closest_dist_sq @ sample_weight) */
  for (int c = 1; c < k; ++c) {
    /* rand_vals = uniform(size=L) * current_pot (float64); candidate_ids = searchsorted(cumsum_f32, rand_vals),
     * clipped to n-1.  np.cumsum of a float32 array adds sequentially in float32. */
    double v[16];
    int cand[16], found[16], pending = L;
    for (int t = 0; t < L; ++t) { v[t] = u[1 + (c - 1) * L + t] * (double)pot; cand[t] = n - 1; found[t] = 0; }
    float S = 0.f;
    for (int i = 0; i < n && pending; ++i) {
      S = S + closest[i];
      for (int t = 0; t < L; ++t)
        if (!found[t] && (double)S >= v[t]) { cand[t] = i; found[t] = 1; --pending; }
    }
    int best_cand = -1;
    float best_pot = 0.f;
    for (int t = 0; t < L; ++t) {
      int64_t p = 0;
      const float* xc = X + (size_t)cand[t] * C;
      for (int i = 0; i < n; ++i)
        p += to_fixed(fminf(closest[i], sqdist_upcast(xc, xx[cand[t]], X + (size_t)i * C, xx[i], C)), sc->S_d);
      const float pf = pot_to_float(p, sc->S_d);     /* candidates_pot is float32; np.argmin: first lowest */
      if (t == 0 || pf < best_pot) { best_pot = pf; best_cand = cand[t]; }
    }
    const float* xb = X + (size_t)best_cand * C;
    for (int i = 0; i < n; ++i)
      closest[i] = fminf(closest[i], sqdist_upcast(xb, xx[best_cand], X + (size_t)i * C, xx[i], C));
    pot = best_pot;
    idx_out[c] = best_cand;
  }
}

/* One Lloyd run from `centers` (k x C, overwritten with the final centres).
 * labels[n] out, returns n_iter; *inertia_q = fixed-point inertia (scale S_d); *strict_out. */
int isa_km_oracle_lloyd(const float* X, int n, int C, int k, int max_iter, float tol_abs, const isa_km_scales* sc,
                        float* centers, int* labels, int64_t* inertia_q, int* strict_out, int* n_reloc_out) {
  int* acct = (int*)malloc(sizeof(int) * n);
  int64_t* sums = (int64_t*)calloc((size_t)k * C, sizeof(int64_t));
  int* cnt = (int*)calloc(k, sizeof(int));
  float* csq = (float*)malloc(sizeof(float) * k);
  float* cnew = (float*)malloc(sizeof(float) * k * C);
  float* dist = NULL;
  for (int i = 0; i < n; ++i) { labels[i] = -1; acct[i] = -1; }
  int strict = 0, n_iter = max_iter, n_reloc = 0;
  for (int it = 0; it < max_iter; ++it) {
    for (int j = 0; j < k; ++j) {
      float a = 0.f;
      for (int f = 0; f < C; ++f) a = fmaf(centers[j * C + f], centers[j * C + f], a);
      csq[j] = a;
    }
    int changed = 0;
    for (int i = 0; i < n; ++i) {
      const float* x = X + (size_t)i * C;
      const int l = nearest(x, centers, csq, k, C);
      if (l != acct[i]) {
        for (int f = 0; f < C; ++f) {
          const int64_t q = to_fixed(x[f], sc->S_x);
          if (acct[i] >= 0) sums[(size_t)acct[i] * C + f] -= q;
          sums[(size_t)l * C + f] += q;
        }
        if (acct[i] >= 0) cnt[acct[i]]--;
        cnt[l]++;
        acct[i] = l;
      }
      if (l != labels[i]) changed = 1;
      labels[i] = l;
    }
    /* empty clusters -> farthest points (_k_means_common.pyx:167-212); the empty list is fixed
     * before any point moves, as in scikit-learn */
    int n_empty = 0;
    int empties[256];
    for (int j = 0; j < k; ++j)
      if (cnt[j] == 0) empties[n_empty++] = j;
    if (n_empty > 0) {
      if (!dist) dist = (float*)malloc(sizeof(float) * n);
      float dmax = 0.f;
      for (int i = 0; i < n; ++i) {
        dist[i] = sqdist(X + (size_t)i * C, centers + (size_t)labels[i] * C, C);
        if (dist[i] > dmax) dmax = dist[i];
      }
      if (dmax > 0.f) {
        for (int e = 0; e < n_empty; ++e) {
          const int j = empties[e];
          /* next farthest point: largest distance, lowest index on ties; taken points are marked -1 */
          int far = -1;
          float fd = -1.f;
          for (int i = 0; i < n; ++i)
            if (dist[i] > fd) { fd = dist[i]; far = i; }
          if (far < 0) break;
          dist[far] = -1.f;
          const int old = acct[far];
          const float* x = X + (size_t)far * C;
          for (int f = 0; f < C; ++f) {
            const int64_t q = to_fixed(x[f], sc->S_x);
            sums[(size_t)old * C + f] -= q;
            sums[(size_t)j * C + f] = q;
          }
          cnt[old]--;
          cnt[j] = 1;
          acct[far] = j;
          ++n_reloc;
        }
      }
    }
    int amax = 0;
    for (int j = 1; j < k; ++j)
      if (cnt[j] > cnt[amax]) amax = j;
    for (int j = 0; j < k; ++j)
      if (cnt[j] > 0)
        for (int f = 0; f < C; ++f)
          cnew[j * C + f] = (float)(ldexp((double)sums[(size_t)j * C + f], -sc->S_x) / (double)cnt[j]);
    for (int j = 0; j < k; ++j)
      if (cnt[j] <= 0)
        for (int f = 0; f < C; ++f) cnew[j * C + f] = cnew[amax * C + f];
    float shift_tot = 0.f;
    for (int j = 0; j < k; ++j) {
      const float s = sqrtf(sqdist(cnew + (size_t)j * C, centers + (size_t)j * C, C));
      shift_tot = fmaf(s, s, shift_tot);
    }
    memcpy(centers, cnew, sizeof(float) * k * C);
    if (!changed) { strict = 1; n_iter = it + 1; break; }
    if (shift_tot <= tol_abs) { n_iter = it + 1; break; }
  }
  if (!strict) {
    for (int j = 0; j < k; ++j) {
      float a = 0.f;
      for (int f = 0; f < C; ++f) a = fmaf(centers[j * C + f], centers[j * C + f], a);
      csq[j] = a;
    }
    for (int i = 0; i < n; ++i) labels[i] = nearest(X + (size_t)i * C, centers, csq, k, C);
  }
  int64_t in = 0;
  for (int i = 0; i < n; ++i) in += to_fixed(sqdist(X + (size_t)i * C, centers + (size_t)labels[i] * C, C), sc->S_d);
  *inertia_q = in;
  *strict_out = strict;
  if (n_reloc_out) *n_reloc_out = n_reloc;
  free(acct); free(sums); free(cnt); free(csq); free(cnew); free(dist);
  return n_iter;
}

/* Full fit.  X_in: n x C row-major fp32 (not modified).  uniforms: n_init * (1+(k-1)*L) doubles drawn from
 * numpy RandomState(seed).random_sample (the stream KMeans consumes).  init_centers (optional,
 * n_init x k x C, NOT centred): use these seeds instead of k-means++ (mean is subtracted like
 * KMeans.fit does for an array init).
 * Outputs: labels[n]; centers[k*C] (mean added back); per-restart inertia (double), n_iter, strict flags;
 * seed_idx[n_init*k] (k-means++ picks, -1 when init_centers given).  Returns status: 0 ok, 1 n<k, 2 non-finite. */
int isa_km_oracle_fit(const float* X_in, int n, int C, int k, int n_init, int max_iter, double tol_rel,
                      const double* uniforms, const float* init_centers,
                      int* labels_out, float* centers_out, double* inertia_out, int* n_iter_out, int* strict_out,
                      int* seed_idx_out, int* best_out, float* tol_abs_out) {
  if (n < k || n <= 0 || k <= 0 || k > 255) return 1;
  float* X = (float*)malloc(sizeof(float) * (size_t)n * C);
  memcpy(X, X_in, sizeof(float) * (size_t)n * C);
  float* mean = (float*)malloc(sizeof(float) * C);
  float tol_abs;
  isa_km_scales sc;
  int st = isa_km_oracle_prepare(X, n, C, tol_rel, mean, &tol_abs, &sc);
  if (st) { free(X); free(mean); return st; }
  if (tol_abs_out) *tol_abs_out = tol_abs;
  const int L = 2 + (int)log((double)k);
  const int per = 1 + (k - 1) * L;
  double* xx = (double*)malloc(sizeof(double) * n);
  for (int i = 0; i < n; ++i) xx[i] = sqnorm64(X + (size_t)i * C, C);
  int* all_labels = (int*)malloc(sizeof(int) * (size_t)n * n_init);
  float* all_centers = (float*)malloc(sizeof(float) * (size_t)k * C * n_init);
  int64_t* inq = (int64_t*)malloc(sizeof(int64_t) * n_init);
#pragma omp parallel for schedule(dynamic, 1)
  for (int r = 0; r < n_init; ++r) {
    float* centers = all_centers + (size_t)r * k * C;
    int* idx = seed_idx_out + (size_t)r * k;
    if (init_centers) {
      for (int j = 0; j < k; ++j) {
        idx[j] = -1;
        for (int f = 0; f < C; ++f) centers[j * C + f] = init_centers[((size_t)r * k + j) * C + f] - mean[f];
      }
    } else {
      float* closest = (float*)malloc(sizeof(float) * n);
      isa_km_oracle_seed(X, n, C, k, L, uniforms + (size_t)r * per, &sc, idx, closest, xx);
      free(closest);
      for (int j = 0; j < k; ++j) memcpy(centers + (size_t)j * C, X + (size_t)idx[j] * C, sizeof(float) * C);
    }
    n_iter_out[r] = isa_km_oracle_lloyd(X, n, C, k, max_iter, tol_abs, &sc, centers, all_labels + (size_t)r * n,
                                        &inq[r], &strict_out[r], NULL);
    inertia_out[r] = ldexp((double)inq[r], -sc.S_d);
  }
  int best = 0;
  for (int r = 1; r < n_init; ++r)
    if (inq[r] < inq[best] && !same_clustering(all_labels + (size_t)r * n, all_labels + (size_t)best * n, n, k)) best = r;
  memcpy(labels_out, all_labels + (size_t)best * n, sizeof(int) * n);
  for (int j = 0; j < k; ++j)
    for (int f = 0; f < C; ++f) centers_out[j * C + f] = all_centers[((size_t)best * k + j) * C + f] + mean[f];
  *best_out = best;
  free(X); free(mean); free(all_labels); free(all_centers); free(inq); free(xx);
  return 0;
}
