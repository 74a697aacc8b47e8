/*
 * ORACLE (test infrastructure, not product code): scalar C restatement of
 * calculate_peak_parameters — reference core/utils/data_loader.py:13-58 — and of the sensitivity
 * S = (f/1.0)*(Q/100.0)*100 its callers derive (data_loader.py:96,105), in float64 like NumPy.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.
 * Pinned against the reference itself by tests/golden/physics_*.npz (tools/make_golden.py).
 */
#include <math.h>
#include <stdint.h>

/* data_loader.py:13-58, one spectrum */
void pigan_oracle_peak_parameters(const double* frequency, const double* t, int len, int peak_idx,
                                  double baseline, double* f_res_out, double* q_out, double* fom_out) {
  const double f_res = frequency[peak_idx];                    /* :14 */
  const double t_min = t[peak_idx];                            /* :15 */
  const double half = t_min + (baseline - t_min) / 2;          /* :16 */
  double f_lower = NAN, f_upper = NAN;                         /* :17 */
  for (int i = peak_idx - 1; i >= 0; --i) {                    /* :20 */
    if ((t[i] >= half && t[i + 1] < half) || (t[i] < half && t[i + 1] >= half)) { /* :21-22 */
      if ((t[i + 1] - t[i]) != 0)                              /* :24 */
        f_lower = frequency[i] + (half - t[i]) * (frequency[i + 1] - frequency[i]) / (t[i + 1] - t[i]);
      else
        f_lower = frequency[i];                                /* :28 */
      break;
    }
  }
  for (int i = peak_idx + 1; i < len - 1; ++i) {               /* :32 */
    if ((t[i] <= half && t[i + 1] > half) || (t[i] > half && t[i + 1] <= half)) { /* :33-34 */
      if ((t[i + 1] - t[i]) != 0)
        f_upper = frequency[i] + (half - t[i]) * (frequency[i + 1] - frequency[i]) / (t[i + 1] - t[i]);
      else
        f_upper = frequency[i];
      break;
    }
  }
  double Q = NAN, FoM = NAN;                                   /* :43-44 */
  if (!isnan(f_lower) && !isnan(f_upper) && f_upper > f_lower) { /* :47 */
    const double delta_f = f_upper - f_lower;
    if (delta_f > 1e-9) Q = f_res / delta_f;                   /* :49-50 */
    if (!isnan(t_min) && fabs(t_min) > 1e-6)                   /* :53 */
      FoM = isnan(Q) ? NAN : Q / fabs(t_min);
    else
      FoM = NAN;
  }
  *f_res_out = f_res;
  *q_out = Q;
  *fom_out = FoM;
}

/* np.argmin semantics: first occurrence of the minimum; a NaN wins and the first NaN is returned */
static int argmin_np(const float* t, int len) {
  int best = 0;
  for (int i = 1; i < len; ++i) {
    if (isnan(t[best])) break;
    if (isnan(t[i]) || t[i] < t[best]) best = i;
  }
  return best;
}

/* Batched driver over fp32 spectra [n, s] (upcast to double per row, as NumPy does when handed a
 * float32 row together with a float64 frequency grid).  peak_idx may be NULL (argmin per row).
 * out_metrics [n,4] = f_res, Q, FoM, S in double. */
void pigan_oracle_physics_batch(const float* spectra, int64_t n, int s, const double* frequency,
                                const int32_t* peak_idx, double baseline, int32_t* out_idx,
                                double* out_metrics) {
  double buf[4096];
  for (int64_t r = 0; r < n; ++r) {
    const float* row = spectra + r * (int64_t)s;
    const int idx = peak_idx ? peak_idx[r] : argmin_np(row, s);
    for (int i = 0; i < s; ++i) buf[i] = (double)row[i];
    double f, q, fom, S = NAN;
    if (idx < 0 || idx >= s) {
      f = q = fom = NAN;
    } else {
      pigan_oracle_peak_parameters(frequency, buf, s, idx, baseline, &f, &q, &fom);
      if (!isnan(q)) S = (f / 1.0) * (q / 100.0) * 100; /* data_loader.py:96 */
    }
    if (out_idx) out_idx[r] = idx;
    out_metrics[4 * r + 0] = f;
    out_metrics[4 * r + 1] = q;
    out_metrics[4 * r + 2] = fom;
    out_metrics[4 * r + 3] = S;
  }
}
