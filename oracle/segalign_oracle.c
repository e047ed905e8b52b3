/* segalign_oracle.c -- CPU restatement of PyPore's cSegmentAligner (TEST INFRASTRUCTURE ONLY: nothing under
 * protpore_b200/ links or loads this file).
 *
 * Follows /root/reference/PyPore/calignment.pyx operation for operation:
 *   20-33   constructor: c_model_dur = numpy.cumsum(model_dur) (sequential float64 sums)
 *   38-50   match[i, j] = -(seq_mean_i - model_mean_j)^2 / (seq_std_i * model_std_j);  first row of `score`
 *   52-68   per row: the skip chain (ascending j), the back-slip chain (descending j), then the row of `score`
 *   70      j = argmax of the last row (double_argmax, 11-18)
 *   72-97   backtrace with the 1e-6 relative tolerance tests, in the reference's order: diagonal, stay, back-slip, skip
 *   99-100  score[s-1, m-1] / numpy.sum(seq_durs)   (numpy's pairwise summation, restated below)
 *
 * Where the reference's behaviour is undefined this file DEFINES it and reports it in *status:
 *   bit 0  no entry of the last row exceeds -1: double_argmax (11-18) returns an uninitialised int.  Defined here: the
 *          first index attaining the row's maximum (what the routine returns whenever the maximum is above -1).
 *   bit 1  the backtrace stands at model position 0 in a row i >= 1: `score[i-1, j-1]` with the unsigned j = 0 (37, 76)
 *          is an out-of-bounds read and the reference raises IndexError (bounds checks are on).  Defined here: the
 *          diagonal test is skipped at j = 0, and a skip that would leave the model on the left stops at position 0.
 * The parity tests compare with the compiled reference only where status == 0.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#define NEGINF (-99999999.0)

/* numpy.add.reduce over a contiguous float64 array (numpy/core/src/umath/loops_utils.h, pairwise_sum) */
static double pairwise_sum(const double *a, int64_t n) {
    if (n < 8) {
        double res = 0.;
        for (int64_t i = 0; i < n; i++) res += a[i];
        return res;
    } else if (n <= 128) {
        double r[8];
        int64_t i;
        for (i = 0; i < 8; i++) r[i] = a[i];
        for (i = 8; i < n - (n % 8); i += 8)
            for (int k = 0; k < 8; k++) r[k] += a[i + k];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; i++) res += a[i];
        return res;
    } else {
        int64_t n2 = n / 2;
        n2 -= n2 % 8;
        return pairwise_sum(a, n2) + pairwise_sum(a + n2, n - n2);
    }
}
double segalign_pairwise_sum(const double *a, int64_t n) { return pairwise_sum(a, n); }

static double dmax(double a, double b) { return a >= b ? a : b; }     /* double_max, 8 */

/* Returns 0, or -1 if out of memory / m < 2 (the reference indexes model position 1 unconditionally, 60). */
int segalign_align(const double *model_means, const double *model_stds, const double *model_dur, int64_t m, double skip_penalty,
                   double backslip_penalty, const double *seq_means, const double *seq_stds, const double *seq_durs, int64_t s,
                   double *score_out, double *order_out, int32_t *status) {
    if (m < 2 || s < 1) return -1;
    double *match = malloc(sizeof(double) * s * m), *score = malloc(sizeof(double) * s * m);
    double *skip = malloc(sizeof(double) * s * m), *back = malloc(sizeof(double) * s * m), *cdur = malloc(sizeof(double) * m);
    if (!match || !score || !skip || !back || !cdur) return -1;
    *status = 0;
    double acc = 0.0;
    for (int64_t j = 0; j < m; j++) { acc += model_dur[j]; cdur[j] = acc; }                     /* 30 */
#define M(i, j) match[(i) * m + (j)]
#define S(i, j) score[(i) * m + (j)]
#define K(i, j) skip[(i) * m + (j)]
#define B(i, j) back[(i) * m + (j)]
    for (int64_t i = 0; i < s; i++)
        for (int64_t j = 0; j < m; j++) {
            const double d = seq_means[i] - model_means[j];
            M(i, j) = -(d * d) / (seq_stds[i] * model_stds[j]);                                    /* 47 */
        }
    for (int64_t j = 0; j < m; j++) S(0, j) = M(0, j) * seq_durs[0] - skip_penalty * (cdur[j] - model_dur[j]);   /* 50 */
    for (int64_t i = 1; i < s; i++) {
        K(i, 0) = NEGINF;
        for (int64_t j = 1; j < m; j++) K(i, j) = dmax(K(i, j - 1), S(i - 1, j - 1)) - model_dur[j] * skip_penalty;       /* 55 */
        B(i, m - 1) = NEGINF;
        for (int64_t j = m - 2; j > 0; j--) B(i, j) = dmax(B(i, j + 1), S(i - 1, j + 1)) - model_dur[j + 1] * backslip_penalty;   /* 58 */
        B(i, 0) = dmax(B(i, 1), S(i - 1, 1)) - model_dur[1] * backslip_penalty;                  /* 59 */
        for (int64_t j = 0; j < m; j++) {
            double prev = S(i - 1, j);
            if (j > 0) {                                                                           /* max(prev, diag, skip): 64 */
                if (S(i - 1, j - 1) > prev) prev = S(i - 1, j - 1);
                if (K(i, j - 1) > prev) prev = K(i, j - 1);
            }
            if (j < m - 1) prev = dmax(prev, B(i, j));
            S(i, j) = prev + M(i, j) * seq_durs[i];                                                /* 67 */
        }
    }
    int64_t j = -1;
    {   /* double_argmax(score[s-1]), 11-18 */
        double maximum = -1;
        for (int64_t t = 0; t < m; t++)
            if (S(s - 1, t) > maximum) { j = t; maximum = S(s - 1, t); }
        if (j < 0) {
            *status |= 1;
            j = 0;
            for (int64_t t = 1; t < m; t++) if (S(s - 1, t) > S(s - 1, j)) j = t;
        }
    }
    for (int64_t i = s - 1; i > 0; i--) {
        order_out[i] = (double)j;
        const double prev_test = S(i, j) - M(i, j) * seq_durs[i];
        const double max_err = 1e-6 * fabs(prev_test);
        if (j == 0) *status |= 2;
        if (j > 0 && fabs(prev_test - S(i - 1, j - 1)) <= max_err) { j -= 1; continue; }           /* 76-78 */
        if (fabs(prev_test - S(i - 1, j)) <= max_err) continue;                                   /* 79-80 */
        if (j < m - 1) {                                                                           /* 81-88 */
            int64_t k = j;
            double back_test = prev_test;
            while (k < m - 1 && fabs(back_test - B(i, k)) <= max_err) {
                k += 1;
                back_test += model_dur[k] * backslip_penalty;
            }
            if (k > j) { j = k; continue; }
        }
        if (j > 0) {                                                                               /* 89-97 */
            int64_t k = j;
            double skip_test = prev_test;
            while (k >= 1 && fabs(skip_test - K(i, k - 1)) <= max_err) {
                k -= 1;
                skip_test += model_dur[k] * skip_penalty;
            }
            if (k < j) {
                if (k == 0) { *status |= 2; j = 0; } else j = k - 1;
                continue;
            }
        }
    }
    order_out[0] = (double)j;
    *score_out = S(s - 1, m - 1) / pairwise_sum(seq_durs, s);                                      /* 100 */
    free(match); free(score); free(skip); free(back); free(cdur);
    return 0;
}
