/* Plain C restatement of the third-party arithmetic the reference steps with -- TEST INFRASTRUCTURE,
 * NOT PRODUCT CODE (see oracle/restate.py; nothing under pyfds_b200/ loads this).
 *
 * scipy.sparse.dia_matrix.dot(x) (scipy 1.18.1, _dia.py:285-298 -> sparsetools dia.h, dia_matvec):
 *   y starts at zero; for every stored diagonal, in stored order, with offset k:
 *     rows i = max(0,-k) .., columns j = max(0,k) .. min(n+k, n)-1:  y[i] += diag[j] * x[j]
 * The data of diagonal d are indexed by COLUMN (data[d*n + j]), every term is one rounded multiply
 * followed by one rounded add. Build with -ffp-contract=off: a fused multiply-add would change bits.
 *
 * Exists as a third, independent implementation of the oracle's inner loop (NumPy slices, scipy's own
 * C++, this file); tests/test_oracle.py requires all three to agree bit for bit on the goldens.
 */
#include <stdint.h>

void fds_oracle_dia_matvec(int64_t n, int64_t n_diags, const int64_t *offsets, const double *data,
                           const double *x, double *y) {
    for (int64_t i = 0; i < n; ++i) y[i] = 0.0;
    for (int64_t d = 0; d < n_diags; ++d) {
        const int64_t k = offsets[d];
        const int64_t i_start = k < 0 ? -k : 0;
        const int64_t j_start = k > 0 ? k : 0;
        int64_t j_end = n + k < n ? n + k : n;
        const double *diag = data + d * n;
        for (int64_t j = j_start, i = i_start; j < j_end; ++j, ++i) {
            const double product = diag[j] * x[j];
            y[i] = y[i] + product;
        }
    }
}
