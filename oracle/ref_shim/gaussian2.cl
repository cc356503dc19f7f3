/* OURS, not the reference's: a 2-D isotropic Gaussian in the reference's model dialect, used only to run the
 * reference's stretch kernel text at DIM = 2 (its own distribution files are all one-dimensional). */
inline REAL gaussian2_logpdf(const uint data_len, const uint params_len, const REAL* params,
                             const uint dim, const REAL* x) {
    const REAL a = x[0] - params[0];
    const REAL b = x[1] - params[1];
    return (a * a + b * b) / (-2.0f * params[2] * params[2]);
}
