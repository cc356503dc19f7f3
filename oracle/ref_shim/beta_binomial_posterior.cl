/* The reference's beta.cl and binomial.cl, verbatim by #include, plus the expansion of its posterior template
 * (distributions/posterior.cl is a format string: name, loglik, prior logpdf — models.clj:102-115).  The three-line
 * expansion below is OURS, written the way `format` fills the template. */
#include "uncomplicate/bayadera/internal/device/opencl/distributions/beta.cl"
#include "uncomplicate/bayadera/internal/device/opencl/distributions/binomial.cl"

inline REAL beta_binomial_mcmc_logpdf(const uint data_len, const uint hyperparams_len, const REAL* params,
                                      const uint dim, const REAL* x) {

    return binomial_loglik(data_len, params, dim, x) +
        beta_logpdf(data_len, hyperparams_len, &params[data_len], dim, x);
}
