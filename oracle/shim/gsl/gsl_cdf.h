/* TEST INFRASTRUCTURE ONLY (oracle/): minimal stand-in for <gsl/gsl_cdf.h>.
 *
 * The reference's src/setComparisons.c includes <gsl/gsl_cdf.h> (line 17) and
 * calls gsl_cdf_tdist_Q (line 117) and gsl_cdf_fdist_Q (line 213).  GSL is not
 * installed in this image, so this header lets that file compile UNMODIFIED
 * from /root/reference/src; the two functions are defined in
 * oracle/wt_special.c (our own regularised-incomplete-beta implementation).
 */
#ifndef WB200_ORACLE_GSL_CDF_SHIM_H
#define WB200_ORACLE_GSL_CDF_SHIM_H
double gsl_cdf_tdist_Q(double x, double nu);
double gsl_cdf_fdist_Q(double x, double nu1, double nu2);
#endif
