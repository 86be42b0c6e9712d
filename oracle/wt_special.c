/* TEST INFRASTRUCTURE ONLY (oracle/).  Nothing under wiggletools_b200/ may
 * include, link or call this file.
 *
 * Student-t and F upper-tail probabilities, standing in for the two GSL
 * functions the reference calls:
 *   gsl_cdf_tdist_Q  <- /root/reference/src/setComparisons.c:117
 *   gsl_cdf_fdist_Q  <- /root/reference/src/setComparisons.c:213
 * GSL (unpinned; Ubuntu 20.04 libgsl-dev ~2.5 per docker/Dockerfile:13) is not
 * vendored in the reference and not installed here, so the published
 * definitions are restated:
 *   Q_t(x; nu)      = 1/2 * I_{nu/(nu+x^2)}(nu/2, 1/2)            (x >= 0)
 *   Q_F(x; n1, n2)  = I_{n2/(n2+n1 x)}(n2/2, n1/2)
 * with I the regularised incomplete beta function, evaluated by its continued
 * fraction (modified Lentz).  PARITY UNPINNED against GSL itself: the
 * reference holds no test vector for these; tests pin this file against closed
 * forms (nu = 1, 2, infinity) instead.
 */
#include <math.h>

static double wt_betacf(double a, double b, double x) {
	const double tiny = 1e-300;
	double qab = a + b, qap = a + 1.0, qam = a - 1.0;
	double c = 1.0, d = 1.0 - qab * x / qap;
	if (fabs(d) < tiny) d = tiny;
	d = 1.0 / d;
	double h = d;
	for (int m = 1; m <= 10000; m++) {
		int m2 = 2 * m;
		double aa = m * (b - m) * x / ((qam + m2) * (a + m2));
		d = 1.0 + aa * d; if (fabs(d) < tiny) d = tiny;
		c = 1.0 + aa / c; if (fabs(c) < tiny) c = tiny;
		d = 1.0 / d;
		h *= d * c;
		aa = -(a + m) * (qab + m) * x / ((a + m2) * (qap + m2));
		d = 1.0 + aa * d; if (fabs(d) < tiny) d = tiny;
		c = 1.0 + aa / c; if (fabs(c) < tiny) c = tiny;
		d = 1.0 / d;
		double del = d * c;
		h *= del;
		if (fabs(del - 1.0) < 1e-16) break;
	}
	return h;
}


/* ln B(a,b) = lgamma(a)+lgamma(b)-lgamma(a+b).  With b = 1/2 and large a the
 * lgamma difference cancels catastrophically; use the asymptotic series of
 * ln[Gamma(a+1/2)/Gamma(a)] there. */
static double wt_lbeta(double a, double b) {
	if (b == 0.5 && a >= 20.0) {
		double r = 1.0 / a, r2 = r * r;
		double ratio = 0.5 * log(a) - r / 8 + r * r2 / 192 - r * r2 * r2 / 640 + 17 * r * r2 * r2 * r2 / 14336;
		return 0.5723649429247001 - ratio; /* ln sqrt(pi) */
	}
	if (a == 0.5 && b >= 20.0) {
		double r = 1.0 / b, r2 = r * r;
		double ratio = 0.5 * log(b) - r / 8 + r * r2 / 192 - r * r2 * r2 / 640 + 17 * r * r2 * r2 * r2 / 14336;
		return 0.5723649429247001 - ratio;
	}
	return lgamma(a) + lgamma(b) - lgamma(a + b);
}

/* I_x(a,b) given both x and xc = 1-x (passed separately so neither suffers
 * cancellation). */
static double wt_incbeta(double a, double b, double x, double xc) {
	if (x <= 0.0) return 0.0;
	if (xc <= 0.0) return 1.0;
	double lbt = -wt_lbeta(a, b) + a * (xc < 0.5 ? log1p(-xc) : log(x)) + b * (x < 0.5 ? log1p(-x) : log(xc));
	double bt = exp(lbt);
	if (x < (a + 1.0) / (a + b + 2.0))
		return bt * wt_betacf(a, b, x) / a;
	return 1.0 - bt * wt_betacf(b, a, xc) / b;
}

double gsl_cdf_tdist_Q(double x, double nu) {
	if (isnan(x) || isnan(nu)) return NAN;
	double x2 = x * x;
	double z = nu / (nu + x2), zc = x2 / (nu + x2);
	if (isinf(x)) { z = 0.0; zc = 1.0; }
	double tail = 0.5 * wt_incbeta(0.5 * nu, 0.5, z, zc);
	return x >= 0 ? tail : 1.0 - tail;
}

double gsl_cdf_fdist_Q(double x, double nu1, double nu2) {
	if (isnan(x) || isnan(nu1) || isnan(nu2)) return NAN;
	if (x <= 0) return 1.0;
	double den = nu2 + nu1 * x;
	return wt_incbeta(0.5 * nu2, 0.5 * nu1, nu2 / den, nu1 * x / den);
}
