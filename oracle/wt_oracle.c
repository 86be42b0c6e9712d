/* TEST INFRASTRUCTURE ONLY (oracle/).  Nothing under wiggletools_b200/ may
 * include, link or call this file; only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs use it, as the checker.
 *
 * CPU restatement, in batch form, of the WiggleTools hot path
 *   N sorted (chrom,start,end,value) tracks -> union-breakpoint segments ->
 *   per-segment N-wide statistic (+ AUC / Pearson scalars)
 * following /root/reference/src/{multiplexer,multiSet,reducers,setComparisons,
 * statistics}.c.  It is NOT the reference's lazy heap-driven iterator graph: the
 * sweep below keeps one cursor per track and takes O(N) per segment.
 *
 * PINNING: validated (tests/test_oracle.py) against
 *   - the reference's own golden vectors: test/expected/pearson.txt
 *     (-0.028968), the AUC answers of test/test.py:95-98, the `sum x x == scale
 *     2 x` identity of test/test.py:76;
 *   - outputs of the UNMODIFIED reference compiled here (oracle/_ref/libwtref.so,
 *     recipe in oracle/Makefile) on seeded random inputs, bit for bit, and the
 *     fixtures generated from it in tests/golden/ (tests/golden/make_golden.py).
 *   - t-test p-values go through oracle/wt_special.c, whose GSL counterpart is
 *     absent: that one function is "parity unpinned" vs GSL and pinned on
 *     closed forms only.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <stdio.h>

double gsl_cdf_tdist_Q(double x, double nu); /* oracle/wt_special.c */

typedef struct {
	const int32_t *chrom;
	const int32_t *start;
	const int32_t *end;
	const double *value;
	int64_t n;
	double default_value;
} wt_track;

typedef struct {
	int32_t *chrom, *start, *end;
	double *value;
	int64_t cap;
} wt_out;

enum {
	WT_SUM = 0, WT_PRODUCT = 1, WT_MEAN = 2, WT_VAR = 3, WT_STDDEV = 4,
	WT_ENTROPY = 5, WT_CV = 6, WT_MIN = 7, WT_MAX = 8, WT_MEDIAN = 9,
	WT_FILLIN = 10, WT_TRIMFILL = 11,
	WT_TTEST = 16, WT_MWU = 17
};

/* ------------------------------------------------------------------------ */
/* The sweep (src/multiplexer.c:37-128 restated).  One call = one segment.  */
/* ------------------------------------------------------------------------ */
typedef struct {
	int n;
	const wt_track *t;
	char **names;
	int64_t *cur;      /* per-track cursor */
	char *inplay;
	double *x;         /* value if in play else default (what reducers read) */
	double *v;         /* raw multiplexer values[] */
	int inplay_count;
	int chrom;         /* index into names, -1 before first */
	int32_t start, finish;
	int done;
} Sweep;

static void sweep_init(Sweep *s, int n, const wt_track *t, char **names) {
	s->n = n; s->t = t; s->names = names;
	s->cur = (int64_t *) calloc(n, sizeof(int64_t));
	s->inplay = (char *) calloc(n, 1);
	s->x = (double *) calloc(n, sizeof(double));
	s->v = (double *) calloc(n, sizeof(double));
	for (int i = 0; i < n; i++) s->x[i] = s->v[i] = t[i].default_value;
	s->inplay_count = 0; s->chrom = -1; s->start = s->finish = 0; s->done = 0;
}

static void sweep_free(Sweep *s) { free(s->cur); free(s->inplay); free(s->x); free(s->v); }

static int on_chrom(const Sweep *s, int i) {
	return s->cur[i] < s->t[i].n && s->t[i].chrom[s->cur[i]] == s->chrom;
}

/* Advance to the next union segment; returns 1 if all tracks are in play. */
static int sweep_step(Sweep *s) {
	int i;
	/* tracks whose interval ends at the current finish leave play (:37-48) */
	for (i = 0; i < s->n; i++)
		if (s->inplay[i] && s->t[i].end[s->cur[i]] == s->finish) {
			s->inplay[i] = 0; s->inplay_count--;
			s->cur[i]++;
			s->x[i] = s->v[i] = s->t[i].default_value;
		}
	/* anything left on this chromosome? */
	int waiting = 0;
	if (s->chrom >= 0)
		for (i = 0; i < s->n; i++)
			if (s->inplay[i] || on_chrom(s, i)) { waiting = 1; break; }
	if (!waiting) {
		/* strcmp-smallest chromosome among unfinished tracks (:50-74) */
		int best = -1;
		for (i = 0; i < s->n; i++)
			if (s->cur[i] < s->t[i].n) {
				int c = s->t[i].chrom[s->cur[i]];
				if (best < 0 || strcmp(s->names[c], s->names[best]) < 0) best = c;
			}
		if (best < 0) { s->done = 1; return 0; }
		s->chrom = best;
	}
	/* new start: abut if something is in play, else jump the gap (:112-115) */
	if (s->inplay_count)
		s->start = s->finish;
	else {
		int32_t m = INT32_MAX;
		for (i = 0; i < s->n; i++)
			if (!s->inplay[i] && on_chrom(s, i) && s->t[i].start[s->cur[i]] < m)
				m = s->t[i].start[s->cur[i]];
		s->start = m;
	}
	/* admit (:76-85) */
	for (i = 0; i < s->n; i++)
		if (!s->inplay[i] && on_chrom(s, i) && s->t[i].start[s->cur[i]] == s->start) {
			s->inplay[i] = 1; s->inplay_count++;
			s->x[i] = s->v[i] = s->t[i].value[s->cur[i]];
		}
	/* new finish (:87-96) */
	int32_t f = INT32_MAX;
	for (i = 0; i < s->n; i++) {
		if (s->inplay[i]) { if (s->t[i].end[s->cur[i]] < f) f = s->t[i].end[s->cur[i]]; }
		else if (on_chrom(s, i) && s->t[i].start[s->cur[i]] < f) f = s->t[i].start[s->cur[i]];
	}
	s->finish = f;
	return s->inplay_count == s->n;
}

static void sweep_pop(Sweep *s, int strict) {
	while (!s->done)
		if (sweep_step(s) || !strict) break;
}

/* ------------------------------------------------------------------------ */
/* Reducers (src/reducers.c restated; quirk ledger SURVEY.md section 8(a))   */
/* ------------------------------------------------------------------------ */
static int cmp_double(const void *a, const void *b) {
	double x = *(const double *) a, y = *(const double *) b;
	return (x < y) ? -1 : (x > y) ? 1 : 0;
}

/* x[i] = value-or-default, inplay[i], raw[i] = multiplexer values[] */
static double reduce_row(int op, int n, const double *x, const char *inplay, double *scratch) {
	int i;
	double acc;
	switch (op) {
	case WT_SUM: case WT_MEAN: /* :259-292, :367-402 */
		acc = 0;
		for (i = 0; i < n; i++) { if (isnan(x[i])) return NAN; acc += x[i]; }
		return op == WT_MEAN ? acc / n : acc;
	case WT_PRODUCT: /* :313-346 */
		acc = 1;
		for (i = 0; i < n; i++) { if (isnan(x[i])) return NAN; acc *= x[i]; }
		return acc;
	case WT_MAX: case WT_MIN: /* :125-168, :192-235: track 0 absent -> literal 0 */
		acc = inplay[0] ? x[0] : 0;
		if (isnan(acc)) return acc;
		for (i = 1; i < n; i++) {
			if (isnan(x[i])) return NAN;
			if (op == WT_MAX ? acc < x[i] : acc > x[i]) acc = x[i];
		}
		return acc;
	case WT_VAR: { /* :428-479: float-rounded mean, absent tracks skipped in the error pass */
		double mean = 0, count = 0;
		for (i = 0; i < n; i++) {
			float f = (float) x[i];
			if (isnan(f)) { mean = NAN; break; }
			mean += f; count++;
		}
		if (count < 2 || isnan(mean)) return NAN;
		mean /= count;
		acc = 0;
		for (i = 0; i < n; i++)
			if (inplay[i]) { double d = mean - x[i]; acc += d * d; }
		return acc / count;
	}
	case WT_STDDEV: case WT_ENTROPY: { /* :511-563; entropy registers the stddev pop (:665) */
		double mean = 0;
		for (i = 0; i < n; i++) {
			float f = (float) x[i];
			if (isnan(f)) return NAN;
			mean += f;
		}
		mean /= n;
		acc = 0;
		for (i = 0; i < n; i++) { double d = mean - x[i]; acc += d * d; }
		return sqrt(acc / n);
	}
	case WT_CV: { /* :672-725 */
		double mean = 0;
		for (i = 0; i < n; i++) {
			float f = (float) x[i];
			mean += f;
			if (isnan(f)) return NAN;
		}
		mean /= n;
		if (mean == 0) return NAN;
		acc = 0;
		for (i = 0; i < n; i++) { double d = mean - x[i]; acc += d * d; }
		return sqrt(acc / n) / mean;
	}
	case WT_FILLIN: case WT_TRIMFILL: /* :78-107: track 1's value (or default) on the union segmentation */
		return x[1];
	case WT_MEDIAN: /* :780-813: upper median, no averaging */
		for (i = 0; i < n; i++) { if (isnan(x[i])) return NAN; scratch[i] = x[i]; }
		qsort(scratch, n, sizeof(double), cmp_double);
		return scratch[n / 2];
	}
	return NAN;
}

/* Derived default_value of each reducer's output (the constructors). */
static double reduce_default(int op, int n, const double *d, double *scratch) {
	int i;
	double sum = 0;
	switch (op) {
	case WT_SUM: /* :294-307 */
		for (i = 0; i < n; i++) { if (isnan(d[i])) return NAN; sum += d[i]; }
		return sum;
	case WT_PRODUCT: { /* :348-361 */
		double p = 1;
		for (i = 0; i < n; i++) { if (isnan(d[i])) return NAN; p *= d[i]; }
		return p;
	}
	case WT_MEAN: { /* :404-422: stored through a float */
		for (i = 0; i < n; i++) { if (isnan(d[i])) return NAN; sum += d[i]; }
		float f = sum / n;
		return f;
	}
	case WT_VAR: case WT_STDDEV: { /* :481-505, :565-589 */
		for (i = 0; i < n; i++) { if (isnan(d[i])) return NAN; sum += d[i]; }
		double mean = sum / n, err = 0;
		for (i = 0; i < n; i++) err += (d[i] - mean) * (d[i] - mean);
		return op == WT_VAR ? err / n : sqrt(err / n);
	}
	case WT_ENTROPY: { /* :644-663: integer division count / n */
		int count = 0;
		for (i = 0; i < n; i++) { if (isnan(d[i])) return NAN; if (d[i] != 0) count++; }
		double p = count / n;
		return p ? -p * log(p) - (1 - p) * log(1 - p) : 0;
	}
	case WT_CV: { /* :727-750: through a float */
		for (i = 0; i < n; i++) { if (isnan(d[i])) return NAN; sum += d[i]; }
		double mean = sum / n, err = 0;
		for (i = 0; i < n; i++) err += (mean - d[i]) * (mean - d[i]);
		float f = sqrt(err / n) / mean;
		return f;
	}
	case WT_MAX: case WT_MIN: { /* :170-186, :237-253 */
		double m = d[0];
		if (isnan(m)) return m;
		for (i = 1; i < n; i++) {
			if (isnan(d[i])) return NAN;
			if (op == WT_MAX ? d[i] > m : d[i] < m) m = d[i];
		}
		return m;
	}
	case WT_FILLIN: case WT_TRIMFILL: /* :117 */
		return d[1];
	case WT_MEDIAN: { /* :815-834: through a float */
		for (i = 0; i < n; i++) { if (isnan(d[i])) return NAN; scratch[i] = d[i]; }
		qsort(scratch, n, sizeof(double), cmp_double);
		float f = scratch[n / 2];
		return f;
	}
	}
	return NAN;
}

static void put(wt_out *out, int64_t k, int chrom, int32_t s, int32_t f, double v) {
	if (out && k < out->cap) {
		out->chrom[k] = chrom; out->start[k] = s; out->end[k] = f; out->value[k] = v;
	}
}

int64_t wto_reduce(int op, int ntracks, const wt_track *tracks, int strict,
                   int nnames, char **names, wt_out *out, double *default_out) {
	(void) nnames;
	Sweep s;
	sweep_init(&s, ntracks, tracks, names);
	double *scratch = (double *) calloc(ntracks, sizeof(double));
	if (default_out) {
		double *d = (double *) calloc(ntracks, sizeof(double));
		for (int i = 0; i < ntracks; i++) d[i] = tracks[i].default_value;
		*default_out = reduce_default(op, ntracks, d, scratch);
		free(d);
	}
	int64_t k = 0;
	for (sweep_pop(&s, strict); !s.done; sweep_pop(&s, strict)) {
		if (op == WT_TRIMFILL && !s.inplay[0]) continue;   /* trim: only where track 0 is in play (:88-96) */
		put(out, k, s.chrom, s.start, s.finish, reduce_row(op, ntracks, s.x, s.inplay, scratch));
		k++;
	}
	free(scratch);
	sweep_free(&s);
	return k;
}

/* Raw multiplexer rows: values[] (default when absent) and inplay[] per segment. */
int64_t wto_rows(int ntracks, const wt_track *tracks, int strict, int nnames, char **names,
                 wt_out *out, double *row_values, char *row_inplay) {
	(void) nnames;
	Sweep s;
	sweep_init(&s, ntracks, tracks, names);
	int64_t k = 0;
	for (sweep_pop(&s, strict); !s.done; sweep_pop(&s, strict)) {
		if (k < out->cap) {
			put(out, k, s.chrom, s.start, s.finish, 0);
			for (int i = 0; i < ntracks; i++) {
				row_values[k * ntracks + i] = s.v[i];
				row_inplay[k * ntracks + i] = s.inplay[i];
			}
		}
		k++;
	}
	sweep_free(&s);
	return k;
}

/* ------------------------------------------------------------------------ */
/* Set comparisons (src/multiSet.c:21-101 + src/setComparisons.c restated).  */
/* The two-level sweep equals one sweep over n1+n2 tracks in which set m is  */
/* "in play" when any of its members is.                                     */
/* ------------------------------------------------------------------------ */
typedef struct { double value; char set; int ord; } VSP;

static int cmp_vsp(const void *a, const void *b) {
	const VSP *x = (const VSP *) a, *y = (const VSP *) b;
	if (x->value < y->value) return -1;
	if (x->value > y->value) return 1;
	return x->ord - y->ord; /* glibc qsort on this size is a stable merge sort */
}

static double ttest_row(int n1, int n2, const double *x, const char *inplay) {
	/* src/setComparisons.c:56-117: sums over in-play members, counts = set sizes */
	double sum1 = 0, sum2 = 0, sq1 = 0, sq2 = 0;
	int i, count1 = n1, count2 = n2;
	for (i = 0; i < n1; i++) if (inplay[i]) { sum1 += x[i]; sq1 += x[i] * x[i]; }
	for (i = 0; i < n2; i++) if (inplay[n1 + i]) { sum2 += x[n1 + i]; sq2 += x[n1 + i] * x[n1 + i]; }
	double mean1 = sum1 / count1, mean2 = sum2 / count2;
	double var1 = sq1 / count1 - mean1 * mean1, var2 = sq2 / count2 - mean2 * mean2;
	if (var1 + var2 == 0) return NAN;
	double t = (mean1 - mean2) / sqrt(var1 / count1 + var2 / count2);
	if (t < 0) t = -t;
	double nu = (var1 / count1 + var2 / count2) * (var1 / count1 + var2 / count2)
	          / ((var1 * var1) / (count1 * count1 * (count1 - 1)) + (var2 * var2) / (count2 * count2 * (count2 - 1)));
	return 2 * gsl_cdf_tdist_Q(t, nu);
}

static double mwu_row(int n1, int n2, const double *x, VSP *tab, double mu_U, double sigma_U, double prev_value) {
	/* src/setComparisons.c:293-366 */
	int N = n1 + n2, i;
	for (i = 0; i < N; i++) {
		if (isnan(x[i])) return NAN;
		tab[i].value = x[i]; tab[i].set = i >= n1; tab[i].ord = i;
	}
	qsort(tab, N, sizeof(VSP), cmp_vsp);
	double U1 = 0;
	int prev = 0, ties = 0, previousTies = 0;
	for (i = 0; i < N && prev < n1; i++) {
		if (!tab[i].set) {
			U1 += i - prev;
			int j;
			if (ties) {
				for (j = i + 1; j < N && tab[j].value == tab[i].value && tab[j].set; j++)
					previousTies++;
				U1 -= previousTies / 2.0;
				U1 += (ties - previousTies) / 2.0;
				if (previousTies == ties) previousTies = ties = 0;
			} else {
				for (j = i + 1; j < N && tab[j].value == tab[i].value; j++)
					if (tab[j].set) ties++;
				if (ties) U1 += ties / 2.0;
			}
			prev++;
		}
	}
	(void) prev_value;
	if (U1 > mu_U) return 2 * erf((mu_U - U1) / sigma_U);
	return 2 * erf((U1 - mu_U) / sigma_U);
}

int64_t wto_setcompare(int op, int n1, const wt_track *t1, int n2, const wt_track *t2,
                       int nnames, char **names, wt_out *out) {
	(void) nnames;
	int N = n1 + n2, i;
	wt_track *all = (wt_track *) calloc(N, sizeof(wt_track));
	memcpy(all, t1, n1 * sizeof(wt_track));
	memcpy(all + n1, t2, n2 * sizeof(wt_track));
	Sweep s;
	sweep_init(&s, N, all, names);
	VSP *tab = (VSP *) calloc(N, sizeof(VSP));
	/* integer division in both moments (src/setComparisons.c:386-387) */
	double mu_U = n1 * n2 / 2;
	double sigma_U = sqrt(n1 * n2 * (n1 + n2 + 1) / 12);
	int64_t k = 0;
	for (sweep_pop(&s, 0); !s.done; sweep_pop(&s, 0)) {
		int c1 = 0, c2 = 0;
		for (i = 0; i < n1; i++) c1 += s.inplay[i];
		for (i = n1; i < N; i++) c2 += s.inplay[i];
		if (!c1 || !c2) continue; /* both sets must be in play (:48-54, :282-288) */
		double v = (op == WT_TTEST) ? ttest_row(n1, n2, s.x, s.inplay)
		                            : mwu_row(n1, n2, s.x, tab, mu_U, sigma_U, 0);
		put(out, k, s.chrom, s.start, s.finish, v);
		k++;
	}
	free(tab); free(all);
	sweep_free(&s);
	return k;
}

/* ------------------------------------------------------------------------ */
/* Whole-genome scalars (src/statistics.c:103-120, :414-458 restated)        */
/* ------------------------------------------------------------------------ */
double wto_auc(int op, int ntracks, const wt_track *tracks, int strict, int nnames, char **names) {
	(void) nnames;
	Sweep s;
	sweep_init(&s, ntracks, tracks, names);
	double *scratch = (double *) calloc(ntracks, sizeof(double));
	double res = 0;
	for (sweep_pop(&s, strict); !s.done; sweep_pop(&s, strict)) {
		double v = reduce_row(op, ntracks, s.x, s.inplay, scratch);
		if (!isnan(v)) res += (s.finish - s.start) * v;
	}
	free(scratch);
	sweep_free(&s);
	return res;
}

double wto_pearson(const wt_track *x, const wt_track *y, int nnames, char **names) {
	(void) nnames;
	wt_track two[2] = { *x, *y };
	Sweep s;
	sweep_init(&s, 2, two, names);
	int count = 0; /* int, as in the reference (:400) */
	double sum_X = 0, sum_Y = 0, T_XX = 0, T_XY = 0, T_YY = 0;
	for (sweep_pop(&s, 0); !s.done; sweep_pop(&s, 0)) {
		double X = s.x[0], Y = s.x[1];
		int length = s.finish - s.start;
		if (count) {
			double old_mean_X = sum_X / count, new_mean_X = sum_X / (count + length);
			double old_mean_Y = sum_Y / count, new_mean_Y = sum_Y / (count + length);
			double ratio = (double) count / (count + length);
			T_XY += (new_mean_X * old_mean_Y + ratio * X * Y - new_mean_X * Y - new_mean_Y * X) * length;
			T_XX += (new_mean_X * (old_mean_X - 2 * X) + ratio * X * X) * length;
			T_YY += (new_mean_Y * (old_mean_Y - 2 * Y) + ratio * Y * Y) * length;
		}
		count += length;
		sum_X += X * length;
		sum_Y += Y * length;
	}
	sweep_free(&s);
	if (T_XX * T_YY != 0) return T_XY / sqrt(T_XX * T_YY);
	return NAN;
}

/* N-dimensional Pearson over two sets of equal size (src/statistics.c:490-556):
 * rank = multi->count = 2 whatever the set size (:551); the sweep covers every
 * segment where any track of either set is in play; an absent track reads its
 * default. */
double wto_ndpearson(int n, const wt_track *a, const wt_track *b, int nnames, char **names) {
	(void) nnames;
	wt_track *all = (wt_track *) calloc(2 * n, sizeof(wt_track));
	memcpy(all, a, n * sizeof(wt_track));
	memcpy(all + n, b, n * sizeof(wt_track));
	Sweep s;
	sweep_init(&s, 2 * n, all, names);
	const int rank = 2;
	int count = 0;
	double sum_X[2] = {0, 0}, sum_Y[2] = {0, 0}, T_XX = 0, T_XY = 0, T_YY = 0;
	for (sweep_pop(&s, 0); !s.done; sweep_pop(&s, 0)) {
		int length = s.finish - s.start;
		double ratio = (double) count / (count + length);
		for (int dim = 0; dim < rank; dim++) {
			double X = s.x[dim], Y = s.x[n + dim];
			if (count) {
				double omx = sum_X[dim] / count, nmx = sum_X[dim] / (count + length);
				double omy = sum_Y[dim] / count, nmy = sum_Y[dim] / (count + length);
				T_XY += (nmx * omy + ratio * X * Y - nmx * Y - nmy * X) * length;
				T_XX += (nmx * (omx - 2 * X) + ratio * X * X) * length;
				T_YY += (nmy * (omy - 2 * Y) + ratio * Y * Y) * length;
			}
			sum_X[dim] += X * length;
			sum_Y[dim] += Y * length;
		}
		count += length;
	}
	sweep_free(&s);
	free(all);
	if (T_XX * T_YY != 0) return T_XY / sqrt(T_XX * T_YY);
	return NAN;
}

/* Writer-side run-length compression (src/unaryOps.c:235-253) of a segment list. */
int64_t wto_compress(int64_t n, const int32_t *chrom, const int32_t *start, const int32_t *end,
                     const double *value, wt_out *out) {
	int64_t k = 0, i = 0;
	while (i < n) {
		int32_t c = chrom[i], s = start[i], f = end[i];
		double v = value[i];
		i++;
		while (i < n && chrom[i] == c && start[i] == f &&
		       ((isnan(value[i]) && isnan(v)) || fabs(value[i] - v) < 0.000001)) {
			f = end[i];
			i++;
		}
		put(out, k, c, s, f, v);
		k++;
	}
	return k;
}

/* The default text sink (src/wigWriter.c) restated for a segment list that is already
 * in memory: `bedgraph` = write_bg (no compression, bedGraph lines only), else `write`
 * (compress first, src/wigWriter.c:261-266).  Records are handed over in blocks of
 * 10000 (BLOCK_LENGTH, :29, :174-186) and every block is printed with fresh state
 * (printBlock, :55-119): point-by-point mode is entered by a record shorter than 2 bp
 * and left by one longer than 5 bp; in that mode a "fixedStep" header is printed when
 * the mode was just entered, the chromosome changed or the record does not abut the
 * previous one, then one "%lf" line per position; otherwise a 0-based bedGraph line.
 * Returns the number of records printed. */
int64_t wto_write(int64_t n, const int32_t *chrom, const int32_t *start, const int32_t *end,
                  const double *value, int nnames, char **names, int bedgraph, const char *path) {
	FILE *f = fopen(path, "w");
	if (!f) return -1;
	(void) nnames;
	int64_t records = 0, i = 0;
	int pbp = 0, last_chrom = -1, last_finish = -1;
	while (i < n) {
		int32_t c = chrom[i], s = start[i], e = end[i];
		double v = value[i];
		i++;
		if (!bedgraph)
			while (i < n && chrom[i] == c && start[i] == e &&
			       ((isnan(value[i]) && isnan(v)) || fabs(value[i] - v) < 0.000001)) {
				e = end[i];
				i++;
			}
		if (records % 10000 == 0) { pbp = 0; last_chrom = -1; last_finish = -1; }
		int make_header = 0;
		if (!bedgraph && e - s < 2 && !pbp) { pbp = 1; make_header = 1; }
		else if (e - s > 5 && pbp) pbp = 0;
		if (pbp) {
			if (make_header || last_chrom != c || s > last_finish)
				fprintf(f, "fixedStep chrom=%s start=%i step=1\n", names[c], s);
			for (int j = 0; j < e - s; j++) fprintf(f, "%lf\n", v);
		} else
			fprintf(f, "%s\t%i\t%i\t%lf\n", names[c], s - 1, e - 1, v);
		last_chrom = c;
		last_finish = e;
		records++;
	}
	fclose(f);
	return records;
}

/* Element-wise unary operators of `map` / the unary_operator grammar (src/commandParser.c:115-211),
 * restated from their pop functions and constructors in src/unaryOps.c: a chain applied to one
 * track, innermost first.  codes: 1 scale s (:650-682), 2 offset s (:722-745), 3 log base, base <= 0
 * meaning `ln` (:760-811; records with value <= 0 are dropped), 4 exp (:823-866, natural), 5 pow s
 * (:873-901), 6 abs (:934-960), 7 default d (:210-229).  Returns the number of records kept;
 * *default_out receives the chain's derived default_value (some constructors round it to float). */
int64_t wto_map(const wt_track *t, int nmaps, const int *codes, const double *args, wt_out *out, double *default_out) {
	double dflt = t->default_value;
	double *v = (double *) malloc(sizeof(double) * (size_t) (t->n > 0 ? t->n : 1));
	char *keep = (char *) malloc((size_t) (t->n > 0 ? t->n : 1));
	for (int64_t j = 0; j < t->n; j++) { v[j] = t->value[j]; keep[j] = 1; }
	for (int c = 0; c < nmaps; c++) {
		const double a = args[c];
		double l = 1;
		switch (codes[c]) {
		case 1:
			for (int64_t j = 0; j < t->n; j++) if (keep[j]) v[j] = isnan(v[j]) ? NAN : a * v[j];
			dflt = isnan(dflt) ? NAN : (double) (float) (dflt * a);
			break;
		case 2:
			for (int64_t j = 0; j < t->n; j++) if (keep[j]) v[j] = a + v[j];
			dflt = isnan(dflt) ? NAN : (double) (float) (dflt + a);
			break;
		case 3:
			if (a > 0) l = log(a);
			for (int64_t j = 0; j < t->n; j++) if (keep[j]) {
				if (v[j] <= 0) keep[j] = 0;
				else v[j] = (isnan(v[j]) || v[j] < 0) ? NAN : log(v[j]) / l;
			}
			dflt = (!isnan(dflt) && dflt > 0) ? log(dflt) / l : NAN;
			break;
		case 4:
			for (int64_t j = 0; j < t->n; j++) if (keep[j]) v[j] = exp(v[j] * 1);
			dflt = isnan(dflt) ? NAN : (double) (float) exp(dflt * 1);
			break;
		case 5:
			for (int64_t j = 0; j < t->n; j++) if (keep[j]) v[j] = ((a < 0 && v[j] <= 0) || isnan(v[j])) ? NAN : pow(v[j], a);
			dflt = (!isnan(dflt) && (dflt > 0 || a > 0)) ? pow(dflt, a) : NAN;
			break;
		case 6:
			for (int64_t j = 0; j < t->n; j++) if (keep[j]) v[j] = isnan(v[j]) ? NAN : fabs(v[j]);
			dflt = isnan(dflt) ? NAN : fabs(dflt);
			break;
		case 7:
			dflt = a;
			break;
		}
	}
	int64_t k = 0;
	for (int64_t j = 0; j < t->n; j++) if (keep[j]) { put(out, k, t->chrom[j], t->start[j], t->end[j], v[j]); k++; }
	free(v); free(keep);
	if (default_out) *default_out = dflt;
	return k;
}

/* `apply` (src/apply.c + the integrators of src/statistics.c:62-326), restated per
 * region.  Every region is served from a per-base buffer (createBufferedWiggleIteratorData always
 * allocates one, so the seek / fill-in branches of computeApplyValues, apply.c:266-273, are never taken):
 * dataset records are written base by base into values[]/set[] (pushDataOnBuffer, :217-237) and read back
 *   - strict: one 1-bp record per set base (StrictBufferedWiggleIteratorPop, :87-101);
 *   - loose : one 1-bp record per set base, one record with the dataset's default value (rounded to
 *             float by createBufferedWiggleIteratorData's parameter type, :41) per run of unset bases
 *             (LooseBufferedWiggleIteratorPop, :66-85);
 * in region-relative coordinates, and the chained integrators fold those records in order.
 * codes: 1 AUC (:103-126), 2 meanI (:62-96), 3 varI, 4 stddevI, 5 CVI (:233-326; a NaN record would hang
 * the reference, :238-239 -- not reproduced, NaN is skipped), 6 maxI (:162-186), 7 minI (:192-216),
 * 8 spanI (:132-156).  out[r * nstats + k] = statistic k of region r.  Regions and the dataset's segments
 * are sorted by (chrom, start); segments do not overlap. */
int64_t wto_apply(int64_t nseg, const int32_t *chrom, const int32_t *start, const int32_t *end, const double *value,
                  double default_value, int64_t nreg, const int32_t *rchrom, const int32_t *rstart, const int32_t *rend,
                  int nstats, const int *codes, int strict, double *out) {
	const double dflt = (double) (float) default_value;
	int64_t first = 0;
	for (int64_t r = 0; r < nreg; r++) {
		const int len = rend[r] - rstart[r];
		double *vals = (double *) calloc(len > 0 ? len : 1, sizeof(double));
		char *set = (char *) calloc(len > 0 ? len : 1, 1);
		/* first segment that can reach this region (regions are sorted, so the scan only moves forward) */
		while (first < nseg && (chrom[first] < rchrom[r] || (chrom[first] == rchrom[r] && end[first] <= rstart[r]))) first++;
		int64_t back = first;      /* an earlier, longer region may have pushed `first` too far: step back over overlaps */
		while (back > 0 && chrom[back - 1] == rchrom[r] && end[back - 1] > rstart[r]) back--;
		for (int64_t k = back; k < nseg && chrom[k] == rchrom[r] && start[k] < rend[r]; k++) {
			int a = start[k] > rstart[r] ? start[k] - rstart[r] : 0;
			int b = end[k] < rend[r] ? end[k] - rstart[r] : len;
			for (int p = a; p < b; p++) { vals[p] = value[k]; set[p] = 1; }
		}
		/* integrator states */
		double auc = 0, msum = 0, mspan = 0, span = 0, mx = NAN, mn = NAN;
		double T = 0, vsum = 0; long vcount = 0;
		int idx = 0;
		while (idx < len) {
			int rs = idx, rf;
			double v;
			if (set[idx]) { v = vals[idx]; idx++; rf = idx; }
			else {
				while (idx < len && !set[idx]) idx++;
				rf = idx;
				if (strict) continue;
				v = dflt;
			}
			const int L = rf - rs;
			if (isnan(v)) continue;
			auc += L * v;
			msum += L * v; mspan += L;
			span += L;
			if (isnan(mx) || v > mx) mx = v;
			if (isnan(mn) || v < mn) mn = v;
			if (vcount) {
				double old_mean = vsum / vcount;
				double new_mean = vsum / (vcount + L);
				double delta_T = old_mean * new_mean - new_mean * 2 * v + ((double) vcount / (vcount + L)) * v * v;
				T += delta_T * L;
			}
			vcount += L;
			vsum += L * v;
		}
		for (int k = 0; k < nstats; k++) {
			double res = NAN;
			switch (codes[k]) {
			case 1: res = auc; break;
			case 2: res = mspan > 0 ? msum / mspan : NAN; break;
			case 3: res = T / (vcount - 1); break;
			case 4: res = sqrt(T / (vcount - 1)); break;
			case 5: res = sqrt(T / (vcount - 1)) / (vsum / vcount); break;
			case 6: res = mx; break;
			case 7: res = mn; break;
			case 8: res = span; break;
			}
			out[r * nstats + k] = res;
		}
		free(vals); free(set);
	}
	return nreg;
}

double wto_tdist_Q(double t, double nu) { return gsl_cdf_tdist_Q(t, nu); }
