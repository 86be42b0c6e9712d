/* The reference-side binding of the device text sink: what a maintainer of
 * WiggleTools would add to src/wigWriter.c (three lines at the top of
 * TeeWiggleIterator / toFile / toStdout) so that `write`, `write_bg` and the
 * default print-to-stdout hand a GPU reduction straight to
 * wb200_sink_run() (include/wiggletools_dropin.h) instead of popping it
 * record by record into the writer thread.
 *
 * /root/reference is read-only here, so the same routing is attached from
 * outside with the linker's --wrap (integration/Makefile):
 * calls from src/commandParser.c to TeeWiggleIterator, runWiggleIterator,
 * toFile and toStdout arrive in the __wrap_ functions below; everything that
 * is not a fresh reduction of the library goes on to the reference's own code
 * (__real_*), unchanged.
 *
 * `write f.wig <reduction>` builds a Tee and then runs it
 * (src/commandParser.c:375-383,987-988); a Tee can also sit in the middle of
 * a pipeline and feed its records on.  The wrapper therefore returns a
 * deferred Tee showing the reduction's first record, as the real one would:
 * if it is simply run to the end the device sink prints everything, and if
 * anybody pops or seeks it, it becomes the reference's Tee at that moment. */
#include <stdio.h>
#include <stdlib.h>
#include "wiggletools.h"
#include "wiggleIterator.h"

int wb200_sink_accepts(WiggleIterator *wi);
void wb200_sink_run(WiggleIterator *wi, FILE *out, bool bedGraph);

WiggleIterator *__real_TeeWiggleIterator(WiggleIterator *, FILE *, bool, bool);
void __real_runWiggleIterator(WiggleIterator *);
void __real_toFile(WiggleIterator *, char *, bool, bool);
void __real_toStdout(WiggleIterator *, bool, bool);

typedef struct {
	WiggleIterator *source;
	FILE *file;
	bool bedGraph;
	WiggleIterator *real;      /* the reference's Tee once somebody consumes records */
} DeferredTee;

static void mirror(WiggleIterator *wi, WiggleIterator *from) {
	wi->chrom = from->chrom;
	wi->start = from->start;
	wi->finish = from->finish;
	wi->value = from->value;
	wi->done = from->done;
}

static void deferred_pop(WiggleIterator *wi) {
	DeferredTee *d = (DeferredTee *) wi->data;
	if (!d->real)
		d->real = __real_TeeWiggleIterator(d->source, d->file, d->bedGraph, false);   /* shows record 1 again */
	pop(d->real);
	mirror(wi, d->real);
}

static void deferred_seek(WiggleIterator *wi, const char *chrom, int start, int finish) {
	DeferredTee *d = (DeferredTee *) wi->data;
	if (!d->real)
		d->real = __real_TeeWiggleIterator(d->source, d->file, d->bedGraph, true);
	seek(d->real, chrom, start, finish);
	mirror(wi, d->real);
}

WiggleIterator *__wrap_TeeWiggleIterator(WiggleIterator *i, FILE *outfile, bool bedGraph, bool holdFire) {
	if (holdFire || !wb200_sink_accepts(i))
		return __real_TeeWiggleIterator(i, outfile, bedGraph, holdFire);
	DeferredTee *d = (DeferredTee *) calloc(1, sizeof(DeferredTee));
	d->source = i;
	d->file = outfile;
	d->bedGraph = bedGraph;
	/* newWiggleIterator (src/wiggleIterator.c:20-34) without its eager pop */
	WiggleIterator *wi = (WiggleIterator *) calloc(1, sizeof(WiggleIterator));
	wi->data = d;
	wi->pop = &deferred_pop;
	wi->seek = &deferred_seek;
	wi->overlaps = i->overlaps;
	wi->default_value = i->default_value;
	mirror(wi, i);
	return wi;
}

void __wrap_runWiggleIterator(WiggleIterator *wi) {
	if (wi->pop == &deferred_pop && !((DeferredTee *) wi->data)->real && !wi->done) {
		DeferredTee *d = (DeferredTee *) wi->data;
		wb200_sink_run(d->source, d->file, d->bedGraph);
		wi->done = true;
		return;
	}
	__real_runWiggleIterator(wi);
}

void __wrap_toFile(WiggleIterator *wi, char *filename, bool bedGraph, bool holdFire) {
	if (holdFire || !wb200_sink_accepts(wi)) {
		__real_toFile(wi, filename, bedGraph, holdFire);
		return;
	}
	FILE *file = fopen(filename, "w");
	if (!file) {
		fprintf(stderr, "Could not open file %s\n", filename);
		exit(1);
	}
	wb200_sink_run(wi, file, bedGraph);
	fclose(file);
}

void __wrap_toStdout(WiggleIterator *wi, bool bedGraph, bool holdFire) {
	if (holdFire || !wb200_sink_accepts(wi)) {
		__real_toStdout(wi, bedGraph, holdFire);
		return;
	}
	wb200_sink_run(wi, stdout, bedGraph);
}

/* ---- element-wise unary operators -------------------------------------------------
 * In src/unaryOps.c the edit would be one line at the top of each constructor, e.g.
 *     WiggleIterator * ScaleWiggleIterator(WiggleIterator * i, double s) {
 *         return wb200_mapped_iterator(i, WB200_MAP_SCALE, s);
 * Popped on the host the returned iterator does what the reference's does; as the input
 * of a GPU reduction its maps run on the device (`mean map ln a.bw b.bw c.bw`). */
WiggleIterator *wb200_mapped_iterator(WiggleIterator *i, int op, double arg);
WiggleIterator *__wrap_ScaleWiggleIterator(WiggleIterator *i, double s) { return wb200_mapped_iterator(i, 1, s); }
WiggleIterator *__wrap_ShiftWiggleIterator(WiggleIterator *i, double s) { return wb200_mapped_iterator(i, 2, s); }
WiggleIterator *__wrap_NaturalLogWiggleIterator(WiggleIterator *i) { return wb200_mapped_iterator(i, 3, 0); }
WiggleIterator *__wrap_LogWiggleIterator(WiggleIterator *i, double s) { return wb200_mapped_iterator(i, 3, s); }
WiggleIterator *__wrap_NaturalExpWiggleIterator(WiggleIterator *i) { return wb200_mapped_iterator(i, 4, 0); }
WiggleIterator *__wrap_PowerWiggleIterator(WiggleIterator *i, double s) { return wb200_mapped_iterator(i, 5, s); }
WiggleIterator *__wrap_AbsWiggleIterator(WiggleIterator *i) { return wb200_mapped_iterator(i, 6, 0); }
WiggleIterator *__wrap_DefaultValueWiggleIterator(WiggleIterator *i, double v) { return wb200_mapped_iterator(i, 0, v); }

/* ---- regional integrators (`apply`, `apply_paste`) ---------------------------------
 * In src/apply.c the edit would be four lines at the top of ApplyMultiplexer (:340):
 *     Multiplexer * gpu = wb200_apply_multiplexer(regions, codes_of(statistics, count), count, dataset, strict);
 *     if (gpu) return gpu;
 * A dataset that is a fresh reduction of the library (`apply_paste out meanI regions.bed mean a.bw b.bw`)
 * is then sought and reduced per run of regions on the device and the integrators run on the
 * resident segments; any other dataset, or a statistic the device does not provide, keeps the
 * reference's per-base buffers. */
Multiplexer *wb200_apply_multiplexer(WiggleIterator *regions, const int *stat_codes, int count, WiggleIterator *dataset, bool strict);
Multiplexer *__real_ApplyMultiplexer(WiggleIterator *regions, WiggleIterator *(**statistics)(WiggleIterator *), int count, WiggleIterator *dataset, bool strict);
WiggleIterator *SpanIntegrator(WiggleIterator *);      /* src/statistics.c:152 (not in wiggletools.h) */

Multiplexer *__wrap_ApplyMultiplexer(WiggleIterator *regions, WiggleIterator *(**statistics)(WiggleIterator *), int count, WiggleIterator *dataset, bool strict) {
	int codes[64];
	bool known = statistics != NULL && count >= 1 && count <= 64;
	for (int i = 0; known && i < count; i++) {
		WiggleIterator *(*f)(WiggleIterator *) = statistics[i];
		if (f == &AUCIntegrator) codes[i] = 1;
		else if (f == &MeanIntegrator) codes[i] = 2;
		else if (f == &VarianceIntegrator) codes[i] = 3;
		else if (f == &StandardDeviationIntegrator) codes[i] = 4;
		else if (f == &CoefficientOfVariationIntegrator) codes[i] = 5;
		else if (f == &MaxIntegrator) codes[i] = 6;
		else if (f == &MinIntegrator) codes[i] = 7;
		else if (f == &SpanIntegrator) codes[i] = 8;
		else known = false;
	}
	if (known) {
		Multiplexer *gpu = wb200_apply_multiplexer(regions, codes, count, dataset, strict);
		if (gpu) return gpu;
	}
	return __real_ApplyMultiplexer(regions, statistics, count, dataset, strict);
}
