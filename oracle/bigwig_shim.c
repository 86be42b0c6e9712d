/* TEST INFRASTRUCTURE ONLY: libBigWig's interface (oracle/shim/bigWig.h) over the library's host
 * BigWig decoder, so that the reference's unmodified src/bigWiggleReader.c can be compiled and run
 * here.  The decoder is used with stretch = 0: it hands out the file's records as they are stored;
 * "overlapping intervals" follows libBigWig's published behaviour -- every stored interval that
 * intersects the 0-based half-open query, with its own (unclipped) coordinates, in file order,
 * `blocksPerIteration` groups at a time. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "bigWig.h"
#include "../include/wiggletools_b200.h"

#define ITEMS_PER_BLOCK 512     /* one iteration hands out blocksPerIteration * 512 records, like data sections of that size */

int bwInit(size_t bufSize) { (void) bufSize; return 0; }

int bwIsBigWig(char *fname, void *callback) {
	(void) callback;
	wb200_bigwig *h = wb200_bw_open(fname);
	if (!h) return 0;
	wb200_bw_close(h);
	return 1;
}

bigWigFile_t *bwOpen(char *fname, void *callback, const char *mode) {
	(void) callback; (void) mode;
	wb200_bigwig *h = wb200_bw_open(fname);
	if (!h) return NULL;
	wb200_bw_set_stretch(h, 0);
	bigWigFile_t *fp = (bigWigFile_t *) calloc(1, sizeof(bigWigFile_t));
	fp->handle = h;
	fp->cached = -1;
	fp->cl = (chromList_t *) calloc(1, sizeof(chromList_t));
	fp->cl->nKeys = wb200_bw_nchrom(h);
	fp->cl->chrom = (char **) calloc(fp->cl->nKeys ? fp->cl->nKeys : 1, sizeof(char *));
	fp->cl->len = (uint32_t *) calloc(fp->cl->nKeys ? fp->cl->nKeys : 1, sizeof(uint32_t));
	for (int64_t i = 0; i < fp->cl->nKeys; i++) {
		fp->cl->chrom[i] = strdup(wb200_bw_chrom_name(h, (int) i));
		fp->cl->len[i] = (uint32_t) wb200_bw_chrom_length(h, (int) i);
	}
	return fp;
}

static int load(bigWigFile_t *fp, int tid) {
	if (fp->cached == tid) return 0;
	wb200_track t;
	if (wb200_bw_decode_chrom((wb200_bigwig *) fp->handle, tid, &t) != 0) return 1;
	free(fp->start); free(fp->end); free(fp->value);
	fp->n = t.n;
	fp->start = (uint32_t *) malloc(sizeof(uint32_t) * (size_t) (t.n ? t.n : 1));
	fp->end = (uint32_t *) malloc(sizeof(uint32_t) * (size_t) (t.n ? t.n : 1));
	fp->value = (float *) malloc(sizeof(float) * (size_t) (t.n ? t.n : 1));
	for (int64_t k = 0; k < t.n; k++) {          /* decoder: 1-based half-open doubles -> libBigWig: 0-based floats */
		fp->start[k] = (uint32_t) (t.start[k] - 1);
		fp->end[k] = (uint32_t) (t.end[k] - 1);
		fp->value[k] = (float) t.value[k];
	}
	fp->cached = tid;
	return 0;
}

static void fill(bwOverlapIterator_t *it) {
	bigWigFile_t *fp = it->bw;
	bwOverlappingIntervals_t *iv = it->intervals;
	iv->l = 0;
	const int64_t want = (int64_t) it->blocksPerIteration * ITEMS_PER_BLOCK;
	while (it->next < it->stop && iv->l < want) {
		const int64_t k = it->next++;
		if (fp->end[k] <= it->start || fp->start[k] >= it->end) continue;
		if (iv->l == iv->m) {
			iv->m = iv->m ? 2 * iv->m : 1024;
			iv->start = (uint32_t *) realloc(iv->start, sizeof(uint32_t) * iv->m);
			iv->end = (uint32_t *) realloc(iv->end, sizeof(uint32_t) * iv->m);
			iv->value = (float *) realloc(iv->value, sizeof(float) * iv->m);
		}
		iv->start[iv->l] = fp->start[k]; iv->end[iv->l] = fp->end[k]; iv->value[iv->l] = fp->value[k];
		iv->l++;
	}
	it->data = iv->l ? (void *) iv : NULL;
}

bwOverlapIterator_t *bwOverlappingIntervalsIterator(bigWigFile_t *fp, char *chrom, uint32_t start, uint32_t end, uint32_t blocksPerIteration) {
	int tid = -1;
	for (int64_t i = 0; i < fp->cl->nKeys; i++) if (!strcmp(fp->cl->chrom[i], chrom)) tid = (int) i;
	if (tid < 0 || load(fp, tid)) return NULL;
	bwOverlapIterator_t *it = (bwOverlapIterator_t *) calloc(1, sizeof(bwOverlapIterator_t));
	it->bw = fp; it->tid = (uint32_t) tid; it->start = start; it->end = end;
	it->blocksPerIteration = blocksPerIteration ? blocksPerIteration : 1;
	it->intervals = (bwOverlappingIntervals_t *) calloc(1, sizeof(bwOverlappingIntervals_t));
	/* records are sorted by start and do not overlap: binary search for the first one that can intersect */
	int64_t lo = 0, hi = fp->n;
	while (lo < hi) { const int64_t mid = (lo + hi) / 2; if (fp->end[mid] <= start) lo = mid + 1; else hi = mid; }
	it->next = lo;
	lo = it->next; hi = fp->n;
	while (lo < hi) { const int64_t mid = (lo + hi) / 2; if (fp->start[mid] < end) lo = mid + 1; else hi = mid; }
	it->stop = lo;
	fill(it);
	return it;
}

bwOverlapIterator_t *bwIteratorNext(bwOverlapIterator_t *it) {
	if (it) fill(it);
	return it;
}

void bwIteratorDestroy(bwOverlapIterator_t *it) {
	if (!it) return;
	if (it->intervals) { free(it->intervals->start); free(it->intervals->end); free(it->intervals->value); free(it->intervals); }
	free(it);
}
