/* TEST INFRASTRUCTURE ONLY (oracle/shim): the part of libBigWig's interface that
 * /root/reference/src/bigWiggleReader.c uses (libBigWig itself -- an unpinned third-party
 * dependency of the reference, src/Makefile:4 -- is not in this image).  Declared from the
 * reference's call sites (bigWiggleReader.c:30,38-46,58,64-70,91-94,116,120), implemented in
 * oracle/bigwig_shim.c on top of the library's block decoder with its stretch emulation switched
 * OFF, so that the reference's own reader code does the 10000-bp stretch queries and the boxing
 * (quirk Q9) -- which is what the decoder's emulation of them is then compared with. */
#ifndef WB200_SHIM_BIGWIG_H
#define WB200_SHIM_BIGWIG_H
#include <stdint.h>

typedef struct {
	int64_t nKeys;
	char **chrom;
	uint32_t *len;
} chromList_t;

typedef struct {
	chromList_t *cl;
	void *handle;              /* wb200_bigwig* */
	/* one decoded chromosome cached at a time (raw records, 0-based half-open) */
	int cached;
	int64_t n;
	uint32_t *start, *end;
	float *value;
} bigWigFile_t;

typedef struct {
	uint32_t l, m;
	uint32_t *start, *end;
	float *value;
} bwOverlappingIntervals_t;

typedef struct {
	bigWigFile_t *bw;
	uint32_t tid, start, end;
	int64_t next, stop;        /* records still to hand out: [next, stop) of the cached chromosome */
	uint32_t blocksPerIteration;
	bwOverlappingIntervals_t *intervals;
	void *entries;             /* bigBed: unused here */
	void *data;                /* non-NULL while `intervals` holds something */
} bwOverlapIterator_t;

int bwInit(size_t bufSize);
int bwIsBigWig(char *fname, void *callback);
bigWigFile_t *bwOpen(char *fname, void *callback, const char *mode);
bwOverlapIterator_t *bwOverlappingIntervalsIterator(bigWigFile_t *fp, char *chrom, uint32_t start, uint32_t end, uint32_t blocksPerIteration);
bwOverlapIterator_t *bwIteratorNext(bwOverlapIterator_t *iter);
void bwIteratorDestroy(bwOverlapIterator_t *iter);
#endif
