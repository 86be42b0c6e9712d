/* TEST INFRASTRUCTURE ONLY (oracle/): link stubs for the reference reader
 * constructors whose third-party libraries (libBigWig, htslib) are absent from
 * this image; used when linking the reference CLI against the GPU library
 * (integration/_build/wiggletools_b200).  Declarations: /root/reference/src/wiggletools.h:36-42,143. */
#include <stdio.h>
#include <stdlib.h>
#include "wiggletools.h"
static void absent(const char *what) {
	fprintf(stderr, "oracle/_ref: %s needs a library that is not in this image\n", what);
	exit(1);
}
#ifndef WB200_HAVE_BIGWIG
WiggleIterator *BigWiggleReader(char *f, bool h) { (void)f; (void)h; absent("BigWiggleReader"); return NULL; }
#endif
WiggleIterator *BigBedReader(char *f, bool h) { (void)f; (void)h; absent("BigBedReader"); return NULL; }
WiggleIterator *BamReader(char *f, bool h, bool r) { (void)f; (void)h; (void)r; absent("BamReader"); return NULL; }
WiggleIterator *BcfReader(char *f, bool h) { (void)f; (void)h; absent("BcfReader"); return NULL; }
void libBigWigInit(int n) { (void)n; }
