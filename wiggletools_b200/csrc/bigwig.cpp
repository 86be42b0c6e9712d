// wiggletools_b200 -- host BigWig batch decoder: file -> pinned SoA tracks.
//
// SURVEY.md section 8(f) next-1.  The reference reads BigWig through libBigWig
// (not vendored, not installed): a producer thread per file queries 10 000-bp
// stretches and pushes one record at a time through a locked buffer
// (src/bigWiggleReader.c:36-108, src/bufferedReader.c).  Here a whole chromosome
// is decoded in one call: the R-tree index is walked once for the chromosome's
// data blocks, the blocks are inflated (zlib) and converted in parallel by a
// small thread pool, and the intervals land in pinned int32/int32/fp64 arrays
// ready for cudaMemcpyAsync.
//
// Record semantics reproduced from the reference reader:
//  - coordinates become 1-based half-open (bigWiggleReader.c:39-40);
//  - the reference queries [1+10000k, 1+10000(k+1)) stretches and boxes every
//    returned interval into the stretch (:43-44, :73-83), so an interval crossing
//    a stretch boundary is emitted as several equal-valued records, and nothing
//    is read at or beyond the last stretch whose start is < chromosome length
//    (SURVEY quirk Q9).  `stretch` = 10000 reproduces that; 0 disables it.
//
// File format (UCSC bigWig v4, little endian): 64-byte header, chromosome B+
// tree, zlib-compressed data sections (bedGraph / variableStep / fixedStep) and
// an R-tree index over them.
#include "../../include/wiggletools_b200.h"

#include <cuda_runtime.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace {

std::string g_bw_error;

template <class T> T rd(const unsigned char *p) { T v; memcpy(&v, p, sizeof(T)); return v; }

struct Block { uint64_t offset, size; };

struct Chrom { std::string name; uint32_t id, length; };

} // namespace

struct wb200_bigwig {
	std::vector<unsigned char> file;   // whole file in memory (inputs are local files; sizes are a few GB at most)
	uint32_t uncompress_buf = 0;
	uint64_t index_offset = 0;
	std::vector<Chrom> chroms;         // strcmp order of the names (the reference's visiting order, bigWiggleReader.c:97)
	int stretch = 10000;
	int threads = 0;
	// pinned result of the last decode
	void *p_start = nullptr, *p_end = nullptr, *p_value = nullptr;
	size_t cap = 0;
	bool pinned = true;
	std::string err;
};

namespace {

bool fail(wb200_bigwig *bw, const std::string &m) { if (bw) bw->err = m; g_bw_error = m; return false; }

// Every offset, count and size in the file is untrusted: reads are bounds-checked against the
// file size, recursion depth and node counts are capped, and anything out of range makes the walk
// fail ("corrupt BigWig index") instead of reading past the buffer.
constexpr int MAX_TREE_DEPTH = 32;
constexpr size_t MAX_CHROMS = 1u << 24, MAX_BLOCKS = 1u << 28;

inline bool in_file(const wb200_bigwig *bw, uint64_t off, uint64_t len) {
	return off <= bw->file.size() && len <= bw->file.size() - off;
}

bool walk_chrom_tree(const wb200_bigwig *bw, uint64_t node, uint32_t key_size, std::vector<Chrom> &out, int depth = 0) {
	if (depth > MAX_TREE_DEPTH || !in_file(bw, node, 4)) return false;
	const unsigned char *p = bw->file.data() + node;
	const bool leaf = p[0] != 0;
	const uint16_t count = rd<uint16_t>(p + 2);
	if (!in_file(bw, node + 4, (uint64_t) count * (key_size + 8))) return false;
	p += 4;
	for (uint16_t k = 0; k < count; k++) {
		if (leaf) {
			if (out.size() >= MAX_CHROMS) return false;
			Chrom c;
			c.name.assign((const char *) p, strnlen((const char *) p, key_size));
			c.id = rd<uint32_t>(p + key_size);
			c.length = rd<uint32_t>(p + key_size + 4);
			out.push_back(c);
			p += key_size + 8;
		} else {
			if (!walk_chrom_tree(bw, rd<uint64_t>(p + key_size), key_size, out, depth + 1)) return false;
			p += key_size + 8;
		}
	}
	return true;
}

// data blocks whose chromosome range includes `id`
bool walk_rtree(const wb200_bigwig *bw, uint64_t node, uint32_t id, std::vector<Block> &out, int depth = 0) {
	if (depth > MAX_TREE_DEPTH || !in_file(bw, node, 4)) return false;
	const unsigned char *p = bw->file.data() + node;
	const bool leaf = p[0] != 0;
	const uint16_t count = rd<uint16_t>(p + 2);
	if (!in_file(bw, node + 4, (uint64_t) count * (leaf ? 32 : 24))) return false;
	p += 4;
	for (uint16_t k = 0; k < count; k++) {
		const uint32_t c0 = rd<uint32_t>(p), c1 = rd<uint32_t>(p + 8);
		const bool hit = c0 <= id && id <= c1;
		if (leaf) {
			if (hit) {
				const Block b{rd<uint64_t>(p + 16), rd<uint64_t>(p + 24)};
				if (!in_file(bw, b.offset, b.size) || out.size() >= MAX_BLOCKS) return false;
				out.push_back(b);
			}
			p += 32;
		} else {
			if (hit && !walk_rtree(bw, rd<uint64_t>(p + 16), id, out, depth + 1)) return false;
			p += 24;
		}
	}
	return true;
}

struct Raw { std::vector<unsigned char> data; int64_t pieces = 0; bool ok = true; };

// number of records the reference would emit for the 1-based interval [s,e)
inline int64_t pieces_of(int64_t s, int64_t e, int stretch, int64_t length) {
	if (stretch <= 0) return 1;
	// stretches start at 1 + k*stretch for k >= 0 while start < length
	const int64_t kmax = (length - 2) / stretch;          // last queried stretch (length >= 2), -1 if none
	if (length < 2) return 0;
	int64_t k0 = (s - 1) / stretch, k1 = (e - 2) / stretch;
	if (k1 > kmax) k1 = kmax;
	return k1 >= k0 ? k1 - k0 + 1 : 0;
}

// items of one data section (1 bedGraph, 2 variableStep, 3 fixedStep); false for a section that is
// truncated or of an unknown type
template <class F>
bool for_items(const unsigned char *d, size_t n, uint32_t id, F &&f) {
	if (n < 24) return false;
	if (rd<uint32_t>(d) != id) return true;           // another chromosome's section: nothing to emit
	const uint32_t cstart = rd<uint32_t>(d + 4), step = rd<uint32_t>(d + 12), span = rd<uint32_t>(d + 16);
	const uint8_t type = d[20];
	const uint16_t count = rd<uint16_t>(d + 22);
	if (type < 1 || type > 3) return false;
	if ((size_t) count * (type == 1 ? 12 : type == 2 ? 8 : 4) > n - 24) return false;
	const unsigned char *p = d + 24;
	for (uint16_t k = 0; k < count; k++) {
		int64_t s, e; float v;
		if (type == 1) { if (p + 12 > d + n) break; s = rd<uint32_t>(p); e = rd<uint32_t>(p + 4); v = rd<float>(p + 8); p += 12; }
		else if (type == 2) { if (p + 8 > d + n) break; s = rd<uint32_t>(p); e = s + span; v = rd<float>(p + 4); p += 8; }
		else { if (p + 4 > d + n) break; s = (int64_t) cstart + (int64_t) k * step; e = s + span; v = rd<float>(p); p += 4; }
		f(s + 1, e + 1, v);      // 0-based -> 1-based, bigWiggleReader.c:39-40
	}
	return true;
}

template <class F>
void parallel_for(int n, int threads, F &&f) {
	if (threads <= 1 || n < 4) { for (int i = 0; i < n; i++) f(i); return; }
	std::atomic<int> next(0);
	std::vector<std::thread> pool;
	const int nt = std::min(threads, n);
	for (int t = 0; t < nt; t++)
		pool.emplace_back([&]() { for (int i = next++; i < n; i = next++) f(i); });
	for (auto &t : pool) t.join();
}

} // namespace

extern "C" {

const char *wb200_bw_last_error(const wb200_bigwig *bw) { return bw ? bw->err.c_str() : g_bw_error.c_str(); }

wb200_bigwig *wb200_bw_open(const char *path) {
	FILE *f = fopen(path, "rb");
	if (!f) { fail(nullptr, std::string("cannot open ") + path); return nullptr; }
	wb200_bigwig *bw = new wb200_bigwig();
	fseek(f, 0, SEEK_END);
	const long size = ftell(f);
	fseek(f, 0, SEEK_SET);
	bw->file.resize(size > 0 ? size : 0);
	if (size > 0 && fread(bw->file.data(), 1, size, f) != (size_t) size) { fclose(f); delete bw; fail(nullptr, "short read"); return nullptr; }
	fclose(f);
	if (size < 64 || rd<uint32_t>(bw->file.data()) != 0x888FFC26u) {
		delete bw;
		fail(nullptr, std::string("File ") + path + " is not in BigWig format");
		return nullptr;
	}
	const unsigned char *h = bw->file.data();
	const uint64_t chrom_tree = rd<uint64_t>(h + 8);
	bw->index_offset = rd<uint64_t>(h + 24);
	bw->uncompress_buf = rd<uint32_t>(h + 52);
	// header offsets are untrusted: B+ tree header (32 bytes) and R-tree header (48 bytes) must lie inside the file
	if (!in_file(bw, chrom_tree, 32) || !in_file(bw, bw->index_offset, 48) ||
	    rd<uint32_t>(h + chrom_tree) != 0x78CA8C91u || rd<uint32_t>(h + bw->index_offset) != 0x2468ACE0u) {
		delete bw;
		fail(nullptr, "corrupt BigWig index");
		return nullptr;
	}
	const uint32_t key_size = rd<uint32_t>(h + chrom_tree + 8);
	if (key_size == 0 || key_size > 4096 || bw->uncompress_buf > (1u << 30) ||
	    !walk_chrom_tree(bw, chrom_tree + 32, key_size, bw->chroms)) {
		delete bw;
		fail(nullptr, "corrupt BigWig index");
		return nullptr;
	}
	std::sort(bw->chroms.begin(), bw->chroms.end(), [](const Chrom &a, const Chrom &b) { return strcmp(a.name.c_str(), b.name.c_str()) < 0; });
	const char *env = getenv("WB200_BW_THREADS");
	bw->threads = env ? atoi(env) : (int) std::min(16u, std::max(1u, std::thread::hardware_concurrency()));
	return bw;
}

static void bw_free(wb200_bigwig *bw) {
	void *p[3] = {bw->p_start, bw->p_end, bw->p_value};
	for (void *q : p) if (q) { if (bw->pinned) cudaFreeHost(q); else free(q); }
	bw->p_start = bw->p_end = bw->p_value = nullptr;
	bw->cap = 0;
}

void wb200_bw_close(wb200_bigwig *bw) {
	if (!bw) return;
	bw_free(bw);
	delete bw;
}

int wb200_bw_nchrom(const wb200_bigwig *bw) { return (int) bw->chroms.size(); }
const char *wb200_bw_chrom_name(const wb200_bigwig *bw, int i) { return bw->chroms[i].name.c_str(); }
int64_t wb200_bw_chrom_length(const wb200_bigwig *bw, int i) { return bw->chroms[i].length; }
void wb200_bw_set_stretch(wb200_bigwig *bw, int stretch_bp) { bw->stretch = stretch_bp; }

int wb200_bw_decode_chrom(wb200_bigwig *bw, int chrom_index, wb200_track *out) {
	if (!bw || chrom_index < 0 || chrom_index >= (int) bw->chroms.size() || !out) return WB200_EINVAL;
	const Chrom &c = bw->chroms[chrom_index];
	std::vector<Block> blocks;
	if (!walk_rtree(bw, bw->index_offset + 48, c.id, blocks)) { fail(bw, "corrupt BigWig index"); return WB200_EINVAL; }
	std::sort(blocks.begin(), blocks.end(), [](const Block &a, const Block &b) { return a.offset < b.offset; });
	const int nb = (int) blocks.size();
	std::vector<Raw> raw(nb);
	const int stretch = bw->stretch;
	const int64_t length = c.length;
	// inflate + count, in parallel
	parallel_for(nb, bw->threads, [&](int b) {
		Raw &r = raw[b];
		const unsigned char *src = bw->file.data() + blocks[b].offset;
		if (blocks[b].offset + blocks[b].size > bw->file.size()) { r.ok = false; return; }
		if (bw->uncompress_buf) {
			r.data.resize(bw->uncompress_buf);
			uLongf n = bw->uncompress_buf;
			if (uncompress(r.data.data(), &n, src, (uLong) blocks[b].size) != Z_OK) { r.ok = false; return; }
			r.data.resize(n);
		} else
			r.data.assign(src, src + blocks[b].size);
		int64_t cnt = 0;
		if (!for_items(r.data.data(), r.data.size(), c.id, [&](int64_t s, int64_t e, float) { cnt += pieces_of(s, e, stretch, length); })) { r.ok = false; return; }
		r.pieces = cnt;
	});
	std::vector<int64_t> off(nb + 1, 0);
	for (int b = 0; b < nb; b++) {
		if (!raw[b].ok) { fail(bw, "corrupt BigWig data block"); return WB200_EINVAL; }
		off[b + 1] = off[b] + raw[b].pieces;
	}
	const int64_t total = off[nb];
	if ((size_t) total > bw->cap) {
		bw_free(bw);
		const size_t want = (size_t) total + total / 8 + 1024;
		bw->pinned = cudaHostAlloc(&bw->p_start, want * 4, cudaHostAllocDefault) == cudaSuccess &&
		             cudaHostAlloc(&bw->p_end, want * 4, cudaHostAllocDefault) == cudaSuccess &&
		             cudaHostAlloc(&bw->p_value, want * 8, cudaHostAllocDefault) == cudaSuccess;
		if (!bw->pinned) {
			// no CUDA driver (decode-only use): the buffers are plain memory; staging them
			// to a GPU then goes through the driver's pageable path
			cudaGetLastError();
			if (bw->p_start) cudaFreeHost(bw->p_start);
			if (bw->p_end) cudaFreeHost(bw->p_end);
			bw->p_start = malloc(want * 4); bw->p_end = malloc(want * 4); bw->p_value = malloc(want * 8);
			if (!bw->p_start || !bw->p_end || !bw->p_value) { fail(bw, "out of host memory"); return WB200_ENOMEM; }
		}
		bw->cap = want;
	}
	int32_t *S = (int32_t *) bw->p_start, *E = (int32_t *) bw->p_end;
	double *V = (double *) bw->p_value;
	parallel_for(nb, bw->threads, [&](int b) {
		int64_t k = off[b];
		for_items(raw[b].data.data(), raw[b].data.size(), c.id, [&](int64_t s, int64_t e, float v) {
			if (stretch <= 0) { S[k] = (int32_t) s; E[k] = (int32_t) e; V[k] = v; k++; return; }
			if (length < 2) return;
			const int64_t kmax = (length - 2) / stretch;
			int64_t k0 = (s - 1) / stretch, k1 = std::min<int64_t>((e - 2) / stretch, kmax);
			for (int64_t q = k0; q <= k1; q++) {
				const int64_t lo = 1 + q * stretch, hi = lo + stretch;     // boxed into the stretch, :43-44
				S[k] = (int32_t) std::max(s, lo); E[k] = (int32_t) std::min(e, hi); V[k] = v; k++;
			}
		});
	});
	out->start = S; out->end = E; out->value = V; out->n = total; out->default_value = 0; out->value32 = nullptr;
	return WB200_OK;
}

} // extern "C"
