// wiggletools_b200 -- the reference's operator API for the merge+reduce path,
// implemented on top of the batch engine (see include/wiggletools_dropin.h).
//
// The reference is a lazy pull graph: newMultiplexer() + MeanReduction() build
// iterators and every pop() advances a heap sweep by one segment
// (multiplexer.c:98-128, reducers.c:367-402).  Here a reduction is BATCH per
// chromosome: on demand the input iterators are drained for the next
// chromosome (strcmp order, multiplexer.c:50-74) into SoA arrays, one
// wb200_reduce_host() call computes all of the chromosome's segments on the
// GPU, and pop() then walks the result buffer.
#include "../../include/wiggletools_dropin.h"
#include "../../include/wiggletools_b200.h"

#include <chrono>
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <set>
#include <string>
#include <vector>

// layouts the reference's callers rely on
static_assert(offsetof(wiggleIterator_st, value) == 16 && offsetof(wiggleIterator_st, done) == 32 &&
              offsetof(wiggleIterator_st, data) == 40 && offsetof(wiggleIterator_st, pop) == 48 &&
              offsetof(wiggleIterator_st, overlaps) == 64 && offsetof(wiggleIterator_st, default_value) == 72 &&
              offsetof(wiggleIterator_st, append) == 80 && sizeof(wiggleIterator_st) == 88, "WiggleIterator layout");
static_assert(offsetof(multiplexer_st, values) == 16 && offsetof(multiplexer_st, count) == 32 &&
              offsetof(multiplexer_st, inplay) == 40 && offsetof(multiplexer_st, iters) == 48 &&
              offsetof(multiplexer_st, done) == 56 && offsetof(multiplexer_st, strict) == 57 &&
              offsetof(multiplexer_st, pop) == 64 && offsetof(multiplexer_st, data) == 96 &&
              sizeof(multiplexer_st) == 104, "Multiplexer layout");
static_assert(offsetof(multiset_st, values) == 16 && offsetof(multiset_st, count) == 24 &&
              offsetof(multiset_st, inplay) == 32 && offsetof(multiset_st, multis) == 40 &&
              offsetof(multiset_st, done) == 48 && sizeof(multiset_st) == 72, "Multiset layout");

namespace {

wb200_ctx *g_ctx = nullptr;

wb200_ctx *engine() {
	if (!g_ctx) {
		const char *dev = getenv("WB200_DEVICE");
		g_ctx = wb200_create(dev ? atoi(dev) : 0);
		if (!g_ctx) {
			fprintf(stderr, "%s\n", wb200_last_error(nullptr));
			exit(1);
		}
	}
	return g_ctx;
}

// one stable char* per chromosome name for the life of the process
char *intern(const char *name) {
	static std::set<std::string> names;
	return const_cast<char *>(names.insert(name).first->c_str());
}

// WB200_PROFILE=1: wall time spent decoding BigWig sections, in engine calls (staging + kernels + read-back)
// and writing text, printed at exit -- where an end-to-end run from files goes
struct Prof {
	double decode = 0, engine = 0, write = 0;
	bool on = false;
	std::chrono::steady_clock::time_point t0;
	Prof() { on = getenv("WB200_PROFILE") != nullptr; t0 = std::chrono::steady_clock::now(); }
	~Prof() {
		if (!on) return;
		const double total = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
		fprintf(stderr, "wiggletools_b200 profile: total %.3f s, BigWig decode %.3f s, engine calls %.3f s, text write %.3f s\n", total, decode, engine, write);
	}
} g_prof;
struct ProfScope {
	double &acc;
	std::chrono::steady_clock::time_point t;
	explicit ProfScope(double &a) : acc(a), t(std::chrono::steady_clock::now()) {}
	~ProfScope() { acc += std::chrono::duration<double>(std::chrono::steady_clock::now() - t).count(); }
};

void die(const char *msg) {
	fprintf(stderr, "wiggletools_b200: %s\n", msg);
	exit(1);
}

void gpu_multiplexer_pop(Multiplexer *);
void gpu_multiplexer_seek(Multiplexer *, const char *, int, int);

bool is_gpu_multiplexer(const Multiplexer *m) { return m->pop == &gpu_multiplexer_pop; }

// ---------------------------------------------------------------------------
// BigWig source: the library's batch decoder behind the WiggleIterator protocol
// (stands where src/bigWiggleReader.c goes through libBigWig)
// ---------------------------------------------------------------------------
struct BwSource {
	wb200_bigwig *bw = nullptr;
	int ci = -1;                 // chromosome being served (index in strcmp order)
	wb200_track t{};             // decoded chromosome (arrays owned by the handle)
	int64_t k = 0;               // next record to serve
	char *chrom = nullptr;
	bool hold = false;           // holdFire: nothing is read before the first seek (commandParser.c:615-624)
	// region restriction after seek(): one boxed query, no 10 000-bp stretches (bigWiggleReader.c:52-71,125-145)
	bool region = false;
	std::vector<int32_t> rs, re;
	std::vector<double> rv;

	bool load_next_chrom() {
		if (region) return false;
		while (++ci < wb200_bw_nchrom(bw)) {
			{ ProfScope ps(g_prof.decode); if (wb200_bw_decode_chrom(bw, ci, &t) != WB200_OK) die(wb200_bw_last_error(bw)); }
			if (t.n > 0) { chrom = intern(wb200_bw_chrom_name(bw, ci)); k = 0; return true; }
		}
		t.n = 0; k = 0;
		return false;
	}
};

void bw_pop(WiggleIterator *wi) {
	BwSource *b = (BwSource *) wi->data;
	if (b->hold) { wi->done = true; return; }
	while (b->k >= b->t.n)
		if (!b->load_next_chrom()) { wi->done = true; return; }
	wi->chrom = b->chrom;
	wi->start = b->t.start[b->k];
	wi->finish = b->t.end[b->k];
	wi->value = b->t.value[b->k];
	b->k++;
}

void bw_seek(WiggleIterator *wi, const char *chrom, int start, int finish) {
	BwSource *b = (BwSource *) wi->data;
	b->hold = false;
	b->region = true;
	b->rs.clear(); b->re.clear(); b->rv.clear();
	b->t = wb200_track{nullptr, nullptr, nullptr, 0, 0};
	b->k = 0;
	if (start < 1) start = 1;
	if (finish < 1) finish = 1;
	for (int i = 0; i < wb200_bw_nchrom(b->bw); i++) {
		if (strcmp(wb200_bw_chrom_name(b->bw, i), chrom) != 0) continue;
		wb200_bw_set_stretch(b->bw, 0);
		wb200_track full;
		if (wb200_bw_decode_chrom(b->bw, i, &full) != WB200_OK) die(wb200_bw_last_error(b->bw));
		wb200_bw_set_stretch(b->bw, 10000);
		for (int64_t q = 0; q < full.n; q++) {
			if (full.start[q] >= finish || full.end[q] <= start) continue;      // overlaps [start, finish)
			b->rs.push_back(std::max(full.start[q], (int32_t) start));          // boxed into the query
			b->re.push_back(std::min(full.end[q], (int32_t) finish));
			b->rv.push_back(full.value[q]);
		}
		b->chrom = intern(chrom);
		b->t = wb200_track{b->rs.data(), b->re.data(), b->rv.data(), (int64_t) b->rs.size(), 0};
		b->ci = i;
	}
	wi->done = false;
	bw_pop(wi);
}

// ---------------------------------------------------------------------------
// element-wise unary operators (unaryOps.c: scale, offset, ln, log, exp, pow, abs,
// default) behind the WiggleIterator protocol.  Popped on the host they behave as the
// reference's own iterators do; drained by a reduction of this library they are seen
// through: the raw records of the source underneath go to the device with the chain of
// maps (wb200_set_input_maps), so `mean map ln a.bw b.bw` maps on the GPU.
// ---------------------------------------------------------------------------
struct RawRecord { char *chrom; int start, finish; double value; bool valid; };

struct MappedSource {
	WiggleIterator *inner = nullptr;
	int op = 0;                 // WB200_MAP_* or 0: records untouched (`default`)
	double arg = 0;             // as the reference's constructor stores it (log base -> log(base), ln / exp -> 1)
	RawRecord raw{nullptr, 0, 0, 0, false};   // the SOURCE's record behind the one currently presented
};

void mapped_pop(WiggleIterator *wi);
void mapped_seek(WiggleIterator *wi, const char *chrom, int start, int finish);
bool is_mapped(const WiggleIterator *wi) { return wi->pop == &mapped_pop; }

RawRecord raw_of(WiggleIterator *it) {
	if (is_mapped(it)) return ((MappedSource *) it->data)->raw;
	return RawRecord{it->chrom, it->start, it->finish, it->value, !it->done};
}

void mapped_pop(WiggleIterator *wi) {
	MappedSource *m = (MappedSource *) wi->data;
	WiggleIterator *iter = m->inner;
	if (m->op == WB200_MAP_LOG)
		while (!iter->done && iter->value <= 0) pop(iter);        // LogWiggleIteratorPop, unaryOps.c:764-765
	if (iter->done) { wi->done = true; m->raw.valid = false; return; }
	wi->chrom = iter->chrom;
	wi->start = iter->start;
	wi->finish = iter->finish;
	const double v = iter->value, a = m->arg;
	switch (m->op) {
	case WB200_MAP_SCALE: wi->value = std::isnan(v) ? NAN : a * v; break;                                  // :650-665
	case WB200_MAP_OFFSET: wi->value = a + v; break;                                                       // :722-734
	case WB200_MAP_LOG: wi->value = (std::isnan(v) || v < 0) ? NAN : std::log(v) / a; break;               // :760-779
	case WB200_MAP_EXP: wi->value = std::exp(v * a); break;                                                // :823-835
	case WB200_MAP_POW: wi->value = ((a < 0 && v <= 0) || std::isnan(v)) ? NAN : std::pow(v, a); break;    // :873-889
	case WB200_MAP_ABS: wi->value = std::isnan(v) ? NAN : std::fabs(v); break;                             // :934-949
	default: wi->value = v; break;                                                                         // :210-223
	}
	m->raw = raw_of(iter);
	pop(iter);
}

void mapped_seek(WiggleIterator *wi, const char *chrom, int start, int finish) {
	MappedSource *m = (MappedSource *) wi->data;
	seek(m->inner, chrom, start, finish);
	wi->done = false;
	pop(wi);
}

// The chain under `it` down to the first iterator that is not a map (innermost map first), and
// the source's records that the wrappers have already consumed, in stream order.
WiggleIterator *see_through(WiggleIterator *it, std::vector<wb200_map> &chain, std::vector<RawRecord> &pending) {
	std::vector<MappedSource *> levels;
	while (is_mapped(it)) { levels.push_back((MappedSource *) it->data); it = ((MappedSource *) it->data)->inner; }
	for (size_t k = levels.size(); k-- > 0; )                    // innermost first
		if (levels[k]->op) chain.push_back(wb200_map{levels[k]->op, 0, levels[k]->arg});
	for (MappedSource *m : levels)                               // outermost wrapper presents the earliest record
		if (m->raw.valid) pending.push_back(m->raw);
	return it;
}

// ---------------------------------------------------------------------------
// batch reduction behind a WiggleIterator
// ---------------------------------------------------------------------------
struct Reducer {
	int op = 0;
	int n1 = 0;
	bool strict = false;
	bool pearson = false;
	int nd_rank = 0;           // > 0: N-dimensional Pearson over nd_rank track pairs (px = d, py = n1 + d)
	double nd_moments[2][6] = {{0, 0, 0, 0, 0, 0}, {0, 0, 0, 0, 0, 0}};
	std::vector<WiggleIterator *> iters;
	// staged inputs of the current chromosome
	std::vector<std::vector<int32_t>> s, e;
	std::vector<std::vector<double>> v;
	std::vector<std::vector<wb200_map>> chains;   // per input: the maps applied on the device (empty: none)
	// results of the current chromosome
	std::vector<int32_t> rs, re;
	std::vector<double> rv;
	size_t k = 0;
	char *chrom = nullptr;
	double auc = 0;            // running whole-genome AUC (device partials, chromosome order)
	double moments[6] = {0, 0, 0, 0, 0, 0};

	// Drain every input positioned on the strcmp-smallest chromosome
	// (multiplexer.c:50-74) into SoA arrays.  false = inputs exhausted.
	bool drain(std::vector<wb200_track> &tracks, std::vector<BwSource *> &direct) {
		const size_t n = iters.size();
		const char *best = nullptr;
		for (WiggleIterator *it : iters)
			if (!it->done && (!best || strcmp(it->chrom, best) < 0)) best = it->chrom;
		if (!best) return false;
		chrom = intern(best);
		s.resize(n); e.resize(n); v.resize(n);
		tracks.assign(n, wb200_track{});
		direct.assign(n, nullptr);
		chains.assign(n, std::vector<wb200_map>());
		for (size_t i = 0; i < n; i++) {
			WiggleIterator *it = iters[i];
			s[i].clear(); e[i].clear(); v[i].clear();
			// (the Pearson entry points stage their tracks themselves: maps stay on the host there)
			if (!pearson && is_mapped(it) && !it->done && strcmp(it->chrom, chrom) == 0 && drain_mapped(i, tracks[i])) continue;
			// A BigWig source positioned on its chromosome's first record hands over the
			// decoded pinned arrays as they are: no per-record pop, no copy.
			if (it->pop == &bw_pop && !it->done) {
				BwSource *b = (BwSource *) it->data;
				if (b->k == 1 && b->chrom == chrom) {
					tracks[i].start = b->t.start; tracks[i].end = b->t.end; tracks[i].value = b->t.value;
					tracks[i].n = b->t.n;
					tracks[i].default_value = it->default_value;
					direct[i] = b;
					continue;
				}
			}
			while (!it->done && strcmp(it->chrom, chrom) == 0) {
				s[i].push_back(it->start);
				e[i].push_back(it->finish);
				v[i].push_back(it->value);
				pop(it);
			}
			tracks[i].start = s[i].data(); tracks[i].end = e[i].data(); tracks[i].value = v[i].data();
			tracks[i].n = (int64_t) s[i].size();
			tracks[i].default_value = it->default_value;
		}
		return true;
	}

	// Input i is a chain of element-wise maps: its source's raw records of this chromosome are
	// staged and the chain goes to the device.  The wrappers are then re-primed on the source's
	// next record (the first of the next chromosome), as if they had been popped all along.
	bool drain_mapped(size_t i, wb200_track &track) {
		std::vector<RawRecord> pending;
		WiggleIterator *base = see_through(iters[i], chains[i], pending);
		if (chains[i].size() > 8) { chains[i].clear(); return false; }
		for (const RawRecord &r : pending)
			if (r.chrom == chrom || strcmp(r.chrom, chrom) == 0) { s[i].push_back(r.start); e[i].push_back(r.finish); v[i].push_back(r.value); }
			else { chains[i].clear(); s[i].clear(); e[i].clear(); v[i].clear(); return false; }   // wrappers straddle a chromosome boundary: host path
		while (!base->done && strcmp(base->chrom, chrom) == 0) {
			s[i].push_back(base->start);
			e[i].push_back(base->finish);
			v[i].push_back(base->value);
			pop(base);
		}
		// re-prime innermost first: each wrapper presents the next record of its (now advanced) input
		std::vector<WiggleIterator *> levels;
		for (WiggleIterator *it = iters[i]; is_mapped(it); it = ((MappedSource *) it->data)->inner) levels.push_back(it);
		for (size_t k = levels.size(); k-- > 0; ) { levels[k]->done = false; ((MappedSource *) levels[k]->data)->raw.valid = false; levels[k]->pop(levels[k]); }
		track.start = s[i].data(); track.end = e[i].data(); track.value = v[i].data();
		track.n = (int64_t) s[i].size();
		track.default_value = iters[i]->default_value;
		return true;
	}

	void arm_maps() {
		bool any = false;
		for (auto &c : chains) any = any || !c.empty();
		if (!any) return;
		std::vector<const wb200_map *> ptrs(chains.size());
		std::vector<int> counts(chains.size());
		for (size_t i = 0; i < chains.size(); i++) { ptrs[i] = chains[i].data(); counts[i] = (int) chains[i].size(); }
		if (wb200_set_input_maps(engine(), (int) chains.size(), ptrs.data(), counts.data()) != WB200_OK) die(wb200_last_error(engine()));
		if (getenv("WB200_VERBOSE")) {
			int mapped = 0;
			for (auto &c : chains) mapped += c.empty() ? 0 : 1;
			fprintf(stderr, "wiggletools_b200: %s: maps of %d of %d inputs run on the device\n", chrom, mapped, (int) chains.size());
		}
	}

	// the decoded arrays have been staged: move the direct sources to their next chromosome
	void release(std::vector<BwSource *> &direct) {
		for (size_t i = 0; i < direct.size(); i++)
			if (direct[i]) { direct[i]->k = direct[i]->t.n; pop(iters[i]); }
	}

	// Next chromosome straight into the text sink: reduction and formatting on the device,
	// only the text comes back (wb200_reduce_format_host).  false = inputs exhausted.
	bool next_chromosome_to_sink(FILE *out, int mode, wb200_sink_state *state) {
		if (!stage()) return false;
		std::vector<wb200_track> &tracks = st_tracks;
		const char *text = nullptr;
		int64_t nbytes = 0, nseg = 0;
		double a = 0;
		arm_maps();
		{
			ProfScope ps(g_prof.engine);
			if (wb200_reduce_format_host(engine(), op, (int) tracks.size(), n1, strict ? 1 : 0, tracks.data(), mode, chrom, state,
			                             &text, &nbytes, &nseg, &a) != WB200_OK)
				die(wb200_last_error(engine()));
		}
		{ ProfScope ps(g_prof.write); if (nbytes > 0 && fwrite(text, 1, (size_t) nbytes, out) != (size_t) nbytes) die("short write to the output file"); }
		auc += a;
		unstage();
		return true;
	}

	// A chromosome drained ahead of its reduction (a strict Multiplexer shows its first row at
	// construction, like the reference's eager pop): the staged inputs wait here until a reduction or
	// the row server computes on them.
	bool staged = false;
	std::vector<wb200_track> st_tracks;
	std::vector<BwSource *> st_direct;

	bool stage() {
		if (staged) return true;
		staged = drain(st_tracks, st_direct);
		return staged;
	}

	// the statistic of this Reducer on the staged chromosome (the inputs stay staged)
	void compute() {
		const size_t n = iters.size();
		rs.clear(); re.clear(); rv.clear(); k = 0;
		wb200_segments out;
		arm_maps();
		{
			ProfScope ps(g_prof.engine);
			if (wb200_reduce_host(engine(), op, (int) n, n1, strict ? 1 : 0, st_tracks.data(), &out) != WB200_OK)
				die(wb200_last_error(engine()));
		}
		rs.assign(out.start, out.start + out.n);
		re.assign(out.end, out.end + out.n);
		rv.assign(out.value, out.value + out.n);
		last_auc = out.auc;
	}
	double last_auc = 0;

	void unstage() {
		if (!staged) return;
		release(st_direct);
		staged = false;
	}

	// Same as drain(), then run the engine on the chromosome.
	bool next_chromosome() {
		if (!stage()) return false;
		std::vector<wb200_track> &tracks = st_tracks;
		const size_t n = iters.size();
		rs.clear(); re.clear(); rv.clear(); k = 0;
		if (pearson && nd_rank > 0) {
			for (int d = 0; d < nd_rank; d++) {
				double m[6];
				if (wb200_pearson_pair_host(engine(), (int) n, tracks.data(), d, n1 + d, m) != WB200_OK) die(wb200_last_error(engine()));
				wb200_pearson_merge(nd_moments[d], m);
			}
			unstage();
			return true;
		}
		if (pearson) {
			double m[6];
			if (wb200_pearson_host(engine(), &tracks[0], &tracks[1], m) != WB200_OK) die(wb200_last_error(engine()));
			wb200_pearson_merge(moments, m);
			unstage();
			return true;
		}
		compute();
		auc += last_auc;
		unstage();
		return true;
	}
};

void reducer_pop(WiggleIterator *wi) {
	Reducer *r = (Reducer *) wi->data;
	while (r->k >= r->rs.size())
		if (!r->next_chromosome()) { wi->done = true; return; }
	wi->chrom = r->chrom;
	wi->start = r->rs[r->k];
	wi->finish = r->re[r->k];
	wi->value = r->rv[r->k];
	r->k++;
}

void reducer_seek(WiggleIterator *wi, const char *chrom, int start, int finish) {
	// WiggleReducerSeek (reducers.c:25-29): seek every input, then pop
	Reducer *r = (Reducer *) wi->data;
	r->staged = false;
	for (WiggleIterator *it : r->iters) seek(it, chrom, start, finish);
	r->rs.clear(); r->re.clear(); r->rv.clear(); r->k = 0;
	r->auc = 0;       // the integral restarts with the sought region (the eager first pop had added the whole first chromosome)
	pop(wi);
}

double host_default(int op, const std::vector<WiggleIterator *> &iters) {
	// the constructors' derived default_value: ask the engine (empty job, same
	// code path as the real one, no kernels needed for zero intervals)
	std::vector<wb200_track> tr(iters.size());
	for (size_t i = 0; i < iters.size(); i++) { tr[i] = wb200_track{nullptr, nullptr, nullptr, 0, iters[i]->default_value}; }
	wb200_segments out;
	if (wb200_reduce_host(engine(), op, (int) tr.size(), (int) tr.size(), 0, tr.data(), &out) != WB200_OK)
		die(wb200_last_error(engine()));
	return out.default_value;
}

struct MuxState;
MuxState *mux_state(Multiplexer *multi);
// the inputs of a GPU multiplexer pass to a reduction / integrator (with the chromosome a strict
// multiplexer staged for its first row, if any); the multiplexer itself is spent
void mux_claim(Multiplexer *multi, Reducer *r);

WiggleIterator *make_reduction(int op, Multiplexer *multi) {
	if (!is_gpu_multiplexer(multi))
		die("reducers run on multiplexers built by newMultiplexer(); a custom multiplexer (apply/mwrite) as reducer input is not supported");
	Reducer *r = new Reducer();
	r->op = op;
	r->strict = multi->strict;
	r->n1 = multi->count;
	mux_claim(multi, r);
	double dflt = host_default(op, r->iters);
	return newWiggleIterator(r, &reducer_pop, &reducer_seek, dflt, false);
}

WiggleIterator *make_setcompare(int op, Multiset *ms) {
	Reducer *r = new Reducer();
	r->op = op;
	for (int m = 0; m < 2; m++) {
		if (!is_gpu_multiplexer(ms->multis[m])) die("set comparisons need multiplexers built by newMultiplexer()");
		Reducer part;
		mux_claim(ms->multis[m], &part);
		if (part.staged) die("set comparisons over strict multiplexers are not supported");   // (the reference's multiset sweeps its children non-strictly)
		r->iters.insert(r->iters.end(), part.iters.begin(), part.iters.end());
	}
	r->n1 = ms->multis[0]->count;
	return newWiggleIterator(r, &reducer_pop, &reducer_seek, NAN, false);
}

// ---------------------------------------------------------------------------
// A GPU-backed Multiplexer is first of all a description of the inputs that a reduction or an
// integrator takes over.  When somebody POPS it instead (mwrite, mwrite_bg, paste, select, any
// user of popMultiplexer) it serves rows: per chromosome the union segments come from one
// reduction on the device and the per-row values[] / inplay[] (multiplexer.c:76-96) from
// wb200_rows_host, fetched in blocks of 10 000 rows (the multi-writer's own block size,
// mWigWriter.c:29).  Like the reference's (multiplexer.c:167) it shows its first row on return
// from newMultiplexer: for a non-strict one that row follows from the inputs' first records and
// nothing is consumed; a strict one stages its first chromosome on the device (the first row
// where all inputs are in play can lie anywhere in it).
// ---------------------------------------------------------------------------
constexpr size_t ROW_BLOCK = 10000;

struct MuxState {
	Reducer r;                 // op = SUM: the segments of a sum are the multiplexer's rows
	bool claimed = false;      // a reduction / integrator took the inputs
	bool started = false;      // row mode: r holds the segments of the current chromosome
	bool eager = false;        // the row on display was derived from the inputs' first records
	long long pops = 0;        // rows the user has moved past
	size_t row = 0;            // index (in r.rs) of the row on display
	size_t blk0 = 0, blkn = 0; // rows [blk0, blk0 + blkn) are in vals / inp
	std::vector<double> vals;
	std::vector<unsigned char> inp;
};

MuxState *mux_state(Multiplexer *multi) { return (MuxState *) multi->data; }

void mux_claim(Multiplexer *multi, Reducer *r) {
	MuxState *ms = mux_state(multi);
	if (ms->claimed) die("this Multiplexer already feeds another reduction");
	if (ms->pops > 0) die("this Multiplexer has been popped; a reduction must be built on a fresh one");
	ms->claimed = true;
	r->iters = ms->r.iters;
	if (ms->r.staged) {        // the chromosome staged for the strict first row goes with the inputs
		r->staged = true;
		r->st_tracks.swap(ms->r.st_tracks); r->st_direct.swap(ms->r.st_direct);
		r->s.swap(ms->r.s); r->e.swap(ms->r.e); r->v.swap(ms->r.v); r->chains.swap(ms->r.chains);
		r->chrom = ms->r.chrom;
		ms->r.staged = false;
	}
}

// the row on display: fields of the Multiplexer, as popCoreMultiplexer2 leaves them
void mux_show(Multiplexer *m, MuxState *ms) {
	const size_t n = (size_t) m->count;
	if (ms->row < ms->blk0 || ms->row >= ms->blk0 + ms->blkn) {
		ms->blk0 = ms->row;
		ms->blkn = std::min(ROW_BLOCK, ms->r.rs.size() - ms->row);
		ms->vals.resize(ms->blkn * n); ms->inp.resize(ms->blkn * n);
		if (wb200_rows_host(engine(), (int64_t) ms->blk0, (int64_t) ms->blkn, ms->vals.data(), ms->inp.data()) != WB200_OK)
			die(wb200_last_error(engine()));    // (e.g. another reduction ran on the engine between two blocks of this multiplexer)
	}
	const size_t o = (ms->row - ms->blk0) * n;
	m->chrom = ms->r.chrom; m->start = ms->r.rs[ms->row]; m->finish = ms->r.re[ms->row];
	int count = 0;
	for (size_t i = 0; i < n; i++) {
		m->values[i] = ms->vals[o + i];
		m->inplay[i] = ms->inp[o + i];
		count += ms->inp[o + i] ? 1 : 0;
	}
	m->inplay_count = count;
}

// next chromosome of the row server: segments on the device, rows fetched on demand
bool mux_next_chromosome(MuxState *ms) {
	while (true) {
		if (!ms->r.stage()) return false;
		ms->r.compute();
		// (the staged inputs stay resident on the device for wb200_rows_host; the sources move on)
		ms->r.unstage();
		ms->row = 0; ms->blk0 = ms->blkn = 0;
		if (!ms->r.rs.empty()) return true;
	}
}

void gpu_multiplexer_pop(Multiplexer *m) {
	MuxState *ms = mux_state(m);
	if (ms->claimed) die("this Multiplexer feeds a reduction; it cannot be popped as well");
	ms->pops++;
	ms->r.unstage();           // a strict multiplexer's first chromosome: its sources move on now (the device keeps the staged copy)
	if (!ms->started) {
		// first pop of a non-strict multiplexer: the row on display came from the inputs' first records;
		// the chromosome is now reduced on the device and its row 0 must be that very row
		if (!mux_next_chromosome(ms)) { m->done = true; return; }
		ms->started = true;
		if (ms->eager && (ms->r.rs[0] != m->start || ms->r.re[0] != m->finish)) die("internal: first multiplexer row differs from the device's");
		ms->eager = false;
	}
	ms->row++;
	while (ms->row >= ms->r.rs.size())
		if (!mux_next_chromosome(ms)) { m->done = true; return; }
	mux_show(m, ms);
}

// the first row on display after construction / seek (multiplexer.c:157-169, 130-141)
void mux_first_row(Multiplexer *m) {
	MuxState *ms = mux_state(m);
	ms->started = false; ms->eager = false; ms->pops = 0; ms->row = 0; ms->blk0 = ms->blkn = 0;
	ms->r.staged = false;
	m->done = false;
	const int n = m->count;
	if (m->strict) {
		// all inputs in play: the first such row can lie anywhere in the chromosome -> stage it now;
		// a reduction built on this multiplexer takes the staged chromosome over (mux_claim)
		while (true) {
			if (!ms->r.stage()) { m->done = true; return; }
			ms->r.compute();
			if (!ms->r.rs.empty()) break;
			ms->r.unstage();
		}
		ms->started = true;
		ms->row = 0; ms->blk0 = ms->blkn = 0;
		mux_show(m, ms);
		return;
	}
	// non-strict: admit the inputs that start first on the smallest chromosome (multiplexer.c:50-96)
	const char *best = nullptr;
	for (int i = 0; i < n; i++) {
		WiggleIterator *it = m->iters[i];
		if (!it->done && (!best || strcmp(it->chrom, best) < 0)) best = it->chrom;
	}
	if (!best) { m->done = true; return; }
	int start = 0; bool any = false;
	for (int i = 0; i < n; i++) {
		WiggleIterator *it = m->iters[i];
		if (!it->done && strcmp(it->chrom, best) == 0 && (!any || it->start < start)) { start = it->start; any = true; }
	}
	int finish = 0; bool have = false; int count = 0;
	for (int i = 0; i < n; i++) {
		WiggleIterator *it = m->iters[i];
		const bool here = !it->done && strcmp(it->chrom, best) == 0;
		const bool in = here && it->start == start;
		m->inplay[i] = in;
		m->values[i] = in ? it->value : m->default_values[i];
		count += in ? 1 : 0;
		if (here) {
			const int edge = in ? it->finish : it->start;
			if (!have || edge < finish) { finish = edge; have = true; }
		}
	}
	m->chrom = intern(best); m->start = start; m->finish = finish; m->inplay_count = count;
	ms->eager = true;
}

void gpu_multiplexer_seek(Multiplexer *multi, const char *chrom, int start, int finish) {
	MuxState *ms = mux_state(multi);
	for (int i = 0; i < multi->count; i++) seek(multi->iters[i], chrom, start, finish);
	if (!ms->claimed) mux_first_row(multi);       // seekCoreMultiplexer pops once (multiplexer.c:140)
}

// ---------------------------------------------------------------------------
// generic pieces restated from the reference (not on the hot path)
// ---------------------------------------------------------------------------
struct UnaryData { WiggleIterator *iter; };

void union_pop(WiggleIterator *wi) {
	// merge overlapping records of one source (unaryOps.c:60-86)
	WiggleIterator *iter = ((UnaryData *) wi->data)->iter;
	if (iter->done) { wi->done = true; return; }
	int count = 0;
	while (!iter->done) {
		if (!count) {
			wi->chrom = iter->chrom; wi->start = iter->start; wi->finish = iter->finish; wi->value = iter->value;
		} else if (wi->chrom == iter->chrom && wi->finish > iter->start) {
			if (iter->finish > wi->finish) wi->finish = iter->finish;
		} else
			break;
		count++;
		pop(iter);
	}
	wi->done = (count == 0);
}

void unary_seek(WiggleIterator *wi, const char *chrom, int start, int finish) {
	seek(((UnaryData *) wi->data)->iter, chrom, start, finish);
	wi->done = false;
	pop(wi);
}

struct SelectData { Multiplexer *multi; int index; };

void select_pop(WiggleIterator *wi) {
	// SelectReductionPop (reducers.c:40-65), generic over any multiplexer
	if (wi->done) return;
	SelectData *d = (SelectData *) wi->data;
	Multiplexer *multi = d->multi;
	if (multi->done) { wi->done = true; return; }
	while (multi->inplay[d->index]) {
		popMultiplexer(multi);
		if (multi->done) { wi->done = true; return; }
	}
	wi->value = multi->values[d->index];
	wi->chrom = multi->chrom; wi->start = multi->start; wi->finish = multi->finish;
	popMultiplexer(multi);
}

void select_seek(WiggleIterator *wi, const char *chrom, int start, int finish) {
	seekMultiplexer(((SelectData *) wi->data)->multi, chrom, start, finish);
	pop(wi);
}

// popped: area of the records passed through since the last seek; carry: what the reference's never-reset
// `res` holds from before the last seek (statistics.c:39-43: StatisticSeek does not clear it)
struct StatData { double res; WiggleIterator *source; Reducer *gpu; double popped, carry; };

void auc_pop(WiggleIterator *wi) {
	// AUCPop (statistics.c:103-120): pass records through, integrate len*value.
	// When the source is one of our reductions the integral is the engine's own
	// (per-chromosome partials computed on the device, folded in order).
	StatData *d = (StatData *) wi->data;
	if (d->source->done) {
		if (d->gpu) d->res = d->carry + d->gpu->auc;
		wi->done = true;
		return;
	}
	wi->chrom = d->source->chrom; wi->start = d->source->start; wi->finish = d->source->finish;
	wi->value = d->source->value;
	if (!std::isnan(wi->value)) {
		const double area = (wi->finish - wi->start) * wi->value;
		if (d->gpu) d->popped += area; else d->res += area;
	}
	pop(d->source);
}

void stat_seek(WiggleIterator *wi, const char *chrom, int start, int finish) {
	// StatisticSeek (statistics.c:39-43): seek the source, pop; the running result is NOT cleared, so
	// what was integrated before the seek (usually the constructor's first record) stays in it
	StatData *d = (StatData *) wi->data;
	d->carry += d->popped;
	d->popped = 0;
	seek(d->source, chrom, start, finish);      // a GPU reduction restarts its own integral here (reducer_seek)
	pop(wi);
}

struct PearsonData { double res; Reducer *r; };

void pearson_pop(WiggleIterator *wi) {
	// PearsonPop (statistics.c:414-458) as a batch: no records are passed
	// through; the result is ready when the iterator reports done.
	PearsonData *d = (PearsonData *) wi->data;
	while (d->r->next_chromosome()) {}
	d->res = wb200_pearson_result(d->r->moments);
	wi->done = true;
}

void ndpearson_pop(WiggleIterator *wi) {
	// NDPearsonPop (statistics.c:490-542) as a batch.  rank = multi->count = 2 in the
	// reference (:551), whatever the set sizes: only the first two track pairs count.
	PearsonData *d = (PearsonData *) wi->data;
	while (d->r->next_chromosome()) {}
	double txx = 0, txy = 0, tyy = 0;
	for (int k = 0; k < d->r->nd_rank; k++) { txx += d->r->nd_moments[k][3]; txy += d->r->nd_moments[k][4]; tyy += d->r->nd_moments[k][5]; }
	d->res = (txx * tyy != 0) ? txy / std::sqrt(txx * tyy) : NAN;
	wi->done = true;
}

void pearson_seek(WiggleIterator *wi, const char *chrom, int start, int finish) {
	PearsonData *d = (PearsonData *) wi->data;
	for (WiggleIterator *it : d->r->iters) seek(it, chrom, start, finish);
	for (double &m : d->r->moments) m = 0;
	for (auto &row : d->r->nd_moments) for (double &m : row) m = 0;
	pop(wi);
}

} // namespace

extern "C" {

// -- wiggleIterator.c ---------------------------------------------------------
WiggleIterator *newWiggleIterator(void *data, void (*popFunction)(WiggleIterator *),
                                  void (*seekFunction)(WiggleIterator *, const char *, int, int),
                                  double default_value, wt_bool overlapping) {
	WiggleIterator *wi = (WiggleIterator *) calloc(1, sizeof(WiggleIterator));
	wi->data = data;
	wi->pop = popFunction;
	wi->seek = seekFunction;
	wi->value = 1;            // default for non-valued bed tracks (wiggleIterator.c:26)
	wi->overlaps = overlapping;
	wi->default_value = default_value;
	pop(wi);                  // eager first record (wiggleIterator.c:32)
	return wi;
}

void pop(WiggleIterator *wi) { if (!wi->done) wi->pop(wi); }
void runWiggleIterator(WiggleIterator *wi) { while (!wi->done) wi->pop(wi); }
void seek(WiggleIterator *wi, const char *chrom, int start, int finish) { wi->done = 0; wi->seek(wi, chrom, start, finish); }
void destroyWiggleIterator(WiggleIterator *wi) { free(wi->data); free(wi); }

WiggleIterator *UnionWiggleIterator(WiggleIterator *i) {
	UnaryData *d = (UnaryData *) calloc(1, sizeof(UnaryData));
	d->iter = i;
	return newWiggleIterator(d, &union_pop, &unary_seek, i->default_value, 0);
}

WiggleIterator *NonOverlappingWiggleIterator(WiggleIterator *i) { return i->overlaps ? UnionWiggleIterator(i) : i; }

// -- bigWiggleReader.c --------------------------------------------------------------
WiggleIterator *BigWiggleReader(char *filename, wt_bool holdFire) {
	BwSource *b = new BwSource();
	b->bw = wb200_bw_open(filename);
	if (!b->bw) {
		printf("File %s is not in BigWig format\n", filename);   // bigWiggleReader.c:116-119
		exit(1);
	}
	b->hold = holdFire != 0;
	return newWiggleIterator(b, &bw_pop, &bw_seek, 0, 0);
}

// -- multiplexer.c --------------------------------------------------------------
void popMultiplexer(Multiplexer *multi) { if (!multi->done) multi->pop(multi); }
void runMultiplexer(Multiplexer *multi) { while (!multi->done) multi->pop(multi); }
void seekMultiplexer(Multiplexer *multi, const char *chrom, int start, int finish) {
	multi->done = 0;
	multi->seek(multi, chrom, start, finish);
}

Multiplexer *newCoreMultiplexer(void *data, int count, void (*popf)(Multiplexer *),
                                void (*seekf)(Multiplexer *, const char *, int, int)) {
	Multiplexer *m = (Multiplexer *) calloc(1, sizeof(Multiplexer));
	m->count = count;
	m->values = (double *) calloc(count, sizeof(double));
	m->default_values = (double *) calloc(count, sizeof(double));
	m->inplay = (wt_bool *) calloc(count, sizeof(wt_bool));
	m->pop = popf;
	m->seek = seekf;
	m->data = data;
	return m;
}

Multiplexer *newMultiplexer(WiggleIterator **iters, int count, wt_bool strict) {
	MuxState *ms = new MuxState();
	Multiplexer *m = newCoreMultiplexer(ms, count, &gpu_multiplexer_pop, &gpu_multiplexer_seek);
	m->strict = strict;
	m->iters = (WiggleIterator **) calloc(count, sizeof(WiggleIterator *));
	for (int i = 0; i < count; i++) {
		m->iters[i] = NonOverlappingWiggleIterator(iters[i]);
		m->default_values[i] = m->values[i] = m->iters[i]->default_value;
	}
	ms->r.op = WB200_SUM;
	ms->r.strict = strict != 0;
	ms->r.n1 = count;
	ms->r.iters.assign(m->iters, m->iters + count);
	// The reference pops once here (multiplexer.c:167): the first row is on display on return.
	mux_first_row(m);
	return m;
}

// -- multiSet.c -------------------------------------------------------------------
Multiset *newMultiset(Multiplexer **multis, int count) {
	Multiset *ms = (Multiset *) calloc(1, sizeof(Multiset));
	ms->count = count;
	ms->multis = multis;   // kept, not copied (multiSet.c:118)
	ms->inplay = (wt_bool *) calloc(count, sizeof(wt_bool));
	ms->values = (double **) calloc(count, sizeof(double *));
	for (int i = 0; i < count; i++) ms->values[i] = multis[i]->values;
	return ms;
}

void popMultiset(Multiset *) { die("a Multiset is batch (GPU): consume it with TTestReduction / MWUReduction"); }

void seekMultiset(Multiset *ms, const char *chrom, int start, int finish) {
	ms->done = 0;
	for (int i = 0; i < ms->count; i++) seekMultiplexer(ms->multis[i], chrom, start, finish);
}

// -- reducers.c ---------------------------------------------------------------------
WiggleIterator *SumReduction(Multiplexer *m) { return make_reduction(WB200_SUM, m); }
WiggleIterator *ProductReduction(Multiplexer *m) { return make_reduction(WB200_PRODUCT, m); }
WiggleIterator *MeanReduction(Multiplexer *m) { return make_reduction(WB200_MEAN, m); }
WiggleIterator *VarianceReduction(Multiplexer *m) { return make_reduction(WB200_VAR, m); }
WiggleIterator *StdDevReduction(Multiplexer *m) { return make_reduction(WB200_STDDEV, m); }
WiggleIterator *EntropyReduction(Multiplexer *m) { return make_reduction(WB200_ENTROPY, m); }
WiggleIterator *CVReduction(Multiplexer *m) { return make_reduction(WB200_CV, m); }
WiggleIterator *MinReduction(Multiplexer *m) { return make_reduction(WB200_MIN, m); }
WiggleIterator *MaxReduction(Multiplexer *m) { return make_reduction(WB200_MAX, m); }
WiggleIterator *MedianReduction(Multiplexer *m) { return make_reduction(WB200_MEDIAN, m); }

WiggleIterator *SelectReduction(Multiplexer *multi, int index) {
	SelectData *d = (SelectData *) calloc(1, sizeof(SelectData));
	d->multi = multi;
	d->index = index;
	return newWiggleIterator(d, &select_pop, &select_seek, multi->default_values[index], 0);
}

WiggleIterator *FillInReduction(Multiplexer *multi, wt_bool trim) {
	if (multi->count != 2) {
		printf("The fill in operator can only work on 2 iterators! Got %i\n", multi->count);
		exit(1);
	}
	return make_reduction(trim ? WB200_TRIMFILL : WB200_FILLIN, multi);
}

// -- setComparisons.c -------------------------------------------------------------------
WiggleIterator *TTestReduction(Multiset *ms) {
	if (ms->count != 2 || ms->multis[0]->count < 3 || ms->multis[1]->count < 3) {
		puts("The t-test function only works for two sets with enough elements to compute variance");
		exit(1);
	}
	return make_setcompare(WB200_TTEST, ms);
}

WiggleIterator *MWUReduction(Multiset *ms) {
	if (ms->count != 2 || ms->multis[0]->count == 0 || ms->multis[1]->count == 0) {
		puts("The Mann-Whitney U function only works for two non-empty sets");
		exit(1);
	}
	return make_setcompare(WB200_MWU, ms);
}

WiggleIterator *FTestReduction(Multiset *) {
	die("ftest is not provided: the reference implementation is defective (setComparisons.c:183,199 loop on the wrong index)");
	return nullptr;
}

// -- statistics.c ---------------------------------------------------------------------------
WiggleIterator *AUCIntegrator(WiggleIterator *wi) {
	StatData *d = (StatData *) calloc(1, sizeof(StatData));
	d->source = NonOverlappingWiggleIterator(wi);
	d->gpu = (d->source->pop == &reducer_pop) ? (Reducer *) d->source->data : nullptr;
	WiggleIterator *res = newWiggleIterator(d, &auc_pop, &stat_seek, wi->default_value, 0);
	res->append = wi;   // newStatisticIterator (statistics.c:45-49)
	return res;
}

WiggleIterator *PearsonIntegrator(Multiplexer *multi) {
	if (!is_gpu_multiplexer(multi) || multi->count != 2) die("pearson takes a multiplexer of exactly two iterators");
	PearsonData *d = (PearsonData *) calloc(1, sizeof(PearsonData));
	d->res = NAN;
	d->r = new Reducer();
	d->r->pearson = true;
	mux_claim(multi, d->r);
	d->r->n1 = 2;
	WiggleIterator *res = newWiggleIterator(d, &pearson_pop, &pearson_seek, multi->iters[1]->default_value, 0);
	res->append = multi->iters[1];
	return res;
}

WiggleIterator *NDPearsonIntegrator(Multiset *ms) {
	if (ms->count != 2 || ms->multis[0]->count != ms->multis[1]->count) {
		fprintf(stderr, "Incorrect number of input tracks to N-dimensional Pearson correlation!\n");
		exit(1);
	}
	if (ms->multis[0]->count < 2) die("N-dimensional Pearson reads two dimensions (statistics.c:551); give at least two tracks per set");
	PearsonData *d = (PearsonData *) calloc(1, sizeof(PearsonData));
	d->res = NAN;
	d->r = new Reducer();
	d->r->pearson = true;
	d->r->nd_rank = ms->count;     // = 2: the reference's quirk, kept
	for (int m = 0; m < 2; m++) {
		if (!is_gpu_multiplexer(ms->multis[m])) die("ndpearson needs multiplexers built by newMultiplexer()");
		Reducer part;
		mux_claim(ms->multis[m], &part);
		d->r->iters.insert(d->r->iters.end(), part.iters.begin(), part.iters.end());
	}
	d->r->n1 = ms->multis[0]->count;
	WiggleIterator *res = newWiggleIterator(d, &ndpearson_pop, &pearson_seek, 0, 0);
	res->append = ms->multis[0]->iters[0];
	return res;
}

// -- apply.c ----------------------------------------------------------------------------------
// ApplyMultiplexer (src/apply.c:340-353) for a dataset that is a fresh reduction of this library.
// The reference seeks the dataset to every batch of regions, copies it into per-base buffers and
// runs the integrator chain base by base (apply.c:259-321).  Here each run of regions on one
// chromosome becomes: one seek of the reduction to the span of the run (the segments are
// computed on the device and stay there), then one wb200_apply_last call -- a segmented
// reduction per region over the resident segments.  stat_codes: WB200_STAT_* per output column.
// Returns NULL when the dataset is not a fresh reduction (the caller keeps the reference's path).
namespace {
struct ApplyState {
	WiggleIterator *regions, *dataset;
	Reducer *r;
	std::vector<int> codes;
	bool strict;
	char *chrom = nullptr;
	std::vector<int32_t> rs, re;
	std::vector<double> out;
	size_t k = 0;
};

void apply_pop(Multiplexer *m) {
	ApplyState *st = (ApplyState *) m->data;
	const int count = m->count;
	if (st->k >= st->rs.size()) {
		if (st->regions->done) { m->done = true; return; }
		// the next run of regions on one chromosome
		st->chrom = intern(st->regions->chrom);
		st->rs.clear(); st->re.clear(); st->k = 0;
		int lo = st->regions->start, hi = st->regions->finish;
		while (!st->regions->done && strcmp(st->regions->chrom, st->chrom) == 0 && st->rs.size() < (1u << 20)) {
			st->rs.push_back(st->regions->start); st->re.push_back(st->regions->finish);
			lo = std::min(lo, st->regions->start); hi = std::max(hi, st->regions->finish);
			pop(st->regions);
		}
		st->out.assign(st->rs.size() * count, NAN);
		// the dataset over the span of the run: reduced on the device, segments stay resident
		seek(st->dataset, st->chrom, lo, hi);
		const bool have = !st->dataset->done && !st->r->rs.empty() && strcmp(st->r->chrom, st->chrom) == 0;
		int rc;
		if (have)
			rc = wb200_apply_last(engine(), st->dataset->default_value, (int64_t) st->rs.size(), st->rs.data(), st->re.data(),
			                      count, st->codes.data(), st->strict ? 1 : 0, st->out.data());
		else        // nothing of the dataset in the span: every base is uncovered
			rc = wb200_apply_host(engine(), nullptr, nullptr, nullptr, 0, st->dataset->default_value, (int64_t) st->rs.size(),
			                      st->rs.data(), st->re.data(), count, st->codes.data(), st->strict ? 1 : 0, st->out.data());
		if (rc != WB200_OK) die(wb200_last_error(engine()));
	}
	m->chrom = st->chrom; m->start = st->rs[st->k]; m->finish = st->re[st->k];
	for (int i = 0; i < count; i++) m->values[i] = st->out[st->k * count + i];
	st->k++;
}

void apply_seek(Multiplexer *m, const char *chrom, int start, int finish) {
	ApplyState *st = (ApplyState *) m->data;
	st->rs.clear(); st->re.clear(); st->k = 0;
	seek(st->regions, chrom, start, finish);      // ApplyMultiplexerSeek (apply.c:323-333): drop the open targets, seek the regions
}
} // namespace

Multiplexer *wb200_apply_multiplexer(WiggleIterator *regions, const int *stat_codes, int count, WiggleIterator *dataset, wt_bool strict) {
	if (!regions || !dataset || !stat_codes || count < 1 || count > 64) return nullptr;
	if (dataset->pop != &reducer_pop || dataset->overlaps) return nullptr;
	Reducer *r = (Reducer *) dataset->data;
	if (r->pearson) return nullptr;
	for (int i = 0; i < count; i++) if (stat_codes[i] < WB200_STAT_AUC || stat_codes[i] > WB200_STAT_SPAN) return nullptr;
	ApplyState *st = new ApplyState();
	st->regions = regions; st->dataset = dataset; st->r = r; st->strict = strict != 0;
	st->codes.assign(stat_codes, stat_codes + count);
	Multiplexer *res = newCoreMultiplexer(st, count, &apply_pop, &apply_seek);
	for (int i = 0; i < count; i++) res->default_values[i] = NAN;       // apply.c:348-349
	popMultiplexer(res);                                                // apply.c:350: the first region is on display
	return res;
}

// -- wigWriter.c -------------------------------------------------------------------------------
// toFile() / toStdout() (src/wigWriter.c:278-290) for an iterator this library produced: each
// chromosome is reduced AND turned into text on the device, the host only fwrites.  The
// reference keeps its own writer for everything else; the binding that routes between the
// two is shown in INTEGRATION.md (oracle/ref_stubs.c holds it for the CLI built here).
int wb200_sink_accepts(WiggleIterator *wi) {
	if (!wi || wi->done || wi->overlaps || wi->pop != &reducer_pop) return 0;
	Reducer *r = (Reducer *) wi->data;
	return !r->pearson && r->k == 1;                       // freshly constructed: positioned on its first record
}

void wb200_sink_run(WiggleIterator *wi, FILE *out, wt_bool bedGraph) {
	if (!wb200_sink_accepts(wi)) die("wb200_sink_run: not a fresh reduction of this library");
	Reducer *r = (Reducer *) wi->data;
	const int mode = bedGraph ? WB200_SINK_BEDGRAPH : WB200_SINK_WIG;
	wb200_sink_state state{0, 0, 0};
	// the constructor's first pop already computed the first chromosome: its segments are on the host
	const char *text = nullptr;
	int64_t nbytes = 0;
	if (wb200_format_host(engine(), mode, r->chrom, r->rs.data(), r->re.data(), r->rv.data(), (int64_t) r->rs.size(), &state,
	                      &text, &nbytes) != WB200_OK)
		die(wb200_last_error(engine()));
	if (nbytes > 0 && fwrite(text, 1, (size_t) nbytes, out) != (size_t) nbytes) die("short write to the output file");
	r->k = r->rs.size();
	while (r->next_chromosome_to_sink(out, mode, &state)) { }
	fflush(out);
	wi->done = true;
}

// -- unaryOps.c: element-wise operators -----------------------------------------------------------
// ScaleWiggleIterator, ShiftWiggleIterator, NaturalLogWiggleIterator, LogWiggleIterator,
// NaturalExpWiggleIterator, PowerWiggleIterator, AbsWiggleIterator, DefaultValueWiggleIterator
// (src/unaryOps.c:210-229, 650-960) in one constructor: op is a WB200_MAP_* code (0: `default`),
// arg the CLI-level argument (scale factor, offset, log base with <= 0 meaning ln, exponent,
// the new default).  oracle/dropin_binding.c routes the reference's constructors here.
WiggleIterator *wb200_mapped_iterator(WiggleIterator *i, int op, double arg) {
	MappedSource *m = new MappedSource();
	double dflt;
	// the reference wraps the input in NonOverlappingWiggleIterator for these (unaryOps.c:675,738,802,858,893,953)
	const bool nonoverlap = op == WB200_MAP_SCALE || op == WB200_MAP_OFFSET || (op == WB200_MAP_LOG && arg > 0) ||
	                        op == WB200_MAP_EXP || op == WB200_MAP_POW || op == WB200_MAP_ABS;
	m->inner = nonoverlap ? NonOverlappingWiggleIterator(i) : i;
	m->op = op;
	if (op == 0) { m->arg = 0; dflt = arg; }
	else {
		m->arg = op == WB200_MAP_LOG ? (arg > 0 ? std::log(arg) : 1.0) : op == WB200_MAP_EXP ? 1.0 : arg;
		dflt = wb200_map_default(op, m->arg, i->default_value);
	}
	return newWiggleIterator(m, &mapped_pop, &mapped_seek, dflt, i->overlaps);
}

} // extern "C"
