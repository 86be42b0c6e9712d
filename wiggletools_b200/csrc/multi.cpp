// wiggletools_b200 -- the whole box from one process: chromosome shards over several GPUs.
//
// The sweep never crosses a chromosome (/root/reference/src/multiplexer.c:50-74), so the shard unit is
// the chromosome: every device gets a set of chromosomes (longest-processing-time
// packing by interval count), one host thread per device drives that device's context through
// wb200_reduce_host_batch / wb200_pearson_host, and no per-segment data ever crosses GPUs.  Only the
// whole-genome scalars do: the AUC partials are summed with ncclAllReduce (ncclDouble, ncclSum, one
// element) and the six Pearson moments of every device are gathered with ncclAllGather and folded
// in device order with the pairwise merge (a plain sum of the T terms would be wrong,
// src/statistics.c:443-456).  NCCL is loaded at run time (libnccl.so.2); the messages are 8 and 48
// bytes, i.e. latency-bound -- NVLink bandwidth plays no role on this path.
//
// This is what the reference's own scale-out does with independent processes per genome region
// (python/wiggletools/parallelWiggleTools.py:66-68), inside one process and one call.
#include "../../include/wiggletools_b200.h"

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace {

struct NcclApi {
	void *lib = nullptr;
	ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
	ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
	ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*GroupStart)() = nullptr;
	ncclResult_t (*GroupEnd)() = nullptr;
	const char *(*GetErrorString)(ncclResult_t) = nullptr;
	bool load(std::string &err) {
		if (lib) return true;
		lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
		if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
		if (!lib) { err = std::string("cannot load NCCL: ") + dlerror(); return false; }
#define WB_SYM(field, name) field = (decltype(field)) dlsym(lib, name); if (!field) { err = std::string("NCCL symbol missing: ") + name; return false; }
		WB_SYM(CommInitAll, "ncclCommInitAll") WB_SYM(CommDestroy, "ncclCommDestroy") WB_SYM(AllReduce, "ncclAllReduce")
		WB_SYM(AllGather, "ncclAllGather") WB_SYM(GroupStart, "ncclGroupStart") WB_SYM(GroupEnd, "ncclGroupEnd")
		WB_SYM(GetErrorString, "ncclGetErrorString")
#undef WB_SYM
		return true;
	}
};

std::string g_multi_error;

} // namespace

struct wb200_multi {
	int ndev = 0;
	std::vector<int> devices;
	std::vector<wb200_ctx *> ctx;
	std::vector<ncclComm_t> comms;
	std::vector<cudaStream_t> streams;
	std::vector<double *> d_scalar;      // per device: 6 local doubles, then 6 * ndev gathered
	NcclApi nccl;
	std::string err;
	std::vector<int> owner;              // device of every shard of the last call
};

namespace {

int mfail(wb200_multi *m, int code, const std::string &msg) { if (m) m->err = msg; else g_multi_error = msg; return code; }

// longest-processing-time packing: owner device of every shard
std::vector<int> lpt(const std::vector<int64_t> &w, int ndev) {
	std::vector<int> order(w.size()), owner(w.size(), 0);
	for (size_t i = 0; i < w.size(); i++) order[i] = (int) i;
	std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return w[a] > w[b]; });
	std::vector<int64_t> load(ndev, 0);
	for (int i : order) {
		int best = 0;
		for (int d = 1; d < ndev; d++) if (load[d] < load[best]) best = d;
		owner[i] = best;
		load[best] += w[i];
	}
	return owner;
}

} // namespace

extern "C" {

const char *wb200_multi_last_error(const wb200_multi *m) { return m ? m->err.c_str() : g_multi_error.c_str(); }

wb200_multi *wb200_multi_create(int ndev, const int *devices) {
	int count = 0;
	if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) { g_multi_error = "wiggletools_b200: no usable CUDA device; this engine has no CPU fallback"; return nullptr; }
	if (ndev <= 0) ndev = count;
	wb200_multi *m = new wb200_multi();
	m->ndev = ndev;
	for (int d = 0; d < ndev; d++) m->devices.push_back(devices ? devices[d] : d);
	for (int d = 0; d < ndev; d++) {
		wb200_ctx *c = wb200_create(m->devices[d]);
		if (!c) { g_multi_error = wb200_last_error(nullptr); wb200_multi_destroy(m); return nullptr; }
		m->ctx.push_back(c);
	}
	if (!m->nccl.load(g_multi_error)) { wb200_multi_destroy(m); return nullptr; }
	m->comms.assign(ndev, nullptr);
	ncclResult_t r = m->nccl.CommInitAll(m->comms.data(), ndev, m->devices.data());
	if (r != ncclSuccess) { g_multi_error = std::string("ncclCommInitAll: ") + m->nccl.GetErrorString(r); m->comms.clear(); wb200_multi_destroy(m); return nullptr; }
	m->streams.assign(ndev, nullptr);
	m->d_scalar.assign(ndev, nullptr);
	for (int d = 0; d < ndev; d++) {
		cudaSetDevice(m->devices[d]);
		if (cudaStreamCreateWithFlags(&m->streams[d], cudaStreamNonBlocking) != cudaSuccess ||
		    cudaMalloc(&m->d_scalar[d], sizeof(double) * 6 * (ndev + 1)) != cudaSuccess) {
			g_multi_error = "wb200_multi_create: stream / scalar buffer allocation failed";
			wb200_multi_destroy(m);
			return nullptr;
		}
	}
	return m;
}

void wb200_multi_destroy(wb200_multi *m) {
	if (!m) return;
	for (size_t d = 0; d < m->comms.size(); d++) if (m->comms[d]) m->nccl.CommDestroy(m->comms[d]);
	for (size_t d = 0; d < m->d_scalar.size(); d++) if (m->d_scalar[d]) { cudaSetDevice(m->devices[d]); cudaFree(m->d_scalar[d]); }
	for (size_t d = 0; d < m->streams.size(); d++) if (m->streams[d]) { cudaSetDevice(m->devices[d]); cudaStreamDestroy(m->streams[d]); }
	for (wb200_ctx *c : m->ctx) wb200_destroy(c);
	delete m;
}

int wb200_multi_ndev(const wb200_multi *m) { return m ? m->ndev : 0; }
int wb200_multi_owner(const wb200_multi *m, int shard) { return (m && shard >= 0 && shard < (int) m->owner.size()) ? m->owner[shard] : -1; }

// `count` doubles per device -> all-reduced (sum) or all-gathered over NCCL; result on the host
static int exchange(wb200_multi *m, const std::vector<std::vector<double>> &local, int count, bool gather, std::vector<double> &result) {
	const int nd = m->ndev;
	for (int d = 0; d < nd; d++) {
		cudaSetDevice(m->devices[d]);
		if (cudaMemcpyAsync(m->d_scalar[d], local[d].data(), count * sizeof(double), cudaMemcpyHostToDevice, m->streams[d]) != cudaSuccess)
			return mfail(m, WB200_ECUDA, "scalar upload failed");
	}
	ncclResult_t r = m->nccl.GroupStart();
	for (int d = 0; d < nd && r == ncclSuccess; d++) {
		cudaSetDevice(m->devices[d]);
		r = gather ? m->nccl.AllGather(m->d_scalar[d], m->d_scalar[d] + 6, count, ncclDouble, m->comms[d], m->streams[d])
		           : m->nccl.AllReduce(m->d_scalar[d], m->d_scalar[d] + 6, count, ncclDouble, ncclSum, m->comms[d], m->streams[d]);
	}
	if (r == ncclSuccess) r = m->nccl.GroupEnd(); else m->nccl.GroupEnd();
	if (r != ncclSuccess) return mfail(m, WB200_ECUDA, std::string("NCCL: ") + m->nccl.GetErrorString(r));
	result.assign(gather ? (size_t) count * nd : (size_t) count, 0.0);
	cudaSetDevice(m->devices[0]);
	if (cudaMemcpyAsync(result.data(), m->d_scalar[0] + 6, result.size() * sizeof(double), cudaMemcpyDeviceToHost, m->streams[0]) != cudaSuccess)
		return mfail(m, WB200_ECUDA, "scalar read-back failed");
	for (int d = 0; d < nd; d++) { cudaSetDevice(m->devices[d]); cudaStreamSynchronize(m->streams[d]); }
	return WB200_OK;
}

int wb200_multi_reduce_host(wb200_multi *m, int op, int nshards, int ntracks, int nset1, int strict,
                            const wb200_track *tracks, wb200_segments *out, double *auc_total) {
	if (!m || !tracks || !out || nshards < 0 || ntracks < 1) return WB200_EINVAL;
	const int nd = m->ndev;
	std::vector<int64_t> weight(nshards, 0);
	for (int c = 0; c < nshards; c++) for (int i = 0; i < ntracks; i++) weight[c] += tracks[(size_t) c * ntracks + i].n;
	m->owner = lpt(weight, nd);
	std::vector<std::vector<int>> mine(nd);
	for (int c = 0; c < nshards; c++) mine[m->owner[c]].push_back(c);
	std::vector<int> rc(nd, WB200_OK);
	std::vector<std::vector<double>> local(nd, std::vector<double>(1, 0.0));
	std::vector<std::thread> pool;
	for (int d = 0; d < nd; d++)
		pool.emplace_back([&, d]() {
			// this device's shards, gathered into one contiguous call (results go straight to the caller's slots)
			const std::vector<int> &sh = mine[d];
			if (sh.empty()) return;
			wb200_bind_host_thread(m->ctx[d]);      // staging buffers of this thread live on the device's NUMA node
			std::vector<wb200_track> tr((size_t) sh.size() * ntracks);
			for (size_t k = 0; k < sh.size(); k++)
				memcpy(&tr[k * ntracks], &tracks[(size_t) sh[k] * ntracks], sizeof(wb200_track) * ntracks);
			std::vector<wb200_segments> res(sh.size());
			rc[d] = wb200_reduce_host_batch(m->ctx[d], op, (int) sh.size(), ntracks, nset1, strict, tr.data(), res.data());
			if (rc[d] != WB200_OK) return;
			double a = 0;
			for (size_t k = 0; k < sh.size(); k++) { out[sh[k]] = res[k]; a += res[k].auc; }
			local[d][0] = a;
		});
	for (auto &t : pool) t.join();
	for (int d = 0; d < nd; d++) if (rc[d] != WB200_OK) return mfail(m, rc[d], std::string("device ") + std::to_string(m->devices[d]) + ": " + wb200_last_error(m->ctx[d]));
	// the whole-genome scalar: the path's only exchange step
	std::vector<double> total;
	int e = exchange(m, local, 1, false, total);
	if (e != WB200_OK) return e;
	if (auc_total) *auc_total = total[0];
	return WB200_OK;
}

int wb200_multi_pearson_host(wb200_multi *m, int nshards, const wb200_track *x, const wb200_track *y, double moments[6]) {
	if (!m || !x || !y || !moments || nshards < 0) return WB200_EINVAL;
	const int nd = m->ndev;
	std::vector<int64_t> weight(nshards, 0);
	for (int c = 0; c < nshards; c++) weight[c] = x[c].n + y[c].n;
	// contiguous runs of shards per device: the moments are folded in shard order within a device and in
	// device order across devices, i.e. in the caller's chromosome order overall
	std::vector<int> first(nd + 1, 0);
	{
		int64_t total = 0, acc = 0;
		for (int64_t w : weight) total += w;
		int d = 0;
		for (int c = 0; c < nshards; c++) {
			if (d + 1 < nd && acc >= (total * (d + 1)) / nd) first[++d] = c;
			acc += weight[c];
		}
		while (d < nd) first[++d] = nshards;
	}
	m->owner.assign(nshards, 0);
	for (int d = 0; d < nd; d++) for (int c = first[d]; c < first[d + 1]; c++) m->owner[c] = d;
	std::vector<int> rc(nd, WB200_OK);
	std::vector<std::vector<double>> local(nd, std::vector<double>(6, 0.0));
	std::vector<std::thread> pool;
	for (int d = 0; d < nd; d++)
		pool.emplace_back([&, d]() {
			for (int c = first[d]; c < first[d + 1]; c++) {
				double mm[6];
				rc[d] = wb200_pearson_host(m->ctx[d], &x[c], &y[c], mm);
				if (rc[d] != WB200_OK) return;
				wb200_pearson_merge(local[d].data(), mm);
			}
		});
	for (auto &t : pool) t.join();
	for (int d = 0; d < nd; d++) if (rc[d] != WB200_OK) return mfail(m, rc[d], std::string("device ") + std::to_string(m->devices[d]) + ": " + wb200_last_error(m->ctx[d]));
	std::vector<double> all;
	int e = exchange(m, local, 6, true, all);
	if (e != WB200_OK) return e;
	for (int k = 0; k < 6; k++) moments[k] = 0;
	for (int d = 0; d < nd; d++) wb200_pearson_merge(moments, &all[(size_t) 6 * d]);
	return WB200_OK;
}

} // extern "C"
