// ig_pool.cu -- a caching device allocator behind every cudaMalloc / cudaFree of the library (ig_ctx.h redirects them).
//
// ig_mcmc_updating() is the drop-in call: create, load, run, destroy -- once per chain.  It allocates and frees ~10 GB at
// config 4, and on these boxes cudaMalloc / cudaFree of multi-GB buffers stall for hundreds of milliseconds now and then
// (round 1: three back-to-back calls took 0.22 / 0.32 / 0.32 s).  Freed blocks are therefore kept, keyed by (device, size),
// and handed out again to the next context: identical chains ask for identical sizes.  Blocks are zero-filled by the
// callers that need it (dalloc), exactly as after a fresh cudaMalloc.  ig_release_cache() returns everything to the driver;
// IG_NO_POOL=1 turns the cache off; when a real allocation fails the cache is emptied and the allocation retried.
#include <cuda_runtime.h>
#include <stdlib.h>
#include <map>
#include <mutex>
#include <unordered_map>
#include <utility>
#include "../../include/instruct_b200.h"

namespace {
std::mutex g_mu;
std::unordered_map<void *, std::pair<int, size_t>> g_live;          // handed-out block -> (device, size)
std::multimap<std::pair<int, size_t>, void *> g_free;               // cached blocks
size_t g_cached = 0;
bool pool_off() { static int off = getenv("IG_NO_POOL") ? 1 : 0; return off != 0; }
size_t cache_limit() { static size_t lim = getenv("IG_POOL_LIMIT_GB") ? (size_t)atof(getenv("IG_POOL_LIMIT_GB")) << 30 : (size_t)64 << 30; return lim; }

void drop_all_locked()
{
	int cur = 0;
	cudaGetDevice(&cur);
	for (auto &kv : g_free) { cudaSetDevice(kv.first.first); cudaFree(kv.second); }
	g_free.clear();
	g_cached = 0;
	cudaSetDevice(cur);
}
}  // namespace

cudaError_t ig_pool_malloc(void **p, size_t bytes)
{
	if (bytes == 0) bytes = 1;
	const size_t sz = (bytes + 511) & ~(size_t)511;
	int dev = 0;
	cudaGetDevice(&dev);
	if (!pool_off()) {
		std::lock_guard<std::mutex> lk(g_mu);
		auto it = g_free.find({dev, sz});
		if (it != g_free.end()) {
			*p = it->second;
			g_free.erase(it);
			g_cached -= sz;
			g_live[*p] = {dev, sz};
			return cudaSuccess;
		}
	}
	cudaError_t e = cudaMalloc(p, sz);
	if (e != cudaSuccess && !pool_off()) {
		cudaGetLastError();
		{ std::lock_guard<std::mutex> lk(g_mu); drop_all_locked(); }
		e = cudaMalloc(p, sz);
	}
	if (e == cudaSuccess && !pool_off()) { std::lock_guard<std::mutex> lk(g_mu); g_live[*p] = {dev, sz}; }
	return e;
}

cudaError_t ig_pool_free(void *p)
{
	if (!p) return cudaSuccess;
	if (!pool_off()) {
		std::lock_guard<std::mutex> lk(g_mu);
		auto it = g_live.find(p);
		if (it != g_live.end()) {
			const std::pair<int, size_t> key = it->second;
			g_live.erase(it);
			if (g_cached + key.second <= cache_limit()) {
				g_free.insert({key, p});
				g_cached += key.second;
				return cudaSuccess;
			}
		}
	}
	return cudaFree(p);
}

extern "C" ig_status ig_release_cache(void)
{
	std::lock_guard<std::mutex> lk(g_mu);
	drop_all_locked();
	return IG_OK;
}
