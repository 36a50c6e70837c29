// ig_api.cu -- the C-ABI of include/instruct_b200.h: context, chain driver, multi-GPU exchange,
// the host-side Dirichlet-process step, and the state hooks the parity tests use.
//
// The chain driver restates the control flow of mcmc_POP_selfing / mcmc_INDV_selfing
// (mcmc.c:182-239, 297-383) around the kernels in ig_kernels.cu.  Nothing here falls back
// to the CPU: every entry point needs a CUDA device.
#include <dlfcn.h>
#include <nccl.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <string>
#include <utility>
#include <vector>
#include <algorithm>
#include <unistd.h>

#include <time.h>
#include "ig_ctx.h"

static double wall_ms()
{
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}
#include "philox.cuh"
#include "samplers.cuh"
#include "sweep_common.cuh"


// --------------------------------------------------------------------------------------
// errors
// --------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
thread_local cudaStream_t ig_alloc_stream = nullptr;
ig_status ig_fail(ig_status st, const char *fmt, ...)
{
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g_err, sizeof(g_err), fmt, ap);
	va_end(ap);
	return st;
}

extern "C" const char *ig_last_error(void) { return g_err; }
extern "C" const char *ig_version(void) { return "instruct_b200 0.2 (sm_100a)"; }
extern "C" int ig_device_count(void)
{
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
	return n;
}

// --------------------------------------------------------------------------------------
// NCCL, bound at run time (the library must load on a box without NCCL or a GPU)
// --------------------------------------------------------------------------------------
struct NcclApi {
	void *h = nullptr;
	ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
	ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
	ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
	ncclResult_t (*CommSplit)(ncclComm_t, int, int, ncclComm_t *, void *) = nullptr;     // optional (NCCL >= 2.18)
	ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*ReduceScatter)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
	const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;
static ig_status nccl_load()
{
	if (g_nccl.h) return IG_OK;
	void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
	if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
	if (!h) return fail(IG_ERR_NCCL, "cannot load libnccl.so.2: %s", dlerror());
#define SYM(f, name)                                                    \
	*(void **)(&g_nccl.f) = dlsym(h, name);                             \
	if (!g_nccl.f) return fail(IG_ERR_NCCL, "libnccl lacks %s", name)
	SYM(GetUniqueId, "ncclGetUniqueId");
	SYM(CommInitRank, "ncclCommInitRank");
	SYM(CommDestroy, "ncclCommDestroy");
	SYM(AllReduce, "ncclAllReduce");
	SYM(AllGather, "ncclAllGather");
	SYM(ReduceScatter, "ncclReduceScatter");
	SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
	*(void **)(&g_nccl.CommSplit) = dlsym(h, "ncclCommSplit");
	g_nccl.h = h;
	return IG_OK;
}
#define NCK(call)                                                                                   \
	do {                                                                                        \
		ncclResult_t r_ = (call);                                                               \
		if (r_ != ncclSuccess) return fail(IG_ERR_NCCL, "%s: %s", #call, g_nccl.GetErrorString(r_)); \
	} while (0)

constexpr int FX_POST = 32, LOCAL_MAX_K = 10;      // layout of ig_ctx::fx: the post-sweep sums, then 2^K subset sums of (value, counts)
constexpr int PT_POINTS = 8, PT_SWEEPS = 64;        // IG_PHASE_TRACE: events per sweep, sweeps traced
static bool getenv_once(const char *name)
{
	// (a handful of debugging switches; looked up per call -- getenv is a short list walk, not on any device-side path)
	const char *v = getenv(name);
	return v && *v && *v != '0';
}
static void ptrace_report(ig_ctx *c);
static void peer_teardown(ig_ctx *c, bool collective);

extern "C" ig_status ig_create(const ig_config *cfg, ig_ctx **out)
{
	if (!cfg || !out) return fail(IG_ERR_ARG, "null argument");
	*out = nullptr;
	if (cfg->ploid != 2 && cfg->ploid != 4) return fail(IG_ERR_UNSUPPORTED, "ploid %d: 2 (diploid) or 4 (autotetraploid)", cfg->ploid);
	if (cfg->ploid == 2 && (cfg->mode < 0 || cfg->mode > 5)) return fail(IG_ERR_UNSUPPORTED, "mode %d: modes 0 to 5 exist", cfg->mode);
	if (cfg->ploid == 2 && cfg->mode == 5 && cfg->prior_flag != 0)
		return fail(IG_ERR_UNSUPPORTED, "mode 5 with the Dirichlet-process prior (-f 1) is not built; use the uniform prior (-f 0)");
	if (cfg->popnum < 1 || cfg->popnum > MAX_K) return fail(IG_ERR_UNSUPPORTED, "popnum %d outside 1..%d", cfg->popnum, MAX_K);
	if (cfg->locinum < 1 || cfg->totalsize < 1) return fail(IG_ERR_ARG, "empty data set (N=%d, L=%d)", cfg->totalsize, cfg->locinum);
	if (cfg->mode == 3 && cfg->prior_flag != 0 && cfg->prior_flag != 1) return fail(IG_ERR_ARG, "prior_flag must be 0 or 1");
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
		cudaGetLastError();
		return fail(IG_ERR_CUDA, "no CUDA device: instruct_b200 has no CPU path");
	}
	if (cfg->device < 0 || cfg->device >= ndev) return fail(IG_ERR_ARG, "device %d of %d", cfg->device, ndev);
	CK(cudaSetDevice(cfg->device));

	ig_ctx *c = new ig_ctx();
	c->cfg = *cfg;
	const int count = cfg->shard_count > 1 ? cfg->shard_count : 1;
	const int rank = count > 1 ? cfg->shard_rank : 0;
	const int N = cfg->totalsize;
	c->shard_cap = (N + count - 1) / count;
	c->Npad = c->shard_cap * count;
	Geometry &g = c->geo;
	g.N = N;
	g.i0 = rank * c->shard_cap;
	g.Nloc = N - g.i0 < c->shard_cap ? N - g.i0 : c->shard_cap;
	if (g.Nloc < 1) { delete c; return fail(IG_ERR_ARG, "shard %d of %d is empty for N=%d", rank, count, N); }
	if (count > 1 && cfg->shard_size != 0 && (cfg->shard_size != g.Nloc || cfg->shard_begin != g.i0)) {
		const int b = g.i0, e = g.i0 + g.Nloc;
		delete c;
		return fail(IG_ERR_ARG, "shard %d must be [%d,%d) (got begin %d size %d)", rank, b, e, cfg->shard_begin, cfg->shard_size);
	}
	g.L = cfg->locinum;
	g.Lpad = (g.L + TILE - 1) / TILE * TILE;
	g.LT = g.Lpad / TILE;
	g.K = cfg->popnum;
	g.KP = (int)pad_k(g.K);
	g.A = 2;                   // fixed when the genotypes are loaded
	g.fmode = (cfg->ploid == 2 && cfg->mode == 5) ? 1 : ((cfg->ploid == 2 && cfg->mode == 4) ? 2 : 0);
	g.REC = g.K + 3 + (g.fmode == 2 ? 2 * g.K : 0);
	if (g.fmode) c->cfg.type_freq = 1;                                // log_ld_F_* have no -y 0 branch (mcmc.c:1776-1847)
	c->ns = (cfg->mode == 3 || cfg->mode == 5) ? N : (cfg->mode <= 1 ? 0 : g.K);       // modes 0 and 1 have no selfing rates
	c->rounds = (cfg->rng_rounds == 10) ? 10 : 7;
	c->key0 = (uint32_t)cfg->seed;
	c->key1 = (uint32_t)(cfg->seed >> 32);
	// the chain's stream gets the highest priority, the side stream of a sharded chain (tally all-reduce + next P draw) the lowest:
	// its 6000 short CTAs otherwise take SM slots from the few small kernels on the sweep's critical path
	int prio_least = 0, prio_greatest = 0;
	cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest);
	if (cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, prio_greatest) != cudaSuccess) { delete c; return fail(IG_ERR_CUDA, "stream creation failed"); }
	ig_alloc_stream = c->stream;
	if (getenv("IG_PHASE_TRACE")) {
		c->ptrace.resize((size_t)PT_POINTS * PT_SWEEPS);
		for (auto &e : c->ptrace) cudaEventCreate(&e);
	}
	if (cfg->ploid == 4) {
		ig_status st = tetra_create(c);
		if (st != IG_OK) { cudaStreamDestroy(c->stream); delete c; return st; }
	}
	*out = c;
	return IG_OK;
}

static void free_all(ig_ctx *c)
{
	cudaFree(c->Es); cudaFree(c->Hs); cudaFree(c->Zs); cudaFree(c->Pc); cudaFree(c->Pcnext);
	cudaFree(c->Xt); cudaFree(c->Zt); cudaFree(c->P); cudaFree(c->P64); cudaFree(c->n); cudaFree(c->allelenum);
	cudaFree(c->ind); cudaFree(c->Qf); cudaFree(c->gprop); cudaFree(c->gpair); cudaFree(c->S); cudaFree(c->state);
	cudaFree(c->sc); cudaFree(c->pcnt); cudaFree(c->plog); cudaFree(c->pnsh); cudaFree(c->nhet); cudaFree(c->nsh); cudaFree(c->cnt); cudaFree(c->llparts); cudaFree(c->initd_dev);
	cudaFree(c->scratch); cudaFree(c->gpart); cudaFree(c->state2); cudaFree(c->S2); cudaFree(c->fx); cudaFree(c->fxg);
	cudaFree(c->fprop); cudaFree(c->hpair); cudaFree(c->ftab); cudaFree(c->pfk);
	cudaFree(c->logP); cudaFree(c->na_pll);
	cudaFree(c->mom.tot); cudaFree(c->mom.indvlkh); cudaFree(c->mom.qq); cudaFree(c->mom.qq2); cudaFree(c->mom.self);
	cudaFree(c->mom.self2); cudaFree(c->mom.gen); cudaFree(c->mom.gen2); cudaFree(c->mom.freq); cudaFree(c->mom.freq2);
	cudaFree(c->mom.convg); cudaFree(c->mom.convg_S);
}

extern "C" void ig_destroy(ig_ctx *c)
{
	if (!c) return;
	cudaSetDevice(c->cfg.device);
	if (c->stream) cudaStreamSynchronize(c->stream);
	ptrace_report(c);
	for (auto e : c->ptrace) cudaEventDestroy(e);
	if (c->stream2) cudaStreamSynchronize(c->stream2);
	peer_teardown(c, c->px);         // before the communicator goes (its barrier) and before Pnext is freed (it may point into the arena)
	if (c->comm2 && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm2);
	if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
	for (auto e : c->ev) cudaEventDestroy(e);
	if (c->graph_exec) cudaGraphExecDestroy(c->graph_exec);
	if (c->stream2) { cudaStreamSynchronize(c->stream2); cudaStreamDestroy(c->stream2); }
	if (c->ev_zq) cudaEventDestroy(c->ev_zq);
	if (c->ev_p) cudaEventDestroy(c->ev_p);
	cudaFree(c->Pnext);
	cudaFree(c->nred);
	cudaFree(c->g8_dev);
	if (c->g8_host) cudaFreeHost(c->g8_host);
	if (c->S_pin) cudaFreeHost(c->S_pin);
	if (c->ev_g) cudaEventDestroy(c->ev_g);
	tetra_destroy(c);
	free_all(c);
	if (c->stream) cudaStreamDestroy(c->stream);
	delete c;
}

// allocate everything that depends on allelenum_max, then tile the genotype store
static ig_status finish_load(ig_ctx *c, const int16_t *x_dev_canon)
{
	Geometry &g = c->geo;
	int amax = 1;
	for (int l = 0; l < g.L; l++) if (c->allelenum_h[l] > amax) amax = c->allelenum_h[l];
	if (amax > 32000) return fail(IG_ERR_ARG, "allelenum_max %d does not fit the int16 genotype store", amax);
	g.A = amax < 2 ? 2 : amax;
	g.snp = 0;
	if (snp_eligible(g, c->cfg.mode, c->cfg.type_freq)) CK(snp_configure(g, c->cfg.device));
	else CK(zq_configure(g, c->cfg.device));
	const size_t tiles = (size_t)g.LT * g.Nloc * TILE * 2;
	const size_t pn = (size_t)g.Lpad * g.A * g.KP;
	CK(dalloc(&c->Xt, tiles));
	if (g.snp) {
		const size_t ne = (size_t)g.nchunks * g.Nloc * g.TL;
		CK(dalloc(&c->Es, ne));
		CK(dalloc(&c->Zs, ne));
		CK(dalloc(&c->Hs, (size_t)g.nchunks * g.Nloc));
		CK(dalloc(&c->Pc, snp_pc_floats(g)));
	} else CK(dalloc(&c->Zt, tiles));
	// a sharded chain reduce-scatters the tally by whole loci and all-gathers P (phase_zq): room for W equal blocks of loci
	const int Wl = c->cfg.shard_count > 1 ? c->cfg.shard_count : 1;
	c->p_lr = (g.Lpad + Wl - 1) / Wl;
	const size_t pn_alloc = (size_t)Wl * c->p_lr * g.A * g.KP;
	CK(dalloc(&c->P, pn_alloc));
	CK(dalloc(&c->n, pn_alloc));
	if (c->cfg.print_freq) CK(dalloc(&c->P64, (size_t)g.K * g.L * g.A));
	CK(dalloc(&c->ind, (size_t)c->Npad * g.REC));
	CK(dalloc(&c->Qf, (size_t)g.Nloc * g.KP));
	CK(dalloc(&c->gprop, (size_t)c->Npad));
	CK(dalloc(&c->gpair, (size_t)g.Nloc));
	CK(dalloc(&c->S, (size_t)(c->ns > g.K ? c->ns : g.K)));
	CK(dalloc(&c->state, (size_t)MAX_K));
	CK(dalloc(&c->sc, 1));
	CK(dalloc(&c->pcnt, (size_t)g.nchunks * g.Nloc * g.KP));
	CK(dalloc(&c->plog, (size_t)g.nchunks * 3 * g.Nloc));
	CK(dalloc(&c->pnsh, (size_t)g.nchunks * g.Nloc));
	CK(dalloc(&c->nhet, (size_t)g.Nloc));
	CK(dalloc(&c->nsh, (size_t)g.Nloc));
	CK(dalloc(&c->cnt, (size_t)g.Nloc * g.K));
	CK(dalloc(&c->llparts, (size_t)g.Nloc * 4));
	CK(dalloc(&c->initd_dev, (size_t)MAX_K));
	CK(dalloc(&c->scratch, (size_t)64));
	CK(dalloc(&c->gpart, (size_t)2 * SC_MAX_CTAS * 20));
	c->grid_pre = scalar_grid(0, g.N, c->cfg.device);
	c->grid_post = scalar_grid(1, g.N, c->cfg.device);
	CK(dalloc(&c->state2, (size_t)MAX_K));
	CK(dalloc(&c->S2, (size_t)MAX_K));
	CK(dalloc(&c->fx, (size_t)FX_POST + 2 * ((size_t)1 << LOCAL_MAX_K)));
	CK(dalloc(&c->fxg, (size_t)FX_POST + 2 * ((size_t)1 << LOCAL_MAX_K)));
	if (c->cfg.mode == 0) { ig_status st0 = na_alloc(c); if (st0 != IG_OK) return st0; }
	if (g.fmode) {
		CK(dalloc(&c->fprop, (size_t)(c->ns > g.K ? c->ns : g.K)));
		CK(dalloc(&c->hpair, (size_t)g.Nloc));
		CK(dalloc(&c->ftab, (size_t)MAX_K * 5));
		if (g.fmode == 2) CK(dalloc(&c->pfk, (size_t)g.nchunks * g.Nloc * 2 * g.KP));
	}
	CK(dalloc(&c->mom.tot, 2));
	CK(dalloc(&c->mom.indvlkh, (size_t)c->Npad));
	CK(dalloc(&c->mom.qq, (size_t)c->Npad * g.K));
	CK(dalloc(&c->mom.qq2, (size_t)c->Npad * g.K));
	CK(dalloc(&c->mom.self, (size_t)c->ns));
	CK(dalloc(&c->mom.self2, (size_t)c->ns));
	CK(dalloc(&c->mom.gen, (size_t)c->Npad));
	CK(dalloc(&c->mom.gen2, (size_t)c->Npad));
	CK(dalloc(&c->mom.convg, (size_t)(c->cfg.ckrep > 0 ? c->cfg.ckrep : 1)));
	if (c->cfg.print_freq) {
		CK(dalloc(&c->mom.freq, (size_t)g.K * g.L * g.A));
		CK(dalloc(&c->mom.freq2, (size_t)g.K * g.L * g.A));
	}
	CK(launch_tile_x(x_dev_canon, c->Xt, c->allelenum, g, c->stream));
	CK(launch_het_counts(c->Xt, nullptr, c->nhet, nullptr, g, c->stream));
	CK(cudaMemsetAsync(c->nsh, 0, (size_t)g.Nloc * sizeof(int32_t), c->stream));
	c->launches += 2;
	if (g.snp) { CK(launch_snp_tile(c->Xt, c->Es, c->Hs, g, c->stream)); c->launches++; }
	CK(cudaStreamSynchronize(c->stream));
	c->loaded = true;
	return IG_OK;
}

extern "C" ig_status ig_load_genotypes(ig_ctx *c, const int16_t *x_host, const int32_t *allelenum_host)
{
	if (!c || !x_host || !allelenum_host) return fail(IG_ERR_ARG, "null argument");
	if (c->loaded) return fail(IG_ERR_STATE, "genotypes already loaded");
	CK(cudaSetDevice(c->cfg.device));
	ig_alloc_stream = c->stream;
	Geometry &g = c->geo;
	c->allelenum_h.assign(allelenum_host, allelenum_host + g.L);
	CK(dalloc(&c->allelenum, (size_t)g.Lpad));
	CK(cudaMemcpyAsync(c->allelenum, allelenum_host, (size_t)g.L * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
	int16_t *tmp = nullptr;
	const size_t bytes = (size_t)g.L * g.Nloc * c->cfg.ploid * sizeof(int16_t);
	const bool trace = getenv("IG_TRACE") != nullptr;
	const double t0 = wall_ms();
	CK(cudaMalloc((void **)&tmp, bytes));
	const double t1 = wall_ms();
	cudaError_t e = cudaMemcpyAsync(tmp, x_host, bytes, cudaMemcpyHostToDevice, c->stream);
	if (e != cudaSuccess) { cudaFree(tmp); CK(e); }
	if (trace) cudaStreamSynchronize(c->stream);
	const double t2 = wall_ms();
	ig_status st = c->tetra ? tetra_load(c, tmp) : finish_load(c, tmp);
	const double t3 = wall_ms();
	cudaFree(tmp);
	if (trace) fprintf(stderr, "[ig_trace] load: malloc %.1f ms, H2D %.1f (%.1f GB/s), allocate+tile %.1f, free %.1f\n", t1 - t0, t2 - t1,
	                   bytes / (t2 - t1) * 1e-6, t3 - t2, wall_ms() - t3);
	return st;
}

extern "C" ig_status ig_load_genotypes_device(ig_ctx *c, const int16_t *x_dev, const int32_t *allelenum_dev)
{
	if (!c || !x_dev || !allelenum_dev) return fail(IG_ERR_ARG, "null argument");
	if (c->loaded) return fail(IG_ERR_STATE, "genotypes already loaded");
	CK(cudaSetDevice(c->cfg.device));
	ig_alloc_stream = c->stream;
	Geometry &g = c->geo;
	c->allelenum_h.resize(g.L);
	// the caller produced x_dev / allelenum_dev on its own stream(s): wait for the whole device once
	CK(cudaDeviceSynchronize());
	CK(cudaMemcpyAsync(c->allelenum_h.data(), allelenum_dev, (size_t)g.L * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
	CK(cudaStreamSynchronize(c->stream));
	CK(dalloc(&c->allelenum, (size_t)g.Lpad));
	CK(cudaMemcpyAsync(c->allelenum, allelenum_dev, (size_t)g.L * sizeof(int32_t), cudaMemcpyDeviceToDevice, c->stream));
	return c->tetra ? tetra_load(c, x_dev) : finish_load(c, x_dev);
}

// --------------------------------------------------------------------------------------
// multi-GPU
// --------------------------------------------------------------------------------------
extern "C" ig_status ig_comm_unique_id(void *id128)
{
	if (!id128) return fail(IG_ERR_ARG, "null argument");
	ig_status st = nccl_load();
	if (st != IG_OK) return st;
	static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
	ncclUniqueId id;
	NCK(g_nccl.GetUniqueId(&id));
	memcpy(id128, &id, 128);
	return IG_OK;
}

// Map every rank's exchange buffer into every rank (CUDA IPC; the 64-byte handles travel through the communicator that was
// just created).  All ranks end up with px == true or all with px == false (then the sums go through ncclAllReduce).
static void peer_teardown(ig_ctx *c, bool collective)
{
	if (c->p_peer) { c->P = c->Pnext = nullptr; c->n = nullptr; c->p_peer = false; }      // they lived in the arena
	for (void *p : c->px_mapped) if (p) cudaIpcCloseMemHandle(p);
	c->px_mapped.clear();
	if (c->px_peers) { cudaFree(c->px_peers); c->px_peers = nullptr; }
	// Freeing memory that another process still has mapped is undefined: every rank closes its mappings first, then all meet
	// in a barrier on the communicator, then free.  `collective`: every rank of the communicator is inside this call (the
	// agreed failure branch of peer_setup; ig_destroy of a context whose set-up succeeded -- which is all ranks or none).
	// A one-rank arena was never exported.
	bool may_free = c->cfg.shard_count <= 1;
	if (c->cfg.shard_count > 1 && collective && c->comm) {
		int32_t *f = nullptr;
		if (ig_pool_malloc((void **)&f, 16) == cudaSuccess) {
			cudaMemsetAsync(f, 0, 16, c->stream);
			may_free = g_nccl.AllReduce(f, f, 1, ncclInt32, ncclSum, c->comm, c->stream) == ncclSuccess &&
			           cudaStreamSynchronize(c->stream) == cudaSuccess;
			ig_pool_free(f);
		}
	}
	if (c->px_buf && may_free) (cudaFree)(c->px_buf);     // otherwise left to process exit (a peer is gone)
	c->px_buf = nullptr;
	c->px = false;
}
static ig_status peer_setup(ig_ctx *c)
{
	const int W = c->cfg.shard_count > 1 ? c->cfg.shard_count : 1, me = W > 1 ? c->cfg.shard_rank : 0;
	// the arena: the scalar exchange buffer, then (when the tally -> P exchange goes over peer memory too) n and both P buffers
	const Geometry &g = c->geo;
	const bool want_p = c->stream2 && !g.snp && !getenv_once("IG_P_NCCL");
	const size_t pn_bytes = (size_t)W * c->p_lr * g.A * g.KP * 4;
	auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
	const size_t sc_bytes = up(px_buffer_words(W) * sizeof(unsigned long long));
	const size_t bytes = sc_bytes + (want_p ? 3 * up(pn_bytes) : 0);
	int bad = 0;
	if ((cudaMalloc)((void **)&c->px_buf, bytes) != cudaSuccess) { cudaGetLastError(); c->px_buf = nullptr; bad = 1; }
	if (!bad) { CK(cudaMemsetAsync(c->px_buf, 0, bytes, c->stream)); CK(cudaStreamSynchronize(c->stream)); }
	// what every rank tells the others: the IPC handle of its arena and its process id.  CUDA IPC maps memory of OTHER
	// processes; ranks that are threads of one process (the `inbreed` CLI's --shard individuals) keep the NCCL exchanges.
	struct PeerBlob { cudaIpcMemHandle_t h; long long pid; long long pad; };
	std::vector<PeerBlob> hs((size_t)W);
	memset(hs.data(), 0, hs.size() * sizeof(PeerBlob));
	hs[(size_t)me].pid = (long long)getpid();
	if (!bad && W > 1 && cudaIpcGetMemHandle(&hs[(size_t)me].h, c->px_buf) != cudaSuccess) { cudaGetLastError(); bad = 1; }
	char *hd = nullptr;
	CK(dalloc(&hd, (size_t)W * sizeof(PeerBlob) + 16));
	CK(cudaMemcpyAsync(hd + (size_t)me * sizeof(PeerBlob), &hs[(size_t)me], sizeof(PeerBlob), cudaMemcpyHostToDevice, c->stream));
	NCK(g_nccl.AllGather(hd + (size_t)me * sizeof(PeerBlob), hd, sizeof(PeerBlob), ncclChar, c->comm, c->stream));
	CK(cudaMemcpyAsync(hs.data(), hd, (size_t)W * sizeof(PeerBlob), cudaMemcpyDeviceToHost, c->stream));
	CK(cudaStreamSynchronize(c->stream));
	std::vector<unsigned long long *> peers((size_t)W, nullptr);
	c->px_mapped.assign((size_t)W, nullptr);
	for (int r = 0; r < W; r++) if (r != me && hs[(size_t)r].pid == hs[(size_t)me].pid) bad = 1;
	for (int r = 0; r < W && !bad; r++) {
		if (r == me) { peers[(size_t)r] = c->px_buf; continue; }
		void *p = nullptr;
		if (cudaIpcOpenMemHandle(&p, hs[(size_t)r].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); bad = 1; break; }
		c->px_mapped[(size_t)r] = p;
		peers[(size_t)r] = (unsigned long long *)p;
	}
	// agree (and make sure every rank's zero fill is complete before anyone stores into a peer): sum of the failures
	int32_t *flag_dev = reinterpret_cast<int32_t *>(hd + (size_t)W * sizeof(PeerBlob));
	CK(cudaMemcpyAsync(flag_dev, &bad, sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
	NCK(g_nccl.AllReduce(flag_dev, flag_dev, 1, ncclInt32, ncclSum, c->comm, c->stream));
	int32_t nbad = 0;
	CK(cudaMemcpyAsync(&nbad, flag_dev, sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
	CK(cudaStreamSynchronize(c->stream));
	cudaFree(hd);
	if (nbad) { peer_teardown(c, true); return IG_OK; }
	CK(dalloc(&c->px_peers, (size_t)W));
	CK(cudaMemcpyAsync(c->px_peers, peers.data(), (size_t)W * sizeof(void *), cudaMemcpyHostToDevice, c->stream));
	CK(cudaStreamSynchronize(c->stream));
	c->px_seq = 0;
	c->px = true;
	if (want_p) {
		// n, P and Pnext move into the arena (their contents are written by ig_chain_init and by every sweep; nothing to copy)
		c->px_n_off = sc_bytes;
		c->px_p_off[0] = sc_bytes + up(pn_bytes);
		c->px_p_off[1] = sc_bytes + 2 * up(pn_bytes);
		cudaFree(c->n); cudaFree(c->P); cudaFree(c->Pnext); cudaFree(c->nred);
		c->nred = nullptr;
		char *base = reinterpret_cast<char *>(c->px_buf);
		c->n = reinterpret_cast<int32_t *>(base + c->px_n_off);
		c->P = reinterpret_cast<float *>(base + c->px_p_off[0]);
		c->Pnext = reinterpret_cast<float *>(base + c->px_p_off[1]);
		c->p_par = 0;
		c->pp_seq = 0;
		c->p_peer = true;
	}
	return IG_OK;
}

extern "C" ig_status ig_comm_init(ig_ctx *c, const void *id128)
{
	if (!c || !id128) return fail(IG_ERR_ARG, "null argument");
	// IG_COMM_SINGLE=1: a one-rank communicator, so that the whole sharded code path (side-stream P draw, local scalar
	// updates, gathers) can be exercised -- and compared with the plain path -- on one GPU (tests/test_gpu_local_scalars.py)
	if (c->cfg.shard_count <= 1 && !getenv_once("IG_COMM_SINGLE")) return IG_OK;
	ig_status st = nccl_load();
	if (st != IG_OK) return st;
	CK(cudaSetDevice(c->cfg.device));
	ig_alloc_stream = c->stream;
	ncclUniqueId id;
	memcpy(&id, id128, 128);
	NCK(g_nccl.CommInitRank(&c->comm, c->cfg.shard_count > 1 ? c->cfg.shard_count : 1, id, c->cfg.shard_count > 1 ? c->cfg.shard_rank : 0));
	// (measured at 2 GPUs: no difference -- the all-gather does not queue behind the tally all-reduce; off by default)
	if (g_nccl.CommSplit && getenv("IG_TWO_COMMS")) {
		if (g_nccl.CommSplit(c->comm, 0, c->cfg.shard_rank, &c->comm2, nullptr) != ncclSuccess) c->comm2 = nullptr;
	}
	if (!c->tetra && c->cfg.mode != 0 && !c->cfg.print_freq && c->loaded && !c->stream2) {
		int prio_least = 0, prio_greatest = 0;
		CK(cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
		CK(cudaStreamCreateWithPriority(&c->stream2, cudaStreamNonBlocking, prio_least));
		CK(cudaEventCreateWithFlags(&c->ev_zq, cudaEventDisableTiming));
		CK(cudaEventCreateWithFlags(&c->ev_p, cudaEventDisableTiming));
		const int Wl = c->cfg.shard_count > 1 ? c->cfg.shard_count : 1;
		CK(dalloc(&c->Pnext, (size_t)Wl * c->p_lr * c->geo.A * c->geo.KP));
		if (!c->geo.snp && g_nccl.ReduceScatter && !getenv_once("IG_P_ALLREDUCE")) CK(dalloc(&c->nred, (size_t)c->p_lr * c->geo.A * c->geo.KP));
		if (c->geo.snp) CK(dalloc(&c->Pcnext, snp_pc_floats(c->geo)));
	}
	// modes 1 and 2 keep the records local (IG_GATHER_RECORDS=1: the all-gather path of round 1, kept for comparison)
	c->local_scalars = !c->tetra && c->loaded && (c->cfg.mode == 1 || c->cfg.mode == 2) && c->geo.fmode == 0 &&
	                   c->geo.K <= LOCAL_MAX_K && !getenv_once("IG_GATHER_RECORDS");
	// ... and all-reduce their sums over peer memory (IG_NCCL_SCALARS=1: through ncclAllReduce, for comparison)
	if (c->local_scalars && !c->px_buf && !getenv_once("IG_NCCL_SCALARS")) return peer_setup(c);
	return IG_OK;
}

ig_status ig_allreduce_int64(ig_ctx *c, int64_t *buf, size_t count)
{
	if (!c->comm) return IG_OK;
	NCK(g_nccl.AllReduce(buf, buf, count, ncclInt64, ncclSum, c->comm, c->stream));
	return IG_OK;
}

static ig_status exchange_tally(ig_ctx *c);
static ig_status exchange_individuals(ig_ctx *c);
ig_status ig_exchange_tally(ig_ctx *c) { return exchange_tally(c); }
ig_status ig_exchange_individuals(ig_ctx *c) { return exchange_individuals(c); }
// all-gather of `per_rank` doubles per shard, in place (rank r's block at buf + r * per_rank)
ig_status ig_allgather_double(ig_ctx *c, double *buf, size_t per_rank)
{
	if (!c->comm) return IG_OK;
	NCK(g_nccl.AllGather(buf + (size_t)c->cfg.shard_rank * per_rank, buf, per_rank, ncclDouble, c->comm, c->stream));
	return IG_OK;
}

ig_status ig_allreduce_int32(ig_ctx *c, int32_t *buf, size_t count)
{
	if (!c->comm) return IG_OK;
	NCK(g_nccl.AllReduce(buf, buf, count, ncclInt32, ncclSum, c->comm, c->stream));
	return IG_OK;
}

static ig_status exchange_tally(ig_ctx *c)
{
	// the one per-sweep reduction of the sharded mode: int32 n[L][A][K] summed over the shards
	// (integer => identical on every rank and for every shard count)
	if (!c->comm) return IG_OK;
	const Geometry &g = c->geo;
	NCK(g_nccl.AllReduce(c->n, c->n, (size_t)g.Lpad * g.A * g.KP, ncclInt32, ncclSum, c->comm, c->stream));
	return IG_OK;
}
static ig_status exchange_individuals(ig_ctx *c)
{
	// (Q, indvlkh, sum log q, G) of every shard, so that the O(N*K) scalar updates run
	// redundantly and identically on every rank
	if (!c->comm) return IG_OK;
	const Geometry &g = c->geo;
	const size_t cnt = (size_t)c->shard_cap * g.REC;
	NCK(g_nccl.AllGather(c->ind + (size_t)g.i0 * g.REC, c->ind, cnt, ncclDouble, c->comm, c->stream));
	return IG_OK;
}

// A sharded chain with local scalar updates leaves other ranks' records stale between sweeps; whatever reads or edits the
// records from outside a sweep (state hooks, phase hooks, stand-alone evaluators) gathers them first.  Collective: in that
// mode every rank has to make the same hook calls (they already must for the phase hooks).
static ig_status ensure_records(ig_ctx *c)
{
	c->tree_ready = false;
	if (!c->comm || !c->ind_stale) return IG_OK;
	c->ind_stale = false;
	return exchange_individuals(c);
}

// --------------------------------------------------------------------------------------
// Dirichlet-process prior on the individual selfing rates -- host side (DPMM.c:124-398).
// The step is a sequential Chinese-restaurant Gibbs scan over individuals; it reads only
// G[0..N) and writes only S[0..N), i.e. N ints down and N doubles up per sweep.  Clusters
// live in a value-sorted index-linked pool.  Randomness: Philox site (j, 0, iter, TAG_DP).
// --------------------------------------------------------------------------------------
static void dp_reset(ig_ctx *c)
{
	const int N = c->geo.N;
	c->dp.assign(N + 1, DpCluster{0.0, 0, -1});
	for (int s = 0; s <= N; s++) c->dp[s].next = s + 1;
	c->dp[N].next = -1;
	c->dp_free = 0; c->dp_head = -1; c->dp_cnt = 0;
	c->dp_of.assign(N, -1);
}
static int dp_create(ig_ctx *c, double v)
{
	const int s = c->dp_free;
	c->dp_free = c->dp[s].next;
	c->dp[s].value = v; c->dp[s].num = 1;
	{	// dgeom(v, g) = v^(g-1) (1-v) for g = 1..50 (mcmc.c:1602), once per cluster instead of a pow() per
		// (individual, cluster) in the Gibbs scan
		if (c->dp_w.size() < c->dp.size() * 51) c->dp_w.resize(c->dp.size() * 51);
		double *w = &c->dp_w[(size_t)s * 51];
		double pw = 1.0 - v;
		w[0] = 0.0;
		for (int gg = 1; gg <= 50; gg++) { w[gg] = pw; pw *= v; }
	}
	if (c->dp_head < 0 || v <= c->dp[c->dp_head].value) { c->dp[s].next = c->dp_head; c->dp_head = s; return s; }
	int q = c->dp_head, p = c->dp[q].next;
	while (p >= 0 && c->dp[p].value <= v) { q = p; p = c->dp[p].next; }
	c->dp[s].next = p;
	c->dp[q].next = s;
	return s;
}
static inline double dp_dgeom(const ig_ctx *c, int p, int gen)     // dgeom(value_p, gen), mcmc.c:1596-1604
{
	if (gen >= 1 && gen <= 50) return c->dp_w[(size_t)p * 51 + gen];
	return pow(c->dp[p].value, (double)(gen - 1)) * (1.0 - c->dp[p].value);
}
// clusters = groups of equal values (state injection: ig_set_state(IG_STATE_S) in mode 3 with the DP prior)
static int dp_create(ig_ctx *c, double v);
static void dp_from_values(ig_ctx *c, const double *S)
{
	const int N = c->geo.N;
	dp_reset(c);
	c->S_h.assign(S, S + N);
	for (int j = 0; j < N; j++) {
		int p = c->dp_head;
		while (p >= 0 && c->dp[p].value != S[j]) p = c->dp[p].next;
		if (p >= 0) { c->dp[p].num++; c->dp_of[j] = p; }
		else { c->dp_of[j] = dp_create(c, S[j]); c->dp_cnt++; }
	}
}
static void dp_leave(ig_ctx *c, int j)
{
	const int s = c->dp_of[j];
	if (--c->dp[s].num > 0) return;
	if (c->dp_head == s) c->dp_head = c->dp[s].next;
	else {
		int q = c->dp_head;
		while (c->dp[q].next != s) q = c->dp[q].next;
		c->dp[q].next = c->dp[s].next;
	}
	c->dp[s].next = c->dp_free;
	c->dp_free = s;
	c->dp_cnt--;
}
// pick an index in [0, n) with probability proportional to w[i] (disc_unif, random.c:403)
static int pick_weighted(const std::vector<double> &cum, int n, double u)
{
	const double t = u * cum[n - 1];
	int k = 0;
	while (k < n - 1 && t > cum[k]) k++;
	return k;
}
static void dp_init(ig_ctx *c)                         // init_DP, DPMM.c:124-161
{
	const int N = c->geo.N;
	const double a = c->cfg.alpha_dpm;
	dp_reset(c);
	c->S_h.assign(N, 0.0);
	std::vector<double> cum(N + 1);
	for (int j = 0; j < N; j++) {
		Stream st((uint32_t)j, 0u, 0u, TAG_DP, c->key0, c->key1);
		cum[0] = a / (a + j);
		int n = 1;
		for (int p = c->dp_head; p >= 0; p = c->dp[p].next, n++) cum[n] = cum[n - 1] + c->dp[p].num / (a + j);
		const int pick = pick_weighted(cum, n, st.uniform());
		if (pick == 0) { c->S_h[j] = st.uniform(); c->dp_of[j] = dp_create(c, c->S_h[j]); c->dp_cnt++; }
		else {
			int p = c->dp_head;
			for (int k = 1; k < pick; k++) p = c->dp[p].next;
			c->dp[p].num++; c->dp_of[j] = p; c->S_h[j] = c->dp[p].value;
		}
	}
}
// The scan is sequential over the individuals (DPMM.c:175-197) and sits on the host's critical path every sweep; one
// Philox4x32-10 block per individual was a third of it.  The scan's uniforms come from xoshiro256** seeded per sweep by
// one Philox block of (sweep, TAG_DP) -- still a pure function of (seed, chain, sweep), still the same on every rank; the
// (rare) Beta(G, 2) draw of a new cluster keeps its own Philox stream keyed by the individual.
struct Xoshiro {
	uint64_t s[4];
	static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
	inline uint64_t next()
	{
		const uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
		s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
		return r;
	}
	inline double uniform() { return ((double)(next() >> 11) + 0.5) * (1.0 / 9007199254740992.0); }
};
static void dp_update(ig_ctx *c, const uint8_t *gen8)   // update_DP, DPMM.c:165-199; gen8[j] = G_j
{
	const Geometry &g = c->geo;
	const int N = g.N;
	std::vector<double> &cum = c->dp_cum;
	if ((int)cum.size() < N + 1) cum.resize(N + 1);
	Xoshiro rng;
	{
		const u32x4 a = philox4x32<10>(u32x4{0u, 0u, c->iter, TAG_DP}, c->key0, c->key1), b = philox4x32<10>(u32x4{1u, 0u, c->iter, TAG_DP}, c->key0, c->key1);
		rng.s[0] = ((uint64_t)a.x << 32) | a.y; rng.s[1] = ((uint64_t)a.z << 32) | a.w;
		rng.s[2] = ((uint64_t)b.x << 32) | b.y; rng.s[3] = (((uint64_t)b.z << 32) | b.w) | 1u;
		for (int w = 0; w < 8; w++) rng.next();
	}
	// T[g] = sum over clusters of num_c * dgeom(S_c, g), kept current as individuals leave and join (two 50-element vector
	// updates per individual), so the total weight of an individual with G = g is alpha/((g+1)g) + T[g] without a pass over
	// the clusters; the pick then walks the value-ordered list only as far as the chosen cluster.  Rebuilt from scratch at
	// the start of every scan, so rounding cannot drift; if rounding leaves the threshold beyond the last cumulative weight,
	// the last cluster is taken (as disc_unif does, random.c:425-429).
	double T[51];
	for (int gg = 0; gg <= 50; gg++) T[gg] = 0.0;
	for (int p = c->dp_head; p >= 0; p = c->dp[p].next) {
		const double *w = &c->dp_w[(size_t)p * 51];
		const double num = c->dp[p].num;
		for (int gg = 1; gg <= 50; gg++) T[gg] += num * w[gg];
	}
	for (int j = 0; j < N; j++) {
		const int gen = (int)gen8[j];
		{
			const double *w = &c->dp_w[(size_t)c->dp_of[j] * 51];
			for (int gg = 1; gg <= 50; gg++) T[gg] -= w[gg];
		}
		dp_leave(c, j);
		const double w0 = c->cfg.alpha_dpm / (gen + 1) / gen;         // gen_post_prob, DPMM.c:369: a new cluster
		int p = -1;
		const double u = rng.uniform();
		if (gen >= 1 && gen <= 50) {
			// num * dgeom(value, G_j), DPMM.c:373, from the per-cluster tables
			const double t = u * (w0 + T[gen]);
			if (t > w0) {
				double cum = w0;
				int last = -1;
				for (int q = c->dp_head; q >= 0; q = c->dp[q].next) {
					cum += c->dp[q].num * c->dp_w[(size_t)q * 51 + gen];
					last = q;
					if (t <= cum) break;
				}
				p = last;                                             // none left (one individual, no cluster): p stays -1 -> new cluster
			}
		} else {
			// the reference's mode-3 initial G is not capped (mcmc.c:329-331): a G beyond the tables is evaluated directly
			double tot = w0;
			for (int q = c->dp_head; q >= 0; q = c->dp[q].next) tot += c->dp[q].num * dp_dgeom(c, q, gen);
			const double t = u * tot;
			if (t > w0) {
				double cum = w0;
				for (int q = c->dp_head; q >= 0; q = c->dp[q].next) { cum += c->dp[q].num * dp_dgeom(c, q, gen); p = q; if (t <= cum) break; }
			}
		}
		if (p < 0) {                                                  // sample_poster, DPMM.c:395: Beta(G, 2)
			Stream st((uint32_t)j, 1u, c->iter, TAG_DP, c->key0, c->key1);
			c->S_h[j] = draw_beta(st, (double)gen, 2.0);
			p = dp_create(c, c->S_h[j]);
			c->dp_of[j] = p;
			c->dp_cnt++;
		} else {
			c->dp[p].num++; c->dp_of[j] = p; c->S_h[j] = c->dp[p].value;
		}
		const double *w = &c->dp_w[(size_t)p * 51];
		for (int gg = 1; gg <= 50; gg++) T[gg] += w[gg];
	}
}

// int64 all-reduce of fx[off .. off + n) on the main stream: peer-memory kernel (optionally with the post-sweep tail fused) or NCCL
static ig_status allreduce_fx(ig_ctx *c, size_t off, size_t n, const PostArgs *tail)
{
	if (c->px) {
		PeerArgs x{c->px_peers, c->cfg.shard_count > 1 ? c->cfg.shard_count : 1, c->cfg.shard_count > 1 ? c->cfg.shard_rank : 0,
		           ++c->px_seq, (int)n, c->fx + off, c->fxg + off, tail ? 1 : 0};
		PostArgs none{};
		CK(launch_peer_allreduce(x, tail ? *tail : none, c->stream));
		c->launches++;
		return IG_OK;
	}
	NCK(g_nccl.AllReduce(c->fx + off, c->fxg + off, n, ncclInt64, ncclSum, c->comm, c->stream));
	CK(cudaMemsetAsync(c->fx + off, 0, n * sizeof(unsigned long long), c->stream));       // the accumulators go back empty
	if (tail) { CK(launch_post_final(*tail, c->fxg, c->stream)); c->launches++; }
	return IG_OK;
}

// --------------------------------------------------------------------------------------
// sweep phases
// --------------------------------------------------------------------------------------
static inline void ptrace_mark(ig_ctx *c, int point)
{
	if (c->ptrace.empty() || c->ptrace_sweeps >= PT_SWEEPS || c->iter_dev) return;
	cudaEventRecord(c->ptrace[(size_t)c->ptrace_sweeps * PT_POINTS + point], c->stream);
}
static ZQArgs zq_args(ig_ctx *c)
{
	ZQArgs a;
	a.Xt = c->Xt; a.Zt = c->Zt; a.P = c->P; a.n = c->n; a.Qf = c->Qf; a.gpair = c->gpair;
	a.pcnt = c->pcnt; a.plog = c->plog; a.pnsh = c->pnsh; a.geo = c->geo; a.iter = c->iter; a.key0 = c->key0; a.key1 = c->key1;
	a.type_freq = c->cfg.type_freq;
	a.iter_dev = c->iter_dev;
	a.fmode = c->geo.fmode;
	a.hpair = (c->geo.fmode == 1) ? c->hpair : nullptr;
	a.ftab = c->ftab; a.pfk = c->pfk;
	a.k_mant = U16_MANT; a.k_one = U16_ONE;         // 16 random bits per allele copy (sweep_common.cuh)
	return a;
}

static ig_status phase_update_P(ig_ctx *c)
{
	if (c->early_p) {
		// drawn behind the previous sweep's kernel on the side stream.  Only zq_sweep reads P: the main stream waits for it
		// there (phase_zq), so update_S / the G proposals of this sweep still overlap the all-reduce and the draw (waiting
		// here cost 46 us per sweep at 1250 individuals per GPU)
		std::swap(c->P, c->Pnext);
		c->p_par ^= 1;
		std::swap(c->Pc, c->Pcnext);
		c->early_p = false;
		c->p_wait = true;
		return IG_OK;
	}
	ig_status st = exchange_tally(c);
	if (st != IG_OK) return st;
	PArgs a{c->n, c->P, c->P64, c->allelenum, c->geo, c->iter, c->key0, c->key1, c->iter_dev, 0, 0, c->Pc, c->geo.TL};
	CK(launch_p_dirichlet(a, c->stream));
	c->launches++;
	return IG_OK;
}

static ig_status phase_update_S(ig_ctx *c)
{
	const Geometry &g = c->geo;
	// mode 1 (mcmc_POP_admixture, mcmc.c:135-180) has neither update_S nor update_G: the (1, 1)
	// generation pairs written by init_chain stay, and with G == 1 the sweep's likelihood is
	// log_ld_noselfing_indv (mcmc.c:1869)
	if (c->cfg.mode <= 1) return IG_OK;
	if (c->cfg.mode == 3 && c->cfg.prior_flag == 1) {
		if (!c->g8_dev) {
			ig_alloc_stream = c->stream;
			CK(dalloc(&c->g8_dev, (size_t)g.N));
			CK(cudaHostAlloc((void **)&c->g8_host, (size_t)g.N, cudaHostAllocDefault));
			CK(cudaHostAlloc((void **)&c->S_pin, (size_t)g.N * sizeof(double), cudaHostAllocDefault));
			CK(cudaEventCreateWithFlags(&c->ev_g, cudaEventDisableTiming));
		}
		if (!c->g8_inflight) {                               // first sweep, or a hook touched the state since the last one
			CK(launch_pack_g(c->ind, c->g8_dev, g, c->stream));
			CK(cudaMemcpyAsync(c->g8_host, c->g8_dev, (size_t)g.N, cudaMemcpyDeviceToHost, c->stream));
			CK(cudaEventRecord(c->ev_g, c->stream));
			c->launches++;
		}
		const double tw0 = c->ptrace.empty() ? 0.0 : wall_ms();
		CK(cudaEventSynchronize(c->ev_g));
		c->g8_inflight = false;
		const double tw1 = c->ptrace.empty() ? 0.0 : wall_ms();
		dp_update(c, c->g8_host);
		if (!c->ptrace.empty()) { c->ptrace_host_ms[5] += tw1 - tw0; c->ptrace_host_ms[6] += wall_ms() - tw1; }
		// the previous sweep's copy out of S_pin has long completed: pre_sweep consumed S before the epilogue whose G we just read
		memcpy(c->S_pin, c->S_h.data(), (size_t)g.N * sizeof(double));
		CK(cudaMemcpyAsync(c->S, c->S_pin, (size_t)g.N * sizeof(double), cudaMemcpyHostToDevice, c->stream));
	}
	if (c->local_now) {
		// update_S_POP + the G proposals on the local records: subset sums (unless they were taken behind the previous
		// sweep's post sums and travelled with that all-reduce), one int64 all-reduce, the K decisions
		TreeArgs t{c->ind, c->S, c->state, c->geo, c->iter, c->key0, c->key1, c->cfg.back_refl, c->fx + FX_POST,
		           c->S2, c->state2, c->gprop, c->gpair, c->sc};
		if (!c->tree_ready) {
			CK(launch_local_sums(&t, nullptr, nullptr, c->stream));
			c->launches++;
			ig_status sta = allreduce_fx(c, FX_POST, 2 * ((size_t)1 << g.K), nullptr);
			if (sta != IG_OK) return sta;
		}
		t.acc = c->fxg + FX_POST;                            // the decisions read the all-reduced table
		c->tree_ready = false;
		CK(launch_spop_decide(t, c->stream));
		std::swap(c->S, c->S2);
		if (c->cfg.back_refl == 0) std::swap(c->state, c->state2);
		c->launches++;
		return IG_OK;
	}
	// UPMCMC.state is read by every CTA and written by one: double-buffered
	PreArgs a{c->ind, c->S, c->state, c->state2, c->gprop, c->gpair, c->sc, c->gpart, c->geo, c->iter, c->key0, c->key1,
	          c->cfg.mode, c->cfg.prior_flag, c->cfg.back_refl, c->iter_dev, c->fprop, c->hpair, c->ftab, c->grid_pre};
	CK(launch_pre_sweep(a, c->stream));
	if (c->cfg.mode == 2 && c->cfg.back_refl == 0) std::swap(c->state, c->state2);
	c->launches++;
	return IG_OK;
}

static ig_status phase_zq(ig_ctx *c, int init)
{
	if (c->cfg.mode == 0) return init ? na_chain_init(c) : na_phase_z(c);     // no admixture: whole-individual labels (noadmix.cu)
	if (c->p_wait) { CK(cudaStreamWaitEvent(c->stream, c->ev_p, 0)); c->p_wait = false; }
	ZQArgs a = zq_args(c);
	if (init) { a.type_freq = 1; a.fmode = 0; a.hpair = nullptr; }     // uniform initial assignment: no likelihood is kept
	const bool timed = c->profile && !init && c->ev_used + 2 <= (int)c->ev.size();
	if (timed) CK(cudaEventRecord(c->ev[c->ev_used], c->stream));
	if (c->geo.snp) {
		SnpArgs sa{c->Es, c->Zs, c->Hs, c->Pc, c->n, c->Qf, c->gpair, c->pcnt, c->plog, c->pnsh, c->geo, c->iter, c->key0, c->key1,
		           c->iter_dev, U16_MANT, U16_ONE};
		CK(launch_zq_snp(sa, c->rounds, c->stream));
		c->zt_stale = true;
	} else CK(launch_zq_sweep(a, c->rounds, c->stream));
	if (timed) { CK(cudaEventRecord(c->ev[c->ev_used + 1], c->stream)); c->ev_used += 2; }
	if (!init) ptrace_mark(c, 3);
	if (!init && c->comm && c->stream2 && c->more_follow) {
		// Sharded chain: n is complete once this kernel ends, and nothing else of this sweep reads it.
		// Its all-reduce and the P draw of the NEXT sweep (same Philox keys, iter + 1: identical on every
		// rank) go to the side stream, overlapping the epilogue, the all-gather and the scalar updates.
		const Geometry &g = c->geo;
		CK(cudaEventRecord(c->ev_zq, c->stream));
		CK(cudaStreamWaitEvent(c->stream2, c->ev_zq, 0));
		ncclComm_t cm = c->comm2 ? c->comm2 : c->comm;
		if (c->p_peer) {
			// the whole tally -> P exchange over peer memory, no NCCL: "my tally is final" to every rank and wait for theirs; pull
			// my block of loci from every rank's n, draw its Dirichlets, push the block into every rank's next-P buffer; "my
			// block is in your buffer" and wait for theirs -- which also tells me every rank is done reading my n: clear it
			const int W = c->cfg.shard_count > 1 ? c->cfg.shard_count : 1, me = W > 1 ? c->cfg.shard_rank : 0;
			PeerPArgs x{c->px_peers, W, me, ++c->pp_seq, c->px_n_off, c->px_p_off[c->p_par ^ 1]};
			PArgs pa{c->n, c->Pnext, nullptr, c->allelenum, c->geo, c->iter + 1, c->key0, c->key1, nullptr, 0, 0,
			         nullptr, c->geo.TL, me * c->p_lr, c->p_lr};
			CK(launch_p_peer_signal(x, 0, c->stream2));
			CK(launch_p_peer_draw(pa, x, c->stream2));
			CK(launch_p_peer_signal(x, 1, c->stream2));
			CK(cudaMemsetAsync(c->n, 0, (size_t)g.Lpad * g.A * g.KP * sizeof(int32_t), c->stream2));
			c->launches += 2;
		} else if (c->nred) {
			// every rank needs all of P but not all of n: the tally is reduce-SCATTERED by blocks of loci, each rank draws the
			// Dirichlets of its block (1/W of the 43 us the full draw takes at config 4) and P is all-gathered -- the same
			// bytes on the wire as the all-reduce, W times less double-precision work behind it
			const int W = c->cfg.shard_count > 1 ? c->cfg.shard_count : 1, me = W > 1 ? c->cfg.shard_rank : 0;
			const size_t cnt = (size_t)c->p_lr * g.A * g.KP;
			NCK(g_nccl.ReduceScatter(c->n, c->nred, cnt, ncclInt32, ncclSum, cm, c->stream2));
			CK(cudaMemsetAsync(c->n, 0, (size_t)g.Lpad * g.A * g.KP * sizeof(int32_t), c->stream2));     // p_dirichlet used to clear it
			PArgs pa{c->nred - (size_t)me * cnt, c->Pnext, nullptr, c->allelenum, c->geo, c->iter + 1, c->key0, c->key1, nullptr, 0, 0,
			         nullptr, c->geo.TL, me * c->p_lr, c->p_lr};
			CK(launch_p_dirichlet(pa, c->stream2));
			NCK(g_nccl.AllGather(c->Pnext + (size_t)me * cnt, c->Pnext, cnt, ncclFloat, cm, c->stream2));
		} else {
			NCK(g_nccl.AllReduce(c->n, c->n, (size_t)g.Lpad * g.A * g.KP, ncclInt32, ncclSum, cm, c->stream2));
			PArgs pa{c->n, c->Pnext, c->P64, c->allelenum, c->geo, c->iter + 1, c->key0, c->key1, nullptr, 0, 0, c->Pcnext, c->geo.TL};
			CK(launch_p_dirichlet(pa, c->stream2));
		}
		CK(cudaEventRecord(c->ev_p, c->stream2));
		c->launches++;
		c->early_p = true;
	}
	EpiArgs e{c->pcnt, c->plog, c->pnsh, c->nhet, c->nsh, c->ind, c->Qf, c->cnt, c->llparts, c->gpair, c->sc, c->geo,
	          c->iter, c->key0, c->key1, init, a.type_freq, init ? nullptr : c->iter_dev, c->geo.fmode, c->S, c->fprop, c->pfk};
	CK(launch_epilogue(e, c->stream));
	c->launches += 2;
	if (c->geo.fmode == 2 && !init) { CK(launch_fk_epilogue(e, c->stream)); c->launches++; }
	if (!init) ptrace_mark(c, 4);
	if (c->local_now) c->ind_stale = true;                  // the records stay on their rank
	else {
		ig_status stx = exchange_individuals(c);
		if (stx != IG_OK) return stx;
	}
	if (c->cfg.mode == 3 && c->cfg.prior_flag == 1 && c->g8_dev && c->more_follow && !c->iter_dev) {
		// the next sweep's Dirichlet-process scan needs this G: send it now, the host reads it while post_sweep and the
		// next p_dirichlet run
		CK(launch_pack_g(c->ind, c->g8_dev, c->geo, c->stream));
		CK(cudaMemcpyAsync(c->g8_host, c->g8_dev, (size_t)c->geo.N, cudaMemcpyDeviceToHost, c->stream));
		CK(cudaEventRecord(c->ev_g, c->stream));
		c->launches++;
		c->g8_inflight = true;
	}
	return IG_OK;
}

static ig_status phase_alpha(ig_ctx *c)
{
	// mode 4: pre_sweep left the proposed adaptive-independence states in state2 (-e 0)
	PostArgs a{c->ind, c->sc, c->gpart, c->geo, c->iter, c->key0, c->key1, c->iter_dev, c->cfg.mode, c->cfg.back_refl,
	           c->S, c->fprop, c->state, c->state2, c->grid_post};
	if (c->local_now) {
		// local sums -> all-reduce -> tail.  When another sweep follows, ITS subset sums of update_S_POP are taken now (S, Q
		// and G are final for this sweep; the proposals are a function of S and the next sweep's Philox streams) and ride
		// the same all-reduce: one exchange of a few KB per sweep.
		const Geometry &g = c->geo;
		const bool ahead = c->more_follow && c->cfg.mode == 2 && !getenv_once("IG_NO_TREE_AHEAD");
		const size_t nt = ahead ? 2 * ((size_t)1 << g.K) : 0;
		TreeArgs t{c->ind, c->S, c->state, c->geo, c->iter + 1, c->key0, c->key1, c->cfg.back_refl, c->fx + FX_POST,
		           nullptr, nullptr, nullptr, nullptr, nullptr};
		CK(launch_local_sums(ahead ? &t : nullptr, &a, c->fx, c->stream));
		ptrace_mark(c, 6);
		ig_status sta = allreduce_fx(c, 0, ahead ? FX_POST + nt : (size_t)(g.K + 4), &a);
		if (sta != IG_OK) return sta;
		c->tree_ready = ahead;
		c->launches++;
		return IG_OK;
	}
	ptrace_mark(c, 6);
	CK(launch_post_sweep(a, c->stream));
	c->launches++;
	return IG_OK;
}

static void ptrace_report(ig_ctx *c)
{
	if (c->ptrace.empty() || c->ptrace_sweeps < 8) return;
	static const char *names[PT_POINTS - 1] = {"update_P (wait / draw)", "update_S + G proposals", "zq_sweep", "epilogue", "all-gather", "local sums", "exchange + update_alpha"};
	double acc[PT_POINTS - 1] = {0, 0, 0, 0, 0, 0, 0}, between = 0;
	std::vector<float> gaps;             // between consecutive sweeps: the median (the caller's own pauses -- a barrier, a sync -- are not the sweep's)
	int n = 0;
	for (int s = 4; s < c->ptrace_sweeps; s++, n++) {            // skip the first sweeps
		for (int p = 0; p + 1 < PT_POINTS; p++) {
			float ms = 0.f;
			cudaEventElapsedTime(&ms, c->ptrace[(size_t)s * PT_POINTS + p], c->ptrace[(size_t)s * PT_POINTS + p + 1]);
			acc[p] += ms;
		}
		if (s + 1 < c->ptrace_sweeps) { float ms = 0.f; cudaEventElapsedTime(&ms, c->ptrace[(size_t)s * PT_POINTS + PT_POINTS - 1], c->ptrace[(size_t)(s + 1) * PT_POINTS]); gaps.push_back(ms); }
	}
	fprintf(stderr, "[ig_phase_trace] rank %d, %d sweeps, us per sweep:", c->cfg.shard_rank, n);
	for (int p = 0; p + 1 < PT_POINTS; p++) fprintf(stderr, " %s %.1f |", names[p], 1e3 * acc[p] / n);
	if (!gaps.empty()) { std::sort(gaps.begin(), gaps.end()); between = gaps[gaps.size() / 2]; }
	fprintf(stderr, " between sweeps (median) %.1f || host enqueue us per sweep:", 1e3 * between);
	for (int p = 0; p < 4; p++) fprintf(stderr, " %.1f", 1e3 * c->ptrace_host_ms[p] / (c->ptrace_host_ms[4] > 0 ? c->ptrace_host_ms[4] : 1.0));
	if (c->ptrace_host_ms[6] > 0) fprintf(stderr, " (DP prior: wait for G %.1f, scan %.1f)", 1e3 * c->ptrace_host_ms[5] / c->ptrace_host_ms[4], 1e3 * c->ptrace_host_ms[6] / c->ptrace_host_ms[4]);
	fprintf(stderr, "\n");
}

static ig_status one_sweep_direct(ig_ctx *c)
{
	ig_status st;
	if (c->tetra) return tetra_one_sweep(c);
	c->iter++;
	c->local_now = c->comm && c->local_scalars;
	if (!c->local_now) c->tree_ready = false;
	struct LocalGuard { ig_ctx *c; ~LocalGuard() { c->local_now = false; } } local_guard{c};
	const bool tr = !c->ptrace.empty();
	double h0 = tr ? wall_ms() : 0.0, h1;
	ptrace_mark(c, 0);
	if ((st = phase_update_P(c)) != IG_OK) return st;      // update_P            mcmc.c:210
	ptrace_mark(c, 1);
	if (tr) { h1 = wall_ms(); c->ptrace_host_ms[0] += h1 - h0; h0 = h1; }
	if ((st = phase_update_S(c)) != IG_OK) return st;      // update_S_* + G proposal  :211-212
	ptrace_mark(c, 2);
	if (tr) { h1 = wall_ms(); c->ptrace_host_ms[1] += h1 - h0; h0 = h1; }
	if ((st = phase_zq(c, 0)) != IG_OK) return st;         // update_G accept, update_ZQ, cal_lkh :212-215
	ptrace_mark(c, 5);
	if (tr) { h1 = wall_ms(); c->ptrace_host_ms[2] += h1 - h0; h0 = h1; }
	st = phase_alpha(c);                                   // update_alpha, totallkh    :214-215
	ptrace_mark(c, 7);
	if (tr) { h1 = wall_ms(); c->ptrace_host_ms[3] += h1 - h0; c->ptrace_host_ms[4] += 1.0; }
	if (!c->ptrace.empty() && c->ptrace_sweeps < PT_SWEEPS && !c->iter_dev) c->ptrace_sweeps++;
	return st;
}

// A sweep of the small configurations is five launches of a few microseconds each: launch latency,
// not the GPU, sets the rate.  Where every kernel argument is constant from sweep to sweep (no host
// step, no buffer ping-pong, no communicator) the sweep is captured ONCE into a CUDA graph; the
// sweep counter that keys the RNG is then read from device memory (DevScalars.iter, advanced by
// post_sweep), and a replay is bit-identical to the direct launches.
static bool graph_eligible(const ig_ctx *c)
{
	if (c->tetra || c->comm || c->profile || c->cfg.use_graph == 2) return false;
	if (c->cfg.mode == 3 && c->cfg.prior_flag == 1) return false;        // host Dirichlet-process step every sweep
	if (c->cfg.mode == 2 && c->cfg.back_refl == 0) return false;         // UPMCMC.state is double-buffered
	return true;
}

static ig_status one_sweep(ig_ctx *c)
{
	if (c->graph_failed || !graph_eligible(c)) return one_sweep_direct(c);
	if (!c->graph_exec) {
		cudaGraph_t graph = nullptr;
		const uint32_t it0 = c->iter;
		const int64_t l0 = c->launches;
		// the device copy of the counter is kept current by post_sweep and ig_set_state; make sure
		const uint32_t next = c->iter + 1;
		CK(cudaMemcpyAsync(&c->sc->iter, &next, sizeof(next), cudaMemcpyHostToDevice, c->stream));
		CK(cudaStreamSynchronize(c->stream));
		c->iter_dev = &c->sc->iter;
		c->dev_iter_valid = true;
		cudaError_t e = cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal);
		ig_status st = IG_OK;
		if (e == cudaSuccess) {
			st = one_sweep_direct(c);
			e = cudaStreamEndCapture(c->stream, &graph);
		}
		c->launches_per_sweep = c->launches - l0;
		c->iter = it0;
		c->launches = l0;
		if (e == cudaSuccess && st == IG_OK) e = cudaGraphInstantiate(&c->graph_exec, graph, 0);
		if (graph) cudaGraphDestroy(graph);
		if (e != cudaSuccess || st != IG_OK || !c->graph_exec) {             // not capturable here: keep launching directly
			cudaGetLastError();
			c->graph_failed = true;
			c->graph_exec = nullptr;
			c->iter_dev = nullptr;
			return one_sweep_direct(c);
		}
		c->iter_dev = nullptr;
	}
	if (!c->dev_iter_valid) {                                               // a phase hook or set_state ran in between
		const uint32_t next = c->iter + 1;
		CK(cudaMemcpyAsync(&c->sc->iter, &next, sizeof(next), cudaMemcpyHostToDevice, c->stream));
		CK(cudaStreamSynchronize(c->stream));
		c->dev_iter_valid = true;
	}
	CK(cudaGraphLaunch(c->graph_exec, c->stream));
	c->iter++;
	c->launches += c->launches_per_sweep;
	return IG_OK;
}

extern "C" ig_status ig_chain_init(ig_ctx *c, int32_t chain_id, const float *initd)
{
	if (!c) return fail(IG_ERR_ARG, "null context");
	if (!c->loaded) return fail(IG_ERR_STATE, "load genotypes first");
	CK(cudaSetDevice(c->cfg.device));
	if (c->tetra) return tetra_chain_init(c, chain_id, initd);
	const Geometry &g = c->geo;
	c->key0 = (uint32_t)c->cfg.seed ^ (0x9E3779B9u * (uint32_t)(chain_id + 1));
	c->key1 = (uint32_t)(c->cfg.seed >> 32) ^ (0x85EBCA6Bu * (uint32_t)(chain_id + 1));
	c->iter = 0;
	c->g8_inflight = false;
	if (c->stream2) CK(cudaStreamSynchronize(c->stream2));
	c->early_p = false;
	c->p_wait = false;
	c->tree_ready = false;
	c->ind_stale = false;
	if (c->graph_exec) { cudaGraphExecDestroy(c->graph_exec); c->graph_exec = nullptr; }   // the chain's RNG key is baked into the captured arguments
	float init_h[MAX_K];
	for (int k = 0; k < MAX_K; k++) init_h[k] = (initd && k < g.K) ? initd[k] : 0.5f;
	if ((c->cfg.mode == 2 || c->cfg.mode == 4) && !initd) {
		// read_init without an -i file draws the starting rates from U(0,1) (initial.c:52-58)
		Stream st(0u, 2u, 0u, TAG_INIT, c->key0, c->key1);
		for (int k = 0; k < g.K; k++) init_h[k] = (float)st.uniform();
	}
	CK(cudaMemcpyAsync(c->initd_dev, init_h, sizeof(init_h), cudaMemcpyHostToDevice, c->stream));
	if (c->cfg.mode == 3 && c->cfg.prior_flag == 1) {
		dp_init(c);
		CK(cudaMemcpyAsync(c->S, c->S_h.data(), (size_t)g.N * sizeof(double), cudaMemcpyHostToDevice, c->stream));
	}
	CK(launch_init_chain(c->ind, c->S, c->state, c->sc, c->initd_dev, c->gprop, c->gpair, g, c->cfg.mode, c->cfg.prior_flag,
	                     c->cfg.back_refl, c->key0, c->key1, c->stream));
	// initial assignment (update_ZQ with init_flag = 1, mcmc.c:206,1143-1144): uniform Z is the
	// categorical draw with all weights equal, so the sweep kernel runs once with P = 1, Q = 1/K
	const size_t pn = (size_t)g.Lpad * g.A * g.KP;
	if (c->cfg.mode != 0) {
		CK(launch_fill_f32(c->P, 1.0f, pn, c->stream));
		CK(launch_fill_q_uniform(c->Qf, g, c->stream));
		CK(cudaMemsetAsync(c->n, 0, pn * sizeof(int32_t), c->stream));
		if (g.snp) {
			CK(launch_fill_f32(c->Pc, 1.0f, snp_pc_floats(g), c->stream));
			CK(cudaMemsetAsync(c->Zs, 0, (size_t)g.nchunks * g.Nloc * g.TL * sizeof(uint16_t), c->stream));
			c->zt_stale = true;
		} else CK(cudaMemsetAsync(c->Zt, 0, (size_t)g.LT * g.Nloc * TILE * 2, c->stream));
		c->launches += 3;
	}
	ig_status st = phase_zq(c, 1);
	if (st != IG_OK) return st;
	c->dev_iter_valid = false;
	c->chain_ready = true;
	return IG_OK;
}

extern "C" ig_status ig_sweep(ig_ctx *c, int32_t nsweeps)
{
	if (!c) return fail(IG_ERR_ARG, "null context");
	if (!c->chain_ready) return fail(IG_ERR_STATE, "call ig_chain_init first");
	CK(cudaSetDevice(c->cfg.device));
	for (int s = 0; s < nsweeps; s++) {
		c->more_follow = s + 1 < nsweeps;
		ig_status st = one_sweep(c);
		c->more_follow = false;
		if (st != IG_OK) return st;
	}
	return IG_OK;
}

extern "C" ig_status ig_time_sweeps(ig_ctx *c, int32_t nsweeps, double *elapsed_ms)
{
	if (!c || !elapsed_ms) return fail(IG_ERR_ARG, "null argument");
	if (!c->chain_ready) return fail(IG_ERR_STATE, "call ig_chain_init first");
	CK(cudaSetDevice(c->cfg.device));
	cudaEvent_t e0, e1;
	CK(cudaEventCreate(&e0));
	CK(cudaEventCreate(&e1));
	CK(cudaStreamSynchronize(c->stream));
	CK(cudaEventRecord(e0, c->stream));
	ig_status st = IG_OK;
	for (int s = 0; s < nsweeps && st == IG_OK; s++) { c->more_follow = s + 1 < nsweeps; st = one_sweep(c); }
	c->more_follow = false;
	cudaEventRecord(e1, c->stream);
	cudaError_t e = cudaEventSynchronize(e1);
	float ms = 0.f;
	if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
	cudaEventDestroy(e0);
	cudaEventDestroy(e1);
	if (st != IG_OK) return st;
	CK(e);
	*elapsed_ms = ms;
	return IG_OK;
}

extern "C" ig_status ig_sync(ig_ctx *c)
{
	if (!c) return fail(IG_ERR_ARG, "null context");
	CK(cudaSetDevice(c->cfg.device));
	CK(cudaStreamSynchronize(c->stream));
	return IG_OK;
}

extern "C" ig_status ig_run_phase(ig_ctx *c, int32_t mask)
{
	if (!c) return fail(IG_ERR_ARG, "null context");
	if (!c->loaded) return fail(IG_ERR_STATE, "load genotypes first");
	CK(cudaSetDevice(c->cfg.device));
	ig_status st;
	if (c->tetra) return tetra_run_phase(c, mask);
	c->dev_iter_valid = false;
	c->g8_inflight = false;
	if ((st = ensure_records(c)) != IG_OK) return st;
	if (mask & IG_PHASE_UPDATE_P) if ((st = phase_update_P(c)) != IG_OK) return st;
	if (mask & IG_PHASE_UPDATE_S) if ((st = phase_update_S(c)) != IG_OK) return st;
	if (mask & IG_PHASE_ZQ) if ((st = phase_zq(c, 0)) != IG_OK) return st;
	if (mask & IG_PHASE_ALPHA) if ((st = phase_alpha(c)) != IG_OK) return st;
	CK(cudaStreamSynchronize(c->stream));
	return IG_OK;
}

// --------------------------------------------------------------------------------------
// the chain driver
// --------------------------------------------------------------------------------------
static ig_status download_result(ig_ctx *c, ig_chain_result *out, double *convg_ld, long stored)
{
	const Geometry &g = c->geo;
	double tot[2];
	if (c->mom_local) {
		// every rank advanced the moments of its own individuals: one all-gather per array, at the chain's end
		ig_status st;
		const size_t cap = (size_t)c->shard_cap;
		if ((st = ig_allgather_double(c, c->mom.indvlkh, cap)) != IG_OK) return st;
		if ((st = ig_allgather_double(c, c->mom.gen, cap)) != IG_OK) return st;
		if ((st = ig_allgather_double(c, c->mom.gen2, cap)) != IG_OK) return st;
		if ((st = ig_allgather_double(c, c->mom.qq, cap * g.K)) != IG_OK) return st;
		if ((st = ig_allgather_double(c, c->mom.qq2, cap * g.K)) != IG_OK) return st;
		c->mom_local = false;
	}
	CK(cudaMemcpyAsync(tot, c->mom.tot, sizeof(tot), cudaMemcpyDeviceToHost, c->stream));
#define DL(dst, src, n) \
	if (dst) CK(cudaMemcpyAsync(dst, src, (size_t)(n) * sizeof(double), cudaMemcpyDeviceToHost, c->stream))
	DL(out->indvlkh, c->mom.indvlkh, g.N);
	DL(out->qq, c->mom.qq, (size_t)g.N * g.K);
	DL(out->qq2, c->mom.qq2, (size_t)g.N * g.K);
	DL(out->self_rates, c->mom.self, c->ns);
	DL(out->self_rates2, c->mom.self2, c->ns);
	DL(out->gen, c->mom.gen, g.N);
	DL(out->gen2, c->mom.gen2, g.N);
	if (c->cfg.print_freq) {
		DL(out->freq, c->mom.freq, (size_t)g.K * g.L * g.A);
		DL(out->freq2, c->mom.freq2, (size_t)g.K * g.L * g.A);
	}
	if (convg_ld && c->cfg.ckrep > 0) {
		const long nv = stored < c->cfg.ckrep ? stored : c->cfg.ckrep;
		if (nv > 0) CK(cudaMemcpyAsync(convg_ld, c->mom.convg, (size_t)nv * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
	}
#undef DL
	CK(cudaStreamSynchronize(c->stream));
	out->totallkh = tot[0];
	out->totallkh2 = tot[1];
	return IG_OK;
}

extern "C" ig_status ig_run_chain(ig_ctx *c, int32_t chain_id, const float *initd, ig_chain_result *out, double *convg_ld)
{
	if (!c || !out) return fail(IG_ERR_ARG, "null argument");
	const ig_config &cf = c->cfg;
	if (cf.update < 1 || cf.burnin < 1 || cf.thinning < 1 || cf.burnin > cf.update)
		return fail(IG_ERR_ARG, "need 1 <= burnin <= update and thinning >= 1 (InStruct.c:299-300)");
	ig_status st = ig_chain_init(c, chain_id, initd);
	if (st != IG_OK) return st;
	const Geometry &g = c->geo;
	out->steps = (cf.update - cf.burnin) / cf.thinning;              // mcmc.c:485
	out->step = 0;
	out->flag_empty_cluster = 0;
	long cnt_step = 0;
	c->trace_rows = 0;
	if (c->ns == g.K && c->ns > 0 && cf.ckrep > 0) {                   // population rates: keep their trace beside the log-likelihood's
		ig_alloc_stream = c->stream;
		if (!c->mom.convg_S) CK(dalloc(&c->mom.convg_S, (size_t)cf.ckrep * g.K));
		else CK(cudaMemsetAsync(c->mom.convg_S, 0, (size_t)cf.ckrep * g.K * sizeof(double), c->stream));
	}
	// records local to their rank: each rank advances the moments of its own individuals, gathered once by download_result
	c->mom_local = c->comm && c->local_scalars;
	MomArgs m{c->ind, c->S, c->sc, c->P, c->mom, g, c->ns, 0, -1, cf.print_freq, cf.print_freq ? c->P64 : nullptr,
	          c->mom_local ? g.i0 : 0, c->mom_local ? g.Nloc : g.N};
	const long print_every = cf.update >= 100 ? cf.update / 100 : 1;   // print_info, mcmc.c:1273 (guards the /0 of App. B #7)
	for (long step = 0; step < cf.update; step++) {
		c->more_follow = step + 1 < cf.update;
		st = one_sweep(c);
		c->more_follow = false;
		if (st != IG_OK) return st;
		if (cf.print_iter == 1 && step % print_every == 0) {
			DevScalars h;
			CK(cudaMemcpyAsync(&h, c->sc, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
			CK(cudaStreamSynchronize(c->stream));
			fprintf(stdout, "\nStep=%ld\tlog_likelihood=%f\n", step + 1, h.totallkh);
			if (cf.mode == 2 || cf.mode == 4) {                          // print_info, mcmc.c:1276-1297
				double sh[MAX_K];
				CK(cudaMemcpy(sh, c->S, g.K * sizeof(double), cudaMemcpyDeviceToHost));
				for (int k = 0; k < g.K; k++) fprintf(stdout, "%c_%d=%f%s", cf.mode == 2 ? 's' : 'f', k, sh[k], k < g.K - 1 ? " " : "");
				fprintf(stdout, "\n");
			}
		}
		if (step == cf.burnin - 1) {                                   // allocate_chn, mcmc.c:218-219
			CK(launch_moments_reset(m, c->stream));
			c->launches++;
		}
		if (step >= cf.burnin && (step + 1 - cf.burnin) % cf.thinning == 0) {    // mcmc.c:220-226
			m.step = cnt_step;
			m.P = c->P;
			m.S = c->S;
			m.convg_slot = (cnt_step < cf.ckrep) ? (int)cnt_step : -1;
			CK(launch_moments(m, c->stream));
			c->launches++;
			cnt_step++;
		}
		// mcmc_POP_no_admixture (mcmc.c:90-131) never calls check_empty_cluster: a surplus cluster in mode 0 simply stays
		// empty (its one-hot column sums to 0) and the chain finishes; every other driver checks (mcmc.c:227-234)
		if (cf.mode != 0 && cnt_step == cf.nstep_check_empty_cluster) {
			DevScalars h;
			CK(cudaMemcpyAsync(&h, c->sc, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
			CK(cudaStreamSynchronize(c->stream));
			int empty = 0;
			for (int k = 0; k < g.K; k++) if (h.qcol[k] < 0.01) empty = 1;   // check_empty_cluster, mcmc.c:1963-1970
			out->flag_empty_cluster = empty;
			if (empty) break;
		}
	}
	out->step = cnt_step;
	c->trace_rows = (int)(cnt_step < cf.ckrep ? cnt_step : cf.ckrep);
	st = download_result(c, out, convg_ld, cnt_step);
	if (st != IG_OK) return st;
	return out->flag_empty_cluster ? IG_EMPTY_CLUSTER : IG_OK;
}

extern "C" ig_status ig_get_rate_trace(ig_ctx *c, double *out, size_t bytes, int32_t *rows)
{
	if (!c || !out) return fail(IG_ERR_ARG, "null argument");
	const Geometry &g = c->geo;
	if (!c->mom.convg_S) return fail(IG_ERR_STATE, "no rate trace: it exists for population rates (modes 2, 4, ploid 4) with ckrep > 0, after ig_run_chain");
	const size_t want = (size_t)c->cfg.ckrep * g.K * sizeof(double);
	if (bytes != want) return fail(IG_ERR_ARG, "rate trace: expected %zu bytes, got %zu", want, bytes);
	CK(cudaSetDevice(c->cfg.device));
	CK(cudaMemcpyAsync(out, c->mom.convg_S, want, cudaMemcpyDeviceToHost, c->stream));
	CK(cudaStreamSynchronize(c->stream));
	if (rows) *rows = c->trace_rows;
	return IG_OK;
}

extern "C" ig_status ig_mcmc_updating(const ig_config *cfg, const int16_t *x_host, const int32_t *allelenum_host,
                                      int32_t chain_id, const float *initd, ig_chain_result *out, double *convg_ld)
{
	ig_ctx *c = nullptr;
	const bool trace = getenv("IG_TRACE") != nullptr;       // wall-clock stages of the drop-in call on stderr
	const double t0 = wall_ms();
	ig_status st = ig_create(cfg, &c);
	if (st != IG_OK) return st;
	const double t1 = wall_ms();
	st = ig_load_genotypes(c, x_host, allelenum_host);
	const double t2 = wall_ms();
	if (st == IG_OK) st = ig_run_chain(c, chain_id, initd, out, convg_ld);
	const double t3 = wall_ms();
	ig_destroy(c);
	if (trace) fprintf(stderr, "[ig_trace] create %.1f ms, load %.1f, run_chain %.1f, destroy %.1f\n", t1 - t0, t2 - t1, t3 - t2, wall_ms() - t3);
	return st;
}

// --------------------------------------------------------------------------------------
// state hooks (parity tests): canonical host layouts <-> device layouts
// --------------------------------------------------------------------------------------
// biallelic path: the micro-tiled Z the hooks read is a view of the class-sorted Z, rebuilt when it is behind
static ig_status ensure_zt(ig_ctx *c)
{
	const Geometry &g = c->geo;
	if (!g.snp) return IG_OK;
	if (!c->Zt) { ig_alloc_stream = c->stream; CK(dalloc(&c->Zt, (size_t)g.LT * g.Nloc * TILE * 2)); c->zt_stale = true; }
	if (c->zt_stale) {
		CK(launch_snp_z_convert(c->Es, c->Zs, c->Zt, g, 0, c->stream));
		CK(cudaStreamSynchronize(c->stream));
		c->zt_stale = false;
	}
	return IG_OK;
}

// Host <-> device copies of the state hooks: on the context's stream and complete on return.  A plain cudaMemcpy runs on the
// legacy default stream, which is not ordered with the (non-blocking) context stream, and a host-to-device copy from
// pageable memory may return before its DMA has landed: a kernel launched on the context stream right after it could
// read the old contents (seen: the chunk-ordered copy of P built from a P that had not arrived yet).
static cudaError_t copy_sync(ig_ctx *c, void *dst, const void *src, size_t bytes, cudaMemcpyKind kind)
{
	cudaError_t e = cudaMemcpyAsync(dst, src, bytes, kind, c->stream);
	if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
	return e;
}

static ig_status need(size_t have, size_t want, const char *what)
{
	if (have != want) return fail(IG_ERR_ARG, "%s: expected %zu bytes, got %zu", what, want, have);
	return IG_OK;
}

extern "C" ig_status ig_get_state(ig_ctx *c, int32_t id, void *host, size_t bytes)
{
	if (!c || !host) return fail(IG_ERR_ARG, "null argument");
	if (!c->loaded) return fail(IG_ERR_STATE, "load genotypes first");
	CK(cudaSetDevice(c->cfg.device));
	const Geometry &g = c->geo;
	ig_status st;
	if ((st = ensure_records(c)) != IG_OK) return st;
	CK(cudaStreamSynchronize(c->stream));
	if (c->tetra) {
		bool handled = false;
		st = tetra_get_state(c, id, host, bytes, &handled);
		if (handled || st != IG_OK) return st;
	}
	switch (id) {
	case IG_STATE_X:
	case IG_STATE_Z: {
		const size_t el = (size_t)g.L * g.Nloc * 2;
		const size_t want = el * (id == IG_STATE_X ? 2 : 1);
		if ((st = need(bytes, want, "X/Z")) != IG_OK) return st;
		if (id == IG_STATE_Z && (st = ensure_zt(c)) != IG_OK) return st;
		void *tmp = nullptr;
		CK(cudaMalloc(&tmp, want));
		cudaError_t e = (id == IG_STATE_X) ? launch_untile_x(c->Xt, (int16_t *)tmp, g, c->stream) : launch_untile_z(c->Zt, (int8_t *)tmp, g, c->stream);
		if (e == cudaSuccess) e = cudaMemcpyAsync(host, tmp, want, cudaMemcpyDeviceToHost, c->stream);
		if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
		cudaFree(tmp);
		CK(e);
		return IG_OK;
	}
	case IG_STATE_MASK: {
		if ((st = need(bytes, (size_t)g.L * g.Nloc, "MASK")) != IG_OK) return st;
		std::vector<int16_t> x((size_t)g.L * g.Nloc * 2);
		if ((st = ig_get_state(c, IG_STATE_X, x.data(), x.size() * 2)) != IG_OK) return st;
		uint8_t *m = (uint8_t *)host;
		for (size_t t = 0; t < (size_t)g.L * g.Nloc; t++) m[t] = (x[2 * t] < 0 || x[2 * t + 1] < 0) ? 1 : 0;
		return IG_OK;
	}
	case IG_STATE_Q: {
		if ((st = need(bytes, (size_t)g.N * g.K * 8, "Q")) != IG_OK) return st;
		std::vector<double> r((size_t)c->Npad * g.REC);
		CK(copy_sync(c, r.data(), c->ind, r.size() * 8, cudaMemcpyDeviceToHost));
		double *q = (double *)host;
		for (int i = 0; i < g.N; i++) for (int k = 0; k < g.K; k++) q[(size_t)i * g.K + k] = r[(size_t)i * g.REC + k];
		return IG_OK;
	}
	case IG_STATE_G:
	case IG_STATE_INDVLKH: {
		std::vector<double> r((size_t)c->Npad * g.REC);
		CK(copy_sync(c, r.data(), c->ind, r.size() * 8, cudaMemcpyDeviceToHost));
		if (id == IG_STATE_G) {
			if ((st = need(bytes, (size_t)g.N * 4, "G")) != IG_OK) return st;
			for (int i = 0; i < g.N; i++) ((int32_t *)host)[i] = (int32_t)r[(size_t)i * g.REC + g.K + 2];
		} else {
			if ((st = need(bytes, (size_t)g.N * 8, "INDVLKH")) != IG_OK) return st;
			for (int i = 0; i < g.N; i++) ((double *)host)[i] = r[(size_t)i * g.REC + g.K];
		}
		return IG_OK;
	}
	case IG_STATE_P:
	case IG_STATE_TALLY: {
		const size_t pn = (size_t)g.Lpad * g.A * g.KP;
		const size_t el = (size_t)g.K * g.L * g.A;
		if ((st = need(bytes, el * (id == IG_STATE_P ? 8 : 4), "P/TALLY")) != IG_OK) return st;
		if (id == IG_STATE_P) {
			std::vector<float> p(pn);
			CK(copy_sync(c, p.data(), c->P, pn * 4, cudaMemcpyDeviceToHost));
			for (int k = 0; k < g.K; k++) for (int l = 0; l < g.L; l++) for (int a = 0; a < g.A; a++)
				((double *)host)[((size_t)k * g.L + l) * g.A + a] = (double)p[((size_t)l * g.A + a) * g.KP + k];
		} else {
			std::vector<int32_t> p(pn);
			CK(copy_sync(c, p.data(), c->n, pn * 4, cudaMemcpyDeviceToHost));
			for (int k = 0; k < g.K; k++) for (int l = 0; l < g.L; l++) for (int a = 0; a < g.A; a++)
				((int32_t *)host)[((size_t)k * g.L + l) * g.A + a] = p[((size_t)l * g.A + a) * g.KP + k];
		}
		return IG_OK;
	}
	case IG_STATE_ALPHA:
	case IG_STATE_TOTALLKH: {
		if ((st = need(bytes, 8, "scalar")) != IG_OK) return st;
		DevScalars h;
		CK(copy_sync(c, &h, c->sc, sizeof(h), cudaMemcpyDeviceToHost));
		*(double *)host = (id == IG_STATE_ALPHA) ? h.alpha : h.totallkh;
		return IG_OK;
	}
	case IG_STATE_S:
		if ((st = need(bytes, (size_t)c->ns * 8, "S")) != IG_OK) return st;
		CK(copy_sync(c, host, c->S, bytes, cudaMemcpyDeviceToHost));
		return IG_OK;
	case IG_STATE_STATE:
		if ((st = need(bytes, (size_t)g.K * 4, "STATE")) != IG_OK) return st;
		CK(copy_sync(c, host, c->state, bytes, cudaMemcpyDeviceToHost));
		return IG_OK;
	case IG_STATE_CNT:
		if ((st = need(bytes, (size_t)g.Nloc * g.K * 4, "CNT")) != IG_OK) return st;
		CK(copy_sync(c, host, c->cnt, bytes, cudaMemcpyDeviceToHost));
		return IG_OK;
	case IG_STATE_GPROP:
		if ((st = need(bytes, (size_t)g.N * 4, "GPROP")) != IG_OK) return st;
		CK(copy_sync(c, host, c->gprop, bytes, cudaMemcpyDeviceToHost));
		return IG_OK;
	case IG_STATE_ITER:
		if ((st = need(bytes, 8, "ITER")) != IG_OK) return st;
		*(int64_t *)host = c->iter;
		return IG_OK;
	case 100:     /* debug: per-individual (d_old, c_new, a_new, b_new) of the last zq pass, double [Nloc][4] */
		if ((st = need(bytes, (size_t)g.Nloc * 32, "LLPARTS")) != IG_OK) return st;
		CK(copy_sync(c, host, c->llparts, bytes, cudaMemcpyDeviceToHost));
		return IG_OK;
	case 102:     /* debug: proposed inbreeding coefficients of the last sweep, double [K] (mode 4) or [N] (mode 5) */
		if (!c->fprop) return fail(IG_ERR_ARG, "FPROP exists in modes 4 and 5 only");
		if ((st = need(bytes, (size_t)c->ns * 8, "FPROP")) != IG_OK) return st;
		CK(copy_sync(c, host, c->fprop, bytes, cudaMemcpyDeviceToHost));
		return IG_OK;
	case 103: {   /* debug, mode 4: per-individual old-Z and new-Z differences per population, double [N][2][K] (nats) */
		if (g.fmode != 2) return fail(IG_ERR_ARG, "FK exists in mode 4 only");
		if ((st = need(bytes, (size_t)g.N * 2 * g.K * 8, "FK")) != IG_OK) return st;
		std::vector<double> r((size_t)c->Npad * g.REC);
		CK(copy_sync(c, r.data(), c->ind, r.size() * 8, cudaMemcpyDeviceToHost));
		for (int i = 0; i < g.N; i++) for (int j = 0; j < 2 * g.K; j++) ((double *)host)[(size_t)i * 2 * g.K + j] = r[(size_t)i * g.REC + g.K + 3 + j];
		return IG_OK;
	}
	case 104: {   /* Dirichlet-process step: for every individual j, the weights gen_post_prob (DPMM.c:361-377) would hand to
		         disc_unif with j taken out of its cluster and everybody else where they are: double [N][N+1],
		         [j][0] = alpha / ((G_j + 1) G_j) (new cluster), [j][1..] = num_c * dgeom(S_c, G_j) in list (value) order, 0 beyond */
		if (!(c->cfg.mode == 3 && c->cfg.prior_flag == 1)) return fail(IG_ERR_ARG, "DP weights exist in mode 3 with the DP prior only");
		if ((st = need(bytes, (size_t)g.N * (g.N + 1) * 8, "DPWEIGHTS")) != IG_OK) return st;
		if ((int)c->dp_of.size() != g.N) return fail(IG_ERR_STATE, "initialise the chain (or set S) first");
		std::vector<double> r((size_t)c->Npad * g.REC);
		CK(copy_sync(c, r.data(), c->ind, r.size() * 8, cudaMemcpyDeviceToHost));
		double *w = (double *)host;
		memset(w, 0, bytes);
		for (int j = 0; j < g.N; j++) {
			const int gen = (int)r[(size_t)j * g.REC + g.K + 2];
			double *wj = w + (size_t)j * (g.N + 1);
			wj[0] = c->cfg.alpha_dpm / (gen + 1) / gen;
			int n = 1;
			for (int p = c->dp_head; p >= 0; p = c->dp[p].next) {
				const int num = c->dp[p].num - (c->dp_of[j] == p ? 1 : 0);
				if (num == 0) continue;                                   // j's own singleton disappears when j leaves (delete, DPMM.c:280)
				wj[n++] = num * dp_dgeom(c, p, gen);
			}
		}
		return IG_OK;
	}
	case 105:     /* number of Dirichlet-process clusters, int64 */
		if ((st = need(bytes, 8, "DPCLUSTERS")) != IG_OK) return st;
		*(int64_t *)host = c->dp_cnt;
		return IG_OK;
	case 101: {   /* debug: launch geometry int32[8] = TL, nchunks, nblk, subs_per_blk, R, smem, KP, A */
		if ((st = need(bytes, 32, "GEOMETRY")) != IG_OK) return st;
		int32_t *o = (int32_t *)host;
		o[0] = g.TL; o[1] = g.nchunks; o[2] = g.nblk; o[3] = g.subs_per_blk; o[4] = g.R; o[5] = (int32_t)g.zq_smem; o[6] = g.KP; o[7] = g.A;
		return IG_OK;
	}
	default:
		return fail(IG_ERR_ARG, "unknown state id %d", id);
	}
}

extern "C" ig_status ig_set_state(ig_ctx *c, int32_t id, const void *host, size_t bytes)
{
	if (!c || !host) return fail(IG_ERR_ARG, "null argument");
	if (!c->loaded) return fail(IG_ERR_STATE, "load genotypes first");
	CK(cudaSetDevice(c->cfg.device));
	const Geometry &g = c->geo;
	ig_status st;
	if ((st = ensure_records(c)) != IG_OK) return st;
	CK(cudaStreamSynchronize(c->stream));
	c->g8_inflight = false;
	if (c->tetra) {
		bool handled = false;
		st = tetra_set_state(c, id, host, bytes, &handled);
		if (handled || st != IG_OK) return st;
	}
	switch (id) {
	case IG_STATE_Z: {
		const size_t want = (size_t)g.L * g.Nloc * 2;
		if ((st = need(bytes, want, "Z")) != IG_OK) return st;
		if ((st = ensure_zt(c)) != IG_OK) return st;
		void *tmp = nullptr;
		CK(cudaMalloc(&tmp, want));
		cudaError_t e = cudaMemcpyAsync(tmp, host, want, cudaMemcpyHostToDevice, c->stream);
		if (e == cudaSuccess) e = launch_tile_z((const int8_t *)tmp, c->Zt, g, c->stream);
		if (e == cudaSuccess) e = launch_tally(c->Xt, c->Zt, c->n, g, c->stream);    // n always mirrors Z
		if (e == cudaSuccess) e = launch_het_counts(c->Xt, c->Zt, nullptr, c->nsh, g, c->stream);   // and so does nsh
		if (e == cudaSuccess && g.snp) e = launch_snp_z_convert(c->Es, c->Zs, c->Zt, g, 1, c->stream);
		if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
		cudaFree(tmp);
		CK(e);
		return IG_OK;
	}
	case IG_STATE_Q: {
		if ((st = need(bytes, (size_t)g.N * g.K * 8, "Q")) != IG_OK) return st;
		std::vector<double> r((size_t)c->Npad * g.REC);
		CK(copy_sync(c, r.data(), c->ind, r.size() * 8, cudaMemcpyDeviceToHost));
		const double *q = (const double *)host;
		for (int i = 0; i < g.N; i++) {
			double slq = 0;
			for (int k = 0; k < g.K; k++) { r[(size_t)i * g.REC + k] = q[(size_t)i * g.K + k]; slq += log(q[(size_t)i * g.K + k]); }
			r[(size_t)i * g.REC + g.K + 1] = slq;
		}
		CK(copy_sync(c, c->ind, r.data(), r.size() * 8, cudaMemcpyHostToDevice));
		CK(launch_qf_from_ind(c->ind, c->Qf, g, c->stream));
		CK(cudaStreamSynchronize(c->stream));
		return IG_OK;
	}
	case IG_STATE_G:
	case IG_STATE_INDVLKH: {
		std::vector<double> r((size_t)c->Npad * g.REC);
		CK(copy_sync(c, r.data(), c->ind, r.size() * 8, cudaMemcpyDeviceToHost));
		if (id == IG_STATE_G) {
			if ((st = need(bytes, (size_t)g.N * 4, "G")) != IG_OK) return st;
			for (int i = 0; i < g.N; i++) r[(size_t)i * g.REC + g.K + 2] = (double)((const int32_t *)host)[i];
		} else {
			if ((st = need(bytes, (size_t)g.N * 8, "INDVLKH")) != IG_OK) return st;
			for (int i = 0; i < g.N; i++) r[(size_t)i * g.REC + g.K] = ((const double *)host)[i];
		}
		CK(copy_sync(c, c->ind, r.data(), r.size() * 8, cudaMemcpyHostToDevice));
		if (id == IG_STATE_G && c->cfg.mode == 0 && c->cfg.ploid == 2) {       // mode 0: G carries the cluster labels zz; n mirrors them
			std::vector<double> r2(r);
			for (int i = 0; i < g.N; i++) for (int k = 0; k < g.K; k++) r2[(size_t)i * g.REC + k] = (k == ((const int32_t *)host)[i]) ? 1.0 : 0.0;
			CK(copy_sync(c, c->ind, r2.data(), r2.size() * 8, cudaMemcpyHostToDevice));
			if ((st = na_retally(c)) != IG_OK) return st;
			CK(cudaStreamSynchronize(c->stream));
		}
		return IG_OK;
	}
	case IG_STATE_P: {
		const size_t pn = (size_t)g.Lpad * g.A * g.KP;
		if ((st = need(bytes, (size_t)g.K * g.L * g.A * 8, "P")) != IG_OK) return st;
		std::vector<float> p(pn, 0.0f);
		for (int k = 0; k < g.K; k++) for (int l = 0; l < g.L; l++) for (int a = 0; a < g.A; a++)
			p[((size_t)l * g.A + a) * g.KP + k] = (float)((const double *)host)[((size_t)k * g.L + l) * g.A + a];
		CK(copy_sync(c, c->P, p.data(), pn * 4, cudaMemcpyHostToDevice));
		if (g.snp) { CK(launch_snp_pc_from_p(c->P, c->Pc, g, c->stream)); CK(cudaStreamSynchronize(c->stream)); }
		return IG_OK;
	}
	case IG_STATE_ALPHA: {
		if ((st = need(bytes, 8, "ALPHA")) != IG_OK) return st;
		CK(copy_sync(c, &c->sc->alpha, host, 8, cudaMemcpyHostToDevice));
		return IG_OK;
	}
	case IG_STATE_S:
		if ((st = need(bytes, (size_t)c->ns * 8, "S")) != IG_OK) return st;
		CK(copy_sync(c, c->S, host, bytes, cudaMemcpyHostToDevice));
		if (c->cfg.ploid == 2 && c->cfg.mode == 3 && c->cfg.prior_flag == 1) dp_from_values(c, (const double *)host);   // the host-side clusters follow
		return IG_OK;
	case IG_STATE_STATE:
		if ((st = need(bytes, (size_t)g.K * 4, "STATE")) != IG_OK) return st;
		CK(copy_sync(c, c->state, host, bytes, cudaMemcpyHostToDevice));
		return IG_OK;
	case IG_STATE_GPROP: {      /* also refreshes the (g, g') pairs the sweep kernel reads */
		if ((st = need(bytes, (size_t)g.N * 4, "GPROP")) != IG_OK) return st;
		CK(copy_sync(c, c->gprop, host, bytes, cudaMemcpyHostToDevice));
		std::vector<double> r((size_t)c->Npad * g.REC);
		CK(copy_sync(c, r.data(), c->ind, r.size() * 8, cudaMemcpyDeviceToHost));
		std::vector<int2> gp(g.Nloc);
		for (int il = 0; il < g.Nloc; il++) gp[il] = make_int2((int)r[(size_t)(g.i0 + il) * g.REC + g.K + 2], ((const int32_t *)host)[g.i0 + il]);
		CK(copy_sync(c, c->gpair, gp.data(), gp.size() * sizeof(int2), cudaMemcpyHostToDevice));
		return IG_OK;
	}
	case IG_STATE_ITER:
		if ((st = need(bytes, 8, "ITER")) != IG_OK) return st;
		c->iter = (uint32_t) * (const int64_t *)host;
		c->chain_ready = true;
		c->dev_iter_valid = false;
		return IG_OK;
	default:
		return fail(IG_ERR_ARG, "state id %d cannot be set", id);
	}
}

extern "C" ig_status ig_loglik(ig_ctx *c, const int32_t *gen, double *out)
{
	if (!c || !gen || !out) return fail(IG_ERR_ARG, "null argument");
	if (!c->loaded) return fail(IG_ERR_STATE, "load genotypes first");
	CK(cudaSetDevice(c->cfg.device));
	const Geometry &g = c->geo;
	{ ig_status stz = ensure_zt(c); if (stz != IG_OK) return stz; }
	int32_t *gd = nullptr;
	double *od = nullptr;
	CK(cudaMalloc((void **)&gd, (size_t)g.Nloc * 4));
	CK(cudaMalloc((void **)&od, (size_t)g.Nloc * 8));
	cudaError_t e = cudaMemcpyAsync(gd, gen, (size_t)g.Nloc * 4, cudaMemcpyHostToDevice, c->stream);
	if (e == cudaSuccess) e = launch_loglik(c->Xt, c->Zt, c->P, c->Qf, gd, od, g, c->cfg.type_freq, c->stream);
	if (e == cudaSuccess) e = cudaMemcpyAsync(out, od, (size_t)g.Nloc * 8, cudaMemcpyDeviceToHost, c->stream);
	if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
	cudaFree(gd);
	cudaFree(od);
	CK(e);
	return IG_OK;
}

extern "C" ig_status ig_proposal_loglik(ig_ctx *c, const double *S, double *out)
{
	if (!c || !S || !out) return fail(IG_ERR_ARG, "null argument");
	if (!c->loaded) return fail(IG_ERR_STATE, "load genotypes first");
	if (c->cfg.mode != 2) return fail(IG_ERR_ARG, "proposal() is the mode-2 likelihood (mcmc.c:1630)");
	CK(cudaSetDevice(c->cfg.device));
	{ ig_status str = ensure_records(c); if (str != IG_OK) return str; }
	CK(cudaMemcpyAsync(c->scratch, S, (size_t)c->geo.K * 8, cudaMemcpyHostToDevice, c->stream));
	CK(launch_proposal_ll(c->ind, c->scratch, c->scratch + 32, c->geo, c->stream));
	CK(cudaMemcpyAsync(out, c->scratch + 32, 8, cudaMemcpyDeviceToHost, c->stream));
	CK(cudaStreamSynchronize(c->stream));
	return IG_OK;
}

extern "C" ig_status ig_alpha_logratio(ig_ctx *c, double ralpha, double *out)
{
	if (!c || !out) return fail(IG_ERR_ARG, "null argument");
	if (!c->loaded) return fail(IG_ERR_STATE, "load genotypes first");
	CK(cudaSetDevice(c->cfg.device));
	// sum log q is maintained per individual in the record; reduce it with the same kernel the
	// sweep uses, on a scratch copy of the scalars so that alpha itself is not advanced
	DevScalars keep;
	{ ig_status str = ensure_records(c); if (str != IG_OK) return str; }
	CK(cudaStreamSynchronize(c->stream));
	CK(copy_sync(c, &keep, c->sc, sizeof(keep), cudaMemcpyDeviceToHost));
	PostArgs a{c->ind, c->sc, c->gpart, c->geo, 0xFFFFFFFFu, c->key0, c->key1, nullptr, c->cfg.mode, c->cfg.back_refl,
	           c->S, c->fprop, c->state, c->state2, c->grid_post};
	CK(launch_post_sweep(a, c->stream));
	DevScalars h;
	CK(cudaMemcpyAsync(&h, c->sc, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
	CK(cudaStreamSynchronize(c->stream));
	*out = (ralpha - keep.alpha) * h.sumlogq;
	CK(copy_sync(c, c->sc, &keep, sizeof(keep), cudaMemcpyHostToDevice));
	return IG_OK;
}

// --------------------------------------------------------------------------------------
// profiling
// --------------------------------------------------------------------------------------
extern "C" ig_status ig_profile(ig_ctx *c, int32_t enable)
{
	if (!c) return fail(IG_ERR_ARG, "null context");
	CK(cudaSetDevice(c->cfg.device));
	c->profile = enable != 0;
	c->ev_used = 0;
	if (enable && c->ev.empty()) {
		c->ev.resize(2 * 4096);
		for (auto &e : c->ev) CK(cudaEventCreate(&e));
	}
	return IG_OK;
}

extern "C" ig_status ig_profile_read(ig_ctx *c, int32_t *launches, double *zq_ms_total, int64_t *kernels_launched)
{
	if (!c) return fail(IG_ERR_ARG, "null context");
	CK(cudaSetDevice(c->cfg.device));
	CK(cudaStreamSynchronize(c->stream));
	double tot = 0.0;
	for (int i = 0; i + 1 < c->ev_used; i += 2) {
		float ms = 0.f;
		CK(cudaEventElapsedTime(&ms, c->ev[i], c->ev[i + 1]));
		tot += ms;
	}
	if (launches) *launches = c->ev_used / 2;
	if (zq_ms_total) *zq_ms_total = tot;
	if (kernels_launched) *kernels_launched = c->launches;
	c->ev_used = 0;
	return IG_OK;
}

extern "C" ig_status ig_algorithmic_bytes(ig_ctx *c, double *bytes_per_sweep, double *copies_per_sweep)
{
	if (!c) return fail(IG_ERR_ARG, "null context");
	const Geometry &g = c->geo;
	// N*L*ploid*4 B (int16 x + int8 old z + int8 new z per allele copy)
	//             + K*L*A*12 B (count write + count read + P write) + N*(16K+16) B (Q r/w, G, lkh)
	double abar = 0.0;
	for (int l = 0; l < g.L; l++) abar += c->allelenum_h.empty() ? g.A : c->allelenum_h[l];
	const double b = (double)g.Nloc * g.L * 2 * 4.0 + (double)g.K * abar * 12.0 + (double)g.Nloc * (16.0 * g.K + 16.0);
	if (bytes_per_sweep) *bytes_per_sweep = b;
	if (copies_per_sweep) *copies_per_sweep = (double)g.Nloc * g.L * 2;
	return IG_OK;
}
