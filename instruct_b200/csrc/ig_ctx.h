// ig_ctx.h -- the context object behind the C-ABI and the error helpers, shared by ig_api.cu
// (diploid driver, hooks) and tetra.cu (autotetraploid driver).  Not part of the C-ABI.
#pragma once
#include <nccl.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <string>
#include <utility>
#include <vector>

#include "ig_internal.h"

// every device allocation of the library goes through the caching allocator of ig_pool.cu
cudaError_t ig_pool_malloc(void **p, size_t bytes);
cudaError_t ig_pool_free(void *p);
#define cudaMalloc(p, n) ig_pool_malloc((void **)(p), (n))
#define cudaFree(p) ig_pool_free((void *)(p))

ig_status ig_fail(ig_status st, const char *fmt, ...);
#define fail ig_fail
#define CK(call)                                                                                           \
	do {                                                                                               \
		cudaError_t e_ = (call);                                                                       \
		if (e_ != cudaSuccess) return ig_fail(IG_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
	} while (0)

namespace ig { struct TetraState; }
using namespace ig;

struct DpCluster { double value; int num; int next; };

struct ig_ctx {
	ig_config cfg;
	Geometry geo;
	int ns;                    // length of S: K (mode 2) or N (mode 3)
	int Npad;                  // records in ind: shard_count * shard_cap
	int shard_cap;
	cudaStream_t stream = nullptr;
	bool loaded = false, chain_ready = false;
	uint32_t iter = 0, key0 = 0, key1 = 0;
	int rounds = 7;
	int trace_rows = 0;        // retained sweeps held by the traces of the last ig_run_chain (<= ckrep)
	int grid_pre = 1, grid_post = 1;   // cooperative grids of the scalar kernels on this context's device
	// device buffers
	int16_t *Xt = nullptr;
	int8_t *Zt = nullptr;           // micro-tiled Z; on the biallelic path only the state hooks' view (allocated on demand)
	// biallelic path (zq_snp.cu): class-sorted store, Z in the same order, chunk-ordered P
	uint32_t *Es = nullptr, *Hs = nullptr;
	uint16_t *Zs = nullptr;
	float *Pc = nullptr, *Pcnext = nullptr;
	bool zt_stale = false;          // Zs is ahead of Zt
	float *P = nullptr;
	double *P64 = nullptr;
	int32_t *n = nullptr;
	int32_t *allelenum = nullptr;
	double *ind = nullptr;
	float *Qf = nullptr;
	int32_t *gprop = nullptr;
	int2 *gpair = nullptr;
	double *S = nullptr;
	int32_t *state = nullptr;
	DevScalars *sc = nullptr;
	uint16_t *pcnt = nullptr;
	double *plog = nullptr;
	uint16_t *pnsh = nullptr;
	int32_t *nhet = nullptr, *nsh = nullptr;
	int32_t *cnt = nullptr;
	double *llparts = nullptr;
	float *initd_dev = nullptr;
	double *scratch = nullptr;      // small device scratch (parity hooks)
	double *gpart = nullptr;        // partials of the cooperative grid sums
	int32_t *state2 = nullptr;      // double buffer of UPMCMC.state (-e 0)
	// modes 4/5 (inbreeding coefficients; S holds UPMCMC.inbreed)
	double *fprop = nullptr;        // proposed coefficients [K] / [N]
	float2 *hpair = nullptr;        // mode 5: (1 - F, 1 - F') per local individual
	float *ftab = nullptr;          // mode 4: the sweep kernel's per-population table
	float *pfk = nullptr;           // mode 4: per-population partials of the sweep kernel
	// mode 0 (no admixture, noadmix.cu)
	float *logP = nullptr;          // log2 P, [Lpad][A][KP]
	double *na_pll = nullptr;       // [nchunks][KP][Nloc] per-chunk log-likelihoods of each individual in each cluster
	Moments mom{};
	// host mirrors
	std::vector<int32_t> allelenum_h;
	std::vector<double> ind_h, S_h;
	// DP prior (host)
	std::vector<DpCluster> dp;
	std::vector<double> dp_w;     // [slot][51] dgeom(value, g), g = 1..50
	std::vector<int> dp_of;
	std::vector<double> dp_cum;   // scratch of the scan
	std::vector<int> dp_node;
	int dp_head = -1, dp_free = -1, dp_cnt = 0;
	// DP prior: the scan reads only G (one byte each, <= 50) and writes only S.  G is packed and copied to pinned host memory
	// right behind the epilogue of the PREVIOUS sweep, so the host scan overlaps post_sweep and the next p_dirichlet;
	// S goes back from pinned memory on the stream.  (Round 1 copied the whole record array both ways, synchronously.)
	uint8_t *g8_dev = nullptr, *g8_host = nullptr;   // [N]
	double *S_pin = nullptr;                         // [N] pinned
	cudaEvent_t ev_g = nullptr;
	bool g8_inflight = false;                        // a copy of the current G is on its way (or has arrived) in g8_host
	// NCCL
	ncclComm_t comm = nullptr;
	ncclComm_t comm2 = nullptr;      // a second communicator over the same ranks for the side stream: operations on ONE communicator
	                                 // are serialised by NCCL even across streams, which put the 6.4 MB tally all-reduce in front of
	                                 // the all-gather of the records on the critical path
	// sharded chains: the tally all-reduce and the NEXT sweep's P draw run on a side stream behind
	// the sweep kernel, off the critical path (ig_api.cu early_update_P)
	cudaStream_t stream2 = nullptr;
	cudaEvent_t ev_zq = nullptr, ev_p = nullptr;
	float *Pnext = nullptr;
	int32_t *nred = nullptr;         // this rank's block of the reduce-scattered tally (p_lr loci); null: all-reduce + full draw on every rank
	int p_lr = 0;                    // loci per rank of that split: ceil(Lpad / W)
	bool early_p = false;            // Pnext holds the P of sweep iter + 1 and n has been consumed
	bool p_wait = false;             // the main stream has not yet waited for that draw (it does before zq_sweep)
	bool more_follow = false;        // another sweep follows inside the current API call
	// Local scalar updates (modes 1-2, K <= 10): a sharded chain keeps each individual's record on its own rank only; the
	// sums update_S_POP / update_alpha / cal_lkh need are fixed-point int64 sums of the local individuals, all-reduced
	// (ig_kernels.cu spop_tree_kernel).  The records are gathered on demand (state hooks) and the moments at the chain's end.
	bool local_scalars = false;      // this context runs its sweeps that way
	bool local_now = false;          // ... and is inside such a sweep (the phase hooks always use the gathered path)
	bool ind_stale = false;          // other ranks' records in `ind` are older than the last sweep
	bool mom_local = false;          // the running moments of other ranks' individuals live on those ranks
	bool tree_ready = false;         // fx holds the all-reduced subset sums for sweep iter + 1 (computed behind the previous post_sweep)
	unsigned long long *fx = nullptr;   // local accumulators: [32] post sums | [2^K][2] subset sums; zero between uses (whoever consumes them clears them)
	unsigned long long *fxg = nullptr;  // the same layout, summed over the ranks
	double *S2 = nullptr;            // double buffer of S (every CTA of the decide kernel reads S, CTA 0 writes it)
	// ... all-reduced over NVLink peer memory (peer_allreduce_kernel) when every rank could map every rank's buffer
	unsigned long long *px_buf = nullptr;             // this rank's buffer (a plain cudaMalloc: it is exported through CUDA IPC)
	std::vector<void *> px_mapped;                    // the peers' buffers as opened here (to close them)
	unsigned long long **px_peers = nullptr;          // device array [W]
	unsigned long long px_seq = 0;
	bool px = false;
	// ... and the tally -> P exchange over the same arena (p_peer_*): n and both P buffers live in it
	bool p_peer = false;
	size_t px_n_off = 0, px_p_off[2] = {0, 0};       // byte offsets inside every rank's arena
	int p_par = 0;                                    // which P buffer c->P is (toggles with the P / Pnext swap; in lockstep on all ranks)
	unsigned long long pp_seq = 0;
	// IG_PHASE_TRACE=1: CUDA events at the phase boundaries of the first 64 sweeps, averages printed by ig_destroy
	std::vector<cudaEvent_t> ptrace;
	int ptrace_sweeps = 0;
	double ptrace_host_ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};     // host time spent enqueueing each phase
	// profiling
	bool profile = false;
	std::vector<cudaEvent_t> ev;
	int ev_used = 0;
	int64_t launches = 0;
	// CUDA-graph replay of one sweep (ig_api.cu one_sweep)
	cudaGraphExec_t graph_exec = nullptr;
	bool graph_failed = false;
	int64_t launches_per_sweep = 0;
	const uint32_t *iter_dev = nullptr;   // non-null only while a sweep is being captured
	bool dev_iter_valid = false;          // DevScalars.iter is known to equal iter + 1
	// autotetraploid driver (tetra.cu); null for ploid 2
	ig::TetraState *tetra = nullptr;
};

static inline uint32_t pad_k(int K) { return K <= 4 ? 4 : (K <= 8 ? 8 : 16); }

// Zero-filled device allocation, complete on return.  A plain cudaMemset runs on the legacy default
// stream, which is NOT ordered with the context's non-blocking stream: the fill could land after the
// first asynchronous write into the buffer (seen: the allele-count vector zeroed after its upload when
// the caller had work queued on stream 0).  Filling on the context's stream alone would leave the
// opposite race with the synchronous cudaMemcpy uploads, so the fill is waited for.
extern thread_local cudaStream_t ig_alloc_stream;      // set by the API entry points that allocate
template <typename T>
static cudaError_t dalloc(T **p, size_t n)
{
	cudaError_t e = cudaMalloc((void **)p, (n ? n : 1) * sizeof(T));
	if (e == cudaSuccess) e = cudaMemsetAsync(*p, 0, (n ? n : 1) * sizeof(T), ig_alloc_stream);
	if (e == cudaSuccess) e = cudaStreamSynchronize(ig_alloc_stream);
	return e;
}


// per-sweep exchanges of the individual-sharded mode (ig_api.cu); no-ops on one GPU
ig_status ig_exchange_tally(ig_ctx *c);
ig_status ig_exchange_individuals(ig_ctx *c);
ig_status ig_allgather_double(ig_ctx *c, double *buf, size_t per_rank);
ig_status ig_allreduce_int32(ig_ctx *c, int32_t *buf, size_t count);
ig_status ig_allreduce_int64(ig_ctx *c, int64_t *buf, size_t count);

// no-admixture driver (noadmix.cu)
ig_status na_alloc(ig_ctx *c);
ig_status na_chain_init(ig_ctx *c);
ig_status na_phase_z(ig_ctx *c);
ig_status na_retally(ig_ctx *c);

// autotetraploid driver (tetra.cu)
ig_status tetra_create(ig_ctx *c);
void tetra_destroy(ig_ctx *c);
ig_status tetra_load(ig_ctx *c, const int16_t *x_dev);
ig_status tetra_chain_init(ig_ctx *c, int32_t chain_id, const float *initd);
ig_status tetra_one_sweep(ig_ctx *c);
ig_status tetra_run_phase(ig_ctx *c, int32_t mask);
ig_status tetra_get_state(ig_ctx *c, int32_t id, void *host, size_t bytes, bool *handled);
ig_status tetra_set_state(ig_ctx *c, int32_t id, const void *host, size_t bytes, bool *handled);
