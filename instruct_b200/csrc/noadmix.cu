// noadmix.cu -- mode 0, the no-admixture model (mcmc_POP_no_admixture, mcmc.c:90-131): every
// individual belongs wholly to one cluster zz[i].
//
//   update_P   mcmc.c:799-861 with the zz branch (:825-831): n[l][a][zz[i]] over both copies
//   update_Z   mcmc.c:1094-1119: P(zz[i] = k) proportional to exp(log_ld_indv_K(i, k))
//   cal_lkh    mcmc.c:1926: indvlkh[i] = log_ld_indv_K(i, zz[i])            (mcmc.c:1893-1913)
//
// One sweep reads the genotype store twice (2 B per allele copy each time): the scan that
// produces the K log-likelihoods of every individual, and -- once zz is drawn from them --
// the tally for the next update_P.  There is no per-copy state (no Z store).  The logarithms
// are taken once per sweep on the L x A x K table (log2 P), so the scan is loads and adds.
#include <math.h>
#include "ig_ctx.h"
#include "philox.cuh"
#include "samplers.cuh"
#include "sweep_common.cuh"

namespace ig {

#define LN2_D 0.69314718055994530942

struct NaArgs {
	const int16_t *Xt;       // [LT][Nloc][TILE][2]
	const float *logP;       // [Lpad][A][KP]  log2 P
	int32_t *n;              // [Lpad][A][KP]
	double *pll;             // [nchunks][KP][Nloc]  per (chunk, individual): log-likelihood in each cluster (nats)
	const int32_t *nhet;     // [Nloc]
	double *ind;             // [Npad][REC]
	Geometry geo;
	uint32_t iter, key0, key1;
	const uint32_t *iter_dev;
	int init;
};

__global__ void na_logp_kernel(const float *P, float *logP, size_t total)
{
	const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (t < total) logP[t] = log2f(P[t]);
}

// scan: grid (locus chunks, individual blocks) x 256 threads, one thread = one individual,
// the chunk's log2 P table in shared memory (one TMA bulk copy).
template <int KP>
__global__ void __launch_bounds__(ZQ_THREADS) na_scan_kernel(const NaArgs a)
{
	extern __shared__ __align__(128) unsigned char smem_raw[];
	__shared__ __align__(8) unsigned long long bar;
	const Geometry &g = a.geo;
	const int tid = threadIdx.x, chunk = blockIdx.x;
	const int l0 = chunk * g.TL;
	const int nl = min(g.TL, g.Lpad - l0);
	const int nmt = nl / TILE;
	const int rowsz = g.A * KP;
	float *Lsm = reinterpret_cast<float *>(smem_raw);
	if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
	__syncthreads();
	if (tid == 0) {
		mbar_expect_tx(&bar, (uint32_t)nl * rowsz * 4u);
		tma_bulk_g2s(Lsm, a.logP + (size_t)l0 * rowsz, (uint32_t)nl * rowsz * 4u, &bar);
	}
	mbar_wait(&bar, 0);
	const int Nloc = g.Nloc, mt0 = l0 / TILE;
	const int nsub_total = (Nloc + ZQ_THREADS - 1) / ZQ_THREADS;
	const int sub0 = blockIdx.y * g.subs_per_blk, sub1 = min(sub0 + g.subs_per_blk, nsub_total);
	for (int sub = sub0; sub < sub1; ++sub) {
		const int il = sub * ZQ_THREADS + tid;
		if (il >= Nloc) continue;
		double acc[KP];
#pragma unroll
		for (int k = 0; k < KP; k++) acc[k] = 0.0;
		const int4 *xp = reinterpret_cast<const int4 *>(a.Xt) + ((size_t)mt0 * Nloc + il) * 2;
		for (int mt = 0; mt < nmt; ++mt) {
			const int4 xa = ldg_stream(xp + (size_t)mt * Nloc * 2), xb = ldg_stream(xp + (size_t)mt * Nloc * 2 + 1);
			const int xw[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
			float m[KP];
#pragma unroll
			for (int k = 0; k < KP; k++) m[k] = 0.0f;
#pragma unroll
			for (int j = 0; j < TILE; j++) {
				if (xw[j] < 0) continue;                             // any copy missing, monomorphic or padded locus
				const int x0 = xw[j] & 0xFFFF, x1 = (xw[j] >> 16) & 0xFFFF;
				const float *r0 = Lsm + ((size_t)(mt * TILE + j) * g.A + x0) * KP;
				const float *r1 = Lsm + ((size_t)(mt * TILE + j) * g.A + x1) * KP;
#pragma unroll
				for (int v = 0; v < KP / 4; v++) {
					const float4 p = reinterpret_cast<const float4 *>(r0)[v], q = reinterpret_cast<const float4 *>(r1)[v];
					m[4 * v] += p.x + q.x; m[4 * v + 1] += p.y + q.y; m[4 * v + 2] += p.z + q.z; m[4 * v + 3] += p.w + q.w;
				}
			}
#pragma unroll
			for (int k = 0; k < KP; k++) acc[k] += (double)m[k];     // two-level: fp32 within a micro-tile, fp64 across
		}
#pragma unroll
		for (int k = 0; k < KP; k++) a.pll[((size_t)chunk * KP + k) * Nloc + il] = acc[k] * LN2_D;
	}
}

// draw: one thread per local individual.  The chunk partials are added in chunk order
// (shard-invariant); the weights are taken relative to the largest log-likelihood instead
// of cluster 0's (mcmc.c:1110-1112) -- the same distribution without the overflow.
__global__ void na_draw_kernel(const NaArgs a)
{
	const Geometry &g = a.geo;
	const int il = blockIdx.x * blockDim.x + threadIdx.x;
	if (il >= g.Nloc) return;
	const uint32_t iter = a.iter_dev ? *a.iter_dev : a.iter;
	const int K = g.K, KP = g.KP;
	const int ig_global = g.i0 + il;
	double ll[MAX_K], cum[MAX_K];
	if (a.init) {
		for (int k = 0; k < K; k++) { ll[k] = 0.0; cum[k] = (double)(k + 1) / K; }       // mcmc.c:1107
	} else {
		const double het = (double)a.nhet[il] * LN2_D;                                     // log 2 per heterozygote, mcmc.c:1908
		double mx = -INFINITY;
		for (int k = 0; k < K; k++) {
			double s = 0.0;
			for (int c = 0; c < g.nchunks; c++) s += a.pll[((size_t)c * KP + k) * g.Nloc + il];
			ll[k] = s + het;
			mx = fmax(mx, ll[k]);
		}
		double run = 0.0;
		for (int k = 0; k < K; k++) { run += exp(ll[k] - mx); cum[k] = run; }
	}
	Stream st((uint32_t)ig_global, 0u, iter, a.init ? TAG_ZINIT : TAG_Z, a.key0, a.key1);
	const double t = st.uniform() * cum[K - 1];                   // disc_unif, random.c:403-430
	int zz = 0;
	for (int k = 0; k < K - 1; k++) zz += (t > cum[k]) ? 1 : 0;
	double *rec = a.ind + (size_t)ig_global * g.REC;
	for (int k = 0; k < K; k++) rec[k] = (k == zz) ? 1.0 : 0.0;   // the indicator: its running mean is CHAIN.z / steps (mcmc.c:1356-1362)
	rec[K] = a.init ? 0.0 : ll[zz];
	rec[K + 1] = 0.0;
	rec[K + 2] = (double)zz;
}

// tally: n[l][a][zz[i]] += 1 for both copies of every usable genotype (mcmc.c:815-845, zz branch)
template <int KP>
__global__ void __launch_bounds__(ZQ_THREADS) na_tally_kernel(const NaArgs a)
{
	extern __shared__ __align__(128) unsigned char smem_raw[];
	const Geometry &g = a.geo;
	const int tid = threadIdx.x, chunk = blockIdx.x;
	const int l0 = chunk * g.TL;
	const int nl = min(g.TL, g.Lpad - l0);
	const int nmt = nl / TILE;
	const int rowsz = g.A * KP;
	const int R = g.R;
	int *hist = reinterpret_cast<int *>(smem_raw);
	const int nbins = nl * rowsz;
	for (int j = tid; j < nbins * R; j += ZQ_THREADS) hist[j] = 0;
	__syncthreads();
	const int Nloc = g.Nloc, mt0 = l0 / TILE;
	const int nsub_total = (Nloc + ZQ_THREADS - 1) / ZQ_THREADS;
	const int sub0 = blockIdx.y * g.subs_per_blk, sub1 = min(sub0 + g.subs_per_blk, nsub_total);
	const int rep = tid & (R - 1);
	for (int sub = sub0; sub < sub1; ++sub) {
		const int il = sub * ZQ_THREADS + tid;
		if (il >= Nloc) continue;
		const int k = (int)a.ind[(size_t)(g.i0 + il) * g.REC + g.K + 2];
		const int4 *xp = reinterpret_cast<const int4 *>(a.Xt) + ((size_t)mt0 * Nloc + il) * 2;
		for (int mt = 0; mt < nmt; ++mt) {
			const int4 xa = ldg_stream(xp + (size_t)mt * Nloc * 2), xb = ldg_stream(xp + (size_t)mt * Nloc * 2 + 1);
			const int xw[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
#pragma unroll
			for (int j = 0; j < TILE; j++) {
				if (xw[j] < 0) continue;
				const int x0 = xw[j] & 0xFFFF, x1 = (xw[j] >> 16) & 0xFFFF;
				const int base = (mt * TILE + j) * g.A;
				atomicAdd(&hist[((base + x0) * KP + k) * R + rep], 1);
				atomicAdd(&hist[((base + x1) * KP + k) * R + rep], 1);
			}
		}
	}
	__syncthreads();
	int32_t *ng = a.n + (size_t)l0 * rowsz;
	for (int b = tid; b < nbins; b += ZQ_THREADS) {
		int s = 0;
		for (int r = 0; r < R; r++) s += hist[b * R + r];
		if (s) atomicAdd(ng + b, s);
	}
}

template <int KP>
static cudaError_t na_launch_big(const NaArgs &a, bool tally, cudaStream_t s)
{
	const Geometry &g = a.geo;
	dim3 grid(g.nchunks, g.nblk), block(ZQ_THREADS);
	const size_t per_locus = (size_t)g.A * KP * 4;
	const size_t sm = tally ? (size_t)g.TL * per_locus * g.R : (size_t)g.TL * per_locus;
	cudaError_t e;
	if (tally) {
		e = cudaFuncSetAttribute(na_tally_kernel<KP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
		if (e != cudaSuccess) return e;
		na_tally_kernel<KP><<<grid, block, sm, s>>>(a);
	} else {
		e = cudaFuncSetAttribute(na_scan_kernel<KP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
		if (e != cudaSuccess) return e;
		na_scan_kernel<KP><<<grid, block, sm, s>>>(a);
	}
	return cudaGetLastError();
}
static cudaError_t na_launch(const NaArgs &a, bool tally, cudaStream_t s)
{
	switch (a.geo.KP) {
	case 4: return na_launch_big<4>(a, tally, s);
	case 8: return na_launch_big<8>(a, tally, s);
	case 16: return na_launch_big<16>(a, tally, s);
	default: return cudaErrorInvalidValue;
	}
}

}  // namespace ig

using namespace ig;

static NaArgs na_args(ig_ctx *c, int init)
{
	NaArgs a;
	a.Xt = c->Xt; a.logP = c->logP; a.n = c->n; a.pll = c->na_pll; a.nhet = c->nhet; a.ind = c->ind; a.geo = c->geo;
	a.iter = c->iter; a.key0 = c->key0; a.key1 = c->key1; a.iter_dev = init ? nullptr : c->iter_dev; a.init = init;
	return a;
}

ig_status na_alloc(ig_ctx *c)
{
	const Geometry &g = c->geo;
	CK(dalloc(&c->logP, (size_t)g.Lpad * g.A * g.KP));
	CK(dalloc(&c->na_pll, (size_t)g.nchunks * g.KP * g.Nloc));
	return IG_OK;
}

// n := tally of the cluster labels held in the records (after a draw, or after state injection)
ig_status na_retally(ig_ctx *c)
{
	const Geometry &g = c->geo;
	CK(cudaMemsetAsync(c->n, 0, (size_t)g.Lpad * g.A * g.KP * sizeof(int32_t), c->stream));
	CK(na_launch(na_args(c, 0), true, c->stream));
	c->launches++;
	return IG_OK;
}

// update_Z with init_flag = 1 (mcmc.c:107): uniform labels, then their tally for the first update_P
ig_status na_chain_init(ig_ctx *c)
{
	const Geometry &g = c->geo;
	na_draw_kernel<<<(g.Nloc + 127) / 128, 128, 0, c->stream>>>(na_args(c, 1));
	CK(cudaGetLastError());
	c->launches++;
	ig_status st = ig_exchange_individuals(c);
	if (st != IG_OK) return st;
	return na_retally(c);
}

// the Z half of a mode-0 sweep: scan, draw, (all-gather of the records), tally
ig_status na_phase_z(ig_ctx *c)
{
	const Geometry &g = c->geo;
	const size_t pn = (size_t)g.Lpad * g.A * g.KP;
	na_logp_kernel<<<(unsigned)((pn + 255) / 256), 256, 0, c->stream>>>(c->P, c->logP, pn);
	CK(cudaGetLastError());
	const bool timed = c->profile && c->ev_used + 2 <= (int)c->ev.size();
	if (timed) CK(cudaEventRecord(c->ev[c->ev_used], c->stream));
	CK(na_launch(na_args(c, 0), false, c->stream));
	if (timed) { CK(cudaEventRecord(c->ev[c->ev_used + 1], c->stream)); c->ev_used += 2; }
	na_draw_kernel<<<(g.Nloc + 127) / 128, 128, 0, c->stream>>>(na_args(c, 0));
	CK(cudaGetLastError());
	c->launches += 3;
	ig_status st = ig_exchange_individuals(c);
	if (st != IG_OK) return st;
	return na_retally(c);
}
