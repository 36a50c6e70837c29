// ig_kernels.cu -- sm_100a kernels of the diploid InStruct sweep (modes 2 and 3).
//
// One sweep of the reference (mcmc.c:208-215 / :334-348) is
//     update_P -> update_S_* -> update_G -> update_ZQ -> update_alpha -> cal_lkh
// with three O(N*L) passes over the data (log_ld_indv twice in update_G, once in cal_lkh)
// plus the O(N*L*A*K) tally at the top of update_P.  Here it is
//     p_dirichlet   (L*K threads)       P | n            mcmc.c:846-857
//     pre_sweep     (1 CTA)             S | G,Q  and the G proposals   mcmc.c:913-983,864-886,1062-1084
//     zq_sweep      (the one big pass, zq_sweep.cu)  Z | P,Q ; per-individual K-counts ;
//                                       n[l][a][k] for the NEXT update_P ; the log-likelihood pieces
//                                       that the G accept and cal_lkh need   mcmc.c:1122-1194,1726-1773,810-845
//     indiv_epilogue(N threads)         G accept, Q | Z,alpha, indvlkh     mcmc.c:1085-1089,1196-1198,1931
//     post_sweep    (1 CTA)             alpha | Q, totallkh, column sums   mcmc.c:1244-1263,1940,1954-1961
//     moments       (N*K threads)       store_chn                         mcmc.c:1320-1456
// A chain sharded over GPUs by individuals (modes 1-2) replaces pre_sweep / post_sweep by sums over its LOCAL individuals
// and two kernels that exchange over NVLink peer memory (no NCCL call inside a sweep):
//     local_sums      2^K subset sums of the next update_S_POP + the post-sweep sums, fixed point
//     peer_allreduce  all-reduce of those few KB through every rank's exported arena, fused with the update_alpha tail
//     spop_decide     the K Metropolis decisions from the summed table, G proposals of the local individuals
//     p_dirichlet<PEER>  pulls its block of every rank's tally, draws, pushes P to every rank (side stream)
//
// update_G reads the OLD Z and update_ZQ never reads G, so both can ride one pass over the
// genotype store: the pass reads x (int16) and old z (int8), writes new z (int8) -- the
// 4 bytes per allele copy of SURVEY.md section 8d.
#include <math.h>
#include <stdio.h>
#include "ig_internal.h"
#include "philox.cuh"
#include <cooperative_groups.h>
#include "samplers.cuh"

namespace cg = cooperative_groups;

namespace ig {

#define LN2_D 0.69314718055994530942

// --------------------------------------------------------------------------------------
// p_dirichlet: P[k][l][.] ~ Dirichlet(n[k][l][.] + 1)  (update_P, mcmc.c:846-857, lambda = 1).
// One thread per (locus, population); writes fp32 P[l][a][k] and clears n for the next sweep.
// --------------------------------------------------------------------------------------
// PEER: the tally is the sum of every rank's n over NVLink and P goes into every rank's buffer (PeerPArgs); n is not cleared
template <bool PEER>
__device__ __forceinline__ int tally_at(const PArgs &a, const PeerPArgs &x, size_t e)
{
	if (!PEER) return a.n[e];
	int v = 0;
	for (int r = 0; r < x.W; r++) v += reinterpret_cast<const int32_t *>(reinterpret_cast<const char *>(x.peers[r]) + x.n_off)[e];
	return v;
}
template <bool PEER>
__device__ __forceinline__ void p_store(const PArgs &a, const PeerPArgs &x, size_t e, float v)
{
	if (!PEER) { a.P[e] = v; return; }
	for (int r = 0; r < x.W; r++) reinterpret_cast<float *>(reinterpret_cast<char *>(x.peers[r]) + x.p_off)[e] = v;
}
template <bool PEER>
__global__ void p_dirichlet_kernel(const PArgs a, const PeerPArgs x)
{
	const Geometry &g = a.geo;
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	const uint32_t iter = a.iter_dev ? *a.iter_dev : a.iter;
	if (t >= (a.nl > 0 ? a.nl : g.Lpad) * g.KP) return;
	const int l = a.l0 + t / g.KP, k = t % g.KP;
	if (l >= g.Lpad) return;
	const size_t base = (size_t)l * g.A * g.KP + k;
	const int Al = (l < g.L) ? a.allelenum[l] : 0;
	if (k >= g.K || Al <= 1) {
		for (int al = 0; al < g.A; al++) {
			const float v = (a.mono_ok && k < g.K && Al == 1 && al == 0) ? 1.0f : 0.0f;
			p_store<PEER>(a, x, base + (size_t)al * g.KP, v);
			if (a.Pc) snp_pc_store(a.Pc, a.tlc, g.KP, l, al, k, v);
			if (!PEER) a.n[base + (size_t)al * g.KP] = 0;
		}
		return;
	}
	Stream st((uint32_t)l, (uint32_t)k + 256u * (uint32_t)a.sub, iter, TAG_P, a.key0, a.key1);
	double sum = 0.0;
	double gam[64];
	// allelenum_max is small (2 for SNPs, tens for microsatellites); larger loci spill to a second pass
	if (Al <= 64) {
		for (int al = 0; al < Al; al++) { gam[al] = draw_gamma(st, (double)tally_at<PEER>(a, x, base + (size_t)al * g.KP) + 1.0); sum += gam[al]; }
		for (int al = 0; al < g.A; al++) {
			const double p = (al < Al) ? gam[al] / sum : 0.0;
			const float pv = (al < Al) ? fmaxf((float)p, P_FLOOR) : 0.0f;
			p_store<PEER>(a, x, base + (size_t)al * g.KP, pv);
			if (a.Pc) snp_pc_store(a.Pc, a.tlc, g.KP, l, al, k, pv);
			if (a.P64) a.P64[((size_t)k * g.L + l) * g.A + al] = p;
			if (!PEER) a.n[base + (size_t)al * g.KP] = 0;
		}
	} else {
		for (int al = 0; al < Al; al++) sum += draw_gamma(st, (double)tally_at<PEER>(a, x, base + (size_t)al * g.KP) + 1.0);
		Stream st2((uint32_t)l, (uint32_t)k + 256u * (uint32_t)a.sub, iter, TAG_P, a.key0, a.key1);     // replay the same stream
		for (int al = 0; al < g.A; al++) {
			const double p = (al < Al) ? draw_gamma(st2, (double)tally_at<PEER>(a, x, base + (size_t)al * g.KP) + 1.0) / sum : 0.0;
			p_store<PEER>(a, x, base + (size_t)al * g.KP, (al < Al) ? fmaxf((float)p, P_FLOOR) : 0.0f);
			if (a.P64) a.P64[((size_t)k * g.L + l) * g.A + al] = p;
			if (!PEER) a.n[base + (size_t)al * g.KP] = 0;
		}
	}
}
cudaError_t launch_p_dirichlet(const PArgs &a, cudaStream_t s)
{
	const int total = (a.nl > 0 ? a.nl : a.geo.Lpad) * a.geo.KP;
	p_dirichlet_kernel<false><<<(total + 127) / 128, 128, 0, s>>>(a, PeerPArgs{});
	return cudaGetLastError();
}
cudaError_t launch_p_peer_draw(const PArgs &a, const PeerPArgs &x, cudaStream_t s)
{
	const int total = a.nl * a.geo.KP;
	p_dirichlet_kernel<true><<<(total + 127) / 128, 128, 0, s>>>(a, x);
	return cudaGetLastError();
}
// one CTA: (after a system-scope fence, so that everything this rank stored before -- its tally, its block of P in the
// peers' buffers -- is visible) publish the sequence number in every rank's flag array, then wait for every rank's number
__global__ void p_peer_signal_kernel(const PeerPArgs x, int which)
{
	const int tid = threadIdx.x;
	const size_t flags = px_pflag_word(x.W, which);
	__threadfence_system();
	if (tid < x.W) {
		*(volatile unsigned long long *)(x.peers[tid] + flags + x.me) = x.seq;
		volatile unsigned long long *f = x.peers[x.me] + flags + tid;
		const long long t0 = clock64();
		while (*f < x.seq)
			if (clock64() - t0 > (1ll << 36)) __trap();
	}
	__threadfence_system();
}
cudaError_t launch_p_peer_signal(const PeerPArgs &x, int which, cudaStream_t s)
{
	p_peer_signal_kernel<<<1, 32, 0, s>>>(x, which);
	return cudaGetLastError();
}

// --------------------------------------------------------------------------------------
// Deterministic block reduction (fixed blockDim => fixed summation tree => identical on
// every rank and for every GPU count).
// --------------------------------------------------------------------------------------
constexpr int RED_THREADS = 1024;
__device__ double block_sum(double v, double *sh)
{
	const int tid = threadIdx.x;
	__syncthreads();
	sh[tid] = v;
	__syncthreads();
	for (int s = RED_THREADS / 2; s > 0; s >>= 1) {
		if (tid < s) sh[tid] += sh[tid + s];
		__syncthreads();
	}
	const double r = sh[0];
	__syncthreads();
	return r;
}

// dt_stat, mcmc.c:1524-1546 (selfing rates are always inside [0,1] here)
__device__ __forceinline__ int sel_state(double v)
{
	const double eps = 0.001;
	if (v <= eps) return 0;
	if (v >= 1.0 - eps) return 2;
	return 1;
}
// log( s^(g-1) (1-s) ): the summand of proposal() (mcmc.c:1645) and log dgeom (mcmc.c:1602)
__device__ __forceinline__ double log_geom(double s, int gen)
{
	const double l1 = log(1.0 - s);
	return (gen > 1) ? (double)(gen - 1) * log(s) + l1 : l1;
}
// q(), mcmc.c:1566-1593
__device__ __forceinline__ double trans_prob(int from, int to)
{
	if (from == 0) return (to == 0 || to == 1) ? 0.5 : 0.0;
	if (from == 2) return (to == 2 || to == 1) ? 0.5 : 0.0;
	return (to == 1) ? 0.90 : 0.05;
}

// --------------------------------------------------------------------------------------
// Deterministic sums over all individuals for the scalar updates.  update_S_POP is K sequential Metropolis steps, each
// needing a sum over all N individuals (proposal(), mcmc.c:1630), update_alpha one more: K + 2 reductions with a barrier
// each.  Two launch shapes, chosen from the GLOBAL N only (so every rank of a sharded chain takes the same one and adds in
// the same order):
//   N <= 1024  one CTA of 1024 threads: the barrier is __syncthreads (configs 1-2 are launch- and latency-bound);
//   larger     a cooperative grid of 256-thread CTAs, one individual per thread: each CTA publishes one partial per value,
//              after grid.sync() every CTA adds the partials in CTA order.  The double-precision logarithms of proposal()
//              need the SMs: a single 8-CTA cluster (tried in round 2: hardware cluster barrier, partials through
//              distributed shared memory) ran 73 us at N = 10^4 where this grid runs 47.
// The cooperative grid size is computed once per context (ig_api.cu) -- round 1 cached it in function statics, which two
// contexts on different devices raced on.
// --------------------------------------------------------------------------------------
constexpr int SC_THREADS = 256;                 // cooperative shape
constexpr int SC1_THREADS = 1024;               // single-CTA shape
constexpr int SC_MAXV = 20;                     // values reduced at once (K + 2 <= 18)

template <typename T>
__device__ __forceinline__ T warp_sum(T v)
{
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
	return v;
}

template <typename T>
struct ScSharedT {
	T warp[SC_MAXV][32];
	T total[SC_MAXV];
};
typedef ScSharedT<double> ScShared;

// reduce nv per-thread values over all threads of the launch; the result lands in out[0..nv) for every thread.
// T = double: fixed shape, fixed order.  T = long long: the fixed-point sums below, exact in any order.
template <bool COOP, typename T>
__device__ void all_sum(const T *v, int nv, T *out, ScSharedT<T> &sh, double *gpart_, int &phase)
{
	T *gpart = reinterpret_cast<T *>(gpart_);
	const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nwarps = blockDim.x >> 5;
	for (int j = 0; j < nv; j++) {
		const T w = warp_sum(v[j]);
		if (lane == 0) sh.warp[j][wid] = w;
	}
	__syncthreads();
	T *buf = COOP ? gpart + (size_t)(phase & 1) * gridDim.x * SC_MAXV : nullptr;
	if (tid < nv) {
		T t = 0;
		for (int w = 0; w < nwarps; w++) t += sh.warp[tid][w];
		if (COOP) buf[(size_t)blockIdx.x * SC_MAXV + tid] = t;
		else sh.total[tid] = t;
	}
	if (COOP) {
		__threadfence();
		cg::this_grid().sync();
		if (tid < nv) {
			T t = 0;
			for (unsigned b = 0; b < gridDim.x; b++) t += buf[(size_t)b * SC_MAXV + tid];
			sh.total[tid] = t;
		}
	}
	__syncthreads();
	for (int j = 0; j < nv; j++) out[j] = sh.total[j];
	__syncthreads();
	phase++;                                     // COOP: the partials are double-buffered, one grid.sync() per reduction suffices
}

// --------------------------------------------------------------------------------------
// Fixed-point sums.  The sums over individuals that DECIDE something (the K Metropolis steps of update_S_POP, the alpha
// step) or that are reported (totallkh, the column sums of Q) are taken over integers: every term is rounded once to a
// multiple of 2^-s and the integers are added -- exact in any order, so the result is the same for every grid shape and for
// every split of the individuals over GPUs.  That is what lets a sharded chain add its LOCAL individuals and all-reduce a
// few int64 instead of all-gathering every individual's record to every rank (ig_api.cu, local scalar updates), and stay
// bit-identical to the one-GPU chain.  -inf and NaN terms (log 0: a selfing rate of exactly 0 or 1, a q that underflowed)
// are carried as counts (low / high half of a second integer), because the reference's accept rules depend on them
// (MIN2(1, NaN) == 1, mcmc.h:10).  Ranges: |sum| < 2^63 / scale: 8.6e9 for the log-scale sums, 5.5e11 for totallkh,
// 8.4e6 individuals for the column sums.
// --------------------------------------------------------------------------------------
constexpr double FX_LL = 1073741824.0;            // 2^30: log( s^(g-1) (1-s) ) terms of proposal()
constexpr double FX_LKH = 16777216.0;             // 2^24: indvlkh (|.| ~ L)
constexpr double FX_SLQ = 268435456.0;            // 2^28: sum_k log q_ik
constexpr double FX_Q = 1099511627776.0;          // 2^40: q_ik in [0, 1]
__device__ __forceinline__ void fx_add(long long &acc, long long &cnt, double x, double scale)
{
	if (x != x) cnt += 1ll << 32;
	else if (x == INFINITY || x == -INFINITY) cnt += (x < 0.0) ? 1ll : (1ll << 32);      // +inf cannot arise from these terms; were it to, it poisons the sum like a NaN
	else acc += __double2ll_rn(x * scale);
}
__device__ __forceinline__ double fx_val(long long acc, long long cnt, double scale)
{
	if (cnt >> 32) return NAN;
	if (cnt) return -INFINITY;
	return (double)acc * (1.0 / scale);
}
// s_i = sum_k q_ik S_k, one fixed sequence of FMAs wherever it is evaluated
__device__ __forceinline__ double mix_rate(const double *rec, const double *Sv, int K)
{
	double s = 0.0;
	for (int k = 0; k < K; k++) s = fma(rec[k], Sv[k], s);
	return s;
}

// one Metropolis step of update_S_POP (mcmc.c:913-983), split so that the sequential kernel and the subset-sum kernels of the
// sharded chain consume the step's Philox stream identically: the proposal first ...
__device__ __forceinline__ void spop_propose(double sj, int cs, int back_refl, Stream &st, double &prop, int &new_state)
{
	new_state = 1;
	if (back_refl == 1) {                           // mcmc.c:939-945
		prop = sj + (st.uniform() * 2.0 * 0.05 - 0.05);
		if (prop <= 0.0) prop = -prop;
		else if (prop >= 1.0) prop = 1.0 - (prop - 1.0);
	} else {                                        // adpt_indp, mcmc.c:1461-1520
		const double u = st.uniform();
		if (cs == 0) { if (u < 0.5) { prop = 0.0; new_state = 0; } else { prop = st.uniform(); new_state = 1; } }
		else if (cs == 2) { if (u < 0.5) { prop = 1.0; new_state = 2; } else { prop = st.uniform(); new_state = 1; } }
		else { if (u <= 0.05) { prop = 0.0; new_state = 0; } else if (u >= 0.95) { prop = 1.0; new_state = 2; } else { prop = st.uniform(); new_state = 1; } }
	}
}
// ... then the accept from proposal() at the current and the proposed rates
__device__ __forceinline__ bool spop_accept_u(double cur, double pl, int cs, int new_state, int back_refl, double u)
{
	double ratio = exp(pl - cur);
	if (back_refl == 0) ratio *= trans_prob(cs, new_state) / trans_prob(new_state, cs);
	// MIN2(1, NaN) == 1 in the reference (mcmc.h:10): a NaN ratio accepts
	return (ratio != ratio) || (u < fmin(1.0, ratio));
}
__device__ __forceinline__ bool spop_accept(double cur, double pl, int cs, int new_state, int back_refl, Stream &st)
{
	return spop_accept_u(cur, pl, cs, new_state, back_refl, st.uniform());
}
// generation proposal of update_G (mcmc.c:1062-1084) for individual i (global index) at selfing rate s
__device__ __forceinline__ int g_propose(double s, int i, uint32_t iter, uint32_t key0, uint32_t key1)
{
	const int stt = sel_state(s);
	if (stt != 1) return (stt == 0) ? 1 : 50;
	Stream st((uint32_t)i, 0u, iter, TAG_GPROP, key0, key1);
	const double v = floor(log(st.uniform()) / log(s)) + 1.0;    // rgeom(1 - s), random.c:311-321
	return (v < 1.0) ? 1 : (v > 50.0 ? 50 : (int)v);
}

// --------------------------------------------------------------------------------------
// pre_sweep (either launch shape, striding over the individuals):
// selfing-rate update, then the generation proposals of update_G.
//   mode 2: update_S_POP (mcmc.c:913-983): K sequential MH steps, each a grid-wide sum of
//           log( s_i^(G_i-1) (1-s_i) ), s_i = sum_k Q_ik S_k  (proposal, mcmc.c:1630)
//   mode 3, uniform prior: update_S_IND (mcmc.c:864-886), independent per individual
//   mode 3, DP prior: S was written by the host step (ig_api.cu) before this launch
// A sharded chain in mode 3 / 4 / 5 runs this redundantly on every rank over the all-gathered (Q, G): identical inputs and
// exact (fixed-point) sums give identical S everywhere, so no broadcast is needed.  Modes 1-2 take the local path below
// (local_sums_kernel / spop_decide_kernel) instead.
// --------------------------------------------------------------------------------------
template <bool COOP>
__global__ void __launch_bounds__(COOP ? SC_THREADS : SC1_THREADS) pre_sweep_kernel(const PreArgs a)
{
	__shared__ ScShared sh;
	__shared__ double Ssh[MAX_K];
	const Geometry &g = a.geo;
	const int tid = threadIdx.x;
	const int K = g.K, REC = g.REC;
	const int gstride = gridDim.x * blockDim.x;
	const int i_first = blockIdx.x * blockDim.x + tid;
	const uint32_t iter = a.iter_dev ? *a.iter_dev : a.iter;
	int phase = 0;

	if (a.mode == 4) {
		// update_inbreedcoff_POP (mcmc.c:986-1050), proposal half: the K Metropolis steps do not
		// interact (a genotype's term depends on the F of one population only), so all K proposals
		// are made here, evaluated by the one sweep pass and accepted in post_sweep.
		if (blockIdx.x == 0 && tid < g.KP) {
			const int j = tid;
			double cur = 0.0, prop = 0.0;
			int new_state = 1;
			if (j < K) {
				cur = a.S[j];
				Stream st((uint32_t)j, 0u, iter, TAG_SPOP, a.key0, a.key1);
				if (a.back_refl == 1) {                         // mcmc.c:1012-1018
					prop = cur + (st.uniform() * 2.0 * 0.05 - 0.05);
					if (prop <= 0.0) prop = -prop;
					else if (prop >= 1.0) prop = 1.0 - (prop - 1.0);
				} else {                                        // adpt_indp, mcmc.c:1461-1520
					const int cs = a.state_in[j];
					const double u = st.uniform();
					if (cs == 0) { if (u < 0.5) { prop = 0.0; new_state = 0; } else { prop = st.uniform(); new_state = 1; } }
					else if (cs == 2) { if (u < 0.5) { prop = 1.0; new_state = 2; } else { prop = st.uniform(); new_state = 1; } }
					else { if (u <= 0.05) { prop = 0.0; new_state = 0; } else if (u >= 0.95) { prop = 1.0; new_state = 2; } else { prop = st.uniform(); new_state = 1; } }
					a.state_out[j] = new_state;
				}
				a.fprop[j] = prop;
			}
			const float omh = (float)(1.0 - cur), omhp = (float)(1.0 - prop);
			a.ftab[4 * j] = 1.0f - omh; a.ftab[4 * j + 1] = omh; a.ftab[4 * j + 2] = 1.0f - omhp; a.ftab[4 * j + 3] = omhp;
			a.ftab[4 * g.KP + j] = (float)(log2((double)omhp) - log2((double)omh));
		}
		return;
	}
	if (a.mode == 5) {
		// update_F_IND (mcmc.c:888-910), proposal half; the accept is taken by indiv_epilogue
		for (int i = i_first; i < g.N; i += gstride) {
			const double f = a.S[i];
			Stream st((uint32_t)i, 0u, iter, TAG_SIND, a.key0, a.key1);
			double prop = f + (st.uniform() * 2.0 * 0.05 - 0.05);
			if (prop <= 0.0) prop = -prop;
			if (prop >= 1.0) prop = 1.0 - (prop - 1.0);
			a.fprop[i] = prop;
			const int il = i - g.i0;
			if (il >= 0 && il < g.Nloc) a.hpair[il] = make_float2((float)(1.0 - f), (float)(1.0 - prop));
		}
		return;
	}
	if (a.mode == 2) {
		ScSharedT<long long> &shl = reinterpret_cast<ScSharedT<long long> &>(sh);
		if (tid < K) Ssh[tid] = a.S[tid];
		__syncthreads();
		double Sl[MAX_K], Sp[MAX_K];
		for (int k = 0; k < K; k++) Sl[k] = Sp[k] = Ssh[k];
		long long part[2] = {0, 0}, tot[2];
		for (int i = i_first; i < g.N; i += gstride) {
			const double *rec = a.ind + (size_t)i * REC;
			fx_add(part[0], part[1], log_geom(mix_rate(rec, Sl, K), (int)rec[K + 2]), FX_LL);
		}
		all_sum<COOP>(part, 2, tot, shl, a.gpart, phase);
		double cur = fx_val(tot[0], tot[1], FX_LL);
		int accepts = 0;
		for (int j = 0; j < K; j++) {
			Stream st((uint32_t)j, 0u, iter, TAG_SPOP, a.key0, a.key1);
			double prop;
			int new_state;
			const int cs = (a.back_refl == 0) ? a.state_in[j] : 1;
			spop_propose(Sl[j], cs, a.back_refl, st, prop, new_state);
			Sp[j] = prop;
			part[0] = part[1] = 0;
			for (int i = i_first; i < g.N; i += gstride) {
				const double *rec = a.ind + (size_t)i * REC;
				fx_add(part[0], part[1], log_geom(mix_rate(rec, Sp, K), (int)rec[K + 2]), FX_LL);
			}
			all_sum<COOP>(part, 2, tot, shl, a.gpart, phase);
			const double pl = fx_val(tot[0], tot[1], FX_LL);
			// every thread of every CTA takes the same decision from the same numbers
			const bool acc = spop_accept(cur, pl, cs, new_state, a.back_refl, st);
			if (acc) {
				cur = pl;
				Sl[j] = prop;
				accepts++;
				if (blockIdx.x == 0 && tid == 0) { a.S[j] = prop; if (a.back_refl == 0) a.state_out[j] = new_state; }
			} else {
				Sp[j] = Sl[j];
				if (blockIdx.x == 0 && tid == 0 && a.back_refl == 0) a.state_out[j] = cs;
			}
		}
		if (blockIdx.x == 0 && tid == 0) { a.sc->cur_prop_ll = cur; a.sc->s_accepts += accepts; }
		// ---- generation proposals with the UPDATED S (update_G follows update_S_POP, mcmc.c:211-212)
		for (int i = i_first; i < g.N; i += gstride) {
			const double *rec = a.ind + (size_t)i * REC;
			const int gp = g_propose(mix_rate(rec, Sl, K), i, iter, a.key0, a.key1);
			a.gprop[i] = gp;
			const int il = i - g.i0;
			if (il >= 0 && il < g.Nloc) a.gpair[il] = make_int2((int)rec[K + 2], gp);
		}
		return;
	}
	// ---- mode 3
	for (int i = i_first; i < g.N; i += gstride) {
		double s = a.S[i];
		const int gen = (int)a.ind[(size_t)i * REC + K + 2];
		if (a.prior_flag == 0) {                             // update_S_IND, mcmc.c:864-886
			Stream st((uint32_t)i, 0u, iter, TAG_SIND, a.key0, a.key1);
			double prop = s + (st.uniform() * 2.0 * 0.05 - 0.05);
			if (prop <= 0.0) prop = -prop;
			if (prop >= 1.0) prop = 1.0 - (prop - 1.0);
			const double ratio = exp(log_geom(prop, gen) - log_geom(s, gen));
			const double u = st.uniform();
			if ((ratio != ratio) || u < fmin(1.0, ratio)) { s = prop; a.S[i] = prop; }
		}
		const int stt = sel_state(s);
		int gp;
		if (stt == 1) {
			Stream st((uint32_t)i, 0u, iter, TAG_GPROP, a.key0, a.key1);
			const double v = floor(log(st.uniform()) / log(s)) + 1.0;
			gp = (v < 1.0) ? 1 : (v > 50.0 ? 50 : (int)v);
		} else gp = (stt == 0) ? 1 : 50;
		a.gprop[i] = gp;
		const int il = i - g.i0;
		if (il >= 0 && il < g.Nloc) a.gpair[il] = make_int2(gen, gp);
	}
}

cudaError_t launch_pre_sweep(const PreArgs &a, cudaStream_t s)
{
	if (a.geo.N <= SC1_THREADS) { pre_sweep_kernel<false><<<1, SC1_THREADS, 0, s>>>(a); return cudaGetLastError(); }
	if (a.mode != 2) {       // only update_S_POP has sums over the individuals: the other modes are independent per individual -- a plain grid
		pre_sweep_kernel<false><<<(a.geo.N + SC1_THREADS - 1) / SC1_THREADS, SC1_THREADS, 0, s>>>(a);
		return cudaGetLastError();
	}
	void *args[] = {(void *)&a};
	return cudaLaunchCooperativeKernel((const void *)pre_sweep_kernel<true>, dim3(a.grid), dim3(SC_THREADS), args, 0, s);
}

// --------------------------------------------------------------------------------------
// indiv_epilogue: CTA = 8 individuals x 32 chunk groups (a shard of 1250 individuals gets 157 CTAs; with the 32 x 8 shape
// of round 1 it got 40 and took 85 us on 148 SMs).  Thread row w adds the partials of chunks w, w+32, ... , the group sums
// are added in group order, then one thread per individual takes the update_G accept decision (mcmc.c:1085-1089), draws
// Q_i ~ Dirichlet(cnt_i + alpha) (mcmc.c:1196-1198) and writes the individual's record
// (Q, indvlkh, sum_k log q, G) into the all-gatherable array.  The summation order depends
// only on the chunk decomposition, which depends only on (L, K, A): shard-invariant.
// --------------------------------------------------------------------------------------
constexpr int EPI_THREADS = 256;
template <int EPI_IPB>              // individuals per CTA: 8 (small shards: more CTAs) or 32; chunk groups = 256 / EPI_IPB
__global__ void __launch_bounds__(EPI_THREADS, EPI_IPB == 32 ? 3 : 2) indiv_epilogue_kernel(const EpiArgs a)
{
	constexpr int EPI_GROUPS = EPI_THREADS / EPI_IPB;
	__shared__ int cnt_sh[EPI_GROUPS][MAX_K][EPI_IPB];
	__shared__ double ll_sh[EPI_GROUPS][3][EPI_IPB];
	__shared__ int nsh_sh[EPI_GROUPS][EPI_IPB];
	__shared__ double gq_sh[MAX_K][EPI_IPB];
	__shared__ int cntt_sh[MAX_K][EPI_IPB];
	__shared__ double llt_sh[3][EPI_IPB];
	__shared__ int nsht_sh[EPI_IPB];
	const Geometry &g = a.geo;
	const int lane = threadIdx.x % EPI_IPB, w = threadIdx.x / EPI_IPB;
	const int il = blockIdx.x * EPI_IPB + lane;
	const uint32_t iter = a.iter_dev ? *a.iter_dev : a.iter;
	const int K = g.K, KP = g.KP;
	const bool live = il < g.Nloc;
	int cnt[MAX_K];
#pragma unroll
	for (int k = 0; k < MAX_K; k++) cnt[k] = 0;
	// The floating-point sums of an individual are taken in an order that depends on the chunk decomposition only (so a
	// chain is bit-identical for every shard count and either CTA shape): chunks c = v (mod 32) form virtual group v, summed
	// in chunk order; T_t = ((s_t + s_t+8) + s_t+16) + s_t+24 for t = 0 .. 7; the T are folded in order.  The 8-row shape
	// walks its four virtual groups in ONE loop with four accumulator sets, the 32-row shape has one group per row.
	constexpr int VG = 32, VPR = VG / EPI_GROUPS;      // virtual groups, and how many of them one thread row walks (1 or 4)
	double dv[VPR], av[VPR], bv[VPR];
#pragma unroll
	for (int vv = 0; vv < VPR; vv++) dv[vv] = av[vv] = bv[vv] = 0.0;
	int nsh_new = 0;
	if (live) {
		for (int c0 = w; c0 < g.nchunks; c0 += VG) {
#pragma unroll
			for (int vv = 0; vv < VPR; vv++) {
				const int c = c0 + vv * EPI_GROUPS;
				if (c >= g.nchunks) break;
				// KP u16 counters per (chunk, individual): 8, 16 or 32 bytes, read as 64 / 128-bit vectors
				const uint16_t *pcb = a.pcnt + ((size_t)c * g.Nloc + il) * KP;
#define EPI_ADD(J, V) do { cnt[2 * (J)] += (V) & 0xFFFFu; cnt[2 * (J) + 1] += (V) >> 16; } while (0)
				if (KP == 4) { const uint2 t = *reinterpret_cast<const uint2 *>(pcb); EPI_ADD(0, t.x); EPI_ADD(1, t.y); }
				else {
					const uint4 t = *reinterpret_cast<const uint4 *>(pcb);
					EPI_ADD(0, t.x); EPI_ADD(1, t.y); EPI_ADD(2, t.z); EPI_ADD(3, t.w);
					if (KP == 16) { const uint4 t2 = *(reinterpret_cast<const uint4 *>(pcb) + 1); EPI_ADD(4, t2.x); EPI_ADD(5, t2.y); EPI_ADD(6, t2.z); EPI_ADD(7, t2.w); }
				}
#undef EPI_ADD
				const double *pl = a.plog + (size_t)c * 3 * g.Nloc + il;
				dv[vv] += pl[0];
				av[vv] += pl[(size_t)g.Nloc];
				bv[vv] += pl[(size_t)2 * g.Nloc];
				nsh_new += a.pnsh[(size_t)c * g.Nloc + il];
			}
		}
	}
	double d_old = dv[0], a_new = av[0], b_new = bv[0];     // VPR == 4: this row's T; VPR == 1: one virtual group, folded below
#pragma unroll
	for (int vv = 1; vv < VPR; vv++) { d_old += dv[vv]; a_new += av[vv]; b_new += bv[vv]; }
#pragma unroll
	for (int k = 0; k < MAX_K; k++) cnt_sh[w][k][lane] = cnt[k];
	ll_sh[w][0][lane] = d_old; ll_sh[w][1][lane] = a_new; ll_sh[w][2][lane] = b_new;
	nsh_sh[w][lane] = nsh_new;
	__syncthreads();
	// ---- group sums in group order (the order depends only on the chunk decomposition: shard-invariant), spread over
	//      the CTA: thread (k, individual) adds population k's counts and draws its gamma of Q_i ~ Dirichlet(cnt_i + alpha)
	//      (mcmc.c:1195-1197; double-precision Marsaglia-Tsang, own Philox stream per (individual, population)); three more
	//      thread rows add the likelihood pieces, one the heterozygote count
	const int ig_global = g.i0 + il;
	if (live)
		for (int task = w; task < K + 4; task += EPI_GROUPS) {
			if (task < K) {
				int ck = 0;
				for (int ww = 0; ww < EPI_GROUPS; ww++) ck += cnt_sh[ww][task][lane];
				cntt_sh[task][lane] = ck;
				Stream sq((uint32_t)ig_global, (uint32_t)task, iter, TAG_Q, a.key0, a.key1);
				gq_sh[task][lane] = draw_gamma(sq, (double)ck + a.sc->alpha);
			} else if (task < K + 3) {
				double t = 0.0;
				for (int tw = 0; tw < 8; tw++) {
					double tt = ll_sh[tw][task - K][lane];
					for (int r = 1; r < EPI_GROUPS / 8; r++) tt += ll_sh[tw + 8 * r][task - K][lane];
					t += tt;
				}
				llt_sh[task - K][lane] = t;
			} else {
				int t = 0;
				for (int ww = 0; ww < EPI_GROUPS; ww++) t += nsh_sh[ww][lane];
				nsht_sh[lane] = t;
			}
		}
	__syncthreads();
	if (w != 0 || !live) return;
#pragma unroll
	for (int k = 0; k < MAX_K; k++) cnt[k] = (k < K) ? cntt_sh[k][lane] : 0;
	d_old = llt_sh[0][lane]; a_new = llt_sh[1][lane]; b_new = llt_sh[2][lane];
	nsh_new = nsht_sh[lane];
	// heterozygotes contribute ln 2 each (genofreq, mcmc.c:1700); the count is data only
	const double c_new = (double)a.nhet[il] * LN2_D;
	double *rec = a.ind + (size_t)ig_global * g.REC;
	if (!a.init && a.fmode == 2) {
		rec[K] = c_new + a_new;                                // at the current F; post_sweep adds the accepted differences
		rec[K + 2] = 0.0;
	} else if (!a.init && a.fmode == 1) {
		// update_F_IND accept (mcmc.c:905-907) and cal_lkh at the accepted F (log_ld_F_indv, mcmc.c:1812)
		const double f = a.S[ig_global], fp = a.fprop[ig_global];
		const double lo = log((double)(float)(1.0 - f)), lp = log((double)(float)(1.0 - fp));   // the sweep kernel's fp32 1 - F
		const int no = a.nsh[il];
		if (no) d_old += (double)no * (lp - lo);
		if (a.llparts) { double *lq = a.llparts + (size_t)il * 4; lq[0] = d_old; lq[1] = c_new; lq[2] = a_new; lq[3] = b_new; }
		Stream sa((uint32_t)ig_global, 0u, iter, TAG_GACC, a.key0, a.key1);
		const double u = sa.uniform();
		const double ratio = exp(d_old);
		const bool acc = (ratio != ratio) || (u < fmin(1.0, ratio));
		rec[K + 2] = acc ? fp : f;
		rec[K] = c_new + (acc ? b_new : a_new);
	} else if (!a.init) {
		const int2 gg = a.gpair[il];
		// same-z heterozygotes carry 2^-(g-1) (genofreq, mcmc.c:1692-1699): on the OLD Z their
		// count is what the previous pass left in nsh (type_freq 0: already inside d_old)
		if (a.type_freq != 0) d_old -= (double)a.nsh[il] * (double)(gg.y - gg.x) * LN2_D;
		if (a.llparts) { double *lp = a.llparts + (size_t)il * 4; lp[0] = d_old; lp[1] = c_new; lp[2] = a_new; lp[3] = b_new; }
		Stream sa((uint32_t)ig_global, 0u, iter, TAG_GACC, a.key0, a.key1);
		const double u = sa.uniform();
		const double ratio = exp(d_old);
		const bool acc = (ratio != ratio) || (u < fmin(1.0, ratio));
		rec[K + 2] = (double)(acc ? gg.y : gg.x);
		rec[K] = c_new + (acc ? b_new : a_new);
	}
	double qv[MAX_K], sum = 0.0;
#pragma unroll
	for (int k = 0; k < MAX_K; k++)
		if (k < K) { qv[k] = gq_sh[k][lane]; sum += qv[k]; }
	double slq = 0.0;
#pragma unroll
	for (int k = 0; k < MAX_K; k++)
		if (k < K) {
			const double qk = qv[k] / sum;
			rec[k] = qk;
			slq += log(qk);
			a.Qf[(size_t)il * KP + k] = (float)qk;
			a.cnt[(size_t)il * K + k] = cnt[k];
		}
	for (int k = K; k < KP; k++) a.Qf[(size_t)il * KP + k] = 0.0f;
	rec[K + 1] = slq;
	a.nsh[il] = nsh_new;
}
cudaError_t launch_epilogue(const EpiArgs &a, cudaStream_t s)
{
	// 32 individuals per CTA once that still gives every SM a few CTAs; 8 for the small shards of a many-GPU chain (the
	// summation order is the same for both shapes, see the kernel)
	if (a.geo.Nloc >= 8192) indiv_epilogue_kernel<32><<<(a.geo.Nloc + 31) / 32, EPI_THREADS, 0, s>>>(a);
	else indiv_epilogue_kernel<8><<<(a.geo.Nloc + 7) / 8, EPI_THREADS, 0, s>>>(a);
	return cudaGetLastError();
}

// mode 4: thread = (individual, j), j < 2 KP: sums the sweep kernel's per-population partials
// over the chunks in chunk order (shard-invariant) into the record, D_k at [K+3+k] (old Z,
// for the K accepts) and E_k at [2K+3+k] (new Z, for cal_lkh at the accepted F), in nats.
__global__ void fk_epilogue_kernel(const EpiArgs a)
{
	const Geometry &g = a.geo;
	const int KP2 = 2 * g.KP;
	const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= (long)g.Nloc * KP2) return;
	const int il = (int)(t / KP2), j = (int)(t % KP2);
	const int k = j % g.KP, which = j / g.KP;
	if (k >= g.K) return;
	double s = 0.0;
	for (int c = 0; c < g.nchunks; c++) s += (double)a.pfk[((size_t)c * g.Nloc + il) * KP2 + j];
	a.ind[(size_t)(g.i0 + il) * g.REC + g.K + 3 + which * g.K + k] = s * LN2_D;
}
cudaError_t launch_fk_epilogue(const EpiArgs &a, cudaStream_t s)
{
	const long n = (long)a.geo.Nloc * 2 * a.geo.KP;
	fk_epilogue_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(a);
	return cudaGetLastError();
}

// --------------------------------------------------------------------------------------
// post_sweep (either launch shape, as pre_sweep): totallkh (cal_lkh, mcmc.c:1940), the alpha MH step
// (update_alpha, mcmc.c:1244-1263, in log form: (alpha'-alpha) * sum log q), and the
// column sums of Q for check_empty_cluster (mcmc.c:1954-1961) -- one grid-wide sum of
// K + 2 values.
// --------------------------------------------------------------------------------------
// the K + 2 sums of post_sweep in fixed point: v[0] totallkh, v[1] sum log q, v[2 + k] column k of Q, v[K + 2] / v[K + 3]
// the -inf / NaN counts of the first two
__device__ __forceinline__ void post_terms(const double *rec, int K, long long *v)
{
	fx_add(v[0], v[K + 2], rec[K], FX_LKH);
	fx_add(v[1], v[K + 3], rec[K + 1], FX_SLQ);
#pragma unroll
	for (int k = 0; k < MAX_K; k++)
		if (k < K) v[2 + k] += __double2ll_rn(rec[k] * FX_Q);
}
__device__ __forceinline__ void post_totals(const long long *t, int K, double *tot)
{
	tot[0] = fx_val(t[0], t[K + 2], FX_LKH);
	tot[1] = fx_val(t[1], t[K + 3], FX_SLQ);
	for (int k = 0; k < K; k++) tot[2 + k] = (double)t[2 + k] * (1.0 / FX_Q);
}

__device__ void post_finish(const PostArgs &a, const double *tot, uint32_t iter);

template <bool COOP>
__global__ void __launch_bounds__(COOP ? SC_THREADS : SC1_THREADS) post_sweep_kernel(const PostArgs a)
{
	__shared__ ScShared sh;
	const Geometry &g = a.geo;
	const int tid = threadIdx.x, K = g.K, REC = g.REC;
	const int gstride = gridDim.x * blockDim.x;
	const uint32_t iter = a.iter_dev ? *a.iter_dev : a.iter;
	double tot[SC_MAXV];
	int phase = 0;
	{
		ScSharedT<long long> &shl = reinterpret_cast<ScSharedT<long long> &>(sh);
		long long v[SC_MAXV], t[SC_MAXV];
#pragma unroll
		for (int j = 0; j < SC_MAXV; j++) v[j] = 0;
		for (int i = blockIdx.x * blockDim.x + tid; i < g.N; i += gstride) post_terms(a.ind + (size_t)i * REC, K, v);
		all_sum<COOP>(v, K + 4, t, shl, a.gpart, phase);
		post_totals(t, K, tot);
	}
	if (a.mode == 5)                                        // the accepted F of every individual, from the (all-gathered) records
		for (int i = blockIdx.x * blockDim.x + tid; i < g.N; i += gstride) a.S[i] = a.ind[(size_t)i * REC + K + 2];
	if (a.mode == 4 && iter != 0xFFFFFFFFu) {
		// update_inbreedcoff_POP accepts (mcmc.c:1038-1047): D_k = sum over individuals of the old-Z
		// differences; the reference multiplies that LOG ratio by the Hastings ratio under -e 0 and
		// accepts when u < exp(min(1, .)) -- reproduced as written (NaN accepts: MIN2(1, NaN) == 1).
		double Fc[MAX_K], Fp[MAX_K];
		int cs[MAX_K], ns_[MAX_K];
		for (int k = 0; k < K; k++) {
			Fc[k] = a.S[k]; Fp[k] = a.fprop[k];
			cs[k] = (a.back_refl == 0) ? a.state[k] : 1;
			ns_[k] = (a.back_refl == 0) ? a.state_prop[k] : 1;
		}
		double dv[SC_MAXV], dt[SC_MAXV];
#pragma unroll
		for (int j = 0; j < SC_MAXV; j++) dv[j] = 0.0;
		for (int i = blockIdx.x * blockDim.x + tid; i < g.N; i += gstride) {
			const double *rec = a.ind + (size_t)i * REC + K + 3;
#pragma unroll
			for (int k = 0; k < MAX_K; k++) if (k < K) dv[k] += rec[k];
		}
		all_sum<COOP>(dv, K, dt, sh, a.gpart, phase);
		bool acc[MAX_K];
		int naccept = 0;
		for (int k = 0; k < K; k++) {
			double mh = dt[k];
			if (a.back_refl == 0) mh *= trans_prob(cs[k], ns_[k]) / trans_prob(ns_[k], cs[k]);
			Stream st((uint32_t)k, 1u, iter, TAG_SPOP, a.key0, a.key1);
			const double u = st.uniform();
			acc[k] = (mh != mh) || (u < exp(fmin(1.0, mh)));
			naccept += acc[k] ? 1 : 0;
		}
		// cal_lkh at the accepted coefficients (log_ld_F_pop, mcmc.c:1776): add the new-Z differences
		double lv = 0.0, lt;
		for (int i = blockIdx.x * blockDim.x + tid; i < g.N; i += gstride) {
			double *rec = a.ind + (size_t)i * REC;
			double l = rec[K];
			for (int k = 0; k < K; k++) if (acc[k]) l += rec[2 * K + 3 + k];
			rec[K] = l;
			lv += l;
		}
		all_sum<COOP>(&lv, 1, &lt, sh, a.gpart, phase);
		tot[0] = lt;
		if (blockIdx.x == 0 && tid == 0) {
			for (int k = 0; k < K; k++) if (acc[k]) { a.S[k] = Fp[k]; if (a.back_refl == 0) a.state[k] = ns_[k]; }
			a.sc->s_accepts += naccept;
		}
	}
	if (blockIdx.x != 0 || tid != 0) return;
	post_finish(a, tot, iter);
}

// the single-thread tail of post_sweep: publish the totals, then update_alpha
__device__ void post_finish(const PostArgs &a, const double *tot, uint32_t iter)
{
	const int K = a.geo.K;
	const double slq = tot[1];
	a.sc->totallkh = tot[0];
	a.sc->sumlogq = slq;
	for (int k = 0; k < K; k++) a.sc->qcol[k] = tot[2 + k];
	if (iter == 0xFFFFFFFFu) return;                        // statistics only (parity hook)
	a.sc->iter = iter + 1;                                  // the next sweep's counter, for graph replay
	if (a.mode == 0) return;                                // no admixture: there is no alpha (mcmc.c:90-131)
	Stream st(0u, 0u, iter, TAG_ALPHA, a.key0, a.key1);
	const double alpha = a.sc->alpha;
	const double ralpha = alpha + draw_normal(st);
	if (ralpha > 0.0) {
		// The reference multiplies pow(q, alpha'+n)/pow(q, n+alpha) over all (i,k) (mcmc.c:1258).
		// A q that underflowed to exactly 0 makes one factor 0/0 = NaN, and MIN2(1, NaN) == 1
		// (mcmc.h:10), so the proposal is accepted.  With near-pure ancestry alpha drifts
		// towards 0 (the ratio has no Gamma normaliser) until this happens, so the rule is
		// what sets alpha's long-run distribution there; it is reproduced: sum log q = -inf
		// <=> some q == 0 <=> accept.
		const double ratio = exp((ralpha - alpha) * slq);
		const double u = st.uniform();
		const bool nan_accept = (slq != slq) || (slq == -INFINITY);
		if (nan_accept || u < fmin(1.0, ratio)) { a.sc->alpha = ralpha; a.sc->alpha_accepts++; }
	}
}
// --------------------------------------------------------------------------------------
// Local scalar updates of a sharded chain (modes 1 and 2; ig_api.cu local_scalars).  A rank holds the records of its own
// individuals only.  update_S_POP is K sequential Metropolis steps, but step j's proposal does not depend on the outcome of
// the earlier steps (it perturbs S_j, which only step j changes) -- only the sums do, through WHICH of S_0 .. S_j-1 were
// replaced.  So the sum of proposal() is taken for every subset B of replaced populations at once (2^K sums over the local
// individuals, one launch), ONE int64 all-reduce makes them global and exact, and a single thread per CTA then walks the K
// decisions through the table: cur = sum[B], proposed = sum[B | 1 << j].  Same Philox streams, same fixed-point sums as the
// sequential kernel: bit-identical S.
// --------------------------------------------------------------------------------------
constexpr int TREE_THREADS = 256;
__device__ __forceinline__ void spop_tree_body(const TreeArgs &a, const unsigned B)
{
	__shared__ double Sc[MAX_K];
	__shared__ long long wsum[2][TREE_THREADS / 32];
	const Geometry &g = a.geo;
	const int tid = threadIdx.x, K = g.K;
	if (tid < K) {
		double sv = a.S[tid];
		if ((B >> tid) & 1u) {
			Stream st((uint32_t)tid, 0u, a.iter, TAG_SPOP, a.key0, a.key1);
			int ns;
			spop_propose(sv, (a.back_refl == 0) ? a.state[tid] : 1, a.back_refl, st, sv, ns);
		}
		Sc[tid] = sv;
	}
	__syncthreads();
	double Sv[MAX_K];
	for (int k = 0; k < K; k++) Sv[k] = Sc[k];
	long long v = 0, cn = 0;
	const int il = blockIdx.x * TREE_THREADS + tid;
	if (il < g.Nloc) {
		const double *rec = a.ind + (size_t)(g.i0 + il) * g.REC;
		fx_add(v, cn, log_geom(mix_rate(rec, Sv, K), (int)rec[K + 2]), FX_LL);
	}
	v = warp_sum(v); cn = warp_sum(cn);
	if ((tid & 31) == 0) { wsum[0][tid >> 5] = v; wsum[1][tid >> 5] = cn; }
	__syncthreads();
	if (tid < 2) {
		long long t = 0;
		for (int w = 0; w < TREE_THREADS / 32; w++) t += wsum[tid][w];
		if (t) atomicAdd(a.acc + 2 * (size_t)B + tid, (unsigned long long)t);
	}
}
__global__ void __launch_bounds__(TREE_THREADS) spop_decide_kernel(const TreeArgs a)
{
	// every CTA repeats the K decisions (they are a function of the all-reduced table and the step's Philox streams), then
	// proposes G for its own 256 individuals.  The table goes to shared memory first and thread j prepares step j's proposal,
	// so the sequential walk of thread 0 is K times (two shared loads, one exp, one compare) -- it was 11 us of dependent
	// global loads and Philox set-ups when one thread did everything.
	extern __shared__ long long tab[];                    // [2^K][2]
	__shared__ double Sn[MAX_K], prop_s[MAX_K], Sold[MAX_K];
	__shared__ int cs_s[MAX_K], ns_s[MAX_K];
	__shared__ double u_s[MAX_K];
	const Geometry &g = a.geo;
	const int tid = threadIdx.x, K = g.K;
	const long long *acc = reinterpret_cast<const long long *>(a.acc);
	for (int t = tid; t < (2 << K); t += TREE_THREADS) tab[t] = acc[t];
	if (tid < K) {
		Stream st((uint32_t)tid, 0u, a.iter, TAG_SPOP, a.key0, a.key1);
		const double sj = a.S[tid];
		const int cs = (a.back_refl == 0) ? a.state[tid] : 1;
		double prop;
		int new_state;
		spop_propose(sj, cs, a.back_refl, st, prop, new_state);
		Sold[tid] = sj; prop_s[tid] = prop; cs_s[tid] = cs; ns_s[tid] = new_state;
		u_s[tid] = st.uniform();                          // the step's accept draw: next in the same stream (spop_accept's order)
	}
	__syncthreads();
	if (tid == 0) {
		unsigned B = 0;
		double cur = fx_val(tab[0], tab[1], FX_LL);
		int accepts = 0;
		for (int j = 0; j < K; j++) {
			const unsigned Bp = B | (1u << j);
			const double pl = fx_val(tab[2 * Bp], tab[2 * Bp + 1], FX_LL);
			const bool ok = spop_accept_u(cur, pl, cs_s[j], ns_s[j], a.back_refl, u_s[j]);
			if (ok) { B = Bp; cur = pl; accepts++; }
			Sn[j] = ok ? prop_s[j] : Sold[j];
			if (blockIdx.x == 0) {
				a.S_out[j] = Sn[j];
				if (a.back_refl == 0) a.state_out[j] = ok ? ns_s[j] : cs_s[j];
			}
		}
		if (blockIdx.x == 0) { a.sc->cur_prop_ll = cur; a.sc->s_accepts += accepts; }
	}
	__syncthreads();
	double Sv[MAX_K];
	for (int k = 0; k < K; k++) Sv[k] = Sn[k];
	const int il = blockIdx.x * TREE_THREADS + tid;
	if (il >= g.Nloc) return;
	const int i = g.i0 + il;
	const double *rec = a.ind + (size_t)i * g.REC;
	const int gp = g_propose(mix_rate(rec, Sv, K), i, a.iter, a.key0, a.key1);
	a.gprop[i] = gp;
	a.gpair[il] = make_int2((int)rec[K + 2], gp);
}
cudaError_t launch_spop_decide(const TreeArgs &a, cudaStream_t s)
{
	spop_decide_kernel<<<(a.geo.Nloc + TREE_THREADS - 1) / TREE_THREADS, TREE_THREADS, (size_t)(2 << a.geo.K) * sizeof(long long), s>>>(a);
	return cudaGetLastError();
}

// post_sweep in two halves around an int64 all-reduce: the local individuals' K + 4 fixed-point sums, then the tail
__device__ __forceinline__ void post_local_body(const PostArgs &a, unsigned long long *acc)
{
	__shared__ long long wsum[SC_MAXV][TREE_THREADS / 32];
	const Geometry &g = a.geo;
	const int tid = threadIdx.x, K = g.K;
	long long v[SC_MAXV];
#pragma unroll
	for (int j = 0; j < SC_MAXV; j++) v[j] = 0;
	for (int il = blockIdx.x * TREE_THREADS + tid; il < g.Nloc; il += gridDim.x * TREE_THREADS)
		post_terms(a.ind + (size_t)(g.i0 + il) * g.REC, K, v);
	for (int j = 0; j < K + 4; j++) {
		const long long w = warp_sum(v[j]);
		if ((tid & 31) == 0) wsum[j][tid >> 5] = w;
	}
	__syncthreads();
	if (tid < K + 4) {
		long long t = 0;
		for (int w = 0; w < TREE_THREADS / 32; w++) t += wsum[tid][w];
		if (t) atomicAdd(acc + tid, (unsigned long long)t);
	}
}
// one launch for all local sums of a sweep: rows [0, n_tree) of the grid are the subset sums of the NEXT update_S_POP (t),
// the last row the post-sweep sums of this sweep (a) -- either part may be absent
__global__ void __launch_bounds__(TREE_THREADS) local_sums_kernel(const TreeArgs t, const PostArgs a, unsigned long long *post_acc, int n_tree)
{
	if ((int)blockIdx.y < n_tree) spop_tree_body(t, blockIdx.y);
	else post_local_body(a, post_acc);
}
cudaError_t launch_local_sums(const TreeArgs *t, const PostArgs *a, unsigned long long *post_acc, cudaStream_t s)
{
	const Geometry &g = t ? t->geo : a->geo;
	const int n_tree = t ? (1 << g.K) : 0;
	const dim3 grid((g.Nloc + TREE_THREADS - 1) / TREE_THREADS, n_tree + (a ? 1 : 0));
	TreeArgs tz{};
	PostArgs az{};
	local_sums_kernel<<<grid, TREE_THREADS, 0, s>>>(t ? *t : tz, a ? *a : az, post_acc, n_tree);
	return cudaGetLastError();
}

__global__ void post_final_kernel(const PostArgs a, const unsigned long long *acc)
{
	if (threadIdx.x != 0 || blockIdx.x != 0) return;
	double tot[SC_MAXV];
	post_totals(reinterpret_cast<const long long *>(acc), a.geo.K, tot);
	post_finish(a, tot, a.iter);
}
cudaError_t launch_post_final(const PostArgs &a, const unsigned long long *acc, cudaStream_t s)
{
	post_final_kernel<<<1, 32, 0, s>>>(a, acc);
	return cudaGetLastError();
}

// --------------------------------------------------------------------------------------
// peer_allreduce: the exchange of a sharded sweep as ONE kernel over NVLink peer memory.  One CTA: (1) stores this rank's
// local sums into its slot of EVERY rank's buffer (plain stores through the IPC mappings; a few KB to each of W - 1 peers
// over NVSwitch), system-scope fence, then its sequence number behind them; (2) spins on its own buffer until every
// rank's sequence number has arrived; (3) adds the W slots -- integers, so any order gives the same bits -- and (4) runs the
// tail of post_sweep (totallkh, update_alpha) on the totals when asked to.  An NCCL all-reduce of the same 4 KB costs ~30 us
// of protocol latency per sweep at 8 GPUs; this is one launch and two NVLink one-way trips.  Buffers alternate by the
// parity of the sequence number: a rank can be at most one all-reduce ahead of the slowest (it needs everyone's number to
// leave), so the slots it overwrites are never ones a peer still has to read.
// --------------------------------------------------------------------------------------
constexpr int PX_THREADS = 512;
__global__ void __launch_bounds__(PX_THREADS) peer_allreduce_kernel(const PeerArgs x, const PostArgs a)
{
	const int tid = threadIdx.x, W = x.W;
	const int par = (int)(x.seq & 1ull);
	const size_t slot = ((size_t)par * W + x.me) * PX_WORDS;
	const size_t flags = (size_t)2 * W * PX_WORDS + (size_t)par * W;
	for (int w = tid; w < x.nwords; w += PX_THREADS) {
		const unsigned long long v = x.acc[w];
		x.acc[w] = 0;                                          // the local accumulators are handed back empty: no memset per sweep
		for (int r = 0; r < W; r++) *(volatile unsigned long long *)(x.peers[r] + slot + w) = v;
	}
	__threadfence_system();
	__syncthreads();
	if (tid < W) *(volatile unsigned long long *)(x.peers[tid] + flags + x.me) = x.seq;
	unsigned long long *mine = x.peers[x.me];
	if (tid < W) {
		volatile unsigned long long *f = mine + flags + tid;
		const long long t0 = clock64();
		while (*f != x.seq)
			if (clock64() - t0 > (1ll << 36)) __trap();        // ~35 s: a peer died; fail the context loudly instead of hanging the GPU
	}
	__threadfence_system();
	__syncthreads();
	for (int w = tid; w < x.nwords; w += PX_THREADS) {
		unsigned long long sum = 0;
		for (int r = 0; r < W; r++) sum += *(volatile unsigned long long *)(mine + ((size_t)par * W + r) * PX_WORDS + w);
		x.out[w] = sum;
	}
	if (!x.do_final) return;
	__threadfence();
	__syncthreads();
	if (tid != 0) return;
	double tot[SC_MAXV];
	post_totals(reinterpret_cast<const long long *>(x.out), a.geo.K, tot);
	post_finish(a, tot, a.iter);
}
cudaError_t launch_peer_allreduce(const PeerArgs &x, const PostArgs &a, cudaStream_t s)
{
	peer_allreduce_kernel<<<1, PX_THREADS, 0, s>>>(x, a);
	return cudaGetLastError();
}

// cooperative grid for n_items individuals on `device` (computed once per context)
int scalar_grid(int which, int n_items, int device)
{
	int per_sm = 1, sms = 148;
	const void *k = which == 0 ? (const void *)pre_sweep_kernel<true> : (const void *)post_sweep_kernel<true>;
	cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, SC_THREADS, 0);
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
	int want = (n_items + SC_THREADS - 1) / SC_THREADS;
	int cap = per_sm * sms;
	if (cap > SC_MAX_CTAS) cap = SC_MAX_CTAS;
	if (want < 1) want = 1;
	return want < cap ? want : cap;
}

cudaError_t launch_post_sweep(const PostArgs &a, cudaStream_t s)
{
	if (a.geo.N <= SC1_THREADS) { post_sweep_kernel<false><<<1, SC1_THREADS, 0, s>>>(a); return cudaGetLastError(); }
	void *args[] = {(void *)&a};
	return cudaLaunchCooperativeKernel((const void *)post_sweep_kernel<true>, dim3(a.grid), dim3(SC_THREADS), args, 0, s);
}

// --------------------------------------------------------------------------------------
// moments: store_chn (mcmc.c:1320-1456).  mean_{n+1} = (n*mean_n + x)/(n+1), which is what
// the reference's m*((step + x/m)/(1+step)) evaluates algebraically.
// --------------------------------------------------------------------------------------
__device__ __forceinline__ void run_mean(double *m, double x, long step) { *m = (*m * (double)step + x) / (double)(step + 1); }

__global__ void moments_kernel(const MomArgs a)
{
	const Geometry &g = a.geo;
	const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
	const int K = g.K, REC = g.REC;
	const long step = a.step;
	if (t < (long)a.n_ind * K) {
		const int i = a.i_lo + (int)(t / K), k = (int)(t % K);
		const double q = a.ind[(size_t)i * REC + k];
		const size_t e = (size_t)i * K + k;
		run_mean(a.m.qq + e, q, step);
		run_mean(a.m.qq2 + e, q * q, step);
		if (k == 0) {
			const double lk = a.ind[(size_t)i * REC + K], gen = a.ind[(size_t)i * REC + K + 2];
			run_mean(a.m.indvlkh + i, lk, step);
			run_mean(a.m.gen + i, gen, step);
			run_mean(a.m.gen2 + i, gen * gen, step);
		}
	}
	if (t < a.ns) {
		const double s = a.S[t];
		run_mean(a.m.self + t, s, step);
		run_mean(a.m.self2 + t, s * s, step);
		if (a.convg_slot >= 0 && a.m.convg_S) a.m.convg_S[(size_t)a.convg_slot * a.ns + t] = s;
	}
	if (t == 0) {
		const double tl = a.sc->totallkh;
		run_mean(a.m.tot, tl, step);
		run_mean(a.m.tot + 1, tl * tl, step);
		if (a.convg_slot >= 0) a.m.convg[a.convg_slot] = tl;
	}
	if (a.print_freq && a.m.freq) {
		const long tot = (long)K * g.L * g.A;
		for (long e = t; e < tot; e += (long)gridDim.x * blockDim.x) {
			const int al = (int)(e % g.A), l = (int)((e / g.A) % g.L), k = (int)(e / ((long)g.A * g.L));
			const double p = a.P64 ? a.P64[e] : (double)a.P[((size_t)l * g.A + al) * g.KP + k];
			run_mean(a.m.freq + e, p, step);
			run_mean(a.m.freq2 + e, p * p, step);
		}
	}
}
cudaError_t launch_moments(const MomArgs &a, cudaStream_t s)
{
	long n = (long)a.n_ind * a.geo.K;
	if (a.ns > n) n = a.ns;
	if (n < 1) n = 1;
	moments_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(a);
	return cudaGetLastError();
}

// --------------------------------------------------------------------------------------
// Layout transforms: canonical [L][Nloc][2] <-> tiled [LT][Nloc][8][2]; loci beyond L and
// monomorphic loci are stored as missing so the sweep skips them (mcmc.c:817,1137).
// --------------------------------------------------------------------------------------
__global__ void tile_x_kernel(const int16_t *xc, int16_t *Xt, const int32_t *allelenum, Geometry g)
{
	const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	const size_t total = (size_t)g.LT * g.Nloc * TILE;
	if (t >= total) return;
	const int j = (int)(t % TILE);
	const int i = (int)((t / TILE) % g.Nloc);
	const int mt = (int)(t / ((size_t)TILE * g.Nloc));
	const int l = mt * TILE + j;
	int16_t v0 = -9, v1 = -9;
	if (l < g.L && allelenum[l] > 1) {
		v0 = xc[((size_t)l * g.Nloc + i) * 2];
		v1 = xc[((size_t)l * g.Nloc + i) * 2 + 1];
		if (v0 < 0 || v1 < 0 || v0 >= g.A || v1 >= g.A) { v0 = -9; v1 = -9; }   // any copy missing drops the genotype (data_interface.c:828-832)
	}
	Xt[t * 2] = v0;
	Xt[t * 2 + 1] = v1;
}
__global__ void untile_x_kernel(const int16_t *Xt, int16_t *xc, Geometry g)
{
	const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	const size_t total = (size_t)g.L * g.Nloc;
	if (t >= total) return;
	const int i = (int)(t % g.Nloc), l = (int)(t / g.Nloc);
	const size_t src = (((size_t)(l / TILE) * g.Nloc + i) * TILE + (l % TILE)) * 2;
	xc[t * 2] = Xt[src];
	xc[t * 2 + 1] = Xt[src + 1];
}
__global__ void tile_z_kernel(const int8_t *zc, int8_t *Zt, Geometry g)
{
	const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	const size_t total = (size_t)g.LT * g.Nloc * TILE;
	if (t >= total) return;
	const int j = (int)(t % TILE);
	const int i = (int)((t / TILE) % g.Nloc);
	const int mt = (int)(t / ((size_t)TILE * g.Nloc));
	const int l = mt * TILE + j;
	int8_t v0 = 0, v1 = 0;
	if (l < g.L) { v0 = zc[((size_t)l * g.Nloc + i) * 2]; v1 = zc[((size_t)l * g.Nloc + i) * 2 + 1]; }
	Zt[t * 2] = v0;
	Zt[t * 2 + 1] = v1;
}
__global__ void untile_z_kernel(const int8_t *Zt, int8_t *zc, Geometry g)
{
	const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	const size_t total = (size_t)g.L * g.Nloc;
	if (t >= total) return;
	const int i = (int)(t % g.Nloc), l = (int)(t / g.Nloc);
	const size_t src = (((size_t)(l / TILE) * g.Nloc + i) * TILE + (l % TILE)) * 2;
	zc[t * 2] = Zt[src];
	zc[t * 2 + 1] = Zt[src + 1];
}
static inline unsigned nblocks(size_t n, int b) { return (unsigned)((n + b - 1) / b); }

// per individual: usable heterozygous genotypes (nhet, data only) and those of them whose two
// copies share an ancestry (nsh, a function of Z).  Run at load and whenever Z is injected;
// during a chain zq_sweep / indiv_epilogue keep nsh current.
__global__ void het_counts_kernel(const int16_t *Xt, const int8_t *Zt, int32_t *nhet, int32_t *nsh, Geometry g)
{
	const int il = blockIdx.x * blockDim.x + threadIdx.x;
	if (il >= g.Nloc) return;
	int h = 0, s = 0;
	for (int mt = 0; mt < g.LT; mt++) {
		const size_t base = ((size_t)mt * g.Nloc + il) * TILE * 2;
		for (int j = 0; j < TILE; j++) {
			const int x0 = Xt[base + 2 * j], x1 = Xt[base + 2 * j + 1];
			if (x0 < 0 || x0 == x1) continue;
			h++;
			if (Zt && Zt[base + 2 * j] == Zt[base + 2 * j + 1]) s++;
		}
	}
	if (nhet) nhet[il] = h;
	if (nsh) nsh[il] = s;
}
cudaError_t launch_het_counts(const int16_t *Xt, const int8_t *Zt, int32_t *nhet, int32_t *nsh, Geometry g, cudaStream_t s)
{
	het_counts_kernel<<<nblocks((size_t)g.Nloc, 128), 128, 0, s>>>(Xt, Zt, nhet, nsh, g);
	return cudaGetLastError();
}
cudaError_t launch_tile_x(const int16_t *xc, int16_t *Xt, const int32_t *an, Geometry g, cudaStream_t s)
{
	tile_x_kernel<<<nblocks((size_t)g.LT * g.Nloc * TILE, 256), 256, 0, s>>>(xc, Xt, an, g);
	return cudaGetLastError();
}
cudaError_t launch_untile_x(const int16_t *Xt, int16_t *xc, Geometry g, cudaStream_t s)
{
	untile_x_kernel<<<nblocks((size_t)g.L * g.Nloc, 256), 256, 0, s>>>(Xt, xc, g);
	return cudaGetLastError();
}
cudaError_t launch_tile_z(const int8_t *zc, int8_t *Zt, Geometry g, cudaStream_t s)
{
	tile_z_kernel<<<nblocks((size_t)g.LT * g.Nloc * TILE, 256), 256, 0, s>>>(zc, Zt, g);
	return cudaGetLastError();
}
cudaError_t launch_untile_z(const int8_t *Zt, int8_t *zc, Geometry g, cudaStream_t s)
{
	untile_z_kernel<<<nblocks((size_t)g.L * g.Nloc, 256), 256, 0, s>>>(Zt, zc, g);
	return cudaGetLastError();
}

// --------------------------------------------------------------------------------------
// Stand-alone evaluators used by the parity hooks (ig_set_state(Z), ig_loglik).  They are
// deliberately simple; the product path is zq_sweep.
// --------------------------------------------------------------------------------------
__global__ void tally_kernel(const int16_t *Xt, const int8_t *Zt, int32_t *n, Geometry g)
{
	const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	const size_t total = (size_t)g.LT * g.Nloc * TILE;
	if (t >= total) return;
	const int j = (int)(t % TILE);
	const int mt = (int)(t / ((size_t)TILE * g.Nloc));
	const int l = mt * TILE + j;
	const int x0 = Xt[t * 2], x1 = Xt[t * 2 + 1];
	if (x0 < 0 || x1 < 0) return;
	atomicAdd(&n[((size_t)l * g.A + x0) * g.KP + Zt[t * 2]], 1);
	atomicAdd(&n[((size_t)l * g.A + x1) * g.KP + Zt[t * 2 + 1]], 1);
}
cudaError_t launch_tally(const int16_t *Xt, const int8_t *Zt, int32_t *n, Geometry g, cudaStream_t s)
{
	cudaError_t e = cudaMemsetAsync(n, 0, (size_t)g.Lpad * g.A * g.KP * sizeof(int32_t), s);
	if (e != cudaSuccess) return e;
	tally_kernel<<<nblocks((size_t)g.LT * g.Nloc * TILE, 256), 256, 0, s>>>(Xt, Zt, n, g);
	return cudaGetLastError();
}

// log_ld_indv (mcmc.c:1726-1773) in double, one thread per individual
__global__ void loglik_kernel(const int16_t *Xt, const int8_t *Zt, const float *P, const float *Qf, const int32_t *gen,
                              double *out, Geometry g, int type_freq)
{
	const int il = blockIdx.x * blockDim.x + threadIdx.x;
	if (il >= g.Nloc) return;
	const int gi = gen[il];
	const double h = 1.0 - exp2(-(double)(gi - 1));
	double ll = 0.0;
	for (int l = 0; l < g.Lpad; l++) {
		const size_t src = (((size_t)(l / TILE) * g.Nloc + il) * TILE + (l % TILE)) * 2;
		const int x0 = Xt[src], x1 = Xt[src + 1];
		if (x0 < 0 || x1 < 0) continue;
		const int z0 = Zt[src], z1 = Zt[src + 1];
		double f0, f1;
		bool same;
		if (type_freq == 0) {
			f0 = 0.0; f1 = 0.0;
			for (int k = 0; k < g.K; k++) {
				f0 += (double)Qf[(size_t)il * g.KP + k] * (double)P[((size_t)l * g.A + x0) * g.KP + k];
				f1 += (double)Qf[(size_t)il * g.KP + k] * (double)P[((size_t)l * g.A + x1) * g.KP + k];
			}
			same = true;
		} else {
			f0 = (double)P[((size_t)l * g.A + x0) * g.KP + z0];
			f1 = (double)P[((size_t)l * g.A + x1) * g.KP + z1];
			same = (z0 == z1);
		}
		if (same) {
			if (x0 == x1) ll += log(f0 * (h + f0 * (1.0 - h)));
			else ll += log(2.0 * f0 * f1) - (double)(gi - 1) * LN2_D;
		} else {
			ll += log(f0) + log(f1);
			if (x0 != x1) ll += LN2_D;
		}
	}
	out[il] = ll;
}
cudaError_t launch_loglik(const int16_t *Xt, const int8_t *Zt, const float *P, const float *Qf, const int32_t *gen,
                          double *out, Geometry g, int type_freq, cudaStream_t s)
{
	loglik_kernel<<<(g.Nloc + 63) / 64, 64, 0, s>>>(Xt, Zt, P, Qf, gen, out, g, type_freq);
	return cudaGetLastError();
}

__global__ void fill_f32_kernel(float *p, float v, size_t n)
{
	const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (t < n) p[t] = v;
}
// Q rows that make the categorical draw uniform over the K real populations (padded ones get 0)
__global__ void fill_q_uniform_kernel(float *Qf, Geometry g)
{
	const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (t < (size_t)g.Nloc * g.KP) Qf[t] = ((int)(t % g.KP) < g.K) ? 1.0f / (float)g.K : 0.0f;
}
cudaError_t launch_fill_q_uniform(float *Qf, Geometry g, cudaStream_t s)
{
	fill_q_uniform_kernel<<<nblocks((size_t)g.Nloc * g.KP, 256), 256, 0, s>>>(Qf, g);
	return cudaGetLastError();
}
cudaError_t launch_fill_f32(float *p, float v, size_t n, cudaStream_t s)
{
	fill_f32_kernel<<<nblocks(n, 256), 256, 0, s>>>(p, v, n);
	return cudaGetLastError();
}

__global__ void qf_from_ind_kernel(const double *ind, float *Qf, Geometry g)
{
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= g.Nloc * g.KP) return;
	const int il = t / g.KP, k = t % g.KP;
	Qf[t] = (k < g.K) ? (float)ind[(size_t)(g.i0 + il) * g.REC + k] : 0.0f;
}
cudaError_t launch_qf_from_ind(const double *ind, float *Qf, Geometry g, cudaStream_t s)
{
	qf_from_ind_kernel<<<nblocks((size_t)g.Nloc * g.KP, 256), 256, 0, s>>>(ind, Qf, g);
	return cudaGetLastError();
}

// the host-side Dirichlet-process step reads nothing but G (1..50)
__global__ void pack_g_kernel(const double *ind, uint8_t *g8, Geometry g)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= g.N) return;
	const double v = ind[(size_t)i * g.REC + g.K + 2];
	g8[i] = (uint8_t)(v < 1.0 ? 1 : (v > 255.0 ? 255 : (int)v));
}
cudaError_t launch_pack_g(const double *ind, uint8_t *g8, Geometry g, cudaStream_t s)
{
	pack_g_kernel<<<(g.N + 255) / 256, 256, 0, s>>>(ind, g8, g);
	return cudaGetLastError();
}

// Chain initialisation (mcmc.c:193-205 mode 2, :315-331 mode 3, initial_chn :479):
// alpha ~ U(0,10); mode 2: G_i = min(rgeom(U), 50), S_k = initd[k]; mode 3 uniform prior:
// S_i ~ U(0,1), G_i = rgeom(1 - S_i) capped at 50 (the reference forgets the cap, App. B #5;
// harmless after burn-in).  With the DP prior S comes from the host (init_DP) beforehand.
__global__ void init_chain_kernel(double *ind, double *S, int32_t *state, DevScalars *sc, const float *initd,
                                  int32_t *gprop, int2 *gpair, Geometry g, int mode, int prior_flag, int back_refl,
                                  uint32_t key0, uint32_t key1)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i == 0) {
		Stream st(0u, 1u, 0u, TAG_INIT, key0, key1);
		sc->alpha = st.uniform() * 10.0;
		sc->totallkh = 0.0; sc->sumlogq = 0.0; sc->alpha_accepts = 0; sc->s_accepts = 0; sc->flags = 0;
		if (mode == 2 || mode == 4)                          // mcmc.c:200-205, :256-260
			for (int k = 0; k < g.K; k++) { S[k] = (double)initd[k]; if (back_refl == 0) state[k] = sel_state(S[k]); }
	}
	if (i >= g.N) return;
	Stream st((uint32_t)i, 0u, 0u, TAG_INIT, key0, key1);
	double *rec = ind + (size_t)i * g.REC;
	int gen;
	double fi = 0.0;
	if (mode == 0 || mode == 1 || mode == 4) gen = 1;            // admixture without selfing: G == 1 makes log_ld_indv the mode-1 likelihood (mcmc.c:1869)
	else if (mode == 5) { gen = 1; fi = st.uniform(); S[i] = fi; }    // mcmc.c:416-419
	else if (mode == 2) {
		const double p = st.uniform(), u = st.uniform();
		const double v = floor(log(u) / log(1.0 - p)) + 1.0;
		gen = (v > 50.0) ? 50 : (v < 1.0 ? 1 : (int)v);
	} else {
		if (prior_flag == 0) S[i] = st.uniform();
		const double s = S[i], u = st.uniform();
		const double v = (s > 0.0 && s < 1.0) ? floor(log(u) / log(s)) + 1.0 : (s <= 0.0 ? 1.0 : 50.0);
		gen = (v > 50.0) ? 50 : (v < 1.0 ? 1 : (int)v);
	}
	for (int k = 0; k < g.K; k++) rec[k] = 1.0 / g.K;
	rec[g.K] = 0.0; rec[g.K + 1] = 0.0; rec[g.K + 2] = (mode == 5) ? fi : (double)gen;
	for (int j = g.K + 3; j < g.REC; j++) rec[j] = 0.0;
	gprop[i] = gen;
	const int il = i - g.i0;
	if (il >= 0 && il < g.Nloc) gpair[il] = make_int2(gen, gen);
}
cudaError_t launch_init_chain(double *ind, double *S, int32_t *state, DevScalars *sc, const float *initd_dev,
                              int32_t *gprop, int2 *gpair, Geometry g, int mode, int prior_flag, int back_refl,
                              uint32_t key0, uint32_t key1, cudaStream_t s)
{
	init_chain_kernel<<<(g.N + 127) / 128 + 1, 128, 0, s>>>(ind, S, state, sc, initd_dev, gprop, gpair, g, mode, prior_flag, back_refl, key0, key1);
	return cudaGetLastError();
}

// proposal() (mcmc.c:1630-1648) at an arbitrary S, for the parity hook
__global__ void __launch_bounds__(RED_THREADS) proposal_ll_kernel(const double *ind, const double *S, double *out, Geometry g)
{
	__shared__ double sh[RED_THREADS];
	double part = 0.0;
	for (int i = threadIdx.x; i < g.N; i += RED_THREADS) {
		const double *rec = ind + (size_t)i * g.REC;
		double s = 0.0;
		for (int k = 0; k < g.K; k++) s += rec[k] * S[k];
		part += log_geom(s, (int)rec[g.K + 2]);
	}
	const double r = block_sum(part, sh);
	if (threadIdx.x == 0) *out = r;
}
cudaError_t launch_proposal_ll(const double *ind, const double *S, double *out, Geometry g, cudaStream_t s)
{
	proposal_ll_kernel<<<1, RED_THREADS, 0, s>>>(ind, S, out, g);
	return cudaGetLastError();
}

__global__ void moments_reset_kernel(const MomArgs a)
{
	// initialize_chn (mcmc.c:644-738) seeds every moment with 1 and step = 0, which the first
	// store overwrites with x; zeroing is equivalent for the (step*m + x)/(step+1) form.
	const Geometry &g = a.geo;
	const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
	if (t < (long)g.N * g.K) { a.m.qq[t] = 0; a.m.qq2[t] = 0; }
	if (t < g.N) { a.m.indvlkh[t] = 0; a.m.gen[t] = 0; a.m.gen2[t] = 0; }
	if (t < a.ns) { a.m.self[t] = 0; a.m.self2[t] = 0; }
	if (t < 2) a.m.tot[t] = 0;
	if (a.m.freq) {
		const long tot = (long)g.K * g.L * g.A;
		for (long e = t; e < tot; e += (long)gridDim.x * blockDim.x) { a.m.freq[e] = 0; a.m.freq2[e] = 0; }
	}
}
cudaError_t launch_moments_reset(const MomArgs &a, cudaStream_t s)
{
	long n = (long)a.geo.N * a.geo.K;
	if (a.ns > n) n = a.ns;
	if (n < 2) n = 2;
	moments_reset_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(a);
	return cudaGetLastError();
}

}  // namespace ig
