// sm_100a kernels of the wave path (see cut_types.h "Wave path" and wave_bodies.h).
//
// One iteration = one wave of commuting cuts:
//   k_wave_begin      (one CTA)       commit of the previous wave (counters, return codes, pending list, progress
//                                     record); decide whether a look-ahead pass runs, give free slots to the next halfspaces
//   k_wave_classify   (whole grid)    look-ahead K1: every live row against the slots of this pass, HBM-bound,
//                                     8*d bytes per row for up to 32 halfspaces instead of for one
//   k_wave_mark       (whole grid)    footprint marks of the candidate cuts (scattered atomics: LSU work for all SMs)
//   k_wave_check      (whole grid)    conflict test; the block that finishes last selects the slots of this wave
//   k_wave_tailA      (cluster / cut) visited list, half-edges, sizes           (phases P0-P4 of k_tail)
//   k_wave_tailB      (cluster / cut) bases, new rows, rewiring, retirement, dead facets, look-ahead lists of the
//                                     still pending halfspaces amended with the new rows, K4 bit matrices (P5-P7)
//   k_wave_k4_filter / k_wave_k4_contain (whole grid) pair test of all cuts of the wave
//   k_wave_tail2      (cluster / cut) adjacency build
// Every kernel returns at once when the scheduler has halted (WaveCtl::halt): the host enqueues iterations ahead
// without reading anything back, and only looks at a progress record in mapped host memory.
#pragma once
#include "cut_kernels.cuh"
#include "wave_bodies.h"

#define WV_TRACE(k) do { if (W.trace && blockIdx.x == 0 && threadIdx.x == 0) W.trace[((W.wc->iter & 255u) << 6) + (k)] = b200_globaltimer(); } while (0)
// phase stamps inside a kernel (block 0 = CTA 0 of the first cluster)
#define WV_TP(k) do { if (W.trace && blockIdx.x == 0 && threadIdx.x == 0) W.trace[((w_iter & 255u) << 6) + (k)] = b200_globaltimer(); } while (0)
static_assert(B200_WV_GROUP == 2 * K_THREADS, "look-ahead row group");
#define WAVE_NC 8              // CTAs per cluster of the per-cut kernels: 16 clusters of 8 are co-resident on 148 SMs

// all threads of the block copy `bytes` (a multiple of 4) -- the caller synchronises
__device__ __forceinline__ void wv_copy_words(void *dst, const void *src, u32 bytes)
{
	for (u32 i = threadIdx.x; i < bytes / 4; i += blockDim.x) ((u32 *)dst)[i] = ((const u32 *)src)[i];
}

// wave_commit and wave_la_plan (wave_bodies.h) by the lanes of one warp: a single thread walking the control block in
// shared memory takes 6 + 4 us for them (every access a dependent 30-cycle round trip), the warp 1 us.  Same
// arithmetic, term by term; lane q owns wave position q in the commit and slot q in the plan.
static_assert(B200_WAVE_SLOTS == 32 && B200_WAVE_MAXW <= 32, "one lane per slot / wave position");
__device__ __forceinline__ u64 wv_sum64(u64 v)
{
#pragma unroll
	for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
	return v;
}
__device__ __forceinline__ void wave_commit_warp(WaveCtl &w, CutCtl &m, const WaveCut *cut, int dim, int *rc)
{
	const u32 lane = threadIdx.x & 31, n_commit = w.n_commit, n_wave = w.n_wave, n_pending = w.n_pending;
	const u64 d = (u64)dim;
	u64 evals = 0, bytes = 0, pt = 0;
	u32 cuts = 0, red = 0, mn = 0, zr = 0, ed = 0, pr = 0, rows = 0, inc = 0, adj = 0, gone = 0, done_slots = 0;
	if (lane < n_commit) {
		const WaveCut c = cut[lane];
		const u64 N = c.live_before;
		evals = N;
		if (c.status & ST_REDUNDANT) {
			red = 1;
			bytes = N * (8 * d + 1);
			rc[lane] = 1;
		} else {
			const u64 nm = c.n_minus, nz = c.n_zero, M = c.n_new, E = M - nz, Wd = ((u64)c.facet + 64) / 64, A = c.n_pairs;
			cuts = 1; mn = c.n_minus; zr = c.n_zero; ed = (u32)E; pr = c.n_pairs;
			pt = M * (M - (M ? 1 : 0)) / 2;
			bytes = N * (8 * d + 1) + N + 4 * (nm + nz) + E * (24 * d + 24 * Wd) + nz * (16 * d + 16 * Wd) + 8 * M * Wd + 8 * A;
			rows = c.n_new; inc = c.inc_new; adj = c.adj_new; gone = c.n_minus + c.n_zero;
			rc[lane] = 0;
		}
		done_slots = 1u << w.wave[lane];
	}
	evals = wv_sum64(evals); bytes = wv_sum64(bytes); pt = wv_sum64(pt);
	cuts = __reduce_add_sync(0xffffffffu, cuts); red = __reduce_add_sync(0xffffffffu, red);
	mn = __reduce_add_sync(0xffffffffu, mn); zr = __reduce_add_sync(0xffffffffu, zr);
	ed = __reduce_add_sync(0xffffffffu, ed); pr = __reduce_add_sync(0xffffffffu, pr);
	rows = __reduce_add_sync(0xffffffffu, rows); inc = __reduce_add_sync(0xffffffffu, inc);
	adj = __reduce_add_sync(0xffffffffu, adj); gone = __reduce_add_sync(0xffffffffu, gone);
	done_slots = __reduce_or_sync(0xffffffffu, done_slots);
	// the committed slots leave the pending list (order kept)
	const u32 slot = lane < n_pending ? w.pending[lane] : 0u;
	const bool keep = lane < n_pending && !((done_slots >> slot) & 1u);
	const u32 kept = __ballot_sync(0xffffffffu, keep);
	if (keep) w.pending[__popc(kept & ((1u << lane) - 1u))] = slot;
	if ((done_slots >> lane) & 1u) w.slot_hs[lane] = B200_NONE;
	if (lane == 0) {
		w.st_evals += evals; w.st_bytes += bytes; w.st_cuts += cuts; w.st_redundant += red; w.st_minus += mn; w.st_zero += zr;
		w.st_edge += ed; w.st_copies += zr; w.st_pair_tests += pt; w.st_pairs += pr;
		w.st_deferred += n_wave - n_commit;
		m.n_live = m.n_live + rows - gone;
		m.nrows += rows;
		m.slot_cnt += rows;
		m.inc_used += inc;
		m.adj_used += adj;
		w.n_pending = __popc(kept);
		w.done_hs += n_commit;
		w.n_wave = w.n_commit = 0;
		if (w.done_hs >= w.n_total) w.halt |= WH_DONE;
		else if (m.nrows != m.n_live && m.nrows >= 4 * B200_TILE && m.nrows >= 2 * m.n_live) w.halt |= WH_COMPACT;
	}
	__syncwarp();
}
__device__ __forceinline__ void wave_la_plan_warp(WaveCtl &w, u32 nrows)
{
	const u32 lane = threadIdx.x & 31;
	u32 n_la = 0, la_new = 0, n_pending = w.n_pending;
	const u32 next_hs = w.next_hs, halt = w.halt, reclassify = w.reclassify;
	__syncwarp();
	if (!halt) {
		if (reclassify) {                    // every pending list is stale: rebuild them in this pass
			if (lane < n_pending) w.la[lane] = w.pending[lane];
			n_la = n_pending;
		}
		if (n_pending < w.refill_below || n_la) {
			const bool is_free = w.slot_hs[lane] == B200_NONE;
			const u32 fm = __ballot_sync(0xffffffffu, is_free);
			const u32 k = min(min((u32)B200_WAVE_SLOTS - n_pending, w.n_total - next_hs), (u32)__popc(fm));
			const u32 r = __popc(fm & ((1u << lane) - 1u));
			if (is_free && r < k) {          // the r-th free slot (ascending) takes the r-th next halfspace
				w.slot_hs[lane] = next_hs + r;
				w.pending[n_pending + r] = lane;
				w.la[n_la + r] = lane;
			}
			la_new = (k >= 32 ? 0xffffffffu : ((1u << k) - 1u)) << n_la;
			n_la += k;
			n_pending += k;
		}
	}
	if (lane == 0) {
		if (!halt) {
			w.reclassify = 0;
			w.next_hs = next_hs + (n_pending - w.n_pending);
			w.n_pending = n_pending;
			if (n_la) { w.st_la_passes++; w.st_rows_scanned += nrows; }
		}
		w.n_la = n_la;
		w.la_new = la_new;
		if (!halt) w.la_rows = nrows;
	}
	__syncwarp();
}

// wave_plan (wave_bodies.h) by the lanes of one warp: lane q owns wave position q, the bases are prefix sums
__device__ __forceinline__ u32 wv_excl_scan(u32 v, u32 lane)
{
	u32 inc = v;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		const u32 t = __shfl_up_sync(0xffffffffu, inc, o);
		if (lane >= (u32)o) inc += t;
	}
	return inc - v;
}
__device__ __forceinline__ void wave_plan_warp(const WaveCtl &w, const WaveCut *cut, u32 nrows, u32 inc_used, u32 n_live, u32 cap_rows, u32 cap_inc, u64 cap_bits, WavePlan &pl)
{
	const u32 lane = threadIdx.x & 31, n_wave = w.n_wave;
	WaveCut c;
	c.status = ST_REDUNDANT; c.n_new = c.inc_new = c.n_minus = c.n_zero = c.facet = 0; c.hs = B200_NONE;
	if (lane < n_wave) c = cut[lane];
	const bool red = (c.status & ST_REDUNDANT) != 0;
	const u32 a_rows = red ? 0u : c.n_new, a_inc = red ? 0u : c.inc_new, a_gone = red ? 0u : c.n_minus + c.n_zero;
	const u32 rows = nrows + wv_excl_scan(a_rows, lane), inc = inc_used + wv_excl_scan(a_inc, lane);
	const u32 live = n_live + wv_excl_scan(a_rows, lane) - wv_excl_scan(a_gone, lane);
	const u32 mpad = (c.n_new + 63) & ~63u, wl_ub = (c.facet + 64) / 64;
	const u64 bits_ub = k4_words(wl_ub, mpad, wl_ub * 64);
	u32 bad = 0;
	if (lane < n_wave && !red) {
		if (c.status & (ST_NEED_BIG | ST_ERR_DEGENERATE | ST_OVF_PADJ)) bad = WH_SERIAL;
		else if ((u64)rows + c.n_new > cap_rows || (u64)rows + c.n_new > B200_WV_ROW_MASK || (u64)inc + c.inc_new > cap_inc || bits_ub > cap_bits) bad = WH_GROW;
	}
	const u32 bad_mask = __ballot_sync(0xffffffffu, bad != 0);
	const u32 first_bad = bad_mask ? (u32)__ffs(bad_mask) - 1u : n_wave;
	if (lane < n_wave) { pl.rows_base[lane] = rows; pl.inc_base[lane] = inc; pl.live_before[lane] = live; }
	if (lane == 0) {
		pl.n_commit = first_bad;
		pl.halt = 0;
		pl.halt_hs = B200_NONE;
		pl.need_rows = pl.need_inc = 0;
		pl.need_bits = 0;
	}
	__syncwarp();
	if (lane == 0 && first_bad == 0 && bad_mask) {       // not even the first cut of the wave fits: the host has to act
		pl.halt = bad;
		pl.halt_hs = c.hs;
		pl.need_rows = rows + c.n_new;
		pl.need_inc = inc + c.inc_new;
		pl.need_bits = bits_ub;
	}
	__syncwarp();
}

__global__ void __launch_bounds__(64) k_wave_begin(DevState S, WaveDev W, const double *vals, const unsigned char *ideal)
{
	__shared__ WaveCtl w;
	__shared__ WaveCut cut[B200_WAVE_MAXW];
	__shared__ CutCtl mctl;
	__shared__ int s_rc[B200_WAVE_MAXW];
	__shared__ u32 s_hs[B200_WAVE_MAXW];
	cudaGridDependencySynchronize();
	const u64 t_begin = b200_globaltimer();
	wv_copy_words(&w, W.wc, sizeof w);
	wv_copy_words(&mctl, S.ctl, sizeof mctl);
	__syncthreads();
	if (w.halt) return;
	u64 *tr = (W.trace && threadIdx.x == 0) ? W.trace + (((w.iter + 1) & 255u) << 6) : nullptr;
	if (tr) tr[40] = b200_globaltimer();
	// ---- the previous wave's adjacency build has completed: commit it
	const u32 n_commit = w.n_commit;
	if (n_commit) {
		if (threadIdx.x < n_commit) { wave_gather_cut(W, w, threadIdx.x, cut[threadIdx.x]); s_hs[threadIdx.x] = cut[threadIdx.x].hs; }
		__syncthreads();
		if (tr) tr[41] = b200_globaltimer();
	}
	if (threadIdx.x < 32) {
		if (n_commit) {
			wave_commit_warp(w, mctl, cut, S.d, s_rc);
			if (threadIdx.x == 0) {
				const u64 now = b200_globaltimer();
				if (!w.t_first) w.t_first = now;
				w.t_last = now;
			}
		}
		if (tr) tr[42] = b200_globaltimer();
		if (threadIdx.x == 0) {
			w.iter++;
			if (W.trace) W.trace[(w.iter & 255u) << 6] = t_begin;
		}
		__syncwarp();
		wave_la_plan_warp(w, mctl.nrows);
		if (threadIdx.x == 0) {
			wave_shard_plan(w, W, mctl.nrows);
			if (w.shard) { W.xsend[0] = 0; W.xsend[1] = w.xseq; }
			if (w.shard_k4 && !w.halt) { wave_k4_record_reset(W, ++w.xseq_k); w.st_sharded_k4++; }    // one pair-test exchange per iteration, empty or not
		}
	}
	__syncthreads();
	if (tr) tr[43] = b200_globaltimer();
	if (n_commit) {
		if (threadIdx.x < n_commit) W.rc[s_hs[threadIdx.x]] = s_rc[threadIdx.x];
		wv_copy_words(S.ctl, &mctl, sizeof mctl);
	}
	if (!w.halt && threadIdx.x < w.n_la) wave_la_init(S, W, w, threadIdx.x, vals, ideal);
	if (threadIdx.x < B200_WAVE_SLOTS) W.wflag[threadIdx.x] = 0;       // wave formation flags of this iteration
	wv_copy_words(W.wc, &w, sizeof w);
	if (tr) tr[44] = b200_globaltimer();
	if (threadIdx.x == 0) wave_publish(W, w, mctl.nrows, mctl.n_live);
	if (tr) tr[45] = b200_globaltimer();
}

// Look-ahead K1.  Thread t of a block owns rows 2t, 2t+1 of a 512-row group (one double2 load per coordinate, a
// contiguous 512-byte run per warp); the coordinates stay in registers while the halfspaces of the pass, staged in
// shared memory, are evaluated against them one after the other.
//
// All but a few hundred of the 10^6 rows are far on the PLUS side of every halfspace, and saying so does not need the
// reference's arithmetic: an FP32 dot product (6 FFMA on the FP32 pipe, twice as wide as the FP64 pipe and otherwise
// idle here) with a rigorous error bound decides them.  |t32 - t| <= (d + 3) 2^-24 * sum|h_j||x_j| <= 7e-7 * ||h||_1 ||x||_inf
// for d <= 8 (inputs rounded to float, products and sums fused); a row is skipped when
//     t32 > hi + 2.5e-6 * (|thr| + ||h||_1 ||x||_inf)          (right-hand side rounded up in every step),
// which implies t > hi + 1e-7 * (|thr| + ||h||_1 ||x||_inf), the guard-band test of wave_code().  Every other row is
// evaluated exactly as before -- strict left-to-right sums of separately rounded FP64 products, as bslv_poly.c:123-125
// computes them -- so the lists are the same sets with or without the filter.  The pass was FP64-pipe-bound
// (30 FP64 instructions per row and halfspace); it is now bound by reading the coordinates.
template <int D>
__global__ void __launch_bounds__(K_THREADS) k_wave_classify(DevState S, WaveDev W)
{
	constexpr int DD = D > 0 ? D : B200_MAXD;
	__shared__ double sh[B200_WAVE_SLOTS][DD];
	__shared__ double s_alpha[B200_WAVE_SLOTS], s_h1[B200_WAVE_SLOTS], s_thr[B200_WAVE_SLOTS][6];
	__shared__ float sh32[B200_WAVE_SLOTS][DD], s_a32[B200_WAVE_SLOTS][2], s_c32[B200_WAVE_SLOTS];
	__shared__ u32 s_slot[B200_WAVE_SLOTS];
	cudaGridDependencySynchronize();
	WV_TRACE(1);
	const WaveCtl *w = W.wc;
	const u32 n_la = w->n_la;
	if (w->halt || n_la == 0) return;
	const int d = D > 0 ? D : S.d;
	for (u32 x = threadIdx.x; x < n_la; x += K_THREADS) {
		const u32 slot = w->la[x];
		const CutParams &P = W.cur[slot];
		s_slot[x] = slot;
		for (int j = 0; j < d; j++) { sh[x][j] = P.h[j]; sh32[x][j] = __double2float_rn(P.h[j]); }
		s_alpha[x] = P.alpha;
		s_h1[x] = P.h1;
		s_thr[x][0] = P.hi[0]; s_thr[x][1] = P.hi[1]; s_thr[x][2] = P.mid[0]; s_thr[x][3] = P.mid[1]; s_thr[x][4] = P.lo[0]; s_thr[x][5] = P.lo[1];
		// FP32 filter: skip when t32 > a32[id] + c32 * xinf32, every constant rounded up
		const float slack = 2.5e-6f;
		s_c32[x] = __fmul_ru(slack, __double2float_ru(P.h1));
		s_a32[x][0] = __fadd_ru(__double2float_ru(P.hi[0]), __fmul_ru(slack, __double2float_ru(fabs(P.alpha))));
		s_a32[x][1] = __double2float_ru(P.hi[1]);
	}
	__syncthreads();
	const size_t cap = S.cap_rows;
	const u32 nrows = w->la_rows, ngroups = (nrows + 2 * K_THREADS - 1) / (2 * K_THREADS);
	const bool shard = w->shard != 0;        // several GPUs: this rank's row groups only, entries into the exchange record
	u32 g_lo = 0, g_hi = ngroups;
	if (shard) wave_shard_range(W, ngroups, g_lo, g_hi);
	for (u32 g = g_lo + blockIdx.x; g < g_hi; g += gridDim.x) {
		const u32 r = g * 2 * K_THREADS + 2 * threadIdx.x;
		const u32 lw = S.live[r >> 5] >> (r & 31), iw = S.ideal[r >> 5] >> (r & 31);
		double2 x[DD];
#pragma unroll
		for (int j = 0; j < DD; j++)
			if (j < d) x[j] = *reinterpret_cast<const double2 *>(S.coord + j * cap + r);
		if (!(lw & 3u)) continue;
		double xi0 = 0, xi1 = 0;
		float x32a[DD], x32b[DD];
#pragma unroll
		for (int j = 0; j < DD; j++)
			if (j < d) {
				xi0 = fmax(xi0, fabs(x[j].x)); xi1 = fmax(xi1, fabs(x[j].y));
				x32a[j] = __double2float_rn(x[j].x); x32b[j] = __double2float_rn(x[j].y);
			}
		const float xf0 = __double2float_ru(xi0), xf1 = __double2float_ru(xi1);
		const u32 id0 = iw & 1u, id1 = (iw >> 1) & 1u;
		for (u32 k = 0; k < n_la; k++) {
			float f0 = sh32[k][0] * x32a[0], f1 = sh32[k][0] * x32b[0];
#pragma unroll
			for (int j = 1; j < DD; j++)
				if (j < d) { f0 = fmaf(sh32[k][j], x32a[j], f0); f1 = fmaf(sh32[k][j], x32b[j], f1); }
			// (a dead row of the pair counts as decided; NaN or inf compares false and takes the exact path)
			const bool far0 = !(lw & 1u) || f0 > __fmaf_ru(s_c32[k], xf0, s_a32[k][id0]);
			const bool far1 = !(lw & 2u) || f1 > __fmaf_ru(s_c32[k], xf1, s_a32[k][id1]);
			if (far0 && far1) continue;
			double t0 = __dmul_rn(sh[k][0], x[0].x), t1 = __dmul_rn(sh[k][0], x[0].y);
#pragma unroll
			for (int j = 1; j < DD; j++)
				if (j < d) {
					t0 = __dadd_rn(t0, __dmul_rn(sh[k][j], x[j].x));
					t1 = __dadd_rn(t1, __dmul_rn(sh[k][j], x[j].y));
				}
			const double t[2] = {t0, t1}, xi[2] = {xi0, xi1};
#pragma unroll
			for (int s = 0; s < 2; s++) {
				if (!((lw >> s) & 1u)) continue;
				const int id = (iw >> s) & 1u;
				const double thr = id ? 0.0 : s_alpha[k];
				const double g2 = B200_WV_GUARD * (fabs(thr) + s_h1[k] * xi[s]);
				if (t[s] > s_thr[k][id] + g2) continue;                                   // safely PLUS
				u32 code = t[s] > s_thr[k][id] ? CLS_PLUS : t[s] > s_thr[k][2 + id] ? CLS_ZP : t[s] > s_thr[k][4 + id] ? CLS_ZERO : CLS_MINUS;
				if (t[s] < s_thr[k][4 + id]) code |= B200_WV_STRICT;
				if (shard) wave_send_append(W, s_slot[k], r + s, code);
				else wave_list_append(W, s_slot[k], r + s, code);
			}
		}
	}
}

// ---------------------------------------------------------------- multi-GPU exchange of a sharded look-ahead pass
__device__ __forceinline__ u32 wv_ld_acquire_sys(const u32 *p)
{
	u32 v;
	asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ void wv_st_release_sys(u32 *p, u32 v)
{
	asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// block p stores this rank's record into rank p's receive area (peer-mapped memory: the stores travel over NVLink)
// and then publishes the pass number in p's flag word for this rank
__global__ void __launch_bounds__(K_THREADS) k_wave_xpush(DevState S, WaveDev W)
{
	cudaGridDependencySynchronize();
	const WaveCtl *w = W.wc;
	if (w->halt || !w->shard) return;
	const u32 p = blockIdx.x;
	if (p >= W.nranks || p == W.rank) return;
	const u32 seq = w->xseq;
	const unsigned long long cnt = W.xsend[0];
	const u32 n = (u32)(cnt < B200_X_CAP ? cnt : B200_X_CAP) + 2;
	unsigned long long *dst = W.xpeer_recv[p] + ((size_t)W.rank * 2 + (seq & 1u)) * B200_X_WORDS;
	for (u32 x = threadIdx.x; x < n; x += K_THREADS) dst[x] = W.xsend[x];
	__threadfence_system();
	__syncthreads();
	if (threadIdx.x == 0) wv_st_release_sys(W.xpeer_flag[p] + W.rank * 32, seq);
}
// blocks (src, part): wait for rank src's record of this pass, append its entries to the lists of their slots
__global__ void __launch_bounds__(K_THREADS) k_wave_xmerge(DevState S, WaveDev W)
{
	__shared__ u32 s_ok;
	cudaGridDependencySynchronize();
	WaveCtl *w = W.wc;
	if (w->halt || !w->shard) return;
	const u32 src = blockIdx.x % W.nranks, part = blockIdx.x / W.nranks, nparts = gridDim.x / W.nranks;
	if (part >= nparts) return;
	const u32 seq = w->xseq;
	const unsigned long long *rec = W.xsend;
	if (src != W.rank) {
		if (threadIdx.x == 0) {
			const u64 t0 = b200_globaltimer();
			u32 ok = 1;
			while ((int)(wv_ld_acquire_sys(W.xflag + src * 32) - seq) < 0) {
				if (b200_globaltimer() - t0 > 4000000000ull) { ok = 0; break; }      // 4 s: the peer is gone
				__nanosleep(100);
			}
			s_ok = ok;
		}
		__syncthreads();
		if (!s_ok) {
			if (threadIdx.x == 0) { const u32 h = atomicOr(&w->halt, (u32)WH_XFAIL) | WH_XFAIL; W.progress->halt = h; }
			return;
		}
		rec = W.xrecv + ((size_t)src * 2 + (seq & 1u)) * B200_X_WORDS;
	}
	const unsigned long long cnt = __ldcg(rec);
	if (cnt > B200_X_CAP) {                 // every rank sees the same count and halts the same way
		if (part == 0 && threadIdx.x == 0) { const u32 h = atomicOr(&w->halt, (u32)WH_XOVER) | WH_XOVER; W.progress->halt = h; }
		return;
	}
	for (u32 x = part * K_THREADS + threadIdx.x; x < (u32)cnt; x += nparts * K_THREADS) wave_merge_entry(W, __ldcg(rec + 2 + x));
}

// ---------------------------------------------------------------- wave formation (two grid-wide kernels)
// candidate p's list occupies [off[p], off[p+1]) of a flat index space every block derives for itself
__device__ __forceinline__ u32 wave_stage_candidates(const WaveDev &W, WaveCtl &w, u32 *off, u32 *nl)
{
	wv_copy_words(&w, W.wc, sizeof w);
	__syncthreads();
	const u32 nc = w.halt ? 0u : wave_candidates(w);
	if (threadIdx.x < nc) nl[threadIdx.x] = min(W.ctl[w.pending[threadIdx.x]].n_list, (u32)B200_WAVE_LIST);
	__syncthreads();
	if (threadIdx.x == 0) {
		u32 t = 0;
		for (u32 p = 0; p < nc; p++) { off[p] = t; t += nl[p]; }
		off[nc] = t;
	}
	__syncthreads();
	return nc;
}
__global__ void __launch_bounds__(K_THREADS) k_wave_mark(DevState S, WaveDev W)
{
	__shared__ WaveCtl w;
	__shared__ u32 off[B200_WAVE_SLOTS + 1], nl[B200_WAVE_SLOTS];
	cudaGridDependencySynchronize();
	WV_TRACE(2);
	const u32 nc = wave_stage_candidates(W, w, off, nl);
	if (!nc) return;
	// (the flags were cleared by k_wave_begin: other blocks of this grid may already be OR-ing bit 1 into them, so a plain
	// store here could lose one)
	if (blockIdx.x == 0 && threadIdx.x < nc && W.ctl[w.pending[threadIdx.x]].n_list > B200_WAVE_LIST) atomicOr(W.wflag + threadIdx.x, 2u);
	const u32 total = off[nc], epoch = w.epoch;
	for (u32 x = blockIdx.x * K_THREADS + threadIdx.x; x < total; x += gridDim.x * K_THREADS) {
		u32 p = 0;
		while (off[p + 1] <= x) p++;
		wave_mark_entry(S, W, w.pending[p], p, epoch, x - off[p]);
	}
}
// a candidate conflicts when a row of its list carries the mark of an earlier candidate; the block that finishes
// last selects the wave
__global__ void __launch_bounds__(K_THREADS) k_wave_check(DevState S, WaveDev W)
{
	__shared__ WaveCtl w;
	__shared__ u32 off[B200_WAVE_SLOTS + 1], nl[B200_WAVE_SLOTS], fl[B200_WAVE_SLOTS], s_last;
	cudaGridDependencySynchronize();
	WV_TRACE(3);
	const u32 nc = wave_stage_candidates(W, w, off, nl);
	if (w.halt) return;
	const u32 total = off[nc], epoch = w.epoch;
	for (u32 x = blockIdx.x * K_THREADS + threadIdx.x; x < total; x += gridDim.x * K_THREADS) {
		u32 p = 0;
		while (off[p + 1] <= x) p++;
		if (p) wave_check_entry(S, W, w.pending[p], p, epoch, x - off[p]);
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		__threadfence();
		s_last = atomicAdd(W.fin_ctr, 1u) == gridDim.x - 1;
	}
	__syncthreads();
	if (!s_last) return;
	__threadfence();
	if (threadIdx.x < nc) fl[threadIdx.x] = __ldcg(W.wflag + threadIdx.x);
	__syncthreads();
	if (threadIdx.x == 0) {
		*W.fin_ctr = 0;
		wave_form_finish(w, fl);
	}
	__syncthreads();
	wv_copy_words(W.wc, &w, sizeof w);
	if (w.halt && threadIdx.x == 0) wave_publish(W, w, S.ctl->nrows, S.ctl->n_live);
}

// ---------------------------------------------------------------- per-cut kernels: one cluster per cut of the wave
// phases P0-P4 of k_tail on the slot's look-ahead list; nothing shared is mutated except the class bytes of the
// visited rows (undone if the cut is deferred)
template <int NC> __global__ void __launch_bounds__(TAIL_THREADS, 1) k_wave_tailA(DevState S0, WaveDev W)
{
	__shared__ u32 ws[99];
	__shared__ u32 slist[B200_VIS_MAX];
	__shared__ u32 soff[B200_VIS_MAX + 1];
	__shared__ u32 s_nvis, s_nstrict;
	const u32 rank = tail_rank<NC>(), ctid = rank * TAIL_THREADS + threadIdx.x, q = blockIdx.x / NC;
	cudaGridDependencySynchronize();
	WV_TRACE(4);
	const WaveCtl *w = W.wc;
	if (w->halt || q >= w->n_wave) return;
	const u32 slot = w->wave[q];
	const DevState S = wave_view(S0, W, slot, q);
	CutCtl *c = S.ctl;
	const u32 w_iter = w->iter;
	WV_TP(16);
	// ---- P0: stage the list; visited rows = live entries classed ZERO or MINUS; trigger = a live strictly violated row
	const u32 n_list = c->n_list;                  // <= B200_WAVE_LIST (longer lists never join a wave)
	if (threadIdx.x == 0) { s_nvis = 0; s_nstrict = 0; }
	__syncthreads();
	{
		u32 nv = 0, ns = 0;
		for (u32 x = threadIdx.x; x < n_list; x += TAIL_THREADS) {
			const u32 ent = S.nplist[x], row = ent & B200_WV_ROW_MASK, code = ent >> B200_WV_ROW_BITS;
			const bool alive = bit_test(S.live, row);
			const bool visited = alive && (code & 3u) >= CLS_ZERO;
			slist[x] = visited ? row : B200_NONE;
			nv += visited;
			ns += alive && (code & B200_WV_STRICT);
		}
		nv = __reduce_add_sync(0xffffffffu, nv);
		ns = __reduce_add_sync(0xffffffffu, ns);
		if ((threadIdx.x & 31) == 0) { if (nv) atomicAdd(&s_nvis, nv); if (ns) atomicAdd(&s_nstrict, ns); }
	}
	__syncthreads();
	const u32 n_vis = s_nvis, n_strict = s_nstrict;
	WV_TP(17);
	if (ctid == 0) {
		c->status = n_strict ? 0u : (u32)ST_REDUNDANT;
		c->n_strict = n_strict;
		c->min_strict_row = c->min_strict_slot = B200_NONE;
		c->n_zp = c->n_zp_projected = 0;
		c->n_vis = n_strict ? n_vis : 0;
		c->n_new = c->inc_new = c->padj_new = 0;
		c->n_minus = c->n_zero = 0;
		c->n_pairs = c->adj_new = c->n_dead_facets = 0;
		c->n_live_scanned = 0;
		c->n_local = c->wl = c->mpad = c->n_surv = 0;
		c->scratch_flag = 0;
		if (!n_strict) S.facet_alive[S.cur->facet] = 0;        // nothing to cut: redundant (bslv_poly.c:132-136)
	}
	if (!n_strict) return;                                      // (every CTA of the cluster sees the same count)
	if (n_vis > n_strict) TAIL_LOOP(x, n_vis * (B200_MAXINC / 64)) S.zmask[x] = 0;
	{
		// rank sort into the ascending visited list; CTA 0 also marks the classes
		const u32 per = (n_list + NC - 1) / NC, e0 = rank * per, e1 = min(n_list, e0 + per);
		const u32 part = threadIdx.x & 7;
		for (u32 eb = e0; eb < e1; eb += TAIL_THREADS / 8) {      // block-uniform trip count
			const u32 i = eb + (threadIdx.x >> 3);
			const u32 key = i < e1 ? slist[i] : B200_NONE;
			u32 rk = 0;
			for (u32 x = part; x < n_list; x += 8) rk += (slist[x] < key);
			rk += __shfl_xor_sync(0xffffffffu, rk, 1);
			rk += __shfl_xor_sync(0xffffffffu, rk, 2);
			rk += __shfl_xor_sync(0xffffffffu, rk, 4);
			if (part == 0 && key != B200_NONE) {
				S.vis[rk] = key;
				S.cls[key] = (u8)((S.nplist[i] >> B200_WV_ROW_BITS) & 3u);
			}
		}
	}
	TAIL_SYNC();
	// ---- P2: half-edge offsets (every CTA, into shared memory)
	WV_TP(18);
	for (u32 x = threadIdx.x; x < n_vis; x += TAIL_THREADS) slist[x] = S.vis[x];
	u32 H = 0;
	for (u32 base = 0; base < n_vis; base += TAIL_THREADS) {
		u32 i = base + threadIdx.x, v = 0, tot;
		if (i < n_vis) {
			v = S.adj_len[slist[i]];
			// an on-plane vertex on more facets than the shared-facet mask holds: this cut runs alone (classic path)
			if (S.inc_len[slist[i]] > B200_MAXINC && S.cls[slist[i]] == CLS_ZERO) s_nstrict = B200_NONE;
		}
		u32 e = block_excl_scan(v, ws, tot);
		if (i < n_vis) {
			soff[i] = H + e;
			if (rank == 0) S.he_off[i] = H + e;
		}
		H += tot;
	}
	if (threadIdx.x == 0) {
		soff[n_vis] = H;
		if (rank == 0) S.he_off[n_vis] = H;
	}
	__syncthreads();
	if (H > S.cap_he || s_nstrict == B200_NONE) {   // too large for the scratch of a wave position: runs alone
		if (ctid == 0) c->status |= ST_NEED_BIG;
		return;
	}
	// ---- P3: evaluate every (visited vertex, neighbour) pair
	WV_TP(19);
	TAIL_SPREAD(e, H) {
		u32 lo = 0, hi = n_vis;
		while (hi - lo > 1) {
			const u32 mid = (lo + hi) >> 1;
			if (soff[mid] <= e) lo = mid;
			else hi = mid;
		}
		S.he_own[e] = lo;
		he_eval_at(S, e, lo, slist[lo], soff[lo]);
	}
	TAIL_SYNC();
	// ---- P4a: sizes per visited vertex
	WV_TP(20);
	{
		u32 nm = 0, nz = 0;
		bool bad = false;
		TAIL_SPREAD(i, n_vis) {
			const u32 v = slist[i];
			const u8 cl = S.cls[v];
			u32 cnt[3];
			bad |= !he_count_core(S, i, v, cl, soff[i], soff[i + 1], cnt);
			S.cnt3[3 * (size_t)i + 0] = cnt[0];
			S.cnt3[3 * (size_t)i + 1] = cnt[1];
			S.cnt3[3 * (size_t)i + 2] = cnt[2];
			nm += (cl == CLS_MINUS);
			nz += (cl == CLS_ZERO);
		}
		nm = __reduce_add_sync(0xffffffffu, nm);
		nz = __reduce_add_sync(0xffffffffu, nz);
		if ((threadIdx.x & 31) == 0) {
			if (nm) atomicAdd(&c->n_minus, nm);
			if (nz) atomicAdd(&c->n_zero, nz);
		}
		if (bad) atomicOr(&c->status, (u32)ST_ERR_DEGENERATE);
	}
	TAIL_SYNC();
	// ---- P4b: offsets (CTA 0 stores them; k_wave_tailB reads them after the launch boundary)
	WV_TP(21);
	if (rank == 0) {
		u32 carry[3] = {0, 0, 0};
		for (u32 base = 0; base < n_vis; base += TAIL_THREADS) {
			const u32 i = base + threadIdx.x;
			u32 v3[3] = {0, 0, 0}, e3[3], tot3[3];
			if (i < n_vis) {
#pragma unroll
				for (int k = 0; k < 3; k++) v3[k] = S.cnt3[3 * (size_t)i + k];
			}
			block_excl_scan3(v3, ws, e3, tot3);
#pragma unroll
			for (int k = 0; k < 3; k++) {
				if (i < n_vis) S.base3[3 * (size_t)i + k] = carry[k] + e3[k];
				carry[k] += tot3[k];
			}
		}
		if (threadIdx.x == 0) {
			c->n_new = carry[0];
			c->inc_new = carry[1];
			c->padj_new = carry[2];
			if (carry[0] > W.cap_new || carry[2] > W.cap_new) atomicOr(&c->status, (u32)ST_OVF_PADJ);
		}
	}
	WV_TP(22);
}

// phases P5-P7 of k_tail for the cuts that are carried out, at the bases wave_plan gives them
template <int NC> __global__ void __launch_bounds__(TAIL_THREADS, 1) k_wave_tailB(DevState S0, WaveDev W)
{
	__shared__ WaveCtl w;
	__shared__ WaveCut cut[B200_WAVE_MAXW];
	__shared__ WavePlan pl;
	__shared__ u32 tgt[B200_WAVE_SLOTS], n_tgt;
	const u32 rank = tail_rank<NC>(), ctid = rank * TAIL_THREADS + threadIdx.x, q = blockIdx.x / NC;
	cudaGridDependencySynchronize();
	WV_TRACE(5);
	wv_copy_words(&w, W.wc, sizeof w);
	__syncthreads();
	if (w.halt || q >= w.n_wave) return;
	const u32 w_iter = w.iter;
	WV_TP(24);
	if (threadIdx.x < w.n_wave) wave_gather_cut(W, w, threadIdx.x, cut[threadIdx.x]);
	if (threadIdx.x == 32) n_tgt = 0;
	__syncthreads();
	if (threadIdx.x < 32) wave_plan_warp(w, cut, S0.ctl->nrows, S0.ctl->inc_used, S0.ctl->n_live, S0.cap_rows, S0.cap_inc, W.cap_bits, pl);
	__syncthreads();
	// look-ahead lists the new rows of this cut are classified into: every pending slot this wave does not carry out
	if (threadIdx.x < w.n_pending) {
		const u32 s2 = w.pending[threadIdx.x];
		bool done = false;
		for (u32 q2 = 0; q2 < pl.n_commit; q2++) done |= (w.wave[q2] == s2);
		if (!done) tgt[atomicAdd(&n_tgt, 1u)] = s2;
	}
	__syncthreads();
	WV_TP(25);
	const u32 slot = w.wave[q];
	const DevState S = wave_view(S0, W, slot, q);
	CutCtl *c = S.ctl;
	const bool redundant = (cut[q].status & ST_REDUNDANT) != 0;
	if (q >= pl.n_commit) {                         // deferred: back to the pending list untouched
		if (!redundant) TAIL_SPREAD(i, c->n_vis) reset_class(S, i);
		if (ctid == 0) {
			c->status |= ST_WAVE_DEFER;
			if (q == 0) {                           // not even the first cut of the wave fits: the host has to act
				WaveCtl *g = W.wc;
				g->n_commit = 0;
				g->halt |= pl.halt;
				g->halt_hs = pl.halt_hs;
				g->halt_rows = pl.need_rows;
				g->halt_inc = pl.need_inc;
				g->halt_bits = pl.need_bits;
				w.halt |= pl.halt;
				wave_publish(W, w, S0.ctl->nrows, S0.ctl->n_live);
			}
		}
		return;
	}
	if (ctid == 0) {
		if (q == 0) W.wc->n_commit = pl.n_commit;
		c->n_live = pl.live_before[q];              // (statistics)
		if (!redundant) {
			c->nrows = pl.rows_base[q];
			c->slot_cnt = S0.ctl->slot_cnt + (pl.rows_base[q] - S0.ctl->nrows);
			c->inc_used = pl.inc_base[q];
			S.facet_cnt[S.cur->facet] = cut[q].n_new;   // every new row lies on the new facet
		}
	}
	if (redundant) return;
	TAIL_SYNC();
	WV_TP(26);
	const CutParams &P = *S.cur;
	const u32 n_vis = c->n_vis, H = S.he_off[n_vis], M = c->n_new;
	// ---- P5: new rows + rewiring (per half-edge) and copies + retirement (per vertex)
	TAIL_SPREAD(x, 2 * H) {
		const int part = x >= H;
		he_emit_part(S, P, part ? x - H : x, part);
	}
	TAIL_SPREAD(i, n_vis) he_finish_vertex(S, P, i);
	TAIL_SYNC();
	WV_TP(27);
	// ---- P6: dead facets ‖ K4 matrix shape, clearing of the column matrix ‖ the new rows against the pending halfspaces
	const u32 wl = (c->n_local + 63) / 64, mpad = (M + 63) & ~63u;
	if (ctid == 0) k4_plan(S);                      // (cannot overflow: wave_plan checked the upper bound)
	{
		u64 *tb = k4_tbits(S, wl, mpad);
		for (u64 x = ctid; x < (u64)c->n_local * (mpad / 64); x += NC * TAIL_THREADS) tb[x] = 0;
	}
	TAIL_SPREAD(i, n_vis) collect_dead_facets(S, i);
	{
		const u32 first = c->nrows, nt = n_tgt;
		for (u32 x = ctid; x < M * nt; x += NC * TAIL_THREADS) wave_classify_row(S, W, tgt[x / M], first + x % M);
	}
	TAIL_SYNC();
	WV_TP(28);
	// ---- P7: K4 bit matrices
	{
		const u32 nrows = c->nrows, f = P.facet;
		TAIL_SPREAD(j, M) k4_build_row_at(S, j, nrows, f, wl, mpad);
	}
	WV_TP(29);
}

// pair test of every carried-out cut of the wave; tile pairs / survivor rounds of all cuts are dealt round-robin
__global__ void __launch_bounds__(K_THREADS) k_wave_k4_filter(DevState S0, WaveDev W)
{
	__shared__ u32 s_slot[B200_WAVE_MAXW], s_M[B200_WAVE_MAXW], s_wl[B200_WAVE_MAXW], s_mpad[B200_WAVE_MAXW], s_n;
	cudaGridDependencySynchronize();
	WV_TRACE(6);
	const WaveCtl *w = W.wc;
	if (threadIdx.x == 0) s_n = w->halt ? 0u : w->n_commit;
	if (threadIdx.x < B200_WAVE_MAXW) {              // shapes of all cuts by parallel threads (no serial walk through global memory)
		const u32 slot = w->wave[threadIdx.x];
		const CutCtl *c = W.ctl + (slot < B200_WAVE_SLOTS ? slot : 0);
		s_slot[threadIdx.x] = slot;
		s_M[threadIdx.x] = (c->status & ST_SKIP_B) ? 0u : c->n_new;
		s_wl[threadIdx.x] = c->wl;
		s_mpad[threadIdx.x] = c->mpad;
	}
	__syncthreads();
	const bool shk = w->shard_k4 != 0;
	const u32 vg = shk ? gridDim.x * W.nranks : gridDim.x, vb = shk ? blockIdx.x * W.nranks + W.rank : blockIdx.x;
	u32 base = 0;
	for (u32 q = 0; q < s_n; q++) {
		if (!s_M[q]) continue;
		const DevState S = wave_view(S0, W, s_slot[q], q);
		// tile pairs of all cuts are dealt round-robin over the blocks -- of every rank when the pair test is sharded
		k4_filter_body(S, k4_threshold(S, true), s_M[q], s_wl[q], s_mpad[q], (vb + vg - base % vg) % vg, vg);
		base += k4_tile_pairs(s_M[q]);
	}
}
__global__ void __launch_bounds__(K_THREADS) k_wave_k4_contain(DevState S0, WaveDev W)
{
	__shared__ u32 s_slot[B200_WAVE_MAXW], s_ns[B200_WAVE_MAXW], s_M[B200_WAVE_MAXW], s_wl[B200_WAVE_MAXW], s_mpad[B200_WAVE_MAXW], s_off[B200_WAVE_MAXW + 1], s_n;
	cudaGridDependencySynchronize();
	WV_TRACE(7);
	const WaveCtl *w = W.wc;
	if (threadIdx.x == 0) s_n = w->halt ? 0u : w->n_commit;
	if (threadIdx.x < B200_WAVE_MAXW) {
		const u32 slot = w->wave[threadIdx.x];
		const CutCtl *c = W.ctl + (slot < B200_WAVE_SLOTS ? slot : 0);
		s_slot[threadIdx.x] = slot;
		s_ns[threadIdx.x] = ((c->status & ST_SKIP_B) || c->n_surv > W.cap_pairs) ? 0u : c->n_surv;   // k_wave_tail2 reports the overflow
		s_M[threadIdx.x] = c->n_new;
		s_wl[threadIdx.x] = c->wl;
		s_mpad[threadIdx.x] = c->mpad;
	}
	__syncthreads();
	const bool shk = w->shard_k4 != 0;
	if (shk && blockIdx.x == 0 && threadIdx.x < s_n) {       // a local survivor list overflowed: every rank must learn it
		const CutCtl *c = W.ctl + s_slot[threadIdx.x];
		if (!(c->status & ST_SKIP_B) && c->n_surv > W.cap_pairs) {
			atomicOr((u32 *)(W.xksend + 2), 1u);
			atomicMax((u32 *)(W.xksend + 3), c->n_surv);
		}
	}
	if (threadIdx.x == 0) {
		u32 t = 0;
		for (u32 q = 0; q < s_n; q++) { s_off[q] = t; t += s_ns[q]; }
		s_off[s_n] = t;
	}
	__syncthreads();
	// one warp per surviving pair, over the survivors of all cuts of the wave; an adjacent pair is flagged in place
	const u32 lane = threadIdx.x & 31, nwarps = gridDim.x * (K_THREADS / 32), total = s_off[s_n];
	u32 cur_q = B200_NONE, found = 0;
	for (u32 g = blockIdx.x * (K_THREADS / 32) + (threadIdx.x >> 5); g < total; g += nwarps) {
		u32 q = 0;
		while (s_off[q + 1] <= g) q++;
		if (q != cur_q) {                            // (warp-uniform) pair count of the cut this warp is leaving
			if (!shk && cur_q != B200_NONE && found && lane == 0) atomicAdd(&W.ctl[s_slot[cur_q]].n_pairs, found);
			cur_q = q;
			found = 0;
		}
		const DevState S = wave_view(S0, W, s_slot[q], q);
		const u32 sv = g - s_off[q], a = S.surv_a[sv], b = S.surv_b[sv];
		if (k4_columns_verdict(S, a, b, lane, s_M[q], s_wl[q], s_mpad[q])) {
			if (lane == 0) {
				if (shk) wave_k4_send_pair(W, q, a, b);   // counted and filed when the records of all ranks are merged
				else wave_flag_adjacent(S, sv, a, b);
			}
			found++;
		}
	}
	if (!shk && cur_q != B200_NONE && found && lane == 0) atomicAdd(&W.ctl[s_slot[cur_q]].n_pairs, found);
}
// exchange of the adjacent pairs (sharded pair test): same protocol as k_wave_xpush / k_wave_xmerge
__global__ void __launch_bounds__(K_THREADS) k_wave_k4_xpush(DevState S, WaveDev W)
{
	cudaGridDependencySynchronize();
	const WaveCtl *w = W.wc;
	if (w->halt || !w->shard_k4) return;
	const u32 p = blockIdx.x;
	if (p >= W.nranks || p == W.rank) return;
	const u32 seq = w->xseq_k;
	const unsigned long long cnt = W.xksend[0] & 0xFFFFFFFFull;
	const u32 n = (u32)(cnt < B200_XK_CAP ? cnt : B200_XK_CAP) + 4;
	unsigned long long *dst = W.xkpeer_recv[p] + ((size_t)W.rank * 2 + (seq & 1u)) * B200_XK_WORDS;
	for (u32 x = threadIdx.x; x < n; x += K_THREADS) dst[x] = W.xksend[x];
	__threadfence_system();
	__syncthreads();
	if (threadIdx.x == 0) wv_st_release_sys(W.xkpeer_flag[p] + W.rank * 32, seq);
}
__global__ void __launch_bounds__(K_THREADS) k_wave_k4_xmerge(DevState S, WaveDev W)
{
	__shared__ u32 s_ok;
	__shared__ WaveCtl sw;
	cudaGridDependencySynchronize();
	WaveCtl *w = W.wc;
	if (w->halt || !w->shard_k4) return;
	const u32 src = blockIdx.x % W.nranks, part = blockIdx.x / W.nranks, nparts = gridDim.x / W.nranks;
	if (part >= nparts) return;
	const u32 seq = w->xseq_k;
	const unsigned long long *rec = W.xksend;
	if (src != W.rank) {
		if (threadIdx.x == 0) {
			const u64 t0 = b200_globaltimer();
			u32 ok = 1;
			while ((int)(wv_ld_acquire_sys(W.xkflag + src * 32) - seq) < 0) {
				if (b200_globaltimer() - t0 > 4000000000ull) { ok = 0; break; }
				__nanosleep(100);
			}
			s_ok = ok;
		}
		__syncthreads();
		if (!s_ok) {
			if (threadIdx.x == 0) { const u32 h = atomicOr(&w->halt, (u32)WH_XFAIL) | WH_XFAIL; W.progress->halt = h; }
			return;
		}
		rec = W.xkrecv + ((size_t)src * 2 + (seq & 1u)) * B200_XK_WORDS;
	}
	const u32 cnt = (u32)__ldcg(rec), flags = (u32)__ldcg(rec + 2), need = (u32)__ldcg(rec + 3);
	if (cnt > B200_XK_CAP || (flags & 1u)) {     // every rank reads the same headers and halts the same way
		if (part == 0 && threadIdx.x == 0) {
			const u32 why = (flags & 1u) ? (u32)WH_GROW_PAIRS : (u32)WH_XOVER_K4;
			if (flags & 1u) atomicMax(&w->halt_pairs, need);
			const u32 h = atomicOr(&w->halt, why) | why;
			W.progress->halt = h;
		}
		return;
	}
	wv_copy_words(&sw, w, sizeof sw);            // (wave positions -> slots)
	__syncthreads();
	for (u32 x = part * K_THREADS + threadIdx.x; x < cnt; x += nparts * K_THREADS) wave_k4_merge_pair(S, W, sw, __ldcg(rec + 4 + x));
}
// before the pair test is redone with larger buffers
__global__ void __launch_bounds__(K_THREADS) k_wave_k4_reset(DevState S0, WaveDev W)
{
	const WaveCtl *w = W.wc;
	if (w->shard_k4 && blockIdx.x == 0 && threadIdx.x == 0) wave_k4_record_reset(W, w->xseq_k);   // (the host advanced the exchange number)
	for (u32 q = 0; q < w->n_commit; q++) {
		const DevState S = wave_view(S0, W, w->wave[q], q);
		if (S.ctl->status & ST_SKIP_B) continue;
		if (blockIdx.x == 0 && threadIdx.x == 0) { S.ctl->n_surv = 0; S.ctl->n_pairs = 0; }
		B200_GRID_STRIDE(j, S.ctl->n_new) S.deg[j] = 0;
	}
}

// adjacency build (k_tail2) at the wave's bases (every cluster derives them from the same counters); the commit is
// left to the first kernel of the next iteration
template <int NC> __global__ void __launch_bounds__(TAIL_THREADS, 1) k_wave_tail2(DevState S0, WaveDev W)
{
	__shared__ WaveCtl w;
	__shared__ WaveCut cut[B200_WAVE_MAXW];
	__shared__ u32 ws[33];
	__shared__ u32 s_base[B200_WAVE_MAXW], s_new[B200_WAVE_MAXW], s_fl;
	const u32 rank = tail_rank<NC>(), ctid = rank * TAIL_THREADS + threadIdx.x, q = blockIdx.x / NC;
	cudaGridDependencySynchronize();
	WV_TRACE(8);
	wv_copy_words(&w, W.wc, sizeof w);
	__syncthreads();
	if (w.halt || q >= w.n_commit) return;
	const u32 w_iter = w.iter;
	WV_TP(32);
	const u32 n_commit = w.n_commit;
	if (threadIdx.x < n_commit) wave_gather_cut(W, w, threadIdx.x, cut[threadIdx.x]);
	__syncthreads();
	const DevState S = wave_view(S0, W, w.wave[q], q);
	CutCtl *c = S.ctl;
	if (threadIdx.x == 0) {
		u32 need_adj, need_pairs;
		s_fl = wave_adj_plan(cut, n_commit, S0.ctl->adj_used, S0.cap_adj, W.cap_pairs, s_base, s_new, need_adj, need_pairs);
		if (rank == 0) {
			c->adj_used = s_base[q];
			c->adj_new = s_new[q];
			if (q == 0 && s_fl) {                    // the host grows the buffer and has this kernel (and the pair test) redone
				WaveCtl *g = W.wc;
				if (s_fl & 8u) { g->halt |= WH_GROW_PAIRS; g->halt_pairs = need_pairs; w.halt |= WH_GROW_PAIRS; }
				else { g->halt |= WH_GROW_ADJ; g->halt_adj = need_adj; w.halt |= WH_GROW_ADJ; }
				wave_publish(W, w, S0.ctl->nrows, S0.ctl->n_live);
			}
		}
	}
	__syncthreads();
	WV_TP(33);
	if (s_fl || (cut[q].status & ST_REDUNDANT)) return;
	const u32 n = cut[q].n_new;
	if (rank == 0) {
		u32 carry = 0;
		for (u32 base = 0; base < n; base += TAIL_THREADS) {
			u32 j = base + threadIdx.x, v = j < n ? S.new_padj_len[j] + S.deg[j] : 0, tot;
			u32 e = block_excl_scan(v, ws, tot);
			if (j < n) S.adj_base[j] = carry + e;
			carry += tot;
		}
	}
	TAIL_SYNC();
	WV_TP(34);
	TAIL_SPREAD(j, n) adj_place(S, j);
	if (w.shard_k4) { TAIL_SPREAD(pp, cut[q].n_pairs) adj_pair_fill(S, pp); }      // the merged pair list of all ranks
	else TAIL_SPREAD(sv, cut[q].n_surv) adj_pair_fill_surv(S, sv);
	TAIL_SYNC();
	WV_TP(35);
	TAIL_SPREAD(j, n) adj_sort(S, j);
	WV_TP(36);
}
